// Parity program for the GPU filters behind the drop-in headers (panslbm2_b200/src/utility/{density,heaviside}filter.h), written
// against the reference API as the drivers use it (production/heatsink3D.cpp:87-103, 249-250; production/ncpump.cpp:83, 257).
//   filter_dump <dim> <lx> <ly> <lz> <R> <beta> <bx> <by> <bz> <dir>     reads <dir>/v.bin, <dir>/d.bin; writes <dir>/{fv,rho,dfds}.out
// bx > 0: the heatsink drivers' design-box weight, else the default cone weight.
#include <cstdio>
#include <string>
#include <vector>
#include "../../panslbm2_b200/src/particle/d2q9.h"
#include "../../panslbm2_b200/src/particle/d3q15.h"
#include "../../panslbm2_b200/src/utility/densityfilter.h"
#include "../../panslbm2_b200/src/utility/heavisidefilter.h"

using namespace PANSLBM2;
static std::string dir;
static std::vector<double> rd(const char* name, size_t n) {
    std::vector<double> v(n);
    FILE* f = fopen((dir + "/" + name).c_str(), "rb");
    if (!f || fread(v.data(), sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", name); exit(2); }
    fclose(f);
    return v;
}
static void wr(const char* name, const std::vector<double>& v) {
    volatile double first = v.empty() ? 0.0 : v[0];     // refresh a stale host copy before the system call reads it
    (void)first;
    FILE* f = fopen((dir + "/" + name + ".out").c_str(), "wb");
    fwrite(v.data(), sizeof(double), v.size(), f);
    fclose(f);
}
template<class P>
static void run(P& pf, double R, double beta, int bx, int by, int bz) {
    std::vector<double> v = rd("v.bin", pf.nxyz), d = rd("d.bin", pf.nxyz);
    auto filterweight = [=](int _i1, int _j1, int _k1, int _i2, int _j2, int _k2) {
        if (_i1 < bx && _j1 < by && _k1 < bz && _i2 < bx && _j2 < by && _k2 < bz) {
            return (R - sqrt(pow(_i1 - _i2, 2.0) + pow(_j1 - _j2, 2.0) + pow(_k1 - _k2, 2.0)))/R;
        } else {
            return (_i1 == _i2 && _j1 == _j2 && _k1 == _k2) ? 1.0 : 0.0;
        }
    };
    for (int rep = 0; rep < 2; ++rep) {        // second round: the cached weight tables are reused
        if (bx > 0) {
            wr("fv", DensityFilter::GetFilteredValue(pf, R, v, filterweight));
            wr("rho", HeavisideFilter::GetFilteredVariable(pf, R, beta, v, filterweight));
            wr("dfds", HeavisideFilter::GetFilteredSensitivity(pf, R, beta, v, d, filterweight));
        } else {
            wr("fv", DensityFilter::GetFilteredValue(pf, R, v));
            wr("rho", HeavisideFilter::GetFilteredVariable(pf, R, beta, v));
            wr("dfds", HeavisideFilter::GetFilteredSensitivity(pf, R, beta, v, d));
        }
    }
}
int main(int argc, char** argv) {
    if (argc != 11) { fprintf(stderr, "usage: filter_dump dim lx ly lz R beta bx by bz dir\n"); return 2; }
    const int dim = atoi(argv[1]), lx = atoi(argv[2]), ly = atoi(argv[3]), lz = atoi(argv[4]);
    const double R = atof(argv[5]), beta = atof(argv[6]);
    const int bx = atoi(argv[7]), by = atoi(argv[8]), bz = atoi(argv[9]);
    dir = argv[10];
    if (dim == 3) { D3Q15<double> pf(lx, ly, lz); run(pf, R, beta, bx, by, bz); }
    else { D2Q9<double> pf(lx, ly); run(pf, R, beta, bx, by, bz); }
    return 0;
}
