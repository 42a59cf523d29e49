// Parity program for closures whose VALUES change while the loop runs — what the reference handles by calling the user's lambdas
// on every boundary site in every call (navierstokes.h:155-157) and the drop-in surface has to notice although it bakes planes:
//   mode 0: the inlet velocity lives in a local the lambda captures BY REFERENCE ([&], the style of test/nssens.cpp, test/fsi.cpp)
//   mode 1: it is captured BY VALUE and the lambda is re-created with another value every step (a time-dependent inlet)
//   mode 2: the lambda reads a profile through a captured POINTER; only the far end of the profile changes (beyond what the
//           closure key looks at: caught by revalidation — the test runs it with PANSLBM_B200_REVALIDATE=1)
// A D2Q9 channel (test/nssens.cpp's loop body: MacroBrinkmanCollide, Stream, bounce-back walls, SetU inlet, SetRho outlet).
// Built twice from this one source: -I<reference>/src -fopenmp (fixtures, tests/golden/make_closure_golden.py) and
// -I panslbm2_b200/src (the program under test).      closure_dump <mode> <lx> <ly> <nt> <dir>   writes <dir>/*.out
#define _USE_AVX_DEFINES
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "particle/d2q9.h"
#include "equation/navierstokes.h"

using namespace PANSLBM2;

static std::string dir;
static void wr(const std::string& name, const double* p, size_t n) {
    volatile double first = n ? p[0] : 0.0;     // user-space touch first (see tests/dropin/heatsink_dump.cpp)
    (void)first;
    FILE* f = fopen((dir + "/" + name + ".out").c_str(), "wb");
    fwrite(p, sizeof(double), n, f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: closure_dump mode lx ly nt dir\n"); return 2; }
    const int mode = atoi(argv[1]), lx = atoi(argv[2]), ly = atoi(argv[3]), nt = atoi(argv[4]);
    dir = argv[5];
    const double nu = 0.1, u0 = 0.03;
    D2Q9<double> pf(lx, ly);
    const int n = pf.nxyz;
    double *rho = new double[n], *ux = new double[n], *uy = new double[n], *alpha = new double[n];
    for (int idx = 0; idx < n; ++idx) { rho[idx] = 1.0; ux[idx] = 0.0; uy[idx] = 0.0; alpha[idx] = 0.0; }
    for (int i = lx/3; i < lx/2; ++i) for (int j = ly/3; j < ly/2; ++j) alpha[pf.Index(i, j)] = 0.4;
    std::vector<double> profile(ly, 0.0);
    const double* prof = profile.data();
    double uin = u0;
    NS::InitialCondition(pf, rho, ux, uy);
    for (int t = 1; t <= nt; ++t) {
        uin = u0*(1.0 + 0.5*std::sin(0.3*t));
        for (int j = 0; j < ly; ++j) profile[j] = u0*(1.0 - std::pow((2.0*j - (ly - 1))/(ly - 1), 2.0))*(j >= ly/2 ? 1.0 + 0.5*std::sin(0.3*t) : 1.0);
        NS::MacroBrinkmanCollide(pf, rho, ux, uy, nu, alpha, true);
        pf.Stream();
        pf.BoundaryCondition([=](int _i, int _j) { return (_j == 0 || _j == ly - 1) ? 1 : 0; });
        if (mode == 0) {
            NS::BoundaryConditionSetU(pf, [&](int _i, int _j) { return uin; }, [&](int _i, int _j) { return 0.0; },
                                      [=](int _i, int _j) { return _i == 0 && _j > 0 && _j < ly - 1; });
        } else if (mode == 1) {
            const double now = uin;
            NS::BoundaryConditionSetU(pf, [=](int _i, int _j) { return now; }, [=](int _i, int _j) { return 0.0; },
                                      [=](int _i, int _j) { return _i == 0 && _j > 0 && _j < ly - 1; });
        } else {
            NS::BoundaryConditionSetU(pf, [=](int _i, int _j) { return prof[_j]; }, [=](int _i, int _j) { return 0.0; },
                                      [=](int _i, int _j) { return _i == 0 && _j > 0 && _j < ly - 1; });
        }
        NS::BoundaryConditionSetRho(pf, [=](int _i, int _j) { return 1.0; }, [=](int _i, int _j) { return 0.0; },
                                    [=](int _i, int _j) { return _i == lx - 1 && _j > 0 && _j < ly - 1; });
    }
    wr("rho", rho, n); wr("ux", ux, n); wr("uy", uy, n);
    wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1));
#ifdef PANSLBM_B200_DROPIN
    uint64_t st[8];
    plh_stats(st);
    double std_[8];
    for (int k = 0; k < 8; ++k) std_[k] = (double)st[k];
    wr("stats", std_, 8);
#endif
    return 0;
}
