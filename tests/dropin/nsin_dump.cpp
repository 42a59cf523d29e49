// Parity program for the incompressible NS model (src/equation/nsincompressible.h): no program of the reference uses NSin, so the
// loop body is the pipe-bend one of production/nsopt.cpp:82-103 with the NSin functions in place of the NS ones — D2Q9
// NSin::MacroBrinkmanCollide, Stream, bounce-back everywhere but the inlet / outlet patches, a parabolic NSin SetU inlet on xmin,
// an NSin SetRho outlet on ymin (through the two edge functions: NSin::BoundaryConditionSetRho itself does not compile in the
// reference, nsincompressible.h:238), SmoothCorner, std::swap of (ux, uxp), Residual every dt steps — followed by a few steps of
// NSin::MacroCollide with a moving lid (NSin::BoundaryConditionSetU on all four edges).
// Built twice from this one source: reference headers -> fixtures (tests/golden/make_nsin_golden.py), drop-in headers -> test.
//   nsin_dump <lx> <ly> <nt> <dt> <dir>        writes <dir>/*.out
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "particle/d2q9.h"
#include "equation/nsincompressible.h"
#include "utility/residual.h"

using namespace PANSLBM2;

static std::string dir;
static void wr(const std::string& name, const double* p, size_t n) {
    volatile double first = n ? p[0] : 0.0;
    (void)first;
    FILE* f = fopen((dir + "/" + name + ".out").c_str(), "wb");
    fwrite(p, sizeof(double), n, f);
    fclose(f);
}

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: nsin_dump lx ly nt dt dir\n"); return 2; }
    const int lx = atoi(argv[1]), ly = atoi(argv[2]), nt = atoi(argv[3]), dt = atoi(argv[4]);
    dir = argv[5];
    double nu = 0.1, u0 = 0.01, q = 0.01, amax = 2e2;
    D2Q9<double> pf(lx, ly);
    const int n = pf.nxyz;
    double *rho = new double[n], *ux = new double[n], *uy = new double[n], *uxp = new double[n], *uyp = new double[n], *alpha = new double[n];
    for (int idx = 0; idx < n; ++idx) { rho[idx] = 1.0; ux[idx] = 0.0;  uy[idx] = 0.0;  uxp[idx] = 0.0; uyp[idx] = 0.0; }
    for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) {
        const double s = 0.6 + 0.35*sin(0.29*i + 0.4)*cos(0.17*j);
        alpha[pf.Index(i, j)] = amax/(double)lx*q*(1.0 - s)/(s + q);
    }
    double res_f = 0.0;

    NSin::InitialCondition(pf, rho, ux, uy);
    for (int td = 1; td <= nt; ++td) {
        NSin::MacroBrinkmanCollide(pf, rho, ux, uy, nu, alpha, true);
        if (td%dt == 0) res_f = Residual(ux, uy, uxp, uyp, pf.nxyz);
        pf.Stream();
        pf.BoundaryCondition([=](int _i, int _j) { return ((_i == 0 && 0.7*ly < _j && _j < 0.9*ly) || (_j == 0 && 0.7*lx < _i && _i < 0.9*lx)) ? 0 : 1; });
        NSin::BoundaryConditionSetU(pf,
            [=](int _i, int _j) { return -u0*(_j - 0.7*ly)*(_j - 0.9*ly)/(0.1*ly*0.1*ly); },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return _i == 0 && 0.7*ly < _j && _j < 0.9*ly; }
        );
        NSin::BoundaryConditionSetRhoAlongYEdge(pf, 0, -1,
            [=](int _i, int _j) { return 1.0; },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return _j == 0 && 0.7*lx < _i && _i < 0.9*lx; }
        );
        pf.SmoothCorner();

        std::swap(ux, uxp);
        std::swap(uy, uyp);
    }
    const char* names[] = {"rho", "ux", "uy", "uxp", "uyp"};
    double* arrs[] = {rho, ux, uy, uxp, uyp};
    for (int a = 0; a < 5; ++a) wr(names[a], arrs[a], n);
    wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1));

    // lid-driven box with NSin::MacroCollide (no Brinkman term), velocities prescribed on all four edges
    for (int td = 1; td <= nt/2; ++td) {
        NSin::MacroCollide(pf, rho, ux, uy, nu, true);
        pf.Stream();
        NSin::BoundaryConditionSetU(pf,
            [=](int _i, int _j) { return _j == ly - 1 ? 5.0*u0 : 0.0; },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return true; }
        );
        pf.SmoothCorner();
    }
    wr("lid.rho", rho, n); wr("lid.ux", ux, n); wr("lid.uy", uy, n);
    wr("lid.f0", pf.f0, n); wr("lid.f", pf.f, (size_t)n*(pf.nc - 1));
    double extra[1] = {res_f};
    wr("extra", extra, 1);
#ifdef PANSLBM_B200_DROPIN
    uint64_t st[8];
    plh_stats(st);
    double std_[8];
    for (int k = 0; k < 8; ++k) std_[k] = (double)st[k];
    wr("stats", std_, 8);
    printf("fused steps %llu, calls one by one %llu, uploads %llu, downloads %llu, faults %llu, plans %llu, settles %llu, stagings %llu\n",
           (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)st[3], (unsigned long long)st[4],
           (unsigned long long)st[5], (unsigned long long)st[6], (unsigned long long)st[7]);
#endif
    return 0;
}
