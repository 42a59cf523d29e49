// Parity program for the pipe-bend loops: production/nsopt.cpp:82-150 written against the reference API exactly as that driver
// does — D2Q9 NS::MacroBrinkmanCollide, Stream, bounce-back everywhere but the inlet / outlet patches, a parabolic SetU inlet on
// xmin, a SetRho outlet on ymin, SmoothCorner, std::swap of (ux, uxp) — then the adjoint loop (ANS::MacroBrinkmanCollide, iStream,
// iBoundaryCondition, iBoundaryConditionSetU with eps = 1, iBoundaryConditionSetRho2D), the pressure-drop objective read straight
// from rho (:132-142), ANS::SensitivityBrinkman and Normalize (:149-150).  Residual every dt steps as in the driver (:85-90,
// :113-118; the convergence break is disabled so that both builds run the same number of steps).  Design: closed form instead of
// the MMA variable; alpha / dads from it with the driver's formulas (:56-59).
// Built twice from this one source (see tests/dropin/transient_dump.cpp): reference headers -> fixtures, drop-in headers -> test.
// nsopt.cpp is the one reference program that leaves _USE_AVX_DEFINES commented out (:2): built as committed it runs the scalar
// templates at every site.  -DNSOPT_AVX builds the same loops with the AVX overloads (the order the drop-in computes in).
//   nsopt_dump <lx> <ly> <nt> <dt> <dir>        writes <dir>/*.out
#ifdef NSOPT_AVX
#define _USE_AVX_DEFINES
#endif
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "particle/d2q9.h"
#include "equation/navierstokes.h"
#include "equation/adjointnavierstokes.h"
#include "utility/residual.h"
#include "utility/normalize.h"

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
    if (argc != 6) { fprintf(stderr, "usage: nsopt_dump lx ly nt dt dir\n"); return 2; }
    const int lx = atoi(argv[1]), ly = atoi(argv[2]), nt = atoi(argv[3]), dt = atoi(argv[4]);
    dir = argv[5];
    double nu = 0.1, u0 = 0.01, q = 0.01, amax = 2e2;
    D2Q9<double> pf(lx, ly);
    const int n = pf.nxyz;
    double *rho = new double[n], *ux = new double[n], *uy = new double[n], *uxp = new double[n], *uyp = new double[n];
    double *irho = new double[n], *iux = new double[n], *iuy = new double[n], *imx = new double[n], *imy = new double[n], *iuxp = new double[n], *iuyp = new double[n];
    double *alpha = new double[n], *dads = new double[n];
    for (int idx = 0; idx < n; ++idx) {
        rho[idx] = 1.0; ux[idx] = 0.0;  uy[idx] = 0.0;  uxp[idx] = 0.0; uyp[idx] = 0.0;
        irho[idx] = 0.0;    iux[idx] = 0.0; iuy[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0; iuxp[idx] = 0.0;    iuyp[idx] = 0.0;
    }
    std::vector<double> s(n, 1.0);
    for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) s[pf.Index(i, j)] = 0.6 + 0.35*sin(0.29*i + 0.4)*cos(0.17*j);
    for (int idx = 0; idx < n; idx++) {
        alpha[idx] = amax/(double)lx*q*(1.0 - s[idx])/(s[idx] + q);
        dads[idx] = -amax/(double)lx*q*(q + 1.0)/pow(q + s[idx], 2.0);
    }
    typedef std::chrono::steady_clock clk;
    double res_f = 0.0, res_a = 0.0;

    NS::InitialCondition(pf, rho, ux, uy);
    clk::time_point t0 = clk::now();
    for (int td = 1; td <= nt; ++td) {
        NS::MacroBrinkmanCollide(pf, rho, ux, uy, nu, alpha, true);
        if (td%dt == 0) res_f = Residual(ux, uy, uxp, uyp, pf.nxyz);
        pf.Stream();
        pf.BoundaryCondition([=](int _i, int _j) { return ((_i == 0 && 0.7*ly < _j && _j < 0.9*ly) || (_j == 0 && 0.7*lx < _i && _i < 0.9*lx)) ? 0 : 1; });
        NS::BoundaryConditionSetU(pf,
            [=](int _i, int _j) { return -u0*(_j - 0.7*ly)*(_j - 0.9*ly)/(0.1*ly*0.1*ly); },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return _i == 0 && 0.7*ly < _j && _j < 0.9*ly; }
        );
        NS::BoundaryConditionSetRho(pf,
            [=](int _i, int _j) { return 1.0; },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return _j == 0 && 0.7*lx < _i && _i < 0.9*lx; }
        );
        pf.SmoothCorner();

        std::swap(ux, uxp);
        std::swap(uy, uyp);
    }
#ifdef PANSLBM_B200_DROPIN
    plh_sync();
#endif
    clk::time_point t1 = clk::now();

    ANS::InitialCondition(pf, ux, uy, irho, iux, iuy);
    clk::time_point t2 = clk::now();
    for (int ti = 1; ti <= nt; ++ti) {
        ANS::MacroBrinkmanCollide(pf, rho, ux, uy, irho, iux, iuy, imx, imy, nu, alpha, true);
        if (ti%dt == 0) res_a = Residual(iux, iuy, iuxp, iuyp, pf.nxyz);
        pf.iStream();
        pf.iBoundaryCondition([=](int _i, int _j) { return ((_i == 0 && 0.7*ly < _j && _j < 0.9*ly) || (_j == 0 && 0.7*lx < _i && _i < 0.9*lx)) ? 0 : 1; });
        ANS::iBoundaryConditionSetU(pf,
            [=](int _i, int _j) { return -u0*(_j - 0.7*ly)*(_j - 0.9*ly)/(0.1*ly*0.1*ly); },
            [=](int _i, int _j) { return 0.0; },
            [=](int _i, int _j) { return _i == 0 && 0.7*ly < _j && _j < 0.9*ly; },
            1.0
        );
        ANS::iBoundaryConditionSetRho2D(pf, [=](int _i, int _j) { return _j == 0 && 0.7*lx < _i && _i < 0.9*lx; });
        pf.SmoothCorner();

        std::swap(iux, iuxp);
        std::swap(iuy, iuyp);
    }
#ifdef PANSLBM_B200_DROPIN
    plh_sync();
#endif
    clk::time_point t3 = clk::now();

    double f_buffer = 0.0;
    for (int j = 0; j < pf.ny; ++j) {
        if (0.7*ly < j + pf.offsety && j + pf.offsety < 0.9*ly && pf.PEx == 0) f_buffer += rho[pf.Index(0, j)]/3.0;
    }
    for (int i = 0; i < pf.nx; i++) {
        if (0.7*lx < i + pf.offsetx && i + pf.offsetx < 0.9*lx && pf.PEy == 0) f_buffer -= rho[pf.Index(i, 0)]/3.0;
    }
    std::vector<double> dfds(n, 0.0), dfds_raw(n, 0.0);
    ANS::SensitivityBrinkman(pf, dfds.data(), ux, uy, imx, imy, dads);
    dfds_raw = dfds;
    Normalize(dfds.data(), pf.nxyz);

    const char* names[] = {"rho", "ux", "uy", "uxp", "uyp", "ip", "iux", "iuy", "iuxp", "iuyp", "imx", "imy"};
    double* arrs[] = {rho, ux, uy, uxp, uyp, irho, iux, iuy, iuxp, iuyp, imx, imy};
    for (int a = 0; a < 12; ++a) wr(names[a], arrs[a], n);
    wr("dfds_raw", dfds_raw.data(), n);
    wr("dfds", dfds.data(), n);
    wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1));
    double extra[3] = {f_buffer, res_f, res_a};
    wr("extra", extra, 3);
    {
        const double fs = std::chrono::duration<double>(t1 - t0).count(), as = std::chrono::duration<double>(t3 - t2).count();
        printf("forward %d steps %.4f ms/step %.1f MLUPS | adjoint %.4f ms/step %.1f MLUPS\n", nt, 1e3*fs/nt, (double)n*nt/fs/1e6, 1e3*as/nt, (double)n*nt/as/1e6);
    }
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
