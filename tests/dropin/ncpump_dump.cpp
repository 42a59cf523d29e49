// Parity program for the natural-convection pump loops: production/ncpump.cpp:112-245 written against the reference API exactly as
// that driver does — walls on the four faces, SetT/SetQ patches, SmoothCorner, then the INTERIOR solid block: bounce-back along its
// four edges (BoundaryConditionAlongX/YEdge), SmoothCornerAt on its four corners, SetQ along its edges on the thermal lattice —
// forward loop (AD::MacroBrinkmanCollideNaturalConvection), adjoint loop (AAD::...NaturalConvectionMassFlow with the "i" versions
// of every closure), AAD::SensitivityBrinkmanDiffusivity, the objective read straight from ux.  Design: closed form instead of
// the filtered MMA variable; alpha/diffusivity/dads/dkds from it with the driver's formulas (ncpump.cpp:92-97).
// Built twice from this one source (see tests/dropin/transient_dump.cpp): reference headers -> fixtures, drop-in headers -> test.
//   ncpump_dump <lx> <ly> <nt> <dir>        writes <dir>/*.out
#define _USE_AVX_DEFINES
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "particle/d2q9.h"
#include "equation/advection.h"
#include "equation/adjointadvection.h"
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
    if (argc != 5) { fprintf(stderr, "usage: ncpump_dump lx ly nt dir\n"); return 2; }
    const int lx = atoi(argv[1]), ly = atoi(argv[2]), nt = atoi(argv[3]);
    dir = argv[4];
    double viscosity = 0.1/6.0, diff_fluid = viscosity/1.0, Th = 1.0, Tl = 0.0, gx = 0.0, gy = 1000*pow(viscosity, 2)/(double)pow(lx - 1, 3);
    double alphamax = 1e5, diff_solid = diff_fluid*10.0, qf = 1e-6, qg = 1e-4;
    D2Q9<double> pf(lx, ly), pg(lx, ly);
    const int n = pf.nxyz;
    double *rho = new double[n], *ux = new double[n], *uy = new double[n], *uxp = new double[n], *uyp = new double[n];
    double *tem = new double[n], *qx = new double[n], *qy = new double[n], *qxp = new double[n], *qyp = new double[n];
    double *irho = new double[n], *iux = new double[n], *iuy = new double[n], *imx = new double[n], *imy = new double[n], *iuxp = new double[n], *iuyp = new double[n];
    double *item = new double[n], *iqx = new double[n], *iqy = new double[n], *iqxp = new double[n], *iqyp = new double[n];
    for (int idx = 0; idx < n; idx++) {
        rho[idx] = 1.0;  ux[idx] = 0.0;  uy[idx] = 0.0;  uxp[idx] = 0.0;  uyp[idx] = 0.0;  tem[idx] = 0.5*(Th + Tl);  qx[idx] = 0.0;  qy[idx] = 0.0;   qxp[idx] = 0.0; qyp[idx] = 0.0;
        irho[idx] = 0.0; iux[idx] = 0.0; iuy[idx] = 0.0; iuxp[idx] = 0.0; iuyp[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0; item[idx] = 0.0; iqx[idx] = 0.0; iqy[idx] = 0.0; iqxp[idx] = 0.0; iqyp[idx] = 0.0;
    }
    double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
    double *gi = new double[n*pg.nc], *igi = new double[n*pg.nc];
    double *directionx = new double[n], *directiony = new double[n];
    for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) {
        int idx = pf.Index(i, j);
        directionx[idx] = (i == lx/2 && j > 9*ly/10) ? -1.0 : 0.0;
        directiony[idx] = 0.0;
        const double ss = j < ly/2 ? 0.5 + 0.4*sin(0.37*i)*cos(0.23*j) : 1.0;
        diffusivity[idx] = diff_solid + (diff_fluid - diff_solid)*ss*(1.0 + qg)/(ss + qg);
        alpha[idx] = alphamax/(double)(ly - 1)*qf*(1.0 - ss)/(ss + qf);
        dkds[idx] = (diff_fluid - diff_solid)*qg*(1.0 + qg)/pow(ss + qg, 2.0);
        dads[idx] = -alphamax/(double)(ly - 1)*qf*(1.0 + qf)/pow(ss + qf, 2.0);
    }
    typedef std::chrono::steady_clock clk;
    double residual = 0.0;

    NS::InitialCondition(pf, rho, ux, uy);
    AD::InitialCondition(pg, tem, ux, uy);
    clk::time_point t0 = clk::now();
    for (int t = 1; t <= nt; ++t) {
        if (t%50 == 0) residual = Residual(ux, uy, uxp, uyp, pf.nxyz);
        AD::MacroBrinkmanCollideNaturalConvection(pf, rho, ux, uy, alpha, viscosity, pg, tem, qx, qy, diffusivity, gx, gy, 0.5*(Th + Tl), true, gi);
        pf.Stream();
        pg.Stream();
        pf.BoundaryCondition([=](int _i, int _j) { return 1; });
        pg.BoundaryCondition([=](int _i, int _j) { return 0; });
        AD::BoundaryConditionSetT(pg, [=](int _i, int _j) { return _i == 0 ? Th : Tl; }, ux, uy,
            [=](int _i, int _j) { return (_i == 0 && _j < ly/2) || (_i == lx - 1 && _j < ly/2); });
        AD::BoundaryConditionSetQ(pg, [=](int _i, int _j) { return 0.0; }, ux, uy, diffusivity,
            [=](int _i, int _j) { return (_i == 0 && ly/2 <= _j) || (_i == lx - 1 && ly/2 <= _j) || _j == 0 || _j == ly - 1; });
        pf.SmoothCorner();
        pg.SmoothCorner();

        pf.BoundaryConditionAlongXEdge(lx/5, 1, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        pf.BoundaryConditionAlongYEdge(ly/2, 1, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pf.BoundaryConditionAlongXEdge(4*lx/5, -1, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        pf.BoundaryConditionAlongYEdge(9*ly/10, 1, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pf.SmoothCornerAt(lx/5, ly/2, -1, -1);
        pf.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
        pf.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
        pf.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);
        AD::BoundaryConditionSetQAlongXEdge(pg, lx/5, 1, [=](int _i, int _j) { return 0.0; }, ux, uy, diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        AD::BoundaryConditionSetQAlongYEdge(pg, ly/2, 1, [=](int _i, int _j) { return 0.0; }, ux, uy, diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        AD::BoundaryConditionSetQAlongXEdge(pg, 4*lx/5, -1, [=](int _i, int _j) { return 0.0; }, ux, uy, diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        AD::BoundaryConditionSetQAlongYEdge(pg, 9*ly/10, 1, [=](int _i, int _j) { return 0.0; }, ux, uy, diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pg.SmoothCornerAt(lx/5, ly/2, -1, -1);
        pg.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
        pg.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
        pg.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);

        std::swap(ux, uxp); std::swap(uy, uyp); std::swap(qx, qxp); std::swap(qy, qyp);
    }
#ifdef PANSLBM_B200_DROPIN
    plh_sync();
#endif
    clk::time_point t1 = clk::now();

    ANS::InitialCondition(pf, ux, uy, irho, iux, iuy);
    AAD::InitialCondition(pg, ux, uy, item, iqx, iqy);
    clk::time_point t2 = clk::now();
    for (int t = 1; t <= nt; ++t) {
        AAD::MacroBrinkmanCollideNaturalConvectionMassFlow(pf, rho, ux, uy, irho, iux, iuy, imx, imy, alpha, viscosity,
            pg, tem, item, iqx, iqy, diffusivity, gx, gy, directionx, directiony, true, igi);
        pf.iStream();
        pg.iStream();
        pf.iBoundaryCondition([=](int _i, int _j) { return 1; });
        pg.iBoundaryCondition([=](int _i, int _j) { return 0; });
        AAD::iBoundaryConditionSetT(pg, ux, uy, [=](int _i, int _j) { return (_i == 0 && _j < ly/2) || (_i == lx - 1 && _j < ly/2); });
        AAD::iBoundaryConditionSetQ(pg, ux, uy, [=](int _i, int _j) { return (_i == 0 && ly/2 <= _j) || (_i == lx - 1 && ly/2 <= _j) || _j == 0 || _j == ly - 1; });
        pf.SmoothCorner();
        pg.SmoothCorner();

        pf.iBoundaryConditionAlongXEdge(lx/5, 1, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        pf.iBoundaryConditionAlongYEdge(ly/2, 1, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pf.iBoundaryConditionAlongXEdge(4*lx/5, -1, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        pf.iBoundaryConditionAlongYEdge(9*ly/10, 1, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pf.SmoothCornerAt(lx/5, ly/2, -1, -1);
        pf.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
        pf.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
        pf.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);
        AAD::iBoundaryConditionSetQAlongXEdge(pg, lx/5, 1, ux, uy, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        AAD::iBoundaryConditionSetQAlongYEdge(pg, ly/2, 1, ux, uy, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        AAD::iBoundaryConditionSetQAlongXEdge(pg, 4*lx/5, -1, ux, uy, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
        AAD::iBoundaryConditionSetQAlongYEdge(pg, 9*ly/10, 1, ux, uy, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
        pg.SmoothCornerAt(lx/5, ly/2, -1, -1);
        pg.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
        pg.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
        pg.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);

        std::swap(iux, iuxp); std::swap(iuy, iuyp); std::swap(iqx, iqxp); std::swap(iqy, iqyp);
    }
#ifdef PANSLBM_B200_DROPIN
    plh_sync();
#endif
    clk::time_point t3 = clk::now();

    double f_buffer = 0.0;
    for (int j = 0; j < pf.ny; ++j) {
        int i = lx/2;
        if (j > 9*ly/10) { int idx = pf.Index(i, j); f_buffer += ux[idx]*directionx[idx] + uy[idx]*directiony[idx]; }
    }
    std::vector<double> dfdss(n, 0.0);
    AAD::SensitivityBrinkmanDiffusivity(pg, dfdss.data(), ux, uy, imx, imy, dads, tem, item, iqx, iqy, gi, igi, diffusivity, dkds);

    const char* names[] = {"rho", "ux", "uy", "tem", "qx", "qy", "ip", "iux", "iuy", "imx", "imy", "item", "iqx", "iqy"};
    double* arrs[] = {rho, ux, uy, tem, qx, qy, irho, iux, iuy, imx, imy, item, iqx, iqy};
    for (int a = 0; a < 14; ++a) wr(names[a], arrs[a], n);
    wr("dfdss", dfdss.data(), n);
    wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
    double extra[2] = {f_buffer, residual};
    wr("extra", extra, 2);
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
