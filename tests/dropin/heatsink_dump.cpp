// Parity program for the drop-in C++ surface (panslbm2_b200/src): one optimisation iteration of the heatsink drivers without
// the filter/MMA stages — forward loop, adjoint loop, sensitivity — written against the reference's API exactly as
// production/heatsink3D.cpp:148-246 and production/heatsink.cpp do (plain `new double[]` arrays, std::swap per step, direct
// reads afterwards), with the step budget and the design of tests/heatsink_case.py.  It dumps every field as raw fp64 so that
// tests/test_gpu_dropin.py can compare them bit for bit with the fixtures generated from the reference build (tests/golden).
//   heatsink_dump <dim> <lx> <ly> <lz> <nt> <dir>      reads <dir>/{alpha,kappa,dads,dkds}.bin, <dir>/params.bin; writes <dir>/*.out
// -DHEATSINK_SCALAR: built like a program that leaves _USE_AVX_DEFINES out (production/nsopt.cpp:2) — the headers then select the
// arithmetic of the reference's scalar templates (pl_set_scalar_order); fixtures: tests/golden/heatsink_scalar.npz.
#ifndef HEATSINK_SCALAR
#define _USE_AVX_DEFINES
#endif
#include <chrono>
#include <cstdio>
#include <string>
#include <vector>
#include "../../panslbm2_b200/src/particle/d2q9.h"
#include "../../panslbm2_b200/src/particle/d3q15.h"
#include "../../panslbm2_b200/src/equation/advection.h"
#include "../../panslbm2_b200/src/equation/adjointadvection.h"
#include "../../panslbm2_b200/src/utility/residual.h"

using namespace PANSLBM2;

static std::string dir;
static void rd(const char* name, double* p, size_t n) {
    FILE* f = fopen((dir + "/" + name).c_str(), "rb");
    if (!f || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", name); exit(2); }
    fclose(f);
}
static void wr(const char* name, const double* p, size_t n) {
    if (getenv("HEATSINK_DUMP_NO_OUTPUT") && std::string(name) != "stats") return;
    // read one element in user space first: a host copy that is stale (device newer) is refreshed by the page-fault path, which
    // a system call reading the buffer (fwrite -> write(2)) cannot trigger — it would just see EFAULT
    volatile double first = n ? p[0] : 0.0;
    (void)first;
    FILE* f = fopen((dir + "/" + name + ".out").c_str(), "wb");
    fwrite(p, sizeof(double), n, f);
    fclose(f);
}
static double* zeros(int n, double v = 0.0) { double* p = new double[n]; for (int i = 0; i < n; ++i) p[i] = v; return p; }

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage: heatsink_dump dim lx ly lz nt dir\n"); return 2; }
    const int dim = atoi(argv[1]), lx = atoi(argv[2]), ly = atoi(argv[3]), lz = atoi(argv[4]), nt = atoi(argv[5]);
    dir = argv[6];
    double prm[7];
    rd("params.bin", prm, 7);
    const double nu = prm[0], gx = prm[1], gy = prm[2], gz = prm[3], tem0 = prm[4], qn0 = prm[5], L = prm[6];
    double residual = 0.0;
    typedef std::chrono::steady_clock clk;
    clk::time_point t0, t1, t2, t3;

    if (dim == 3) {
        D3Q15<double> pf(lx, ly, lz), pg(lx, ly, lz);
        const int n = pf.nxyz;
        double *rho = zeros(n, 1.0), *ux = zeros(n), *uy = zeros(n), *uz = zeros(n), *uxp = zeros(n), *uyp = zeros(n), *uzp = zeros(n);
        double *tem = zeros(n), *qx = zeros(n), *qy = zeros(n), *qz = zeros(n), *qxp = zeros(n), *qyp = zeros(n), *qzp = zeros(n);
        double *irho = zeros(n), *iux = zeros(n), *iuy = zeros(n), *iuz = zeros(n), *imx = zeros(n), *imy = zeros(n), *imz = zeros(n), *iuxp = zeros(n), *iuyp = zeros(n), *iuzp = zeros(n);
        double *item = zeros(n), *iqx = zeros(n), *iqy = zeros(n), *iqz = zeros(n), *iqxp = zeros(n), *iqyp = zeros(n), *iqzp = zeros(n);
        double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
        double *gi = new double[n*pg.nc], *igi = new double[n*pg.nc];
        rd("alpha.bin", alpha, n); rd("kappa.bin", diffusivity, n); rd("dads.bin", dads, n); rd("dkds.bin", dkds, n);

        NS::InitialCondition(pf, rho, ux, uy, uz);
        AD::InitialCondition(pg, tem, ux, uy, uz);
        plh_sync(); t0 = clk::now();
        for (int t = 1; t <= nt; t++) {
            AD::MacroBrinkmanCollideNaturalConvection(pf, rho, ux, uy, uz, alpha, nu, pg, tem, qx, qy, qz, diffusivity, gx, gy, gz, tem0, true, gi);
            if (t%5 == 0) residual = Residual(ux, uy, uz, uxp, uyp, uzp, pf.nxyz);      // an observation in the middle of the loop body
            pf.Stream();
            pg.Stream();
            pf.BoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 1; });
            AD::BoundaryConditionSetT(pg, [=](int _i, int _j, int _k) { return tem0; }, ux, uy, uz,
                [=](int _i, int _j, int _k) { return _i == lx - 1 || _j == ly - 1 || _k == lz - 1; });
            AD::BoundaryConditionSetQ(pg, [=](int _i, int _j, int _k) { return (_j == 0 && _i < L && _k < L) ? qn0 : 0.0; }, ux, uy, uz, diffusivity,
                [=](int _i, int _j, int _k) { return _j == 0; });
            pg.BoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 0; });
            pf.SmoothCorner();
            pg.SmoothCorner();
            std::swap(ux, uxp); std::swap(uy, uyp); std::swap(uz, uzp); std::swap(qx, qxp); std::swap(qy, qyp); std::swap(qz, qzp);
        }
        plh_sync(); t1 = clk::now();
        ANS::InitialCondition(pf, ux, uy, uz, irho, iux, iuy, iuz);
        AAD::InitialCondition(pg, ux, uy, uz, item, iqx, iqy, iqz);
        plh_sync(); t2 = clk::now();
        for (int t = 1; t <= nt; t++) {
            AAD::MacroBrinkmanCollideNaturalConvection(pf, rho, ux, uy, uz, irho, iux, iuy, iuz, imx, imy, imz, alpha, nu,
                                                       pg, tem, item, iqx, iqy, iqz, diffusivity, gx, gy, gz, true, igi);
            pf.iStream();
            pg.iStream();
            AAD::iBoundaryConditionSetT(pg, ux, uy, uz, [=](int _i, int _j, int _k) { return _i == lx - 1 || _j == ly - 1 || _k == lz - 1; });
            AAD::iBoundaryConditionSetQ(pg, ux, uy, uz, [=](int _i, int _j, int _k) { return _j == 0; });
            AAD::iBoundaryConditionSetQ(pg, ux, uy, uz, [=](int _i, int _j, int _k) { return _j == 0 && _i < L && _k < L; }, 1.0);
            pg.iBoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 0; });
            pf.iBoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 1; });
            pf.SmoothCorner();
            pg.SmoothCorner();
            std::swap(iux, iuxp); std::swap(iuy, iuyp); std::swap(iuz, iuzp); std::swap(iqx, iqxp); std::swap(iqy, iqyp); std::swap(iqz, iqzp);
        }
        plh_sync(); t3 = clk::now();
        // objective read straight from the array, as production/heatsink3D.cpp:227-235 does
        double f_buffer = 0.0;
        for (int k = 0; k < pf.nz; ++k) for (int i = 0; i < pf.nx; ++i) if (i < L && k < L) f_buffer += tem[pf.Index(i, 0, k)];
        std::vector<double> dfdss(n, 0.0);
        AAD::SensitivityTemperatureAtHeatSource(pg, dfdss.data(), ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gi, igi, diffusivity, dkds,
            [=](int _i, int _j, int _k) { return (_j == 0 && _i < L && _k < L) ? qn0 : 0.0; },
            [=](int _i, int _j, int _k) { return _j == 0 && _i < L && _k < L; });
        const char* names[] = {"rho", "ux", "uy", "uz", "tem", "qx", "qy", "qz", "ip", "iux", "iuy", "iuz", "imx", "imy", "imz", "item", "iqx", "iqy", "iqz"};
        double* arrs[] = {rho, ux, uy, uz, tem, qx, qy, qz, irho, iux, iuy, iuz, imx, imy, imz, item, iqx, iqy, iqz};
        for (int a = 0; a < 19; ++a) wr(names[a], arrs[a], n);
        wr("dfdss", dfdss.data(), n);
        wr("gsnap", gi, (size_t)n*pg.nc); wr("igsnap", igi, (size_t)n*pg.nc);     // the reference's layout: [pack][c][lane], tail [idx][c]
        wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
        double extra[2] = {f_buffer, residual};
        wr("extra", extra, 2);
    } else {
        D2Q9<double> pf(lx, ly), pg(lx, ly);
        const int n = pf.nxyz;
        double *rho = zeros(n, 1.0), *ux = zeros(n), *uy = zeros(n), *uxp = zeros(n), *uyp = zeros(n);
        double *tem = zeros(n), *qx = zeros(n), *qy = zeros(n), *qxp = zeros(n), *qyp = zeros(n);
        double *irho = zeros(n), *iux = zeros(n), *iuy = zeros(n), *imx = zeros(n), *imy = zeros(n), *iuxp = zeros(n), *iuyp = zeros(n);
        double *item = zeros(n), *iqx = zeros(n), *iqy = zeros(n), *iqxp = zeros(n), *iqyp = zeros(n);
        double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
        double *gi = new double[n*pg.nc], *igi = new double[n*pg.nc];
        rd("alpha.bin", alpha, n); rd("kappa.bin", diffusivity, n); rd("dads.bin", dads, n); rd("dkds.bin", dkds, n);

        NS::InitialCondition(pf, rho, ux, uy);
        AD::InitialCondition(pg, tem, ux, uy);
        plh_sync(); t0 = clk::now();
        for (int t = 1; t <= nt; t++) {
            AD::MacroBrinkmanCollideNaturalConvection(pf, rho, ux, uy, alpha, nu, pg, tem, qx, qy, diffusivity, gx, gy, tem0, true, gi);
            if (t%5 == 0) residual = Residual(ux, uy, uxp, uyp, pf.nxyz);
            pf.Stream();
            pg.Stream();
            pf.BoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 1; });
            AD::BoundaryConditionSetT(pg, [=](int _i, int _j) { return tem0; }, ux, uy, [=](int _i, int _j) { return _i == lx - 1 || _j == ly - 1; });
            AD::BoundaryConditionSetQ(pg, [=](int _i, int _j) { return (_j == 0 && _i < L) ? qn0 : 0.0; }, ux, uy, diffusivity, [=](int _i, int _j) { return _j == 0; });
            pg.BoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 0; });
            pf.SmoothCorner();
            pg.SmoothCorner();
            std::swap(ux, uxp); std::swap(uy, uyp); std::swap(qx, qxp); std::swap(qy, qyp);
        }
        plh_sync(); t1 = clk::now();
        ANS::InitialCondition(pf, ux, uy, irho, iux, iuy);
        AAD::InitialCondition(pg, ux, uy, item, iqx, iqy);
        plh_sync(); t2 = clk::now();
        for (int t = 1; t <= nt; t++) {
            AAD::MacroBrinkmanCollideNaturalConvection(pf, rho, ux, uy, irho, iux, iuy, imx, imy, alpha, nu, pg, tem, item, iqx, iqy, diffusivity, gx, gy, true, igi);
            pf.iStream();
            pg.iStream();
            AAD::iBoundaryConditionSetT(pg, ux, uy, [=](int _i, int _j) { return _i == lx - 1 || _j == ly - 1; });
            AAD::iBoundaryConditionSetQ(pg, ux, uy, [=](int _i, int _j) { return _j == 0; });
            AAD::iBoundaryConditionSetQ(pg, ux, uy, [=](int _i, int _j) { return _j == 0 && _i < L; }, 1.0);
            pg.iBoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 0; });
            pf.iBoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 1; });
            pf.SmoothCorner();
            pg.SmoothCorner();
            std::swap(iux, iuxp); std::swap(iuy, iuyp); std::swap(iqx, iqxp); std::swap(iqy, iqyp);
        }
        plh_sync(); t3 = clk::now();
        double f_buffer = 0.0;
        for (int i = 0; i < pf.nx; ++i) if (i < L) f_buffer += tem[pf.Index(i, 0)];
        std::vector<double> dfdss(n, 0.0);
        AAD::SensitivityTemperatureAtHeatSource(pg, dfdss.data(), ux, uy, imx, imy, dads, tem, item, iqx, iqy, gi, igi, diffusivity, dkds,
            [=](int _i, int _j) { return (_j == 0 && _i < L) ? qn0 : 0.0; }, [=](int _i, int _j) { return _j == 0 && _i < L; });
        const char* names[] = {"rho", "ux", "uy", "tem", "qx", "qy", "ip", "iux", "iuy", "imx", "imy", "item", "iqx", "iqy"};
        double* arrs[] = {rho, ux, uy, tem, qx, qy, irho, iux, iuy, imx, imy, item, iqx, iqy};
        for (int a = 0; a < 14; ++a) wr(names[a], arrs[a], n);
        wr("dfdss", dfdss.data(), n);
        wr("gsnap", gi, (size_t)n*pg.nc); wr("igsnap", igi, (size_t)n*pg.nc);
        wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
        double extra[2] = {f_buffer, residual};
        wr("extra", extra, 2);
    }
    {
        const double fs = std::chrono::duration<double>(t1 - t0).count(), as = std::chrono::duration<double>(t3 - t2).count();
        const double sites = (double)lx*ly*lz;
        printf("forward %d steps %.3f ms/step %.1f MLUPS | adjoint %.3f ms/step %.1f MLUPS\n", nt, 1e3*fs/nt, sites*nt/fs/1e6, 1e3*as/nt, sites*nt/as/1e6);
    }
    uint64_t st[8];
    plh_stats(st);
    double std_[8];
    for (int k = 0; k < 8; ++k) std_[k] = (double)st[k];
    wr("stats", std_, 8);
    printf("fused steps %llu, calls one by one %llu, uploads %llu, downloads %llu, faults %llu, plans %llu, settles %llu, stagings %llu\n",
           (unsigned long long)st[0], (unsigned long long)st[1], (unsigned long long)st[2], (unsigned long long)st[3], (unsigned long long)st[4],
           (unsigned long long)st[5], (unsigned long long)st[6], (unsigned long long)st[7]);
    return 0;
}
