// Parity program for the time-periodic natural-convection pump: production/ncpump_periodic.cpp:107-305 written against the
// reference API exactly as that driver does.  What it adds to the steady pump (tests/dropin/ncpump_dump.cpp) and to the transient
// heatsink (tests/dropin/transient_dump.cpp):
//   * a run-in loop of nt0 steps on the arrays of step 0 (:123-170, first optimisation iteration only), then one stored period:
//     rho[t], ux[t], uy[t], tem[t], gi[t] per step, qx / qy shared (:177-228);
//   * a wall temperature that changes EVERY step — the SetT value lambda captures the loop counter, tembc(t) = Th(1 - cos(2 pi t /
//     period)) (:71, :137-142, :190-195);
//   * the objective read on the HOST from ux[t] / uy[t] right after every forward step (:217-227);
//   * an adjoint loop that REWRITES directionxt / directionyt on the host before every step from f[t] (:251-254), runs the
//     MassFlow adjoint collide and AAD::SensitivityBrinkmanDiffusivity every step (:256-262), masks dfds and calls Normalize (:291-297);
//   * later optimisation iterations restart from the last stored step by a host copy loop (:172-175).
// Design: closed form instead of the MMA variable.  Built twice from this one source: reference headers -> fixtures
// (tests/golden/make_ncpump_periodic_golden.py), drop-in headers -> test (tests/test_gpu_ncpump.py).
//   ncpump_periodic_dump <lx> <ly> <nt0> <nt> <nk> <dir>        writes <dir>/*.out
#define _USE_MATH_DEFINES
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
    if (argc != 7) { fprintf(stderr, "usage: ncpump_periodic_dump lx ly nt0 nt nk dir\n"); return 2; }
    const int lx = atoi(argv[1]), ly = atoi(argv[2]), nt0 = atoi(argv[3]), nt = atoi(argv[4]), nk = atoi(argv[5]);
    const int period = nt;
    dir = argv[6];
    double viscosity = 0.1/6.0, diff_fluid = viscosity/1.0, Th = 1.0, Tl = 0.0, gx = 0.0, gy = 1000*pow(viscosity, 2)/(double)pow(lx - 1, 3);
    double alphamax = 1e5, diff_solid = diff_fluid*10.0, qf = 1e-6, qg = 1e-4, ratio = 0.5;
    D2Q9<double> pf(lx, ly), pg(lx, ly);
    const int n = pf.nxyz;
    double **rho = new double*[nt], **ux = new double*[nt], **uy = new double*[nt];
    double **tem = new double*[nt], *qx = new double[n], *qy = new double[n];
    double **gi = new double*[nt];
    double *f = new double[nt];
    f[0] = 0.0;
    for (int t = 0; t < nt; ++t) {
        rho[t] = new double[n];   ux[t] = new double[n];    uy[t] = new double[n];
        tem[t] = new double[n];   gi[t] = new double[n*pg.nc];
    }
    double *irho = new double[n], *iux = new double[n], *iuy = new double[n], *imx = new double[n], *imy = new double[n];
    double *item = new double[n], *iqx = new double[n], *iqy = new double[n];
    for (int idx = 0; idx < n; idx++) {
        rho[0][idx] = 1.0; ux[0][idx] = 0.0; uy[0][idx] = 0.0; tem[0][idx] = 0.5*(Tl + Th); qx[idx] = 0.0; qy[idx] = 0.0;
        irho[idx] = 1.0; iux[idx] = 0.0; iuy[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0; item[idx] = 0.0; iqx[idx] = 0.0; iqy[idx] = 0.0;
    }
    double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
    double *igi = new double[n*pg.nc];
    double *directionx = new double[n], *directiony = new double[n], *directionxt = new double[n], *directionyt = new double[n];
    std::vector<double> s(n, 1.0);
    for (int i = 0; i < pf.nx; ++i) {
        for (int j = 0; j < pf.ny; ++j) {
            int idx = pf.Index(i, j);
            directionx[idx] = ((i + pf.offsetx) == lx/2 && (j + pf.offsety) > 9*ly/10) ? -1.0 : 0.0;
            directiony[idx] = 0.0;
            s[idx] = j < ly/2 ? 0.5 + 0.4*sin(0.37*i)*cos(0.23*j) : 1.0;
        }
    }
    auto tembc = [=](int _t) { return Th*(1 - cos(2*M_PI*_t/period)); };
    typedef std::chrono::steady_clock clk;
    double fwd_s = 0.0, adj_s = 0.0, F = 0.0, faverage = 0.0, variance = 0.0;
    std::vector<double> dfds(n, 0.0), dfds_raw(n, 0.0);

    for (int k = 1; k <= nk; k++) {
        for (int idx = 0; idx < n; idx++) {
            diffusivity[idx] = diff_solid + (diff_fluid - diff_solid)*s[idx]*(1.0 + qg)/(s[idx] + qg);
            alpha[idx] = alphamax/(double)(ly - 1)*qf*(1.0 - s[idx])/(s[idx] + qf);
            dkds[idx] = (diff_fluid - diff_solid)*qg*(1.0 + qg)/pow(s[idx] + qg, 2.0);
            dads[idx] = -alphamax/(double)(ly - 1)*qf*(1.0 + qf)/pow(s[idx] + qf, 2.0);
        }

        //********************Direct analyze********************
        if (k == 1) {
            NS::InitialCondition(pf, rho[0], ux[0], uy[0]);
            AD::InitialCondition(pg, tem[0], ux[0], uy[0]);
            for (int t = 1; t < nt0; ++t) {
                AD::MacroBrinkmanCollideNaturalConvection(
                    pf, rho[0], ux[0], uy[0], alpha, viscosity,
                    pg, tem[0], qx, qy, diffusivity, gx, gy, 0.5*(Th + Tl), true, gi[0]
                );

                pf.Stream();
                pg.Stream();
                pf.BoundaryCondition([=](int _i, int _j) { return 1; });
                pg.BoundaryCondition([=](int _i, int _j) { return 0; });
                AD::BoundaryConditionSetT(pg,
                    [=](int _i, int _j) { return _i == 0 ? tembc(t) : Tl; },
                    ux[0], uy[0],
                    [=](int _i, int _j) { return (_i == 0 && _j < ly/2) || (_i == lx - 1 && _j < ly/2); }
                );
                AD::BoundaryConditionSetQ(pg,
                    [=](int _i, int _j) { return 0.0; },
                    ux[0], uy[0], diffusivity,
                    [=](int _i, int _j) { return (_i == 0 && ly/2 <= _j) || (_i == lx - 1 && ly/2 <= _j) || _j == 0 || _j == ly - 1; }
                );
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
                AD::BoundaryConditionSetQAlongXEdge(pg, lx/5, 1, [=](int _i, int _j) { return 0.0; }, ux[0], uy[0], diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
                AD::BoundaryConditionSetQAlongYEdge(pg, ly/2, 1, [=](int _i, int _j) { return 0.0; }, ux[0], uy[0], diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
                AD::BoundaryConditionSetQAlongXEdge(pg, 4*lx/5, -1, [=](int _i, int _j) { return 0.0; }, ux[0], uy[0], diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
                AD::BoundaryConditionSetQAlongYEdge(pg, 9*ly/10, 1, [=](int _i, int _j) { return 0.0; }, ux[0], uy[0], diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
                pg.SmoothCornerAt(lx/5, ly/2, -1, -1);
                pg.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
                pg.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
                pg.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);
            }
        } else {
            for (int idx = 0; idx < n; idx++) {
                rho[0][idx] = rho[nt - 1][idx]; ux[0][idx] = ux[nt - 1][idx]; uy[0][idx] = uy[nt - 1][idx]; tem[0][idx] = tem[nt - 1][idx];
            }
        }
        double faverage_buffer = 0.0, fsquare_buffer = 0.0;
        NS::InitialCondition(pf, rho[0], ux[0], uy[0]);
        AD::InitialCondition(pg, tem[0], ux[0], uy[0]);
        clk::time_point t0 = clk::now();
        for (int t = 1; t < nt; ++t) {
            AD::MacroBrinkmanCollideNaturalConvection(
                pf, rho[t], ux[t], uy[t], alpha, viscosity,
                pg, tem[t], qx, qy, diffusivity, gx, gy, 0.5*(Th + Tl), true, gi[t]
            );

            pf.Stream();
            pg.Stream();
            pf.BoundaryCondition([=](int _i, int _j) { return 1; });
            pg.BoundaryCondition([=](int _i, int _j) { return 0; });
            AD::BoundaryConditionSetT(pg,
                [=](int _i, int _j) { return _i == 0 ? tembc(t) : Tl; },
                ux[t], uy[t],
                [=](int _i, int _j) { return (_i == 0 && _j < ly/2) || (_i == lx - 1 && _j < ly/2); }
            );
            AD::BoundaryConditionSetQ(pg,
                [=](int _i, int _j) { return 0.0; },
                ux[t], uy[t], diffusivity,
                [=](int _i, int _j) { return (_i == 0 && ly/2 <= _j) || (_i == lx - 1 && ly/2 <= _j) || _j == 0 || _j == ly - 1; }
            );
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
            AD::BoundaryConditionSetQAlongXEdge(pg, lx/5, 1, [=](int _i, int _j) { return 0.0; }, ux[t], uy[t], diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
            AD::BoundaryConditionSetQAlongYEdge(pg, ly/2, 1, [=](int _i, int _j) { return 0.0; }, ux[t], uy[t], diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
            AD::BoundaryConditionSetQAlongXEdge(pg, 4*lx/5, -1, [=](int _i, int _j) { return 0.0; }, ux[t], uy[t], diffusivity, [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
            AD::BoundaryConditionSetQAlongYEdge(pg, 9*ly/10, 1, [=](int _i, int _j) { return 0.0; }, ux[t], uy[t], diffusivity, [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
            pg.SmoothCornerAt(lx/5, ly/2, -1, -1);
            pg.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
            pg.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
            pg.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);

            f[t] = 0.0;
            for (int j = 0; j < pf.ny; ++j) {
                int i = lx/2 - pf.offsetx;
                if (0 <= i && i < pf.nx && (j + pf.offsety) > 9*ly/10) {
                    int idx = pf.Index(i, j);
                    f[t] += ux[t][idx]*directionx[idx] + uy[t][idx]*directiony[idx];
                }
            }
            f[t] /= (double)((ly - 1)/10);
            faverage_buffer += f[t];
            fsquare_buffer += pow(f[t], 2.0);
        }
#ifdef PANSLBM_B200_DROPIN
        plh_sync();
#endif
        clk::time_point t1 = clk::now();
        faverage = faverage_buffer/(double)nt;
        variance = fsquare_buffer/(double)nt - pow(faverage, 2.0);
        double coef = (1.0 - ratio)/sqrt(variance);
        F = ratio*faverage + (1.0 - ratio)*sqrt(variance);

        //********************Inverse analyze********************
        dfds.assign(n, 0.0);
        ANS::InitialCondition(pf, ux[nt - 1], uy[nt - 1], irho, iux, iuy);
        AAD::InitialCondition(pg, ux[nt - 1], uy[nt - 1], item, iqx, iqy);
        clk::time_point t2 = clk::now();
        for (int t = nt - 2; t >= 0; --t) {
            for (int idx = 0; idx < n; ++idx) {
                directionxt[idx] = (ratio + coef*(f[t] - faverage))*directionx[idx];
                directionyt[idx] = (ratio + coef*(f[t] - faverage))*directiony[idx];
            }

            AAD::MacroBrinkmanCollideNaturalConvectionMassFlow(
                pf, rho[t], ux[t], uy[t], irho, iux, iuy, imx, imy, alpha, viscosity,
                pg, tem[t], item, iqx, iqy, diffusivity, gx, gy,
                directionxt, directionyt, true, igi
            );

            AAD::SensitivityBrinkmanDiffusivity(pg, dfds.data(), ux[t], uy[t], imx, imy, dads, tem[t], item, iqx, iqy, gi[t], igi, diffusivity, dkds);

            pf.iStream();
            pg.iStream();
            pf.iBoundaryCondition([=](int _i, int _j) { return 1; });
            pg.iBoundaryCondition([=](int _i, int _j) { return 0; });
            AAD::iBoundaryConditionSetT(pg, ux[t], uy[t], [=](int _i, int _j) { return (_i == 0 && _j < ly/2) || (_i == lx - 1 && _j < ly/2); });
            AAD::iBoundaryConditionSetQ(pg, ux[t], uy[t], [=](int _i, int _j) { return (_i == 0 && ly/2 <= _j) || (_i == lx - 1 && ly/2 <= _j) || _j == 0 || _j == ly - 1; });
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
            AAD::iBoundaryConditionSetQAlongXEdge(pg, lx/5, 1, ux[t], uy[t], [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
            AAD::iBoundaryConditionSetQAlongYEdge(pg, ly/2, 1, ux[t], uy[t], [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
            AAD::iBoundaryConditionSetQAlongXEdge(pg, 4*lx/5, -1, ux[t], uy[t], [=](int _i, int _j) { return ly/2 <= _j && _j < 9*ly/10; });
            AAD::iBoundaryConditionSetQAlongYEdge(pg, 9*ly/10, 1, ux[t], uy[t], [=](int _i, int _j) { return lx/5 <= _i && _i < 4*lx/5; });
            pg.SmoothCornerAt(lx/5, ly/2, -1, -1);
            pg.SmoothCornerAt(4*lx/5, ly/2, 1, -1);
            pg.SmoothCornerAt(4*lx/5, 9*ly/10, 1, 1);
            pg.SmoothCornerAt(lx/5, 9*ly/10, -1, 1);
        }
#ifdef PANSLBM_B200_DROPIN
        plh_sync();
#endif
        clk::time_point t3 = clk::now();
        fwd_s = std::chrono::duration<double>(t1 - t0).count();
        adj_s = std::chrono::duration<double>(t3 - t2).count();
        for (int i = 0; i < pf.nx; ++i) {
            for (int j = 0; j < pf.ny; ++j) {
                int idx = pf.Index(i, j);
                dfds[idx] = (j + pf.offsety) < ly/2 ? dfds[idx] : 0.0;
            }
        }
        dfds_raw = dfds;
        Normalize(dfds.data(), pg.nxyz);
        // stands in for the MMA update of the driver (:314): a closed-form move of the design along the sensitivity
        for (int idx = 0; idx < n; ++idx) s[idx] = std::min(1.0, std::max(0.0, s[idx] - 0.05*dfds[idx]));
        for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) { int idx = pf.Index(i, j); s[idx] = (j + pf.offsety) < ly/2 ? s[idx] : 1.0; }
    }

    const int tm = nt/2;
    wr("rho_last", rho[nt - 1], n); wr("ux_last", ux[nt - 1], n); wr("uy_last", uy[nt - 1], n); wr("tem_last", tem[nt - 1], n);
    wr("rho_mid", rho[tm], n); wr("ux_mid", ux[tm], n); wr("uy_mid", uy[tm], n); wr("tem_mid", tem[tm], n);
    wr("rho_0", rho[0], n); wr("tem_0", tem[0], n);
    wr("qx", qx, n); wr("qy", qy, n);
    const char* names[] = {"ip", "iux", "iuy", "imx", "imy", "item", "iqx", "iqy"};
    double* arrs[] = {irho, iux, iuy, imx, imy, item, iqx, iqy};
    for (int a = 0; a < 8; ++a) wr(names[a], arrs[a], n);
    wr("dfds_raw", dfds_raw.data(), n);
    wr("dfds", dfds.data(), n);
    wr("s", s.data(), n);
    wr("fobj", f, nt);
    wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
    double extra[3] = {F, faverage, variance};
    wr("extra", extra, 3);
    printf("forward %d steps %.4f ms/step | adjoint %.4f ms/step (last iteration)\n", nt - 1, 1e3*fwd_s/(nt - 1), 1e3*adj_s/(nt - 1));
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
