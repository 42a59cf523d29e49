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

// Geometry of the pump (ncpump_periodic.cpp:129-168): heated / cooled lower halves of the side walls, adiabatic rest, and a solid
// block [lx/5, 4lx/5) x [ly/2, 9ly/10) whose four edges carry bounce-back (flow) and zero flux (heat), corners smoothed.
struct Pump {
    int lx, ly;
    int x0() const { return lx/5; }
    int x1() const { return 4*lx/5; }
    int y0() const { return ly/2; }
    int y1() const { return 9*ly/10; }
    // the four block edges in the driver's order: (x0, +1), (y0, +1), (x1, -1), (y1, +1); fx / fy get (coordinate, direction)
    template<class FX, class FY> void edges(FX fx, FY fy) const { fx(x0(), 1); fy(y0(), 1); fx(x1(), -1); fy(y1(), 1); }
    // the four block corners in the driver's order
    template<class L> void corners(L& l) const {
        l.SmoothCornerAt(x0(), y0(), -1, -1); l.SmoothCornerAt(x1(), y0(), 1, -1); l.SmoothCornerAt(x1(), y1(), 1, 1); l.SmoothCornerAt(x0(), y1(), -1, 1);
    }
};

// everything between the two Streams and the end of a forward step (:133-168 = :186-221); twall(i, j) = wall temperature of this step
template<class TW>
static void forward_closures(const Pump& G, D2Q9<double>& pf, D2Q9<double>& pg, TW twall, double* ux, double* uy, const double* diffusivity) {
    const int lx = G.lx, ly = G.ly;
    auto along_y = [=](int, int _j) { return G.y0() <= _j && _j < G.y1(); };      // sites of a vertical block edge
    auto along_x = [=](int _i, int) { return G.x0() <= _i && _i < G.x1(); };      // sites of a horizontal block edge
    auto noflux = [=](int, int) { return 0.0; };
    pf.BoundaryCondition([=](int, int) { return 1; });
    pg.BoundaryCondition([=](int, int) { return 0; });
    AD::BoundaryConditionSetT(pg, twall, ux, uy, [=](int _i, int _j) { return (_i == 0 || _i == lx - 1) && _j < ly/2; });
    AD::BoundaryConditionSetQ(pg, noflux, ux, uy, diffusivity, [=](int _i, int _j) { return ((_i == 0 || _i == lx - 1) && ly/2 <= _j) || _j == 0 || _j == ly - 1; });
    pf.SmoothCorner();
    pg.SmoothCorner();
    G.edges([&](int c, int d) { pf.BoundaryConditionAlongXEdge(c, d, along_y); }, [&](int c, int d) { pf.BoundaryConditionAlongYEdge(c, d, along_x); });
    G.corners(pf);
    G.edges([&](int c, int d) { AD::BoundaryConditionSetQAlongXEdge(pg, c, d, noflux, ux, uy, diffusivity, along_y); },
            [&](int c, int d) { AD::BoundaryConditionSetQAlongYEdge(pg, c, d, noflux, ux, uy, diffusivity, along_x); });
    G.corners(pg);
}

// the same for an adjoint step (:265-290)
static void adjoint_closures(const Pump& G, D2Q9<double>& pf, D2Q9<double>& pg, const double* ux, const double* uy) {
    const int lx = G.lx, ly = G.ly;
    auto along_y = [=](int, int _j) { return G.y0() <= _j && _j < G.y1(); };
    auto along_x = [=](int _i, int) { return G.x0() <= _i && _i < G.x1(); };
    pf.iBoundaryCondition([=](int, int) { return 1; });
    pg.iBoundaryCondition([=](int, int) { return 0; });
    AAD::iBoundaryConditionSetT(pg, ux, uy, [=](int _i, int _j) { return (_i == 0 || _i == lx - 1) && _j < ly/2; });
    AAD::iBoundaryConditionSetQ(pg, ux, uy, [=](int _i, int _j) { return ((_i == 0 || _i == lx - 1) && ly/2 <= _j) || _j == 0 || _j == ly - 1; });
    pf.SmoothCorner();
    pg.SmoothCorner();
    G.edges([&](int c, int d) { pf.iBoundaryConditionAlongXEdge(c, d, along_y); }, [&](int c, int d) { pf.iBoundaryConditionAlongYEdge(c, d, along_x); });
    G.corners(pf);
    G.edges([&](int c, int d) { AAD::iBoundaryConditionSetQAlongXEdge(pg, c, d, ux, uy, along_y); }, [&](int c, int d) { AAD::iBoundaryConditionSetQAlongYEdge(pg, c, d, ux, uy, along_x); });
    G.corners(pg);
}

int main(int argc, char** argv) {
    if (argc != 7) { fprintf(stderr, "usage: ncpump_periodic_dump lx ly nt0 nt nk dir\n"); return 2; }
    const int lx = atoi(argv[1]), ly = atoi(argv[2]), nt0 = atoi(argv[3]), nt = atoi(argv[4]), nk = atoi(argv[5]);
    const int period = nt;
    dir = argv[6];
    const Pump G{lx, ly};
    const double viscosity = 0.1/6.0, diff_fluid = viscosity/1.0, Th = 1.0, Tl = 0.0, gx = 0.0, gy = 1000*pow(viscosity, 2)/(double)pow(lx - 1, 3);
    const double alphamax = 1e5, diff_solid = diff_fluid*10.0, qf = 1e-6, qg = 1e-4, ratio = 0.5, tem_ref = 0.5*(Th + Tl);
    D2Q9<double> pf(lx, ly), pg(lx, ly);
    const int n = pf.nxyz;
    // one set of arrays per stored step (:42-52); qx, qy and every adjoint field are shared
    std::vector<double*> rho(nt), ux(nt), uy(nt), tem(nt), gi(nt);
    for (int t = 0; t < nt; ++t) { rho[t] = new double[n]; ux[t] = new double[n]; uy[t] = new double[n]; tem[t] = new double[n]; gi[t] = new double[n*pg.nc]; }
    double *qx = new double[n], *qy = new double[n], *f = new double[nt];
    double *irho = new double[n], *iux = new double[n], *iuy = new double[n], *imx = new double[n], *imy = new double[n], *item = new double[n], *iqx = new double[n], *iqy = new double[n];
    double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n], *igi = new double[n*pg.nc];
    double *directionx = new double[n], *directiony = new double[n], *directionxt = new double[n], *directionyt = new double[n];
    f[0] = 0.0;
    for (int idx = 0; idx < n; idx++) {
        rho[0][idx] = 1.0; ux[0][idx] = 0.0; uy[0][idx] = 0.0; tem[0][idx] = tem_ref; qx[idx] = 0.0; qy[idx] = 0.0;
        irho[idx] = 1.0; iux[idx] = 0.0; iuy[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0; item[idx] = 0.0; iqx[idx] = 0.0; iqy[idx] = 0.0;
    }
    std::vector<double> s(n, 1.0);
    for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) {
        const int idx = pf.Index(i, j);
        directionx[idx] = (i == lx/2 && j > G.y1()) ? -1.0 : 0.0;       // the mass flow through the gap above the block (:59-65)
        directiony[idx] = 0.0;
        s[idx] = j < ly/2 ? 0.5 + 0.4*sin(0.37*i)*cos(0.23*j) : 1.0;    // closed-form design in the lower half (stands in for the MMA variable)
    }
    auto tembc = [=](int _t) { return Th*(1 - cos(2*M_PI*_t/period)); };
    typedef std::chrono::steady_clock clk;
    double fwd_s = 0.0, adj_s = 0.0, F = 0.0, faverage = 0.0, variance = 0.0;
    std::vector<double> dfds(n, 0.0), dfds_raw(n, 0.0);

    // one forward step on the arrays of slot `slot` with the wall temperature of time `t`
    auto forward_step = [&](int slot, int t) {
        AD::MacroBrinkmanCollideNaturalConvection(pf, rho[slot], ux[slot], uy[slot], alpha, viscosity, pg, tem[slot], qx, qy, diffusivity, gx, gy, tem_ref, true, gi[slot]);
        pf.Stream();
        pg.Stream();
        forward_closures(G, pf, pg, [=](int _i, int) { return _i == 0 ? tembc(t) : Tl; }, ux[slot], uy[slot], diffusivity);
    };

    for (int k = 1; k <= nk; k++) {
        for (int idx = 0; idx < n; idx++) {     // :96-101
            diffusivity[idx] = diff_solid + (diff_fluid - diff_solid)*s[idx]*(1.0 + qg)/(s[idx] + qg);
            alpha[idx] = alphamax/(double)(ly - 1)*qf*(1.0 - s[idx])/(s[idx] + qf);
            dkds[idx] = (diff_fluid - diff_solid)*qg*(1.0 + qg)/pow(s[idx] + qg, 2.0);
            dads[idx] = -alphamax/(double)(ly - 1)*qf*(1.0 + qf)/pow(s[idx] + qf, 2.0);
        }
        // ---- direct analysis: run-in on slot 0 in the first iteration (:123-170), else restart from the last stored step (:172-175)
        if (k == 1) {
            NS::InitialCondition(pf, rho[0], ux[0], uy[0]);
            AD::InitialCondition(pg, tem[0], ux[0], uy[0]);
            for (int t = 1; t < nt0; ++t) forward_step(0, t);
        } else {
            for (int idx = 0; idx < n; idx++) { rho[0][idx] = rho[nt - 1][idx]; ux[0][idx] = ux[nt - 1][idx]; uy[0][idx] = uy[nt - 1][idx]; tem[0][idx] = tem[nt - 1][idx]; }
        }
        double fsum = 0.0, fsq = 0.0;
        NS::InitialCondition(pf, rho[0], ux[0], uy[0]);
        AD::InitialCondition(pg, tem[0], ux[0], uy[0]);
        clk::time_point t0 = clk::now();
        for (int t = 1; t < nt; ++t) {          // one stored period (:177-228)
            forward_step(t, t);
            f[t] = 0.0;                         // the objective of this step, read on the host (:217-227)
            for (int j = G.y1() + 1; j < pf.ny; ++j) { const int idx = pf.Index(lx/2, j); f[t] += ux[t][idx]*directionx[idx] + uy[t][idx]*directiony[idx]; }
            f[t] /= (double)((ly - 1)/10);
            fsum += f[t];
            fsq += pow(f[t], 2.0);
        }
#ifdef PANSLBM_B200_DROPIN
        plh_sync();
#endif
        clk::time_point t1 = clk::now();
        faverage = fsum/(double)nt;
        variance = fsq/(double)nt - pow(faverage, 2.0);
        const double coef = (1.0 - ratio)/sqrt(variance);
        F = ratio*faverage + (1.0 - ratio)*sqrt(variance);

        // ---- inverse analysis (:243-290): backwards over the stored period, sensitivity accumulated every step
        dfds.assign(n, 0.0);
        ANS::InitialCondition(pf, ux[nt - 1], uy[nt - 1], irho, iux, iuy);
        AAD::InitialCondition(pg, ux[nt - 1], uy[nt - 1], item, iqx, iqy);
        clk::time_point t2 = clk::now();
        for (int t = nt - 2; t >= 0; --t) {
            const double w = ratio + coef*(f[t] - faverage);        // the direction fields are rewritten on the host every step (:251-254)
            for (int idx = 0; idx < n; ++idx) { directionxt[idx] = w*directionx[idx]; directionyt[idx] = w*directiony[idx]; }
            AAD::MacroBrinkmanCollideNaturalConvectionMassFlow(pf, rho[t], ux[t], uy[t], irho, iux, iuy, imx, imy, alpha, viscosity,
                                                               pg, tem[t], item, iqx, iqy, diffusivity, gx, gy, directionxt, directionyt, true, igi);
            AAD::SensitivityBrinkmanDiffusivity(pg, dfds.data(), ux[t], uy[t], imx, imy, dads, tem[t], item, iqx, iqy, gi[t], igi, diffusivity, dkds);
            pf.iStream();
            pg.iStream();
            adjoint_closures(G, pf, pg, ux[t], uy[t]);
        }
#ifdef PANSLBM_B200_DROPIN
        plh_sync();
#endif
        clk::time_point t3 = clk::now();
        fwd_s = std::chrono::duration<double>(t1 - t0).count();
        adj_s = std::chrono::duration<double>(t3 - t2).count();
        for (int i = 0; i < pf.nx; ++i) for (int j = ly/2; j < pf.ny; ++j) dfds[pf.Index(i, j)] = 0.0;      // design region = lower half (:291-296)
        dfds_raw = dfds;
        Normalize(dfds.data(), pg.nxyz);
        // stands in for the MMA update of the driver (:314): a closed-form move of the design along the sensitivity
        for (int i = 0; i < pf.nx; ++i) for (int j = 0; j < pf.ny; ++j) {
            const int idx = pf.Index(i, j);
            s[idx] = j < ly/2 ? std::min(1.0, std::max(0.0, s[idx] - 0.05*dfds[idx])) : 1.0;
        }
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
