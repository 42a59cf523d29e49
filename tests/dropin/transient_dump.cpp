// Parity program for the transient heatsink loops (BASELINE configs[4]): production/heatsink3D_transient.cpp:145-232 and
// production/heatsink_transient.cpp:136-215 written against the reference API exactly as those drivers do — one set of
// macroscopic arrays and one thermal snapshot PER TIME STEP (rho[t], ux[t], ..., gi[t]), the adjoint loop walking them
// backwards and accumulating AAD::SensitivityTemperatureAtHeatSource every step, the objective summed over tem[t] afterwards —
// with a small nt and the design of tests/heatsink_case.py.  Built twice from this one source:
//   * with -I<reference>/src and -fopenmp            -> the fixture generator (tests/golden/make_transient_golden.py)
//   * with -I panslbm2_b200/src (drop-in headers)    -> the program under test (tests/test_gpu_transient.py)
// -DTRANSIENT_DIM=2|3 selects the lattice at compile time.
//   transient_dump <dim> <lx> <ly> <lz> <nt> <dir> [iterations]   reads <dir>/{alpha,kappa,dads,dkds}.bin, <dir>/params.bin; writes <dir>/*.out
// iterations > 1 repeats the forward + adjoint loops over the same arrays as the optimisation loop of the drivers does
// (heatsink3D_transient.cpp:97, nitr = 500); the printed timings are those of the last repetition (the first one also pays for
// the first-touch allocation of the per-step arrays).
#define _USE_AVX_DEFINES
// Multi-rank runs (drop-in build only; the container has no MPI for the reference): TRANSIENT_PE="mx,my,mz" with one process per
// GPU under tools/mpiexec_b200 decomposes the 3-D lattice as the MPI build of production/heatsink3D_transient.cpp:28-48 does
// (rank = PEid, MPI_Allreduce of the objective); every rank reads the GLOBAL input fields, takes its block, and writes its
// block of every output as <name>.r<rank>.out plus block.r<rank>.out = (offsets, extents).  The single-rank fixture is the oracle.
#if defined(PANSLBM_B200_DROPIN) && defined(TRANSIENT_MPI)
#define _USE_MPI_DEFINES
#include "mpi/mpi.h"
#endif
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#ifndef TRANSIENT_DIM
#define TRANSIENT_DIM 3      // the reference's d2q9.h and d3q15.h cannot share a translation unit (both define BARRIER/MIRROR)
#endif
#if TRANSIENT_DIM == 3
#include "particle/d3q15.h"
#else
#include "particle/d2q9.h"
#endif
#include "equation/advection.h"
#include "equation/adjointadvection.h"

using namespace PANSLBM2;

static std::string dir;
static void rd(const char* name, double* p, size_t n) {
    FILE* f = fopen((dir + "/" + name).c_str(), "rb");
    if (!f || fread(p, sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", name); exit(2); }
    fclose(f);
}
static std::string rank_suffix;
static void wr(const std::string& name, const double* p, size_t n) {
    volatile double first = n ? p[0] : 0.0;     // user-space touch first (see tests/dropin/heatsink_dump.cpp)
    (void)first;
    FILE* f = fopen((dir + "/" + name + rank_suffix + ".out").c_str(), "wb");
    fwrite(p, sizeof(double), n, f);
    fclose(f);
}
static double* filled(int n, double v) { double* p = new double[n]; for (int i = 0; i < n; ++i) p[i] = v; return p; }
static void sync_device() {
#ifdef PANSLBM_B200_DROPIN
    plh_sync();
#endif
}

int main(int argc, char** argv) {
    if (argc != 7 && argc != 8) { fprintf(stderr, "usage: transient_dump dim lx ly lz nt dir [iterations]\n"); return 2; }
    const int iterations = argc == 8 ? atoi(argv[7]) : 1;
    const int dim = atoi(argv[1]), lx = atoi(argv[2]), ly = atoi(argv[3]), lz = atoi(argv[4]), nt = atoi(argv[5]);
    dir = argv[6];
    double prm[7];
    rd("params.bin", prm, 7);
    const double nu = prm[0], gx = prm[1], gy = prm[2], gz = prm[3], tem0 = prm[4], qn = prm[5], L = prm[6];
    typedef std::chrono::steady_clock clk;
    clk::time_point t0, t1, t2, t3;
    double f_buffer = 0.0;

    if (dim != TRANSIENT_DIM) { fprintf(stderr, "built for dim %d\n", TRANSIENT_DIM); return 2; }
#if TRANSIENT_DIM == 3
    {
        int MyRank = 0, pex = 1, pey = 1, pez = 1;
#ifdef _USE_MPI_DEFINES
        {
            int PeTot;
            MPI_Init(&argc, &argv);
            MPI_Comm_size(MPI_COMM_WORLD, &PeTot);
            MPI_Comm_rank(MPI_COMM_WORLD, &MyRank);
            const char* pe = getenv("TRANSIENT_PE");
            if (!pe || sscanf(pe, "%d,%d,%d", &pex, &pey, &pez) != 3 || pex*pey*pez != PeTot) { fprintf(stderr, "TRANSIENT_PE=mx,my,mz must match the number of ranks\n"); return 2; }
            if (PeTot > 1) rank_suffix = ".r" + std::to_string(MyRank);
        }
#endif
        D3Q15<double> pf(lx, ly, lz, MyRank, pex, pey, pez), pg(lx, ly, lz, MyRank, pex, pey, pez);
        const int n = pf.nxyz;
        double **rho = new double*[nt], **ux = new double*[nt], **uy = new double*[nt], **uz = new double*[nt];
        double **tem = new double*[nt], **qx = new double*[nt], **qy = new double*[nt], **qz = new double*[nt];
        double **gi = new double*[nt];
        for (int t = 0; t < nt; ++t) {
            rho[t] = new double[n]; ux[t] = new double[n]; uy[t] = new double[n]; uz[t] = new double[n];
            tem[t] = new double[n]; qx[t] = new double[n]; qy[t] = new double[n]; qz[t] = new double[n];
            gi[t] = new double[n*pg.nc];
        }
        double *irho = filled(n, 0.0), *iux = filled(n, 0.0), *iuy = filled(n, 0.0), *iuz = filled(n, 0.0), *imx = filled(n, 0.0), *imy = filled(n, 0.0), *imz = filled(n, 0.0);
        double *item = filled(n, 0.0), *iqx = filled(n, 0.0), *iqy = filled(n, 0.0), *iqz = filled(n, 0.0);
        double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
        double *igi = new double[n*pg.nc];
        {
            // the input files hold the fields of the GLOBAL domain: every rank takes its block
            const size_t gn = (size_t)lx*ly*lz;
            std::vector<double> gbuf(gn);
            const char* files[4] = {"alpha.bin", "kappa.bin", "dads.bin", "dkds.bin"};
            double* dst[4] = {alpha, diffusivity, dads, dkds};
            for (int a = 0; a < 4; ++a) {
                rd(files[a], gbuf.data(), gn);
                for (int k = 0; k < pf.nz; ++k) for (int j = 0; j < pf.ny; ++j) for (int i = 0; i < pf.nx; ++i)
                    dst[a][pf.Index(i, j, k)] = gbuf[(size_t)(i + pf.offsetx) + (size_t)lx*((size_t)(j + pf.offsety) + (size_t)ly*(size_t)(k + pf.offsetz))];
            }
        }

        std::vector<double> dfdss(n, 0.0);
        for (int it = 0; it < iterations; ++it) {
        for (int idx = 0; idx < n; idx++) {
            rho[0][idx] = 1.0; ux[0][idx] = 0.0; uy[0][idx] = 0.0; uz[0][idx] = 0.0;
            tem[0][idx] = 0.0; qx[0][idx] = 0.0; qy[0][idx] = 0.0; qz[0][idx] = 0.0;
        }
        NS::InitialCondition(pf, rho[0], ux[0], uy[0], uz[0]);
        AD::InitialCondition(pg, tem[0], ux[0], uy[0], uz[0]);
        sync_device(); t0 = clk::now();
        for (int t = 1; t < nt; ++t) {
            AD::MacroBrinkmanCollideNaturalConvection(pf, rho[t], ux[t], uy[t], uz[t], alpha, nu, pg, tem[t], qx[t], qy[t], qz[t], diffusivity,
                                                      gx, gy, gz, tem0, true, gi[t]);
            pf.Stream();
            pg.Stream(30);
            pf.BoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 1; });
            AD::BoundaryConditionSetT(pg, [=](int _i, int _j, int _k) { return tem0; }, ux[t], uy[t], uz[t],
                [=](int _i, int _j, int _k) { return _i == lx - 1 || _j == ly - 1 || _k == lz - 1; });
            AD::BoundaryConditionSetQ(pg, [=](int _i, int _j, int _k) { return (_j == 0 && _i < L && _k < L) ? qn : 0.0; }, ux[t], uy[t], uz[t], diffusivity,
                [=](int _i, int _j, int _k) { return _j == 0; });
            pg.BoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 0; });
            pf.SmoothCorner();
            pg.SmoothCorner();
        }
        sync_device(); t1 = clk::now();
        if (it > 0) {       // heatsink3D_transient.cpp:180-184
            for (int idx = 0; idx < n; idx++) {
                dfdss[idx] = 0.0; irho[idx] = 0.0; iux[idx] = 0.0; iuy[idx] = 0.0; iuz[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0; imz[idx] = 0.0;
                item[idx] = 0.0; iqx[idx] = 0.0; iqy[idx] = 0.0; iqz[idx] = 0.0;
            }
        }
        ANS::InitialCondition(pf, ux[nt - 1], uy[nt - 1], uz[nt - 1], irho, iux, iuy, iuz);
        AAD::InitialCondition(pg, ux[nt - 1], uy[nt - 1], uz[nt - 1], item, iqx, iqy, iqz);
        sync_device(); t2 = clk::now();
        for (int t = nt - 2; t >= 0; --t) {
            AAD::MacroBrinkmanCollideNaturalConvection(pf, rho[t], ux[t], uy[t], uz[t], irho, iux, iuy, iuz, imx, imy, imz, alpha, nu,
                                                       pg, tem[t], item, iqx, iqy, iqz, diffusivity, gx, gy, gz, true, igi);
            AAD::SensitivityTemperatureAtHeatSource(pg, dfdss.data(), ux[t], uy[t], uz[t], imx, imy, imz, dads, tem[t], item, iqx, iqy, iqz, gi[t], igi, diffusivity, dkds,
                [=](int _i, int _j, int _k) { return (_j == 0 && _i < L && _k < L) ? qn : 0.0; },
                [=](int _i, int _j, int _k) { return _j == 0 && _i < L && _k < L; });
            pf.iStream();
            pg.iStream(30);
            AAD::iBoundaryConditionSetT(pg, ux[t], uy[t], uz[t], [=](int _i, int _j, int _k) { return _i == lx - 1 || _j == ly - 1 || _k == lz - 1; });
            AAD::iBoundaryConditionSetQ(pg, ux[t], uy[t], uz[t], [=](int _i, int _j, int _k) { return _j == 0; });
            AAD::iBoundaryConditionSetQ(pg, ux[t], uy[t], uz[t], [=](int _i, int _j, int _k) { return _j == 0 && _i < L && _k < L; }, 1.0);
            pg.iBoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 0; });
            pf.iBoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 1; });
            pf.SmoothCorner();
            pg.SmoothCorner();
        }
        sync_device(); t3 = clk::now();
        }
        // objective: heat-patch temperature summed over every stored step (heatsink3D_transient.cpp:221-231)
        // t = 0 holds the initial condition the driver wrote; steps 1..nt-1 what the collides saved
        for (int t = 0; t < nt; ++t)
            for (int i = 0; i < pf.nx; ++i) for (int k = 0; k < pf.nz; ++k) if ((i + pf.offsetx) < L && (k + pf.offsetz) < L && pf.PEy == 0) f_buffer += tem[t][pf.Index(i, 0, k)];
#ifdef _USE_MPI_DEFINES
        { double f_all = 0.0; MPI_Allreduce(&f_buffer, &f_all, 1, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD); f_buffer = f_all; }
        { double blk[6] = {(double)pf.offsetx, (double)pf.offsety, (double)pf.offsetz, (double)pf.nx, (double)pf.ny, (double)pf.nz}; wr("block", blk, 6); }
#endif
        const int tq[3] = {1, nt/2, nt - 1};
        for (int q = 0; q < 3; ++q) {
            const std::string s = "@" + std::to_string(q);
            wr("rho" + s, rho[tq[q]], n); wr("ux" + s, ux[tq[q]], n); wr("uz" + s, uz[tq[q]], n); wr("tem" + s, tem[tq[q]], n); wr("qy" + s, qy[tq[q]], n);
        }
        const char* names[] = {"ip", "iux", "iuy", "iuz", "imx", "imy", "imz", "item", "iqx", "iqy", "iqz"};
        double* arrs[] = {irho, iux, iuy, iuz, imx, imy, imz, item, iqx, iqy, iqz};
        for (int a = 0; a < 11; ++a) wr(names[a], arrs[a], n);
        wr("dfdss", dfdss.data(), n);
        wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
    }
#else
    {
        (void)gz; (void)lz;
        D2Q9<double> pf(lx, ly), pg(lx, ly);
        const int n = pf.nxyz;
        double **rho = new double*[nt], **ux = new double*[nt], **uy = new double*[nt], **tem = new double*[nt], **gi = new double*[nt];
        for (int t = 0; t < nt; ++t) {
            rho[t] = new double[n]; ux[t] = new double[n]; uy[t] = new double[n]; tem[t] = new double[n]; gi[t] = new double[n*pg.nc];
        }
        double *qx = filled(n, 0.0), *qy = filled(n, 0.0);
        double *irho = filled(n, 0.0), *iux = filled(n, 0.0), *iuy = filled(n, 0.0), *imx = filled(n, 0.0), *imy = filled(n, 0.0);
        double *item = filled(n, 0.0), *iqx = filled(n, 0.0), *iqy = filled(n, 0.0);
        double *alpha = new double[n], *diffusivity = new double[n], *dads = new double[n], *dkds = new double[n];
        double *igi = new double[n*pg.nc];
        rd("alpha.bin", alpha, n); rd("kappa.bin", diffusivity, n); rd("dads.bin", dads, n); rd("dkds.bin", dkds, n);

        std::vector<double> dfdss(n, 0.0);
        for (int it = 0; it < iterations; ++it) {
        for (int idx = 0; idx < n; idx++) { rho[0][idx] = 1.0; ux[0][idx] = 0.0; uy[0][idx] = 0.0; tem[0][idx] = 0.0; }
        NS::InitialCondition(pf, rho[0], ux[0], uy[0]);
        AD::InitialCondition(pg, tem[0], ux[0], uy[0]);
        sync_device(); t0 = clk::now();
        for (int t = 1; t < nt; ++t) {
            AD::MacroBrinkmanCollideNaturalConvection(pf, rho[t], ux[t], uy[t], alpha, nu, pg, tem[t], qx, qy, diffusivity, gx, gy, tem0, true, gi[t]);
            pf.Stream();
            pg.Stream();
            pf.BoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 1; });
            AD::BoundaryConditionSetT(pg, [=](int _i, int _j) { return tem0; }, ux[t], uy[t], [=](int _i, int _j) { return _i == lx - 1 || _j == ly - 1; });
            AD::BoundaryConditionSetQ(pg, [=](int _i, int _j) { return (_j == 0 && _i < L) ? qn : 0.0; }, ux[t], uy[t], diffusivity, [=](int _i, int _j) { return _j == 0; });
            pg.BoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 0; });
            pf.SmoothCorner();
            pg.SmoothCorner();
        }
        sync_device(); t1 = clk::now();
        if (it > 0) {
            for (int idx = 0; idx < n; idx++) {
                dfdss[idx] = 0.0; qx[idx] = 0.0; qy[idx] = 0.0; irho[idx] = 0.0; iux[idx] = 0.0; iuy[idx] = 0.0; imx[idx] = 0.0; imy[idx] = 0.0;
                item[idx] = 0.0; iqx[idx] = 0.0; iqy[idx] = 0.0;
            }
        }
        ANS::InitialCondition(pf, ux[nt - 1], uy[nt - 1], irho, iux, iuy);
        AAD::InitialCondition(pg, ux[nt - 1], uy[nt - 1], item, iqx, iqy);
        sync_device(); t2 = clk::now();
        for (int t = nt - 2; t >= 0; --t) {
            AAD::MacroBrinkmanCollideNaturalConvection(pf, rho[t], ux[t], uy[t], irho, iux, iuy, imx, imy, alpha, nu, pg, tem[t], item, iqx, iqy, diffusivity, gx, gy, true, igi);
            AAD::SensitivityTemperatureAtHeatSource(pg, dfdss.data(), ux[t], uy[t], imx, imy, dads, tem[t], item, iqx, iqy, gi[t], igi, diffusivity, dkds,
                [=](int _i, int _j) { return (_j == 0 && _i < L) ? qn : 0.0; }, [=](int _i, int _j) { return _j == 0 && _i < L; });
            pf.iStream();
            pg.iStream();
            AAD::iBoundaryConditionSetT(pg, ux[t], uy[t], [=](int _i, int _j) { return _i == lx - 1 || _j == ly - 1; });
            AAD::iBoundaryConditionSetQ(pg, ux[t], uy[t], [=](int _i, int _j) { return _j == 0; });
            AAD::iBoundaryConditionSetQ(pg, ux[t], uy[t], [=](int _i, int _j) { return _j == 0 && _i < L; }, 1.0);
            pg.iBoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 0; });
            pf.iBoundaryCondition([=](int _i, int _j) { return _i == 0 ? 2 : 1; });
            pf.SmoothCorner();
            pg.SmoothCorner();
        }
        sync_device(); t3 = clk::now();
        }
        for (int t = 0; t < nt; ++t) for (int i = 0; i < pf.nx; ++i) if (i < L) f_buffer += tem[t][pf.Index(i, 0)];
        const int tq[3] = {1, nt/2, nt - 1};
        for (int q = 0; q < 3; ++q) {
            const std::string s = "@" + std::to_string(q);
            wr("rho" + s, rho[tq[q]], n); wr("ux" + s, ux[tq[q]], n); wr("uy" + s, uy[tq[q]], n); wr("tem" + s, tem[tq[q]], n);
        }
        const char* names[] = {"qx", "qy", "ip", "iux", "iuy", "imx", "imy", "item", "iqx", "iqy"};
        double* arrs[] = {qx, qy, irho, iux, iuy, imx, imy, item, iqx, iqy};
        for (int a = 0; a < 10; ++a) wr(names[a], arrs[a], n);
        wr("dfdss", dfdss.data(), n);
        wr("f.f0", pf.f0, n); wr("f.f", pf.f, (size_t)n*(pf.nc - 1)); wr("g.f0", pg.f0, n); wr("g.f", pg.f, (size_t)n*(pg.nc - 1));
    }
#endif
    double extra[1] = {f_buffer};
    wr("extra", extra, 1);
    {
        const double fs = std::chrono::duration<double>(t1 - t0).count(), as = std::chrono::duration<double>(t3 - t2).count();
        const double sites = (double)lx*ly*lz;
        printf("forward %d steps %.3f ms/step %.1f MLUPS | adjoint+sensitivity %.3f ms/step %.1f MLUPS\n", nt - 1, 1e3*fs/(nt - 1), sites*(nt - 1)/fs/1e6,
               1e3*as/(nt - 1), sites*(nt - 1)/as/1e6);
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
    uint64_t ss[4];
    plh_store_stats(ss);
    printf("state store: spilled %llu mirrors, restored %llu, device bytes now %llu, peak %llu\n", (unsigned long long)ss[0], (unsigned long long)ss[1],
           (unsigned long long)ss[2], (unsigned long long)ss[3]);
#endif
#ifdef _USE_MPI_DEFINES
    MPI_Finalize();
#endif
    return 0;
}
