// Several fused time steps in ONE cooperative launch, for lattices that live in L2.
// A 141 x 161 D2Q9 lattice (production/heatsink.cpp:41) or a 41 x 81 x 41 block (production/heatsink3D.cpp on 2 x 2 x 2) is a few
// 10^4 - 10^5 sites: a step of four dependent kernel launches (k_xclose -> k_fused, k_shell -> k_tubes, join) costs 30 us of
// launch and dependency latency for a few us of work.  k_steps keeps the grid resident and walks through the same three
// phases per step — x-plane closures on the compact wall buffers | interior sites + boundary-pass sites | SmoothCorner tubes —
// separated by grid-wide barriers instead of kernel boundaries, for as many steps as the caller asks, alternating the two
// argument sets, the gather / local pass and the wall-buffer phase exactly as pl_plan_advance does pass by pass.  Same site
// functions, same arithmetic, same memory locations as k_xclose / k_fused / k_shell / k_tubes (lbm_kernels.cuh).
#pragma once
#include "lbm_kernels.cuh"
#include <cooperative_groups.h>

namespace plb {

struct StepsArgs {
    Geom G;
    double *f, *g;                          // population buffers, updated in place (g == nullptr: one lattice)
    CollideParams P[2];                     // collide arguments by argument-set parity
    const ClosureArgs* prog[2];             // closure program by argument-set parity
    ShellMask S;
    int inverse;
    const int* list; const unsigned long long* ent; int nlist, ndirect;      // boundary pass
    double *tube_f, *tube_g; const TubeSite* tube_info;
    const int* xlist; const unsigned long long* xent; int nxlist; XNeed xneed;   // x planes on the compact wall buffers
    double *xout_f[2], *xout_g[2], *xres_f, *xres_g;
    int np, xon[2];
    int nsteps;        // fused passes to run
    int parity;        // argument set of the last collide executed (the closures of the first pass use it, its collide the other one)
    int mode;          // PASS_GATHER / PASS_LOCAL of the first pass (they alternate)
    int xphase;        // wall `out` buffer the first pass reads
    int save_last;     // the last save_last passes store macros / snapshot at every site (< 0: all of them)
};
constexpr int STEPS_THREADS = 256;

template <int D, int M, int MODE>
PL_D void steps_pass(const StepsArgs& A, const CollideParams& P, int issave, const ClosureArgs* prog, int xphase, double* tile, cooperative_groups::grid_group& grid) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr bool HASG = (FL & F_G) != 0;
    constexpr int NC = LT<D>::nc;
    const Geom& G = A.G;
    const int tid = threadIdx.x;
    XWall W;
    W.out_f = A.nxlist ? A.xout_f[xphase ^ 1] : nullptr; W.out_g = A.nxlist ? A.xout_g[xphase ^ 1] : nullptr;
    W.res_f = A.xres_f; W.res_g = A.xres_g; W.np = A.np; W.on[0] = A.xon[0]; W.on[1] = A.xon[1];
    // ---- phase 1: closures of the x boundary planes (k_xclose)
    if (tid < SHELL_THREADS) {
        const double *in_f = A.xout_f[xphase], *in_g = A.xout_g[xphase];
        for (int t = blockIdx.x*SHELL_THREADS + tid; t < A.nxlist; t += gridDim.x*SHELL_THREADS) {
            const long long idx = A.xlist[t];
            const unsigned long long entries = A.xent[t];
            int i, j, k;
            decompose(G, idx, i, j, k);
            const int side = i == 0 ? 0 : 1;
            const size_t tp = (size_t)(j + G.ny*k);
            double f[NC], g[NC];
            sfor<0, NC>([&](auto C) {
                constexpr int c = decltype(C)::value;
                const size_t o = (size_t)(side*NC + c)*A.np + tp;
                const bool wrapped = (A.inverse ? -LT<D>::cx(c) : LT<D>::cx(c)) == (side ? -1 : 1);
                f[c] = (((A.xneed.f[side] >> c) & 1u) || wrapped) ? in_f[o] : 0.0;
                if constexpr (HASG) g[c] = (((A.xneed.g[side] >> c) & 1u) || wrapped) ? in_g[o] : 0.0;
            });
            boundary_path_sh<D, HASG>(f, g, tile + tid, prog, entries, i, j, k, idx);
            const int want = side ? -1 : 1;
            sfor<1, NC>([&](auto C) {
                constexpr int c = decltype(C)::value;
                constexpr int X = LT<D>::cx(c);
                if constexpr (X != 0) {
                    if ((A.inverse ? -X : X) == want) {
                        const size_t o = (size_t)(side*NC + c)*A.np + tp;
                        A.xres_f[o] = f[c];
                        if constexpr (HASG) A.xres_g[o] = g[c];
                    }
                }
            });
        }
    }
    grid.sync();
    // ---- phase 2a: interior sites (k_fused)
    for (long long idx = (long long)blockIdx.x*STEPS_THREADS + tid; idx < G.npacked; idx += (long long)gridDim.x*STEPS_THREADS) {
        int i, j, k, wside;
        unsigned long long entries;
        decompose(G, idx, i, j, k);
        if (!interior_site(A.S, i, j, k, entries, wside)) continue;
        Nbr n = neighbours(G, i, j, k);
        orient(n, A.inverse);
        double f[NC], g[NC];
        pass_load_wall<D, MODE, HASG>(f, g, A.f, A.g, G.pitch, idx, n, W, wside, (size_t)(j + G.ny*k), A.inverse);
        collide_site<D, FL, false>(f, g, P, (size_t)idx, issave == 1 || (issave == 2 && entries != 0ull));
        pass_store<D, MODE>(f, A.f, G.pitch, idx, n);
        if constexpr (HASG) pass_store<D, MODE>(g, A.g, G.pitch, idx, n);
        wall_scatter<D, HASG>(W, f, g, G, i, j, k, A.inverse);
    }
    // ---- phase 2b: boundary-pass sites (k_shell): closure planes and the scalar tail collide here, tube sites go to the tube buffer
    if (tid < SHELL_THREADS) {
        for (int t = blockIdx.x*SHELL_THREADS + tid; t < A.nlist; t += gridDim.x*SHELL_THREADS) {
            const long long idx = A.list[t];
            const unsigned long long entries = A.ent[t];
            int i, j, k;
            decompose(G, idx, i, j, k);
            Nbr n = neighbours(G, i, j, k);
            orient(n, A.inverse);
            double f[NC], g[NC];
            pass_load<D, MODE>(f, A.f, G.pitch, idx, n);
            if constexpr (HASG) pass_load<D, MODE>(g, A.g, G.pitch, idx, n);
            if (entries) boundary_path_sh<D, HASG>(f, g, tile + tid, prog, entries, i, j, k, idx);
            if (t < A.ndirect) {
                if (idx < G.npacked) collide_site<D, FL, false>(f, g, P, (size_t)idx, issave != 0);
                else collide_site<D, FL, true>(f, g, P, (size_t)idx, issave != 0);
                pass_store<D, MODE>(f, A.f, G.pitch, idx, n);
                if constexpr (HASG) pass_store<D, MODE>(g, A.g, G.pitch, idx, n);
                wall_scatter<D, HASG>(W, f, g, G, i, j, k, A.inverse);
            } else {
                const size_t nt = (size_t)(A.nlist - A.ndirect), tt = (size_t)(t - A.ndirect);
                sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; A.tube_f[c*nt + tt] = f[c]; if constexpr (HASG) A.tube_g[c*nt + tt] = g[c]; });
            }
        }
    }
    const int ntube = A.nlist - A.ndirect;
    if (ntube > 0) {
        grid.sync();
        // ---- phase 3: SmoothCorner + collide of the tube sites (k_tubes)
        for (int tt = blockIdx.x*STEPS_THREADS + tid; tt < ntube; tt += gridDim.x*STEPS_THREADS) {
            const TubeSite T = A.tube_info[tt];
            double f[NC], g[NC];
            tube_load<D>(f, A.tube_f, (size_t)ntube, (size_t)tt, T.kind[0], T.a[0]);
            if constexpr (HASG) tube_load<D>(g, A.tube_g, (size_t)ntube, (size_t)tt, T.kind[1], T.a[1]);
            const long long idx = T.idx;
            if (idx < G.npacked) collide_site<D, FL, false>(f, g, P, (size_t)idx, issave != 0);
            else collide_site<D, FL, true>(f, g, P, (size_t)idx, issave != 0);
            int i, j, k;
            decompose(G, idx, i, j, k);
            Nbr n = neighbours(G, i, j, k);
            orient(n, A.inverse);
            pass_store<D, MODE>(f, A.f, G.pitch, idx, n);
            if constexpr (HASG) pass_store<D, MODE>(g, A.g, G.pitch, idx, n);
            wall_scatter<D, HASG>(W, f, g, G, i, j, k, A.inverse);
        }
    }
    grid.sync();
}

template <int D, int M>
__global__ void __launch_bounds__(STEPS_THREADS, 2) k_steps(const __grid_constant__ StepsArgs A) {
    constexpr bool HASG = (ModelFlags<M>::v & F_G) != 0;
    __shared__ double tile[(HASG ? 2 : 1)*LT<D>::nc*SHELL_THREADS];
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    int par = A.parity, mode = A.mode, xphase = A.xphase;
    for (int step = 0; step < A.nsteps; ++step) {
        // the pass applies the closures recorded with argument set `par` and the collide of the other set
        const CollideParams& P = A.P[par ^ 1];
        const bool full = A.save_last < 0 || A.nsteps - step <= A.save_last;
        const int issave = P.issave ? (full ? 1 : 2) : 0;
        if (mode == PASS_GATHER) steps_pass<D, M, PASS_GATHER>(A, P, issave, A.prog[par], xphase, tile, grid);
        else steps_pass<D, M, PASS_LOCAL>(A, P, issave, A.prog[par], xphase, tile, grid);
        par ^= 1; xphase ^= 1;
        mode = mode == PASS_GATHER ? PASS_LOCAL : PASS_GATHER;
    }
}

}  // namespace plb
