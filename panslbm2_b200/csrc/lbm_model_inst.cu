// One (lattice, collide model) pair of the per-model kernels: compiled once per pair with -DPLI_DIM=<2|3> -DPLI_MODEL=<1..12>
// (panslbm2_b200/build.py), so that the 23 pairs build in parallel.  See lbm_launch.h.
#include "lbm_launch.h"
#include <algorithm>

#if !defined(PLI_DIM) || !defined(PLI_MODEL)
#error "compile with -DPLI_DIM=<2|3> -DPLI_MODEL=<model>"
#endif

namespace plb {
namespace {

inline unsigned blocks(long long n, int bs) { return (unsigned)((n + bs - 1)/bs); }

cudaError_t collide_(cudaStream_t st, const Geom& G, double* fb, double* gb, const CollideParams& P, const int* list, long long count) {
    if (count == 0) return cudaSuccess;
    k_collide<PLI_DIM, PLI_MODEL><<<blocks(count, 256), 256, 0, st>>>(G, fb, gb, P, list, count);
    return cudaGetLastError();
}

template <int MODE> cudaError_t fused_m(cudaStream_t st, const FusedArgs& A) {
    k_fused<PLI_DIM, PLI_MODEL, MODE><<<blocks(A.G.npacked, PLK_FUSED_THREADS), PLK_FUSED_THREADS, 0, st>>>(A.G, A.fs, A.fd, A.gs, A.gd, A.P, A.S, A.prog, A.inverse, A.W);
    return cudaGetLastError();
}
template <int MODE> cudaError_t fused_pipe_m(cudaStream_t st, const FusedArgs& A) {
    constexpr bool hasg = (ModelFlags<PLI_MODEL>::v & F_G) != 0;
    constexpr size_t smem = (size_t)(hasg ? 2 : 1)*LT<PLI_DIM>::nc*PIPE_THREADS*sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_fused_pipe<PLI_DIM, PLI_MODEL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const long long ntiles = (A.G.npacked + PIPE_THREADS - 1)/PIPE_THREADS;
    const unsigned grid = (unsigned)std::min<long long>(ntiles, 2LL*(A.sms > 0 ? A.sms : 148));
    k_fused_pipe<PLI_DIM, PLI_MODEL, MODE><<<grid, PIPE_THREADS, smem, st>>>(A.G, A.fs, A.fd, A.gs, A.gd, A.P, A.S, A.inverse, A.W);
    return cudaGetLastError();
}
cudaError_t fused_(cudaStream_t st, const FusedArgs& A, int mode) {
    if (A.G.npacked == 0) return cudaSuccess;
    if (A.pipe && A.prog == nullptr) {
        switch (mode) {
            case PASS_GATHER: return fused_pipe_m<PASS_GATHER>(st, A);
            case PASS_LOCAL: return fused_pipe_m<PASS_LOCAL>(st, A);
            default: return fused_pipe_m<PASS_COPY>(st, A);
        }
    }
    switch (mode) {
        case PASS_GATHER: return fused_m<PASS_GATHER>(st, A);
        case PASS_LOCAL: return fused_m<PASS_LOCAL>(st, A);
        default: return fused_m<PASS_COPY>(st, A);
    }
}

template <int MODE> cudaError_t shell_m(cudaStream_t st, const FusedArgs& A) {
    k_shell<PLI_DIM, PLI_MODEL, MODE><<<blocks(A.nlist, SHELL_THREADS), SHELL_THREADS, 0, st>>>(A.G, A.fs, A.fd, A.gs, A.gd, A.P, A.S, A.prog, A.list, A.ent, A.nlist, A.ndirect,
                                                                                       A.inverse, A.tube_f, A.tube_g, A.HF, A.HG, A.W);
    cudaError_t e = cudaGetLastError();
    // SmoothCorner + collide of the tube sites, right behind the boundary pass on the same stream
    const int ntube = A.nlist - A.ndirect;
    if (e == cudaSuccess && ntube > 0) {
        k_tubes<PLI_DIM, PLI_MODEL, MODE><<<blocks(ntube, 128), 128, 0, st>>>(A.G, A.tube_f, A.tube_g, A.fd, A.gd, A.P, A.tube_info, ntube, A.W, A.inverse);
        e = cudaGetLastError();
    }
    return e;
}
cudaError_t shell_(cudaStream_t st, const FusedArgs& A, int mode) {
    if (A.nlist == 0) return cudaSuccess;
    switch (mode) {
        case PASS_GATHER: return shell_m<PASS_GATHER>(st, A);
        case PASS_LOCAL: return shell_m<PASS_LOCAL>(st, A);
        default: return shell_m<PASS_COPY>(st, A);
    }
}

cudaError_t steps_(cudaStream_t st, const StepsArgs& A, int sms, int* grid_used) {
    static int per_sm = -1;
    if (per_sm < 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_steps<PLI_DIM, PLI_MODEL>, STEPS_THREADS, 0);
        if (e != cudaSuccess) { per_sm = -1; return e; }
    }
    if (per_sm < 1) return cudaErrorCooperativeLaunchTooLarge;
    const long long want = std::max<long long>(1, (std::max<long long>(A.G.npacked, (long long)A.nlist*2) + STEPS_THREADS - 1)/STEPS_THREADS);
    const int grid = (int)std::min<long long>(want, (long long)per_sm*(sms > 0 ? sms : 148));
    if (grid_used) *grid_used = grid;
    void* args[] = {(void*)&A};
    return cudaLaunchCooperativeKernel((const void*)k_steps<PLI_DIM, PLI_MODEL>, dim3(grid), dim3(STEPS_THREADS), args, 0, st);
}

}  // namespace

#define PL_CAT_(a, b, c) a##b##_##c
#define PL_CAT(a, b, c) PL_CAT_(a, b, c)
extern const ModelLaunch PL_CAT(model_launch_, PLI_DIM, PLI_MODEL);
const ModelLaunch PL_CAT(model_launch_, PLI_DIM, PLI_MODEL) = {collide_, fused_, shell_, steps_};

}  // namespace plb
