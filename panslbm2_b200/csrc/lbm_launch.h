// Launch entry points of the per-model kernels (k_collide, k_fused, k_shell, k_tubes).  Every (lattice, collide model) pair is
// compiled in its own translation unit (lbm_model_inst.cu with -DPLI_DIM=.. -DPLI_MODEL=..) so that the library builds in parallel;
// panslbm_api.cu reaches them through this table.  Host-side plumbing only.
#pragma once
#include "lbm_kernels.cuh"
#include "lbm_steps.cuh"
#include <cuda_runtime.h>

namespace plb {

struct FusedArgs {
    Geom G;
    const double *fs, *gs;       // populations the pass reads  (gs == nullptr: one lattice)
    double *fd, *gd;             // ... and writes (== fs/gs for the in-place passes)
    CollideParams P;
    ShellMask S;
    const ClosureArgs* prog;     // closure program of the step (k_fused: only with PANSLBM_XINLINE)
    int inverse;
    XWall W;
    // boundary pass
    const int* list; const unsigned long long* ent; int nlist, ndirect;
    double *tube_f, *tube_g; const TubeSite* tube_info;
    HaloView HF, HG;
    int pipe;                    // interior kernel: the software-pipelined persistent form (k_fused_pipe) on `sms` multiprocessors
    int sms;
};
struct ModelLaunch {
    cudaError_t (*collide)(cudaStream_t, const Geom&, double* fb, double* gb, const CollideParams&, const int* list, long long count);
    cudaError_t (*fused)(cudaStream_t, const FusedArgs&, int mode);
    cudaError_t (*shell)(cudaStream_t, const FusedArgs&, int mode);     // k_shell (+ k_tubes right behind it when the plan has tube sites)
    // k_steps: A.nsteps fused passes in one cooperative launch; *max_blocks = co-resident CTAs the device can hold (0: unsupported)
    cudaError_t (*steps)(cudaStream_t, const StepsArgs&, int sms, int* grid_used);
};
// nullptr: the model does not exist for this lattice (PL_AAD_NAT_CONV_MASSFLOW is D2Q9 only)
const ModelLaunch* model_launch(int D, int M);

}  // namespace plb
