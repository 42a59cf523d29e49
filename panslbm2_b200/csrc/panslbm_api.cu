// C-ABI of libpanslbm_b200.so (see include/panslbm_c.h).  Host-side orchestration only: every numerical
// operation is a CUDA kernel from lbm_kernels.cuh / lbm_closures.cuh / lbm_reduce.cuh.  No CPU fallback.
#include "../../include/panslbm_c.h"
#include "lbm_kernels.cuh"
#include "lbm_launch.h"
#include "lbm_reduce.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

using namespace plb;

namespace plb {
#define PL_ML(D, M) extern const ModelLaunch model_launch_##D##_##M;
PL_ML(2, 1) PL_ML(2, 2) PL_ML(2, 3) PL_ML(2, 4) PL_ML(2, 5) PL_ML(2, 6) PL_ML(2, 7) PL_ML(2, 8) PL_ML(2, 9) PL_ML(2, 10) PL_ML(2, 11) PL_ML(2, 12) PL_ML(2, 13) PL_ML(2, 14)
PL_ML(3, 1) PL_ML(3, 2) PL_ML(3, 3) PL_ML(3, 4) PL_ML(3, 5) PL_ML(3, 6) PL_ML(3, 7) PL_ML(3, 8) PL_ML(3, 9) PL_ML(3, 10) PL_ML(3, 11)
#undef PL_ML
const ModelLaunch* model_launch(int D, int M) {
    static const ModelLaunch* const t2[15] = {nullptr, &model_launch_2_1, &model_launch_2_2, &model_launch_2_3, &model_launch_2_4, &model_launch_2_5, &model_launch_2_6,
                                              &model_launch_2_7, &model_launch_2_8, &model_launch_2_9, &model_launch_2_10, &model_launch_2_11, &model_launch_2_12,
                                              &model_launch_2_13, &model_launch_2_14};
    static const ModelLaunch* const t3[15] = {nullptr, &model_launch_3_1, &model_launch_3_2, &model_launch_3_3, &model_launch_3_4, &model_launch_3_5, &model_launch_3_6,
                                              &model_launch_3_7, &model_launch_3_8, &model_launch_3_9, &model_launch_3_10, &model_launch_3_11, nullptr, nullptr, nullptr};
    if (M < 1 || M > 14) return nullptr;
    return D == 2 ? t2[M] : (D == 3 ? t3[M] : nullptr);
}
}  // namespace plb

// -------------------------------------------------------------------------------------------------
namespace {
thread_local std::string g_err;
cudaStream_t g_stream = 0;          // legacy default stream unless the caller installs another one
uint64_t g_launches = 0;

int fail(int code, const std::string& msg) { g_err = msg; return code; }
// entry points that hand CALLER host pointers to the CUDA runtime mark themselves: the coherence layer of the host-pointer surface
// (panslbm_host.cpp) must not try to serve a page fault taken in there
thread_local int g_in_call = 0;
struct InCall { InCall() { ++g_in_call; } ~InCall() { --g_in_call; } };
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(PL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));             \
    } while (0)
#define CUP(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            fail(PL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                     \
            return nullptr;                                                                             \
        }                                                                                               \
    } while (0)
#define LAUNCH_ON(stream, kernel, grid, block, ...)                                                     \
    do {                                                                                                \
        kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                          \
        ++g_launches;                                                                                   \
        CU(cudaGetLastError());                                                                         \
    } while (0)
#define LAUNCH(kernel, grid, block, ...) LAUNCH_ON(g_stream, kernel, grid, block, __VA_ARGS__)

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1)/bs); }

// tuning knobs (environment, read once): width of the aligned x group the boundary pass takes around an x closure plane,
// and the queueing order of interior kernel / boundary pass on a single block
int env_int(const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; }
int opt_xslab() { static int w = std::max(1, std::min(32, env_int("PANSLBM_XSLAB", 4))); return w; }
bool opt_fused_first() { static int v = env_int("PANSLBM_FUSED_FIRST", 0); return v != 0; }
// x closure planes of an undecomposed axis: 1 = the interior kernel runs their closures inline, 0 = the boundary pass takes
// the aligned x group around them
int opt_prefetch() { static int v = std::max(0, std::min(3, env_int("PANSLBM_PREFETCH", 0))); return v; }
// single block only: 1 = the boundary pass is queued behind the interior kernel on the same stream instead of beside it
bool opt_shell_serial() { static int v = env_int("PANSLBM_SHELL_SERIAL", 0); return v != 0; }
// replay a fused step as one captured CUDA graph (1 launch instead of 5 launches + 4 event operations); matters for the
// launch-bound 2-D domains (production/heatsink.cpp: 141 x 161 sites)
bool opt_graph() { static int v = env_int("PANSLBM_GRAPH", 0); return v != 0; }
// closures of the x boundary planes run ahead of the fused pass and leave their results in the periodic wrap slots of the
// source buffer (k_xclose); 0 = the boundary pass takes the aligned x groups around those planes
bool opt_xghost() { static int v = env_int("PANSLBM_XGHOST", 1); return v != 0; }
bool opt_xinline() { static int v = env_int("PANSLBM_XINLINE", 0); return v != 0; }
// fused passes update the ONE population buffer of a lattice in place (AA pattern: gather pass, local pass, ...);
// 0 = every pass goes from the buffer to a second one borrowed from the spare pool (the reference's f / fnext scheme)
bool opt_inplace() { static int v = env_int("PANSLBM_INPLACE", 1); return v != 0; }
// interior kernel as a persistent, software-pipelined kernel (cp.async prefetch of the next tile; k_fused_pipe); 0 = one thread per site
bool opt_pipe() { static int v = env_int("PANSLBM_PIPE", 0); return v != 0; }
// interior kernel: L2 prefetch distance in CTAs (the CTA that follows on the same SM slot is 2*SMs CTAs further on); 0 = off
int opt_l2_ahead() { static int v = std::max(0, env_int("PANSLBM_L2_AHEAD", 148)); return v; }
// lattices of up to this many sites run several fused passes per cooperative launch (k_steps: grid barriers instead of kernel
// boundaries); 0 = never, the default: measured SLOWER than the launches it replaces (heatsink 141 x 161: 35 vs 29 us per step,
// 41 x 81 x 41: 1 718 vs 2 457 MLUPS — three grid barriers per step and no overlap of the boundary pass with the interior)
long long opt_coop_sites() { static long long v = std::max(0, env_int("PANSLBM_COOP_SITES", 0)); return v; }
int device_sms() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148; }
    return n;
}

// Spare population buffers.  A lattice owns ONE buffer; the operations that cannot work in place — a standalone Stream()/iStream(),
// the conversion of the streamed layout back to the natural one, the two-buffer passes of PANSLBM_INPLACE=0 — write into a buffer
// borrowed here and hand their old one back.  Lattices of one shape share the spares (f and g of a driver stream one after the
// other through the same one), and the pool is emptied once a run of 64 in-place passes shows that nobody needs them.
struct SparePool {
    std::multimap<size_t, double*> free_;
    int streak = 0;                    // fused in-place passes since the last borrow
    uint64_t borrows = 0;
    size_t held() const { size_t b = 0; for (auto& kv : free_) b += kv.first; return b; }
    double* get(size_t bytes) {
        streak = 0; ++borrows;
        auto it = free_.find(bytes);
        if (it != free_.end()) { double* p = it->second; free_.erase(it); return p; }
        double* p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); trim(); if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } }
        return p;
    }
    void put(double* p, size_t bytes) { free_.emplace(bytes, p); }
    void trim() {
        if (free_.empty()) return;
        cudaStreamSynchronize(g_stream);
        for (auto& kv : free_) cudaFree(kv.second);
        free_.clear();
    }
    // 64 passes without a borrow: a production loop (thousands of steps) sheds its spares early, a short loop that closes with a
    // standalone Stream() every few steps keeps them (cudaMalloc / cudaFree of a 5 GB buffer inside a loop costs milliseconds)
    void inplace_pass() { if (!free_.empty() && ++streak >= 64) trim(); }
} g_spares;
int g_scalar_order = 0;      // pl_set_scalar_order: the caller is a build WITHOUT _USE_AVX_DEFINES (scalar templates at every site)
inline long long packed_sites(long long nxyz) { return g_scalar_order ? 0 : 4*(nxyz/4); }
uint64_t g_lattice_bytes = 0, g_conversions = 0;      // population buffers owned by live lattices; streamed -> natural conversions so far

// grow-only device scratch for the reductions: cudaMalloc/cudaFree per call would cost milliseconds next to tens of GB of
// live allocations and synchronise the device
struct Scratch {
    void* p = nullptr; size_t cap = 0;
    void* get(size_t bytes) {
        if (bytes > cap) {
            if (p) { cudaStreamSynchronize(g_stream); cudaFree(p); p = nullptr; cap = 0; }
            if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return nullptr; }
            cap = bytes;
        }
        return p;
    }
} g_scratch;
}  // namespace

struct pl_lattice {
    int kind;                      // 2 / 3
    int nc;
    int lx, ly, lz, peid, mx, my, mz, pex, pey, pez;
    Geom g;
    double* buf = nullptr;         // the populations: ONE buffer, fp64 SoA [c][pitch]
    int rep = 0;                   // layout: 0 = natural, 1 = streamed (left by an in-place gather pass; lbm_kernels.cuh, pass modes)
    int rep_inverse = 0;           // ... of a Stream (0) / iStream (1)
    int streamed = 1;              // phase: 1 = populations are "pre-collision" (after init / Stream+closures), 0 = just collided
    uint64_t version = 0;          // bumped by everything that changes the populations (plans check whether their wall buffers still describe them)
    double* current() const { return buf; }
    size_t bytes() const { return g.pitch*(size_t)nc*sizeof(double); }
    // ---- halo of a block-decomposed lattice (lbm_halo.cuh) ----
    struct Halo {
        bool on = false;
        int e[3] = {0, 0, 0};
        int nmsg = 0;
        HaloMsgDesc msg[2][26];              // [inverse][message]: same codes, peers and sizes, different population sets
        size_t off[26];                      // offset (doubles) of message m inside a send / receive buffer
        int index_of_code[27];               // message index of a direction code, -1 if none
        size_t total = 0;                    // doubles per buffer
        double* send[2] = {nullptr, nullptr};// double-buffered by epoch parity (loopback peers read the previous one late)
        double* recv = nullptr;
        long long epoch = 0;                 // number of packs so far
        bool packed = false;                 // send[epoch&1] holds the outgoing populations of the current state
        int packed_dir = 0;                  // ... for Stream (0) / iStream (1)
        bool exchanged = false;              // the exchange of this epoch has been posted
        cudaEvent_t ev_ready = nullptr;      // NCCL mode: completes when the receive buffers hold this epoch's messages
        int seq = -1;                        // loopback mode: n-th lattice created with this PEid (pairs f with f, g with g)
    } halo;
};

struct pl_bc {
    pl_lattice* lat;
    int type, axis, coord, dir;
    bool empty;
    Plane pl;
    uint8_t* mask = nullptr;
    double *v0 = nullptr, *v1 = nullptr, *v2 = nullptr;
    std::vector<uint8_t> hmask;     // host copy of the mask (plan validation)
};


// -------------------------------------------------------------------------------------------------
// Communicator: NCCL over NVLink/NVSwitch (one process per GPU), or a process-local "loopback" world in which the
// blocks of every rank live on one device (parity tests of the decomposed path on a single GPU).
namespace {
struct NcclId { char internal[128]; };
struct NcclApi {
    void* h = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
} g_nccl;
constexpr int NCCL_F64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

enum { COMM_NONE = 0, COMM_NCCL = 1, COMM_LOOPBACK = 2 };
struct Comm {
    int mode = COMM_NONE;
    int rank = 0, nranks = 1;
    void* nccl = nullptr;
    cudaStream_t stream = nullptr;           // exchanges run here, beside the interior kernel
    cudaEvent_t ev_post = nullptr;
    std::vector<std::vector<pl_lattice*>> loop;   // loopback: lattices by PEid in creation order
    double* red = nullptr;                   // 4 doubles of device scratch for the global reductions
} g_comm;

int load_nccl() {
    if (g_nccl.h) return PL_OK;
    const char* cands[] = {getenv("PANSLBM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* c : cands) {
        if (!c) continue;
        g_nccl.h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return fail(PL_ERR_UNSUPPORTED, std::string("NCCL library not found (set PANSLBM_NCCL_LIB): ") + dlerror());
    auto sym = [&](const char* n) { return dlsym(g_nccl.h, n); };
    g_nccl.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    g_nccl.GroupStart = (int (*)())sym("ncclGroupStart");
    g_nccl.GroupEnd = (int (*)())sym("ncclGroupEnd");
    g_nccl.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))sym("ncclSend");
    g_nccl.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))sym("ncclRecv");
    g_nccl.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))sym("ncclAllReduce");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.GetErrorString || !g_nccl.GroupStart || !g_nccl.GroupEnd ||
        !g_nccl.Send || !g_nccl.Recv || !g_nccl.AllReduce) {
        g_nccl.h = nullptr;
        return fail(PL_ERR_UNSUPPORTED, "NCCL library lacks a required symbol");
    }
    return PL_OK;
}
#define NC(call)                                                                                        \
    do {                                                                                                \
        int e__ = (call);                                                                               \
        if (e__ != 0) return fail(PL_ERR_CUDA, std::string(#call) + ": " + g_nccl.GetErrorString(e__)); \
    } while (0)

int comm_streams() {
    if (!g_comm.stream) {
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&g_comm.stream, cudaStreamNonBlocking, hi));
        CU(cudaEventCreateWithFlags(&g_comm.ev_post, cudaEventDisableTiming));
        CU(cudaMalloc(&g_comm.red, 4*sizeof(double)));
    }
    return PL_OK;
}

template <int D> void describe_t(const pl_lattice* l, int inverse, HaloMsgDesc* out, int& cnt) {
    const int n[3] = {l->g.nx, l->g.ny, l->g.nz}, m[3] = {l->mx, l->my, l->mz}, pe[3] = {l->pex, l->pey, l->pez};
    cnt = halo_describe<D>(n, m, pe, inverse, out);
}

int halo_setup(pl_lattice* l) {
    pl_lattice::Halo& h = l->halo;
    for (int a = 0; a < 3; ++a) h.e[a] = 0;
    h.e[0] = l->mx > 1; h.e[1] = l->my > 1; h.e[2] = l->kind == PL_D3Q15 && l->mz > 1;
    h.on = h.e[0] || h.e[1] || h.e[2];
    if (!h.on) return PL_OK;
    for (int inv = 0; inv < 2; ++inv) {
        if (l->kind == PL_D2Q9) describe_t<2>(l, inv, h.msg[inv], h.nmsg); else describe_t<3>(l, inv, h.msg[inv], h.nmsg);
    }
    for (int c = 0; c < 27; ++c) h.index_of_code[c] = -1;
    h.total = 0;
    for (int m = 0; m < h.nmsg; ++m) {
        h.index_of_code[h.msg[0][m].code] = m;
        h.off[m] = h.total;
        h.total += (size_t)((h.msg[0][m].rsize*h.msg[0][m].npop + 15)/16*16);
    }
    for (int b = 0; b < 2; ++b) CU(cudaMalloc(&h.send[b], h.total*sizeof(double)));
    CU(cudaMalloc(&h.recv, h.total*sizeof(double)));
    CU(cudaMemsetAsync(h.recv, 0, h.total*sizeof(double), g_stream));
    CU(cudaEventCreateWithFlags(&h.ev_ready, cudaEventDisableTiming));
    if (g_comm.mode == COMM_LOOPBACK) {
        if (l->mx*l->my*l->mz != g_comm.nranks) return fail(PL_ERR_ARG, "pl_lattice_create: PE grid does not match the loopback world size");
        auto& v = g_comm.loop[l->peid];
        h.seq = (int)v.size();
        v.push_back(l);
    } else if (g_comm.mode == COMM_NCCL) {
        if (l->mx*l->my*l->mz != g_comm.nranks || l->peid != g_comm.rank)
            return fail(PL_ERR_ARG, "pl_lattice_create: PEid / PE grid do not match the communicator (rank must equal PEid)");
    }
    return PL_OK;
}
void halo_release(pl_lattice* l) {
    pl_lattice::Halo& h = l->halo;
    if (!h.on) return;
    if (g_comm.stream) cudaStreamSynchronize(g_comm.stream);
    cudaFree(h.send[0]); cudaFree(h.send[1]); cudaFree(h.recv);
    if (h.ev_ready) cudaEventDestroy(h.ev_ready);
    if (g_comm.mode == COMM_LOOPBACK && h.seq >= 0 && l->peid < (int)g_comm.loop.size() && h.seq < (int)g_comm.loop[l->peid].size())
        g_comm.loop[l->peid][h.seq] = nullptr;
}
inline void halo_touch(pl_lattice* l) { l->halo.packed = false; ++l->version; }

int halo_pack(pl_lattice* l, int inverse) {
    pl_lattice::Halo& h = l->halo;
    ++h.epoch;
    PackList L;
    L.count = h.nmsg;
    long long maxr = 1;
    for (int m = 0; m < h.nmsg; ++m) {
        const HaloMsgDesc& d = h.msg[inverse][m];
        PackMsg& M = L.m[m];
        M.dst = h.send[h.epoch & 1] + h.off[m];
        M.base = d.base; M.s1 = d.s1; M.s2 = d.s2; M.n1 = d.n1; M.n2 = d.n2; M.npop = d.npop;
        for (int s = 0; s < 5; ++s) M.pop[s] = s < d.npop ? d.pop[s] : 0;
        maxr = std::max(maxr, d.rsize);
    }
    dim3 grid((unsigned)std::min<long long>((maxr + 127)/128, 1024), (unsigned)h.nmsg);
    if (l->kind == PL_D2Q9) LAUNCH(k_halo_pack<2>, grid, 128, l->current(), l->g, L, l->rep, l->rep_inverse);
    else LAUNCH(k_halo_pack<3>, grid, 128, l->current(), l->g, L, l->rep, l->rep_inverse);
    h.packed = true; h.packed_dir = inverse; h.exchanged = false;
    return PL_OK;
}
pl_lattice* loop_peer(const pl_lattice* l, int peer) {
    if (peer < 0 || peer >= (int)g_comm.loop.size()) return nullptr;
    auto& v = g_comm.loop[peer];
    return l->halo.seq < (int)v.size() ? v[l->halo.seq] : nullptr;
}
// post the exchange of the current epoch (NCCL) / bring lagging peers to this epoch (loopback)
int halo_exchange(pl_lattice* l) {
    pl_lattice::Halo& h = l->halo;
    if (g_comm.mode == COMM_NCCL) {
        int r = comm_streams();
        if (r) return r;
        CU(cudaEventRecord(g_comm.ev_post, g_stream));
        CU(cudaStreamWaitEvent(g_comm.stream, g_comm.ev_post, 0));
        NC(g_nccl.GroupStart());
        for (int m = 0; m < h.nmsg; ++m) {
            const HaloMsgDesc& d = h.msg[h.packed_dir][m];
            const size_t count = (size_t)(d.rsize*d.npop);
            NC(g_nccl.Send(h.send[h.epoch & 1] + h.off[m], count, NCCL_F64, d.peer, g_comm.nccl, g_comm.stream));
            // the neighbour on the opposite side sends its message `code`; it arrives from direction opposite(code)
            const int from = h.index_of_code[halo_opposite(d.code)];
            NC(g_nccl.Recv(h.recv + h.off[from], count, NCCL_F64, h.msg[h.packed_dir][from].peer, g_comm.nccl, g_comm.stream));
        }
        NC(g_nccl.GroupEnd());
        ++g_launches;
        CU(cudaEventRecord(h.ev_ready, g_comm.stream));
    } else if (g_comm.mode == COMM_LOOPBACK) {
        for (int m = 0; m < h.nmsg; ++m) {
            pl_lattice* p = loop_peer(l, h.msg[0][m].peer);
            if (!p) return fail(PL_ERR_ARG, "halo exchange: the block of a neighbouring PE has not been created in this loopback world");
            if (p == l) continue;
            if (p->halo.epoch < h.epoch) { int r = halo_pack(p, h.packed_dir); if (r) return r; p->halo.exchanged = false; }
            if (p->halo.epoch > h.epoch + 1) return fail(PL_ERR_ARG, "halo exchange: loopback ranks must advance in lockstep (a neighbour is more than one exchange ahead)");
        }
    } else {
        return fail(PL_ERR_ARG, "this lattice is block-decomposed (mx*my*mz > 1): call pl_comm_init / pl_comm_init_loopback before streaming it");
    }
    h.exchanged = true;
    return PL_OK;
}
// make the receive side of `l` valid for a Stream (inverse = 0) / iStream (1) of its current populations
// eager = called right after a collide, ahead of the Stream that will need it: in loopback mode only this block is packed
// then (the neighbours pack themselves when their own collide is done; a Stream call catches up the ones that have not)
int halo_prepare(pl_lattice* l, int inverse, bool eager = false) {
    pl_lattice::Halo& h = l->halo;
    if (!h.on) return PL_OK;
    int r;
    if (!h.packed || h.packed_dir != inverse) { if ((r = halo_pack(l, inverse))) return r; }
    if (eager && g_comm.mode != COMM_NCCL) return PL_OK;
    if (!h.exchanged && (r = halo_exchange(l))) return r;
    return PL_OK;
}
int halo_view(const pl_lattice* l, HaloView& V) {
    const pl_lattice::Halo& h = l->halo;
    memset(&V, 0, sizeof(V));
    if (!h.on) return PL_OK;
    V.on = 1; V.e[0] = h.e[0]; V.e[1] = h.e[1]; V.e[2] = h.e[2];
    for (int m = 0; m < h.nmsg; ++m) {
        const int code = h.msg[0][m].code;          // data arriving from the neighbour in this direction ...
        const int theirs = h.index_of_code[halo_opposite(code)];   // ... is its message towards the opposite direction
        if (g_comm.mode == COMM_LOOPBACK) {
            const pl_lattice* p = loop_peer(l, h.msg[0][m].peer);
            if (!p) return fail(PL_ERR_ARG, "halo view: missing loopback neighbour");
            if (p->halo.epoch < h.epoch || p->halo.epoch > h.epoch + 1) return fail(PL_ERR_ARG, "halo view: loopback ranks out of lockstep");
            V.r[code] = p->halo.send[h.epoch & 1] + p->halo.off[theirs];
        } else {
            V.r[code] = h.recv + h.off[m];
        }
    }
    return PL_OK;
}
// order `stream` after the arrival of the current epoch's messages
int halo_wait(const pl_lattice* l, cudaStream_t stream) {
    if (l->halo.on && g_comm.mode == COMM_NCCL) CU(cudaStreamWaitEvent(stream, l->halo.ev_ready, 0));
    return PL_OK;
}
// bring the populations back to the natural layout (every function but the fused passes works on that one).  The content does not
// change: neither the halo buffers nor a plan's wall buffers go stale.
int make_natural(pl_lattice* l) {
    if (!l || l->rep == 0) return PL_OK;
    double* dst = g_spares.get(l->bytes());
    if (!dst) return fail(PL_ERR_CUDA, "out of device memory for the spare population buffer (streamed -> natural layout)");
    if (l->kind == PL_D2Q9) LAUNCH(k_unstream<2>, blocks_for(l->g.nxyz, 256), 256, l->g, l->buf, dst, l->rep_inverse);
    else LAUNCH(k_unstream<3>, blocks_for(l->g.nxyz, 256), 256, l->g, l->buf, dst, l->rep_inverse);
    g_spares.put(l->buf, l->bytes());
    l->buf = dst;
    l->rep = 0;
    ++g_conversions;
    return PL_OK;
}
}  // namespace

// -------------------------------------------------------------------------------------------------
extern "C" {

const char* pl_last_error(void) { return g_err.c_str(); }
int pl_in_call(void) { return g_in_call; }
const char* pl_version(void) { return "panslbm_b200 0.1 (sm_100a, fp64, fmad=off)"; }
int pl_set_scalar_order(int on) {
    const int v = on ? 1 : 0;
    // lattices fix their packed / scalar split when they are created: one order per process, chosen before the first lattice
    if (v != g_scalar_order && g_lattice_bytes != 0) return fail(PL_ERR_ARG, "pl_set_scalar_order: lattices exist already (the operation order is chosen once, before the first lattice)");
    g_scalar_order = v;
    return PL_OK;
}
int pl_scalar_order(void) { return g_scalar_order; }
int pl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int pl_set_device(int device) { CU(cudaSetDevice(device)); return PL_OK; }
int pl_synchronize(void) { CU(cudaStreamSynchronize(g_stream)); return PL_OK; }
void* pl_get_stream(void) { return (void*)g_stream; }
int pl_set_stream(void* s) { g_stream = (cudaStream_t)s; return PL_OK; }
uint64_t pl_launch_count(void) { return g_launches; }
void pl_launch_count_reset(void) { g_launches = 0; }

double* pl_array_alloc(size_t n) {
    double* p = nullptr;
    CUP(cudaMalloc(&p, std::max<size_t>(n, 1)*sizeof(double)));
    return p;
}
int pl_array_free(double* dev) { CU(cudaFree(dev)); return PL_OK; }
int pl_array_upload(double* dev, const double* host, size_t n) {
    InCall in_call_;
    CU(cudaMemcpyAsync(dev, host, n*sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
int pl_array_download(double* host, const double* dev, size_t n) {
    InCall in_call_;
    CU(cudaMemcpyAsync(host, dev, n*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
namespace {
cudaStream_t g_copy = nullptr;
cudaEvent_t g_copy_ev = nullptr, g_compute_ev = nullptr;
int copy_stream() {
    if (g_copy) return PL_OK;
    CU(cudaStreamCreateWithFlags(&g_copy, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&g_copy_ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g_compute_ev, cudaEventDisableTiming));
    return PL_OK;
}
int copy_after_compute() {
    int r = copy_stream();
    if (r) return r;
    CU(cudaEventRecord(g_compute_ev, g_stream));
    CU(cudaStreamWaitEvent(g_copy, g_compute_ev, 0));
    return PL_OK;
}
}  // namespace
extern "C" {
int pl_array_upload_async(double* dev, const double* host, size_t n) {
    if (!dev || !host) return fail(PL_ERR_ARG, "pl_array_upload_async: null");
    int r = copy_after_compute();
    if (r) return r;
    CU(cudaMemcpyAsync(dev, host, n*sizeof(double), cudaMemcpyHostToDevice, g_copy));
    return PL_OK;
}
int pl_array_download_async(double* host, const double* dev, size_t n) {
    if (!dev || !host) return fail(PL_ERR_ARG, "pl_array_download_async: null");
    int r = copy_after_compute();
    if (r) return r;
    CU(cudaMemcpyAsync(host, dev, n*sizeof(double), cudaMemcpyDeviceToHost, g_copy));
    return PL_OK;
}
int pl_copy_fence(void) {
    if (!g_copy) return PL_OK;
    CU(cudaEventRecord(g_copy_ev, g_copy));
    CU(cudaStreamWaitEvent(g_stream, g_copy_ev, 0));
    return PL_OK;
}
int pl_copy_wait(void) {
    if (!g_copy) return PL_OK;
    CU(cudaStreamSynchronize(g_copy));
    return PL_OK;
}
}
int pl_array_fill(double* dev, double value, size_t n) {
    if (n == 0) return PL_OK;
    LAUNCH(k_fill, blocks_for((long long)n, 256), 256, dev, value, (long long)n);
    return PL_OK;
}

// ---- lattices -------------------------------------------------------------------------------------
pl_lattice* pl_lattice_create(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz) {
    if ((kind != PL_D2Q9 && kind != PL_D3Q15) || lx <= 0 || ly <= 0 || lz <= 0 || peid < 0 || mx <= 0 || my <= 0 || mz <= 0) {
        fail(PL_ERR_ARG, "pl_lattice_create: bad arguments");   // the reference asserts (d3q15.h:37)
        return nullptr;
    }
    if (kind == PL_D2Q9) { lz = 1; mz = 1; }
    if (peid >= mx*my*mz) { fail(PL_ERR_ARG, "pl_lattice_create: PEid outside the PE grid"); return nullptr; }
    pl_lattice* l = new pl_lattice();
    l->kind = kind; l->nc = kind == PL_D2Q9 ? 9 : 15;
    l->lx = lx; l->ly = ly; l->lz = lz; l->peid = peid; l->mx = mx; l->my = my; l->mz = mz;
    // block decomposition: d3q15.h:29-35, d2q9.h:29-35
    l->pex = peid%mx;
    l->pey = kind == PL_D2Q9 ? peid/mx : (peid/mx)%my;
    l->pez = kind == PL_D2Q9 ? 0 : peid/(mx*my);
    Geom& g = l->g;
    g.lx = lx; g.ly = ly; g.lz = lz;
    g.nx = (lx + l->pex)/mx; g.ny = (ly + l->pey)/my; g.nz = kind == PL_D2Q9 ? 1 : (lz + l->pez)/mz;
    g.offx = mx - l->pex > lx%mx ? l->pex*g.nx : lx - (mx - l->pex)*g.nx;
    g.offy = my - l->pey > ly%my ? l->pey*g.ny : ly - (my - l->pey)*g.ny;
    g.offz = kind == PL_D2Q9 ? 0 : (mz - l->pez > lz%mz ? l->pez*g.nz : lz - (mz - l->pez)*g.nz);
    g.nxyz = (long long)g.nx*g.ny*g.nz;
    if (g.nxyz <= 0 || g.nxyz >= (1LL << 31)) { delete l; fail(PL_ERR_ARG, "pl_lattice_create: block must hold 1..2^31-1 sites"); return nullptr; }
    g.npacked = packed_sites(g.nxyz);
    g.pitch = (size_t)((g.nxyz + 15)/16*16);
    {
        cudaError_t e = cudaMalloc(&l->buf, l->bytes());
        if (e != cudaSuccess) {
            fail(PL_ERR_CUDA, std::string("pl_lattice_create: cudaMalloc: ") + cudaGetErrorString(e));
            delete l;
            return nullptr;
        }
    }
    cudaMemsetAsync(l->buf, 0, l->bytes(), g_stream);
    g_lattice_bytes += l->bytes();
    if (halo_setup(l) != PL_OK) {
        std::string keep = g_err;
        l->halo.on = true;   // release whatever halo_setup allocated
        halo_release(l);
        cudaFree(l->buf);
        g_lattice_bytes -= l->bytes();
        delete l;
        g_err = keep;
        return nullptr;
    }
    return l;
}
int pl_lattice_destroy(pl_lattice* l) {
    if (!l) return PL_OK;
    cudaStreamSynchronize(g_stream);
    halo_release(l);
    cudaFree(l->buf);
    g_lattice_bytes -= l->bytes();
    g_spares.trim();
    delete l;
    return PL_OK;
}
int pl_lattice_info(const pl_lattice* l, int* o) {
    if (!l || !o) return fail(PL_ERR_ARG, "pl_lattice_info: null");
    int v[18] = {l->lx, l->ly, l->lz, l->peid, l->mx, l->my, l->mz, l->pex, l->pey, l->pez, l->g.nx, l->g.ny, l->g.nz, (int)l->g.nxyz,
                 l->g.offx, l->g.offy, l->g.offz, l->nc};
    memcpy(o, v, sizeof(v));
    return PL_OK;
}
int pl_lattice_set_host(pl_lattice* l, const double* f0, const double* f) {
    InCall in_call_;
    if (!l || !f0 || !f) return fail(PL_ERR_ARG, "pl_lattice_set_host: null");
    size_t n = (size_t)l->g.nxyz, nf = n*(l->nc - 1);
    double *d0 = nullptr, *d1 = nullptr;
    CU(cudaMalloc(&d0, n*sizeof(double)));
    CU(cudaMalloc(&d1, nf*sizeof(double)));
    CU(cudaMemcpyAsync(d0, f0, n*sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CU(cudaMemcpyAsync(d1, f, nf*sizeof(double), cudaMemcpyHostToDevice, g_stream));
    if (l->kind == PL_D2Q9) LAUNCH(k_from_aos<2>, blocks_for(l->g.nxyz, 256), 256, l->g, d0, d1, l->current());
    else LAUNCH(k_from_aos<3>, blocks_for(l->g.nxyz, 256), 256, l->g, d0, d1, l->current());
    CU(cudaStreamSynchronize(g_stream));
    cudaFree(d0); cudaFree(d1);
    l->rep = 0;
    halo_touch(l);
    return PL_OK;
}
int pl_lattice_get_host(pl_lattice* l, double* f0, double* f) {
    InCall in_call_;
    if (!l || !f0 || !f) return fail(PL_ERR_ARG, "pl_lattice_get_host: null");
    { int r = make_natural(l); if (r) return r; }
    size_t n = (size_t)l->g.nxyz, nf = n*(l->nc - 1);
    double *d0 = nullptr, *d1 = nullptr;
    CU(cudaMalloc(&d0, n*sizeof(double)));
    CU(cudaMalloc(&d1, nf*sizeof(double)));
    if (l->kind == PL_D2Q9) LAUNCH(k_to_aos<2>, blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), d0, d1);
    else LAUNCH(k_to_aos<3>, blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), d0, d1);
    CU(cudaMemcpyAsync(f0, d0, n*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaMemcpyAsync(f, d1, nf*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    cudaFree(d0); cudaFree(d1);
    return PL_OK;
}
int pl_lattice_streamed(const pl_lattice* l) { return l ? l->streamed : 0; }

// ---- population checkpoints (checkpoint-recompute for transient adjoints) ----
struct pl_checkpoint {
    double* buf = nullptr;
    size_t bytes = 0;
    int kind = 0;
    long long nxyz = 0;
    int rep = 0, rep_inverse = 0, streamed = 1;
    bool valid = false;
};
pl_checkpoint* pl_checkpoint_create(const pl_lattice* l) {
    if (!l) { fail(PL_ERR_ARG, "pl_checkpoint_create: null"); return nullptr; }
    pl_checkpoint* c = new pl_checkpoint();
    c->bytes = l->bytes(); c->kind = l->kind; c->nxyz = l->g.nxyz;
    if (cudaMalloc(&c->buf, c->bytes) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        fail(PL_ERR_CUDA, "pl_checkpoint_create: out of device memory");
        return nullptr;
    }
    return c;
}
int pl_checkpoint_destroy(pl_checkpoint* c) {
    if (!c) return PL_OK;
    cudaStreamSynchronize(g_stream);
    cudaFree(c->buf);
    delete c;
    return PL_OK;
}
// the buffer is copied as it is — in whichever layout the last pass left it — together with the layout and phase flags
int pl_checkpoint_save(pl_checkpoint* c, const pl_lattice* l) {
    if (!c || !l) return fail(PL_ERR_ARG, "pl_checkpoint_save: null");
    if (c->kind != l->kind || c->nxyz != l->g.nxyz || c->bytes != l->bytes()) return fail(PL_ERR_ARG, "pl_checkpoint_save: the checkpoint was created for a lattice of another shape");
    CU(cudaMemcpyAsync(c->buf, l->buf, c->bytes, cudaMemcpyDeviceToDevice, g_stream));
    c->rep = l->rep; c->rep_inverse = l->rep_inverse; c->streamed = l->streamed;
    c->valid = true;
    return PL_OK;
}
int pl_checkpoint_restore(const pl_checkpoint* c, pl_lattice* l) {
    if (!c || !l) return fail(PL_ERR_ARG, "pl_checkpoint_restore: null");
    if (!c->valid) return fail(PL_ERR_ARG, "pl_checkpoint_restore: nothing has been saved into this checkpoint");
    if (c->kind != l->kind || c->nxyz != l->g.nxyz || c->bytes != l->bytes()) return fail(PL_ERR_ARG, "pl_checkpoint_restore: the checkpoint belongs to a lattice of another shape");
    CU(cudaMemcpyAsync(l->buf, c->buf, c->bytes, cudaMemcpyDeviceToDevice, g_stream));
    l->rep = c->rep; l->rep_inverse = c->rep_inverse; l->streamed = c->streamed;
    halo_touch(l);      // new content: plans refill their wall buffers, a decomposed block packs and exchanges again
    return PL_OK;
}
int pl_memory_stats(uint64_t* out4) {
    if (!out4) return fail(PL_ERR_ARG, "pl_memory_stats: null");
    out4[0] = g_lattice_bytes; out4[1] = g_spares.held(); out4[2] = g_spares.borrows; out4[3] = g_conversions;
    return PL_OK;
}
int pl_memory_trim(void) { g_spares.trim(); return PL_OK; }
int pl_lattice_device_view(pl_lattice* l, double** base, size_t* pitch) {
    if (!l) return fail(PL_ERR_ARG, "pl_lattice_device_view: null");
    { int r = make_natural(l); if (r) return r; }
    if (base) *base = l->current();
    if (pitch) *pitch = l->g.pitch;
    return PL_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// internal launch helpers (C++ linkage)
namespace {

int do_stream_all(pl_lattice* l, int inverse) {
    HaloView H;
    int r;
    if ((r = make_natural(l)) || (r = halo_prepare(l, inverse)) || (r = halo_wait(l, g_stream)) || (r = halo_view(l, H))) return r;
    double* dst = g_spares.get(l->bytes());
    if (!dst) return fail(PL_ERR_CUDA, "out of device memory for the spare population buffer (Stream)");
    if (l->halo.on) {
        if (l->kind == PL_D2Q9) LAUNCH((k_stream<2, true>), blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), dst, inverse, H);
        else LAUNCH((k_stream<3, true>), blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), dst, inverse, H);
    } else {
        if (l->kind == PL_D2Q9) LAUNCH((k_stream<2, false>), blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), dst, inverse, H);
        else LAUNCH((k_stream<3, false>), blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), dst, inverse, H);
    }
    g_spares.put(l->buf, l->bytes());
    l->buf = dst;
    halo_touch(l);
    return PL_OK;
}
// the local sites of the global plane axis=coord
bool make_plane(const pl_lattice* l, int axis, int coord, int dir, Plane& pl) {
    const Geom& g = l->g;
    int off[3] = {g.offx, g.offy, g.offz}, n[3] = {g.nx, g.ny, g.nz};
    long long st[3] = {1, g.nx, (long long)g.nx*g.ny};
    if (axis < 0 || axis >= l->kind) return false;
    int loc = coord - off[axis];
    if (loc < 0 || loc >= n[axis]) return false;
    int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
    pl.axis = axis; pl.dir = dir;
    pl.n1 = n[a1]; pl.s1 = st[a1];
    pl.n2 = n[a2]; pl.s2 = st[a2];
    pl.base = loc*st[axis];
    return true;
}

int smooth_lists(const pl_lattice* l, SmoothList& edges, SmoothList& corners) {
    const Geom& g = l->g;
    edges.count = 0; edges.maxlen = 0; corners.count = 0; corners.maxlen = 1;
    long long st[3] = {1, g.nx, (long long)g.nx*g.ny};
    int n[3] = {g.nx, g.ny, g.nz};
    int lo[3] = {0 - g.offx, 0 - g.offy, 0 - g.offz};                       // local coordinate of the global min faces
    int hi[3] = {g.lx - 1 - g.offx, g.ly - 1 - g.offy, g.lz - 1 - g.offz};  // ... of the global max faces
    auto in = [&](int d, int v) { return 0 <= v && v < n[d]; };
    // inward neighbour delta along axis d for a site on the min (dirn=-1) / max (dirn=+1) face: x - dirn (periodic wrap as Index())
    auto inward = [&](int d, int v, int dirn) -> long long {
        int w = v - dirn;
        if (w == -1) w = n[d] - 1; else if (w == n[d]) w = 0;
        return (long long)(w - v)*st[d];
    };
    if (l->kind == PL_D2Q9) {
        int cs[4][4] = {{lo[0], lo[1], -1, -1}, {lo[0], hi[1], -1, 1}, {hi[0], lo[1], 1, -1}, {hi[0], hi[1], 1, 1}};
        for (auto& c : cs) if (in(0, c[0]) && in(1, c[1])) {
            SmoothItem it{};
            it.base = c[0] + c[1]*st[1]; it.stride = 0; it.len = 1;
            it.n0 = inward(0, c[0], c[2]); it.n1 = inward(1, c[1], c[3]); it.n2 = 0;
            corners.it[corners.count++] = it;
        }
        return PL_OK;
    }
    // 12 edges (d3q15.h:200-211): line along axis `al`, fixed on the faces of the two other axes (p, q)
    for (int al = 0; al < 3; ++al) {
        int p = (al + 1)%3, q = (al + 2)%3;
        int combos[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (auto& cb : combos) {
            int vp = cb[0] < 0 ? lo[p] : hi[p], vq = cb[1] < 0 ? lo[q] : hi[q];
            if (!in(p, vp) || !in(q, vq)) continue;
            // skip the global corner sites at the ends: they are overwritten by SmoothCornerAt afterwards (d3q15.h:212-219)
            int a0 = 0, a1 = n[al] - 1;
            if (lo[al] == 0) a0 = 1;
            if (hi[al] == n[al] - 1) a1 = n[al] - 2;
            if (a1 < a0) continue;
            SmoothItem it{};
            it.base = vp*st[p] + vq*st[q] + a0*st[al]; it.stride = st[al]; it.len = a1 - a0 + 1;
            it.n0 = inward(p, vp, cb[0]); it.n1 = inward(q, vq, cb[1]); it.n2 = 0;
            edges.it[edges.count++] = it;
            edges.maxlen = std::max(edges.maxlen, it.len);
        }
    }
    int dirs[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
    for (auto& d : dirs) {
        int v[3];
        for (int a = 0; a < 3; ++a) v[a] = d[a] < 0 ? lo[a] : hi[a];
        if (!in(0, v[0]) || !in(1, v[1]) || !in(2, v[2])) continue;
        SmoothItem it{};
        it.base = v[0] + v[1]*st[1] + v[2]*st[2]; it.stride = 0; it.len = 1;
        it.n0 = inward(0, v[0], d[0]); it.n1 = inward(1, v[1], d[1]); it.n2 = inward(2, v[2], d[2]);
        corners.it[corners.count++] = it;
    }
    return PL_OK;
}

// SmoothCorner of one lattice, or of the two lattices of a plan together (same shape): all edge lines in one launch, then all corners
int do_smooth(pl_lattice* l, pl_lattice* l2 = nullptr) {
    SmoothList e, c;
    smooth_lists(l, e, c);
    const int ne = e.count, ncn = c.count;
    for (pl_lattice* q : {l, l2}) {
        if (!q) continue;
        { int r = make_natural(q); if (r) return r; }
        halo_touch(q);
        for (int k = 0; k < ne; ++k) { SmoothItem& it = e.it[q == l ? k : ne + k]; it = e.it[k]; it.fb = q->current(); }
        for (int k = 0; k < ncn; ++k) { SmoothItem& it = c.it[q == l ? k : ncn + k]; it = c.it[k]; it.fb = q->current(); }
    }
    if (l2) { e.count = 2*ne; c.count = 2*ncn; }
    if (e.count > 0) {
        dim3 grid(blocks_for(e.maxlen, 128), e.count);
        if (l->kind == PL_D2Q9) LAUNCH(k_smooth<2>, grid, 128, l->g, e); else LAUNCH(k_smooth<3>, grid, 128, l->g, e);
    }
    if (c.count > 0) {
        dim3 grid(1, c.count);
        if (l->kind == PL_D2Q9) LAUNCH(k_smooth<2>, grid, 128, l->g, c); else LAUNCH(k_smooth<3>, grid, 128, l->g, c);
    }
    return PL_OK;
}

// plane closure -> kernel argument block (+ validation of the arrays the closure type dereferences)
int make_closure_args(const pl_lattice* l, const pl_lattice* other, const pl_bc* bc, const pl_bc_aux* aux, ClosureArgs& A) {
    memset(&A, 0, sizeof(A));
    A.type = bc->type; A.pl = bc->pl; A.mask = bc->mask; A.v0 = bc->v0; A.v1 = bc->v1; A.v2 = bc->v2;
    if (aux) {
        A.rho = aux->rho; A.ux = aux->ux; A.uy = aux->uy; A.uz = aux->uz; A.tem = aux->tem; A.kappa = aux->diffusivity;
        A.kconst = aux->diffusivity_const; A.eps = aux->eps;
    }
    const bool d3 = l->kind == PL_D3Q15;
    const bool vel = A.ux && A.uy && (!d3 || A.uz);
    switch (bc->type) {
        case PL_BC_BOUNCE: case PL_BC_IBOUNCE: case PL_BC_ANS_ISET_RHO: break;
        case PL_BC_NSIN_SET_U: case PL_BC_NSIN_SET_RHO:
            if (d3) return fail(PL_ERR_UNSUPPORTED, "closure: the reference's NSin closures exist for D2Q9 only (nsincompressible.h:46-154)");
            if (!bc->v0 || !bc->v1) return fail(PL_ERR_ARG, "closure: NSin SetU / SetRho need two plane values (ux,uy / rho,us)");
            break;
        case PL_BC_NS_SET_U: case PL_BC_ANS_ISET_U:
            if (!bc->v0 || !bc->v1 || (d3 && !bc->v2)) return fail(PL_ERR_ARG, "closure: SetU needs ux,uy(,uz) plane values");
            break;
        case PL_BC_NS_SET_RHO:
            if (!bc->v0 || !bc->v1 || (d3 && !bc->v2)) return fail(PL_ERR_ARG, "closure: SetRho needs rho,us(,ut) plane values");
            break;
        case PL_BC_AD_SET_T: case PL_BC_AD_SET_Q:
            if (!bc->v0) return fail(PL_ERR_ARG, "closure: SetT/SetQ needs the plane values of T / qn");
            if (!vel) return fail(PL_ERR_ARG, "closure: SetT/SetQ needs the velocity fields (pl_bc_aux ux,uy(,uz))");
            break;
        case PL_BC_AAD_ISET_T: case PL_BC_AAD_ISET_Q:
            if (!vel) return fail(PL_ERR_ARG, "closure: iSetT/iSetQ needs the velocity fields (pl_bc_aux ux,uy(,uz))");
            break;
        case PL_BC_AAD_ISET_RHO:
            if (d3) return fail(PL_ERR_UNSUPPORTED, "closure: AAD::iBoundaryConditionSetRho is D2Q9 only");
            if (!vel || !A.rho || !A.tem) return fail(PL_ERR_ARG, "closure: AAD iSetRho needs rho,ux,uy,tem fields");
            if (!other) return fail(PL_ERR_ARG, "closure: AAD iSetRho needs the thermal lattice");
            break;
        default: return fail(PL_ERR_ARG, "closure: unknown type");
    }
    return PL_OK;
}

bool same_shape(const pl_lattice* a, const pl_lattice* b) {
    return a->kind == b->kind && a->g.nxyz == b->g.nxyz && a->g.nx == b->g.nx && a->g.ny == b->g.ny && a->g.offx == b->g.offx &&
           a->g.offy == b->g.offy && a->g.offz == b->g.offz;
}

int do_bc(pl_lattice* l, pl_lattice* other, const pl_bc* bc, const pl_bc_aux* aux) {
    if (bc->empty) return PL_OK;
    if (bc->lat != l && !same_shape(bc->lat, l)) return fail(PL_ERR_ARG, "pl_bc_apply: closure was created for a lattice of another shape");
    if (other && !same_shape(other, l)) return fail(PL_ERR_ARG, "pl_bc_apply: the two lattices differ in shape");
    ClosureArgs A;
    int r = make_closure_args(l, other, bc, aux, A);
    if (r) return r;
    int np = bc->pl.n1*bc->pl.n2;
    if ((r = make_natural(l)) || (other && (r = make_natural(other)))) return r;
    halo_touch(l);
    const double* qb = (other && bc->type == PL_BC_AAD_ISET_RHO) ? other->current() : nullptr;
    if (l->kind == PL_D2Q9) LAUNCH(k_closure<2>, blocks_for(np, 128), 128, l->g, l->current(), qb, A);
    else LAUNCH(k_closure<3>, blocks_for(np, 128), 128, l->g, l->current(), qb, A);
    return PL_OK;
}

// pl_collide_args -> kernel argument block
int make_params(const pl_lattice* f, const pl_lattice* g, const pl_collide_args* a, CollideParams& P, unsigned& flags) {
    static const unsigned FLAGS[15] = {0, ModelFlags<1>::v, ModelFlags<2>::v, ModelFlags<3>::v, ModelFlags<4>::v, ModelFlags<5>::v, ModelFlags<6>::v,
                                       ModelFlags<7>::v, ModelFlags<8>::v, ModelFlags<9>::v, ModelFlags<10>::v, ModelFlags<11>::v, ModelFlags<12>::v,
                                       ModelFlags<13>::v, ModelFlags<14>::v};
    if (!f || !a) return fail(PL_ERR_ARG, "pl_collide: null");
    if (a->model < 1 || a->model > 14) return fail(PL_ERR_ARG, "pl_collide: unknown model");
    flags = FLAGS[a->model];
    const bool d3 = f->kind == PL_D3Q15;
    if ((flags & F_INCOMP) && d3) return fail(PL_ERR_UNSUPPORTED, "pl_collide: the reference's NSin equations exist for D2Q9 only (nsincompressible.h)");
    if (a->model == PL_AAD_NAT_CONV_MASSFLOW && d3)
        return fail(PL_ERR_UNSUPPORTED, "pl_collide: the reference's D3Q15 NaturalConvectionMassFlow does not compile (adjointadvection_avx.h:1161); D2Q9 only");
    if ((flags & F_G) && (!g || g->kind != f->kind || g->g.nxyz != f->g.nxyz)) return fail(PL_ERR_ARG, "pl_collide: model needs a thermal lattice of the same shape");
    memset(&P, 0, sizeof(P));
    P.issave = a->issave;
    P.omegaf = 1.0/(3.0*a->viscosity + 0.5); P.iomegaf = 1.0 - P.omegaf;
    P.omegag = 1.0/(3.0*a->diffusivity_const + 0.5); P.iomegag = 1.0 - P.omegag;
    P.gx = a->gx; P.gy = a->gy; P.gz = d3 ? a->gz : 0.0; P.tem0 = a->tem0;
    for (int c = 0; c < f->nc; ++c) {
        double cx = d3 ? LT<3>::cx(c) : LT<2>::cx(c), cy = d3 ? LT<3>::cy(c) : LT<2>::cy(c), cz = d3 ? LT<3>::cz(c) : 0;
        double ei = d3 ? LT<3>::ei(c) : LT<2>::ei(c);
        volatile double px = cx*P.gx, py = cy*P.gy, pz = cz*P.gz;   // volatile: keep the host compiler from contracting
        volatile double s = px + py;
        if (d3) s = s + pz;
        P.cg[c] = s;
        volatile double e = ei*s;
        P.eicg[c] = e;
    }
    P.alpha = a->alpha; P.kappa = a->diffusivity; P.beta = a->beta; P.dirx = a->dirx; P.diry = a->diry; P.dirz = a->dirz;
    P.rho = a->rho; P.ux = a->ux; P.uy = a->uy; P.uz = a->uz; P.tem = a->tem; P.qx = a->qx; P.qy = a->qy; P.qz = a->qz;
    P.ip = a->ip; P.iux = a->iux; P.iuy = a->iuy; P.iuz = a->iuz; P.imx = a->imx; P.imy = a->imy; P.imz = a->imz;
    P.item = a->item; P.iqx = a->iqx; P.iqy = a->iqy; P.iqz = a->iqz;
    P.snap = (flags & F_SNAP) ? a->snapshot : nullptr; P.snap_pitch = (size_t)f->g.nxyz;
    P.scalar_build = (f->g.npacked == 0 && g_scalar_order) ? 1 : 0;
    // argument validation: every array the selected model dereferences must be present
    auto need = [&](const void* p, const char* what) { if (!p) { g_err = std::string("pl_collide: missing array ") + what; return false; } return true; };
    bool ok = true;
    const bool adj = flags & F_ADJ, two = flags & F_G;
    if (adj || a->issave) {
        ok = ok && need(a->rho, "rho") && need(a->ux, "ux") && need(a->uy, "uy") && (!d3 || need(a->uz, "uz"));
        if (two && (adj ? true : a->issave)) ok = ok && need(a->tem, "tem");
    }
    if (!adj && two && a->issave) ok = ok && need(a->qx, "qx") && need(a->qy, "qy") && (!d3 || need(a->qz, "qz"));
    if (adj && a->issave) {
        ok = ok && need(a->ip, "ip") && need(a->iux, "iux") && need(a->iuy, "iuy") && need(a->imx, "imx") && need(a->imy, "imy") && (!d3 || (need(a->iuz, "iuz") && need(a->imz, "imz")));
        if (two) ok = ok && need(a->item, "item") && need(a->iqx, "iqx") && need(a->iqy, "iqy") && (!d3 || need(a->iqz, "iqz"));
    }
    if ((flags & F_BRINK) || (adj && two)) ok = ok && need(a->alpha, "alpha");
    if (flags & F_KFIELD) ok = ok && need(a->diffusivity, "diffusivity");
    if (flags & F_HEATEX) ok = ok && need(a->beta, "beta");
    if (flags & F_MASSFLOW) ok = ok && need(a->dirx, "dirx") && need(a->diry, "diry");
    return ok ? PL_OK : PL_ERR_ARG;
}

// the kernels of one (lattice, collide model) pair: lbm_model_inst.cu, one translation unit per pair
const ModelLaunch* launcher(const pl_lattice* f, int model) {
    const ModelLaunch* ml = model_launch(f->kind == PL_D2Q9 ? 2 : 3, model);
    if (!ml) fail(PL_ERR_UNSUPPORTED, "collide: model not available for this lattice");
    return ml;
}
int dispatch_collide(int model, pl_lattice* f, pl_lattice* g, const CollideParams& P, const int* list, long long count) {
    const ModelLaunch* ml = launcher(f, model);
    if (!ml) return PL_ERR_UNSUPPORTED;
    int r;
    if ((r = make_natural(f)) || (g && (r = make_natural(g)))) return r;
    if (count == 0) return PL_OK;
    ++g_launches;
    CU(ml->collide(g_stream, f->g, f->current(), g ? g->current() : nullptr, P, list, count));
    return PL_OK;
}

}  // namespace

// -------------------------------------------------------------------------------------------------
extern "C" {

int pl_stream(pl_lattice* l, int inverse) {
    if (!l) return fail(PL_ERR_ARG, "pl_stream: null");
    int r = do_stream_all(l, inverse);
    l->streamed = 1;
    return r;
}
int pl_smooth_corner(pl_lattice* l) {
    if (!l) return fail(PL_ERR_ARG, "pl_smooth_corner: null");
    return do_smooth(l);
}

int pl_smooth_corner_at(pl_lattice* l, int gi, int gj, int gk, int dx, int dy, int dz) {
    if (!l) return fail(PL_ERR_ARG, "pl_smooth_corner_at: null");
    const Geom& g = l->g;
    const int D = l->kind;
    int d[3] = {dx, dy, D == 3 ? dz : 0}, v[3] = {gi - g.offx, gj - g.offy, D == 3 ? gk - g.offz : 0}, n[3] = {g.nx, g.ny, g.nz};
    long long st[3] = {1, g.nx, (long long)g.nx*g.ny};
    int nzero = 0, line = -1;
    for (int a = 0; a < 3; ++a) {
        if (d[a] != 0 && d[a] != 1 && d[a] != -1) return fail(PL_ERR_ARG, "pl_smooth_corner_at: directions are -1, 0 or +1");
        if (d[a] == 0) { ++nzero; if (a < D) line = a; }
    }
    if ((D == 3 && nzero > 1) || (D == 2 && nzero != 1)) return fail(PL_ERR_ARG, "pl_smooth_corner_at: need two (edge line / 2-D corner) or three (3-D corner) directions");
    for (int a = 0; a < D; ++a) if (d[a] != 0 && (v[a] < 0 || v[a] >= n[a])) return PL_OK;   // not on this rank's block (d3q15.h:1244-1245)
    auto inward = [&](int a) -> long long {
        int w = v[a] - d[a];
        if (w == -1) w = n[a] - 1; else if (w == n[a]) w = 0;
        return (long long)(w - v[a])*st[a];
    };
    SmoothList L; L.count = 1;
    SmoothItem& it = L.it[0];
    it = SmoothItem{};
    it.base = 0; it.stride = 0; it.len = 1;
    long long nb[3]; int k = 0;
    for (int a = 0; a < D; ++a) if (d[a] != 0) { it.base += v[a]*st[a]; nb[k++] = inward(a); }
    if (D == 3 && line >= 0) { it.stride = st[line]; it.len = n[line]; }
    it.n0 = nb[0]; it.n1 = nb[1]; it.n2 = k == 3 ? nb[2] : 0;
    L.maxlen = it.len;
    { int r = make_natural(l); if (r) return r; }
    halo_touch(l);
    it.fb = l->current();
    dim3 grid(blocks_for(it.len, 128), 1);
    if (D == 2) LAUNCH(k_smooth<2>, grid, 128, l->g, L); else LAUNCH(k_smooth<3>, grid, 128, l->g, L);
    return PL_OK;
}

pl_bc* pl_bc_create(pl_lattice* l, int type, int axis, int coord, int dir, const uint8_t* mask, const double* v0, const double* v1, const double* v2) {
    if (!l || type < 1 || type > 13 || (dir != -1 && dir != 1) || axis < 0 || axis >= l->kind) { fail(PL_ERR_ARG, "pl_bc_create: bad arguments"); return nullptr; }
    if (type == PL_BC_AAD_ISET_RHO && l->kind == PL_D3Q15) {
        fail(PL_ERR_UNSUPPORTED, "pl_bc_create: the reference's D3Q15 AAD::iBoundaryConditionSetRho does not compile (adjointadvection.h:583); D2Q9 only");
        return nullptr;
    }
    if ((type == PL_BC_NSIN_SET_U || type == PL_BC_NSIN_SET_RHO) && l->kind == PL_D3Q15) {
        fail(PL_ERR_UNSUPPORTED, "pl_bc_create: the reference's NSin closures exist for D2Q9 only (nsincompressible.h:46-154)");
        return nullptr;
    }
    pl_bc* bc = new pl_bc();
    bc->lat = l; bc->type = type; bc->axis = axis; bc->coord = coord; bc->dir = dir;
    bc->empty = !make_plane(l, axis, coord, dir, bc->pl);
    if (!bc->empty) {
        if (!mask) { delete bc; fail(PL_ERR_ARG, "pl_bc_create: mask is required"); return nullptr; }
        size_t np = (size_t)bc->pl.n1*bc->pl.n2;
        bool any = false;
        for (size_t t = 0; t < np; ++t) any = any || mask[t] != 0;
        if (!any) bc->empty = true;
        else {
            auto up8 = [&](const uint8_t* h) -> uint8_t* { uint8_t* d = nullptr; if (cudaMalloc(&d, np) != cudaSuccess) return nullptr; cudaMemcpy(d, h, np, cudaMemcpyHostToDevice); return d; };
            auto upd = [&](const double* h) -> double* { if (!h) return nullptr; double* d = nullptr; if (cudaMalloc(&d, np*sizeof(double)) != cudaSuccess) return nullptr; cudaMemcpy(d, h, np*sizeof(double), cudaMemcpyHostToDevice); return d; };
            bc->mask = up8(mask); bc->v0 = upd(v0); bc->v1 = upd(v1); bc->v2 = upd(v2);
            bc->hmask.assign(mask, mask + np);
            if (!bc->mask || (v0 && !bc->v0) || (v1 && !bc->v1) || (v2 && !bc->v2)) { pl_bc_destroy(bc); fail(PL_ERR_CUDA, "pl_bc_create: device allocation failed"); return nullptr; }
        }
    }
    return bc;
}
int pl_bc_destroy(pl_bc* bc) {
    if (!bc) return PL_OK;
    cudaStreamSynchronize(g_stream);
    cudaFree(bc->mask); cudaFree(bc->v0); cudaFree(bc->v1); cudaFree(bc->v2);
    delete bc;
    return PL_OK;
}
int pl_bc_is_empty(const pl_bc* bc) { return bc ? (bc->empty ? 1 : 0) : 1; }
int pl_bc_update_values(pl_bc* bc, const double* v0, const double* v1, const double* v2) {
    InCall in_call_;
    if (!bc) return fail(PL_ERR_ARG, "pl_bc_update_values: null");
    if (bc->empty) return PL_OK;
    if ((v0 != nullptr) != (bc->v0 != nullptr) || (v1 != nullptr) != (bc->v1 != nullptr) || (v2 != nullptr) != (bc->v2 != nullptr))
        return fail(PL_ERR_ARG, "pl_bc_update_values: the plane was created with other value arrays");
    const size_t nb = (size_t)bc->pl.n1*bc->pl.n2*sizeof(double);
    // pageable source: the runtime stages it before returning, the device copy is ordered behind every pass already queued
    if (v0) CU(cudaMemcpyAsync(bc->v0, v0, nb, cudaMemcpyHostToDevice, g_stream));
    if (v1) CU(cudaMemcpyAsync(bc->v1, v1, nb, cudaMemcpyHostToDevice, g_stream));
    if (v2) CU(cudaMemcpyAsync(bc->v2, v2, nb, cudaMemcpyHostToDevice, g_stream));
    return PL_OK;
}
int pl_bc_apply(pl_lattice* l, pl_lattice* other, const pl_bc* bc, const pl_bc_aux* aux) {
    if (!l || !bc) return fail(PL_ERR_ARG, "pl_bc_apply: null");
    return do_bc(l, other, bc, aux);
}

int pl_collide(pl_lattice* f, pl_lattice* g, const pl_collide_args* a) {
    CollideParams P; unsigned flags;
    int r = make_params(f, g, a, P, flags);
    if (r) return r;
    if (!(flags & F_G)) g = nullptr;
    r = dispatch_collide(a->model, f, g, P, nullptr, f->g.nxyz);
    if (r) return r;
    f->streamed = 0; if (g) g->streamed = 0;
    halo_touch(f); if (g) halo_touch(g);
    return PL_OK;
}

int pl_snapshot_to_host(const pl_lattice* l, const double* snap, double* out) {
    InCall in_call_;
    if (!l || !snap || !out) return fail(PL_ERR_ARG, "pl_snapshot_to_host: null");
    size_t n = (size_t)l->g.nxyz*l->nc;
    double* d = nullptr;
    CU(cudaMalloc(&d, n*sizeof(double)));
    if (l->kind == PL_D2Q9) LAUNCH(k_snapshot_to_ref<2>, blocks_for(l->g.nxyz, 256), 256, l->g, snap, (size_t)l->g.nxyz, d);
    else LAUNCH(k_snapshot_to_ref<3>, blocks_for(l->g.nxyz, 256), 256, l->g, snap, (size_t)l->g.nxyz, d);
    CU(cudaMemcpyAsync(out, d, n*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    cudaFree(d);
    return PL_OK;
}

// snapshot layout conversion for a lattice shape given by (kind, nxyz) alone: the host-pointer surface keeps converting a caller's
// snapshot array after the lattice that produced it is gone
int pl_snapshot_convert(int kind, long long nxyz, const double* in, double* out, int to_host) {
    InCall in_call_;
    if ((kind != PL_D2Q9 && kind != PL_D3Q15) || nxyz <= 0 || !in || !out) return fail(PL_ERR_ARG, "pl_snapshot_convert: bad arguments");
    Geom g;
    memset(&g, 0, sizeof(g));
    g.nxyz = nxyz; g.npacked = packed_sites(nxyz);
    const size_t n = (size_t)nxyz*(kind == PL_D2Q9 ? 9 : 15);
    double* d = nullptr;
    CU(cudaMalloc(&d, n*sizeof(double)));
    if (to_host) {
        if (kind == PL_D2Q9) LAUNCH(k_snapshot_to_ref<2>, blocks_for(nxyz, 256), 256, g, in, (size_t)nxyz, d);
        else LAUNCH(k_snapshot_to_ref<3>, blocks_for(nxyz, 256), 256, g, in, (size_t)nxyz, d);
        CU(cudaMemcpyAsync(out, d, n*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    } else {
        CU(cudaMemcpyAsync(d, in, n*sizeof(double), cudaMemcpyHostToDevice, g_stream));
        if (kind == PL_D2Q9) LAUNCH(k_snapshot_from_ref<2>, blocks_for(nxyz, 256), 256, g, d, out, (size_t)nxyz);
        else LAUNCH(k_snapshot_from_ref<3>, blocks_for(nxyz, 256), 256, g, d, out, (size_t)nxyz);
    }
    CU(cudaStreamSynchronize(g_stream));
    cudaFree(d);
    return PL_OK;
}
int pl_snapshot_from_host(const pl_lattice* l, const double* in_host, double* snap) {
    if (!l) return fail(PL_ERR_ARG, "pl_snapshot_from_host: null");
    return pl_snapshot_convert(l->kind, l->g.nxyz, in_host, snap, 0);
}

int pl_initial_condition(pl_lattice* l, int family, const double* const* a, int na) {
    if (!l || !a || family < 1 || family > 5) return fail(PL_ERR_ARG, "pl_initial_condition: bad arguments");
    if (family == 5 && l->kind == PL_D3Q15) return fail(PL_ERR_UNSUPPORTED, "pl_initial_condition: the reference's NSin equations exist for D2Q9 only");
    const int want = (family <= 2 || family == 5) ? 4 : 7;
    if (na < want) return fail(PL_ERR_ARG, "pl_initial_condition: too few arrays");
    const double* p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int n = 0; n < want; ++n) p[n] = a[n];
    const bool d3 = l->kind == PL_D3Q15;
    // z-components may be null on D2Q9
    for (int n = 0; n < want; ++n) {
        bool zslot = (family <= 2 || family == 5) ? n == 3 : (n == 2 || n == 6);
        if (!p[n] && !(zslot && !d3)) return fail(PL_ERR_ARG, "pl_initial_condition: null array");
    }
    l->rep = 0;      // every population is overwritten: whatever layout the buffer was in is irrelevant
    if (d3) LAUNCH(k_init<3>, blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), family, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
    else LAUNCH(k_init<2>, blocks_for(l->g.nxyz, 256), 256, l->g, l->current(), family, p[0], p[1], p[2], p[3], p[4], p[5], p[6]);
    l->streamed = 1;
    halo_touch(l);
    return PL_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// plans
struct PlanBC { int on_g; const pl_bc* bc; pl_bc_aux aux[2]; bool has_aux; int pos; };
struct PlanSmoothAt { int on_g, i, j, k, dx, dy, dz, pos; };      // SmoothCornerAt(i, j, k, dx, dy, dz) in global coordinates
struct pl_plan {
    pl_lattice *f, *g;
    pl_collide_args args[2];
    bool have_collide = false;
    int inverse = 0;
    std::vector<PlanBC> bcs;
    int smooth_f = 0, smooth_g = 0;
    // call order of the loop body: every closure, SmoothCorner and SmoothCornerAt gets the next sequence number
    int seq = 0;
    int smooth_pos[2] = {-1, -1};
    std::vector<PlanSmoothAt> ats;
    bool finalized = false;
    int parity = 0;
    // per-coordinate plane words (see ShellMask), the site list of k_shell (closure-plane and AVX-tail sites first,
    // SmoothCorner tube sites last), the closure program per argument-set parity
    unsigned long long *mx = nullptr, *my = nullptr, *mz = nullptr;
    int* list = nullptr;
    unsigned long long* ent = nullptr;     // closure entries of each listed site (the plane words of its coordinates, OR-ed)
    double *tube_f = nullptr, *tube_g = nullptr;   // [c][ntube] streamed+closed populations of the SmoothCorner tube sites
    TubeSite* tube_info = nullptr;
    int* xlist = nullptr;                  // sites of the x boundary planes whose closures run ahead of the pass (k_xclose)
    unsigned long long* xent = nullptr;
    int nxlist = 0;
    XNeed xneed = {};
    // compact wall buffers of those planes (XWall): per lattice two `out` buffers, alternating with the population buffer the pass
    // reads (a pass writes the one the next pass reads while k_xclose of this pass may still be reading the other), one `res`
    double *xout[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}}, *xres[2] = {nullptr, nullptr};
    int xon[2] = {0, 0};
    uint64_t xver[2] = {~0ull, ~0ull};     // lattice versions the buffers were left for by the plan's own last pass
    int nlist = 0, ndirect = 0;
    ClosureArgs* prog[2] = {nullptr, nullptr};
    int nprog = 0;
    // pl_plan_rebind: pinned staging ring for the re-built closure programs (stream-ordered copies, no host synchronisation
    // unless the ring wraps onto a copy still in flight)
    static constexpr int NSTAGE = 8;
    ClosureArgs* stage = nullptr;
    cudaEvent_t stage_ev[NSTAGE] = {};
    int stage_next = 0;
    // the boundary pass runs beside the interior kernel on its own (high-priority) stream
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // captured fused steps, by [argument set of the closures][pass mode][wall-buffer phase][every site stores / only the closure planes]
    struct Graph { cudaGraphExec_t exec = nullptr; uint64_t launches = 0; const double *fbuf = nullptr, *gbuf = nullptr; };
    Graph graphs[2][3][2][2];
    int xphase = 0;                // which of the two wall `out` buffers the next pass reads
    cudaStream_t cap = nullptr;
    int graph_cooldown = 0;        // steps to run ungraphed after the arguments were re-bound (per-step arrays: nothing to replay)
    // measurement hook
    bool profile = false;
    struct ProfEv { cudaEvent_t a, b; int cls; long long sites; };   // cls 0 = a pass that stores only on the closure planes, 1 = every site stores
    std::vector<ProfEv> events;
};

namespace {
int plan_stream_bc_smooth_full(pl_plan* p, int parity) {     // standalone S: every site
    int r;
    if ((r = do_stream_all(p->f, p->inverse))) return r;
    if (p->g && (r = do_stream_all(p->g, p->inverse))) return r;
    for (auto& b : p->bcs) {
        pl_lattice* l = b.on_g ? p->g : p->f;
        pl_lattice* o = b.bc->type == PL_BC_AAD_ISET_RHO ? p->g : nullptr;
        if ((r = do_bc(l, o, b.bc, b.has_aux ? &b.aux[parity] : nullptr))) return r;
    }
    if (p->smooth_f && (r = do_smooth(p->f))) return r;
    if (p->g && p->smooth_g && (r = do_smooth(p->g))) return r;
    // pl_plan_finalize accepted the body only if this order (closures, SmoothCorner, SmoothCornerAt) equals the call order
    for (auto& a : p->ats) if ((r = pl_smooth_corner_at(a.on_g ? p->g : p->f, a.i, a.j, a.k, a.dx, a.dy, a.dz))) return r;
    p->f->streamed = 1; if (p->g) p->g->streamed = 1;
    return PL_OK;
}
int plan_collide_full(pl_plan* p, int parity) {              // standalone C: every site, in place
    CollideParams P; unsigned flags;
    int r = make_params(p->f, p->g, &p->args[parity], P, flags);
    if (r) return r;
    pl_lattice* g = (flags & F_G) ? p->g : nullptr;
    if ((r = dispatch_collide(p->args[parity].model, p->f, g, P, nullptr, p->f->g.nxyz))) return r;
    p->f->streamed = 0; if (p->g) p->g->streamed = 0;
    // post the exchange the Stream of this step needs right away: it overlaps whatever is queued next
    halo_touch(p->f); if (p->g) halo_touch(p->g);
    if ((r = halo_prepare(p->f, p->inverse, true))) return r;
    if (p->g && (r = halo_prepare(p->g, p->inverse, true))) return r;
    return PL_OK;
}
void drop_graphs(pl_plan* p) {
    for (auto& a : p->graphs) for (auto& b : a) for (auto& c : b) for (auto& g : c) if (g.exec) { cudaGraphExecDestroy(g.exec); g = pl_plan::Graph(); }
}
// which kind of pass comes next (lbm_kernels.cuh, pass modes), bringing the two lattices to a common layout first if needed
int plan_pass_mode(pl_plan* p, int& mode) {
    int r;
    const bool xstale = p->nxlist && (p->xver[0] != p->f->version || (p->g && p->xver[1] != p->g->version));
    if (!opt_inplace() || xstale) {      // the wall buffers are refilled from the natural layout
        if ((r = make_natural(p->f)) || (p->g && (r = make_natural(p->g)))) return r;
    }
    if (!opt_inplace()) { mode = PASS_COPY; return PL_OK; }
    for (pl_lattice* l : {p->f, p->g}) if (l && l->rep && l->rep_inverse != p->inverse && (r = make_natural(l))) return r;
    if (p->g && p->f->rep != p->g->rep && ((r = make_natural(p->f)) || (r = make_natural(p->g)))) return r;
    mode = p->f->rep ? PASS_LOCAL : PASS_GATHER;
    return PL_OK;
}
void plan_pass_done(pl_plan* p, int mode) {
    for (pl_lattice* l : {p->f, p->g}) {
        if (!l) continue;
        if (mode == PASS_GATHER) { l->rep = 1; l->rep_inverse = p->inverse; }
        else if (mode == PASS_LOCAL) l->rep = 0;
        l->streamed = 0;
    }
    p->xphase ^= 1;
}
int plan_fused_body(pl_plan* p, int bc_parity, int col_parity, bool full_save, int mode);
// fused F: Stream + closures + SmoothCorner of step t (argument set `bc_parity`) followed by the collide of step t+1 —
// through a captured graph where that is possible: single block (no NCCL inside), in place, not being profiled, arguments stable.
// full_save: every site stores what _issave asks for; else only the sites on closure planes do (pl_plan_advance_observed)
int plan_fused(pl_plan* p, int bc_parity, int col_parity, bool full_save) {
    const bool xstale = p->nxlist && (p->xver[0] != p->f->version || (p->g && p->xver[1] != p->g->version));      // the one-off refill must not be captured
    int mode, r;
    if ((r = plan_pass_mode(p, mode))) return r;
    if (mode != PASS_COPY) g_spares.inplace_pass();      // (never under stream capture: it may synchronise and free)
    if (!opt_graph() || p->f->halo.on || p->profile || opt_shell_serial() || xstale || mode == PASS_COPY) return plan_fused_body(p, bc_parity, col_parity, full_save, mode);
    if (p->graph_cooldown > 0) { --p->graph_cooldown; return plan_fused_body(p, bc_parity, col_parity, full_save, mode); }
    pl_plan::Graph& G = p->graphs[bc_parity][mode][p->xphase][full_save ? 1 : 0];
    if (G.exec && (G.fbuf != p->f->buf || G.gbuf != (p->g ? p->g->buf : nullptr))) { cudaGraphExecDestroy(G.exec); G = pl_plan::Graph(); }   // a Stream() swapped the buffers
    if (G.exec) {
        CU(cudaGraphLaunch(G.exec, g_stream));
        g_launches += G.launches;
        plan_pass_done(p, mode);      // the host-side state changes of plan_fused_body (the content version stays in step with xver)
        return PL_OK;
    }
    if (!p->cap) CU(cudaStreamCreateWithFlags(&p->cap, cudaStreamNonBlocking));
    cudaStream_t user = g_stream;
    const uint64_t before = g_launches;
    G.fbuf = p->f->buf; G.gbuf = p->g ? p->g->buf : nullptr;
    g_stream = p->cap;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(p->cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) { g_stream = user; cudaGetLastError(); return plan_fused_body(p, bc_parity, col_parity, full_save, mode); }
    r = plan_fused_body(p, bc_parity, col_parity, full_save, mode);
    cudaError_t e = cudaStreamEndCapture(p->cap, &graph);
    g_stream = user;
    if (r) { if (graph) cudaGraphDestroy(graph); return r; }
    if (e != cudaSuccess || !graph) return fail(PL_ERR_CUDA, std::string("fused step: graph capture failed: ") + cudaGetErrorString(e));
    e = cudaGraphInstantiate(&G.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { G.exec = nullptr; return fail(PL_ERR_CUDA, std::string("fused step: cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
    G.launches = g_launches - before;
    CU(cudaGraphLaunch(G.exec, g_stream));      // the body ran under capture: this launch is its execution
    return PL_OK;
}
int plan_fused_body(pl_plan* p, int bc_parity, int col_parity, bool full_save, int mode) {
    CollideParams P; unsigned flags;
    int r = make_params(p->f, p->g, &p->args[col_parity], P, flags);
    if (r) return r;
    if (P.issave && !full_save) P.issave = 2;
    const int model = p->args[col_parity].model;
    pl_lattice* g = (flags & F_G) ? p->g : nullptr;
    if (p->g && !g) return fail(PL_ERR_ARG, "plan: a single-lattice collide cannot drive a two-lattice plan");
    const ModelLaunch* ml = launcher(p->f, model);
    if (!ml) return PL_ERR_UNSUPPORTED;
    // halo of a decomposed block: normally posted already by the collide that produced these populations
    if ((r = halo_prepare(p->f, p->inverse))) return r;
    if (p->g && (r = halo_prepare(p->g, p->inverse))) return r;
    FusedArgs A;
    memset(&A, 0, sizeof(A));
    A.G = p->f->g; A.P = P; A.S = ShellMask{p->mx, p->my, p->mz, opt_prefetch(), opt_l2_ahead()*PLK_FUSED_THREADS}; A.inverse = p->inverse;
    A.fs = p->f->buf; A.gs = g ? g->buf : nullptr;
    double *fdst = p->f->buf, *gdst = g ? g->buf : nullptr;
    if (mode == PASS_COPY) {
        fdst = g_spares.get(p->f->bytes());
        gdst = g ? g_spares.get(g->bytes()) : nullptr;
        if (!fdst || (g && !gdst)) return fail(PL_ERR_CUDA, "out of device memory for the second population buffer (PANSLBM_INPLACE=0)");
    }
    A.fd = fdst; A.gd = gdst;
    A.pipe = opt_pipe() ? 1 : 0; A.sms = device_sms();
    A.list = p->list; A.ent = p->ent; A.nlist = p->nlist; A.ndirect = p->ndirect; A.tube_f = p->tube_f; A.tube_g = p->tube_g; A.tube_info = p->tube_info;
    if ((r = halo_view(p->f, A.HF))) return r;
    if (g && (r = halo_view(g, A.HG))) return r;
    // compact wall buffers of the x boundary planes (XWall): this pass reads xout[.][xphase] and fills xout[.][xphase ^ 1]
    if (p->nxlist) {
        A.W.out_f = p->xout[0][p->xphase ^ 1]; A.W.res_f = p->xres[0];
        if (g) { A.W.out_g = p->xout[1][p->xphase ^ 1]; A.W.res_g = p->xres[1]; }
        A.W.np = p->f->g.ny*p->f->g.nz; A.W.on[0] = p->xon[0]; A.W.on[1] = p->xon[1];
    }
    // boundary pass on the side stream (closure planes, block faces, SmoothCorner tubes, AVX-tail sites): it touches only
    // locations the interior kernel leaves alone, so the two run concurrently
    const bool serial = !p->f->halo.on && opt_shell_serial();
    if (!serial) {
        CU(cudaEventRecord(p->ev_fork, g_stream));
        CU(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
        if ((r = halo_wait(p->f, p->side))) return r;
        if (p->g && (r = halo_wait(p->g, p->side))) return r;
    }
    auto shell = [&](cudaStream_t st) -> int {
        A.prog = p->prog[bc_parity];
        if (p->nlist == 0) return PL_OK;
        g_launches += p->nlist > p->ndirect ? 2 : 1;
        CU(ml->shell(st, A, mode));
        return PL_OK;
    };
    // closures of the x boundary planes on the compact wall buffers (the boundary pass may run beside this)
    if (p->nxlist) {
        const int rd = p->xphase, np = A.W.np, nc = p->f->nc;
        // wall buffers left by something else than this plan's own last pass (first pass after a standalone collide): refill
        for (int l = 0; l < (g ? 2 : 1); ++l) {
            pl_lattice* q = l ? g : p->f;
            if (p->xver[l] == q->version) continue;
            if (q->rep) return fail(PL_ERR_ARG, "plan: internal error (wall buffers stale on a streamed layout)");
            dim3 grid(blocks_for(np, 128), 2*nc);
            if (q->kind == PL_D2Q9) LAUNCH(k_xfill<2>, grid, 128, q->g, q->current(), p->xout[l][rd], np, p->inverse, p->xon[0], p->xon[1]);
            else LAUNCH(k_xfill<3>, grid, 128, q->g, q->current(), p->xout[l][rd], np, p->inverse, p->xon[0], p->xon[1]);
        }
        const int nb = (int)blocks_for(p->nxlist, SHELL_THREADS);
        const double *inf = p->xout[0][rd], *ing = g ? p->xout[1][rd] : nullptr;
        double *rf = p->xres[0], *rg = g ? p->xres[1] : nullptr;
        if (p->f->kind == PL_D2Q9) {
            if (g) LAUNCH((k_xclose<2, true>), nb, SHELL_THREADS, p->f->g, inf, ing, rf, rg, np, p->prog[bc_parity], p->xlist, p->xent, p->nxlist, p->inverse, p->xneed);
            else LAUNCH((k_xclose<2, false>), nb, SHELL_THREADS, p->f->g, inf, ing, rf, rg, np, p->prog[bc_parity], p->xlist, p->xent, p->nxlist, p->inverse, p->xneed);
        } else {
            if (g) LAUNCH((k_xclose<3, true>), nb, SHELL_THREADS, p->f->g, inf, ing, rf, rg, np, p->prog[bc_parity], p->xlist, p->xent, p->nxlist, p->inverse, p->xneed);
            else LAUNCH((k_xclose<3, false>), nb, SHELL_THREADS, p->f->g, inf, ing, rf, rg, np, p->prog[bc_parity], p->xlist, p->xent, p->nxlist, p->inverse, p->xneed);
        }
    }
    // a decomposed block wants its faces first (the next exchange hangs on them); a single block may queue the interior first
    const bool shell_first = !serial && (p->f->halo.on || !opt_fused_first());
    if (shell_first) {
        if ((r = shell(p->side))) return r;
        CU(cudaEventRecord(p->ev_join, p->side));
    }
    // interior: one pass
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (p->profile) {
        CU(cudaEventCreate(&ev0)); CU(cudaEventCreate(&ev1));
        CU(cudaEventRecord(ev0, g_stream));
    }
    A.prog = opt_xinline() ? p->prog[bc_parity] : nullptr;
    if (p->f->g.npacked > 0) { ++g_launches; CU(ml->fused(g_stream, A, mode)); }
    if (p->profile) {
        CU(cudaEventRecord(ev1, g_stream));
        p->events.push_back(pl_plan::ProfEv{ev0, ev1, P.issave == 2 ? 0 : 1, (long long)(p->f->g.nxyz - p->nlist)});
    }
    if (serial) {
        if ((r = shell(g_stream))) return r;
    } else {
        if (!shell_first) {
            if ((r = shell(p->side))) return r;
            CU(cudaEventRecord(p->ev_join, p->side));
        }
        CU(cudaStreamWaitEvent(g_stream, p->ev_join, 0));
    }
    if (mode == PASS_COPY) {
        g_spares.put(p->f->buf, p->f->bytes()); p->f->buf = fdst;
        if (g) { g_spares.put(g->buf, g->bytes()); g->buf = gdst; }
    }
    plan_pass_done(p, mode);
    // every block-face site is final: pack and post the next exchange now, it overlaps the next interior kernel
    halo_touch(p->f); if (p->g) halo_touch(p->g);
    p->xver[0] = p->f->version; if (p->g) p->xver[1] = p->g->version;      // the wall buffers describe exactly these populations
    if ((r = halo_prepare(p->f, p->inverse, true))) return r;
    if (p->g && (r = halo_prepare(p->g, p->inverse, true))) return r;
    return PL_OK;
}
// `n` fused passes in one cooperative launch (lbm_steps.cuh), for single blocks that live in L2.  false in *done: not applicable
// (the caller runs the passes one by one).
int plan_steps(pl_plan* p, int n, int save_last_of_call, int remaining_after, bool* done) {
    *done = false;
    if (n < 2 || opt_coop_sites() <= 0 || p->f->g.nxyz > opt_coop_sites() || !opt_inplace() || p->f->halo.on || p->profile || opt_graph() || opt_xinline() ||
        opt_shell_serial() || opt_pipe()) return PL_OK;
    int mode, r;
    if ((r = plan_pass_mode(p, mode))) return r;
    if (mode == PASS_COPY) return PL_OK;
    if (p->nxlist && (p->xver[0] != p->f->version || (p->g && p->xver[1] != p->g->version))) return PL_OK;     // the first pass refills the wall buffers
    CollideParams P0, P1; unsigned flags;
    if ((r = make_params(p->f, p->g, &p->args[0], P0, flags)) || (r = make_params(p->f, p->g, &p->args[1], P1, flags))) return r;
    pl_lattice* g = (flags & F_G) ? p->g : nullptr;
    if (p->g && !g) return fail(PL_ERR_ARG, "plan: a single-lattice collide cannot drive a two-lattice plan");
    const ModelLaunch* ml = launcher(p->f, p->args[0].model);
    if (!ml) return PL_ERR_UNSUPPORTED;
    StepsArgs A;
    memset(&A, 0, sizeof(A));
    A.G = p->f->g; A.f = p->f->buf; A.g = g ? g->buf : nullptr;
    A.P[0] = P0; A.P[1] = P1; A.prog[0] = p->prog[0]; A.prog[1] = p->prog[1];
    A.S = ShellMask{p->mx, p->my, p->mz, 0, 0}; A.inverse = p->inverse;
    A.list = p->list; A.ent = p->ent; A.nlist = p->nlist; A.ndirect = p->ndirect; A.tube_f = p->tube_f; A.tube_g = p->tube_g; A.tube_info = p->tube_info;
    A.xlist = p->xlist; A.xent = p->xent; A.nxlist = p->nxlist; A.xneed = p->xneed;
    for (int b = 0; b < 2; ++b) { A.xout_f[b] = p->xout[0][b]; A.xout_g[b] = g ? p->xout[1][b] : nullptr; }
    A.xres_f = p->xres[0]; A.xres_g = g ? p->xres[1] : nullptr;
    A.np = p->f->g.ny*p->f->g.nz; A.xon[0] = p->xon[0]; A.xon[1] = p->xon[1];
    A.nsteps = n; A.parity = p->parity; A.mode = mode; A.xphase = p->xphase;
    // passes of this launch that store everywhere: the last save_last collides of the whole pl_plan_advance call
    A.save_last = save_last_of_call < 0 ? -1 : std::max(0, save_last_of_call - remaining_after);
    g_spares.inplace_pass();
    int grid = 0;
    cudaError_t e = ml->steps(g_stream, A, device_sms(), &grid);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) { cudaGetLastError(); return PL_OK; }
    CU(e);
    ++g_launches;
    for (int s = 0; s < n; ++s) { plan_pass_done(p, mode); mode = mode == PASS_GATHER ? PASS_LOCAL : PASS_GATHER; p->parity ^= 1; }
    halo_touch(p->f); if (p->g) halo_touch(p->g);
    p->xver[0] = p->f->version; if (p->g) p->xver[1] = p->g->version;
    *done = true;
    return PL_OK;
}
}  // namespace

extern "C" {

pl_plan* pl_plan_create(pl_lattice* f, pl_lattice* g) {
    if (!f) { fail(PL_ERR_ARG, "pl_plan_create: null lattice"); return nullptr; }
    if (g && !same_shape(f, g)) { fail(PL_ERR_ARG, "pl_plan_create: lattices differ in shape"); return nullptr; }
    pl_plan* p = new pl_plan();
    p->f = f; p->g = g;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        fail(PL_ERR_CUDA, std::string("pl_plan_create: ") + cudaGetErrorString(cudaGetLastError()));
        delete p;
        return nullptr;
    }
    return p;
}
int pl_plan_destroy(pl_plan* p) {
    if (!p) return PL_OK;
    cudaStreamSynchronize(g_stream);
    cudaFree(p->mx); cudaFree(p->my); cudaFree(p->mz); cudaFree(p->list); cudaFree(p->ent); cudaFree(p->xlist); cudaFree(p->xent); cudaFree(p->prog[0]); cudaFree(p->prog[1]);
    cudaFree(p->tube_f); cudaFree(p->tube_g); cudaFree(p->tube_info);
    for (int l = 0; l < 2; ++l) { cudaFree(p->xout[l][0]); cudaFree(p->xout[l][1]); cudaFree(p->xres[l]); }
    if (p->stage) cudaFreeHost(p->stage);
    for (auto& e : p->stage_ev) if (e) cudaEventDestroy(e);
    drop_graphs(p);
    if (p->cap) cudaStreamDestroy(p->cap);
    for (auto& e : p->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (p->side) { cudaStreamSynchronize(p->side); cudaStreamDestroy(p->side); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    if (p->ev_join) cudaEventDestroy(p->ev_join);
    delete p;
    return PL_OK;
}
int pl_plan_set_collide(pl_plan* p, const pl_collide_args* even, const pl_collide_args* odd) {
    if (!p || !even) return fail(PL_ERR_ARG, "pl_plan_set_collide: null");
    drop_graphs(p);
    p->args[0] = *even; p->args[1] = odd ? *odd : *even;
    if (p->args[0].model != p->args[1].model) return fail(PL_ERR_ARG, "pl_plan_set_collide: the two argument sets must use the same model");
    p->have_collide = true;
    return PL_OK;
}
int pl_plan_set_stream(pl_plan* p, int inverse) { if (!p) return fail(PL_ERR_ARG, "null plan"); p->inverse = inverse ? 1 : 0; return PL_OK; }
int pl_plan_add_bc(pl_plan* p, int on_g, const pl_bc* bc, const pl_bc_aux* even, const pl_bc_aux* odd) {
    if (!p || !bc) return fail(PL_ERR_ARG, "pl_plan_add_bc: null");
    if (p->finalized) return fail(PL_ERR_ARG, "pl_plan_add_bc: plan already finalized");
    if (on_g && !p->g) return fail(PL_ERR_ARG, "pl_plan_add_bc: plan has no thermal lattice");
    if (bc->type == PL_BC_AAD_ISET_RHO && (on_g || !p->g)) return fail(PL_ERR_ARG, "pl_plan_add_bc: AAD iSetRho acts on the flow lattice of a two-lattice plan");
    if (!same_shape(bc->lat, p->f)) return fail(PL_ERR_ARG, "pl_plan_add_bc: closure was created for a lattice of another shape");
    PlanBC b{};
    b.on_g = on_g ? 1 : 0; b.bc = bc; b.has_aux = even != nullptr; b.pos = p->seq++;
    if (even) { b.aux[0] = *even; b.aux[1] = odd ? *odd : *even; }
    p->bcs.push_back(b);
    return PL_OK;
}
int pl_plan_set_smooth_corner(pl_plan* p, int on_f, int on_g) {
    if (!p) return fail(PL_ERR_ARG, "null plan");
    if (p->finalized) return fail(PL_ERR_ARG, "pl_plan_set_smooth_corner: plan already finalized");
    if (on_f && !p->smooth_f) p->smooth_pos[0] = p->seq++;       // position in the loop body = where the flag is first raised
    if (on_g && !p->smooth_g) p->smooth_pos[1] = p->seq++;
    p->smooth_f = on_f; p->smooth_g = on_g;
    return PL_OK;
}
int pl_plan_add_smooth_corner_at(pl_plan* p, int on_g, int i, int j, int k, int dx, int dy, int dz) {
    if (!p) return fail(PL_ERR_ARG, "null plan");
    if (p->finalized) return fail(PL_ERR_ARG, "pl_plan_add_smooth_corner_at: plan already finalized");
    if (on_g && !p->g) return fail(PL_ERR_ARG, "pl_plan_add_smooth_corner_at: plan has no thermal lattice");
    p->ats.push_back(PlanSmoothAt{on_g ? 1 : 0, i, j, k, dx, dy, dz, p->seq++});
    return PL_OK;
}
}  // extern "C"

namespace {
// geometry of one SmoothCornerAt on this rank's block (the arithmetic of pl_smooth_corner_at): the sites it writes (a point, or in
// 3-D with one zero direction a whole line), the index deltas to the 2 or 3 inward neighbours, and per axis the two local
// coordinates involved.  false: not on this block (d3q15.h:1244-1245) or malformed.
struct AtGeom { std::vector<long long> sites; long long nb[3]; int nnb; int co[3][2]; bool used[3]; };
bool smooth_at_geometry(const pl_lattice* l, const PlanSmoothAt& a, AtGeom& G) {
    const Geom& g = l->g;
    const int D = l->kind;
    int d[3] = {a.dx, a.dy, D == 3 ? a.dz : 0}, v[3] = {a.i - g.offx, a.j - g.offy, D == 3 ? a.k - g.offz : 0}, n[3] = {g.nx, g.ny, g.nz};
    long long st[3] = {1, g.nx, (long long)g.nx*g.ny};
    int nzero = 0, line = -1;
    for (int x = 0; x < 3; ++x) {
        if (d[x] != 0 && d[x] != 1 && d[x] != -1) return false;
        if (d[x] == 0) { ++nzero; if (x < D) line = x; }
    }
    if ((D == 3 && nzero > 1) || (D == 2 && nzero != 1)) return false;
    for (int x = 0; x < D; ++x) if (d[x] != 0 && (v[x] < 0 || v[x] >= n[x])) return false;
    long long base = 0;
    G.nnb = 0;
    for (int x = 0; x < 3; ++x) G.used[x] = false;
    for (int x = 0; x < D; ++x) if (d[x] != 0) {
        int w = v[x] - d[x];
        if (w == -1) w = n[x] - 1; else if (w == n[x]) w = 0;
        base += v[x]*st[x];
        G.nb[G.nnb++] = (long long)(w - v[x])*st[x];
        G.used[x] = true; G.co[x][0] = v[x]; G.co[x][1] = w;
    }
    G.sites.clear();
    if (D == 3 && line >= 0) for (int t = 0; t < n[line]; ++t) G.sites.push_back(base + (long long)t*st[line]);
    else G.sites.push_back(base);
    return true;
}
}  // namespace
extern "C" {
int pl_plan_finalize(pl_plan* p) {
    if (!p || !p->have_collide) return fail(PL_ERR_ARG, "pl_plan_finalize: no collide set");
    const Geom& g = p->f->g;
    std::vector<unsigned long long> hx(g.nx, 0), hy(g.ny, 0), hz(g.nz, 0);
    std::vector<unsigned long long>* h[3] = {&hx, &hy, &hz};
    int off[3] = {g.offx, g.offy, g.offz};
    // closure program: the non-empty closures in call order, one copy per argument-set parity
    std::vector<ClosureArgs> prog[2];
    for (auto& b : p->bcs) {
        if (b.bc->empty) continue;
        if (prog[0].size() >= (size_t)MAX_PROGRAM) return fail(PL_ERR_UNSUPPORTED, "pl_plan_finalize: more than 60 non-empty closures in one loop body");
        (*h[b.bc->axis])[b.bc->coord - off[b.bc->axis]] |= 1ull << prog[0].size();
        for (int par = 0; par < 2; ++par) {
            ClosureArgs A;
            pl_lattice* l = b.on_g ? p->g : p->f;
            int r = make_closure_args(l, b.bc->type == PL_BC_AAD_ISET_RHO ? p->g : nullptr, b.bc, b.has_aux ? &b.aux[par] : nullptr, A);
            if (r) return r;
            A.on_g = b.on_g; A.loc = b.bc->coord - off[b.bc->axis];
            prog[par].push_back(A);
        }
    }
    // faces of a decomposed block: their sites pull from the halo receive buffers
    {
        int nn[3] = {g.nx, g.ny, g.nz};
        for (int a = 0; a < p->f->kind; ++a)
            if (p->f->halo.on && p->f->halo.e[a]) { (*h[a])[0] |= HALO_BIT; (*h[a])[nn[a] - 1] |= HALO_BIT; }
    }
    // x planes the boundary pass owns (block faces; closure planes unless the interior kernel takes them inline): it takes the
    // aligned group of x-coordinates around each (see ShellMask)
    // ... except the x boundary planes whose closures can run ahead of the pass (k_xclose): a plane at the wall of an
    // undecomposed x axis, every closure on it rebuilding exactly the populations the plan's Stream direction pulls through the
    // periodic wrap (forward closures with Stream, the "i" closures with iStream)
    std::vector<char> ghost(g.nx, 0);
    if (opt_xghost() && !opt_xinline() && g.nx >= 4 && g.nx == g.lx && !(p->f->halo.on && p->f->halo.e[0])) {
        for (int i : {0, g.nx - 1}) {
            if (!(hx[i] & ENTRY_BITS)) continue;
            bool ok = true;
            for (size_t e = 0; e < prog[0].size() && ok; ++e) {
                if (!((hx[i] >> e) & 1ull)) continue;
                const ClosureArgs& A = prog[0][e];
                const bool fwd = A.type == BC_BOUNCE || A.type == BC_NS_SET_U || A.type == BC_NS_SET_RHO || A.type == BC_AD_SET_T || A.type == BC_AD_SET_Q ||
                                 A.type == BC_NSIN_SET_U || A.type == BC_NSIN_SET_RHO;
                ok = A.pl.axis == 0 && A.pl.dir == (i == 0 ? -1 : 1) && fwd == (p->inverse == 0);
            }
            ghost[i] = ok;
        }
    }
    for (int i = 0; i < g.nx; ++i)
        if ((hx[i] & HALO_BIT) || ((hx[i] & ENTRY_BITS) && !opt_xinline() && !ghost[i])) {
            const int w = opt_xslab(), lo = i/w*w;
            for (int v = lo; v < std::min(g.nx, lo + w); ++v) hx[v] |= SLAB_BIT;
        }
    for (int i : {0, g.nx - 1}) if (hx[i] & SLAB_BIT) ghost[i] = 0;     // a neighbouring plane pulled it into a group
    for (int i : {0, g.nx - 1}) if (ghost[i]) hx[i] |= GHOST_BIT;
    // the directions those closures read (lbm_closures.cuh): bounce-back its sources, SetU/SetRho/SetT everything but the
    // incoming set, SetQ the outgoing set, the adjoint closures their known set K
    p->xneed = XNeed{};
    for (int side = 0; side < 2; ++side) {
        const int i = side ? g.nx - 1 : 0, dir = side ? 1 : -1;
        if (!ghost[i]) continue;
        auto set_of = [&](int want, bool equal) {      // directions with c_x == want (equal) or c_x != want
            unsigned m = 0;
            for (int c = 0; c < p->f->nc; ++c) {
                const int cx = p->f->kind == PL_D2Q9 ? LT<2>::cx(c) : LT<3>::cx(c);
                if ((cx == want) == equal) m |= 1u << c;
            }
            return m;
        };
        for (size_t e = 0; e < prog[0].size(); ++e) {
            if (!((hx[i] >> e) & 1ull)) continue;
            const ClosureArgs& A = prog[0][e];
            unsigned own = 0, other = 0;
            switch (A.type) {
                case BC_BOUNCE: own = set_of(dir, true); break;
                case BC_IBOUNCE: own = set_of(-dir, true); break;
                case BC_NS_SET_U: case BC_NS_SET_RHO: case BC_AD_SET_T: case BC_NSIN_SET_U: case BC_NSIN_SET_RHO: own = set_of(-dir, false); break;
                case BC_AD_SET_Q: own = set_of(dir, true); break;
                case BC_AAD_ISET_RHO: own = set_of(-dir, true); other = own; break;
                default: own = set_of(-dir, true); break;      // ANS iSetU/iSetRho, AAD iSetT/iSetQ: the known set K
            }
            if (A.on_g) { p->xneed.g[side] |= own; p->xneed.f[side] |= other; }
            else { p->xneed.f[side] |= own; p->xneed.g[side] |= other; }
        }
    }
    if (env_int("PANSLBM_XNEED_ALL", 0)) p->xneed = XNeed{{0x7fffu, 0x7fffu}, {0x7fffu, 0x7fffu}};
    // SmoothCorner: flag the global boundary planes and their inward neighbours (bit 1); sites with two flagged coordinates
    // form the edge tubes.  Every site SmoothCorner writes (edge lines, corners) or reads (their inward neighbours) must lie
    // in a tube: collide is deferred there until k_smooth has run.
    int n[3] = {g.nx, g.ny, g.nz}, ext[3] = {g.lx, g.ly, g.lz};
    if (p->smooth_f || p->smooth_g) {
        for (int a = 0; a < p->f->kind; ++a) {
            int lo = 0 - off[a], hi = ext[a] - 1 - off[a];
            for (int v : {lo, lo + 1, hi - 1, hi}) if (0 <= v && v < n[a]) (*h[a])[v] |= TUBE_BIT;
        }
    }
    // SmoothCornerAt: the coordinates of each point (line) and of its inward neighbours are flagged the same way, which puts
    // the point and its neighbours (and a few more sites, harmlessly) into the tubes
    for (auto& a : p->ats) {
        AtGeom ag;
        if (!smooth_at_geometry(a.on_g ? p->g : p->f, a, ag)) continue;
        for (int x = 0; x < p->f->kind; ++x) if (ag.used[x]) { (*h[x])[ag.co[x][0]] |= TUBE_BIT; (*h[x])[ag.co[x][1]] |= TUBE_BIT; }
    }
    auto coords = [&](long long idx, int& i, int& j, int& k) {
        k = (int)(idx/((long long)g.nx*g.ny)); int r = (int)(idx - (long long)k*g.nx*g.ny); j = r/g.nx; i = r - j*g.nx;
    };
    auto tube = [&](int i, int j, int k) { return (hx[i] >> 63) + (hy[j] >> 63) + (hz[k] >> 63) >= 2; };
    if (p->smooth_f || p->smooth_g) {
        SmoothList e, c;
        smooth_lists(p->f, e, c);
        for (SmoothList* L : {&e, &c})
            for (int m = 0; m < L->count; ++m) {
                const SmoothItem& it = L->it[m];
                for (int t = 0; t < it.len; ++t) {
                    long long idx = it.base + (long long)t*it.stride;
                    for (long long s : {idx, idx + it.n0, idx + it.n1, idx + it.n2}) {
                        int i, j, k;
                        coords(s, i, j, k);
                        if (!tube(i, j, k)) return fail(PL_ERR_UNSUPPORTED, "pl_plan_finalize: SmoothCorner on a block this thin is not supported by the fused plan");
                    }
                }
            }
    }
    // list = [sites on closure planes and sites of the last incomplete AVX pack, outside the tubes | tube sites]
    std::vector<int> list, tubes, xlist;
    for (int k = 0; k < g.nz; ++k)
        for (int j = 0; j < g.ny; ++j) {
            const unsigned long long wyz = hy[j] | hz[k];
            const int two = (int)((hy[j] >> 63) + (hz[k] >> 63));
            const long long row = (long long)g.nx*(j + (long long)g.ny*k);
            const bool tailrow = row + g.nx > g.npacked;
            if (!(wyz & ~TUBE_BIT) && two == 0 && !tailrow) {   // only the x planes can put a site of this row on the list
                for (int i = 0; i < g.nx; ++i) if (hx[i] & SLAB_BIT) list.push_back((int)(row + i));
                // the sites the interior kernel takes as ordinary ones once k_xclose has run (neither listed nor in a tube:
                // with one flagged coordinate at most they cannot be)
                for (int i : {0, g.nx - 1}) if (ghost[i]) xlist.push_back((int)(row + i));
                continue;
            }
            for (int i = 0; i < g.nx; ++i) {
                if (two + (int)(hx[i] >> 63) >= 2) tubes.push_back((int)(row + i));
                else if ((wyz & ~TUBE_BIT) || (hx[i] & SLAB_BIT) || row + i >= g.npacked) list.push_back((int)(row + i));
                else if (ghost[i]) xlist.push_back((int)(row + i));
            }
        }
    p->ndirect = (int)list.size();
    list.insert(list.end(), tubes.begin(), tubes.end());
    cudaFree(p->mx); cudaFree(p->my); cudaFree(p->mz); cudaFree(p->list); cudaFree(p->ent); cudaFree(p->xlist); cudaFree(p->xent); cudaFree(p->prog[0]); cudaFree(p->prog[1]);
    cudaFree(p->tube_f); cudaFree(p->tube_g); cudaFree(p->tube_info);
    p->tube_f = p->tube_g = nullptr; p->tube_info = nullptr;
    p->mx = p->my = p->mz = nullptr; p->list = nullptr; p->ent = nullptr; p->xlist = nullptr; p->xent = nullptr; p->nxlist = 0; p->prog[0] = p->prog[1] = nullptr;
    CU(cudaMalloc(&p->mx, g.nx*8)); CU(cudaMalloc(&p->my, g.ny*8)); CU(cudaMalloc(&p->mz, g.nz*8));
    CU(cudaMemcpy(p->mx, hx.data(), g.nx*8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->my, hy.data(), g.ny*8, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(p->mz, hz.data(), g.nz*8, cudaMemcpyHostToDevice));
    p->nlist = (int)list.size();
    if (p->nlist) {
        CU(cudaMalloc(&p->list, list.size()*sizeof(int)));
        CU(cudaMemcpy(p->list, list.data(), list.size()*sizeof(int), cudaMemcpyHostToDevice));
        std::vector<unsigned long long> ent(list.size());
        for (size_t t = 0; t < list.size(); ++t) {
            int i, j, k;
            coords(list[t], i, j, k);
            ent[t] = (hx[i] | hy[j] | hz[k]) & ENTRY_BITS;
        }
        CU(cudaMalloc(&p->ent, ent.size()*sizeof(unsigned long long)));
        CU(cudaMemcpy(p->ent, ent.data(), ent.size()*sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    if (!tubes.empty()) {
        // per tube site: what SmoothCorner makes of it (smooth_lists: the 12 edge lines and 8 corners of the global domain, 4 corners
        // in 2-D), with the neighbours as indices into the tube buffer
        const int ntube = (int)tubes.size();
        std::vector<TubeSite> info(ntube);
        std::unordered_map<int, int> where;
        for (int t = 0; t < ntube; ++t) { where[tubes[t]] = t; memset(&info[t], 0, sizeof(TubeSite)); info[t].idx = tubes[t]; }
        auto tube_of = [&](long long site) { auto it = where.find((int)site); return it == where.end() ? -1 : it->second; };
        const char* thin = "pl_plan_finalize: SmoothCorner on a block this thin is not supported by the fused plan";
        const char* order = "pl_plan_finalize: this loop body cannot be reordered into closures, SmoothCorner, SmoothCornerAt";
        std::vector<long long> smooth_sites[2];      // per lattice: every site SmoothCorner writes or reads
        if (p->smooth_f || p->smooth_g) {
            SmoothList e, c;
            smooth_lists(p->f, e, c);
            std::unordered_map<long long, std::pair<long long, long long>> pair_of;      // edge-line site -> its two face neighbours
            for (int m = 0; m < e.count; ++m)
                for (int t = 0; t < e.it[m].len; ++t) {
                    const long long s0 = e.it[m].base + (long long)t*e.it[m].stride;
                    pair_of[s0] = {s0 + e.it[m].n0, s0 + e.it[m].n1};
                }
            TubeSite proto;
            for (int l = 0; l < 2; ++l) {
                if (!(l == 0 ? p->smooth_f : (p->g && p->smooth_g))) continue;
                for (auto& kv : pair_of) {
                    const int t = tube_of(kv.first), a = tube_of(kv.second.first), b = tube_of(kv.second.second);
                    if (t < 0 || a < 0 || b < 0) return fail(PL_ERR_UNSUPPORTED, thin);
                    info[t].kind[l] = 1; info[t].a[l][0] = a; info[t].a[l][1] = b;
                    smooth_sites[l].insert(smooth_sites[l].end(), {kv.first, kv.second.first, kv.second.second});
                }
                for (int m = 0; m < c.count; ++m) {
                    const SmoothItem& it = c.it[m];
                    const int t = tube_of(it.base);
                    if (t < 0) return fail(PL_ERR_UNSUPPORTED, thin);
                    smooth_sites[l].push_back(it.base);
                    if (it.n2 == 0) {      // 2-D corner: the mean of two sites no SmoothCorner touches
                        const int a = tube_of(it.base + it.n0), b = tube_of(it.base + it.n1);
                        if (a < 0 || b < 0) return fail(PL_ERR_UNSUPPORTED, thin);
                        info[t].kind[l] = 1; info[t].a[l][0] = a; info[t].a[l][1] = b;
                        smooth_sites[l].insert(smooth_sites[l].end(), {it.base + it.n0, it.base + it.n1});
                    } else {
                        info[t].kind[l] = 2;
                        const long long nb3[3] = {it.base + it.n0, it.base + it.n1, it.base + it.n2};
                        for (int q = 0; q < 3; ++q) {
                            auto pe = pair_of.find(nb3[q]);
                            if (pe == pair_of.end()) return fail(PL_ERR_UNSUPPORTED, thin);      // the neighbour of a corner is an edge-line site
                            const int a = tube_of(pe->second.first), b = tube_of(pe->second.second);
                            if (a < 0 || b < 0) return fail(PL_ERR_UNSUPPORTED, thin);
                            info[t].a[l][2*q] = a; info[t].a[l][2*q + 1] = b;
                        }
                    }
                }
            }
            (void)proto;
        }
        // SmoothCornerAt: the point (line) becomes the mean of its neighbours' streamed + closed populations
        struct AtSites { int l, pos; std::vector<long long> touched; };
        std::vector<AtSites> at_sites;
        for (auto& a : p->ats) {
            AtGeom ag;
            const int l = a.on_g;
            if (!smooth_at_geometry(l ? p->g : p->f, a, ag)) continue;
            if (p->smooth_pos[l] >= 0 && a.pos < p->smooth_pos[l]) return fail(PL_ERR_UNSUPPORTED, order);
            AtSites as{l, a.pos, {}};
            for (long long site : ag.sites) {
                const int t = tube_of(site);
                if (t < 0 || info[t].kind[l] != 0) return fail(PL_ERR_UNSUPPORTED, order);      // not a tube site / smoothed twice
                info[t].kind[l] = ag.nnb == 2 ? 1 : 3;
                as.touched.push_back(site);
                for (int q = 0; q < ag.nnb; ++q) {
                    const int b = tube_of(site + ag.nb[q]);
                    if (b < 0) return fail(PL_ERR_UNSUPPORTED, order);
                    info[t].a[l][q] = b;
                    as.touched.push_back(site + ag.nb[q]);
                }
            }
            at_sites.push_back(as);
        }
        // every mean is taken over sites that are not smoothed themselves (k_tubes reads the un-smoothed tube buffer)
        for (int t = 0; t < ntube; ++t)
            for (int l = 0; l < 2; ++l) {
                const int kd = info[t].kind[l];
                if (kd != 1 && kd != 3) continue;
                for (int q = 0; q < (kd == 1 ? 2 : 3); ++q) if (info[info[t].a[l][q]].kind[l] != 0) return fail(PL_ERR_UNSUPPORTED, order);
            }
        // a closure called AFTER a SmoothCorner / SmoothCornerAt of its lattice must not touch a site that one wrote or read:
        // the kernels run all closures first
        for (auto& b : p->bcs) {
            if (b.bc->empty) continue;
            const int l = b.on_g;
            auto hits = [&](const std::vector<long long>& sites) {
                const Plane& pl = b.bc->pl;
                const int a1 = pl.axis == 0 ? 1 : 0, a2 = pl.axis == 2 ? 1 : 2;
                const int loc = b.bc->coord - off[pl.axis];
                for (long long site : sites) {
                    int c3[3];
                    coords(site, c3[0], c3[1], c3[2]);
                    if (c3[pl.axis] != loc) continue;
                    if (b.bc->hmask[(size_t)c3[a1] + (size_t)pl.n1*c3[a2]]) return true;
                }
                return false;
            };
            if (p->smooth_pos[l] >= 0 && b.pos > p->smooth_pos[l] && hits(smooth_sites[l])) return fail(PL_ERR_UNSUPPORTED, order);
            for (auto& as : at_sites) if (as.l == l && as.pos < b.pos && hits(as.touched)) return fail(PL_ERR_UNSUPPORTED, order);
        }
        const size_t nb = (size_t)p->f->nc*ntube*sizeof(double);
        CU(cudaMalloc(&p->tube_f, nb));
        if (p->g) CU(cudaMalloc(&p->tube_g, nb));
        CU(cudaMalloc(&p->tube_info, (size_t)ntube*sizeof(TubeSite)));
        CU(cudaMemcpy(p->tube_info, info.data(), (size_t)ntube*sizeof(TubeSite), cudaMemcpyHostToDevice));
    }
    p->nxlist = (int)xlist.size();
    for (int l = 0; l < 2; ++l) { cudaFree(p->xout[l][0]); cudaFree(p->xout[l][1]); cudaFree(p->xres[l]); p->xout[l][0] = p->xout[l][1] = p->xres[l] = nullptr; p->xver[l] = ~0ull; }
    p->xon[0] = ghost[0]; p->xon[1] = ghost[g.nx - 1];
    if (p->nxlist) {
        const size_t wb = (size_t)2*p->f->nc*g.ny*g.nz*sizeof(double);
        for (int l = 0; l < (p->g ? 2 : 1); ++l) {
            CU(cudaMalloc(&p->xout[l][0], wb)); CU(cudaMalloc(&p->xout[l][1], wb)); CU(cudaMalloc(&p->xres[l], wb));
            CU(cudaMemsetAsync(p->xout[l][0], 0, wb, g_stream)); CU(cudaMemsetAsync(p->xout[l][1], 0, wb, g_stream)); CU(cudaMemsetAsync(p->xres[l], 0, wb, g_stream));
        }
        std::vector<unsigned long long> xent(xlist.size());
        for (size_t t = 0; t < xlist.size(); ++t) {
            int i, j, k;
            coords(xlist[t], i, j, k);
            xent[t] = hx[i] & ENTRY_BITS;
        }
        CU(cudaMalloc(&p->xlist, xlist.size()*sizeof(int)));
        CU(cudaMemcpy(p->xlist, xlist.data(), xlist.size()*sizeof(int), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&p->xent, xent.size()*sizeof(unsigned long long)));
        CU(cudaMemcpy(p->xent, xent.data(), xent.size()*sizeof(unsigned long long), cudaMemcpyHostToDevice));
    }
    p->nprog = (int)prog[0].size();
    for (int par = 0; par < 2; ++par) {
        CU(cudaMalloc(&p->prog[par], std::max<size_t>(1, prog[par].size())*sizeof(ClosureArgs)));
        if (p->nprog) CU(cudaMemcpy(p->prog[par], prog[par].data(), prog[par].size()*sizeof(ClosureArgs), cudaMemcpyHostToDevice));
    }
    p->finalized = true;
    return PL_OK;
}
int pl_plan_rebind(pl_plan* p, int parity, const pl_collide_args* collide, const pl_bc_aux* aux, int naux) {
    if (!p || !p->finalized) return fail(PL_ERR_ARG, "pl_plan_rebind: plan not finalized");
    if (parity != 0 && parity != 1) return fail(PL_ERR_ARG, "pl_plan_rebind: parity 0 / 1");
    drop_graphs(p);               // captured steps hold the old addresses
    p->graph_cooldown = 8;
    if (collide) {
        if (collide->model != p->args[parity].model) return fail(PL_ERR_ARG, "pl_plan_rebind: the collide model of a plan cannot change");
        p->args[parity] = *collide;
    }
    if (!aux || naux <= 0) return PL_OK;
    int k = 0;
    for (auto& b : p->bcs) {
        if (!b.has_aux) continue;
        if (k >= naux) break;
        b.aux[parity] = aux[k++];
    }
    if (p->nprog == 0) return PL_OK;
    // re-build the closure program of this argument set and queue its copy behind everything already queued
    if (!p->stage) {
        CU(cudaMallocHost(&p->stage, (size_t)pl_plan::NSTAGE*MAX_PROGRAM*sizeof(ClosureArgs)));
        for (auto& e : p->stage_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const int slot = p->stage_next;
    p->stage_next = (slot + 1)%pl_plan::NSTAGE;
    CU(cudaEventSynchronize(p->stage_ev[slot]));      // returns at once unless the ring wrapped onto a copy still queued
    ClosureArgs* h = p->stage + (size_t)slot*MAX_PROGRAM;
    const Geom& g = p->f->g;
    const int off[3] = {g.offx, g.offy, g.offz};
    int e = 0;
    for (auto& b : p->bcs) {
        if (b.bc->empty) continue;
        pl_lattice* l = b.on_g ? p->g : p->f;
        int r = make_closure_args(l, b.bc->type == PL_BC_AAD_ISET_RHO ? p->g : nullptr, b.bc, b.has_aux ? &b.aux[parity] : nullptr, h[e]);
        if (r) return r;
        h[e].on_g = b.on_g; h[e].loc = b.bc->coord - off[b.bc->axis];
        ++e;
    }
    CU(cudaMemcpyAsync(p->prog[parity], h, (size_t)p->nprog*sizeof(ClosureArgs), cudaMemcpyHostToDevice, g_stream));
    CU(cudaEventRecord(p->stage_ev[slot], g_stream));
    return PL_OK;
}
int pl_plan_parity(const pl_plan* p) { return p ? p->parity : 0; }
int pl_plan_set_parity(pl_plan* p, int parity) { if (!p) return fail(PL_ERR_ARG, "null plan"); p->parity = parity ? 1 : 0; return PL_OK; }
int pl_plan_profile(pl_plan* p, int enable) { if (!p) return fail(PL_ERR_ARG, "null plan"); p->profile = enable != 0; return PL_OK; }
int pl_plan_profile_read2(pl_plan* p, double* ms2, int* launches2, long long* sites2) {
    if (!p) return fail(PL_ERR_ARG, "null plan");
    CU(cudaStreamSynchronize(g_stream));
    double ms[2] = {0.0, 0.0}; int n[2] = {0, 0}; long long sites[2] = {0, 0};
    for (auto& e : p->events) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, e.a, e.b));
        ms[e.cls] += t; ++n[e.cls]; sites[e.cls] += e.sites;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    for (int c = 0; c < 2; ++c) {
        if (ms2) ms2[c] = ms[c];
        if (launches2) launches2[c] = n[c];
        if (sites2) sites2[c] = sites[c];
    }
    p->events.clear();
    return PL_OK;
}
int pl_plan_profile_read(pl_plan* p, double* total_ms, int* launches, long long* total_sites) {
    double ms[2]; int n[2]; long long sites[2];
    int r = pl_plan_profile_read2(p, ms, n, sites);
    if (r) return r;
    if (total_ms) *total_ms = ms[0] + ms[1];
    if (launches) *launches = n[0] + n[1];
    if (total_sites) *total_sites = sites[0] + sites[1];
    return PL_OK;
}

int pl_plan_advance(pl_plan* p, int ncollides, int end_streamed) { return pl_plan_advance_observed(p, ncollides, end_streamed, -1); }
int pl_plan_advance_observed(pl_plan* p, int ncollides, int end_streamed, int save_last) {
    if (!p || !p->finalized) return fail(PL_ERR_ARG, "pl_plan_advance: plan not finalized");
    if (ncollides < 0) return fail(PL_ERR_ARG, "pl_plan_advance: negative count");
    int r;
    int done = 0;
    // The argument set of step t serves its collide and the closures that follow it (they read the velocities that
    // collide saved, production/heatsink3D.cpp:164-173); the driver swaps sets after the closures (:178-183).
    if (ncollides > 0 && p->f->streamed) {
        if ((r = plan_collide_full(p, p->parity))) return r;
        done = 1;
    }
    if (done < ncollides && ncollides - done >= 3) {
        // small lattices: the first pass on its own (it may refill the wall buffers), the rest in one cooperative launch
        if ((r = plan_fused(p, p->parity, p->parity ^ 1, save_last < 0 || ncollides - done <= save_last))) return r;
        p->parity ^= 1;
        ++done;
        bool coop = false;
        if ((r = plan_steps(p, ncollides - done, save_last, 0, &coop))) return r;
        if (coop) done = ncollides;
    }
    while (done < ncollides) {
        // state: just collided with set `parity`; fuse S(parity) with C(parity^1)
        // only the last `save_last` collides of the call leave their macroscopic fields / snapshot behind at every site
        if ((r = plan_fused(p, p->parity, p->parity ^ 1, save_last < 0 || ncollides - done <= save_last))) return r;
        p->parity ^= 1;
        ++done;
    }
    if (end_streamed && !p->f->streamed) {
        if ((r = plan_stream_bc_smooth_full(p, p->parity))) return r;
        p->parity ^= 1;
    }
    return PL_OK;
}

// ---- communicator ---------------------------------------------------------------------------------
int pl_comm_unique_id(char* out128) {
    if (!out128) return fail(PL_ERR_ARG, "pl_comm_unique_id: null");
    int r = load_nccl();
    if (r) return r;
    NcclId id;
    NC(g_nccl.GetUniqueId(&id));
    memcpy(out128, id.internal, 128);
    return PL_OK;
}
int pl_comm_init(const char* id128, int rank, int nranks) {
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(PL_ERR_ARG, "pl_comm_init: bad arguments");
    if (g_comm.mode != COMM_NONE) return fail(PL_ERR_ARG, "pl_comm_init: a communicator already exists");
    int r = load_nccl();
    if (r || (r = comm_streams())) return r;
    NcclId id;
    memcpy(id.internal, id128, 128);
    NC(g_nccl.CommInitRank(&g_comm.nccl, nranks, id, rank));
    g_comm.mode = COMM_NCCL; g_comm.rank = rank; g_comm.nranks = nranks;
    return PL_OK;
}
int pl_comm_init_loopback(int nranks) {
    if (nranks < 1) return fail(PL_ERR_ARG, "pl_comm_init_loopback: bad world size");
    if (g_comm.mode != COMM_NONE) return fail(PL_ERR_ARG, "pl_comm_init_loopback: a communicator already exists");
    g_comm.mode = COMM_LOOPBACK; g_comm.rank = 0; g_comm.nranks = nranks;
    g_comm.loop.assign(nranks, std::vector<pl_lattice*>());
    return PL_OK;
}
int pl_comm_destroy(void) {
    if (g_comm.mode == COMM_NCCL) {
        cudaStreamSynchronize(g_comm.stream);
        cudaStreamSynchronize(g_stream);
        if (g_comm.nccl) g_nccl.CommDestroy(g_comm.nccl);
        g_comm.nccl = nullptr;
    }
    g_comm.loop.clear();
    g_comm.mode = COMM_NONE; g_comm.rank = 0; g_comm.nranks = 1;
    return PL_OK;
}
int pl_comm_info(int* mode, int* rank, int* nranks) {
    if (mode) *mode = g_comm.mode;
    if (rank) *rank = g_comm.rank;
    if (nranks) *nranks = g_comm.nranks;
    return PL_OK;
}
int pl_comm_allreduce_v(void* inout_host, size_t n, int dtype, int op) {
    InCall in_call_;
    if (!inout_host || (dtype != 0 && dtype != 1) || op < 0 || op > 2) return fail(PL_ERR_ARG, "pl_comm_allreduce_v: dtype 0 (f64) / 1 (i32), op 0 (sum) / 1 (max) / 2 (min)");
    if (g_comm.mode != COMM_NCCL || n == 0) return PL_OK;   // a world of one
    const size_t bytes = n*(dtype == 0 ? sizeof(double) : sizeof(int));
    void* d = g_scratch.get(bytes);
    if (!d) return fail(PL_ERR_CUDA, "pl_comm_allreduce_v: scratch allocation failed");
    CU(cudaMemcpyAsync(d, inout_host, bytes, cudaMemcpyHostToDevice, g_stream));
    NC(g_nccl.AllReduce(d, d, n, dtype == 0 ? NCCL_F64 : 2 /* ncclInt32 */, op == 0 ? NCCL_SUM : (op == 1 ? NCCL_MAX : 3 /* ncclMin */), g_comm.nccl, g_stream));
    ++g_launches;
    CU(cudaMemcpyAsync(inout_host, d, bytes, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
int pl_comm_p2p(const pl_p2p_op* ops, int n) {
    InCall in_call_;
    if (n < 0 || (n > 0 && !ops)) return fail(PL_ERR_ARG, "pl_comm_p2p: bad arguments");
    if (n == 0) return PL_OK;
    if (g_comm.mode != COMM_NCCL) return fail(PL_ERR_ARG, "pl_comm_p2p: no NCCL communicator");
    size_t total = 0;
    std::vector<size_t> off(n);
    for (int k = 0; k < n; ++k) { off[k] = total; total += (ops[k].bytes + 255)/256*256; }
    char* d = (char*)g_scratch.get(std::max<size_t>(total, 256));
    if (!d) return fail(PL_ERR_CUDA, "pl_comm_p2p: scratch allocation failed");
    for (int k = 0; k < n; ++k)
        if (ops[k].is_send && ops[k].bytes) CU(cudaMemcpyAsync(d + off[k], ops[k].host, ops[k].bytes, cudaMemcpyHostToDevice, g_stream));
    NC(g_nccl.GroupStart());
    for (int k = 0; k < n; ++k) {
        if (ops[k].peer < 0 || ops[k].peer >= g_comm.nranks) { g_nccl.GroupEnd(); return fail(PL_ERR_ARG, "pl_comm_p2p: peer out of range"); }
        if (ops[k].is_send) NC(g_nccl.Send(d + off[k], ops[k].bytes, 0 /* ncclInt8 */, ops[k].peer, g_comm.nccl, g_stream));
        else NC(g_nccl.Recv(d + off[k], ops[k].bytes, 0, ops[k].peer, g_comm.nccl, g_stream));
    }
    NC(g_nccl.GroupEnd());
    ++g_launches;
    for (int k = 0; k < n; ++k)
        if (!ops[k].is_send && ops[k].bytes) CU(cudaMemcpyAsync(ops[k].host, d + off[k], ops[k].bytes, cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
int pl_comm_allreduce(double* inout_host, int n, int op) {
    InCall in_call_;
    if (!inout_host || n < 1 || n > 4 || (op != 0 && op != 1)) return fail(PL_ERR_ARG, "pl_comm_allreduce: 1..4 values, op 0 (sum) / 1 (max)");
    if (g_comm.mode != COMM_NCCL) return PL_OK;   // a world of one
    CU(cudaMemcpyAsync(g_comm.red, inout_host, n*sizeof(double), cudaMemcpyHostToDevice, g_stream));
    NC(g_nccl.AllReduce(g_comm.red, g_comm.red, (size_t)n, NCCL_F64, op == 0 ? NCCL_SUM : NCCL_MAX, g_comm.nccl, g_stream));
    ++g_launches;
    CU(cudaMemcpyAsync(inout_host, g_comm.red, n*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
int pl_halo_describe(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz, int inverse, int* out, int* count) {
    if ((kind != PL_D2Q9 && kind != PL_D3Q15) || !out || !count || lx <= 0 || ly <= 0 || lz <= 0 || mx <= 0 || my <= 0 || mz <= 0 || peid < 0)
        return fail(PL_ERR_ARG, "pl_halo_describe: bad arguments");
    if (kind == PL_D2Q9) { lz = 1; mz = 1; }
    if (peid >= mx*my*mz) return fail(PL_ERR_ARG, "pl_halo_describe: PEid outside the PE grid");
    const int pe[3] = {peid%mx, kind == PL_D2Q9 ? peid/mx : (peid/mx)%my, kind == PL_D2Q9 ? 0 : peid/(mx*my)};
    const int m[3] = {mx, my, mz};
    const int n[3] = {(lx + pe[0])/mx, (ly + pe[1])/my, kind == PL_D2Q9 ? 1 : (lz + pe[2])/mz};
    HaloMsgDesc d[26];
    const int cnt = kind == PL_D2Q9 ? halo_describe<2>(n, m, pe, inverse, d) : halo_describe<3>(n, m, pe, inverse, d);
    for (int k = 0; k < cnt; ++k) {
        int* o = out + 16*k;
        o[0] = d[k].code; o[1] = d[k].peer; o[2] = (int)d[k].rsize; o[3] = d[k].npop;
        for (int q = 0; q < 5; ++q) o[4 + q] = q < d[k].npop ? d[k].pop[q] : -1;
        o[9] = (int)d[k].base; o[10] = (int)d[k].s1; o[11] = (int)d[k].s2; o[12] = d[k].n1; o[13] = d[k].n2;
        o[14] = halo_opposite(d[k].code); o[15] = 0;
    }
    *count = cnt;
    return PL_OK;
}

// ---- reductions ---------------------------------------------------------------------------------
int pl_residual(const double* ux, const double* uy, const double* uz, const double* uxp, const double* uyp, const double* uzp, size_t n, double* out) {
    if (!ux || !uxp || !out) return fail(PL_ERR_ARG, "pl_residual: null");
    const int nb = 1024;
    double* scratch = (double*)g_scratch.get(2*(nb + 1)*sizeof(double));
    if (!scratch) return fail(PL_ERR_CUDA, "pl_residual: scratch allocation failed");
    LAUNCH(k_residual_partial, nb, 256, ux, uy, uz, uxp, uyp, uzp, (long long)n, scratch);
    LAUNCH(k_sum_final, 1, 256, scratch, nb, 2, scratch + 2*nb);
    double h[2];
    // MPI build of the reference: MPI_Allreduce(SUM) of the two partial sums (residual.h:16, 31, 46)
    if (g_comm.mode == COMM_NCCL) { NC(g_nccl.AllReduce(scratch + 2*nb, scratch + 2*nb, 2, NCCL_F64, NCCL_SUM, g_comm.nccl, g_stream)); ++g_launches; }
    CU(cudaMemcpyAsync(h, scratch + 2*nb, 2*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    *out = sqrt(h[0]/h[1]);
    return PL_OK;
}
int pl_reduce_sum(const double* v, size_t n, double* out) {
    if (!v || !out) return fail(PL_ERR_ARG, "pl_reduce_sum: null");
    const int nb = 1024;
    double* scratch = (double*)g_scratch.get((nb + 1)*sizeof(double));
    if (!scratch) return fail(PL_ERR_CUDA, "pl_reduce_sum: scratch allocation failed");
    LAUNCH(k_sum_partial, nb, 256, v, (long long)n, scratch);
    LAUNCH(k_sum_final, 1, 256, scratch, nb, 1, scratch + nb);
    CU(cudaMemcpyAsync(out, scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
static int absmax_impl(const double* v, size_t n, double* out, bool global);
int pl_reduce_absmax(const double* v, size_t n, double* out) { return absmax_impl(v, n, out, false); }
static int absmax_impl(const double* v, size_t n, double* out, bool global) {
    if (!v || !out) return fail(PL_ERR_ARG, "pl_reduce_absmax: null");
    const int nb = 1024;
    double* scratch = (double*)g_scratch.get((nb + 1)*sizeof(double));
    if (!scratch) return fail(PL_ERR_CUDA, "pl_reduce_absmax: scratch allocation failed");
    LAUNCH(k_absmax_partial, nb, 256, v, (long long)n, scratch);
    LAUNCH(k_absmax_final, 1, 256, scratch, nb, scratch + nb);
    // normalize.h:17 — MPI_Allreduce(MAX) in the reference's MPI build
    if (global && g_comm.mode == COMM_NCCL) { NC(g_nccl.AllReduce(scratch + nb, scratch + nb, 1, NCCL_F64, NCCL_MAX, g_comm.nccl, g_stream)); ++g_launches; }
    CU(cudaMemcpyAsync(out, scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
int pl_normalize(double* v, size_t n) {
    double m = 0.0;
    int r = absmax_impl(v, n, &m, true);
    if (r) return r;
    LAUNCH(k_divide, blocks_for((long long)n, 256), 256, v, m, (long long)n);
    return PL_OK;
}

// ---- filters ----------------------------------------------------------------------------------------
struct pl_filter {
    FilterGeom F;
    bool global = false;           // decomposed block: fields are assembled over all ranks first
    double* wtab = nullptr;        // [npat][K] weight patterns
    int* pid = nullptr;            // pattern of each site
    int npat = 0;
    double* tmp = nullptr;         // first pass of the sensitivity filter (block)
    // decomposed lattice: the block with an nR-wide ghost layer along the decomposed axes, filled from the neighbouring ranks
    // (heavisidefilter.h:291-400 exchanges the same layer with 26 MPI messages; here axis by axis, 2 messages each, the slabs of
    // the later axes carrying the ghosts of the earlier ones, which delivers the edge and corner regions as well)
    double* gh = nullptr;
    double *sbuf[2] = {nullptr, nullptr}, *rbuf[2] = {nullptr, nullptr};
    int dec[3] = {0, 0, 0}, pe[3] = {0, 0, 0}, m[3] = {1, 1, 1};
};
static pl_filter* filter_from_patterns(pl_lattice* l, int nR, const double* patterns, int npat, const int* pattern_of_site) {
    pl_filter* f = new pl_filter();
    FilterGeom& F = f->F;
    F.nx = l->g.nx; F.ny = l->g.ny; F.nz = l->g.nz; F.nR = nR; F.nxyz = l->g.nxyz;
    F.gx = l->g.lx; F.gy = l->g.ly; F.gz = l->g.lz; F.ox = l->g.offx; F.oy = l->g.offy; F.oz = l->g.offz;
    f->global = l->halo.on;
    f->npat = npat;
    f->dec[0] = l->mx > 1; f->dec[1] = l->my > 1; f->dec[2] = l->kind == PL_D3Q15 && l->mz > 1;
    f->pe[0] = l->pex; f->pe[1] = l->pey; f->pe[2] = l->pez; f->m[0] = l->mx; f->m[1] = l->my; f->m[2] = l->mz;
    F.ax = f->dec[0] ? nR : 0; F.ay = f->dec[1] ? nR : 0; F.az = f->dec[2] ? nR : 0;
    F.fx = F.nx + 2*F.ax; F.fy = F.ny + 2*F.ay; F.fz = F.nz + 2*F.az;
    const size_t side = 2*(size_t)nR + 1, K = side*side*side, n = (size_t)l->g.nxyz, gn = (size_t)F.fx*F.fy*F.fz;
    bool ok = cudaMalloc(&f->wtab, std::max<size_t>(1, K*npat)*sizeof(double)) == cudaSuccess && cudaMalloc(&f->pid, n*sizeof(int)) == cudaSuccess &&
              cudaMalloc(&f->tmp, n*sizeof(double)) == cudaSuccess;
    if (ok && f->global) {
        const size_t slab = (size_t)std::max(1, nR)*std::max({(size_t)F.fy*F.fz, (size_t)F.fx*F.fz, (size_t)F.fx*F.fy});
        ok = cudaMalloc(&f->gh, gn*sizeof(double)) == cudaSuccess;
        for (int b = 0; ok && b < 2; ++b) ok = cudaMalloc(&f->sbuf[b], slab*sizeof(double)) == cudaSuccess && cudaMalloc(&f->rbuf[b], slab*sizeof(double)) == cudaSuccess;
    }
    ok = ok && cudaMemcpy(f->wtab, patterns, K*npat*sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(f->pid, pattern_of_site, n*sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        fail(PL_ERR_CUDA, std::string("pl_filter_create: ") + cudaGetErrorString(cudaGetLastError()));
        cudaFree(f->wtab); cudaFree(f->pid); cudaFree(f->tmp); cudaFree(f->gh);
        for (int b = 0; b < 2; ++b) { cudaFree(f->sbuf[b]); cudaFree(f->rbuf[b]); }
        delete f;
        return nullptr;
    }
    return f;
}
static bool filter_args_ok(pl_lattice* l, int nR, const void* a, const void* b) {
    if (!l || nR < 0 || nR > 8 || !a || !b) { fail(PL_ERR_ARG, "pl_filter_create: bad arguments"); return false; }
    if (l->halo.on && g_comm.mode != COMM_NCCL) {
        fail(PL_ERR_UNSUPPORTED, "pl_filter_create: filters on a block-decomposed lattice need the NCCL communicator (pl_comm_init)");
        return false;
    }
    return true;
}
pl_filter* pl_filter_create_patterns(pl_lattice* l, int nR, const double* patterns, int npatterns, const int* pattern_of_site) {
    InCall in_call_;
    if (!filter_args_ok(l, nR, patterns, pattern_of_site) || npatterns < 1) { if (npatterns < 1) fail(PL_ERR_ARG, "pl_filter_create_patterns: no pattern"); return nullptr; }
    for (long long i = 0; i < l->g.nxyz; ++i)
        if (pattern_of_site[i] < 0 || pattern_of_site[i] >= npatterns) { fail(PL_ERR_ARG, "pl_filter_create_patterns: pattern index out of range"); return nullptr; }
    return filter_from_patterns(l, nR, patterns, npatterns, pattern_of_site);
}
// dense per-site table (weights_host[o*nxyz + idx]): sites with the same K weights share one pattern
pl_filter* pl_filter_create(pl_lattice* l, int nR, const double* weights_host) {
    InCall in_call_;
    if (!filter_args_ok(l, nR, weights_host, weights_host)) return nullptr;
    const size_t side = 2*(size_t)nR + 1, K = side*side*side, n = (size_t)l->g.nxyz;
    std::unordered_map<std::string, int> seen;
    std::vector<double> patterns, row(K);
    std::vector<int> pid(n);
    for (size_t i = 0; i < n; ++i) {
        for (size_t o = 0; o < K; ++o) row[o] = weights_host[o*n + i];
        auto r = seen.emplace(std::string(reinterpret_cast<const char*>(row.data()), K*sizeof(double)), (int)seen.size());
        if (r.second) patterns.insert(patterns.end(), row.begin(), row.end());
        pid[i] = r.first->second;
    }
    return filter_from_patterns(l, nR, patterns.data(), (int)seen.size(), pid.data());
}
int pl_filter_patterns(const pl_filter* f) { return f ? f->npat : 0; }
int pl_filter_destroy(pl_filter* f) {
    if (!f) return PL_OK;
    cudaStreamSynchronize(g_stream);
    cudaFree(f->wtab); cudaFree(f->pid); cudaFree(f->tmp); cudaFree(f->gh);
    for (int b = 0; b < 2; ++b) { cudaFree(f->sbuf[b]); cudaFree(f->rbuf[b]); }
    delete f;
    return PL_OK;
}
// the field the filter kernel reads: the block itself, or (decomposed) the block with its nR-wide ghost layer exchanged with the
// neighbouring ranks — O(block) memory and traffic per rank.  The domain is NOT periodic for the filters (neighbours outside the
// global domain do not count, heavisidefilter.h:470-556): ranks on a domain face have no partner there.
static int filter_field(pl_filter* f, const double* v, const double** out) {
    if (!f->global) { *out = v; return PL_OK; }
    const FilterGeom& F = f->F;
    const int nR = F.nR;
    CU(cudaMemsetAsync(f->gh, 0, (size_t)F.fx*F.fy*F.fz*sizeof(double), g_stream));
    LAUNCH(k_box_copy, blocks_for(F.nxyz, 256), 256, v, F.nx, F.ny, 0, 0, 0, f->gh, F.fx, F.fy, F.ax, F.ay, F.az, F.nx, F.ny, F.nz);
    if (nR == 0) { *out = f->gh; return PL_OK; }
    const int n[3] = {F.nx, F.ny, F.nz}, a[3] = {F.ax, F.ay, F.az}, fd[3] = {F.fx, F.fy, F.fz};
    const int stride[3] = {1, f->m[0], f->m[0]*f->m[1]};
    for (int ax = 0; ax < 3; ++ax) {
        if (!f->dec[ax]) continue;
        // the slab spans the full (ghosted) extent of the axes already exchanged, the interior of the later ones
        int ext[3], org[3];
        for (int d = 0; d < 3; ++d) { ext[d] = d < ax ? fd[d] : n[d]; org[d] = d < ax ? 0 : a[d]; }
        ext[ax] = nR;
        const long long cnt = (long long)ext[0]*ext[1]*ext[2];
        const bool lo = f->pe[ax] > 0, hi = f->pe[ax] < f->m[ax] - 1;
        const int rank = g_comm.rank;
        // pack: towards the low neighbour the first nR interior layers, towards the high neighbour the last nR
        for (int side = 0; side < 2; ++side) {
            if (!(side ? hi : lo)) continue;
            int so[3] = {org[0], org[1], org[2]};
            so[ax] = side ? a[ax] + n[ax] - nR : a[ax];
            LAUNCH(k_box_copy, blocks_for(cnt, 256), 256, f->gh, fd[0], fd[1], so[0], so[1], so[2], f->sbuf[side], ext[0], ext[1], 0, 0, 0, ext[0], ext[1], ext[2]);
        }
        NC(g_nccl.GroupStart());
        if (lo) { NC(g_nccl.Send(f->sbuf[0], (size_t)cnt, NCCL_F64, rank - stride[ax], g_comm.nccl, g_stream)); NC(g_nccl.Recv(f->rbuf[0], (size_t)cnt, NCCL_F64, rank - stride[ax], g_comm.nccl, g_stream)); }
        if (hi) { NC(g_nccl.Send(f->sbuf[1], (size_t)cnt, NCCL_F64, rank + stride[ax], g_comm.nccl, g_stream)); NC(g_nccl.Recv(f->rbuf[1], (size_t)cnt, NCCL_F64, rank + stride[ax], g_comm.nccl, g_stream)); }
        NC(g_nccl.GroupEnd());
        ++g_launches;
        // unpack into the ghost layers
        for (int side = 0; side < 2; ++side) {
            if (!(side ? hi : lo)) continue;
            int dorg[3] = {org[0], org[1], org[2]};
            dorg[ax] = side ? a[ax] + n[ax] : 0;
            LAUNCH(k_box_copy, blocks_for(cnt, 256), 256, f->rbuf[side], ext[0], ext[1], 0, 0, 0, f->gh, fd[0], fd[1], dorg[0], dorg[1], dorg[2], ext[0], ext[1], ext[2]);
        }
    }
    *out = f->gh;
    return PL_OK;
}
int pl_filter_apply(pl_filter* f, int mode, double beta, const double* v, const double* dfdrho, double* out) {
    if (!f || !v || !out) return fail(PL_ERR_ARG, "pl_filter_apply: null");
    if (mode < 0 || mode > 2) return fail(PL_ERR_ARG, "pl_filter_apply: mode 0 (density), 1 (Heaviside variable), 2 (Heaviside sensitivity)");
    if (mode == 2 && !dfdrho) return fail(PL_ERR_ARG, "pl_filter_apply: the sensitivity filter needs dfdrho");
    const unsigned nb = blocks_for(f->F.nxyz, 256);
    const double* field;
    int r = filter_field(f, v, &field);
    if (r) return r;
    if (mode < 2) LAUNCH(k_filter, nb, 256, f->F, f->wtab, f->pid, field, nullptr, beta, mode, out);
    else {
        LAUNCH(k_filter, nb, 256, f->F, f->wtab, f->pid, field, dfdrho, beta, 2, f->tmp);
        if ((r = filter_field(f, f->tmp, &field))) return r;
        LAUNCH(k_filter, nb, 256, f->F, f->wtab, f->pid, field, nullptr, beta, 3, out);
    }
    return PL_OK;
}
int pl_design_map(const double* ss, size_t n, double diff_fluid, double diff_solid, double qg, double alpha0, double qf, double* diffusivity, double* alpha,
                  double* dkds, double* dads) {
    if (!ss || !diffusivity || !alpha || !dkds || !dads) return fail(PL_ERR_ARG, "pl_design_map: null");
    if (n == 0) return PL_OK;
    LAUNCH(k_design_map, blocks_for((long long)n, 256), 256, ss, (long long)n, diff_fluid, diff_solid, qg, alpha0, qf, diffusivity, alpha, dkds, dads);
    return PL_OK;
}
int pl_reduce_box_sum(const pl_lattice* l, const double* v, int i0, int i1, int j0, int j1, int k0, int k1, double* out) {
    if (!l || !v || !out) return fail(PL_ERR_ARG, "pl_reduce_box_sum: null");
    const Geom& g = l->g;
    // global coordinates, clipped to this rank's block (the drivers' `(i + offsetx) < L` tests, heatsink3D.cpp:229-235)
    i0 = std::max(i0 - g.offx, 0); i1 = std::min(i1 - g.offx, g.nx); j0 = std::max(j0 - g.offy, 0); j1 = std::min(j1 - g.offy, g.ny);
    k0 = std::max(k0 - g.offz, 0); k1 = std::min(k1 - g.offz, g.nz);
    const int nb = 256;
    double* scratch = (double*)g_scratch.get((nb + 1)*sizeof(double));
    if (!scratch) return fail(PL_ERR_CUDA, "pl_reduce_box_sum: scratch allocation failed");
    if (i1 <= i0 || j1 <= j0 || k1 <= k0) { *out = 0.0; }
    else {
        LAUNCH(k_box_sum_partial, nb, 256, v, g.nx, g.ny, i0, i1, j0, j1, k0, k1, scratch);
        LAUNCH(k_sum_final, 1, 256, scratch, nb, 1, scratch + nb);
    }
    if (i1 <= i0 || j1 <= j0 || k1 <= k0) CU(cudaMemsetAsync(scratch + nb, 0, sizeof(double), g_stream));
    // the drivers' MPI_Allreduce(SUM) of the partial objective (heatsink3D.cpp:236)
    if (g_comm.mode == COMM_NCCL) { NC(g_nccl.AllReduce(scratch + nb, scratch + nb, 1, NCCL_F64, NCCL_SUM, g_comm.nccl, g_stream)); ++g_launches; }
    CU(cudaMemcpyAsync(out, scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}
// every rank's block of a per-site field into the field of the GLOBAL domain, on every rank (what the VTK writers of the
// reference gather block by block with MPI_Isend/Irecv, vtkxmlexport.h:172-214): device-side scatter + one all-reduce
int pl_comm_gather_field(const pl_lattice* l, const double* v, double* out_host_global) {
    InCall in_call_;
    if (!l || !v || !out_host_global) return fail(PL_ERR_ARG, "pl_comm_gather_field: null");
    const Geom& g = l->g;
    const size_t gn = (size_t)g.lx*g.ly*g.lz;
    if (!l->halo.on) { CU(cudaMemcpyAsync(out_host_global, v, gn*sizeof(double), cudaMemcpyDeviceToHost, g_stream)); CU(cudaStreamSynchronize(g_stream)); return PL_OK; }
    if (g_comm.mode != COMM_NCCL) return fail(PL_ERR_UNSUPPORTED, "pl_comm_gather_field: a block-decomposed lattice needs the NCCL communicator");
    double* gv = (double*)g_scratch.get(gn*sizeof(double));
    if (!gv) return fail(PL_ERR_CUDA, "pl_comm_gather_field: scratch allocation failed");
    FilterGeom F;
    F.nx = g.nx; F.ny = g.ny; F.nz = g.nz; F.nR = 0; F.nxyz = g.nxyz; F.gx = g.lx; F.gy = g.ly; F.gz = g.lz; F.ox = g.offx; F.oy = g.offy; F.oz = g.offz;
    CU(cudaMemsetAsync(gv, 0, gn*sizeof(double), g_stream));
    LAUNCH(k_filter_scatter, blocks_for(g.nxyz, 256), 256, F, v, gv);
    NC(g_nccl.AllReduce(gv, gv, gn, NCCL_F64, NCCL_SUM, g_comm.nccl, g_stream));
    ++g_launches;
    CU(cudaMemcpyAsync(out_host_global, gv, gn*sizeof(double), cudaMemcpyDeviceToHost, g_stream));
    CU(cudaStreamSynchronize(g_stream));
    return PL_OK;
}

int pl_sensitivity(pl_lattice* l, const pl_sens_args* a) {
    if (!l || !a) return fail(PL_ERR_ARG, "pl_sensitivity: null");
    const bool d3 = l->kind == PL_D3Q15;
    if (a->kind < 1 || a->kind > 3) return fail(PL_ERR_ARG, "pl_sensitivity: unknown kind");
    // test/nssens3D.cpp:105 hands a D3Q15 lattice to the 2-D overload of ANS::SensitivityBrinkman (no uz, imz): the reference then
    // evaluates the two-component expression at every site (adjointnavierstokes_avx.h:262-278) — so does this
    const bool planar = d3 && a->kind == PL_SENS_ANS_BRINKMAN && !a->uz && !a->imz;
    if (!a->dfds || !a->ux || !a->uy || !a->imx || !a->imy || !a->dads || (d3 && !planar && (!a->uz || !a->imz)))
        return fail(PL_ERR_ARG, "pl_sensitivity: missing dfds / u / im / dads arrays");
    if (a->kind == PL_SENS_AAD_HEATEX && (!a->tem || !a->item || !a->dbds)) return fail(PL_ERR_ARG, "pl_sensitivity: HeatExchange needs tem, item, dbds");
    if (a->kind == PL_SENS_AAD_BRINKMAN_DIFF &&
        (!a->tem || !a->item || !a->iqx || !a->iqy || (d3 && !a->iqz) || !a->gsnap || !a->igsnap || !a->diffusivity || !a->dkds))
        return fail(PL_ERR_ARG, "pl_sensitivity: BrinkmanDiffusivity needs tem, item, iq, the two snapshots, diffusivity and dkds");
    SensArgs A{};
    A.kind = a->kind; A.dfds = a->dfds; A.ux = a->ux; A.uy = a->uy; A.uz = a->uz; A.imx = a->imx; A.imy = a->imy; A.imz = a->imz; A.dads = a->dads;
    A.tem = a->tem; A.item = a->item; A.iqx = a->iqx; A.iqy = a->iqy; A.iqz = a->iqz; A.gsnap = a->gsnap; A.igsnap = a->igsnap;
    A.kappa = a->diffusivity; A.dkds = a->dkds; A.dbds = a->dbds; A.pitch = (size_t)l->g.nxyz;
    if (d3 && !planar) LAUNCH(k_sensitivity<3>, blocks_for(l->g.nxyz, 256), 256, l->g, A);
    else LAUNCH(k_sensitivity<2>, blocks_for(l->g.nxyz, 256), 256, l->g, A);      // (the Brinkman kind reads nothing lattice-specific but nxyz / npacked)
    return PL_OK;
}
int pl_sensitivity_heat_source(pl_lattice* l, const pl_bc* plane, double* dfds, const double* ux, const double* uy, const double* uz,
                                const double* igsnap, const double* diffusivity, const double* dkds) {
    if (!l || !plane || !dfds || !ux || !uy || (l->kind == PL_D3Q15 && !uz) || !igsnap || !diffusivity || !dkds)
        return fail(PL_ERR_ARG, "pl_sensitivity_heat_source: null argument");
    if (plane->empty) return PL_OK;
    if (!same_shape(plane->lat, l)) return fail(PL_ERR_ARG, "pl_sensitivity_heat_source: plane was created for a lattice of another shape");
    if (!plane->v0) return fail(PL_ERR_ARG, "pl_sensitivity_heat_source: the plane carries no qn values");
    ClosureArgs A{};
    A.type = 0; A.pl = plane->pl; A.mask = plane->mask; A.v0 = plane->v0; A.ux = ux; A.uy = uy; A.uz = uz; A.kappa = diffusivity;
    int np = plane->pl.n1*plane->pl.n2;
    if (l->kind == PL_D2Q9) LAUNCH(k_sens_heat_source<2>, blocks_for(np, 128), 128, l->g, A, igsnap, dkds, dfds);
    else LAUNCH(k_sens_heat_source<3>, blocks_for(np, 128), 128, l->g, A, igsnap, dkds, dfds);
    return PL_OK;
}

}  // extern "C"
