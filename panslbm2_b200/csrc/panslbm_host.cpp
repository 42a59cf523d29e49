// Host-pointer surface of libpanslbm_b200.so (plh_*, see include/panslbm_c.h): what the drop-in C++ headers in
// panslbm2_b200/src/ call.  The reference's drivers own plain host arrays (`new double[nxyz]`, production/heatsink3D.cpp:50-59),
// pass them to every call, std::swap them between steps (:178-183) and read them directly whenever they like (:231).
// This file keeps that contract on top of the device-pointer C-ABI (pl_*):
//
//   1. coherence  every array handed out by plh_alloc (the headers route operator new[] here) has a lazily created
//      device mirror and one of three states, enforced with page protection on the host copy:
//        HOST   host copy current, device stale      host pages read/write
//        SHARED both current                         host pages read-only  (first host write faults -> HOST)
//        DEVICE device copy current, host stale      host pages no access  (first host touch faults -> sync, copy back -> SHARED)
//      A kernel reading an array needs SHARED/DEVICE (upload if HOST); a kernel writing it moves it to DEVICE.  In the
//      steady state of a time loop no array changes state, so no copy and no mprotect happens per step.
//      Pointers that were not allocated here (std::vector storage, stack arrays) are staged through a transient device
//      buffer around the call (synchronous; correct but slow — never on the time-loop arrays of the reference drivers).
//      A block the host has not touched since plh_alloc (no page of it present or swapped, /proc/self/pagemap) is mirrored
//      by a zero-filled device buffer without any upload: the transient drivers allocate one set of arrays per time step
//      that only the collide ever writes (production/heatsink3D_transient.cpp:50-57).
//   1b. state store  the mirrors are the HBM-resident store of those per-step states.  When a device allocation would
//      exceed the budget (PANSLBM_B200_DEVICE_BUDGET_MB, else when cudaMalloc fails), mirrors that were not used in the
//      current or the previous loop iteration are spilled to their host copies: the ones farthest behind the direction in
//      which the loop walks the arrays (allocation order == time order in the drivers) — the oldest states in the forward
//      loop, the already consumed ones in the time-reversed adjoint loop (heatsink3D_transient.cpp:190-215) — and come
//      back on demand.
//   2. fusion     collide / Stream / closures / SmoothCorner arrive as separate calls per lattice.  The engine executes them
//      one by one while it LEARNs two consecutive loop iterations, builds a pl_plan from them (the two argument sets the
//      driver alternates between), and then REPLAYs: Stream/closure/SmoothCorner calls that match the recorded iteration are
//      only checked off, and the next collide call executes "stream + closures + SmoothCorner + collide" as ONE fused pass
//      (pl_plan_advance).  Anything unexpected — a different call, an observation of the populations, the end of the
//      loop — settles the checked-off calls first (one standalone pass, or call by call) and falls back to LEARN.
//      Results are identical to call-by-call execution.  "Match" means the same calls with the same scalars and the same
//      arrays PRESENT; the array addresses themselves may differ from step to step (the transient drivers pass rho[t],
//      ux[t], ..., gi[t], production/heatsink3D_transient.cpp:156-176): the plan is then re-bound (pl_plan_rebind) with the
//      arrays of the step before the fused pass is queued.
//
// Single-threaded callers, as everywhere in this library.  TEST NOTE: no CPU arithmetic lives here; every number is
// produced by the CUDA kernels behind pl_*.
#include "../../include/panslbm_c.h"

#include <fcntl.h>
#include <pthread.h>
#include <signal.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

enum { ST_HOST = 0, ST_SHARED = 1, ST_DEVICE = 2 };
enum { BK_ARRAY = 0, BK_POP0 = 1, BK_POPF = 2 };

struct Block {
    char* base = nullptr;
    size_t bytes = 0, map_bytes = 0;
    double* dev = nullptr;
    int state = ST_HOST;
    int kind = BK_ARRAY;
    pl_lattice* lat = nullptr;      // population views only
    uint64_t seq = 0;               // allocation order (the drivers allocate their per-step arrays in time order)
    uint64_t last = 0;              // loop iteration (g_tick) of the last device use
    bool maybe_fresh = true;        // no device use yet: the host may never have touched it (see untouched())
    bool spilled = false;           // its mirror was given up under memory pressure
    // a thermal snapshot (`_g` / `_ig` of the collides, consumed by Sensitivity*): the device keeps it SoA [c][nxyz], the reference's
    // host layout is [pack][c][lane] / [idx][c] (advection_avx.h:1046-1052, 1093-1098) — converted whenever it crosses, so that a
    // driver that reads, checkpoints or provides one sees the reference's layout
    int snap_kind = 0;              // 0 = not a snapshot, else PL_D2Q9 / PL_D3Q15 of the lattice it belongs to (its nxyz = bytes/8/nc)
    long long snap_n() const { return (long long)(bytes/sizeof(double)/(snap_kind == PL_D2Q9 ? 9 : 15)); }
};
std::map<uintptr_t, Block> g_blocks;              // by base address
// The allocation hook (operator new of the drop-in headers) may be reached from any thread of the caller — an OpenMP region, a
// library thread — while the fault handler and every call of this file walk the map: a spin lock around the map itself (never
// held across anything that can fault or block; map nodes are stable, so a Block* stays valid after the lock is dropped).
// Recursive: the map's own node allocations go through the program's replaced operator new / delete, and the delete hook asks
// plh_owns() — on the thread that already holds the lock.
std::atomic<unsigned long> g_blocks_owner{0};
int g_blocks_depth = 0;
struct BlocksLock {
    BlocksLock() {
        const unsigned long me = (unsigned long)pthread_self();
        if (g_blocks_owner.load(std::memory_order_acquire) == me) { ++g_blocks_depth; return; }
        unsigned long none = 0;
        while (!g_blocks_owner.compare_exchange_weak(none, me, std::memory_order_acquire)) none = 0;
        g_blocks_depth = 1;
    }
    ~BlocksLock() { if (--g_blocks_depth == 0) g_blocks_owner.store(0, std::memory_order_release); }
};
struct Views { Block *f0 = nullptr, *f = nullptr; };
std::map<pl_lattice*, Views> g_views;
uint64_t g_stat[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // fused steps, unfused ops, uploads, downloads, faults, plans, settles, transient stagings
uint64_t g_store[4] = {0, 0, 0, 0};              // mirrors spilled to the host, mirrors restored, device bytes held, peak device bytes
uint64_t g_seq = 0, g_tick = 0;
uint64_t g_last_new_seq = 0;                      // seq of the last block that needed a new mirror, and the direction of travel
int g_direction = 1;
size_t g_budget = 0;                              // bytes; 0 = unlimited (spill only when cudaMalloc fails)
bool g_budget_read = false;
size_t g_page = 4096;
bool g_handler = false;
struct sigaction g_prev;
std::string g_herr;

// PANSLBM_B200_PROFILE=1: wall time spent inside each entry point of this file, printed at exit (host-side tuning aid)
enum { T_COLLIDE, T_STREAM, T_SMOOTH, T_BC, T_INIT, T_RESIDUAL, T_SENS, T_SENS_HS, T_FILTER, T_SYNC, T_FETCH, T_MIRROR, T_NTIMERS };
const char* const g_tname[T_NTIMERS] = {"collide", "stream", "smooth_corner", "bc", "initial_condition", "residual", "sensitivity", "sensitivity_heat_source",
                                        "filter", "sync", "fetch(fault)", "mirror alloc/spill"};
double g_tms[T_NTIMERS];
uint64_t g_tcalls[T_NTIMERS];
bool g_profile = false;
struct HostTimer {
    int id; std::chrono::steady_clock::time_point t0;
    explicit HostTimer(int i) : id(i) { if (g_profile) t0 = std::chrono::steady_clock::now(); }
    ~HostTimer() { if (g_profile) { g_tms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++g_tcalls[id]; } }
};
void profile_dump() {
    for (int k = 0; k < T_NTIMERS; ++k) if (g_tcalls[k]) fprintf(stderr, "panslbm_b200 host profile: %-24s %10llu calls %12.3f ms\n", g_tname[k], (unsigned long long)g_tcalls[k], g_tms[k]);
}
struct ProfileInit { ProfileInit() { const char* v = getenv("PANSLBM_B200_PROFILE"); if (v && *v && *v != '0') { g_profile = true; atexit(profile_dump); } } } g_profile_init;

int hfail(const char* what) {
    g_herr = std::string(what) + ": " + pl_last_error();
    return PL_ERR_CUDA;
}

// Device memory of the mirrors: carved from slabs and recycled by size.  cudaMalloc/cudaFree per array would cost milliseconds
// and synchronise the device — the transient loops mirror nine new arrays per time step.  All device work is ordered on the
// library's stream, so a recycled buffer cannot be overtaken by its previous user.
struct DevPool {
    std::map<size_t, std::vector<double*>> free_;
    char* slab = nullptr;
    size_t left = 0, next_slab = (size_t)64 << 20;
    double* get(size_t bytes) {
        auto it = free_.find(bytes);
        if (it != free_.end() && !it->second.empty()) { double* p = it->second.back(); it->second.pop_back(); return p; }
        const size_t need = (bytes + 255)/256*256;
        if (left < need) {
            if (left >= 4096) free_[left/256*256].push_back((double*)slab);     // the tail of the old slab stays usable
            size_t want = need > next_slab ? need : next_slab;
            double* p = pl_array_alloc(want/sizeof(double));
            if (!p && want > need) { want = need; p = pl_array_alloc(want/sizeof(double)); }
            if (!p) { left = 0; return nullptr; }
            slab = (char*)p; left = want;
            if (next_slab < ((size_t)4 << 30)) next_slab *= 2;
        }
        double* p = (double*)slab;
        slab += need; left -= need;
        return p;
    }
    void put(double* p, size_t bytes) { free_[bytes].push_back(p); }
} g_pool;

Block* find_block(const void* p) {
    BlocksLock lock_;
    if (g_blocks.empty()) return nullptr;
    auto it = g_blocks.upper_bound((uintptr_t)p);
    if (it == g_blocks.begin()) return nullptr;
    --it;
    Block& b = it->second;
    return ((const char*)p < b.base + b.map_bytes) ? &b : nullptr;
}
void protect(Block* b, int prot) { mprotect(b->base, b->map_bytes, prot); }

void settle_all();
int flush_pending();
int acquire_mirror(Block* b);

// bring the host copy of a block up to date (it is in state DEVICE) and make it readable
void fetch(Block* b) {
    HostTimer timer_(T_FETCH);
    flush_pending();      // fused passes the engine still holds back may write this array
    pl_synchronize();
    if (b->kind == BK_ARRAY) {
        protect(b, PROT_READ | PROT_WRITE);
        if (b->snap_kind) pl_snapshot_convert(b->snap_kind, b->snap_n(), b->dev, (double*)b->base, 1);
        else pl_array_download((double*)b->base, b->dev, b->bytes/sizeof(double));
        protect(b, PROT_READ);
        b->state = ST_SHARED;
    } else {
        settle_all();     // checked-off Stream/closure calls change what the populations are
        Views& v = g_views[b->lat];
        protect(v.f0, PROT_READ | PROT_WRITE); protect(v.f, PROT_READ | PROT_WRITE);
        pl_lattice_get_host(b->lat, (double*)v.f0->base, (double*)v.f->base);
        protect(v.f0, PROT_READ); protect(v.f, PROT_READ);
        v.f0->state = v.f->state = ST_SHARED;
    }
    ++g_stat[3];
}

std::atomic_flag g_fault_lock = ATOMIC_FLAG_INIT;     // host threads of the caller (OpenMP loops over its arrays) may fault together
void on_fault(int sig, siginfo_t* si, void* uc) {
    // A fault taken while this thread is INSIDE a library call (the CUDA runtime reading a caller's buffer whose host copy is
    // stale, e.g. a raw pointer handed to pl_comm_* or pl_array_upload) cannot be served: fetching would re-enter the CUDA
    // runtime from its own signal context.  Such pointers must go through plh_host_acquire first (the mpi.h shim does).
    if (pl_in_call() && find_block(si->si_addr)) {
        static const char msg[] = "panslbm_b200: a host array whose current copy lives on the device was handed to a device-pointer entry point "
                                  "(pl_*); call plh_host_acquire(ptr, bytes, for_write) on it first\n";
        ssize_t w = write(2, msg, sizeof(msg) - 1); (void)w;
        abort();
    }
    while (g_fault_lock.test_and_set(std::memory_order_acquire)) {}
    Block* b = find_block(si->si_addr);
    if (b && b->state == ST_DEVICE) { ++g_stat[4]; fetch(b); g_fault_lock.clear(std::memory_order_release); return; }
    if (b && b->state == ST_SHARED) {
        flush_pending();      // ... or read it: they must see the content it had when the collide was called
        ++g_stat[4]; protect(b, PROT_READ | PROT_WRITE); b->state = ST_HOST;
        g_fault_lock.clear(std::memory_order_release);
        return;
    }
    g_fault_lock.clear(std::memory_order_release);
    // not ours: hand over to whoever was there before (default action: re-raise and die as usual)
    if (g_prev.sa_flags & SA_SIGINFO) { if (g_prev.sa_sigaction) { g_prev.sa_sigaction(sig, si, uc); return; } }
    else if (g_prev.sa_handler != SIG_DFL && g_prev.sa_handler != SIG_IGN) { g_prev.sa_handler(sig); return; }
    signal(SIGSEGV, SIG_DFL);
}
void install_handler() {
    if (g_handler) return;
    g_page = (size_t)sysconf(_SC_PAGESIZE);
    struct sigaction sa;
    memset(&sa, 0, sizeof(sa));
    sa.sa_sigaction = on_fault;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGSEGV, &sa, &g_prev);
    g_handler = true;
}

Block* new_block(size_t bytes, int kind, pl_lattice* lat, int state, int prot) {
    install_handler();
    size_t mb = (bytes + g_page - 1)/g_page*g_page;
    if (mb == 0) mb = g_page;
    // large blocks: 2 MB aligned and advised for transparent huge pages, so that the first host touch of a block the device
    // filled (a spilled state, a field read after the loops) costs one fault per 2 MB instead of one per 4 KB
    const size_t huge = (size_t)2 << 20;
    const bool big = mb >= 2*huge;
    void* raw = mmap(nullptr, big ? mb + huge : mb, prot, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (raw == MAP_FAILED) return nullptr;
    void* p = raw;
    if (big) {
        const uintptr_t a = ((uintptr_t)raw + huge - 1)/huge*huge;
        if (a > (uintptr_t)raw) munmap(raw, a - (uintptr_t)raw);
        const size_t tail = (uintptr_t)raw + mb + huge - (a + mb);
        if (tail) munmap((void*)(a + mb), tail);
        p = (void*)a;
        madvise(p, mb, MADV_HUGEPAGE);
    }
    Block b;
    BlocksLock lock_;
    b.base = (char*)p; b.bytes = bytes; b.map_bytes = mb; b.kind = kind; b.lat = lat; b.state = state; b.seq = ++g_seq;
    auto r = g_blocks.emplace((uintptr_t)p, b);
    return &r.first->second;
}
void drop_block(Block* b) {
    if (b->dev) { g_pool.put(b->dev, b->map_bytes); g_store[2] -= b->map_bytes; }
    char* base = b->base; size_t mb = b->map_bytes;
    { BlocksLock lock_; g_blocks.erase((uintptr_t)base); }
    munmap(base, mb);
}

// ---- translation of one array argument ----------------------------------------------------------------------
struct Staged { double* dev; double* host; size_t n; bool write; };
std::vector<Staged> g_staged;      // transient mirrors of the call being translated

// true if the host never touched any page of the block since plh_alloc: every page-map entry is neither present nor swapped,
// so the content is still the zero pages of the anonymous mapping.  Any doubt (no pagemap, short read) answers false.
bool untouched(const Block* b) {
    static int fd = -2;
    if (fd == -2) fd = open("/proc/self/pagemap", O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    const size_t npages = b->map_bytes/g_page;
    static uint64_t buf[8192];
    size_t done = 0;
    while (done < npages) {
        const size_t m = npages - done < 8192 ? npages - done : 8192;
        const ssize_t got = pread(fd, buf, m*8, (off_t)(((uintptr_t)b->base/g_page + done)*8));
        if (got != (ssize_t)(m*8)) return false;
        for (size_t k = 0; k < m; ++k) if (buf[k] & (3ull << 62)) return false;      // bit 63 present, bit 62 swapped
        done += m;
    }
    return true;
}

// ---- the state store: device mirrors under a budget ------------------------------------------------------------------
// give up the mirror of `v`: bring the host copy up to date first if the device holds the only current one
int spill(Block* v, double** keep) {
    flush_pending();
    if (v->state == ST_DEVICE) {
        pl_synchronize();
        protect(v, PROT_READ | PROT_WRITE);
        if (v->snap_kind ? pl_snapshot_convert(v->snap_kind, v->snap_n(), v->dev, (double*)v->base, 1) : pl_array_download((double*)v->base, v->dev, v->bytes/sizeof(double))) return hfail("spill download");
        ++g_stat[3];
    } else if (v->state == ST_SHARED) {
        pl_synchronize();
        protect(v, PROT_READ | PROT_WRITE);
    }
    v->state = ST_HOST;
    if (keep) *keep = v->dev;            // handed over to the block that needs a mirror of the same size
    else { g_pool.put(v->dev, v->map_bytes); g_store[2] -= v->map_bytes; }
    v->dev = nullptr;
    v->spilled = true;
    ++g_store[0];
    return PL_OK;
}
// the mirror farthest behind the direction of travel among those not used in this or the previous iteration
Block* pick_victim(const Block* need, bool same_size) {
    Block* best = nullptr;
    BlocksLock lock_;
    for (auto& kv : g_blocks) {
        Block& c = kv.second;
        if (c.kind != BK_ARRAY || !c.dev || &c == need || c.last + 1 >= g_tick) continue;
        if (same_size && c.map_bytes != need->map_bytes) continue;
        if (!best || (g_direction > 0 ? c.seq < best->seq : c.seq > best->seq)) best = &c;
    }
    return best;
}
int acquire_mirror(Block* b) {
    HostTimer timer_(T_MIRROR);
    if (!g_budget_read) {
        const char* v = getenv("PANSLBM_B200_DEVICE_BUDGET_MB");
        g_budget = v && *v ? (size_t)atoll(v) << 20 : 0;
        g_budget_read = true;
    }
    if (b->seq != g_last_new_seq) { g_direction = b->seq > g_last_new_seq ? 1 : -1; g_last_new_seq = b->seq; }
    const bool restore = b->spilled;
    b->spilled = false;
    if (g_budget && g_store[2] + b->map_bytes > g_budget) {
        // over budget: take over the mirror of a victim of the same size, else free victims until the new one fits
        if (Block* v = pick_victim(b, true)) {
            int rc = spill(v, &b->dev);
            if (rc) return rc;
        } else {
            while (g_store[2] + b->map_bytes > g_budget) {
                Block* w = pick_victim(b, false);
                if (!w) break;           // everything left is in use: exceed the budget rather than fail
                int rc = spill(w, nullptr);
                if (rc) return rc;
            }
        }
    }
    while (!b->dev) {
        b->dev = g_pool.get(b->map_bytes);
        if (b->dev) { g_store[2] += b->map_bytes; break; }
        Block* w = pick_victim(b, false);        // device memory exhausted: spill and retry
        if (!w) return hfail("device mirror");
        int rc = spill(w, nullptr);
        if (rc) return rc;
    }
    if (g_store[2] > g_store[3]) g_store[3] = g_store[2];
    if (restore) ++g_store[1];
    return PL_OK;
}

// device address of host pointer `h` (n doubles) for a kernel that reads it (rd) and/or writes it (wr).  wr && !rd: the kernel
// overwrites EVERY element (the macroscopic outputs and the snapshot of a storing collide, a filter result): whatever the array
// held before is irrelevant, so its mirror needs neither an upload nor a zero-fill — the transient drivers hand nine such
// arrays (194 MB at 81 x 161 x 81) to every step of the forward loop (production/heatsink3D_transient.cpp:156-160).
int xlate(const double* h, size_t n, bool rd, bool wr, double** out) {
    *out = nullptr;
    if (!h) return PL_OK;
    Block* b = find_block(h);
    if (b && b->kind == BK_ARRAY) {
        b->last = g_tick;
        if (!b->dev) {
            int rc = acquire_mirror(b);
            if (rc) return rc;
        }
        const bool overwrite = wr && !rd;
        if (!overwrite && b->state == ST_HOST && b->maybe_fresh && untouched(b)) {
            // never touched by the host: its content is the zero pages mmap would hand out — no upload
            if (pl_array_fill(b->dev, 0.0, b->map_bytes/sizeof(double))) return hfail("mirror fill");
            b->state = ST_SHARED;
            if (!wr) protect(b, PROT_READ);
        }
        b->maybe_fresh = false;
        if (!overwrite && b->state == ST_HOST) {      // also before a partial write: the rest of the array must survive
            flush_pending();            // passes held back were called with the previous content of the mirror
            if (b->snap_kind ? pl_snapshot_convert(b->snap_kind, b->snap_n(), (const double*)b->base, b->dev, 0) : pl_array_upload(b->dev, (const double*)b->base, b->bytes/sizeof(double))) return hfail("upload");
            ++g_stat[2];
            b->state = ST_SHARED;
            if (!wr) protect(b, PROT_READ);
        }
        if (wr && b->state != ST_DEVICE) { protect(b, PROT_NONE); b->state = ST_DEVICE; }
        *out = b->dev + ((const char*)h - b->base)/sizeof(double);
        return PL_OK;
    }
    // foreign memory: stage through a transient device buffer
    Staged s;
    s.host = const_cast<double*>(h); s.n = n; s.write = wr;
    s.dev = pl_array_alloc(n);
    if (!s.dev) return hfail("staging buffer");
    if (pl_array_upload(s.dev, h, n)) return hfail("staging upload");
    (void)rd;
    ++g_stat[7];
    g_staged.push_back(s);
    *out = s.dev;
    return PL_OK;
}
// `h` is a thermal snapshot of lattice `l` (only whole blocks of exactly nc*nxyz doubles can be converted)
void tag_snapshot(const double* h, pl_lattice* l, size_t n_nc) {
    if (!h || !l) return;
    Block* b = find_block(h);
    if (b && b->kind == BK_ARRAY && (const char*)h == b->base && b->bytes == n_nc*sizeof(double)) {
        int info[18];
        pl_lattice_info(l, info);
        b->snap_kind = info[17] == 9 ? PL_D2Q9 : PL_D3Q15;
    }
}
int unstage() {
    int rc = PL_OK;
    for (auto& s : g_staged) {
        if (s.write && pl_array_download(s.host, s.dev, s.n)) rc = hfail("staging download");
        else if (!s.write) pl_synchronize();
        pl_array_free(s.dev);
    }
    g_staged.clear();
    return rc;
}

// the populations of `l` are about to change on the device: its host views (public f0/f) go stale
void pops_written(pl_lattice* l) {
    auto it = g_views.find(l);
    if (it == g_views.end()) return;
    for (Block* b : {it->second.f0, it->second.f})
        if (b->state != ST_DEVICE) { protect(b, PROT_NONE); b->state = ST_DEVICE; }
}
// the host wrote into f0/f since the last device operation: import them
int pops_sync_in(pl_lattice* l) {
    auto it = g_views.find(l);
    if (it == g_views.end()) return PL_OK;
    Views& v = it->second;
    if (v.f0->state == ST_HOST || v.f->state == ST_HOST) {
        if (pl_lattice_set_host(l, (const double*)v.f0->base, (const double*)v.f->base)) return hfail("pl_lattice_set_host");
        for (Block* b : {v.f0, v.f}) { protect(b, PROT_READ); b->state = ST_SHARED; }
        ++g_stat[2];
    }
    return PL_OK;
}

// ---- the fusion engine -------------------------------------------------------------------------------------------
enum { OP_STREAM = 0, OP_BC = 1, OP_SMOOTH = 2, OP_SMOOTH_AT = 3 };
struct Op {
    int kind = 0;
    pl_lattice* l = nullptr;
    pl_lattice* other = nullptr;
    const pl_bc* bc = nullptr;
    pl_bc_aux aux;
    bool has_aux = false;
    int inverse = 0;
    int at[6] = {0, 0, 0, 0, 0, 0};      // OP_SMOOTH_AT: i, j, k, dx, dy, dz
};
bool same_aux(const Op& a, const Op& b) { return a.has_aux == b.has_aux && (!a.has_aux || memcmp(&a.aux, &b.aux, sizeof(pl_bc_aux)) == 0); }
bool same_shape(const Op& a, const Op& b) {
    return a.kind == b.kind && a.l == b.l && a.other == b.other && a.bc == b.bc && a.inverse == b.inverse && memcmp(a.at, b.at, sizeof(a.at)) == 0;
}
bool same_args(const pl_collide_args& a, const pl_collide_args& b) { return memcmp(&a, &b, sizeof(pl_collide_args)) == 0; }
// equal up to the array addresses: same model, flags and scalars, and the same arrays present
bool like_args(const pl_collide_args& a, const pl_collide_args& b) {
    if (a.model != b.model || a.issave != b.issave || a.viscosity != b.viscosity || a.diffusivity_const != b.diffusivity_const ||
        a.gx != b.gx || a.gy != b.gy || a.gz != b.gz || a.tem0 != b.tem0) return false;
#define SAME_NULL(f) if ((a.f == nullptr) != (b.f == nullptr)) return false
    SAME_NULL(alpha); SAME_NULL(diffusivity); SAME_NULL(beta); SAME_NULL(dirx); SAME_NULL(diry); SAME_NULL(dirz);
    SAME_NULL(rho); SAME_NULL(ux); SAME_NULL(uy); SAME_NULL(uz); SAME_NULL(tem); SAME_NULL(qx); SAME_NULL(qy); SAME_NULL(qz);
    SAME_NULL(ip); SAME_NULL(iux); SAME_NULL(iuy); SAME_NULL(iuz); SAME_NULL(imx); SAME_NULL(imy); SAME_NULL(imz);
    SAME_NULL(item); SAME_NULL(iqx); SAME_NULL(iqy); SAME_NULL(iqz); SAME_NULL(snapshot);
#undef SAME_NULL
    return true;
}
bool like_aux(const Op& a, const Op& b) {
    if (a.has_aux != b.has_aux) return false;
    if (!a.has_aux) return true;
    const pl_bc_aux &x = a.aux, &y = b.aux;
    return x.diffusivity_const == y.diffusivity_const && x.eps == y.eps && (x.rho == nullptr) == (y.rho == nullptr) && (x.ux == nullptr) == (y.ux == nullptr) &&
           (x.uy == nullptr) == (y.uy == nullptr) && (x.uz == nullptr) == (y.uz == nullptr) && (x.tem == nullptr) == (y.tem == nullptr) &&
           (x.diffusivity == nullptr) == (y.diffusivity == nullptr);
}

struct Iter {
    bool have_c = false;
    pl_lattice *f = nullptr, *g = nullptr;
    pl_collide_args c;
    std::vector<Op> ops;
};
struct Plan {
    pl_plan* p = nullptr;
    pl_lattice *f = nullptr, *g = nullptr;
    pl_collide_args c[2];         // what the caller passes now (the device plan holds the same unless flagged below)
    std::vector<Op> ops[2];
    bool dirty_c[2] = {false, false}, dirty_aux[2] = {false, false};
};
// hand changed array bindings of argument set `par` to the device plan
int flush_bindings(Plan* pl, int par) {
    if (!pl->dirty_c[par] && !pl->dirty_aux[par]) return PL_OK;
    std::vector<pl_bc_aux> aux;
    if (pl->dirty_aux[par]) for (const Op& o : pl->ops[par]) if (o.kind == OP_BC && o.has_aux) aux.push_back(o.aux);
    if (pl_plan_rebind(pl->p, par, pl->dirty_c[par] ? &pl->c[par] : nullptr, aux.empty() ? nullptr : aux.data(), (int)aux.size())) return hfail("pl_plan_rebind");
    pl->dirty_c[par] = pl->dirty_aux[par] = false;
    return PL_OK;
}
struct Engine {
    Iter hist[2];
    int nhist = 0;
    Iter cur;
    Plan* active = nullptr;
    int par = 0;          // REPLAY: argument set of the last collide executed
    int pos = 0;          // REPLAY: Stream/closure/SmoothCorner calls of the current iteration checked off so far
    // REPLAY: fused passes accepted but not yet queued on the device (0..2, the most recent ones).  The reference stores its
    // macroscopic fields and the thermal snapshot at every site on every step (production/heatsink3D.cpp:151); a pass is queued
    // WITHOUT those stores (pl_plan_advance_observed, save_last = 0) once the collide call two steps later has arrived with no
    // observation in between: by then both alternating argument sets are being overwritten by newer steps, so nothing can
    // ever see what that pass would have stored.  Any observation — a call that reads an array (Residual, Sensitivity*,
    // filters), a host access to a mirrored array (page fault), the end of the loop — first queues the held-back passes WITH
    // their stores, which leaves every array exactly as step-by-step execution would.
    int pending = 0;
    // how many passes may be held back: 2 proves non-observation; small lattices (PANSLBM_COOP_SITES) hold back more, because
    // the library runs a batch of passes in ONE cooperative launch (k_steps) where a step is bound by launch latency
    int lag = 2;
    std::vector<Plan*> plans;
} E;

int exec_op(const Op& o) {
    ++g_stat[1];
    pops_written(o.l);
    switch (o.kind) {
        case OP_STREAM: return pl_stream(o.l, o.inverse);
        case OP_BC: return pl_bc_apply(o.l, o.other, o.bc, o.has_aux ? &o.aux : nullptr);
        case OP_SMOOTH_AT: return pl_smooth_corner_at(o.l, o.at[0], o.at[1], o.at[2], o.at[3], o.at[4], o.at[5]);
        default: return pl_smooth_corner(o.l);
    }
}
int flush_pending() {
    if (!E.active || E.pending == 0) return PL_OK;
    const int n = E.pending;
    E.pending = 0;
    // only the two most recent passes can still be observed: older ones of the batch are overwritten by them
    if (pl_plan_advance_observed(E.active->p, n, 0, 2)) return hfail("pl_plan_advance");
    return PL_OK;
}
int settle() {
    if (!E.active) return PL_OK;
    Plan* pl = E.active;
    int rc = flush_pending();
    if (rc) return rc;
    const int nops = (int)pl->ops[0].size();
    if (E.pos == nops && nops > 0) { rc = flush_bindings(pl, E.par); if (!rc) rc = pl_plan_advance(pl->p, 0, 1); }
    else for (int k = 0; k < E.pos && !rc; ++k) rc = exec_op(pl->ops[E.par][k]);
    ++g_stat[6];
    E.active = nullptr; E.pos = 0; E.nhist = 0; E.cur = Iter();
    return rc ? hfail("settle") : PL_OK;
}
void settle_all() { settle(); }

// can [collide, ops] be expressed as a pl_plan?  streams of every lattice first (one direction), then closures, then SmoothCorner
bool fusable(const Iter& a, const Iter& b) {
    if (!a.have_c || !b.have_c || a.f != b.f || a.g != b.g || a.c.model != b.c.model || a.ops.size() != b.ops.size() || a.ops.empty()) return false;
    const size_t nlat = a.g ? 2 : 1;
    if (a.ops.size() < nlat) return false;
    for (size_t k = 0; k < a.ops.size(); ++k) {
        if (!same_shape(a.ops[k], b.ops[k])) return false;
        const Op& o = a.ops[k];
        if (o.l != a.f && o.l != a.g) return false;
        if (k < nlat) { if (o.kind != OP_STREAM || o.inverse != a.ops[0].inverse) return false; }
        else if (o.kind == OP_STREAM) return false;
    }
    if (nlat == 2 && a.ops[0].l == a.ops[1].l) return false;
    // closures, SmoothCorner (once per lattice) and SmoothCornerAt in any order: pl_plan_finalize decides whether the body can
    // be run as "closures, SmoothCorner, SmoothCornerAt" and refuses otherwise (the plan is then kept as a tombstone)
    bool sf = false, sg = false;
    for (size_t k = nlat; k < a.ops.size(); ++k) {
        const Op& o = a.ops[k];
        if (o.kind == OP_SMOOTH) {
            bool& s = o.l == a.f ? sf : sg;
            if (s) return false;
            s = true;
        }
    }
    return true;
}
Plan* build_plan(const Iter& a, const Iter& b) {
    Plan* pl = new Plan();
    pl->f = a.f; pl->g = a.g; pl->c[0] = a.c; pl->c[1] = b.c; pl->ops[0] = a.ops; pl->ops[1] = b.ops;
    pl->p = pl_plan_create(a.f, a.g);
    bool ok = pl->p != nullptr;
    ok = ok && pl_plan_set_collide(pl->p, &a.c, &b.c) == PL_OK && pl_plan_set_stream(pl->p, a.ops[0].inverse) == PL_OK;
    int sf = 0, sg = 0;
    for (size_t k = 0; ok && k < a.ops.size(); ++k) {
        const Op &o = a.ops[k], &o2 = b.ops[k];
        const int on_g = o.l == a.g && a.g ? 1 : 0;
        if (o.kind == OP_BC) ok = pl_plan_add_bc(pl->p, on_g, o.bc, o.has_aux ? &o.aux : nullptr, o2.has_aux ? &o2.aux : nullptr) == PL_OK;
        else if (o.kind == OP_SMOOTH) {      // raised where it is called: the plan records the call order
            if (o.l == a.f) sf = 1; else sg = 1;
            ok = pl_plan_set_smooth_corner(pl->p, sf, sg) == PL_OK;
        } else if (o.kind == OP_SMOOTH_AT) ok = pl_plan_add_smooth_corner_at(pl->p, on_g, o.at[0], o.at[1], o.at[2], o.at[3], o.at[4], o.at[5]) == PL_OK;
    }
    ok = ok && pl_plan_finalize(pl->p) == PL_OK;
    if (!ok) { if (pl->p) pl_plan_destroy(pl->p); pl->p = nullptr; }     // kept as a tombstone: do not try this shape again
    E.plans.push_back(pl);
    ++g_stat[5];
    return pl->p ? pl : nullptr;
}
bool same_plan_shape(const Plan* pl, const Iter& a, const Iter& b) {
    if (pl->f != a.f || pl->g != a.g || pl->ops[0].size() != a.ops.size()) return false;
    for (int par = 0; par < 2; ++par) {
        const Iter& it = par ? b : a;
        if (!like_args(pl->c[par], it.c)) return false;
        for (size_t k = 0; k < a.ops.size(); ++k) if (!same_shape(pl->ops[par][k], it.ops[k]) || !like_aux(pl->ops[par][k], it.ops[k])) return false;
    }
    return true;
}
// a known plan meets the same loop body with other arrays: take the arrays of the two learned iterations
void adopt_bindings(Plan* pl, const Iter& a, const Iter& b) {
    for (int par = 0; par < 2; ++par) {
        const Iter& it = par ? b : a;
        if (!same_args(pl->c[par], it.c)) { pl->c[par] = it.c; pl->dirty_c[par] = true; }
        for (size_t k = 0; k < it.ops.size(); ++k)
            if (!same_aux(pl->ops[par][k], it.ops[k])) { pl->ops[par][k].aux = it.ops[k].aux; pl->dirty_aux[par] = true; }
    }
}
void drop_plans_of(pl_lattice* l) {
    for (size_t k = 0; k < E.plans.size();) {
        if (E.plans[k]->f == l || E.plans[k]->g == l) {
            if (E.plans[k]->p) pl_plan_destroy(E.plans[k]->p);
            delete E.plans[k];
            E.plans.erase(E.plans.begin() + k);
        } else ++k;
    }
}

int enter_replay(Plan* pl, int par, const pl_collide_args& d) {
    if (!same_args(pl->c[par], d)) { pl->c[par] = d; pl->dirty_c[par] = true; }
    int rc = flush_bindings(pl, par);
    if (rc) return rc;
    if (pl_plan_set_parity(pl->p, par)) return hfail("pl_plan_set_parity");
    if (pl_plan_advance(pl->p, 1, 0)) return hfail("pl_plan_advance");
    ++g_stat[0];
    E.active = pl; E.par = par; E.pos = 0; E.nhist = 0; E.cur = Iter();
    {
        static long long coop_sites = -1;
        if (coop_sites < 0) { const char* v = getenv("PANSLBM_COOP_SITES"); coop_sites = v && *v ? atoll(v) : 0; }
        int info[18];
        pl_lattice_info(pl->f, info);
        E.lag = (long long)info[13] <= coop_sites ? 32 : 2;
    }
    return PL_OK;
}

int do_collide(pl_lattice* f, pl_lattice* g, const pl_collide_args& d, bool staged) {
    pops_written(f); if (g) pops_written(g);
    if (E.active) {
        Plan* pl = E.active;
        const int nops = (int)pl->ops[0].size();
        const int next = E.par ^ 1;
        if (!staged && pl->f == f && pl->g == g && E.pos == nops && like_args(d, pl->c[next])) {
            // the fused pass runs the closures of the finished iteration (set E.par) and this collide (set next)
            int rc = PL_OK;
            if (!same_args(d, pl->c[next])) { rc = flush_pending(); pl->c[next] = d; pl->dirty_c[next] = true; }
            if (pl->dirty_c[E.par] || pl->dirty_aux[E.par] || pl->dirty_c[next] || pl->dirty_aux[next]) {
                // arrays that change from step to step (transient drivers): every step's stores are read later
                if (!rc) rc = flush_pending();
                if (!rc) rc = flush_bindings(pl, E.par);
                if (!rc) rc = flush_bindings(pl, next);
                if (rc) return rc;
                if (pl_plan_advance(pl->p, 1, 0)) return hfail("pl_plan_advance");
            } else if (++E.pending > E.lag) {
                // everything but the two most recent passes: their stores can no longer be observed (see Engine::pending)
                const int n = E.pending - 2;
                E.pending = 2;
                if (pl_plan_advance_observed(pl->p, n, 0, 0)) return hfail("pl_plan_advance");
            }
            ++g_stat[0];
            E.par = next; E.pos = 0;
            return PL_OK;
        }
        int rc = settle();
        if (rc) return rc;
    }
    // LEARN: the previous iteration (collide + what followed) is complete now
    if (E.cur.have_c) {
        if (E.nhist == 2) E.hist[0] = E.hist[1], E.nhist = 1;
        E.hist[E.nhist++] = E.cur;
        E.cur = Iter();
    }
    if (!staged) {
        for (Plan* pl : E.plans) {
            // speculative re-entry: a known plan whose collide arguments match; every following call is still verified
            if (pl->p && pl->f == f && pl->g == g && E.nhist == 0 && pl_lattice_streamed(f) && (!g || pl_lattice_streamed(g))) {
                for (int par = 0; par < 2; ++par) if (same_args(d, pl->c[par])) return enter_replay(pl, par, d);
            }
        }
        for (Plan* pl : E.plans) {
            // the same loop body met again with other arrays (per-step arrays of the transient drivers): re-bind
            if (pl->p && pl->f == f && pl->g == g && E.nhist == 0 && pl_lattice_streamed(f) && (!g || pl_lattice_streamed(g)) && like_args(d, pl->c[0]))
                return enter_replay(pl, 0, d);
        }
        if (E.nhist == 2 && fusable(E.hist[0], E.hist[1]) && E.hist[0].f == f && E.hist[0].g == g && like_args(d, E.hist[0].c) &&
            pl_lattice_streamed(f) && (!g || pl_lattice_streamed(g))) {
            Plan* found = nullptr;
            bool tomb = false;
            for (Plan* pl : E.plans) if (same_plan_shape(pl, E.hist[0], E.hist[1])) { found = pl->p ? pl : nullptr; tomb = !pl->p; break; }
            if (found) adopt_bindings(found, E.hist[0], E.hist[1]);
            if (!found && !tomb) found = build_plan(E.hist[0], E.hist[1]);
            if (found) return enter_replay(found, 0, d);
        }
    }
    ++g_stat[1];
    if (pl_collide(f, g, &d)) return hfail("pl_collide");
    E.cur.have_c = !staged; E.cur.f = f; E.cur.g = g; E.cur.c = d; E.cur.ops.clear();
    return PL_OK;
}

int do_op(const Op& o, bool staged) {
    if (E.active) {
        Plan* pl = E.active;
        const int nops = (int)pl->ops[0].size();
        if (!staged && E.pos < nops && same_shape(o, pl->ops[E.par][E.pos]) && like_aux(o, pl->ops[E.par][E.pos])) {
            if (!same_aux(o, pl->ops[E.par][E.pos])) {
                int rc = flush_pending();
                if (rc) return rc;
                pl->ops[E.par][E.pos].aux = o.aux; pl->dirty_aux[E.par] = true;
            }
            ++E.pos;
            pops_written(o.l);
            return PL_OK;
        }
        int rc = settle();
        if (rc) return rc;
    }
    if (exec_op(o)) return hfail("boundary/stream call");
    if (E.cur.have_c) { if (staged) E.cur = Iter(); else E.cur.ops.push_back(o); }
    return PL_OK;
}

// settle if `l` takes part in the loop being replayed / learned (its populations are about to be observed or replaced)
int quiesce(pl_lattice* l) {
    if (E.active && (E.active->f == l || E.active->g == l)) { int rc = settle(); if (rc) return rc; }
    if (E.cur.have_c && (E.cur.f == l || E.cur.g == l)) { E.cur = Iter(); E.nhist = 0; }
    return PL_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

const char* plh_last_error(void) { return g_herr.c_str(); }

void* plh_alloc(size_t bytes) {
    Block* b = new_block(bytes, BK_ARRAY, nullptr, ST_HOST, PROT_READ | PROT_WRITE);
    return b ? b->base : nullptr;
}
int plh_owns(const void* p) {
    BlocksLock lock_;
    auto it = g_blocks.find((uintptr_t)p);
    return it != g_blocks.end() && it->second.kind == BK_ARRAY;
}
int plh_owns_range(const void* p) { return find_block(p) != nullptr; }
int plh_bc_update_values(pl_bc* bc, const double* v0, const double* v1, const double* v2) {
    int rc = flush_pending();      // they were called with the previous values
    if (rc) return rc;
    return pl_bc_update_values(bc, v0, v1, v2) ? hfail("pl_bc_update_values") : PL_OK;
}
void plh_free(void* p) {
    flush_pending();
    Block* b = nullptr;
    { BlocksLock lock_; auto it = g_blocks.find((uintptr_t)p); if (it != g_blocks.end()) b = &it->second; }
    if (b) drop_block(b);
}

// Make [p, p + bytes) readable (and, with for_write, writable) by ordinary host code, system calls and other libraries — which,
// unlike a load or store of the program itself, cannot be served by the fault handler: an fwrite of a stale array fails with
// EFAULT, an MPI / NCCL staging copy faults inside the CUDA runtime.  Brings the host copy up to date if the device holds the
// current one; for_write also makes the host copy the only current one.  The range may span several blocks; foreign memory is
// left alone.
int plh_host_acquire(const void* p, size_t bytes, int for_write) {
    if (!p || bytes == 0) return PL_OK;
    const char* a = (const char*)p;
    const char* end = a + bytes;
    while (a < end) {
        Block* b = find_block(a);
        if (!b) {      // not ours: skip to the next block that starts inside the range, if any
            const char* next = nullptr;
            { BlocksLock lock_; auto it = g_blocks.upper_bound((uintptr_t)a); if (it != g_blocks.end()) next = it->second.base; }
            if (!next || next >= end) break;
            a = next;
            continue;
        }
        if (b->state == ST_DEVICE) { ++g_stat[4]; fetch(b); }
        if (for_write && b->kind == BK_ARRAY && b->state == ST_SHARED) { flush_pending(); protect(b, PROT_READ | PROT_WRITE); b->state = ST_HOST; }
        if (for_write && b->kind != BK_ARRAY) {
            Views& v = g_views[b->lat];
            for (Block* q : {v.f0, v.f}) if (q->state == ST_SHARED) { protect(q, PROT_READ | PROT_WRITE); q->state = ST_HOST; }
        }
        a = b->base + b->map_bytes;
    }
    return PL_OK;
}

int plh_lattice_attach_views(pl_lattice* l, double** f0, double** f) {
    if (!l || !f0 || !f) { g_herr = "plh_lattice_attach_views: null"; return PL_ERR_ARG; }
    int info[18];
    if (pl_lattice_info(l, info)) return hfail("pl_lattice_info");
    const size_t n = (size_t)info[13], nc = (size_t)info[17];
    Views v;
    // no access until somebody looks: the device populations (all zero after creation) are the truth
    v.f0 = new_block(n*sizeof(double), BK_POP0, l, ST_DEVICE, PROT_NONE);
    v.f = new_block(n*(nc - 1)*sizeof(double), BK_POPF, l, ST_DEVICE, PROT_NONE);
    if (!v.f0 || !v.f) { g_herr = "plh_lattice_attach_views: mmap failed"; return PL_ERR_CUDA; }
    g_views[l] = v;
    *f0 = (double*)v.f0->base; *f = (double*)v.f->base;
    return PL_OK;
}
int plh_lattice_detach(pl_lattice* l) {
    int rc = quiesce(l);
    drop_plans_of(l);
    auto it = g_views.find(l);
    if (it != g_views.end()) {
        drop_block(it->second.f0); drop_block(it->second.f);
        g_views.erase(it);
    }
    return rc;
}

// collide: every pointer of `h` is a HOST pointer; which arrays the model reads / writes follows the reference signatures
int plh_collide(pl_lattice* f, pl_lattice* g, const pl_collide_args* h) {
    HostTimer timer_(T_COLLIDE);
    if (!f || !h) { g_herr = "plh_collide: null"; return PL_ERR_ARG; }
    ++g_tick;       // one loop iteration per collide: the recency window of the state store
    int info[18];
    pl_lattice_info(f, info);
    const size_t n = (size_t)info[13], nc = (size_t)info[17];
    int rc;
    if ((rc = pops_sync_in(f)) || (g && (rc = pops_sync_in(g)))) return rc;
    pl_collide_args d;
    memset(&d, 0, sizeof(d));
    d.model = h->model; d.issave = h->issave; d.viscosity = h->viscosity; d.diffusivity_const = h->diffusivity_const;
    d.gx = h->gx; d.gy = h->gy; d.gz = h->gz; d.tem0 = h->tem0;
    const bool adj = h->model >= PL_ANS_BRINKMAN && h->model <= PL_AAD_NAT_CONV_MASSFLOW, save = h->issave != 0;
    g_staged.clear();
#define RD(field) if ((rc = xlate(h->field, n, true, false, (double**)&d.field))) return rc
#define WR(field) if ((rc = xlate(h->field, n, false, save, (double**)&d.field))) return rc
    RD(alpha); RD(diffusivity); RD(beta); RD(dirx); RD(diry); RD(dirz);
    if (adj) { RD(rho); RD(ux); RD(uy); RD(uz); RD(tem); } else { WR(rho); WR(ux); WR(uy); WR(uz); WR(tem); }
    WR(qx); WR(qy); WR(qz);
    WR(ip); WR(iux); WR(iuy);
    // the 3-D scalar tail of AAD::MacroBrinkmanCollideForceConvection does not store _iuz (adjointadvection_avx.h:725-735): a partial write
    if (h->model == PL_AAD_FORCE_CONV) { if ((rc = xlate(h->iuz, n, true, save, (double**)&d.iuz))) return rc; } else WR(iuz);
    WR(imx); WR(imy); WR(imz); WR(item); WR(iqx); WR(iqy); WR(iqz);
#undef RD
#undef WR
    tag_snapshot(h->snapshot, g ? g : f, n*nc);
    if ((rc = xlate(h->snapshot, n*nc, false, save, &d.snapshot))) return rc;
    const bool staged = !g_staged.empty();
    rc = do_collide(f, g, d, staged);
    int rc2 = unstage();
    return rc ? rc : rc2;
}

int plh_stream(pl_lattice* l, int inverse) {
    HostTimer timer_(T_STREAM);
    if (!l) { g_herr = "plh_stream: null"; return PL_ERR_ARG; }
    int rc = pops_sync_in(l);
    if (rc) return rc;
    Op o; o.kind = OP_STREAM; o.l = l; o.inverse = inverse ? 1 : 0;
    return do_op(o, false);
}
int plh_smooth_corner(pl_lattice* l) {
    HostTimer timer_(T_SMOOTH);
    if (!l) { g_herr = "plh_smooth_corner: null"; return PL_ERR_ARG; }
    int rc = pops_sync_in(l);
    if (rc) return rc;
    Op o; o.kind = OP_SMOOTH; o.l = l;
    return do_op(o, false);
}
int plh_smooth_corner_at(pl_lattice* l, int i, int j, int k, int dx, int dy, int dz) {
    HostTimer timer_(T_SMOOTH);
    if (!l) { g_herr = "plh_smooth_corner_at: null"; return PL_ERR_ARG; }
    int rc = pops_sync_in(l);
    if (rc) return rc;
    Op o; o.kind = OP_SMOOTH_AT; o.l = l;
    o.at[0] = i; o.at[1] = j; o.at[2] = k; o.at[3] = dx; o.at[4] = dy; o.at[5] = dz;
    return do_op(o, false);
}
int plh_bc(pl_lattice* l, pl_lattice* other, const pl_bc* bc, const pl_bc_aux* h) {
    HostTimer timer_(T_BC);
    if (!l || !bc) { g_herr = "plh_bc: null"; return PL_ERR_ARG; }
    if (pl_bc_is_empty(bc)) return PL_OK;      // the reference's `if (0 <= i && i < nx)` guard: nothing to do on this rank
    int info[18];
    pl_lattice_info(l, info);
    const size_t n = (size_t)info[13];
    int rc;
    if ((rc = pops_sync_in(l)) || (other && (rc = pops_sync_in(other)))) return rc;
    Op o; o.kind = OP_BC; o.l = l; o.other = other; o.bc = bc; o.has_aux = h != nullptr;
    memset(&o.aux, 0, sizeof(o.aux));
    g_staged.clear();
    if (h) {
        o.aux.diffusivity_const = h->diffusivity_const; o.aux.eps = h->eps;
        if ((rc = xlate(h->rho, n, true, false, (double**)&o.aux.rho)) || (rc = xlate(h->ux, n, true, false, (double**)&o.aux.ux)) ||
            (rc = xlate(h->uy, n, true, false, (double**)&o.aux.uy)) || (rc = xlate(h->uz, n, true, false, (double**)&o.aux.uz)) ||
            (rc = xlate(h->tem, n, true, false, (double**)&o.aux.tem)) || (rc = xlate(h->diffusivity, n, true, false, (double**)&o.aux.diffusivity)))
            return rc;
    }
    const bool staged = !g_staged.empty();
    rc = do_op(o, staged);
    int rc2 = unstage();
    return rc ? rc : rc2;
}

int plh_initial_condition(pl_lattice* l, int family, const double* const* h, int na) {
    HostTimer timer_(T_INIT);
    if (!l || !h) { g_herr = "plh_initial_condition: null"; return PL_ERR_ARG; }
    int info[18];
    pl_lattice_info(l, info);
    const size_t n = (size_t)info[13];
    int rc = flush_pending();
    if (!rc) rc = quiesce(l);
    if (rc) return rc;
    const double* d[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    g_staged.clear();
    for (int k = 0; k < na && k < 8; ++k) if ((rc = xlate(h[k], n, true, false, (double**)&d[k]))) return rc;
    pops_written(l);
    ++g_stat[1];
    rc = pl_initial_condition(l, family, d, na) ? hfail("pl_initial_condition") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}

int plh_residual(const double* ux, const double* uy, const double* uz, const double* uxp, const double* uyp, const double* uzp, size_t n, double* out) {
    HostTimer timer_(T_RESIDUAL);
    const double* h[6] = {ux, uy, uz, uxp, uyp, uzp};
    double* d[6];
    int rc;
    if ((rc = flush_pending())) return rc;
    g_staged.clear();
    for (int k = 0; k < 6; ++k) if ((rc = xlate(h[k], n, true, false, &d[k]))) return rc;
    rc = pl_residual(d[0], d[1], d[2], d[3], d[4], d[5], n, out) ? hfail("pl_residual") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}
int plh_normalize(double* v, size_t n) {
    double* d;
    int rc;
    if ((rc = flush_pending())) return rc;
    g_staged.clear();
    if ((rc = xlate(v, n, true, true, &d))) return rc;
    rc = pl_normalize(d, n) ? hfail("pl_normalize") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}
int plh_sensitivity(pl_lattice* l, const pl_sens_args* h) {
    HostTimer timer_(T_SENS);
    if (!l || !h) { g_herr = "plh_sensitivity: null"; return PL_ERR_ARG; }
    int info[18];
    pl_lattice_info(l, info);
    const size_t n = (size_t)info[13], nc = (size_t)info[17];
    pl_sens_args d;
    memset(&d, 0, sizeof(d));
    d.kind = h->kind;
    int rc;
    if ((rc = flush_pending())) return rc;
    g_staged.clear();
    tag_snapshot(h->gsnap, l, n*nc); tag_snapshot(h->igsnap, l, n*nc);
    if ((rc = xlate(h->dfds, n, true, true, &d.dfds))) return rc;
#define RD(field, len) if ((rc = xlate(h->field, len, true, false, (double**)&d.field))) return rc
    RD(ux, n); RD(uy, n); RD(uz, n); RD(imx, n); RD(imy, n); RD(imz, n); RD(dads, n); RD(tem, n); RD(item, n); RD(iqx, n); RD(iqy, n); RD(iqz, n);
    RD(gsnap, n*nc); RD(igsnap, n*nc); RD(diffusivity, n); RD(dkds, n); RD(dbds, n);
#undef RD
    rc = pl_sensitivity(l, &d) ? hfail("pl_sensitivity") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}
int plh_sensitivity_heat_source(pl_lattice* l, const pl_bc* plane, double* dfds, const double* ux, const double* uy, const double* uz,
                                const double* igsnap, const double* diffusivity, const double* dkds) {
    HostTimer timer_(T_SENS_HS);
    if (!l || !plane) { g_herr = "plh_sensitivity_heat_source: null"; return PL_ERR_ARG; }
    if (pl_bc_is_empty(plane)) return PL_OK;
    int info[18];
    pl_lattice_info(l, info);
    const size_t n = (size_t)info[13], nc = (size_t)info[17];
    double *d_dfds, *d_ux, *d_uy, *d_uz, *d_ig, *d_k, *d_dk;
    int rc;
    if ((rc = flush_pending())) return rc;
    tag_snapshot(igsnap, l, n*nc);
    g_staged.clear();
    if ((rc = xlate(dfds, n, true, true, &d_dfds)) || (rc = xlate(ux, n, true, false, &d_ux)) || (rc = xlate(uy, n, true, false, &d_uy)) ||
        (rc = xlate(uz, n, true, false, &d_uz)) || (rc = xlate(igsnap, n*nc, true, false, &d_ig)) || (rc = xlate(diffusivity, n, true, false, &d_k)) ||
        (rc = xlate(dkds, n, true, false, &d_dk)))
        return rc;
    rc = pl_sensitivity_heat_source(l, plane, d_dfds, d_ux, d_uy, d_uz, d_ig, d_k, d_dk) ? hfail("pl_sensitivity_heat_source") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}

int plh_filter_apply(pl_filter* f, int mode, double beta, const double* v, const double* dfdrho, double* out, size_t n) {
    HostTimer timer_(T_FILTER);
    if (!f) { g_herr = "plh_filter_apply: null"; return PL_ERR_ARG; }
    double *dv, *dd, *dout;
    int rc;
    if ((rc = flush_pending())) return rc;
    g_staged.clear();
    if ((rc = xlate(v, n, true, false, &dv)) || (rc = xlate(dfdrho, n, true, false, &dd)) || (rc = xlate(out, n, false, true, &dout))) return rc;
    rc = pl_filter_apply(f, mode, beta, dv, dd, dout) ? hfail("pl_filter_apply") : PL_OK;
    int rc2 = unstage();
    return rc ? rc : rc2;
}

int plh_sync(void) {
    HostTimer timer_(T_SYNC);
    int rc = settle();
    pl_synchronize();
    return rc;
}
int plh_stats(uint64_t* out8) {
    if (!out8) return PL_ERR_ARG;
    memcpy(out8, g_stat, sizeof(g_stat));
    return PL_OK;
}
int plh_store_stats(uint64_t* out4) {
    if (!out4) return PL_ERR_ARG;
    memcpy(out4, g_store, sizeof(g_store));
    return PL_OK;
}

}  // extern "C"
