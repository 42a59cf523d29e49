// CUDA kernels of the sweep: population streaming, the fused stream+collide pass, in-place collide,
// bounce-back / specular planes, SmoothCorner, layout conversion.  fp64 SoA: population c of site idx lives at
// base[c*pitch + idx], idx = i + nx*(j + ny*k)  (same site order as the reference, d3q15.h:136-141).
#pragma once
#include "lbm_sens.cuh"
#include "lbm_halo.cuh"

namespace plb {

// signed index deltas to the periodic neighbours of one site (reference Index(): d3q15.h:136-141)
struct Nbr {
    int m[3], p[3];   // delta to coordinate-1 / coordinate+1 along x,y,z (a block holds fewer than 2^31 sites: 32-bit index arithmetic
                      // keeps the thirty load / store addresses of the fused pass cheap to recompute instead of living in registers)
};
PL_D void decompose(const Geom& G, long long idx, int& i, int& j, int& k) {
    unsigned u = (unsigned)idx, nxy = (unsigned)(G.nx*G.ny);
    unsigned kk = u/nxy, r = u - kk*nxy, jj = r/(unsigned)G.nx;
    k = (int)kk; j = (int)jj; i = (int)(r - jj*(unsigned)G.nx);
}
PL_D Nbr neighbours(const Geom& G, int i, int j, int k) {
    Nbr n;
    const int sx = 1, sy = G.nx, sz = G.nx*G.ny;
    n.m[0] = i == 0 ? (G.nx - 1)*sx : -sx;  n.p[0] = i == G.nx - 1 ? -(G.nx - 1)*sx : sx;
    n.m[1] = j == 0 ? (G.ny - 1)*sy : -sy;  n.p[1] = j == G.ny - 1 ? -(G.ny - 1)*sy : sy;
    n.m[2] = k == 0 ? (G.nz - 1)*sz : -sz;  n.p[2] = k == G.nz - 1 ? -(G.nz - 1)*sz : sz;
    return n;
}
// offset of the site population c is pulled from: x - c (Stream) or x + c (iStream, n has m/p swapped by the caller)
template <int D, int c> PL_D int pull_offset(const Nbr& n) {
    constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
    int o = 0;
    if constexpr (X > 0) o += n.m[0]; else if constexpr (X < 0) o += n.p[0];
    if constexpr (Y > 0) o += n.m[1]; else if constexpr (Y < 0) o += n.p[1];
    if constexpr (Z > 0) o += n.m[2]; else if constexpr (Z < 0) o += n.p[2];
    return o;
}
PL_D void orient(Nbr& n, int inverse) {
    if (inverse) {
        for (int d = 0; d < 3; ++d) { int t = n.m[d]; n.m[d] = n.p[d]; n.p[d] = t; }
    }
}

template <int D> PL_D void pull(double (&f)[LT<D>::nc], const double* __restrict__ src, size_t pitch, long long idx, const Nbr& n) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        f[c] = __ldg(src + (size_t)c*pitch + (size_t)(unsigned)((int)idx + pull_offset<D, c>(n)));
    });
}
// the same for a block-decomposed lattice: sources beyond a decomposed block face come from the receive buffers
template <int D> PL_D void pull_h(double (&f)[LT<D>::nc], const double* __restrict__ src, size_t pitch, long long idx, int i, int j, int k,
                                  const Geom& G, const Nbr& n, const HaloView& H, int inverse) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        f[c] = pull_halo<D, c>(src, idx, i, j, k, G, n, (size_t)c*pitch + (size_t)(unsigned)((int)idx + pull_offset<D, c>(n)), H, inverse);
    });
}
template <int D> PL_D void load_site(double (&f)[LT<D>::nc], const double* __restrict__ src, size_t pitch, long long idx) {
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = src[(size_t)c*pitch + (size_t)idx]; });
}
template <int D> PL_D void store_site(const double (&f)[LT<D>::nc], double* __restrict__ dst, size_t pitch, long long idx) {
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; dst[(size_t)c*pitch + (size_t)idx] = f[c]; });
}

// ---------------------------------------------------------------------------------------------------------
// Addressing of one fused stream+collide pass.  The reference keeps two population buffers per lattice (f and the hidden fnext,
// d3q15.h:41-45, 238); the fused pass needs only ONE, AA-pattern style: every thread writes exactly the locations it has read.
//   natural layout N  : location (c, x) holds the post-collision population c of site x            (= what every other
//                       function of this library reads and writes)
//   streamed layout P : location (opp(c), x + s*c) holds the post-collision population c of site x, s = +1 Stream / -1 iStream;
//                       i.e. location (opp(c), y) holds the population c ARRIVING at y — Stream() has happened, slots swapped
//   PASS_GATHER  N -> P : population c is pulled from (c, x - s*c) — the classic pull — and, after closures and collide, the new
//                         population c goes to (opp(c), x + s*c): the very location the thread pulled population opp(c) from
//   PASS_LOCAL   P -> N : population c is read from (opp(c), x), the new one goes to (c, x): no shifted access at all
//   PASS_COPY    N -> N': the two-buffer pass (pull from the source buffer, store to the destination buffer; PANSLBM_INPLACE=0)
// Each location belongs to exactly one thread per pass (x - s*c is a bijection on the periodic block), so the interior kernel,
// the boundary pass and the tube kernel can still run side by side.
enum { PASS_COPY = 0, PASS_GATHER = 1, PASS_LOCAL = 2 };
template <int D, int MODE, int c> PL_D size_t read_loc(size_t pitch, long long idx, const Nbr& n) {
    if constexpr (MODE == PASS_LOCAL) return (size_t)LT<D>::opp(c)*pitch + (size_t)(unsigned)(int)idx;
    else return (size_t)c*pitch + (size_t)(unsigned)((int)idx + pull_offset<D, c>(n));
}
template <int D, int MODE, int c> PL_D size_t write_loc(size_t pitch, long long idx, const Nbr& n) {
    if constexpr (MODE == PASS_GATHER) return (size_t)LT<D>::opp(c)*pitch + (size_t)(unsigned)((int)idx + pull_offset<D, LT<D>::opp(c)>(n));
    else return (size_t)c*pitch + (size_t)(unsigned)(int)idx;
}
template <int D, int MODE> PL_D void pass_load(double (&f)[LT<D>::nc], const double* __restrict__ src, size_t pitch, long long idx, const Nbr& n) {
    // the in-place passes write what they read: no read-only (non-coherent) loads there
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        if constexpr (MODE == PASS_COPY) f[c] = __ldg(src + read_loc<D, MODE, c>(pitch, idx, n));
        else f[c] = src[read_loc<D, MODE, c>(pitch, idx, n)];
    });
}
template <int D, int MODE> PL_D void pass_store(const double (&f)[LT<D>::nc], double* __restrict__ dst, size_t pitch, long long idx, const Nbr& n) {
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; dst[write_loc<D, MODE, c>(pitch, idx, n)] = f[c]; });
}
// the same load for a block-decomposed lattice: populations whose source site lies beyond a decomposed block face come from the
// receive buffers whatever the layout (what the neighbour packed is the population itself, not a location)
template <int D, int MODE> PL_D void pass_load_h(double (&f)[LT<D>::nc], const double* __restrict__ src, size_t pitch, long long idx, int i, int j, int k,
                                                const Geom& G, const Nbr& n, const HaloView& H, int inverse) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        f[c] = pull_halo<D, c>(src, idx, i, j, k, G, n, read_loc<D, MODE, c>(pitch, idx, n), H, inverse);
    });
}

// ---------------------------------------------------------------------------------------------------------
// Stream()/iStream(): dst(x,c) = src(x -/+ c, c) for every site (d3q15.h:601-616, 964-979)
template <int D, bool HALO>
__global__ void __launch_bounds__(256) k_stream(Geom G, const double* __restrict__ src, double* __restrict__ dst, int inverse, HaloView H) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    int i, j, k;
    decompose(G, idx, i, j, k);
    Nbr n = neighbours(G, i, j, k);
    orient(n, inverse);
    double f[LT<D>::nc];
    if constexpr (HALO) pull_h<D>(f, src, G.pitch, idx, i, j, k, G, n, H, inverse);
    else pull<D>(f, src, G.pitch, idx, n);
    store_site<D>(f, dst, G.pitch, idx);
}
// streamed layout P -> natural layout N (see the pass modes above): dst(c, x) = src(opp(c), x + s*c), the post-collision
// populations back at their own sites.  Only when something other than the next fused pass wants to look at the lattice.
template <int D>
__global__ void __launch_bounds__(256) k_unstream(Geom G, const double* __restrict__ src, double* __restrict__ dst, int inverse) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    int i, j, k;
    decompose(G, idx, i, j, k);
    Nbr n = neighbours(G, i, j, k);
    orient(n, inverse);
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        dst[(size_t)c*G.pitch + (size_t)idx] = src[write_loc<D, PASS_GATHER, c>(G.pitch, idx, n)];
    });
}
// ---------------------------------------------------------------------------------------------------------
// In-place Macro*Collide* over all sites (list == nullptr) or over a site list.  Tail sites
// (idx >= 4*(nxyz/4)) take the scalar operation order exactly as the reference does (navierstokes_avx.h:180-200).
template <int D, int M>
__global__ void __launch_bounds__(256) k_collide(Geom G, double* __restrict__ fb, double* __restrict__ gb, CollideParams P,
                                                 const int* __restrict__ list, long long count) {
    constexpr unsigned FL = ModelFlags<M>::v;
    long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= count) return;
    long long idx = list ? (long long)list[t] : t;
    double f[LT<D>::nc], g[LT<D>::nc];
    load_site<D>(f, fb, G.pitch, idx);
    if constexpr ((FL & F_G) != 0) load_site<D>(g, gb, G.pitch, idx);
    if (idx < G.npacked) collide_site<D, FL, false>(f, g, P, (size_t)idx, P.issave != 0);
    else collide_site<D, FL, true>(f, g, P, (size_t)idx, P.issave != 0);
    store_site<D>(f, fb, G.pitch, idx);
    if constexpr ((FL & F_G) != 0) store_site<D>(g, gb, G.pitch, idx);
}

// ---------------------------------------------------------------------------------------------------------
// Plane descriptor shared by every closure kernel.  n1 x n2 local plane sites, site index = base + a*s1 + b*s2.
struct Plane {
    int axis, dir;          // normal axis 0/1/2, outward direction -1/+1
    int n1, n2;             // extents of the two in-plane axes (lower axis first)
    long long base, s1, s2; // index of plane site (0,0) and the strides of the in-plane axes
};

// Arguments of one plane closure launch (k_closure) / one entry of a plan's closure program (k_shell).
struct ClosureArgs {
    int type;                      // BC_* (PL_BC_* of the C-ABI)
    int on_g;                      // program entries only: 0 = acts on the flow lattice, 1 = on the thermal lattice
    int loc;                       // program entries only: local coordinate of the plane along its axis
    Plane pl;
    const uint8_t* mask;           // per plane site
    const double *v0, *v1, *v2;    // per plane site values
    const double *rho, *ux, *uy, *uz, *tem, *kappa;   // per lattice site fields (device, may be null)
    double kconst, eps;
};
PL_D SiteVals site_vals(const ClosureArgs& A, int t, long long idx) {
    SiteVals V;
    V.v0 = A.v0 ? A.v0[t] : 0.0; V.v1 = A.v1 ? A.v1[t] : 0.0; V.v2 = A.v2 ? A.v2[t] : 0.0;
    V.rho = A.rho ? A.rho[idx] : 0.0; V.ux = A.ux ? A.ux[idx] : 0.0; V.uy = A.uy ? A.uy[idx] : 0.0; V.uz = A.uz ? A.uz[idx] : 0.0;
    V.tem = A.tem ? A.tem[idx] : 0.0;
    V.kappa = A.kappa ? A.kappa[idx] : A.kconst;
    V.eps = A.eps;
    return V;
}

// One closure on one plane, in place (the call-by-call path: P::BoundaryConditionAlong*, NS::BoundaryConditionSetU, ...).
// pb = populations of the lattice the closure acts on, qb = the other lattice (BC_AAD_ISET_RHO only, else null).
template <int D>
__global__ void __launch_bounds__(128) k_closure(Geom G, double* __restrict__ pb, const double* __restrict__ qb, ClosureArgs A) {
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= A.pl.n1*A.pl.n2) return;
    const int m = A.mask[t];
    if (!m) return;
    int a = t%A.pl.n1, b = t/A.pl.n1;
    long long idx = A.pl.base + a*A.pl.s1 + b*A.pl.s2;
    double p[LT<D>::nc], q[LT<D>::nc];
    #pragma unroll
    for (int c = 0; c < LT<D>::nc; ++c) { p[c] = pb[(size_t)c*G.pitch + idx]; q[c] = qb ? qb[(size_t)c*G.pitch + idx] : 0.0; }
    apply_closure<D>(A.type, A.pl.axis, A.pl.dir, m, p, q, site_vals(A, t, idx));
    #pragma unroll
    for (int c = 1; c < LT<D>::nc; ++c) pb[(size_t)c*G.pitch + idx] = p[c];
}

// heat-source boundary term of AAD::SensitivityTemperatureAtHeatSource on one plane (adjointadvection_avx.h:16-185);
// A.v0 = qn, A.ux/uy/uz/kappa = fields; igsnap = adjoint thermal snapshot (SoA [c][nxyz]).
template <int D>
__global__ void __launch_bounds__(128) k_sens_heat_source(Geom G, ClosureArgs A, const double* __restrict__ igsnap, const double* __restrict__ dkds,
                                                          double* __restrict__ dfds) {
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= A.pl.n1*A.pl.n2) return;
    if (!A.mask[t]) return;
    int a = t%A.pl.n1, b = t/A.pl.n1;
    long long idx = A.pl.base + a*A.pl.s1 + b*A.pl.s2;
    double ig[LT<D>::nc];
    #pragma unroll
    for (int c = 0; c < LT<D>::nc; ++c) ig[c] = igsnap[(size_t)c*(size_t)G.nxyz + idx];
    dfds[idx] = dfds[idx] + sens_heat_source_term<D>(ig, A.pl.axis, A.pl.dir, site_vals(A, t, idx), dkds[idx]);
}

// ---------------------------------------------------------------------------------------------------------
// Per-coordinate plane words of a plan (one 64-bit word per local x / y / z coordinate):
//   bits 0..59: entry e of the closure program acts on this plane
//   bit 60    : (x only) the closures of this x boundary plane run AHEAD of the pass (k_xclose) on the compact wall buffers; the
//               interior kernel takes the plane's sites as ordinary ones and picks the rebuilt populations up from there (XWall)
//   bit 61    : the plane is a face of this rank's block along a decomposed axis: its sites pull from the halo receive
//               buffers and are the first to finish, so that the next exchange overlaps the interior kernel
//   bit 62    : (x only) the coordinate shares an aligned group of 4 sites (one 32-byte sector; PANSLBM_XSLAB) with an x plane
//               the boundary pass owns (a decomposed block face; any x closure plane with PANSLBM_XINLINE=0): the boundary
//               pass takes the whole group.  Splitting a sector between the two kernels, or between
//               warps of one, costs far more than the extra sites do (measured: profiles/r01_tuning.md)
//   bit 63    : the plane is a global boundary plane or next to one and the plan has SmoothCorner; sites with two such
//               coordinates form the edge "tubes" SmoothCorner reads and writes.
struct ShellMask {
    const unsigned long long *x, *y, *z;
    int prefetch;      // closure inputs ahead of the pull (PANSLBM_PREFETCH): 0 = off, 1 = prefetch.global.L2, 2 = prefetch.global.L1, 3 = plain loads
    int l2_ahead;      // interior kernel: sites ahead of its own for which a thread asks L2 to fetch the lines it will need (0 = off)
};
constexpr unsigned long long TUBE_BIT = 1ull << 63, SLAB_BIT = 1ull << 62, HALO_BIT = 1ull << 61, GHOST_BIT = 1ull << 60,
                             ENTRY_BITS = ~(TUBE_BIT | SLAB_BIT | HALO_BIT | GHOST_BIT);
constexpr int MAX_PROGRAM = 60;

// Compact staging of the x boundary planes whose closures run ahead of the pass.  x planes are the expensive ones for any
// boundary treatment: consecutive plane sites lie nx*8 bytes apart, so every 8-byte access of a per-site kernel moves a whole
// 32-byte sector (ncu, round 1: 3.9 DRAM sectors per requested sector in the old k_xclose).  Instead every kernel that produces
// post-collision populations also drops the ones that Stream() will carry ONTO such a plane into a compact buffer, at the plane
// index of the site they arrive at:
//     out[side][c][t],  t = j + ny*k of the TARGET site, side 0: x = 0, side 1: x = nx-1
// (written by the threads with x in {0, 1} / {nx-2, nx-1}: a handful of extra stores in the warps that hold such a lane).
// k_xclose of the next pass reads `in` = the `out` of this one — fully coalesced, no shifts: the values already sit where
// they arrive — runs the closure program of the plane site and leaves the rebuilt populations (the ones Stream() would have
// pulled through the periodic wrap, whose wrapped-around value is dead) in res[side][c][t]; the interior kernel of that pass
// overwrites what it pulled through the wrap with them.  Nothing strided is left on the critical path.
struct XWall {
    double *out_f, *out_g;            // [2][nc][np]: written by this pass for the next one
    const double *res_f, *res_g;      // [2][nc][np]: rebuilt populations for this pass (k_xclose)
    int np;                           // ny*nz
    int on[2];                        // the plane x = 0 / x = nx-1 is handled this way
};
// after the collide of site (i,j,k): the populations that stream onto a compact x plane — also through the periodic wrap: a
// population no closure of the plane rebuilds keeps the wrapped-around value, exactly as after Stream() (a plane whose closure mask
// covers only part of it, a lattice without closures on a plane the other lattice has them on).
// XV = the x component of the populations that take the step (one call per step direction: straight-line code for its five
// populations, no per-population test — at nx = 81 almost every other warp holds a lane that comes through here)
template <int D, bool HASG, int XV>
PL_D void wall_scatter_set(const XWall& W, const double (&f)[LT<D>::nc], const double (&g)[LT<D>::nc], const Geom& G, int side, int j, int k, int inverse) {
    constexpr int NC = LT<D>::nc;
    sfor<0, NC>([&](auto C) {
        constexpr int c = decltype(C)::value;
        constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
        if constexpr (X == XV) {
            int jt = j, kt = k;
            if constexpr (Y != 0) { jt = j + (inverse ? -Y : Y); jt = jt < 0 ? G.ny - 1 : (jt >= G.ny ? 0 : jt); }
            if constexpr (Z != 0) { kt = k + (inverse ? -Z : Z); kt = kt < 0 ? G.nz - 1 : (kt >= G.nz ? 0 : kt); }
            const size_t o = (size_t)(side*NC + c)*W.np + (size_t)(jt + G.ny*kt);
            W.out_f[o] = f[c];
            if constexpr (HASG) W.out_g[o] = g[c];
        }
    });
}
template <int D, bool HASG>
PL_D void wall_scatter(const XWall& W, const double (&f)[LT<D>::nc], const double (&g)[LT<D>::nc], const Geom& G, int i, int j, int k, int inverse) {
    if (W.out_f == nullptr || (i > 1 && i < G.nx - 2)) return;
    // (side, di): the plane this site feeds and the x step s*c_x that lands on it (nx >= 4: the four cases are distinct)
    auto emit = [&](int side, int di) {
        if (!W.on[side]) return;
        const int xv = inverse ? -di : di;
        if (xv == 0) wall_scatter_set<D, HASG, 0>(W, f, g, G, side, j, k, inverse);
        else if (xv > 0) wall_scatter_set<D, HASG, 1>(W, f, g, G, side, j, k, inverse);
        else wall_scatter_set<D, HASG, -1>(W, f, g, G, side, j, k, inverse);
    };
    if (i == 0) { emit(0, 0); emit(1, -1); }                 // its own plane; the far plane through the periodic wrap
    else if (i == 1) emit(0, -1);
    else if (i == G.nx - 2) emit(1, 1);
    else { emit(1, 0); emit(0, 1); }
}
// the load of the interior kernel: a site on a compact x plane (wside = 0: x = 0, 1: x = nx-1; -1: any other site) takes the
// populations that would come through the periodic wrap from what the closures of the plane made of them (XWall::res) instead —
// chosen by ADDRESS, one load per population either way (overwriting them after the pull would make the warp wait for its loads twice)
template <int D, int MODE, bool HASG>
PL_D void pass_load_wall(double (&f)[LT<D>::nc], double (&g)[LT<D>::nc], const double* fs, const double* gs, size_t pitch, long long idx, const Nbr& n,
                         const XWall& W, int wside, size_t wt, int inverse) {
    constexpr int NC = LT<D>::nc;
    sfor<0, NC>([&](auto C) {
        constexpr int c = decltype(C)::value;
        constexpr int X = LT<D>::cx(c);
        const size_t loc = read_loc<D, MODE, c>(pitch, idx, n);
        const double *pf = fs + loc, *pg = HASG ? gs + loc : nullptr;
        if constexpr (X != 0) {
            if (wside >= 0 && (inverse ? -X : X) == (wside ? -1 : 1)) {
                const size_t o = (size_t)(wside*NC + c)*W.np + wt;
                pf = W.res_f + o;
                if constexpr (HASG) pg = W.res_g + o;
            }
        }
        if constexpr (MODE == PASS_COPY) { f[c] = __ldg(pf); if constexpr (HASG) g[c] = __ldg(pg); }
        else { f[c] = *pf; if constexpr (HASG) g[c] = *pg; }
    });
}

PL_D bool in_tube(unsigned long long wx, unsigned long long wy, unsigned long long wz) { return (wx >> 63) + (wy >> 63) + (wz >> 63) >= 2ull; }

// The closure program is a chain of dependent loads (program entry -> mask -> plane values / saved fields of the site), one
// DRAM round trip each, behind the pull.  Issuing prefetches for all of them before the pull turns the chain into cache hits.
PL_D void touch(const void* p, int level) {
    if (level == 3) {      // a real load whose value nobody reads: the line is in L1 when the closure asks for it
        unsigned long long v;
        asm volatile("ld.global.nc.L1::evict_last.b64 %0, [%1];" : "=l"(v) : "l"((unsigned long long)p & ~7ull));
    } else if (level == 2) asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
    else asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}
template <unsigned FL> PL_D void prefetch_collide(const CollideParams& P, long long idx, int level) {
    if constexpr ((FL & F_KFIELD) != 0) touch(P.kappa + idx, level);
    if constexpr ((FL & F_BRINK) != 0 || ((FL & F_ADJ) != 0 && (FL & F_G) != 0)) touch(P.alpha + idx, level);
    if constexpr ((FL & F_HEATEX) != 0) touch(P.beta + idx, level);
    if constexpr ((FL & F_ADJ) != 0) {
        touch(P.rho + idx, level); touch(P.ux + idx, level); touch(P.uy + idx, level);
        if (P.uz) touch(P.uz + idx, level);
        if constexpr ((FL & F_G) != 0) touch(P.tem + idx, level);
    }
}
PL_D void prefetch_program(const ClosureArgs* __restrict__ prog, unsigned long long entries, int i, int j, int k, long long idx, int level) {
    const int co[3] = {i, j, k};
    while (entries) {
        const int e = __ffsll((long long)entries) - 1;
        entries &= entries - 1;
        const ClosureArgs& A = prog[e];
        const int axis = A.pl.axis;
        const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        const int pt = co[a1] + A.pl.n1*co[a2];
        touch(A.mask + pt, level);
        if (A.v0) touch(A.v0 + pt, level);
        if (A.v1) touch(A.v1 + pt, level);
        if (A.v2) touch(A.v2 + pt, level);
        if (A.rho) touch(A.rho + idx, level);
        if (A.ux) touch(A.ux + idx, level);
        if (A.uy) touch(A.uy + idx, level);
        if (A.uz) touch(A.uz + idx, level);
        if (A.tem) touch(A.tem + idx, level);
        if (A.kappa) touch(A.kappa + idx, level);
    }
}

// A plan's closure program on one site: the recorded closures whose plane passes through the site (bits of `entries`), in
// call order, each applied where its baked mask is set.  Deliberately NOT inlined and fed through an addressable copy of
// the populations: the closures index p[] dynamically, which would otherwise drag the populations of the collide into
// local memory.
template <int D, bool HASG>
__device__ __noinline__ void run_program(double* __restrict__ fg, const ClosureArgs* __restrict__ prog, unsigned long long entries,
                                         int i, int j, int k, long long idx) {
    constexpr int NC = LT<D>::nc;
    const int co[3] = {i, j, k};
    while (entries) {
        const int e = __ffsll((long long)entries) - 1;
        entries &= entries - 1;
        const ClosureArgs& A = prog[e];
        const int axis = A.pl.axis;
        const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        const int pt = co[a1] + A.pl.n1*co[a2];
        const int m = A.mask[pt];
        if (!m) continue;
        // one call site (the specialised closure bodies are instantiated once per face inside it): select the lattice by pointer
        const bool og = HASG && A.on_g;
        double (&p)[NC] = *reinterpret_cast<double (*)[NC]>(og ? fg + NC : fg);
        const double (&q)[NC] = *reinterpret_cast<const double (*)[NC]>(og ? fg : fg + NC);
        apply_closure<D>(A.type, axis, A.pl.dir, m, p, q, site_vals(A, pt, idx));
    }
}
template <int D, bool HASG>
PL_D void boundary_path(double (&f)[LT<D>::nc], double (&g)[LT<D>::nc], const ClosureArgs* __restrict__ prog, unsigned long long entries,
                        int i, int j, int k, long long idx) {
    constexpr int NC = LT<D>::nc;
    double t[2*NC];
    sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; t[c] = f[c]; t[NC + c] = HASG ? g[c] : 0.0; });
    run_program<D, HASG>(t, prog, entries, i, j, k, idx);
    sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = t[c]; if constexpr (HASG) g[c] = t[NC + c]; });
}

// The boundary pass keeps the populations of a site in shared memory while the closure program runs on them: one column of
// the CTA's [2*nc][SHELL_THREADS] tile per thread (stride = the CTA width, conflict-free for 8-byte accesses).  Thread-local
// scratch for the same purpose lives in L1-backed local memory, which the streaming loads of the pass keep evicting (ncu:
// half of the local loads missed L1, profiles/r01_tuning.md).
constexpr int SHELL_THREADS = 128;
struct SPop {
    double* b;
    PL_D double& operator[](int c) const { return b[c*SHELL_THREADS]; }
};
template <int D, bool HASG>
__device__ __noinline__ void run_program_sh(double* __restrict__ col, const ClosureArgs* __restrict__ prog, unsigned long long entries,
                                            int i, int j, int k, long long idx) {
    constexpr int NC = LT<D>::nc;
    const int co[3] = {i, j, k};
    while (entries) {
        const int e = __ffsll((long long)entries) - 1;
        entries &= entries - 1;
        const ClosureArgs& A = prog[e];
        const int axis = A.pl.axis;
        const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
        const int pt = co[a1] + A.pl.n1*co[a2];
        const int m = A.mask[pt];
        if (!m) continue;
        const bool og = HASG && A.on_g;
        SPop p{og ? col + NC*SHELL_THREADS : col};
        const SPop q{og ? col : col + NC*SHELL_THREADS};
        apply_closure<D>(A.type, axis, A.pl.dir, m, p, q, site_vals(A, pt, idx));
    }
}
template <int D, bool HASG>
PL_D void boundary_path_sh(double (&f)[LT<D>::nc], double (&g)[LT<D>::nc], double* __restrict__ col, const ClosureArgs* __restrict__ prog,
                           unsigned long long entries, int i, int j, int k, long long idx) {
    constexpr int NC = LT<D>::nc;
    sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; col[c*SHELL_THREADS] = f[c]; if constexpr (HASG) col[(NC + c)*SHELL_THREADS] = g[c]; });
    run_program_sh<D, HASG>(col, prog, entries, i, j, k, idx);
    sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = col[c*SHELL_THREADS]; if constexpr (HASG) g[c] = col[(NC + c)*SHELL_THREADS]; });
}

// The hot kernel: one fused Stream + Macro*Collide* pass, source buffer -> destination buffer, for every packed site that
// lies on no y/z closure plane, in no x group of the boundary pass and in no SmoothCorner tube.  Each population is read once
// and written once.  Sites of an x closure plane that the plan left to this kernel (no SLAB bit) are either ordinary sites
// here (prog == nullptr: k_xclose has already put the closure results where this kernel pulls from) or run the closure
// program between pull and collide (PANSLBM_XINLINE: one lane of the warp diverges; measured slower, profiles/r01_tuning.md).
// CTA shape of the interior kernel: 256 threads, 124-128 registers -> 2 CTAs (16 warps) per SM.  -DPLK_FUSED_THREADS=128
// -DPLK_FUSED_MINB=5 caps the registers at 96 for 5 CTAs (20 warps) per SM at the price of a few spilled doubles (build.py,
// PANSLBM_BUILD_TAG: an A/B variant of the library)
#ifndef PLK_FUSED_THREADS
#define PLK_FUSED_THREADS 256
#define PLK_FUSED_MINB 2
#endif
// one-lattice models need half the registers: 3 CTAs per SM (an explicit bound of 2 would let ptxas spend 128 registers on them)
template <int M> constexpr int fused_min_blocks() { return (ModelFlags<M>::v & F_G) != 0 ? PLK_FUSED_MINB : (PLK_FUSED_MINB*3)/2; }
// The pass is latency-bound, not bandwidth-bound, once it moves 496 B per site: 124-128 registers allow 16 warps per SM, and a warp
// spends more than half of its life waiting for its thirty loads (ncu: long-scoreboard stall 5.3 of 9.2 cycles per issue, 5.7 of
// 6.5 TB/s).  Registers cannot hold more loads in flight — but L2 can: every thread also asks L2 for the lines the thread
// `l2_ahead` sites further on will load (prefetch.global.L2: no register, no L1 line), i.e. for the CTA that follows on this SM;
// when that CTA issues its loads they cost an L2 hit instead of a DRAM round trip.  DRAM traffic is unchanged (each line is
// fetched once, the 126 MB L2 holds the ~20 MB in flight).
PL_D void l2_fetch(const double* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
template <int D, int MODE, unsigned FL>
PL_D void l2_ahead(const double* fs, const double* gs, const CollideParams& P, const Geom& G, long long idx, const Nbr& n, int ahead) {
    const long long a = idx + ahead;
    if (ahead <= 0 || a >= G.npacked) return;
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        const size_t loc = read_loc<D, MODE, c>(G.pitch, a, n);      // the offsets of THIS site: right for every interior site
        l2_fetch(fs + loc);
        if constexpr ((FL & F_G) != 0) l2_fetch(gs + loc);
    });
    if constexpr ((FL & F_KFIELD) != 0) l2_fetch(P.kappa + a);
    if constexpr ((FL & F_BRINK) != 0 || ((FL & F_ADJ) != 0 && (FL & F_G) != 0)) l2_fetch(P.alpha + a);
    if constexpr ((FL & F_ADJ) != 0) {
        l2_fetch(P.rho + a); l2_fetch(P.ux + a); l2_fetch(P.uy + a);
        if constexpr (D == 3) l2_fetch(P.uz + a);
        if constexpr ((FL & F_G) != 0) l2_fetch(P.tem + a);
    }
}
template <int D, int M, int MODE>
__global__ void __launch_bounds__(PLK_FUSED_THREADS, fused_min_blocks<M>()) k_fused(Geom G, const double* fs, double* fd, const double* gs, double* gd,
                                               CollideParams P, ShellMask S, const ClosureArgs* __restrict__ prog, int inverse, XWall W) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr bool HASG = (FL & F_G) != 0;
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.npacked) return;
    int i, j, k;
    decompose(G, idx, i, j, k);
    unsigned long long entries = 0ull;
    int wside = -1;
    if ((S.x[i] | S.y[j] | S.z[k]) != 0ull) {
        const unsigned long long wx = S.x[i], wy = S.y[j], wz = S.z[k];
        if (((wy | wz) & ~TUBE_BIT) != 0ull || (wx & (SLAB_BIT | HALO_BIT)) != 0ull || in_tube(wx, wy, wz)) return;
        entries = wx & ENTRY_BITS;
        if ((wx & GHOST_BIT) != 0ull) wside = i == 0 ? 0 : 1;
        if (entries && prog && S.prefetch) { prefetch_program(prog, entries, i, j, k, idx, S.prefetch); prefetch_collide<FL>(P, idx, S.prefetch); }
    }
    Nbr n = neighbours(G, i, j, k);
    orient(n, inverse);
    double f[LT<D>::nc], g[LT<D>::nc];
    pass_load_wall<D, MODE, HASG>(f, g, fs, gs, G.pitch, idx, n, W, wside, (size_t)(j + G.ny*k), inverse);
    l2_ahead<D, MODE, FL>(fs, gs, P, G, idx, n, S.l2_ahead);
    if (entries && prog) boundary_path<D, HASG>(f, g, prog, entries, i, j, k, idx);
    // a step nobody observes (issave == 2) stores its macros only where the closures of the next step read them: on the x
    // closure planes this kernel owns (the sites of every other closure plane belong to the boundary pass, which always stores)
    collide_site<D, FL, false>(f, g, P, (size_t)idx, P.issave == 1 || (P.issave == 2 && entries != 0ull));
    pass_store<D, MODE>(f, fd, G.pitch, idx, n);
    if constexpr (HASG) pass_store<D, MODE>(g, gd, G.pitch, idx, n);
    wall_scatter<D, HASG>(W, f, g, G, i, j, k, inverse);
}

// The same pass, software-pipelined.  ncu on k_fused (round 2): 124-128 registers allow 2 CTAs = 16 warps per SM; every warp issues
// its thirty loads, waits for them (long-scoreboard stall: two thirds of its life), computes, stores — with 14 warps resident the
// loads in flight per SM do not cover the DRAM latency once the pass moves only 496 B per site (5.4 of 6.5 TB/s).  Here a
// PERSISTENT CTA walks over tiles of 256 sites and copies the populations of its NEXT tile into shared memory with cp.async
// (8 bytes per thread and population, no registers held while in flight) before it computes the current one: 61 KB per CTA
// are always on their way, whatever the register budget.  Every thread reads back only the column it copied itself, so the
// pipeline needs no barrier.  Same arithmetic, same locations, same XWall traffic as k_fused.
PL_D void cp_async8(double* smem, const double* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(s), "l"(gmem) : "memory");
}
PL_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
PL_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
constexpr int PIPE_THREADS = 256;
// what the interior kernel does with site (i,j,k): false = the boundary pass owns it
PL_D bool interior_site(const ShellMask& S, int i, int j, int k, unsigned long long& entries, int& wside) {
    entries = 0ull; wside = -1;
    const unsigned long long wx = S.x[i], wy = S.y[j], wz = S.z[k];
    if ((wx | wy | wz) == 0ull) return true;
    if (((wy | wz) & ~TUBE_BIT) != 0ull || (wx & (SLAB_BIT | HALO_BIT)) != 0ull || in_tube(wx, wy, wz)) return false;
    entries = wx & ENTRY_BITS;
    if ((wx & GHOST_BIT) != 0ull) wside = i == 0 ? 0 : 1;
    return true;
}
template <int D, int M, int MODE>
__global__ void __launch_bounds__(PIPE_THREADS, 2) k_fused_pipe(Geom G, const double* fs, double* fd, const double* gs, double* gd,
                                                                CollideParams P, ShellMask S, int inverse, XWall W) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr bool HASG = (FL & F_G) != 0;
    constexpr int NC = LT<D>::nc;
    extern __shared__ double stage[];      // [(HASG ? 2 : 1)*NC][PIPE_THREADS]: one column per thread
    double* col = stage + threadIdx.x;
    const long long ntiles = (G.npacked + PIPE_THREADS - 1)/PIPE_THREADS;
    auto prefetch = [&](long long tile) {
        const long long idx = tile*PIPE_THREADS + threadIdx.x;
        if (idx >= G.npacked) return;
        int i, j, k, wside;
        unsigned long long entries;
        decompose(G, idx, i, j, k);
        if (!interior_site(S, i, j, k, entries, wside)) return;
        Nbr n = neighbours(G, i, j, k);
        orient(n, inverse);
        const size_t wt = (size_t)(j + G.ny*k);
        sfor<0, NC>([&](auto C) {
            constexpr int c = decltype(C)::value;
            constexpr int X = LT<D>::cx(c);
            const size_t loc = read_loc<D, MODE, c>(G.pitch, idx, n);
            const double *pf = fs + loc, *pg = HASG ? gs + loc : nullptr;
            if constexpr (X != 0) {
                if (wside >= 0 && (inverse ? -X : X) == (wside ? -1 : 1)) {
                    const size_t o = (size_t)(wside*NC + c)*W.np + wt;
                    pf = W.res_f + o;
                    if constexpr (HASG) pg = W.res_g + o;
                }
            }
            cp_async8(col + c*PIPE_THREADS, pf);
            if constexpr (HASG) cp_async8(col + (NC + c)*PIPE_THREADS, pg);
        });
    };
    long long tile = blockIdx.x;
    if (tile < ntiles) prefetch(tile);
    cp_async_commit();
    for (; tile < ntiles; tile += gridDim.x) {
        cp_async_wait_all();
        const long long idx = tile*PIPE_THREADS + threadIdx.x;
        int i = 0, j = 0, k = 0, wside = -1;
        unsigned long long entries = 0ull;
        bool mine = idx < G.npacked;
        if (mine) { decompose(G, idx, i, j, k); mine = interior_site(S, i, j, k, entries, wside); }
        double f[NC], g[NC];
        if (mine) sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = col[c*PIPE_THREADS]; if constexpr (HASG) g[c] = col[(NC + c)*PIPE_THREADS]; });
        // this thread's column is in registers: the next tile may land in it while the current one is computed
        const long long next = tile + gridDim.x;
        if (next < ntiles) prefetch(next);
        cp_async_commit();
        if (mine) {
            Nbr n = neighbours(G, i, j, k);
            orient(n, inverse);
            collide_site<D, FL, false>(f, g, P, (size_t)idx, P.issave == 1 || (P.issave == 2 && entries != 0ull));
            pass_store<D, MODE>(f, fd, G.pitch, idx, n);
            if constexpr (HASG) pass_store<D, MODE>(g, gd, G.pitch, idx, n);
            wall_scatter<D, HASG>(W, f, g, G, i, j, k, inverse);
        }
    }
}

// The boundary pass of a fused step, one thread per listed site: Stream (pull), the closure program, and then either the
// collide of the next step (t < ndirect: sites on closure planes, sites of the last incomplete AVX pack) or, for the
// SmoothCorner tubes (t >= ndirect), a store of the streamed+closed populations into the tube buffer, which k_tubes finishes
// right behind this kernel.  Both run beside k_fused on their own stream.
template <int D, int M, int MODE>
__global__ void __launch_bounds__(SHELL_THREADS) k_shell(Geom G, const double* fs, double* fd, const double* gs, double* gd, CollideParams P, ShellMask S,
                                                         const ClosureArgs* __restrict__ prog, const int* __restrict__ list,
                                                         const unsigned long long* __restrict__ ent, int nlist, int ndirect, int inverse,
                                                         double* __restrict__ tube_f, double* __restrict__ tube_g, HaloView HF, HaloView HG, XWall W) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr bool HASG = (FL & F_G) != 0;
    constexpr int NC = LT<D>::nc;
    __shared__ double tile[(HASG ? 2 : 1)*NC*SHELL_THREADS];
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= nlist) return;
    // one round trip for the site and its closure entries (baked per listed site: no dependent look-up of the plane words)
    const long long idx = list[t];
    const unsigned long long entries = ent[t];
    int i, j, k;
    decompose(G, idx, i, j, k);
    if (entries && S.prefetch) {
        prefetch_program(prog, entries, i, j, k, idx, S.prefetch);
        if (t < ndirect) prefetch_collide<FL>(P, idx, S.prefetch);
    }
    Nbr n = neighbours(G, i, j, k);
    orient(n, inverse);
    double f[NC], g[NC];
    if (HF.on) {
        pass_load_h<D, MODE>(f, fs, G.pitch, idx, i, j, k, G, n, HF, inverse);
        if constexpr (HASG) pass_load_h<D, MODE>(g, gs, G.pitch, idx, i, j, k, G, n, HG, inverse);
    } else {
        pass_load<D, MODE>(f, fs, G.pitch, idx, n);
        if constexpr (HASG) pass_load<D, MODE>(g, gs, G.pitch, idx, n);
    }
    if (entries) boundary_path_sh<D, HASG>(f, g, tile + threadIdx.x, prog, entries, i, j, k, idx);
    if (t < ndirect) {
        if (idx < G.npacked) collide_site<D, FL, false>(f, g, P, (size_t)idx, P.issave != 0);
        else collide_site<D, FL, true>(f, g, P, (size_t)idx, P.issave != 0);
        pass_store<D, MODE>(f, fd, G.pitch, idx, n);
        if constexpr (HASG) pass_store<D, MODE>(g, gd, G.pitch, idx, n);
        wall_scatter<D, HASG>(W, f, g, G, i, j, k, inverse);
    } else {
        // SmoothCorner tube: the streamed + closed populations go to the compact tube buffer [c][ntube]; k_tubes finishes them
        const size_t nt = (size_t)(nlist - ndirect), tt = (size_t)(t - ndirect);
        sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; tube_f[c*nt + tt] = f[c]; if constexpr (HASG) tube_g[c*nt + tt] = g[c]; });
    }
}

// SmoothCorner / SmoothCornerAt + collide of the tube sites, one thread per site, reading the tube buffer k_shell filled and
// writing the destination populations.  kind 0: the site's own populations; kind 1: an edge-line site, a 2-D corner or a
// SmoothCornerAt point/line = the mean of its two inward neighbours (d3q15.h:1242-1290, d2q9.h:578-587); kind 2: a 3-D corner of
// SmoothCorner = the mean of its three neighbouring edge sites (d3q15.h:1291-1303), each of which is the mean of two face sites
// — recomputed here from the face sites with the same operations, so no pass has to wait for another; kind 3: a 3-D
// SmoothCornerAt corner = the mean of three plain sites.  a[] holds tube indices.
struct TubeSite { int idx; int kind[2]; int a[2][6]; };     // kind / neighbours per lattice (0 = flow, 1 = thermal)
template <int D>
PL_D void tube_load(double (&p)[LT<D>::nc], const double* __restrict__ scr, size_t nt, size_t tt, int kind, const int (&a)[6]) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        const double* s = scr + (size_t)c*nt;
        if (kind == 0) p[c] = s[tt];
        else if (kind == 1) p[c] = 0.5*(s[a[0]] + s[a[1]]);
        else if (kind == 3) p[c] = (s[a[0]] + s[a[1]] + s[a[2]])/3.0;
        else {
            const double e0 = 0.5*(s[a[0]] + s[a[1]]), e1 = 0.5*(s[a[2]] + s[a[3]]), e2 = 0.5*(s[a[4]] + s[a[5]]);
            p[c] = (e0 + e1 + e2)/3.0;
        }
    });
}
template <int D, int M, int MODE>
__global__ void __launch_bounds__(128) k_tubes(Geom G, const double* __restrict__ tube_f, const double* __restrict__ tube_g, double* fd,
                                               double* gd, CollideParams P, const TubeSite* __restrict__ info, int nt, XWall W, int inverse) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr bool HASG = (FL & F_G) != 0;
    int tt = blockIdx.x*blockDim.x + threadIdx.x;
    if (tt >= nt) return;
    const TubeSite T = info[tt];
    double f[LT<D>::nc], g[LT<D>::nc];
    tube_load<D>(f, tube_f, (size_t)nt, (size_t)tt, T.kind[0], T.a[0]);
    if constexpr (HASG) tube_load<D>(g, tube_g, (size_t)nt, (size_t)tt, T.kind[1], T.a[1]);
    const long long idx = T.idx;
    if (idx < G.npacked) collide_site<D, FL, false>(f, g, P, (size_t)idx, P.issave != 0);
    else collide_site<D, FL, true>(f, g, P, (size_t)idx, P.issave != 0);
    int i, j, k;
    decompose(G, idx, i, j, k);
    Nbr n = neighbours(G, i, j, k);
    orient(n, inverse);
    pass_store<D, MODE>(f, fd, G.pitch, idx, n);
    if constexpr (HASG) pass_store<D, MODE>(g, gd, G.pitch, idx, n);
    wall_scatter<D, HASG>(W, f, g, G, i, j, k, inverse);
}

// Closures of the x boundary planes of an undecomposed axis, ahead of the fused pass.  On such a plane the populations a
// closure rebuilds are exactly the ones Stream() pulls through the periodic wrap (x - c beyond the wall), and the wrapped-
// around value is dead: the closure overwrites it.  So the closure can run BEFORE the streaming pass, on the compact wall
// buffers (XWall): this kernel reads the populations that arrive at the plane site from `in` (left there by the kernels of the
// previous pass), runs the site's closure program and stores the rebuilt populations in `res`, from where the interior kernel
// patches its pull.  The interior kernel then treats the plane like any other site — no divergent closure code, no strided x
// groups in the boundary pass, no 32-byte sector shared between two kernels, and every access of this kernel is coalesced.
// One thread per plane site (sites that also lie on a y/z closure plane, in a SmoothCorner tube or in the AVX tail stay with
// the boundary pass, which runs their whole program on what it pulls itself).
// Directions the closures of the plane at x = 0 ([0]) / x = nx-1 ([1]) read, per lattice (bit c).
struct XNeed { unsigned f[2], g[2]; };
template <int D, bool HASG>
__global__ void __launch_bounds__(SHELL_THREADS) k_xclose(Geom G, const double* __restrict__ in_f, const double* __restrict__ in_g, double* __restrict__ res_f,
                                                          double* __restrict__ res_g, int np, const ClosureArgs* __restrict__ prog,
                                                          const int* __restrict__ xlist, const unsigned long long* __restrict__ xent, int n, int inverse,
                                                          XNeed need) {
    constexpr int NC = LT<D>::nc;
    __shared__ double tile[(HASG ? 2 : 1)*NC*SHELL_THREADS];
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= n) return;
    const long long idx = xlist[t];
    const unsigned long long entries = xent[t];
    int i, j, k;
    decompose(G, idx, i, j, k);
    const int side = i == 0 ? 0 : 1;
    const size_t tp = (size_t)(j + G.ny*k);
    double f[NC], g[NC];
    sfor<0, NC>([&](auto C) {
        constexpr int c = decltype(C)::value;
        const size_t o = (size_t)(side*NC + c)*np + tp;
        // what the closures read, and the populations that arrive through the wrap (the default of whatever they do not rebuild)
        const bool wrapped = (inverse ? -LT<D>::cx(c) : LT<D>::cx(c)) == (side ? -1 : 1);
        f[c] = (((need.f[side] >> c) & 1u) || wrapped) ? in_f[o] : 0.0;
        if constexpr (HASG) g[c] = (((need.g[side] >> c) & 1u) || wrapped) ? in_g[o] : 0.0;
    });
    boundary_path_sh<D, HASG>(f, g, tile + threadIdx.x, prog, entries, i, j, k, idx);
    const int want = side ? -1 : 1;
    sfor<1, NC>([&](auto C) {
        constexpr int c = decltype(C)::value;
        constexpr int X = LT<D>::cx(c);
        if constexpr (X != 0) {
            if ((inverse ? -X : X) == want) {
                const size_t o = (size_t)(side*NC + c)*np + tp;
                res_f[o] = f[c];
                if constexpr (HASG) res_g[o] = g[c];
            }
        }
    });
}
// fill the compact wall buffer from the (just collided) populations of the lattice: what wall_scatter of a fused pass would have
// left — needed once, when something other than a fused pass of the plan produced the current populations
template <int D>
__global__ void __launch_bounds__(128) k_xfill(Geom G, const double* __restrict__ src, double* __restrict__ out, int np, int inverse, int on0, int on1) {
    constexpr int NC = LT<D>::nc;
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    const int c = blockIdx.y % NC, side = blockIdx.y / NC;
    if (t >= np || !(side ? on1 : on0)) return;
    const int jt = t % G.ny, kt = t / G.ny;
    const int X = rdir<D>(c, 0), Y = rdir<D>(c, 1), Z = rdir<D>(c, 2);
    const int s = inverse ? -1 : 1;
    int is = (side ? G.nx - 1 : 0) - s*X;
    is = is < 0 ? G.nx - 1 : (is >= G.nx ? 0 : is);
    int js = jt - s*Y, ks = kt - s*Z;
    js = js < 0 ? G.ny - 1 : (js >= G.ny ? 0 : js);
    ks = ks < 0 ? G.nz - 1 : (ks >= G.nz ? 0 : ks);
    out[(size_t)(side*NC + c)*np + t] = src[(size_t)c*G.pitch + (size_t)(is + (long long)G.nx*(js + (long long)G.ny*ks))];
}

// SmoothCorner (d3q15.h:199-220, 1242-1303; d2q9.h:127-132, 578-587).  A line/point list is built on the host.
struct SmoothItem {
    double* fb;               // populations of the lattice the item belongs to (one launch serves the flow and the thermal lattice)
    long long base, stride;   // first site of the line and stride along it (corner: a single site)
    int len;                  // number of sites on the line (1 for a corner)
    long long n0, n1, n2;     // index deltas to the 2 (edge / 2-D corner) or 3 (3-D corner) inward neighbours; n2 == 0: two neighbours
};
struct SmoothList { SmoothItem it[24]; int count; int maxlen; };
template <int D>
__global__ void __launch_bounds__(128) k_smooth(Geom G, SmoothList L) {
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    int which = blockIdx.y;
    if (which >= L.count) return;
    const SmoothItem it = L.it[which];
    if (t >= it.len) return;
    long long idx = it.base + (long long)t*it.stride;
    #pragma unroll
    for (int c = 0; c < LT<D>::nc; ++c) {
        double* p = it.fb + (size_t)c*G.pitch;
        if (it.n2 == 0) p[idx] = 0.5*(p[idx + it.n0] + p[idx + it.n1]);
        else p[idx] = (p[idx + it.n0] + p[idx + it.n1] + p[idx + it.n2])/3.0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// layout conversion host(reference AoS) <-> device SoA:  f0[idx], f[(nc-1)*idx + c-1]
template <int D>
__global__ void k_from_aos(Geom G, const double* __restrict__ f0, const double* __restrict__ f, double* __restrict__ dst) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    dst[idx] = f0[idx];
    #pragma unroll
    for (int c = 1; c < LT<D>::nc; ++c) dst[(size_t)c*G.pitch + idx] = f[(size_t)(LT<D>::nc - 1)*idx + (c - 1)];
}
template <int D>
__global__ void k_to_aos(Geom G, const double* __restrict__ src, double* __restrict__ f0, double* __restrict__ f) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    f0[idx] = src[idx];
    #pragma unroll
    for (int c = 1; c < LT<D>::nc; ++c) f[(size_t)(LT<D>::nc - 1)*idx + (c - 1)] = src[(size_t)c*G.pitch + idx];
}
// snapshot SoA -> reference host layout: [pack][c][lane] for packed sites, [idx][c] for the tail (adjointadvection_avx.h:20-22)
template <int D>
__global__ void k_snapshot_to_ref(Geom G, const double* __restrict__ snap, size_t spitch, double* __restrict__ out) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    constexpr int NC = LT<D>::nc;
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
        size_t o = idx < G.npacked ? (size_t)(idx/4)*4*NC + 4*c + idx%4 : (size_t)NC*idx + c;
        out[o] = snap[(size_t)c*spitch + idx];
    }
}

// ... and back: a snapshot the host holds in the reference layout -> device SoA
template <int D>
__global__ void k_snapshot_from_ref(Geom G, const double* __restrict__ in, double* __restrict__ snap, size_t spitch) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    constexpr int NC = LT<D>::nc;
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
        size_t o = idx < G.npacked ? (size_t)(idx/4)*4*NC + 4*c + idx%4 : (size_t)NC*idx + c;
        snap[(size_t)c*spitch + idx] = in[o];
    }
}

// InitialCondition: populations = scalar-order equilibrium (navierstokes.h:550-572, advection.h:1048-1070,
// adjointnavierstokes.h:474-498, adjointadvection.h:1359-1381).  a0..a6 by family as in pl_initial_condition.
template <int D>
__global__ void __launch_bounds__(256) k_init(Geom G, double* __restrict__ dst, int family, const double* a0, const double* a1, const double* a2,
                                              const double* a3, const double* a4, const double* a5, const double* a6) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    double eq[LT<D>::nc];
    if (family == 1) {          // NS: rho, ux, uy, uz
        ns_eq_sc<D>(eq, a0[idx], a1[idx], a2[idx], D == 3 ? a3[idx] : 0.0);
    } else if (family == 2) {   // AD: tem, ux, uy, uz
        ad_eq_sc<D>(eq, a0[idx], a1[idx], a2[idx], D == 3 ? a3[idx] : 0.0);
    } else if (family == 3) {   // ANS: ux, uy, uz, ip, iux, iuy, iuz
        ans_eq<D>(eq, a0[idx], a1[idx], D == 3 ? a2[idx] : 0.0, a3[idx], a4[idx], a5[idx], D == 3 ? a6[idx] : 0.0);
    } else if (family == 5) {   // NSin: rho, ux, uy (nsincompressible.h:212-223)
        nsin_eq<D>(eq, a0[idx], a1[idx], a2[idx], D == 3 ? a3[idx] : 0.0);
    } else {                    // AAD: ux, uy, uz, item, iqx, iqy, iqz
        double ux = a0[idx], uy = a1[idx], uz = D == 3 ? a2[idx] : 0.0;
        // scalar order: item + 3*(ux*iqx + uy*iqy + uz*iqz)  (adjointadvection.h:54-68)
        double ge = a3[idx] + 3.0*dot<D>(ux, uy, uz, a4[idx], a5[idx], D == 3 ? a6[idx] : 0.0);
        #pragma unroll
        for (int c = 0; c < LT<D>::nc; ++c) eq[c] = ge;
    }
    store_site<D>(eq, dst, G.pitch, idx);
}

}  // namespace plb
