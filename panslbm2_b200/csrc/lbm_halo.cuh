// Halo exchange of a block-decomposed lattice: what replaces D3Q15::Communicate / D2Q9::Communicate
// (reference d3q15.h:1306-1407, d2q9.h:590-630) and the pack/unpack loops around it (d3q15.h:296-436, 458-598).
//
// The reference keeps no ghost layers: each rank streams with a LOCAL periodic wrap, packs the populations that
// wrapped around (5 per face site, 2 per edge site, 1 per corner), sends them to the 26 neighbours of a periodic PE
// grid and overwrites the wrapped slots with what it receives.  Here the same messages exist (same sets, same
// counts), but nothing is unpacked: a population whose source site lies beyond a decomposed block face is pulled
// straight out of the receive buffer by the streaming kernel (pull_halo).  Axes with m == 1 are not exchanged; the
// local wrap of neighbours() already is the global periodic wrap there.
//
// Message for the neighbour in direction o = (ox,oy,oz), o != 0, nonzero only along decomposed axes:
//   sites    the local sites with x_a = (o_a > 0 ? n_a-1 : 0) on every axis a with o_a != 0, free along the others,
//            ordered lower free axis fastest ("region", rsize sites);
//   pops     Stream:  every c with c_a ==  o_a on all those axes (they leave through that face/edge/corner),
//            iStream: every c with c_a == -o_a, in ascending c ("slot");
//   layout   [slot][region site]  (SoA: pack and pull are coalesced along the fastest free axis).
#pragma once
#include "lbm_traits.cuh"

namespace plb {

PL_HD constexpr int halo_code(int ox, int oy, int oz) { return (ox + 1) + 3*(oy + 1) + 9*(oz + 1); }

// slot of population c inside the message crossing the axes of `mask` (bit a set = axis a crosses): the number of
// c' < c that agree with c on every crossing axis.  Packed 3 bits per mask value so that a run-time mask costs ALU only.
template <int D> PL_HD constexpr int halo_slot_of(int c, int mask) {
    int s = 0;
    for (int d = 0; d < c; ++d) {
        bool same = true;
        for (int a = 0; a < 3; ++a) if (((mask >> a) & 1) && cdir<D>(d, a) != cdir<D>(c, a)) same = false;
        if (same) ++s;
    }
    return s;
}
template <int D, int c> PL_HD constexpr unsigned halo_slot_word() {
    unsigned w = 0;
    for (int m = 1; m < 8; ++m) {
        bool ok = true;   // only masks whose axes all have c_a != 0 can occur
        for (int a = 0; a < 3; ++a) if (((m >> a) & 1) && cdir<D>(c, a) == 0) ok = false;
        if (ok) w |= (unsigned)halo_slot_of<D>(c, m) << (3*m);
    }
    return w;
}

struct HaloView {
    const double* r[27];   // receive buffer of the message from the neighbour in direction code (null: none)
    int e[3];              // axis is decomposed (m > 1)
    int on;                // any axis decomposed
};

#ifdef __CUDACC__
// population c of the site (i,j,k) after Stream (inverse = 0: source x - c) / iStream (inverse = 1: source x + c);
// `n` holds the local periodic neighbour deltas already oriented for `inverse` (see orient()).
template <int D, int c, class NbrT>
PL_D double pull_halo(const double* __restrict__ src, long long idx, int i, int j, int k, const Geom& G, const NbrT& n,
                      size_t local_loc, const HaloView& H, int inverse) {
    constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
    const int dx = inverse ? X : -X, dy = inverse ? Y : -Y, dz = inverse ? Z : -Z;   // where the source lies
    const bool bx = X != 0 && H.e[0] && (dx < 0 ? i == 0 : i == G.nx - 1);
    const bool by = Y != 0 && H.e[1] && (dy < 0 ? j == 0 : j == G.ny - 1);
    const bool bz = D == 3 && Z != 0 && H.e[2] && (dz < 0 ? k == 0 : k == G.nz - 1);
    if (!(bx || by || bz)) return src[local_loc];      // inside the block: wherever the layout of the pass keeps it (lbm_kernels.cuh: read_loc)
    const int code = halo_code(bx ? dx : 0, by ? dy : 0, bz ? dz : 0);
    const int mask = (bx ? 1 : 0) | (by ? 2 : 0) | (bz ? 4 : 0);
    // source coordinates along the free axes (local periodic wrap as Index(), d3q15.h:136-141)
    int si = i + dx; si = si < 0 ? G.nx - 1 : (si >= G.nx ? 0 : si);
    int sj = j + dy; sj = sj < 0 ? G.ny - 1 : (sj >= G.ny ? 0 : sj);
    int sk = k + dz; sk = sk < 0 ? G.nz - 1 : (sk >= G.nz ? 0 : sk);
    long long ridx = 0, rs = 1;
    if (!bx) { ridx += si*rs; rs *= G.nx; }
    if (!by) { ridx += sj*rs; rs *= G.ny; }
    if (D == 3 && !bz) { ridx += sk*rs; rs *= G.nz; }
    constexpr unsigned W = halo_slot_word<D, c>();
    const int slot = (int)((W >> (3*mask)) & 7u);
    return H.r[code][(size_t)slot*(size_t)rs + (size_t)ridx];
}

// One message of the pack kernel.
struct PackMsg {
    double* dst;             // send buffer
    long long base;          // index of region site 0
    long long s1, s2;        // strides of the (up to two) free axes, lower first
    int n1, n2;              // extents of the free axes (1 when absent)
    int npop;
    int pop[5];
};
struct PackList { PackMsg m[26]; int count; };

// gather the outgoing populations of every message from the current populations (one launch per lattice; grid.y = message).
// streamed != 0: the lattice is in the streamed layout P of the in-place passes (lbm_kernels.cuh): population c of site x sits at
// (opp(c), x + s*c), the step taken with the LOCAL periodic wrap — for an outgoing population that is a site of the opposite face.
template <int D>
__global__ void __launch_bounds__(128) k_halo_pack(const double* __restrict__ cur, Geom G, PackList L, int streamed, int inverse) {
    const PackMsg& M = L.m[blockIdx.y];
    const long long rsize = (long long)M.n1*M.n2;
    const size_t pitch = G.pitch;
    for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < rsize; t += (long long)gridDim.x*blockDim.x) {
        const long long a = t%M.n1, b = t/M.n1;
        const long long site = M.base + a*M.s1 + b*M.s2;
        if (!streamed) {
            for (int s = 0; s < M.npop; ++s) M.dst[(size_t)s*rsize + t] = cur[(size_t)M.pop[s]*pitch + site];
        } else {
            const long long nxy = (long long)G.nx*G.ny;
            const int k = (int)(site/nxy), j = (int)((site - k*nxy)/G.nx), i = (int)(site - k*nxy - (long long)j*G.nx);
            const int sg = inverse ? -1 : 1;
            for (int s = 0; s < M.npop; ++s) {
                const int c = M.pop[s];
                int ti = i + sg*rdir<D>(c, 0), tj = j + sg*rdir<D>(c, 1), tk = k + sg*rdir<D>(c, 2);
                ti = ti < 0 ? G.nx - 1 : (ti >= G.nx ? 0 : ti);
                tj = tj < 0 ? G.ny - 1 : (tj >= G.ny ? 0 : tj);
                tk = tk < 0 ? G.nz - 1 : (tk >= G.nz ? 0 : tk);
                M.dst[(size_t)s*rsize + t] = cur[(size_t)ropp<D>(c)*pitch + (size_t)(ti + (long long)G.nx*(tj + (long long)G.ny*tk))];
            }
        }
    }
}
#endif

// ---- host-side description of the exchange (pure arithmetic: also used without a device, pl_halo_describe) ----------
struct HaloMsgDesc {
    int code;                // direction code of the neighbour this message goes to / comes from
    int o[3];
    int peer;                // rank (PEid) of that neighbour on the periodic PE grid (IndexPE, d3q15.h:145-150)
    long long rsize;         // region sites
    int npop;
    int pop[5];              // ascending c
    long long base, s1, s2;  // send region geometry (see PackMsg)
    int n1, n2;
};

// messages of one rank for Stream (inverse = 0) / iStream (inverse = 1), in the fixed direction order every rank uses
// (ascending code).  Rank r SENDS message `code` to peer(code) and RECEIVES its message `code` from peer(opposite code):
// with that rule two ranks that are each other's neighbour in several directions (m == 2) still pair their messages.
template <int D>
inline int halo_describe(const int n[3], const int m[3], const int pe[3], int inverse, HaloMsgDesc out[26]) {
    int cnt = 0;
    const long long st[3] = {1, n[0], (long long)n[0]*n[1]};
    for (int code = 0; code < 27; ++code) {
        const int o[3] = {code%3 - 1, (code/3)%3 - 1, code/9 - 1};
        if (o[0] == 0 && o[1] == 0 && o[2] == 0) continue;
        bool ok = true;
        for (int a = 0; a < 3; ++a) if (o[a] != 0 && (a >= D || m[a] <= 1)) ok = false;
        if (!ok) continue;
        HaloMsgDesc& d = out[cnt];
        d.code = code; d.o[0] = o[0]; d.o[1] = o[1]; d.o[2] = o[2];
        int q[3];
        for (int a = 0; a < 3; ++a) { q[a] = pe[a] + o[a]; q[a] = q[a] < 0 ? m[a] - 1 : (q[a] >= m[a] ? 0 : q[a]); }
        d.peer = q[0] + m[0]*(q[1] + m[1]*q[2]);
        d.npop = 0;
        for (int c = 1; c < LT<D>::nc; ++c) {
            bool in = true;
            for (int a = 0; a < 3; ++a) if (o[a] != 0 && cdir<D>(c, a) != (inverse ? -o[a] : o[a])) in = false;
            if (in) d.pop[d.npop++] = c;
        }
        d.base = 0; d.s1 = d.s2 = 0; d.n1 = d.n2 = 1;
        int nfree = 0;
        for (int a = 0; a < 3; ++a) {
            if (o[a] != 0) d.base += (long long)(o[a] > 0 ? n[a] - 1 : 0)*st[a];
            else if (nfree == 0) { d.s1 = st[a]; d.n1 = n[a]; ++nfree; }
            else if (nfree == 1) { d.s2 = st[a]; d.n2 = n[a]; ++nfree; }
            else { d.n2 *= n[a]; }   // unreachable: o != 0 leaves at most two free axes
        }
        d.rsize = (long long)d.n1*d.n2;
        ++cnt;
    }
    return cnt;
}
inline int halo_opposite(int code) { return 26 - code; }

}  // namespace plb
