// Lattice constants and compile-time helpers for the D2Q9 / D3Q15 kernels.
// Velocity sets and weights: reference src/particle/d2q9.h:160-163, src/particle/d3q15.h:251-254.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>

// The site-local math (this header, lbm_equations.cuh, lbm_closures.cuh, lbm_sens.cuh) also compiles as plain
// host C++ (tests/hostmath: the not-gpu suite checks the arithmetic of the product code against the reference
// build); kernels are guarded by __CUDACC__.
#ifdef __CUDACC__
#define PL_HD __host__ __device__ __forceinline__
#define PL_D __device__ __forceinline__
#define PL_UNROLL _Pragma("unroll")
#else
#define PL_HD inline
#define PL_D inline
#define PL_UNROLL
#endif

namespace plb {

template <int D> struct LT;

template <> struct LT<2> {
    static constexpr int nc = 9, nd = 2;
    PL_HD static constexpr int cx(int c) { constexpr int t[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1}; return t[c]; }
    PL_HD static constexpr int cy(int c) { constexpr int t[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1}; return t[c]; }
    PL_HD static constexpr int cz(int) { return 0; }
    PL_HD static constexpr double ei(int c) {
        constexpr double t[9] = {4.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/36.0, 1.0/36.0, 1.0/36.0, 1.0/36.0};
        return t[c];
    }
    PL_HD static constexpr int opp(int c) { constexpr int t[9] = {0, 3, 4, 1, 2, 7, 8, 5, 6}; return t[c]; }
};

template <> struct LT<3> {
    static constexpr int nc = 15, nd = 3;
    PL_HD static constexpr int cx(int c) { constexpr int t[15] = {0, 1, 0, 0, -1, 0, 0, 1, -1, 1, 1, -1, 1, -1, -1}; return t[c]; }
    PL_HD static constexpr int cy(int c) { constexpr int t[15] = {0, 0, 1, 0, 0, -1, 0, 1, 1, -1, 1, -1, -1, 1, -1}; return t[c]; }
    PL_HD static constexpr int cz(int c) { constexpr int t[15] = {0, 0, 0, 1, 0, 0, -1, 1, 1, 1, -1, -1, -1, -1, 1}; return t[c]; }
    PL_HD static constexpr double ei(int c) {
        constexpr double t[15] = {2.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0,
                                  1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0};
        return t[c];
    }
    PL_HD static constexpr int opp(int c) { constexpr int t[15] = {0, 4, 5, 6, 1, 2, 3, 11, 12, 13, 14, 7, 8, 9, 10}; return t[c]; }
};

template <int D> PL_HD constexpr int cdir(int c, int axis) { return axis == 0 ? LT<D>::cx(c) : (axis == 1 ? LT<D>::cy(c) : LT<D>::cz(c)); }

// Run-time lookups (c not a compile-time constant: the boundary closures).  Indexing the constexpr tables above with a
// run-time c makes the compiler rebuild the table on the thread's stack at every call; these bit-packed words are pure ALU.
template <int D> PL_HD constexpr unsigned pack_dirs(int axis) {
    unsigned v = 0;
    for (int c = 0; c < LT<D>::nc; ++c) v |= (unsigned)(cdir<D>(c, axis) + 1) << (2*c);
    return v;
}
template <int D> PL_HD constexpr unsigned long long pack_opps() {
    unsigned long long v = 0;
    for (int c = 0; c < LT<D>::nc; ++c) v |= (unsigned long long)LT<D>::opp(c) << (4*c);
    return v;
}
template <int D> PL_HD int rdir(int c, int axis) {
    constexpr unsigned X = pack_dirs<D>(0), Y = pack_dirs<D>(1), Z = pack_dirs<D>(2);
    const unsigned w = axis == 0 ? X : (axis == 1 ? Y : Z);
    return (int)((w >> (2*c)) & 3u) - 1;
}
template <int D> PL_HD int ropp(int c) {
    constexpr unsigned long long O = pack_opps<D>();
    return (int)((O >> (4*c)) & 15ull);
}

// index of the direction with the given integer velocity, -1 if none (run-time helper for table-driven closures)
template <int D> PL_HD int find_dir(int x, int y, int z) {
    int r = -1;
    PL_UNROLL
    for (int c = LT<D>::nc - 1; c >= 0; --c)
        if (rdir<D>(c, 0) == x && rdir<D>(c, 1) == y && rdir<D>(c, 2) == z) r = c;
    return r;
}

// compile-time loop: f(std::integral_constant<int, c>) for c in [B, E)
template <int B, int E, class F> PL_HD void sfor(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        sfor<B + 1, E>(f);
    }
}

// acc + S*v for S in {-1,0,+1} with the exact value the reference's literal multiply-by-constant gives
template <int S> PL_HD double sadd(double acc, double v) {
    if constexpr (S == 0) return acc;
    else if constexpr (S > 0) return acc + v;
    else return acc - v;
}

// (cx*vx + cy*vy) [+ cz*vz] with the zero terms dropped (adding an exact 0.0 never changes a value)
template <int D, int c> PL_HD double cdot(double vx, double vy, double vz) {
    constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
    if constexpr (X == 0 && Y == 0 && Z == 0) return 0.0;
    double s;
    if constexpr (X != 0) { s = X > 0 ? vx : -vx; s = sadd<Y>(s, vy); }
    else if constexpr (Y != 0) { s = Y > 0 ? vy : -vy; }
    if constexpr (X == 0 && Y == 0) { s = Z > 0 ? vz : -vz; }
    else s = sadd<Z>(s, vz);
    return s;
}

// Local geometry of one rank's block (reference D3Q15 public ints, d3q15.h:222)
struct Geom {
    int nx, ny, nz;
    int lx, ly, lz;
    int offx, offy, offz;
    long long nxyz;      // nx*ny*nz
    long long npacked;   // 4*(nxyz/4): sites the reference's AVX overloads handle (navierstokes_avx.h:96)
    size_t pitch;        // doubles between consecutive population planes (>= nxyz, multiple of 16)
};

}  // namespace plb
