// Boundary closures of the NS / AD / ANS / AAD equations on one lattice plane (the reference's
// "...AlongXFace/YFace/ZFace" and "...AlongXEdge/YEdge" helpers).  They have no AVX version in the reference,
// so the operation order is the scalar one of src/equation/*.h.  One thread per plane site; the host
// callables of the reference are baked into per-plane mask/value arrays (see include/panslbm_c.h).
#pragma once
#include "lbm_kernels.cuh"

namespace plb {

struct ClosureArgs {
    int type;                      // PL_BC_*
    Plane pl;
    const uint8_t* mask;           // per plane site
    const double *v0, *v1, *v2;    // per plane site values
    const double *rho, *ux, *uy, *uz, *tem, *kappa;   // per lattice site fields (device, may be null)
    double kconst, eps;
};

// NS::BoundaryConditionSetU / SetRho (navierstokes.h:92-426).  Normal axis a, outward direction dir:
//   "in"  = populations entering the domain (c_a == -dir): the unknowns,
//   "out" = their opposites (c_a == dir), "tan" = c_a == 0.
//   SetU:   rho0 = (f0 + sum(tan) + 2*sum(out))/(1 + dir*u_a)
//   SetRho: u_a  = -dir*(1 - (f0 + sum(tan) + 2*sum(out))/rho)
//   m_a = rho0*u_a/(6|12), m_t = (1/2|1/4)*(f_{+t} - f_{-t} - rho0*u_t)
//   f_in(axis) = f_out -dir*(4|8)*m_a,  f_in(diagonal) = f_opp + sum_d s_d m_d with s_a = c_a, s_t = -c_t,
// sums in ascending c and x,y,z order as written in the reference.
template <int D>
PL_D void closure_ns(double (&p)[LT<D>::nc], const ClosureArgs& A, int t, bool setrho) {
    constexpr int NC = LT<D>::nc;
    const int axis = A.pl.axis, dir = A.pl.dir;
    double s = p[0];
    for (int c = 1; c < NC; ++c) if (cdir<D>(c, axis) == 0) s = s + p[c];
    double o = 0.0; bool first = true;
    for (int c = 1; c < NC; ++c) if (cdir<D>(c, axis) == dir) { o = first ? p[c] : o + p[c]; first = false; }
    const double tot = s + 2.0*o;
    double u[3] = {0.0, 0.0, 0.0}, rho0;
    const int t1 = D == 2 ? 1 - axis : (axis + 1)%3, t2 = D == 2 ? -1 : (axis + 2)%3;
    if (!setrho) {
        u[0] = A.v0[t]; u[1] = A.v1[t]; if (D == 3) u[2] = A.v2[t];
        rho0 = dir == -1 ? tot/(1.0 - u[axis]) : tot/(1.0 + u[axis]);
    } else {
        // reference argument order (rho, us, ut): (uy,uz) on X, (uz,ux) on Y, (ux,uy) on Z faces (navierstokes.h:326,363,400)
        rho0 = A.v0[t];
        u[t1] = A.v1[t];
        if (D == 3) u[t2] = A.v2[t];
        u[axis] = dir == -1 ? 1.0 - tot/rho0 : -1.0 + tot/rho0;
    }
    const double kn = D == 2 ? 6.0 : 12.0, kt = D == 2 ? 0.5 : 0.25, ka = D == 2 ? 4.0 : 8.0;
    double m[3] = {0.0, 0.0, 0.0};
    m[axis] = rho0*u[axis]/kn;
    for (int d = 0; d < D; ++d) if (d != axis) {
        int cp = find_dir<D>(d == 0, d == 1, d == 2), cm = find_dir<D>(-(d == 0), -(d == 1), -(d == 2));
        m[d] = kt*(p[cp] - p[cm] - rho0*u[d]);
    }
    double out[NC];
    for (int c = 1; c < NC; ++c) {
        out[c] = p[c];
        if (cdir<D>(c, axis) != -dir) continue;
        int nz = abs(LT<D>::cx(c)) + abs(LT<D>::cy(c)) + abs(LT<D>::cz(c));
        double val;
        if (nz == 1) val = dir == -1 ? p[LT<D>::opp(c)] + ka*m[axis] : p[LT<D>::opp(c)] - ka*m[axis];
        else {
            val = p[LT<D>::opp(c)];
            for (int d = 0; d < D; ++d) {
                int sg = d == axis ? cdir<D>(c, d) : -cdir<D>(c, d);
                val = sg > 0 ? val + m[d] : val - m[d];
            }
        }
        out[c] = val;
    }
    for (int c = 1; c < NC; ++c) p[c] = out[c];
}

template <int D>
__global__ void __launch_bounds__(128) k_closure(Geom G, double* __restrict__ fb, double* __restrict__ gb, ClosureArgs A) {
    int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= A.pl.n1*A.pl.n2) return;
    if (!A.mask[t]) return;
    int a = t%A.pl.n1, b = t/A.pl.n1;
    long long idx = A.pl.base + a*A.pl.s1 + b*A.pl.s2;
    double p[LT<D>::nc];
    #pragma unroll
    for (int c = 0; c < LT<D>::nc; ++c) p[c] = fb[(size_t)c*G.pitch + idx];
    switch (A.type) {
        case 3: closure_ns<D>(p, A, t, false); break;
        case 4: closure_ns<D>(p, A, t, true); break;
        default: return;
    }
    #pragma unroll
    for (int c = 1; c < LT<D>::nc; ++c) fb[(size_t)c*G.pitch + idx] = p[c];
}

}  // namespace plb
