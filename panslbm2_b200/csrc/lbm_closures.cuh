// Boundary closures of the particle classes and of the NS / AD / ANS / AAD equations on one lattice plane (the
// reference's "...AlongXFace/YFace/ZFace" and "...AlongXEdge/YEdge" helpers), as site-local functions over the
// register copy p[] of one site's populations.  They have no AVX version in the reference, so the operation
// order is the scalar one of src/equation/*.h and src/particle/*.h.  The host callables of the reference are
// baked into per-plane mask/value arrays (see include/panslbm_c.h).
//
// The population arguments are generic (`PA`): a plain `double[nc]` (registers / the host build) or any object with
// operator[](int) -> double& (the strided shared-memory view of the boundary pass, lbm_kernels.cuh).
//
// Vocabulary for a plane with normal axis a and outward direction dir:
//   K = populations with c_a == -dir, ascending c (the axis-aligned one first, then the diagonals);
//       for the forward closures these are the unknowns entering the domain, for the adjoint ("i") closures
//       they are the known ones and the unknowns are their opposites.
#pragma once
#include "lbm_equations.cuh"

namespace plb {

// closure types (PL_BC_* of include/panslbm_c.h)
enum : int {
    BC_BOUNCE = 1, BC_IBOUNCE = 2, BC_NS_SET_U = 3, BC_NS_SET_RHO = 4, BC_AD_SET_T = 5, BC_AD_SET_Q = 6,
    BC_ANS_ISET_U = 7, BC_ANS_ISET_RHO = 8, BC_AAD_ISET_T = 9, BC_AAD_ISET_Q = 10, BC_AAD_ISET_RHO = 11,
    BC_NSIN_SET_U = 12, BC_NSIN_SET_RHO = 13
};

// per-site inputs of one closure application
struct SiteVals {
    double v0, v1, v2;              // baked plane values (meaning by type)
    double rho, ux, uy, uz, tem;    // macroscopic fields at the site (saved by the collide of the same step)
    double kappa;                   // diffusivity at the site (field or scalar overload)
    double eps;
};

template <int D> struct FaceK { static constexpr int n = D == 2 ? 3 : 5; };

// K list: c with c_axis == sgn, ascending
template <int D> PL_HD void face_list(int axis, int sgn, int (&K)[FaceK<D>::n]) {
    int m = 0;
    PL_UNROLL
    for (int c = 1; c < LT<D>::nc; ++c)
        if (rdir<D>(c, axis) == sgn && m < FaceK<D>::n) K[m++] = c;
}
// sum_i sign(c_b(K_i)) p[K_i] over the diagonals, left to right
template <int D, class PA> PL_HD double signed_diag_sum(const PA& p, const int (&K)[FaceK<D>::n], int b) {
    double s = rdir<D>(K[1], b) > 0 ? p[K[1]] : -p[K[1]];
    PL_UNROLL
    for (int i = 2; i < FaceK<D>::n; ++i) s = rdir<D>(K[i], b) > 0 ? s + p[K[i]] : s - p[K[i]];
    return s;
}
// w*p[K0] + p[K1] + ... left to right
template <int D, class PA> PL_HD double weighted_face_sum(const PA& p, const int (&K)[FaceK<D>::n], double w) {
    double s = w*p[K[0]];
    PL_UNROLL
    for (int i = 1; i < FaceK<D>::n; ++i) s = s + p[K[i]];
    return s;
}
// 1.0 + cx*3ux + cy*3uy (+ cz*3uz), left to right, zero components skipped
template <int D> PL_HD double one_plus_3cu(int c, double ux, double uy, double uz) {
    double t = 1.0;
    int x = rdir<D>(c, 0), y = rdir<D>(c, 1), z = rdir<D>(c, 2);
    if (x) t = x > 0 ? t + 3.0*ux : t - 3.0*ux;
    if (y) t = y > 0 ? t + 3.0*uy : t - 3.0*uy;
    if (D == 3 && z) t = z > 0 ? t + 3.0*uz : t - 3.0*uz;
    return t;
}
PL_HD double pick(int axis, double x, double y, double z) { return axis == 0 ? x : (axis == 1 ? y : z); }

// ---------------------------------------------------------------------------------------------------------
// BARRIER / MIRROR (d3q15.h:984-1239, d2q9.h:431-575): forward rebuilds the populations entering the domain
// (c_a == -dir) from their opposite (BARRIER=1) or mirror image (MIRROR=2); inverse the leaving ones.
template <int D, class PA> PL_HD void closure_bounce(PA& p, int axis, int dir, int type, bool inverse) {
    if (type != 1 && type != 2) return;
    const int want = inverse ? dir : -dir;
    PL_UNROLL
    for (int c = 1; c < LT<D>::nc; ++c) {
        if (rdir<D>(c, axis) != want) continue;
        int src;
        if (type == 1) src = ropp<D>(c);
        else {
            int x = rdir<D>(c, 0), y = rdir<D>(c, 1), z = rdir<D>(c, 2);
            if (axis == 0) x = -x; else if (axis == 1) y = -y; else z = -z;
            src = find_dir<D>(x, y, z);
        }
        p[c] = p[src];   // sources have c_a == -want: never overwritten in this loop
    }
}

// NS::BoundaryConditionSetU / SetRho (navierstokes.h:92-426).
//   "in"  = populations entering the domain (c_a == -dir): the unknowns,
//   "out" = their opposites (c_a == dir), "tan" = c_a == 0.
//   SetU:   rho0 = (f0 + sum(tan) + 2*sum(out))/(1 + dir*u_a)
//   SetRho: u_a  = -dir*(1 - (f0 + sum(tan) + 2*sum(out))/rho)
//   m_a = rho0*u_a/(6|12), m_t = (1/2|1/4)*(f_{+t} - f_{-t} - rho0*u_t)
//   f_in(axis) = f_out -dir*(4|8)*m_a,  f_in(diagonal) = f_opp + sum_d s_d m_d with s_a = c_a, s_t = -c_t,
// sums in ascending c and x,y,z order as written in the reference.
template <int D, class PA> PL_HD void closure_ns(PA& p, int axis, int dir, const SiteVals& V, bool setrho) {
    constexpr int NC = LT<D>::nc;
    double s = p[0];
    PL_UNROLL
    for (int c = 1; c < NC; ++c) if (rdir<D>(c, axis) == 0) s = s + p[c];
    double o = 0.0; bool first = true;
    PL_UNROLL
    for (int c = 1; c < NC; ++c) if (rdir<D>(c, axis) == dir) { o = first ? p[c] : o + p[c]; first = false; }
    const double tot = s + 2.0*o;
    double u[3] = {0.0, 0.0, 0.0}, rho0;
    const int t1 = D == 2 ? 1 - axis : (axis + 1)%3, t2 = D == 2 ? -1 : (axis + 2)%3;
    if (!setrho) {
        u[0] = V.v0; u[1] = V.v1; if (D == 3) u[2] = V.v2;
        rho0 = dir == -1 ? tot/(1.0 - u[axis]) : tot/(1.0 + u[axis]);
    } else {
        // reference argument order (rho, us, ut): (uy,uz) on X, (uz,ux) on Y, (ux,uy) on Z faces (navierstokes.h:326,363,400)
        rho0 = V.v0;
        u[t1] = V.v1;
        if (D == 3) u[t2] = V.v2;
        u[axis] = dir == -1 ? 1.0 - tot/rho0 : -1.0 + tot/rho0;
    }
    const double kn = D == 2 ? 6.0 : 12.0, kt = D == 2 ? 0.5 : 0.25, ka = D == 2 ? 4.0 : 8.0;
    double m[3] = {0.0, 0.0, 0.0};
    m[axis] = rho0*u[axis]/kn;
    PL_UNROLL
    for (int d = 0; d < D; ++d) if (d != axis) {
        int cp = find_dir<D>(d == 0, d == 1, d == 2), cm = find_dir<D>(-(d == 0), -(d == 1), -(d == 2));
        m[d] = kt*(p[cp] - p[cm] - rho0*u[d]);
    }
    double out[NC];
    PL_UNROLL
    for (int c = 1; c < NC; ++c) {
        out[c] = p[c];
        if (rdir<D>(c, axis) != -dir) continue;
        int nz = abs(rdir<D>(c, 0)) + abs(rdir<D>(c, 1)) + abs(rdir<D>(c, 2));
        double val;
        if (nz == 1) val = dir == -1 ? p[ropp<D>(c)] + ka*m[axis] : p[ropp<D>(c)] - ka*m[axis];
        else {
            val = p[ropp<D>(c)];
            PL_UNROLL
            for (int d = 0; d < D; ++d) {
                int sg = d == axis ? rdir<D>(c, d) : -rdir<D>(c, d);
                val = sg > 0 ? val + m[d] : val - m[d];
            }
        }
        out[c] = val;
    }
    PL_UNROLL
    for (int c = 1; c < NC; ++c) p[c] = out[c];
}

// NSin::BoundaryConditionSetU / SetRho along an edge of a D2Q9 lattice (nsincompressible.h:46-154).  a = normal axis, t = the other.
//   SetU:   u given.   SetRho: u_t given, u_a = -dir*(rho - (f0 + sum(tan) + 2*sum(out)))  (v0 = rho, v1 = u_t)
//   f_in(axis)     = f_out - dir*(2 u_a/3)
//   f_in(diagonal) = f_opp - dir*(u_a/6) - c_t*0.5*(f_{+t} - f_{-t} - u_t)
template <int D, class PA> PL_HD void closure_nsin(PA& p, int axis, int dir, const SiteVals& V, bool setrho) {
    if constexpr (D == 2) {
        const int t = 1 - axis;
        double ua, ut;
        if (!setrho) { ua = axis == 0 ? V.v0 : V.v1; ut = axis == 0 ? V.v1 : V.v0; }
        else {
            double s = p[0];
            PL_UNROLL
            for (int c = 1; c < 9; ++c) if (rdir<2>(c, axis) == 0) s = s + p[c];
            double o = 0.0; bool first = true;
            PL_UNROLL
            for (int c = 1; c < 9; ++c) if (rdir<2>(c, axis) == dir) { o = first ? p[c] : o + p[c]; first = false; }
            const double tot = s + 2.0*o;
            ua = dir == -1 ? V.v0 - tot : -V.v0 + tot;
            ut = V.v1;
        }
        const int cp = find_dir<2>(t == 0, t == 1, 0), cm = find_dir<2>(-(t == 0), -(t == 1), 0);
        const double T = p[cp] - p[cm] - ut;
        double out[9];
        PL_UNROLL
        for (int c = 1; c < 9; ++c) {
            out[c] = p[c];
            if (rdir<2>(c, axis) != -dir) continue;
            const int ct = rdir<2>(c, t);
            double val;
            if (ct == 0) val = dir == -1 ? p[ropp<2>(c)] + 2.0*ua/3.0 : p[ropp<2>(c)] - 2.0*ua/3.0;
            else {
                val = dir == -1 ? p[ropp<2>(c)] + ua/6.0 : p[ropp<2>(c)] - ua/6.0;
                val = ct > 0 ? val - 0.5*T : val + 0.5*T;
            }
            out[c] = val;
        }
        PL_UNROLL
        for (int c = 1; c < 9; ++c) p[c] = out[c];
    }
}

// AD::BoundaryConditionSetT (advection.h:99-238) and SetQ (advection.h:242-524; scalar and per-cell diffusivity):
//   SetT: tem0 = 6*(T - g0 - sum_{c_a != -dir} g_c)/(1 - dir*3u_a)
//   SetQ: tem0 = 6*((1 + 1/(6 kappa))*qn + sum_{c_a == dir} g_c)/(1 + dir*3u_a)
//   g_in = tem0*(1 + 3 c.u)/(9 | 36 | 72)
template <int D, class PA> PL_HD void closure_ad(PA& g, int axis, int dir, const SiteVals& V, bool setq) {
    constexpr int NC = LT<D>::nc;
    const double ua = pick(axis, V.ux, V.uy, V.uz);
    double tem0;
    if (!setq) {
        double s = V.v0 - g[0];
        PL_UNROLL
        for (int c = 1; c < NC; ++c) if (rdir<D>(c, axis) != -dir) s = s - g[c];
        tem0 = dir == -1 ? 6.0*s/(1.0 + 3.0*ua) : 6.0*s/(1.0 - 3.0*ua);
    } else {
        double s = (1.0 + 1.0/(6.0*V.kappa))*V.v0;
        PL_UNROLL
        for (int c = 1; c < NC; ++c) if (rdir<D>(c, axis) == dir) s = s + g[c];
        tem0 = dir == -1 ? 6.0*s/(1.0 - 3.0*ua) : 6.0*s/(1.0 + 3.0*ua);
    }
    const double wd = D == 2 ? 36.0 : 72.0;
    PL_UNROLL
    for (int c = 1; c < NC; ++c) {
        if (rdir<D>(c, axis) != -dir) continue;
        int nz = abs(rdir<D>(c, 0)) + abs(rdir<D>(c, 1)) + abs(rdir<D>(c, 2));
        g[c] = tem0*one_plus_3cu<D>(c, V.ux, V.uy, V.uz)/(nz == 1 ? 9.0 : wd);
    }
}

// ANS::iBoundaryConditionSetU (adjointnavierstokes.h:97-254).  v0,v1,v2 = prescribed ux,uy,uz.
//   rho0 = (-(2|4) eps -dir*u_a*((4|8) f_K0 + sum f_Kdiag) + sum_t 3 u_t sum_i c_t(K_i) f_Ki)/((3|6)(1 + dir*u_a)),  f_opp(K) = f_K + rho0
// 3-D: terms in x,y,z order; 2-D: normal term first.  The reference's 2-D y-edge version reads ux where uy is
// meant (adjointnavierstokes.h:134,139); reproduced.
template <int D, class PA> PL_HD void closure_ans_isetu(PA& f, int axis, int dir, const SiteVals& V) {
    int K[FaceK<D>::n];
    face_list<D>(axis, -dir, K);
    double u[3] = {V.v0, V.v1, V.v2};
    if (D == 2 && axis == 1) u[1] = u[0];
    const double ua = u[axis];
    double acc, rho0;
    if (D == 2) {
        acc = -2.0*V.eps;
        double tn = ua*weighted_face_sum<D>(f, K, 4.0);
        acc = dir == -1 ? acc + tn : acc - tn;
        const int b = 1 - axis;
        const double ut = axis == 1 ? u[0] : u[1];
        acc = acc + 3.0*ut*signed_diag_sum<D>(f, K, b);
        rho0 = dir == -1 ? acc/(3.0*(1.0 - ua)) : acc/(3.0*(1.0 + ua));
    } else {
        acc = -4.0*V.eps;
        PL_UNROLL
        for (int b = 0; b < 3; ++b) {
            if (b == axis) {
                double tn = ua*weighted_face_sum<D>(f, K, 8.0);
                acc = dir == -1 ? acc + tn : acc - tn;
            } else acc = acc + 3.0*u[b]*signed_diag_sum<D>(f, K, b);
        }
        rho0 = dir == -1 ? acc/(6.0*(1.0 - ua)) : acc/(6.0*(1.0 + ua));
    }
    double nv[FaceK<D>::n];
    PL_UNROLL
    for (int i = 0; i < FaceK<D>::n; ++i) nv[i] = f[K[i]] + rho0;
    PL_UNROLL
    for (int i = 0; i < FaceK<D>::n; ++i) f[ropp<D>(K[i])] = nv[i];
}

// ANS::iBoundaryConditionSetRho (adjointnavierstokes.h:258-392): rho0 = ((4|8) f_K0 + sum f_Kdiag)/(3|6), f_opp(K) = f_K - rho0
template <int D, class PA> PL_HD void closure_ans_isetrho(PA& f, int axis, int dir) {
    int K[FaceK<D>::n];
    face_list<D>(axis, -dir, K);
    const double rho0 = D == 2 ? weighted_face_sum<D>(f, K, 4.0)/3.0 : weighted_face_sum<D>(f, K, 8.0)/6.0;
    PL_UNROLL
    for (int i = 0; i < FaceK<D>::n; ++i) f[ropp<D>(K[i])] = f[K[i]] - rho0;
}

// AAD::iBoundaryConditionSetT (adjointadvection.h:154-300): every unknown (opposites of K) takes the same value
//   3-D: -(8 g_K0 + sum g_Kdiag)/12 - sum_t u_t*S_t/(4(1 - dir*3u_a)), tangential axes in cyclic order (a+1, a+2)
//   2-D: -(4(1 - dir*3u_a) g_K0 + sum (1 + 3 c.u) g_Kdiag)/(6(1 - dir*3u_a))
template <int D, class PA> PL_HD void closure_aad_isett(PA& g, int axis, int dir, const SiteVals& V) {
    int K[FaceK<D>::n];
    face_list<D>(axis, -dir, K);
    const double ua = pick(axis, V.ux, V.uy, V.uz);
    const double one3 = dir == -1 ? 1.0 + 3.0*ua : 1.0 - 3.0*ua;
    double r;
    if (D == 2) {
        double a = 4.0*one3*g[K[0]];
        PL_UNROLL
        for (int i = 1; i < 3; ++i) a = a + one_plus_3cu<D>(K[i], V.ux, V.uy, V.uz)*g[K[i]];
        r = -a/(6.0*one3);
    } else {
        r = -weighted_face_sum<D>(g, K, 8.0)/12.0;
        PL_UNROLL
        for (int t = 1; t <= 2; ++t) {
            int b = (axis + t)%3;
            r = r - pick(b, V.ux, V.uy, V.uz)*signed_diag_sum<D>(g, K, b)/(4.0*one3);
        }
    }
    PL_UNROLL
    for (int i = 0; i < FaceK<D>::n; ++i) g[ropp<D>(K[i])] = r;
}

// the bracket shared by AAD::iBoundaryConditionSetQ (adjointadvection.h:304-484) and the heat-source term of
// AAD::SensitivityTemperatureAtHeatSource (adjointadvection_avx.h:16-185):
//   (1 - dir*3u_a)*(lead + (4|8) g_K0 + sum g_Kdiag) + sum_t w_t u_t S_t,  tangential axes in cyclic order,
//   w_t = 3 except on the 3-D ymax, zmin and zmax faces where the reference writes the bare u_t
//   (adjointadvection.h:458-459,470-471,...; adjointadvection_avx.h:141-142,171-172,176-177); reproduced.
template <int D, class PA> PL_HD double aad_q_bracket(const PA& g, const int (&K)[FaceK<D>::n], int axis, int dir,
                                            const SiteVals& V, bool with_lead, double lead) {
    const double ua = pick(axis, V.ux, V.uy, V.uz);
    const double one3 = dir == -1 ? 1.0 + 3.0*ua : 1.0 - 3.0*ua;
    double s;
    if (with_lead) {
        s = lead + (D == 2 ? 4.0 : 8.0)*g[K[0]];
        PL_UNROLL
        for (int i = 1; i < FaceK<D>::n; ++i) s = s + g[K[i]];
    }
    else s = weighted_face_sum<D>(g, K, D == 2 ? 4.0 : 8.0);
    double acc = one3*s;
    if (D == 2) {
        const int b = 1 - axis;
        acc = acc + 3.0*pick(b, V.ux, V.uy, V.uz)*signed_diag_sum<D>(g, K, b);
    } else {
        const bool three = axis == 0 || (axis == 1 && dir == -1);
        PL_UNROLL
        for (int t = 1; t <= 2; ++t) {
            int b = (axis + t)%3;
            double ub = pick(b, V.ux, V.uy, V.uz);
            acc = three ? acc + 3.0*ub*signed_diag_sum<D>(g, K, b) : acc + ub*signed_diag_sum<D>(g, K, b);
        }
    }
    return acc;
}
template <int D, class PA> PL_HD void closure_aad_isetq(PA& g, int axis, int dir, const SiteVals& V) {
    int K[FaceK<D>::n];
    face_list<D>(axis, -dir, K);
    const double ua = pick(axis, V.ux, V.uy, V.uz);
    double acc = aad_q_bracket<D>(g, K, axis, dir, V, false, 0.0);
    acc = acc - (D == 2 ? 12.0 : 24.0)*V.eps;
    const double den = (D == 2 ? 6.0 : 12.0)*(dir == -1 ? 1.0 - 3.0*ua : 1.0 + 3.0*ua);
    const double r = acc/den;
    PL_UNROLL
    for (int i = 0; i < FaceK<D>::n; ++i) g[ropp<D>(K[i])] = r;
}

// AAD::iBoundaryConditionSetRho for D2Q9 (adjointadvection.h:488-575): f and g lattices together; the mask value
// selects the thermal closure the edge carries (1 = SetT, 2 = SetQ; adjointadvection.h:16-17).
template <class PA, class QA> PL_HD void closure_aad_isetrho2d(PA& f, const QA& g, int axis, int dir, int kind, const SiteVals& V) {
    int K[3];
    face_list<2>(axis, -dir, K);
    const double ua = axis == 0 ? V.ux : V.uy, ut = axis == 0 ? V.uy : V.ux;
    const double rho0 = -weighted_face_sum<2>(f, K, 4.0)/3.0;
    const double sd = signed_diag_sum<2>(g, K, 1 - axis);
    const double onep = dir == -1 ? 1.0 + 3.0*ua : 1.0 - 3.0*ua;   // 1 - dir*3u_a
    const double onem = dir == -1 ? 1.0 - 3.0*ua : 1.0 + 3.0*ua;   // 1 + dir*3u_a
    double flux0 = 0.0;
    if (kind == 1) flux0 = V.tem*ut*sd/(2.0*onep*V.rho);
    else if (kind == 2) flux0 = -V.tem*(weighted_face_sum<2>(g, K, 4.0)/3.0 + ut*sd/2.0)/(onem*V.rho);
    const double obj0 = V.eps*2.0*V.tem/(onem*V.rho);
    double nv[3];
    PL_UNROLL
    for (int i = 0; i < 3; ++i) nv[i] = f[K[i]] + rho0 + flux0 + obj0;
    PL_UNROLL
    for (int i = 0; i < 3; ++i) f[ropp<2>(K[i])] = nv[i];
}

// One closure application on a site.  p = populations of the lattice the closure acts on; q = the other lattice's
// populations at the same site (only BC_AAD_ISET_RHO reads it).  maskval = baked mask byte (non-zero).
// The closure bodies above are written over run-time (axis, dir) with table lookups per direction; on the device they are
// instantiated once per face with literal arguments, so that after inlining and unrolling every population index is a
// constant and each closure is a few dozen straight-line fp64 instructions in the order the generic loops define.
template <int D, class PA, class QA>
PL_HD void apply_closure_on(int type, int axis, int dir, int maskval, PA& p, const QA& q, const SiteVals& V) {
    switch (type) {
        case BC_BOUNCE: closure_bounce<D>(p, axis, dir, maskval, false); break;
        case BC_IBOUNCE: closure_bounce<D>(p, axis, dir, maskval, true); break;
        case BC_NS_SET_U: closure_ns<D>(p, axis, dir, V, false); break;
        case BC_NS_SET_RHO: closure_ns<D>(p, axis, dir, V, true); break;
        case BC_AD_SET_T: closure_ad<D>(p, axis, dir, V, false); break;
        case BC_AD_SET_Q: closure_ad<D>(p, axis, dir, V, true); break;
        case BC_ANS_ISET_U: closure_ans_isetu<D>(p, axis, dir, V); break;
        case BC_ANS_ISET_RHO: closure_ans_isetrho<D>(p, axis, dir); break;
        case BC_AAD_ISET_T: closure_aad_isett<D>(p, axis, dir, V); break;
        case BC_AAD_ISET_Q: closure_aad_isetq<D>(p, axis, dir, V); break;
        case BC_AAD_ISET_RHO: if constexpr (D == 2) closure_aad_isetrho2d(p, q, axis, dir, maskval, V); break;
        case BC_NSIN_SET_U: closure_nsin<D>(p, axis, dir, V, false); break;
        case BC_NSIN_SET_RHO: closure_nsin<D>(p, axis, dir, V, true); break;
        default: break;
    }
}
template <int D, class PA, class QA>
PL_HD void apply_closure(int type, int axis, int dir, int maskval, PA& p, const QA& q, const SiteVals& V) {
#ifdef __CUDA_ARCH__
    switch (2*axis + (dir > 0 ? 1 : 0)) {
        case 0: apply_closure_on<D>(type, 0, -1, maskval, p, q, V); break;
        case 1: apply_closure_on<D>(type, 0, 1, maskval, p, q, V); break;
        case 2: apply_closure_on<D>(type, 1, -1, maskval, p, q, V); break;
        case 3: apply_closure_on<D>(type, 1, 1, maskval, p, q, V); break;
        case 4: if constexpr (D == 3) apply_closure_on<D>(type, 2, -1, maskval, p, q, V); break;
        case 5: if constexpr (D == 3) apply_closure_on<D>(type, 2, 1, maskval, p, q, V); break;
        default: break;
    }
#else
    apply_closure_on<D>(type, axis, dir, maskval, p, q, V);
#endif
}

// heat-source boundary term of AAD::SensitivityTemperatureAtHeatSource on one plane site
// (adjointadvection_avx.h:16-185): ig = adjoint thermal snapshot at the site, V.v0 = qn, returns the increment of dfds.
template <int D, class PA> PL_HD double sens_heat_source_term(const PA& ig, int axis, int dir, const SiteVals& V, double dkds) {
    int K[FaceK<D>::n];
    face_list<D>(axis, -dir, K);
    const double ua = pick(axis, V.ux, V.uy, V.uz);
    const double e = aad_q_bracket<D>(ig, K, axis, dir, V, true, D == 2 ? -6.0 : -12.0);
    const double den = (D == 2 ? 36.0 : 72.0)*(dir == -1 ? 1.0 - 3.0*ua : 1.0 + 3.0*ua)*(V.kappa*V.kappa);
    return V.v0*dkds*e/den;
}

}  // namespace plb
