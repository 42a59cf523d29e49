// Site-local lattice-Boltzmann updates (moments, forces, equilibria, relaxation) for the NS / AD / ANS / AAD
// equations.  Everything is IEEE fp64 with FMA contraction disabled (-fmad=false) and follows the
// operation order of the reference so that results are bit-identical to its OpenMP+AVX build:
//   * "avx" order = the __m256d overloads in src/equation_avx/*.h (sites idx < 4*(nxyz/4)),
//   * "sc"  order = the scalar templates in src/equation/*.h (tail sites, InitialCondition, closures).
// Multiplications by the lattice constants 0/±1 are folded at compile time (exact).
#pragma once
#include "lbm_traits.cuh"

namespace plb {

// model feature flags
enum : unsigned {
    F_G = 1u,         // second (thermal) lattice g
    F_ADJ = 2u,       // adjoint equations
    F_NATCONV = 4u,   // buoyancy coupling
    F_BRINK = 8u,     // Brinkman force on f (forward models; the adjoint coupling force is always present)
    F_HEATEX = 16u,   // heat-exchange source on g
    F_MASSFLOW = 32u, // mass-flow objective source (adjoint, D2Q9)
    F_KFIELD = 64u,   // per-cell diffusivity field (else scalar)
    F_SNAP = 128u,    // model has a snapshot parameter (_g / _ig)
    F_INCOMP = 256u   // incompressible NS (NSin, nsincompressible.h): momentum moments, rho only in the rest term of feq; D2Q9, scalar templates only
};

template <int M> struct ModelFlags;
template <> struct ModelFlags<1> { static constexpr unsigned v = 0u; };
template <> struct ModelFlags<2> { static constexpr unsigned v = F_BRINK; };
template <> struct ModelFlags<3> { static constexpr unsigned v = F_G; };
template <> struct ModelFlags<4> { static constexpr unsigned v = F_G | F_NATCONV; };
template <> struct ModelFlags<5> { static constexpr unsigned v = F_G | F_BRINK | F_HEATEX; };
template <> struct ModelFlags<6> { static constexpr unsigned v = F_G | F_BRINK | F_KFIELD | F_SNAP; };
template <> struct ModelFlags<7> { static constexpr unsigned v = F_G | F_BRINK | F_NATCONV | F_KFIELD | F_SNAP; };
template <> struct ModelFlags<8> { static constexpr unsigned v = F_ADJ | F_BRINK; };
template <> struct ModelFlags<9> { static constexpr unsigned v = F_ADJ | F_G | F_HEATEX; };
template <> struct ModelFlags<10> { static constexpr unsigned v = F_ADJ | F_G | F_KFIELD | F_SNAP; };
template <> struct ModelFlags<11> { static constexpr unsigned v = F_ADJ | F_G | F_NATCONV | F_KFIELD | F_SNAP; };
template <> struct ModelFlags<12> { static constexpr unsigned v = F_ADJ | F_G | F_NATCONV | F_MASSFLOW | F_KFIELD | F_SNAP; };
template <> struct ModelFlags<13> { static constexpr unsigned v = F_INCOMP; };
template <> struct ModelFlags<14> { static constexpr unsigned v = F_INCOMP | F_BRINK; };

// kernel-side argument block of one collide (built on the host from pl_collide_args)
struct CollideParams {
    int issave;               // 0 = nothing is stored, 1 = every site stores its macros / snapshot, 2 = only the sites on closure planes do
    double omegaf, iomegaf;   // 1/(3 nu + 1/2), 1 - omegaf   (navierstokes_avx.h:151)
    double omegag, iomegag;   // scalar-diffusivity models
    double gx, gy, gz, tem0;
    double eicg[15];          // ei[c]*((cx*gx + cy*gy) + cz*gz)   (advection_avx.h:87-92), uniform per launch
    double cg[15];            // (cx*gx + cy*gy) + cz*gz           (scalar order needs it separately)
    const double *alpha, *kappa, *beta, *dirx, *diry, *dirz;
    double *rho, *ux, *uy, *uz, *tem, *qx, *qy, *qz;
    double *ip, *iux, *iuy, *iuz, *imx, *imy, *imz, *item, *iqx, *iqy, *iqz;
    double *snap;             // SoA [c][snap_pitch]
    size_t snap_pitch;
    int scalar_build;         // the caller was built WITHOUT _USE_AVX_DEFINES: every site runs the scalar templates of src/equation/*.h
                              // (all sites are "tail" sites, and the quirks of the tail code inside the *_avx.h files do not apply)
};

template <int D> PL_D double dot(double ax, double ay, double az, double bx, double by, double bz) {
    double s = ax*bx + ay*by;
    if constexpr (D == 3) s = s + az*bz;
    return s;
}

// ---------------------------------------------------------------------------------------------------------
// NS  (navierstokes.h:17-88, navierstokes_avx.h:24-91)
template <int D> PL_D void ns_macro(const double (&f)[LT<D>::nc], double& rho, double& ux, double& uy, double& uz) {
    rho = f[0]; ux = 0.0; uy = 0.0; uz = 0.0;
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        rho = rho + f[c];
        ux = sadd<LT<D>::cx(c)>(ux, f[c]);
        uy = sadd<LT<D>::cy(c)>(uy, f[c]);
        if constexpr (D == 3) uz = sadd<LT<D>::cz(c)>(uz, f[c]);
    });
    double inv = 1.0/rho;
    ux = ux*inv; uy = uy*inv; uz = uz*inv;
}
template <int D> PL_D void ns_eq_avx(double (&feq)[LT<D>::nc], double rho, double ux, double uy, double uz) {
    double a = 1.0 - 1.5*dot<D>(ux, uy, uz, ux, uy, uz);
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double cu = cdot<D, c>(ux, uy, uz);
        feq[c] = LT<D>::ei(c)*(rho*(a + (3.0*cu + 4.5*(cu*cu))));
    });
}
template <int D> PL_D void ns_eq_sc(double (&feq)[LT<D>::nc], double rho, double ux, double uy, double uz) {
    double uu = 1.0 - 1.5*dot<D>(ux, uy, uz, ux, uy, uz);
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double ciu = cdot<D, c>(ux, uy, uz);
        feq[c] = LT<D>::ei(c)*rho*(3.0*ciu + 4.5*ciu*ciu + uu);
    });
}
template <int D> PL_D void ns_brinkman(double (&f)[LT<D>::nc], double rho, double ux, double uy, double uz, double alpha) {
    double coef = 3.0*alpha*rho/(rho + alpha);
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        f[c] = f[c] - coef*LT<D>::ei(c)*cdot<D, c>(ux, uy, uz);
    });
}

// NSin  (nsincompressible.h:11-44): u is the momentum sum itself (no division by rho), rho enters feq only through the rest term.
// The reference has scalar templates only (no AVX overloads), so every site computes in this one order.
template <int D> PL_D void nsin_macro(const double (&f)[LT<D>::nc], double& rho, double& ux, double& uy, double& uz) {
    rho = f[0]; ux = 0.0; uy = 0.0; uz = 0.0;
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        rho = rho + f[c];
        ux = sadd<LT<D>::cx(c)>(ux, f[c]);
        uy = sadd<LT<D>::cy(c)>(uy, f[c]);
        if constexpr (D == 3) uz = sadd<LT<D>::cz(c)>(uz, f[c]);
    });
}
template <int D> PL_D void nsin_eq(double (&feq)[LT<D>::nc], double rho, double ux, double uy, double uz) {
    double rhouu = rho - 1.5*dot<D>(ux, uy, uz, ux, uy, uz);
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double ciu = cdot<D, c>(ux, uy, uz);
        feq[c] = LT<D>::ei(c)*(3.0*ciu + 4.5*ciu*ciu + rhouu);
    });
}

// ---------------------------------------------------------------------------------------------------------
// AD  (advection.h:18-95, advection_avx.h:25-102)
template <int D> PL_D void ad_macro(const double (&g)[LT<D>::nc], double ux, double uy, double uz, double omegag,
                                    double& tem, double& qx, double& qy, double& qz) {
    tem = g[0]; qx = 0.0; qy = 0.0; qz = 0.0;
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        tem = tem + g[c];
        qx = sadd<LT<D>::cx(c)>(qx, g[c]);
        qy = sadd<LT<D>::cy(c)>(qy, g[c]);
        if constexpr (D == 3) qz = sadd<LT<D>::cz(c)>(qz, g[c]);
    });
    double coef = 1.0 - 0.5*omegag;
    qx = coef*(qx - tem*ux);
    qy = coef*(qy - tem*uy);
    if constexpr (D == 3) qz = coef*(qz - tem*uz);
}
template <int D> PL_D void ad_eq_avx(double (&geq)[LT<D>::nc], double tem, double ux, double uy, double uz) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        geq[c] = LT<D>::ei(c)*(tem*(1.0 + 3.0*cdot<D, c>(ux, uy, uz)));
    });
}
template <int D> PL_D void ad_eq_sc(double (&geq)[LT<D>::nc], double tem, double ux, double uy, double uz) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        geq[c] = LT<D>::ei(c)*tem*(1.0 + 3.0*cdot<D, c>(ux, uy, uz));
    });
}
template <int D, bool SC> PL_D void ad_natconv(double (&f)[LT<D>::nc], double tem, const CollideParams& P) {
    if constexpr (!SC) {
        double coef = 3.0*(tem - P.tem0);
        sfor<1, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = f[c] + coef*P.eicg[c]; });
    } else {
        double dt = tem - P.tem0;
        sfor<1, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; f[c] = f[c] + 3.0*LT<D>::ei(c)*P.cg[c]*dt; });
    }
}
template <int D, bool SC> PL_D void ad_heatex(double (&g)[LT<D>::nc], double tem, double beta) {
    double coef;
    if constexpr (!SC) coef = beta*((1.0 - tem)/(1.0 + beta));
    else coef = beta*(1.0 - tem)/(1.0 + beta);
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; g[c] = g[c] + LT<D>::ei(c)*coef; });
}

// ---------------------------------------------------------------------------------------------------------
// ANS  (adjointnavierstokes.h:17-93, adjointnavierstokes_avx.h:27-110)
template <int D, bool SC> PL_D void ans_macro(const double (&f)[LT<D>::nc], double ux, double uy, double uz,
                                              double& ip, double& iux, double& iuy, double& iuz, double& imx, double& imy, double& imz) {
    const double uu = dot<D>(ux, uy, uz, ux, uy, uz);
    if constexpr (!SC) {
        ip = 0.0; iux = 0.0; iuy = 0.0; iuz = 0.0; imx = 0.0; imy = 0.0; imz = 0.0;
        const double a2 = 1.0 - 1.5*uu;   // 2-D association (adjointnavierstokes_avx.h:34-38)
        const double b3 = 1.5*uu;         // 3-D association (adjointnavierstokes_avx.h:58-64)
        sfor<0, LT<D>::nc>([&](auto C) {
            constexpr int c = decltype(C)::value;
            constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
            double fei = f[c]*LT<D>::ei(c);
            double cu = cdot<D, c>(ux, uy, uz);
            double c3 = 3.0*cu;
            if constexpr (D == 2) ip = ip + fei*(a2 + (c3 + 4.5*(cu*cu)));
            else ip = ip + fei*(1.0 + (c3 + (4.5*(cu*cu) - b3)));
            // cx + ((3*(cu*cx)) - ux)
            iux = iux + fei*((double)X + ((X == 0 ? 0.0 : (X > 0 ? c3 : -c3)) - ux));
            iuy = iuy + fei*((double)Y + ((Y == 0 ? 0.0 : (Y > 0 ? c3 : -c3)) - uy));
            if constexpr (D == 3) iuz = iuz + fei*((double)Z + ((Z == 0 ? 0.0 : (Z > 0 ? c3 : -c3)) - uz));
            imx = sadd<X>(imx, fei);
            imy = sadd<Y>(imy, fei);
            if constexpr (D == 3) imz = sadd<Z>(imz, fei);
        });
    } else {
        double f0ei = f[0]*LT<D>::ei(0);
        ip = f0ei*(1.0 - 1.5*uu);
        double nf0ei = -f[0]*LT<D>::ei(0);
        iux = nf0ei*ux; iuy = nf0ei*uy; iuz = nf0ei*uz;
        imx = 0.0; imy = 0.0; imz = 0.0;
        sfor<1, LT<D>::nc>([&](auto C) {
            constexpr int c = decltype(C)::value;
            constexpr int X = LT<D>::cx(c), Y = LT<D>::cy(c), Z = LT<D>::cz(c);
            double ciu = cdot<D, c>(ux, uy, uz);
            double fei = f[c]*LT<D>::ei(c);
            ip = ip + fei*(1.0 + 3.0*ciu + 4.5*ciu*ciu - 1.5*uu);
            double c3 = 3.0*ciu;
            iux = iux + fei*((double)X + (X == 0 ? 0.0 : (X > 0 ? c3 : -c3)) - ux);
            iuy = iuy + fei*((double)Y + (Y == 0 ? 0.0 : (Y > 0 ? c3 : -c3)) - uy);
            if constexpr (D == 3) iuz = iuz + fei*((double)Z + (Z == 0 ? 0.0 : (Z > 0 ? c3 : -c3)) - uz);
            imx = sadd<X>(imx, fei);
            imy = sadd<Y>(imy, fei);
            if constexpr (D == 3) imz = sadd<Z>(imz, fei);
        });
    }
}
// feq_c = ip + 3*((iux*(cx-ux) + iuy*(cy-uy)) + iuz*(cz-uz))   (same in both orders)
template <int D> PL_D void ans_eq(double (&feq)[LT<D>::nc], double ux, double uy, double uz, double ip, double iux, double iuy, double iuz) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double s = iux*((double)LT<D>::cx(c) - ux) + iuy*((double)LT<D>::cy(c) - uy);
        if constexpr (D == 3) s = s + iuz*((double)LT<D>::cz(c) - uz);
        feq[c] = ip + 3.0*s;
    });
}
template <int D, bool SC> PL_D void ans_brinkman(double (&f)[LT<D>::nc], double rho, double ux, double uy, double uz,
                                                 double imx, double imy, double imz, double alpha) {
    double coef;
    if constexpr (!SC) {
        coef = 3.0*(alpha/(rho + alpha));
        f[0] = f[0] + coef*dot<D>(ux, uy, uz, imx, imy, imz);
    } else {
        coef = 3.0*alpha/(rho + alpha);
        f[0] = f[0] - (-coef*dot<D>(ux, uy, uz, imx, imy, imz));
    }
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double s = ((double)LT<D>::cx(c) - ux)*imx + ((double)LT<D>::cy(c) - uy)*imy;
        if constexpr (D == 3) s = s + ((double)LT<D>::cz(c) - uz)*imz;
        f[c] = f[c] - coef*s;
    });
}

// ---------------------------------------------------------------------------------------------------------
// AAD  (adjointadvection.h:24-150, adjointadvection_avx.h:203-322)
template <int D> PL_D void aad_macro(const double (&g)[LT<D>::nc], double& item, double& iqx, double& iqy, double& iqz) {
    item = LT<D>::ei(0)*g[0]; iqx = 0.0; iqy = 0.0; iqz = 0.0;
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double gei = LT<D>::ei(c)*g[c];
        item = item + gei;
        iqx = sadd<LT<D>::cx(c)>(iqx, gei);
        iqy = sadd<LT<D>::cy(c)>(iqy, gei);
        if constexpr (D == 3) iqz = sadd<LT<D>::cz(c)>(iqz, gei);
    });
}
template <int D> PL_D double aad_eq(double item, double iqx, double iqy, double iqz, double ux, double uy, double uz) {
    return item + 3.0*dot<D>(iqx, iqy, iqz, ux, uy, uz);
}
// coupling force of the adjoint flow equation (adjointadvection_avx.h:250-279; the scalar version gives the same values)
template <int D> PL_D void aad_brinkman(double (&f)[LT<D>::nc], double rho, double ux, double uy, double uz, double imx, double imy, double imz,
                                        double tem, double iqx, double iqy, double iqz, double omegag, double alpha) {
    double coef = 3.0/(rho + alpha);
    double kx = tem*iqx*omegag - alpha*imx;
    double ky = tem*iqy*omegag - alpha*imy;
    double kz = 0.0;
    if constexpr (D == 3) kz = tem*iqz*omegag - alpha*imz;
    f[0] = f[0] - coef*dot<D>(kx, ky, kz, ux, uy, uz);
    sfor<1, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double s = kx*((double)LT<D>::cx(c) - ux) + ky*((double)LT<D>::cy(c) - uy);
        if constexpr (D == 3) s = s + kz*((double)LT<D>::cz(c) - uz);
        f[c] = f[c] + coef*s;
    });
}
template <int D, bool SC> PL_D void aad_heatex(double (&g)[LT<D>::nc], double item, double beta) {
    double coef;
    if constexpr (!SC) coef = beta*((1.0 + item)/(1.0 + beta));
    else coef = beta*(1.0 + item)/(1.0 + beta);
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; g[c] = g[c] - coef; });
}
template <int D> PL_D void aad_natconv(double (&g)[LT<D>::nc], double imx, double imy, double imz, const CollideParams& P) {
    double coef = 3.0*dot<D>(imx, imy, imz, P.gx, P.gy, P.gz);
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; g[c] = g[c] + coef; });
}
// adjointadvection_avx.h:308-313 (the scalar version adjointadvection.h:136-141 gives the same values)
template <int D> PL_D void aad_massflow(double (&f)[LT<D>::nc], double rho, double ux, double uy, double uz, double dx, double dy, double dz) {
    sfor<0, LT<D>::nc>([&](auto C) {
        constexpr int c = decltype(C)::value;
        double s = ((double)LT<D>::cx(c) - ux)*dx + ((double)LT<D>::cy(c) - uy)*dy;
        if constexpr (D == 3) s = s + ((double)LT<D>::cz(c) - uz)*dz;
        f[c] = f[c] - s/rho;
    });
}

// ---------------------------------------------------------------------------------------------------------
template <int D> PL_D void relax(double (&p)[LT<D>::nc], const double (&eq)[LT<D>::nc], double omega, double iomega) {
    sfor<0, LT<D>::nc>([&](auto C) { constexpr int c = decltype(C)::value; p[c] = iomega*p[c] + omega*eq[c]; });
}

// One site of any Macro*Collide*: f (and g) hold the pre-collision populations on entry and the
// post-collision ones on exit.  FL = ModelFlags, SC = scalar (tail) order.
// `save`: store the macroscopic fields (and the thermal snapshot) of this site — P.issave for every site of a saving step; on a
// step whose outputs nobody can observe (pl_plan_advance_observed) only the sites a closure of the plan reads them at.
template <int D, unsigned FL, bool SC>
PL_D void collide_site(double (&f)[LT<D>::nc], double (&g)[LT<D>::nc], const CollideParams& P, size_t idx, bool save) {
    constexpr int NC = LT<D>::nc;
    constexpr bool G = (FL & F_G) != 0;
    double omegag = P.omegag, iomegag = P.iomegag;
    if constexpr (G && (FL & F_KFIELD)) {
        omegag = 1.0/(3.0*P.kappa[idx] + 0.5);
        iomegag = 1.0 - omegag;
    }
    auto snapshot = [&]() {
        if constexpr (G && (FL & F_SNAP)) {
            if (P.snap) sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; P.snap[(size_t)c*P.snap_pitch + idx] = g[c]; });
        }
    };
    if constexpr (!(FL & F_ADJ)) {
        // ---- forward: advection_avx.h:1011-1066 (and the variants :106-883), navierstokes_avx.h:148-329
        double rho, ux, uy, uz, tem = 0.0, qx = 0.0, qy = 0.0, qz = 0.0;
        constexpr bool INC = (FL & F_INCOMP) != 0;      // NSin::MacroCollide / MacroBrinkmanCollide (nsincompressible.h:158-210)
        auto fmacro = [&]() { if constexpr (INC) nsin_macro<D>(f, rho, ux, uy, uz); else ns_macro<D>(f, rho, ux, uy, uz); };
        fmacro();
        if constexpr (G) ad_macro<D>(g, ux, uy, uz, omegag, tem, qx, qy, qz);
        // quirk: the 2-D scalar tail of NS::MacroBrinkmanCollide stores the macros before the force (navierstokes_avx.h:246-254)
        // (the scalar template proper stores them after the force like every other version, navierstokes.h:494-503)
        constexpr bool early = SC && D == 2 && FL == F_BRINK;
        if constexpr (early) { if (save && !P.scalar_build) { P.rho[idx] = rho; P.ux[idx] = ux; P.uy[idx] = uy; } }
        if constexpr ((FL & F_NATCONV) != 0) ad_natconv<D, SC>(f, tem, P);
        if constexpr ((FL & F_BRINK) != 0) ns_brinkman<D>(f, rho, ux, uy, uz, P.alpha[idx]);
        if constexpr ((FL & (F_NATCONV | F_BRINK)) != 0) fmacro();
        if constexpr ((FL & F_HEATEX) != 0) ad_heatex<D, SC>(g, tem, P.beta[idx]);
        if constexpr (G && (FL & (F_NATCONV | F_BRINK)) != 0) ad_macro<D>(g, ux, uy, uz, omegag, tem, qx, qy, qz);
        if (save) {
            if constexpr (!early) {
                P.rho[idx] = rho; P.ux[idx] = ux; P.uy[idx] = uy;
                if constexpr (D == 3) P.uz[idx] = uz;
            } else {
                if (P.scalar_build) { P.rho[idx] = rho; P.ux[idx] = ux; P.uy[idx] = uy; }
            }
            if constexpr (G) {
                P.tem[idx] = tem; P.qx[idx] = qx; P.qy[idx] = qy;
                if constexpr (D == 3) P.qz[idx] = qz;
                snapshot();
            }
        }
        double eq[NC];
        if constexpr (INC) nsin_eq<D>(eq, rho, ux, uy, uz);
        else if constexpr (SC) ns_eq_sc<D>(eq, rho, ux, uy, uz);
        else ns_eq_avx<D>(eq, rho, ux, uy, uz);
        relax<D>(f, eq, P.omegaf, P.iomegaf);
        if constexpr (G) {
            if constexpr (SC) ad_eq_sc<D>(eq, tem, ux, uy, uz); else ad_eq_avx<D>(eq, tem, ux, uy, uz);
            relax<D>(g, eq, omegag, iomegag);
        }
    } else {
        // ---- adjoint: adjointadvection_avx.h:893-952 (and variants), adjointnavierstokes_avx.h:114-259
        const double rho = P.rho[idx], ux = P.ux[idx], uy = P.uy[idx];
        double uz = 0.0;
        if constexpr (D == 3) uz = P.uz[idx];
        double ip, iux, iuy, iuz, imx, imy, imz, item = 0.0, iqx = 0.0, iqy = 0.0, iqz = 0.0;
        ans_macro<D, SC>(f, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz);
        if constexpr (G) aad_macro<D>(g, item, iqx, iqy, iqz);
        if constexpr ((FL & F_MASSFLOW) != 0) aad_massflow<D>(f, rho, ux, uy, uz, P.dirx[idx], P.diry[idx], D == 3 ? P.dirz[idx] : 0.0);
        if constexpr (G) aad_brinkman<D>(f, rho, ux, uy, uz, imx, imy, imz, P.tem[idx], iqx, iqy, iqz, omegag, P.alpha[idx]);
        else ans_brinkman<D, SC>(f, rho, ux, uy, uz, imx, imy, imz, P.alpha[idx]);
        ans_macro<D, SC>(f, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz);
        if constexpr ((FL & F_HEATEX) != 0) aad_heatex<D, SC>(g, item, P.beta[idx]);
        if constexpr ((FL & F_NATCONV) != 0) aad_natconv<D>(g, imx, imy, imz, P);
        if constexpr (G && (FL & (F_HEATEX | F_NATCONV)) != 0) aad_macro<D>(g, item, iqx, iqy, iqz);
        if (save) {
            P.ip[idx] = ip; P.iux[idx] = iux; P.iuy[idx] = iuy; P.imx[idx] = imx; P.imy[idx] = imy;
            if constexpr (D == 3) {
                // quirk: the 3-D scalar tail of AAD::MacroBrinkmanCollideForceConvection does not store _iuz (adjointadvection_avx.h:725-735)
                constexpr bool skip_iuz = SC && FL == ModelFlags<10>::v;
                if constexpr (!skip_iuz) P.iuz[idx] = iuz;
                P.imz[idx] = imz;
            }
            if constexpr (G) {
                P.item[idx] = item; P.iqx[idx] = iqx; P.iqy[idx] = iqy;
                if constexpr (D == 3) P.iqz[idx] = iqz;
                snapshot();
            }
        }
        double eq[NC];
        ans_eq<D>(eq, ux, uy, uz, ip, iux, iuy, iuz);
        relax<D>(f, eq, P.omegaf, P.iomegaf);
        if constexpr (G) {
            double ge = aad_eq<D>(item, iqx, iqy, iqz, ux, uy, uz);
            sfor<0, NC>([&](auto C) { constexpr int c = decltype(C)::value; g[c] = iomegag*g[c] + omegag*ge; });
        }
    }
}

}  // namespace plb
