// Sensitivity accumulation of the adjoint equations: ANS::SensitivityBrinkman (adjointnavierstokes_avx.h:262-293),
// AAD::SensitivityHeatExchange (adjointadvection_avx.h:1257-1300), AAD::SensitivityBrinkmanDiffusivity
// (:1302-1401) and the volume part of AAD::SensitivityTemperatureAtHeatSource (:1403-1513; its boundary part is
// sens_heat_source_term in lbm_closures.cuh).  One value of dfds per site; "packed" sites (idx < 4*(nxyz/4)) follow
// the operation order of the __m256d code, tail sites the scalar order.
#pragma once
#include "lbm_closures.cuh"

namespace plb {

enum : int { SENS_ANS_BRINKMAN = 1, SENS_AAD_HEATEX = 2, SENS_AAD_BRINKMAN_DIFF = 3 };

struct SensSite {
    double dfds, ux, uy, uz, imx, imy, imz, dads;
    double tem, item, iqx, iqy, iqz, kappa, dkds, dbds;
};

// dfds + 3 dads (u . im)
template <int D, bool SC> PL_HD double sens_brinkman(const SensSite& s) {
    if constexpr (!SC) {
        double d = D == 3 ? s.ux*s.imx + (s.uy*s.imy + s.uz*s.imz) : s.ux*s.imx + s.uy*s.imy;
        return s.dfds + 3.0*(s.dads*d);
    } else {
        double d = s.ux*s.imx + s.uy*s.imy;
        if (D == 3) d = d + s.uz*s.imz;
        return s.dfds + 3.0*s.dads*d;
    }
}
// dfds + (3 dads (u . im) - dbds (1 - T)(1 + iT))      (both orders agree: adjointadvection_avx.h:1272,1277)
template <int D> PL_HD double sens_heatex(const SensSite& s) {
    double d = s.ux*s.imx + s.uy*s.imy;
    if (D == 3) d = d + s.uz*s.imz;
    return s.dfds + (3.0*s.dads*d - s.dbds*(1.0 - s.tem)*(1.0 + s.item));
}
// Brinkman term, then the diffusivity term  -3 dkds (sum_c g_c ig_c - T (iT + 3 u . iq))/(3 kappa + 1/2)^2
template <int D, bool SC> PL_HD double sens_brinkman_diffusivity(const SensSite& s, const double (&g)[LT<D>::nc], const double (&ig)[LT<D>::nc]) {
    double v = sens_brinkman<D, SC>(s);
    double sumg = 0.0;
    for (int c = 0; c < LT<D>::nc; ++c) sumg = sumg + g[c]*ig[c];
    if constexpr (!SC) {
        const double taug = 3.0*s.kappa + 0.5;
        double d = D == 3 ? s.ux*s.iqx + (s.uy*s.iqy + s.uz*s.iqz) : s.ux*s.iqx + s.uy*s.iqy;
        return v - (3.0*(s.dkds*(sumg - s.tem*(s.item + 3.0*d))))/(taug*taug);
    } else {
        const double taug = 3.0*s.kappa + 0.5;
        double d = s.ux*s.iqx + s.uy*s.iqy;
        if (D == 3) d = d + s.uz*s.iqz;
        return v + -3.0/(taug*taug)*s.dkds*(sumg - s.tem*(s.item + 3.0*d));
    }
}

#ifdef __CUDACC__
struct SensArgs {
    int kind;
    double* dfds;
    const double *ux, *uy, *uz, *imx, *imy, *imz, *dads, *tem, *item, *iqx, *iqy, *iqz;
    const double *gsnap, *igsnap;   // SoA [c][pitch]
    const double *kappa, *dkds, *dbds;
    size_t pitch;
};
template <int D>
__global__ void __launch_bounds__(256) k_sensitivity(Geom G, SensArgs A) {
    long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= G.nxyz) return;
    SensSite s{};
    s.dfds = A.dfds[idx]; s.ux = A.ux[idx]; s.uy = A.uy[idx]; s.imx = A.imx[idx]; s.imy = A.imy[idx]; s.dads = A.dads[idx];
    if (D == 3) { s.uz = A.uz[idx]; s.imz = A.imz[idx]; }
    const bool tail = idx >= G.npacked;
    double out;
    if (A.kind == SENS_ANS_BRINKMAN) out = tail ? sens_brinkman<D, true>(s) : sens_brinkman<D, false>(s);
    else if (A.kind == SENS_AAD_HEATEX) {
        s.tem = A.tem[idx]; s.item = A.item[idx]; s.dbds = A.dbds[idx];
        out = sens_heatex<D>(s);
    } else {
        s.tem = A.tem[idx]; s.item = A.item[idx]; s.iqx = A.iqx[idx]; s.iqy = A.iqy[idx];
        if (D == 3) s.iqz = A.iqz[idx];
        s.kappa = A.kappa[idx]; s.dkds = A.dkds[idx];
        double g[LT<D>::nc], ig[LT<D>::nc];
        #pragma unroll
        for (int c = 0; c < LT<D>::nc; ++c) { g[c] = A.gsnap[(size_t)c*A.pitch + idx]; ig[c] = A.igsnap[(size_t)c*A.pitch + idx]; }
        out = tail ? sens_brinkman_diffusivity<D, true>(s, g, ig) : sens_brinkman_diffusivity<D, false>(s, g, ig);
    }
    A.dfds[idx] = out;
}
#endif

}  // namespace plb
