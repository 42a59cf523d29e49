// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Op-level C wrapper around the UNMODIFIED PANSLBM2 reference headers (OpenMP + AVX build).
// Compiled only where /root/reference exists (see oracle/Makefile), output goes to
// oracle/_ref/libpanslbm_ref.so. No reference source is copied: this file only #includes the
// headers where they lie and forwards every call. It exists so that
//   (1) the plain-C restatement in oracle/lbm_oracle.c can be pinned function by function, and
//   (2) bench.py --impl reference / cpu_baseline can time the reference's own CPU path.
//
// Conventions shared with oracle/lbm_oracle.c (orc_*) so that tests can drive both identically:
//   * populations are exchanged in the reference layout  f0[nxyz], f[(nc-1)*idx + (c-1)]
//     (src/particle/d3q15.h:142-144, d2q9.h:86-88)
//   * boundary predicates / values are DENSE GLOBAL arrays indexed g = i + lx*(j + ly*k)
//     (the lambdas handed to the reference read them with the global coordinates the
//      reference passes in, e.g. src/equation/navierstokes.h:155-157)
//   * the two lattice headers cannot share a translation unit (both define BARRIER/MIRROR in an
//     anonymous namespace, d2q9.h:19-22 / d3q15.h:19-22), so this file is compiled twice:
//     -DDIM=2 -> libpanslbm_ref2d.so (D2Q9), -DDIM=3 -> libpanslbm_ref3d.so (D3Q15).
//     For DIM 2 every *z pointer argument is ignored.
// -DREF_SCALAR: the build of a program that leaves _USE_AVX_DEFINES out (production/nsopt.cpp:2): the scalar templates of
// src/equation/*.h at every site -> libpanslbm_ref{2d,3d}_scalar.so (checker of pl_set_scalar_order).
#ifndef REF_SCALAR
#define _USE_AVX_DEFINES
#endif
#include <cmath>
#include <cstring>
#include <chrono>
#include <vector>
#include <omp.h>

#if DIM == 2
#include "src/particle/d2q9.h"
typedef PANSLBM2::D2Q9<double> PT;
#define Z(x)
#define ZL(x)
#define IJK int i, int j
#define GIDX g(i, j)
#else
#include "src/particle/d3q15.h"
typedef PANSLBM2::D3Q15<double> PT;
#define Z(x) x,
#define ZL(x) , x
#define IJK int i, int j, int k
#define GIDX g(i, j, k)
#endif
#include "src/equation/navierstokes.h"
#include "src/equation/advection.h"
#include "src/equation/adjointnavierstokes.h"
#include "src/equation/adjointadvection.h"
#if DIM == 2
#include "src/equation/nsincompressible.h"
#endif
#include "src/utility/residual.h"
#include "src/utility/normalize.h"
#include "src/utility/densityfilter.h"
#include "src/utility/heavisidefilter.h"

using namespace PANSLBM2;

namespace {
struct Lat {
    PT* p;
    int lx, ly, lz;
};
inline Lat* L(void* h) { return static_cast<Lat*>(h); }
struct GI {  // dense global indexer
    int lx, ly;
    int operator()(int i, int j) const { return i + lx*j; }
    int operator()(int i, int j, int k) const { return i + lx*(j + ly*k); }
};
#define GDEF(l) GI g{(l)->lx, (l)->ly}
#define FV(arr) [=](IJK) { return arr[GIDX]; }
#define FM(arr) [=](IJK) { return arr[GIDX] != 0; }
}  // namespace

extern "C" {

int ref_dim() { return DIM; }

void* ref_lattice_create(int lx, int ly, int lz, int peid, int mx, int my, int mz) {
#if DIM == 2
    return new Lat{new PT(lx, ly, peid, mx, my), lx, ly, 1};
#else
    return new Lat{new PT(lx, ly, lz, peid, mx, my, mz), lx, ly, lz};
#endif
}
void ref_lattice_destroy(void* h) { delete L(h)->p; delete L(h); }
// out[0..16] = lx ly lz PEid mx my mz PEx PEy PEz nx ny nz nxyz offsetx offsety offsetz ; out[17]=nc
void ref_lattice_info(void* h, int* out) {
    PT* p = L(h)->p;
    int v[18] = {p->lx, p->ly, p->lz, p->PEid, p->mx, p->my, p->mz, p->PEx, p->PEy, p->PEz, p->nx, p->ny, p->nz, p->nxyz, p->offsetx, p->offsety, p->offsetz, PT::nc};
    std::memcpy(out, v, sizeof(v));
}
void ref_lattice_get(void* h, double* f0, double* f) {
    PT* p = L(h)->p;
    std::memcpy(f0, p->f0, sizeof(double)*p->nxyz);
    std::memcpy(f, p->f, sizeof(double)*(size_t)p->nxyz*(PT::nc - 1));
}
void ref_lattice_set(void* h, const double* f0, const double* f) {
    PT* p = L(h)->p;
    std::memcpy(p->f0, f0, sizeof(double)*p->nxyz);
    std::memcpy(p->f, f, sizeof(double)*(size_t)p->nxyz*(PT::nc - 1));
}

//---------------------------------------------------------------- particle ops
void ref_stream(void* h) { L(h)->p->Stream(); }
void ref_istream(void* h) { L(h)->p->iStream(); }
void ref_smooth_corner(void* h) { L(h)->p->SmoothCorner(); }
void ref_bc(void* h, const int* bct, int inverse) {
    GDEF(L(h));
    if (inverse) L(h)->p->iBoundaryCondition(FV(bct)); else L(h)->p->BoundaryCondition(FV(bct));
}
// single plane: axis 0/1/2, global coordinate, direction -1/+1
void ref_bc_plane(void* h, int axis, int coord, int dir, const int* bct, int inverse) {
    GDEF(L(h));
    PT* p = L(h)->p;
#if DIM == 2
    if (axis == 0) { if (inverse) p->iBoundaryConditionAlongXEdge(coord, dir, FV(bct)); else p->BoundaryConditionAlongXEdge(coord, dir, FV(bct)); }
    else           { if (inverse) p->iBoundaryConditionAlongYEdge(coord, dir, FV(bct)); else p->BoundaryConditionAlongYEdge(coord, dir, FV(bct)); }
#else
    if (axis == 0)      { if (inverse) p->iBoundaryConditionAlongXFace(coord, dir, FV(bct)); else p->BoundaryConditionAlongXFace(coord, dir, FV(bct)); }
    else if (axis == 1) { if (inverse) p->iBoundaryConditionAlongYFace(coord, dir, FV(bct)); else p->BoundaryConditionAlongYFace(coord, dir, FV(bct)); }
    else                { if (inverse) p->iBoundaryConditionAlongZFace(coord, dir, FV(bct)); else p->BoundaryConditionAlongZFace(coord, dir, FV(bct)); }
#endif
}

//---------------------------------------------------------------- NS
void ref_ns_init(void* h, const double* rho, const double* ux, const double* uy, const double* uz) {
    NS::InitialCondition(*L(h)->p, rho, ux, uy ZL(uz));
}
void ref_ns_macro_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    NS::MacroCollide(*L(h)->p, rho, ux, uy, Z(uz) nu, issave != 0);
}
void ref_ns_macro_brinkman_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    NS::MacroBrinkmanCollide(*L(h)->p, rho, ux, uy, Z(uz) nu, alpha, issave != 0);
}
void ref_ns_bc_set_u(void* h, const double* uxg, const double* uyg, const double* uzg, const int* mask) {
    GDEF(L(h));
    NS::BoundaryConditionSetU(*L(h)->p, FV(uxg), FV(uyg), Z(FV(uzg)) FM(mask));
}
// v0 = rho, v1 = _usbc, v2 = _utbc exactly as the reference names them (navierstokes.h:596-612)
void ref_ns_bc_set_rho(void* h, const double* v0, const double* v1, const double* v2, const int* mask) {
    GDEF(L(h));
    NS::BoundaryConditionSetRho(*L(h)->p, FV(v0), FV(v1), Z(FV(v2)) FM(mask));
}

//---------------------------------------------------------------- NSin (D2Q9 only, scalar templates only: nsincompressible.h)
#if DIM == 2
void ref_nsin_init(void* h, const double* rho, const double* ux, const double* uy, const double* uz) {
    NSin::InitialCondition(*L(h)->p, rho, ux, uy);
}
void ref_nsin_macro_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    NSin::MacroCollide(*L(h)->p, rho, ux, uy, nu, issave != 0);
}
void ref_nsin_macro_brinkman_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    NSin::MacroBrinkmanCollide(*L(h)->p, rho, ux, uy, nu, alpha, issave != 0);
}
void ref_nsin_bc_set_u(void* h, const double* uxg, const double* uyg, const double* uzg, const int* mask) {
    GDEF(L(h));
    NSin::BoundaryConditionSetU(*L(h)->p, FV(uxg), FV(uyg), FM(mask));
}
// NSin::BoundaryConditionSetRho itself cannot be instantiated (it calls a misspelt helper, nsincompressible.h:238): the four
// edge calls it stands for, in its order.  v0 = rho, v1 = _usbc.
void ref_nsin_bc_set_rho(void* h, const double* v0, const double* v1, const double* v2, const int* mask) {
    GDEF(L(h));
    PT& p = *L(h)->p;
    NSin::BoundaryConditionSetRhoAlongXEdge(p, 0, -1, FV(v0), FV(v1), FM(mask));
    NSin::BoundaryConditionSetRhoAlongXEdge(p, p.lx - 1, 1, FV(v0), FV(v1), FM(mask));
    NSin::BoundaryConditionSetRhoAlongYEdge(p, 0, -1, FV(v0), FV(v1), FM(mask));
    NSin::BoundaryConditionSetRhoAlongYEdge(p, p.ly - 1, 1, FV(v0), FV(v1), FM(mask));
}
#endif

//---------------------------------------------------------------- AD
void ref_ad_init(void* hg, const double* tem, const double* ux, const double* uy, const double* uz) {
    AD::InitialCondition(*L(hg)->p, tem, ux, uy ZL(uz));
}
void ref_ad_macro_collide_force_convection(void* hf, double* rho, double* ux, double* uy, double* uz, double nu,
                                           void* hg, double* tem, double* qx, double* qy, double* qz, double diffusivity, int issave) {
    AD::MacroCollideForceConvection(*L(hf)->p, rho, ux, uy, Z(uz) nu, *L(hg)->p, tem, qx, qy, Z(qz) diffusivity, issave != 0);
}
void ref_ad_macro_collide_natural_convection(void* hf, double* rho, double* ux, double* uy, double* uz, double nu,
                                             void* hg, double* tem, double* qx, double* qy, double* qz, double diffusivity,
                                             double gx, double gy, double gz, double tem0, int issave) {
    AD::MacroCollideNaturalConvection(*L(hf)->p, rho, ux, uy, Z(uz) nu, *L(hg)->p, tem, qx, qy, Z(qz) diffusivity, gx, gy, Z(gz) tem0, issave != 0);
}
void ref_ad_macro_brinkman_collide_heat_exchange(void* hf, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
                                                 void* hg, double* tem, double* qx, double* qy, double* qz, const double* beta, double diffusivity, int issave) {
    AD::MacroBrinkmanCollideHeatExchange(*L(hf)->p, rho, ux, uy, Z(uz) alpha, nu, *L(hg)->p, tem, qx, qy, Z(qz) beta, diffusivity, issave != 0);
}
void ref_ad_macro_brinkman_collide_force_convection(void* hf, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
                                                    void* hg, double* tem, double* qx, double* qy, double* qz, const double* diffusivity, int issave, double* gsnap) {
    AD::MacroBrinkmanCollideForceConvection(*L(hf)->p, rho, ux, uy, Z(uz) alpha, nu, *L(hg)->p, tem, qx, qy, Z(qz) diffusivity, issave != 0, gsnap);
}
void ref_ad_macro_brinkman_collide_natural_convection(void* hf, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
                                                      void* hg, double* tem, double* qx, double* qy, double* qz, const double* diffusivity,
                                                      double gx, double gy, double gz, double tem0, int issave, double* gsnap) {
    AD::MacroBrinkmanCollideNaturalConvection(*L(hf)->p, rho, ux, uy, Z(uz) alpha, nu, *L(hg)->p, tem, qx, qy, Z(qz) diffusivity, gx, gy, Z(gz) tem0, issave != 0, gsnap);
}
void ref_ad_bc_set_t(void* hg, const double* temg, const double* ux, const double* uy, const double* uz, const int* mask) {
    GDEF(L(hg));
    AD::BoundaryConditionSetT(*L(hg)->p, FV(temg), ux, uy, Z(uz) FM(mask));
}
// diffusivity: if kfield != nullptr the per-cell overload is used, else the scalar kconst overload
void ref_ad_bc_set_q(void* hg, const double* qng, const double* ux, const double* uy, const double* uz, const double* kfield, double kconst, const int* mask) {
    GDEF(L(hg));
    if (kfield) AD::BoundaryConditionSetQ(*L(hg)->p, FV(qng), ux, uy, Z(uz) kfield, FM(mask));
    else AD::BoundaryConditionSetQ(*L(hg)->p, FV(qng), ux, uy, Z(uz) kconst, FM(mask));
}

//---------------------------------------------------------------- ANS
void ref_ans_init(void* h, const double* ux, const double* uy, const double* uz, const double* ip, const double* iux, const double* iuy, const double* iuz) {
    ANS::InitialCondition(*L(h)->p, ux, uy, Z(uz) ip, iux, iuy ZL(iuz));
}
void ref_ans_macro_brinkman_collide(void* h, const double* rho, const double* ux, const double* uy, const double* uz,
                                    double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz,
                                    double nu, const double* alpha, int issave) {
    ANS::MacroBrinkmanCollide(*L(h)->p, rho, ux, uy, Z(uz) ip, iux, iuy, Z(iuz) imx, imy, Z(imz) nu, alpha, issave != 0);
}
void ref_ans_ibc_set_u(void* h, const double* uxg, const double* uyg, const double* uzg, const int* mask, double eps) {
    GDEF(L(h));
    ANS::iBoundaryConditionSetU(*L(h)->p, FV(uxg), FV(uyg), Z(FV(uzg)) FM(mask), eps);
}
void ref_ans_ibc_set_rho(void* h, const int* mask) {
    GDEF(L(h));
#if DIM == 2
    ANS::iBoundaryConditionSetRho2D(*L(h)->p, FM(mask));
#else
    ANS::iBoundaryConditionSetRho3D(*L(h)->p, FM(mask));
#endif
}
void ref_ans_sensitivity_brinkman(void* h, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz, const double* dads) {
    ANS::SensitivityBrinkman(*L(h)->p, dfds, ux, uy, Z(uz) imx, imy, Z(imz) dads);
}

// the two-component overload on whatever lattice this library was built for: test/nssens3D.cpp:105 calls it with its D3Q15 lattice
void ref_ans_sensitivity_brinkman_planar(void* h, double* dfds, const double* ux, const double* uy, const double* imx, const double* imy, const double* dads) {
    ANS::SensitivityBrinkman(*L(h)->p, dfds, ux, uy, imx, imy, dads);
}

//---------------------------------------------------------------- AAD
void ref_aad_init(void* hg, const double* ux, const double* uy, const double* uz, const double* item, const double* iqx, const double* iqy, const double* iqz) {
    AAD::InitialCondition(*L(hg)->p, ux, uy, Z(uz) item, iqx, iqy ZL(iqz));
}
void ref_aad_macro_brinkman_collide_heat_exchange(void* hf, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* hg, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* beta, double diffusivity, int issave) {
    AAD::MacroBrinkmanCollideHeatExchange(*L(hf)->p, rho, ux, uy, Z(uz) ip, iux, iuy, Z(iuz) imx, imy, Z(imz) alpha, nu, *L(hg)->p, tem, item, iqx, iqy, Z(iqz) beta, diffusivity, issave != 0);
}
void ref_aad_macro_brinkman_collide_force_convection(void* hf, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* hg, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity, int issave, double* igsnap) {
    AAD::MacroBrinkmanCollideForceConvection(*L(hf)->p, rho, ux, uy, Z(uz) ip, iux, iuy, Z(iuz) imx, imy, Z(imz) alpha, nu, *L(hg)->p, tem, item, iqx, iqy, Z(iqz) diffusivity, issave != 0, igsnap);
}
void ref_aad_macro_brinkman_collide_natural_convection(void* hf, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* hg, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity,
        double gx, double gy, double gz, int issave, double* igsnap) {
    AAD::MacroBrinkmanCollideNaturalConvection(*L(hf)->p, rho, ux, uy, Z(uz) ip, iux, iuy, Z(iuz) imx, imy, Z(imz) alpha, nu, *L(hg)->p, tem, item, iqx, iqy, Z(iqz) diffusivity, gx, gy, Z(gz) issave != 0, igsnap);
}
#if DIM == 2
// D2Q9 only: the D3Q15 AVX overload (adjointadvection_avx.h:1131-1255) cannot be instantiated — it calls
// ExternalForceMassFlow with a missing __uz argument (adjointadvection_avx.h:1161).
void ref_aad_macro_brinkman_collide_natural_convection_massflow(void* hf, const double* rho, const double* ux, const double* uy,
        double* ip, double* iux, double* iuy, double* imx, double* imy, const double* alpha, double nu,
        void* hg, const double* tem, double* item, double* iqx, double* iqy, const double* diffusivity,
        double gx, double gy, const double* dirx, const double* diry, int issave, double* igsnap) {
    AAD::MacroBrinkmanCollideNaturalConvectionMassFlow(*L(hf)->p, rho, ux, uy, ip, iux, iuy, imx, imy, alpha, nu, *L(hg)->p, tem, item, iqx, iqy, diffusivity, gx, gy, dirx, diry, issave != 0, igsnap);
}
// D2Q9 only: the D3Q15 overload (adjointadvection.h:1434-1441) cannot be instantiated — its helpers reference an
// undeclared `_bctype` (adjointadvection.h:583,642,701).
void ref_aad_ibc_set_rho(void* hf, void* hg, const double* rho, const double* ux, const double* uy, const double* tem, const int* mask, double eps) {
    GDEF(L(hf));
    AAD::iBoundaryConditionSetRho(*L(hf)->p, *L(hg)->p, rho, ux, uy, tem, FV(mask), eps);   // 0 / SetT=1 / SetQ=2 (adjointadvection.h:16-17)
}
#endif
void ref_aad_ibc_set_t(void* hg, const double* ux, const double* uy, const double* uz, const int* mask) {
    GDEF(L(hg));
    AAD::iBoundaryConditionSetT(*L(hg)->p, ux, uy, Z(uz) FM(mask));
}
void ref_aad_ibc_set_q(void* hg, const double* ux, const double* uy, const double* uz, const int* mask, double eps) {
    GDEF(L(hg));
    AAD::iBoundaryConditionSetQ(*L(hg)->p, ux, uy, Z(uz) FM(mask), eps);
}
void ref_aad_sensitivity_heat_exchange(void* hg, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
                                       const double* dads, const double* tem, const double* item, const double* dbds) {
    AAD::SensitivityHeatExchange(*L(hg)->p, dfds, ux, uy, Z(uz) imx, imy, Z(imz) dads, tem, item, dbds);
}
void ref_aad_sensitivity_brinkman_diffusivity(void* hg, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* iqx, const double* iqy, const double* iqz, const double* gsnap, const double* igsnap,
        const double* diffusivity, const double* dkds) {
    AAD::SensitivityBrinkmanDiffusivity(*L(hg)->p, dfds, ux, uy, Z(uz) imx, imy, Z(imz) dads, tem, item, iqx, iqy, Z(iqz) gsnap, igsnap, diffusivity, dkds);
}
void ref_aad_sensitivity_temperature_at_heat_source(void* hg, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* iqx, const double* iqy, const double* iqz, const double* gsnap, const double* igsnap,
        const double* diffusivity, const double* dkds, const double* qng, const int* mask) {
    GDEF(L(hg));
    AAD::SensitivityTemperatureAtHeatSource(*L(hg)->p, dfds, ux, uy, Z(uz) imx, imy, Z(imz) dads, tem, item, iqx, iqy, Z(iqz) gsnap, igsnap, diffusivity, dkds, FV(qng), FM(mask));
}

//---------------------------------------------------------------- utilities
double ref_residual3(const double* ux, const double* uy, const double* uz, const double* uxp, const double* uyp, const double* uzp, int n) { return Residual(ux, uy, uz, uxp, uyp, uzp, n); }
double ref_residual2(const double* ux, const double* uy, const double* uxp, const double* uyp, int n) { return Residual(ux, uy, uxp, uyp, n); }
double ref_residual1(const double* ux, const double* uxp, int n) { return Residual(ux, uxp, n); }
void ref_normalize(double* v, int n) { Normalize(v, n); }

//---------------------------------------------------------------- timed loops (CPU baseline, bench.py --impl reference)
// Filters of the reference (src/utility/densityfilter.h, heavisidefilter.h; single-rank path).  mode 0: DensityFilter::GetFilteredValue,
// 1: HeavisideFilter::GetFilteredVariable, 2: HeavisideFilter::GetFilteredSensitivity.  bx > 0 selects the heatsink drivers'
// design-box weight (production/heatsink3D.cpp:87-93, heatsink.cpp:83-89), else the default cone weight of the headers.
void ref_filter(void* h, int mode, double R, double beta, const double* v, const double* dfdrho, double* out, int bx, int by, int bz) {
    PT& p = *L(h)->p;
    std::vector<double> s(v, v + p.nxyz), d, res;
    if (dfdrho) d.assign(dfdrho, dfdrho + p.nxyz);
    auto boxw = [=](int _i1, int _j1, int _k1, int _i2, int _j2, int _k2) {
        if (_i1 < bx && _j1 < by && _k1 < bz && _i2 < bx && _j2 < by && _k2 < bz) {
            return (R - sqrt(pow(_i1 - _i2, 2.0) + pow(_j1 - _j2, 2.0) + pow(_k1 - _k2, 2.0)))/R;
        } else {
            return (_i1 == _i2 && _j1 == _j2 && _k1 == _k2) ? 1.0 : 0.0;
        }
    };
    if (bx > 0) {
        if (mode == 0) res = DensityFilter::GetFilteredValue(p, R, s, boxw);
        else if (mode == 1) res = HeavisideFilter::GetFilteredVariable(p, R, beta, s, boxw);
        else res = HeavisideFilter::GetFilteredSensitivity(p, R, beta, s, d, boxw);
    } else {
        if (mode == 0) res = DensityFilter::GetFilteredValue(p, R, s);
        else if (mode == 1) res = HeavisideFilter::GetFilteredVariable(p, R, beta, s);
        else res = HeavisideFilter::GetFilteredSensitivity(p, R, beta, s, d);
    }
    memcpy(out, res.data(), sizeof(double)*p.nxyz);
}

int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { omp_set_num_threads(n); }

#if DIM == 3
// test/cavityflow3D.cpp:32-59 call sequence on an lx*ly*lz box; returns seconds for `steps` steps after `warmup`.
double ref_time_cavity3d(int lx, int ly, int lz, int steps, int warmup, double* rho_out, double* ux_out, double* uy_out, double* uz_out) {
    double nu = 0.1, u0 = 0.1, theta = 90.0;
    PT pf(lx, ly, lz);
    std::vector<double> rho(pf.nxyz, 1.0), ux(pf.nxyz, 0.0), uy(pf.nxyz, 0.0), uz(pf.nxyz, 0.0);
    NS::InitialCondition(pf, rho.data(), ux.data(), uy.data(), uz.data());
    auto step = [&]() {
        NS::MacroCollide(pf, rho.data(), ux.data(), uy.data(), uz.data(), nu, true);
        pf.Stream();
        pf.BoundaryCondition([=](int _i, int _j, int _k) { return (_i == 0 || _i == lx - 1 || _j == 0 || _j == ly - 1 || _k == 0) ? 1 : 0; });
        NS::BoundaryConditionSetU(pf,
            [=](int, int, int) { return u0*cos(theta*M_PI/180.0); },
            [=](int, int, int) { return u0*sin(theta*M_PI/180.0); },
            [=](int, int, int) { return 0.0; },
            [=](int, int, int _k) { return _k == lz - 1; });
        pf.SmoothCorner();
    };
    for (int t = 0; t < warmup; ++t) step();
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < steps; ++t) step();
    auto t1 = std::chrono::steady_clock::now();
    if (rho_out) {
        std::memcpy(rho_out, rho.data(), sizeof(double)*pf.nxyz); std::memcpy(ux_out, ux.data(), sizeof(double)*pf.nxyz);
        std::memcpy(uy_out, uy.data(), sizeof(double)*pf.nxyz); std::memcpy(uz_out, uz.data(), sizeof(double)*pf.nxyz);
    }
    return std::chrono::duration<double>(t1 - t0).count();
}
#endif

// production/heatsink3D.cpp:148-224 (production/heatsink.cpp:140-215 for DIM 2) forward and adjoint time loops on an
// lx*ly*lz box with caller-supplied alpha / diffusivity fields, convergence `break` disabled; secs[0] = seconds for `steps`
// forward steps, secs[1] = seconds for `steps` adjoint steps, each after `warmup` untimed steps.
void ref_time_heatsink(int lx, int ly, int lz, const double* alpha, const double* diffusivity, double nu, double gx, double gy, double gz,
                       double tem0, double qn0, double L, int steps, int warmup, double* secs) {
#if DIM == 2
    PT pf(lx, ly), pg(lx, ly);
#else
    PT pf(lx, ly, lz), pg(lx, ly, lz);
#endif
    const int n = pf.nxyz;
    std::vector<double> rho(n, 1.0), ux(n, 0.0), uy(n, 0.0), uz(n, 0.0), uxp(n, 0.0), uyp(n, 0.0), uzp(n, 0.0);
    std::vector<double> tem(n, 0.0), qx(n, 0.0), qy(n, 0.0), qz(n, 0.0), qxp(n, 0.0), qyp(n, 0.0), qzp(n, 0.0);
    std::vector<double> irho(n, 0.0), iux(n, 0.0), iuy(n, 0.0), iuz(n, 0.0), imx(n, 0.0), imy(n, 0.0), imz(n, 0.0), iuxp(n, 0.0), iuyp(n, 0.0), iuzp(n, 0.0);
    std::vector<double> item(n, 0.0), iqx(n, 0.0), iqy(n, 0.0), iqz(n, 0.0), iqxp(n, 0.0), iqyp(n, 0.0), iqzp(n, 0.0);
    std::vector<double> gi((size_t)n*PT::nc), igi((size_t)n*PT::nc);
    double *pux = ux.data(), *puy = uy.data(), *puz = uz.data(), *puxp = uxp.data(), *puyp = uyp.data(), *puzp = uzp.data();
    double *pqx = qx.data(), *pqy = qy.data(), *pqz = qz.data(), *pqxp = qxp.data(), *pqyp = qyp.data(), *pqzp = qzp.data();
    double *piux = iux.data(), *piuy = iuy.data(), *piuz = iuz.data(), *piuxp = iuxp.data(), *piuyp = iuyp.data(), *piuzp = iuzp.data();
    double *piqx = iqx.data(), *piqy = iqy.data(), *piqz = iqz.data(), *piqxp = iqxp.data(), *piqyp = iqyp.data(), *piqzp = iqzp.data();
#if DIM == 2
#define WALLF [=](int _i, int _j) { return _i == 0 ? 2 : 1; }
#define WALLG [=](int _i, int _j) { return _i == 0 ? 2 : 0; }
#define SETT [=](int _i, int _j) { return _i == lx - 1 || _j == ly - 1; }
#define SETQ [=](int _i, int _j) { return _j == 0; }
#define SRC [=](int _i, int _j) { return _j == 0 && _i < L; }
#define QN [=](int _i, int _j) { return (_j == 0 && _i < L) ? qn0 : 0.0; }
#define TEM [=](int _i, int _j) { return tem0; }
#else
#define WALLF [=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 1; }
#define WALLG [=](int _i, int _j, int _k) { return (_i == 0 || _k == 0) ? 2 : 0; }
#define SETT [=](int _i, int _j, int _k) { return _i == lx - 1 || _j == ly - 1 || _k == lz - 1; }
#define SETQ [=](int _i, int _j, int _k) { return _j == 0; }
#define SRC [=](int _i, int _j, int _k) { return _j == 0 && _i < L && _k < L; }
#define QN [=](int _i, int _j, int _k) { return (_j == 0 && _i < L && _k < L) ? qn0 : 0.0; }
#define TEM [=](int _i, int _j, int _k) { return tem0; }
#endif
    NS::InitialCondition(pf, rho.data(), pux, puy ZL(puz));
    AD::InitialCondition(pg, tem.data(), pux, puy ZL(puz));
    auto fwd = [&]() {
        AD::MacroBrinkmanCollideNaturalConvection(pf, rho.data(), pux, puy, Z(puz) alpha, nu, pg, tem.data(), pqx, pqy, Z(pqz) diffusivity, gx, gy, Z(gz) tem0, true, gi.data());
        pf.Stream(); pg.Stream();
        pf.BoundaryCondition(WALLF);
        AD::BoundaryConditionSetT(pg, TEM, pux, puy, Z(puz) SETT);
        AD::BoundaryConditionSetQ(pg, QN, pux, puy, Z(puz) diffusivity, SETQ);
        pg.BoundaryCondition(WALLG);
        pf.SmoothCorner(); pg.SmoothCorner();
        std::swap(pux, puxp); std::swap(puy, puyp); std::swap(puz, puzp); std::swap(pqx, pqxp); std::swap(pqy, pqyp); std::swap(pqz, pqzp);
    };
    for (int t = 0; t < warmup; ++t) fwd();
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < steps; ++t) fwd();
    auto t1 = std::chrono::steady_clock::now();
    secs[0] = std::chrono::duration<double>(t1 - t0).count();
    ANS::InitialCondition(pf, pux, puy, Z(puz) irho.data(), piux, piuy ZL(piuz));
    AAD::InitialCondition(pg, pux, puy, Z(puz) item.data(), piqx, piqy ZL(piqz));
    auto adj = [&]() {
        AAD::MacroBrinkmanCollideNaturalConvection(pf, rho.data(), pux, puy, Z(puz) irho.data(), piux, piuy, Z(piuz) imx.data(), imy.data(), Z(imz.data()) alpha, nu,
                                                   pg, tem.data(), item.data(), piqx, piqy, Z(piqz) diffusivity, gx, gy, Z(gz) true, igi.data());
        pf.iStream(); pg.iStream();
        AAD::iBoundaryConditionSetT(pg, pux, puy, Z(puz) SETT);
        AAD::iBoundaryConditionSetQ(pg, pux, puy, Z(puz) SETQ);
        AAD::iBoundaryConditionSetQ(pg, pux, puy, Z(puz) SRC, 1.0);
        pg.iBoundaryCondition(WALLG);
        pf.iBoundaryCondition(WALLF);
        pf.SmoothCorner(); pg.SmoothCorner();
        std::swap(piux, piuxp); std::swap(piuy, piuyp); std::swap(piuz, piuzp); std::swap(piqx, piqxp); std::swap(piqy, piqyp); std::swap(piqz, piqzp);
    };
    for (int t = 0; t < warmup; ++t) adj();
    t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < steps; ++t) adj();
    t1 = std::chrono::steady_clock::now();
    secs[1] = std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
