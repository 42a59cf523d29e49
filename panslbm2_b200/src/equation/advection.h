// AD namespace of PANSLBM2 (reference src/equation/advection.h + src/equation_avx/advection_avx.h), B200 edition: the thermal
// lattice g next to the flow lattice f.  Both lattices are advanced by ONE kernel per call (and by one fused
// stream+collide pass per time step once the runtime has recognised the loop), each site read and written once.
#pragma once
#include "navierstokes.h"

namespace PANSLBM2 {
    namespace AD {
        namespace detail {
            inline void flow(pl_collide_args& a, double* rho, double* ux, double* uy, double* uz) { a.rho = rho; a.ux = ux; a.uy = uy; a.uz = uz; }
            inline void heat(pl_collide_args& a, double* tem, double* qx, double* qy, double* qz) { a.tem = tem; a.qx = qx; a.qy = qy; a.qz = qz; }
            template<class P, class Q> inline void run(P& p, Q& q, const pl_collide_args& a, const char* what) {
                b200::check(plh_collide(p.b200_handle(), q.b200_handle(), &a), what);
            }
        }

        // ---- Dirichlet temperature planes (advection.h:99-238); they read the velocity the collide of this step saved ----
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetTAlongXEdge(Q<T>& _q, int _i, int _directionx, Fv _tembc, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_T, 0, _i, _directionx, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetTAlongYEdge(Q<T>& _q, int _j, int _directiony, Fv _tembc, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_T, 1, _j, _directiony, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetTAlongXFace(Q<T>& _q, int _i, int _directionx, Fv _tembc, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_T, 0, _i, _directionx, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetTAlongYFace(Q<T>& _q, int _j, int _directiony, Fv _tembc, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_T, 1, _j, _directiony, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetTAlongZFace(Q<T>& _q, int _k, int _directionz, Fv _tembc, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_T, 2, _k, _directionz, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }

        // ---- Neumann heat-flux planes, scalar diffusivity (advection.h:242-381) ----
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongXEdge(Q<T>& _q, int _i, int _directionx, Fv _qnbc, const T *_ux, const T *_uy, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, _diffusivity, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 0, _i, _directionx, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongYEdge(Q<T>& _q, int _j, int _directiony, Fv _qnbc, const T *_ux, const T *_uy, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, _diffusivity, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 1, _j, _directiony, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongXFace(Q<T>& _q, int _i, int _directionx, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, _diffusivity, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 0, _i, _directionx, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongYFace(Q<T>& _q, int _j, int _directiony, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, _diffusivity, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 1, _j, _directiony, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongZFace(Q<T>& _q, int _k, int _directionz, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, _diffusivity, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 2, _k, _directionz, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        // ---- ... per-cell diffusivity (advection.h:385-524) ----
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongXEdge(Q<T>& _q, int _i, int _directionx, Fv _qnbc, const T *_ux, const T *_uy, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, _diffusivity, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 0, _i, _directionx, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongYEdge(Q<T>& _q, int _j, int _directiony, Fv _qnbc, const T *_ux, const T *_uy, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, _diffusivity, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 1, _j, _directiony, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongXFace(Q<T>& _q, int _i, int _directionx, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, _diffusivity, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 0, _i, _directionx, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongYFace(Q<T>& _q, int _j, int _directiony, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, _diffusivity, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 1, _j, _directiony, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQAlongZFace(Q<T>& _q, int _k, int _directionz, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, _diffusivity, 0.0, 0.0);
            b200::plane(_q, PL_BC_AD_SET_Q, 2, _k, _directionz, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }

        // ---- two-lattice collides (advection_avx.h:104-1116) ----
        template<class T, template<class>class P, template<class>class Q>
        void MacroCollideForceConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, Q<T>& _q, T *_tem, T *_qx, T *_qy, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_FORCE_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, nullptr); detail::heat(a, _tem, _qx, _qy, nullptr); a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AD::MacroCollideForceConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroCollideForceConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, T _viscosity, Q<T>& _q, T *_tem, T *_qx, T *_qy, T *_qz, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_FORCE_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, _uz); detail::heat(a, _tem, _qx, _qy, _qz); a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AD::MacroCollideForceConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroCollideNaturalConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, Q<T>& _q, T *_tem, T *_qx, T *_qy, T _diffusivity,
                                           T _gx, T _gy, T _tem0, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_NAT_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, nullptr); detail::heat(a, _tem, _qx, _qy, nullptr); a.diffusivity_const = _diffusivity;
            a.gx = _gx; a.gy = _gy; a.tem0 = _tem0;
            detail::run(_p, _q, a, "AD::MacroCollideNaturalConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroCollideNaturalConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, T _viscosity, Q<T>& _q, T *_tem, T *_qx, T *_qy, T *_qz, T _diffusivity,
                                           T _gx, T _gy, T _gz, T _tem0, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_NAT_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, _uz); detail::heat(a, _tem, _qx, _qy, _qz); a.diffusivity_const = _diffusivity;
            a.gx = _gx; a.gy = _gy; a.gz = _gz; a.tem0 = _tem0;
            detail::run(_p, _q, a, "AD::MacroCollideNaturalConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideHeatExchange(P<T>& _p, T *_rho, T *_ux, T *_uy, const T *_alpha, T _viscosity,
                                              Q<T>& _q, T *_tem, T *_qx, T *_qy, const T *_beta, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_HEATEX, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, nullptr); detail::heat(a, _tem, _qx, _qy, nullptr); a.alpha = _alpha; a.beta = _beta; a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideHeatExchange");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideHeatExchange(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, const T *_alpha, T _viscosity,
                                              Q<T>& _q, T *_tem, T *_qx, T *_qy, T *_qz, const T *_beta, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_HEATEX, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, _uz); detail::heat(a, _tem, _qx, _qy, _qz); a.alpha = _alpha; a.beta = _beta; a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideHeatExchange");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideForceConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, const T *_alpha, T _viscosity,
                                                 Q<T>& _q, T *_tem, T *_qx, T *_qy, const T *_diffusivity, bool _issave = false, T *_g = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_FORCE_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, nullptr); detail::heat(a, _tem, _qx, _qy, nullptr); a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _g;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideForceConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideForceConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, const T *_alpha, T _viscosity,
                                                 Q<T>& _q, T *_tem, T *_qx, T *_qy, T *_qz, const T *_diffusivity, bool _issave = false, T *_g = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_FORCE_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, _uz); detail::heat(a, _tem, _qx, _qy, _qz); a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _g;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideForceConvection");
        }
        // The forward step of the heatsink drivers (production/heatsink3D.cpp:151; advection_avx.h:1001-1116).  `_g` receives the
        // pre-relaxation thermal populations for the sensitivity; it is opaque to callers here as it is in the reference.
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideNaturalConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, const T *_alpha, T _viscosity,
                                                   Q<T>& _q, T *_tem, T *_qx, T *_qy, const T *_diffusivity,
                                                   T _gx, T _gy, T _tem0, bool _issave = false, T *_g = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_NAT_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, nullptr); detail::heat(a, _tem, _qx, _qy, nullptr); a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _g;
            a.gx = _gx; a.gy = _gy; a.tem0 = _tem0;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideNaturalConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideNaturalConvection(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, const T *_alpha, T _viscosity,
                                                   Q<T>& _q, T *_tem, T *_qx, T *_qy, T *_qz, const T *_diffusivity,
                                                   T _gx, T _gy, T _gz, T _tem0, bool _issave = false, T *_g = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AD_BRINKMAN_NAT_CONV, _issave, _viscosity);
            detail::flow(a, _rho, _ux, _uy, _uz); detail::heat(a, _tem, _qx, _qy, _qz); a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _g;
            a.gx = _gx; a.gy = _gy; a.gz = _gz; a.tem0 = _tem0;
            detail::run(_p, _q, a, "AD::MacroBrinkmanCollideNaturalConvection");
        }

        // ---- initial condition (advection.h:1048-1070) ----
        template<class T, template<class>class Q>
        void InitialCondition(Q<T>& _q, const T *_tem, const T *_ux, const T *_uy) {
            const double* a[4] = { _tem, _ux, _uy, nullptr };
            b200::check(plh_initial_condition(_q.b200_handle(), 2, a, 4), "AD::InitialCondition");
        }
        template<class T, template<class>class Q>
        void InitialCondition(Q<T>& _q, const T *_tem, const T *_ux, const T *_uy, const T *_uz) {
            const double* a[4] = { _tem, _ux, _uy, _uz };
            b200::check(plh_initial_condition(_q.b200_handle(), 2, a, 4), "AD::InitialCondition");
        }

        // ---- closures on all faces of the global domain (advection.h:1074-1130) ----
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetT(Q<T>& _q, Fv _tembc, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::faces(_q, PL_BC_AD_SET_T, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetT(Q<T>& _q, Fv _tembc, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::faces(_q, PL_BC_AD_SET_T, _bctype, _tembc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQ(Q<T>& _q, Fv _qnbc, const T *_ux, const T *_uy, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, _diffusivity, 0.0);
            b200::faces(_q, PL_BC_AD_SET_Q, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQ(Q<T>& _q, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, T _diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, _diffusivity, 0.0);
            b200::faces(_q, PL_BC_AD_SET_Q, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQ(Q<T>& _q, Fv _qnbc, const T *_ux, const T *_uy, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, _diffusivity, 0.0, 0.0);
            b200::faces(_q, PL_BC_AD_SET_Q, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void BoundaryConditionSetQ(Q<T>& _q, Fv _qnbc, const T *_ux, const T *_uy, const T *_uz, const T *_diffusivity, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, _diffusivity, 0.0, 0.0);
            b200::faces(_q, PL_BC_AD_SET_Q, _bctype, _qnbc, b200::none_t(), b200::none_t(), &a);
        }
    }
}
