// NS namespace of PANSLBM2 (reference src/equation/navierstokes.h + src/equation_avx/navierstokes_avx.h), B200 edition.
// Same function names, argument order and defaults; every call ends in a CUDA kernel of libpanslbm_b200.so that reproduces
// the arithmetic of the reference's AVX overloads (and of their scalar tail) operation by operation.
#pragma once
#include "../b200/bind.h"

namespace PANSLBM2 {
    namespace NS {
        // ---- boundary closures: Zou-He-type velocity / density planes (navierstokes.h:92-426) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetUAlongXEdge(P<T>& _p, int _i, int _directionx, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_U, 0, _i, _directionx, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetUAlongYEdge(P<T>& _p, int _j, int _directiony, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_U, 1, _j, _directiony, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetUAlongXFace(P<T>& _p, int _i, int _directionx, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_U, 0, _i, _directionx, _bctype, _uxbc, _uybc, _uzbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetUAlongYFace(P<T>& _p, int _j, int _directiony, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_U, 1, _j, _directiony, _bctype, _uxbc, _uybc, _uzbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetUAlongZFace(P<T>& _p, int _k, int _directionz, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_U, 2, _k, _directionz, _bctype, _uxbc, _uybc, _uzbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRhoAlongXEdge(P<T>& _p, int _i, int _directionx, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_RHO, 0, _i, _directionx, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRhoAlongYEdge(P<T>& _p, int _j, int _directiony, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_RHO, 1, _j, _directiony, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetRhoAlongXFace(P<T>& _p, int _i, int _directionx, Fv0 _rhobc, Fv1 _usbc, Fv2 _utbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_RHO, 0, _i, _directionx, _bctype, _rhobc, _usbc, _utbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetRhoAlongYFace(P<T>& _p, int _j, int _directiony, Fv0 _rhobc, Fv1 _usbc, Fv2 _utbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_RHO, 1, _j, _directiony, _bctype, _rhobc, _usbc, _utbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetRhoAlongZFace(P<T>& _p, int _k, int _directionz, Fv0 _rhobc, Fv1 _usbc, Fv2 _utbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NS_SET_RHO, 2, _k, _directionz, _bctype, _rhobc, _usbc, _utbc, nullptr);
        }

        // ---- collides (navierstokes_avx.h:93-329) ----
        template<class T, template<class>class P>
        void MacroCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NS_COLLIDE, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NS::MacroCollide");
        }
        template<class T, template<class>class P>
        void MacroCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, T _viscosity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NS_COLLIDE, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy; a.uz = _uz;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NS::MacroCollide");
        }
        template<class T, template<class>class P>
        void MacroBrinkmanCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, const T *_alpha, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NS_BRINKMAN, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy; a.alpha = _alpha;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NS::MacroBrinkmanCollide");
        }
        template<class T, template<class>class P>
        void MacroBrinkmanCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T *_uz, T _viscosity, const T *_alpha, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NS_BRINKMAN, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy; a.uz = _uz; a.alpha = _alpha;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NS::MacroBrinkmanCollide");
        }

        // ---- initial condition: populations = equilibrium (navierstokes.h:550-572) ----
        template<class T, template<class>class P>
        void InitialCondition(P<T>& _p, const T *_rho, const T *_ux, const T *_uy) {
            const double* a[4] = { _rho, _ux, _uy, nullptr };
            b200::check(plh_initial_condition(_p.b200_handle(), 1, a, 4), "NS::InitialCondition");
        }
        template<class T, template<class>class P>
        void InitialCondition(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, const T *_uz) {
            const double* a[4] = { _rho, _ux, _uy, _uz };
            b200::check(plh_initial_condition(_p.b200_handle(), 1, a, 4), "NS::InitialCondition");
        }

        // ---- closures on all faces of the global domain (navierstokes.h:576-612) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetU(P<T>& _p, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::faces(_p, PL_BC_NS_SET_U, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetU(P<T>& _p, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype) {
            b200::faces(_p, PL_BC_NS_SET_U, _bctype, _uxbc, _uybc, _uzbc, nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRho(P<T>& _p, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::faces(_p, PL_BC_NS_SET_RHO, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void BoundaryConditionSetRho(P<T>& _p, Fv0 _rhobc, Fv1 _usbc, Fv2 _utbc, Ff _bctype) {
            b200::faces(_p, PL_BC_NS_SET_RHO, _bctype, _rhobc, _usbc, _utbc, nullptr);
        }
    }
}
