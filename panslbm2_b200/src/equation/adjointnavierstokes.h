// ANS namespace of PANSLBM2 (reference src/equation/adjointnavierstokes.h + src/equation_avx/adjointnavierstokes_avx.h),
// B200 edition: adjoint Navier-Stokes equation on the flow lattice.  Same names, argument order and defaults.
#pragma once
#include "../b200/bind.h"

namespace PANSLBM2 {
    namespace ANS {
        // ---- adjoint velocity planes (adjointnavierstokes.h:97-254) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void iBoundaryConditionSetUAlongXEdge(P<T>& _p, int _i, int _directionx, Fv0 _uxbc, Fv1 _uybc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_ANS_ISET_U, 0, _i, _directionx, _bctype, _uxbc, _uybc, b200::none_t(), &a);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void iBoundaryConditionSetUAlongYEdge(P<T>& _p, int _j, int _directiony, Fv0 _uxbc, Fv1 _uybc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_ANS_ISET_U, 1, _j, _directiony, _bctype, _uxbc, _uybc, b200::none_t(), &a);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void iBoundaryConditionSetUAlongXFace(P<T>& _p, int _i, int _directionx, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_ANS_ISET_U, 0, _i, _directionx, _bctype, _uxbc, _uybc, _uzbc, &a);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void iBoundaryConditionSetUAlongYFace(P<T>& _p, int _j, int _directiony, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_ANS_ISET_U, 1, _j, _directiony, _bctype, _uxbc, _uybc, _uzbc, &a);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void iBoundaryConditionSetUAlongZFace(P<T>& _p, int _k, int _directionz, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_ANS_ISET_U, 2, _k, _directionz, _bctype, _uxbc, _uybc, _uzbc, &a);
        }
        // ---- adjoint pressure planes (adjointnavierstokes.h:258-392) ----
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRhoAlongXEdge(P<T>& _p, int _i, int _directionx, Ff _bctype) {
            b200::plane(_p, PL_BC_ANS_ISET_RHO, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRhoAlongYEdge(P<T>& _p, int _j, int _directiony, Ff _bctype) {
            b200::plane(_p, PL_BC_ANS_ISET_RHO, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRhoAlongXFace(P<T>& _p, int _i, int _directionx, Ff _bctype) {
            b200::plane(_p, PL_BC_ANS_ISET_RHO, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRhoAlongYFace(P<T>& _p, int _j, int _directiony, Ff _bctype) {
            b200::plane(_p, PL_BC_ANS_ISET_RHO, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRhoAlongZFace(P<T>& _p, int _k, int _directionz, Ff _bctype) {
            b200::plane(_p, PL_BC_ANS_ISET_RHO, 2, _k, _directionz, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }

        // ---- collide (adjointnavierstokes_avx.h:112-259): rho,u are the frozen forward fields ----
        template<class T, template<class>class P>
        void MacroBrinkmanCollide(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, T *_ip, T *_iux, T *_iuy, T *_imx, T *_imy,
                                  T _viscosity, const T *_alpha, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_ANS_BRINKMAN, _issave, _viscosity);
            a.rho = const_cast<T*>(_rho); a.ux = const_cast<T*>(_ux); a.uy = const_cast<T*>(_uy);
            a.ip = _ip; a.iux = _iux; a.iuy = _iuy; a.imx = _imx; a.imy = _imy; a.alpha = _alpha;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "ANS::MacroBrinkmanCollide");
        }
        template<class T, template<class>class P>
        void MacroBrinkmanCollide(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, const T *_uz, T *_ip, T *_iux, T *_iuy, T *_iuz, T *_imx, T *_imy, T *_imz,
                                  T _viscosity, const T *_alpha, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_ANS_BRINKMAN, _issave, _viscosity);
            a.rho = const_cast<T*>(_rho); a.ux = const_cast<T*>(_ux); a.uy = const_cast<T*>(_uy); a.uz = const_cast<T*>(_uz);
            a.ip = _ip; a.iux = _iux; a.iuy = _iuy; a.iuz = _iuz; a.imx = _imx; a.imy = _imy; a.imz = _imz; a.alpha = _alpha;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "ANS::MacroBrinkmanCollide");
        }

        // ---- initial condition (adjointnavierstokes.h:474-498) ----
        template<class T, template<class>class P>
        void InitialCondition(P<T>& _p, const T *_ux, const T *_uy, const T *_ip, const T *_iux, const T *_iuy) {
            const double* a[7] = { _ux, _uy, nullptr, _ip, _iux, _iuy, nullptr };
            b200::check(plh_initial_condition(_p.b200_handle(), 3, a, 7), "ANS::InitialCondition");
        }
        template<class T, template<class>class P>
        void InitialCondition(P<T>& _p, const T *_ux, const T *_uy, const T *_uz, const T *_ip, const T *_iux, const T *_iuy, const T *_iuz) {
            const double* a[7] = { _ux, _uy, _uz, _ip, _iux, _iuy, _iuz };
            b200::check(plh_initial_condition(_p.b200_handle(), 3, a, 7), "ANS::InitialCondition");
        }

        // ---- closures on all faces of the global domain (adjointnavierstokes.h:502-538) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void iBoundaryConditionSetU(P<T>& _p, Fv0 _uxbc, Fv1 _uybc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::faces(_p, PL_BC_ANS_ISET_U, _bctype, _uxbc, _uybc, b200::none_t(), &a);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Fv2, class Ff>
        void iBoundaryConditionSetU(P<T>& _p, Fv0 _uxbc, Fv1 _uybc, Fv2 _uzbc, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::faces(_p, PL_BC_ANS_ISET_U, _bctype, _uxbc, _uybc, _uzbc, &a);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRho2D(P<T>& _p, Ff _bctype) {
            b200::faces(_p, PL_BC_ANS_ISET_RHO, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Ff>
        void iBoundaryConditionSetRho3D(P<T>& _p, Ff _bctype) {
            b200::faces(_p, PL_BC_ANS_ISET_RHO, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }

        // ---- sensitivity of the Brinkman term (adjointnavierstokes_avx.h:262-293): dfds += 3 dads (u . im) ----
        template<class T, template<class>class P>
        void SensitivityBrinkman(P<T>& _p, T *_dfds, const T *_ux, const T *_uy, const T *_imx, const T *_imy, const T *_dads) {
            pl_sens_args s;
            std::memset(&s, 0, sizeof(s));
            s.kind = PL_SENS_ANS_BRINKMAN; s.dfds = _dfds; s.ux = _ux; s.uy = _uy; s.imx = _imx; s.imy = _imy; s.dads = _dads;
            b200::check(plh_sensitivity(_p.b200_handle(), &s), "ANS::SensitivityBrinkman");
        }
        template<class T, template<class>class P>
        void SensitivityBrinkman(P<T>& _p, T *_dfds, const T *_ux, const T *_uy, const T *_uz, const T *_imx, const T *_imy, const T *_imz, const T *_dads) {
            pl_sens_args s;
            std::memset(&s, 0, sizeof(s));
            s.kind = PL_SENS_ANS_BRINKMAN; s.dfds = _dfds; s.ux = _ux; s.uy = _uy; s.uz = _uz; s.imx = _imx; s.imy = _imy; s.imz = _imz; s.dads = _dads;
            b200::check(plh_sensitivity(_p.b200_handle(), &s), "ANS::SensitivityBrinkman");
        }
    }
}
