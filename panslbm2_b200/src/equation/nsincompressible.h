// NSin namespace of PANSLBM2 (reference src/equation/nsincompressible.h), B200 edition: the incompressible Navier-Stokes model on
// D2Q9 — momentum moments without the division by rho, rho only in the rest term of the equilibrium.  The reference has scalar
// templates only (no AVX overloads) and 2-D signatures only; same function names, argument order and defaults here, every call
// ends in a CUDA kernel of libpanslbm_b200.so (collide models PL_NSIN_COLLIDE / PL_NSIN_BRINKMAN, closures PL_BC_NSIN_SET_U /
// PL_BC_NSIN_SET_RHO, InitialCondition family 5).
#pragma once
#include "../b200/bind.h"

namespace PANSLBM2 {
    namespace NSin {
        // ---- boundary closures along an edge (nsincompressible.h:46-154) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetUAlongXEdge(P<T>& _p, int _i, int _directionx, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::plane(_p, PL_BC_NSIN_SET_U, 0, _i, _directionx, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetUAlongYEdge(P<T>& _p, int _j, int _directiony, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::plane(_p, PL_BC_NSIN_SET_U, 1, _j, _directiony, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRhoAlongXEdge(P<T>& _p, int _i, int _directionx, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NSIN_SET_RHO, 0, _i, _directionx, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRhoAlongYEdge(P<T>& _p, int _j, int _directiony, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::plane(_p, PL_BC_NSIN_SET_RHO, 1, _j, _directiony, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }

        // ---- collides (nsincompressible.h:158-210) ----
        template<class T, template<class>class P>
        void MacroCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NSIN_COLLIDE, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NSin::MacroCollide");
        }
        template<class T, template<class>class P>
        void MacroBrinkmanCollide(P<T>& _p, T *_rho, T *_ux, T *_uy, T _viscosity, const T *_alpha, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_NSIN_BRINKMAN, _issave, _viscosity);
            a.rho = _rho; a.ux = _ux; a.uy = _uy; a.alpha = _alpha;
            b200::check(plh_collide(_p.b200_handle(), nullptr, &a), "NSin::MacroBrinkmanCollide");
        }

        // ---- initial condition: populations = equilibrium (nsincompressible.h:212-223) ----
        template<class T, template<class>class P>
        void InitialCondition(P<T>& _p, const T *_rho, const T *_ux, const T *_uy) {
            const double* a[4] = { _rho, _ux, _uy, nullptr };
            b200::check(plh_initial_condition(_p.b200_handle(), 5, a, 4), "NSin::InitialCondition");
        }

        // ---- closures on all four edges of the global domain (nsincompressible.h:225-241) ----
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetU(P<T>& _p, Fv0 _uxbc, Fv1 _uybc, Ff _bctype) {
            b200::faces(_p, PL_BC_NSIN_SET_U, _bctype, _uxbc, _uybc, b200::none_t(), nullptr);
        }
        // The reference's NSin::BoundaryConditionSetRho calls a misspelt helper (BoundaryConditionSetRHoAlongYEdge,
        // nsincompressible.h:238) and cannot be instantiated; this one does what the four lines say.
        template<class T, template<class>class P, class Fv0, class Fv1, class Ff>
        void BoundaryConditionSetRho(P<T>& _p, Fv0 _rhobc, Fv1 _usbc, Ff _bctype) {
            b200::faces(_p, PL_BC_NSIN_SET_RHO, _bctype, _rhobc, _usbc, b200::none_t(), nullptr);
        }
    }
}
