// AAD namespace of PANSLBM2 (reference src/equation/adjointadvection.h + src/equation_avx/adjointadvection_avx.h), B200 edition:
// adjoint advection-diffusion equation on the thermal lattice coupled to the adjoint flow lattice, its closures and the
// sensitivities of the heatsink / ncpump objectives.  Same names, argument order and defaults.
#pragma once
#include "adjointnavierstokes.h"

namespace {
    const int SetT = 1;
    const int SetQ = 2;
}

namespace PANSLBM2 {
    namespace AAD {
        namespace detail {
            inline void fwd(pl_collide_args& a, const double* rho, const double* ux, const double* uy, const double* uz, const double* tem) {
                a.rho = const_cast<double*>(rho); a.ux = const_cast<double*>(ux); a.uy = const_cast<double*>(uy); a.uz = const_cast<double*>(uz);
                a.tem = const_cast<double*>(tem);
            }
            inline void adj(pl_collide_args& a, double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz,
                            double* item, double* iqx, double* iqy, double* iqz) {
                a.ip = ip; a.iux = iux; a.iuy = iuy; a.iuz = iuz; a.imx = imx; a.imy = imy; a.imz = imz; a.item = item; a.iqx = iqx; a.iqy = iqy; a.iqz = iqz;
            }
            template<class P, class Q> inline void run(P& p, Q& q, const pl_collide_args& a, const char* what) {
                b200::check(plh_collide(p.b200_handle(), q.b200_handle(), &a), what);
            }
        }

        // ---- adjoint temperature planes (adjointadvection.h:154-300) ----
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetTAlongXEdge(Q<T>& _q, int _i, int _directionx, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AAD_ISET_T, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetTAlongYEdge(Q<T>& _q, int _j, int _directiony, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AAD_ISET_T, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetTAlongXFace(Q<T>& _q, int _i, int _directionx, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AAD_ISET_T, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetTAlongYFace(Q<T>& _q, int _j, int _directiony, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AAD_ISET_T, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetTAlongZFace(Q<T>& _q, int _k, int _directionz, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::plane(_q, PL_BC_AAD_ISET_T, 2, _k, _directionz, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        // ---- adjoint heat-flux planes (adjointadvection.h:304-484); _eps adds the objective's source term ----
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQAlongXEdge(Q<T>& _q, int _i, int _directionx, const T *_ux, const T *_uy, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_q, PL_BC_AAD_ISET_Q, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQAlongYEdge(Q<T>& _q, int _j, int _directiony, const T *_ux, const T *_uy, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::plane(_q, PL_BC_AAD_ISET_Q, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQAlongXFace(Q<T>& _q, int _i, int _directionx, const T *_ux, const T *_uy, const T *_uz, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, _eps);
            b200::plane(_q, PL_BC_AAD_ISET_Q, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQAlongYFace(Q<T>& _q, int _j, int _directiony, const T *_ux, const T *_uy, const T *_uz, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, _eps);
            b200::plane(_q, PL_BC_AAD_ISET_Q, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQAlongZFace(Q<T>& _q, int _k, int _directionz, const T *_ux, const T *_uy, const T *_uz, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, _eps);
            b200::plane(_q, PL_BC_AAD_ISET_Q, 2, _k, _directionz, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        // ---- adjoint pressure planes coupled to the thermal lattice, D2Q9 (adjointadvection.h:488-575); _bctype returns 0 / SetT / SetQ.
        //      (The reference's D3Q15 versions, :578-751, do not compile when instantiated; they are not provided.) ----
        template<class T, template<class>class P, template<class>class Q, class Ff>
        void iBoundaryConditionSetRhoAlongXEdge(P<T>& _p, Q<T>& _q, int _i, int _directionx, const T *_rho, const T *_ux, const T *_uy, const T *_tem, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(_rho, _ux, _uy, nullptr, _tem, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_AAD_ISET_RHO, 0, _i, _directionx, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a, _q.b200_handle());
        }
        template<class T, template<class>class P, template<class>class Q, class Ff>
        void iBoundaryConditionSetRhoAlongYEdge(P<T>& _p, Q<T>& _q, int _j, int _directiony, const T *_rho, const T *_ux, const T *_uy, const T *_tem, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(_rho, _ux, _uy, nullptr, _tem, nullptr, 0.0, _eps);
            b200::plane(_p, PL_BC_AAD_ISET_RHO, 1, _j, _directiony, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a, _q.b200_handle());
        }

        // ---- two-lattice adjoint collides (adjointadvection_avx.h:324-1129) ----
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideHeatExchange(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, T *_ip, T *_iux, T *_iuy, T *_imx, T *_imy, const T *_alpha, T _viscosity,
                                              Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, const T *_beta, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AAD_HEATEX, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, nullptr, _tem); detail::adj(a, _ip, _iux, _iuy, nullptr, _imx, _imy, nullptr, _item, _iqx, _iqy, nullptr);
            a.alpha = _alpha; a.beta = _beta; a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideHeatExchange");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideHeatExchange(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, const T *_uz, T *_ip, T *_iux, T *_iuy, T *_iuz, T *_imx, T *_imy, T *_imz,
                                              const T *_alpha, T _viscosity,
                                              Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, T *_iqz, const T *_beta, T _diffusivity, bool _issave = false) {
            pl_collide_args a = b200::collide_args(PL_AAD_HEATEX, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, _uz, _tem); detail::adj(a, _ip, _iux, _iuy, _iuz, _imx, _imy, _imz, _item, _iqx, _iqy, _iqz);
            a.alpha = _alpha; a.beta = _beta; a.diffusivity_const = _diffusivity;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideHeatExchange");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideForceConvection(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, T *_ip, T *_iux, T *_iuy, T *_imx, T *_imy, const T *_alpha, T _viscosity,
                                                 Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, const T *_diffusivity, bool _issave = false, T *_ig = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AAD_FORCE_CONV, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, nullptr, _tem); detail::adj(a, _ip, _iux, _iuy, nullptr, _imx, _imy, nullptr, _item, _iqx, _iqy, nullptr);
            a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _ig;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideForceConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideForceConvection(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, const T *_uz, T *_ip, T *_iux, T *_iuy, T *_iuz, T *_imx, T *_imy, T *_imz,
                                                 const T *_alpha, T _viscosity,
                                                 Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, T *_iqz, const T *_diffusivity, bool _issave = false, T *_ig = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AAD_FORCE_CONV, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, _uz, _tem); detail::adj(a, _ip, _iux, _iuy, _iuz, _imx, _imy, _imz, _item, _iqx, _iqy, _iqz);
            a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _ig;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideForceConvection");
        }
        // The adjoint step of the heatsink drivers (production/heatsink3D.cpp:194-198; adjointadvection_avx.h:884-1005).
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideNaturalConvection(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, T *_ip, T *_iux, T *_iuy, T *_imx, T *_imy, const T *_alpha, T _viscosity,
                                                   Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, const T *_diffusivity, T _gx, T _gy, bool _issave = false, T *_ig = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AAD_NAT_CONV, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, nullptr, _tem); detail::adj(a, _ip, _iux, _iuy, nullptr, _imx, _imy, nullptr, _item, _iqx, _iqy, nullptr);
            a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _ig; a.gx = _gx; a.gy = _gy;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideNaturalConvection");
        }
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideNaturalConvection(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, const T *_uz, T *_ip, T *_iux, T *_iuy, T *_iuz, T *_imx, T *_imy, T *_imz,
                                                   const T *_alpha, T _viscosity,
                                                   Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, T *_iqz, const T *_diffusivity, T _gx, T _gy, T _gz,
                                                   bool _issave = false, T *_ig = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AAD_NAT_CONV, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, _uz, _tem); detail::adj(a, _ip, _iux, _iuy, _iuz, _imx, _imy, _imz, _item, _iqx, _iqy, _iqz);
            a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _ig; a.gx = _gx; a.gy = _gy; a.gz = _gz;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideNaturalConvection");
        }
        // ncpump objective (production/ncpump.cpp:193-197).  D2Q9 only: the reference's D3Q15 overload (:1295) does not compile.
        template<class T, template<class>class P, template<class>class Q>
        void MacroBrinkmanCollideNaturalConvectionMassFlow(P<T>& _p, const T *_rho, const T *_ux, const T *_uy, T *_ip, T *_iux, T *_iuy, T *_imx, T *_imy, const T *_alpha, T _viscosity,
                                                           Q<T>& _q, const T *_tem, T *_item, T *_iqx, T *_iqy, const T *_diffusivity, T _gx, T _gy,
                                                           const T *_directionx, const T *_directiony, bool _issave = false, T *_ig = nullptr) {
            pl_collide_args a = b200::collide_args(PL_AAD_NAT_CONV_MASSFLOW, _issave, _viscosity);
            detail::fwd(a, _rho, _ux, _uy, nullptr, _tem); detail::adj(a, _ip, _iux, _iuy, nullptr, _imx, _imy, nullptr, _item, _iqx, _iqy, nullptr);
            a.alpha = _alpha; a.diffusivity = _diffusivity; a.snapshot = _ig; a.gx = _gx; a.gy = _gy; a.dirx = _directionx; a.diry = _directiony;
            detail::run(_p, _q, a, "AAD::MacroBrinkmanCollideNaturalConvectionMassFlow");
        }

        // ---- initial condition (adjointadvection.h:1359-1381) ----
        template<class T, template<class>class Q>
        void InitialCondition(Q<T>& _q, const T *_ux, const T *_uy, const T *_item, const T *_iqx, const T *_iqy) {
            const double* a[7] = { _ux, _uy, nullptr, _item, _iqx, _iqy, nullptr };
            b200::check(plh_initial_condition(_q.b200_handle(), 4, a, 7), "AAD::InitialCondition");
        }
        template<class T, template<class>class Q>
        void InitialCondition(Q<T>& _q, const T *_ux, const T *_uy, const T *_uz, const T *_item, const T *_iqx, const T *_iqy, const T *_iqz) {
            const double* a[7] = { _ux, _uy, _uz, _item, _iqx, _iqy, _iqz };
            b200::check(plh_initial_condition(_q.b200_handle(), 4, a, 7), "AAD::InitialCondition");
        }

        // ---- closures on all faces of the global domain (adjointadvection.h:1385-1441) ----
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetT(Q<T>& _q, const T *_ux, const T *_uy, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, 0.0);
            b200::faces(_q, PL_BC_AAD_ISET_T, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetT(Q<T>& _q, const T *_ux, const T *_uy, const T *_uz, Ff _bctype) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, 0.0);
            b200::faces(_q, PL_BC_AAD_ISET_T, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQ(Q<T>& _q, const T *_ux, const T *_uy, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, nullptr, nullptr, nullptr, 0.0, _eps);
            b200::faces(_q, PL_BC_AAD_ISET_Q, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class Q, class Ff>
        void iBoundaryConditionSetQ(Q<T>& _q, const T *_ux, const T *_uy, const T *_uz, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(nullptr, _ux, _uy, _uz, nullptr, nullptr, 0.0, _eps);
            b200::faces(_q, PL_BC_AAD_ISET_Q, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a);
        }
        template<class T, template<class>class P, template<class>class Q, class Ff>
        void iBoundaryConditionSetRho(P<T>& _p, Q<T>& _q, const T *_rho, const T *_ux, const T *_uy, const T *_tem, Ff _bctype, T _eps = T()) {
            pl_bc_aux a = b200::aux(_rho, _ux, _uy, nullptr, _tem, nullptr, 0.0, _eps);
            b200::faces(_p, PL_BC_AAD_ISET_RHO, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), &a, _q.b200_handle());
        }

        // ---- sensitivities (adjointadvection_avx.h:1257-1513) ----
        namespace detail {
            inline pl_sens_args sens(int kind, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
                                     const double* dads) {
                pl_sens_args s;
                std::memset(&s, 0, sizeof(s));
                s.kind = kind; s.dfds = dfds; s.ux = ux; s.uy = uy; s.uz = uz; s.imx = imx; s.imy = imy; s.imz = imz; s.dads = dads;
                return s;
            }
        }
        template<class T, template<class>class Q>
        void SensitivityHeatExchange(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_imx, const T *_imy, const T *_dads, const T *_tem, const T *_item, const T *_dbds) {
            pl_sens_args s = detail::sens(PL_SENS_AAD_HEATEX, _dfds, _ux, _uy, nullptr, _imx, _imy, nullptr, _dads);
            s.tem = _tem; s.item = _item; s.dbds = _dbds;
            b200::check(plh_sensitivity(_q.b200_handle(), &s), "AAD::SensitivityHeatExchange");
        }
        template<class T, template<class>class Q>
        void SensitivityHeatExchange(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_uz, const T *_imx, const T *_imy, const T *_imz, const T *_dads,
                                     const T *_tem, const T *_item, const T *_dbds) {
            pl_sens_args s = detail::sens(PL_SENS_AAD_HEATEX, _dfds, _ux, _uy, _uz, _imx, _imy, _imz, _dads);
            s.tem = _tem; s.item = _item; s.dbds = _dbds;
            b200::check(plh_sensitivity(_q.b200_handle(), &s), "AAD::SensitivityHeatExchange");
        }
        template<class T, template<class>class Q>
        void SensitivityBrinkmanDiffusivity(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_imx, const T *_imy, const T *_dads,
                                            const T *_tem, const T *_item, const T *_iqx, const T *_iqy, const T *_g, const T *_ig, const T *_diffusivity, const T *_dkds) {
            pl_sens_args s = detail::sens(PL_SENS_AAD_BRINKMAN_DIFF, _dfds, _ux, _uy, nullptr, _imx, _imy, nullptr, _dads);
            s.tem = _tem; s.item = _item; s.iqx = _iqx; s.iqy = _iqy; s.gsnap = _g; s.igsnap = _ig; s.diffusivity = _diffusivity; s.dkds = _dkds;
            b200::check(plh_sensitivity(_q.b200_handle(), &s), "AAD::SensitivityBrinkmanDiffusivity");
        }
        template<class T, template<class>class Q>
        void SensitivityBrinkmanDiffusivity(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_uz, const T *_imx, const T *_imy, const T *_imz, const T *_dads,
                                            const T *_tem, const T *_item, const T *_iqx, const T *_iqy, const T *_iqz, const T *_g, const T *_ig,
                                            const T *_diffusivity, const T *_dkds) {
            pl_sens_args s = detail::sens(PL_SENS_AAD_BRINKMAN_DIFF, _dfds, _ux, _uy, _uz, _imx, _imy, _imz, _dads);
            s.tem = _tem; s.item = _item; s.iqx = _iqx; s.iqy = _iqy; s.iqz = _iqz; s.gsnap = _g; s.igsnap = _ig; s.diffusivity = _diffusivity; s.dkds = _dkds;
            b200::check(plh_sensitivity(_q.b200_handle(), &s), "AAD::SensitivityBrinkmanDiffusivity");
        }
        // volume term as above + the heat-source boundary term on every face of the global domain (adjointadvection_avx.h:1403-1513, 16-185)
        template<class T, template<class>class Q, class Fv, class Ff>
        void SensitivityTemperatureAtHeatSource(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_imx, const T *_imy, const T *_dads,
                                                const T *_tem, const T *_item, const T *_iqx, const T *_iqy, const T *_g, const T *_ig,
                                                const T *_diffusivity, const T *_dkds, Fv _qnbc, Ff _bctype) {
            SensitivityBrinkmanDiffusivity(_q, _dfds, _ux, _uy, _imx, _imy, _dads, _tem, _item, _iqx, _iqy, _g, _ig, _diffusivity, _dkds);
            const int ext[2] = { _q.lx, _q.ly };
            for (int axis = 0; axis < 2; ++axis)
                for (int side = 0; side < 2; ++side) {
                    const pl_bc* pln = b200::baked(_q, PL_BC_AD_SET_Q, axis, side ? ext[axis] - 1 : 0, side ? 1 : -1, _bctype, _qnbc, b200::none_t(), b200::none_t());
                    b200::check(plh_sensitivity_heat_source(_q.b200_handle(), pln, _dfds, _ux, _uy, nullptr, _ig, _diffusivity, _dkds), "AAD::SensitivityTemperatureAtHeatSource");
                }
        }
        template<class T, template<class>class Q, class Fv, class Ff>
        void SensitivityTemperatureAtHeatSource(Q<T>& _q, T *_dfds, const T *_ux, const T *_uy, const T *_uz, const T *_imx, const T *_imy, const T *_imz, const T *_dads,
                                                const T *_tem, const T *_item, const T *_iqx, const T *_iqy, const T *_iqz, const T *_g, const T *_ig,
                                                const T *_diffusivity, const T *_dkds, Fv _qnbc, Ff _bctype) {
            SensitivityBrinkmanDiffusivity(_q, _dfds, _ux, _uy, _uz, _imx, _imy, _imz, _dads, _tem, _item, _iqx, _iqy, _iqz, _g, _ig, _diffusivity, _dkds);
            const int ext[3] = { _q.lx, _q.ly, _q.lz };
            for (int axis = 0; axis < 3; ++axis)
                for (int side = 0; side < 2; ++side) {
                    const pl_bc* pln = b200::baked(_q, PL_BC_AD_SET_Q, axis, side ? ext[axis] - 1 : 0, side ? 1 : -1, _bctype, _qnbc, b200::none_t(), b200::none_t());
                    b200::check(plh_sensitivity_heat_source(_q.b200_handle(), pln, _dfds, _ux, _uy, _uz, _ig, _diffusivity, _dkds), "AAD::SensitivityTemperatureAtHeatSource");
                }
        }
    }
}
