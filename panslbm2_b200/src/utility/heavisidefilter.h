// HeavisideFilter of PANSLBM2 (reference src/utility/heavisidefilter.h): cone density filter followed by the tanh projection
// (GetFilteredVariable, :404-579) and its chain rule (GetFilteredSensitivity, :586-894).  B200 edition: weights baked once per
// lattice into a device table, one kernel per call (two for the sensitivity) instead of serial host loops.
#pragma once
#include <cassert>
#include <cmath>
#include <vector>
#include "../b200/bind.h"

namespace PANSLBM2 {
    namespace HeavisideFilter {
        template<class T, template<class>class P, class F>
        std::vector<T> GetFilteredVariable(P<T>& _p, T _R, T _beta, const std::vector<T> &_s, F _weight) {
            assert(_R > T());
            std::vector<T> rho(_p.nxyz, T());
            b200::check(plh_filter_apply(b200::filter(_p, _R, _weight), 1, _beta, _s.data(), nullptr, rho.data(), (size_t)_p.nxyz), "HeavisideFilter::GetFilteredVariable");
            return rho;
        }
        template<class T, template<class>class P>
        std::vector<T> GetFilteredVariable(P<T>& _p, T _R, T _beta, const std::vector<T> &_s) {   // heavisidefilter.h:582-584
            return GetFilteredVariable(_p, _R, _beta, _s, [=](int _i1, int _j1, int _k1, int _i2, int _j2, int _k2) {
                return (_R - sqrt(pow(_i1 - _i2, 2.0) + pow(_j1 - _j2, 2.0) + pow(_k1 - _k2, 2.0)))/_R;
            });
        }
        template<class T, template<class>class P, class F>
        std::vector<T> GetFilteredSensitivity(P<T>& _p, T _R, T _beta, const std::vector<T> &_s, const std::vector<T> &_dfdrho, F _weight) {
            assert(_R > T());
            std::vector<T> dfds(_p.nxyz, T());
            b200::check(plh_filter_apply(b200::filter(_p, _R, _weight), 2, _beta, _s.data(), _dfdrho.data(), dfds.data(), (size_t)_p.nxyz), "HeavisideFilter::GetFilteredSensitivity");
            return dfds;
        }
        template<class T, template<class>class P>
        std::vector<T> GetFilteredSensitivity(P<T>& _p, T _R, T _beta, const std::vector<T> &_s, const std::vector<T> &_dfdrho) {   // heavisidefilter.h:896-899
            return GetFilteredSensitivity(_p, _R, _beta, _s, _dfdrho, [=](int _i1, int _j1, int _k1, int _i2, int _j2, int _k2) {
                return (_R - sqrt(pow(_i1 - _i2, 2.0) + pow(_j1 - _j2, 2.0) + pow(_k1 - _k2, 2.0)))/_R;
            });
        }
    }
}
