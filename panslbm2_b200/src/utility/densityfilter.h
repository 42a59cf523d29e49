// DensityFilter of PANSLBM2 (reference src/utility/densityfilter.h): cone-weighted average over the (2nR+1)^3 neighbourhood.
// B200 edition: the weight callable is baked once per lattice into a device table, every call is one CUDA kernel
// (the reference loops serially on the host and re-evaluates the callable for every pair on every call).
#pragma once
#include <cassert>
#include <cmath>
#include <vector>
#include "../b200/bind.h"

namespace PANSLBM2 {
    namespace DensityFilter {
        template<class T, template<class>class P, class F>
        std::vector<T> GetFilteredValue(P<T>& _p, T _R, const std::vector<T> &_v, F _weight) {     // densityfilter.h:10-11
            assert(_R > T());
            std::vector<T> fv(_p.nxyz, T());
            b200::check(plh_filter_apply(b200::filter(_p, _R, _weight), 0, 0.0, _v.data(), nullptr, fv.data(), (size_t)_p.nxyz), "DensityFilter::GetFilteredValue");
            return fv;
        }
        template<class T, template<class>class P>
        std::vector<T> GetFilteredValue(P<T>& _p, T _R, const std::vector<T> &_v) {                // densityfilter.h:512-515
            return GetFilteredValue(_p, _R, _v, [=](int _i1, int _j1, int _k1, int _i2, int _j2, int _k2) {
                return (_R - sqrt(pow(_i1 - _i2, 2.0) + pow(_j1 - _j2, 2.0) + pow(_k1 - _k2, 2.0)))/_R;
            });
        }
    }
}
