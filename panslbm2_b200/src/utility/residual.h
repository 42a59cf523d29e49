// Residual of PANSLBM2 (reference src/utility/residual.h:8-50): sqrt(sum|u-up|^2 / sum|u|^2), here a warp-shuffle block
// reduction on the device.  With a communicator (the reference's _USE_MPI_DEFINES build) the two sums are reduced over all ranks.
#pragma once
#include <cmath>
#include "../b200/bind.h"

namespace PANSLBM2 {
    template<class T>
    T Residual(const T *_ux, const T *_uxp, int _nxy) {
        double r = 0.0;
        b200::check(plh_residual(_ux, nullptr, nullptr, _uxp, nullptr, nullptr, (size_t)_nxy, &r), "Residual");
        return r;
    }
    template<class T>
    T Residual(const T *_ux, const T *_uy, const T *_uxp, const T *_uyp, int _nxy) {
        double r = 0.0;
        b200::check(plh_residual(_ux, _uy, nullptr, _uxp, _uyp, nullptr, (size_t)_nxy, &r), "Residual");
        return r;
    }
    template<class T>
    T Residual(const T *_ux, const T *_uy, const T *_uz, const T *_uxp, const T *_uyp, const T *_uzp, int _nxyz) {
        double r = 0.0;
        b200::check(plh_residual(_ux, _uy, _uz, _uxp, _uyp, _uzp, (size_t)_nxyz, &r), "Residual");
        return r;
    }
}
