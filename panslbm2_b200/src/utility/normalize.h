// Normalize of PANSLBM2 (reference src/utility/normalize.h:8-24): divide by the (global) maximum magnitude, on the device.
#pragma once
#include <cmath>
#include "../b200/bind.h"

namespace PANSLBM2 {
    template<class T>
    void Normalize(T *_v, int _size) {
        b200::check(plh_normalize(_v, (size_t)_size), "Normalize");
    }
}
