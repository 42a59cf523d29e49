// D2Q9 lattice of PANSLBM2 (reference src/particle/d2q9.h:24-158), B200 edition: same public surface, populations in HBM as
// fp64 structure-of-arrays inside libpanslbm_b200.so.  See d3q15.h of this tree.
#pragma once
#include <cassert>
#include <utility>
#ifdef _USE_AVX_DEFINES
    #include <immintrin.h>
#endif
#include "../b200/bind.h"

namespace PANSLBM2 {
    template<class T>
    class D2Q9 {
public:
        D2Q9() = delete;
        D2Q9(int _lx, int _ly, int _PEid = 0, int _mx = 1, int _my = 1) :       // d2q9.h:28-35
            lx(_lx), ly(_ly), lz(1), PEid(_PEid), mx(_mx), my(_my), mz(1),
            PEx(_PEid%_mx), PEy(_PEid/_mx), PEz(0),
            nx((_lx + PEx)/_mx), ny((_ly + PEy)/_my), nz(1), nxyz(nx*ny),
            offsetx(_mx - PEx > _lx%_mx ? PEx*nx : _lx - (_mx - PEx)*nx),
            offsety(_my - PEy > _ly%_my ? PEy*ny : _ly - (_my - PEy)*ny),
            offsetz(0), f0(nullptr), f(nullptr)
        {
            b200::only_double<T>();
            assert(0 < _lx && 0 < _ly && 0 <= _PEid && 0 < _mx && 0 < _my);
            core.create(PL_D2Q9, _lx, _ly, 1, _PEid, _mx, _my, 1, &f0, &f);
#ifdef _USE_AVX_DEFINES
            LoadCxCyCzEi();
#endif
        }
        D2Q9(const D2Q9<T>&) = delete;
        ~D2Q9() { core.destroy(); }

        int Index(int _i, int _j) const {
            const int i = _i == -1 ? nx - 1 : (_i == nx ? 0 : _i), j = _j == -1 ? ny - 1 : (_j == ny ? 0 : _j);
            return i + nx*j;
        }
        int Index(int _i, int _j, int) const { return Index(_i, _j); }
        static int IndexF(int _idx, int _c) { return (nc - 1)*_idx + (_c - 1); }
        int IndexPE(int _i, int _j) const {
            const int i = _i == -1 ? mx - 1 : (_i == mx ? 0 : _i), j = _j == -1 ? my - 1 : (_j == my ? 0 : _j);
            return i + mx*j;
        }
        int IndexPE(int _i, int _j, int) const { return IndexPE(_i, _j); }

        void Stream() { b200::check(plh_stream(core.h, 0), "Stream"); }
        void iStream() { b200::check(plh_stream(core.h, 1), "iStream"); }

        template<class Ff> void BoundaryConditionAlongXEdge(int _i, int _directionx, Ff _bctype) { bounce(PL_BC_BOUNCE, 0, _i, _directionx, _bctype); }
        template<class Ff> void BoundaryConditionAlongYEdge(int _j, int _directiony, Ff _bctype) { bounce(PL_BC_BOUNCE, 1, _j, _directiony, _bctype); }
        template<class Ff> void iBoundaryConditionAlongXEdge(int _i, int _directionx, Ff _bctype) { bounce(PL_BC_IBOUNCE, 0, _i, _directionx, _bctype); }
        template<class Ff> void iBoundaryConditionAlongYEdge(int _j, int _directiony, Ff _bctype) { bounce(PL_BC_IBOUNCE, 1, _j, _directiony, _bctype); }
        void SmoothCornerAt(int _i, int _j, int _directionx, int _directiony) {
            b200::check(plh_smooth_corner_at(core.h, _i, _j, 0, _directionx, _directiony, 0), "SmoothCornerAt");
        }

        template<class Ff>
        void BoundaryCondition(Ff _bctype) { b200::faces(*this, PL_BC_BOUNCE, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr); }
        template<class Ff>
        void iBoundaryCondition(Ff _bctype) { b200::faces(*this, PL_BC_IBOUNCE, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr); }
        void SmoothCorner() { b200::check(plh_smooth_corner(core.h), "SmoothCorner"); }     // the 4 corners (d2q9.h:127-132)

        const int lx, ly, lz, PEid, mx, my, mz, PEx, PEy, PEz, nx, ny, nz, nxyz, offsetx, offsety, offsetz;
        static const int nc = 9, nd = 2, cx[nc], cy[nc], cz[nc];
        static const T ei[nc];
        T *f0, *f;

#ifdef _USE_AVX_DEFINES
        static const int packsize = 32/sizeof(T);
        static __m256d __cx[nc], __cy[nc], __cz[nc], __ei[nc];
        static void LoadCxCyCzEi() {
            for (int c = 0; c < nc; ++c) {
                __cx[c] = _mm256_set1_pd((double)cx[c]); __cy[c] = _mm256_set1_pd((double)cy[c]);
                __cz[c] = _mm256_set1_pd((double)cz[c]); __ei[c] = _mm256_set1_pd((double)ei[c]);
            }
        }
        template<class mmT> void LoadF(int _idx, mmT *__f) {       // pack layout of d2q9.h:650-708
            __f[0] = _mm256_set_pd(f0[_idx + 3], f0[_idx + 2], f0[_idx + 1], f0[_idx]);
            for (int c = 1; c < nc; ++c)
                __f[c] = _mm256_set_pd(f[IndexF(_idx + 3, c)], f[IndexF(_idx + 2, c)], f[IndexF(_idx + 1, c)], f[IndexF(_idx, c)]);
        }
        template<class mmT> void StoreF(int _idx, const mmT *__f) {
            for (int c = 0; c < nc; ++c) {
                alignas(32) double lane[4];
                _mm256_store_pd(lane, __f[c]);
                for (int s = 0; s < 4; ++s) { if (c == 0) f0[_idx + s] = lane[s]; else f[IndexF(_idx + s, c)] = lane[s]; }
            }
        }
#endif
        b200::Core& b200_core() { return core; }
        pl_lattice* b200_handle() const { return core.h; }

private:
        b200::Core core;
        template<class Ff> void bounce(int _type, int _axis, int _coord, int _dir, Ff _bctype) {
            b200::plane(*this, _type, _axis, _coord, _dir, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr);
        }
    };

    template<class T>const int D2Q9<T>::cx[D2Q9<T>::nc] = { 0, 1, 0, -1, 0, 1, -1, -1, 1 };
    template<class T>const int D2Q9<T>::cy[D2Q9<T>::nc] = { 0, 0, 1, 0, -1, 1, 1, -1, -1 };
    template<class T>const int D2Q9<T>::cz[D2Q9<T>::nc] = { 0, 0, 0, 0, 0, 0, 0, 0, 0 };
    template<class T>const T D2Q9<T>::ei[D2Q9<T>::nc] = { 4.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/36.0, 1.0/36.0, 1.0/36.0, 1.0/36.0 };
#ifdef _USE_AVX_DEFINES
    template<class T>__m256d D2Q9<T>::__cx[D2Q9<T>::nc];
    template<class T>__m256d D2Q9<T>::__cy[D2Q9<T>::nc];
    template<class T>__m256d D2Q9<T>::__cz[D2Q9<T>::nc];
    template<class T>__m256d D2Q9<T>::__ei[D2Q9<T>::nc];
#endif
}
