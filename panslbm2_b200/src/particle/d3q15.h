// D3Q15 lattice of PANSLBM2 (reference src/particle/d3q15.h:24-249), B200 edition: the same public surface — constructor,
// geometry members, velocity set, Index helpers, Stream/iStream, bounce-back planes, SmoothCorner, public f0/f — in front
// of populations that live in HBM as fp64 structure-of-arrays inside libpanslbm_b200.so.  Every member function ends in a
// CUDA kernel; there is no host implementation behind it.
#pragma once
#include <cassert>
#include <utility>
#ifdef _USE_AVX_DEFINES
    #include <immintrin.h>
#endif
#include "../b200/bind.h"

namespace PANSLBM2 {
    template<class T>
    class D3Q15 {
public:
        D3Q15() = delete;
        // block decomposition of the reference (d3q15.h:28-35): the remainder goes to the high ranks
        D3Q15(int _lx, int _ly, int _lz, int _PEid = 0, int _mx = 1, int _my = 1, int _mz = 1) :
            lx(_lx), ly(_ly), lz(_lz), PEid(_PEid), mx(_mx), my(_my), mz(_mz),
            PEx(_PEid%_mx), PEy((_PEid/_mx)%_my), PEz(_PEid/(_mx*_my)),
            nx((_lx + PEx)/_mx), ny((_ly + PEy)/_my), nz((_lz + PEz)/_mz), nxyz(nx*ny*nz),
            offsetx(_mx - PEx > _lx%_mx ? PEx*nx : _lx - (_mx - PEx)*nx),
            offsety(_my - PEy > _ly%_my ? PEy*ny : _ly - (_my - PEy)*ny),
            offsetz(_mz - PEz > _lz%_mz ? PEz*nz : _lz - (_mz - PEz)*nz),
            f0(nullptr), f(nullptr)
        {
            b200::only_double<T>();
            assert(0 < _lx && 0 < _ly && 0 < _lz && 0 <= _PEid && 0 < _mx && 0 < _my && 0 < _mz);
            core.create(PL_D3Q15, _lx, _ly, _lz, _PEid, _mx, _my, _mz, &f0, &f);
#ifdef _USE_AVX_DEFINES
            LoadCxCyCzEi();
#endif
        }
        D3Q15(const D3Q15<T>&) = delete;
        ~D3Q15() { core.destroy(); }

        int Index(int _i, int _j, int _k) const {      // periodic wrap of the first out-of-range layer (d3q15.h:136-141)
            const int i = _i == -1 ? nx - 1 : (_i == nx ? 0 : _i), j = _j == -1 ? ny - 1 : (_j == ny ? 0 : _j), k = _k == -1 ? nz - 1 : (_k == nz ? 0 : _k);
            return i + nx*j + nx*ny*k;
        }
        static int IndexF(int _idx, int _c) { return (nc - 1)*_idx + (_c - 1); }
        int IndexPE(int _i, int _j, int _k) const {
            const int i = _i == -1 ? mx - 1 : (_i == mx ? 0 : _i), j = _j == -1 ? my - 1 : (_j == my ? 0 : _j), k = _k == -1 ? mz - 1 : (_k == mz ? 0 : _k);
            return i + mx*j + mx*my*k;
        }
        int IndexBCx(int _j, int _k) const { return _j + ny*_k; }
        int IndexBCy(int _k, int _i) const { return _k + nz*_i; }
        int IndexBCz(int _i, int _j) const { return _i + nx*_j; }

        // _offset was the MPI tag base separating the f and g messages (d3q15.h:1309); NCCL pairs messages by order
        void Stream(int _offset = 0) { (void)_offset; b200::check(plh_stream(core.h, 0), "Stream"); }
        void iStream(int _offset = 0) { (void)_offset; b200::check(plh_stream(core.h, 1), "iStream"); }

        template<class Ff> void BoundaryConditionAlongXFace(int _i, int _directionx, Ff _bctype) { bounce(PL_BC_BOUNCE, 0, _i, _directionx, _bctype); }
        template<class Ff> void BoundaryConditionAlongYFace(int _j, int _directiony, Ff _bctype) { bounce(PL_BC_BOUNCE, 1, _j, _directiony, _bctype); }
        template<class Ff> void BoundaryConditionAlongZFace(int _k, int _directionz, Ff _bctype) { bounce(PL_BC_BOUNCE, 2, _k, _directionz, _bctype); }
        template<class Ff> void iBoundaryConditionAlongXFace(int _i, int _directionx, Ff _bctype) { bounce(PL_BC_IBOUNCE, 0, _i, _directionx, _bctype); }
        template<class Ff> void iBoundaryConditionAlongYFace(int _j, int _directiony, Ff _bctype) { bounce(PL_BC_IBOUNCE, 1, _j, _directiony, _bctype); }
        template<class Ff> void iBoundaryConditionAlongZFace(int _k, int _directionz, Ff _bctype) { bounce(PL_BC_IBOUNCE, 2, _k, _directionz, _bctype); }
        void SmoothCornerAlongYZ(int _j, int _k, int _directiony, int _directionz) { corner(0, _j, _k, 0, _directiony, _directionz); }
        void SmoothCornerAlongZX(int _k, int _i, int _directionz, int _directionx) { corner(_i, 0, _k, _directionx, 0, _directionz); }
        void SmoothCornerAlongXY(int _i, int _j, int _directionx, int _directiony) { corner(_i, _j, 0, _directionx, _directiony, 0); }
        void SmoothCornerAt(int _i, int _j, int _k, int _directionx, int _directiony, int _directionz) { corner(_i, _j, _k, _directionx, _directiony, _directionz); }

        template<class Ff>
        void BoundaryCondition(Ff _bctype) { b200::faces(*this, PL_BC_BOUNCE, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr); }
        template<class Ff>
        void iBoundaryCondition(Ff _bctype) { b200::faces(*this, PL_BC_IBOUNCE, _bctype, b200::none_t(), b200::none_t(), b200::none_t(), nullptr); }
        // 12 edges, then 8 corners (d3q15.h:199-220) in one device pass
        void SmoothCorner() { b200::check(plh_smooth_corner(core.h), "SmoothCorner"); }

        const int lx, ly, lz, PEid, mx, my, mz, PEx, PEy, PEz, nx, ny, nz, nxyz, offsetx, offsety, offsetz;
        static const int nc = 15, nd = 3, cx[nc], cy[nc], cz[nc];
        static const T ei[nc];
        T *f0, *f;      // host view of the populations in the reference layout, refreshed from / written back to the device on access

#ifdef _USE_AVX_DEFINES
        static const int packsize = 32/sizeof(T);
        static __m256d __cx[nc], __cy[nc], __cz[nc], __ei[nc];
        static void LoadCxCyCzEi() {
            for (int c = 0; c < nc; ++c) {
                __cx[c] = _mm256_set1_pd((double)cx[c]); __cy[c] = _mm256_set1_pd((double)cy[c]);
                __cz[c] = _mm256_set1_pd((double)cz[c]); __ei[c] = _mm256_set1_pd((double)ei[c]);
            }
        }
        // pack c holds population c of sites _idx.._idx+3 (d3q15.h:1427-1511); host-side view helpers only
        template<class mmT> void LoadF(int _idx, mmT *__f) {
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
        void corner(int _i, int _j, int _k, int _dx, int _dy, int _dz) { b200::check(plh_smooth_corner_at(core.h, _i, _j, _k, _dx, _dy, _dz), "SmoothCorner*"); }
    };

    template<class T>const int D3Q15<T>::cx[D3Q15<T>::nc] = { 0, 1, 0, 0, -1, 0, 0, 1, -1, 1, 1, -1, 1, -1, -1 };
    template<class T>const int D3Q15<T>::cy[D3Q15<T>::nc] = { 0, 0, 1, 0, 0, -1, 0, 1, 1, -1, 1, -1, -1, 1, -1 };
    template<class T>const int D3Q15<T>::cz[D3Q15<T>::nc] = { 0, 0, 0, 1, 0, 0, -1, 1, 1, 1, -1, -1, -1, -1, 1 };
    template<class T>const T D3Q15<T>::ei[D3Q15<T>::nc] = { 2.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0 };
#ifdef _USE_AVX_DEFINES
    template<class T>__m256d D3Q15<T>::__cx[D3Q15<T>::nc];
    template<class T>__m256d D3Q15<T>::__cy[D3Q15<T>::nc];
    template<class T>__m256d D3Q15<T>::__cz[D3Q15<T>::nc];
    template<class T>__m256d D3Q15<T>::__ei[D3Q15<T>::nc];
#endif
}
