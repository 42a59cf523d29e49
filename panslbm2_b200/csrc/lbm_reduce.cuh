// Reductions: Residual (src/utility/residual.h:8-50), Normalize (src/utility/normalize.h:8-24), sums.
// Grid-stride accumulation per thread, warp-shuffle + shared-memory block reduction, fixed grid => the
// summation order (and therefore the result) is deterministic from run to run.
#pragma once
#include "lbm_traits.cuh"

namespace plb {

PL_D double warp_sum(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
PL_D double warp_max(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// sum over the block; result valid in thread 0
PL_D double block_sum(double v) {
    __shared__ double sh[32];
    __syncthreads();
    v = warp_sum(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = threadIdx.x < nw ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}
PL_D double block_max(double v) {
    __shared__ double shm[32];
    __syncthreads();
    v = warp_max(v);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) shm[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = threadIdx.x < nw ? shm[threadIdx.x] : 0.0;
    if (w == 0) v = warp_max(v);
    return v;
}

__global__ void k_fill(double* __restrict__ p, double v, long long n) {
    long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_divide(double* __restrict__ p, double d, long long n) {
    long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i]/d;
}

// out[2*b+0] = sum |u-up|^2, out[2*b+1] = sum |u|^2 over this block's share; uy/uz may be null (1- and 2-component overloads)
__global__ void __launch_bounds__(256) k_residual_partial(const double* __restrict__ ux, const double* __restrict__ uy, const double* __restrict__ uz,
                                                          const double* __restrict__ uxp, const double* __restrict__ uyp, const double* __restrict__ uzp,
                                                          long long n, double* __restrict__ out) {
    double d = 0.0, s = 0.0;
    for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) {
        double a = ux[i], e = a - uxp[i];
        double dd = e*e, ss = a*a;
        if (uy) { double b = uy[i], eb = b - uyp[i]; dd = dd + eb*eb; ss = ss + b*b; }
        if (uz) { double c = uz[i], ec = c - uzp[i]; dd = dd + ec*ec; ss = ss + c*c; }
        d += dd; s += ss;
    }
    d = block_sum(d);
    s = block_sum(s);
    if (threadIdx.x == 0) { out[2*blockIdx.x] = d; out[2*blockIdx.x + 1] = s; }
}
__global__ void __launch_bounds__(256) k_sum_partial(const double* __restrict__ v, long long n, double* __restrict__ out) {
    double s = 0.0;
    for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) s += v[i];
    s = block_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}
// final pass: `ncomp` interleaved partials per block -> out[0..ncomp)
__global__ void __launch_bounds__(256) k_sum_final(const double* __restrict__ part, int nb, int ncomp, double* __restrict__ out) {
    for (int c = 0; c < ncomp; ++c) {
        double s = 0.0;
        for (int b = threadIdx.x; b < nb; b += blockDim.x) s += part[(size_t)ncomp*b + c];
        s = block_sum(s);
        if (threadIdx.x == 0) out[c] = s;
    }
}
__global__ void __launch_bounds__(256) k_absmax_partial(const double* __restrict__ v, long long n, double* __restrict__ out) {
    double m = 0.0;
    for (long long i = (long long)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) m = fmax(m, fabs(v[i]));
    m = block_max(m);
    if (threadIdx.x == 0) out[blockIdx.x] = m;
}
__global__ void __launch_bounds__(256) k_absmax_final(const double* __restrict__ part, int nb, double* __restrict__ out) {
    double m = 0.0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) m = fmax(m, part[b]);
    m = block_max(m);
    if (threadIdx.x == 0) out[0] = m;
}

// ---------------------------------------------------------------------------------------------------------
// Cone density filter / Heaviside projection (src/utility/densityfilter.h:389-497, heavisidefilter.h:459-563, 641-857).
// The weight callable of the reference is baked into PATTERNS: wtab[p*K + o] is the weight of the pair (site, neighbour at
// offset o) for every site whose pattern is p = pid[site], o running over the K = (2nR+1)^3 cube in the reference's loop order (i2
// outermost, k2 innermost).  The weights of the reference drivers depend on the offset and on which side of the design box
// the two sites lie (production/heatsink3D.cpp:87-93): a few hundred distinct patterns whatever the lattice size, so the table
// lives in L1/L2 and a call moves ~20 B per site (a dense per-site table was 8K B per site: 1 GB at 81 x 161 x 81).  Pairs
// beyond R or outside the domain carry weight 0 and are skipped, which leaves both running sums bit-identical to the
// reference's.  `v` is a field of the GLOBAL domain.
struct FilterGeom {
    int nx, ny, nz, nR;          // this rank's block
    long long nxyz;
    int gx, gy, gz;              // global domain (neighbours outside it do not count)
    int ox, oy, oz;              // offset of the block in it
    // the field the kernel reads: fx*fy*fz doubles with the block's site (0,0,0) at (ax,ay,az) — the block itself (a = 0), or on a
    // decomposed lattice the block with an nR-wide ghost layer filled from the neighbouring ranks (a = nR along decomposed axes)
    int fx, fy, fz, ax, ay, az;
};
// copy the box [0,ex) x [0,ey) x [0,ez) at origin (sx,sy,sz) of a field of dims (sdx,sdy,.) to origin (dx,dy,dz) of a field of dims (ddx,ddy,.)
__global__ void __launch_bounds__(256) k_box_copy(const double* __restrict__ src, int sdx, int sdy, int sx, int sy, int sz,
                                                  double* __restrict__ dst, int ddx, int ddy, int dx, int dy, int dz, int ex, int ey, int ez) {
    const long long total = (long long)ex*ey*ez;
    for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x*blockDim.x) {
        const int i = (int)(t%ex), j = (int)((t/ex)%ey), k = (int)(t/((long long)ex*ey));
        dst[(size_t)(dx + i) + (size_t)ddx*((size_t)(dy + j) + (size_t)ddy*(size_t)(dz + k))] =
            src[(size_t)(sx + i) + (size_t)sdx*((size_t)(sy + j) + (size_t)sdy*(size_t)(sz + k))];
    }
}
// block -> its place in a zeroed field of the global domain (summed over the ranks afterwards: x + 0.0 == x)
__global__ void __launch_bounds__(256) k_filter_scatter(FilterGeom F, const double* __restrict__ v, double* __restrict__ gv) {
    const long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= F.nxyz) return;
    const int nxy = F.nx*F.ny;
    const int k1 = (int)(idx/nxy), r = (int)(idx - (long long)k1*nxy), j1 = r/F.nx, i1 = r - j1*F.nx;
    gv[(size_t)(i1 + F.ox) + (size_t)F.gx*((size_t)(j1 + F.oy) + (size_t)F.gy*(size_t)(k1 + F.oz))] = v[idx];
}
// mode 0: out = sum(w v)/sum(w)                                   DensityFilter::GetFilteredValue
// mode 1: out = 0.5 (tanh(b/2) + tanh(b (sum(w v)/sum(w) - 1/2)))/tanh(b/2)     HeavisideFilter::GetFilteredVariable
// mode 2: out = aux * 0.5 b (1 - tanh(b (sum(w v)/sum(w) - 1/2))^2)/tanh(b/2)   first pass of GetFilteredSensitivity (aux = dfdrho)
// mode 3: out = sum(w v)/sum(w) with out accumulated from 0 and divided last    second pass of GetFilteredSensitivity
__global__ void __launch_bounds__(256) k_filter(FilterGeom F, const double* __restrict__ wtab, const int* __restrict__ pid, const double* __restrict__ v,
                                                const double* __restrict__ aux, double beta, int mode, double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= F.nxyz) return;
    const int nxy = F.nx*F.ny;
    const int k1 = (int)(idx/nxy), r = (int)(idx - (long long)k1*nxy), j1 = r/F.nx, i1 = r - j1*F.nx;
    const int side = 2*F.nR + 1;
    const double* __restrict__ w = wtab + (size_t)pid[idx]*(size_t)(side*side*side);
    double wv = 0.0, ws = 0.0;
    int o = 0;
    for (int di = -F.nR; di <= F.nR; ++di)
        for (int dj = -F.nR; dj <= F.nR; ++dj)
            for (int dk = -F.nR; dk <= F.nR; ++dk, ++o) {
                const int i2 = i1 + F.ox + di, j2 = j1 + F.oy + dj, k2 = k1 + F.oz + dk;     // global coordinates of the neighbour
                if (i2 < 0 || i2 >= F.gx || j2 < 0 || j2 >= F.gy || k2 < 0 || k2 >= F.gz) continue;
                const double wt = w[o];
                if (wt == 0.0) continue;
                wv = wv + wt*v[(size_t)(i1 + F.ax + di) + (size_t)F.fx*((size_t)(j1 + F.ay + dj) + (size_t)F.fy*(size_t)(k1 + F.az + dk))];
                ws = ws + wt;
            }
    double res;
    if (mode == 0 || mode == 3) res = wv/ws;
    else if (mode == 1) res = 0.5*(tanh(0.5*beta) + tanh(beta*(wv/ws - 0.5)))/tanh(0.5*beta);
    else { const double th = tanh(beta*(wv/ws - 0.5)); res = aux[idx]*(0.5*beta*(1.0 - th*th)/tanh(0.5*beta)); }
    out[idx] = res;
}

// the design map of the heatsink drivers (production/heatsink3D.cpp:114-119, heatsink.cpp:106-111), site by site:
//   diffusivity = ks + (kf - ks) ss (1 + qg)/(ss + qg)          alpha = a0 qf (1 - ss)/(ss + qf)
//   dkds = (kf - ks) qg (1 + qg)/pow(ss + qg, 2)                dads = -a0 qf (1 + qf)/pow(ss + qf, 2)      (a0 = alphamax/(ly - 1))
// in the reference's operation order; pow(x, 2.0) is x*x in glibc as on the device (exact).
__global__ void __launch_bounds__(256) k_design_map(const double* __restrict__ ss, long long n, double kf, double ks, double qg, double a0, double qf,
                                                    double* __restrict__ diffusivity, double* __restrict__ alpha, double* __restrict__ dkds, double* __restrict__ dads) {
    const long long idx = (long long)blockIdx.x*blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const double s = ss[idx];
    diffusivity[idx] = ks + (kf - ks)*s*(1.0 + qg)/(s + qg);
    alpha[idx] = a0*qf*(1.0 - s)/(s + qf);
    const double pg = s + qg, pf = s + qf;
    dkds[idx] = (kf - ks)*qg*(1.0 + qg)/(pg*pg);
    dads[idx] = -a0*qf*(1.0 + qf)/(pf*pf);
}
// sum of v over the box [i0,i1) x [j0,j1) x [k0,k1) of LOCAL coordinates (the objective of the heatsink drivers: the mean
// temperature of the heat patch, heatsink3D.cpp:227-240); fixed grid, deterministic
__global__ void __launch_bounds__(256) k_box_sum_partial(const double* __restrict__ v, int nx, int ny, int i0, int i1, int j0, int j1, int k0, int k1,
                                                         double* __restrict__ out) {
    const long long bi = i1 - i0, bj = j1 - j0, total = bi*bj*(long long)(k1 - k0);
    double s = 0.0;
    for (long long t = (long long)blockIdx.x*blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x*blockDim.x) {
        const long long i = t%bi, r = t/bi, j = r%bj, k = r/bj;
        s += v[(size_t)(i0 + i) + (size_t)nx*((size_t)(j0 + j) + (size_t)ny*(size_t)(k0 + k))];
    }
    s = block_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

}  // namespace plb
