// TEST INFRASTRUCTURE — NOT PRODUCT CODE, NOT REACHABLE FROM THE PACKAGE.
//
// Compiles the PRODUCT's site-local math (panslbm2_b200/csrc/lbm_{traits,equations,closures,sens}.cuh — the
// __host__ __device__ functions the CUDA kernels call per lattice site) as plain host C++, behind the same op-level
// interface as oracle/lbm_oracle.h (prefix hm_ instead of orc_).  The not-gpu suite drives it next to the reference
// build (oracle/_ref) so that every collide model, closure and sensitivity formula of the CUDA code is checked
// bit-for-bit in the build container, where no GPU exists.  Kernels, streaming, plans and the C-ABI are NOT
// exercised here — those are the -m gpu tests.  Built by tests/hostmath/build.py with -ffp-contract=off.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../panslbm2_b200/csrc/lbm_sens.cuh"

using namespace plb;

namespace {
struct Lat {
    int kind, nc;
    int lx, ly, lz, peid, mx, my, mz, pex, pey, pez;
    int nx, ny, nz, offx, offy, offz;
    long long nxyz, npacked;
    std::vector<double> f0, f;
};
inline Lat* L(void* h) { return static_cast<Lat*>(h); }

template <int D> void load(const Lat* l, long long idx, double (&p)[LT<D>::nc]) {
    p[0] = l->f0[idx];
    for (int c = 1; c < LT<D>::nc; ++c) p[c] = l->f[(size_t)(LT<D>::nc - 1)*idx + (c - 1)];
}
template <int D> void store(Lat* l, long long idx, const double (&p)[LT<D>::nc]) {
    l->f0[idx] = p[0];
    for (int c = 1; c < LT<D>::nc; ++c) l->f[(size_t)(LT<D>::nc - 1)*idx + (c - 1)] = p[c];
}

struct FaceJob {
    int type;
    const int* mask;                       // dense global
    const double *v0, *v1, *v2;            // dense global (may be null)
    const double *rho, *ux, *uy, *uz, *tem, *kfield;   // local per-site fields (may be null)
    double kconst, eps;
};

// the 2*D global boundary planes in the reference's order xmin,xmax,ymin,ymax,zmin,zmax
template <int D> void for_faces(Lat* l, Lat* other, const FaceJob& J, double* dfds = nullptr, const double* igsnap = nullptr, const double* dkds = nullptr) {
    const int ext[3] = {l->lx, l->ly, l->lz}, off[3] = {l->offx, l->offy, l->offz}, n[3] = {l->nx, l->ny, l->nz};
    for (int axis = 0; axis < D; ++axis)
        for (int side = 0; side < 2; ++side) {
            const int dir = side == 0 ? -1 : 1, coord = side == 0 ? 0 : ext[axis] - 1;
            const int loc = coord - off[axis];
            if (loc < 0 || loc >= n[axis]) continue;
            const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
            for (int b = 0; b < n[a2]; ++b)
                for (int a = 0; a < n[a1]; ++a) {
                    int co[3];
                    co[axis] = loc; co[a1] = a; co[a2] = b;
                    const long long idx = co[0] + (long long)l->nx*(co[1] + (long long)l->ny*co[2]);
                    const long long gi = (co[0] + off[0]) + (long long)l->lx*((co[1] + off[1]) + (long long)l->ly*(co[2] + off[2]));
                    const int m = J.mask[gi];
                    if (!m) continue;
                    SiteVals V{};
                    V.v0 = J.v0 ? J.v0[gi] : 0.0; V.v1 = J.v1 ? J.v1[gi] : 0.0; V.v2 = J.v2 ? J.v2[gi] : 0.0;
                    V.rho = J.rho ? J.rho[idx] : 0.0; V.ux = J.ux ? J.ux[idx] : 0.0; V.uy = J.uy ? J.uy[idx] : 0.0;
                    V.uz = (D == 3 && J.uz) ? J.uz[idx] : 0.0; V.tem = J.tem ? J.tem[idx] : 0.0;
                    V.kappa = J.kfield ? J.kfield[idx] : J.kconst; V.eps = J.eps;
                    if (dfds) {   // heat-source sensitivity term: reads the snapshot in the reference layout
                        double ig[LT<D>::nc];
                        for (int c = 0; c < LT<D>::nc; ++c)
                            ig[c] = igsnap[idx < l->npacked ? (size_t)(idx/4)*4*LT<D>::nc + 4*c + idx%4 : (size_t)LT<D>::nc*idx + c];
                        dfds[idx] = dfds[idx] + sens_heat_source_term<D>(ig, axis, dir, V, dkds[idx]);
                        continue;
                    }
                    double p[LT<D>::nc], q[LT<D>::nc];
                    load<D>(l, idx, p);
                    if (other) load<D>(other, idx, q); else for (int c = 0; c < LT<D>::nc; ++c) q[c] = 0.0;
                    apply_closure<D>(J.type, axis, dir, m, p, q, V);
                    store<D>(l, idx, p);
                }
        }
}
void faces(Lat* l, Lat* other, const FaceJob& J) { if (l->kind == 2) for_faces<2>(l, other, J); else for_faces<3>(l, other, J); }

// snapshot device layout in this shim = the reference host layout; collide_site writes P.snap[c*pitch + idx] (SoA), so the
// collides run with a scratch SoA snapshot that is re-laid out afterwards.
template <int D, int M> void collide_all(Lat* f, Lat* g, CollideParams P, double* snap_ref) {
    constexpr unsigned FL = ModelFlags<M>::v;
    constexpr int NC = LT<D>::nc;
    std::vector<double> soa;
    if ((FL & F_SNAP) && snap_ref) { soa.assign((size_t)NC*f->nxyz, 0.0); P.snap = soa.data(); P.snap_pitch = (size_t)f->nxyz; }
    else P.snap = nullptr;
    for (long long idx = 0; idx < f->nxyz; ++idx) {
        double p[NC], q[NC];
        load<D>(f, idx, p);
        if constexpr ((FL & F_G) != 0) load<D>(g, idx, q); else for (int c = 0; c < NC; ++c) q[c] = 0.0;
        if (idx < f->npacked) collide_site<D, FL, false>(p, q, P, (size_t)idx, P.issave != 0);
        else collide_site<D, FL, true>(p, q, P, (size_t)idx, P.issave != 0);
        store<D>(f, idx, p);
        if constexpr ((FL & F_G) != 0) store<D>(g, idx, q);
    }
    if (P.snap && P.issave)
        for (long long idx = 0; idx < f->nxyz; ++idx)
            for (int c = 0; c < NC; ++c)
                snap_ref[idx < f->npacked ? (size_t)(idx/4)*4*NC + 4*c + idx%4 : (size_t)NC*idx + c] = soa[(size_t)c*f->nxyz + idx];
}
template <int M> void collide(Lat* f, Lat* g, const CollideParams& P, double* snap_ref) {
    if (f->kind == 2) collide_all<2, M>(f, g, P, snap_ref);
    else if constexpr (M < 12) collide_all<3, M>(f, g, P, snap_ref);
}
CollideParams params(const Lat* f, double nu, double kconst, double gx, double gy, double gz, double tem0, int issave) {
    CollideParams P;
    memset(&P, 0, sizeof(P));
    P.issave = issave;
    P.scalar_build = (f->npacked == 0 && f->nxyz >= 4) ? 1 : 0;      // hm_set_scalar_build(1): as pl_set_scalar_order does in the product
    P.omegaf = 1.0/(3.0*nu + 0.5); P.iomegaf = 1.0 - P.omegaf;
    P.omegag = 1.0/(3.0*kconst + 0.5); P.iomegag = 1.0 - P.omegag;
    const bool d3 = f->kind == 3;
    P.gx = gx; P.gy = gy; P.gz = d3 ? gz : 0.0; P.tem0 = tem0;
    for (int c = 0; c < f->nc; ++c) {
        double cx = d3 ? LT<3>::cx(c) : LT<2>::cx(c), cy = d3 ? LT<3>::cy(c) : LT<2>::cy(c), cz = d3 ? LT<3>::cz(c) : 0;
        double ei = d3 ? LT<3>::ei(c) : LT<2>::ei(c);
        double s = cx*P.gx + cy*P.gy;
        if (d3) s = s + cz*P.gz;
        P.cg[c] = s; P.eicg[c] = ei*s;
    }
    return P;
}
}  // namespace

// a program built WITHOUT _USE_AVX_DEFINES runs the reference's scalar templates at every site: no packed sites
static int g_scalar_build = 0;

extern "C" {

void hm_set_scalar_build(int on) { g_scalar_build = on; }
void* hm_lattice_create(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz) {
    Lat* l = new Lat();
    if (kind == 2) { lz = 1; mz = 1; }
    l->kind = kind; l->nc = kind == 2 ? 9 : 15;
    l->lx = lx; l->ly = ly; l->lz = lz; l->peid = peid; l->mx = mx; l->my = my; l->mz = mz;
    l->pex = peid%mx; l->pey = kind == 2 ? peid/mx : (peid/mx)%my; l->pez = kind == 2 ? 0 : peid/(mx*my);
    l->nx = (lx + l->pex)/mx; l->ny = (ly + l->pey)/my; l->nz = kind == 2 ? 1 : (lz + l->pez)/mz;
    l->offx = mx - l->pex > lx%mx ? l->pex*l->nx : lx - (mx - l->pex)*l->nx;
    l->offy = my - l->pey > ly%my ? l->pey*l->ny : ly - (my - l->pey)*l->ny;
    l->offz = kind == 2 ? 0 : (mz - l->pez > lz%mz ? l->pez*l->nz : lz - (mz - l->pez)*l->nz);
    l->nxyz = (long long)l->nx*l->ny*l->nz; l->npacked = g_scalar_build ? 0 : 4*(l->nxyz/4);
    l->f0.assign(l->nxyz, 0.0); l->f.assign((size_t)l->nxyz*(l->nc - 1), 0.0);
    return l;
}
void hm_lattice_destroy(void* h) { delete L(h); }
void hm_lattice_info(void* h, int* o) {
    Lat* l = L(h);
    int v[18] = {l->lx, l->ly, l->lz, l->peid, l->mx, l->my, l->mz, l->pex, l->pey, l->pez, l->nx, l->ny, l->nz, (int)l->nxyz, l->offx, l->offy, l->offz, l->nc};
    memcpy(o, v, sizeof(v));
}
void hm_lattice_get(void* h, double* f0, double* f) { memcpy(f0, L(h)->f0.data(), sizeof(double)*L(h)->f0.size()); memcpy(f, L(h)->f.data(), sizeof(double)*L(h)->f.size()); }
void hm_lattice_set(void* h, const double* f0, const double* f) { memcpy(L(h)->f0.data(), f0, sizeof(double)*L(h)->f0.size()); memcpy(L(h)->f.data(), f, sizeof(double)*L(h)->f.size()); }

// ---- closures (global boundary planes in the reference's order)
void hm_bc(void* h, const int* bct, int inverse) { FaceJob J{}; J.type = inverse ? BC_IBOUNCE : BC_BOUNCE; J.mask = bct; faces(L(h), nullptr, J); }
void hm_ns_bc_set_u(void* h, const double* ux, const double* uy, const double* uz, const int* mask) {
    FaceJob J{}; J.type = BC_NS_SET_U; J.mask = mask; J.v0 = ux; J.v1 = uy; J.v2 = uz; faces(L(h), nullptr, J);
}
void hm_ns_bc_set_rho(void* h, const double* v0, const double* v1, const double* v2, const int* mask) {
    FaceJob J{}; J.type = BC_NS_SET_RHO; J.mask = mask; J.v0 = v0; J.v1 = v1; J.v2 = v2; faces(L(h), nullptr, J);
}
void hm_ad_bc_set_t(void* h, const double* temg, const double* ux, const double* uy, const double* uz, const int* mask) {
    FaceJob J{}; J.type = BC_AD_SET_T; J.mask = mask; J.v0 = temg; J.ux = ux; J.uy = uy; J.uz = uz; faces(L(h), nullptr, J);
}
void hm_ad_bc_set_q(void* h, const double* qng, const double* ux, const double* uy, const double* uz, const double* kfield, double kconst, const int* mask) {
    FaceJob J{}; J.type = BC_AD_SET_Q; J.mask = mask; J.v0 = qng; J.ux = ux; J.uy = uy; J.uz = uz; J.kfield = kfield; J.kconst = kconst; faces(L(h), nullptr, J);
}
void hm_ans_ibc_set_u(void* h, const double* ux, const double* uy, const double* uz, const int* mask, double eps) {
    FaceJob J{}; J.type = BC_ANS_ISET_U; J.mask = mask; J.v0 = ux; J.v1 = uy; J.v2 = uz; J.eps = eps; faces(L(h), nullptr, J);
}
void hm_ans_ibc_set_rho(void* h, const int* mask) { FaceJob J{}; J.type = BC_ANS_ISET_RHO; J.mask = mask; faces(L(h), nullptr, J); }
void hm_aad_ibc_set_t(void* h, const double* ux, const double* uy, const double* uz, const int* mask) {
    FaceJob J{}; J.type = BC_AAD_ISET_T; J.mask = mask; J.ux = ux; J.uy = uy; J.uz = uz; faces(L(h), nullptr, J);
}
void hm_aad_ibc_set_q(void* h, const double* ux, const double* uy, const double* uz, const int* mask, double eps) {
    FaceJob J{}; J.type = BC_AAD_ISET_Q; J.mask = mask; J.ux = ux; J.uy = uy; J.uz = uz; J.eps = eps; faces(L(h), nullptr, J);
}
void hm_aad_ibc_set_rho(void* hf, void* hg, const double* rho, const double* ux, const double* uy, const double* tem, const int* mask, double eps) {
    FaceJob J{}; J.type = BC_AAD_ISET_RHO; J.mask = mask; J.rho = rho; J.ux = ux; J.uy = uy; J.tem = tem; J.eps = eps; faces(L(hf), L(hg), J);
}

// ---- collides (argument order of oracle/lbm_oracle.h)
#define SETM(P) P.rho = rho; P.ux = ux; P.uy = uy; P.uz = uz;
void hm_ns_macro_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    CollideParams P = params(L(h), nu, 0, 0, 0, 0, 0, issave); SETM(P) collide<1>(L(h), nullptr, P, nullptr);
}
void hm_ns_macro_brinkman_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    CollideParams P = params(L(h), nu, 0, 0, 0, 0, 0, issave); SETM(P) P.alpha = alpha; collide<2>(L(h), nullptr, P, nullptr);
}
// NSin (D2Q9 only; models 13 / 14, closures 12 / 13, InitialCondition family 5)
void hm_nsin_macro_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    CollideParams P = params(L(h), nu, 0, 0, 0, 0, 0, issave); SETM(P) collide<13>(L(h), nullptr, P, nullptr);
}
void hm_nsin_macro_brinkman_collide(void* h, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    CollideParams P = params(L(h), nu, 0, 0, 0, 0, 0, issave); SETM(P) P.alpha = alpha; collide<14>(L(h), nullptr, P, nullptr);
}
void hm_nsin_bc_set_u(void* h, const double* ux, const double* uy, const double* uz, const int* mask) {
    FaceJob J{}; J.type = BC_NSIN_SET_U; J.mask = mask; J.v0 = ux; J.v1 = uy; faces(L(h), nullptr, J);
}
void hm_nsin_bc_set_rho(void* h, const double* v0, const double* v1, const double* v2, const int* mask) {
    FaceJob J{}; J.type = BC_NSIN_SET_RHO; J.mask = mask; J.v0 = v0; J.v1 = v1; faces(L(h), nullptr, J);
}
#define SETQ(P) P.tem = tem; P.qx = qx; P.qy = qy; P.qz = qz;
void hm_ad_macro_collide_force_convection(void* f, double* rho, double* ux, double* uy, double* uz, double nu,
        void* g, double* tem, double* qx, double* qy, double* qz, double diffusivity, int issave) {
    CollideParams P = params(L(f), nu, diffusivity, 0, 0, 0, 0, issave); SETM(P) SETQ(P) collide<3>(L(f), L(g), P, nullptr);
}
void hm_ad_macro_collide_natural_convection(void* f, double* rho, double* ux, double* uy, double* uz, double nu,
        void* g, double* tem, double* qx, double* qy, double* qz, double diffusivity, double gx, double gy, double gz, double tem0, int issave) {
    CollideParams P = params(L(f), nu, diffusivity, gx, gy, gz, tem0, issave); SETM(P) SETQ(P) collide<4>(L(f), L(g), P, nullptr);
}
void hm_ad_macro_brinkman_collide_heat_exchange(void* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        void* g, double* tem, double* qx, double* qy, double* qz, const double* beta, double diffusivity, int issave) {
    CollideParams P = params(L(f), nu, diffusivity, 0, 0, 0, 0, issave); SETM(P) SETQ(P) P.alpha = alpha; P.beta = beta; collide<5>(L(f), L(g), P, nullptr);
}
void hm_ad_macro_brinkman_collide_force_convection(void* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        void* g, double* tem, double* qx, double* qy, double* qz, const double* diffusivity, int issave, double* gsnap) {
    CollideParams P = params(L(f), nu, 0, 0, 0, 0, 0, issave); SETM(P) SETQ(P) P.alpha = alpha; P.kappa = diffusivity; collide<6>(L(f), L(g), P, gsnap);
}
void hm_ad_macro_brinkman_collide_natural_convection(void* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        void* g, double* tem, double* qx, double* qy, double* qz, const double* diffusivity, double gx, double gy, double gz, double tem0, int issave, double* gsnap) {
    CollideParams P = params(L(f), nu, 0, gx, gy, gz, tem0, issave); SETM(P) SETQ(P) P.alpha = alpha; P.kappa = diffusivity; collide<7>(L(f), L(g), P, gsnap);
}
#define SETA(P) P.rho = (double*)rho; P.ux = (double*)ux; P.uy = (double*)uy; P.uz = (double*)uz; P.ip = ip; P.iux = iux; P.iuy = iuy; P.iuz = iuz; P.imx = imx; P.imy = imy; P.imz = imz;
#define SETAG(P) P.tem = (double*)tem; P.item = item; P.iqx = iqx; P.iqy = iqy; P.iqz = iqz;
void hm_ans_macro_brinkman_collide(void* h, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, double nu, const double* alpha, int issave) {
    CollideParams P = params(L(h), nu, 0, 0, 0, 0, 0, issave); SETA(P) P.alpha = alpha; collide<8>(L(h), nullptr, P, nullptr);
}
void hm_aad_macro_brinkman_collide_heat_exchange(void* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* beta, double diffusivity, int issave) {
    CollideParams P = params(L(f), nu, diffusivity, 0, 0, 0, 0, issave); SETA(P) SETAG(P) P.alpha = alpha; P.beta = beta; collide<9>(L(f), L(g), P, nullptr);
}
void hm_aad_macro_brinkman_collide_force_convection(void* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity, int issave, double* igsnap) {
    CollideParams P = params(L(f), nu, 0, 0, 0, 0, 0, issave); SETA(P) SETAG(P) P.alpha = alpha; P.kappa = diffusivity; collide<10>(L(f), L(g), P, igsnap);
}
void hm_aad_macro_brinkman_collide_natural_convection(void* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        void* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity, double gx, double gy, double gz, int issave, double* igsnap) {
    CollideParams P = params(L(f), nu, 0, gx, gy, gz, 0, issave); SETA(P) SETAG(P) P.alpha = alpha; P.kappa = diffusivity; collide<11>(L(f), L(g), P, igsnap);
}
void hm_aad_macro_brinkman_collide_natural_convection_massflow(void* f, const double* rho, const double* ux, const double* uy,
        double* ip, double* iux, double* iuy, double* imx, double* imy, const double* alpha, double nu,
        void* g, const double* tem, double* item, double* iqx, double* iqy, const double* diffusivity,
        double gx, double gy, const double* dirx, const double* diry, int issave, double* igsnap) {
    const double* uz = nullptr; double *iuz = nullptr, *imz = nullptr, *iqz = nullptr;
    CollideParams P = params(L(f), nu, 0, gx, gy, 0, 0, issave); SETA(P) SETAG(P) P.alpha = alpha; P.kappa = diffusivity; P.dirx = dirx; P.diry = diry;
    collide<12>(L(f), L(g), P, igsnap);
}

// ---- InitialCondition: scalar-order equilibria
void hm_ns_init(void* h, const double* rho, const double* ux, const double* uy, const double* uz) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        if (l->kind == 2) { double e[9]; ns_eq_sc<2>(e, rho[i], ux[i], uy[i], 0.0); store<2>(l, i, e); }
        else { double e[15]; ns_eq_sc<3>(e, rho[i], ux[i], uy[i], uz[i]); store<3>(l, i, e); }
    }
}
void hm_nsin_init(void* h, const double* rho, const double* ux, const double* uy, const double* uz) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) { double e[9]; nsin_eq<2>(e, rho[i], ux[i], uy[i], 0.0); store<2>(l, i, e); }
}
void hm_ad_init(void* h, const double* tem, const double* ux, const double* uy, const double* uz) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        if (l->kind == 2) { double e[9]; ad_eq_sc<2>(e, tem[i], ux[i], uy[i], 0.0); store<2>(l, i, e); }
        else { double e[15]; ad_eq_sc<3>(e, tem[i], ux[i], uy[i], uz[i]); store<3>(l, i, e); }
    }
}
void hm_ans_init(void* h, const double* ux, const double* uy, const double* uz, const double* ip, const double* iux, const double* iuy, const double* iuz) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        if (l->kind == 2) { double e[9]; ans_eq<2>(e, ux[i], uy[i], 0.0, ip[i], iux[i], iuy[i], 0.0); store<2>(l, i, e); }
        else { double e[15]; ans_eq<3>(e, ux[i], uy[i], uz[i], ip[i], iux[i], iuy[i], iuz[i]); store<3>(l, i, e); }
    }
}
void hm_aad_init(void* h, const double* ux, const double* uy, const double* uz, const double* item, const double* iqx, const double* iqy, const double* iqz) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        if (l->kind == 2) { double e[9]; double v = item[i] + 3.0*dot<2>(ux[i], uy[i], 0.0, iqx[i], iqy[i], 0.0); for (double& x : e) x = v; store<2>(l, i, e); }
        else { double e[15]; double v = item[i] + 3.0*dot<3>(ux[i], uy[i], uz[i], iqx[i], iqy[i], iqz[i]); for (double& x : e) x = v; store<3>(l, i, e); }
    }
}

// ---- sensitivities
static SensSite site(long long i, int d3, const double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz, const double* dads) {
    SensSite s{};
    s.dfds = dfds[i]; s.ux = ux[i]; s.uy = uy[i]; s.imx = imx[i]; s.imy = imy[i]; s.dads = dads[i];
    if (d3) { s.uz = uz[i]; s.imz = imz[i]; }
    return s;
}
void hm_ans_sensitivity_brinkman(void* h, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz, const double* dads) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        SensSite s = site(i, l->kind == 3, dfds, ux, uy, uz, imx, imy, imz, dads);
        const bool tail = i >= l->npacked;
        if (l->kind == 2) dfds[i] = tail ? sens_brinkman<2, true>(s) : sens_brinkman<2, false>(s);
        else dfds[i] = tail ? sens_brinkman<3, true>(s) : sens_brinkman<3, false>(s);
    }
}
// test/nssens3D.cpp:105 — the two-component overload handed a D3Q15 lattice: what pl_sensitivity does when uz / imz are absent
void hm_ans_sensitivity_brinkman_planar(void* h, double* dfds, const double* ux, const double* uy, const double* imx, const double* imy, const double* dads) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        SensSite s = site(i, false, dfds, ux, uy, nullptr, imx, imy, nullptr, dads);
        dfds[i] = i >= l->npacked ? sens_brinkman<2, true>(s) : sens_brinkman<2, false>(s);
    }
}
void hm_aad_sensitivity_heat_exchange(void* h, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* dbds) {
    Lat* l = L(h);
    for (long long i = 0; i < l->nxyz; ++i) {
        SensSite s = site(i, l->kind == 3, dfds, ux, uy, uz, imx, imy, imz, dads);
        s.tem = tem[i]; s.item = item[i]; s.dbds = dbds[i];
        dfds[i] = l->kind == 2 ? sens_heatex<2>(s) : sens_heatex<3>(s);
    }
}
}  // extern "C"
template <int D> static void sens_bd(Lat* l, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* iqx, const double* iqy, const double* iqz, const double* gs, const double* igs,
        const double* kappa, const double* dkds) {
    constexpr int NC = LT<D>::nc;
    for (long long i = 0; i < l->nxyz; ++i) {
        SensSite s = site(i, D == 3, dfds, ux, uy, uz, imx, imy, imz, dads);
        s.tem = tem[i]; s.item = item[i]; s.iqx = iqx[i]; s.iqy = iqy[i]; if (D == 3) s.iqz = iqz[i];
        s.kappa = kappa[i]; s.dkds = dkds[i];
        double g[NC], ig[NC];
        for (int c = 0; c < NC; ++c) {
            size_t o = i < l->npacked ? (size_t)(i/4)*4*NC + 4*c + i%4 : (size_t)NC*i + c;
            g[c] = gs[o]; ig[c] = igs[o];
        }
        dfds[i] = i >= l->npacked ? sens_brinkman_diffusivity<D, true>(s, g, ig) : sens_brinkman_diffusivity<D, false>(s, g, ig);
    }
}
extern "C" {
void hm_aad_sensitivity_brinkman_diffusivity(void* h, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* iqx, const double* iqy, const double* iqz, const double* gs, const double* igs,
        const double* kappa, const double* dkds) {
    if (L(h)->kind == 2) sens_bd<2>(L(h), dfds, ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gs, igs, kappa, dkds);
    else sens_bd<3>(L(h), dfds, ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gs, igs, kappa, dkds);
}
void hm_aad_sensitivity_temperature_at_heat_source(void* h, double* dfds, const double* ux, const double* uy, const double* uz, const double* imx, const double* imy, const double* imz,
        const double* dads, const double* tem, const double* item, const double* iqx, const double* iqy, const double* iqz, const double* gs, const double* igs,
        const double* kappa, const double* dkds, const double* qng, const int* mask) {
    hm_aad_sensitivity_brinkman_diffusivity(h, dfds, ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gs, igs, kappa, dkds);
    FaceJob J{}; J.type = 0; J.mask = mask; J.v0 = qng; J.ux = ux; J.uy = uy; J.uz = uz; J.kfield = kappa;
    if (L(h)->kind == 2) for_faces<2>(L(h), nullptr, J, dfds, igs, dkds); else for_faces<3>(L(h), nullptr, J, dfds, igs, dkds);
}

}  // extern "C"
