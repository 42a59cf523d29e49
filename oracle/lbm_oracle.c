/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See lbm_oracle.h for the contract and parity status.
 *
 * Everything here works on a per-site local copy  p[0..nc-1]  of the populations (p[0] = f0),
 * gathered from / scattered to the reference layout.  Each primitive exists in the operation
 * order of the reference's AVX overloads ("_avx") and, where it differs, of its scalar templates
 * ("_sc"); the collides pick per site exactly as the reference does (packed sites vs tail sites).
 */
#define _GNU_SOURCE
#include "lbm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* lattice constants: d2q9.h:160-163, d3q15.h:251-254                                          */
static const double CX2[9] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const double CY2[9] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const double CZ2[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
static const double EI2[9] = {4.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/36.0, 1.0/36.0, 1.0/36.0, 1.0/36.0};
static const double CX3[15] = {0, 1, 0, 0, -1, 0, 0, 1, -1, 1, 1, -1, 1, -1, -1};
static const double CY3[15] = {0, 0, 1, 0, 0, -1, 0, 1, 1, -1, 1, -1, -1, 1, -1};
static const double CZ3[15] = {0, 0, 0, 1, 0, 0, -1, 1, 1, 1, -1, -1, -1, -1, 1};
static const double EI3[15] = {2.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0, 1.0/9.0,
                               1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0, 1.0/72.0};
#define NCMAX 15

struct orc_lattice {
    int nd, nc;
    int lx, ly, lz, peid, mx, my, mz, pex, pey, pez, nx, ny, nz, nxyz, offx, offy, offz;
    const double *cx, *cy, *cz, *ei;
    int ci[NCMAX][3];  /* integer velocities */
    int opp[NCMAX];    /* opposite direction */
    double *f0, *f, *fnext;
};
typedef struct orc_lattice L;

static int find_dir(const L* l, int x, int y, int z) {
    for (int c = 0; c < l->nc; ++c) if (l->ci[c][0] == x && l->ci[c][1] == y && l->ci[c][2] == z) return c;
    return -1;
}

/* block decomposition rule: d3q15.h:29-35, d2q9.h:29-35 */
orc_lattice* orc_lattice_create(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz) {
    L* l = (L*)calloc(1, sizeof(L));
    l->nd = kind; l->nc = kind == 2 ? 9 : 15;
    if (kind == 2) { lz = 1; mz = 1; }
    l->lx = lx; l->ly = ly; l->lz = lz; l->peid = peid; l->mx = mx; l->my = my; l->mz = mz;
    l->pex = peid%mx;
    l->pey = kind == 2 ? peid/mx : (peid/mx)%my;
    l->pez = kind == 2 ? 0 : peid/(mx*my);
    l->nx = (lx + l->pex)/mx; l->ny = (ly + l->pey)/my; l->nz = kind == 2 ? 1 : (lz + l->pez)/mz;
    l->nxyz = l->nx*l->ny*l->nz;
    l->offx = mx - l->pex > lx%mx ? l->pex*l->nx : lx - (mx - l->pex)*l->nx;
    l->offy = my - l->pey > ly%my ? l->pey*l->ny : ly - (my - l->pey)*l->ny;
    l->offz = kind == 2 ? 0 : (mz - l->pez > lz%mz ? l->pez*l->nz : lz - (mz - l->pez)*l->nz);
    l->cx = kind == 2 ? CX2 : CX3; l->cy = kind == 2 ? CY2 : CY3; l->cz = kind == 2 ? CZ2 : CZ3; l->ei = kind == 2 ? EI2 : EI3;
    for (int c = 0; c < l->nc; ++c) { l->ci[c][0] = (int)l->cx[c]; l->ci[c][1] = (int)l->cy[c]; l->ci[c][2] = (int)l->cz[c]; }
    for (int c = 0; c < l->nc; ++c) l->opp[c] = find_dir(l, -l->ci[c][0], -l->ci[c][1], -l->ci[c][2]);
    size_t n = (size_t)l->nxyz;
    l->f0 = (double*)calloc(n, sizeof(double));
    l->f = (double*)calloc(n*(l->nc - 1), sizeof(double));
    l->fnext = (double*)calloc(n*(l->nc - 1), sizeof(double));
    return l;
}
void orc_lattice_destroy(orc_lattice* l) { free(l->f0); free(l->f); free(l->fnext); free(l); }
void orc_lattice_info(const orc_lattice* l, int* o) {
    int v[18] = {l->lx, l->ly, l->lz, l->peid, l->mx, l->my, l->mz, l->pex, l->pey, l->pez, l->nx, l->ny, l->nz, l->nxyz, l->offx, l->offy, l->offz, l->nc};
    memcpy(o, v, sizeof(v));
}
void orc_lattice_get(const orc_lattice* l, double* f0, double* f) {
    memcpy(f0, l->f0, sizeof(double)*l->nxyz); memcpy(f, l->f, sizeof(double)*(size_t)l->nxyz*(l->nc - 1));
}
void orc_lattice_set(orc_lattice* l, const double* f0, const double* f) {
    memcpy(l->f0, f0, sizeof(double)*l->nxyz); memcpy(l->f, f, sizeof(double)*(size_t)l->nxyz*(l->nc - 1));
}

/* periodic-wrap index: d3q15.h:136-141 */
static inline int wrap(int v, int n) { return v == -1 ? n - 1 : (v == n ? 0 : v); }
static inline int IDX(const L* l, int i, int j, int k) { return wrap(i, l->nx) + l->nx*(wrap(j, l->ny) + l->ny*wrap(k, l->nz)); }
static inline size_t IF(const L* l, int idx, int c) { return (size_t)(l->nc - 1)*idx + (c - 1); }
static inline int GIDX(const L* l, int i, int j, int k) { return (i + l->offx) + l->lx*((j + l->offy) + l->ly*(k + l->offz)); }

static inline void gather(const L* l, int idx, double* p) {
    p[0] = l->f0[idx];
    for (int c = 1; c < l->nc; ++c) p[c] = l->f[IF(l, idx, c)];
}
static inline void scatter(L* l, int idx, const double* p) {
    l->f0[idx] = p[0];
    for (int c = 1; c < l->nc; ++c) l->f[IF(l, idx, c)] = p[c];
}

/* ------------------------------------------------------------------------------------------ */
/* Stream / iStream (single-rank path): d3q15.h:601-616, 964-979; d2q9.h:284-295                */
static void stream_dir(L* l, int sgn) {
    #pragma omp parallel for
    for (int k = 0; k < l->nz; ++k)
        for (int j = 0; j < l->ny; ++j)
            for (int i = 0; i < l->nx; ++i) {
                int idx = IDX(l, i, j, k);
                for (int c = 1; c < l->nc; ++c) {
                    int src = IDX(l, i - sgn*l->ci[c][0], j - sgn*l->ci[c][1], k - sgn*l->ci[c][2]);
                    l->fnext[IF(l, idx, c)] = l->f[IF(l, src, c)];
                }
            }
    double* t = l->f; l->f = l->fnext; l->fnext = t;
}
void orc_stream(orc_lattice* l) { stream_dir(l, 1); }
void orc_istream(orc_lattice* l) { stream_dir(l, -1); }

/* Iterate the local sites of the global plane  axis = coord  in the reference's loop order and
 * call fn(l, idx, gidx, ctx).  (d3q15.h:984-990 etc.: X face loops j then k, Y face k then i, Z face
 * i then j; the order is irrelevant to the result because every update is site-local.) */
typedef void (*site_fn)(L* l, int idx, int gidx, int axis, int dir, void* ctx);
static void for_plane(L* l, int axis, int coord, int dir, site_fn fn, void* ctx) {
    int off[3] = {l->offx, l->offy, l->offz}, n[3] = {l->nx, l->ny, l->nz};
    int loc = coord - off[axis];
    if (!(0 <= loc && loc < n[axis])) return;
    int a1 = (axis + 1)%3, a2 = (axis + 2)%3;
    if (l->nd == 2) { a1 = 1 - axis; a2 = 2; }
    for (int p = 0; p < n[a1]; ++p)
        for (int q = 0; q < n[a2]; ++q) {
            int ijk[3]; ijk[axis] = loc; ijk[a1] = p; ijk[a2] = q;
            fn(l, IDX(l, ijk[0], ijk[1], ijk[2]), GIDX(l, ijk[0], ijk[1], ijk[2]), axis, dir, ctx);
        }
}
/* the 4 (2D) / 6 (3D) outer planes in the reference's order xmin,xmax,ymin,ymax,zmin,zmax (d3q15.h:182-189) */
static void for_all_faces(L* l, site_fn fn, void* ctx) {
    int ext[3] = {l->lx, l->ly, l->lz};
    for (int axis = 0; axis < l->nd; ++axis) {
        for_plane(l, axis, 0, -1, fn, ctx);
        for_plane(l, axis, ext[axis] - 1, 1, fn, ctx);
    }
}

/* bounce-back (BARRIER=1) / specular (MIRROR=2): d3q15.h:984-1239, d2q9.h:431-575.
 * forward: the populations entering the domain (c_axis == -dir) are rebuilt from their partners;
 * inverse: the ones leaving (c_axis == dir). */
typedef struct { const int* bct; int inverse; } bounce_ctx;
static void bounce_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    bounce_ctx* b = (bounce_ctx*)vctx;
    int t = b->bct[gidx];
    if (t != 1 && t != 2) return;
    int want = b->inverse ? dir : -dir;
    /* quirk: MIRROR on an X face with direction +1 always addresses the plane i = nx-1 (d3q15.h:1013) —
       identical to idx whenever the plane is the xmax face, which is the only way the drivers call it. */
    for (int c = 1; c < l->nc; ++c) {
        if (l->ci[c][axis] != want) continue;
        int src;
        if (t == 1) src = l->opp[c];
        else { int v[3] = {l->ci[c][0], l->ci[c][1], l->ci[c][2]}; v[axis] = -v[axis]; src = find_dir(l, v[0], v[1], v[2]); }
        l->f[IF(l, idx, c)] = l->f[IF(l, idx, src)];
    }
}
void orc_bc(orc_lattice* l, const int* bct, int inverse) { bounce_ctx b = {bct, inverse}; for_all_faces(l, bounce_site, &b); }
void orc_bc_plane(orc_lattice* l, int axis, int coord, int dir, const int* bct, int inverse) {
    bounce_ctx b = {bct, inverse}; for_plane(l, axis, coord, dir, bounce_site, &b);
}

/* SmoothCorner: d3q15.h:199-220, 1242-1303; d2q9.h:127-132, 578-587 */
static void smooth2(L* l, int idx, int a, int b) {
    l->f0[idx] = 0.5*(l->f0[a] + l->f0[b]);
    for (int c = 1; c < l->nc; ++c) l->f[IF(l, idx, c)] = 0.5*(l->f[IF(l, a, c)] + l->f[IF(l, b, c)]);
}
static void smooth3(L* l, int idx, int a, int b, int d) {
    l->f0[idx] = (l->f0[a] + l->f0[b] + l->f0[d])/3.0;
    for (int c = 1; c < l->nc; ++c) l->f[IF(l, idx, c)] = (l->f[IF(l, a, c)] + l->f[IF(l, b, c)] + l->f[IF(l, d, c)])/3.0;
}
static int inloc(int v, int n) { return 0 <= v && v < n; }
void orc_smooth_corner(orc_lattice* l) {
    int X0 = 0 - l->offx, X1 = l->lx - 1 - l->offx, Y0 = 0 - l->offy, Y1 = l->ly - 1 - l->offy, Z0 = 0 - l->offz, Z1 = l->lz - 1 - l->offz;
    if (l->nd == 2) {
        /* order: (xmin,ymin) (xmin,ymax) (xmax,ymin) (xmax,ymax) */
        int cs[4][4] = {{X0, Y0, -1, -1}, {X0, Y1, -1, 1}, {X1, Y0, 1, -1}, {X1, Y1, 1, 1}};
        for (int n = 0; n < 4; ++n) {
            int i = cs[n][0], j = cs[n][1], dx = cs[n][2], dy = cs[n][3];
            if (inloc(i, l->nx) && inloc(j, l->ny)) smooth2(l, IDX(l, i, j, 0), IDX(l, i - dx, j, 0), IDX(l, i, j - dy, 0));
        }
        return;
    }
    /* 12 edges: YZ lines, ZX lines, XY lines, each (min,min) (max,min) (max,max) (min,max) */
    int yz[4][4] = {{Y0, Z0, -1, -1}, {Y1, Z0, 1, -1}, {Y1, Z1, 1, 1}, {Y0, Z1, -1, 1}};
    for (int n = 0; n < 4; ++n) {
        int j = yz[n][0], k = yz[n][1], dy = yz[n][2], dz = yz[n][3];
        if (inloc(j, l->ny) && inloc(k, l->nz))
            for (int i = 0; i < l->nx; ++i) smooth2(l, IDX(l, i, j, k), IDX(l, i, j - dy, k), IDX(l, i, j, k - dz));
    }
    int zx[4][4] = {{Z0, X0, -1, -1}, {Z1, X0, 1, -1}, {Z1, X1, 1, 1}, {Z0, X1, -1, 1}};
    for (int n = 0; n < 4; ++n) {
        int k = zx[n][0], i = zx[n][1], dz = zx[n][2], dx = zx[n][3];
        if (inloc(k, l->nz) && inloc(i, l->nx))
            for (int j = 0; j < l->ny; ++j) smooth2(l, IDX(l, i, j, k), IDX(l, i, j, k - dz), IDX(l, i - dx, j, k));
    }
    int xy[4][4] = {{X0, Y0, -1, -1}, {X1, Y0, 1, -1}, {X1, Y1, 1, 1}, {X0, Y1, -1, 1}};
    for (int n = 0; n < 4; ++n) {
        int i = xy[n][0], j = xy[n][1], dx = xy[n][2], dy = xy[n][3];
        if (inloc(i, l->nx) && inloc(j, l->ny))
            for (int k = 0; k < l->nz; ++k) smooth2(l, IDX(l, i, j, k), IDX(l, i - dx, j, k), IDX(l, i, j - dy, k));
    }
    int cn[8][6] = {{X0, Y0, Z0, -1, -1, -1}, {X1, Y0, Z0, 1, -1, -1}, {X1, Y1, Z0, 1, 1, -1}, {X0, Y1, Z0, -1, 1, -1},
                    {X0, Y0, Z1, -1, -1, 1}, {X1, Y0, Z1, 1, -1, 1}, {X1, Y1, Z1, 1, 1, 1}, {X0, Y1, Z1, -1, 1, 1}};
    for (int n = 0; n < 8; ++n) {
        int i = cn[n][0], j = cn[n][1], k = cn[n][2];
        if (inloc(i, l->nx) && inloc(j, l->ny) && inloc(k, l->nz))
            smooth3(l, IDX(l, i, j, k), IDX(l, i - cn[n][3], j, k), IDX(l, i, j - cn[n][4], k), IDX(l, i, j, k - cn[n][5]));
    }
}

/* ------------------------------------------------------------------------------------------ */
/* shared helpers                                                                              */
/* (ax*bx + ay*by) [+ az*bz] in the reference's association */
static inline double dot3(const L* l, double ax, double ay, double az, double bx, double by, double bz) {
    double s = ax*bx + ay*by;
    if (l->nd == 3) s = s + az*bz;
    return s;
}
static inline double cdot(const L* l, int c, double vx, double vy, double vz) { return dot3(l, l->cx[c], l->cy[c], l->cz[c], vx, vy, vz); }
static inline int npacked(const L* l) { return 4*(l->nxyz/4); }  /* sites handled by the AVX overloads */
static inline void relax(const L* l, double* p, const double* eq, double omega, double iomega) {
    for (int c = 0; c < l->nc; ++c) p[c] = iomega*p[c] + omega*eq[c];
}

/* ------------------------------------------------------------------------------------------ */
/* NS primitives                                                                               */
/* Macro: navierstokes.h:17-50 == navierstokes_avx.h:24-55 (same order) */
static void ns_macro(const L* l, const double* p, double* rho, double* ux, double* uy, double* uz) {
    double r = p[0], x = 0.0, y = 0.0, z = 0.0;
    for (int c = 1; c < l->nc; ++c) {
        r = r + p[c];
        x = x + p[c]*l->cx[c];
        y = y + p[c]*l->cy[c];
        if (l->nd == 3) z = z + p[c]*l->cz[c];
    }
    double inv = 1.0/r;
    *rho = r; *ux = x*inv; *uy = y*inv; *uz = z*inv;
}
/* Equilibrium, AVX order: navierstokes_avx.h:58-75 */
static void ns_eq_avx(const L* l, double* feq, double rho, double ux, double uy, double uz) {
    double a = 1.0 - 1.5*dot3(l, ux, uy, uz, ux, uy, uz);
    for (int c = 0; c < l->nc; ++c) {
        double cu = cdot(l, c, ux, uy, uz);
        feq[c] = l->ei[c]*(rho*(a + (3.0*cu + 4.5*(cu*cu))));
    }
}
/* Equilibrium, scalar order: navierstokes.h:53-70 */
static void ns_eq_sc(const L* l, double* feq, double rho, double ux, double uy, double uz) {
    double uu = 1.0 - 1.5*dot3(l, ux, uy, uz, ux, uy, uz);
    for (int c = 0; c < l->nc; ++c) {
        double ciu = cdot(l, c, ux, uy, uz);
        feq[c] = l->ei[c]*rho*(3.0*ciu + 4.5*ciu*ciu + uu);
    }
}
/* Brinkman force: navierstokes.h:73-88 == navierstokes_avx.h:77-91 */
static void ns_brinkman(const L* l, double* p, double rho, double ux, double uy, double uz, double alpha) {
    double coef = 3.0*alpha*rho/(rho + alpha);
    for (int c = 1; c < l->nc; ++c) p[c] = p[c] - coef*l->ei[c]*cdot(l, c, ux, uy, uz);
}

void orc_ns_init(orc_lattice* l, const double* rho, const double* ux, const double* uy, const double* uz) {
    double feq[NCMAX];
    for (int idx = 0; idx < l->nxyz; ++idx) {   /* navierstokes.h:550-572 (scalar Equilibrium) */
        ns_eq_sc(l, feq, rho[idx], ux[idx], uy[idx], l->nd == 3 ? uz[idx] : 0.0);
        scatter(l, idx, feq);
    }
}

/* MacroCollide / MacroBrinkmanCollide: navierstokes_avx.h:93-329 */
static void ns_collide(L* l, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    double omega = 1.0/(3.0*nu + 0.5), iomega = 1.0 - omega;
    int ne = npacked(l);
    #pragma omp parallel for
    for (int idx = 0; idx < l->nxyz; ++idx) {
        double p[NCMAX], feq[NCMAX], r, x, y, z;
        int tail = idx >= ne;
        gather(l, idx, p);
        ns_macro(l, p, &r, &x, &y, &z);
        /* quirk: the 2-D Brinkman tail saves the macros BEFORE the force (navierstokes_avx.h:246-254) */
        int save_early = alpha && tail && l->nd == 2;
        if (issave && save_early) { rho[idx] = r; ux[idx] = x; uy[idx] = y; }
        if (alpha) {
            ns_brinkman(l, p, r, x, y, z, alpha[idx]);
            ns_macro(l, p, &r, &x, &y, &z);
        }
        if (issave && !save_early) { rho[idx] = r; ux[idx] = x; uy[idx] = y; if (l->nd == 3) uz[idx] = z; }
        if (tail) ns_eq_sc(l, feq, r, x, y, z); else ns_eq_avx(l, feq, r, x, y, z);
        relax(l, p, feq, omega, iomega);
        scatter(l, idx, p);
    }
}
void orc_ns_macro_collide(orc_lattice* l, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    ns_collide(l, rho, ux, uy, uz, nu, NULL, issave);
}
void orc_ns_macro_brinkman_collide(orc_lattice* l, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    ns_collide(l, rho, ux, uy, uz, nu, alpha, issave);
}

/* ---- velocity / pressure closures on a face: navierstokes.h:90-426 ------------------------------
 * Face with normal axis a and outward direction dir.  "in" = populations entering the domain
 * (c_a == -dir, the unknowns), "out" = their opposites, "tan" = populations with c_a == 0 (c != 0).
 *   SetU:   rho0 = (f0 + sum(tan) + 2*sum(out)) / (1 + dir*u_a)
 *   SetRho: u_a  = -dir*(1 - (f0 + sum(tan) + 2*sum(out))/rho)
 *   m_a = rho0*u_a/(6|12);  m_t = (0.5|0.25)*(f_{+t} - f_{-t} - rho0*u_t)
 *   f_in(axis pop) = f_out -dir*(4|8)*m_a ;  f_in(diag) = f_opp + sum_d s_d*m_d, s_a = c_a, s_t = -c_t,
 * all sums in ascending c and x,y,z order exactly as written in the reference. */
typedef struct { const double *v0, *v1, *v2; const int* mask; int setrho; } nsbc_ctx;
static void ns_bc_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    nsbc_ctx* b = (nsbc_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double p[NCMAX];
    gather(l, idx, p);
    double s = p[0];
    for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] == 0) s = s + p[c];
    double o = 0.0; int first = 1;
    for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] == dir) { o = first ? p[c] : o + p[c]; first = 0; }
    double tot = s + 2.0*o;
    double u[3], rho0;
    int t1 = l->nd == 2 ? 1 - axis : (axis + 1)%3, t2 = l->nd == 2 ? -1 : (axis + 2)%3;
    if (!b->setrho) {
        u[0] = b->v0[gidx]; u[1] = b->v1[gidx]; u[2] = l->nd == 3 ? b->v2[gidx] : 0.0;
        rho0 = dir == -1 ? tot/(1.0 - u[axis]) : tot/(1.0 + u[axis]);
    } else {
        /* reference argument order: (rho, us, ut) with (us,ut) = (uy,uz) on X, (uz,ux) on Y, (ux,uy) on Z faces
           (navierstokes.h:326, 363, 400); in 2-D the single tangential velocity (navierstokes.h:267, 296) */
        rho0 = b->v0[gidx];
        u[0] = u[1] = u[2] = 0.0;
        if (l->nd == 2) u[t1] = b->v1[gidx];
        else { u[t1] = b->v1[gidx]; u[t2] = b->v2[gidx]; }
        u[axis] = dir == -1 ? 1.0 - tot/rho0 : -1.0 + tot/rho0;
    }
    double m[3] = {0.0, 0.0, 0.0};
    double kn = l->nd == 2 ? 6.0 : 12.0, kt = l->nd == 2 ? 0.5 : 0.25, ka = l->nd == 2 ? 4.0 : 8.0;
    m[axis] = rho0*u[axis]/kn;
    for (int d = 0; d < l->nd; ++d) if (d != axis) {
        int v[3] = {0, 0, 0}; v[d] = 1; int cp = find_dir(l, v[0], v[1], v[2]); v[d] = -1; int cm = find_dir(l, v[0], v[1], v[2]);
        m[d] = kt*(p[cp] - p[cm] - rho0*u[d]);
    }
    for (int c = 1; c < l->nc; ++c) {
        if (l->ci[c][axis] != -dir) continue;
        int nz = abs(l->ci[c][0]) + abs(l->ci[c][1]) + abs(l->ci[c][2]);
        double val;
        if (nz == 1) val = dir == -1 ? p[l->opp[c]] + ka*m[axis] : p[l->opp[c]] - ka*m[axis];
        else {
            val = p[l->opp[c]];
            for (int d = 0; d < l->nd; ++d) {
                int sgn = d == axis ? l->ci[c][d] : -l->ci[c][d];
                val = sgn > 0 ? val + m[d] : val - m[d];
            }
        }
        l->f[IF(l, idx, c)] = val;
    }
}
void orc_ns_bc_set_u(orc_lattice* l, const double* uxg, const double* uyg, const double* uzg, const int* mask) {
    nsbc_ctx b = {uxg, uyg, uzg, mask, 0}; for_all_faces(l, ns_bc_site, &b);
}
void orc_ns_bc_set_rho(orc_lattice* l, const double* v0, const double* v1, const double* v2, const int* mask) {
    nsbc_ctx b = {v0, v1, v2, mask, 1}; for_all_faces(l, ns_bc_site, &b);
}

/* ------------------------------------------------------------------------------------------ */
/* NSin — incompressible NS on D2Q9: nsincompressible.h (scalar templates only, no AVX overloads: one order for every site) */
/* Macro: nsincompressible.h:11-24 — rho = sum f, u = sum c f (NOT divided by rho) */
static void nsin_macro(const L* l, const double* p, double* rho, double* ux, double* uy) {
    double r = p[0], x = 0.0, y = 0.0;
    for (int c = 1; c < l->nc; ++c) {
        r = r + p[c];
        x = x + l->cx[c]*p[c];
        y = y + l->cy[c]*p[c];
    }
    *rho = r; *ux = x; *uy = y;
}
/* Equilibrium: nsincompressible.h:26-34 — feq_c = ei_c*(3 c.u + 4.5 (c.u)^2 + (rho - 1.5 u.u)) */
static void nsin_eq(const L* l, double* feq, double rho, double ux, double uy) {
    double rhouu = rho - 1.5*(ux*ux + uy*uy);
    for (int c = 0; c < l->nc; ++c) {
        double ciu = l->cx[c]*ux + l->cy[c]*uy;
        feq[c] = l->ei[c]*(3.0*ciu + 4.5*ciu*ciu + rhouu);
    }
}
/* InitialCondition: nsincompressible.h:212-223 */
void orc_nsin_init(orc_lattice* l, const double* rho, const double* ux, const double* uy, const double* uz) {
    (void)uz;
    double feq[NCMAX];
    for (int idx = 0; idx < l->nxyz; ++idx) {
        nsin_eq(l, feq, rho[idx], ux[idx], uy[idx]);
        scatter(l, idx, feq);
    }
}
/* MacroCollide: nsincompressible.h:158-181; MacroBrinkmanCollide: :183-210 (force = NS's ExternalForceBrinkman, :36-44; macros
 * re-evaluated after the force and stored after it) */
static void nsin_collide(L* l, double* rho, double* ux, double* uy, double nu, const double* alpha, int issave) {
    double omega = 1.0/(3.0*nu + 0.5), iomega = 1.0 - omega;
    #pragma omp parallel for
    for (int idx = 0; idx < l->nxyz; ++idx) {
        double p[NCMAX], feq[NCMAX], r, x, y;
        gather(l, idx, p);
        nsin_macro(l, p, &r, &x, &y);
        if (alpha) {
            double coef = 3.0*alpha[idx]*r/(r + alpha[idx]);
            for (int c = 1; c < l->nc; ++c) p[c] = p[c] - coef*l->ei[c]*(l->cx[c]*x + l->cy[c]*y);
            nsin_macro(l, p, &r, &x, &y);
        }
        if (issave) { rho[idx] = r; ux[idx] = x; uy[idx] = y; }
        nsin_eq(l, feq, r, x, y);
        relax(l, p, feq, omega, iomega);
        scatter(l, idx, p);
    }
}
void orc_nsin_macro_collide(orc_lattice* l, double* rho, double* ux, double* uy, double* uz, double nu, int issave) {
    (void)uz; nsin_collide(l, rho, ux, uy, nu, NULL, issave);
}
void orc_nsin_macro_brinkman_collide(orc_lattice* l, double* rho, double* ux, double* uy, double* uz, double nu, const double* alpha, int issave) {
    (void)uz; nsin_collide(l, rho, ux, uy, nu, alpha, issave);
}
/* Edge closures: nsincompressible.h:46-94 (SetU), :96-154 (SetRho).  a = normal axis, t = the other one.
 *   SetRho first derives the normal velocity: u_a = -dir*(rho - (f0 + f_{+t} + f_{-t} + 2*(sum of the three populations leaving)))
 *   axis population:  f_in = f_out -dir*2u_a/3
 *   diagonals:        f_in = f_opp -dir*u_a/6 - c_t*0.5*(f_{+t} - f_{-t} - u_t) */
static void nsin_bc_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    nsbc_ctx* b = (nsbc_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double p[NCMAX];
    gather(l, idx, p);
    int t = 1 - axis, v[3] = {0, 0, 0};
    v[t] = 1; int cp = find_dir(l, v[0], v[1], v[2]); v[t] = -1; int cm = find_dir(l, v[0], v[1], v[2]);
    double ua, ut;
    if (!b->setrho) { ua = axis == 0 ? b->v0[gidx] : b->v1[gidx]; ut = axis == 0 ? b->v1[gidx] : b->v0[gidx]; }
    else {
        double s = p[0] + p[cp] + p[cm];
        double o = 0.0; int first = 1;
        for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] == dir) { o = first ? p[c] : o + p[c]; first = 0; }
        double tot = s + 2.0*o;
        ua = dir == -1 ? b->v0[gidx] - tot : -b->v0[gidx] + tot;
        ut = b->v1[gidx];
    }
    double T = p[cp] - p[cm] - ut;
    for (int c = 1; c < l->nc; ++c) {
        if (l->ci[c][axis] != -dir) continue;
        double val;
        if (l->ci[c][t] == 0) val = dir == -1 ? p[l->opp[c]] + 2.0*ua/3.0 : p[l->opp[c]] - 2.0*ua/3.0;
        else {
            val = dir == -1 ? p[l->opp[c]] + ua/6.0 : p[l->opp[c]] - ua/6.0;
            val = l->ci[c][t] > 0 ? val - 0.5*T : val + 0.5*T;
        }
        l->f[IF(l, idx, c)] = val;
    }
}
void orc_nsin_bc_set_u(orc_lattice* l, const double* uxg, const double* uyg, const double* uzg, const int* mask) {
    (void)uzg; nsbc_ctx b = {uxg, uyg, NULL, mask, 0}; for_all_faces(l, nsin_bc_site, &b);
}
void orc_nsin_bc_set_rho(orc_lattice* l, const double* v0, const double* v1, const double* v2, const int* mask) {
    (void)v2; nsbc_ctx b = {v0, v1, NULL, mask, 1}; for_all_faces(l, nsin_bc_site, &b);
}

/* ------------------------------------------------------------------------------------------ */
/* AD (thermal lattice) primitives: advection.h:18-95 (scalar), advection_avx.h:25-102 (AVX)   */
/* Macro: the same order in both (advection.h:18-51 == advection_avx.h:25-56) */
static void ad_macro(const L* l, const double* g, double ux, double uy, double uz, double omegag, double* tem, double* qx, double* qy, double* qz) {
    double t = g[0], x = 0.0, y = 0.0, z = 0.0;
    for (int c = 1; c < l->nc; ++c) {
        t = t + g[c];
        x = x + l->cx[c]*g[c];
        y = y + l->cy[c]*g[c];
        if (l->nd == 3) z = z + l->cz[c]*g[c];
    }
    double coef = 1.0 - 0.5*omegag;
    *tem = t;
    *qx = coef*(x - t*ux);
    *qy = coef*(y - t*uy);
    *qz = l->nd == 3 ? coef*(z - t*uz) : 0.0;
}
/* Equilibrium, AVX order: advection_avx.h:58-74   geq = ei*(tem*(1 + 3*cu)) */
static void ad_eq_avx(const L* l, double* geq, double tem, double ux, double uy, double uz) {
    for (int c = 0; c < l->nc; ++c) geq[c] = l->ei[c]*(tem*(1.0 + 3.0*cdot(l, c, ux, uy, uz)));
}
/* Equilibrium, scalar order: advection.h:53-68    geq = (ei*tem)*(1 + 3*ciu) */
static void ad_eq_sc(const L* l, double* geq, double tem, double ux, double uy, double uz) {
    for (int c = 0; c < l->nc; ++c) geq[c] = l->ei[c]*tem*(1.0 + 3.0*cdot(l, c, ux, uy, uz));
}
/* buoyancy on the flow lattice, AVX order: advection_avx.h:76-92   f += (3*(T - T0))*(ei*(c.g)) */
static void ad_natconv_avx(const L* l, double* p, double tem, double gx, double gy, double gz, double tem0) {
    double coef = 3.0*(tem - tem0);
    for (int c = 1; c < l->nc; ++c) p[c] = p[c] + coef*(l->ei[c]*cdot(l, c, gx, gy, gz));
}
/* ... scalar order: advection.h:70-84   f += ((3*ei)*(c.g))*(T - T0) */
static void ad_natconv_sc(const L* l, double* p, double tem, double gx, double gy, double gz, double tem0) {
    for (int c = 1; c < l->nc; ++c) p[c] += 3.0*l->ei[c]*cdot(l, c, gx, gy, gz)*(tem - tem0);
}

/* AD::InitialCondition: advection.h:1046-1070 (scalar Equilibrium at every site) */
void orc_ad_init(orc_lattice* g, const double* tem, const double* ux, const double* uy, const double* uz) {
    double geq[NCMAX];
    for (int idx = 0; idx < g->nxyz; ++idx) {
        ad_eq_sc(g, geq, tem[idx], ux[idx], uy[idx], g->nd == 3 ? uz[idx] : 0.0);
        scatter(g, idx, geq);
    }
}

/* AD::MacroBrinkmanCollideNaturalConvection: advection_avx.h:886-998 (2-D), 1001-1116 (3-D).
 * Per site: moments of f and g -> buoyancy on f (pre-force T) -> Brinkman on f (pre-force rho, u) -> moments again (g unchanged,
 * new u) -> save rho,u,T,q and the snapshot of g -> relax f, relax g with omega_g = 1/(3 kappa[idx] + 1/2).
 * Snapshot layout: [pack][c][lane] for packed sites (:1047-1052), [idx][c] for the tail (:1093-1098). */
void orc_ad_macro_brinkman_collide_natural_convection(orc_lattice* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        orc_lattice* g, double* tem, double* qx, double* qy, double* qz, const double* diffusivity,
        double gx, double gy, double gz, double tem0, int issave, double* gsnap) {
    const double omegaf = 1.0/(3.0*nu + 0.5), iomegaf = 1.0 - omegaf;
    const int ne = npacked(f), nc = g->nc;
    if (f->nd == 2) gz = 0.0;
    #pragma omp parallel for
    for (int idx = 0; idx < f->nxyz; ++idx) {
        const int tail = idx >= ne;
        const double omegag = 1.0/(3.0*diffusivity[idx] + 0.5), iomegag = 1.0 - omegag;
        double p[NCMAX], q[NCMAX], feq[NCMAX], geq[NCMAX], r, x, y, z, t, hx, hy, hz;
        gather(f, idx, p);
        gather(g, idx, q);
        ns_macro(f, p, &r, &x, &y, &z);
        ad_macro(g, q, x, y, z, omegag, &t, &hx, &hy, &hz);
        if (tail) ad_natconv_sc(f, p, t, gx, gy, gz, tem0); else ad_natconv_avx(f, p, t, gx, gy, gz, tem0);
        ns_brinkman(f, p, r, x, y, z, alpha[idx]);
        ns_macro(f, p, &r, &x, &y, &z);
        ad_macro(g, q, x, y, z, omegag, &t, &hx, &hy, &hz);
        if (issave) {
            rho[idx] = r; ux[idx] = x; uy[idx] = y; tem[idx] = t; qx[idx] = hx; qy[idx] = hy;
            if (f->nd == 3) { uz[idx] = z; qz[idx] = hz; }
            if (gsnap) {
                if (!tail) { const int base = idx - idx%4, lane = idx%4; for (int c = 0; c < nc; ++c) gsnap[(size_t)nc*base + 4*c + lane] = q[c]; }
                else for (int c = 0; c < nc; ++c) gsnap[(size_t)nc*idx + c] = q[c];
            }
        }
        if (tail) ns_eq_sc(f, feq, r, x, y, z); else ns_eq_avx(f, feq, r, x, y, z);
        relax(f, p, feq, omegaf, iomegaf);
        if (tail) ad_eq_sc(g, geq, t, x, y, z); else ad_eq_avx(g, geq, t, x, y, z);
        relax(g, q, geq, omegag, iomegag);
        scatter(f, idx, p);
        scatter(g, idx, q);
    }
}

/* ---- temperature / heat-flux closures on a face: advection.h:99-524 -----------------------------------
 * Face with normal axis a and outward direction dir; "in" = populations entering the domain (c_a == -dir).
 *   SetT: tem0 = 6*(T - g0 - g_c ... over the c with c_a != -dir, ascending)/(1 - dir*3u_a)          (e.g. :106, :159)
 *   SetQ: tem0 = 6*((1 + 1/(6 kappa))*qn + g_c ... over the c with c_a == dir, ascending)/(1 + dir*3u_a) (e.g. :250, :301)
 *   g_in(c) = tem0*(1 +/- 3ux +/- 3uy +/- 3uz)/(9 | 36 | 72), the velocity terms in x,y,z order, absent components skipped */
typedef struct { const double *val, *ux, *uy, *uz, *kfield; double kconst; const int* mask; int setq; } adbc_ctx;
static void ad_bc_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    adbc_ctx* b = (adbc_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double g[NCMAX];
    gather(l, idx, g);
    const double u[3] = {b->ux[idx], b->uy[idx], l->nd == 3 ? b->uz[idx] : 0.0};
    double tem0;
    if (!b->setq) {
        double s = b->val[gidx] - g[0];
        for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] != -dir) s = s - g[c];
        tem0 = dir == -1 ? 6.0*s/(1.0 + 3.0*u[axis]) : 6.0*s/(1.0 - 3.0*u[axis]);
    } else {
        const double kappa = b->kfield ? b->kfield[idx] : b->kconst;
        double s = (1.0 + 1.0/(6.0*kappa))*b->val[gidx];
        for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] == dir) s = s + g[c];
        tem0 = dir == -1 ? 6.0*s/(1.0 - 3.0*u[axis]) : 6.0*s/(1.0 + 3.0*u[axis]);
    }
    for (int c = 1; c < l->nc; ++c) {
        if (l->ci[c][axis] != -dir) continue;
        double t = 1.0;
        for (int d = 0; d < l->nd; ++d) if (l->ci[c][d]) t = l->ci[c][d] > 0 ? t + 3.0*u[d] : t - 3.0*u[d];
        const int nz = abs(l->ci[c][0]) + abs(l->ci[c][1]) + abs(l->ci[c][2]);
        l->f[IF(l, idx, c)] = tem0*t/(nz == 1 ? 9.0 : (l->nd == 2 ? 36.0 : 72.0));
    }
}
void orc_ad_bc_set_t(orc_lattice* g, const double* temg, const double* ux, const double* uy, const double* uz, const int* mask) {
    adbc_ctx b = {temg, ux, uy, uz, NULL, 0.0, mask, 0}; for_all_faces(g, ad_bc_site, &b);
}
void orc_ad_bc_set_q(orc_lattice* g, const double* qng, const double* ux, const double* uy, const double* uz, const double* kfield, double kconst, const int* mask) {
    adbc_ctx b = {qng, ux, uy, uz, kfield, kconst, mask, 1}; for_all_faces(g, ad_bc_site, &b);
}

/* ------------------------------------------------------------------------------------------ */
/* ANS / AAD (adjoint flow / adjoint thermal) primitives                                       */
/* ANS::Macro, AVX order: adjointnavierstokes_avx.h:27-72 (the 2-D and 3-D versions associate the ip term differently) */
static void ans_macro_avx(const L* l, const double* f, double ux, double uy, double uz,
                          double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz) {
    double p = 0.0, ax = 0.0, ay = 0.0, az = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
    const double uu = dot3(l, ux, uy, uz, ux, uy, uz);
    const double one_uu = 1.0 - 1.5*uu;                 /* 2-D: :35 */
    for (int c = 0; c < l->nc; ++c) {
        const double fei = f[c]*l->ei[c];
        const double cu = cdot(l, c, ux, uy, uz);
        if (l->nd == 2) p = p + fei*(one_uu + (3.0*cu + 4.5*(cu*cu)));
        else p = p + fei*(1.0 + (3.0*cu + (4.5*(cu*cu) - 1.5*uu)));
        ax = ax + fei*(l->cx[c] + (3.0*(cu*l->cx[c]) - ux));
        ay = ay + fei*(l->cy[c] + (3.0*(cu*l->cy[c]) - uy));
        if (l->nd == 3) az = az + fei*(l->cz[c] + (3.0*(cu*l->cz[c]) - uz));
        mx = mx + fei*l->cx[c];
        my = my + fei*l->cy[c];
        if (l->nd == 3) mz = mz + fei*l->cz[c];
    }
    *ip = p; *iux = ax; *iuy = ay; *iuz = az; *imx = mx; *imy = my; *imz = mz;
}
/* ANS::Macro, scalar order: adjointnavierstokes.h:17-61 */
static void ans_macro_sc(const L* l, const double* f, double ux, double uy, double uz,
                         double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz) {
    const double uu = l->nd == 3 ? ux*ux + uy*uy + uz*uz : ux*ux + uy*uy;
    double p = f[0]*l->ei[0]*(1.0 - 1.5*uu);
    double ax = -f[0]*l->ei[0]*ux, ay = -f[0]*l->ei[0]*uy, az = l->nd == 3 ? -f[0]*l->ei[0]*uz : 0.0;
    double mx = 0.0, my = 0.0, mz = 0.0;
    for (int c = 1; c < l->nc; ++c) {
        const double ciu = l->nd == 3 ? l->cx[c]*ux + l->cy[c]*uy + l->cz[c]*uz : l->cx[c]*ux + l->cy[c]*uy;
        const double fei = f[c]*l->ei[c];
        p += fei*(1.0 + 3.0*ciu + 4.5*ciu*ciu - 1.5*uu);
        ax += fei*(l->cx[c] + 3.0*ciu*l->cx[c] - ux);
        ay += fei*(l->cy[c] + 3.0*ciu*l->cy[c] - uy);
        if (l->nd == 3) az += fei*(l->cz[c] + 3.0*ciu*l->cz[c] - uz);
        mx += fei*l->cx[c];
        my += fei*l->cy[c];
        if (l->nd == 3) mz += fei*l->cz[c];
    }
    *ip = p; *iux = ax; *iuy = ay; *iuz = az; *imx = mx; *imy = my; *imz = mz;
}
/* ANS::Equilibrium: adjointnavierstokes.h:63-77 == adjointnavierstokes_avx.h:74-88 */
static void ans_eq(const L* l, double* feq, double ux, double uy, double uz, double ip, double iux, double iuy, double iuz) {
    for (int c = 0; c < l->nc; ++c) {
        double s = iux*(l->cx[c] - ux) + iuy*(l->cy[c] - uy);
        if (l->nd == 3) s = s + iuz*(l->cz[c] - uz);
        feq[c] = ip + 3.0*s;
    }
}
/* AAD::Macro: adjointadvection.h:24-50 == adjointadvection_avx.h:203-229 */
static void aad_macro(const L* l, const double* g, double* item, double* iqx, double* iqy, double* iqz) {
    double t = l->ei[0]*g[0], x = 0.0, y = 0.0, z = 0.0;
    for (int c = 1; c < l->nc; ++c) {
        const double gei = l->ei[c]*g[c];
        t = t + gei;
        x = x + gei*l->cx[c];
        y = y + gei*l->cy[c];
        if (l->nd == 3) z = z + gei*l->cz[c];
    }
    *item = t; *iqx = x; *iqy = y; *iqz = z;
}
/* AAD::Equilibrium (one value for every c): adjointadvection.h:52-68 == adjointadvection_avx.h:231-247 */
static double aad_eq(const L* l, double item, double iqx, double iqy, double iqz, double ux, double uy, double uz) {
    return item + 3.0*dot3(l, ux, uy, uz, iqx, iqy, iqz);
}
/* coupling force on the adjoint flow lattice: adjointadvection.h:70-106 / adjointadvection_avx.h:249-279 (same values) */
static void aad_brinkman(const L* l, double* f, double rho, double ux, double uy, double uz, double imx, double imy, double imz,
                         double tem, double iqx, double iqy, double iqz, double omegag, double alpha) {
    const double coef = 3.0/(rho + alpha);
    const double kx = tem*iqx*omegag - alpha*imx, ky = tem*iqy*omegag - alpha*imy, kz = l->nd == 3 ? tem*iqz*omegag - alpha*imz : 0.0;
    f[0] = f[0] - coef*dot3(l, kx, ky, kz, ux, uy, uz);
    for (int c = 1; c < l->nc; ++c) {
        double s = kx*(l->cx[c] - ux) + ky*(l->cy[c] - uy);
        if (l->nd == 3) s = s + kz*(l->cz[c] - uz);
        f[c] = f[c] + coef*s;
    }
}
/* adjoint buoyancy on the adjoint thermal lattice: adjointadvection.h:114-132 == adjointadvection_avx.h:292-306 */
static void aad_natconv(const L* l, double* g, double imx, double imy, double imz, double gx, double gy, double gz) {
    const double coef = 3.0*dot3(l, imx, imy, imz, gx, gy, gz);
    for (int c = 0; c < l->nc; ++c) g[c] = g[c] + coef;
}

/* InitialCondition: adjointnavierstokes.h:474-498, adjointadvection.h:1359-1381 */
void orc_ans_init(orc_lattice* l, const double* ux, const double* uy, const double* uz, const double* ip, const double* iux, const double* iuy, const double* iuz) {
    double feq[NCMAX];
    for (int idx = 0; idx < l->nxyz; ++idx) {
        ans_eq(l, feq, ux[idx], uy[idx], l->nd == 3 ? uz[idx] : 0.0, ip[idx], iux[idx], iuy[idx], l->nd == 3 ? iuz[idx] : 0.0);
        scatter(l, idx, feq);
    }
}
void orc_aad_init(orc_lattice* g, const double* ux, const double* uy, const double* uz, const double* item, const double* iqx, const double* iqy, const double* iqz) {
    double geq[NCMAX];
    for (int idx = 0; idx < g->nxyz; ++idx) {
        const double v = aad_eq(g, item[idx], iqx[idx], iqy[idx], g->nd == 3 ? iqz[idx] : 0.0, ux[idx], uy[idx], g->nd == 3 ? uz[idx] : 0.0);
        for (int c = 0; c < g->nc; ++c) geq[c] = v;
        scatter(g, idx, geq);
    }
}

/* AAD::MacroBrinkmanCollideNaturalConvection: adjointadvection_avx.h:765-880 (2-D), 884-1005 (3-D).
 * Per site: adjoint moments of f and g -> coupling force on f -> adjoint flow moments again -> adjoint buoyancy on g (new im) ->
 * adjoint thermal moments again -> save ip,iu,im,iT,iq and the snapshot of g -> relax f (ANS equilibrium), relax g. */
void orc_aad_macro_brinkman_collide_natural_convection(orc_lattice* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        orc_lattice* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity,
        double gx, double gy, double gz, int issave, double* igsnap) {
    const double omegaf = 1.0/(3.0*nu + 0.5), iomegaf = 1.0 - omegaf;
    const int ne = npacked(f), nc = g->nc;
    if (f->nd == 2) gz = 0.0;
    #pragma omp parallel for
    for (int idx = 0; idx < f->nxyz; ++idx) {
        const int tail = idx >= ne;
        const double omegag = 1.0/(3.0*diffusivity[idx] + 0.5), iomegag = 1.0 - omegag;
        const double r = rho[idx], x = ux[idx], y = uy[idx], z = f->nd == 3 ? uz[idx] : 0.0;
        double p[NCMAX], q[NCMAX], feq[NCMAX], geq[NCMAX], a, bx, by, bz, mx, my, mz, t, hx, hy, hz;
        gather(f, idx, p);
        gather(g, idx, q);
        if (tail) ans_macro_sc(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz); else ans_macro_avx(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz);
        aad_macro(g, q, &t, &hx, &hy, &hz);
        aad_brinkman(f, p, r, x, y, z, mx, my, mz, tem[idx], hx, hy, hz, omegag, alpha[idx]);
        if (tail) ans_macro_sc(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz); else ans_macro_avx(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz);
        aad_natconv(g, q, mx, my, mz, gx, gy, gz);
        aad_macro(g, q, &t, &hx, &hy, &hz);
        if (issave) {
            ip[idx] = a; iux[idx] = bx; iuy[idx] = by; imx[idx] = mx; imy[idx] = my; item[idx] = t; iqx[idx] = hx; iqy[idx] = hy;
            if (f->nd == 3) { iuz[idx] = bz; imz[idx] = mz; iqz[idx] = hz; }
            if (igsnap) {
                if (!tail) { const int base = idx - idx%4, lane = idx%4; for (int c = 0; c < nc; ++c) igsnap[(size_t)nc*base + 4*c + lane] = q[c]; }
                else for (int c = 0; c < nc; ++c) igsnap[(size_t)nc*idx + c] = q[c];
            }
        }
        ans_eq(f, feq, x, y, z, a, bx, by, bz);
        relax(f, p, feq, omegaf, iomegaf);
        const double ge = aad_eq(g, t, hx, hy, hz, x, y, z);
        for (int c = 0; c < nc; ++c) geq[c] = ge;
        relax(g, q, geq, omegag, iomegag);
        scatter(f, idx, p);
        scatter(g, idx, q);
    }
}

/* ---- adjoint thermal closures on a face: adjointadvection.h:154-484 ------------------------------------
 * K = the populations with c_a == -dir, ascending (the axis-aligned one first, then the four / two diagonals): the known ones.
 * Every unknown (the opposites of K) takes one value r.  S_t = sum over the diagonals of sign(c_t(K_i)) g_Ki.
 *   iSetT 2-D: r = -(4 (1 - dir 3u_a) g_K0 + sum_i (1 +/- 3ux +/- 3uy) g_Ki)/(6 (1 - dir 3u_a))                       (:161, :166)
 *   iSetT 3-D: r = -(8 g_K0 + sum g_Ki)/12 - u_t1 S_t1/(4 (1 - dir 3u_a)) - u_t2 S_t2/(4 (1 - dir 3u_a)), t1,t2 = a+1,a+2 cyclic   (:209-211)
 *   iSetQ:     r = ((1 - dir 3u_a)((4|8) g_K0 + sum g_Ki) + w u_t1 S_t1 [+ w u_t2 S_t2] - (12|24) eps)/((6|12)(1 + dir 3u_a)),
 *              w = 3, except that the 3-D ymax, zmin and zmax faces carry the bare u_t (:428-429, :456-457, :468-469) */
static int face_K(const L* l, int axis, int sgn, int* K) {
    int m = 0;
    for (int c = 1; c < l->nc; ++c) if (l->ci[c][axis] == sgn) K[m++] = c;
    return m;
}
static double signed_diag(const L* l, const double* g, const int* K, int m, int t) {
    double s = l->ci[K[1]][t] > 0 ? g[K[1]] : -g[K[1]];
    for (int i = 2; i < m; ++i) s = l->ci[K[i]][t] > 0 ? s + g[K[i]] : s - g[K[i]];
    return s;
}
static double one_pm_3cu(const L* l, int c, const double* u) {
    double t = 1.0;
    for (int d = 0; d < l->nd; ++d) if (l->ci[c][d]) t = l->ci[c][d] > 0 ? t + 3.0*u[d] : t - 3.0*u[d];
    return t;
}
/* (1 - dir 3u_a)(lead + (4|8) g_K0 + sum g_Ki) + w u_t S_t ...: shared by iSetQ and the heat-source sensitivity term */
static double aad_q_bracket(const L* l, const double* g, const int* K, int m, int axis, int dir, const double* u, int with_lead, double lead) {
    const double one3 = dir == -1 ? 1.0 + 3.0*u[axis] : 1.0 - 3.0*u[axis];
    double s = with_lead ? lead + (l->nd == 2 ? 4.0 : 8.0)*g[K[0]] : (l->nd == 2 ? 4.0 : 8.0)*g[K[0]];
    for (int i = 1; i < m; ++i) s = s + g[K[i]];
    double acc = one3*s;
    if (l->nd == 2) {
        const int t = 1 - axis;
        acc = acc + 3.0*u[t]*signed_diag(l, g, K, m, t);
    } else {
        const int three = axis == 0 || (axis == 1 && dir == -1);
        for (int n = 1; n <= 2; ++n) {
            const int t = (axis + n)%3;
            acc = three ? acc + 3.0*u[t]*signed_diag(l, g, K, m, t) : acc + u[t]*signed_diag(l, g, K, m, t);
        }
    }
    return acc;
}
typedef struct { const double *ux, *uy, *uz; const int* mask; double eps; int setq; } aadbc_ctx;
static void aad_ibc_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    aadbc_ctx* b = (aadbc_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double g[NCMAX];
    int K[NCMAX];
    gather(l, idx, g);
    const int m = face_K(l, axis, -dir, K);
    const double u[3] = {b->ux[idx], b->uy[idx], l->nd == 3 ? b->uz[idx] : 0.0};
    double r;
    if (!b->setq) {
        const double one3 = dir == -1 ? 1.0 + 3.0*u[axis] : 1.0 - 3.0*u[axis];
        if (l->nd == 2) {
            double a = 4.0*one3*g[K[0]];
            for (int i = 1; i < m; ++i) a = a + one_pm_3cu(l, K[i], u)*g[K[i]];
            r = -a/(6.0*one3);
        } else {
            double a = 8.0*g[K[0]];
            for (int i = 1; i < m; ++i) a = a + g[K[i]];
            r = -a/12.0;
            for (int n = 1; n <= 2; ++n) { const int t = (axis + n)%3; r = r - u[t]*signed_diag(l, g, K, m, t)/(4.0*one3); }
        }
    } else {
        double acc = aad_q_bracket(l, g, K, m, axis, dir, u, 0, 0.0);
        acc = acc - (l->nd == 2 ? 12.0 : 24.0)*b->eps;
        r = acc/((l->nd == 2 ? 6.0 : 12.0)*(dir == -1 ? 1.0 - 3.0*u[axis] : 1.0 + 3.0*u[axis]));
    }
    for (int i = 0; i < m; ++i) l->f[IF(l, idx, l->opp[K[i]])] = r;
}
void orc_aad_ibc_set_t(orc_lattice* g, const double* ux, const double* uy, const double* uz, const int* mask) {
    aadbc_ctx b = {ux, uy, uz, mask, 0.0, 0}; for_all_faces(g, aad_ibc_site, &b);
}
void orc_aad_ibc_set_q(orc_lattice* g, const double* ux, const double* uy, const double* uz, const int* mask, double eps) {
    aadbc_ctx b = {ux, uy, uz, mask, eps, 1}; for_all_faces(g, aad_ibc_site, &b);
}

/* AAD::SensitivityTemperatureAtHeatSource: adjointadvection_avx.h:1403-1513 (volume terms, AVX order for packed sites and the
 * scalar order with pow() for the tail), :16-185 (heat-source face terms, scalar).  gsnap / igsnap in the reference layout. */
static size_t index_g(const L* l, int idx, int c) {
    const int ne = npacked(l);
    return idx < ne ? (size_t)(idx/4)*4*l->nc + 4*c + idx%4 : (size_t)l->nc*idx + c;
}
typedef struct { double* dfds; const double *ux, *uy, *uz, *ig, *kappa, *dkds, *qn; const int* mask; } sens_ctx;
static void sens_face_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    sens_ctx* b = (sens_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double ig[NCMAX];
    int K[NCMAX];
    for (int c = 0; c < l->nc; ++c) ig[c] = b->ig[index_g(l, idx, c)];
    const int m = face_K(l, axis, -dir, K);
    const double u[3] = {b->ux[idx], b->uy[idx], l->nd == 3 ? b->uz[idx] : 0.0};
    const double e = aad_q_bracket(l, ig, K, m, axis, dir, u, 1, l->nd == 2 ? -6.0 : -12.0);
    const double den = (l->nd == 2 ? 36.0 : 72.0)*(dir == -1 ? 1.0 - 3.0*u[axis] : 1.0 + 3.0*u[axis])*pow(b->kappa[idx], 2.0);
    b->dfds[idx] += b->qn[gidx]*b->dkds[idx]*e/den;
}
void orc_aad_sensitivity_temperature_at_heat_source(orc_lattice* g, double* dfds, const double* ux, const double* uy, const double* uz,
        const double* imx, const double* imy, const double* imz, const double* dads, const double* tem, const double* item,
        const double* iqx, const double* iqy, const double* iqz, const double* gsnap, const double* igsnap, const double* diffusivity, const double* dkds,
        const double* qng, const int* mask) {
    const int ne = npacked(g), nc = g->nc, d3 = g->nd == 3;
    for (int idx = 0; idx < g->nxyz; ++idx) {
        double sumg = 0.0;
        for (int c = 0; c < nc; ++c) sumg = sumg + gsnap[index_g(g, idx, c)]*igsnap[index_g(g, idx, c)];
        if (idx < ne) {
            const double um = d3 ? ux[idx]*imx[idx] + (uy[idx]*imy[idx] + uz[idx]*imz[idx]) : ux[idx]*imx[idx] + uy[idx]*imy[idx];
            double v = dfds[idx] + 3.0*(dads[idx]*um);
            const double taug = 3.0*diffusivity[idx] + 0.5;
            const double uq = d3 ? ux[idx]*iqx[idx] + (uy[idx]*iqy[idx] + uz[idx]*iqz[idx]) : ux[idx]*iqx[idx] + uy[idx]*iqy[idx];
            v = v - (3.0*(dkds[idx]*(sumg - tem[idx]*(item[idx] + 3.0*uq))))/(taug*taug);
            dfds[idx] = v;
        } else {
            if (d3) dfds[idx] += 3.0*dads[idx]*(ux[idx]*imx[idx] + uy[idx]*imy[idx] + uz[idx]*imz[idx]);
            else dfds[idx] += 3.0*dads[idx]*(ux[idx]*imx[idx] + uy[idx]*imy[idx]);
            if (d3) dfds[idx] += -3.0/pow(3.0*diffusivity[idx] + 0.5, 2.0)*dkds[idx]*(sumg - tem[idx]*(item[idx] + 3.0*(ux[idx]*iqx[idx] + uy[idx]*iqy[idx] + uz[idx]*iqz[idx])));
            else dfds[idx] += -3.0/pow(3.0*diffusivity[idx] + 0.5, 2.0)*dkds[idx]*(sumg - tem[idx]*(item[idx] + 3.0*(ux[idx]*iqx[idx] + uy[idx]*iqy[idx])));
        }
    }
    sens_ctx b = {dfds, ux, uy, uz, igsnap, diffusivity, dkds, qng, mask};
    for_all_faces(g, sens_face_site, &b);
}

/* ------------------------------------------------------------------------------------------ */
/* the other collide variants, from the same primitives                                        */
/* heat exchange on the thermal lattice: advection.h:86-95 (scalar), advection_avx.h:94-102 (AVX) */
static void ad_heatex(const L* l, double* g, double tem, double beta, int sc) {
    const double coef = sc ? beta*(1.0 - tem)/(1.0 + beta) : beta*((1.0 - tem)/(1.0 + beta));
    for (int c = 0; c < l->nc; ++c) g[c] = g[c] + l->ei[c]*coef;
}
/* forward family (advection_avx.h:106-1116): moments -> [buoyancy] -> [Brinkman] -> flow moments again if forced -> [heat exchange]
 * -> thermal moments again if the flow was forced -> save -> relax both.  kfield / kconst: per-cell or scalar diffusivity. */
static void ad_collide(L* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
                       L* g, double* tem, double* qx, double* qy, double* qz, const double* kfield, double kconst, const double* beta,
                       int natconv, double gx, double gy, double gz, double tem0, int issave, double* gsnap) {
    const double omegaf = 1.0/(3.0*nu + 0.5), iomegaf = 1.0 - omegaf;
    const int ne = npacked(f), nc = g->nc;
    if (f->nd == 2) gz = 0.0;
    #pragma omp parallel for
    for (int idx = 0; idx < f->nxyz; ++idx) {
        const int tail = idx >= ne;
        const double omegag = 1.0/(3.0*(kfield ? kfield[idx] : kconst) + 0.5), iomegag = 1.0 - omegag;
        double p[NCMAX], q[NCMAX], feq[NCMAX], geq[NCMAX], r, x, y, z, t, hx, hy, hz;
        gather(f, idx, p);
        gather(g, idx, q);
        ns_macro(f, p, &r, &x, &y, &z);
        ad_macro(g, q, x, y, z, omegag, &t, &hx, &hy, &hz);
        if (natconv) { if (tail) ad_natconv_sc(f, p, t, gx, gy, gz, tem0); else ad_natconv_avx(f, p, t, gx, gy, gz, tem0); }
        if (alpha) ns_brinkman(f, p, r, x, y, z, alpha[idx]);
        if (natconv || alpha) ns_macro(f, p, &r, &x, &y, &z);
        if (beta) ad_heatex(g, q, t, beta[idx], tail);
        if (natconv || alpha) ad_macro(g, q, x, y, z, omegag, &t, &hx, &hy, &hz);
        if (issave) {
            rho[idx] = r; ux[idx] = x; uy[idx] = y; tem[idx] = t; qx[idx] = hx; qy[idx] = hy;
            if (f->nd == 3) { uz[idx] = z; qz[idx] = hz; }
            if (gsnap) {
                if (!tail) { const int base = idx - idx%4, lane = idx%4; for (int c = 0; c < nc; ++c) gsnap[(size_t)nc*base + 4*c + lane] = q[c]; }
                else for (int c = 0; c < nc; ++c) gsnap[(size_t)nc*idx + c] = q[c];
            }
        }
        if (tail) ns_eq_sc(f, feq, r, x, y, z); else ns_eq_avx(f, feq, r, x, y, z);
        relax(f, p, feq, omegaf, iomegaf);
        if (tail) ad_eq_sc(g, geq, t, x, y, z); else ad_eq_avx(g, geq, t, x, y, z);
        relax(g, q, geq, omegag, iomegag);
        scatter(f, idx, p);
        scatter(g, idx, q);
    }
}
void orc_ad_macro_collide_force_convection(orc_lattice* f, double* rho, double* ux, double* uy, double* uz, double nu,
        orc_lattice* g, double* tem, double* qx, double* qy, double* qz, double diffusivity, int issave) {
    ad_collide(f, rho, ux, uy, uz, NULL, nu, g, tem, qx, qy, qz, NULL, diffusivity, NULL, 0, 0, 0, 0, 0, issave, NULL);
}
void orc_ad_macro_collide_natural_convection(orc_lattice* f, double* rho, double* ux, double* uy, double* uz, double nu,
        orc_lattice* g, double* tem, double* qx, double* qy, double* qz, double diffusivity,
        double gx, double gy, double gz, double tem0, int issave) {
    ad_collide(f, rho, ux, uy, uz, NULL, nu, g, tem, qx, qy, qz, NULL, diffusivity, NULL, 1, gx, gy, gz, tem0, issave, NULL);
}
void orc_ad_macro_brinkman_collide_heat_exchange(orc_lattice* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        orc_lattice* g, double* tem, double* qx, double* qy, double* qz, const double* beta, double diffusivity, int issave) {
    ad_collide(f, rho, ux, uy, uz, alpha, nu, g, tem, qx, qy, qz, NULL, diffusivity, beta, 0, 0, 0, 0, 0, issave, NULL);
}
void orc_ad_macro_brinkman_collide_force_convection(orc_lattice* f, double* rho, double* ux, double* uy, double* uz, const double* alpha, double nu,
        orc_lattice* g, double* tem, double* qx, double* qy, double* qz, const double* diffusivity, int issave, double* gsnap) {
    ad_collide(f, rho, ux, uy, uz, alpha, nu, g, tem, qx, qy, qz, diffusivity, 0.0, NULL, 0, 0, 0, 0, 0, issave, gsnap);
}

/* ANS::ExternalForceBrinkman: adjointnavierstokes_avx.h:90-110 (AVX), adjointnavierstokes.h:79-93 (scalar) */
static void ans_brinkman(const L* l, double* f, double rho, double ux, double uy, double uz, double imx, double imy, double imz, double alpha, int sc) {
    double coef;
    if (!sc) { coef = 3.0*(alpha/(rho + alpha)); f[0] = f[0] + coef*dot3(l, ux, uy, uz, imx, imy, imz); }
    else { coef = 3.0*alpha/(rho + alpha); f[0] -= -coef*dot3(l, ux, uy, uz, imx, imy, imz); }
    for (int c = 1; c < l->nc; ++c) {
        double s = (l->cx[c] - ux)*imx + (l->cy[c] - uy)*imy;
        if (l->nd == 3) s = s + (l->cz[c] - uz)*imz;
        f[c] = f[c] - coef*s;
    }
}
/* adjoint heat exchange: adjointadvection.h:108-112 (scalar), adjointadvection_avx.h:281-290 (AVX) */
static void aad_heatex(const L* l, double* g, double item, double beta, int sc) {
    const double coef = sc ? beta*(1.0 + item)/(1.0 + beta) : beta*((1.0 + item)/(1.0 + beta));
    for (int c = 0; c < l->nc; ++c) g[c] = g[c] - coef;
}
/* mass-flow objective force: adjointadvection.h:134-150 == adjointadvection_avx.h:308-322 (same values) */
static void aad_massflow(const L* l, double* f, double rho, double ux, double uy, double uz, double dx, double dy, double dz) {
    for (int c = 0; c < l->nc; ++c) {
        double s = (l->cx[c] - ux)*dx + (l->cy[c] - uy)*dy;
        if (l->nd == 3) s = s + (l->cz[c] - uz)*dz;
        f[c] = f[c] - s/rho;
    }
}
/* adjoint family (adjointnavierstokes_avx.h:114-259, adjointadvection_avx.h:326-1255): adjoint moments -> [mass flow] -> coupling /
 * Brinkman force on f -> adjoint flow moments again -> [heat exchange | buoyancy on g -> adjoint thermal moments again] -> save ->
 * relax.  g == NULL: the flow-only ANS::MacroBrinkmanCollide. */
static void aad_collide(L* f, const double* rho, const double* ux, const double* uy, const double* uz,
                        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
                        L* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* kfield, double kconst, const double* beta,
                        int natconv, double gx, double gy, double gz, const double* dirx, const double* diry, int issave, double* igsnap, int skip_iuz_tail) {
    const double omegaf = 1.0/(3.0*nu + 0.5), iomegaf = 1.0 - omegaf;
    const int ne = npacked(f), nc = f->nc;
    if (f->nd == 2) gz = 0.0;
    #pragma omp parallel for
    for (int idx = 0; idx < f->nxyz; ++idx) {
        const int tail = idx >= ne;
        const double omegag = g ? 1.0/(3.0*(kfield ? kfield[idx] : kconst) + 0.5) : 0.0, iomegag = 1.0 - omegag;
        const double r = rho[idx], x = ux[idx], y = uy[idx], z = f->nd == 3 ? uz[idx] : 0.0;
        double p[NCMAX], q[NCMAX], feq[NCMAX], geq[NCMAX], a, bx, by, bz, mx, my, mz, t = 0.0, hx = 0.0, hy = 0.0, hz = 0.0;
        gather(f, idx, p);
        if (g) gather(g, idx, q);
        if (tail) ans_macro_sc(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz); else ans_macro_avx(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz);
        if (g) aad_macro(g, q, &t, &hx, &hy, &hz);
        if (dirx) aad_massflow(f, p, r, x, y, z, dirx[idx], diry[idx], 0.0);
        if (g) aad_brinkman(f, p, r, x, y, z, mx, my, mz, tem[idx], hx, hy, hz, omegag, alpha[idx]);
        else ans_brinkman(f, p, r, x, y, z, mx, my, mz, alpha[idx], tail);
        if (tail) ans_macro_sc(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz); else ans_macro_avx(f, p, x, y, z, &a, &bx, &by, &bz, &mx, &my, &mz);
        if (beta) aad_heatex(g, q, t, beta[idx], tail);
        if (natconv) aad_natconv(g, q, mx, my, mz, gx, gy, gz);
        if (g && (beta || natconv)) aad_macro(g, q, &t, &hx, &hy, &hz);
        if (issave) {
            ip[idx] = a; iux[idx] = bx; iuy[idx] = by; imx[idx] = mx; imy[idx] = my;
            if (f->nd == 3) { if (!(tail && skip_iuz_tail)) iuz[idx] = bz; imz[idx] = mz; }
            if (g) {
                item[idx] = t; iqx[idx] = hx; iqy[idx] = hy;
                if (f->nd == 3) iqz[idx] = hz;
                if (igsnap) {
                    if (!tail) { const int base = idx - idx%4, lane = idx%4; for (int c = 0; c < nc; ++c) igsnap[(size_t)nc*base + 4*c + lane] = q[c]; }
                    else for (int c = 0; c < nc; ++c) igsnap[(size_t)nc*idx + c] = q[c];
                }
            }
        }
        ans_eq(f, feq, x, y, z, a, bx, by, bz);
        relax(f, p, feq, omegaf, iomegaf);
        scatter(f, idx, p);
        if (g) {
            const double ge = aad_eq(g, t, hx, hy, hz, x, y, z);
            for (int c = 0; c < nc; ++c) geq[c] = ge;
            relax(g, q, geq, omegag, iomegag);
            scatter(g, idx, q);
        }
    }
}
void orc_ans_macro_brinkman_collide(orc_lattice* l, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, double nu, const double* alpha, int issave) {
    aad_collide(l, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0.0, NULL, 0, 0, 0, 0, NULL, NULL, issave, NULL, 0);
}
void orc_aad_macro_brinkman_collide_heat_exchange(orc_lattice* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        orc_lattice* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* beta, double diffusivity, int issave) {
    aad_collide(f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz, NULL, diffusivity, beta, 0, 0, 0, 0, NULL, NULL, issave, NULL, 0);
}
/* quirk: the 3-D scalar tail of this one does not store _iuz (adjointadvection_avx.h:725-735) */
void orc_aad_macro_brinkman_collide_force_convection(orc_lattice* f, const double* rho, const double* ux, const double* uy, const double* uz,
        double* ip, double* iux, double* iuy, double* iuz, double* imx, double* imy, double* imz, const double* alpha, double nu,
        orc_lattice* g, const double* tem, double* item, double* iqx, double* iqy, double* iqz, const double* diffusivity, int issave, double* igsnap) {
    aad_collide(f, rho, ux, uy, uz, ip, iux, iuy, iuz, imx, imy, imz, alpha, nu, g, tem, item, iqx, iqy, iqz, diffusivity, 0.0, NULL, 0, 0, 0, 0, NULL, NULL, issave, igsnap, 1);
}
void orc_aad_macro_brinkman_collide_natural_convection_massflow(orc_lattice* f, const double* rho, const double* ux, const double* uy,
        double* ip, double* iux, double* iuy, double* imx, double* imy, const double* alpha, double nu,
        orc_lattice* g, const double* tem, double* item, double* iqx, double* iqy, const double* diffusivity,
        double gx, double gy, const double* dirx, const double* diry, int issave, double* igsnap) {
    aad_collide(f, rho, ux, uy, NULL, ip, iux, iuy, NULL, imx, imy, NULL, alpha, nu, g, tem, item, iqx, iqy, NULL, diffusivity, 0.0, NULL, 1, gx, gy, 0.0, dirx, diry, issave, igsnap, 0);
}

/* ---- adjoint flow closures: adjointnavierstokes.h:97-392; adjoint thermal-flow coupling closure (D2Q9): adjointadvection.h:488-575 */
static double weighted_face_sum(const double* p, const int* K, int m, double w) {
    double s = w*p[K[0]];
    for (int i = 1; i < m; ++i) s = s + p[K[i]];
    return s;
}
typedef struct { const double *v0, *v1, *v2; const int* mask; double eps; int setrho; } ansbc_ctx;
static void ans_ibc_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    ansbc_ctx* b = (ansbc_ctx*)vctx;
    if (!b->mask[gidx]) return;
    double f[NCMAX], nv[NCMAX];
    int K[NCMAX];
    gather(l, idx, f);
    const int m = face_K(l, axis, -dir, K);
    double rho0;
    if (b->setrho) {        /* iSetRho: rho0 = ((4|8) f_K0 + sum f_Ki)/(3|6), f_opp(K) = f_K - rho0   (:258-392) */
        rho0 = l->nd == 2 ? weighted_face_sum(f, K, m, 4.0)/3.0 : weighted_face_sum(f, K, m, 8.0)/6.0;
        for (int i = 0; i < m; ++i) nv[i] = f[K[i]] - rho0;
    } else {                /* iSetU (:97-254); the 2-D y-edge version reads ux where uy is meant (:134, :139): reproduced */
        double u[3] = {b->v0[gidx], b->v1[gidx], l->nd == 3 ? b->v2[gidx] : 0.0};
        if (l->nd == 2 && axis == 1) u[1] = u[0];
        const double ua = u[axis];
        double acc;
        if (l->nd == 2) {
            acc = -2.0*b->eps;
            const double tn = ua*weighted_face_sum(f, K, m, 4.0);
            acc = dir == -1 ? acc + tn : acc - tn;
            const int t = 1 - axis;
            const double ut = axis == 1 ? u[0] : u[1];
            acc = acc + 3.0*ut*signed_diag(l, f, K, m, t);
            rho0 = dir == -1 ? acc/(3.0*(1.0 - ua)) : acc/(3.0*(1.0 + ua));
        } else {
            acc = -4.0*b->eps;
            for (int t = 0; t < 3; ++t) {
                if (t == axis) { const double tn = ua*weighted_face_sum(f, K, m, 8.0); acc = dir == -1 ? acc + tn : acc - tn; }
                else acc = acc + 3.0*u[t]*signed_diag(l, f, K, m, t);
            }
            rho0 = dir == -1 ? acc/(6.0*(1.0 - ua)) : acc/(6.0*(1.0 + ua));
        }
        for (int i = 0; i < m; ++i) nv[i] = f[K[i]] + rho0;
    }
    for (int i = 0; i < m; ++i) l->f[IF(l, idx, l->opp[K[i]])] = nv[i];
}
void orc_ans_ibc_set_u(orc_lattice* l, const double* uxg, const double* uyg, const double* uzg, const int* mask, double eps) {
    ansbc_ctx b = {uxg, uyg, uzg, mask, eps, 0}; for_all_faces(l, ans_ibc_site, &b);
}
void orc_ans_ibc_set_rho(orc_lattice* l, const int* mask) {
    ansbc_ctx b = {NULL, NULL, NULL, mask, 0.0, 1}; for_all_faces(l, ans_ibc_site, &b);
}
/* AAD::iBoundaryConditionSetRho (D2Q9): the mask value selects the thermal closure the edge carries (1 = SetT, 2 = SetQ) */
typedef struct { L* g; const double *rho, *ux, *uy, *tem; const int* mask; double eps; } aadrho_ctx;
static void aad_isetrho_site(L* l, int idx, int gidx, int axis, int dir, void* vctx) {
    aadrho_ctx* b = (aadrho_ctx*)vctx;
    const int kind = b->mask[gidx];
    if (!kind) return;
    double f[NCMAX], g[NCMAX], nv[3];
    int K[NCMAX];
    gather(l, idx, f);
    gather(b->g, idx, g);
    const int m = face_K(l, axis, -dir, K);
    const double ua = axis == 0 ? b->ux[idx] : b->uy[idx], ut = axis == 0 ? b->uy[idx] : b->ux[idx];
    const double rho0 = -weighted_face_sum(f, K, m, 4.0)/3.0;
    const double sd = signed_diag(l, g, K, m, 1 - axis);
    const double onep = dir == -1 ? 1.0 + 3.0*ua : 1.0 - 3.0*ua;
    const double onem = dir == -1 ? 1.0 - 3.0*ua : 1.0 + 3.0*ua;
    double flux0 = 0.0;
    if (kind == 1) flux0 = b->tem[idx]*ut*sd/(2.0*onep*b->rho[idx]);
    else if (kind == 2) flux0 = -b->tem[idx]*(weighted_face_sum(g, K, m, 4.0)/3.0 + ut*sd/2.0)/(onem*b->rho[idx]);
    const double obj0 = b->eps*2.0*b->tem[idx]/(onem*b->rho[idx]);
    for (int i = 0; i < m; ++i) nv[i] = f[K[i]] + rho0 + flux0 + obj0;
    for (int i = 0; i < m; ++i) l->f[IF(l, idx, l->opp[K[i]])] = nv[i];
}
void orc_aad_ibc_set_rho(orc_lattice* f, orc_lattice* g, const double* rho, const double* ux, const double* uy, const double* tem, const int* mask, double eps) {
    aadrho_ctx b = {g, rho, ux, uy, tem, mask, eps}; for_all_faces(f, aad_isetrho_site, &b);
}

/* ---- the other sensitivities: adjointnavierstokes_avx.h:262-293, adjointadvection_avx.h:1257-1401 */
void orc_ans_sensitivity_brinkman(orc_lattice* l, double* dfds, const double* ux, const double* uy, const double* uz,
        const double* imx, const double* imy, const double* imz, const double* dads) {
    const int ne = npacked(l), d3 = l->nd == 3;
    for (int idx = 0; idx < l->nxyz; ++idx) {
        if (idx < ne) {
            const double um = d3 ? ux[idx]*imx[idx] + (uy[idx]*imy[idx] + uz[idx]*imz[idx]) : ux[idx]*imx[idx] + uy[idx]*imy[idx];
            dfds[idx] = dfds[idx] + 3.0*(dads[idx]*um);
        } else if (d3) dfds[idx] += 3.0*dads[idx]*(ux[idx]*imx[idx] + uy[idx]*imy[idx] + uz[idx]*imz[idx]);
        else dfds[idx] += 3.0*dads[idx]*(ux[idx]*imx[idx] + uy[idx]*imy[idx]);
    }
}
void orc_aad_sensitivity_heat_exchange(orc_lattice* g, double* dfds, const double* ux, const double* uy, const double* uz,
        const double* imx, const double* imy, const double* imz, const double* dads, const double* tem, const double* item, const double* dbds) {
    for (int idx = 0; idx < g->nxyz; ++idx) {
        double d = ux[idx]*imx[idx] + uy[idx]*imy[idx];
        if (g->nd == 3) d = d + uz[idx]*imz[idx];
        dfds[idx] = dfds[idx] + (3.0*dads[idx]*d - dbds[idx]*(1.0 - tem[idx])*(1.0 + item[idx]));
    }
}
void orc_aad_sensitivity_brinkman_diffusivity(orc_lattice* g, double* dfds, const double* ux, const double* uy, const double* uz,
        const double* imx, const double* imy, const double* imz, const double* dads, const double* tem, const double* item,
        const double* iqx, const double* iqy, const double* iqz, const double* gsnap, const double* igsnap, const double* diffusivity, const double* dkds) {
    /* the volume terms of SensitivityTemperatureAtHeatSource without the face terms (adjointadvection_avx.h:1302-1401) */
    int* nomask = (int*)calloc((size_t)g->lx*g->ly*g->lz, sizeof(int));
    orc_aad_sensitivity_temperature_at_heat_source(g, dfds, ux, uy, uz, imx, imy, imz, dads, tem, item, iqx, iqy, iqz, gsnap, igsnap, diffusivity, dkds, NULL, nomask);
    free(nomask);
}

/* ------------------------------------------------------------------------------------------ */
/* utilities: residual.h:8-50, normalize.h:8-24 (serial loops)                                  */
double orc_residual3(const double* ux, const double* uy, const double* uz, const double* uxp, const double* uyp, const double* uzp, int n) {
    double unorm = 0.0, dunorm = 0.0;
    for (int i = 0; i < n; ++i) {
        unorm += pow(ux[i], 2.0) + pow(uy[i], 2.0) + pow(uz[i], 2.0);
        dunorm += pow(ux[i] - uxp[i], 2.0) + pow(uy[i] - uyp[i], 2.0) + pow(uz[i] - uzp[i], 2.0);
    }
    return sqrt(dunorm/unorm);
}
double orc_residual2(const double* ux, const double* uy, const double* uxp, const double* uyp, int n) {
    double unorm = 0.0, dunorm = 0.0;
    for (int i = 0; i < n; ++i) {
        unorm += pow(ux[i], 2.0) + pow(uy[i], 2.0);
        dunorm += pow(ux[i] - uxp[i], 2.0) + pow(uy[i] - uyp[i], 2.0);
    }
    return sqrt(dunorm/unorm);
}
double orc_residual1(const double* ux, const double* uxp, int n) {
    double unorm = 0.0, dunorm = 0.0;
    for (int i = 0; i < n; ++i) { dunorm += pow(ux[i] - uxp[i], 2.0); unorm += pow(ux[i], 2.0); }
    return sqrt(dunorm/unorm);
}
void orc_normalize(double* v, int n) {
    double m = 0.0;
    for (int i = 0; i < n; ++i) if (m < fabs(v[i])) m = fabs(v[i]);
    for (int i = 0; i < n; ++i) v[i] /= m;
}

/* ------------------------------------------------------------------------------------------ */
/* cpu_baseline: the test/cavityflow3D.cpp:44-59 call sequence on the oracle port              */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9*ts.tv_nsec; }
double orc_time_cavity3d(int lx, int ly, int lz, int steps, int warmup, double* rho, double* ux, double* uy, double* uz) {
    L* l = orc_lattice_create(3, lx, ly, lz, 0, 1, 1, 1);
    size_t n = (size_t)l->nxyz;
    int* bct = (int*)malloc(n*sizeof(int)); int* lid = (int*)malloc(n*sizeof(int));
    double* uxg = (double*)malloc(n*sizeof(double)); double* uyg = (double*)malloc(n*sizeof(double)); double* uzg = (double*)malloc(n*sizeof(double));
    double u0 = 0.1, theta = 90.0, nu = 0.1;
    for (int k = 0; k < lz; ++k) for (int j = 0; j < ly; ++j) for (int i = 0; i < lx; ++i) {
        size_t g = i + (size_t)lx*(j + (size_t)ly*k);
        bct[g] = (i == 0 || i == lx - 1 || j == 0 || j == ly - 1 || k == 0) ? 1 : 0;
        lid[g] = k == lz - 1;
        uxg[g] = u0*cos(theta*M_PI/180.0); uyg[g] = u0*sin(theta*M_PI/180.0); uzg[g] = 0.0;
        rho[g] = 1.0; ux[g] = 0.0; uy[g] = 0.0; uz[g] = 0.0;
    }
    orc_ns_init(l, rho, ux, uy, uz);
    double t0 = 0.0;
    for (int t = 0; t < warmup + steps; ++t) {
        if (t == warmup) t0 = now_s();
        orc_ns_macro_collide(l, rho, ux, uy, uz, nu, 1);
        orc_stream(l);
        orc_bc(l, bct, 0);
        orc_ns_bc_set_u(l, uxg, uyg, uzg, lid);
        orc_smooth_corner(l);
    }
    double t1 = now_s();
    free(bct); free(lid); free(uxg); free(uyg); free(uzg);
    orc_lattice_destroy(l);
    return t1 - t0;
}
