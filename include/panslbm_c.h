/* panslbm_c.h — C-ABI of libpanslbm_b200.so: the B200 (sm_100a) implementation of the PANSLBM2
 * lattice-Boltzmann forward/adjoint sweep.
 *
 * The reference (PANFACTORY/PANSLBM2) has no FFI: its "plugin API" is the header-only C++ template surface
 * in src/particle and src/equation.  The drop-in headers in panslbm2_b200/src/ keep that surface (same names,
 * argument order and defaults) and forward every call to the entry points below; each entry point cites the
 * reference interface it replaces.  Plain C types only; every per-site array argument is a DEVICE pointer to
 * nxyz doubles unless a parameter is explicitly named *_host.  All calls are asynchronous on the library's
 * stream unless they return data to the host.  Single-threaded callers, one CUDA device per process (rank).
 *
 * Return convention: 0 = success, non-zero = failure with a message available from pl_last_error().
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with PL_ERR_CUDA.
 */
#ifndef PANSLBM_C_H
#define PANSLBM_C_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PL_OK 0
#define PL_ERR_ARG 1
#define PL_ERR_CUDA 2
#define PL_ERR_UNSUPPORTED 3

typedef struct pl_lattice pl_lattice;
typedef struct pl_bc pl_bc;
typedef struct pl_plan pl_plan;

/* ---- library ------------------------------------------------------------------------------- */
const char* pl_last_error(void);
/* != 0 while the calling thread is inside an entry point that passes caller host pointers to the CUDA runtime (used by the
 * coherence layer of the host-pointer surface to refuse serving a page fault from in there) */
int pl_in_call(void);
const char* pl_version(void);
int pl_device_count(void);
int pl_set_device(int device);
int pl_synchronize(void);
/* The CUstream all work is queued on (as a void*); callers that share buffers with another runtime
 * (e.g. torch) must order against it. pl_set_stream(NULL) restores the library's own stream. */
void* pl_get_stream(void);
int pl_set_stream(void* custream);
/* Number of kernels this library has launched since load / since the last reset (bench.py "gpu_launches"). */
uint64_t pl_launch_count(void);
void pl_launch_count_reset(void);

/* ---- communicator: replaces MPI_COMM_WORLD of the reference's _USE_MPI_DEFINES build --------------------------
 * One process per GPU; rank == PEid of every lattice created afterwards; nranks == mx*my*mz.  The 128-byte id is
 * NCCL's unique id: rank 0 obtains it and the host program distributes it (torch.distributed, MPI, a file ...).
 * With a communicator, Stream()/iStream() of a decomposed lattice exchange the outgoing populations with the 26
 * neighbours of the periodic PE grid (d3q15.h:1306-1407: 5 per face site, 2 per edge site, 1 per corner; d2q9.h:590-630)
 * by ncclSend/ncclRecv on a dedicated stream, and pl_residual / pl_normalize reduce over all ranks as the reference's MPI
 * build does (residual.h:16, normalize.h:17). */
int pl_comm_unique_id(char* out128);
int pl_comm_init(const char* id128, int rank, int nranks);
/* Process-local world for parity tests on ONE device: the blocks of all `nranks` PEs are created in this process and
 * exchange through device memory.  The caller must advance the ranks in lockstep (same call sequence on every rank, one
 * call at a time per rank).  The n-th lattice created with a PEid pairs with the n-th lattice of every other PEid. */
int pl_comm_init_loopback(int nranks);
int pl_comm_destroy(void);
int pl_comm_info(int* mode, int* rank, int* nranks);   /* mode: 0 none, 1 NCCL, 2 loopback */
/* MPI_Allreduce(MPI_IN_PLACE, v, n<=4, MPI_DOUBLE, op) of the drivers (heatsink3D.cpp:136, 236, 272): op 0 = SUM, 1 = MAX */
int pl_comm_allreduce(double* inout_host, int n, int op);
/* The general form (any count; MMA's distributed sums in src/utility/mma.h:256-585 reduce m and (m+1)^2 doubles, the drivers also
 * reduce ints: heatsink3D.cpp:302-305): dtype 0 = double, 1 = int32; op 0 = SUM, 1 = MAX, 2 = MIN.  In place, host memory. */
int pl_comm_allreduce_v(void* inout_host, size_t n, int dtype, int op);
/* A batch of point-to-point messages between HOST buffers, matched per pair of ranks in issue order (what MPI_Isend /
 * MPI_Irecv / MPI_Waitall of the reference's VTK writer, src/utility/vtkxmlexport.h:172-214, amount to).  Blocking. */
typedef struct pl_p2p_op { void* host; size_t bytes; int peer; int is_send; } pl_p2p_op;
int pl_comm_p2p(const pl_p2p_op* ops, int n);
/* Pure host arithmetic (no device needed): the messages rank `peid` sends per Stream (inverse=0) / iStream (1), in issue
 * order.  out: 16 ints per message = code, peer, region sites, npop, pop[5], base, s1, s2, n1, n2, code of the message
 * received in the same step, 0; the count in doubles is region sites * npop.  Every rank sends message `code` to `peer`
 * and receives the message of equal size from the peer of the opposite code. */
int pl_halo_describe(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz, int inverse, int* out, int* count);

/* ---- device arrays (caller-owned macroscopic fields: `new double[nxyz]` in the drivers,
 *      e.g. production/heatsink3D.cpp:50-59) ------------------------------------------------- */
double* pl_array_alloc(size_t n);                 /* NULL on failure */
int pl_array_free(double* dev);
int pl_array_upload(double* dev, const double* host, size_t n);     /* synchronous wrt the host buffer */
int pl_array_download(double* host, const double* dev, size_t n);   /* synchronises the stream */
/* Copies on the library's own copy stream, beside the kernels (host buffers should be pinned, else CUDA serialises them).
 * An optimisation iteration moves whole fields between the host optimiser and the sweep (production/heatsink3D.cpp:114-119:
 * alpha, diffusivity, dads, dkds in; :227-246: tem, dfdss out); only alpha/diffusivity gate the first step.
 *   upload_async    starts after everything queued so far (earlier kernels may still read `dev`); kernels queued after the
 *                   next pl_copy_fence() see the data
 *   download_async  starts after everything queued so far has finished (the data are final); the host buffer is valid after
 *                   pl_copy_wait() */
int pl_array_upload_async(double* dev, const double* host, size_t n);
int pl_array_download_async(double* host, const double* dev, size_t n);
int pl_copy_fence(void);     /* the compute stream waits (on the device) for the copies issued so far */
int pl_copy_wait(void);      /* the host waits for the copies issued so far */
int pl_array_fill(double* dev, double value, size_t n);

/* Operation order.  The reference selects its arithmetic per program: with `#define _USE_AVX_DEFINES` before the includes
 * (production/heatsink3D.cpp:2 and every other program but one) the __m256d overloads of the src/equation_avx headers handle the first
 * 4*(nxyz/4) sites and scalar tail code inside those files the rest; without it (production/nsopt.cpp:2) the scalar templates of
 * the src/equation headers handle every site.  The two differ in the association of a few sums (results agree to rounding) and in one
 * place in what they store: the 2-D tail of NS::MacroBrinkmanCollide saves rho, u BEFORE the Brinkman force
 * (navierstokes_avx.h:246-254), the scalar template after it (navierstokes.h:494-503).  The default here is the order of the
 * _USE_AVX_DEFINES build; pl_set_scalar_order(1) — process-wide, before the first lattice is created; the drop-in headers call
 * it when they are compiled without _USE_AVX_DEFINES — makes every site a scalar-order site (they all take the boundary-pass
 * kernels: correct, bit-identical to the reference's scalar build, several times slower than the packed path).  One known
 * exception: the reference's scalar 3-D SensitivityTemperatureAtHeatSource passes `_uz` and `_ig` to its face helpers in swapped
 * order (adjointadvection.h:1536 vs :805) and reads out of bounds; the intended expression (= the AVX overload) is computed. */
int pl_set_scalar_order(int on);
int pl_scalar_order(void);

/* ---- lattices: D2Q9<double> (src/particle/d2q9.h:24-158), D3Q15<double> (src/particle/d3q15.h:24-249) ---- */
#define PL_D2Q9 2
#define PL_D3Q15 3
/* ctor (d2q9.h:28, d3q15.h:28): same block-decomposition rule (d3q15.h:29-35). For PL_D2Q9 pass lz=mz=1. */
pl_lattice* pl_lattice_create(int kind, int lx, int ly, int lz, int peid, int mx, int my, int mz);
int pl_lattice_destroy(pl_lattice*);
/* out[0..17] = lx ly lz PEid mx my mz PEx PEy PEz nx ny nz nxyz offsetx offsety offsetz nc (d3q15.h:222-223) */
int pl_lattice_info(const pl_lattice*, int* out18);
/* Host view in the reference layout f0[nxyz], f[(nc-1)*idx + (c-1)] (d3q15.h:142-144, public members f0/f:225).
 * Device storage is fp64 SoA [c][nxyz]; these two convert (test/d2q9.cpp, test/d3q15.cpp touch f0/f directly). */
int pl_lattice_set_host(pl_lattice*, const double* f0_host, const double* f_host);
int pl_lattice_get_host(pl_lattice*, double* f0_host, double* f_host);
/* Phase of the populations: 1 = pre-collision (after InitialCondition / Stream and its closures), 0 = just collided. */
int pl_lattice_streamed(const pl_lattice*);
/* Device-side checkpoint of a lattice's populations (with their layout and phase): the building block of checkpoint-recompute
 * for transient adjoints.  The reference's transient drivers keep the macroscopic fields and the thermal snapshot of EVERY time
 * step (production/heatsink3D_transient.cpp:50-57: 23 doubles per site and step) and walk them backwards (:190-215); with a
 * checkpoint of both lattices every K steps the states in between can be recomputed segment by segment during the adjoint loop
 * (panslbm2_b200/transient.py), which holds nt/K + K states instead of nt.  save / restore are stream-ordered device copies;
 * restore invalidates what plans and the halo exchange derived from the old content.  On a decomposed lattice every rank must
 * restore at the same point of its loop. */
typedef struct pl_checkpoint pl_checkpoint;
pl_checkpoint* pl_checkpoint_create(const pl_lattice*);
int pl_checkpoint_save(pl_checkpoint*, const pl_lattice*);
int pl_checkpoint_restore(const pl_checkpoint*, pl_lattice*);
int pl_checkpoint_destroy(pl_checkpoint*);
/* Device SoA view of the current populations: c-th plane at base + c*pitch (pitch in doubles). */
int pl_lattice_device_view(pl_lattice*, double** base, size_t* pitch);
/* Population memory.  A lattice owns ONE buffer of nc*pitch doubles (the reference keeps two, f and the hidden fnext, d3q15.h:41-45,
 * 238): the fused passes of a plan update it in place.  Only the operations that cannot — a standalone Stream()/iStream(), bringing
 * a lattice a fused pass left in its streamed layout back to the natural one — borrow a spare buffer, shared by all lattices of one
 * shape and given back to the device once a few in-place passes have gone by without a borrow (or by pl_memory_trim).
 * out[0] = bytes of population buffers owned by live lattices, out[1] = spare bytes held right now, out[2] = borrows so far,
 * out[3] = layout conversions so far. */
int pl_memory_stats(uint64_t* out4);
int pl_memory_trim(void);

/* Stream()/iStream() single-rank path (d3q15.h:601-616, 964-979; d2q9.h:284-295): pull with periodic wrap. */
int pl_stream(pl_lattice*, int inverse);
/* SmoothCorner() (d3q15.h:199-220, 1242-1303; d2q9.h:127-132, 578-587) */
int pl_smooth_corner(pl_lattice*);

/* One SmoothCornerAlong{YZ,ZX,XY} / SmoothCornerAt call (d3q15.h:1242-1303, d2q9.h:578-587) at GLOBAL coordinates, as
 * production/ncpump.cpp:159-170 issues them on interior corners.  A zero direction marks the axis the edge line runs
 * along (its coordinate is ignored): (0,dy,dz) = AlongYZ(j,k), (dx,0,dz) = AlongZX(k,i), (dx,dy,0) = AlongXY(i,j) and, for
 * D2Q9, the corner (i,j); three non-zero directions = the 3-D corner.  The whole line is processed, end sites included,
 * exactly as the reference helper does; a site outside this rank's block is a no-op. */
int pl_smooth_corner_at(pl_lattice*, int i, int j, int k, int dirx, int diry, int dirz);

/* ---- boundary conditions ------------------------------------------------------------------------
 * One pl_bc = one call of a reference "...AlongXFace/YFace/ZFace (XEdge/YEdge)" helper: the plane
 * `axis` = `coord` (GLOBAL coordinate, may be interior: production/ncpump.cpp:155-170), outward `dir` = -1/+1.
 * The host callables of the reference (bctype / value lambdas, evaluated with global coordinates,
 * navierstokes.h:155-157) are baked by the caller into per-plane arrays over the LOCAL plane sites in
 * natural order (lower axis fastest): X plane [j + ny*k], Y plane [i + nx*k], Z plane [i + nx*j].
 * mask_host: uint8 per plane site. BOUNCE/IBOUNCE: 0 none, 1 BARRIER, 2 MIRROR (d3q15.h:19-22); others: 0/1.
 * v0..v2_host: fp64 per plane site, meaning by type (unused = NULL). The arrays are copied at creation.
 * A plane that does not intersect this rank's block, or whose mask is all zero, yields an empty pl_bc
 * whose application is a no-op — exactly the reference's `if (0 <= i && i < nx)` guard (d3q15.h:987). */
#define PL_BC_BOUNCE 1        /* P::BoundaryConditionAlong*      d3q15.h:984-1110, d2q9.h:431-501 */
#define PL_BC_IBOUNCE 2       /* P::iBoundaryConditionAlong*     d3q15.h:1114-1239, d2q9.h:505-575 */
#define PL_BC_NS_SET_U 3      /* NS::BoundaryConditionSetUAlong*   navierstokes.h:92-257;  v0,v1,v2 = ux,uy,uz */
#define PL_BC_NS_SET_RHO 4    /* NS::BoundaryConditionSetRhoAlong* navierstokes.h:261-426; v0=rho, v1=_usbc, v2=_utbc */
#define PL_BC_AD_SET_T 5      /* AD::BoundaryConditionSetTAlong*   advection.h:99-238;  v0 = T; aux ux,uy,uz */
#define PL_BC_AD_SET_Q 6      /* AD::BoundaryConditionSetQAlong*   advection.h:242-524; v0 = qn; aux ux,uy,uz,diffusivity */
#define PL_BC_ANS_ISET_U 7    /* ANS::iBoundaryConditionSetUAlong* adjointnavierstokes.h:97-254; v0,v1,v2 = ux,uy,uz; eps */
#define PL_BC_ANS_ISET_RHO 8  /* ANS::iBoundaryConditionSetRhoAlong* adjointnavierstokes.h:258-392 */
#define PL_BC_AAD_ISET_T 9    /* AAD::iBoundaryConditionSetTAlong* adjointadvection.h:154-300; aux ux,uy,uz */
#define PL_BC_AAD_ISET_Q 10   /* AAD::iBoundaryConditionSetQAlong* adjointadvection.h:304-484; aux ux,uy,uz; eps */
#define PL_BC_AAD_ISET_RHO 11 /* AAD::iBoundaryConditionSetRhoAlong*Edge (D2Q9 only) adjointadvection.h:488-575 */
#define PL_BC_NSIN_SET_U 12   /* NSin::BoundaryConditionSetUAlong*Edge (D2Q9 only)   nsincompressible.h:46-94;   v0,v1 = ux,uy */
#define PL_BC_NSIN_SET_RHO 13 /* NSin::BoundaryConditionSetRhoAlong*Edge (D2Q9 only) nsincompressible.h:96-154;  v0 = rho, v1 = _usbc */

pl_bc* pl_bc_create(pl_lattice*, int type, int axis, int coord, int dir,
                    const uint8_t* mask_host, const double* v0_host, const double* v1_host, const double* v2_host);
int pl_bc_destroy(pl_bc*);
int pl_bc_is_empty(const pl_bc*);
/* Replace the per-site VALUES of a plane (same arrays as at creation, host pointers; the mask stays): a time-dependent inlet
 * profile keeps its handle — and every plan that holds it — while its numbers change.  Stream-ordered behind the passes queued. */
int pl_bc_update_values(pl_bc*, const double* v0, const double* v1, const double* v2);

/* Per-site fields some closures read at the boundary site (device pointers, nxyz doubles; unused = NULL):
 * the velocities saved by the collide of the same step (advection.h:1083-1090), the per-cell diffusivity
 * (advection.h:1114-1130) or its scalar overload (advection.h:1094-1110), eps (adjointadvection.h:1405). */
typedef struct pl_bc_aux {
    const double *rho, *ux, *uy, *uz, *tem, *diffusivity;
    double diffusivity_const;   /* used when diffusivity == NULL */
    double eps;
} pl_bc_aux;
/* Apply one plane closure to the lattice's current populations. `other` is the second lattice for
 * PL_BC_AAD_ISET_RHO (f first, g second as in adjointadvection.h:1425) and NULL otherwise. */
int pl_bc_apply(pl_lattice*, pl_lattice* other, const pl_bc*, const pl_bc_aux* aux);

/* ---- collides -----------------------------------------------------------------------------------
 * One entry point for every Macro*Collide* of the reference; `model` selects the function. */
#define PL_NS_COLLIDE 1                  /* NS::MacroCollide                         navierstokes_avx.h:93-201 */
#define PL_NS_BRINKMAN 2                 /* NS::MacroBrinkmanCollide                 navierstokes_avx.h:203-329 */
#define PL_AD_FORCE_CONV 3               /* AD::MacroCollideForceConvection          advection_avx.h:104-270 */
#define PL_AD_NAT_CONV 4                 /* AD::MacroCollideNaturalConvection        advection_avx.h:272-460 */
#define PL_AD_BRINKMAN_HEATEX 5          /* AD::MacroBrinkmanCollideHeatExchange     advection_avx.h:462-654 */
#define PL_AD_BRINKMAN_FORCE_CONV 6      /* AD::MacroBrinkmanCollideForceConvection  advection_avx.h:656-882 */
#define PL_AD_BRINKMAN_NAT_CONV 7        /* AD::MacroBrinkmanCollideNaturalConvection advection_avx.h:884-1116 */
#define PL_ANS_BRINKMAN 8                /* ANS::MacroBrinkmanCollide                adjointnavierstokes_avx.h:112-259 */
#define PL_AAD_HEATEX 9                  /* AAD::MacroBrinkmanCollideHeatExchange    adjointadvection_avx.h:324-528 */
#define PL_AAD_FORCE_CONV 10             /* AAD::MacroBrinkmanCollideForceConvection adjointadvection_avx.h:530-761 */
#define PL_AAD_NAT_CONV 11               /* AAD::MacroBrinkmanCollideNaturalConvection adjointadvection_avx.h:763-1005 */
#define PL_AAD_NAT_CONV_MASSFLOW 12      /* AAD::...NaturalConvectionMassFlow (D2Q9)  adjointadvection_avx.h:1007-1129 */
#define PL_NSIN_COLLIDE 13               /* NSin::MacroCollide (D2Q9, scalar templates only)          nsincompressible.h:158-181 */
#define PL_NSIN_BRINKMAN 14              /* NSin::MacroBrinkmanCollide (D2Q9)                         nsincompressible.h:183-210 */

typedef struct pl_collide_args {
    int model;
    int issave;                      /* _issave */
    double viscosity;                /* _viscosity */
    double diffusivity_const;        /* scalar _diffusivity overloads (models 3,4,5,9) */
    double gx, gy, gz, tem0;         /* buoyancy / reference temperature */
    const double *alpha;             /* Brinkman coefficient field */
    const double *diffusivity;       /* per-cell diffusivity field (models 6,7,10,11,12) */
    const double *beta;              /* heat-exchange coefficient field (models 5,9) */
    const double *dirx, *diry, *dirz;/* mass-flow direction fields (model 12) */
    /* forward macros: outputs of models 1-7, 13, 14 (written when issave), inputs of models 8-12 */
    double *rho, *ux, *uy, *uz, *tem, *qx, *qy, *qz;
    /* adjoint macros: outputs of models 8-12 (written when issave) */
    double *ip, *iux, *iuy, *iuz, *imx, *imy, *imz, *item, *iqx, *iqy, *iqz;
    /* optional snapshot of the thermal populations before relaxation (`_g` / `_ig`, advection_avx.h:1047-1052):
     * device buffer of nc*nxyz doubles (the size the drivers allocate, production/heatsink3D.cpp:59), stored SoA
     * [c][nxyz]; opaque to callers exactly as in the reference. */
    double *snapshot;
} pl_collide_args;
/* f = flow lattice (or the only lattice), g = thermal lattice (NULL for models 1,2,8). */
int pl_collide(pl_lattice* f, pl_lattice* g, const pl_collide_args*);

/* Export a device snapshot (SoA) into the reference's host layout ([pack][c][lane] for idx < 4*(nxyz/4),
 * [idx][c] for the tail; adjointadvection_avx.h:20-22) — used only by parity tests. */
int pl_snapshot_to_host(const pl_lattice*, const double* snapshot_dev, double* out_host);
/* ... and back: a snapshot held by the host in the reference's layout into the device layout.  The host-pointer surface uses
 * the pair to keep a caller's `_g` / `_ig` array in the REFERENCE layout whenever the host looks at it or has written it. */
int pl_snapshot_from_host(const pl_lattice*, const double* in_host, double* snapshot_dev);
/* the same for a lattice shape given by kind (PL_D2Q9 / PL_D3Q15) and site count alone; to_host: device layout -> reference layout */
int pl_snapshot_convert(int kind, long long nxyz, const double* in, double* out, int to_host);

/* InitialCondition of NS / AD / ANS / AAD (navierstokes.h:550-572, advection.h:1048-1070,
 * adjointnavierstokes.h:474-498, adjointadvection.h:1359-1381). family: 1=NS(rho,ux,uy,uz) 2=AD(tem,ux,uy,uz)
 * 3=ANS(ux,uy,uz,ip,iux,iuy,iuz) 4=AAD(ux,uy,uz,item,iqx,iqy,iqz) 5=NSin(rho,ux,uy,-; nsincompressible.h:212-223, D2Q9 only);
 * a[] holds the device arrays in that order. */
int pl_initial_condition(pl_lattice*, int family, const double* const* a, int na);

/* ---- fused time stepping ------------------------------------------------------------------------
 * A plan records one iteration of a driver time loop — collide; Stream/iStream; closures in call order;
 * SmoothCorner (test/cavityflow3D.cpp:48-58, production/heatsink3D.cpp:150-176, 194-216) — and advances it
 * with ONE fused stream+closure+collide pass per step instead of one pass per call. Results are identical
 * to issuing the calls one by one. args[2]/aux[2]: the two argument sets a driver alternates between by
 * std::swap of its array pointers (heatsink3D.cpp:178-183); pass the same pointer twice if it does not swap. */
pl_plan* pl_plan_create(pl_lattice* f, pl_lattice* g);
int pl_plan_destroy(pl_plan*);
int pl_plan_set_collide(pl_plan*, const pl_collide_args* even, const pl_collide_args* odd);
int pl_plan_set_stream(pl_plan*, int inverse);
int pl_plan_add_bc(pl_plan*, int on_g, const pl_bc*, const pl_bc_aux* even, const pl_bc_aux* odd);
int pl_plan_set_smooth_corner(pl_plan*, int on_f, int on_g);
/* SmoothCornerAt(i, j[, k], dx, dy[, dz]) of lattice f / g inside the loop body (d2q9.h:127-132, d3q15.h:212-220; the interior
 * corners of production/ncpump.cpp:159-162, 173-176).  The body is recorded in CALL ORDER (closures, SmoothCorner, SmoothCornerAt,
 * each numbered as it is added; pl_plan_set_smooth_corner counts where a flag is first raised); pl_plan_finalize accepts it when it
 * is equivalent to "all closures, then SmoothCorner, then the SmoothCornerAt points" — i.e. no closure called after a smoothing
 * touches a site that smoothing wrote or read — and answers PL_ERR_UNSUPPORTED otherwise. */
int pl_plan_add_smooth_corner_at(pl_plan*, int on_g, int i, int j, int k, int dx, int dy, int dz);
int pl_plan_finalize(pl_plan*);
/* Execute `ncollides` collides starting from the lattices' current phase (streamed or just collided);
 * every stream+closures+SmoothCorner between two collides is fused with the collide that follows.
 * end_streamed != 0 appends the trailing Stream+closures+SmoothCorner (loop ran to nt);
 * end_streamed == 0 stops right after the last collide (the drivers' convergence `break`). */
int pl_plan_advance(pl_plan*, int ncollides, int end_streamed);
/* The same for a caller that knows when it will look at the saved fields.  The reference stores rho, u, T, q (and the thermal
 * snapshot `_g`) at every site on every step because that is free on a CPU (production/heatsink3D.cpp:151,
 * advection_avx.h:1040-1052); nobody reads the interior values except Residual every `dt` steps (heatsink3D.cpp:152-160) and the code
 * after the loop (:227-246).  Only the LAST `save_last` collides of this call store what `issave` asks for at every site; the
 * earlier ones store it on the closure planes only, where the closures of the same step read the saved velocities
 * (advection.h:1083-1090).  With the two argument sets a driver alternates between (heatsink3D.cpp:178-183) save_last = 2 leaves
 * every array exactly as running all steps with stores would: both sets were last written by the final two collides.
 * save_last < 0: every collide stores everywhere (= pl_plan_advance). */
int pl_plan_advance_observed(pl_plan*, int ncollides, int end_streamed, int save_last);
/* Re-bind the ARRAY arguments of one argument set of a finalized plan; the loop body, its closures, masks and scalars stay.
 * For loops whose arrays change every step instead of alternating: the transient drivers keep one set of macroscopic arrays
 * and one snapshot per time step (production/heatsink3D_transient.cpp:50-57, 156-160, 196-200: rho[t], ux[t], ..., gi[t]).
 * `collide` (may be null = keep) replaces the collide arguments of set `parity` (same model and scalars as before);
 * `aux[k]` (k < naux, in pl_plan_add_bc order; aux == null = keep) replaces the field arrays of the k-th closure that was added
 * with fields.  Stream-ordered: takes effect for the passes queued after the call. */
int pl_plan_rebind(pl_plan*, int parity, const pl_collide_args* collide, const pl_bc_aux* aux, int naux);
/* 0/1: the argument set of the last collide if the lattices are in the just-collided phase, else of the next one */
int pl_plan_parity(const pl_plan*);
int pl_plan_set_parity(pl_plan*, int parity);
/* Measurement hook (bench.py "roofline"): when enabled, every launch of the fused interior kernel is bracketed by CUDA
 * events on the launching stream.  pl_plan_profile_read synchronises, returns the accumulated kernel milliseconds, the
 * number of launches and the total number of lattice sites those launches updated, and clears the accumulators. */
int pl_plan_profile(pl_plan*, int enable);
int pl_plan_profile_read(pl_plan*, double* total_ms, int* launches, long long* total_sites);
/* the same split by pass kind: [0] = passes that store on the closure planes only (pl_plan_advance_observed), [1] = passes in
 * which every site stores its macros / snapshot */
int pl_plan_profile_read2(pl_plan*, double* ms2, int* launches2, long long* sites2);

/* ---- reductions and sensitivities ---------------------------------------------------------------- */
/* Residual (src/utility/residual.h:8-50): sqrt(sum|u-up|^2 / sum|u|^2) over 1, 2 or 3 components. */
int pl_residual(const double* ux, const double* uy, const double* uz,
                const double* uxp, const double* uyp, const double* uzp, size_t n, double* out_host);
int pl_reduce_sum(const double* v, size_t n, double* out_host);
int pl_reduce_absmax(const double* v, size_t n, double* out_host);
/* Sum of a per-site field over the box [i0,i1) x [j0,j1) x [k0,k1) of GLOBAL coordinates, clipped to this rank's block and summed
 * over the ranks: the objective of the heatsink drivers (mean temperature of the heat patch, production/heatsink3D.cpp:227-240,
 * with its MPI_Allreduce) without bringing the whole field to the host. */
int pl_reduce_box_sum(const pl_lattice*, const double* v, int i0, int i1, int j0, int j1, int k0, int k1, double* out_host);
/* The design map of the heatsink drivers on the device (production/heatsink3D.cpp:114-119): from the filtered design ss the
 * diffusivity, the Brinkman coefficient and their derivatives, alpha0 = alphamax/(ly - 1); the reference's operation order. */
int pl_design_map(const double* ss, size_t n, double diff_fluid, double diff_solid, double qg, double alpha0, double qf,
                  double* diffusivity, double* alpha, double* dkds, double* dads);
/* Every rank's block of a per-site field assembled into the field of the global domain (lx*ly*lz doubles, host) on every rank:
 * what the reference's VTK writers gather with MPI_Isend/Irecv (src/utility/vtkxmlexport.h:172-214). */
int pl_comm_gather_field(const pl_lattice*, const double* v_dev, double* out_host_global);
/* Normalize (src/utility/normalize.h:8-24) */
int pl_normalize(double* v, size_t n);

#define PL_SENS_ANS_BRINKMAN 1            /* ANS::SensitivityBrinkman              adjointnavierstokes_avx.h:262-293 */
#define PL_SENS_AAD_HEATEX 2              /* AAD::SensitivityHeatExchange          adjointadvection_avx.h:1257-1300 */
#define PL_SENS_AAD_BRINKMAN_DIFF 3       /* AAD::SensitivityBrinkmanDiffusivity   adjointadvection_avx.h:1302-1401 */
typedef struct pl_sens_args {
    int kind;
    double* dfds;
    const double *ux, *uy, *uz, *imx, *imy, *imz, *dads, *tem, *item, *iqx, *iqy, *iqz;
    const double *gsnap, *igsnap;   /* device snapshots, SoA */
    const double *diffusivity, *dkds, *dbds;
} pl_sens_args;
int pl_sensitivity(pl_lattice*, const pl_sens_args*);
/* Heat-source boundary term of AAD::SensitivityTemperatureAtHeatSource (adjointadvection_avx.h:16-185) on one plane.
 * `plane` is a pl_bc created with type PL_BC_AD_SET_Q for that plane: mask = the _bctype predicate, v0 = the _qnbc values
 * (baked once and reusable every optimisation iteration). The volume term is pl_sensitivity(PL_SENS_AAD_BRINKMAN_DIFF). */
int pl_sensitivity_heat_source(pl_lattice*, const pl_bc* plane, double* dfds, const double* ux, const double* uy, const double* uz,
                               const double* igsnap, const double* diffusivity, const double* dkds);

/* ---- filters (src/utility/densityfilter.h:389-497, heavisidefilter.h:459-563, 641-857) ---------------------------------
 * The weight callable of the reference (`_weight(i1,j1,k1,i2,j2,k2)`, global coordinates) is baked once by the caller:
 * weights_host[o*nxyz + idx] for the site idx and its neighbour at offset o = ((di+nR)*(2nR+1) + (dj+nR))*(2nR+1) + (dk+nR)
 * (the reference's loop order: i2 outermost, k2 innermost), 0 for neighbours farther than R or outside the domain.
 * mode 0: DensityFilter::GetFilteredValue(v); 1: HeavisideFilter::GetFilteredVariable(s = v, beta);
 * 2: HeavisideFilter::GetFilteredSensitivity(s = v, dfdrho, beta).  Device pointers, nxyz doubles each.
 * On a block-decomposed lattice (NCCL communicator) the neighbours may lie in other ranks' blocks: the field is assembled
 * over the ranks first (one all-reduce of lx*ly*lz doubles per pass, instead of the reference's 26-neighbour nR-wide halo,
 * heavisidefilter.h:291-400); the weights then cover neighbours anywhere in the GLOBAL domain. */
typedef struct pl_filter pl_filter;
pl_filter* pl_filter_create(pl_lattice*, int nR, const double* weights_host);
/* The same from weight PATTERNS: patterns[p*K + o] (K = (2nR+1)^3) for the sites with pattern_of_site[idx] == p.  The weights of the
 * reference's drivers depend on the offset and on which side of the design box the two sites lie (production/heatsink3D.cpp:87-93) —
 * a few hundred distinct patterns whatever the lattice size; the drop-in headers bake the callable straight into this form (no
 * dense table is ever built), pl_filter_create folds its dense input into it.  pl_filter_patterns: how many patterns a filter holds. */
pl_filter* pl_filter_create_patterns(pl_lattice*, int nR, const double* patterns, int npatterns, const int* pattern_of_site);
int pl_filter_patterns(const pl_filter*);
int pl_filter_destroy(pl_filter*);
int pl_filter_apply(pl_filter*, int mode, double beta, const double* v, const double* dfdrho, double* out);

/* ==== host-pointer surface (csrc/panslbm_host.cpp) ====================================================================
 * What the drop-in C++ headers (panslbm2_b200/src/particle, src/equation, src/utility) bind.  Same operations as above, but
 * every array argument is a HOST pointer owned by the caller exactly as in the reference (`new double[nxyz]`,
 * production/heatsink3D.cpp:50-59): the runtime keeps a device mirror per array, moves data only when the other side
 * actually touched it (page protection on the host copy), and fuses the per-call sequence of a time loop
 * (collide; Stream; closures; SmoothCorner) into one pass per step once it has seen the loop body twice.
 * Results are identical to the device-pointer calls issued one by one. */
const char* plh_last_error(void);
/* Host memory with a device mirror; the headers route `operator new[]` / `delete[]` of large blocks here. */
void* plh_alloc(size_t bytes);
void plh_free(void* p);
int plh_owns(const void* p);
int plh_owns_range(const void* p);   /* p lies anywhere inside a mirrored block (arrays and population views) */
/* pl_bc_update_values behind the passes the fusion engine still holds back */
int plh_bc_update_values(pl_bc*, const double* v0, const double* v1, const double* v2);
/* The public `T *f0, *f` members of D2Q9/D3Q15 (d3q15.h:225): host views in the reference layout, kept coherent with
 * the device populations on demand (test/d2q9.cpp, test/d3q15.cpp read and write them directly). */
int plh_lattice_attach_views(pl_lattice*, double** f0, double** f);
int plh_lattice_detach(pl_lattice*);    /* before pl_lattice_destroy */
int plh_collide(pl_lattice* f, pl_lattice* g, const pl_collide_args* host_args);
int plh_stream(pl_lattice*, int inverse);
int plh_smooth_corner(pl_lattice*);
int plh_smooth_corner_at(pl_lattice*, int i, int j, int k, int dirx, int diry, int dirz);
int plh_bc(pl_lattice*, pl_lattice* other, const pl_bc*, const pl_bc_aux* host_aux);
int plh_initial_condition(pl_lattice*, int family, const double* const* host_arrays, int na);
int plh_residual(const double* ux, const double* uy, const double* uz, const double* uxp, const double* uyp, const double* uzp, size_t n, double* out);
int plh_normalize(double* v, size_t n);
int plh_sensitivity(pl_lattice*, const pl_sens_args* host_args);
int plh_sensitivity_heat_source(pl_lattice*, const pl_bc* plane, double* dfds, const double* ux, const double* uy, const double* uz,
                                const double* igsnap, const double* diffusivity, const double* dkds);
int plh_filter_apply(pl_filter*, int mode, double beta, const double* v_host, const double* dfdrho_host, double* out_host, size_t n);
/* Before handing a mirrored host array to code that is not this program's own loads and stores — a system call (fwrite of a whole
 * field), another library, a device-pointer entry point of this one (pl_comm_*, pl_array_upload): make [p, p + bytes) current and
 * readable on the host (for_write: writable, the host copy becomes the only current one).  The page-fault path serves ordinary
 * accesses transparently; those callers it cannot serve (EFAULT, or a fault inside the CUDA runtime, which aborts with a message). */
int plh_host_acquire(const void* p, size_t bytes, int for_write);
/* Execute whatever is still checked off and wait for the device. */
int plh_sync(void);
/* out[0..7] = fused steps, calls executed one by one, uploads, downloads, page faults served, plans built, settles, stagings */
int plh_stats(uint64_t* out8);
/* The device mirrors of the caller's arrays double as the HBM-resident store of the per-step states the transient drivers keep
 * (production/heatsink3D_transient.cpp:50-57: rho[t] ... gi[t], nt of each).  Under the budget PANSLBM_B200_DEVICE_BUDGET_MB
 * (unset: when device memory runs out) mirrors not used in the current or previous loop iteration are spilled to their host
 * copies and restored on demand.  out[0..3] = mirrors spilled, mirrors restored, device bytes held now, peak device bytes. */
int plh_store_stats(uint64_t* out4);

#ifdef __cplusplus
}
#endif
#endif
