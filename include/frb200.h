/* frb200.h -- C ABI of libfrb200.so: a Blackwell (sm_100a) implementation of the
 * FluxReconstruction.jl semi-discrete residual and explicit time step.
 *
 * The reference has NO foreign-function boundary on this path: it is pure Julia behind
 * the SciML in-place right-hand side  f!(du, u, p, t)  (src/Equation/eq_euler.jl:29,
 * src/Equation/eq_advection.jl:55, example/euler2d_wave.jl:35, example/bgk_wave.jl:69)
 * driven by OrdinaryDiffEq init/step!/solve.  Each entry point below names the
 * reference interface it stands in for; INTEGRATION.md shows the Julia `ccall` shim.
 *
 * Conventions
 *  - every function returns an int32 status: 0 = ok, negative = error
 *    (frb_last_error() holds the message; nothing throws, nothing calls exit);
 *  - all arrays are Float64 in the reference's own memory layout: Julia column-major,
 *    first index fastest, ghost cells included where the reference's array has them
 *    (OffsetArray 0:nx+1 is passed as the plain parent array);
 *  - host pointers are borrowed for the duration of the call only;
 *  - operator arrays (ll, lr, lpdm, dgl, dgr) are inputs: FRPSpace1D/FRPSpace2D
 *    (src/struct.jl:40-88,130-245) stay the single source of truth.  lpdm is the
 *    reference's `ps.dl`, nsp x nsp column-major: element [m,k] at lpdm[m + nsp*k];
 *  - there is no CPU fallback: without a CUDA device every compute call fails with
 *    FRB_ERR_CUDA.
 */
#ifndef FRB200_H
#define FRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRB_OK 0
#define FRB_ERR_ARG (-1)     /* bad argument / unsupported size */
#define FRB_ERR_CUDA (-2)    /* CUDA runtime or driver error, or no device */
#define FRB_ERR_STATE (-3)   /* call not valid for this problem kind / state */
#define FRB_ERR_NUMERIC (-4) /* limiter parameter outside (0,1] (reference: @assert) */
#define FRB_ERR_PEER (-5)    /* halo peer mapping / wait timeout */

#define FRB_MAX_NSP 8

typedef struct frb_ctx_s *frb_ctx_t;
typedef struct frb_prob_s *frb_prob_t;

/* eq_euler.jl:68-70 / eq_advection.jl:72-74: bc::Symbol -> :dirichlet | :period */
enum { FRB_BC_DIRICHLET = 0, FRB_BC_PERIOD = 1 };
/* OrdinaryDiffEq algorithms used by the reference at fixed dt:
 * Euler() (ns_cavity.jl:380), Midpoint() (euler2d_wave.jl:123); SSPRK3 = Shu-Osher */
enum { FRB_SCHEME_EULER = 0, FRB_SCHEME_MIDPOINT = 1, FRB_SCHEME_SSPRK3 = 2 };
/* per-step ghost fill of the user loop: euler2d_wave.jl:127-132 (x wave),
 * :159-164 (y wave), shock-vortex.jl:324-326 (copy) */
enum { FRB_GHOST_NONE = -1, FRB_GHOST_WAVE_X = 0, FRB_GHOST_WAVE_Y = 1, FRB_GHOST_COPY = 2,
       /* plain periodic copies in both directions: dev/parallelogram.jl:201-205 */
       FRB_GHOST_PERIODIC = 3,
       /* dev/cylinder2.jl:176-187: theta ghost rows = point-reversed mirror images with flipped y momentum,
        * outer radial column copied from its neighbour over the first half of the rows */
       FRB_GHOST_CYLINDER = 4 };
/* eq_advection.jl (seam epsilon 1e-6) vs example/advection_lowlevel.jl (1e-8) */
enum { FRB_ADV_PACKAGED = 0, FRB_ADV_LOWLEVEL = 1 };
/* kernel selection (frb_set_kernel).  2-D Euler: AUTO = the row-chunk streaming kernel for frb_step and the
 * reference-image marching kernel for f! with host buffers (deg 2-3), the generic per-element kernel otherwise */
enum { FRB_KERNEL_AUTO = 0, FRB_KERNEL_GENERIC = 1, FRB_KERNEL_MARCH = 2, FRB_KERNEL_RC = 3,
       /* bgk1d problems only: the one-pass register-tile kernel instead of the two launches (same results) */
       FRB_KERNEL_BGK_ONE_PASS = 4,
       /* curvilinear euler2d problems (frb_euler2d_curv_create): AUTO = one launch per stage, every block evaluates
        * the common fluxes of its own faces; GENERIC = face kernel + element kernel with the common fluxes through
        * device memory (also what the vertex metric runs); CURV_MARCH = one launch that marches over the rows of a
        * 30-element strip (stored metric, arrays below 2^32 elements).  Same results. */
       FRB_KERNEL_CURV_MARCH = 5 };

/* common (Riemann) flux of the Euler problems.  HLL is what the reference calls (flux_hll!,
 * eq_euler.jl:53, euler2d_wave.jl:73,80) and what the marching kernels implement; LF (Rusanov)
 * and ROE (Harten entropy fix below 0.1 a~) are the north star's extras: no reference
 * implementation exists, DESIGN.md section 2 specifies them; served by the generic kernels. */
enum { FRB_FLUX_HLL = 0, FRB_FLUX_LF = 1, FRB_FLUX_ROE = 2 };

/* the constant operator arrays of one FR space: ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr
 * (struct.jl:49-51,63-64,182,193), plus ps.dll/ps.dlr (struct.jl:55-61, may be NULL
 * unless the equation needs interface slopes) */
typedef struct {
  int32_t deg;
  const double *ll, *lr, *lpdm, *dgl, *dgr, *dll, *dlr;
} frb_operators;

/* ---- context ------------------------------------------------------------------ */
/* one context = one CUDA device + one stream; replaces nothing in the reference
 * (Julia owns no device).  device < 0 selects the current device. */
int32_t frb_ctx_create(int32_t device, frb_ctx_t *out);
int32_t frb_ctx_destroy(frb_ctx_t ctx);
/* message of the last failing call on this thread (ctx may be NULL) */
const char *frb_last_error(frb_ctx_t ctx);
/* library/ABI version, SM count and device name for logs */
int32_t frb_device_info(frb_ctx_t ctx, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor,
                        char *name, int32_t name_len);

/* ---- problems: one per packaged/example RHS of the reference ------------------- */
/* FRAdvectionProblem(u, tspan, ps, a, bc)  src/Equation/eq_advection.jl:1-26 and the
 * rhs! of example/advection_lowlevel.jl:4-47.  State u[ncell, nsp]; J[ncell]. */
int32_t frb_advection1d_create(frb_ctx_t ctx, int32_t ncell, const frb_operators *ops,
                               const double *J, double a, int32_t bc, int32_t variant,
                               frb_prob_t *out);
/* FREulerProblem(u, tspan, ps, gamma, bc)  src/Equation/eq_euler.jl:1-27,
 * RHS frode_euler! :29-98.  State u[ncell, nsp, 3]; J[ncell]. */
int32_t frb_euler1d_create(frb_ctx_t ctx, int32_t ncell, const frb_operators *ops,
                           const double *J, double gamma, int32_t bc, frb_prob_t *out);
/* dudt! of example/euler2d_wave.jl:35-107 (== shock-vortex.jl:26-118) on the uniform
 * rectangular FRPSpace2D with one ghost ring.  State u[nx+2, ny+2, nsp, nsp, 4];
 * Jx = dx/2, Jy = dy/2 are the diagonal of ps.J[i,j][k,l] (geo_jacobi.jl:77-108). */
int32_t frb_euler2d_create(frb_ctx_t ctx, int32_t nx, int32_t ny, const frb_operators *ops,
                           double Jx, double Jy, double gamma, frb_prob_t *out);
/* dudt! on a curvilinear structured quadrilateral FRPSpace2D(base, deg) (src/struct.jl:130-227): the
 * scratch scripts dev/parallelogram.jl:80-165 and dev/cylinder2.jl:52-164.  State as in
 * frb_euler2d_create, u[nx+2, ny+2, nsp, nsp, 4].  Metric and normals, Julia column-major:
 *   iJ [nx+2, ny+2, nsp, nsp, 2, 2] = ps.iJ[i,j][k,l][a,b] flattened (struct.jl:137-142; ghost entries unused)
 *   n1 [nx+1, ny, 2], n2 [nx, ny+1, 2] = the scripts' unit face normals (parallelogram.jl:176-186)
 *   fpc[nx, ny, nsp, 4] or NULL: correction factors at the flux points,
 *       (inv(Ji[i,j][4,l]) n1[i,j])[1], (inv(Ji[i,j][2,l]) n1[i+1,j])[1],
 *       (inv(Ji[i,j][1,k]) n2[i,j])[2], (inv(Ji[i,j][3,k]) n2[i,j+1])[2]      (cylinder2.jl:155-158);
 *       NULL: factors from the solution-point inverse Jacobian, (iJ[i,j][k,l] n)[c] (parallelogram.jl:145-148)
 * flags: FRB_CURV_FY_ROW_INDEX reproduces the scripts' fy_interaction[i,j,l,m] (row index where the
 * rectangular scripts use the flux-point index k, euler2d_wave.jl:100-103); FRB_CURV_WALL_XLO makes x face 1
 * the mirror wall of cylinder2.jl:100-120 (ghost column 0 is then never read).  deg 1..3, single GPU; every
 * frb_* call of an euler2d problem applies (f!, step, tableau, limiter, filter, ghost fill, frb_set_flux:
 * the common flux acts in the face frame of the unit normal). */
#define FRB_CURV_FY_ROW_INDEX 1
#define FRB_CURV_WALL_XLO 2
int32_t frb_euler2d_curv_create(frb_ctx_t ctx, int32_t nx, int32_t ny, const frb_operators *ops,
                                const double *iJ, const double *n1, const double *n2, const double *fpc,
                                int32_t flags, double gamma, frb_prob_t *out);
/* Optional: hand the library what ps.iJ is computed from -- vertices[nx+2, ny+2, 4, 2] = ps.base.vertices and
 * r[nsp] = ps.xpl (struct.jl:133-142) -- and the kernels evaluate iJ[i,j][k,l] = inv(rs_jacobi(r_k, r_l,
 * vertices[i,j])) (geo_jacobi.jl:77-88) on the fly: 8 doubles per element instead of 4 per solution point
 * cross HBM.  Equal to the stored metric up to rounding.  vertices == NULL returns to the stored copy. */
int32_t frb_euler2d_curv_set_vertices(frb_prob_t prob, const double *vertices, const double *r);
/* mol! of example/bgk_wave.jl:69-129 with its periodic e2f/f2e tables (:42-67).
 * State u[ncell, nu, nsp]; dx[ncell]; velo, weights [nu] (VSpace1D). */
int32_t frb_bgk1d_create(frb_ctx_t ctx, int32_t ncell, int32_t nu, const frb_operators *ops,
                         const double *dx, const double *velo, const double *weights,
                         double tau, frb_prob_t *out);
/* relaxation model of a bgk1d problem.  FRB_BGK_WAVE: Maxwellian from the three moments, conserve_prim(w, 3.0)
 * (bgk_wave.jl:77-81).  FRB_BGK_KINETIC_ADVECTION: the mol! of example/advection_kinetic.jl:73-128 -- the same
 * residual with rho = sum(u .* weights), prim = [rho, a, 1.0] (:80-88); tau is the one given at creation
 * (the script: 2.0 * 0.001, :89). */
enum { FRB_BGK_WAVE = 0, FRB_BGK_KINETIC_ADVECTION = 1 };
int32_t frb_bgk1d_set_model(frb_prob_t prob, int32_t model, double a);
/* dudt! + boundary! of example/ns_cavity.jl:147-344 (gas-kinetic flux, :49-145).
 * State u[4, nsp, nsp, ny+2, nx+2] (variable fastest).  ops->dll/dlr required.
 * gas = (K, gamma, mu_ref, omega); dt enters the time-averaged interface flux; lid_u is
 * the wall speed pb[2] of :337; lambda_wall the wall 1/T of boundary!(u,p,1.0). */
int32_t frb_ns2d_create(frb_ctx_t ctx, int32_t nx, int32_t ny, const frb_operators *ops,
                        double Jx, double Jy, double inK, double gamma, double mu_ref,
                        double omega, double dt, double lid_u, double lambda_wall,
                        frb_prob_t *out);
/* 2-D Euler on triangles: dudt! of dev/sod.jl:31-123 with the tuple it receives,
 * p = (ps.cellType, ps.J, ps.lf, ps.cellNormals, ps.fpn, ps.dl, ps.phi, gamma) of a TriFRPSpace
 * (struct.jl:305-352).  Julia layouts: cell_type[ncell] (0 interior, 1 frozen boundary, 2 mirror wall);
 * J[ncell,2,2] (geo_jacobi.jl:32-43, flattened from the vector of matrices); normals[ncell,3,2];
 * fpn[3, ncell, 3, deg+1] (the tuples (cell, face, point), 1-based, <= 0: no neighbour);
 * lf[3,deg+1,np]; dl[np,np,2]; phi[3,deg+1,np]; state u[ncell, np, 4]; np = (deg+1)(deg+2)/2, deg 1..3. */
int32_t frb_tri_euler_create(frb_ctx_t ctx, int32_t ncell, int32_t deg, const int32_t *cell_type, const double *J,
                             const double *normals, const int32_t *fpn, const double *lf, const double *dl,
                             const double *phi, double gamma, frb_prob_t *out);
int32_t frb_prob_destroy(frb_prob_t prob);

/* number of Float64 in the state array / of interior degrees of freedom */
int64_t frb_state_len(frb_prob_t prob);
int64_t frb_interior_dofs(frb_prob_t prob);

/* ---- state movement (itg.u of the reference lives in host memory) --------------- */
int32_t frb_state_upload(frb_prob_t prob, const double *u_host);
int32_t frb_state_download(frb_prob_t prob, double *u_host);
/* raw device address of the resident state (for zero-copy interop / peer mapping) */
int32_t frb_state_device_ptr(frb_prob_t prob, double **dptr);

/* ---- f!(du, u, p, t) ----------------------------------------------------------- */
/* The SciML in-place RHS.  u_host == NULL evaluates the resident state; du_host ==
 * NULL leaves du on the device.  With host pointers the call uploads u, evaluates
 * and downloads du (H2D + D2H inside the call).  f! is PURE with respect to the
 * integrator, as in the reference (eq_euler.jl:29: du is the only output): a
 * caller-supplied u is evaluated from a scratch buffer and the resident state that
 * frb_step advances is left as the last upload / step put it.  (ns2d: boundary!
 * rewrites the ghost cells of the device copy only, never the host array.)
 * With FRB_KERNEL_RC selected -- and, under AUTO, whenever the resident state already
 * lives in the row-chunk layout -- the 2-D Euler residual is evaluated by the same
 * euler2d_rc_kernel that frb_step launches (its rhs_only form).  t is accepted for
 * signature parity and unused (all reference RHS are autonomous). */
int32_t frb_rhs(frb_prob_t prob, const double *u_host, double *du_host, double t);
/* As frb_rhs with host pointers, but streams the 2-D state through the device in
 * row slabs so that H2D, compute and D2H overlap (euler2d only). */
int32_t frb_rhs_pipelined(frb_prob_t prob, const double *u_host, double *du_host, int32_t nslab);
/* step!(itg) with itg.u on the host between steps (the user loop of euler2d_wave.jl:125-135: the user fills the
 * ghost cells / limits / filters the HOST array, then steps): u_out = one step of `scheme` from u_in, both host
 * arrays in the reference image (they may be the same array).  2-D Euler, HLL, deg 2-3, no device step hooks: the
 * state streams through the device in `nslab` row slabs -- upload of the next slab, the stages of the slabs that
 * have arrived and the download of finished slabs overlap (PCIe full duplex); the ghost cells are the caller's and
 * stay frozen through the step.  Every other case: upload + frb_step + download.  The result is also the resident
 * state afterwards. */
int32_t frb_step_host(frb_prob_t prob, const double *u_in_host, double *u_out_host, int32_t scheme, double dt,
                      int32_t nslab);

/* ---- step!(itg) ---------------------------------------------------------------- */
/* what the user loop does around step! (euler2d_wave.jl:125-135, shock-vortex.jl:296-303):
 * before every step, optionally positive_limiter (weights != NULL: nsp or nsp*nsp
 * quadrature weights, already normalised as the callers do with ps.wp ./ 2 or ./ 4)
 * and optionally a ghost fill.  Persisted on the problem. */
int32_t frb_set_step_hooks(frb_prob_t prob, int32_t ghost_mode, const double *limiter_weights);
/* nsteps fixed steps of the resident state, fused RHS + stage update kernels, no host
 * synchronisation between steps.  Synchronous at return. */
int32_t frb_step(frb_prob_t prob, int32_t scheme, double dt, int32_t nsteps);
/* the same loop for any explicit Runge-Kutta scheme given by its Butcher tableau -- the fixed-step
 * Tsit5() of example/advection_highlevel.jl:26 and example/euler1d_convergence.jl:133
 * (adaptive=false, dt=dt), RK4, Heun, ...:  k_i = L(u + dt * sum_{j<i} A[i*nstage + j] k_j),
 * u <- u + dt * sum_i b[i] k_i.  A is row-major nstage x nstage (strictly lower part read),
 * nstage <= FRB_RK_MAX_STAGES.  Same per-step hooks as frb_step; single-GPU problems only. */
#define FRB_RK_MAX_STAGES 8
int32_t frb_step_tableau(frb_prob_t prob, int32_t nstage, const double *A, const double *b, double dt,
                         int32_t nsteps);
/* the hooks as stand-alone calls on the resident state */
int32_t frb_ghost_fill(frb_prob_t prob, int32_t ghost_mode);
/* positive_limiter(u, gamma, weights, ll, lr) src/dissipation.jl:61-123,125-206 on every
 * interior cell; *nbad = cells whose parameter left (0,1] (reference: @assert) */
int32_t frb_limiter_positivity(frb_prob_t prob, const double *weights, int32_t *nbad);
/* Shock sensor + modal filter on every element of the resident state (Euler problems), the pass
 * the reference's shock cases run on the host between two step! calls
 * (example/euler_highlevel.jl:37-52, example/shock-vortex.jl:308-321):
 *   u_hat = iV * u[element, density];  su = u_hat[end]^2 / (sum(u_hat.^2) + eps)
 *   if shock_detector(log10(su), deg, S0, kappa)  (src/dissipation.jl:13-23):
 *       every variable  u <- F * u,   F = V * diag(filter) * iV   (KitBase modal_filter! on the modes)
 * iV, F: np x np column-major (np = deg+1 in 1-D, (deg+1)^2 in 2-D, Julia's [:] order of the
 * element block); include_ghosts != 0 runs over the ghost cells too (2-D script).
 * *nfiltered = elements the sensor flagged. */
int32_t frb_filter_modal(frb_prob_t prob, const double *iV, const double *F, int32_t np, double eps,
                         double S0, double kappa, int32_t include_ghosts, int32_t *nfiltered);
/* the same pass as a hook of frb_step: when = 1 before every step, 2 after every step, 0 off */
int32_t frb_set_filter_hook(frb_prob_t prob, int32_t when, const double *iV, const double *F, int32_t np,
                            double eps, double S0, double kappa, int32_t include_ghosts);

/* ---- measurement (CUDA events on the library's own stream) ---------------------- */
/* Launch `iters` back-to-back RK stages of kind `stage_kind` (0: u' = u + dt L(u), 16 B/DOF;
 * 1: u' = a u_n + b u + c dt L(u), 24 B/DOF) on scratch buffers and return the mean
 * device time per launch in milliseconds.  The state is not modified. */
int32_t frb_time_stage(frb_prob_t prob, int32_t stage_kind, int32_t iters, float *ms_per_launch);
/* device time of the last frb_step / frb_rhs call in milliseconds (events around the
 * kernels only, excluding host<->device copies) and kernels launched by it */
int32_t frb_last_timing(frb_prob_t prob, float *ms, int64_t *kernel_launches);
int32_t frb_set_kernel(frb_prob_t prob, int32_t kernel_kind);
/* FRB_FLUX_*: Euler problems only (1-D and 2-D).  2-D, deg 2-3: every flux has its own instantiation of the
 * row-chunk stage kernel (frb_step, and f! through it); other degrees and FRB_KERNEL_GENERIC run the generic kernels */
int32_t frb_set_flux(frb_prob_t prob, int32_t flux_kind);
/* per-launch timing of the fused stage kernels inside frb_step / frb_rhs: when enabled,
 * every stage launch is bracketed by CUDA events on the library stream; after the call
 * returns, frb_stage_timing reports the summed device time of those launches and their
 * count (the dominant-kernel figure of the roofline report). */
int32_t frb_set_profiling(frb_prob_t prob, int32_t enabled);
int32_t frb_stage_timing(frb_prob_t prob, float *ms_total, int64_t *stage_launches);

/* ---- pinned host memory (for the host-buffer paths; replaces nothing in the reference) */
int32_t frb_host_alloc(int64_t bytes, void **ptr);
int32_t frb_host_free(void *ptr);

/* ---- multi-GPU: element-slab partition along the slowest cell index ------------- */
/* One process per GPU.  A rank's problem is created on its slab (2-D: ny_local rows,
 * 1-D: ncell_local cells) and told its neighbours.  Halo rows are the neighbours'
 * boundary rows of the *current* stage; they are written straight into the
 * neighbour's memory by the stage kernel over NVLink (peer mapping via CUDA IPC
 * handles the host exchanges out of band).  On the row-chunk path (2-D Euler, deg 2-3)
 * the whole exchange lives INSIDE the stage kernel: rows 1 / ny are computed first and
 * stored into a ring of halo rows on the neighbour, the last strip raises the
 * neighbour's mailbox (st.release.sys), and only the CTAs that need a halo row poll
 * for it (ld.acquire.sys) -- every other CTA runs at once, so the exchange overlaps
 * the interior rows and a stage stays ONE launch.  The other kernels push their rows
 * after the stage and hand epochs over with one-thread signal / wait kernels.
 * Curvilinear euler2d problems (frb_euler2d_curv_create) take the same row slabs: every rank creates its problem
 * from its own rows of iJ / n1 / n2 / fpc, ghost mode FRB_GHOST_PERIODIC; rows pushed after the stage.
 * ns2d problems (cfg5) are split the same way along the slowest index of their layout, i: a rank owns nx_local
 * columns, columns 0 / nx_local+1 of an interior slab boundary are halo columns that replace the wall ghosts
 * of boundary! there (no periodic seam); the same four calls apply.
 * The reference has no distributed path: this is new surface. */
#define FRB_IPC_HANDLE_BYTES 64
/* blob = 6 IPC handles (u, s1, s2, mailbox, row-chunk buffers, halo ring of the row-chunk stage kernel)
 * + int32 ny_local + int32 has_rc */
#define FRB_HALO_BLOB_BYTES (6 * FRB_IPC_HANDLE_BYTES + 8)
/* export this rank's blob; the host exchanges blobs out of band (torch.distributed, MPI, ...) */
int32_t frb_halo_export(frb_prob_t prob, unsigned char *blob_out);
/* map the blobs of the rank below (rank-1 mod nranks) and above (rank+1 mod nranks) and send
 * this rank's boundary rows of the resident state; call after every rank uploaded its slab.
 * The global y seam (rank 0 <-> rank nranks-1) is the frozen per-step ghost row of the ghost
 * mode set with frb_set_step_hooks; all other slab boundaries are exchanged every stage. */
int32_t frb_halo_connect(frb_prob_t prob, int32_t rank, int32_t nranks,
                         const unsigned char *blob_lo, const unsigned char *blob_hi);
/* resend the boundary rows after a new frb_state_upload (every rank calls it) */
int32_t frb_halo_sync(frb_prob_t prob);
int32_t frb_halo_disconnect(frb_prob_t prob);

#ifdef __cplusplus
}
#endif
#endif /* FRB200_H */
