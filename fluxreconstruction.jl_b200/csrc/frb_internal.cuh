// Internal declarations shared by the libfrb200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/frb200.h"

#define FRB_NSPMAX FRB_MAX_NSP

// Operator arrays of one FR space, passed by value as a kernel parameter so that
// every access is a uniform constant-bank load.  lpdm is row-major here:
// lpdm[m * FRB_NSPMAX + k] = ps.dl[m, k].
struct FrbOps {
  double ll[FRB_NSPMAX], lr[FRB_NSPMAX], dgl[FRB_NSPMAX], dgr[FRB_NSPMAX];
  double dll[FRB_NSPMAX], dlr[FRB_NSPMAX];
  double lpdm[FRB_NSPMAX * FRB_NSPMAX];
  // lpdm with the flux-trace part of the correction folded in:
  //   dmod[k][q] = lpdm[k][q] - dgl[k]*ll[q] - dgr[k]*lr[q]
  // so that  sum_q lpdm[k][q] f[q] + (fhatL - f.ll) dgl[k] + (fhatR - f.lr) dgr[k]
  //        = sum_q dmod[k][q] f[q] + fhatL dgl[k] + fhatR dgr[k]
  double dmod[FRB_NSPMAX * FRB_NSPMAX];
};

// One Runge-Kutta stage in the fused form  out = ca*ua + cb*u + cdt*L(u).
//   rhs_only : out = L(u)   (the f!(du,u,p,t) shape)
//   use_a    : ua is read (24 B/DOF stage) or skipped (16 B/DOF stage)
//   nested   : the combination is  ca*ua + cb*(u + cdt*L(u))  (Shu-Osher form, the order
//              OrdinaryDiffEq's SSPRK33 and the oracle evaluate it in)
struct FrbStage {
  double ca, cb, cdt;
  int use_a, rhs_only, nested;
};

enum FrbKind { K_ADV1D = 0, K_EULER1D = 1, K_EULER2D = 2, K_BGK1D = 3, K_NS2D = 4, K_TRI_EULER = 5 };

void frb_set_error(const std::string &msg);
int frb_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FRB_CUDA(call)                                                     \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return frb_cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

struct frb_ctx_s {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_in = nullptr, copy_out = nullptr;
  std::string name;
};

struct FrbHalo;

struct frb_prob_s {
  frb_ctx_t ctx = nullptr;
  FrbKind kind = K_ADV1D;
  int nsp = 0;
  FrbOps ops;
  // sizes
  int ncell = 0, nu = 0;  // 1-D
  int nx = 0, ny = 0;     // 2-D (interior)
  int64_t len = 0, dofs = 0;
  // physics
  double a = 0, gamma = 0, Jx = 0, Jy = 0, tau = 0;
  int bc = 0, variant = 0;
  int bgk_model = 0;  // FRB_BGK_* (frb_bgk1d_set_model); the advection speed of the kinetic model is `a`
  int flux = FRB_FLUX_HLL;  // common flux of the Euler problems (frb_set_flux)
  double gks_K = 0, gks_mu = 0, gks_omega = 0, gks_dt = 0, lid_u = 0, lambda_wall = 1.0;
  // device buffers
  double *u = nullptr;              // resident state  (u_n)
  double *s1 = nullptr, *s2 = nullptr;  // stage buffers
  double *du = nullptr;             // f! output (lazy)
  double *s3 = nullptr;             // fourth state buffer of the pipelined host step (frb_step_host, lazy)
  std::vector<double *> rk_k;       // stage derivatives of frb_step_tableau (lazy)
  // row-chunk mirror of the 2-D Euler state (frb_rc.cuh): the layout frb_step streams in.  u is
  // the reference image the ABI exposes; exactly one of the two (or both) is current.
  double *rc_base = nullptr;        // one allocation: ru | rs1 | rs2
  double *ru = nullptr, *rs1 = nullptr, *rs2 = nullptr;
  size_t rc_len = 0;                // doubles per RC buffer
  bool ref_valid = true, rc_valid = false;
  double *J = nullptr;              // 1-D per-cell Jacobian / bgk dx
  double *velo = nullptr, *weights = nullptr, *prim = nullptr;  // bgk
  // triangles: operators (lf | dl | phi), flux-point traces, connectivity
  double *tri_ops = nullptr, *tri_uf = nullptr, *tri_normals = nullptr;
  int *tri_type = nullptr, *tri_fpn = nullptr;
  double *ns_flux = nullptr;        // ns2d: common fluxes on the x | y faces (lazy)
  // curvilinear quadrilaterals (frb_euler2d_curv_create): metric planes, face normals, optional
  // flux-point correction factors; curv_iJ != nullptr marks the problem
  double *curv_iJ = nullptr, *curv_n1 = nullptr, *curv_n2 = nullptr, *curv_fpc = nullptr;
  double *curv_vert = nullptr;      // cell vertices (frb_euler2d_curv_set_vertices): metric evaluated on the fly
  double curv_r[FRB_NSPMAX] = {0};  // solution points
  double *curv_flux = nullptr;      // common fluxes on the x | y faces (lazy)
  int curv_flags = 0;
  double *lim_w = nullptr;          // limiter weights (device)
  // shock sensor + modal filter hook (frb_set_filter_hook): iV | F on the device, when = 0 off,
  // 1 before every step (euler_highlevel.jl:37-52), 2 after every step (shock-vortex.jl:308-321)
  double *filt = nullptr;
  int filt_np = 0, filt_when = 0, filt_ghosts = 0;
  double filt_eps = 0, filt_S0 = 0, filt_kappa = 0;
  int *flag = nullptr;              // device int for limiter nbad
  unsigned *loop_bar = nullptr;     // grid barrier of the one-launch 1-D time loop
  // hooks
  int ghost_mode = FRB_GHOST_NONE;
  bool limiter_on = false;
  int kernel_kind = FRB_KERNEL_AUTO;
  int row_lo = 0, row_hi = 0;  // sub-range of rows for the next 2-D stage launch (0 = all)
  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int64_t last_launches = 0;
  int64_t launches = 0;
  // per-stage profiling (event pairs around every stage launch)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;  // pool
  std::vector<cudaEvent_t> pipe_events;  // per-slab events of frb_rhs_pipelined (created once)
  size_t prof_used = 0;
  float stage_ms = 0.f;
  int64_t stage_count = 0;
  // TMA descriptors for the marching kernel (opaque 128-byte CUtensorMap images)
  void *tmaps = nullptr;  // host copy, keyed by device pointer
  FrbHalo *halo = nullptr;
  bool halo_pending = false;  // a signal was sent; wait for the neighbours before the next stage
  bool halo_pending_legacy = false;  // ... by a signal kernel (rows pushed into the arrays themselves, not the ring)
};

// ---- kernel launchers (each returns the number of kernels launched or <0) --------
int frb_launch_adv1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_euler1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_bgk1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_euler2d_generic(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_euler2d_curv(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_ghost_cylinder(frb_prob_t p, double *u);
int frb_launch_euler2d_march(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
bool frb_euler2d_march_supported(frb_prob_t p);
// row-chunk path (frb_euler2d_rc.cu, frb_rc.cu); every pointer is an RC buffer unless named ref
bool frb_euler2d_rc_supported(frb_prob_t p);
struct RcHalo;
int frb_launch_euler2d_rc(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st,
                          const RcHalo *halo = nullptr);
int frb_rc_from_ref(frb_prob_t p, const double *ref, double *rc);
int frb_rc_to_ref(frb_prob_t p, const double *rc, double *ref, bool interior_only = false);
int frb_rc_ghost_fill(frb_prob_t p, double *u, int mode);
int frb_rc_ghost_x(frb_prob_t p, double *u, int mode);
int frb_rc_ring_copy(frb_prob_t p, const double *src, double *dst, bool row0, bool rowN);
int frb_rc_ghost_x_ring(frb_prob_t p, double *u, double *d1, double *d2, int mode);
int frb_rc_row_push3(frb_prob_t p, const double *src, double *const *dst_lo, double *const *dst_hi, int nyl_lo,
                     int flip_var);
int frb_halo_push_seam_all(frb_prob_t p, const double *src, int flip_var);
int frb_rc_limiter2d(frb_prob_t p, double *u);
int frb_rc_row_push(frb_prob_t p, const double *src, double *dst_lo, double *dst_hi, int nyl_lo, int flip_var);
int frb_launch_tri_euler(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_ns2d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);
int frb_launch_ghost_fill2d(frb_prob_t p, double *u, int mode);
int frb_launch_ring_copy2d(frb_prob_t p, const double *src, double *dst, bool row0 = true, bool rowN = true);
int frb_launch_ghost_x2d(frb_prob_t p, double *u, int mode);
int frb_launch_ring_rows2d(frb_prob_t p, const double *src, double *d1, double *d2, double *d3, int ra, int rb);
int frb_launch_lincomb(frb_prob_t p, double *out, const double *u, int n, double *const *k, const double *c,
                       double dt);
// frb_halo.cu
bool frb_halo_active(frb_prob_t p);
void frb_halo_swap_roles(frb_prob_t p, int a, int b, bool rc = false);
void frb_halo_stage_targets(frb_prob_t p, const double *out, double **dst_lo, double **dst_hi, int *nyl_lo,
                            int *nyl_hi);
// in-kernel exchange of the row-chunk stage kernel (frb_rc.cuh: RcHalo)
bool frb_halo_rc_inkernel(frb_prob_t p);
bool frb_halo_rc_input_in_array(frb_prob_t p, const double *u);
int frb_halo_rc_stage(frb_prob_t p, const double *u, const double *out, RcHalo *h);
int frb_halo_rc_flush(frb_prob_t p, double *U);
void frb_halo_rc_reset(frb_prob_t p);
int frb_halo_push(frb_prob_t p, const double *src, int dst_role, bool seam, int flip_var);
int frb_halo_role(frb_prob_t p, const double *ptr);
int frb_halo_signal(frb_prob_t p);
int frb_halo_wait(frb_prob_t p);
int frb_halo_check_timeout(frb_prob_t p);
int frb_halo_rank(frb_prob_t p, int *nranks);
int frb_launch_modal_filter(frb_prob_t p, double *u, const double *iV_dev, const double *F_dev, double eps,
                            double S0, double kappa, bool include_ghosts, int *count);
int frb_rc_modal_filter(frb_prob_t p, double *u, const double *iV_dev, const double *F_dev, double eps, double S0,
                        double kappa, bool include_ghosts);
int frb_launch_limiter1d(frb_prob_t p, double *u);
int frb_launch_loop1d(frb_prob_t p, int scheme, double dt, int nsteps);
int frb_launch_limiter2d(frb_prob_t p, double *u);
int frb_launch_dirichlet_copy1d(frb_prob_t p, const double *src, double *dst);
void frb_march_release(frb_prob_t p);
