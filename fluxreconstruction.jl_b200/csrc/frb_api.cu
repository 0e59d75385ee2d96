// C ABI of libfrb200 (see include/frb200.h): contexts, problems, state movement,
// f!(du,u,p,t), step!, hooks and measurement.  No CPU fallback anywhere: every compute
// entry point launches sm_100a kernels or fails.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "frb_internal.cuh"
#include "frb_rc.cuh"

static thread_local std::string g_err;

void frb_set_error(const std::string &msg) { g_err = msg; }

int frb_cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file,
           line, what);
  g_err = buf;
  return FRB_ERR_CUDA;
}

extern "C" const char *frb_last_error(frb_ctx_t) { return g_err.c_str(); }

#define FRB_REQUIRE(cond, code, msg) \
  do {                               \
    if (!(cond)) {                   \
      frb_set_error(msg);            \
      return (code);                 \
    }                                \
  } while (0)

// ---- context ------------------------------------------------------------------------
extern "C" int32_t frb_ctx_create(int32_t device, frb_ctx_t *out) {
  FRB_REQUIRE(out, FRB_ERR_ARG, "frb_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    frb_set_error(std::string("no CUDA device available (libfrb200 has no CPU fallback): ") +
                  cudaGetErrorString(e));
    return FRB_ERR_CUDA;
  }
  if (device < 0) FRB_CUDA(cudaGetDevice(&device));
  FRB_REQUIRE(device < ndev, FRB_ERR_ARG, "frb_ctx_create: device index out of range");
  FRB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FRB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    frb_set_error(std::string("libfrb200 is built for sm_100a only; device is ") + prop.name);
    return FRB_ERR_CUDA;
  }
  frb_ctx_t c = new frb_ctx_s();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->name = prop.name;
  FRB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  FRB_CUDA(cudaStreamCreateWithFlags(&c->copy_in, cudaStreamNonBlocking));
  FRB_CUDA(cudaStreamCreateWithFlags(&c->copy_out, cudaStreamNonBlocking));
  *out = c;
  return FRB_OK;
}

extern "C" int32_t frb_ctx_destroy(frb_ctx_t c) {
  if (!c) return FRB_OK;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_in) cudaStreamDestroy(c->copy_in);
  if (c->copy_out) cudaStreamDestroy(c->copy_out);
  delete c;
  return FRB_OK;
}

extern "C" int32_t frb_device_info(frb_ctx_t c, int32_t *sm_count, int32_t *cc_major,
                                   int32_t *cc_minor, char *name, int32_t name_len) {
  FRB_REQUIRE(c, FRB_ERR_ARG, "frb_device_info: ctx is NULL");
  if (sm_count) *sm_count = c->sm_count;
  if (cc_major) *cc_major = c->cc_major;
  if (cc_minor) *cc_minor = c->cc_minor;
  if (name && name_len > 0) {
    strncpy(name, c->name.c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  return FRB_OK;
}

// ---- problems -----------------------------------------------------------------------
static int fill_ops(const frb_operators *o, FrbOps *dst, int *nsp_out, bool need_slopes) {
  FRB_REQUIRE(o && o->ll && o->lr && o->lpdm && o->dgl && o->dgr, FRB_ERR_ARG,
              "operators: ll, lr, lpdm, dgl, dgr must be non-NULL");
  const int nsp = o->deg + 1;
  FRB_REQUIRE(o->deg >= 1 && nsp <= FRB_NSPMAX, FRB_ERR_ARG, "operators: deg must be in 1..7");
  FRB_REQUIRE(!need_slopes || (o->dll && o->dlr), FRB_ERR_ARG, "operators: dll/dlr required");
  memset(dst, 0, sizeof(FrbOps));
  for (int q = 0; q < nsp; ++q) {
    dst->ll[q] = o->ll[q];
    dst->lr[q] = o->lr[q];
    dst->dgl[q] = o->dgl[q];
    dst->dgr[q] = o->dgr[q];
    if (o->dll) dst->dll[q] = o->dll[q];
    if (o->dlr) dst->dlr[q] = o->dlr[q];
    for (int k = 0; k < nsp; ++k) dst->lpdm[q * FRB_NSPMAX + k] = o->lpdm[q + nsp * k];  // [m,k] col-major
  }
  for (int k = 0; k < nsp; ++k)
    for (int q = 0; q < nsp; ++q)
      dst->dmod[k * FRB_NSPMAX + q] =
          dst->lpdm[k * FRB_NSPMAX + q] - dst->dgl[k] * dst->ll[q] - dst->dgr[k] * dst->lr[q];
  *nsp_out = nsp;
  return FRB_OK;
}

static int alloc_common(frb_prob_t p) {
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  FRB_CUDA(cudaMalloc(&p->u, sizeof(double) * p->len));
  FRB_CUDA(cudaMalloc(&p->s1, sizeof(double) * p->len));
  FRB_CUDA(cudaMalloc(&p->s2, sizeof(double) * p->len));
  FRB_CUDA(cudaMemsetAsync(p->u, 0, sizeof(double) * p->len, p->ctx->stream));
  FRB_CUDA(cudaMemsetAsync(p->s1, 0, sizeof(double) * p->len, p->ctx->stream));
  FRB_CUDA(cudaMemsetAsync(p->s2, 0, sizeof(double) * p->len, p->ctx->stream));
  FRB_CUDA(cudaMalloc(&p->flag, sizeof(int)));
  FRB_CUDA(cudaEventCreate(&p->ev0));
  FRB_CUDA(cudaEventCreate(&p->ev1));
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return FRB_OK;
}

// row-chunk mirror of the 2-D Euler state (the layout frb_step streams in, frb_rc.cuh); allocated
// with the problem so that the multi-GPU path can export it
static int alloc_rc(frb_prob_t p) {
  if (!frb_euler2d_rc_supported(p) || getenv("FRB_NO_RC")) return FRB_OK;
  if (p->ny + 2 > 65535) return FRB_OK;  // grid.z / grid.y of the rc_* kernels index rows: stay on the reference image
  const RcGeom g = rc_geom(p->nx, p->ny, p->nsp);
  p->rc_len = g.len;
  FRB_CUDA(cudaMalloc(&p->rc_base, sizeof(double) * 3 * g.len));
  FRB_CUDA(cudaMemset(p->rc_base, 0, sizeof(double) * 3 * g.len));
  p->ru = p->rc_base;
  p->rs1 = p->rc_base + g.len;
  p->rs2 = p->rc_base + 2 * g.len;
  return FRB_OK;
}

static int upload_vec(frb_prob_t p, double **dst, const double *src, size_t n) {
  FRB_CUDA(cudaMalloc(dst, sizeof(double) * n));
  FRB_CUDA(cudaMemcpy(*dst, src, sizeof(double) * n, cudaMemcpyHostToDevice));
  (void)p;
  return FRB_OK;
}

extern "C" int32_t frb_prob_destroy(frb_prob_t p) {
  if (!p) return FRB_OK;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  frb_halo_disconnect(p);
  frb_march_release(p);
  cudaFree(p->u); cudaFree(p->s1); cudaFree(p->s2); cudaFree(p->s3); cudaFree(p->du); cudaFree(p->rc_base);
  cudaFree(p->J); cudaFree(p->velo); cudaFree(p->weights); cudaFree(p->prim);
  cudaFree(p->lim_w); cudaFree(p->flag); cudaFree(p->loop_bar); cudaFree(p->filt); cudaFree(p->ns_flux);
  cudaFree(p->curv_iJ); cudaFree(p->curv_n1); cudaFree(p->curv_n2); cudaFree(p->curv_fpc); cudaFree(p->curv_flux); cudaFree(p->curv_vert);
  cudaFree(p->tri_ops); cudaFree(p->tri_uf); cudaFree(p->tri_normals); cudaFree(p->tri_type); cudaFree(p->tri_fpn);
  if (p->ev0) cudaEventDestroy(p->ev0);
  if (p->ev1) cudaEventDestroy(p->ev1);
  for (cudaEvent_t e : p->prof_events) cudaEventDestroy(e);
  for (cudaEvent_t e : p->pipe_events) cudaEventDestroy(e);
  for (double *k : p->rk_k) cudaFree(k);
  delete p;
  return FRB_OK;
}

#define FRB_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ < 0) {              \
      frb_prob_destroy(p);       \
      return rc__;               \
    }                            \
  } while (0)

extern "C" int32_t frb_advection1d_create(frb_ctx_t ctx, int32_t ncell, const frb_operators *ops,
                                          const double *J, double a, int32_t bc, int32_t variant,
                                          frb_prob_t *out) {
  FRB_REQUIRE(ctx && out && J, FRB_ERR_ARG, "frb_advection1d_create: NULL argument");
  FRB_REQUIRE(ncell >= 2, FRB_ERR_ARG, "frb_advection1d_create: ncell must be >= 2");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_ADV1D; p->ncell = ncell; p->a = a; p->bc = bc; p->variant = variant;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, false));
  p->len = (int64_t)ncell * p->nsp;
  p->dofs = p->len;
  FRB_TRY(alloc_common(p));
  FRB_TRY(upload_vec(p, &p->J, J, ncell));
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_euler1d_create(frb_ctx_t ctx, int32_t ncell, const frb_operators *ops,
                                      const double *J, double gamma, int32_t bc, frb_prob_t *out) {
  FRB_REQUIRE(ctx && out && J, FRB_ERR_ARG, "frb_euler1d_create: NULL argument");
  FRB_REQUIRE(ncell >= 2, FRB_ERR_ARG, "frb_euler1d_create: ncell must be >= 2");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_EULER1D; p->ncell = ncell; p->gamma = gamma; p->bc = bc;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, false));
  p->len = (int64_t)ncell * p->nsp * 3;
  p->dofs = p->len;
  FRB_TRY(alloc_common(p));
  FRB_TRY(upload_vec(p, &p->J, J, ncell));
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_euler2d_create(frb_ctx_t ctx, int32_t nx, int32_t ny,
                                      const frb_operators *ops, double Jx, double Jy, double gamma,
                                      frb_prob_t *out) {
  FRB_REQUIRE(ctx && out, FRB_ERR_ARG, "frb_euler2d_create: NULL argument");
  FRB_REQUIRE(nx >= 1 && ny >= 1, FRB_ERR_ARG, "frb_euler2d_create: nx, ny must be >= 1");
  FRB_REQUIRE(Jx > 0 && Jy > 0, FRB_ERR_ARG, "frb_euler2d_create: Jacobian must be positive");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_EULER2D; p->nx = nx; p->ny = ny; p->Jx = Jx; p->Jy = Jy; p->gamma = gamma;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, false));
  if (p->nsp > 6) {
    frb_set_error("frb_euler2d_create: deg must be in 1..5");
    frb_prob_destroy(p);
    return FRB_ERR_ARG;
  }
  p->len = (int64_t)(nx + 2) * (ny + 2) * p->nsp * p->nsp * 4;
  p->dofs = (int64_t)nx * ny * p->nsp * p->nsp * 4;
  FRB_TRY(alloc_common(p));
  FRB_TRY(alloc_rc(p));
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_euler2d_curv_create(frb_ctx_t ctx, int32_t nx, int32_t ny, const frb_operators *ops,
                                           const double *iJ, const double *n1, const double *n2,
                                           const double *fpc, int32_t flags, double gamma, frb_prob_t *out) {
  FRB_REQUIRE(ctx && out && iJ && n1 && n2, FRB_ERR_ARG, "frb_euler2d_curv_create: NULL argument");
  FRB_REQUIRE(nx >= 1 && ny >= 1, FRB_ERR_ARG, "frb_euler2d_curv_create: nx, ny must be >= 1");
  FRB_REQUIRE((flags & ~(FRB_CURV_FY_ROW_INDEX | FRB_CURV_WALL_XLO)) == 0, FRB_ERR_ARG,
              "frb_euler2d_curv_create: unknown flag");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_EULER2D; p->nx = nx; p->ny = ny; p->Jx = 1.0; p->Jy = 1.0; p->gamma = gamma;
  p->curv_flags = flags;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, false));
  if (p->nsp > 4) {
    frb_set_error("frb_euler2d_curv_create: deg must be in 1..3");
    frb_prob_destroy(p);
    return FRB_ERR_ARG;
  }
  const size_t npp = (size_t)p->nsp * p->nsp, ne = (size_t)(nx + 2) * (ny + 2);
  p->len = (int64_t)(ne * npp * 4);
  p->dofs = (int64_t)nx * ny * (int64_t)npp * 4;
  FRB_TRY(alloc_common(p));
  FRB_TRY(upload_vec(p, &p->curv_iJ, iJ, ne * npp * 4));
  FRB_TRY(upload_vec(p, &p->curv_n1, n1, (size_t)(nx + 1) * ny * 2));
  FRB_TRY(upload_vec(p, &p->curv_n2, n2, (size_t)nx * (ny + 1) * 2));
  if (fpc) FRB_TRY(upload_vec(p, &p->curv_fpc, fpc, (size_t)nx * ny * p->nsp * 4));
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_euler2d_curv_set_vertices(frb_prob_t p, const double *vertices, const double *r) {
  FRB_REQUIRE(p && p->curv_iJ, FRB_ERR_STATE, "frb_euler2d_curv_set_vertices: curvilinear euler2d problems only");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (!vertices) {  // back to the stored metric
    cudaFree(p->curv_vert);
    p->curv_vert = nullptr;
    return FRB_OK;
  }
  FRB_REQUIRE(r, FRB_ERR_ARG, "frb_euler2d_curv_set_vertices: r is NULL");
  const size_t n = (size_t)(p->nx + 2) * (p->ny + 2) * 8;
  if (!p->curv_vert) FRB_CUDA(cudaMalloc(&p->curv_vert, sizeof(double) * n));
  FRB_CUDA(cudaMemcpy(p->curv_vert, vertices, sizeof(double) * n, cudaMemcpyHostToDevice));
  for (int q = 0; q < p->nsp; ++q) p->curv_r[q] = r[q];
  return FRB_OK;
}

extern "C" int32_t frb_bgk1d_create(frb_ctx_t ctx, int32_t ncell, int32_t nu,
                                    const frb_operators *ops, const double *dx, const double *velo,
                                    const double *weights, double tau, frb_prob_t *out) {
  FRB_REQUIRE(ctx && out && dx && velo && weights, FRB_ERR_ARG, "frb_bgk1d_create: NULL argument");
  FRB_REQUIRE(ncell >= 2 && nu >= 1 && nu <= 65535, FRB_ERR_ARG, "frb_bgk1d_create: bad sizes");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_BGK1D; p->ncell = ncell; p->nu = nu; p->tau = tau;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, false));
  p->len = (int64_t)ncell * nu * p->nsp;
  p->dofs = p->len;
  FRB_TRY(alloc_common(p));
  {
    std::vector<double> inv_j(ncell);  // 1 / J = 2 / dx: the kernel multiplies (bgk_wave.jl:85-88 divides)
    for (int i = 0; i < ncell; ++i) inv_j[i] = 1.0 / (0.5 * dx[i]);
    FRB_TRY(upload_vec(p, &p->J, inv_j.data(), ncell));
  }
  FRB_TRY(upload_vec(p, &p->velo, velo, nu));
  FRB_TRY(upload_vec(p, &p->weights, weights, nu));
  if (cudaMalloc(&p->prim, sizeof(double) * (size_t)ncell * p->nsp * 3) != cudaSuccess) {
    frb_prob_destroy(p);
    frb_set_error("frb_bgk1d_create: cudaMalloc failed");
    return FRB_ERR_CUDA;
  }
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_bgk1d_set_model(frb_prob_t p, int32_t model, double a) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_bgk1d_set_model: prob is NULL");
  FRB_REQUIRE(p->kind == K_BGK1D, FRB_ERR_STATE, "frb_bgk1d_set_model: bgk1d problems only");
  FRB_REQUIRE(model == FRB_BGK_WAVE || model == FRB_BGK_KINETIC_ADVECTION, FRB_ERR_ARG,
              "frb_bgk1d_set_model: unknown model");
  p->bgk_model = model;
  p->a = a;
  return FRB_OK;
}

extern "C" int32_t frb_ns2d_create(frb_ctx_t ctx, int32_t nx, int32_t ny, const frb_operators *ops,
                                   double Jx, double Jy, double inK, double gamma, double mu_ref,
                                   double omega, double dt, double lid_u, double lambda_wall,
                                   frb_prob_t *out) {
  FRB_REQUIRE(ctx && out, FRB_ERR_ARG, "frb_ns2d_create: NULL argument");
  FRB_REQUIRE(nx >= 1 && ny >= 1 && Jx > 0 && Jy > 0 && dt > 0, FRB_ERR_ARG, "frb_ns2d_create: bad sizes");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_NS2D; p->nx = nx; p->ny = ny; p->Jx = Jx; p->Jy = Jy; p->gamma = gamma;
  p->gks_K = inK; p->gks_mu = mu_ref; p->gks_omega = omega; p->gks_dt = dt; p->lid_u = lid_u;
  p->lambda_wall = lambda_wall;
  FRB_TRY(fill_ops(ops, &p->ops, &p->nsp, true));
  if (p->nsp > 4) {
    frb_set_error("frb_ns2d_create: deg must be in 1..3");
    frb_prob_destroy(p);
    return FRB_ERR_ARG;
  }
  p->len = (int64_t)(nx + 2) * (ny + 2) * p->nsp * p->nsp * 4;
  p->dofs = (int64_t)nx * ny * p->nsp * p->nsp * 4;
  FRB_TRY(alloc_common(p));
  *out = p;
  return FRB_OK;
}

extern "C" int32_t frb_tri_euler_create(frb_ctx_t ctx, int32_t ncell, int32_t deg, const int32_t *cell_type,
                                        const double *J, const double *normals, const int32_t *fpn,
                                        const double *lf, const double *dl, const double *phi, double gamma,
                                        frb_prob_t *out) {
  FRB_REQUIRE(ctx && out && cell_type && J && normals && fpn && lf && dl && phi, FRB_ERR_ARG,
              "frb_tri_euler_create: NULL argument");
  FRB_REQUIRE(ncell >= 1 && deg >= 1 && deg <= 3, FRB_ERR_ARG, "frb_tri_euler_create: ncell >= 1, deg in 1..3");
  frb_prob_t p = new frb_prob_s();
  p->ctx = ctx; p->kind = K_TRI_EULER; p->ncell = ncell; p->nsp = deg + 1; p->gamma = gamma;
  const int np = (deg + 1) * (deg + 2) / 2, nf = deg + 1;
  p->len = (int64_t)ncell * np * 4;
  p->dofs = p->len;
  FRB_TRY(alloc_common(p));
  auto up = [&](auto **dst, const auto *src, size_t n) -> int {
    FRB_CUDA(cudaMalloc(dst, sizeof(**dst) * n));
    FRB_CUDA(cudaMemcpy(*dst, src, sizeof(**dst) * n, cudaMemcpyHostToDevice));
    return FRB_OK;
  };
  FRB_TRY(up(&p->tri_type, cell_type, (size_t)ncell));
  FRB_TRY(up(&p->J, J, (size_t)ncell * 4));
  FRB_TRY(up(&p->tri_normals, normals, (size_t)ncell * 6));
  FRB_TRY(up(&p->tri_fpn, fpn, (size_t)ncell * 3 * nf * 3));
  std::vector<double> ops;
  ops.insert(ops.end(), lf, lf + 3 * nf * np);
  ops.insert(ops.end(), dl, dl + 2 * np * np);
  ops.insert(ops.end(), phi, phi + 3 * nf * np);
  FRB_TRY(up(&p->tri_ops, ops.data(), ops.size()));
  if (cudaMalloc(&p->tri_uf, sizeof(double) * (size_t)ncell * 3 * nf * 4) != cudaSuccess) {
    frb_prob_destroy(p);
    frb_set_error("frb_tri_euler_create: cudaMalloc failed");
    return FRB_ERR_CUDA;
  }
  *out = p;
  return FRB_OK;
}

extern "C" int64_t frb_state_len(frb_prob_t p) { return p ? p->len : 0; }
extern "C" int64_t frb_interior_dofs(frb_prob_t p) { return p ? p->dofs : 0; }

// ---- layout bookkeeping ---------------------------------------------------------------
// p->u (reference image) and p->ru (row-chunk mirror) are two representations of one state.
// Every entry point that touches p->u first brings it up to date; the ones that may change it
// invalidate the mirror.  frb_step on the RC path does the opposite.
static int halo_flush_rc(frb_prob_t p, double *U);

static int need_ref(frb_prob_t p, bool will_write) {
  if (!p->ref_valid) {
    int n = halo_flush_rc(p, p->ru);  // slab-parallel: halo rows out of the stage kernel's ring first
    if (n < 0) return n;
    n = frb_rc_to_ref(p, p->ru, p->u);
    if (n < 0) return n;
    p->launches += n;
    p->ref_valid = true;
  }
  if (will_write) p->rc_valid = false;
  return FRB_OK;
}

// ---- state movement -------------------------------------------------------------------
extern "C" int32_t frb_state_upload(frb_prob_t p, const double *u_host) {
  FRB_REQUIRE(p && u_host, FRB_ERR_ARG, "frb_state_upload: NULL argument");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  FRB_CUDA(cudaMemcpyAsync(p->u, u_host, sizeof(double) * p->len, cudaMemcpyHostToDevice, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  p->ref_valid = true;
  p->rc_valid = false;
  return FRB_OK;
}

extern "C" int32_t frb_state_download(frb_prob_t p, double *u_host) {
  FRB_REQUIRE(p && u_host, FRB_ERR_ARG, "frb_state_download: NULL argument");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  if (int rc = need_ref(p, false)) return rc;
  FRB_CUDA(cudaMemcpyAsync(u_host, p->u, sizeof(double) * p->len, cudaMemcpyDeviceToHost, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  return FRB_OK;
}

extern "C" int32_t frb_state_device_ptr(frb_prob_t p, double **dptr) {
  FRB_REQUIRE(p && dptr, FRB_ERR_ARG, "frb_state_device_ptr: NULL argument");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (int rc = need_ref(p, true)) return rc;  // the caller may write through the pointer
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  *dptr = p->u;
  return FRB_OK;
}

// ---- stage dispatch ---------------------------------------------------------------------
static bool use_march(frb_prob_t p) {
  if (p->kind != K_EULER2D || p->flux != FRB_FLUX_HLL) return false;
  if (p->kernel_kind == FRB_KERNEL_GENERIC) return false;
  return frb_euler2d_march_supported(p);
}

// frb_step streams in the row-chunk layout when it can: 2-D Euler, deg 2..3
static bool use_rc(frb_prob_t p) {
  if (p->kind != K_EULER2D || !p->rc_base) return false;  // every common flux (HLL / LF / Roe) has its instantiation
  return p->kernel_kind == FRB_KERNEL_AUTO || p->kernel_kind == FRB_KERNEL_RC;
}

static cudaEvent_t prof_event(frb_prob_t p) {
  if (p->prof_used == p->prof_events.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    p->prof_events.push_back(e);
  }
  return p->prof_events[p->prof_used++];
}

static void prof_begin(frb_prob_t p) { p->prof_used = 0; p->stage_ms = 0.f; p->stage_count = 0; }

static void prof_collect(frb_prob_t p) {  // call after the stream has been synchronised
  for (size_t q = 0; q + 1 < p->prof_used; q += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p->prof_events[q], p->prof_events[q + 1]) == cudaSuccess) {
      p->stage_ms += ms;
      p->stage_count += 1;
    }
  }
  p->prof_used = 0;
}

static int launch_stage_inner(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st);

static int launch_stage(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  if (!p->profiling) return launch_stage_inner(p, u, ua, out, st);
  cudaEvent_t e0 = prof_event(p), e1 = prof_event(p);
  if (e0 && e1) cudaEventRecord(e0, p->ctx->stream);
  int n = launch_stage_inner(p, u, ua, out, st);
  if (e0 && e1) cudaEventRecord(e1, p->ctx->stream);
  return n;
}

static int launch_stage_inner(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  int n;
  switch (p->kind) {
    case K_ADV1D: n = frb_launch_adv1d(p, u, ua, out, st); break;
    case K_EULER1D: n = frb_launch_euler1d(p, u, ua, out, st); break;
    case K_BGK1D: n = frb_launch_bgk1d(p, u, ua, out, st); break;
    case K_EULER2D:
      n = p->curv_iJ   ? frb_launch_euler2d_curv(p, u, ua, out, st)
          : use_march(p) ? frb_launch_euler2d_march(p, u, ua, out, st)
                         : frb_launch_euler2d_generic(p, u, ua, out, st);
      break;
    case K_NS2D: n = frb_launch_ns2d(p, u, ua, out, st); break;
    case K_TRI_EULER: n = frb_launch_tri_euler(p, u, ua, out, st); break;
    default: frb_set_error("unknown problem kind"); return FRB_ERR_STATE;
  }
  if (n > 0) p->launches += n;
  return n;
}

static bool is2d(frb_prob_t p) { return p->kind == K_EULER2D || p->kind == K_NS2D; }

static int ensure_du(frb_prob_t p) {
  if (!p->du) {
    FRB_CUDA(cudaMalloc(&p->du, sizeof(double) * p->len));
    FRB_CUDA(cudaMemsetAsync(p->du, 0, sizeof(double) * p->len, p->ctx->stream));  // du = 0 in ghosts
  }
  return FRB_OK;
}

static int halo_wait_if_pending(frb_prob_t p);
static int halo_republish(frb_prob_t p, const double *U);

// Does f!(du,u,p,t) of the resident state go through the row-chunk kernel (the one frb_step and bench.py
// run)?  Always when FRB_KERNEL_RC is selected; under AUTO when the resident state already lives in that
// layout (a call between steps: no conversion of the state, only of du).
static bool rhs_via_rc(frb_prob_t p, bool host_u) {
  if (!use_rc(p) || frb_halo_active(p)) return false;
  if (p->kernel_kind == FRB_KERNEL_RC) return true;
  if (p->kernel_kind == FRB_KERNEL_AUTO && p->flux != FRB_FLUX_HLL) return true;  // the reference-image marching
  // kernel is HLL-only: LF / Roe take the row-chunk kernel (two conversions) rather than the generic one
  return !host_u && p->rc_valid;
}

extern "C" int32_t frb_rhs(frb_prob_t p, const double *u_host, double *du_host, double t) {
  (void)t;
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_rhs: prob is NULL");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  if (int rc = ensure_du(p)) return rc;
  const bool via_rc = rhs_via_rc(p, u_host != nullptr);
  // f! is pure with respect to the integrator (SciML contract): a caller-supplied u goes into the idle
  // stage buffer s1, the resident state p->u / p->ru stays what the last upload / step left
  const double *src = p->u;
  if (u_host) {
    FRB_CUDA(cudaMemcpyAsync(p->s1, u_host, sizeof(double) * p->len, cudaMemcpyHostToDevice, s));
    src = p->s1;
  } else if (!via_rc || !p->rc_valid) {
    if (int rc = need_ref(p, false)) return rc;
  }
  const int64_t l0 = p->launches;
  prof_begin(p);
  FRB_CUDA(cudaEventRecord(p->ev0, s));
  FrbStage st = {0.0, 0.0, 1.0, 0, 1};
  if (int rc = halo_wait_if_pending(p)) return rc;  // slab-parallel: the neighbours' rows of p->u have landed
  int n;
  if (via_rc) {
    // row-chunk kernel: state in RC (resident mirror, or the caller's u converted into rs1), L(u) into
    // rs2, interior of rs2 back into the reference image of du (du = 0 in the ghosts)
    const double *rsrc = p->ru;
    if (u_host || !p->rc_valid) {
      if ((n = frb_rc_from_ref(p, src, p->rs1)) < 0) return n;
      p->launches += n;
      rsrc = p->rs1;
    }
    cudaEvent_t e0 = p->profiling ? prof_event(p) : nullptr, e1 = p->profiling ? prof_event(p) : nullptr;
    if (e0 && e1) cudaEventRecord(e0, s);
    n = frb_launch_euler2d_rc(p, rsrc, nullptr, p->rs2, st);
    if (e0 && e1) cudaEventRecord(e1, s);
    if (n < 0) return n;
    p->launches += n;
    if ((n = frb_rc_to_ref(p, p->rs2, p->du, true)) < 0) return n;
    p->launches += n;
  } else {
    n = launch_stage(p, src, nullptr, p->du, st);
    if (n < 0) return n;
  }
  FRB_CUDA(cudaEventRecord(p->ev1, s));
  if (du_host)
    FRB_CUDA(cudaMemcpyAsync(du_host, p->du, sizeof(double) * p->len, cudaMemcpyDeviceToHost, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  FRB_CUDA(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
  p->last_launches = p->launches - l0;
  prof_collect(p);
  return FRB_OK;
}

// f!(du,u,p,t) with host buffers, streamed: the 2-D state goes through the device in row slabs so
// that the H2D copy of slab s+1, the residual of slab s and the D2H copy of slab s-1 overlap
// (three streams, per-slab events).  Same result as frb_rhs; the win is PCIe full duplex.
extern "C" int32_t frb_rhs_pipelined(frb_prob_t p, const double *u_host, double *du_host, int32_t nslab) {
  FRB_REQUIRE(p && u_host && du_host, FRB_ERR_ARG, "frb_rhs_pipelined: NULL argument");
  FRB_REQUIRE(p->kind == K_EULER2D, FRB_ERR_STATE, "frb_rhs_pipelined: euler2d problems only");
  if (!use_march(p) || nslab <= 1 || p->ny < 2 * nslab) return frb_rhs(p, u_host, du_host, 0.0);
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (int rc = ensure_du(p)) return rc;
  cudaStream_t sc = p->ctx->stream, si = p->ctx->copy_in, so = p->ctx->copy_out;
  double *const U = p->s1;  // the caller's u streams through the idle stage buffer: the resident state survives
  FRB_CUDA(cudaStreamSynchronize(sc));  // du memset / earlier work
  const size_t NXG = p->nx + 2, NE = NXG * (size_t)(p->ny + 2);
  const int nplanes = 4 * p->nsp * p->nsp;
  const size_t pitch = NE * sizeof(double);
  const int rows = (p->ny + nslab - 1) / nslab;
  // per-slab events live with the problem (created once, reused by every call)
  while ((int)p->pipe_events.size() < 2 * nslab) {
    cudaEvent_t e = nullptr;
    FRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->pipe_events.push_back(e);
  }
  cudaEvent_t *const ev_in = p->pipe_events.data(), *const ev_c = p->pipe_events.data() + nslab;
  const int64_t l0 = p->launches;
  prof_begin(p);
  FRB_CUDA(cudaEventRecord(p->ev0, sc));
  int up_next = 0;  // first row not yet uploaded
  int rc = FRB_OK;
  for (int s = 0; s < nslab && rc == FRB_OK; ++s) {
    const int a = 1 + s * rows, b = std::min(p->ny, a + rows - 1);
    if (a > p->ny) break;
    const int up_to = b + 1;  // rows a-1 .. b+1 are needed
    if (up_to >= up_next) {
      const size_t off = NXG * (size_t)up_next, w = NXG * (size_t)(up_to - up_next + 1) * sizeof(double);
      FRB_CUDA(cudaMemcpy2DAsync(U + off, pitch, u_host + off, pitch, w, nplanes, cudaMemcpyHostToDevice, si));
      up_next = up_to + 1;
    }
    FRB_CUDA(cudaEventRecord(ev_in[s], si));
    FRB_CUDA(cudaStreamWaitEvent(sc, ev_in[s], 0));
    p->row_lo = a;
    p->row_hi = b;
    FrbStage st = {0.0, 0.0, 1.0, 0, 1, 0};
    int n = launch_stage(p, U, nullptr, p->du, st);
    p->row_lo = p->row_hi = 0;
    if (n < 0) { rc = n; break; }
    FRB_CUDA(cudaEventRecord(ev_c[s], sc));
    FRB_CUDA(cudaStreamWaitEvent(so, ev_c[s], 0));
    // du rows a..b (plus the zero ghost row next to the first / last slab)
    const int d0 = a == 1 ? 0 : a, d1 = b == p->ny ? p->ny + 1 : b;
    const size_t off = NXG * (size_t)d0, w = NXG * (size_t)(d1 - d0 + 1) * sizeof(double);
    FRB_CUDA(cudaMemcpy2DAsync(du_host + off, pitch, p->du + off, pitch, w, nplanes, cudaMemcpyDeviceToHost, so));
  }
  FRB_CUDA(cudaEventRecord(p->ev1, sc));
  FRB_CUDA(cudaStreamSynchronize(si));
  FRB_CUDA(cudaStreamSynchronize(sc));
  FRB_CUDA(cudaStreamSynchronize(so));
  if (rc != FRB_OK) return rc;
  FRB_CUDA(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
  p->last_launches = p->launches - l0;
  prof_collect(p);
  return FRB_OK;
}

// ---- hooks ------------------------------------------------------------------------------
static int set_limiter_weights(frb_prob_t p, const double *w) {
  const size_t n = is2d(p) ? (size_t)p->nsp * p->nsp : (size_t)p->nsp;
  if (!p->lim_w) FRB_CUDA(cudaMalloc(&p->lim_w, sizeof(double) * n));
  FRB_CUDA(cudaMemcpy(p->lim_w, w, sizeof(double) * n, cudaMemcpyHostToDevice));
  return FRB_OK;
}

extern "C" int32_t frb_set_step_hooks(frb_prob_t p, int32_t ghost_mode, const double *limiter_weights) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_set_step_hooks: prob is NULL");
  FRB_REQUIRE(ghost_mode >= FRB_GHOST_NONE && ghost_mode <= FRB_GHOST_CYLINDER, FRB_ERR_ARG,
              "frb_set_step_hooks: unknown ghost mode");
  FRB_REQUIRE(ghost_mode != FRB_GHOST_CYLINDER || p->curv_iJ, FRB_ERR_STATE,
              "frb_set_step_hooks: the cylinder ghost fill applies to curvilinear euler2d problems");
  FRB_REQUIRE(ghost_mode == FRB_GHOST_NONE || p->kind == K_EULER2D, FRB_ERR_STATE,
              "frb_set_step_hooks: ghost fill applies to euler2d problems");
  FRB_REQUIRE(!limiter_weights || p->kind == K_EULER1D || p->kind == K_EULER2D, FRB_ERR_STATE,
              "frb_set_step_hooks: the positivity limiter applies to Euler problems");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  p->ghost_mode = ghost_mode;
  p->limiter_on = limiter_weights != nullptr;
  if (limiter_weights)
    if (int rc = set_limiter_weights(p, limiter_weights)) return rc;
  return FRB_OK;
}

static int run_limiter(frb_prob_t p, bool rc = false) {
  int n = p->kind == K_EULER1D ? frb_launch_limiter1d(p, p->u)
          : rc                 ? frb_rc_limiter2d(p, p->ru)
                               : frb_launch_limiter2d(p, p->u);
  if (n > 0) p->launches += n;
  return n;
}

extern "C" int32_t frb_ghost_fill(frb_prob_t p, int32_t ghost_mode) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_ghost_fill: prob is NULL");
  FRB_REQUIRE(p->kind == K_EULER2D, FRB_ERR_STATE, "frb_ghost_fill: euler2d problems only");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (int rc = need_ref(p, true)) return rc;
  int n;
  if (frb_halo_active(p)) {
    // slab-parallel: rows 0 / ny+1 are the neighbours' rows, not ghosts -- the x half runs locally, the y half
    // is the exchange across the global seam (first / last rank), as in the per-step hook of frb_step
    if ((n = halo_wait_if_pending(p)) < 0) return n;
    if ((n = frb_launch_ghost_x2d(p, p->u, ghost_mode)) < 0) return n;
    p->launches += n;
    if ((n = frb_halo_push(p, p->u, 0, true, ghost_mode == FRB_GHOST_WAVE_X ? 2 : -1)) < 0) return n;
    p->launches += n;
    if ((n = halo_republish(p, p->u)) < 0) return n;
  } else {
    if ((n = frb_launch_ghost_fill2d(p, p->u, ghost_mode)) < 0) return n;
    p->launches += n;
  }
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return FRB_OK;
}

extern "C" int32_t frb_limiter_positivity(frb_prob_t p, const double *weights, int32_t *nbad) {
  FRB_REQUIRE(p && weights, FRB_ERR_ARG, "frb_limiter_positivity: NULL argument");
  FRB_REQUIRE(p->kind == K_EULER1D || p->kind == K_EULER2D, FRB_ERR_STATE,
              "frb_limiter_positivity: Euler problems only");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (int rc = set_limiter_weights(p, weights)) return rc;
  if (int rc = need_ref(p, true)) return rc;
  FRB_CUDA(cudaMemsetAsync(p->flag, 0, sizeof(int), p->ctx->stream));
  int n = run_limiter(p);
  if (n < 0) return n;
  if ((n = halo_republish(p, p->u)) < 0) return n;
  int bad = 0;
  FRB_CUDA(cudaMemcpyAsync(&bad, p->flag, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (nbad) *nbad = bad;
  if (bad) {
    frb_set_error("incorrect range of limiter parameter t");  // dissipation.jl:84
    return FRB_ERR_NUMERIC;
  }
  return FRB_OK;
}

// ---- shock sensor + modal filter ---------------------------------------------------------
static int set_filter(frb_prob_t p, const double *iV, const double *F, int np) {
  const int want = p->kind == K_EULER1D ? p->nsp : p->nsp * p->nsp;
  FRB_REQUIRE(p->kind == K_EULER1D || p->kind == K_EULER2D, FRB_ERR_STATE, "modal filter: Euler problems only");
  FRB_REQUIRE(iV && F && np == want, FRB_ERR_ARG, "modal filter: iV, F must be np x np with np = points per element");
  if (!p->filt || p->filt_np != np) {
    cudaFree(p->filt);
    p->filt = nullptr;
    FRB_CUDA(cudaMalloc(&p->filt, sizeof(double) * 2 * np * np));
    p->filt_np = np;
  }
  FRB_CUDA(cudaMemcpy(p->filt, iV, sizeof(double) * np * np, cudaMemcpyHostToDevice));
  FRB_CUDA(cudaMemcpy(p->filt + np * np, F, sizeof(double) * np * np, cudaMemcpyHostToDevice));
  return FRB_OK;
}

static int run_filter(frb_prob_t p, int *count = nullptr, bool rc = false) {
  const double *iV = p->filt, *F = p->filt + p->filt_np * p->filt_np;
  int n = rc ? frb_rc_modal_filter(p, p->ru, iV, F, p->filt_eps, p->filt_S0, p->filt_kappa, p->filt_ghosts != 0)
             : frb_launch_modal_filter(p, p->u, iV, F, p->filt_eps, p->filt_S0, p->filt_kappa, p->filt_ghosts != 0,
                                       count);
  if (n > 0) p->launches += n;
  return n;
}

extern "C" int32_t frb_filter_modal(frb_prob_t p, const double *iV, const double *F, int32_t np, double eps,
                                    double S0, double kappa, int32_t include_ghosts, int32_t *nfiltered) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_filter_modal: prob is NULL");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (int rc = set_filter(p, iV, F, np)) return rc;
  if (int rc = need_ref(p, true)) return rc;
  const int when = p->filt_when;  // a stand-alone call does not disturb the hook's parameters
  const double e0 = p->filt_eps, s0 = p->filt_S0, k0 = p->filt_kappa;
  const int g0 = p->filt_ghosts;
  p->filt_eps = eps; p->filt_S0 = S0; p->filt_kappa = kappa; p->filt_ghosts = include_ghosts;
  FRB_CUDA(cudaMemsetAsync(p->flag, 0, sizeof(int), p->ctx->stream));
  int n = run_filter(p, p->flag);
  if (n >= 0) {
    int r = halo_republish(p, p->u);
    if (r < 0) n = r;
  }
  p->filt_eps = e0; p->filt_S0 = s0; p->filt_kappa = k0; p->filt_ghosts = g0; p->filt_when = when;
  if (n < 0) return n;
  int cnt = 0;
  FRB_CUDA(cudaMemcpyAsync(&cnt, p->flag, sizeof(int), cudaMemcpyDeviceToHost, p->ctx->stream));
  FRB_CUDA(cudaStreamSynchronize(p->ctx->stream));
  if (nfiltered) *nfiltered = cnt;
  return FRB_OK;
}

extern "C" int32_t frb_set_filter_hook(frb_prob_t p, int32_t when, const double *iV, const double *F, int32_t np,
                                       double eps, double S0, double kappa, int32_t include_ghosts) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_set_filter_hook: prob is NULL");
  FRB_REQUIRE(when >= 0 && when <= 2, FRB_ERR_ARG, "frb_set_filter_hook: when must be 0, 1 or 2");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  if (when != 0) {
    if (int rc = set_filter(p, iV, F, np)) return rc;
    p->filt_eps = eps; p->filt_S0 = S0; p->filt_kappa = kappa; p->filt_ghosts = include_ghosts;
  }
  p->filt_when = when;
  return FRB_OK;
}

// ---- step! ------------------------------------------------------------------------------
static int halo_wait_if_pending(frb_prob_t p) {
  if (!p->halo_pending) return 0;
  int n = frb_halo_wait(p);
  if (n < 0) return n;
  p->launches += n;
  p->halo_pending = false;
  p->halo_pending_legacy = false;
  return 0;
}

// everything that reads a row-chunk buffer as a whole array (conversion to the reference image, f! of the
// resident slab, the limiter / filter passes) needs the interior halo rows in the array, not in the ring
static int halo_flush_rc(frb_prob_t p, double *U) {
  if (!frb_halo_active(p) || frb_halo_rc_input_in_array(p, U)) return 0;
  int n;
  if ((n = halo_wait_if_pending(p)) < 0) return n;
  if ((n = frb_halo_rc_flush(p, U)) < 0) return n;
  return 0;
}

// Slab-parallel path: something other than a stage (limiter, modal filter) has just rewritten rows of U that
// the neighbours hold copies of in their halo rows.  A neighbour barrier (everybody is done rewriting -- the
// filter may also touch its own halo cells), then the boundary rows go out again, then the usual epoch.
static int halo_republish(frb_prob_t p, const double *U) {
  if (!frb_halo_active(p)) return 0;
  int n;
  if ((n = halo_wait_if_pending(p)) < 0) return n;
  if ((n = frb_halo_signal(p)) < 0) return n;
  p->launches += n;
  p->halo_pending = true;
  if ((n = halo_wait_if_pending(p)) < 0) return n;
  if ((n = frb_halo_push(p, U, 0, false, -1)) < 0) return n;
  p->launches += n;
  if ((n = frb_halo_signal(p)) < 0) return n;
  p->launches += n;
  p->halo_pending = true;
  return 0;
}

// one fused stage; on the slab-parallel path preceded by the wait for the neighbours' rows of
// the previous stage and followed by the push of this stage's boundary rows + the flag
static int stage_x(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st, bool rc) {
  int n;
  if (rc) {
    // Row-chunk path.  Slab-parallel: the exchange is inside the stage kernel (RcHalo) -- no wait kernel unless
    // rows were pushed into the arrays since the last one (per-step seam rows, a fresh upload), no signal kernel.
    RcHalo h;
    const bool inker = frb_halo_rc_inkernel(p);
    if (inker) {
      if (p->halo_pending_legacy && (n = halo_wait_if_pending(p)) < 0) return n;
      if ((n = frb_halo_rc_stage(p, u, out, &h)) < 0) return n;
    } else if ((n = halo_wait_if_pending(p)) < 0) {
      return n;
    }
    cudaEvent_t e0 = p->profiling ? prof_event(p) : nullptr, e1 = p->profiling ? prof_event(p) : nullptr;
    if (e0 && e1) cudaEventRecord(e0, p->ctx->stream);
    n = frb_launch_euler2d_rc(p, u, ua, out, st, inker ? &h : nullptr);
    if (e0 && e1) cudaEventRecord(e1, p->ctx->stream);
    if (n < 0) return n;
    p->launches += n;
    if (inker) {
      p->halo_pending = true;  // the epoch was raised by the kernel; array-wide readers wait for it (halo_flush_rc)
      return 0;
    }
    if (frb_halo_active(p)) {  // FRB_HALO_LEGACY: rows pushed after the stage, epochs by signal / wait kernels
      if ((n = frb_halo_push(p, out, frb_halo_role(p, out), false, -1)) < 0) return n;
      p->launches += n;
      if ((n = frb_halo_signal(p)) < 0) return n;
      p->launches += n;
      p->halo_pending = true;
    }
    return 0;
  }
  if ((n = halo_wait_if_pending(p)) < 0) return n;
  if ((n = launch_stage(p, u, ua, out, st)) < 0) return n;
  if (frb_halo_active(p)) {
    if (!use_march(p)) {  // the marching kernel stores its boundary rows to the peers itself
      if ((n = frb_halo_push(p, out, frb_halo_role(p, out), false, -1)) < 0) return n;
      p->launches += n;
    }
    if ((n = frb_halo_signal(p)) < 0) return n;
    p->launches += n;
    p->halo_pending = true;
  }
  return 0;
}

static int one_step(frb_prob_t p, int scheme, double dt, bool rc) {
  int n;
  const bool par = frb_halo_active(p);
  double *&U = rc ? p->ru : p->u, *&S1 = rc ? p->rs1 : p->s1, *&S2 = rc ? p->rs2 : p->s2;
  if (rc && par && (p->filt_when != 0 || p->limiter_on)) {
    if ((n = halo_flush_rc(p, U)) < 0) return n;  // the hooks work on the array as a whole
  }
  if (p->filt_when == 1) {
    if ((n = run_filter(p, nullptr, rc)) < 0) return n;
  }
  if (p->limiter_on) {
    if ((n = run_limiter(p, rc)) < 0) return n;
  }
  if (par && (p->filt_when == 1 || p->limiter_on)) {
    if ((n = halo_republish(p, U)) < 0) return n;  // the neighbours' halo copies of my rows 1 / ny are stale
  }
  bool ring_done = false;
  if (p->ghost_mode != FRB_GHOST_NONE) {
    if (par) {
      if ((n = halo_wait_if_pending(p)) < 0) return n;  // neighbours are done with the last stage
      if (rc) {  // x ghosts of u_n and the frozen ghost columns of the stage buffers in one launch
        n = frb_rc_ghost_x_ring(p, U, S1, scheme != FRB_SCHEME_EULER ? S2 : nullptr, p->ghost_mode);
        ring_done = true;
      } else {
        n = frb_launch_ghost_x2d(p, U, p->ghost_mode);
      }
    } else {
      n = rc ? frb_rc_ghost_fill(p, U, p->ghost_mode) : frb_launch_ghost_fill2d(p, U, p->ghost_mode);
    }
    if (n < 0) return n;
    p->launches += n;
  }
  if (p->kind == K_EULER2D && !ring_done) {
    // ghosts are frozen across the stages of a step (du = 0 there): give the stage buffers the
    // same ring as u_n.  Rows owned by a neighbouring rank are excluded (the neighbour writes them).
    int nr = 1, rk = frb_halo_rank(p, &nr);
    const bool seam_local = !par || p->ghost_mode == FRB_GHOST_NONE;
    const bool row0 = !par || (rk == 0 && seam_local), rowN = !par || (rk == nr - 1 && seam_local);
    n = rc ? frb_rc_ring_copy(p, U, S1, row0, rowN) : frb_launch_ring_copy2d(p, U, S1, row0, rowN);
    if (n < 0) return n;
    p->launches += n;
    if (scheme != FRB_SCHEME_EULER) {
      n = rc ? frb_rc_ring_copy(p, U, S2, row0, rowN) : frb_launch_ring_copy2d(p, U, S2, row0, rowN);
      if (n < 0) return n;
      p->launches += n;
    }
  }
  if (par && p->ghost_mode != FRB_GHOST_NONE) {
    // the y half of the ghost fill crosses the global seam: first/last rank exchange their
    // boundary rows (frozen for the step, so they go into all three buffers of the peer)
    const int flip = p->ghost_mode == FRB_GHOST_WAVE_X ? 2 : -1;
    if ((n = frb_halo_push_seam_all(p, U, flip)) < 0) return n;
    p->launches += n;
    if ((n = frb_halo_signal(p)) < 0) return n;
    p->launches += n;
    p->halo_pending = true;
  }
  if (scheme == FRB_SCHEME_EULER) {
    FrbStage st = {0.0, 1.0, dt, 0, 0};
    if ((n = stage_x(p, U, nullptr, S1, st, rc)) < 0) return n;
    std::swap(U, S1);
    frb_halo_swap_roles(p, 0, 1, rc);
  } else if (scheme == FRB_SCHEME_MIDPOINT) {
    FrbStage a = {0.0, 1.0, 0.5 * dt, 0, 0};
    if ((n = stage_x(p, U, nullptr, S1, a, rc)) < 0) return n;
    // the last stage of a step never runs in place (u_n read and u' written through the same
    // DRAM pages cost ~14 % of the launch): it writes the free buffer and the pointers rotate
    FrbStage b = {1.0, 0.0, dt, 1, 0};
    double *O = p->kind == K_EULER2D ? S2 : U;
    if ((n = stage_x(p, S1, U, O, b, rc)) < 0) return n;
    if (O != U) {
      std::swap(U, S2);
      frb_halo_swap_roles(p, 0, 2, rc);
    }
  } else if (scheme == FRB_SCHEME_SSPRK3) {
    FrbStage a = {0.0, 1.0, dt, 0, 0};
    if ((n = stage_x(p, U, nullptr, S1, a, rc)) < 0) return n;
    FrbStage b = {0.75, 0.25, dt, 1, 0, 1};
    if ((n = stage_x(p, S1, U, S2, b, rc)) < 0) return n;
    FrbStage c = {1.0 / 3.0, 2.0 / 3.0, dt, 1, 0, 1};
    double *O = p->kind == K_EULER2D ? S1 : U;  // s1 is free again; see the Midpoint branch
    if ((n = stage_x(p, S2, U, O, c, rc)) < 0) return n;
    if (O != U) {
      std::swap(U, S1);
      frb_halo_swap_roles(p, 0, 1, rc);
    }
  } else {
    frb_set_error("frb_step: unknown scheme");
    return FRB_ERR_ARG;
  }
  if (p->filt_when == 2) {
    if ((n = run_filter(p, nullptr, rc)) < 0) return n;
    if (par && (n = halo_republish(p, U)) < 0) return n;
  }
  return FRB_OK;
}

// One step of an explicit Runge-Kutta scheme given by its tableau, in the reference image:
// k_i = L(u + dt sum_j a_ij k_j) through the rhs_only form of the stage kernels, combinations by
// frb_launch_lincomb over the whole array (k_j = 0 in the ghosts: they stay frozen through the step).
struct RkTab {
  int ns;
  double A[FRB_RK_MAX_STAGES * FRB_RK_MAX_STAGES], b[FRB_RK_MAX_STAGES];
};

static int tableau_step(frb_prob_t p, const RkTab &tab, double dt) {
  int n;
  if (p->filt_when == 1) {
    if ((n = run_filter(p)) < 0) return n;
  }
  if (p->limiter_on) {
    if ((n = run_limiter(p)) < 0) return n;
  }
  if (p->ghost_mode != FRB_GHOST_NONE) {
    if ((n = frb_launch_ghost_fill2d(p, p->u, p->ghost_mode)) < 0) return n;
    p->launches += n;
  }
  const FrbStage rhs = {0.0, 0.0, 1.0, 0, 1, 0};
  for (int i = 0; i < tab.ns; ++i) {
    const double *row = tab.A + i * tab.ns;
    const double *src = p->u;
    bool any = false;
    for (int j = 0; j < i; ++j) any = any || row[j] != 0.0;
    if (any) {
      if ((n = frb_launch_lincomb(p, p->s1, p->u, i, p->rk_k.data(), row, dt)) < 0) return n;
      p->launches += n;
      src = p->s1;
    }
    if ((n = launch_stage(p, src, nullptr, p->rk_k[i], rhs)) < 0) return n;
  }
  if ((n = frb_launch_lincomb(p, p->u, p->u, tab.ns, p->rk_k.data(), tab.b, dt)) < 0) return n;
  p->launches += n;
  if (p->filt_when == 2) {
    if ((n = run_filter(p)) < 0) return n;
  }
  return FRB_OK;
}

// bring the row-chunk mirror up to date (first RC step after an upload / a reference-image call)
static int need_rc(frb_prob_t p) {
  if (p->rc_valid) return FRB_OK;
  int n;
  if ((n = halo_wait_if_pending(p)) < 0) return n;  // the neighbours' rows of p->u have landed
  if ((n = frb_rc_from_ref(p, p->u, p->ru)) < 0) return n;
  p->launches += n;
  p->rc_valid = true;
  frb_halo_rc_reset(p);  // the halo rows came with the conversion: they are in the arrays
  if (frb_halo_active(p)) {
    // the neighbours may store into my RC halo rows only after this conversion: one more epoch
    if ((n = frb_halo_signal(p)) < 0) return n;
    p->launches += n;
    p->halo_pending = true;
  }
  return FRB_OK;
}

// The time loop.  Long runs of small problems are launch-bound (cfg1: 2.4 KB of state, 5 us per
// launch), so from 16 steps on a pair of steps is captured once into a CUDA graph and replayed:
// two steps because the stage buffers rotate with period 2.  The first two steps run eagerly --
// they do the lazy allocations, descriptor caches and attribute opt-ins that must not happen inside
// a capture.  Not used with per-stage profiling (event pairs) or the slab-parallel path (the halo
// epochs are kernel arguments that change every stage).
static int run_steps(frb_prob_t p, int scheme, double dt, bool rc, int nsteps, const RkTab *tab = nullptr) {
  cudaStream_t s = p->ctx->stream;
  int it = 0, r;
  const bool graphable = nsteps >= 16 && !p->profiling && !frb_halo_active(p) && !getenv("FRB_NO_GRAPH");
  auto step = [&]() { return tab ? tableau_step(p, *tab, dt) : one_step(p, scheme, dt, rc); };
  if (graphable && !tab && p->filt_when == 0 && !getenv("FRB_NO_LOOP1D")) {
    // small 1-D problems: the whole loop in one cooperative launch (grid barrier instead of launches)
    r = frb_launch_loop1d(p, scheme, dt, nsteps);
    if (r < 0) return r;
    if (r > 0) {
      p->launches += r;
      return FRB_OK;
    }
  }
  if (graphable) {
    for (; it < 2; ++it)
      if ((r = step()) < 0) return r;
    const int pairs = (nsteps - it) / 2;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    FRB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    const int64_t l0 = p->launches;
    r = step();
    if (r >= 0) r = step();
    cudaError_t ce = cudaStreamEndCapture(s, &graph);  // always: leaves the stream usable
    if (r < 0) {
      if (graph) cudaGraphDestroy(graph);
      return r;
    }
    if (ce != cudaSuccess) return frb_cuda_fail(ce, "cudaStreamEndCapture", __FILE__, __LINE__);
    const int64_t per_pair = p->launches - l0;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) {
      cudaGraphDestroy(graph);
      return frb_cuda_fail(ce, "cudaGraphInstantiate", __FILE__, __LINE__);
    }
    for (int q = 0; q < pairs && ce == cudaSuccess; ++q) ce = cudaGraphLaunch(exec, s);
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) return frb_cuda_fail(ce, "cudaGraphLaunch", __FILE__, __LINE__);
    p->launches += per_pair * (pairs - 1);
    it += 2 * pairs;
  }
  for (; it < nsteps; ++it)
    if ((r = step()) < 0) return r;
  return FRB_OK;
}

extern "C" int32_t frb_step(frb_prob_t p, int32_t scheme, double dt, int32_t nsteps) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_step: prob is NULL");
  FRB_REQUIRE(nsteps >= 0, FRB_ERR_ARG, "frb_step: nsteps must be >= 0");
  FRB_REQUIRE(scheme >= FRB_SCHEME_EULER && scheme <= FRB_SCHEME_SSPRK3, FRB_ERR_ARG, "frb_step: unknown scheme");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  const int64_t l0 = p->launches;
  const bool rc = use_rc(p);
  prof_begin(p);
  if (nsteps > 0) {
    if (rc) {
      if (int r = need_rc(p)) return r;
      p->ref_valid = false;
    } else if (int r = need_ref(p, true)) {
      return r;
    }
  }
  if (p->limiter_on) FRB_CUDA(cudaMemsetAsync(p->flag, 0, sizeof(int), s));
  FRB_CUDA(cudaEventRecord(p->ev0, s));
  if (int r = run_steps(p, scheme, dt, rc, nsteps)) return r;
  FRB_CUDA(cudaEventRecord(p->ev1, s));
  if (rc) {
    if (int r = halo_flush_rc(p, p->ru)) return r;  // outside the timed region: two row copies per call
  }
  int bad = 0;
  if (p->limiter_on)
    FRB_CUDA(cudaMemcpyAsync(&bad, p->flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  FRB_CUDA(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
  p->last_launches = p->launches - l0;
  prof_collect(p);
  if (int r = frb_halo_check_timeout(p)) return r;
  if (bad) {
    frb_set_error("incorrect range of limiter parameter t");
    return FRB_ERR_NUMERIC;
  }
  return FRB_OK;
}

extern "C" int32_t frb_step_tableau(frb_prob_t p, int32_t ns, const double *A, const double *b, double dt,
                                    int32_t nsteps) {
  FRB_REQUIRE(p && A && b, FRB_ERR_ARG, "frb_step_tableau: NULL argument");
  FRB_REQUIRE(ns >= 1 && ns <= FRB_RK_MAX_STAGES, FRB_ERR_ARG, "frb_step_tableau: nstage must be 1..8");
  FRB_REQUIRE(nsteps >= 0, FRB_ERR_ARG, "frb_step_tableau: nsteps must be >= 0");
  FRB_REQUIRE(!frb_halo_active(p), FRB_ERR_STATE, "frb_step_tableau: not available on the slab-parallel path");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  RkTab tab;
  tab.ns = ns;
  for (int i = 0; i < ns; ++i) {
    tab.b[i] = b[i];
    for (int j = 0; j < ns; ++j) tab.A[i * ns + j] = j < i ? A[i * ns + j] : 0.0;
  }
  const int64_t l0 = p->launches;
  prof_begin(p);
  if (nsteps > 0) {
    if (int r = need_ref(p, true)) return r;
    while ((int)p->rk_k.size() < ns) {
      double *k = nullptr;
      FRB_CUDA(cudaMalloc(&k, sizeof(double) * (size_t)p->len));
      p->rk_k.push_back(k);
      FRB_CUDA(cudaMemsetAsync(k, 0, sizeof(double) * (size_t)p->len, s));
    }
  }
  if (p->limiter_on) FRB_CUDA(cudaMemsetAsync(p->flag, 0, sizeof(int), s));
  FRB_CUDA(cudaEventRecord(p->ev0, s));
  if (int r = run_steps(p, -1, dt, false, nsteps, &tab)) return r;
  FRB_CUDA(cudaEventRecord(p->ev1, s));
  int bad = 0;
  if (p->limiter_on) FRB_CUDA(cudaMemcpyAsync(&bad, p->flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  FRB_CUDA(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
  p->last_launches = p->launches - l0;
  prof_collect(p);
  if (bad) {
    frb_set_error("incorrect range of limiter parameter t");
    return FRB_ERR_NUMERIC;
  }
  return FRB_OK;
}

// step!(itg) with the state on the HOST between steps (the reference's user loop, euler2d_wave.jl:125-135: the user
// touches itg.u -- ghost fill, limiter, filter -- and calls step!).  The plain form is upload + frb_step + download:
// the two copies run one after the other (2 x 39 ms at cfg3 for 3.3 ms of kernels).  Here the step streams through
// the device in row slabs: slab k + 1 is on its way up while the stages run on the slabs that have arrived -- stage
// s of slab k needs stage s - 1 of slabs k - 1 .. k + 1, so the last stage of slab k can run once slab k + n_stages
// is on the device -- and finished slabs are on their way down: H2D and D2H overlap (PCIe full duplex).
// The ghost cells are the caller's (filled on the host, as the reference's loop does) and frozen through the step;
// step hooks (device ghost fill, limiter, filter), other fluxes / degrees and slab-parallel problems take the
// plain path.  The new state is also the resident state afterwards.
extern "C" int32_t frb_step_host(frb_prob_t p, const double *u_in, double *u_out, int32_t scheme, double dt,
                                 int32_t nslab) {
  FRB_REQUIRE(p && u_in && u_out, FRB_ERR_ARG, "frb_step_host: NULL argument");
  FRB_REQUIRE(scheme >= FRB_SCHEME_EULER && scheme <= FRB_SCHEME_SSPRK3, FRB_ERR_ARG, "frb_step_host: unknown scheme");
  const int nst = scheme == FRB_SCHEME_EULER ? 1 : scheme == FRB_SCHEME_MIDPOINT ? 2 : 3;
  const bool streamed = p->kind == K_EULER2D && use_march(p) && !frb_halo_active(p) && p->ghost_mode == FRB_GHOST_NONE &&
                        !p->limiter_on && p->filt_when == 0 && nslab > 1 && p->ny >= 2 * nslab;
  if (!streamed) {
    if (int rc = frb_state_upload(p, u_in)) return rc;
    if (frb_halo_active(p)) {
      // slab-parallel: every rank makes this call; the neighbours' halo rows of the new state are exchanged behind a
      // neighbour barrier on the mailbox (nobody pushes into a buffer its owner is still uploading)
      FRB_CUDA(cudaSetDevice(p->ctx->device));
      if (int rc = halo_republish(p, p->u)) return rc;
    }
    if (int rc = frb_step(p, scheme, dt, 1)) return rc;
    return frb_state_download(p, u_out);
  }
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t sc = p->ctx->stream, si = p->ctx->copy_in, so = p->ctx->copy_out;
  if (!p->s3) FRB_CUDA(cudaMalloc(&p->s3, sizeof(double) * p->len));
  while ((int)p->pipe_events.size() < 2 * nslab) {
    cudaEvent_t e = nullptr;
    FRB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->pipe_events.push_back(e);
  }
  cudaEvent_t *const ev_in = p->pipe_events.data(), *const ev_c = p->pipe_events.data() + nslab;
  FRB_CUDA(cudaStreamSynchronize(sc));
  double *const U = p->u, *const S1 = p->s1, *const S2 = p->s2, *const OUT = p->s3;
  const size_t NXG = p->nx + 2, NE = NXG * (size_t)(p->ny + 2);
  const int nplanes = 4 * p->nsp * p->nsp;
  const size_t pitch = NE * sizeof(double);
  const int rows = (p->ny + nslab - 1) / nslab;
  auto lo = [&](int k) { return 1 + k * rows; };
  auto hi = [&](int k) { return std::min(p->ny, lo(k) + rows - 1); };
  int ns = 0;  // slabs that exist
  while (ns < nslab && lo(ns) <= p->ny) ++ns;
  // stage s of the scheme: input, u_n, output, coefficients (out = ca u_n + cb (u + cdt L(u)))
  struct Stg { const double *in; const double *ua; double *out; FrbStage st; };
  Stg stg[3];
  if (scheme == FRB_SCHEME_EULER) {
    stg[0] = {U, nullptr, OUT, {0.0, 1.0, dt, 0, 0, 0}};
  } else if (scheme == FRB_SCHEME_MIDPOINT) {
    stg[0] = {U, nullptr, S1, {0.0, 1.0, 0.5 * dt, 0, 0, 0}};
    stg[1] = {S1, U, OUT, {1.0, 0.0, dt, 1, 0, 0}};
  } else {
    stg[0] = {U, nullptr, S1, {0.0, 1.0, dt, 0, 0, 0}};
    stg[1] = {S1, U, S2, {0.75, 0.25, dt, 1, 0, 1}};
    stg[2] = {S2, U, OUT, {1.0 / 3.0, 2.0 / 3.0, dt, 1, 0, 1}};
  }
  const int64_t l0 = p->launches;
  prof_begin(p);
  FRB_CUDA(cudaEventRecord(p->ev0, sc));
  // uploads: slab k = rows lo(k) .. hi(k), plus the ghost row next to the first / last slab
  for (int k = 0; k < ns; ++k) {
    const int a = k == 0 ? 0 : lo(k), b = k == ns - 1 ? p->ny + 1 : hi(k);
    const size_t off = NXG * (size_t)a, w = NXG * (size_t)(b - a + 1) * sizeof(double);
    FRB_CUDA(cudaMemcpy2DAsync(U + off, pitch, u_in + off, pitch, w, nplanes, cudaMemcpyHostToDevice, si));
    FRB_CUDA(cudaEventRecord(ev_in[k], si));
  }
  int rc = FRB_OK;
  auto ring = [&](int k) -> int {  // frozen ghosts of slab k's rows into the stage buffers and the result
    const int a = k == 0 ? 0 : lo(k), b = k == ns - 1 ? p->ny + 1 : hi(k);
    int n = frb_launch_ring_rows2d(p, U, nst > 1 ? S1 : nullptr, nst > 2 ? S2 : nullptr, OUT, a, b);
    if (n > 0) p->launches += n;
    return n < 0 ? n : 0;
  };
  for (int w = 0; w < ns + nst - 1 && rc == FRB_OK; ++w) {
    if (w == 0) {
      FRB_CUDA(cudaStreamWaitEvent(sc, ev_in[0], 0));
      if ((rc = ring(0)) != FRB_OK) break;
    }
    if (w + 1 < ns) {
      FRB_CUDA(cudaStreamWaitEvent(sc, ev_in[w + 1], 0));
      if ((rc = ring(w + 1)) != FRB_OK) break;
    }
    for (int s = 0; s < nst; ++s) {
      const int k = w - s;
      if (k < 0 || k >= ns) continue;
      p->row_lo = lo(k);
      p->row_hi = hi(k);
      int n = launch_stage(p, stg[s].in, stg[s].ua, stg[s].out, stg[s].st);
      p->row_lo = p->row_hi = 0;
      if (n < 0) { rc = n; break; }
      if (s == nst - 1) {  // slab k is final: on its way down
        FRB_CUDA(cudaEventRecord(ev_c[k], sc));
        FRB_CUDA(cudaStreamWaitEvent(so, ev_c[k], 0));
        const int a = k == 0 ? 0 : lo(k), b = k == ns - 1 ? p->ny + 1 : hi(k);
        const size_t off = NXG * (size_t)a, wd = NXG * (size_t)(b - a + 1) * sizeof(double);
        FRB_CUDA(cudaMemcpy2DAsync(u_out + off, pitch, OUT + off, pitch, wd, nplanes, cudaMemcpyDeviceToHost, so));
      }
    }
  }
  FRB_CUDA(cudaEventRecord(p->ev1, sc));
  FRB_CUDA(cudaStreamSynchronize(si));
  FRB_CUDA(cudaStreamSynchronize(sc));
  FRB_CUDA(cudaStreamSynchronize(so));
  if (rc != FRB_OK) return rc;
  std::swap(p->u, p->s3);  // the result is the resident state
  p->ref_valid = true;
  p->rc_valid = false;
  FRB_CUDA(cudaEventElapsedTime(&p->last_ms, p->ev0, p->ev1));
  p->last_launches = p->launches - l0;
  prof_collect(p);
  return FRB_OK;
}

// ---- measurement ------------------------------------------------------------------------
extern "C" int32_t frb_time_stage(frb_prob_t p, int32_t stage_kind, int32_t iters, float *ms) {
  FRB_REQUIRE(p && ms && iters > 0, FRB_ERR_ARG, "frb_time_stage: bad argument");
  FRB_CUDA(cudaSetDevice(p->ctx->device));
  cudaStream_t s = p->ctx->stream;
  const bool rc = use_rc(p) && !frb_halo_active(p);
  // scratch: u_n := s1 (copy of the state so that the arithmetic sees finite data)
  if (rc) {
    if (int r = need_rc(p)) return r;
    FRB_CUDA(cudaMemcpyAsync(p->rs1, p->ru, sizeof(double) * p->rc_len, cudaMemcpyDeviceToDevice, s));
  } else {
    if (int r = need_ref(p, false)) return r;
    FRB_CUDA(cudaMemcpyAsync(p->s1, p->u, sizeof(double) * p->len, cudaMemcpyDeviceToDevice, s));
  }
  FrbStage st = stage_kind == 0 ? FrbStage{0.0, 1.0, 1e-9, 0, 0} : FrbStage{0.75, 0.25, 0.25e-9, 1, 0};
  const int64_t l0 = p->launches;
  FRB_CUDA(cudaEventRecord(p->ev0, s));
  for (int it = 0; it < iters; ++it) {
    int n = rc ? frb_launch_euler2d_rc(p, p->ru, p->rs1, p->rs2, st)
               : launch_stage(p, p->u, p->s1, p->s2, st);
    if (n < 0) return n;
    if (rc) p->launches += n;
  }
  FRB_CUDA(cudaEventRecord(p->ev1, s));
  FRB_CUDA(cudaStreamSynchronize(s));
  float total = 0.f;
  FRB_CUDA(cudaEventElapsedTime(&total, p->ev0, p->ev1));
  *ms = total / iters;
  p->last_ms = total;
  p->last_launches = p->launches - l0;
  return FRB_OK;
}

extern "C" int32_t frb_last_timing(frb_prob_t p, float *ms, int64_t *kernel_launches) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_last_timing: prob is NULL");
  if (ms) *ms = p->last_ms;
  if (kernel_launches) *kernel_launches = p->last_launches;
  return FRB_OK;
}

extern "C" int32_t frb_set_kernel(frb_prob_t p, int32_t kind) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_set_kernel: prob is NULL");
  FRB_REQUIRE(kind >= FRB_KERNEL_AUTO && kind <= FRB_KERNEL_CURV_MARCH, FRB_ERR_ARG,
              "frb_set_kernel: unknown kernel kind");
  if (kind == FRB_KERNEL_CURV_MARCH)
    FRB_REQUIRE(p->kind == K_EULER2D && p->curv_iJ, FRB_ERR_STATE,
                "frb_set_kernel: the one-launch marching kernel belongs to curvilinear euler2d problems");
  if (kind == FRB_KERNEL_BGK_ONE_PASS)
    FRB_REQUIRE(p->kind == K_BGK1D, FRB_ERR_STATE, "frb_set_kernel: the one-pass kernel belongs to bgk1d problems");
  if (kind == FRB_KERNEL_RC)
    FRB_REQUIRE(p->rc_base != nullptr, FRB_ERR_STATE, "frb_set_kernel: row-chunk kernel needs euler2d, deg 2..3");
  if (kind == FRB_KERNEL_MARCH)
    FRB_REQUIRE(frb_euler2d_march_supported(p), FRB_ERR_STATE,
                "frb_set_kernel: marching kernel needs euler2d, deg 2..3 and even nx");
  p->kernel_kind = kind;
  return FRB_OK;
}

extern "C" int32_t frb_set_flux(frb_prob_t p, int32_t kind) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_set_flux: prob is NULL");
  FRB_REQUIRE(p->kind == K_EULER1D || p->kind == K_EULER2D, FRB_ERR_STATE, "frb_set_flux: Euler problems only");
  FRB_REQUIRE(kind >= FRB_FLUX_HLL && kind <= FRB_FLUX_ROE, FRB_ERR_ARG, "frb_set_flux: unknown flux kind");
  p->flux = kind;
  return FRB_OK;
}

extern "C" int32_t frb_set_profiling(frb_prob_t p, int32_t enabled) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_set_profiling: prob is NULL");
  p->profiling = enabled != 0;
  return FRB_OK;
}

extern "C" int32_t frb_stage_timing(frb_prob_t p, float *ms_total, int64_t *stage_launches) {
  FRB_REQUIRE(p, FRB_ERR_ARG, "frb_stage_timing: prob is NULL");
  if (ms_total) *ms_total = p->stage_ms;
  if (stage_launches) *stage_launches = p->stage_count;
  return FRB_OK;
}

extern "C" int32_t frb_host_alloc(int64_t bytes, void **ptr) {
  FRB_REQUIRE(ptr && bytes > 0, FRB_ERR_ARG, "frb_host_alloc: bad argument");
  FRB_CUDA(cudaMallocHost(ptr, (size_t)bytes));
  return FRB_OK;
}

extern "C" int32_t frb_host_free(void *ptr) {
  if (ptr) FRB_CUDA(cudaFreeHost(ptr));
  return FRB_OK;
}
