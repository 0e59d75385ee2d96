// 1-D fused residual + RK-stage kernels: advection, Euler (HLL), BGK, positivity limiter.
// One kernel evaluates  out = ca*ua + cb*u + cdt*L(u)  for every cell: the element's
// nsp-point block lives in registers, the neighbour traces are recomputed from the
// neighbour's block (L1/L2 hits), so a stage is one pass over HBM.
//
// Reference semantics: src/Equation/eq_advection.jl:55-175 + eq_scalar.jl:1-12,
// example/advection_lowlevel.jl:4-47, src/Equation/eq_euler.jl:29-98,
// example/bgk_wave.jl:69-129, src/dissipation.jl:61-123.
#include <cstdlib>
#include <utility>

#include "frb_internal.cuh"
#include "frb_physics.cuh"

namespace {

// The 1-D problems are latency-bound (a few hundred KB of state), so these kernels keep
// the reference's literal operation order -- sequential dot products, true divisions --
// and this file is compiled with -fmad=false: the residual then agrees with the CPU
// restatement to the last bit or two, which matters for the ill-conditioned wave-speed
// quotient au = df/(du + 1e-8) of eq_advection.jl:130.
template <int NSP>
__device__ __forceinline__ double dotn(const double *a, const double *l) {
  double s = a[0] * l[0];
#pragma unroll
  for (int q = 1; q < NSP; ++q) s = s + a[q] * l[q];
  return s;
}

__device__ __forceinline__ double stage_out(const FrbStage &st, const double *ua, size_t idx,
                                            double u, double du) {
  if (st.rhs_only) return du;
  double r = st.nested ? st.cb * (u + st.cdt * du) : st.cb * u + st.cdt * du;
  if (st.use_a) r = st.ca * ua[idx] + r;
  return r;
}

// [KB] conserve_prim / euler_flux / flux_hll! in the reference's operation order
__device__ __forceinline__ double lambda3(double w0, double w1, double w2, double gamma) {
  return 0.5 * w0 / (gamma - 1.0) / (w2 - 0.5 * (w1 * w1) / w0);
}
__device__ __forceinline__ frb::Flux3 flux3_lit(double w0, double w1, double w2, double gamma) {
  double lam = lambda3(w0, w1, w2, gamma);
  double p = 0.5 * w0 / lam;
  return {w1, (w1 * w1) / w0 + p, (w2 + p) * w1 / w0};
}
__device__ __forceinline__ frb::Flux3 hll3_lit(double l0, double l1, double l2, double r0, double r1,
                                               double r2, double gamma) {
  double lamL = lambda3(l0, l1, l2, gamma), lamR = lambda3(r0, r1, r2, gamma);
  double aL = sqrt(0.5 * gamma / lamL), aR = sqrt(0.5 * gamma / lamR);
  double lmin = l1 / l0 - aL, lmax = r1 / r0 + aR;
  frb::Flux3 f1 = flux3_lit(l0, l1, l2, gamma), f2 = flux3_lit(r0, r1, r2, gamma);
  if (lmin >= 0.0) return f1;
  if (lmax <= 0.0) return f2;
  double fac = 1.0 / (lmax - lmin), mm = lmax * lmin;
  return {fac * (lmax * f1.f0 - lmin * f2.f0 + mm * (r0 - l0)),
          fac * (lmax * f1.f1 - lmin * f2.f1 + mm * (r1 - l1)),
          fac * (lmax * f1.f2 - lmin * f2.f2 + mm * (r2 - l2))};
}

// ---------------------------------------------------------------- advection ----
struct Adv1dArgs {
  const double *J;
  int ncell;
  double a;
  int bc;
  double eps_seam;
};

template <int NSP>
__device__ __forceinline__ void adv1d_cell(int i, const double *__restrict__ u, const double *__restrict__ ua,
                                           double *__restrict__ out, const Adv1dArgs &A, const FrbOps &ops,
                                           const FrbStage &st) {
  const double *J = A.J;
  const int ncell = A.ncell, bc = A.bc;
  const double a = A.a, eps_seam = A.eps_seam;
  const bool periodic = bc == FRB_BC_PERIOD;
  int im = i == 0 ? ncell - 1 : i - 1;
  int ip = i == ncell - 1 ? 0 : i + 1;
  double uc[NSP], f[NSP], um[NSP], up[NSP], fm[NSP], fp[NSP];
  double Jc = J[i], Jm = J[im], Jp = J[ip];
#pragma unroll
  for (int q = 0; q < NSP; ++q) {
    uc[q] = u[i + (size_t)ncell * q];
    um[q] = u[im + (size_t)ncell * q];
    up[q] = u[ip + (size_t)ncell * q];
    f[q] = a * uc[q] / Jc;   // advection_dflux!  eq_advection.jl:103-111
    fm[q] = a * um[q] / Jm;
    fp[q] = a * up[q] / Jp;
  }
  double uL = dotn<NSP>(uc, ops.ll), uR = dotn<NSP>(uc, ops.lr);
  double fL = dotn<NSP>(f, ops.ll), fR = dotn<NSP>(f, ops.lr);
  double uRm = dotn<NSP>(um, ops.lr), fRm = dotn<NSP>(fm, ops.lr);
  double uLp = dotn<NSP>(up, ops.ll), fLp = dotn<NSP>(fp, ops.ll);
  // interface fluxes: advection_iflux! :128-137, seam period_advection! :163-166
  double e0 = (i == 0) ? eps_seam : 1e-8;
  double e1 = (i == ncell - 1) ? eps_seam : 1e-8;
  double au0 = (fL - fRm) / (uL - uRm + e0);
  double fi0 = 0.5 * (fL + fRm) - 0.5 * fabs(au0) * (uL - uRm);
  double au1 = (fLp - fR) / (uLp - uR + e1);
  double fi1 = 0.5 * (fLp + fR) - 0.5 * fabs(au1) * (uLp - uR);
  const bool frozen = !periodic && (i == 0 || i == ncell - 1);  // dirichlet_advection! :153-156
#pragma unroll
  for (int p = 0; p < NSP; ++p) {
    double rhs1 = dotn<NSP>(f, &ops.lpdm[p * FRB_NSPMAX]);
    double du = -(rhs1 + (fi0 - fL) * ops.dgl[p] + (fi1 - fR) * ops.dgr[p]);  // eq_scalar.jl:4-8
    if (frozen) du = 0.0;
    size_t idx = i + (size_t)ncell * p;
    out[idx] = stage_out(st, ua, idx, uc[p], du);
  }
}

template <int NSP>
__global__ void __launch_bounds__(128)
adv1d_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out, Adv1dArgs A,
             FrbOps ops, FrbStage st) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.ncell) adv1d_cell<NSP>(i, u, ua, out, A, ops, st);
}

// -------------------------------------------------------------------- Euler ----
template <int NSP>
__device__ __forceinline__ void load_cell3(const double *__restrict__ u, int i, int ncell,
                                           double (&w)[3][NSP]) {
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int q = 0; q < NSP; ++q) w[k][q] = u[i + (size_t)ncell * (q + NSP * k)];
}

struct Euler1dArgs {
  const double *J;
  int ncell;
  double gamma;
  int bc, flux;
};

template <int NSP>
__device__ __forceinline__ void euler1d_cell(int i, const double *__restrict__ u, const double *__restrict__ ua,
                                             double *__restrict__ out, const Euler1dArgs &A, const FrbOps &ops,
                                             const FrbStage &st) {
  const double *J = A.J;
  const int ncell = A.ncell, bc = A.bc, flux = A.flux;
  const double gamma = A.gamma;
  const bool periodic = bc == FRB_BC_PERIOD;
  int im = i == 0 ? ncell - 1 : i - 1;
  int ip = i == ncell - 1 ? 0 : i + 1;
  double w[3][NSP], wm[3][NSP], wp[3][NSP], f[3][NSP];
  load_cell3<NSP>(u, i, ncell, w);
  load_cell3<NSP>(u, im, ncell, wm);
  load_cell3<NSP>(u, ip, ncell, wp);
  const double Jc = J[i];
#pragma unroll
  for (int q = 0; q < NSP; ++q) {  // eq_euler.jl:35-39
    frb::Flux3 F = flux3_lit(w[0][q], w[1][q], w[2][q], gamma);
    f[0][q] = F.f0 / Jc; f[1][q] = F.f1 / Jc; f[2][q] = F.f2 / Jc;
  }
  double uL[3], uR[3], fL[3], fR[3], uRm[3], uLp[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // :41-49
    uL[k] = dotn<NSP>(w[k], ops.ll); uR[k] = dotn<NSP>(w[k], ops.lr);
    fL[k] = dotn<NSP>(f[k], ops.ll); fR[k] = dotn<NSP>(f[k], ops.lr);
    uRm[k] = dotn<NSP>(wm[k], ops.lr); uLp[k] = dotn<NSP>(wp[k], ops.ll);
  }
  // :51-54 (interior faces) and :86-89 (periodic seam): HLL with dt = 1
  // (flux != HLL: the oracle-defined LF / Roe extras, frb_physics.cuh)
  frb::Flux3 h0 = flux == FRB_FLUX_HLL ? hll3_lit(uRm[0], uRm[1], uRm[2], uL[0], uL[1], uL[2], gamma)
                                       : frb::riemann3(flux, uRm[0], uRm[1], uRm[2], uL[0], uL[1], uL[2], gamma);
  frb::Flux3 h1 = flux == FRB_FLUX_HLL ? hll3_lit(uR[0], uR[1], uR[2], uLp[0], uLp[1], uLp[2], gamma)
                                       : frb::riemann3(flux, uR[0], uR[1], uR[2], uLp[0], uLp[1], uLp[2], gamma);
  double fi0[3] = {h0.f0, h0.f1, h0.f2}, fi1[3] = {h1.f0, h1.f1, h1.f2};
  const bool frozen = !periodic && (i == 0 || i == ncell - 1);  // dirichlet_euler! :75-78
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int p = 0; p < NSP; ++p) {
      double rhs1 = dotn<NSP>(f[k], &ops.lpdm[p * FRB_NSPMAX]);  // :56-58
      double du = -(rhs1 + (fi0[k] / Jc - fL[k]) * ops.dgl[p] + (fi1[k] / Jc - fR[k]) * ops.dgr[p]);
      if (frozen) du = 0.0;
      size_t idx = i + (size_t)ncell * (p + NSP * k);
      out[idx] = stage_out(st, ua, idx, w[k][p], du);
    }
}

template <int NSP>
__global__ void __launch_bounds__(128)
euler1d_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
               Euler1dArgs A, FrbOps ops, FrbStage st) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < A.ncell) euler1d_cell<NSP>(i, u, ua, out, A, ops, st);
}

// positive_limiter(u::Matrix, gamma, weights, ll, lr)  dissipation.jl:61-123, density branch.
template <int NSP>
__device__ __forceinline__ void limiter1d_cell(int i, double *__restrict__ u, int ncell, double gamma,
                                               const double *__restrict__ wts, const FrbOps &ops,
                                               int *__restrict__ nbad) {
  double w[3][NSP];
  load_cell3<NSP>(u, i, ncell, w);
  double um[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) s += w[k][q] * wts[q];  // :71 (mean not normalised)
    um[k] = s;
  }
  double lam = 0.5 * um[0] / (gamma - 1.0) / (um[2] - 0.5 * um[1] * um[1] / um[0]);
  double p_mean = 0.5 * um[0] * (1.0 / lam);
  double rl = dotn<NSP>(w[0], ops.ll), rr = dotn<NSP>(w[0], ops.lr);
  double eps = fmin(fmin(1e-13, um[0]), p_mean);  // :81
  double rmin = fmin(rl, rr);
#pragma unroll
  for (int q = 0; q < NSP; ++q) rmin = fmin(rmin, w[0][q]);
  double t1 = fmin((um[0] - eps) / (um[0] - rmin + 1e-8), 1.0);  // :83
  if (!(t1 > 0.0 && t1 <= 1.0)) atomicAdd(nbad, 1);              // :84 @assert
#pragma unroll
  for (int q = 0; q < NSP; ++q) u[i + (size_t)ncell * q] = t1 * (w[0][q] - um[0]) + um[0];
}

template <int NSP>
__global__ void __launch_bounds__(128)
limiter1d_kernel(double *__restrict__ u, int ncell, double gamma, const double *__restrict__ wts,
                 FrbOps ops, int *__restrict__ nbad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ncell) limiter1d_cell<NSP>(i, u, ncell, gamma, wts, ops, nbad);
}

// ---------------------------------------------------------------- whole time loop in one launch ----
// The smallest 1-D problems of the reference (cfg1: 100 cells, 2.4 KB) are pure launch latency: a stage
// runs for about a microsecond, a launch costs 5 us eagerly and ~3 us as a graph node.  Up to 512 cells
// the whole time loop runs in ONE CTA with a block barrier after every pass (cfg1, SSPRK3: 9.3 -> 2.9 us
// per step).  Same cell routines, same stage sequence and buffer roles as frb_step's host loop.
// Two multi-CTA forms exist behind environment switches, both measured slower than the CUDA-graph replay
// of the host loop at cfg2 (4096 cells, Midpoint + limiter: 26.7 us per step as a graph):
//   FRB_LOOP1D_CLUSTER  one thread-block cluster of up to 16 CTAs, hardware cluster barrier
//                       (barrier.cluster.arrive.release / wait.acquire) after every pass: 34.6 us per step
//                       with 8 CTAs x 512 cells, 32.2 us with 16 x 256 -- the barrier is cheap, but a pass
//                       is one long FP64 dependency chain per cell (IEEE divisions and square roots in the
//                       reference's operation order) and the 512-thread register budget (128, spills)
//                       lengthens it: ~11 us per pass against ~6 us for the 128-thread stage kernel;
//   FRB_LOOP1D_GRID     cooperative launch, grid barrier through a global counter: 28.0 us per step.
struct Loop1dArgs {
  double *u, *s1, *s2;       // state and stage buffers (roles as on the host)
  unsigned *bar;             // grid barrier: arrival counter
  int cluster;               // the grid is one cluster: hardware barrier
  int nsteps, scheme;
  double dt;
  const double *lim_w;       // limiter hook (Euler) or nullptr
  int *nbad;
};

// release on arrival, acquire on the poll: the CTA's writes (ordered before thread 0 by bar.sync) are
// visible to every thread that leaves the barrier; no stand-alone membar on either side
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned &round, int cluster) {
  if (cluster) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    return;
  }
  __syncthreads();
  if (gridDim.x > 1) {
    if (threadIdx.x == 0) {
      round += gridDim.x;
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      } while (v < round);
    }
    __syncthreads();
  }
}

template <int NSP, bool EULER>
__global__ void __launch_bounds__(512)
loop1d_kernel(Loop1dArgs L, Adv1dArgs AA, Euler1dArgs EA, FrbOps ops) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncell = EULER ? EA.ncell : AA.ncell;
  const bool live = i < ncell;
  unsigned round = 0;
  double *U = L.u, *S1 = L.s1, *S2 = L.s2;
  auto stage = [&](const double *in, const double *ua, double *out, FrbStage st) {
    if (live) {
      if (EULER) euler1d_cell<NSP>(i, in, ua, out, EA, ops, st);
      else adv1d_cell<NSP>(i, in, ua, out, AA, ops, st);
    }
    grid_barrier(L.bar, round, L.cluster);
  };
  for (int it = 0; it < L.nsteps; ++it) {
    if (EULER && L.lim_w) {
      if (live) limiter1d_cell<NSP>(i, U, ncell, EA.gamma, L.lim_w, ops, L.nbad);
      grid_barrier(L.bar, round, L.cluster);
    }
    const double dt = L.dt;
    if (L.scheme == FRB_SCHEME_EULER) {
      stage(U, nullptr, S1, FrbStage{0.0, 1.0, dt, 0, 0, 0});
      double *t = U; U = S1; S1 = t;
    } else if (L.scheme == FRB_SCHEME_MIDPOINT) {
      stage(U, nullptr, S1, FrbStage{0.0, 1.0, 0.5 * dt, 0, 0, 0});
      stage(S1, U, U, FrbStage{1.0, 0.0, dt, 1, 0, 0});
    } else {
      stage(U, nullptr, S1, FrbStage{0.0, 1.0, dt, 0, 0, 0});
      stage(S1, U, S2, FrbStage{0.75, 0.25, dt, 1, 0, 1});
      stage(S2, U, U, FrbStage{1.0 / 3.0, 2.0 / 3.0, dt, 1, 0, 1});
    }
  }
}

// dirichlet cells of the stage buffers are never written by L(u)=0 updates when the
// stage is rhs_only; nothing to do here (kept for symmetry with the 2-D ring copy).

// ---------------------------------------------------------------------- BGK ----
}  // namespace

#define FRB_NSP_SWITCH(nsp, CALL)                       \
  switch (nsp) {                                        \
    case 2: { constexpr int N = 2; CALL; } break;       \
    case 3: { constexpr int N = 3; CALL; } break;       \
    case 4: { constexpr int N = 4; CALL; } break;       \
    case 5: { constexpr int N = 5; CALL; } break;       \
    case 6: { constexpr int N = 6; CALL; } break;       \
    case 7: { constexpr int N = 7; CALL; } break;       \
    case 8: { constexpr int N = 8; CALL; } break;       \
    default: frb_set_error("unsupported polynomial degree"); return FRB_ERR_ARG; \
  }

static int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, what, __FILE__, __LINE__);
  return 0;
}

int frb_launch_adv1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  dim3 blk(128), grd((p->ncell + 127) / 128);
  double eps = p->variant == FRB_ADV_LOWLEVEL ? 1e-8 : 1e-6;
  int bc = p->variant == FRB_ADV_LOWLEVEL ? FRB_BC_PERIOD : p->bc;
  const Adv1dArgs A = {p->J, p->ncell, p->a, bc, eps};
  FRB_NSP_SWITCH(p->nsp, (adv1d_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(u, ua, out, A, p->ops, st)));
  if (int rc = check_launch("adv1d_kernel")) return rc;
  return 1;
}

int frb_launch_euler1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  dim3 blk(128), grd((p->ncell + 127) / 128);
  const Euler1dArgs A = {p->J, p->ncell, p->gamma, p->bc, p->flux};
  FRB_NSP_SWITCH(p->nsp, (euler1d_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(u, ua, out, A, p->ops, st)));
  if (int rc = check_launch("euler1d_kernel")) return rc;
  return 1;
}

int frb_launch_limiter1d(frb_prob_t p, double *u) {
  dim3 blk(128), grd((p->ncell + 127) / 128);
  FRB_NSP_SWITCH(p->nsp, (limiter1d_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(
                             u, p->ncell, p->gamma, p->lim_w, p->ops, p->flag)));
  if (int rc = check_launch("limiter1d_kernel")) return rc;
  return 1;
}


int frb_launch_dirichlet_copy1d(frb_prob_t, const double *, double *) { return 0; }

// The whole time loop of a small 1-D problem in one cooperative launch.  Returns 0 when the problem
// does not qualify (too many cells for one CTA per SM, or the device refuses a cooperative launch):
// the caller then runs the ordinary loop.  On success the state is in p->u with the host's buffer
// roles (the forward-Euler scheme swaps u and s1 every step).
int frb_launch_loop1d(frb_prob_t p, int scheme, double dt, int nsteps) {
  if (p->kind != K_ADV1D && p->kind != K_EULER1D) return 0;
  const bool grid_form = getenv("FRB_LOOP1D_GRID") != nullptr;
  const bool cluster_form = !grid_form && p->ncell > 512 && getenv("FRB_LOOP1D_CLUSTER") != nullptr;
  int threads = grid_form ? 128 : ((p->ncell + 31) / 32) * 32;
  int blocks = grid_form ? (p->ncell + 127) / 128 : 1;
  if (cluster_form) {
    const char *e = getenv("FRB_CLUSTER1D_CTAS");
    const int want = e ? atoi(e) : 16;
    blocks = (p->ncell + 511) / 512;
    if (blocks > 16) return 0;
    if (want > blocks && want <= 16) blocks = want;  // spread over more SMs: the passes are FP64 bound
    threads = (((p->ncell + blocks - 1) / blocks + 31) / 32) * 32;
  }
  if (threads > 512 || blocks > p->ctx->sm_count || nsteps <= 0) return 0;
  static int coop = -1;
  if (coop < 0) cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, p->ctx->device);
  if (!coop && !cluster_form) return 0;
  if (!p->loop_bar) FRB_CUDA(cudaMalloc(&p->loop_bar, sizeof(unsigned)));
  FRB_CUDA(cudaMemsetAsync(p->loop_bar, 0, sizeof(unsigned), p->ctx->stream));
  Loop1dArgs L = {p->u, p->s1, p->s2, p->loop_bar, cluster_form ? 1 : 0, nsteps, scheme, dt,
                  (p->kind == K_EULER1D && p->limiter_on) ? p->lim_w : nullptr, p->flag};
  const double eps = p->variant == FRB_ADV_LOWLEVEL ? 1e-8 : 1e-6;
  Adv1dArgs AA = {p->J, p->ncell, p->a, p->variant == FRB_ADV_LOWLEVEL ? FRB_BC_PERIOD : p->bc, eps};
  Euler1dArgs EA = {p->J, p->ncell, p->gamma, p->bc, p->flux};
  void *args[] = {&L, &AA, &EA, &p->ops};
  const void *fn = nullptr;
#define FRB_LOOP_CASE(N)                                                                        \
  case N: fn = p->kind == K_EULER1D ? (const void *)loop1d_kernel<N, true> : (const void *)loop1d_kernel<N, false>; break;
  switch (p->nsp) {
    FRB_LOOP_CASE(2) FRB_LOOP_CASE(3) FRB_LOOP_CASE(4) FRB_LOOP_CASE(5) FRB_LOOP_CASE(6) FRB_LOOP_CASE(7) FRB_LOOP_CASE(8)
    default: return 0;
  }
#undef FRB_LOOP_CASE
  cudaError_t e;
  if (cluster_form) {
    if (blocks > 8) {  // more than the portable cluster size: opt in, fall back if the device refuses
      if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
      }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(threads);
    cfg.stream = p->ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = blocks;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int fits = 0;  // can one such cluster be resident?
    if (cudaOccupancyMaxActiveClusters(&fits, fn, &cfg) != cudaSuccess || fits < 1) {
      (void)cudaGetLastError();
      return 0;
    }
    e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e == cudaErrorInvalidClusterSize || e == cudaErrorLaunchOutOfResources) {
      (void)cudaGetLastError();
      return 0;
    }
  } else {
    e = cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(threads), args, 0, p->ctx->stream);
  }
  if (e == cudaErrorCooperativeLaunchTooLarge) {
    (void)cudaGetLastError();
    return 0;
  }
  if (e != cudaSuccess) return frb_cuda_fail(e, "loop1d_kernel", __FILE__, __LINE__);
  if (scheme == FRB_SCHEME_EULER && (nsteps & 1)) std::swap(p->u, p->s1);
  return 1;
}
