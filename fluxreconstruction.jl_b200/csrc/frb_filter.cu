// Shock sensor + modal filter between steps: the per-element pass every shock case of
// the reference runs on the host between two step! calls --
//   1-D  example/euler_highlevel.jl:37-52      u_hat = iV * u[i,:,1];  su = u_hat[end]^2 / (sum(u_hat.^2) + 1e-6)
//   2-D  example/shock-vortex.jl:308-321       u_hat = iV * u[i,j,:,:,1][:];  su = u_hat[end]^2 / sum(u_hat.^2)
//   isShock = shock_detector(log10(su), deg, S0, kappa)                     src/dissipation.jl:13-23
//   if isShock: for every variable  u <- V * filter(iV * u)                 (KitBase modal_filter!)
// A modal filter is a diagonal damping of the modes, so V * filter(iV * u) = F * u with
// F = V * diag(sigma) * iV built on the host from whatever filter the caller picked (:l2, :exp, ...):
// the kernel needs iV (sensor) and F (np x np each, np = nsp or nsp^2), nothing of KitBase.
// One thread per element; state in the reference memory image.
#include "frb_internal.cuh"
#include "frb_rc.cuh"

namespace {

// dissipation.jl:13-23
__device__ __forceinline__ bool shock_detector(double Se, double S0, double kappa) {
  double sigma;
  if (Se < S0 - kappa) sigma = 1.0;
  else if (S0 - kappa <= Se && Se < S0 + kappa) sigma = 0.5 * (1.0 - sin(0.5 * 3.14159265358979323846 * (Se - S0) / kappa));
  else sigma = 0.0;
  return sigma < 0.99;
}

// One element: u[e + stride*q + vstride*var], q = 0..NP-1 (the element's points in Julia's [:]
// order); dup >= 0: offset of the element's copy in the neighbouring chunk (row-chunk layout).
template <int NP>
__device__ __forceinline__ void filter_element(double *__restrict__ u, long long e, long long dup, long long stride,
                                               long long vstride, int nvar, const double *__restrict__ iV,
                                               const double *__restrict__ F, double eps, double S0, double kappa,
                                               int *__restrict__ count) {
  double w[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) w[q] = u[e + stride * q];
  double last = 0.0, sum = 0.0;
#pragma unroll
  for (int m = 0; m < NP; ++m) {  // u_hat = iV * rho  (iV column-major: iV[m + NP*q])
    double a = 0.0;
#pragma unroll
    for (int q = 0; q < NP; ++q) a += iV[m + NP * q] * w[q];
    sum += a * a;
    last = a;
  }
  const double su = (last * last) / (sum + eps);
  if (!shock_detector(log10(su), S0, kappa)) return;
  if (count) atomicAdd(count, 1);
  for (int s = 0; s < nvar; ++s) {
    double *p = u + e + vstride * s;
    if (s > 0) {
#pragma unroll
      for (int q = 0; q < NP; ++q) w[q] = p[stride * q];
    }
    double r[NP];
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < NP; ++q) a += F[m + NP * q] * w[q];
      r[m] = a;
    }
#pragma unroll
    for (int m = 0; m < NP; ++m) {
      p[stride * m] = r[m];
      if (dup >= 0) u[dup + vstride * s + stride * m] = r[m];
    }
  }
}

template <int NP>
__global__ void __launch_bounds__(128)
modal_filter_kernel(double *__restrict__ u, long long nelem, int row, long long pitch, long long e0,
                    long long stride, long long vstride, int nvar, const double *__restrict__ iV,
                    const double *__restrict__ F, double eps, double S0, double kappa, int *__restrict__ count) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= nelem) return;
  // elements are enumerated row by row: `row` consecutive ones, then a jump of `pitch`
  filter_element<NP>(u, e0 + (t % row) + (t / row) * pitch, -1, stride, vstride, nvar, iV, F, eps, S0, kappa, count);
}

// the same pass on the row-chunk layout (frb_rc.cuh): thread = (lane, strip, row); every copy of a
// column is filtered by the thread that owns it (ghost columns: the lane that holds them)
template <int NP>
__global__ void __launch_bounds__(128)
modal_filter_rc_kernel(double *__restrict__ u, RcGeom g, int ghosts, const double *__restrict__ iV,
                       const double *__restrict__ F, double eps, double S0, double kappa) {
  const int lane = threadIdx.x, s = blockIdx.x * blockDim.y + threadIdx.y;
  const int j = blockIdx.y + (ghosts ? 0 : 1);
  if (s >= g.ns) return;
  const int i = kRcOwn * s + lane;
  if (i > g.nx + 1) return;
  int sp, lp;
  rc_primary(g, i, &sp, &lp);
  if (sp != s) return;                                    // a copy: its owner filters it
  if (!ghosts && (i == 0 || i == g.nx + 1)) return;
  int s2, l2;
  long long dup = -1;
  if (rc_duplicate(g, s, lane, &s2, &l2)) dup = (long long)rc_index(g, j, s2, 0, l2);
  filter_element<NP>(u, (long long)rc_index(g, j, s, 0, lane), dup, 32, 32LL * NP, 4, iV, F, eps, S0, kappa, nullptr);
}

}  // namespace

// iV_dev, F_dev: np*np doubles each on the device.  include_ghosts: 2-D loops of the reference run
// over axes(itg.u, 1:2), ghost cells included (shock-vortex.jl:309).
int frb_launch_modal_filter(frb_prob_t p, double *u, const double *iV_dev, const double *F_dev, double eps,
                            double S0, double kappa, bool include_ghosts, int *count) {
  long long nelem, pitch, e0, stride, vstride;
  int row, nvar, np;
  if (p->kind == K_EULER1D) {  // u[cell, sp, var]
    np = p->nsp; nvar = 3; row = p->ncell; nelem = p->ncell; pitch = 0; e0 = 0;
    stride = p->ncell; vstride = (long long)p->ncell * p->nsp;
  } else if (p->kind == K_EULER2D) {  // u[i, j, k, l, m]; [:] of u[i,j,:,:,m] runs k fastest
    const long long NXG = p->nx + 2, NYG = p->ny + 2;
    np = p->nsp * p->nsp; nvar = 4;
    if (include_ghosts) { row = (int)NXG; nelem = NXG * NYG; e0 = 0; }
    else { row = p->nx; nelem = (long long)p->nx * p->ny; e0 = 1 + NXG; }
    pitch = NXG; stride = NXG * NYG; vstride = NXG * NYG * np;
  } else {
    frb_set_error("modal filter: Euler problems only");
    return FRB_ERR_STATE;
  }
  dim3 blk(128), grd((unsigned)((nelem + 127) / 128));
  cudaStream_t s = p->ctx->stream;
#define FRB_FILTER_CASE(N)                                                                               \
  case N:                                                                                                \
    modal_filter_kernel<N><<<grd, blk, 0, s>>>(u, nelem, row, pitch, e0, stride, vstride, nvar, iV_dev, \
                                               F_dev, eps, S0, kappa, count);                          \
    break;
  switch (np) {
    FRB_FILTER_CASE(2) FRB_FILTER_CASE(3) FRB_FILTER_CASE(4) FRB_FILTER_CASE(5) FRB_FILTER_CASE(6)
    FRB_FILTER_CASE(7) FRB_FILTER_CASE(8) FRB_FILTER_CASE(9) FRB_FILTER_CASE(16) FRB_FILTER_CASE(25)
    FRB_FILTER_CASE(36)
    default: frb_set_error("modal filter: unsupported number of points per element"); return FRB_ERR_ARG;
  }
#undef FRB_FILTER_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "modal_filter_kernel", __FILE__, __LINE__);
  return 1;
}

// the pass on the row-chunk mirror of a 2-D Euler problem (hook of frb_step)
int frb_rc_modal_filter(frb_prob_t p, double *u, const double *iV_dev, const double *F_dev, double eps, double S0,
                        double kappa, bool include_ghosts) {
  const RcGeom g = rc_geom(p->nx, p->ny, p->nsp);
  dim3 blk(32, 4), grd((g.ns + 3) / 4, include_ghosts ? p->ny + 2 : p->ny);
  cudaStream_t s = p->ctx->stream;
  switch (p->nsp * p->nsp) {
    case 9: modal_filter_rc_kernel<9><<<grd, blk, 0, s>>>(u, g, include_ghosts, iV_dev, F_dev, eps, S0, kappa); break;
    case 16: modal_filter_rc_kernel<16><<<grd, blk, 0, s>>>(u, g, include_ghosts, iV_dev, F_dev, eps, S0, kappa); break;
    default: frb_set_error("row-chunk modal filter: deg 2..3"); return FRB_ERR_ARG;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "modal_filter_rc_kernel", __FILE__, __LINE__);
  return 1;
}
