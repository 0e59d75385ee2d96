// Linear combination of stage derivatives for explicit Runge-Kutta schemes given by a Butcher
// tableau (frb_step_tableau: Tsit5 of example/advection_highlevel.jl:26 and
// example/euler1d_convergence.jl:133, RK4, ...):   out = u + dt * sum_j c[j] * k_j   over the whole
// state array, ghosts included (k_j = 0 there, so the ghosts stay frozen through the stages like
// the reference's du = 0).  Pure streaming: (n + 2) * 8 bytes per value.
#include "frb_internal.cuh"

namespace {

struct LinComb {
  const double *k[FRB_RK_MAX_STAGES];
  double c[FRB_RK_MAX_STAGES];
  int n;
};

__global__ void __launch_bounds__(256) lincomb_kernel(double *__restrict__ out, const double *__restrict__ u,
                                                      LinComb lc, double dt, size_t len) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += stride) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < FRB_RK_MAX_STAGES; ++j)
      if (j < lc.n) acc += lc.c[j] * lc.k[j][i];
    out[i] = u[i] + dt * acc;
  }
}

}  // namespace

// out = u + dt * sum_j c[j] k[j]; zero coefficients are skipped.  out may alias u.
int frb_launch_lincomb(frb_prob_t p, double *out, const double *u, int n, double *const *k, const double *c,
                       double dt) {
  LinComb lc;
  lc.n = 0;
  for (int j = 0; j < n; ++j) {
    if (c[j] == 0.0) continue;
    lc.k[lc.n] = k[j];
    lc.c[lc.n] = c[j];
    ++lc.n;
  }
  for (int j = lc.n; j < FRB_RK_MAX_STAGES; ++j) {
    lc.k[j] = nullptr;
    lc.c[j] = 0.0;
  }
  const size_t len = (size_t)p->len;
  const size_t want = (len + 255) / 256;
  const size_t cap = (size_t)p->ctx->sm_count * 8;
  const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
  lincomb_kernel<<<grid, 256, 0, p->ctx->stream>>>(out, u, lc, dt, len);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "lincomb_kernel", __FILE__, __LINE__);
  return 1;
}
