// 1-D BGK kinetic equation, fused residual + RK stage (config 4: 8192 cells x 256 velocities,
// data-parallel over phase space).  Reference: mol! of example/bgk_wave.jl:69-129 -- Maxwellian from
// the moments (:77-81), flux v_j u / (dx/2) (:85-88), interp_face! (:99-101), upwind interface flux
// through the periodic f2e / e2f tables (:42-67, :103-107), poly_derivative! (:114-116), correction
// + relaxation (M - u)/tau (:122-128).  State u[cell, velocity, sp], cell fastest.
//
// Two launches per stage:
//   bgk_moments_kernel  thread = (cell, sp): the three moments over the nu velocities in the
//                       reference's summation order, 16 loads in flight per thread (the sum is a
//                       serial chain, the loads are not); writes rho*sqrt(lambda/pi), U, lambda.
//   bgk1d_kernel        thread = (cell, velocity): own block + the upwind neighbour's block, 1 / J
//                       precomputed per cell, one exp per point.
// model FRB_BGK_KINETIC_ADVECTION is the mol! of example/advection_kinetic.jl:73-128: the same residual with the
// Maxwellian of prim = [rho, a, 1] (:80-88) -- only the epilogue of the moments kernel differs.
// Both are L2 / HBM streams of the 50 MB state; nothing here is GEMM-shaped.
#include "frb_internal.cuh"

namespace {

template <int NSP>
__device__ __forceinline__ double dotn(const double *a, const double *l) {
  double s = a[0] * l[0];
#pragma unroll
  for (int q = 1; q < NSP; ++q) s = fma(a[q], l[q], s);
  return s;
}

__device__ __forceinline__ double stage_out(const FrbStage &st, const double *ua, size_t idx, double u, double du) {
  if (st.rhs_only) return du;
  double r = st.nested ? st.cb * (u + st.cdt * du) : st.cb * u + st.cdt * du;
  if (st.use_a) r = st.ca * ua[idx] + r;
  return r;
}

// moments_conserve + conserve_prim(w, 3.0): bgk_wave.jl:77-79.  prim[3][nsp][ncell].
// Block = 32 cells (lanes, coalesced) x kMomGroups velocity groups (warps): each warp sums its
// contiguous slice of the velocity grid in order, the slices are combined in order through smem
// (deterministic; the reference's `sum` has no specified order either).
constexpr int kMomGroups = 8;
__global__ void __launch_bounds__(32 * kMomGroups)
bgk_moments_kernel(const double *__restrict__ u, double *__restrict__ prim, int ncell, int nu, int nsp,
                   const double *__restrict__ velo, const double *__restrict__ wts, int model, double a) {
  __shared__ double part[kMomGroups][3][32];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int k = blockIdx.y;
  const int per = (nu + kMomGroups - 1) / kMomGroups;
  const int j0 = grp * per, j1 = min(nu, j0 + per);
  double w0 = 0.0, w1 = 0.0, w2 = 0.0;
  if (i < ncell) {
    const double *p = u + i + (size_t)ncell * nu * k;
    int j = j0;
    for (; j + 16 <= j1; j += 16) {
      double f[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) f[q] = p[(size_t)ncell * (j + q)];
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const double v = velo[j + q], wt = wts[j + q];
        w0 += wt * f[q];
        w1 += wt * v * f[q];
        w2 += wt * (v * v) * f[q];
      }
    }
    for (; j < j1; ++j) {
      const double f = p[(size_t)ncell * j], v = velo[j], wt = wts[j];
      w0 += wt * f;
      w1 += wt * v * f;
      w2 += wt * (v * v) * f;
    }
  }
  part[grp][0][lane] = w0;
  part[grp][1][lane] = w1;
  part[grp][2][lane] = w2;
  __syncthreads();
  if (grp != 0 || i >= ncell) return;
  for (int g = 1; g < kMomGroups; ++g) {
    w0 += part[g][0][lane];
    w1 += part[g][1][lane];
    w2 += part[g][2][lane];
  }
  const size_t o = i + (size_t)ncell * k;
  if (model == FRB_BGK_KINETIC_ADVECTION) {
    // example/advection_kinetic.jl:80-88: rho = sum(u .* weights), prim = [rho, a, 1.0]
    prim[o] = w0 * sqrt(1.0 / 3.14159265358979323846);
    prim[o + (size_t)ncell * nsp] = a;
    prim[o + 2 * (size_t)ncell * nsp] = 1.0;
    return;
  }
  w2 *= 0.5;
  const double lam = 0.5 * w0 / (3.0 - 1.0) / (w2 - 0.5 * w1 * w1 / w0);
  prim[o] = w0 * sqrt(lam / 3.14159265358979323846);  // maxwellian prefactor rho*sqrt(lambda/pi)
  prim[o + (size_t)ncell * nsp] = w1 / w0;
  prim[o + 2 * (size_t)ncell * nsp] = lam;
}

template <int NSP>
__global__ void __launch_bounds__(128)
bgk1d_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
             const double *__restrict__ prim, const double *__restrict__ inv_j,
             const double *__restrict__ velo, int ncell, int nu, double inv_tau, FrbOps ops, FrbStage st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= ncell) return;
  const double v = velo[j];
  const bool pos = v >= 0.0;  // heaviside delta, bgk_wave.jl:26
  const size_t vs = (size_t)ncell * nu;
  // only the upwind neighbour contributes: left cell for v >= 0, right cell otherwise (periodic)
  const int in = pos ? (i == 0 ? ncell - 1 : i - 1) : (i == ncell - 1 ? 0 : i + 1);
  const double sc = v * inv_j[i], sn = v * inv_j[in];  // v / J
  double uc[NSP], f[NSP], fn[NSP];
#pragma unroll
  for (int q = 0; q < NSP; ++q) {
    uc[q] = u[i + (size_t)ncell * j + vs * q];
    f[q] = sc * uc[q];  // :85-88
    fn[q] = sn * u[in + (size_t)ncell * j + vs * q];
  }
  const double fL = dotn<NSP>(f, ops.ll), fR = dotn<NSP>(f, ops.lr);  // interp_face! :99-101
  // f_interaction at the left / right face of cell i (:103-107): own trace downwind, neighbour's upwind
  const double cl = pos ? dotn<NSP>(fn, ops.lr) - fL : 0.0;   // fi0 - fL
  const double cr = pos ? 0.0 : dotn<NSP>(fn, ops.ll) - fR;   // fi1 - fR
#pragma unroll
  for (int p = 0; p < NSP; ++p) {
    const size_t po = i + (size_t)ncell * p;
    const double pre = prim[po], U = prim[po + (size_t)ncell * NSP], lam = prim[po + 2 * (size_t)ncell * NSP];
    const double c = v - U;
    const double M = pre * exp(-lam * (c * c));  // maxwellian
    const double rhs1 = dotn<NSP>(f, &ops.lpdm[p * FRB_NSPMAX]);  // poly_derivative! :114-116
    const double du = -(rhs1 + cl * ops.dgl[p] + cr * ops.dgr[p]) + (M - uc[p]) * inv_tau;
    const size_t idx = i + (size_t)ncell * j + vs * p;
    out[idx] = stage_out(st, ua, idx, uc[p], du);
  }
}

}  // namespace

int frb_launch_bgk1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  dim3 blk(128), g1((p->ncell + 31) / 32, p->nsp), g2((p->ncell + 127) / 128, p->nu);
  bgk_moments_kernel<<<g1, 32 * kMomGroups, 0, p->ctx->stream>>>(u, p->prim, p->ncell, p->nu, p->nsp, p->velo, p->weights,
                                                                  p->bgk_model, p->a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "bgk_moments_kernel", __FILE__, __LINE__);
  const double it = 1.0 / p->tau;
  switch (p->nsp) {
    case 2: bgk1d_kernel<2><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 3: bgk1d_kernel<3><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 4: bgk1d_kernel<4><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 5: bgk1d_kernel<5><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 6: bgk1d_kernel<6><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 7: bgk1d_kernel<7><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    case 8: bgk1d_kernel<8><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    default: frb_set_error("bgk1d: deg must be in 1..7"); return FRB_ERR_ARG;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "bgk1d_kernel", __FILE__, __LINE__);
  return 2;
}
