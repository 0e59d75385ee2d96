// 2-D Euler on TRIANGLES (UnstructFRPSpace), fused residual + RK stage: dudt! of dev/sod.jl:31-123
// (the same loop is dev/euler.jl and dev/euler_naca.jl).  What the reference passes in `p` comes in
// through the ABI unchanged -- cellType, J, lf, cell_normal, fpn, dl, phi (struct.jl:305-352,
// geo_jacobi.jl:32-43, geo_neighbor.jl:8-60) -- so TriFRPSpace stays the single source of truth.
// State u[ncell, np, 4], cell fastest (dev/sod.jl:19): one thread per cell, every access coalesced.
//
//   tri_trace_kernel   u_face[i, j, k, :] = sum_p u[i, p, :] lf[j, k, p]             (:48-54) -> HBM
//   tri_rhs_kernel     point fluxes / J, flux traces, the common flux on the cell's three faces over
//                      the flux-point connectivity fpn (gather of the neighbour's trace, HLL in the
//                      face frame, :62-99; wall cells mirror the tangential... momentum as :79-82 do),
//                      -f . dl - (fhat - fn) . phi (:101-115), cells of type 1 frozen, RK stage.
// Common fluxes are evaluated by both neighbours (a gather kernel has no owner of a face); the
// meshes of the reference's triangle cases are 1e3..1e5 cells, i.e. latency- not bandwidth-bound.
#include "frb_internal.cuh"
#include "frb_physics.cuh"

namespace {

struct TriOps {  // device pointers, Julia layouts
  const double *lf;   // [3, deg+1, np]
  const double *dl;   // [np, np, 2]
  const double *phi;  // [3, deg+1, np]
};

template <int DEG>
__global__ void __launch_bounds__(128)
tri_trace_kernel(const double *__restrict__ u, double *__restrict__ uf, int ncell, TriOps ops) {
  constexpr int NP = (DEG + 1) * (DEG + 2) / 2, NF = DEG + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  double w[NP][4];
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int p = 0; p < NP; ++p) w[p][l] = u[i + (size_t)ncell * (p + NP * l)];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int k = 0; k < NF; ++k)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double a = 0.0;
#pragma unroll
        for (int p = 0; p < NP; ++p) a += w[p][l] * ops.lf[j + 3 * (k + NF * p)];
        uf[i + (size_t)ncell * (j + 3 * (k + NF * l))] = a;
      }
}

template <int DEG>
__global__ void __launch_bounds__(128)
tri_rhs_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
               const double *__restrict__ uf, int ncell, const int *__restrict__ cell_type,
               const double *__restrict__ J, const double *__restrict__ normals, const int *__restrict__ fpn,
               TriOps ops, double gamma, FrbStage st) {
  constexpr int NP = (DEG + 1) * (DEG + 2) / 2, NF = DEG + 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const double gm1 = gamma - 1.0;
  // inv(J[i]),  J = [xr xs; yr ys]  stored [ncell, 2, 2] column-major
  const double xr = J[i], yr = J[i + (size_t)ncell], xs = J[i + (size_t)ncell * 2], ys = J[i + (size_t)ncell * 3];
  const double idet = 1.0 / (xr * ys - xs * yr);
  const double i00 = ys * idet, i01 = -xs * idet, i10 = -yr * idet, i11 = xr * idet;
  const int type = cell_type[i];
  double w[NP][4], f[NP][4][2];
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int p = 0; p < NP; ++p) w[p][l] = u[i + (size_t)ncell * (p + NP * l)];
#pragma unroll
  for (int p = 0; p < NP; ++p) {  // :37-46
    frb::Flux4 F, G;
    frb::euler_flux4(w[p][0], w[p][1], w[p][2], w[p][3], gm1, F, G);
    const double Fv[4] = {F.f0, F.f1, F.f2, F.f3}, Gv[4] = {G.f0, G.f1, G.f2, G.f3};
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      f[p][l][0] = i00 * Fv[l] + i01 * Gv[l];
      f[p][l][1] = i10 * Fv[l] + i11 * Gv[l];
    }
  }
  double du[NP][4];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int l = 0; l < 4; ++l) {  // rhs1 :101-105
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < NP; ++q) a += f[q][l][0] * ops.dl[p + NP * q] + f[q][l][1] * ops.dl[p + NP * (q + NP)];
      du[p][l] = -a;
    }
  const double nref[3][2] = {{0.0, -1.0}, {0.70710678118654757, 0.70710678118654757}, {-1.0, 0.0}};  // :56
  for (int j = 0; j < 3; ++j) {
    const double c = normals[i + (size_t)ncell * j], s = normals[i + (size_t)ncell * (j + 3)];
    for (int k = 0; k < NF; ++k) {
      double uo[4], fn[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double a = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const double lfp = ops.lf[j + 3 * (k + NF * p)];
          a += w[p][l] * lfp;
          b0 += f[p][l][0] * lfp;
          b1 += f[p][l][1] * lfp;
        }
        uo[l] = a;
        fn[l] = b0 * nref[j][0] + b1 * nref[j][1];  // fn_face :58-61
      }
      // fpn[c + 3*(i + ncell*(j + 3*k))], 1-based as the reference stores it; <= 0: no neighbour
      const size_t fo = 3 * ((size_t)i + (size_t)ncell * (j + 3 * k));
      const int ni = fpn[fo] - 1, nj = fpn[fo + 1] - 1, nk = fpn[fo + 2] - 1;
      double fl[4] = {0.0, 0.0, 0.0, 0.0};
      if (ni >= 0 || type == 2) {
        double un[4];
        if (ni >= 0) {
#pragma unroll
          for (int l = 0; l < 4; ++l) un[l] = uf[ni + (size_t)ncell * (nj + 3 * (nk + NF * l))];
        } else {  // :79-82
          un[0] = uo[0]; un[1] = uo[1]; un[2] = -uo[2]; un[3] = uo[3];
        }
        // local_frame(w, c, s) = (w0, w1 c + w2 s, w2 c - w1 s, w3)
        const frb::Flux4 h = frb::hll4(uo[0], uo[1] * c + uo[2] * s, uo[2] * c - uo[1] * s, uo[3], un[0],
                                       un[1] * c + un[2] * s, un[2] * c - un[1] * s, un[3], gamma);
        fl[0] = h.f0; fl[1] = h.f1 * c - h.f2 * s; fl[2] = h.f1 * s + h.f2 * c; fl[3] = h.f3;  // global_frame
      }
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const double fx = fl[l] * c, fy = fl[l] * s;                          // fwn_xy :86-89
        const double fr = i00 * fx + i01 * fy, fs = i10 * fx + i11 * fy;      // inv(J) * .  :91-94
        const double d = (fr * nref[j][0] + fs * nref[j][1]) - fn[l];         // fwns - fn_face
#pragma unroll
        for (int p = 0; p < NP; ++p) du[p][l] -= d * ops.phi[j + 3 * (k + NF * p)];  // rhs2 :107-115
      }
    }
  }
  const bool active = type == 0 || type == 2;
#pragma unroll
  for (int l = 0; l < 4; ++l)
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const size_t idx = i + (size_t)ncell * (p + NP * l);
      const double d = active ? du[p][l] : 0.0;
      double r;
      if (st.rhs_only) r = d;
      else {
        r = st.nested ? st.cb * (w[p][l] + st.cdt * d) : st.cb * w[p][l] + st.cdt * d;
        if (st.use_a) r = st.ca * ua[idx] + r;
      }
      out[idx] = r;
    }
}

}  // namespace

int frb_launch_tri_euler(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  const int np = (p->nsp) * (p->nsp + 1) / 2, nf = p->nsp;  // nsp = deg + 1
  TriOps ops = {p->tri_ops, p->tri_ops + 3 * nf * np, p->tri_ops + 3 * nf * np + 2 * np * np};
  dim3 blk(128), grd((p->ncell + 127) / 128);
  cudaStream_t s = p->ctx->stream;
#define FRB_TRI_CASE(D)                                                                                       \
  case D:                                                                                                     \
    tri_trace_kernel<D><<<grd, blk, 0, s>>>(u, p->tri_uf, p->ncell, ops);                                     \
    tri_rhs_kernel<D><<<grd, blk, 0, s>>>(u, ua, out, p->tri_uf, p->ncell, p->tri_type, p->J, p->tri_normals, \
                                          p->tri_fpn, ops, p->gamma, st);                                     \
    break;
  switch (p->nsp - 1) {
    FRB_TRI_CASE(1) FRB_TRI_CASE(2) FRB_TRI_CASE(3)
    default: frb_set_error("triangle kernels support deg 1..3"); return FRB_ERR_ARG;
  }
#undef FRB_TRI_CASE
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "tri_euler kernels", __FILE__, __LINE__);
  return 2;
}
