// 2-D Euler on the structured quadrilateral FRPSpace2D: generic per-element fused kernel
// (any degree), ghost-ring utilities and the positivity limiter.
//
// Reference semantics: dudt! of example/euler2d_wave.jl:35-107 (== shock-vortex.jl:26-118),
// the ghost fill of euler2d_wave.jl:127-132,159-164 / shock-vortex.jl:324-326 and
// positive_limiter(::AbstractArray{T,3}) of src/dissipation.jl:125-206.
//
// State layout (Julia column-major, one ghost ring): u[i + NXG*j + NE*(k + NSP*(l + NSP*m))]
// with i in 0..nx+1 fastest, k <-> r (x), l <-> s (y), m = variable.  Every (k,l,m) is a
// contiguous "plane" of NE = (nx+2)(ny+2) doubles, so a warp that maps lanes to consecutive
// i reads 256 contiguous bytes per plane.
#include "frb_internal.cuh"
#include "frb_rc.cuh"
#include "frb_physics.cuh"

namespace {

template <int NSP>
__device__ __forceinline__ size_t pidx(int k, int l, int m) {
  return (size_t)(k + NSP * (l + NSP * m));
}

// The generic kernel: one thread per interior element, everything fused; neighbour traces
// are recomputed from the neighbour blocks.  Correct for every degree; it is the
// fallback for deg != 3 and the on-device cross-check of the marching kernel.
template <int NSP>
__global__ void __launch_bounds__(128)
euler2d_generic_kernel(const double *__restrict__ u, const double *__restrict__ ua,
                       double *__restrict__ out, int nx, int ny, double iJx, double iJy,
                       double gamma, int flux, FrbOps ops, FrbStage st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > nx || j > ny) return;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  const size_t e = i + NXG * j;
  const double gm1 = gamma - 1.0;

  double w[4][NSP][NSP];   // [m][l][k]
  double f1[4][NSP][NSP];  // iJx * F
  double f2[4][NSP][NSP];  // iJy * G
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int l = 0; l < NSP; ++l)
#pragma unroll
      for (int k = 0; k < NSP; ++k) w[m][l][k] = u[e + NE * pidx<NSP>(k, l, m)];
#pragma unroll
  for (int l = 0; l < NSP; ++l)
#pragma unroll
    for (int k = 0; k < NSP; ++k) {  // euler2d_wave.jl:45-50
      frb::Flux4 F, G;
      frb::euler_flux4(w[0][l][k], w[1][l][k], w[2][l][k], w[3][l][k], gm1, F, G);
      f1[0][l][k] = F.f0 * iJx; f1[1][l][k] = F.f1 * iJx; f1[2][l][k] = F.f2 * iJx; f1[3][l][k] = F.f3 * iJx;
      f2[0][l][k] = G.f0 * iJy; f2[1][l][k] = G.f1 * iJy; f2[2][l][k] = G.f2 * iJy; f2[3][l][k] = G.f3 * iJy;
    }

  double du[4][NSP][NSP];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int l = 0; l < NSP; ++l)
#pragma unroll
      for (int k = 0; k < NSP; ++k) {  // :84-91
        double a = f1[m][l][0] * ops.lpdm[k * FRB_NSPMAX];
        double b = f2[m][0][k] * ops.lpdm[l * FRB_NSPMAX];
#pragma unroll
        for (int q = 1; q < NSP; ++q) {
          a = fma(f1[m][l][q], ops.lpdm[k * FRB_NSPMAX + q], a);
          b = fma(f2[m][q][k], ops.lpdm[l * FRB_NSPMAX + q], b);
        }
        du[m][l][k] = a + b;
      }

  // x faces (:68-74): row l of the left (face 4) and right (face 2) edges
#pragma unroll
  for (int l = 0; l < NSP; ++l) {
    double uL[4], uR[4], fL[4], fR[4], nL[4], nR[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double a = 0, b = 0, c = 0, d = 0, g = 0, h = 0;
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        a = fma(w[m][l][q], ops.ll[q], a);
        b = fma(w[m][l][q], ops.lr[q], b);
        c = fma(f1[m][l][q], ops.ll[q], c);
        d = fma(f1[m][l][q], ops.lr[q], d);
        g = fma(u[e - 1 + NE * pidx<NSP>(q, l, m)], ops.lr[q], g);  // u_face[i-1,j,2,l,m]
        h = fma(u[e + 1 + NE * pidx<NSP>(q, l, m)], ops.ll[q], h);  // u_face[i+1,j,4,l,m]
      }
      uL[m] = a; uR[m] = b; fL[m] = c; fR[m] = d; nL[m] = g; nR[m] = h;
    }
    frb::Flux4 hl = frb::riemann4(flux, nL[0], nL[1], nL[2], nL[3], uL[0], uL[1], uL[2], uL[3], gamma);
    frb::Flux4 hr = frb::riemann4(flux, uR[0], uR[1], uR[2], uR[3], nR[0], nR[1], nR[2], nR[3], gamma);
    const double cl[4] = {hl.f0 * iJx - fL[0], hl.f1 * iJx - fL[1], hl.f2 * iJx - fL[2], hl.f3 * iJx - fL[3]};
    const double cr[4] = {hr.f0 * iJx - fR[0], hr.f1 * iJx - fR[1], hr.f2 * iJx - fR[2], hr.f3 * iJx - fR[3]};
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int k = 0; k < NSP; ++k)
        du[m][l][k] += cl[m] * ops.dgl[k] + cr[m] * ops.dgr[k];  // :96-99
  }
  // y faces (:75-82): column k of the bottom (face 1) and top (face 3) edges
#pragma unroll
  for (int k = 0; k < NSP; ++k) {
    double uB[4], uT[4], gB[4], gT[4], nB[4], nT[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      double a = 0, b = 0, c = 0, d = 0, g = 0, h = 0;
#pragma unroll
      for (int q = 0; q < NSP; ++q) {
        a = fma(w[m][q][k], ops.ll[q], a);
        b = fma(w[m][q][k], ops.lr[q], b);
        c = fma(f2[m][q][k], ops.ll[q], c);
        d = fma(f2[m][q][k], ops.lr[q], d);
        g = fma(u[e - NXG + NE * pidx<NSP>(k, q, m)], ops.lr[q], g);  // u_face[i,j-1,3,k,m]
        h = fma(u[e + NXG + NE * pidx<NSP>(k, q, m)], ops.ll[q], h);  // u_face[i,j+1,1,k,m]
      }
      uB[m] = a; uT[m] = b; gB[m] = c; gT[m] = d; nB[m] = g; nT[m] = h;
    }
    frb::Flux4 hb = frb::riemann4_y(flux, nB[0], nB[1], nB[2], nB[3], uB[0], uB[1], uB[2], uB[3], gamma);
    frb::Flux4 ht = frb::riemann4_y(flux, uT[0], uT[1], uT[2], uT[3], nT[0], nT[1], nT[2], nT[3], gamma);
    const double cb[4] = {hb.f0 * iJy - gB[0], hb.f1 * iJy - gB[1], hb.f2 * iJy - gB[2], hb.f3 * iJy - gB[3]};
    const double ct[4] = {ht.f0 * iJy - gT[0], ht.f1 * iJy - gT[1], ht.f2 * iJy - gT[2], ht.f3 * iJy - gT[3]};
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int l = 0; l < NSP; ++l)
        du[m][l][k] += cb[m] * ops.dgl[l] + ct[m] * ops.dgr[l];  // :100-103
  }

#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int l = 0; l < NSP; ++l)
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        size_t idx = e + NE * pidx<NSP>(k, l, m);
        double d = -du[m][l][k];
        double r;
        if (st.rhs_only) r = d;
        else {
          r = fma(st.cdt, d, st.cb * w[m][l][k]);
          if (st.use_a) r = fma(st.ca, ua[idx], r);
        }
        out[idx] = r;
      }
}

// ---- ghost ring utilities ---------------------------------------------------------
// x pass: u[0,j,p] = sL * u[srcL,j,p], u[nx+1,j,p] = sR * u[srcR,j,p] for all j, planes.
__global__ void ghost_x_kernel(double *__restrict__ u, int nx, int ny, int nplanes, int npp,
                               int srcL, int srcR, int flip_var, int do_left) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int p = blockIdx.y;
  if (j > ny + 1 || p >= nplanes) return;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  double s = (p / npp == flip_var) ? -1.0 : 1.0;
  double *row = u + NXG * j + NE * p;
  if (do_left) row[0] = s * row[srcL];
  row[nx + 1] = s * row[srcR];
}
// y pass: u[i,0,p] = s*u[i,srcB,p], u[i,ny+1,p] = s*u[i,srcT,p] for all i.
__global__ void ghost_y_kernel(double *__restrict__ u, int nx, int ny, int nplanes, int npp,
                               int srcB, int srcT, int flip_var) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int p = blockIdx.y;
  if (i > nx + 1 || p >= nplanes) return;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  double s = (p / npp == flip_var) ? -1.0 : 1.0;
  double *pl = u + NE * p;
  pl[i] = s * pl[i + NXG * srcB];
  pl[i + NXG * (ny + 1)] = s * pl[i + NXG * srcT];
}
// copy (or zero) the ghost ring of every plane
__global__ void ring_copy_kernel(const double *__restrict__ src, double *__restrict__ dst, int nx,
                                 int ny, int nplanes, int row0, int rowN) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int p = blockIdx.y;
  const int NXG = nx + 2, NYG = ny + 2;
  const int nring = 2 * NXG + 2 * ny;
  if (t >= nring || p >= nplanes) return;
  int i, j;
  if (t < NXG) { i = t; j = 0; if (!row0) return; }
  else if (t < 2 * NXG) { i = t - NXG; j = NYG - 1; if (!rowN) return; }
  else if (t < 2 * NXG + ny) { i = 0; j = t - 2 * NXG + 1; }
  else { i = NXG - 1; j = t - 2 * NXG - ny + 1; }
  size_t idx = i + (size_t)NXG * j + (size_t)NXG * NYG * p;
  dst[idx] = src ? src[idx] : 0.0;
}

// ghost cells of rows ra..rb of every plane (rows 0 / ny+1 entirely, otherwise columns 0 and nx+1): src -> up to three
// buffers.  The slab-wise form of ring_copy_kernel for the pipelined host step (frb_step_host).
__global__ void ring_rows_kernel(const double *__restrict__ src, double *__restrict__ d1, double *__restrict__ d2,
                                 double *__restrict__ d3, int nx, int ny, int ra) {
  const int r = ra + blockIdx.x, p = blockIdx.y;
  const int NXG = nx + 2;
  const size_t base = (size_t)NXG * r + (size_t)NXG * (ny + 2) * p;
  if (r == 0 || r == ny + 1) {
    for (int i = threadIdx.x; i < NXG; i += blockDim.x) {
      const double v = src[base + i];
      if (d1) d1[base + i] = v;
      if (d2) d2[base + i] = v;
      if (d3) d3[base + i] = v;
    }
  } else if (threadIdx.x < 2) {
    const int i = threadIdx.x == 0 ? 0 : NXG - 1;
    const double v = src[base + i];
    if (d1) d1[base + i] = v;
    if (d2) d2[base + i] = v;
    if (d3) d3[base + i] = v;
  }
}

// positive_limiter(u[nsp,nsp,4], gamma, weights, ll, lr): dissipation.jl:125-206, density branch
// One element at offset e with plane stride NE (reference image: NE = (nx+2)(ny+2); row-chunk
// layout: NE = 32); dup >= 0 is the offset of the element's copy in the neighbouring chunk.
template <int NSP>
__device__ __forceinline__ void limit_element(double *__restrict__ u, size_t e, size_t NE, long long dup,
                                              double gamma, const double *__restrict__ wts,
                                              const FrbOps &ops, int *__restrict__ nbad) {
  double rho[NSP][NSP];  // [l][k]
  double um[4];
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    double s = 0.0;
#pragma unroll
    for (int l = 0; l < NSP; ++l)
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double v = u[e + NE * pidx<NSP>(k, l, m)];
        if (m == 0) rho[l][k] = v;
        s += v * wts[k + NSP * l];  // :135 (one mean per variable)
      }
    um[m] = s;
  }
  double lam = 0.5 * um[0] / (gamma - 1.0) / (um[3] - 0.5 * (um[1] * um[1] + um[2] * um[2]) / um[0]);
  double p_mean = 0.5 * um[0] * (1.0 / lam);
  double eps = fmin(fmin(1e-13, um[0]), p_mean);  // :165
  double rmin = INFINITY;
#pragma unroll
  for (int a = 0; a < NSP; ++a) {
    double b1 = 0, b2 = 0, b3 = 0, b4 = 0;
#pragma unroll
    for (int q = 0; q < NSP; ++q) {
      // u[k=a, l=q] traces along s (faces 1,3); u[k=q, l=a] traces along r (faces 2,4)
      b1 = fma(rho[q][a], ops.ll[q], b1);
      b3 = fma(rho[q][a], ops.lr[q], b3);
      b2 = fma(rho[a][q], ops.lr[q], b2);
      b4 = fma(rho[a][q], ops.ll[q], b4);
      rmin = fmin(rmin, rho[a][q]);
    }
    rmin = fmin(rmin, fmin(fmin(b1, b2), fmin(b3, b4)));
  }
  double t1 = fmin((um[0] - eps) / (um[0] - rmin + 1e-8), 1.0);  // :167
  if (!(t1 > 0.0 && t1 <= 1.0)) atomicAdd(nbad, 1);
#pragma unroll
  for (int l = 0; l < NSP; ++l)
#pragma unroll
    for (int k = 0; k < NSP; ++k)
    {
      const double v = t1 * (rho[l][k] - um[0]) + um[0];
      u[e + NE * pidx<NSP>(k, l, 0)] = v;
      if (dup >= 0) u[(size_t)dup + NE * pidx<NSP>(k, l, 0)] = v;
    }
}

template <int NSP>
__global__ void __launch_bounds__(128)
limiter2d_kernel(double *__restrict__ u, int nx, int ny, double gamma,
                 const double *__restrict__ wts, FrbOps ops, int *__restrict__ nbad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > nx || j > ny) return;
  const size_t NXG = nx + 2, NE = NXG * (size_t)(ny + 2);
  limit_element<NSP>(u, i + NXG * j, NE, -1, gamma, wts, ops, nbad);
}

// the same limiter on the row-chunk layout (frb_rc.cuh): thread = (lane, strip, row)
template <int NSP>
__global__ void __launch_bounds__(128)
limiter2d_rc_kernel(double *__restrict__ u, RcGeom g, double gamma, const double *__restrict__ wts,
                    FrbOps ops, int *__restrict__ nbad) {
  const int lane = threadIdx.x, s = blockIdx.x * blockDim.y + threadIdx.y, j = blockIdx.y + 1;
  const int i = kRcOwn * s + lane;
  if (s >= g.ns || lane < 1 || lane > kRcOwn || i > g.nx) return;
  int s2, l2;
  long long dup = -1;
  if (rc_duplicate(g, s, lane, &s2, &l2)) dup = (long long)rc_index(g, j, s2, 0, l2);
  limit_element<NSP>(u, rc_index(g, j, s, 0, lane), 32, dup, gamma, wts, ops, nbad);
}

}  // namespace

static int check_launch2(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, what, __FILE__, __LINE__);
  return 0;
}

#define FRB_NSP2_SWITCH(nsp, CALL)                      \
  switch (nsp) {                                        \
    case 2: { constexpr int N = 2; CALL; } break;       \
    case 3: { constexpr int N = 3; CALL; } break;       \
    case 4: { constexpr int N = 4; CALL; } break;       \
    case 5: { constexpr int N = 5; CALL; } break;       \
    case 6: { constexpr int N = 6; CALL; } break;       \
    default: frb_set_error("2-D kernels support deg 1..5"); return FRB_ERR_ARG; \
  }

int frb_launch_euler2d_generic(frb_prob_t p, const double *u, const double *ua, double *out,
                               FrbStage st) {
  dim3 blk(32, 4), grd((p->nx + 31) / 32, (p->ny + 3) / 4);
  if (st.nested) { st.cdt *= st.cb; st.nested = 0; }
  FRB_NSP2_SWITCH(p->nsp, (euler2d_generic_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(
                              u, ua, out, p->nx, p->ny, 1.0 / p->Jx, 1.0 / p->Jy, p->gamma, p->flux, p->ops, st)));
  if (int rc = check_launch2("euler2d_generic_kernel")) return rc;
  return 1;
}

int frb_launch_ghost_fill2d(frb_prob_t p, double *u, int mode) {
  const int npp = p->nsp * p->nsp, nplanes = 4 * npp;
  cudaStream_t s = p->ctx->stream;
  dim3 blk(128), gx((p->ny + 2 + 127) / 128, nplanes), gy((p->nx + 2 + 127) / 128, nplanes);
  if (mode == FRB_GHOST_WAVE_X) {  // euler2d_wave.jl:127-132
    ghost_x_kernel<<<gx, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->nx, 1, -1, 1);
    ghost_y_kernel<<<gy, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->ny, 1, 2);
  } else if (mode == FRB_GHOST_WAVE_Y) {  // :159-164
    ghost_y_kernel<<<gy, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->ny, 1, -1);
    ghost_x_kernel<<<gx, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->nx, 1, 1, 1);
  } else if (mode == FRB_GHOST_COPY) {  // shock-vortex.jl:324-326
    ghost_y_kernel<<<gy, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, 1, p->ny, -1);
    ghost_x_kernel<<<gx, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, 0, p->nx, -1, 0);
  } else if (mode == FRB_GHOST_PERIODIC) {  // dev/parallelogram.jl:201-205
    ghost_x_kernel<<<gx, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->nx, 1, -1, 1);
    ghost_y_kernel<<<gy, blk, 0, s>>>(u, p->nx, p->ny, nplanes, npp, p->ny, 1, -1);
  } else if (mode == FRB_GHOST_CYLINDER) {  // dev/cylinder2.jl:176-187
    return frb_launch_ghost_cylinder(p, u);
  } else {
    frb_set_error("unknown ghost mode");
    return FRB_ERR_ARG;
  }
  if (int rc = check_launch2("ghost kernels")) return rc;
  return 2;
}

// row0 / rowN: whether the bottom / top ghost row is copied too (false for rows that a
// neighbouring rank writes, see frb_halo.cu)
int frb_launch_ring_copy2d(frb_prob_t p, const double *src, double *dst, bool row0, bool rowN) {
  const int nplanes = 4 * p->nsp * p->nsp;
  const int nring = 2 * (p->nx + 2) + 2 * p->ny;
  dim3 blk(128), grd((nring + 127) / 128, nplanes);
  ring_copy_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, dst, p->nx, p->ny, nplanes, row0, rowN);
  if (int rc = check_launch2("ring_copy_kernel")) return rc;
  return 1;
}

int frb_launch_ring_rows2d(frb_prob_t p, const double *src, double *d1, double *d2, double *d3, int ra, int rb) {
  if (rb < ra) return 0;
  dim3 blk(128), grd(rb - ra + 1, 4 * p->nsp * p->nsp);
  ring_rows_kernel<<<grd, blk, 0, p->ctx->stream>>>(src, d1, d2, d3, p->nx, p->ny, ra);
  if (int rc = check_launch2("ring_rows_kernel")) return rc;
  return 1;
}

// the x half of the ghost fill only (all rows, halo rows included): used by the slab-parallel
// path, where the y half is an exchange with the neighbouring ranks
int frb_launch_ghost_x2d(frb_prob_t p, double *u, int mode) {
  const int npp = p->nsp * p->nsp, nplanes = 4 * npp;
  dim3 blk(128), gx((p->ny + 2 + 127) / 128, nplanes);
  if (mode == FRB_GHOST_WAVE_X || mode == FRB_GHOST_PERIODIC)  // x half of :325 / :334 (the y half is the seam push)
    ghost_x_kernel<<<gx, blk, 0, p->ctx->stream>>>(u, p->nx, p->ny, nplanes, npp, p->nx, 1, -1, 1);
  else if (mode == FRB_GHOST_WAVE_Y)
    ghost_x_kernel<<<gx, blk, 0, p->ctx->stream>>>(u, p->nx, p->ny, nplanes, npp, p->nx, 1, 1, 1);
  else {
    frb_set_error("slab-parallel ghost fill supports the periodic modes (wave_x, wave_y, periodic)");
    return FRB_ERR_ARG;
  }
  if (int rc = check_launch2("ghost_x_kernel")) return rc;
  return 1;
}

int frb_launch_limiter2d(frb_prob_t p, double *u) {
  dim3 blk(32, 4), grd((p->nx + 31) / 32, (p->ny + 3) / 4);
  FRB_NSP2_SWITCH(p->nsp, (limiter2d_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(
                              u, p->nx, p->ny, p->gamma, p->lim_w, p->ops, p->flag)));
  if (int rc = check_launch2("limiter2d_kernel")) return rc;
  return 1;
}

int frb_rc_limiter2d(frb_prob_t p, double *u) {
  const RcGeom g = rc_geom(p->nx, p->ny, p->nsp);
  dim3 blk(32, 4), grd((g.ns + 3) / 4, p->ny);
  FRB_NSP2_SWITCH(p->nsp, (limiter2d_rc_kernel<N><<<grd, blk, 0, p->ctx->stream>>>(
                              u, g, p->gamma, p->lim_w, p->ops, p->flag)));
  if (int rc = check_launch2("limiter2d_rc_kernel")) return rc;
  return 1;
}
