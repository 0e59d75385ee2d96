// 2-D Euler on curvilinear structured quadrilaterals, one ROW-MARCHING launch per stage
// (FRB_KERNEL_CURV_MARCH; not the default: frb_euler2d_curv_fused.cu is faster, profiles/r02_summary.md section F).
// A block owns a strip of 30 elements and marches over a segment of element rows, so that the state is read once,
// the metric once, every y flux is evaluated once and the common fluxes never leave the SM (1.68 GB of DRAM traffic
// per 16-B stage at 1024^2 p3 against the 2.71 GB of round 1's two launches; 1.61 GB algorithmic).
//
// Reference semantics: dudt! of dev/parallelogram.jl:80-165 and dev/cylinder2.jl:52-164, as restated in
// frb_euler2d_curv_elem.cuh (same layouts, same factors, the formulas of row_xpass / row_ypass in their FOLD form).
//
// thread = (lane = element of the strip, l = point row); lanes 1..30 own their element, lanes 0 / 31 are the
// neighbours whose traces the x faces of the strip's edge elements need (the duplicate-lane form of the
// rectangular row-chunk kernel).  Every global input arrives by cp.async as a shared-memory image [plane][lane]
// (16 KB per row at p3) at least one phase before its use: the state of three consecutive rows (ring, two rows
// ahead), the metric of two rows (one row ahead), u_n of the row in flight; only the normals (4 doubles per
// thread and row) are plain loads, one row ahead in registers.  104 KB per block at p3: 2 blocks / SM.
// Per row jj = j0 .. j1-1:
//   A  row jj+1 has landed; its y traces at flux point p = l; common flux of y face jj+1 from the carried top
//      trace of row jj and the bottom trace of row jj+1 -> fyt[(jj+1) & 1][p][m][lane] (fyt[jj & 1] holds face jj
//      from the previous iteration);
//      x pass of row jj: traces in registers, the neighbours' by __shfl, x-face fluxes, point fluxes iJ [F; G],
//      r-derivative + x corrections; f2 of the row goes over the state values the row owner has just read (the
//      ring slot of row jj becomes the f2 tile)
//   B  y pass of row jj: s-derivative, y corrections from both fyt halves, stage update, store
//
// Arithmetic form: the flux traces of the correction are folded into the derivative matrix (FrbOps::dmod, as in
// the rectangular marching kernels) and the stage coefficients into the operators,
//   out = ca u_n + cb u + sum_q f1[q] M[k][q] + cxL[k] gl[k] FxL + cxR[k] gr[k] FxR
//                        + sum_q f2[q] M[l][q] + cyL[k] gl[l] FyB + cyR[k] gr[l] FyT,
//   M = -cdt dmod, gl = -cdt dgl, gr = -cdt dgr  (CurvMarchOps, built by the launcher);
// the common flux is a compile-time choice (riemann4_fast<FLUX>), reciprocals are MUFU-seeded, element indices are
// unsigned 32-bit (the launcher checks the array size).
#include "frb_euler2d_curv_elem.cuh"

namespace {

constexpr int kOwn = 30;
#ifndef FRB_CURV_MINB
#define FRB_CURV_MINB 2  // blocks per SM: 104 KB of shared memory per block at p3, up to 255 registers
#endif

struct CurvMarchOps {
  double ll[FRB_NSPMAX], lr[FRB_NSPMAX];
  double M[FRB_NSPMAX * FRB_NSPMAX];       // -cdt dmod[k][q]
  double gl[FRB_NSPMAX], gr[FRB_NSPMAX];   // -cdt dgl, -cdt dgr
  double ca, cb;
  int use_a;
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// local_frame -> common flux -> global_frame about the unit normal (c, s) (parallelogram.jl:119-123), the
// branch-free fluxes of the marching kernels; wall: the mirror state of cylinder2.jl:103-114 as the left state
template <int FLUX>
__device__ __forceinline__ void flux_frame(const double (&L)[4], const double (&R)[4], double c, double s, double gamma,
                                           double gm1, bool wall, double (&h)[4]) {
  double l0 = L[0], l1 = fma(L[2], s, L[1] * c), l2 = fma(-L[1], s, L[2] * c), l3 = L[3];
  const double r0 = R[0], r1 = fma(R[2], s, R[1] * c), r2 = fma(-R[1], s, R[2] * c), r3 = R[3];
  if (wall) {
    const double ir = frb::rcp_fast(r0), U = r1 * ir, V = r2 * ir;
    const double lam = 0.5 * r0 * frb::rcp_fast(gm1 * (r3 - 0.5 * (r1 * r1 + r2 * r2) * ir));
    const double t = lam - 1.0;
    const double rn = (1.0 - t) * frb::rcp_fast(1.0 + t) * r0, ln = 2.0 - lam;
    l0 = rn;
    l1 = rn * (-U);
    l2 = rn * V;
    l3 = 0.5 * rn * frb::rcp_fast(ln * gm1) + 0.5 * rn * (U * U + V * V);
  }
  const frb::Flux4 f = frb::riemann4_fast<FLUX>(l0, l1, l2, l3, r0, r1, r2, r3, gamma, gm1);
  h[0] = f.f0;
  h[1] = fma(-f.f2, s, f.f1 * c);
  h[2] = fma(f.f1, s, f.f2 * c);
  h[3] = f.f3;
}

template <int NSP, int FLUX>
__global__ void __launch_bounds__(32 * NSP, FRB_CURV_MINB)
euler2d_curv_march_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
                          CurvGeom g, double gamma, CurvMarchOps ops, int rows) {
  using frbcurv::plane;
  constexpr int NPL = NSP * NSP * 4;
  extern __shared__ double sm[];
  double *ring = sm;                 // [3][plane][lane]  state images of rows jj-1 / jj / jj+1 (mod 3); the slot of
                                     //                   row jj becomes the f2 tile once its owner has read it
  double *mJ = sm + 3 * NPL * 32;    // [2][plane][lane]  metric images of rows jj / jj+1, by row parity
  double *sA = mJ + 2 * NPL * 32;    // [plane][lane]     u_n image of row jj
  double *fyt = sA + NPL * 32;       // [2][p][m][lane]   common fluxes of y faces jj and jj+1, by face parity
  const int lane = threadIdx.x, l = threadIdx.y;
  const int nx = g.nx, ny = g.ny;
  const unsigned NXG = nx + 2, NE = NXG * (unsigned)(ny + 2);  // 32-bit element indices: the launcher checks the size
  const int iraw = blockIdx.x * kOwn + lane;
  const int i = iraw <= nx + 1 ? iraw : nx + 1;  // lanes past the mesh re-read the last ghost column
  const bool owned = lane >= 1 && lane <= kOwn && iraw <= nx;
  const int j0 = 1 + blockIdx.y * rows;
  const int j1 = min(j0 + rows, ny + 1);       // rows j0 .. j1-1
  const int ifx = min(max(i, 1), nx + 1) - 1;  // x face left of element i (clamped for the lanes that own none)
  const int ify = min(max(i, 1), nx) - 1;      // column of the y faces
  const unsigned s1 = (unsigned)(nx + 1) * ny, s2 = (unsigned)nx * (ny + 1), sfp = (unsigned)nx * ny;
  const double gm1 = gamma - 1.0;
  const bool wall = g.wall_xlo && i == 1;

  // row r of an array [nx+2, ny+2, NSP, NSP, 4] -> image [plane][lane]: every thread moves the 4 NSP planes of its
  // own point row
  auto fetch = [&](const double *__restrict__ src, int r, double *img) {
    const unsigned e = i + NXG * (unsigned)r;
    double *dst = img + lane;
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
      for (int k = 0; k < NSP; ++k) cp_async8(dst + plane<NSP>(k, l, m) * 32, src + (e + NE * plane<NSP>(k, l, m)));
  };
  // y traces of the row in ring slot `slot` at flux point p = l: bottom (ll) and top (lr)
  auto traces_y = [&](int slot, double (&bot)[4], double (&tp)[4]) {
    const double *S = ring + slot * NPL * 32 + lane;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double v0 = S[plane<NSP>(l, 0, m) * 32];
      double b = v0 * ops.ll[0], t = v0 * ops.lr[0];
#pragma unroll
      for (int q = 1; q < NSP; ++q) {
        const double v = S[plane<NSP>(l, q, m) * 32];
        b = fma(v, ops.ll[q], b);
        t = fma(v, ops.lr[q], t);
      }
      bot[m] = b;
      tp[m] = t;
    }
  };
  // common flux of y face jf (1..ny+1) of the thread's column -> fyt[jf & 1]
  auto flux_y = [&](int jf, const double (&lo)[4], const double (&hi)[4], double c, double s) {
    double h[4];
    flux_frame<FLUX>(lo, hi, c, s, gamma, gm1, false, h);
    double *F = fyt + (jf & 1) * NSP * 4 * 32 + lane;
#pragma unroll
    for (int m = 0; m < 4; ++m) F[(l * 4 + m) * 32] = h[m];
  };

  int sc = j0 % 3;  // slot of row jj; rows jj-1 / jj+1 sit in the slots before / after it (mod 3)
  double top[4];    // top trace of row jj at flux point l
  // normals, one row ahead of their use: x face left of the element in row jj, y faces jj (b) and jj+1 (t)
  double n1c = g.n1[ifx + (unsigned)(nx + 1) * (j0 - 1)], n1s = g.n1[ifx + (unsigned)(nx + 1) * (j0 - 1) + s1];
  double nbc = g.n2[ify + (unsigned)nx * (j0 - 1)], nbs = g.n2[ify + (unsigned)nx * (j0 - 1) + s2];
  double ntc = g.n2[ify + (unsigned)nx * j0], nts = g.n2[ify + (unsigned)nx * j0 + s2];
  {
    const int sp = sc == 0 ? 2 : sc - 1, sn = sc == 2 ? 0 : sc + 1;
    fetch(u, j0 - 1, ring + sp * NPL * 32);
    fetch(u, j0, ring + sc * NPL * 32);
    cp_commit();
    fetch(u, j0 + 1, ring + sn * NPL * 32);
    fetch(g.iJ, j0, mJ + (j0 & 1) * NPL * 32);
    cp_commit();
    cp_wait_but_one();
    __syncthreads();
    double b0[4], t0[4], b1[4];
    traces_y(sp, b0, t0);
    traces_y(sc, b1, top);
    flux_y(j0, t0, b1, nbc, nbs);
  }
  unsigned eoff = i + NXG * (unsigned)j0;  // element (i, jj)

  for (int jj = j0; jj < j1; ++jj, eoff += NXG) {
    const int sp = sc == 0 ? 2 : sc - 1, sn = sc == 2 ? 0 : sc + 1;
    cp_wait_all();
    __syncthreads();  // A: row jj+1 is in ring[sn], the metric of row jj in mJ[jj & 1]; ring[sp], mJ[(jj+1) & 1],
                      //    sA and fyt[(jj+1) & 1] are free
    if (ops.use_a) fetch(ua, jj, sA);
    cp_commit();
    if (jj + 2 <= j1) fetch(u, jj + 2, ring + sp * NPL * 32);
    if (jj + 1 < j1) fetch(g.iJ, jj + 1, mJ + ((jj + 1) & 1) * NPL * 32);
    cp_commit();
    // normals of the next row (clamped on the last one), used after the next barrier A
    const int jn = jj + 1 < j1 ? jj + 1 : jj;
    const double n1c_n = g.n1[ifx + (unsigned)(nx + 1) * (jn - 1)], n1s_n = g.n1[ifx + (unsigned)(nx + 1) * (jn - 1) + s1];
    const double ntc_n = g.n2[ify + (unsigned)nx * jn], nts_n = g.n2[ify + (unsigned)nx * jn + s2];

    // ---- y face jj+1
    {
      double bot[4], tp[4];
      traces_y(sn, bot, tp);
      flux_y(jj + 1, top, bot, ntc, nts);
#pragma unroll
      for (int m = 0; m < 4; ++m) top[m] = tp[m];
    }

    // ---- x pass of row jj (every lane: the edge lanes supply traces and the flux of their left face)
    double d[NSP][4], gyL[NSP], gyR[NSP];
    {
      double *S = ring + sc * NPL * 32 + lane;
      const double *J = mJ + (jj & 1) * NPL * 32 + lane;
      const double gll = ops.gl[l], grl = ops.gr[l];
      const unsigned ifp = (unsigned)(owned ? i - 1 : 0) + (unsigned)nx * (jj - 1);
      double w[NSP][4], f1[NSP][4], gxL[NSP], gxR[NSP];
      double nrc = 0.0, nrs = 0.0, fx0 = 0.0, fx1 = 0.0;
      if (g.fpc) {
        fx0 = g.fpc[ifp + sfp * (l + NSP * 0)];
        fx1 = g.fpc[ifp + sfp * (l + NSP * 1)];
      } else {
        nrc = __shfl_down_sync(0xffffffffu, n1c, 1);
        nrs = __shfl_down_sync(0xffffffffu, n1s, 1);
      }
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double a[4];  // a11, a21, a12, a22
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          w[k][m] = S[plane<NSP>(k, l, m) * 32];
          a[m] = J[plane<NSP>(k, l, m) * 32];
        }
        const double w0 = w[k][0], w1 = w[k][1], w2 = w[k][2], w3 = w[k][3];
        const double r = frb::rcp_fast(w0), vx = w1 * r, vy = w2 * r;
        const double p = gm1 * fma(-0.5, fma(w1, vx, w2 * vy), w3);
        const double h = w3 + p;
        const double F[4] = {w1, fma(w1, vx, p), w1 * vy, h * vx};
        const double G[4] = {w2, w2 * vx, fma(w2, vy, p), h * vy};
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          f1[k][m] = fma(a[2], G[m], a[0] * F[m]);
          S[plane<NSP>(k, l, m) * 32] = fma(a[3], G[m], a[1] * F[m]);  // f2 over the state value just read
        }
        if (g.fpc) {  // cylinder2.jl:155-158
          gxL[k] = fx0 * ops.gl[k];
          gxR[k] = fx1 * ops.gr[k];
          gyL[k] = g.fpc[ifp + sfp * (k + NSP * 2)] * gll;
          gyR[k] = g.fpc[ifp + sfp * (k + NSP * 3)] * grl;
        } else {  // parallelogram.jl:145-148
          gxL[k] = fma(a[2], n1s, a[0] * n1c) * ops.gl[k];
          gxR[k] = fma(a[2], nrs, a[0] * nrc) * ops.gr[k];
          gyL[k] = fma(a[3], nbs, a[1] * nbc) * gll;
          gyR[k] = fma(a[3], nts, a[1] * ntc) * grl;
        }
      }
      double tl[4], tr[4], FxL[4], FxR[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double x = w[0][m] * ops.ll[0], y = w[0][m] * ops.lr[0];
#pragma unroll
        for (int q = 1; q < NSP; ++q) {
          x = fma(w[q][m], ops.ll[q], x);
          y = fma(w[q][m], ops.lr[q], y);
        }
        tl[m] = x;
        tr[m] = __shfl_up_sync(0xffffffffu, y, 1);  // u_face[i-1, j, 2, l, m]
      }
      flux_frame<FLUX>(tr, tl, n1c, n1s, gamma, gm1, wall, FxL);
#pragma unroll
      for (int m = 0; m < 4; ++m) FxR[m] = __shfl_down_sync(0xffffffffu, FxL[m], 1);
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < NSP; ++k) {
          double x = ops.cb * w[k][m];
#pragma unroll
          for (int q = 0; q < NSP; ++q) x = fma(f1[q][m], ops.M[k * FRB_NSPMAX + q], x);
          x = fma(gxL[k], FxL[m], x);
          d[k][m] = fma(gxR[k], FxR[m], x);
        }
    }
    cp_wait_but_one();  // this thread's part of the u_n image (the row prefetches stay in flight)
    __syncthreads();    // B: ring[sc] = f2 of row jj, fyt = fluxes of faces jj and jj+1, sA = u_n of row jj

    // ---- y pass of row jj
    if (owned) {
      const double *T = ring + sc * NPL * 32 + lane, *A = sA + lane;
      const double *FB = fyt + (jj & 1) * NSP * 4 * 32 + lane, *FT = fyt + ((jj + 1) & 1) * NSP * 4 * 32 + lane;
      double Ml[NSP];
#pragma unroll
      for (int q = 0; q < NSP; ++q) Ml[q] = ops.M[l * FRB_NSPMAX + q];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < NSP; ++k) {
          const int yi = g.fy_row ? l : k;
          double x = d[k][m];
#pragma unroll
          for (int r = 0; r < NSP; ++r) x = fma(T[plane<NSP>(k, r, m) * 32], Ml[r], x);
          x = fma(gyL[k], FB[(yi * 4 + m) * 32], x);
          x = fma(gyR[k], FT[(yi * 4 + m) * 32], x);
          if (ops.use_a) x = fma(ops.ca, A[plane<NSP>(k, l, m) * 32], x);
          out[eoff + NE * plane<NSP>(k, l, m)] = x;
        }
    }
    sc = sn;
    n1c = n1c_n;
    n1s = n1s_n;
    nbc = ntc;
    nbs = nts;
    ntc = ntc_n;
    nts = nts_n;
  }
}

int rows_override() {  // FRB_CURV_ROWS = rows per segment (tests, tuning); read per launch
  const char *s = getenv("FRB_CURV_ROWS");
  return s ? atoi(s) : 0;
}

template <int NSP, int FLUX>
int launch(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g, const FrbStage &st) {
  constexpr size_t smem = sizeof(double) * 32 * (6 * NSP * NSP * 4 + 2 * NSP * 4);  // ring[3] + mJ[2] + sA + fyt[2]
  static int per_sm = 0;
  if (!per_sm) {
    FRB_CUDA(cudaFuncSetAttribute(euler2d_curv_march_kernel<NSP, FLUX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    const char *cv = getenv("FRB_CURV_CARVEOUT");  // percent of the unified L1 / shared memory given to shared memory
    FRB_CUDA(cudaFuncSetAttribute(euler2d_curv_march_kernel<NSP, FLUX>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout, cv ? atoi(cv) : 100));
    int n = 0;
    FRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, euler2d_curv_march_kernel<NSP, FLUX>, 32 * NSP, smem));
    per_sm = n > 0 ? n : 1;
  }
  CurvMarchOps mo;
  for (int k = 0; k < FRB_NSPMAX; ++k) {
    mo.ll[k] = p->ops.ll[k];
    mo.lr[k] = p->ops.lr[k];
    mo.gl[k] = -st.cdt * p->ops.dgl[k];
    mo.gr[k] = -st.cdt * p->ops.dgr[k];
    for (int q = 0; q < FRB_NSPMAX; ++q) mo.M[k * FRB_NSPMAX + q] = -st.cdt * p->ops.dmod[k * FRB_NSPMAX + q];
  }
  mo.ca = st.use_a ? st.ca : 0.0;
  mo.cb = st.cb;
  mo.use_a = st.use_a && ua != nullptr;
  // segments: as many as fit on the device at once (one wave of equal blocks), never shorter than one row
  const int strips = (p->nx + kOwn - 1) / kOwn;
  const int cap = p->ctx->sm_count * per_sm;
  int segs = cap / strips;
  if (segs < 1) segs = 1;
  if (segs > p->ny) segs = p->ny;
  int rows = (p->ny + segs - 1) / segs;
  if (rows_override() > 0) rows = rows_override();
  if (rows > p->ny) rows = p->ny;
  segs = (p->ny + rows - 1) / rows;
  if (segs > 65535) {
    frb_set_error("euler2d_curv: too many row segments");
    return FRB_ERR_ARG;
  }
  if ((double)(p->nx + 2) * (p->ny + 2) * NSP * NSP * 4 >= 4294967296.0) {
    frb_set_error("euler2d_curv: the marching kernel indexes with 32 bits (state arrays below 2^32 elements)");
    return FRB_ERR_ARG;
  }
  dim3 blk(32, NSP), grd(strips, segs);
  euler2d_curv_march_kernel<NSP, FLUX><<<grd, blk, smem, p->ctx->stream>>>(u, ua, out, g, p->gamma, mo, rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "euler2d_curv_march_kernel", __FILE__, __LINE__);
  return 1;
}

template <int NSP>
int launch_flux(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g, const FrbStage &st) {
  switch (g.flux) {
    case FRB_FLUX_HLL: return launch<NSP, FRB_FLUX_HLL>(p, u, ua, out, g, st);
    case FRB_FLUX_LF: return launch<NSP, FRB_FLUX_LF>(p, u, ua, out, g, st);
    case FRB_FLUX_ROE: return launch<NSP, FRB_FLUX_ROE>(p, u, ua, out, g, st);
    default: frb_set_error("euler2d_curv: unknown flux"); return FRB_ERR_ARG;
  }
}

}  // namespace

// st: the branch-free form of frb_launch_euler2d_curv (rhs_only and nested already mapped onto ca, cb, cdt)
int frb_launch_euler2d_curv_march(frb_prob_t p, const double *u, const double *ua, double *out, const CurvGeom &g,
                                  const FrbStage &st) {
  switch (p->nsp) {
    case 2: return launch_flux<2>(p, u, ua, out, g, st);
    case 3: return launch_flux<3>(p, u, ua, out, g, st);
    case 4: return launch_flux<4>(p, u, ua, out, g, st);
    default: frb_set_error("euler2d_curv: deg must be in 1..3"); return FRB_ERR_ARG;
  }
}
