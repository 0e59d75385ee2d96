// 1-D BGK kinetic equation, fused residual + RK stage (config 4: 8192 cells x 256 velocities,
// data-parallel over phase space).  Reference: mol! of example/bgk_wave.jl:69-129 -- Maxwellian from
// the moments (:77-81), flux v_j u / (dx/2) (:85-88), interp_face! (:99-101), upwind interface flux
// through the periodic f2e / e2f tables (:42-67, :103-107), poly_derivative! (:114-116), correction
// + relaxation (M - u)/tau (:122-128).  State u[cell, velocity, sp], cell fastest.
//
// Two launches per stage (the default), or one where the velocity grid fits a register tile (bgk1d_fused_kernel,
// below: even ncell, nu <= 256, deg 1..3, FRB_BGK_ONE_PASS=1).  The two launches:
//   bgk_moments_kernel  thread = (cell, sp): the three moments over the nu velocities in the
//                       reference's summation order, 16 loads in flight per thread (the sum is a
//                       serial chain, the loads are not); writes rho*sqrt(lambda/pi), U, lambda.
//   bgk1d_kernel        thread = (cell, velocity): own block + the upwind neighbour's block, 1 / J
//                       precomputed per cell, one exp per point.
// model FRB_BGK_KINETIC_ADVECTION is the mol! of example/advection_kinetic.jl:73-128: the same residual with the
// Maxwellian of prim = [rho, a, 1] (:80-88) -- only the epilogue of the moments kernel differs.
// Both are L2 / HBM streams of the 50 MB state; nothing here is GEMM-shaped.
#include <atomic>
#include <cstdlib>

#include "frb_internal.cuh"

namespace {

template <int NSP>
__device__ __forceinline__ double dotn(const double *a, const double *l) {
  double s = a[0] * l[0];
#pragma unroll
  for (int q = 1; q < NSP; ++q) s = fma(a[q], l[q], s);
  return s;
}

// exp(x) for x <= 0 (the Maxwellian's exponent -lambda (v - U)^2), branch-free: k = round(x log2 e), r = x - k ln 2
// in two pieces, Taylor polynomial of degree 13 on |r| <= 0.347 (truncation 4e-18), scaling by 2^k through the
// exponent field; 0 below -700 (the true value is < 1e-304).  Within 1 ulp of the library exp on this range;
// no convergence barriers, so the six exponentials of a velocity interleave in the FP64 pipe.
__device__ __forceinline__ double exp_neg(double x) {
  const double t = fma(x, 1.4426950408889634074, 6755399441055744.0);  // 2^52 + 2^51: k in the low word
  const int k = __double2loint(t);
  const double kd = t - 6755399441055744.0;
  double r = fma(kd, -6.93147180369123816490e-01, x);
  r = fma(kd, -1.90821492927058770002e-10, r);
  double p = 1.6059043836821613e-10;             // 1/13!
  p = fma(p, r, 2.08767569878680989792e-09);     // 1/12!
  p = fma(p, r, 2.50521083854417187751e-08);
  p = fma(p, r, 2.75573192239858906526e-07);
  p = fma(p, r, 2.75573192239858906526e-06);
  p = fma(p, r, 2.48015873015873015873e-05);
  p = fma(p, r, 1.98412698412698412698e-04);
  p = fma(p, r, 1.38888888888888888889e-03);
  p = fma(p, r, 8.33333333333333333333e-03);
  p = fma(p, r, 4.16666666666666666667e-02);
  p = fma(p, r, 1.66666666666666666667e-01);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double s = __hiloint2double((k + 1023) << 20, 0);  // 2^k, k >= -1010 here
  return x < -700.0 ? 0.0 : p * s;
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
#ifndef FRB_BGK_MOMGROUPS
#define FRB_BGK_MOMGROUPS 8  // kernel experiment switch (scripts/build_variants.py)
#endif
constexpr int kMomGroups = FRB_BGK_MOMGROUPS;
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

// VPT velocities per thread (j, j + nu / VPT, ...): the Maxwellian parameters of the cell, the neighbour indices and
// the constants of the exponential are fetched once for all of them -- the kernel is instruction-issue bound (ncu:
// issue slots 71 % busy, 2/3 of the executed instructions are not FP64), so instructions per DOF are what counts.
#ifndef FRB_BGK_VPT
#define FRB_BGK_VPT 2
#endif
template <int NSP, int VPT>
__global__ void __launch_bounds__(128)
bgk1d_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
             const double *__restrict__ prim, const double *__restrict__ inv_j,
             const double *__restrict__ velo, int ncell, int nu, double inv_tau, FrbOps ops, FrbStage st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const size_t vs = (size_t)ncell * nu;
  const int il = i == 0 ? ncell - 1 : i - 1, ir = i == ncell - 1 ? 0 : i + 1;  // periodic neighbours
  const double ij = inv_j[i], ijl = inv_j[il], ijr = inv_j[ir];
  double pre[NSP], U[NSP], lam[NSP];
#pragma unroll
  for (int p = 0; p < NSP; ++p) {
    const size_t po = i + (size_t)ncell * p;
    pre[p] = prim[po];
    U[p] = prim[po + (size_t)ncell * NSP];
    lam[p] = prim[po + 2 * (size_t)ncell * NSP];
  }
#pragma unroll
  for (int s = 0; s < VPT; ++s) {
    const int j = blockIdx.y + s * (nu / VPT);
    const double v = velo[j];
    const bool pos = v >= 0.0;  // heaviside delta, bgk_wave.jl:26
    // only the upwind neighbour contributes: left cell for v >= 0, right cell otherwise (periodic)
    const int in = pos ? il : ir;
    const double sc = v * ij, sn = v * (pos ? ijl : ijr);  // v / J
    const size_t row = (size_t)ncell * j;
    double uc[NSP], f[NSP], fn[NSP];
#pragma unroll
    for (int q = 0; q < NSP; ++q) {
      uc[q] = u[i + row + vs * q];
      f[q] = sc * uc[q];  // :85-88
      fn[q] = sn * u[in + row + vs * q];
    }
    const double fL = dotn<NSP>(f, ops.ll), fR = dotn<NSP>(f, ops.lr);  // interp_face! :99-101
    // f_interaction at the left / right face of cell i (:103-107): own trace downwind, neighbour's upwind
    const double cl = pos ? dotn<NSP>(fn, ops.lr) - fL : 0.0;   // fi0 - fL
    const double cr = pos ? 0.0 : dotn<NSP>(fn, ops.ll) - fR;   // fi1 - fR
#pragma unroll
    for (int p = 0; p < NSP; ++p) {
      const double c = v - U[p];
      const double M = pre[p] * exp_neg(-lam[p] * (c * c));  // maxwellian
      const double rhs1 = dotn<NSP>(f, &ops.lpdm[p * FRB_NSPMAX]);  // poly_derivative! :114-116
      const double du = -(rhs1 + cl * ops.dgl[p] + cr * ops.dgr[p]) + (M - uc[p]) * inv_tau;
      const size_t idx = i + row + vs * p;
      out[idx] = stage_out(st, ua, idx, uc[p], du);
    }
  }
}

// Two adjacent cells per thread (even ncell): own values, u_n, the Maxwellian parameters and the result move as
// 16-byte accesses, and one of the two upwind neighbours is the other cell of the pair (no load, its flux trace is
// computed anyway) -- half the load, store and address instructions per DOF of the kernel above.
template <int NSP>
__global__ void __launch_bounds__(128)
bgk1d_pair_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
                  const double *__restrict__ prim, const double *__restrict__ inv_j,
                  const double *__restrict__ velo, int ncell, int nu, double inv_tau, FrbOps ops, FrbStage st) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);  // cells i, i + 1
  if (i >= ncell) return;
  const int j = blockIdx.y;
  const size_t vs = (size_t)ncell * nu, row = (size_t)ncell * j;
  const double v = velo[j];
  const bool pos = v >= 0.0;  // heaviside delta, bgk_wave.jl:26
  // the one upwind neighbour outside the pair: cell i - 1 for v >= 0, cell i + 2 otherwise (periodic)
  const int in = pos ? (i == 0 ? ncell - 1 : i - 1) : (i + 2 >= ncell ? 0 : i + 2);
  const double2 ij = *reinterpret_cast<const double2 *>(inv_j + i);
  const double s0 = v * ij.x, s1 = v * ij.y, sn = v * inv_j[in];  // v / J
  double2 uc[NSP];
  double f0[NSP], f1[NSP], fn[NSP];
#pragma unroll
  for (int q = 0; q < NSP; ++q) {
    uc[q] = *reinterpret_cast<const double2 *>(u + i + row + vs * q);
    fn[q] = sn * u[in + row + vs * q];
  }
#pragma unroll
  for (int q = 0; q < NSP; ++q) { f0[q] = s0 * uc[q].x; f1[q] = s1 * uc[q].y; }  // :85-88
  const double fL0 = dotn<NSP>(f0, ops.ll), fR0 = dotn<NSP>(f0, ops.lr);  // interp_face! :99-101
  const double fL1 = dotn<NSP>(f1, ops.ll), fR1 = dotn<NSP>(f1, ops.lr);
  // f_interaction - own trace at the left / right face (:103-107): zero on the downwind side
  const double cl0 = pos ? dotn<NSP>(fn, ops.lr) - fL0 : 0.0, cr0 = pos ? 0.0 : fL1 - fR0;
  const double cl1 = pos ? fR0 - fL1 : 0.0, cr1 = pos ? 0.0 : dotn<NSP>(fn, ops.ll) - fR1;
  const bool with_a = st.use_a && !st.rhs_only;
#pragma unroll
  for (int p = 0; p < NSP; ++p) {
    const size_t po = i + (size_t)ncell * p;
    const double2 pre = *reinterpret_cast<const double2 *>(prim + po);
    const double2 U = *reinterpret_cast<const double2 *>(prim + po + (size_t)ncell * NSP);
    const double2 lam = *reinterpret_cast<const double2 *>(prim + po + 2 * (size_t)ncell * NSP);
    const double c0 = v - U.x, c1 = v - U.y;
    const double M0 = pre.x * exp_neg(-lam.x * (c0 * c0)), M1 = pre.y * exp_neg(-lam.y * (c1 * c1));  // maxwellian
    const double d0 = -(dotn<NSP>(f0, &ops.lpdm[p * FRB_NSPMAX]) + cl0 * ops.dgl[p] + cr0 * ops.dgr[p]) +
                      (M0 - uc[p].x) * inv_tau;
    const double d1 = -(dotn<NSP>(f1, &ops.lpdm[p * FRB_NSPMAX]) + cl1 * ops.dgl[p] + cr1 * ops.dgr[p]) +
                      (M1 - uc[p].y) * inv_tau;
    const size_t idx = i + row + vs * p;
    double2 o;
    if (st.rhs_only) {
      o = make_double2(d0, d1);
    } else {
      o.x = st.nested ? st.cb * (uc[p].x + st.cdt * d0) : st.cb * uc[p].x + st.cdt * d0;
      o.y = st.nested ? st.cb * (uc[p].y + st.cdt * d1) : st.cb * uc[p].y + st.cdt * d1;
      if (with_a) {
        const double2 an = *reinterpret_cast<const double2 *>(ua + idx);
        o.x = st.ca * an.x + o.x;
        o.y = st.ca * an.y + o.y;
      }
    }
    *reinterpret_cast<double2 *>(out + idx) = o;
  }
}

// ---- one pass: moments, Maxwellian and the stage update from ONE read of u ----------------------------------
// A CTA owns kFC = 8 cells x all nu velocities x NSP points and keeps them in REGISTERS: thread = (cell pair,
// velocity group) holds u of 2 cells (16-byte loads: a warp instruction covers 8 rows of 64 bytes) x VR
// velocities x NSP points (24 doubles at cfg4), all loads issued before the first use.  The three moments are
// summed per thread over its velocities, across the 8 velocity groups of a warp by xor shuffles and across the 8
// warps through 4.6 KB of shared memory, in a fixed order (deterministic; the reference's `sum` has no specified
// order, bgk_wave.jl:78); then every thread updates its own values from the registers.  The upwind neighbour's
// flux trace is the adjacent cell of the pair, the adjacent lane's (shuffle), or -- for the two cells at the
// CTA's edges -- a halo cell fetched with the first loads (L2 hits: the neighbouring CTA streams that line).
// u crosses HBM / L2 once per stage instead of twice, and `prim` never leaves the SM.
constexpr int kFC = 8;        // cells per CTA
constexpr int kFG = 64;       // velocity groups per CTA (8 per warp x 8 warps)
constexpr int kFusedMaxVR = 4;

#ifndef FRB_BGK_MINB
#define FRB_BGK_MINB 2  // kernel experiment switch (scripts/build_variants.py)
#endif
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Persistent: one CTA per SM slot walks over the cell blocks (tiles) t = blockIdx.x, + gridDim.x, ...  While a
// tile is being worked on, the NEXT tile of u (and of u_n for a 24-B stage) streams into shared memory with
// cp.async (LDGSTS, 16 B per thread and piece, every thread into its own slots: no barrier needed to read them
// back).  The exposed DRAM / L2 latency of the first loads -- 37 % of all stall samples of the non-pipelined
// version (profiles/r02_summary.md) -- is paid once per CTA instead of once per tile.
template <int NSP, int VR>
__global__ void __launch_bounds__(256, FRB_BGK_MINB)
bgk1d_fused_kernel(const double *__restrict__ u, const double *__restrict__ ua, double *__restrict__ out,
                   const double *__restrict__ inv_j, const double *__restrict__ velo,
                   const double *__restrict__ wts, int ncell, int nu, double inv_tau, FrbOps ops, FrbStage st,
                   int model, double a) {
  extern __shared__ __align__(16) double2 stage[];  // [VR * NSP][256] for u, then the same for u_n
  __shared__ double red[8][kFC / 2][2][NSP][3];
  __shared__ double prim_s[kFC][NSP][3];
  // flux trace of the halo cell next to the CTA's first / last pair, upwind side only: computed by the edge lanes
  // with the first loads, parked in shared memory until the update (registers are the scarce resource here)
  __shared__ double halo_s[2][VR][kFG];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cp = lane & 3;                // cell pair within the CTA
  const int vg = warp * 8 + (lane >> 2);  // velocity group: velocities vg, vg + 64, ...
  const size_t vs = (size_t)ncell * nu;
  const int ntiles = (ncell + kFC - 1) / kFC;
  const bool with_a = st.use_a && !st.rhs_only;
  double2 *const stage_a = stage + VR * NSP * 256;

  auto prefetch = [&](const double *src, double2 *dst, int tile) {
    const int i = tile * kFC + 2 * cp;
    if (i < ncell) {
#pragma unroll
      for (int r = 0; r < VR; ++r) {
        const int j = vg + kFG * r;
        if (j < nu) {
#pragma unroll
          for (int q = 0; q < NSP; ++q) cp_async16(&dst[(r * NSP + q) * 256 + tid], src + i + (size_t)ncell * j + vs * q);
        }
      }
    }
    cp_async_commit();  // always: the sequence of groups is the same for every thread and every tile
  };
  // Groups in flight, oldest first, at the top of tile t:  u(t) [, u_n(t)].  A 24-B stage waits for all but the
  // youngest (u_n(t) keeps streaming through the moment phase), issues u(t + 1), and before the update waits
  // again for all but the youngest (now u(t + 1)), i.e. for u_n(t); u_n(t + 1) is issued after the update.
  int tile = blockIdx.x;
  if (tile < ntiles) {
    prefetch(u, stage, tile);
    if (with_a) prefetch(ua, stage_a, tile);
  }
  for (; tile < ntiles; tile += gridDim.x) {
    const int i = tile * kFC + 2 * cp;  // first cell of the pair (ncell is even: in range or out as a pair)
    const bool cell_ok = i < ncell;
    const bool edge_lo = cp == 0, edge_hi = cell_ok && (cp == kFC / 2 - 1 || i + 2 >= ncell);
    const int il = i == 0 ? ncell - 1 : i - 1, ir = i + 2 >= ncell ? 0 : i + 2;  // periodic halo cells

    if (with_a) cp_async_wait_but_one();  // this thread's pieces of u(t) have landed
    else cp_async_wait_all();
    double2 w[VR][NSP];
#pragma unroll
    for (int r = 0; r < VR; ++r) {
      const bool ok = cell_ok && vg + kFG * r < nu;
#pragma unroll
      for (int q = 0; q < NSP; ++q) w[r][q] = ok ? stage[(r * NSP + q) * 256 + tid] : make_double2(0.0, 0.0);
    }
    const int next = tile + gridDim.x;
    prefetch(u, stage, next < ntiles ? next : ntiles);  // the slots were just read by their owner (past the end: an empty group)
    if (cell_ok && (edge_lo || edge_hi)) {
#pragma unroll
      for (int r = 0; r < VR; ++r) {
        const int j = vg + kFG * r;
        if (j >= nu) continue;
        const double v = velo[j];
        if (edge_lo && v >= 0.0) {  // right trace of the flux of cell il
          double fn[NSP];
          const double sn = v * inv_j[il];
#pragma unroll
          for (int q = 0; q < NSP; ++q) fn[q] = sn * u[il + (size_t)ncell * j + vs * q];
          halo_s[0][r][vg] = dotn<NSP>(fn, ops.lr);
        }
        if (edge_hi && !(v >= 0.0)) {  // left trace of the flux of cell ir
          double fn[NSP];
          const double sn = v * inv_j[ir];
#pragma unroll
          for (int q = 0; q < NSP; ++q) fn[q] = sn * u[ir + (size_t)ncell * j + vs * q];
          halo_s[1][r][vg] = dotn<NSP>(fn, ops.ll);
        }
      }
    }

    // ---- moments (bgk_wave.jl:77-79): per thread over its velocities, then over the velocity groups
    {
      double m[2][NSP][3];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < NSP; ++q) m[c][q][0] = m[c][q][1] = m[c][q][2] = 0.0;
#pragma unroll
      for (int r = 0; r < VR; ++r) {
        const int j = vg + kFG * r;
        const double v = j < nu ? velo[j] : 0.0, wt = j < nu ? wts[j] : 0.0;
        const double wv = wt * v, wvv = wt * (v * v);
#pragma unroll
        for (int q = 0; q < NSP; ++q) {
          m[0][q][0] += wt * w[r][q].x; m[0][q][1] += wv * w[r][q].x; m[0][q][2] += wvv * w[r][q].x;
          m[1][q][0] += wt * w[r][q].y; m[1][q][1] += wv * w[r][q].y; m[1][q][2] += wvv * w[r][q].y;
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int q = 0; q < NSP; ++q)
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            double x = m[c][q][k];
            x += __shfl_xor_sync(0xffffffffu, x, 4);
            x += __shfl_xor_sync(0xffffffffu, x, 8);
            x += __shfl_xor_sync(0xffffffffu, x, 16);
            if (lane < 4) red[warp][cp][c][q][k] = x;
          }
    }
    __syncthreads();
    if (tid < kFC * NSP) {
      const int c = tid % kFC, q = tid / kFC;
      double w0 = 0.0, w1 = 0.0, w2 = 0.0;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        w0 += red[g][c >> 1][c & 1][q][0];
        w1 += red[g][c >> 1][c & 1][q][1];
        w2 += red[g][c >> 1][c & 1][q][2];
      }
      if (model == FRB_BGK_KINETIC_ADVECTION) {  // example/advection_kinetic.jl:80-88: prim = [rho, a, 1.0]
        prim_s[c][q][0] = w0 * sqrt(1.0 / 3.14159265358979323846);
        prim_s[c][q][1] = a;
        prim_s[c][q][2] = 1.0;
      } else {
        w2 *= 0.5;
        const double lam = 0.5 * w0 / (3.0 - 1.0) / (w2 - 0.5 * w1 * w1 / w0);  // conserve_prim(w, 3.0)
        prim_s[c][q][0] = w0 * sqrt(lam / 3.14159265358979323846);               // rho * sqrt(lambda / pi)
        prim_s[c][q][1] = w1 / w0;
        prim_s[c][q][2] = lam;
      }
    }
    __syncthreads();

    // ---- update: every thread its own (2 cells, VR velocities, NSP points)
    if (with_a) cp_async_wait_but_one();  // u_n(t)
    const double (*const pr)[NSP][3] = &prim_s[2 * cp];  // read from shared memory where used: 18 registers less
    const double ij0 = cell_ok ? inv_j[i] : 0.0, ij1 = cell_ok ? inv_j[i + 1] : 0.0;
#pragma unroll
    for (int r = 0; r < VR; ++r) {
      const int j = vg + kFG * r;
      const bool ok = cell_ok && j < nu;
      const double v = j < nu ? velo[j] : 0.0;
      const bool pos = v >= 0.0;  // heaviside delta, bgk_wave.jl:26
      const size_t base = (size_t)i + (size_t)ncell * j;
      const double s0 = v * ij0, s1 = v * ij1;  // v / J
      double f0[NSP], f1[NSP];
#pragma unroll
      for (int q = 0; q < NSP; ++q) { f0[q] = s0 * w[r][q].x; f1[q] = s1 * w[r][q].y; }  // :85-88
      const double fL0 = dotn<NSP>(f0, ops.ll), fR0 = dotn<NSP>(f0, ops.lr);              // interp_face! :99-101
      const double fL1 = dotn<NSP>(f1, ops.ll), fR1 = dotn<NSP>(f1, ops.lr);
      // upwind neighbour traces: the other cell of the pair, the adjacent lane's pair, or the halo cell
      double upL = __shfl_up_sync(0xffffffffu, fR1, 1, 4);    // right trace of cell i - 1
      double upR = __shfl_down_sync(0xffffffffu, fL0, 1, 4);  // left trace of cell i + 2
      if (edge_lo && pos) upL = halo_s[0][r][vg];   // written by this very thread before the barriers
      if (edge_hi && !pos) upR = halo_s[1][r][vg];
      // f_interaction - own trace at the left / right face (:103-107): zero on the downwind side
      const double cl0 = pos ? upL - fL0 : 0.0, cr0 = pos ? 0.0 : fL1 - fR0;
      const double cl1 = pos ? fR0 - fL1 : 0.0, cr1 = pos ? 0.0 : upR - fR1;
      double2 o[NSP];
#pragma unroll
      for (int p = 0; p < NSP; ++p) {
        const double c0 = v - pr[0][p][1], c1 = v - pr[1][p][1];
        const double M0 = pr[0][p][0] * exp_neg(-pr[0][p][2] * (c0 * c0));  // maxwellian
        const double M1 = pr[1][p][0] * exp_neg(-pr[1][p][2] * (c1 * c1));
        const double d0 = -(dotn<NSP>(f0, &ops.lpdm[p * FRB_NSPMAX]) + cl0 * ops.dgl[p] + cr0 * ops.dgr[p]) +
                          (M0 - w[r][p].x) * inv_tau;
        const double d1 = -(dotn<NSP>(f1, &ops.lpdm[p * FRB_NSPMAX]) + cl1 * ops.dgl[p] + cr1 * ops.dgr[p]) +
                          (M1 - w[r][p].y) * inv_tau;
        if (st.rhs_only) {
          o[p] = make_double2(d0, d1);
        } else {
          double r0 = st.nested ? st.cb * (w[r][p].x + st.cdt * d0) : st.cb * w[r][p].x + st.cdt * d0;
          double r1 = st.nested ? st.cb * (w[r][p].y + st.cdt * d1) : st.cb * w[r][p].y + st.cdt * d1;
          if (with_a) {
            const double2 an = stage_a[(r * NSP + p) * 256 + tid];  // u_n: landed with the tile (cp.async)
            r0 = st.ca * an.x + r0;
            r1 = st.ca * an.y + r1;
          }
          o[p] = make_double2(r0, r1);
        }
      }
      if (ok) {
#pragma unroll
        for (int p = 0; p < NSP; ++p) *reinterpret_cast<double2 *>(out + base + vs * p) = o[p];
      }
    }
    if (with_a) prefetch(ua, stage_a, next < ntiles ? next : ntiles);  // u_n of the next tile: its slots are free now
  }
}

template <int NSP, int VR>
int launch_fused_vr(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  const bool with_a = st.use_a && !st.rhs_only;
  const size_t smem = sizeof(double2) * VR * NSP * 256 * (with_a ? 2 : 1);
  static std::atomic<unsigned long long> attr_done{0};
  const unsigned long long dev_bit = 1ull << (p->ctx->device & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    FRB_CUDA(cudaFuncSetAttribute(bgk1d_fused_kernel<NSP, VR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double2) * VR * NSP * 256 * 2)));
    // two CTAs of a 24-B stage need 2 x 107 KB: without the carve-out hint the driver may leave room for one only
    FRB_CUDA(cudaFuncSetAttribute(bgk1d_fused_kernel<NSP, VR>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  const int ntiles = (p->ncell + kFC - 1) / kFC;
  const int slots = p->ctx->sm_count * FRB_BGK_MINB;  // persistent: one CTA per resident slot
  const dim3 grd(ntiles < slots ? ntiles : slots), blk(256);
  bgk1d_fused_kernel<NSP, VR><<<grd, blk, smem, p->ctx->stream>>>(u, ua, out, p->J, p->velo, p->weights, p->ncell,
                                                                   p->nu, 1.0 / p->tau, p->ops, st, p->bgk_model, p->a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "bgk1d_fused_kernel", __FILE__, __LINE__);
  return 1;
}

template <int NSP>
int launch_fused(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  switch ((p->nu + kFG - 1) / kFG) {
    case 1: return launch_fused_vr<NSP, 1>(p, u, ua, out, st);
    case 2: return launch_fused_vr<NSP, 2>(p, u, ua, out, st);
    case 3: return launch_fused_vr<NSP, 3>(p, u, ua, out, st);
    default: return launch_fused_vr<NSP, 4>(p, u, ua, out, st);
  }
}

// the one-pass kernel needs 16-byte aligned cell pairs (even ncell), the velocity grid in registers
// (nu <= 256) and deg 1..3; everything else runs the two-launch form
bool fused_ok(frb_prob_t p, const double *u, const double *ua, const double *out) {
  // Default: the two-launch form.  Measured at cfg4 (B200, 16-B / 24-B stage): two launches 45.7 / 56.3 us, one pass
  // 46-50 / 64.5 us -- the state is L2-resident and neither form is memory-bound, so
  // saving a pass over u buys nothing while the register tile costs occupancy.  frb_set_kernel(FRB_KERNEL_BGK_ONE_PASS) or
  // FRB_BGK_ONE_PASS=1 select the one-pass kernel (same results; tests run both).
  static const bool on = getenv("FRB_BGK_ONE_PASS") != nullptr && getenv("FRB_BGK_TWO_PASS") == nullptr;
  if (!(on || p->kernel_kind == FRB_KERNEL_BGK_ONE_PASS) || p->ncell % 2 || p->nu > kFG * kFusedMaxVR || p->nsp < 2 || p->nsp > 4) return false;
  return u != out && (((uintptr_t)u | (uintptr_t)out | (uintptr_t)(ua ? ua : u)) & 15) == 0;
}

}  // namespace

int frb_launch_bgk1d(frb_prob_t p, const double *u, const double *ua, double *out, FrbStage st) {
  if (fused_ok(p, u, ua, out)) {
    switch (p->nsp) {
      case 2: return launch_fused<2>(p, u, ua, out, st);
      case 3: return launch_fused<3>(p, u, ua, out, st);
      default: return launch_fused<4>(p, u, ua, out, st);
    }
  }
  dim3 blk(128), g1((p->ncell + 31) / 32, p->nsp), g2((p->ncell + 127) / 128, p->nu);
  // (a two-cells-per-lane form of the moments kernel with 16-byte loads was measured: 39.9 / 51.6 us per stage
  // against 39.5 / 50.0 with this one -- dropped)
  bgk_moments_kernel<<<g1, 32 * kMomGroups, 0, p->ctx->stream>>>(u, p->prim, p->ncell, p->nu, p->nsp, p->velo, p->weights,
                                                                  p->bgk_model, p->a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "bgk_moments_kernel", __FILE__, __LINE__);
  const double it = 1.0 / p->tau;
#ifndef FRB_BGK_NO_PAIR
  // two cells per thread where the pairs are 16-byte aligned: even ncell, aligned buffers, deg 1..3
  if (p->ncell % 2 == 0 && p->nsp <= 4 && u != out &&
      (((uintptr_t)u | (uintptr_t)out | (uintptr_t)(ua ? ua : u) | (uintptr_t)p->prim | (uintptr_t)p->J) & 15) == 0) {
    dim3 gp((p->ncell / 2 + 127) / 128, p->nu);
    switch (p->nsp) {
      case 2: bgk1d_pair_kernel<2><<<gp, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
      case 3: bgk1d_pair_kernel<3><<<gp, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
      default: bgk1d_pair_kernel<4><<<gp, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it, p->ops, st); break;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return frb_cuda_fail(e, "bgk1d_pair_kernel", __FILE__, __LINE__);
    return 2;
  }
#endif
  // VPT velocities per thread where it pays: measured at cfg4 (us per stage, VPT = 1 / 2 / 4): 16-B stage 45.2 /
  // 43.5 / 43.8, 24-B stage 55.9 / 59.6-60.9 / 62.5 -- the u_n loads of a 24-B stage want the occupancy back
  const bool multi = p->nu % FRB_BGK_VPT == 0 && !(st.use_a && !st.rhs_only);
  if (multi) g2.y = p->nu / FRB_BGK_VPT;
#define FRB_BGK_MAIN(N)                                                                                              \
  case N:                                                                                                            \
    if (multi)                                                                                                       \
      bgk1d_kernel<N, FRB_BGK_VPT><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell,     \
                                                                   p->nu, it, p->ops, st);                          \
    else                                                                                                             \
      bgk1d_kernel<N, 1><<<g2, blk, 0, p->ctx->stream>>>(u, ua, out, p->prim, p->J, p->velo, p->ncell, p->nu, it,    \
                                                         p->ops, st);                                               \
    break;
  switch (p->nsp) {
    FRB_BGK_MAIN(2)
    FRB_BGK_MAIN(3)
    FRB_BGK_MAIN(4)
    FRB_BGK_MAIN(5)
    FRB_BGK_MAIN(6)
    FRB_BGK_MAIN(7)
    FRB_BGK_MAIN(8)
    default: frb_set_error("bgk1d: deg must be in 1..7"); return FRB_ERR_ARG;
  }
#undef FRB_BGK_MAIN
  e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "bgk1d_kernel", __FILE__, __LINE__);
  return 2;
}
