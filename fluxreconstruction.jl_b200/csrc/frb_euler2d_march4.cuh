// Variant 4 of the row-marching 2-D Euler stage kernel (included inside the anonymous namespace
// of frb_euler2d_march.cu, after the helpers and MarchParams).
//
// Differences to euler2d_march_kernel:
//  * the top-face term of row j is applied one step later, when the bottom face of row j+1 is
//    computed (its HLL needs row j's top trace, carried in registers) -> a step touches only ONE
//    tile, so the TMA ring is 2 deep (32 KB instead of 48 KB);
//  * the freed 16 KB hold u_n: each thread cp.async's the 16 u_n values it will need for row j
//    into private shared-memory slots at the end of step j and reads them back at the end of step
//    j+1 -- a full step of latency cover, no registers, no scoreboard slot held by the loads.
// Carried per thread between steps: acc[NSP][4] (= cb*u + cxs*dux + cys*(duy without the top-face
// term)) and the top trace uT[4].
template <int NSP>
struct Smem4 {
  static constexpr int kPlanes = 4 * NSP * NSP;
  static constexpr int kTile = kPlanes * 32;  // doubles
  alignas(128) double tile[2][kTile];
  alignas(128) double xd[kTile];
  alignas(128) double xrp[2 * NSP * NSP * 32];
  alignas(128) double un[kTile];  // u_n staging: slot (t + NSP*(l + NSP*m))*32 + lane, thread-private
  alignas(8) uint64_t bar[2];
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NSP, int MINB>
__global__ void __launch_bounds__(NSP * 32, MINB)
euler2d_march4_kernel(const __grid_constant__ CUtensorMap tmap, MarchParams P, FrbOps ops) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = Smem4<NSP>;
  SM &S = *reinterpret_cast<SM *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  constexpr int kTile = SM::kTile;
  constexpr uint32_t kTileBytes = kTile * sizeof(double);

  const int lane = threadIdx.x & 31;
  const int t = threadIdx.x >> 5;
  const int i = blockIdx.x * kOwn + lane;
  const int ja = P.jlo + blockIdx.y * P.rows_per_seg;
  const int jb = min(P.jhi, ja + P.rows_per_seg - 1);
  if (ja > P.jhi) return;
  const int ntiles = jb - ja + 3;  // rows ja-1 .. jb+1; tile q holds row ja-1+q in buffer q & 1
  const size_t NXG = P.nx + 2, NE = NXG * (size_t)(P.ny + 2);
  const bool owner = lane >= 1 && lane <= kOwn && i <= P.nx;
  const double gamma = P.gamma, gm1 = gamma - 1.0;
  const int c0 = blockIdx.x * kOwn;

  if (threadIdx.x == 0) {
    mbar_init(&S.bar[0], 1);
    mbar_init(&S.bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < 2; ++q) {
      mbar_expect_tx(&S.bar[q], kTileBytes);
      tma_load_3d(S.tile[q], &tmap, &S.bar[q], c0, ja - 1 + q, 0);
    }
  }

  const int offx = 32 * NSP * t + lane;  // row view   + 32*(k + NSP*NSP*m)
  const int offy = 32 * t + lane;        // column view + 32*NSP*(l + NSP*m)
  double *const xdx = S.xd + offx;
  const double *const xdy = S.xd + offy;
  double *const xrpx = S.xrp + offx;
  const double *const xrpy = S.xrp + offy;
  double *const uns = S.un + offy;
  const size_t pstep = NE * NSP;
  const size_t goff = i + NE * (size_t)t;

  double acc[NSP][4];  // carried: everything of row j-1 except the top-face term and u_n
  double uTp[4];       // carried: top trace of row j-1
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    uTp[m] = 0.0;
#pragma unroll
    for (int l = 0; l < NSP; ++l) acc[l][m] = 0.0;
  }

  for (int q = 0; q < ntiles; ++q) {  // tile q = row j
    const int j = ja - 1 + q;
    const int buf = q & 1;
    const bool full = q >= 1 && q <= ntiles - 2;  // ja <= j <= jb: this row is updated
    const bool finish = q >= 2;                    // row j-1 is an updated row: complete it now
    const double *const Ux = S.tile[0] + buf * kTile + offx;
    const double *const Uy = S.tile[0] + buf * kTile + offy;
    mbar_wait(&S.bar[buf], (q >> 1) & 1);

    if (full) {
      // ------------------------------------------------------------ x pass: row l = t
      double w[NSP][4], f[NSP][4];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < NSP; ++k) w[k][m] = Ux[32 * (k + NSP * NSP * m)];
#pragma unroll
      for (int k = 0; k < NSP; ++k) {
        double rr = frb::rcp_fast(w[k][0]);
        double vx = w[k][1] * rr, vy = w[k][2] * rr;
        double p = gm1 * fma(-0.5, fma(w[k][1], vx, w[k][2] * vy), w[k][3]);
        f[k][0] = w[k][1];
        f[k][1] = fma(w[k][1], vx, p);
        f[k][2] = w[k][1] * vy;
        f[k][3] = (w[k][3] + p) * vx;
        xrpx[32 * k] = rr;
        xrpx[32 * (NSP * NSP + k)] = p;
      }
      double uL[4], uR[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double a = w[0][m] * ops.ll[0], b = w[0][m] * ops.lr[0];
#pragma unroll
        for (int q2 = 1; q2 < NSP; ++q2) {
          a = fma(w[q2][m], ops.ll[q2], a);
          b = fma(w[q2][m], ops.lr[q2], b);
        }
        uL[m] = a; uR[m] = b;
      }
      double n0 = __shfl_up_sync(0xffffffffu, uR[0], 1), n1 = __shfl_up_sync(0xffffffffu, uR[1], 1);
      double n2 = __shfl_up_sync(0xffffffffu, uR[2], 1), n3 = __shfl_up_sync(0xffffffffu, uR[3], 1);
      frb::Flux4 hl = frb::hll4_fast(n0, n1, n2, n3, uL[0], uL[1], uL[2], uL[3], gamma, gm1);
      const double hL[4] = {hl.f0, hl.f1, hl.f2, hl.f3};
      const double hR[4] = {__shfl_down_sync(0xffffffffu, hl.f0, 1), __shfl_down_sync(0xffffffffu, hl.f1, 1),
                            __shfl_down_sync(0xffffffffu, hl.f2, 1), __shfl_down_sync(0xffffffffu, hl.f3, 1)};
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int k = 0; k < NSP; ++k) {
          double d = f[0][m] * ops.dmod[k * FRB_NSPMAX];
#pragma unroll
          for (int q2 = 1; q2 < NSP; ++q2) d = fma(f[q2][m], ops.dmod[k * FRB_NSPMAX + q2], d);
          d = fma(hL[m], ops.dgl[k], d);
          d = fma(hR[m], ops.dgr[k], d);
          xdx[32 * (k + NSP * NSP * m)] = d;
        }
    }
    __syncthreads();  // (A) xd / xrp of this row visible

    // -------------------------------------------------------------- y pass: column k = t
    {
      double w[NSP][4];  // [l][m]
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int l = 0; l < NSP; ++l) w[l][m] = Uy[32 * NSP * (l + NSP * m)];
      double uB[4], uT[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double a = w[0][m] * ops.ll[0], b = w[0][m] * ops.lr[0];
#pragma unroll
        for (int q2 = 1; q2 < NSP; ++q2) {
          a = fma(w[q2][m], ops.ll[q2], a);
          b = fma(w[q2][m], ops.lr[q2], b);
        }
        uB[m] = a; uT[m] = b;
      }
      double hb[4] = {0.0, 0.0, 0.0, 0.0};
      if (q >= 1) {  // face between rows j-1 and j  (euler2d_wave.jl:75-82)
        frb::Flux4 h = frb::hll4_y_fast(uTp[0], uTp[1], uTp[2], uTp[3], uB[0], uB[1], uB[2], uB[3], gamma, gm1);
        hb[0] = h.f0; hb[1] = h.f1; hb[2] = h.f2; hb[3] = h.f3;
      }
      const size_t grow = goff + NXG * (size_t)j;
      if (finish) {
        // row j-1: add its top-face term and u_n, store u'
        if (P.use_a) cp_async_wait_all();
        double *po = P.out + grow - NXG;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const double z = P.cys * hb[m];
#pragma unroll
          for (int l = 0; l < NSP; ++l, po += pstep) {
            double v = fma(z, ops.dgr[l], acc[l][m]);
            if (P.use_a) v = fma(P.ca, uns[32 * NSP * (l + NSP * m)], v);
            if (owner) __stcs(po, v);
          }
        }
        // slab-parallel path: forward the finished row to the neighbour rank's halo row
        const int jf = j - 1;
        if ((jf == 1 && P.peer_lo) || (jf == P.ny && P.peer_hi)) {
          if (owner) {
            const double *src = P.out + grow - NXG;
            if (jf == 1 && P.peer_lo) {
              const size_t NEl = NXG * (size_t)(P.nyl_lo + 2);
              double *dst = P.peer_lo + i + NXG * (size_t)(P.nyl_lo + 1) + NEl * (size_t)t;
              for (int c = 0; c < 4 * NSP; ++c) dst[NEl * NSP * c] = src[pstep * c];
            }
            if (jf == P.ny && P.peer_hi) {
              const size_t NEh = NXG * (size_t)(P.nyl_hi + 2);
              double *dst = P.peer_hi + i + NEh * (size_t)t;
              for (int c = 0; c < 4 * NSP; ++c) dst[NEh * NSP * c] = src[pstep * c];
            }
          }
        }
      }
      if (full) {
        if (P.use_a && owner) {  // u_n of THIS row, needed one step from now
          const double *pa = P.ua + grow;
#pragma unroll
          for (int c = 0; c < 4 * NSP; ++c, pa += pstep) cp_async8(uns + 32 * NSP * c, pa);
          cp_async_commit();
        }
        double g[NSP][4];
#pragma unroll
        for (int l = 0; l < NSP; ++l) {
          double rr = xrpy[32 * NSP * l];
          double p = xrpy[32 * NSP * (NSP + l)];
          double vy = w[l][2] * rr;
          g[l][0] = w[l][2];
          g[l][1] = w[l][1] * vy;
          g[l][2] = fma(w[l][2], vy, p);
          g[l][3] = (w[l][3] + p) * vy;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
#pragma unroll
          for (int l = 0; l < NSP; ++l) {
            double d = g[0][m] * ops.dmod[l * FRB_NSPMAX];
#pragma unroll
            for (int q2 = 1; q2 < NSP; ++q2) d = fma(g[q2][m], ops.dmod[l * FRB_NSPMAX + q2], d);
            d = fma(hb[m], ops.dgl[l], d);
            double dx = xdy[32 * NSP * (l + NSP * m)];
            acc[l][m] = fma(P.cys, d, fma(P.cxs, dx, P.cb * w[l][m]));
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 4; ++m) uTp[m] = uT[m];
    }
    __syncthreads();  // (B) every read of tile[buf], xd, xrp is done

    if (threadIdx.x == 0 && q + 2 < ntiles) {
      mbar_expect_tx(&S.bar[buf], kTileBytes);
      tma_load_3d(S.tile[buf], &tmap, &S.bar[buf], c0, ja - 1 + q + 2, 0);
    }
  }
}
