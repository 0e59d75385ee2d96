// Fused 2-D Euler residual + RK stage: the row-marching TMA kernel (the roofline path).
//
// One HBM pass per stage: read u (once, plus a 1-row/1-column halo that lives in L2),
// optionally read u_n, write u'.  Everything between -- point fluxes, the four trace
// interpolations, the HLL common fluxes on all four faces, the lpdm derivative and the
// dgl/dgr correction, division by the Jacobian and the stage combination -- happens on
// chip (reference: dudt! of example/euler2d_wave.jl:35-107 + the OrdinaryDiffEq stage
// axpys around it).
//
// Decomposition
//   grid.x = strips of 30 owned elements in x (a warp's 32 lanes are elements
//            i0-1 .. i0+30: lanes 0 and 31 are x-halo lanes that only feed traces),
//   grid.y = row segments; the CTA marches j = ja .. jb through its segment.
//   CTA    = NSP warps; warp t handles row  l = t in the x pass and column k = t in the
//            y pass of every element of the strip.
// Per row step
//   TMA (cp.async.bulk.tensor.3d) lands the [4*NSP^2 planes] x [32 elements] block of a row
//   in shared memory (NBUF-deep mbarrier ring), so both the row view and the column view
//   of an element are plain conflict-free LDS -- the transpose is free.
//   x pass: F at the row's points, x traces, left-face HLL with the left neighbour's trace
//           by __shfl_up, right-face flux by __shfl_down, d/dr + correction -> smem.
//   y pass: G at the column's points (1/rho and p come from the x pass through smem),
//           y traces, top-face HLL against the bottom trace of row j+1 (already resident
//           in the ring), d/ds + corrections; the bottom-face flux is the top-face flux
//           carried in registers from row j-1.  u' is written fused with the RK combination.
#include <cuda.h>

#include <cstdlib>
#include <map>

#include "frb_internal.cuh"
#include "frb_physics.cuh"
#include "frb_euler2d_passes.cuh"
#include "frb_ptx.cuh"

namespace {

constexpr int kOwn = 30;  // owned elements per strip (32 lanes - 2 halo lanes)

using namespace frbptx;
using namespace frbpass;

struct MarchParams {
  const double *ua;  // u_n (may alias out)
  double *out;
  int nx, ny;
  int rows_per_seg;
  int jlo, jhi;  // rows to update (1..ny for a whole-mesh launch; a sub-range for the pipelined path)
  double gamma;
  double ca, cb;
  int use_a;
  // slab-parallel path: the first / last owned row is also stored straight into the halo row
  // of the rank below / above (peer memory over NVLink); NULL when there is no such neighbour
  double *peer_lo, *peer_hi;
  int nyl_lo, nyl_hi;
  int pfdist;  // rows of look-ahead of the u_n L2 prefetch
};

template <int NSP, int NBUF>
struct Smem {
  static constexpr int kPlanes = 4 * NSP * NSP;
  static constexpr int kTile = kPlanes * 32;  // doubles
  alignas(128) double tile[NBUF][kTile];
  alignas(128) double xd[kTile];              // x-pass derivative + correction (no 1/Jx yet)
  alignas(128) double xrp[2 * NSP * NSP * 32];  // v_y and p at the points
  alignas(8) uint64_t bar[NBUF];
};

// MINB = resident CTAs per SM the register budget is compiled for
template <int NSP, int NBUF, int MINB, bool PREFETCH, bool USEA, bool SAMEJ, bool COPYONLY = false>
__global__ void __launch_bounds__(NSP * 32, MINB)
euler2d_march_kernel(const __grid_constant__ CUtensorMap tmap, MarchParams P, MarchOps ops) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = Smem<NSP, NBUF>;
  // keep the pointer in the shared window (LDS/STS, not generic LD/ST): align by offset
  SM &S = *reinterpret_cast<SM *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  constexpr int kPlanes = SM::kPlanes;
  constexpr int kTile = SM::kTile;
  constexpr uint32_t kTileBytes = kPlanes * 32 * sizeof(double);

  const int lane = threadIdx.x & 31;
  const int t = threadIdx.x >> 5;  // row index in the x pass, column index in the y pass
  const int i = blockIdx.x * kOwn + lane;  // element column of this lane (0 = ghost)
  const int ja = P.jlo + blockIdx.y * P.rows_per_seg;
  const int jb = min(P.jhi, ja + P.rows_per_seg - 1);
  if (ja > P.jhi) return;
  const int ntiles = jb - ja + 3;  // rows ja-1 .. jb+1; tile q holds row ja-1+q in buffer q % NBUF
  const size_t NXG = P.nx + 2, NE = NXG * (size_t)(P.ny + 2);
  const bool owner = lane >= 1 && lane <= kOwn && i <= P.nx;
  const int own = owner ? 1 : 0;
  const double gamma = P.gamma, gm1 = gamma - 1.0;
  const int c0 = blockIdx.x * kOwn;

  if (threadIdx.x == 0) {
    for (int b = 0; b < NBUF; ++b) mbar_init(&S.bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int npre = ntiles < NBUF ? ntiles : NBUF;
    for (int q = 0; q < npre; ++q) {
      mbar_expect_tx(&S.bar[q], kTileBytes);
      tma_load_3d(S.tile[q], &tmap, &S.bar[q], c0, ja - 1 + q, 0);
    }
  }

  // per-thread views: row view (x pass, l = t) and column view (y pass, k = t)
  const int offx = 32 * NSP * t + lane;  // + 32*(k + NSP*NSP*m)
  const int offy = 32 * t + lane;        // + 32*NSP*(l + NSP*m)
  double *const xdx = S.xd + offx;
  const double *const xdy = S.xd + offy;
  double *const xrpx = S.xrp + offx;
  const double *const xrpy = S.xrp + offy;
  // global plane walk: plane(t, l, m) = t + NSP*(l + NSP*m)  ->  base + NE*t, step NE*NSP
  const size_t pstep = NE * NSP;
  // non-owner lanes never touch global memory; give them the address of an owner lane so that
  // every address the loop forms is in bounds
  const size_t goff = (owner ? i : (size_t)(blockIdx.x * kOwn + 1)) + NE * (size_t)t;

  if (PREFETCH && USEA && owner) {  // u_n rows of the first pfdist steps -> L2
    for (int r = 0; r < P.pfdist && ja + r <= jb; ++r) {
      const double *pa = P.ua + goff + NXG * (size_t)(ja + r);
      for (int c = 0; c < 4 * NSP; ++c, pa += pstep) prefetch_l2(pa);
    }
  }

  // ---- prologue: common flux on the bottom face of row ja from tiles 0 (row ja-1) and 1
  double hb[4];
  {
    mbar_wait(&S.bar[0], 0);
    mbar_wait(&S.bar[1 % NBUF], 0);
    double uT[4];
    col_trace<NSP>(S.tile[0] + offy, ops.lr, uT);
    face_flux_y<NSP>(uT, S.tile[1 % NBUF] + offy, ops, gamma, gm1, hb);
  }
  {
    // tile 0 is dead after the prologue: refill its buffer with tile NBUF
    __syncthreads();
    if (threadIdx.x == 0 && ntiles > NBUF) {
      mbar_expect_tx(&S.bar[0], kTileBytes);
      tma_load_3d(S.tile[0], &tmap, &S.bar[0], c0, ja - 1 + NBUF, 0);
    }
  }

  for (int q = 1; q <= ntiles - 2; ++q) {  // tile q = row j
    const int j = ja - 1 + q;
    const int buf = q % NBUF, nbuf = (q + 1) % NBUF;
    const double *const Ux = S.tile[0] + buf * kTile + offx;
    const double *const Uy = S.tile[0] + buf * kTile + offy;
    size_t grow = goff + NXG * (size_t)j;
    // opaque to the optimiser: otherwise it keeps one 64-bit induction variable per plane
    // (16 planes x {load, store} -> 60+ registers of loop-carried offsets)
    asm volatile("" : "+l"(grow));
    // pull the u_n row needed pfdist steps from now into L2 (one bulk prefetch by one thread);
    // the loads at the top of that step's y pass then hit L2 instead of waiting on DRAM
    if (PREFETCH && USEA && owner && j + P.pfdist <= jb) {
      const double *pa = P.ua + grow + NXG * (size_t)P.pfdist;
#pragma unroll
      for (int c = 0; c < 4 * NSP; ++c, pa += pstep) prefetch_l2(pa);
    }
    // tile q was already waited on as the "next" tile of step q-1 (or in the prologue)

    if (COPYONLY) {  // measurement aid: the memory-access skeleton of the kernel without the math
      __syncthreads();
      mbar_wait(&S.bar[nbuf], ((q + 1) / NBUF) & 1);
      const double *pa = P.ua + grow;
      double *po = P.out + grow;
#pragma unroll
      for (int c = 0; c < 4 * NSP; ++c, pa += pstep, po += pstep) {
        double v = P.cb * Uy[32 * NSP * c];
        if (USEA && owner) v = fma(P.ca, __ldcs(pa), v);
        if (owner) __stcs(po, v);
      }
      __syncthreads();
      if (threadIdx.x == 0 && q + NBUF < ntiles) {
        mbar_expect_tx(&S.bar[buf], kTileBytes);
        tma_load_3d(S.tile[buf], &tmap, &S.bar[buf], c0, ja - 1 + q + NBUF, 0);
      }
      continue;
    }
    // -------------------------------------------------------------- x pass: row l = t
    x_pass<NSP, false>(Ux, xdx, xrpx, ops, P.cb, gamma, gm1);
    __syncthreads();  // (A) xd / xrp of this row visible

    // -------------------------------------------------------------- y pass: column k = t
    {
      double un[NSP][4];
      if (USEA) {
        const double *pa = P.ua + grow;
#pragma unroll
        for (int m = 0; m < 4; ++m)
#pragma unroll
          for (int l = 0; l < NSP; ++l, pa += pstep) un[l][m] = __ldcs(pa);
      }
      double g[NSP][4];  // G at the column's points, [l][m]
      double uT[4], ht[4];
      y_fluxes<NSP>(Uy, xrpy, ops, g, uT);
      // top face of row j: HLL between this row's top trace and row j+1's bottom trace
      mbar_wait(&S.bar[nbuf], ((q + 1) / NBUF) & 1);
      face_flux_y<NSP>(uT, S.tile[0] + nbuf * kTile + offy, ops, gamma, gm1, ht);
      // four independent FMA chains per variable, stored as soon as they retire
      double *po = P.out + grow;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double v[NSP];
#pragma unroll
        for (int l = 0; l < NSP; ++l) {
          double d = y_value<NSP, SAMEJ>(xdy[32 * NSP * (l + NSP * m)], g, hb[m], ht[m], ops, l, m);
          if (USEA) d = fma(P.ca, un[l][m], d);
          v[l] = d;
        }
#pragma unroll
        for (int l = 0; l < NSP; ++l, po += pstep) st_cs_if(po, v[l], own);
        hb[m] = ht[m];
      }
    }
    __syncthreads();  // (B) every read of tile[buf], xd, xrp is done

    if (threadIdx.x == 0 && q + NBUF < ntiles) {
      mbar_expect_tx(&S.bar[buf], kTileBytes);
      tma_load_3d(S.tile[buf], &tmap, &S.bar[buf], c0, ja - 1 + q + NBUF, 0);
    }
  }
  // slab-parallel path: the first / last owned row also goes straight into the halo row of the
  // rank below / above (peer memory over NVLink).  Each thread forwards the values it stored
  // itself (program order, L2-hot), once per segment that owns row 1 / row ny.
  if (owner && ((ja == 1 && P.peer_lo) || (jb == P.ny && P.peer_hi))) {
    const size_t go = i + NE * (size_t)t;
    if (ja == 1 && P.peer_lo) {
      const double *src = P.out + go + NXG;
      const size_t NEl = NXG * (size_t)(P.nyl_lo + 2);
      double *dst = P.peer_lo + i + NXG * (size_t)(P.nyl_lo + 1) + NEl * (size_t)t;
      for (int c = 0; c < 4 * NSP; ++c) dst[NEl * NSP * c] = src[pstep * c];
    }
    if (jb == P.ny && P.peer_hi) {
      const double *src = P.out + go + NXG * (size_t)P.ny;
      const size_t NEh = NXG * (size_t)(P.nyl_hi + 2);
      double *dst = P.peer_hi + i + NEh * (size_t)t;
      for (int c = 0; c < 4 * NSP; ++c) dst[NEh * NSP * c] = src[pstep * c];
    }
  }
}



// ---- host side: tensor-map cache and launch ---------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct MapCache {
  std::map<const void *, CUtensorMap> maps;
};

int env_int(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

}  // namespace

bool frb_euler2d_march_supported(frb_prob_t p) {
  // TMA needs 16-byte global strides: (nx+2)*8 bytes per row -> nx even
  return p->kind == K_EULER2D && !p->curv_iJ && (p->nsp == 4 || p->nsp == 3) && (p->nx % 2 == 0);
}

void frb_march_release(frb_prob_t p) {
  delete static_cast<MapCache *>(p->tmaps);
  p->tmaps = nullptr;
}

static int march_rows_per_seg(frb_prob_t p, int ctas_per_sm) {
  // Segments of about 32 rows: every segment re-reads two halo rows (6 % extra reads at 32), and
  // the grid (strips x segments) is then many waves of SMs x resident CTAs, so the tail is short.
  // Measured at 2048^2, p3 on B200: 16..32 rows are within 1 % of each other, 108 rows (3 waves)
  // is 3-4 % slower.  FRB_MARCH_ROWS overrides.
  int forced = env_int("FRB_MARCH_ROWS", 0);
  if (forced > 0) return forced < p->ny ? forced : p->ny;
  const int strips = (p->nx + kOwn - 1) / kOwn;
  const int slots = p->ctx->sm_count * ctas_per_sm;
  int nseg = (p->ny + 31) / 32;
  // small meshes: make sure every SM slot gets a CTA if the row count allows it
  while ((long)strips * nseg < slots && (p->ny + nseg) / (nseg + 1) >= 4) ++nseg;
  return (p->ny + nseg - 1) / nseg;
}

template <int NSP, int NBUF, int MINB, bool PREFETCH, bool USEA, bool SAMEJ, bool COPYONLY = false>
static int launch_march(frb_prob_t p, const CUtensorMap &map, MarchParams mp, const MarchOps &mo) {
  mp.rows_per_seg = march_rows_per_seg(p, MINB);
  const int strips = (p->nx + kOwn - 1) / kOwn;
  const int segs = (mp.jhi - mp.jlo + 1 + mp.rows_per_seg - 1) / mp.rows_per_seg;
  const size_t smem = sizeof(Smem<NSP, NBUF>) + 128;
  static unsigned long long attr_done = 0;  // per device: the attribute belongs to the device's copy of the kernel
  const unsigned long long dev_bit = 1ull << (p->ctx->device & 63);
  if (!(attr_done & dev_bit)) {
    FRB_CUDA(cudaFuncSetAttribute(euler2d_march_kernel<NSP, NBUF, MINB, PREFETCH, USEA, SAMEJ, COPYONLY>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FRB_CUDA(cudaFuncSetAttribute(euler2d_march_kernel<NSP, NBUF, MINB, PREFETCH, USEA, SAMEJ, COPYONLY>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done |= dev_bit;
  }
  dim3 grd(strips, segs), blk(NSP * 32);
  euler2d_march_kernel<NSP, NBUF, MINB, PREFETCH, USEA, SAMEJ, COPYONLY><<<grd, blk, smem, p->ctx->stream>>>(map, mp, mo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "euler2d_march_kernel", __FILE__, __LINE__);
  return 1;
}

// TMA descriptor of one state buffer: dims (x, y, plane), box = 32 elements x 1 row x all planes
static int march_tensor_map(frb_prob_t p, const double *u, const CUtensorMap **out) {
  if (!p->tmaps) p->tmaps = new MapCache();
  MapCache *mc = static_cast<MapCache *>(p->tmaps);
  auto it = mc->maps.find(u);
  if (it == mc->maps.end()) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
      frb_set_error("cuTensorMapEncodeTiled not available from the driver");
      return FRB_ERR_CUDA;
    }
    const cuuint64_t NXG = p->nx + 2, NYG = p->ny + 2, NPL = 4 * p->nsp * p->nsp;
    cuuint64_t gdim[3] = {NXG, NYG, NPL};
    cuuint64_t gstr[2] = {NXG * 8, NXG * NYG * 8};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)NPL};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(u), gdim, gstr,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      frb_set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
      return FRB_ERR_CUDA;
    }
    it = mc->maps.emplace(u, m).first;
  }
  *out = &it->second;
  return 0;
}

int frb_launch_euler2d_march(frb_prob_t p, const double *u, const double *ua, double *out,
                             FrbStage st) {
  if (!frb_euler2d_march_supported(p)) {
    frb_set_error("marching kernel needs deg 2 or 3 and even nx");
    return FRB_ERR_ARG;
  }
  const CUtensorMap *tmp = nullptr;
  if (int rc = march_tensor_map(p, u, &tmp)) return rc;
  MarchParams mp;
  mp.ua = ua;
  mp.out = out;
  mp.nx = p->nx;
  mp.ny = p->ny;
  mp.rows_per_seg = 0;
  mp.jlo = p->row_lo > 0 ? p->row_lo : 1;
  mp.jhi = p->row_hi > 0 ? p->row_hi : p->ny;
  mp.pfdist = env_int("FRB_MARCH_PFDIST", 0);
  mp.gamma = p->gamma;
  frb_halo_stage_targets(p, out, &mp.peer_lo, &mp.peer_hi, &mp.nyl_lo, &mp.nyl_hi);
  double cxs, cys;  // -cdt/Jx, -cdt/Jy  (L = -(dF/dr / Jx + dG/ds / Jy))
  if (st.rhs_only) {
    mp.ca = 0.0; mp.cb = 0.0; mp.use_a = 0;
    cxs = -1.0 / p->Jx; cys = -1.0 / p->Jy;
  } else {
    const double cdt = st.nested ? st.cb * st.cdt : st.cdt;
    mp.ca = st.ca; mp.cb = st.cb; mp.use_a = st.use_a;
    cxs = -cdt / p->Jx; cys = -cdt / p->Jy;
  }
  MarchOps mo;
  const int n = p->nsp;
  for (int k = 0; k < 4; ++k) {
    const bool in = k < n;
    mo.ll[k] = in ? p->ops.ll[k] : 0.0;
    mo.lr[k] = in ? p->ops.lr[k] : 0.0;
    mo.glx[k] = in ? cxs * p->ops.dgl[k] : 0.0;
    mo.grx[k] = in ? cxs * p->ops.dgr[k] : 0.0;
    mo.gly[k] = in ? cys * p->ops.dgl[k] : 0.0;
    mo.gry[k] = in ? cys * p->ops.dgr[k] : 0.0;
    for (int q = 0; q < 4; ++q) {
      const double d = (in && q < n) ? p->ops.dmod[k * FRB_NSPMAX + q] : 0.0;
      mo.dmx[k * 4 + q] = cxs * d;
      mo.dmy[k * 4 + q] = cys * d;
    }
  }
  const bool pf = env_int("FRB_MARCH_PREFETCH", 1) != 0;  // L2 prefetch of the u_n row (+10 % on 24-B stages)
  const bool samej = cxs == cys;  // one operator table set: no uniform-register spills
#define FRB_MARCH_GO(NSP, NBUF, MINB, PF, UA, ...)                                         \
  (samej ? launch_march<NSP, NBUF, MINB, PF, UA, true, ##__VA_ARGS__>(p, *tmp, mp, mo)     \
         : launch_march<NSP, NBUF, MINB, PF, UA, false, ##__VA_ARGS__>(p, *tmp, mp, mo))
  if (p->nsp == 4) {  // 72 KB smem -> 3 CTAs/SM
    if (env_int("FRB_MARCH_COPYONLY", 0))
      return mp.use_a ? launch_march<4, 3, 3, true, true, true, true>(p, *tmp, mp, mo)
                      : launch_march<4, 3, 3, true, false, true, true>(p, *tmp, mp, mo);
    if (mp.use_a) return pf ? FRB_MARCH_GO(4, 3, 3, true, true) : FRB_MARCH_GO(4, 3, 3, false, true);
    return FRB_MARCH_GO(4, 3, 3, false, false);
  }
  return mp.use_a ? FRB_MARCH_GO(3, 3, 4, false, true) : FRB_MARCH_GO(3, 3, 4, false, false);
#undef FRB_MARCH_GO
}
