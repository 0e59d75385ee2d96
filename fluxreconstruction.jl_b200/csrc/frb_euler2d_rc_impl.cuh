// The row-chunk stage kernel of the 2-D Euler path, templated on the common flux (see frb_euler2d_rc.cu for the
// description).  Included by one translation unit per flux.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "frb_internal.cuh"
#include "frb_physics.cuh"
#include "frb_euler2d_passes.cuh"
#include "frb_ptx.cuh"
#include "frb_rc.cuh"

// HALO (template parameter of the kernel): the slab-parallel code -- halo ring, mailbox polls, peer stores -- is
// compiled into its own instantiations; single-domain launches run kernels without it (11 registers less, ~1 %)
#define RC_HALO_ACTIVE(P) (HALO)
// kernel experiment switches (scripts/build_variants.py): ring depth and CTAs per SM of the 16-B stage at p3
#ifndef FRB_RC_NBUF16
#define FRB_RC_NBUF16 3
#endif
#ifndef FRB_RC_MINB16
#define FRB_RC_MINB16 3
#endif

namespace frbrc {


using namespace frbptx;
using namespace frbpass;

struct RcParams {
  const double *u;   // stage input (RC)
  const double *ua;  // u_n (RC; may alias out)
  double *out;
  RcGeom g;
  int rows_per_seg;
  double gamma, ca, cb;
  RcHalo h;  // slab-parallel path: the exchange with the neighbouring ranks (frb_rc.cuh)
};

// thread 0 of a boundary CTA: wait until the neighbour has raised mailbox[side] to `want` (its boundary row of
// the stage that produced this stage's input is in the local slot).  Never hangs the box: after 5 s the time-out
// is recorded in mailbox[2] (frb_halo_check_timeout -> FRB_ERR_PEER) and the launch carries on.
static __device__ __noinline__ void halo_poll(unsigned long long *mailbox, int side, unsigned long long want) {
  if (!want) return;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(mailbox + side) < want) {
    if (globaltimer_ns() - t0 > 5000000000ull) {
      mailbox[2] = want;
      break;
    }
    __nanosleep(100);
  }
}

// thread 0, after the CTA's stores of a boundary row (peer stores included) and a block barrier: count the strip;
// the last strip of the launch raises the neighbour's mailbox (fence / atomic / fence / release: the peer stores
// of every strip are ordered before the flag)
__device__ __forceinline__ void halo_raise(unsigned int *count, int ns, unsigned long long *flag,
                                           unsigned long long epoch) {
  const unsigned int old = atomicAdd(count, 1u);
  if ((old + 1u) % (unsigned)ns == 0u) {
    __threadfence_system();
    st_release_sys(flag, epoch);
  }
}

template <int NSP, int NBUF, bool USEA>
struct SmemRc {
  static constexpr int kTile = 4 * NSP * NSP * 32;  // doubles
  alignas(128) double tile[NBUF][kTile];
  alignas(128) double un[USEA ? kTile : 16];
  alignas(128) double xd[kTile];                 // x-pass output: cb*u (+) x derivative + correction
  alignas(128) double xrp[2 * NSP * NSP * 32];   // v_y and p at the points
  alignas(8) uint64_t bar[NBUF + 1];
};

// CB1: the stage has cb == 1 (u' = u + dt L(u), first stage of every scheme): no multiply
template <int NSP, bool USEA, bool SAMEJ, int MINB, bool CB1, int FLUX, bool HALO>
__global__ void __launch_bounds__(NSP * 32, MINB) euler2d_rc_kernel(RcParams P, MarchOps ops) {
  constexpr int NBUF = USEA ? 2 : FRB_RC_NBUF16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = SmemRc<NSP, NBUF, USEA>;
  SM &S = *reinterpret_cast<SM *>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
  constexpr int kTile = SM::kTile;
  constexpr uint32_t kTileBytes = kTile * sizeof(double);
  uint64_t *const bar_un = &S.bar[NBUF];

  const RcGeom &g = P.g;
  const int lane = threadIdx.x & 31;
  const int t = threadIdx.x >> 5;  // point row l in the x pass, point column k in the y pass
  const int s = blockIdx.x;
  const int i = kRcOwn * s + lane;  // element column of this lane
  // row segment of this CTA.  Slab-parallel launches: rows 1 and ny are one-row segments of their own, first in
  // launch order, so that the rows the neighbours wait for leave in the first microseconds of the launch
  int ja, jb;
  if (!RC_HALO_ACTIVE(P)) {
    ja = 1 + blockIdx.y * P.rows_per_seg;
    jb = min(g.ny, ja + P.rows_per_seg - 1);
  } else if (blockIdx.y == 0) {
    ja = jb = 1;
  } else if (blockIdx.y == 1) {
    ja = jb = g.ny > 1 ? g.ny : 0;
    if (g.ny < 2) return;
  } else {
    ja = 2 + (blockIdx.y - 2) * P.rows_per_seg;
    jb = min(g.ny - 1, ja + P.rows_per_seg - 1);
  }
  if (ja > jb) return;
  // Odd segments march DOWNWARDS (rows jb .. ja): two neighbouring segments then meet at their common boundary --
  // both at the start or both at the end of their march -- and the halo rows each reads of the other arrive while
  // the owner streams them: the second read is an L2 hit instead of a DRAM read a whole CTA lifetime later.
#ifdef FRB_RC_UP_ONLY
  const bool down = false;
#else
  const bool down = ((RC_HALO_ACTIVE(P) ? (int)blockIdx.y - 2 : (int)blockIdx.y) & 1) != 0 && jb > ja;
#endif
  const int row0 = down ? jb + 1 : ja - 1, rstep = down ? -1 : 1;  // tile q holds row row0 + rstep * q
  const int ntiles = jb - ja + 3;  // rows ja-1 .. jb+1; tile q in buffer q % NBUF
  const bool owner = lane >= 1 && lane <= kRcOwn && i <= g.nx;
  const int own = owner ? 1 : 0;
  // the copy of my column in the neighbouring chunk (lane 1 -> lane 31 of strip s-1, lane 30 ->
  // lane 0 of strip s+1)
  const int dup = (owner && ((lane == 1 && s > 0) || (lane == kRcOwn && s < g.ns - 1))) ? 1 : 0;
  const ptrdiff_t dup_off = lane == 1 ? (ptrdiff_t)(kRcOwn - g.chunk) : (ptrdiff_t)(g.chunk - kRcOwn);
  const double gamma = P.gamma, gm1 = gamma - 1.0;
  const size_t strip_off = (size_t)s * g.chunk;

  // Where row r of the input comes from: the array itself, or -- rows 0 / ny+1 of a slab whose neighbour stores
  // its boundary row into this rank's halo ring -- the local slot of the stage that produced the input.
  auto is_slot = [&](int r) -> bool {
    return RC_HALO_ACTIVE(P) && ((r == 0 && P.h.src_lo != nullptr) || (r == g.ny + 1 && P.h.src_hi != nullptr));
  };
  // thread 0: one bulk copy of the strip's chunk of row r into ring buffer b
  auto issue_bulk = [&](int b, int r) {
    mbar_expect_tx(&S.bar[b], kTileBytes);
    bulk_load(S.tile[b], P.u + (size_t)r * g.row + strip_off, kTileBytes, &S.bar[b]);
  };
  // ALL threads (CTA-uniform call): a halo slot row.  Thread 0 waits for the neighbour's flag, then the CTA copies
  // the chunk with L2 loads (ld.global.cg: written by the peer GPU during this launch, never through L1) and
  // completes the tile's barrier phase by hand.
  auto fetch_slot = [&](int b, int r) {
    if (threadIdx.x == 0) halo_poll(P.h.mailbox, r == 0 ? 0 : 1, r == 0 ? P.h.wait_lo : P.h.wait_hi);
    __syncthreads();
    const double2 *src = reinterpret_cast<const double2 *>((r == 0 ? P.h.src_lo : P.h.src_hi) + strip_off);
    double2 *dst = reinterpret_cast<double2 *>(S.tile[b]);
    for (int q = threadIdx.x; q < kTile / 2; q += NSP * 32) dst[q] = __ldcg(src + q);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // a bulk copy may refill this buffer later
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(&S.bar[b]);
  };

  if (threadIdx.x == 0) {
    for (int b = 0; b < NBUF + 1; ++b) mbar_init(&S.bar[b], 1);
    mbar_fence_init();
  }
  __syncthreads();
  {
    const int npre = ntiles < NBUF ? ntiles : NBUF;
    if (threadIdx.x == 0) {
      for (int q = 0; q < npre; ++q)
        if (!is_slot(row0 + rstep * q)) issue_bulk(q, row0 + rstep * q);
      if (USEA) {
        mbar_expect_tx(bar_un, kTileBytes);
        bulk_load(S.un, P.ua + (size_t)(row0 + rstep) * g.row + strip_off, kTileBytes, bar_un);
      }
    }
    for (int q = 0; q < npre; ++q)  // after the bulk copies are in flight: these may have to wait for a neighbour
      if (is_slot(row0 + rstep * q)) fetch_slot(q, row0 + rstep * q);
  }

  // per-thread views of a tile (= of a chunk): row view (x pass, l = t), column view (y pass, k = t)
  const int offx = 32 * NSP * t + lane;  // + 32*(k + NSP*NSP*m)
  const int offy = 32 * t + lane;        // + 32*NSP*(l + NSP*m)
  double *const xdx = S.xd + offx;
  const double *const xdy = S.xd + offy;
  double *const xrpx = S.xrp + offx;
  const double *const xrpy = S.xrp + offy;

  // trace operators along the march: the current row's trace on the face towards the next row, the next row's on
  // the same face (upwards: lr / ll; downwards: ll / lr)
  const double *const lown = down ? ops.ll : ops.lr, *const lnb = down ? ops.lr : ops.ll;
  // ---- prologue: common flux on the face behind the first row (tiles 0 and 1), carried along the march
  double hc[4];
  {
    mbar_wait(&S.bar[0], 0);
    mbar_wait(&S.bar[1], 0);
    double u0[4];
    col_trace<NSP>(S.tile[0] + offy, lown, u0);  // tile 0 seen from its successor: same roles as in the loop
    face_flux_y_dir<NSP, FLUX>(u0, S.tile[1] + offy, lnb, down, gamma, gm1, hc);
  }
  {
    // tile 0 is dead after the prologue: refill its buffer with tile NBUF
    __syncthreads();
    if (ntiles > NBUF) {
      if (is_slot(row0 + rstep * NBUF)) fetch_slot(0, row0 + rstep * NBUF);
      else if (threadIdx.x == 0) issue_bulk(0, row0 + rstep * NBUF);
    }
  }

  for (int q = 1; q <= ntiles - 2; ++q) {  // tile q = row j
    const int j = row0 + rstep * q;
    const int buf = q % NBUF, nbuf = (q + 1) % NBUF;
    const double *const Ux = S.tile[0] + buf * kTile + offx;
    const double *const Uy = S.tile[0] + buf * kTile + offy;
    // tile q was already waited on as the "next" tile of step q-1 (or in the prologue)

    // -------------------------------------------------------------- x pass: row l = t
    x_pass<NSP, CB1, FLUX>(Ux, xdx, xrpx, ops, P.cb, gamma, gm1);
    __syncthreads();  // (A) xd / xrp of this row visible

    // -------------------------------------------------------------- y pass: column k = t
    {
      double g4[NSP][4];  // G at the column's points, [l][m]
      double uT[4], hf[4];
      y_fluxes_dir<NSP>(Uy, xrpy, lown, g4, uT);
      // the face ahead of row j: common flux between this row's trace and the next row's (already in the ring)
      mbar_wait(&S.bar[nbuf], ((q + 1) / NBUF) & 1);
      face_flux_y_dir<NSP, FLUX>(uT, S.tile[0] + nbuf * kTile + offy, lnb, down, gamma, gm1, hf);
      if (USEA) mbar_wait(bar_un, (q - 1) & 1);  // u_n row j (requested one row step ago)
      // four independent FMA chains per variable, stored as soon as they retire; the chunk offset
      // of a value equals its tile offset
      double *const po = P.out + (size_t)j * g.row + strip_off + offy;
      // slab-parallel: row 1 / row ny also goes into the neighbour's halo ring (peer memory), warp-uniform
      const bool to_lo = RC_HALO_ACTIVE(P) && P.h.dst_lo != nullptr && j == 1;
      const bool to_hi = RC_HALO_ACTIVE(P) && P.h.dst_hi != nullptr && j == g.ny;
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        double v[NSP];
#pragma unroll
        for (int l = 0; l < NSP; ++l) {
          // bottom / top face of the row: the carried flux and the fresh one, by direction
          double d = y_value<NSP, SAMEJ>(xdy[32 * NSP * (l + NSP * m)], g4, down ? hf[m] : hc[m],
                                         down ? hc[m] : hf[m], ops, l, m);
          if (USEA) d = fma(P.ca, S.un[offy + 32 * NSP * (l + NSP * m)], d);
          v[l] = d;
        }
#pragma unroll
        for (int l = 0; l < NSP; ++l) {
          st_cs_if(po + 32 * NSP * (l + NSP * m), v[l], own);
          st_cs_if(po + 32 * NSP * (l + NSP * m) + dup_off, v[l], dup);
        }
        if (to_lo) {
          double *const pp = P.h.dst_lo + strip_off + offy;
#pragma unroll
          for (int l = 0; l < NSP; ++l) {
            st_if(pp + 32 * NSP * (l + NSP * m), v[l], own);
            st_if(pp + 32 * NSP * (l + NSP * m) + dup_off, v[l], dup);
          }
        }
        if (to_hi) {
          double *const pp = P.h.dst_hi + strip_off + offy;
#pragma unroll
          for (int l = 0; l < NSP; ++l) {
            st_if(pp + 32 * NSP * (l + NSP * m), v[l], own);
            st_if(pp + 32 * NSP * (l + NSP * m) + dup_off, v[l], dup);
          }
        }
        hc[m] = hf[m];
      }
    }
    __syncthreads();  // (B) every read of tile[buf], un, xd, xrp is done; every store of row j is issued

    if (RC_HALO_ACTIVE(P) && threadIdx.x == 0 && (j == 1 || j == g.ny)) {
      // this strip's part of a boundary row is out (locally and in the neighbour's ring), and the halo row next
      // to it has been consumed: count it, the last strip raises the neighbour's flag
      __threadfence_system();
      if (j == 1) halo_raise(P.h.count + 0, g.ns, P.h.flag_lo, P.h.epoch);
      if (j == g.ny) halo_raise(P.h.count + 1, g.ns, P.h.flag_hi, P.h.epoch);
    }
    if (threadIdx.x == 0 && USEA && q + 1 <= ntiles - 2) {
      mbar_expect_tx(bar_un, kTileBytes);
      bulk_load(S.un, P.ua + (size_t)(j + rstep) * g.row + strip_off, kTileBytes, bar_un);
    }
    if (q + NBUF < ntiles) {
      const int r = row0 + rstep * (q + NBUF);
      if (is_slot(r)) fetch_slot(buf, r);
      else if (threadIdx.x == 0) issue_bulk(buf, r);
    }
  }

}

static inline int env_int_rc(const char *name, int dflt) {
  const char *s = getenv(name);
  return s && *s ? atoi(s) : dflt;
}

static inline int rc_rows_per_seg(frb_prob_t p, const RcGeom &g, int ctas_per_sm, bool usea) {
  // Short segments on purpose.  The kernel is fastest while the CTAs that run at the same time work on
  // ADJACENT chunks -- all strips of a row are one contiguous 1.1 MB run -- which is the state the CTAs
  // of a wave start in and drift out of.  Per-CTA globaltimer stamps: 2.5 us per row step in the first
  // wave whatever the segment length, ~3.6 us in later waves of 64-row segments.  Measured at 2048^2, p3
  // on B200 (16-B / 24-B stage, ms): 6 rows 0.91 / 1.085, 12: 0.88 / 1.13, 32: 1.01 / 1.16, 64: 1.11 /
  // 1.22; the extra halo rows (2 per segment) are mostly L2 hits.  What does NOT explain it (each was
  // built and measured): the strip-edge duplicate stores (no change without them), a tail wave alone,
  // start-up lock-step (per-CTA jitter: no change), TLB reach (strip-major chunk order: slower).
  // FRB_MARCH_ROWS overrides.
  static const int forced = env_int_rc("FRB_MARCH_ROWS", 0);  // read once (thread-safe static initialisation)
  if (forced > 0) return forced < g.ny ? forced : g.ny;
  const int slots = p->ctx->sm_count * ctas_per_sm;
  int nseg = (g.ny + (usea ? 5 : 11)) / (usea ? 6 : 12);
  // small meshes: enough CTAs for every SM slot if the row count allows it
  while ((long)g.ns * nseg < slots && (g.ny + nseg) / (nseg + 1) >= 4) ++nseg;
  return (g.ny + nseg - 1) / nseg;
}

template <int NSP, bool USEA, bool SAMEJ, int MINB, bool CB1, int FLUX, bool HALO>
int launch_rc(frb_prob_t p, RcParams rp, const MarchOps &mo) {
  constexpr int NBUF = USEA ? 2 : FRB_RC_NBUF16;
  rp.rows_per_seg = rc_rows_per_seg(p, rp.g, MINB, USEA);
  int segs = (rp.g.ny + rp.rows_per_seg - 1) / rp.rows_per_seg;
  if (rp.h.active) {  // rows 1 and ny as one-row segments (blockIdx.y 0, 1), rows 2 .. ny-1 in normal segments
    const int inner = rp.g.ny - 2;
    segs = (rp.g.ny >= 2 ? 2 : 1) + (inner > 0 ? (inner + rp.rows_per_seg - 1) / rp.rows_per_seg : 0);
  }
  const size_t smem = sizeof(SmemRc<NSP, NBUF, USEA>) + 128;
  // per device: the attribute belongs to the device's copy of the kernel.  Atomic bit mask: problems on different
  // handles may launch from different host threads (frb200.h: re-entrant across handles); setting the attribute
  // twice is harmless, missing it is not.
  static std::atomic<unsigned long long> attr_done{0};
  const unsigned long long dev_bit = 1ull << (p->ctx->device & 63);
  if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
    FRB_CUDA(cudaFuncSetAttribute(euler2d_rc_kernel<NSP, USEA, SAMEJ, MINB, CB1, FLUX, HALO>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FRB_CUDA(cudaFuncSetAttribute(euler2d_rc_kernel<NSP, USEA, SAMEJ, MINB, CB1, FLUX, HALO>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done.fetch_or(dev_bit, std::memory_order_release);
  }
  dim3 grd(rp.g.ns, segs), blk(NSP * 32);
  euler2d_rc_kernel<NSP, USEA, SAMEJ, MINB, CB1, FLUX, HALO><<<grd, blk, smem, p->ctx->stream>>>(rp, mo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return frb_cuda_fail(e, "euler2d_rc_kernel", __FILE__, __LINE__);
  return 1;
}

template <int NSP, int MINB, int FLUX, bool HALO>
int dispatch_rc(frb_prob_t p, const RcParams &rp, const MarchOps &mo, bool usea, bool samej) {
  if (usea)
    return samej ? launch_rc<NSP, true, true, MINB, false, FLUX, HALO>(p, rp, mo)
                 : launch_rc<NSP, true, false, MINB, false, FLUX, HALO>(p, rp, mo);
  constexpr int M16 = NSP == 4 ? FRB_RC_MINB16 : MINB;  // 16-B stages: no u_n tile, room for another CTA
  if (rp.cb == 1.0)
    return samej ? launch_rc<NSP, false, true, M16, true, FLUX, HALO>(p, rp, mo)
                 : launch_rc<NSP, false, false, M16, true, FLUX, HALO>(p, rp, mo);
  return samej ? launch_rc<NSP, false, true, M16, false, FLUX, HALO>(p, rp, mo)
               : launch_rc<NSP, false, false, M16, false, FLUX, HALO>(p, rp, mo);
}

// one translation unit per common flux instantiates this (frb_euler2d_rc.cu: HLL, frb_euler2d_rc_lf.cu,
// frb_euler2d_rc_roe.cu): the three builds run in parallel
template <int FLUX>
int dispatch_rc_flux(frb_prob_t p, const RcParams &rp, const MarchOps &mo, bool usea, bool samej) {
  if (rp.h.active) {
    if (p->nsp == 4) return dispatch_rc<4, 3, FLUX, true>(p, rp, mo, usea, samej);
    return dispatch_rc<3, 4, FLUX, true>(p, rp, mo, usea, samej);
  }
  if (p->nsp == 4) return dispatch_rc<4, 3, FLUX, false>(p, rp, mo, usea, samej);
  return dispatch_rc<3, 4, FLUX, false>(p, rp, mo, usea, samej);
}


}  // namespace frbrc

// per-flux entry points (defined in frb_euler2d_rc.cu / _lf.cu / _roe.cu)
int frb_rc_dispatch_hll(frb_prob_t p, const frbrc::RcParams &rp, const MarchOps &mo, bool usea, bool samej);
int frb_rc_dispatch_lf(frb_prob_t p, const frbrc::RcParams &rp, const MarchOps &mo, bool usea, bool samej);
int frb_rc_dispatch_roe(frb_prob_t p, const frbrc::RcParams &rp, const MarchOps &mo, bool usea, bool samej);
