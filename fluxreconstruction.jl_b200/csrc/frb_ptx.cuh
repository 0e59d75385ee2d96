// Inline-PTX helpers shared by the marching stage kernels (sm_100a): mbarrier, TMA tensor and
// bulk copies, predicated streaming stores; and the operator block those kernels consume.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace frbptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 3-D tiled tensor copy global -> shared (TMA, SASS UTMALDG)
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// contiguous bulk copy global -> shared (TMA, SASS UBLKCP); 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// one thread pulls a whole box into L2 (no smem destination)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap *map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// predicated streaming store: keeps the value computation out of a divergent branch (nvcc
// otherwise sinks each output's FMA chain into its own `if (owner)` block, serialising them)
__device__ __forceinline__ void st_cs_if(double *p, double v, int pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q st.global.cs.f64 [%0], %1;\n"
      "}\n" ::"l"(p),
      "d"(v), "r"(pred)
      : "memory");
}

// predicated plain store (peer memory: no cache hint)
__device__ __forceinline__ void st_if(double *p, double v, int pred) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.s32 q, %2, 0;\n"
      "@q st.global.f64 [%0], %1;\n"
      "}\n" ::"l"(p),
      "d"(v), "r"(pred)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

}  // namespace frbptx

// Operators as the marching kernels consume them: the derivative matrix with the flux-trace
// part of the correction folded in (FrbOps::dmod) and the correction vectors, pre-multiplied on
// the host by -cdt/Jx (x pass) and -cdt/Jy (y pass).  The stage is then a pure FMA chain:
//   x pass : xd = cb*u + sum_q dmx[k][q] F[q] + glx[k] Fhat_L + grx[k] Fhat_R
//   y pass : u' = xd   + sum_q dmy[l][q] G[q] + gly[l] Ghat_B + gry[l] Ghat_T  (+ ca*u_n)
// When Jx == Jy the kernels are instantiated to read only the x tables (half the constants:
// the second set overflows the uniform register file and costs spills).
struct MarchOps {
  double ll[4], lr[4];
  double dmx[16], glx[4], grx[4];
  double dmy[16], gly[4], gry[4];
};
