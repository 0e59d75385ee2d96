// Memory-skeleton experiment for a row-chunk ("RC") device layout of the 2-D Euler state:
//   chunk(j, s) = [64 planes][32 elements] doubles = 16 KB contiguous, chunks ordered [j][s].
// A CTA (4 warps) marches over rows of one strip: bulk-copies chunk(j, s) of u into a smem ring
// (cp.async.bulk, 1-D), optionally reads the u_n chunk, writes the out chunk.  No math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rc_skeleton rc_skeleton.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_store(void *dst, const void *src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}

constexpr int kChunk = 64 * 32;  // doubles
constexpr uint32_t kBytes = kChunk * 8;

// MODE: 0 = u -> out (16 B), 1 = u + u_n(LDG) -> out (24 B), 2 = u + u_n(bulk to smem) -> out,
//       3 = like 2 but the out chunk goes through smem + bulk store
template <int MODE, int NBUF>
__global__ void __launch_bounds__(128, 3) rc_kernel(const double *u, const double *un, double *out, int ns, int ny, int rows) {
  extern __shared__ __align__(128) unsigned char raw[];
  double *tile = reinterpret_cast<double *>(raw);                  // NBUF chunks
  double *tun = tile + NBUF * kChunk;                              // MODE>=2: 2 chunks
  double *tout = tun + (MODE >= 2 ? 2 * kChunk : 0);               // MODE==3: 1 chunk
  uint64_t *bar = reinterpret_cast<uint64_t *>(tout + (MODE == 3 ? kChunk : 0));  // NBUF + 2
  const int s = blockIdx.x, ja = blockIdx.y * rows, jb = min(ny, ja + rows);
  const int lane = threadIdx.x & 31, t = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int b = 0; b < NBUF + 2; ++b) mbar_init(&bar[b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int n = jb - ja;
  if (threadIdx.x == 0) {
    for (int q = 0; q < NBUF && q < n; ++q) { mbar_expect(&bar[q], kBytes); bulk_load(tile + q * kChunk, u + ((size_t)(ja + q) * ns + s) * kChunk, kBytes, &bar[q]); }
    if (MODE >= 2) for (int q = 0; q < 2 && q < n; ++q) { mbar_expect(&bar[NBUF + q], kBytes); bulk_load(tun + q * kChunk, un + ((size_t)(ja + q) * ns + s) * kChunk, kBytes, &bar[NBUF + q]); }
  }
  for (int q = 0; q < n; ++q) {
    const int buf = q % NBUF;
    const size_t base = ((size_t)(ja + q) * ns + s) * kChunk;
    mbar_wait(&bar[buf], (q / NBUF) & 1);
    if (MODE >= 2) mbar_wait(&bar[NBUF + (q & 1)], (q / 2) & 1);
    if (MODE == 3 && q > 0) { if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); __syncthreads(); }
    const double *T = tile + buf * kChunk + 32 * t + lane;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      double v = 0.25 * T[128 * c];
      if (MODE == 1) v = fma(0.75, __ldcs(un + base + 32 * t + lane + 128 * c), v);
      if (MODE >= 2) v = fma(0.75, tun[(q & 1) * kChunk + 32 * t + lane + 128 * c], v);
      if (MODE == 3) tout[32 * t + lane + 128 * c] = v;
      else __stcs(out + base + 32 * t + lane + 128 * c, v);
    }
    if (MODE == 3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      if (MODE == 3) { bulk_store(out + base, tout, kBytes); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
      if (q + NBUF < n) { mbar_expect(&bar[buf], kBytes); bulk_load(tile + buf * kChunk, u + ((size_t)(ja + q + NBUF) * ns + s) * kChunk, kBytes, &bar[buf]); }
      if (MODE >= 2 && q + 2 < n) { mbar_expect(&bar[NBUF + (q & 1)], kBytes); bulk_load(tun + (q & 1) * kChunk, un + ((size_t)(ja + q + 2) * ns + s) * kChunk, kBytes, &bar[NBUF + (q & 1)]); }
    }
  }
  if (MODE == 3 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, int NBUF>
float run(const double *u, const double *un, double *out, int ns, int ny, int rows, int iters) {
  size_t smem = (size_t)(NBUF + (MODE >= 2 ? 2 : 0) + (MODE == 3 ? 1 : 0)) * kBytes + 64;
  CK(cudaFuncSetAttribute(rc_kernel<MODE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 g(ns, (ny + rows - 1) / rows);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) rc_kernel<MODE, NBUF><<<g, 128, smem>>>(u, un, out, ns, ny, rows);
  CK(cudaGetLastError());
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) rc_kernel<MODE, NBUF><<<g, 128, smem>>>(u, un, out, ns, ny, rows);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / iters;
}

int main(int argc, char **argv) {
  const int ns = 64, ny = 2048;  // 2048 x 2048 elements, 64 strips of 32
  const size_t n = (size_t)ns * ny * kChunk;
  double *u, *un, *out;
  CK(cudaMalloc(&u, n * 8)); CK(cudaMalloc(&un, n * 8)); CK(cudaMalloc(&out, n * 8));
  CK(cudaMemset(u, 0, n * 8)); CK(cudaMemset(un, 0, n * 8));
  const double gb16 = 2.0 * n * 8 / 1e9, gb24 = 3.0 * n * 8 / 1e9;
  for (int rows : {16, 32, 64}) {
    float m0 = run<0, 3>(u, un, out, ns, ny, rows, 20);
    float m1 = run<1, 3>(u, un, out, ns, ny, rows, 20);
    float m2 = run<2, 2>(u, un, out, ns, ny, rows, 20);
    float m3 = run<3, 2>(u, un, out, ns, ny, rows, 20);
    printf("rows %3d | 16B: %.3f ms %.0f GB/s | 24B ldg: %.3f ms %.0f GB/s | 24B bulk-un (2+2 bufs): %.3f ms %.0f GB/s | 24B bulk in+out: %.3f ms %.0f GB/s\n",
           rows, m0, gb16 / m0 * 1e3, m1, gb24 / m1 * 1e3, m2, gb24 / m2 * 1e3, m3, gb24 / m3 * 1e3);
  }
  // reference: plain device-to-device copy of the same bytes
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemcpy(out, u, n * 8, cudaMemcpyDeviceToDevice);
  cudaEventRecord(a); for (int i = 0; i < 10; ++i) cudaMemcpyAsync(out, u, n * 8, cudaMemcpyDeviceToDevice); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b); printf("cudaMemcpy D2D: %.3f ms %.0f GB/s\n", ms / 10, gb16 / (ms / 10) * 1e3);
  return 0;
}
