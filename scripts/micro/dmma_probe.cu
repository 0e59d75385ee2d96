// Micro-benchmark: FP64 FMA pipe vs FP64 tensor pipe (mma.sync.m8n8k4.f64) on one B200, alone and
// interleaved in one warp.  Question behind it ("next"): the 16-B stage of the marching
// kernel is co-limited by FP64 issue; could the 4x4 derivative / trace products move to DMMA?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0 = DFMA only, 1 = DMMA only, 2 = both interleaved
__global__ void __launch_bounds__(256) probe(double *out, int iters, double seed) {
  double f[8], c0[2] = {0, 0}, c1[2] = {0, 0}, c2[2] = {0, 0}, c3[2] = {0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = seed * (threadIdx.x + i);
  double a = seed + threadIdx.x, b = seed * 0.5;
  for (int it = 0; it < iters; ++it) {
    if (MODE != 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], b, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE != 0) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(a), "d"(b));
    }
  }
  double s = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, double *out, int ctas) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  probe<MODE><<<ctas, 256>>>(out, 100, 1e-9);
  cudaEventRecord(e0);
  probe<MODE><<<ctas, 256>>>(out, iters, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double thr = (double)ctas * 256;
  const double fma_flops = MODE != 1 ? thr * iters * 16 * 2 : 0;                // 16 DFMA / thread / iter
  const double mma_flops = MODE != 0 ? (thr / 32) * iters * 4 * (8 * 8 * 4 * 2) : 0;  // 4 DMMA / warp / iter
  printf("%-18s %8.3f ms   DFMA %7.2f TF/s   DMMA %7.2f TF/s   sum %7.2f TF/s\n", name, ms,
         fma_flops / ms * 1e-9, mma_flops / ms * 1e-9, (fma_flops + mma_flops) / ms * 1e-9);
}

int main() {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  for (int per = 2; per <= 8; per *= 2) {
    printf("-- %d CTAs x 256 threads per SM\n", per);
    run<0>("DFMA only", out, sms * per);
    run<1>("DMMA only", out, sms * per);
    run<2>("DFMA + DMMA", out, sms * per);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
