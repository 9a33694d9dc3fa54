// Microbenchmark: DMMA (mma.sync.m8n8k4.f64) throughput per SM as a function of resident warps and independent
// accumulators per warp, operands in registers or re-loaded from shared memory every step.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_probe tools/micro/dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int U, bool SMEM>
__global__ void k(int iters, double* sink, long long* cyc) {
  extern __shared__ double sm[];
  double c[U][2];
#pragma unroll
  for (int u = 0; u < U; ++u) { c[u][0] = threadIdx.x * 1e-9 + u; c[u][1] = c[u][0] + 0.5; }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  double a = 1.0000001, b = 1e-9;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (SMEM) {
        // 2 A fragments + 2 B fragments feed 4 DMMAs (the 2 x 2 register blocking of the contractions)
#pragma unroll
        for (int u = 0; u < U; u += 4) {
          const double a0 = sm[((r * 4 + q) * 100 + g + 16 * u) & 4095], a1 = sm[((r * 4 + q) * 100 + g + 8 + 16 * u) & 4095];
          const double b0 = sm[(2048 + (r * 4 + q) * 100 + g + 16 * u) & 4095], b1 = sm[(2048 + (r * 4 + q) * 100 + g + 8 + 16 * u) & 4095];
          dmma(c[u][0], c[u][1], a0, b0);
          if (U > 1) dmma(c[(u + 1) % U][0], c[(u + 1) % U][1], a0, b1);
          if (U > 2) dmma(c[(u + 2) % U][0], c[(u + 2) % U][1], a1, b0);
          if (U > 3) dmma(c[(u + 3) % U][0], c[(u + 3) % U][1], a1, b1);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) dmma(c[u][0], c[u][1], a, b);
      }
    }
  }
  const long long t1 = clock64();
  double r = 0.0;
#pragma unroll
  for (int u = 0; u < U; ++u) r += c[u][0] + c[u][1];
  if (r == 12345.678) sink[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int U, bool SMEM>
void run(int warps, int ctas_per_sm) {
  double* sink; long long* cyc; cudaMalloc(&sink, 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  const size_t smem = ctas_per_sm == 1 ? 120 * 1024 : (ctas_per_sm == 2 ? 100 * 1024 : 40 * 1024);
  cudaFuncSetAttribute(k<U, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<U, SMEM><<<148 * ctas_per_sm, warps * 32, smem>>>(10, sink, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<U, SMEM><<<148 * ctas_per_sm, warps * 32, smem>>>(iters, sink, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const int per_iter = SMEM ? 8 * ((U + 3) / 4) * (U >= 4 ? 4 : U) : 8 * U;
  const double dm = (double)iters * per_iter * warps * ctas_per_sm;   // DMMAs per SM
  printf("U=%d smem=%d warps/CTA=%2d CTAs/SM=%d: %8.3f ms  %6.2f TFLOP/s  %.3f DMMA/clk/SM (clock64 %.3f)  err=%s\n", U, (int)SMEM, warps, ctas_per_sm, ms,
         dm * 148 * 512 / (ms * 1e-3) / 1e12, dm / (ms * 1e-3 * 1.965e9), dm / (double)c, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink); cudaFree(cyc);
}
int main() {
  for (int w : {4, 8, 16, 32}) { run<1, false>(w, 1); run<2, false>(w, 1); run<4, false>(w, 1); run<8, false>(w, 1); run<12, false>(w, 1); }
  for (int w : {8}) { run<4, false>(w, 2); run<8, false>(w, 2); run<12, false>(w, 2); }
  for (int w : {8, 16}) { run<4, true>(w, 1); run<8, true>(w, 1); run<12, true>(w, 1); run<4, true>(w, 2); run<8, true>(w, 2); run<12, true>(w, 2); }
  return 0;
}
