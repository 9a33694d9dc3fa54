// What does ncu's l1tex__data_bank_conflicts_pipe_lsu_mem_shared count for a warp-wide 64-bit shared-memory load?
// One CTA; every warp issues `iters` x 8 LDS.64 whose 32 addresses are consecutive doubles (conflict-free by
// construction: two 128-byte wavefronts) or lie in a column of an odd-stride matrix (the layout of DESIGN.md 5.1).
// run: ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__sass_inst_executed_op_shared_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ./lds64_probe.bin
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int iters, double* sink) {
  __shared__ double sm[64 * 91 + 64];
  for (int i = threadIdx.x; i < 64 * 91 + 64; i += blockDim.x) sm[i] = i * 1e-3;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  double acc = 0.0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = MODE == 0 ? (i * 8 + u + lane) % 5000             // consecutive doubles (a row walk)
                                : ((lane + u) * 91 + (i % 60));          // a column of an odd-stride (91) matrix
      acc += sm[idx];
    }
  }
  if (acc == 1.2345) sink[0] = acc;
}
int main() {
  double* sink; cudaMalloc(&sink, 8);
  k<0><<<1, 256>>>(1000, sink);
  k<1><<<1, 256>>>(1000, sink);
  cudaDeviceSynchronize();
  printf("done: %s (8 warps x 1000 x 8 LDS.64 per kernel = 64000 warp-level loads each)\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
