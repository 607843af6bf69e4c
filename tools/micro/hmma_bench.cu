// Micro-benchmark (development aid): latency / issue rate of mma.sync.m16n8k16 (bf16, fp32 accumulate) on sm_100a,
// as used by the decode GEMV.  nvcc -gencode arch=compute_100a,code=sm_100a -o hmma_bench hmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters) {
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  uint32_t a = 0x3c003c00u + threadIdx.x, b = 0x3c003c00u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(d[c], a, a, a, a, b, b);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  if (s == 12345.f) sink[0] = s;
  if (threadIdx.x % 32 == 0) out[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 1024 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  long long h[64];
#define RUN(CH, WARPS) { k<CH><<<1, 32 * WARPS>>>(out, sink, iters); cudaDeviceSynchronize(); cudaMemcpy(h, out, WARPS * 8, cudaMemcpyDeviceToHost); \
    printf("chains %d warps %2d: %.1f cycles per MMA per warp; SM rate %.2f MMA/cycle\n", CH, WARPS, (double)h[0] / iters / CH, (double)WARPS * CH * iters / h[0]); }
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(2, 4) RUN(4, 4) RUN(2, 8) RUN(4, 8) RUN(2, 12)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
