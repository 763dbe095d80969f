// Microbenchmark: legacy mma.sync m16n8k16 (HMMA) throughput on sm_100a, per SM, vs warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma tools/ubench_mma.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int NACC>
__global__ void __launch_bounds__(512) kern(float* out, int iters) {
  float d[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = threadIdx.x;
  uint32_t a0 = threadIdx.x * 0x3c003c00u, a1 = 0x3c003c00u, a2 = 0x38003800u, a3 = 0x3c003800u, b0 = 0x3c003c00u, b1 = 0x34003400u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      kern<14><<<148, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double mmas_per_sm = (double)(threads / 32) * iters * 14;
      if (rep == 2)
        printf("warps/SM %2d: %.3f ms  %.1f ns per MMA per SM  -> %.0f MAC/ns/SM (%.1f dense TFLOP/s chip)  [nominal clk %d kHz]\n",
               threads / 32, ms, ms * 1e6 / mmas_per_sm, mmas_per_sm * 2048 / (ms * 1e6), mmas_per_sm * 2048 * 2 * 148 / (ms * 1e9), clk);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
