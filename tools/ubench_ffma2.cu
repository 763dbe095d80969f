// Microbenchmark: fp32 FMA throughput with FFMA (scalar) vs FFMA2 (fma.rn.f32x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_ffma2 tools/ubench_ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) kern(float2* out, float a, float b, int iters) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 x = make_float2(a, a * 1.0001f), y = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {
        acc[i].x = fmaf(acc[i].x, x.x, y.x);
        acc[i].y = fmaf(acc[i].y, x.y, y.y);
      } else {
        unsigned long long d, aa = *reinterpret_cast<unsigned long long*>(&acc[i]);
        unsigned long long xx = *reinterpret_cast<const unsigned long long*>(&x), yy = *reinterpret_cast<const unsigned long long*>(&y);
        asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(aa), "l"(xx), "l"(yy));
        acc[i] = *reinterpret_cast<float2*>(&d);
      }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s.x += acc[i].x; s.y += acc[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float2* out;
  const int blocks = 148 * 8, iters = 4096;
  cudaMalloc(&out, blocks * 256 * sizeof(float2));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) kern<0><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
      else kern<1><<<blocks, 256>>>(out, 1.0001f, 0.5f, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double fma = (double)blocks * 256 * iters * 32;
      if (rep == 2) printf("%s: %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms / 1e9, 2 * fma / ms / 1e9);
    }
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
