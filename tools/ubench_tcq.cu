// How deep is the tcgen05.mma issue queue?  One thread issues 48 K=16 MMAs (M128 N256, cta_group::1) back to back and
// records clock64 after every issue; then a commit + wait.  Deltas ~0 = queued, ~128 = blocked behind the pipe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../genconvit_b200/csrc/tc_ptx.cuh"
using namespace gcv;
__global__ void __launch_bounds__(128, 1) kern(int n, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) w[i] = 0x3c003800u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(false, 128, n);
    const uint64_t da = umma_desc_kmajor<128>(base), db = umma_desc_kmajor<128>(base + 16384);
    long long t[49];
    t[0] = clock64();
#pragma unroll
    for (int i = 0; i < 48; ++i) {
      tc_mma(tmem, da, db, idesc, 1u);
      t[i + 1] = clock64();
    }
    tc_commit(smem_u32(&bar));
    const long long tc = clock64();
    mbar_wait(smem_u32(&bar), 0);
    const long long td = clock64();
    if (blockIdx.x == 0) {
      for (int i = 0; i < 48; ++i) out[i] = t[i + 1] - t[i];
      out[48] = tc - t[48];
      out[49] = td - t[0];
    }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}
int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 51 * 1024);
  for (int n : {256, 128, 64}) {
    for (int rep = 0; rep < 2; ++rep) { kern<<<1, 128, 51 * 1024>>>(n, d); cudaDeviceSynchronize(); }
    long long h[50]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d issue deltas:", n);
    for (int i = 0; i < 48; ++i) printf(" %lld", h[i]);
    printf("\n  commit %lld, total %lld clk for 48 MMAs (%.1f per MMA) err=%s\n", h[48], h[49], h[49] / 48.0, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
