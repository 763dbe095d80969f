// Microbenchmark: raw tcgen05.mma (kind::f16, SS operands, fp32 TMEM accumulate) issue/execute rate per SM on sm_100a,
// no loads, no epilogue: one thread per CTA (pair) issues `iters` batches of 4 K=16 MMAs on resident smem, one commit
// per batch (like a GEMM k-block), and the CTA clocks the whole run with clock64 and globaltimer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tcgen05 tools/ubench_tcgen05.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../genconvit_b200/csrc/tc_ptx.cuh"

using namespace gcv;

__device__ __forceinline__ uint64_t gtime() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// DUO: cta_group::2, cluster of 2; M per CTA 128.  n = MMA N.  garbage: fill smem with nonzero fp16 pattern.
template <bool DUO>
__global__ void __launch_bounds__(128, 1) kern(int n, int iters, int fill, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // A: 128 rows x 64 (16 KB); B: up to 256 rows x 64 (32 KB)
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x)
    w[i] = fill ? (0x3c003800u ^ (uint32_t)(i * 2654435761u) & 0x83ff83ffu) : 0u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if constexpr (DUO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool leader = !DUO || cluster_ctarank() == 0;
  long long c0 = 0, c1 = 0;
  uint64_t g0 = 0, g1 = 0;
  if (threadIdx.x == 0 && leader) {
    const uint32_t idesc = umma_idesc_f16(false, DUO ? 256 : 128, n);
    const uint32_t sa = base, sb = base + 16384;
    const int acc_stages = 512 / n;
    c0 = clock64();
    g0 = gtime();
    for (int it = 0; it < iters; ++it) {
      const uint32_t d = tmem + (uint32_t)((it % acc_stages) * n);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (DUO) tc_mma_2sm(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
        else tc_mma(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
      }
    }
    if constexpr (DUO) tc_commit_2sm(smem_u32(&bar));
    else tc_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    c1 = clock64();
    g1 = gtime();
    out[blockIdx.x * 2] = (unsigned long long)(c1 - c0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)(g1 - g0);
  } else if (threadIdx.x == 0 && DUO) {
    mbar_wait(smem_u32(&bar), 0);      // commit multicast arrives here too
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  if (threadIdx.x < 32) {
    if constexpr (DUO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// The GEMM's issue-loop structure without loads: ring of `stages` full/empty barriers, a relay thread standing in
// for the TMA producer (waits empty, arrives full), issuer waits full -> 4 MMAs -> commit(empty).
// mode 1: commit per batch only (no waits); 2: issuer waits its own commit `stages` batches back; 3: relay thread + fence
template <bool DUO>
__global__ void __launch_bounds__(128, 1) kern_ring(int n, int iters, int mode, int stages, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) w[i] = (0x3c003800u ^ (uint32_t)(i * 2654435761u) & 0x83ff83ffu);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if constexpr (DUO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool leader = !DUO || cluster_ctarank() == 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0 && (mode == 3 || mode == 4 || mode == 6)) {
    // relay ("producer"): in DUO mode only the leader's full barrier is used, but both CTAs run the relay like the GEMM
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
      if (leader) mbar_arrive(smem_u32(full_bar + stage));
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    const uint32_t idesc = umma_idesc_f16(false, DUO ? 256 : 128, n);
    const uint32_t sa = base, sb = base + 16384;
    const int acc_stages = 512 / n;
    int stage = 0; uint32_t phase = 0;
    const long long c0 = clock64();
    const uint64_t g0 = gtime();
    bool ready = false;                 // mode 4: result of the try_wait issued during the previous batch
    for (int it = 0; it < iters; ++it) {
      if (mode == 3) { mbar_wait(smem_u32(full_bar + stage), phase); tc_fence_after(); }
      else if (mode == 6) { mbar_wait(smem_u32(full_bar + stage), phase); }
      else if (mode == 4) { if (!ready) mbar_wait(smem_u32(full_bar + stage), phase); tc_fence_after(); }
      else if (mode == 2) mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
      else if (mode == 5 && (it & 1) == 0) mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
      const uint32_t d = tmem + (uint32_t)(((it >> 5) % acc_stages) * n);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if constexpr (DUO) tc_mma_2sm(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
        else tc_mma(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
        if (mode == 4 && k == 1) {
          const int ns = stage + 1 == stages ? 0 : stage + 1;
          ready = mbar_try_wait(smem_u32(full_bar + ns), stage + 1 == stages ? phase ^ 1 : phase);
        }
      }
      if constexpr (DUO) tc_commit_2sm(smem_u32(empty_bar + stage));
      else tc_commit(smem_u32(empty_bar + stage));
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if constexpr (DUO) tc_commit_2sm(smem_u32(&done_bar));
    else tc_commit(smem_u32(&done_bar));
    mbar_wait(smem_u32(&done_bar), 0);
    out[blockIdx.x * 2] = (unsigned long long)(clock64() - c0);
    out[blockIdx.x * 2 + 1] = (unsigned long long)(gtime() - g0);
  } else if (warp == 1 && lane == 0 && DUO) {
    mbar_wait(smem_u32(&done_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  if (threadIdx.x < 32) {
    if constexpr (DUO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// Same ring as mode 3, but the whole MMA warp runs the loop (warp-uniform control flow and operands) and only the
// tcgen05 instructions are predicated on elect.sync -- the compiler can then keep descriptors in uniform registers.
template <bool DUO>
__global__ void __launch_bounds__(128, 1) kern_ring_uni(int n, int iters, int stages, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[8], empty_bar[8], done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t* w = reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) w[i] = (0x3c003800u ^ (uint32_t)(i * 2654435761u) & 0x83ff83ffu);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) { mbar_init(smem_u32(full_bar + s), 1); mbar_init(smem_u32(empty_bar + s), 1); }
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    if constexpr (DUO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const bool leader = !DUO || cluster_ctarank() == 0;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
      if (leader && elect_one()) mbar_arrive(smem_u32(full_bar + stage));
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && leader) {
    const uint32_t idesc = umma_idesc_f16(false, DUO ? 256 : 128, n);
    const uint32_t sa = base, sb = base + 16384;
    const int acc_stages = 512 / n;
    int stage = 0; uint32_t phase = 0;
    const long long c0 = clock64();
    const uint64_t g0 = gtime();
    for (int it = 0; it < iters; ++it) {
      mbar_wait(smem_u32(full_bar + stage), phase);
      tc_fence_after();
      const uint32_t d = tmem + (uint32_t)(((it >> 5) % acc_stages) * n);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if constexpr (DUO) tc_mma_2sm(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
          else tc_mma(d, umma_desc_kmajor<128>(sa + k * 32), umma_desc_kmajor<128>(sb + k * 32), idesc, 1u);
        }
        if constexpr (DUO) tc_commit_2sm(smem_u32(empty_bar + stage));
        else tc_commit(smem_u32(empty_bar + stage));
      }
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) {
      if constexpr (DUO) tc_commit_2sm(smem_u32(&done_bar));
      else tc_commit(smem_u32(&done_bar));
    }
    mbar_wait(smem_u32(&done_bar), 0);
    if ((threadIdx.x & 31) == 0) {
      out[blockIdx.x * 2] = (unsigned long long)(clock64() - c0);
      out[blockIdx.x * 2 + 1] = (unsigned long long)(gtime() - g0);
    }
  } else if (warp == 1 && DUO) {
    mbar_wait(smem_u32(&done_bar), 0);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();
  if (threadIdx.x < 32) {
    if constexpr (DUO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <bool DUO>
void run_ring_uni(int n, int iters, int stages, unsigned long long* dout) {
  const int smem = 50 * 1024 + 1024, ctas = 148;
  cudaFuncSetAttribute(kern_ring_uni<DUO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (DUO) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kern_ring_uni<DUO>, n, iters, stages, dout);
    } else {
      kern_ring_uni<DUO><<<ctas, 128, smem>>>(n, iters, stages, dout);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[296];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  const double flop_sm = 2.0 * 128 * n * 64 * (double)iters;
  const double clk = (double)h[0], ns = (double)h[1];
  printf("ring-uniform %s N=%3d stages=%d: %.3f ms  %.0f FLOP/clk/SM  SM clock %.0f MHz  chip %.0f TFLOP/s err=%s\n",
         DUO ? "cta_group::2" : "cta_group::1", n, stages, best, flop_sm / clk, clk / ns * 1e3,
         flop_sm * ctas / (best * 1e9), cudaGetErrorString(cudaGetLastError()));
}

template <bool DUO>
void run_ring(int n, int iters, int mode, int stages, unsigned long long* dout) {
  const int smem = 50 * 1024 + 1024, ctas = 148;
  cudaFuncSetAttribute(kern_ring<DUO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (DUO) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kern_ring<DUO>, n, iters, mode, stages, dout);
    } else {
      kern_ring<DUO><<<ctas, 128, smem>>>(n, iters, mode, stages, dout);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[296];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  const double flop_sm = 2.0 * 128 * n * 64 * (double)iters;
  const double clk = (double)h[0], ns = (double)h[1];
  printf("ring %s N=%3d mode=%d stages=%d: %.3f ms  %.0f FLOP/clk/SM  SM clock %.0f MHz  chip %.0f TFLOP/s err=%s\n",
         DUO ? "cta_group::2" : "cta_group::1", n, mode, stages, best, flop_sm / clk, clk / ns * 1e3,
         flop_sm * ctas / (best * 1e9), cudaGetErrorString(cudaGetLastError()));
}

template <bool DUO>
void run(int n, int iters, int fill, int ctas, unsigned long long* dout) {
  const int smem = 50 * 1024 + 1024;
  cudaFuncSetAttribute(kern<DUO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemset(dout, 0, 148 * 2 * 8);
    cudaEventRecord(e0);
    if (DUO) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, kern<DUO>, n, iters, fill, dout);
    } else {
      kern<DUO><<<ctas, 128, smem>>>(n, iters, fill, dout);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  unsigned long long h[296];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  // per-SM work: 128 rows x n x 64 K per batch
  const double flop_sm = 2.0 * 128 * n * 64 * (double)iters;
  const int slot = 0;
  const double clk = (double)h[slot * 2], ns = (double)h[slot * 2 + 1];
  printf("%s N=%3d fill=%d ctas=%3d: %.3f ms  %.0f FLOP/clk/SM  SM clock %.0f MHz  chip %.0f TFLOP/s  (cyc per K16 MMA %.1f) err=%s\n",
         DUO ? "cta_group::2 M256" : "cta_group::1 M128", n, fill, ctas, best, flop_sm / clk, clk / ns * 1e3,
         flop_sm * ctas / (best * 1e9), clk / (4.0 * iters), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  unsigned long long* dout;
  cudaMalloc(&dout, 148 * 2 * 8);
  const int iters = 40000;
  for (int fill = 0; fill < 2; ++fill) {
    for (int n : {64, 128, 256}) run<false>(n, iters, fill, 148, dout);
    for (int n : {64, 128, 256}) run<true>(n, iters, fill, 148, dout);
  }
  for (int n : {64, 128, 192, 256}) { run_ring_uni<true>(n, iters, 5, dout); run_ring_uni<false>(n, iters, 5, dout); }
  for (int mode = 1; mode <= 3; ++mode)
    for (int st : {5}) { run_ring<true>(256, iters, mode, st, dout); run_ring<false>(256, iters, mode, st, dout); }
  run<false>(256, iters, 1, 1, dout);
  run<true>(256, iters, 1, 2, dout);
  return 0;
}
