// Does TMA multicast lift the L2 -> shared-memory delivery rate?  (the long-K GEMMs stream their
// operands at 10.4-11.6 TB/s whatever the tile shape; sharing an operand slice between CTAs only helps if that ceiling is
// on the L2 side, not in the per-SM delivery path.)
//
// Every CTA receives rounds x STAGES boxes of 128 rows x 128 bytes (16 KB, 128B-swizzled, the GEMM's A slice) from an
// L2-resident matrix and does nothing with them.  mode 0: each CTA fetches its own boxes.  mode 1: 2-CTA clusters, each
// CTA fetches HALF of every box (64 rows) and multicasts it to both CTAs: the same bytes land in every SM's shared
// memory, half as many leave L2.  mode 2: the same with 4-CTA clusters (a quarter box each).  A round is STAGES loads
// in flight, then a cluster barrier (so that nobody's barrier is signalled for the next phase before it is re-armed).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma_mcast tools/ubench_tma_mcast.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../genconvit_b200/csrc/tc_ptx.cuh"
using namespace gcv;

constexpr int MAX_STAGES = 13, BOX_ROWS = 128, BOX_BYTES = BOX_ROWS * 128;

__global__ void __launch_bounds__(128, 1) kern(const __grid_constant__ CUtensorMap tm, int rounds, int csize, int rows_total, int STAGES) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bars[MAX_STAGES];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  const uint32_t rank = csize > 1 ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(bars + s), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  const int part = BOX_ROWS / csize;                    // rows this CTA fetches of every box
  const uint16_t mask = (uint16_t)((1u << csize) - 1u);
  const int cluster_id = (int)blockIdx.x / csize;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) {
        const uint32_t bar = smem_u32(bars + s);
        mbar_expect_tx(bar, BOX_BYTES);
        // a different box per (cluster, round, stage), wrapping inside the L2-resident matrix
        const int row0 = (int)(((long long)(cluster_id * 977 + r * STAGES + s) * BOX_ROWS) % (rows_total - BOX_ROWS));
        const uint32_t dst = base + s * BOX_BYTES + rank * part * 128;
        if (csize > 1) tma_load_2d_mcast(dst, &tm, bar, 0, row0 + (int)rank * part, mask);
        else tma_load_2d(dst, &tm, bar, 0, row0);
      }
    }
    for (int s = 0; s < STAGES; ++s) mbar_wait(smem_u32(bars + s), (uint32_t)r & 1u);
    if (csize > 1) cluster_sync_all();
    else __syncthreads();
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 256 * 1024;                          // 256 K rows x 128 B = 32 MB: L2-resident
  void* buf;
  cudaMalloc(&buf, (size_t)rows * 128);
  cudaMemset(buf, 1, (size_t)rows * 128);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int rounds = 2000;
  // modes 0-2: all SMs, cluster size 1 / 2 / 4; modes 3-4: no multicast on a half / a quarter of the SMs (is the ceiling per SM
  // or in the L2?); modes 5-7: no multicast, all SMs, 4 / 12 / 13 boxes in flight instead of 8 (bandwidth or latency?)
  for (int mode = 0; mode < 8; ++mode) {
    const int csize = mode < 3 ? 1 << mode : 1;
    const int STAGES = mode == 5 ? 4 : (mode == 6 ? 12 : (mode == 7 ? 13 : 8));
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(BOX_ROWS / csize)};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, estr,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
    const int grid = mode == 3 ? sms / 2 : (mode == 4 ? sms / 4 : sms / csize * csize);
    const size_t smem = (size_t)MAX_STAGES * BOX_BYTES + 1024;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      cudaError_t err = cudaLaunchKernelEx(&cfg, kern, tm, rounds, csize, rows, STAGES);
      cudaEventRecord(e1);
      cudaError_t e2 = cudaDeviceSynchronize();
      if (err != cudaSuccess || e2 != cudaSuccess) {
        printf("mode %d: %s / %s\n", mode, cudaGetErrorString(err), cudaGetErrorString(e2));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double delivered = (double)grid * rounds * STAGES * BOX_BYTES;
      if (rep == 2)
        printf("cluster %d (%s): %d CTAs, %d x 16 KB in flight, %.2f us per round, delivered to shared memory %.2f TB/s "
               "(%.1f GB/s per SM), read from L2 %.2f TB/s\n", csize, csize == 1 ? "no multicast" : "multicast", grid, STAGES,
               ms * 1e3 / rounds, delivered / ms / 1e9, delivered / ms / 1e6 / grid, delivered / csize / ms / 1e9);
    }
  }
  return 0;
}
