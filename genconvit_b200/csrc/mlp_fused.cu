// Fused ConvNeXt MLP for the wide-token stages (C = 96, 192):
//
//     x[m, :] += gamma * ( GELU( y[m, :] * W1^T + b1 ) * W2^T + b2 )          y = LN(dwconv(x))
//
// i.e. timm ConvNeXtBlock's  mlp.fc1 -> GELU -> mlp.fc2 -> gamma -> + shortcut  in ONE kernel.  Unfused, the
// [M, 4C] hidden activation makes a round trip through HBM (1.2 GB per block at 224x224, bs 256 for C = 96)
// and both GEMMs are store/load bound; here the hidden never leaves the SM.
//
// Per 128-row tile (persistent CTA, 576 threads):
//   warp 0    TMA producer: the tile's y block (C/32 K-blocks of 128 rows x 64 B, 64B-swizzled) and, through a
//             ring of slots, the K-blocks of W1 (128 hidden rows) and W2 (C output rows) for each 128-wide hidden
//             chunk -- the weights stream from L2, the ring keeps ~100 KB of loads in flight;
//   warp 1    MMA issuer (one lane): S[j%2] = y * W1_j^T (tcgen05.mma M128 x N128, K = C) into one of two TMEM
//             chunk accumulators, and, one chunk behind, O += H_j * W2_j^T (M128 x N = C, K = 128) where H_j is the
//             bf16 GELU output the epilogue warps wrote to shared memory in the UMMA K-major layout;
//   warps 2-17  epilogue: per chunk, each warp owns a 32-row x 32-column block: tcgen05.ld from S, + b1, GELU,
//             16-bit pack, swizzled st.shared into H (then fence.proxy.async + mbarrier arrive); per tile, the
//             O accumulator goes through the same bias / layer-scale / residual / staged coalesced store path as
//             the stand-alone GEMM.
// TMEM: S0 [0,128), S1 [128,256), O [256, 256+C).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gcv {

namespace {

constexpr int FM = 128;              // token rows per tile
constexpr int FCH = 128;             // hidden chunk width
constexpr int FKB = 32;              // K block: 32 x 16-bit = one 64-byte swizzle row
constexpr int FBLK = FM * 64;        // one 128-row K-block in smem (8 KB)
constexpr int kFEpiWarps = 16;
constexpr int kFThreads = 64 + 32 * kFEpiWarps;
constexpr int kFCtrl = 1024;
constexpr uint32_t kFTmemCols = 512;

template <int C>
struct Cfg {
  static constexpr int HC = 4 * C;
  static constexpr int NCH = HC / FCH;                       // hidden chunks per tile (3 / 6)
  static constexpr int XKB = C / FKB;                        // K-blocks of y / W1 (3 / 6)
  static constexpr int HKB = FCH / FKB;                      // K-blocks of H / W2 (4)
  static constexpr int X_BYTES = XKB * FBLK;
  static constexpr int H_BYTES = HKB * FBLK;
  // ring slot = 24 KB of weights: 3 K-blocks of W1 (128 hidden rows x 64 B each) or KB2 K-blocks of W2 (C rows x 64 B
  // each); one mbarrier wait per slot keeps the single MMA-issuing thread off the critical path
  static constexpr int SLOT = 3 * FBLK;
  static constexpr int KB1 = 3, P1 = XKB / KB1;              // W1 chunk = P1 slots
  static constexpr int KB2 = SLOT / (C * 64), P2 = HKB / KB2; // W2 chunk = P2 slots
  static constexpr int RING = C <= 96 ? 5 : 4;
  static_assert(XKB % KB1 == 0 && KB2 * C * 64 == SLOT && HKB % KB2 == 0, "slot geometry");
  static constexpr int VEC_BYTES = ((HC + 2 * C) * 4 + 1023) / 1024 * 1024;
  static constexpr int SMEM = kFCtrl + VEC_BYTES + X_BYTES + 2 * H_BYTES + RING * SLOT + 1024;
  static_assert(C % 32 == 0 && C <= 192 && HC % FCH == 0, "unsupported width");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct FParams {
  int64_t M;
  int tiles;
  uint32_t idesc1, idesc2;
  const float* b1;
  const float* b2;
  const float* gamma;
  void* x;                 // [M, C] residual in / result out (in place)
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, int C>
__global__ void __launch_bounds__(kFThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const FParams p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // control block
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase);
  uint64_t* x_full = bars + 0;
  uint64_t* x_empty = bars + 1;
  uint64_t* o_full = bars + 2;
  uint64_t* o_empty = bars + 3;
  uint64_t* s_full = bars + 4;        // [2]
  uint64_t* s_empty = bars + 6;       // [2]
  uint64_t* h_full = bars + 8;        // [2]
  uint64_t* h_empty = bars + 10;      // [2]
  uint64_t* ring_full = bars + 12;    // [RING]
  uint64_t* ring_empty = bars + 12 + K::RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12 + 2 * K::RING);
  float* vec_b1 = reinterpret_cast<float*>(gbase + kFCtrl);           // [HC]
  float* vec_b2 = vec_b1 + K::HC;                                      // [C]
  float* vec_g = vec_b2 + C;                                           // [C]
  const uint32_t x_smem = base + kFCtrl + K::VEC_BYTES;
  const uint32_t h_smem = x_smem + K::X_BYTES;                         // H0, H1
  const uint32_t ring_smem = h_smem + 2 * K::H_BYTES;
  uint8_t* h_gen = gbase + kFCtrl + K::VEC_BYTES + K::X_BYTES;        // generic pointer to H0

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(x_full), 1);
    mbar_init(smem_u32(x_empty), 1);
    mbar_init(smem_u32(o_full), 1);
    mbar_init(smem_u32(o_empty), kFEpiWarps);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(s_full + i), 1);
      mbar_init(smem_u32(s_empty + i), kFEpiWarps);
      mbar_init(smem_u32(h_full + i), kFEpiWarps);
      mbar_init(smem_u32(h_empty + i), 1);
    }
    for (int i = 0; i < K::RING; ++i) {
      mbar_init(smem_u32(ring_full + i), 1);
      mbar_init(smem_u32(ring_empty + i), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kFTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < K::HC; i += kFThreads) vec_b1[i] = __ldg(p.b1 + i);
  for (int i = threadIdx.x; i < C; i += kFThreads) {
    vec_b2[i] = __ldg(p.b2 + i);
    vec_g[i] = __ldg(p.gamma + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 2 * FCH;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t rphase = 0, x_use = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++x_use) {
        mbar_wait(smem_u32(x_empty), (x_use & 1) ^ 1);
        mbar_expect_tx(smem_u32(x_full), K::X_BYTES);
#pragma unroll 1
        for (int kb = 0; kb < K::XKB; ++kb)
          tma_load_2d(x_smem + kb * FBLK, &tm_y, smem_u32(x_full), kb * FKB, tile * FM);
#pragma unroll 1
        for (int j = 0; j <= K::NCH; ++j) {
          if (j < K::NCH) {
#pragma unroll 1
            for (int part = 0; part < K::P1; ++part) {
              mbar_wait(smem_u32(ring_empty + slot), rphase ^ 1);
              mbar_expect_tx(smem_u32(ring_full + slot), K::SLOT);
#pragma unroll
              for (int kb = 0; kb < K::KB1; ++kb)
                tma_load_2d(ring_smem + slot * K::SLOT + kb * FBLK, &tm_w1, smem_u32(ring_full + slot),
                            (part * K::KB1 + kb) * FKB, j * FCH);
              if (++slot == K::RING) { slot = 0; rphase ^= 1; }
            }
          }
          if (j >= 1) {
#pragma unroll 1
            for (int part = 0; part < K::P2; ++part) {
              mbar_wait(smem_u32(ring_empty + slot), rphase ^ 1);
              mbar_expect_tx(smem_u32(ring_full + slot), K::SLOT);
#pragma unroll
              for (int kb = 0; kb < K::KB2; ++kb)
                tma_load_2d(ring_smem + slot * K::SLOT + kb * (C * 64), &tm_w2, smem_u32(ring_full + slot),
                            (j - 1) * FCH + (part * K::KB2 + kb) * FKB, 0);
              if (++slot == K::RING) { slot = 0; rphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int slot = 0;
      // x_use = tiles done by this CTA; buffer b of S / H is used (NCH + 1 - b) / 2 times per tile, so its
      // n-th use (n = x_use * uses_per_tile + j / 2) has mbarrier parity n & 1
      uint32_t rphase = 0, x_use = 0, o_use = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++x_use) {
        mbar_wait(smem_u32(x_full), x_use & 1);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j <= K::NCH; ++j) {
          if (j < K::NCH) {
            const int b = j & 1;
            const uint32_t n_use = x_use * ((K::NCH + 1 - b) >> 1) + (j >> 1);
            mbar_wait(smem_u32(s_empty + b), (n_use & 1) ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + b * FCH;
#pragma unroll 1
            for (int part = 0; part < K::P1; ++part) {
              mbar_wait(smem_u32(ring_full + slot), rphase);
              tc_fence_after();
#pragma unroll
              for (int kb = 0; kb < K::KB1; ++kb) {
                const uint32_t a = x_smem + (part * K::KB1 + kb) * FBLK, w = ring_smem + slot * K::SLOT + kb * FBLK;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  tc_mma(d, umma_desc_kmajor<64>(a + k * 32), umma_desc_kmajor<64>(w + k * 32), p.idesc1,
                         (part | kb | k) ? 1u : 0u);
              }
              tc_commit(smem_u32(ring_empty + slot));
              if (++slot == K::RING) { slot = 0; rphase ^= 1; }
            }
            tc_commit(smem_u32(s_full + b));
            if (j == K::NCH - 1) tc_commit(smem_u32(x_empty));      // y block may be overwritten by the next tile
          }
          if (j >= 1) {
            const int jj = j - 1, hb = jj & 1;
            const uint32_t n_use = x_use * ((K::NCH + 1 - hb) >> 1) + (jj >> 1);
            mbar_wait(smem_u32(h_full + hb), n_use & 1);
            tc_fence_after();
            if (jj == 0) {
              mbar_wait(smem_u32(o_empty), (o_use & 1) ^ 1);
              tc_fence_after();
            }
#pragma unroll 1
            for (int part = 0; part < K::P2; ++part) {
              mbar_wait(smem_u32(ring_full + slot), rphase);
              tc_fence_after();
#pragma unroll
              for (int kb = 0; kb < K::KB2; ++kb) {
                const uint32_t a = h_smem + hb * K::H_BYTES + (part * K::KB2 + kb) * FBLK;
                const uint32_t w = ring_smem + slot * K::SLOT + kb * (C * 64);
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  tc_mma(tmem_o, umma_desc_kmajor<64>(a + k * 32), umma_desc_kmajor<64>(w + k * 32), p.idesc2,
                         (jj | part | kb | k) ? 1u : 0u);
              }
              tc_commit(smem_u32(ring_empty + slot));
              if (++slot == K::RING) { slot = 0; rphase ^= 1; }
            }
            tc_commit(smem_u32(h_empty + hb));
            if (jj == K::NCH - 1) {
              tc_commit(smem_u32(o_full));
              ++o_use;
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                  // TMEM lane quarter = rows 32q .. 32q+31 of the tile
    const int s = (warp - 2) >> 2;           // K-block of the hidden chunk / column-chunk phase of O
    const int row = q * 32 + lane;
    const int sw = (row >> 1) & 3;           // 64B-swizzle XOR for this row
    uint32_t o_use = 0, t_use = 0;          // t_use = tiles done by this CTA (same use-count arithmetic as the MMA warp)
    uint8_t* my_stage = h_gen + s * FBLK + (q * 32) * 64;    // this warp's 32 x 64 B slice of H0 doubles as O staging
    T* xg = reinterpret_cast<T*>(p.x);
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++t_use) {
      const int64_t m_warp = (int64_t)tile * FM + q * 32;
      const int64_t m = m_warp + lane;
#pragma unroll 1
      for (int j = 0; j < K::NCH; ++j) {
        const int b = j & 1;
        const uint32_t n_use = t_use * ((K::NCH + 1 - b) >> 1) + (j >> 1);
        mbar_wait(smem_u32(s_full + b), n_use & 1);
        tc_fence_after();
        float v[32];
        {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + b * FCH + s * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(s_empty + b));
        const float* bj = vec_b1 + j * FCH + s * 32;
        uint4 pk[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float4 b0 = *reinterpret_cast<const float4*>(bj + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bj + g * 8 + 4);
          float* w = v + g * 8;
          w[0] = gelu_fast(w[0] + b0.x); w[1] = gelu_fast(w[1] + b0.y);
          w[2] = gelu_fast(w[2] + b0.z); w[3] = gelu_fast(w[3] + b0.w);
          w[4] = gelu_fast(w[4] + b1.x); w[5] = gelu_fast(w[5] + b1.y);
          w[6] = gelu_fast(w[6] + b1.z); w[7] = gelu_fast(w[7] + b1.w);
          pk[g].x = pack2<T>(w[0], w[1]); pk[g].y = pack2<T>(w[2], w[3]);
          pk[g].z = pack2<T>(w[4], w[5]); pk[g].w = pack2<T>(w[6], w[7]);
        }
        mbar_wait(smem_u32(h_empty + b), (n_use & 1) ^ 1);          // fc2 of the chunk that last used this buffer retired
        uint8_t* hrow = h_gen + b * K::H_BYTES + s * FBLK + row * 64;
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(hrow + ((g ^ sw) << 4)) = pk[g];
        fence_async_smem();                                          // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(h_full + b));
      }
      // ---- output accumulator: + b2, * gamma, + residual, staged coalesced store (in place on x) ----
      // the residual rows come from HBM: fetch them before blocking on the accumulator
      uint4 res[2][4];
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int c = s + 4 * ci;
          res[ci][g] = make_uint4(0, 0, 0, 0);
          if (c < C / 32 && m < p.M) res[ci][g] = *reinterpret_cast<const uint4*>(xg + m * C + c * 32 + g * 8);
        }
      mbar_wait(smem_u32(o_full), o_use & 1);
      tc_fence_after();
      ++o_use;
      if (s >= C / 32) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(o_empty));
      }
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = s + 4 * ci;
        if (c >= C / 32) break;
        float v[32];
        {
          uint32_t r[32];
          tc_ld32(tmem_o + ((uint32_t)(q * 32) << 16) + c * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        }
        if (c + 4 >= C / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(o_empty));
        }
        const int n0 = c * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n0 + g * 8;
          float* w = v + g * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(vec_b2 + n), b1 = *reinterpret_cast<const float4*>(vec_b2 + n + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(vec_g + n), g1 = *reinterpret_cast<const float4*>(vec_g + n + 4);
          float rr[8];
          {
            const uint4 rq = res[ci][g];
            float2 f;
            f = unpack2<T>(rq.x); rr[0] = f.x; rr[1] = f.y;
            f = unpack2<T>(rq.y); rr[2] = f.x; rr[3] = f.y;
            f = unpack2<T>(rq.z); rr[4] = f.x; rr[5] = f.y;
            f = unpack2<T>(rq.w); rr[6] = f.x; rr[7] = f.y;
          }
          w[0] = fmaf(w[0] + b0.x, g0.x, rr[0]); w[1] = fmaf(w[1] + b0.y, g0.y, rr[1]);
          w[2] = fmaf(w[2] + b0.z, g0.z, rr[2]); w[3] = fmaf(w[3] + b0.w, g0.w, rr[3]);
          w[4] = fmaf(w[4] + b1.x, g1.x, rr[4]); w[5] = fmaf(w[5] + b1.y, g1.y, rr[5]);
          w[6] = fmaf(w[6] + b1.z, g1.z, rr[6]); w[7] = fmaf(w[7] + b1.w, g1.w, rr[7]);
          uint4 pk;
          pk.x = pack2<T>(w[0], w[1]); pk.y = pack2<T>(w[2], w[3]);
          pk.z = pack2<T>(w[4], w[5]); pk.w = pack2<T>(w[6], w[7]);
          *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = pk;
        }
        __syncwarp();
        const int piece = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          const int64_t mm = m_warp + r;
          if (mm < p.M) {
            const uint4 pk = *reinterpret_cast<const uint4*>(my_stage + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4));
            *reinterpret_cast<uint4*>(xg + mm * C + n0 + piece * 8) = pk;
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn fused_get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int fused_map(CUtensorMap* map, int dtype, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
  EncodeTiledFn enc = fused_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not resolvable (no CUDA driver?)");
    return GCV_ERR_NO_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)FKB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fused MLP) failed: CUresult %d", (int)r);
    return GCV_ERR_CUDA;
  }
  return GCV_OK;
}

template <typename T, int C>
int launch_fused(int dtype, const void* y, const void* w1, const void* w2, const FParams& p, cudaStream_t stream) {
  using K = Cfg<C>;
  CUtensorMap my, m1, m2;
  int rc = fused_map(&my, dtype, y, p.M, C, FM);
  if (rc) return rc;
  if ((rc = fused_map(&m1, dtype, w1, K::HC, C, FCH))) return rc;
  if ((rc = fused_map(&m2, dtype, w2, C, K::HC, C))) return rc;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<T, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(fused MLP smem=%d): %s", K::SMEM, cudaGetErrorString(e));
      return GCV_ERR_CUDA;
    }
    attr = true;
  }
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int grid = p.tiles < sms ? p.tiles : sms;
  mlp_fused_kernel<T, C><<<grid, kFThreads, K::SMEM, stream>>>(my, m1, m2, p);
  return check_launch("mlp_fused");
}

}  // namespace

bool mlp_fused_supported(int dtype, int C) { return (dtype == GCV_BF16 || dtype == GCV_F16) && (C == 96 || C == 192); }

int mlp_fused(int dtype, const void* y, const void* w1, const float* b1, const void* w2, const float* b2,
              const float* gamma, void* x, int64_t M, int C, cudaStream_t stream) {
  GCV_REQUIRE(mlp_fused_supported(dtype, C), "mlp_fused: bf16/fp16 and C in {96,192} only (dtype=%d C=%d)", dtype, C);
  GCV_REQUIRE(M > 0 && y && w1 && w2 && b1 && b2 && gamma && x, "mlp_fused: bad arguments");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GCV_REQUIRE(al16(y) && al16(w1) && al16(w2) && al16(x), "mlp_fused: pointers must be 16-byte aligned");
  FParams p{};
  p.M = M;
  p.tiles = (int)((M + FM - 1) / FM);
  p.idesc1 = umma_idesc_f16(dtype == GCV_BF16, FM, FCH);
  p.idesc2 = umma_idesc_f16(dtype == GCV_BF16, FM, C);
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.x = x;
  if (dtype == GCV_BF16) {
    return C == 96 ? launch_fused<__nv_bfloat16, 96>(dtype, y, w1, w2, p, stream)
                   : launch_fused<__nv_bfloat16, 192>(dtype, y, w1, w2, p, stream);
  }
  return C == 96 ? launch_fused<__half, 96>(dtype, y, w1, w2, p, stream)
                 : launch_fused<__half, 192>(dtype, y, w1, w2, p, stream);
}

}  // namespace gcv
