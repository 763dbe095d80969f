// TMA-fed tcgen05 GEMM with TMEM accumulators for sm_100a.
//
//   D[M,N] = epilogue(A[M,K] * B[N,K]^T)      A, B: bf16 or fp16, K-major; fp32 accumulate
//
// Replaces every nn.Linear / patchify-conv / im2col'd conv / k2s2 transposed conv
// contraction of the GenConViT forward (see include/genconvit_b200.h, gcv_gemm).
//
// Structure (one persistent CTA per SM, or a cta_group::2 pair of them; 640 threads):
//   warp 0        TMA producer: cp.async.bulk.tensor 2D loads of the A (128x64) and B (block_n x 64) K-slices into a
//                 ring of 128B-swizzled smem stages (whole warp runs the loop, one elected lane issues)
//   warp 1        MMA issuer: tcgen05.mma (M=128 / 256 per pair, N=block_n, K=16) into a ring of 512/block_n TMEM
//                 accumulator stages; tcgen05.commit frees smem stages and publishes finished accumulators
//   warps 2..17   epilogue (4 per TMEM lane quarter): tcgen05.ld of the fp32 accumulator (thread = row, 2 x 16 columns
//                 per 32-column chunk), fused bias / activation / folded LayerNorm / layer-scale+residual / VAE
//                 reparameterisation, 16-bit result staged in smem and sent off as one TMA store per 32 x 32 piece
//   warps 18..19  row statistics of the folded LayerNorm (fc1): reduce the per-chunk partial sums of a tile's 128 rows
//                 one tile ahead of its epilogue (idle in the other modes)
// The accumulator ring lets tile i's epilogue overlap the following tiles' MMAs.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gcv {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                       // 64 x 16-bit = one 128-byte swizzle row
constexpr int kEpiWarps = 16;                 // 4 per TMEM lane quarter: TLP hides tcgen05.ld / MUFU / smem latency
constexpr int kStatWarp = 2 + kEpiWarps;       // MODE 3: two warps reduce the tile rows' LayerNorm partial sums one tile ahead
constexpr int kStatWarps = 2;                 // (640 threads leave exactly the 96 registers per thread the epilogues need)
constexpr int kThreads = 64 + 32 * kEpiWarps + 32 * kStatWarps;
constexpr int kStatStages = 4;                // ring of per-tile row statistics, indexed like the accumulator stages
constexpr int kMaxStages = 8;
constexpr int kMaxAccStages = 8;              // TMEM accumulator ring: 512 columns / block_n
constexpr int kTileSmem = 160 * 1024;            // 227 KB budget minus control block, epilogue staging, vectors, slack
constexpr int kVecMaxN = 3072;                   // bias / layer-scale vectors up to this N are staged in smem
constexpr int kVecSmem = 2 * kVecMaxN * 4;
constexpr int kCtrlSmem = 1024 + kStatStages * 128 * 8;   // barriers + TMEM slot (1 KB), row-statistics ring (4 KB)
constexpr int kStageRow = 64;                            // staged row: 32 x 16-bit, 16-byte pieces XOR-swizzled (TMA SWIZZLE_64B)
constexpr int kStageWarp = 32 * kStageRow;               // per epilogue warp: 32 rows x 32 columns = one TMA store box
constexpr int kStageSmem = kEpiWarps * kStageWarp;
constexpr int kDynSmem = kCtrlSmem + kStageSmem + kVecSmem + 1024 + kTileSmem;   // +1024 alignment slack
constexpr uint32_t kTmemCols = 512;

// GCV_DEBUG what-if switches (timing experiments, results are garbage) only exist in builds with -DGCV_GEMM_WHATIF
// (GCV_NVCC_FLAGS=-DGCV_GEMM_WHATIF python -m genconvit_b200.build): in the shipped kernel they were ~10 branches per
// 32-column chunk of every epilogue warp.
#ifdef GCV_GEMM_WHATIF
constexpr bool kWhatIf = true;
#else
constexpr bool kWhatIf = false;
#endif

struct Params {
  int64_t M;
  int N, K;
  int block_n, num_stages, acc_stages;
  int n_sub, mma_n;    // a tile's accumulator is n_sub MMAs of width mma_n side by side (block_n = n_sub * mma_n)
  int tiles_m, tiles_n;
  uint32_t idesc;
  int vec_ok;          // epilogue may use 16-byte row-chunk loads/stores
  int vec_smem;        // bias / gamma are staged in shared memory (N <= kVecMaxN)
  int tma_store;       // row-major 16-bit output leaves through TMA stores (tmap_d) instead of per-lane st.global
  int debug;           // profiling knob (GCV_DEBUG): 1 = epilogue skips all work, 2 = skips stores
  gcv_epilogue ep;
};

// K-major, 128B-swizzled operand tile: rows are 128 B, 8-row atoms are 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO [16,30), SBO [32,46), version=1 [46,48), SWIZZLE_128B=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                  // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row atoms
  d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                  // SWIZZLE_128B
  return d;
}

// Cold path, kept out of line so the hot epilogue loop stays small in the instruction cache: element-wise
// epilogue for 8 consecutive columns (ragged right edge, fp32 / unaligned outputs, VAE reparameterisation).
template <typename T>
__device__ __noinline__ void epilogue_slow8(const gcv_epilogue& ep, int64_t m, int n, int N, float v0, float v1,
                                            float v2, float v3, float v4, float v5, float v6, float v7, void* D) {
  const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
#pragma unroll 1
  for (int e = 0; e < 8; ++e)
    if (n + e < N) epilogue_one<T>(ep, m, n + e, N, v[e], D);
}

// MODE: 0 = staged epilogue, general (bias / act / layer-scale / residual); 1 = staged, bias + GELU only (packed fp16
// math); 2 = element-wise cold path (ragged / fp32 / reparameterisation outputs); 3 = mode 1 with the LayerNorm of
// the A rows folded in (gcv_epilogue.ln_stats: per-row rstd / mean from partial sums, column sums in vec_gamma);
// 4 = staged, bias + layer-scale + residual only (ConvNeXt fc2): the residual rows are fetched a chunk ahead (the
// first chunk before the accumulator wait), so their HBM latency is off the epilogue's critical path.  Separate instantiations keep each
// epilogue within the 96 registers a 640-thread CTA leaves per thread.
// DUO: the CTA pair of a 2-CTA cluster computes one 256 x block_n tile with cta_group::2 MMAs: each CTA stages its
// own 128 rows of A and HALF of the B tile (so a stage is 16 KB + block_n/2 x 128 B instead of 16 KB + block_n x 128 B:
// one third less shared-memory and L2 traffic per MAC, deeper pipeline), the leader CTA issues the MMAs, and each
// CTA drains its own 128 accumulator rows from its own TMEM.
template <typename T, int MODE, bool DUO>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_d, void* D, const Params p) {
  extern __shared__ uint8_t smem_dyn[];
  // 1024-byte aligned base: TMA / UMMA swizzle atoms (operand stages, the epilogue's store boxes)
  uint8_t* smem_raw = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  // control block: barriers + TMEM base pointer
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tmem_full = empty_bar + kMaxStages;
  uint64_t* tmem_empty = tmem_full + kMaxAccStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + kMaxAccStages);
  uint64_t* stat_full = reinterpret_cast<uint64_t*>(smem_raw + 512);      // MODE 3 with partial sums (ep.ln_chunks > 0)
  uint64_t* stat_empty = stat_full + kStatStages;
  float2* stat_ring = reinterpret_cast<float2*>(smem_raw + 1024);         // [kStatStages][BM] (rstd, -mean * rstd)
  uint8_t* stage_base = smem_raw + kCtrlSmem;            // epilogue staging, kStageWarp bytes per warp
  float* vec_bias = reinterpret_cast<float*>(smem_raw + kCtrlSmem + kStageSmem);   // [kVecMaxN] bias, then gamma
  float* vec_gamma = vec_bias + kVecMaxN;
  const uint32_t tiles_base = smem_u32(smem_raw) + kCtrlSmem + kStageSmem + kVecSmem;

  const int dbg = kWhatIf ? p.debug : 0;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);      // provably warp-uniform (see elect_one)
  const int lane = threadIdx.x & 31;
  const uint32_t b_rows = DUO ? (uint32_t)p.mma_n >> 1 : (uint32_t)p.mma_n;     // B rows this CTA stages per MMA
  const uint32_t a_bytes = BM * BK * 2, sub_bytes = b_rows * BK * 2, b_bytes = (uint32_t)p.n_sub * sub_bytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_tiles = p.tiles_m * p.tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.num_stages; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int s = 0; s < kStatStages; ++s) {
      mbar_init(smem_u32(stat_full + s), kStatWarps);
      mbar_init(smem_u32(stat_empty + s), kEpiWarps);
    }
    for (int s = 0; s < p.acc_stages; ++s) {
      mbar_init(smem_u32(tmem_full + s), 1);
      mbar_init(smem_u32(tmem_empty + s), DUO ? 2 * kEpiWarps : kEpiWarps);   // DUO: both CTAs' epilogues release the leader
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (DUO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.vec_smem) {
    // bias / layer-scale vectors are read by every epilogue thread for every tile: keep them on chip
    for (int i = threadIdx.x; i < p.N; i += kThreads) {
      // the GELU epilogues (MODE 1, 3) work on half the pre-activation: their staged bias is bias / 2
      vec_bias[i] = p.ep.bias ? ((MODE == 1 || MODE == 3) ? 0.5f : 1.0f) * __ldg(p.ep.bias + i) : 0.0f;
      vec_gamma[i] = MODE == 3 ? __ldg(p.ep.ln_colsum + i) : (p.ep.gamma ? __ldg(p.ep.gamma + i) : 1.0f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();     // barrier inits / TMEM allocation of both CTAs visible before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t crank = DUO ? cluster_ctarank() : 0;
  const bool leader = crank == 0;
  // tile walk: a CTA (or a pair, taking m-tiles 2i and 2i+1 of one n-tile) strides over the tile list
  const int walker = DUO ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int walkers = DUO ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int walk_tiles = DUO ? ((p.tiles_m + 1) >> 1) * p.tiles_n : num_tiles;

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = walker; tile < walk_tiles; tile += walkers) {
        const int mw = tile / p.tiles_n, n_blk = tile - mw * p.tiles_n;
        const int m_blk = DUO ? 2 * mw + (int)crank : mw;           // may be one past the end: TMA zero-fills, nothing is stored
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
          const uint32_t fb = smem_u32(full_bar + stage);
          const uint32_t sa = tiles_base + stage * stage_bytes;
          if (!elect_one()) {
            // nothing to issue
          } else if (dbg >= 3 && dbg <= 5) {
            // mainloop experiments (results are garbage): 3 = no A loads, 4 = no B loads, 5 = no loads at all
            const bool ldA = dbg == 4, ldB = dbg == 3;
            const uint32_t bytes = (ldA ? a_bytes : 0u) + (ldB ? b_bytes : 0u);
            if (!DUO || leader) mbar_expect_tx(fb, DUO ? 2 * bytes : bytes);
            if constexpr (DUO) {
              if (ldA) tma_load_2d_2sm(sa, &tmap_a, fb, kb * BK, m_blk * BM);
              if (ldB)
                for (int j = 0; j < p.n_sub; ++j)
                  tma_load_2d_2sm(sa + a_bytes + j * sub_bytes, &tmap_b, fb, kb * BK,
                                  n_blk * p.block_n + j * p.mma_n + (int)(crank * b_rows));
            } else {
              if (ldA) tma_load_2d(sa, &tmap_a, fb, kb * BK, m_blk * BM);
              if (ldB) tma_load_2d(sa + a_bytes, &tmap_b, fb, kb * BK, n_blk * p.block_n);
            }
          } else if constexpr (DUO) {
            // both CTAs' loads complete on the leader's barrier, which therefore expects two stages' worth of bytes
            if (leader) mbar_expect_tx(fb, 2 * stage_bytes);
            tma_load_2d_2sm(sa, &tmap_a, fb, kb * BK, m_blk * BM);
            for (int j = 0; j < p.n_sub; ++j)
              tma_load_2d_2sm(sa + a_bytes + j * sub_bytes, &tmap_b, fb, kb * BK,
                              n_blk * p.block_n + j * p.mma_n + (int)(crank * b_rows));
          } else {
            mbar_expect_tx(fb, stage_bytes);
            tma_load_2d(sa, &tmap_a, fb, kb * BK, m_blk * BM);
            tma_load_2d(sa + a_bytes, &tmap_b, fb, kb * BK, n_blk * p.block_n);
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop, one elected lane issues) =====================
    if (!DUO || leader) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = walker; tile < walk_tiles; tile += walkers) {
        mbar_wait(smem_u32(tmem_empty + as), aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(full_bar + stage), phase);
          tc_fence_after();
          const uint32_t sa = tiles_base + stage * stage_bytes;
          const int k_left = p.K - kb * BK;
          const int kmma = k_left >= BK ? BK / 16 : (k_left + 15) / 16;
          if (elect_one()) {
            // descriptors advance by 32 B (= 2 in the >>4 address field) per K=16 slice
            const uint64_t ad0 = umma_desc(sa), bd0 = umma_desc(sa + a_bytes);
            if (kmma == BK / 16 && p.n_sub == 1) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {
                if constexpr (DUO) tc_mma_2sm(d_tmem, ad0 + 2 * k, bd0 + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
                else tc_mma(d_tmem, ad0 + 2 * k, bd0 + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
              }
            } else {
              for (int k = 0; k < kmma; ++k) {
                if constexpr (DUO) {
                  for (int j = 0; j < p.n_sub; ++j)
                    tc_mma_2sm(d_tmem + (uint32_t)(j * p.mma_n), ad0 + 2 * k, bd0 + ((j * sub_bytes) >> 4) + 2 * k, p.idesc,
                               (kb | k) ? 1u : 0u);
                } else {
                  tc_mma(d_tmem, ad0 + 2 * k, bd0 + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
                }
              }
            }
            if constexpr (DUO) tc_commit_2sm(smem_u32(empty_bar + stage));   // frees this stage in both CTAs
            else tc_commit(smem_u32(empty_bar + stage));  // frees the smem stage once these MMAs retire
            if (kb == num_kb - 1) {
              if constexpr (DUO) tc_commit_2sm(smem_u32(tmem_full + as));      // accumulator complete, in both CTAs' TMEM
              else tc_commit(smem_u32(tmem_full + as));        // accumulator complete
            }
          }
          __syncwarp();
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
        if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= kStatWarp) {
    // ===================== row statistics (folded LayerNorm from per-chunk partial sums) =====================
    // One tile ahead of the epilogue: (rstd, -mean * rstd) of the tile's 128 rows (64 per warp, two per lane), summed in
    // chunk order (the values gcv_ln_finalize writes), into the ring slot of the tile's accumulator stage.  The partial
    // sums are 8 * chunks bytes per row, read once per tile from L2 here instead of by every epilogue thread -- and the
    // 24 reduction launches per step between the depthwise convolutions and the fc1 GEMMs are gone.  All of a lane's
    // loads are in flight together: one L2 round trip per tile (two for 24 chunks), well inside a tile period.
    if constexpr (MODE == 3) {
      const int chunks = p.ep.ln_chunks;
      if (chunks > 0) {
        const int sw = warp - kStatWarp;
        const float inv = 1.0f / (float)p.K;
        const float eps_ln = p.ep.ln_eps;
        auto finish = [&](float sum, float sq) {
          const float mean = sum * inv;
          const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv), 0.0f) + eps_ln);
          return make_float2(rstd, -mean * rstd);
        };
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = walker; tile < walk_tiles; tile += walkers) {
          const int mw = tile / p.tiles_n;
          const int r0 = sw * 64 + lane, r1 = r0 + 32;                     // this lane's two rows of the tile
          const int64_t m0 = (int64_t)(DUO ? 2 * mw + (int)crank : mw) * BM + r0, m1 = m0 + 32;
          mbar_wait(smem_u32(stat_empty + as), aphase ^ 1);
          float2 rs0 = make_float2(1.0f, 0.0f), rs1 = rs0;
          if ((chunks & 1) == 0 && chunks <= 12) {
            const int nv = chunks >> 1;
            const float4* s0 = reinterpret_cast<const float4*>(p.ep.ln_stats) + (m0 < p.M ? m0 : 0) * nv;
            const float4* s1 = reinterpret_cast<const float4*>(p.ep.ln_stats) + (m1 < p.M ? m1 : 0) * nv;
            float4 v0[6], v1[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) {
              v0[c] = c < nv ? __ldg(s0 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
              v1[c] = c < nv ? __ldg(s1 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float a = 0.0f, b = 0.0f, c2 = 0.0f, d = 0.0f;
#pragma unroll
            for (int c = 0; c < 6; ++c)
              if (c < nv) {
                a += v0[c].x; b += v0[c].y; a += v0[c].z; b += v0[c].w;
                c2 += v1[c].x; d += v1[c].y; c2 += v1[c].z; d += v1[c].w;
              }
            rs0 = finish(a, b);
            rs1 = finish(c2, d);
          } else {
#pragma unroll 1
            for (int i = 0; i < 2; ++i) {
              const int64_t m = i ? m1 : m0;
              float sum = 0.0f, sq = 0.0f;
              if (m < p.M) {
                if ((chunks & 1) == 0 && chunks <= 24) {
                  const float4* src = reinterpret_cast<const float4*>(p.ep.ln_stats) + m * (chunks >> 1);
                  float4 v[12];
#pragma unroll
                  for (int c = 0; c < 12; ++c) v[c] = c < (chunks >> 1) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int c = 0; c < 12; ++c)
                    if (c < (chunks >> 1)) { sum += v[c].x; sq += v[c].y; sum += v[c].z; sq += v[c].w; }
                } else {
                  const float2* src = reinterpret_cast<const float2*>(p.ep.ln_stats) + m * chunks;
                  for (int c = 0; c < chunks; ++c) { const float2 v = __ldg(src + c); sum += v.x; sq += v.y; }
                }
              }
              (i ? rs1 : rs0) = finish(sum, sq);
            }
          }
          stat_ring[as * BM + r0] = rs0;
          stat_ring[as * BM + r1] = rs1;
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(stat_full + as));
          if (++as == p.acc_stages) { as = 0; aphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    // Phase A (thread = accumulator row): tcgen05.ld 32 columns, bias / activation / layer-scale /
    // residual in registers, 16-bit result parked in this warp's smem staging tile.
    // Phase B (lane = 16-byte piece): the 32x32 staged tile leaves as 8 rows x 64 contiguous bytes per
    // store instruction -- full 32-byte sectors instead of 32 scattered 16-byte row fragments.
    const int ew = warp - 2;
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may read
    const int sub = ew >> 2;                             // the 4 warps of a quarter take chunks sub, sub+4, ...
    const int chunks = p.block_n / 32;
    constexpr bool vec_ok = MODE != 2;
    constexpr bool gelu_only = MODE == 1 || MODE == 3;
    uint8_t* my_stage = stage_base + ew * kStageWarp;
    const gcv_epilogue& ep = p.ep;
    // hand an accumulator stage back to the MMA warp (DUO: the leader CTA's, from either CTA)
    auto release_acc = [&](int a) {
      if (DUO && !leader) mbar_arrive_remote(smem_u32(tmem_empty + a), 0);
      else mbar_arrive(smem_u32(tmem_empty + a));
    };
    int as = -1;
    uint32_t aphase = 1;
    for (int tile = walker; tile < walk_tiles; tile += walkers) {
      const int mw = tile / p.tiles_n, n_blk = tile - mw * p.tiles_n;
      const int m_blk = DUO ? 2 * mw + (int)crank : mw;
      if (++as == p.acc_stages) as = 0;
      if (as == 0) aphase ^= 1;
      const int64_t m_warp = (int64_t)m_blk * BM + quarter * 32;
      const int64_t m = m_warp + lane;
      float2 lnrs = make_float2(1.0f, 0.0f);             // folded LayerNorm: (rstd, -mean * rstd) of this thread's row,
      if constexpr (MODE == 3) {
        if (ep.ln_chunks > 0) {                          // reduced by the statistics warp while the previous tile drained
          mbar_wait(smem_u32(stat_full + as), aphase);
          lnrs = stat_ring[as * BM + quarter * 32 + lane];
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(stat_empty + as));
        } else if (m < p.M && sub < chunks) {            // one pair per row (gcv_ln_finalize), fetched before blocking
          lnrs = __ldg(reinterpret_cast<const float2*>(ep.ln_stats) + m);
        }
        lnrs.x *= 0.5f; lnrs.y *= 0.5f;                  // GELU is evaluated from x / 2 (ln_bias_gelu_pack8)
      }
      uint4 res[4];                                      // MODE 4: residual of this thread's row, 32 columns of a chunk
      auto fetch_residual = [&](int c) {
        const int nb = n_blk * p.block_n + c * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          res[j] = make_uint4(0u, 0u, 0u, 0u);
          if (m < p.M && nb + j * 8 + 8 <= p.N)
            res[j] = *reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(ep.residual) + m * ep.ldr + nb + j * 8);
        }
      };
      if constexpr (MODE == 4) {
        if (sub < chunks) fetch_residual(sub);
      }
      mbar_wait(smem_u32(tmem_full + as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.block_n);
      if (sub >= chunks) {                               // narrow tile: this warp has no chunk, release immediately
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(as);
      }
      if (dbg == 1 || (dbg >= 3 && dbg <= 5)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0 && sub < chunks) release_acc(as);
        continue;
      }
      for (int c = sub; c < chunks; c += 4) {
        const int n0 = n_blk * p.block_n + c * 32;
        const bool full = n0 + 32 <= p.N;

        // two 16-column TMEM loads per 32-column chunk: half the live registers of one x32 load (the 640-thread CTA
        // leaves 96 registers per thread)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float v[16];
          // this half's 16 columns of the per-column vectors (bias; layer scale / LayerNorm column sums), fetched from
          // shared memory while the TMEM load is in flight: the asm statements around the load are memory barriers for
          // the compiler, so left to itself it issues these loads after the wait and every FFMA2 stalls on them
          float4 pb[4], pg[4];
          {
            uint32_t r[16];
            tc_ld16(t_row + c * 32 + h * 16, r);
            if constexpr (MODE == 1 || MODE == 3 || MODE == 4) {
              if (full) {
                const int nh = n0 + h * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  pb[q] = *reinterpret_cast<const float4*>(vec_bias + nh + 4 * q);
                  if constexpr (MODE != 1) pg[q] = *reinterpret_cast<const float4*>(vec_gamma + nh + 4 * q);
                }
              }
            }
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
          }
          if (h == 1 && c + 4 >= chunks) {
            tc_fence_before();                           // all TMEM reads of this tile by this warp are done:
            __syncwarp();                                // hand the accumulator stage back to the MMA warp early
            if (lane == 0) release_acc(as);
          }
          if (dbg == 6) {                            // experiment: TMEM drain only
            if (v[0] == 123.456f) *reinterpret_cast<float*>(my_stage) = v[5];
            continue;
          }
          if (!vec_ok) {
            if (m < p.M) {
#pragma unroll
              for (int j = 0; j < 2; ++j)
                epilogue_slow8<T>(ep, m, n0 + h * 16 + j * 8, p.N, v[j * 8], v[j * 8 + 1], v[j * 8 + 2], v[j * 8 + 3],
                                  v[j * 8 + 4], v[j * 8 + 5], v[j * 8 + 6], v[j * 8 + 7], D);
            }
            continue;
          }
          // ---- phase A ----
          if (h == 0 && p.tma_store) {                   // the previous bulk store of this warp must have drained the staging
            if (elect_one()) tma_store_wait_read();      // tile: waited for as late as possible (after this chunk's TMEM load)
            __syncwarp();
          }
          // interior chunks (all 32 columns inside N) take a copy of the loop without the per-8-column bounds checks
          auto phase_a = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            const int j = h * 2 + jj;
            const int n = n0 + j * 8;
            float* w = v + jj * 8;
            uint4 q;
            if constexpr (MODE == 4) {
              // fc2 path: (acc + bias) * gamma + residual; columns past N (tile padding, N % 8 == 0) are never stored
              if (FULL || n + 8 <= p.N) {
                const float4 b0 = FULL ? pb[2 * jj] : *reinterpret_cast<const float4*>(vec_bias + n);
                const float4 b1 = FULL ? pb[2 * jj + 1] : *reinterpret_cast<const float4*>(vec_bias + n + 4);
                const float4 g0 = FULL ? pg[2 * jj] : *reinterpret_cast<const float4*>(vec_gamma + n);
                const float4 g1 = FULL ? pg[2 * jj + 1] : *reinterpret_cast<const float4*>(vec_gamma + n + 4);
                const float2 r0 = unpack2<T>(res[j].x), r1 = unpack2<T>(res[j].y), r2 = unpack2<T>(res[j].z), r3 = unpack2<T>(res[j].w);
                // (w + b) * g + r as w * g + (b * g + r): two FFMA2 per column pair
                const float2 y0 = fma2(make_float2(w[0], w[1]), make_float2(g0.x, g0.y), fma2(make_float2(b0.x, b0.y), make_float2(g0.x, g0.y), r0));
                const float2 y1 = fma2(make_float2(w[2], w[3]), make_float2(g0.z, g0.w), fma2(make_float2(b0.z, b0.w), make_float2(g0.z, g0.w), r1));
                const float2 y2 = fma2(make_float2(w[4], w[5]), make_float2(g1.x, g1.y), fma2(make_float2(b1.x, b1.y), make_float2(g1.x, g1.y), r2));
                const float2 y3 = fma2(make_float2(w[6], w[7]), make_float2(g1.z, g1.w), fma2(make_float2(b1.z, b1.w), make_float2(g1.z, g1.w), r3));
                q.x = pack2<T>(y0.x, y0.y); q.y = pack2<T>(y1.x, y1.y); q.z = pack2<T>(y2.x, y2.y); q.w = pack2<T>(y3.x, y3.y);
              } else {
                q = make_uint4(0u, 0u, 0u, 0u);
              }
            } else if (gelu_only && (FULL || n + 8 <= p.N)) {
              // fc1 fast path: bias + GELU in packed fp16 arithmetic
              if constexpr (MODE == 3)
                q = ln_bias_gelu_pack8<T>(w, lnrs.x, lnrs.y,
                                          FULL ? pg[2 * jj] : *reinterpret_cast<const float4*>(vec_gamma + n),
                                          FULL ? pg[2 * jj + 1] : *reinterpret_cast<const float4*>(vec_gamma + n + 4),
                                          FULL ? pb[2 * jj] : *reinterpret_cast<const float4*>(vec_bias + n),
                                          FULL ? pb[2 * jj + 1] : *reinterpret_cast<const float4*>(vec_bias + n + 4));
              else
                q = bias_gelu_pack8<T>(w, FULL ? pb[2 * jj] : *reinterpret_cast<const float4*>(vec_bias + n),
                                       FULL ? pb[2 * jj + 1] : *reinterpret_cast<const float4*>(vec_bias + n + 4));
            } else {
              if (FULL || n + 8 <= p.N) {
                if (p.vec_smem) {
                  if (ep.bias) {
                    const float4 b0 = *reinterpret_cast<const float4*>(vec_bias + n);
                    const float4 b1 = *reinterpret_cast<const float4*>(vec_bias + n + 4);
                    w[0] += b0.x; w[1] += b0.y; w[2] += b0.z; w[3] += b0.w;
                    w[4] += b1.x; w[5] += b1.y; w[6] += b1.z; w[7] += b1.w;
                  }
                } else if (ep.bias) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.bias + n + 4));
                  w[0] += b0.x; w[1] += b0.y; w[2] += b0.z; w[3] += b0.w;
                  w[4] += b1.x; w[5] += b1.y; w[6] += b1.z; w[7] += b1.w;
                }
                if (ep.act != GCV_ACT_NONE) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] = apply_act_fast(w[e], ep.act);
                }
                if (ep.eps) {
                  // VAE reparameterisation z = eps * exp(mu / 2) + mu (genconvit_vae.py:43-49) in the staged epilogue:
                  // eps is in the reference's latent order c * eps_hw + hw, the 8 columns here are 8 channels of one hw.
                  // (The per-element path this replaces wrote 2-byte scattered stores: 0.135 of the layer's 0.29 ms.)
                  if (m < p.M) {
                    const int hw = n / ep.eps_c, c0 = n - hw * ep.eps_c;
                    const float* er = ep.eps + m * (int64_t)p.N + (int64_t)c0 * ep.eps_hw + hw;
#pragma unroll
                    for (int e = 0; e < 8; ++e) w[e] = __ldg(er + (int64_t)e * ep.eps_hw) * expf(0.5f * w[e]) + w[e];
                  }
                }
                if (ep.gamma) {
                  const float4 g0 = p.vec_smem ? *reinterpret_cast<const float4*>(vec_gamma + n)
                                               : __ldg(reinterpret_cast<const float4*>(ep.gamma + n));
                  const float4 g1 = p.vec_smem ? *reinterpret_cast<const float4*>(vec_gamma + n + 4)
                                               : __ldg(reinterpret_cast<const float4*>(ep.gamma + n + 4));
                  w[0] *= g0.x; w[1] *= g0.y; w[2] *= g0.z; w[3] *= g0.w;
                  w[4] *= g1.x; w[5] *= g1.y; w[6] *= g1.z; w[7] *= g1.w;
                }
                if (ep.residual && m < p.M) {
                  float rr[8];
                  load8<T>(reinterpret_cast<const T*>(ep.residual) + m * ep.ldr + n, rr);
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] += rr[e];
                }
              } else {
                // ragged right edge (N % 8 != 0): finish these columns element-wise, nothing staged
                if (m < p.M && n < p.N)
                  epilogue_slow8<T>(ep, m, n, p.N, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7], D);
              }
              q.x = pack2<T>(w[0], w[1]); q.y = pack2<T>(w[2], w[3]);
              q.z = pack2<T>(w[4], w[5]); q.w = pack2<T>(w[6], w[7]);
            }
            *reinterpret_cast<uint4*>(my_stage + lane * kStageRow + ((j ^ ((lane >> 1) & 3)) << 4)) = q;
          }
          };
          if (full) phase_a(std::true_type{});
          else phase_a(std::false_type{});
        }
        if (!vec_ok || dbg == 6) continue;
        if constexpr (MODE == 4) {
          if (c + 4 < chunks) fetch_residual(c + 4);     // in flight during phase B and the next chunk's TMEM loads
        }
        if (p.tma_store) {
          // ---- phase B, TMA form: the 32 x 32 staged tile is one bulk tensor store (rows / columns past M / N are clipped)
          fence_proxy_async_smem();
          __syncwarp();
          if (dbg != 2 && dbg != 7 && elect_one()) {
            tma_store_2d(&tmap_d, smem_u32(my_stage), n0, (int)m_warp);
            tma_store_commit();
          }
          __syncwarp();
          continue;
        }
        __syncwarp();
        if (dbg == 2) { __syncwarp(); continue; }
        // ---- phase B ----
        const int piece = lane & 3;
        const int n = n0 + piece * 8;
        if (n + 8 <= p.N) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2);
            const int64_t mm = m_warp + row;
            if (mm < p.M) {
              const uint4 q = *reinterpret_cast<const uint4*>(my_stage + row * kStageRow + ((piece ^ ((row >> 1) & 3)) << 4));
              int64_t off;
              if (ep.store == GCV_STORE_PIXEL_SHUFFLE2) {
                const int ij = n / ep.ps_co, co = n - ij * ep.ps_co;
                const int64_t hw = (int64_t)ep.ps_h * ep.ps_w;
                const int64_t b = mm / hw;
                const int rem = (int)(mm - b * hw);
                const int h = rem / ep.ps_w, wq = rem - h * ep.ps_w;
                off = ((b * (2 * ep.ps_h) + 2 * h + (ij >> 1)) * (int64_t)(2 * ep.ps_w) + 2 * wq + (ij & 1)) * ep.ps_co + co;
              } else {
                off = mm * ep.ldd + n;
              }
              if (dbg != 7 || q.x == 0x12345678u)      // 7: everything but the global store itself
                *reinterpret_cast<uint4*>(reinterpret_cast<T*>(D) + off) = q;
            }
          }
        }
        __syncwarp();                                    // staging tile is reused by the next chunk
      }
    }
    if (p.tma_store) {                                   // shared memory must outlive the last bulk stores' reads
      if (elect_one()) tma_store_wait_read();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (DUO) cluster_sync_all();   // the peer may still be signalling this CTA's barriers / reading its smem
  if (warp == 1) {
    if constexpr (DUO)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_map(CUtensorMap* map, int dtype, const void* base, int64_t rows, int64_t k, int64_t ld, int box_rows,
             int box_cols = BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not resolvable (no CUDA driver?)");
    return GCV_ERR_NO_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows=%lld k=%lld ld=%lld box_rows=%d)", (int)r,
              (long long)rows, (long long)k, (long long)ld, box_rows);
    return GCV_ERR_CUDA;
  }
  return GCV_OK;
}

int pick_block_n(int64_t M, int N, int sms) {
  // Largest tile that still gives every SM work; N tiles must be multiples of 32 (epilogue chunk) <= 256.
  // 256 and 128 give every epilogue warp the same number of 32-column chunks (4 warps per lane quarter)
  // (measured after the warp-uniform issue fix: for N = 384 two 192-wide tiles with two TMEM stages beat both 3 x 128
  // and the single-stage 384-wide pair tile by ~12 %)
  const int cands[] = {256, 192, 128, 96, 64, 32};
  const int64_t tiles_m = (M + BM - 1) / BM;
  int best = 32;
  for (int c : cands) {
    if (c > ((N + 31) / 32) * 32) continue;          // tile wider than the problem
    const int tn = (N + c - 1) / c;
    const int waste = tn * c - N;
    if (waste * 8 > N && c > 32) continue;           // >12.5% padded columns
    best = c;
    if (tiles_m * tn >= sms) break;                  // big enough to fill the machine
  }
  return best;
}

}  // namespace

bool tcgen05_eligible(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t K) {
  if (dtype != GCV_BF16 && dtype != GCV_F16) return false;
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
  if (lda % 8 || ldb % 8 || K % 8 || K < 8) return false;
  return true;
}

int gemm_tcgen05(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t M, int64_t N,
                 int64_t K, const gcv_epilogue* ep, int force_block_n, cudaStream_t stream) {
  GCV_REQUIRE(tcgen05_eligible(dtype, A, lda, B, ldb, K),
              "tcgen05 GEMM needs bf16/fp16, 16-byte aligned A/B and K, lda, ldb multiples of 8 (K=%lld lda=%lld ldb=%lld)",
              (long long)K, (long long)lda, (long long)ldb);
  GCV_REQUIRE(M > 0 && N > 0 && N < (1 << 30), "bad GEMM shape");
  const int sms = device_sms();
  Params p{};
  p.M = M; p.N = (int)N; p.K = (int)K;
  p.block_n = force_block_n > 0 ? force_block_n : pick_block_n(M, (int)N, sms);
  GCV_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && (p.block_n <= 256 || p.block_n == 384),
              "block_n must be a multiple of 32 in [32,256], or 384");
  p.tiles_m = (int)((M + BM - 1) / BM);
  p.tiles_n = (int)((N + p.block_n - 1) / p.block_n);
  // DUO (cta_group::2 pairs) for the large contractions: they are shared-memory / L2 bandwidth bound with one CTA per
  // 128 x block_n tile.  GCV_GEMM_DUO=0 disables it (A/B timing).
  static int duo_env = -1;
  if (duo_env < 0) { const char* e = getenv("GCV_GEMM_DUO"); duo_env = e ? atoi(e) : 1; }
  // Wide pair tiles (256 x 384 as two N=192 MMAs) for the large-K, N = 384k contractions (ConvNeXt fc2 / downsample):
  // A -- the big operand there -- is then fetched once per 384 instead of once per 128 output columns, which is what
  // bounds these GEMMs (L2 -> SM bandwidth).  One TMEM accumulator stage (384 of 512 columns).  GCV_GEMM_WIDE=0 disables.
  static int wide_env = -1;
  if (wide_env < 0) { const char* e = getenv("GCV_GEMM_WIDE"); wide_env = e ? atoi(e) : 0; }   // off: its single accumulator stage serialises mainloop and epilogue
  if (force_block_n <= 0 && wide_env && duo_env && N % 384 == 0 && K >= 512 && p.tiles_m >= 2 * sms) {
    p.block_n = 384;
    p.tiles_n = (int)(N / 384);
  }
  // Weight-streaming shape (exactly one CTA pair of rows, long K: the VAE's 25088 -> 12544 `mu` layer at bs 256): with
  // single-CTA tiles the two row tiles of a column block run at different times and the 629 MB weight matrix is
  // streamed from HBM twice.  As ONE pair tile per column block B is read once; the tile width is the one that fills
  // a single round of the SM pairs best.  (Measured: 0.377 -> 0.316 ms; storing B K-block-major so that every TMA box
  // is contiguous changed nothing -- the remaining bound is every pair re-reading the same A tiles from L2.)
  bool stream_b = false;
  if (force_block_n <= 0 && duo_env && p.tiles_m == 2 && K >= 4096 && N >= 1024) {
    const int pairs = sms / 2;
    double best_eff = 0.0;
    int best_bn = 0;
    for (int bn : {256, 192, 128}) {
      const int t = (int)((N + bn - 1) / bn);
      const int rounds = (t + pairs - 1) / pairs;
      const double eff = (double)N / ((double)rounds * pairs * bn);      // useful columns per column slot
      if (eff > best_eff) { best_eff = eff; best_bn = bn; }
    }
    if (best_bn) {
      stream_b = true;
      p.block_n = best_bn;
      p.tiles_n = (int)((N + best_bn - 1) / best_bn);
    }
  }
  const bool duo = duo_env && p.tiles_m >= 2 && K >= 256 && p.block_n >= 64 &&
                   ((int64_t)p.tiles_m * p.tiles_n >= 2 * sms || stream_b);
  GCV_REQUIRE(p.block_n <= 256 || duo, "block_n = 384 needs the paired (cta_group::2) mode");
  p.n_sub = p.block_n > 256 ? 2 : 1;
  p.mma_n = p.block_n / p.n_sub;
  const int stage_bytes = BM * BK * 2 + (duo ? p.block_n / 2 : p.block_n) * BK * 2;
  p.acc_stages = 512 / p.block_n;
  if (p.acc_stages > kMaxAccStages) p.acc_stages = kMaxAccStages;
  if (ep->ln_stats && ep->ln_chunks > 0 && p.acc_stages > kStatStages) p.acc_stages = kStatStages;   // one statistics slot per stage
  p.num_stages = kTileSmem / stage_bytes;
  if (p.num_stages > kMaxStages) p.num_stages = kMaxStages;
  {
    static int cap = -1;                                   // GCV_GEMM_STAGES: pipeline-depth experiments
    if (cap < 0) { const char* e = getenv("GCV_GEMM_STAGES"); cap = e ? atoi(e) : 0; }
    if (cap > 0 && p.num_stages > cap) p.num_stages = cap;
  }
  p.idesc = umma_idesc_f16(dtype == GCV_BF16, duo ? 2 * BM : BM, p.mma_n);
  p.ep = *ep;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GCV_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  {
    const size_t es = ep->out_f32 ? 4 : 2;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    // the reparameterisation runs in the staged path when no fp32 copy of mu is wanted and 8 columns never straddle an hw
    const bool eps_ok = ep->eps == nullptr || (ep->mu_out == nullptr && ep->eps_c % 8 == 0 && ep->act == GCV_ACT_NONE);
    bool ok = eps_ok && !ep->out_f32 && al16(D) && (!ep->bias || al16(ep->bias)) && (!ep->gamma || al16(ep->gamma));
    if (ep->store == GCV_STORE_ROWS) ok = ok && (ep->ldd * es) % 16 == 0;
    else ok = ok && ep->ps_co % 8 == 0;
    if (ep->residual) ok = ok && al16(ep->residual) && ep->ldr % 8 == 0;
    p.vec_ok = ok ? 1 : 0;
    p.vec_smem = (ok && N <= kVecMaxN && (ep->bias || ep->gamma)) ? 1 : 0;
  }

  CUtensorMap ma, mb;
  int rc = make_map(&ma, dtype, A, M, K, lda, BM);
  if (rc) return rc;
  rc = make_map(&mb, dtype, B, N, K, ldb, duo ? p.mma_n / 2 : p.mma_n);
  if (rc) return rc;
  // row-major 16-bit outputs leave through TMA stores: one 32 x 32 box per epilogue warp and chunk.  GCV_GEMM_TMA_STORE=0
  // keeps the per-lane st.global path (A/B timing).
  static int tst_env = -1;
  if (tst_env < 0) { const char* e = getenv("GCV_GEMM_TMA_STORE"); tst_env = e ? atoi(e) : 1; }
  p.tma_store = (tst_env && p.vec_ok && ep->store == GCV_STORE_ROWS && N % 8 == 0) ? 1 : 0;
  CUtensorMap md = ma;
  if (p.tma_store) {
    rc = make_map(&md, dtype, D, M, N, ep->ldd, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }

  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  int grid = (int)(tiles < sms ? tiles : sms);
  if (duo) grid &= ~1;
  int mode = !p.vec_ok ? 2
             : (p.vec_smem && ep->act == GCV_ACT_GELU && ep->bias && !ep->gamma && !ep->residual) ? 1 : 0;
  if (mode == 0 && p.vec_smem && ep->act == GCV_ACT_NONE && ep->bias && ep->gamma && ep->residual &&
      ep->store == GCV_STORE_ROWS && N % 8 == 0)
    mode = 4;
  if (ep->ln_stats) {
    GCV_REQUIRE(mode == 1 && ep->ln_colsum && ep->ln_chunks >= 0 && N % 8 == 0,
                "tcgen05 GEMM: the folded LayerNorm needs bias + GELU, a 16-bit row-major output and N %% 8 == 0, N <= %d",
                kVecMaxN);
    mode = 3;
  }
  cudaError_t le = cudaSuccess;
  static unsigned long long attr_set[2][5][2] = {};         // per kernel instantiation: devices whose smem limit is raised
  unsigned long long& attr_devs = attr_set[dtype == GCV_BF16 ? 0 : 1][mode][duo ? 1 : 0];
  auto launch = [&](auto kernel) {
    if (first_on_device(attr_devs)) {
      le = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem);
      if (le != cudaSuccess) return;
    }
    if (!duo) {
      kernel<<<grid, kThreads, kDynSmem, stream>>>(ma, mb, md, D, p);
      le = cudaGetLastError();
    } else {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(kThreads);
      cfg.dynamicSmemBytes = kDynSmem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      le = cudaLaunchKernelEx(&cfg, kernel, ma, mb, md, D, p);
    }
  };
  if (dtype == GCV_BF16) {
    if (duo) {
      if (mode == 0) launch(gemm_tcgen05_kernel<__nv_bfloat16, 0, true>);
      else if (mode == 1) launch(gemm_tcgen05_kernel<__nv_bfloat16, 1, true>);
      else if (mode == 3) launch(gemm_tcgen05_kernel<__nv_bfloat16, 3, true>);
      else if (mode == 4) launch(gemm_tcgen05_kernel<__nv_bfloat16, 4, true>);
      else launch(gemm_tcgen05_kernel<__nv_bfloat16, 2, true>);
    } else {
      if (mode == 0) launch(gemm_tcgen05_kernel<__nv_bfloat16, 0, false>);
      else if (mode == 1) launch(gemm_tcgen05_kernel<__nv_bfloat16, 1, false>);
      else if (mode == 3) launch(gemm_tcgen05_kernel<__nv_bfloat16, 3, false>);
      else if (mode == 4) launch(gemm_tcgen05_kernel<__nv_bfloat16, 4, false>);
      else launch(gemm_tcgen05_kernel<__nv_bfloat16, 2, false>);
    }
  } else {
    if (duo) {
      if (mode == 0) launch(gemm_tcgen05_kernel<__half, 0, true>);
      else if (mode == 1) launch(gemm_tcgen05_kernel<__half, 1, true>);
      else if (mode == 3) launch(gemm_tcgen05_kernel<__half, 3, true>);
      else if (mode == 4) launch(gemm_tcgen05_kernel<__half, 4, true>);
      else launch(gemm_tcgen05_kernel<__half, 2, true>);
    } else {
      if (mode == 0) launch(gemm_tcgen05_kernel<__half, 0, false>);
      else if (mode == 1) launch(gemm_tcgen05_kernel<__half, 1, false>);
      else if (mode == 3) launch(gemm_tcgen05_kernel<__half, 3, false>);
      else if (mode == 4) launch(gemm_tcgen05_kernel<__half, 4, false>);
      else launch(gemm_tcgen05_kernel<__half, 2, false>);
    }
  }
  if (le != cudaSuccess) {
    set_error("gemm_tcgen05 launch: %s", cudaGetErrorString(le));
    return GCV_ERR_CUDA;
  }
  return check_launch("gemm_tcgen05");
}

}  // namespace gcv
