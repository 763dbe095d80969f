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
// TMEM: S0 [0,128), S1 [128,256), H [256,320), O0 [320, 320+C) (, O1 [320+C, 320+2C) for C = 96).
//
// Second generation (what the first version's profile showed: l1tex 70 % busy, tensor pipe 25 % -- per 128-row tile the
// kernel moved ~730 KB through shared memory: weights re-streamed by TMA for every tile, both operands of every MMA,
// the GELU output written by st.shared and read back by the MMA):
//   * the GELU output H never touches shared memory: the epilogue warps write it to TENSOR memory (tcgen05.st, two
//     16-bit values per 32-bit column, lane = row) and fc2 takes it as its A operand from there (tcgen05.mma with
//     A in TMEM) -- no st.shared of H, no A-operand reads for fc2;
//   * C = 96: W1 and W2 (147 KB) are loaded ONCE per CTA and stay resident in shared memory; only the y block of the
//     next tile (double-buffered) is streamed.  C = 192 keeps streaming its 590 KB of weights through the slot ring,
//     which the freed H buffers make deeper.
#include <stdlib.h>

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
  // C = 96: both weight matrices resident in shared memory for the whole launch
  static constexpr bool RESIDENT = C <= 96;
  static constexpr int W1_CHUNK = XKB * FBLK;                // one 128-row hidden chunk of W1: XKB K-blocks of 128 x 64 B
  static constexpr int W2_CHUNK = HKB * C * 64;              // the matching K = 128 slab of W2: 4 K-blocks of C x 64 B
  static constexpr int W_BYTES = RESIDENT ? NCH * (W1_CHUNK + W2_CHUNK) : 0;
  // streamed weights (C = 192): ring slot = 24 KB: 3 K-blocks of W1 (128 hidden rows x 64 B each) or KB2 K-blocks of W2
  // (C rows x 64 B each); one mbarrier wait per slot keeps the single MMA-issuing thread off the critical path
  static constexpr int SLOT = 3 * FBLK;
  static constexpr int KB1 = 3, P1 = XKB / KB1;              // W1 chunk = P1 slots
  static constexpr int KB2 = SLOT / (C * 64), P2 = HKB / KB2; // W2 chunk = P2 slots
  static constexpr int RING = RESIDENT ? 0 : 5;
  // the chunk pipeline runs across tile boundaries: that needs the next tile's y block and (TMEM permitting) a
  // second output accumulator
  static constexpr int XBUF = RESIDENT ? 2 : 1;  // streamed weights: fc1 runs two chunks ahead, which hides the reload of the single y buffer
  static constexpr int OBUF = (2 * FCH + FCH / 2 + 2 * C <= 512) ? 2 : 1;
  static constexpr int STAGE_WARPS = 4 * (C / 32 < 4 ? C / 32 : 4);     // epilogue warps that drain output columns
  // one 32-row x 64-byte staging tile per (32-column chunk of the output, TMEM lane quarter): the TMA producer drops the
  // tile's RESIDUAL rows there (same swizzle as the drain uses), the drain adds the accumulator in place and stores
  static constexpr int STAGE_TILES = 4 * (C / 32);
  static constexpr int STAGE_BYTES = STAGE_TILES * 32 * 64;
  static constexpr int VEC_BYTES = ((2 * HC + 2 * C) * 4 + 1023) / 1024 * 1024;   // b1, colsum1 (LN fold), b2, gamma
  static constexpr int SMEM = kFCtrl + VEC_BYTES + XBUF * X_BYTES + W_BYTES + RING * SLOT + STAGE_BYTES + 1024;
  static_assert(C % 32 == 0 && C <= 192 && HC % FCH == 0 && NCH >= 2, "unsupported width");
  static_assert(XKB % KB1 == 0 && KB2 * C * 64 == SLOT && HKB % KB2 == 0, "slot geometry");
  static_assert(2 * FCH + FCH / 2 + OBUF * C <= 512, "tensor memory budget");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// GCV_FUSED_DEBUG what-if switches only exist in builds with -DGCV_GEMM_WHATIF (GCV_NVCC_FLAGS, see build.py)
#ifdef GCV_GEMM_WHATIF
constexpr bool kFWhatIf = true;
#else
constexpr bool kFWhatIf = false;
#endif

struct FParams {
  int64_t M;
  int tiles;
  uint32_t idesc1, idesc2;
  const float* b1;
  const float* b2;
  const float* gamma;
  void* x;                 // [M, C] residual in / result out (in place)
  const float* ln_stats;   // folded LayerNorm (LN = true): [M][C/32] x (sum, sumsq) of the y rows; colsum1 [4C]
  const float* colsum1;
  float ln_eps;
  int debug;               // GCV_FUSED_DEBUG what-if switches (results are garbage): 1 = no GELU math, 2 = no residual
                           // loads, 3 = no output stores, 4 = no drain at all, 5 = no LN-statistics loads
};

#ifdef GCV_FUSED_TRACE
// debug build: per-role cycle counters of CTA 0 (wait sites + total), dumped to gcv_fused_trace[]
__device__ long long gcv_fused_trace[64];
#define TR_DECL long long tr_t0 = 0; long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_start = clock64();
#define TR_BEGIN tr_t0 = clock64();
#define TR_END(i) tr_acc[i] += clock64() - tr_t0;
#define TR_DUMP(base) if (blockIdx.x == 0) { for (int i_ = 0; i_ < 7; ++i_) gcv_fused_trace[(base) + i_] = tr_acc[i_]; gcv_fused_trace[(base) + 7] = clock64() - tr_start; }
#else
#define TR_DECL
#define TR_BEGIN
#define TR_END(i)
#define TR_DUMP(base)
#endif

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tcgen05.mma with the A operand in tensor memory (16-bit elements, two per 32-bit column, lane = row)
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// 16 consecutive 32-bit columns of this warp's 32 lanes
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint4& a, const uint4& b, const uint4& c, const uint4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
      "r"(d.x), "r"(d.y), "r"(d.z), "r"(d.w)
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <typename T, int C, bool LN>
__global__ void __launch_bounds__(kFThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_w1,
                 const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_x, const FParams p) {
  const int dbg = kFWhatIf ? p.debug : 0;
  using K = Cfg<C>;
  constexpr int NRING = K::RING > 0 ? K::RING : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  // control block
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase);
  uint64_t* x_full = bars + 0;        // [2]
  uint64_t* x_empty = bars + 2;       // [2]
  uint64_t* o_full = bars + 4;        // [2]
  uint64_t* o_empty = bars + 6;       // [2]
  uint64_t* s_full = bars + 8;        // [2]
  uint64_t* s_empty = bars + 10;      // [2]
  uint64_t* h_full = bars + 12;       // [1]  the single H buffer in tensor memory: chunk g's GELU output is in place
  uint64_t* h_empty = bars + 13;      // [2]  fc2 of chunk g has read it: barrier g & 1 (each epilogue group then waits on
                                      //      consecutive phases of ONE barrier -- parity waits must not skip a phase)
  uint64_t* w_full = bars + 15;       // [1]  resident weights have landed
  uint64_t* res_full = bars + 16 + 2 * NRING + 1;    // [1]  the tile's residual rows are in the staging tiles
  uint64_t* res_empty = res_full + 1;                // [1]  every draining warp has stored the tile
  uint64_t* ring_full = bars + 16;    // [RING]
  uint64_t* ring_empty = bars + 16 + NRING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16 + 2 * NRING);
  float* vec_b1 = reinterpret_cast<float*>(gbase + kFCtrl);           // [HC]
  float* vec_s1 = vec_b1 + K::HC;                                      // [HC] column sums of W1 (LN fold)
  float* vec_b2 = vec_s1 + K::HC;                                      // [C]
  float* vec_g = vec_b2 + C;                                           // [C]
  const uint32_t x_smem = base + kFCtrl + K::VEC_BYTES;               // X0 (, X1)
  const uint32_t w_smem = x_smem + K::XBUF * K::X_BYTES;              // resident: W1 chunks, then W2 chunks
  const uint32_t ring_smem = w_smem + K::W_BYTES;
  uint8_t* stage_gen = gbase + kFCtrl + K::VEC_BYTES + K::XBUF * K::X_BYTES + K::W_BYTES + K::RING * K::SLOT;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (tc_ptx.cuh elect_one)
  const int lane = threadIdx.x & 31;
  const int my_tiles = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = my_tiles * K::NCH;            // hidden chunks this CTA processes, numbered g = 0 .. total-1

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(x_full + i), 1);
      mbar_init(smem_u32(x_empty + i), 1);
      mbar_init(smem_u32(o_full + i), 1);
      mbar_init(smem_u32(o_empty + i), kFEpiWarps);
      mbar_init(smem_u32(s_full + i), 1);
      mbar_init(smem_u32(s_empty + i), kFEpiWarps / 2);     // chunk buffer i belongs to epilogue group i (8 warps)
    }
    mbar_init(smem_u32(h_full), kFEpiWarps / 2);            // the group that owns the chunk
    mbar_init(smem_u32(h_empty + 0), 1);
    mbar_init(smem_u32(h_empty + 1), 1);
    mbar_init(smem_u32(w_full), 1);
    mbar_init(smem_u32(res_full), 1);
    mbar_init(smem_u32(res_empty), K::STAGE_WARPS);
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
  for (int i = threadIdx.x; i < K::HC; i += kFThreads) {
    vec_b1[i] = 0.5f * __ldg(p.b1 + i);            // the GELU helpers work on half the pre-activation
    if constexpr (LN) vec_s1[i] = __ldg(p.colsum1 + i);
  }
  for (int i = threadIdx.x; i < C; i += kFThreads) {
    vec_b2[i] = __ldg(p.b2 + i);
    vec_g[i] = __ldg(p.gamma + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_h = tmem_base + 2 * FCH;                 // 64 columns: 128 hidden values of a chunk, two per column
  const uint32_t tmem_o = tmem_base + 2 * FCH + FCH / 2;       // O0 (, O1 = O0 + C)

  // Buffer bookkeeping shared by all roles.  Tile i of this CTA uses y buffer i % XBUF for the (i / XBUF)-th time and
  // output accumulator i % OBUF for the (i / OBUF)-th time; chunk g uses S buffer g & 1 for the (g >> 1)-th time and the
  // single H buffer for the g-th time.
  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop; one elected lane issues) =====================
    {
      int slot = 0;
      uint32_t rphase = 0;
      auto issue_x = [&](int ti) {
        const int xb = ti % K::XBUF;
        const uint32_t use = (uint32_t)(ti / K::XBUF);
        mbar_wait(smem_u32(x_empty + xb), (use & 1) ^ 1);
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        if (elect_one()) {
          mbar_expect_tx(smem_u32(x_full + xb), K::X_BYTES);
#pragma unroll 1
          for (int kb = 0; kb < K::XKB; ++kb)
            tma_load_2d(x_smem + xb * K::X_BYTES + kb * FBLK, &tm_y, smem_u32(x_full + xb), kb * FKB, tile * FM);
        }
        __syncwarp();
      };
      // the residual rows of tile ti ([128, C] of x) go into the staging tiles once the drain of tile ti-1 has left them
      const uint32_t stage_smem = ring_smem + K::RING * K::SLOT;
      auto issue_res = [&](int ti) {
        mbar_wait(smem_u32(res_empty), ((uint32_t)ti & 1) ^ 1);
        const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
        if (elect_one()) {
          mbar_expect_tx(smem_u32(res_full), K::STAGE_BYTES);
#pragma unroll 1
          for (int t = 0; t < K::STAGE_TILES; ++t)       // tile t = (column chunk t >> 2, lane quarter t & 3)
            tma_load_2d(stage_smem + t * 2048, &tm_x, smem_u32(res_full), (t >> 2) * 32, tile * FM + (t & 3) * 32);
        }
        __syncwarp();
      };
      if (my_tiles > 0) issue_x(0);
      if constexpr (K::RESIDENT) {
        // both weight matrices, once: W1 chunk j = rows [128 j, 128 j + 128) in XKB K-blocks; W2 chunk j = columns
        // [128 j, 128 j + 128) of all C rows in 4 K-blocks
        if (my_tiles > 0 && elect_one()) {
          mbar_expect_tx(smem_u32(w_full), K::W_BYTES);
#pragma unroll 1
          for (int j = 0; j < K::NCH; ++j) {
#pragma unroll 1
            for (int kb = 0; kb < K::XKB; ++kb)
              tma_load_2d(w_smem + j * K::W1_CHUNK + kb * FBLK, &tm_w1, smem_u32(w_full), kb * FKB, j * FCH);
#pragma unroll 1
            for (int kb = 0; kb < K::HKB; ++kb)
              tma_load_2d(w_smem + K::NCH * K::W1_CHUNK + j * K::W2_CHUNK + kb * (C * 64), &tm_w2, smem_u32(w_full),
                          j * FCH + kb * FKB, 0);
          }
        }
        __syncwarp();
        // y blocks stay two tiles ahead of the residual rows: residual(ti) has to wait for the drain of tile ti-1 (which
        // runs a chunk into tile ti), and nothing the MMA warp needs may queue behind that wait
        if (my_tiles > 1) issue_x(1);
#pragma unroll 1
        for (int ti = 0; ti < my_tiles; ++ti) {
          issue_res(ti);
          if (ti + 2 < my_tiles) issue_x(ti + 2);
        }
      } else {
#pragma unroll 1
        for (int g = 0; g < total + 2; ++g) {          // same order as the MMA warp: W1 of chunk g, then W2 of chunk g-2
          if (g < total) {
            const int ti = g / K::NCH, j = g - ti * K::NCH;
#pragma unroll 1
            for (int part = 0; part < K::P1; ++part) {
              mbar_wait(smem_u32(ring_empty + slot), rphase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(smem_u32(ring_full + slot), K::SLOT);
#pragma unroll
                for (int kb = 0; kb < K::KB1; ++kb)
                  tma_load_2d(ring_smem + slot * K::SLOT + kb * FBLK, &tm_w1, smem_u32(ring_full + slot),
                              (part * K::KB1 + kb) * FKB, j * FCH);
              }
              __syncwarp();
              if (++slot == NRING) { slot = 0; rphase ^= 1; }
            }
            // next tile's y block: as soon as this tile's fc1s can retire
            if (ti + 1 < my_tiles && j == K::NCH - 1) issue_x(ti + 1);
            // this tile's residual rows, two chunks before its drain starts (the previous tile's drain is long over, so
            // the wait inside never holds up the weight stream)
            if (j == K::NCH - 2) issue_res(ti);
          }
          if (g >= 2) {
            const int gj = g - 2, jj = gj % K::NCH;
#pragma unroll 1
            for (int part = 0; part < K::P2; ++part) {
              mbar_wait(smem_u32(ring_empty + slot), rphase ^ 1);
              if (elect_one()) {
                mbar_expect_tx(smem_u32(ring_full + slot), K::SLOT);
#pragma unroll
                for (int kb = 0; kb < K::KB2; ++kb)
                  tma_load_2d(ring_smem + slot * K::SLOT + kb * (C * 64), &tm_w2, smem_u32(ring_full + slot),
                              jj * FCH + (part * K::KB2 + kb) * FKB, 0);
              }
              __syncwarp();
              if (++slot == NRING) { slot = 0; rphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp runs the loop; one elected lane issues) =====================
    {
      int slot = 0;
      uint32_t rphase = 0;
      // descriptor bases: only the 14-bit start-address field changes between operands
      const uint64_t dx0 = umma_desc_kmajor<64>(x_smem), dw0 = umma_desc_kmajor<64>(w_smem), dr0 = umma_desc_kmajor<64>(ring_smem);
      if constexpr (K::RESIDENT) {
        if (my_tiles > 0) mbar_wait(smem_u32(w_full), 0);
        tc_fence_after();
      }
      // fc1 runs two chunks ahead of fc2: S[g & 1] is free as soon as the epilogue has pulled chunk g-2 out of TMEM,
      // long before that chunk's GELU output is back in H -- so the next chunk's accumulator is always waiting for
      // the epilogue group instead of the other way round
#pragma unroll 1
      for (int g = 0; g < total + 2; ++g) {
        if (g < total) {
          const int ti = g / K::NCH, j = g - ti * K::NCH;
          const int xb = ti % K::XBUF;
          if (j == 0) mbar_wait(smem_u32(x_full + xb), (uint32_t)(ti / K::XBUF) & 1);
          const int b = g & 1;
          mbar_wait(smem_u32(s_empty + b), ((uint32_t)(g >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d = tmem_base + b * FCH;
          if constexpr (K::RESIDENT) {
            const uint64_t da = dx0 + (uint64_t)((xb * K::X_BYTES) >> 4);
            const uint64_t dw = dw0 + (uint64_t)((j * K::W1_CHUNK) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int kb = 0; kb < K::XKB; ++kb)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  tc_mma(d, da + ((kb * FBLK + k * 32) >> 4), dw + ((kb * FBLK + k * 32) >> 4), p.idesc1, (kb | k) ? 1u : 0u);
              tc_commit(smem_u32(s_full + b));
              if (j == K::NCH - 1) tc_commit(smem_u32(x_empty + xb));        // y buffer may be refilled
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int part = 0; part < K::P1; ++part) {
              mbar_wait(smem_u32(ring_full + slot), rphase);
              tc_fence_after();
              const uint64_t da = dx0 + (uint64_t)((xb * K::X_BYTES + part * K::KB1 * FBLK) >> 4);
              const uint64_t dw = dr0 + (uint64_t)((slot * K::SLOT) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < K::KB1; ++kb)
#pragma unroll
                  for (int k = 0; k < 2; ++k)
                    tc_mma(d, da + ((kb * FBLK + k * 32) >> 4), dw + ((kb * FBLK + k * 32) >> 4), p.idesc1,
                           (part | kb | k) ? 1u : 0u);
                tc_commit(smem_u32(ring_empty + slot));
                if (part == K::P1 - 1) {
                  tc_commit(smem_u32(s_full + b));
                  if (j == K::NCH - 1) tc_commit(smem_u32(x_empty + xb));      // y buffer may be refilled
                }
              }
              __syncwarp();
              if (++slot == NRING) { slot = 0; rphase ^= 1; }
            }
          }
        }
        if (g >= 2) {
          const int gj = g - 2, ti = gj / K::NCH, jj = gj - ti * K::NCH;
          const int ob = ti % K::OBUF;
          mbar_wait(smem_u32(h_full), (uint32_t)gj & 1);
          tc_fence_after();
          if (jj == 0) {
            mbar_wait(smem_u32(o_empty + ob), ((uint32_t)(ti / K::OBUF) & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t dout = tmem_o + ob * C;
          // A = H from tensor memory: K slice k (16 hidden values) = 8 columns
          if constexpr (K::RESIDENT) {
            const uint64_t dw = dw0 + (uint64_t)((K::NCH * K::W1_CHUNK + jj * K::W2_CHUNK) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int kb = 0; kb < K::HKB; ++kb)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  tc_mma_ts(dout, tmem_h + (uint32_t)((kb * 2 + k) * 8), dw + ((kb * C * 64 + k * 32) >> 4), p.idesc2,
                            (jj | kb | k) ? 1u : 0u);
              tc_commit(smem_u32(h_empty + (gj & 1)));
              if (jj == K::NCH - 1) tc_commit(smem_u32(o_full + ob));
            }
            __syncwarp();
          } else {
#pragma unroll 1
            for (int part = 0; part < K::P2; ++part) {
              mbar_wait(smem_u32(ring_full + slot), rphase);
              tc_fence_after();
              const uint64_t dw = dr0 + (uint64_t)((slot * K::SLOT) >> 4);
              if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < K::KB2; ++kb)
#pragma unroll
                  for (int k = 0; k < 2; ++k)
                    tc_mma_ts(dout, tmem_h + (uint32_t)(((part * K::KB2 + kb) * 2 + k) * 8), dw + ((kb * C * 64 + k * 32) >> 4),
                              p.idesc2, (jj | part | kb | k) ? 1u : 0u);
                tc_commit(smem_u32(ring_empty + slot));
                if (part == K::P2 - 1) {
                  tc_commit(smem_u32(h_empty + (gj & 1)));
                  if (jj == K::NCH - 1) tc_commit(smem_u32(o_full + ob));
                }
              }
              __syncwarp();
              if (++slot == NRING) { slot = 0; rphase ^= 1; }
            }
          }
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    // The 16 warps form two groups of 8; group g owns the chunks with (chunk index & 1) == g, i.e. TMEM buffer S[g].
    // While one group sits in a TMEM-load / barrier latency the other one issues math.
    const int q = warp & 3;                  // TMEM lane quarter = rows 32q .. 32q+31 of the tile
    const int grp = (warp - 2) >> 3;
    const int half = ((warp - 2) >> 2) & 1;  // which 64 hidden columns (2 K-blocks) of the group's chunk
    const int s = (warp - 2) >> 2;           // column-chunk phase of the O epilogue
    const int row = q * 32 + lane;
    T* xg = reinterpret_cast<T*>(p.x);
    // this warp's share of tile ti's output accumulator: + b2, * gamma, + residual, staged coalesced store (in place)
    auto drain_output = [&](int ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int ob = ti % K::OBUF;
      const int64_t m_warp = (int64_t)tile * FM + q * 32;
      if (dbg == 4) {
        mbar_wait(smem_u32(o_full + ob), (uint32_t)(ti / K::OBUF) & 1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(o_empty + ob));
        if (s < C / 32) {
          mbar_wait(smem_u32(res_full), (uint32_t)ti & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(res_empty));
        }
        return;
      }
      // ---- + b2, * gamma, + residual, staged coalesced store (in place on x) ----
      mbar_wait(smem_u32(o_full + ob), (uint32_t)(ti / K::OBUF) & 1);
      tc_fence_after();
      if (s >= C / 32) {                     // this warp owns no output columns: release the accumulator and leave
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(o_empty + ob));
        return;
      }
      mbar_wait(smem_u32(res_full), (uint32_t)ti & 1);     // the residual rows of this tile sit in the staging tiles
#pragma unroll
      for (int ci = 0; ci < 2; ++ci) {
        const int c = s + 4 * ci;
        if (c >= C / 32) break;
        uint8_t* my_stage = stage_gen + (c * 4 + q) * (32 * 64);
        float v[32];
        {
          uint32_t r[32];
          tc_ld32(tmem_o + ob * C + ((uint32_t)(q * 32) << 16) + c * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        }
        if (c + 4 >= C / 32) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(o_empty + ob));
        }
        const int n0 = c * 32;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          const int n = n0 + gq * 8;
          float* w = v + gq * 8;
          const float4 b0 = *reinterpret_cast<const float4*>(vec_b2 + n), b1 = *reinterpret_cast<const float4*>(vec_b2 + n + 4);
          const float4 g0 = *reinterpret_cast<const float4*>(vec_g + n), g1 = *reinterpret_cast<const float4*>(vec_g + n + 4);
          uint4* slot = reinterpret_cast<uint4*>(my_stage + lane * 64 + ((gq ^ ((lane >> 1) & 3)) << 4));
          float rr[8];
          {
            const uint4 rq = dbg == 2 ? make_uint4(0, 0, 0, 0) : *slot;      // residual row `lane`, columns n .. n+7
            float2 f;
            f = unpack2<T>(rq.x); rr[0] = f.x; rr[1] = f.y;
            f = unpack2<T>(rq.y); rr[2] = f.x; rr[3] = f.y;
            f = unpack2<T>(rq.z); rr[4] = f.x; rr[5] = f.y;
            f = unpack2<T>(rq.w); rr[6] = f.x; rr[7] = f.y;
          }
          // (w + b) * g + r as w * g + (b * g + r): two FFMA2 per column pair
          const float2 y0 = fma2(make_float2(w[0], w[1]), make_float2(g0.x, g0.y), fma2(make_float2(b0.x, b0.y), make_float2(g0.x, g0.y), make_float2(rr[0], rr[1])));
          const float2 y1 = fma2(make_float2(w[2], w[3]), make_float2(g0.z, g0.w), fma2(make_float2(b0.z, b0.w), make_float2(g0.z, g0.w), make_float2(rr[2], rr[3])));
          const float2 y2 = fma2(make_float2(w[4], w[5]), make_float2(g1.x, g1.y), fma2(make_float2(b1.x, b1.y), make_float2(g1.x, g1.y), make_float2(rr[4], rr[5])));
          const float2 y3 = fma2(make_float2(w[6], w[7]), make_float2(g1.z, g1.w), fma2(make_float2(b1.z, b1.w), make_float2(g1.z, g1.w), make_float2(rr[6], rr[7])));
          uint4 pk;
          pk.x = pack2<T>(y0.x, y0.y); pk.y = pack2<T>(y1.x, y1.y);
          pk.z = pack2<T>(y2.x, y2.y); pk.w = pack2<T>(y3.x, y3.y);
          *slot = pk;                                                        // in place: same thread, same 16 bytes
        }
        __syncwarp();
        const int piece = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + (lane >> 2);
          const int64_t mm = m_warp + r;
          if (mm < p.M && dbg != 3) {
            const uint4 pk = *reinterpret_cast<const uint4*>(my_stage + r * 64 + ((piece ^ ((r >> 1) & 3)) << 4));
            *reinterpret_cast<uint4*>(xg + mm * C + n0 + piece * 8) = pk;
          }
        }
        __syncwarp();
      }
      // the staging tiles of this warp may take the next tile's residual rows: this warp's generic-proxy writes / reads
      // are ordered before the producer's next bulk copy into them
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(res_empty));
    };
    int pending = -1;
    int ln_ti = -1;
    float2 lnrs = make_float2(1.0f, 0.0f);   // folded LayerNorm: (rstd, -mean * rstd) of this thread's row of tile ln_ti
    float2 ln_raw[8];                        // partial sums of the FOLLOWING tile's row: loaded a tile ahead, reduced at the tile change
#pragma unroll
    for (int i = 0; i < 8; ++i) ln_raw[i] = make_float2(0.0f, 1.0f);
#pragma unroll 1
    for (int g = grp; g < total; g += 2) {
      const int ti = g / K::NCH, j = g - ti * K::NCH;
      const int b = grp;
      if constexpr (K::OBUF == 1) {
        // One output accumulator and one H buffer: fc2 of the next tile's first chunk waits for EVERY warp's drain of this
        // tile, and the chunk after it (owned by the group that had the tile's LAST chunk) waits (h_empty) for that fc2 --
        // a circular wait if that group drained one chunk late like the other one does.  It drains first instead; its
        // accumulator wait is short (the tile's last fc2 was issued as soon as this group stored its H).
        if (pending >= 0 && grp == ((K::NCH - 1) & 1)) { drain_output(pending); pending = -1; }
      }
      if constexpr (LN) {
        if (ti != ln_ti) {
          if (ln_ti < 0) {
            const int64_t m = (int64_t)((int)blockIdx.x + ti * (int)gridDim.x) * FM + row;
            if (m < p.M && dbg != 5) ln_row_load(p.ln_stats, m, C / 32, ln_raw);
          }
          lnrs = ln_row_finish(ln_raw, C / 32, C, p.ln_eps);
          lnrs.x *= 0.5f; lnrs.y *= 0.5f;
          ln_ti = ti;
          const int64_t mn = (int64_t)((int)blockIdx.x + (ti + 1) * (int)gridDim.x) * FM + row;
          if (mn < p.M && dbg != 5) ln_row_load(p.ln_stats, mn, C / 32, ln_raw);      // consumed at the next tile change
        }
      }
      const uint32_t n_use = (uint32_t)(g >> 1);
      mbar_wait(smem_u32(s_full + b), n_use & 1);
      tc_fence_after();
#pragma unroll 1
      for (int hk = 0; hk < 2; ++hk) {
        const int kb = half * 2 + hk;        // K-block of the chunk = 32 hidden columns
        float v[32];
        const float* bj = vec_b1 + j * FCH + kb * 32;
        // the first 16 columns' bias / column-sum vectors are fetched while the TMEM load is in flight (the asm statements
        // around it are compiler barriers: otherwise every shared-memory load is issued after the wait)
        float4 pb[4], ps[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pb[e] = *reinterpret_cast<const float4*>(bj + 4 * e);
          if constexpr (LN) ps[e] = *reinterpret_cast<const float4*>(bj + K::HC + 4 * e);
        }
        {
          uint32_t r[32];
          tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + b * FCH + kb * 32, r);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
        }
        if (hk == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(s_empty + b));
        }
        uint4 pk[4];
        if (dbg == 1) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq)
            pk[gq] = make_uint4(__float_as_uint(v[gq * 8]), __float_as_uint(v[gq * 8 + 2]), __float_as_uint(v[gq * 8 + 4]), __float_as_uint(v[gq * 8 + 6]));
        } else
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          if constexpr (LN) {
            const float* sj = bj + K::HC;    // vec_s1 follows vec_b1
            pk[gq] = ln_bias_gelu_pack8<T>(v + gq * 8, lnrs.x, lnrs.y,
                                           gq < 2 ? ps[2 * gq] : *reinterpret_cast<const float4*>(sj + gq * 8),
                                           gq < 2 ? ps[2 * gq + 1] : *reinterpret_cast<const float4*>(sj + gq * 8 + 4),
                                           gq < 2 ? pb[2 * gq] : *reinterpret_cast<const float4*>(bj + gq * 8),
                                           gq < 2 ? pb[2 * gq + 1] : *reinterpret_cast<const float4*>(bj + gq * 8 + 4));
          } else {
            pk[gq] = bias_gelu_pack8<T>(v + gq * 8, gq < 2 ? pb[2 * gq] : *reinterpret_cast<const float4*>(bj + gq * 8),
                                        gq < 2 ? pb[2 * gq + 1] : *reinterpret_cast<const float4*>(bj + gq * 8 + 4));
          }
        }
        if (hk == 0) {                       // the single H buffer: fc2 of the previous chunk (g - 1) has retired;
          // that is completion number (g - 1) / 2 of barrier (g - 1) & 1 (g = 0: parity 1 of a fresh barrier passes)
          mbar_wait(smem_u32(h_empty + ((g - 1) & 1)), (uint32_t)((g - 1) >> 1) & 1u);
          tc_fence_after();
        }
        // 32 hidden values of this thread's row = 16 packed columns of H
        tc_st16(tmem_h + ((uint32_t)(q * 32) << 16) + (uint32_t)(kb * 16), pk[0], pk[1], pk[2], pk[3]);
      }
      tc_wait_st();
      tc_fence_before();                                           // TMEM writes ordered before the MMA warp's reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(h_full));

      // The tile's output accumulator is complete only after the OTHER group's last chunk went through fc2, so this
      // warp drains tile ti one chunk late (after its first chunk of the next tile) instead of idling on o_full.
      if (pending >= 0) { drain_output(pending); pending = -1; }
      if ((g + 2) / K::NCH != ti) {                   // that was this group's last chunk of tile ti
        pending = ti;
      }
    }
    if (pending >= 0) drain_output(pending);
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

template <typename T, int C, bool LN>
int launch_fused_impl(int dtype, const void* y, const void* w1, const void* w2, const FParams& p, cudaStream_t stream) {
  using K = Cfg<C>;
  CUtensorMap my, m1, m2;
  int rc = fused_map(&my, dtype, y, p.M, C, FM);
  if (rc) return rc;
  if ((rc = fused_map(&m1, dtype, w1, K::HC, C, FCH))) return rc;
  if ((rc = fused_map(&m2, dtype, w2, C, K::HC, C))) return rc;
  CUtensorMap mx;                      // residual rows: 32-column x 32-row boxes of x [M, C], same 64B swizzle as the drain's staging
  if ((rc = fused_map(&mx, dtype, p.x, p.M, C, 32))) return rc;
  static unsigned long long attr_devs = 0;
  if (first_on_device(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fused_kernel<T, C, LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(fused MLP smem=%d): %s", K::SMEM, cudaGetErrorString(e));
      return GCV_ERR_CUDA;
    }
  }
  const int sms = device_sms();
  const int grid = p.tiles < sms ? p.tiles : sms;
  mlp_fused_kernel<T, C, LN><<<grid, kFThreads, K::SMEM, stream>>>(my, m1, m2, mx, p);
  return check_launch("mlp_fused");
}

template <typename T, int C>
int launch_fused(int dtype, const void* y, const void* w1, const void* w2, const FParams& p, cudaStream_t stream) {
  return p.ln_stats ? launch_fused_impl<T, C, true>(dtype, y, w1, w2, p, stream)
                    : launch_fused_impl<T, C, false>(dtype, y, w1, w2, p, stream);
}

}  // namespace

int mlp_fused_trace(long long* out64) {
#ifdef GCV_FUSED_TRACE
  return cudaMemcpyFromSymbol(out64, gcv_fused_trace, 64 * sizeof(long long)) == cudaSuccess ? 64 : -1;
#else
  (void)out64;
  return 0;
#endif
}

bool mlp_fused_supported(int dtype, int C) { return (dtype == GCV_BF16 || dtype == GCV_F16) && (C == 96 || C == 192); }

int mlp_fused(int dtype, const void* y, const float* ln_stats, float ln_eps, const void* w1, const float* b1,
              const float* colsum1, const void* w2, const float* b2, const float* gamma, void* x, int64_t M, int C,
              cudaStream_t stream) {
  GCV_REQUIRE(mlp_fused_supported(dtype, C), "mlp_fused: bf16/fp16 and C in {96,192} only (dtype=%d C=%d)", dtype, C);
  GCV_REQUIRE(M > 0 && y && w1 && w2 && b1 && b2 && gamma && x, "mlp_fused: bad arguments");
  GCV_REQUIRE(!ln_stats || colsum1, "mlp_fused: the folded LayerNorm needs the column sums of w1");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  GCV_REQUIRE(al16(y) && al16(w1) && al16(w2) && al16(x), "mlp_fused: pointers must be 16-byte aligned");
  FParams p{};
  p.M = M;
  p.tiles = (int)((M + FM - 1) / FM);
  p.idesc1 = umma_idesc_f16(dtype == GCV_BF16, FM, FCH);
  // (measured: kind::f16 with an fp16 A operand in tensor memory and a bf16 B operand is an illegal instruction on
  // sm_100a -- H has to be converted to the model's 16-bit type)
  p.idesc2 = umma_idesc_f16(dtype == GCV_BF16, FM, C);
  p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.x = x;
  p.ln_stats = ln_stats; p.colsum1 = colsum1; p.ln_eps = ln_eps;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("GCV_FUSED_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  if (dtype == GCV_BF16) {
    return C == 96 ? launch_fused<__nv_bfloat16, 96>(dtype, y, w1, w2, p, stream)
                   : launch_fused<__nv_bfloat16, 192>(dtype, y, w1, w2, p, stream);
  }
  return C == 96 ? launch_fused<__half, 96>(dtype, y, w1, w2, p, stream)
                 : launch_fused<__half, 192>(dtype, y, w1, w2, p, stream);
}

}  // namespace gcv
