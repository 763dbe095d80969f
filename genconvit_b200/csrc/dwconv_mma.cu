// Depthwise 7x7 convolution (pad 3, + bias) on the tensor cores, 16-bit activations, NHWC.
// timm==0.6.5 ConvNeXtBlock.conv_dw as reached from reference model/genconvit_ed.py:68,
// model/genconvit_vae.py:97 (arithmetic restated in oracle/backbones.py).
//
// A depthwise convolution has no contraction over channels, so the usual implicit GEMM does not
// exist; on CUDA cores it costs 49 FMAs per output and is instruction-issue bound (the FFMA2
// column kernel in convnext_ops.cu sits at 35 % of the FMA roof).  Here every channel is treated
// as its own small banded problem instead: for one image row and one channel,
//     out[y][x0 + n] = sum_dy  in[y + dy - 3][x0 - 3 + k] * T_dy[k][n],   T_dy[k][n] = w[dy][k - n]
// with T_dy a 16 x 8 banded Toeplitz matrix (7 non-zero diagonals).  Stacking 16 images in the M
// dimension gives a dense m16n8k16 tensor-core MMA per (channel, dy): 7/16 of its MACs are useful,
// which is still ~6x fewer issued instructions per output than FFMA2, and the fp32 accumulation is
// exact for 16-bit inputs.  One warp owns one channel pair (a 32-bit shared-memory word holds
// both channels of a pixel) and marches down the image: each input row's A fragments are loaded
// once and feed the 7 output rows in flight (7 x 2 MMAs), so all data reuse happens in registers.
//
// CTA = 16 warps = one 32-channel chunk of a (16 images x 8 output columns) tile, all rows.
// Input rows arrive through a TMA ring (box = 32 ch x 17 px x 1 row x 16 images, 64B-swizzled;
// out-of-image rows/columns/images are zero-filled by TMA, which is the conv padding).  Finished
// output rows go through a padded shared-memory tile so that global stores are 16 bytes per lane
// and 64 contiguous bytes per pixel.  The taps are rounded to the activation type (like every
// other weight of the 16-bit modes); LayerNorm runs as a separate row pass over the stored result.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace gcv {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn mma_get_encode() {
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

constexpr int MM_IMGS = 16;                 // M rows of the MMA = images
constexpr int MM_XOUT = 8;                  // N = output columns per tile
constexpr int MM_XIN = 17;                  // staged input columns per tile (16 used; 17 spreads the swizzle phase)
constexpr int MM_CCH = 32;                  // channels per CTA pass = 16 warps x 2
constexpr int MM_THREADS = 512;
constexpr uint32_t MM_SLOT_BYTES = MM_IMGS * MM_XIN * MM_CCH * 2;      // 17408 = 17 KiB (keeps slots 1024-aligned)
constexpr int MM_STAGE_STRIDE = 17;         // words per staged output pixel (16 + 1 pad: conflict-light both ways)
constexpr uint32_t MM_STAGE_BYTES = MM_IMGS * MM_XOUT * MM_STAGE_STRIDE * 4;   // 8704
constexpr int MM_DEPTH = 8;                 // TMA ring slots (power of two)

__device__ __forceinline__ void mm_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CTA b works on channel chunk b % n_chunks for the whole launch (its Toeplitz fragments are loaded once) and walks
// the (image group, column tile) tiles b / n_chunks + i * (gridDim / n_chunks): the CTAs that run side by side cover
// all chunks of the same pixels, so DRAM and L2 see whole pixels rather than 64 of every 2C bytes.
// Pipeline step n of a CTA = input row (n % H) of its (n / H)-th tile.  The loop body is kept to ~70 instructions per
// warp and step (the first version spent 400, mostly index arithmetic, and was issue-bound on exactly that).
template <typename T>
__device__ __forceinline__ void mma16816_init(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c));
}

template <typename T>
__global__ void __launch_bounds__(MM_THREADS, 1)
dwconv7_mma_kernel(const __grid_constant__ CUtensorMap tm_x, T* __restrict__ y, float2* __restrict__ stats,
                   const float* __restrict__ taps,
                   const float* __restrict__ bias, int B, int H, int W, int C, int xtiles, int n_tiles,
                   int n_chunks) {
  extern __shared__ uint8_t msm_dyn[];
  // layout (1024-aligned so that the swizzle phase of a slot starts at 0): [ring 8 x 17408][stage 2 x 8704][full barriers]
  uint8_t* msm_raw = msm_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(msm_dyn) & 1023u)) & 1023u);
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(msm_raw);
  const uint32_t stage_s = smem_base + MM_DEPTH * MM_SLOT_BYTES;
  const uint32_t bar_s = stage_s + 2 * MM_STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cc = (int)blockIdx.x % n_chunks, tile0 = (int)blockIdx.x / n_chunks, tile_step = (int)gridDim.x / n_chunks;
  const int my_subs = tile0 < n_tiles ? (n_tiles - tile0 + tile_step - 1) / tile_step : 0;
  const int total_steps = my_subs * H;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MM_DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s + 8 * i), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer state (thread 0 only; kept in shared memory to save registers in the other 511 threads): the next
  // pipeline step to fetch and where it lies
  int* pst = reinterpret_cast<int*>(msm_raw + MM_DEPTH * MM_SLOT_BYTES + 2 * MM_STAGE_BYTES + 8 * MM_DEPTH);
  if (threadIdx.x == 0) {
    pst[0] = 0; pst[1] = 0; pst[2] = tile0; pst[3] = (tile0 % xtiles) * MM_XOUT - 3; pst[4] = (tile0 / xtiles) * MM_IMGS;
  }
  auto issue_next = [&]() {
    int p_n = pst[0], p_yin = pst[1], p_tile = pst[2];
    const int p_x = pst[3], p_b = pst[4];
    if (p_n >= total_steps) return;
    const uint32_t fb = bar_s + 8 * (p_n & (MM_DEPTH - 1));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(MM_SLOT_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_base + (p_n & (MM_DEPTH - 1)) * MM_SLOT_BYTES), "l"(reinterpret_cast<uint64_t>(&tm_x)), "r"(fb),
          "r"(cc * MM_CCH), "r"(p_x), "r"(p_yin), "r"(p_b)
        : "memory");
    pst[0] = p_n + 1;
    if (++p_yin == H) {
      p_yin = 0;
      p_tile += tile_step;
      pst[2] = p_tile;
      pst[3] = (p_tile % xtiles) * MM_XOUT - 3;
      pst[4] = (p_tile / xtiles) * MM_IMGS;
    }
    pst[1] = p_yin;
  };
  if (threadIdx.x == 0)
    for (int k = 0; k < MM_DEPTH; ++k) issue_next();

  // A-fragment byte offsets inside a ring slot (64B swizzle: byte-address bits 4-5 ^= bits 7-8):
  // index r*4 + q: image g + 8r, input column 2t + {0, 1, 8, 9}[q]
  uint32_t aoff[8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 2 * t + (q & 1) + ((q >> 1) << 3);
      const uint32_t lin = (uint32_t)(((g + 8 * r) * MM_XIN + k) * (MM_CCH * 2) + warp * 4);
      aoff[r * 4 + q] = smem_base + (lin ^ (((lin >> 7) & 3u) << 4));
    }
  // output staging: this lane's accumulators are pixels (image g / g+8, column 2t / 2t+1) of channel pair `warp`;
  // read-back role: thread i -> staged pixel i >> 2, channels 8*(i & 3) .. +7 of the chunk
  uint32_t st_s = stage_s + 4u * (uint32_t)((g * MM_XOUT + 2 * t) * MM_STAGE_STRIDE + warp);
  const int rb_px = threadIdx.x >> 2, rb_q = threadIdx.x & 3;
  uint32_t rb_s = stage_s + 4u * (uint32_t)(rb_px * MM_STAGE_STRIDE + 4 * rb_q);

  const int c0 = cc * MM_CCH + 2 * warp;
  // Toeplitz B fragments: b[h] holds T_dy[k = 2t + 8h + {0,1}][n = g] = w[dy][k - n] (zero off the band)
  uint32_t bfrag[2][7][2];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch)
#pragma unroll
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int dx0 = 2 * t + 8 * h - g;
        const float v0 = (dx0 >= 0 && dx0 <= 6) ? __ldg(taps + (dy * 7 + dx0) * C + c0 + ch) : 0.0f;
        const float v1 = (dx0 + 1 >= 0 && dx0 + 1 <= 6) ? __ldg(taps + (dy * 7 + dx0 + 1) * C + c0 + ch) : 0.0f;
        bfrag[ch][dy][h] = pack2<T>(v0, v1);
      }
  const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + c0));

  // The words of pipeline step n are fetched from the ring one step ahead, so their shared-memory latency and bank
  // conflicts overlap the previous step's MMAs.
  int n = 0;
  uint32_t wn[8];
  auto prefetch = [&](int nn) {
    if (nn < total_steps) {
      const uint32_t slot = (uint32_t)nn & (MM_DEPTH - 1);
      mm_mbar_wait(bar_s + 8 * slot, ((uint32_t)nn / MM_DEPTH) & 1u);
      const uint32_t so = slot * MM_SLOT_BYTES;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(wn[j]) : "r"(aoff[j] + so));
    }
  };
  prefetch(0);

  uint32_t st_flip = MM_STAGE_BYTES, rb_flip = MM_STAGE_BYTES;
  const uint32_t row_bytes = (uint32_t)W * (uint32_t)C * 2u;
  bool pending = false, rb_ok = false;
  uint32_t rb_off = 0;
  // staging tile -> global memory for the most recently parked output row (+ its LayerNorm partial sums)
  auto readback = [&]() {
    uint4 q;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(q.x) : "r"(rb_s));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(q.y) : "r"(rb_s + 4));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(q.z) : "r"(rb_s + 8));
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(q.w) : "r"(rb_s + 12));
    if (rb_ok) *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(y) + rb_off) = q;
    if (stats) {
      // LayerNorm partials of this chunk: sum and sum of squares over its 32 channels, taken from the rounded
      // values the consumer will read; the 4 lanes of a pixel hold 8 channels each
      const float2 f0 = unpack2<T>(q.x), f1 = unpack2<T>(q.y), f2 = unpack2<T>(q.z), f3 = unpack2<T>(q.w);
      float sm = ((f0.x + f0.y) + (f1.x + f1.y)) + ((f2.x + f2.y) + (f3.x + f3.y));
      float sq = fmaf(f0.x, f0.x, f0.y * f0.y) + fmaf(f1.x, f1.x, f1.y * f1.y) +
                 (fmaf(f2.x, f2.x, f2.y * f2.y) + fmaf(f3.x, f3.x, f3.y * f3.y));
      sm += __shfl_xor_sync(0xffffffffu, sm, 1); sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sm += __shfl_xor_sync(0xffffffffu, sm, 2); sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      // for the lane with rb_q == 0, rb_off = 64 * (pixel * n_chunks + cc)
      if (rb_ok && rb_q == 0)
        *reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats) + (rb_off >> 3)) = make_float2(sm, sq);
    }
    rb_off += row_bytes;
    rb_s += rb_flip; rb_flip = 0u - rb_flip;
  };
  const int nsteps7 = ((H + 6 + 6) / 7) * 7;   // H + 6 steps, rounded up to whole turns of the 7-slot accumulator ring
  for (int i = 0; i < my_subs; ++i) {
    const int tile = tile0 + i * tile_step;
    const int ig = tile / xtiles, xt = tile - ig * xtiles;
    const int rb_b = ig * MM_IMGS + (rb_px >> 3), rb_x = xt * MM_XOUT + (rb_px & 7);
    rb_ok = rb_b < B && rb_x < W;
    rb_off = ((uint32_t)(rb_b * H) * (uint32_t)W + (uint32_t)rb_x) * (uint32_t)(C * 2) + (uint32_t)(cc * MM_CCH + 8 * rb_q) * 2u;   // output row 0

    float acc[7][2][4];                        // [output-row slot][channel][fragment]
    // rows 0..2 never see a dy = 0 tap (their first input row is row 0), so they start from the bias here; every
    // other output row is initialised by its dy = 0 MMA (C operand = bias)
#pragma unroll
    for (int sl = 0; sl < 3; ++sl) {
      acc[sl][0][0] = acc[sl][0][1] = acc[sl][0][2] = acc[sl][0][3] = bv.x;
      acc[sl][1][0] = acc[sl][1][1] = acc[sl][1][2] = acc[sl][1][3] = bv.y;
    }
    // step s consumes input row s - 3 (when inside the image) and retires output row s - 6.  Order inside a step:
    // issue this step's MMAs, then (while they execute) move the row retired in the PREVIOUS step from the staging tile
    // to global memory, then park the row retired now in the other staging tile, barrier.
    for (int s0 = 0; s0 < nsteps7; s0 += 7) {
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int s = s0 + u;
        const bool load = s >= 3 && s < H + 3;
        if (load) {
          uint32_t a0[4], a1[4];               // channel c0 / c0 + 1
          // fragment order: (row g, k lo), (row g+8, k lo), (row g, k hi), (row g+8, k hi)
          a0[0] = __byte_perm(wn[0], wn[1], 0x5410); a1[0] = __byte_perm(wn[0], wn[1], 0x7632);
          a0[1] = __byte_perm(wn[4], wn[5], 0x5410); a1[1] = __byte_perm(wn[4], wn[5], 0x7632);
          a0[2] = __byte_perm(wn[2], wn[3], 0x5410); a1[2] = __byte_perm(wn[2], wn[3], 0x7632);
          a0[3] = __byte_perm(wn[6], wn[7], 0x5410); a1[3] = __byte_perm(wn[6], wn[7], 0x7632);
          prefetch(n + 1);
          // dy = 6 first: it completes the row that is parked below, so that store waits for one MMA, not fourteen
#pragma unroll
          for (int dy = 6; dy >= 1; --dy) {
            const int sl = (u - dy + 7) % 7;   // output row s - dy lives in slot (s - dy) mod 7
            mma16816<T>(acc[sl][0], a0, bfrag[0][dy][0], bfrag[0][dy][1]);
            mma16816<T>(acc[sl][1], a1, bfrag[1][dy][0], bfrag[1][dy][1]);
          }
          mma16816_init<T>(acc[u][0], a0, bfrag[0][0][0], bfrag[0][0][1], bv.x);      // output row s, slot s mod 7
          mma16816_init<T>(acc[u][1], a1, bfrag[1][0][0], bfrag[1][0][1], bv.y);
          ++n;
        }
        if (pending) { readback(); pending = false; }
        const int done = (u + 1) % 7;          // slot of output row s - 6
        const bool emit = s >= 6 && s < H + 6;
        if (emit) {
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_s), "r"(pack2<T>(acc[done][0][0], acc[done][1][0])) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_s + 4 * MM_STAGE_STRIDE), "r"(pack2<T>(acc[done][0][1], acc[done][1][1])) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_s + 4 * 8 * MM_XOUT * MM_STAGE_STRIDE), "r"(pack2<T>(acc[done][0][2], acc[done][1][2])) : "memory");
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(st_s + 4 * (8 * MM_XOUT + 1) * MM_STAGE_STRIDE), "r"(pack2<T>(acc[done][0][3], acc[done][1][3])) : "memory");
          st_s += st_flip; st_flip = 0u - st_flip;                 // other staging buffer next time
          pending = true;
        }
        __syncthreads();
        // every warp has turned this step's input row into fragments: refill its ring slot (step n - 1 + 8)
        if (threadIdx.x == 0 && load) issue_next();
      }
    }
    if (pending) { readback(); pending = false; }    // the tile's last row (rb_off / rb_ok change with the tile)
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Second generation (default): same banded-Toeplitz formulation and the same (16 images x 8 output columns x 32
// channels) CTA tile, rebuilt around what the first kernel's profile showed (ncu source page, stage 0): 2780 warp
// instructions per pipeline step of which 224 are MMAs, issue slots 47 % busy, stalls spread over short-scoreboard,
// dependency waits and the barrier -- an instruction-count / latency problem, not a bandwidth one.  Per step:
//   * a warp owns FOUR channels (two adjacent channel-pair words = one 8-byte piece of the pixel): one LDS.64 feeds
//     two channel pairs, 8 compute warps (up to 255 registers: 7 rows x 4 channels x 4 accumulators in flight);
//   * the bias rides in the MMA: input columns 14 / 15 of the K = 16 window only ever meet zero taps, so they are
//     replaced by 1.0 in registers and row 14 / 15 of the centre-row (dy = 3) Toeplitz matrix carries the bias split
//     into a 16-bit high and low part (~2^-17 relative); a fresh accumulator is then started by its dy = 0 MMA with
//     C = 0 -- no per-step accumulator initialisation, and nothing behind the 15-pixel TMA box can leak in
//     (12 % less L2 -> SM traffic than the 17-pixel box);
//   * finished rows are parked in a TMA-swizzled staging tile and leave as ONE bulk tensor store per step (the
//     hardware clips columns / images past the tensor);
//   * the LayerNorm partial sums come from the tensor cores as well: per 16 staged pixels, ldmatrix gives the
//     [pixel x channel] tile as an A operand; A x ones = channel sums, and diag(A x A^T) = sums of squares, where
//     the B operand A^T is a subset of the very same registers (rows of the staged tile are K-contiguous).  Exact
//     products, fp32 accumulation, ~20 instructions per warp instead of ~75 of unpack / add / shuffle code.
constexpr int M4_WARPS = 8;
constexpr int M4_THREADS = 32 * M4_WARPS;
constexpr int M4_XIN = 15;
constexpr uint32_t M4_SLOT_BYTES = MM_IMGS * M4_XIN * MM_CCH * 2;     // 15360 (a multiple of 1024: swizzle phase 0)
constexpr uint32_t M4_STAGE_BYTES = MM_IMGS * MM_XOUT * MM_CCH * 2;   // 8192: 128 pixels x 64 B, TMA SWIZZLE_64B layout
constexpr int M4_STAGES = 3;

// CTA barrier as plain PTX: the two warp groups reach it from different program points (warp-uniform control flow)
__device__ __forceinline__ void m4_cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// The MMAs of this kernel are plain (non-volatile) asm statements: pure register operations, so the compiler is free to
// interleave the step's conversion / shared-memory / address instructions between them (with two warps per scheduler the
// MMA issue cadence leaves most issue slots empty, and the scalar code's dependency latencies are otherwise exposed).
template <typename T>
__device__ __forceinline__ void mma16816_zero(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  const float z = 0.0f;
  if constexpr (std::is_same<T, __half>::value)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(z));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(z));
}
template <typename T>
__device__ __forceinline__ void mma16816_acc(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
__global__ void __launch_bounds__(M4_THREADS, 1)
dwconv7_mma4_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                    float2* __restrict__ stats, const float* __restrict__ taps, const float* __restrict__ bias, int B,
                    int H, int W, int C, int xtiles, int n_tiles, int n_chunks) {
  extern __shared__ uint8_t msm_dyn[];
  // [ring 8 x 15360][stage 3 x 8192][full barriers 8 x 8]; 1024-aligned (TMA swizzle phase 0 for slots and staging)
  uint8_t* msm_raw = msm_dyn + ((1024u - ((uint32_t)__cvta_generic_to_shared(msm_dyn) & 1023u)) & 1023u);
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(msm_raw);
  const uint32_t stage_s = smem_base + MM_DEPTH * M4_SLOT_BYTES;
  const uint32_t bar_s = stage_s + M4_STAGES * M4_STAGE_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int cc = (int)blockIdx.x % n_chunks, tile0 = (int)blockIdx.x / n_chunks, tile_step = (int)gridDim.x / n_chunks;
  const int my_subs = tile0 < n_tiles ? (n_tiles - tile0 + tile_step - 1) / tile_step : 0;
  const int total_steps = my_subs * H;
  if (total_steps == 0) return;

  if (threadIdx.x == 0) {
    for (int i = 0; i < MM_DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s + 8 * i), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // producer (thread 0): next pipeline step to fetch and where it lies; pipeline step n = input row n % H of tile n / H
  int p_n = 0, p_yin = 0, p_tile = tile0, p_x = (tile0 % xtiles) * MM_XOUT - 3, p_b = (tile0 / xtiles) * MM_IMGS;
  auto issue_next = [&]() {
    if (p_n >= total_steps) return;
    const uint32_t fb = bar_s + 8 * (p_n & (MM_DEPTH - 1));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(M4_SLOT_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_base + (p_n & (MM_DEPTH - 1)) * M4_SLOT_BYTES), "l"(reinterpret_cast<uint64_t>(&tm_x)), "r"(fb),
          "r"(cc * MM_CCH), "r"(p_x), "r"(p_yin), "r"(p_b)
        : "memory");
    ++p_n;
    if (++p_yin == H) {
      p_yin = 0;
      p_tile += tile_step;
      p_x = (p_tile % xtiles) * MM_XOUT - 3;
      p_b = (p_tile / xtiles) * MM_IMGS;
    }
  };
  if (threadIdx.x == 0)
    for (int k = 0; k < MM_DEPTH; ++k) issue_next();

  // A-fragment source offsets inside a ring slot (64B swizzle: byte-address bits 4-5 ^= bits 7-8); index r*4 + q:
  // image g + 8r, input column 2t + {0, 1, 8, 9}[q]; the 8 bytes at +8*warp hold this warp's two channel pairs
  uint32_t aoff[8];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int k = 2 * t + (q & 1) + ((q >> 1) << 3);
      const uint32_t lin = (uint32_t)(((g + 8 * r) * M4_XIN + k) * (MM_CCH * 2) + warp * 8);
      aoff[r * 4 + q] = smem_base + (lin ^ (((lin >> 7) & 3u) << 4));
    }
  const bool t3 = t == 3;                    // this lane's high-K fragments are input columns 14, 15: replaced by 1.0
  const uint32_t one2 = pack2<T>(1.0f, 1.0f);

  // output staging (TMA SWIZZLE_64B: pixel p owns bytes [64 p, 64 p + 64), its 16-byte piece c sits at c ^ ((p >> 1) & 3)).
  // This lane's accumulators are pixels (image g / g + 8, column 2t / 2t + 1) of channels 4*warp..+3 = 8 bytes per
  // pixel.  Lanes with odd g park column 2t + 1 first and 2t second (even g the other way round): the 16 lanes of a
  // store phase then split over both 64-byte bank halves (2-way instead of 4-way conflicts).
  const uint32_t e1 = (uint32_t)(g & 1);
  uint32_t st_a, st_b;            // first / second parked pixel of image g (image g + 8: + 64 * 64 bytes, same swizzle)
  {
    const uint32_t pa = (uint32_t)(g * MM_XOUT + 2 * t) + e1, pb = (uint32_t)(g * MM_XOUT + 2 * t) + (1u - e1);
    st_a = stage_s + pa * 64u + ((((uint32_t)warp >> 1) ^ ((pa >> 1) & 3u)) << 4) + ((uint32_t)warp & 1u) * 8u;
    st_b = stage_s + pb * 64u + ((((uint32_t)warp >> 1) ^ ((pb >> 1) & 3u)) << 4) + ((uint32_t)warp & 1u) * 8u;
  }
  // statistics role: warp w reduces staged pixels 16w .. 16w+15 (images 2w, 2w+1 of the tile).  ldmatrix row addresses:
  // lane l supplies row (l & 7) + 8 ((l >> 3) & 1) of 16-byte piece (l >> 4) [channels 0-15] / 2 + (l >> 4) [16-31]
  uint32_t lm0, lm1;
  {
    const uint32_t p = (uint32_t)(16 * warp + (lane & 7) + 8 * ((lane >> 3) & 1)), c = (uint32_t)(lane >> 4);
    lm0 = stage_s + p * 64u + ((c ^ ((p >> 1) & 3u)) << 4);
    lm1 = stage_s + p * 64u + (((c + 2u) ^ ((p >> 1) & 3u)) << 4);
  }
  // lanes (g, t = g >> 1) end up holding the diagonal of A x A^T: they write the partial sums of pixels g and g + 8
  const bool st_lane = t == (g >> 1);
  const bool st_odd = (g & 1) != 0;

  const int c0 = cc * MM_CCH + 4 * warp;
  // Toeplitz B fragments: b[h] holds T_dy[k = 2t + 8h + {0,1}][n = g] = w[dy][k - n] (zero off the band); rows 14 / 15
  // of T_3 (lanes t == 3, h == 1) carry (bias - hi, hi) with hi = bias rounded to the activation type
  uint32_t bfrag[4][7][2];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
    for (int dy = 0; dy < 7; ++dy)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int dx0 = 2 * t + 8 * h - g;
        const float v0 = (dx0 >= 0 && dx0 <= 6) ? __ldg(taps + (dy * 7 + dx0) * C + c0 + ch) : 0.0f;
        const float v1 = (dx0 + 1 >= 0 && dx0 + 1 <= 6) ? __ldg(taps + (dy * 7 + dx0 + 1) * C + c0 + ch) : 0.0f;
        bfrag[ch][dy][h] = pack2<T>(v0, v1);
      }
    if (t3) {
      const float bf = __ldg(bias + c0 + ch);
      const float hi = to_f<T>(from_f<T>(bf));
      bfrag[ch][3][1] = pack2<T>(bf - hi, hi);
    }
  }

  // The 8-byte words of pipeline step n are fetched from the ring one step ahead of their use.  Past the last step the
  // last one is fetched again (its barrier phase stays complete): the loads are unconditional, nobody reads them.
  int n = 0;
  uint2 wn[8];
  auto prefetch_wait = [&](int nn) {          // the spin loop is kept out of the step's main basic block
    nn = nn < total_steps ? nn : total_steps - 1;
    mm_mbar_wait(bar_s + 8 * ((uint32_t)nn & (MM_DEPTH - 1)), ((uint32_t)nn / MM_DEPTH) & 1u);
  };
  auto prefetch = [&](int nn) {
    nn = nn < total_steps ? nn : total_steps - 1;
    const uint32_t so = ((uint32_t)nn & (MM_DEPTH - 1)) * M4_SLOT_BYTES;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(wn[j].x), "=r"(wn[j].y) : "r"(aoff[j] + so));
  };
  prefetch_wait(0);
  prefetch(0);

  // Three staging buffers: the row parked in step s leaves (bulk store + partial sums) during step s + 1 and its
  // buffer is written again in step s + 3, so thread 0 only ever waits for the store issued a whole step earlier.
  uint32_t st_buf = 0, rd_buf = 0;           // byte offsets (0, 8192, 16384) of the buffer parked next / retired next
  bool pending = false;
  // the tile the parked rows belong to (a tile's last row is retired while the next tile already runs)
  int r_row = 0, r_x = 0, r_b = 0;           // output row of the next row to retire, tile origin (column, image)
  bool ok0 = false, ok1 = false;
  uint32_t st_off0 = 0, st_off1 = 0;         // byte offsets of this lane's two pixels in the partial-sum array
  const uint32_t st_row = (uint32_t)W * (uint32_t)n_chunks * 8u;

  // the row parked in the previous step: bulk store to y (one elected thread), partial sums (every warp, 16 pixels)
  auto retire = [&]() {
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                   ::"l"(reinterpret_cast<uint64_t>(&tm_y)), "r"(stage_s + rd_buf), "r"(cc * MM_CCH), "r"(r_x), "r"(r_row),
                     "r"(r_b)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // all but the store just issued have read their staging buffer
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    if (stats) {
      uint32_t a[4], b[4];
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(lm0 + rd_buf));
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                   : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(lm1 + rd_buf));
      float ds[4], q0[4], q1[4];
      mma16816_zero<T>(ds, a, one2, one2);           // channel sums (every column of D the same)
      mma16816_acc<T>(ds, b, one2, one2);
      mma16816_zero<T>(q0, a, a[0], a[2]);           // A x (rows 0-7 of A)^T: diagonal = sums of squares of pixels 0-7
      mma16816_acc<T>(q0, b, b[0], b[2]);
      mma16816_zero<T>(q1, a, a[1], a[3]);           // A x (rows 8-15)^T: entries (8 + n, n) = pixels 8-15
      mma16816_acc<T>(q1, b, b[1], b[3]);
      if (st_lane) {
        if (ok0) *reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats) + st_off0) = make_float2(ds[0], st_odd ? q0[1] : q0[0]);
        if (ok1) *reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(stats) + st_off1) = make_float2(ds[2], st_odd ? q1[3] : q1[2]);
      }
      st_off0 += st_row; st_off1 += st_row;
    }
    ++r_row;
    rd_buf = rd_buf == 2 * M4_STAGE_BYTES ? 0u : rd_buf + M4_STAGE_BYTES;
  };

  const int nsteps7 = ((H + 6 + 6) / 7) * 7;   // H + 6 steps, rounded up to whole turns of the 7-slot accumulator ring
  for (int i = 0; i < my_subs; ++i) {
    const int tile = tile0 + i * tile_step;
    const int ig = tile / xtiles, xt = tile - ig * xtiles;

    float acc[7][4][4];                        // [output-row slot][channel][fragment]
    // rows 0..2 never see a dy = 0 tap (their first input row is row 0): they start from zero here; every other output
    // row is started by its dy = 0 MMA (C = 0); the bias arrives with the dy = 3 MMA every row receives
#pragma unroll
    for (int sl = 0; sl < 3; ++sl)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[sl][ch][0] = acc[sl][ch][1] = acc[sl][ch][2] = acc[sl][ch][3] = 0.0f;

    // step s consumes input row s - 3 (when inside the image) and retires output row s - 6
    for (int s0 = 0; s0 < nsteps7; s0 += 7) {
#pragma unroll
      for (int u = 0; u < 7; ++u) {
        const int s = s0 + u;
        const bool load = s >= 3 && s < H + 3;
        const int done = (u + 1) % 7;          // slot of output row s - 6
        const bool emit = s >= 6 && s < H + 6;
        // everything that is not an MMA: retiring the row parked in the previous step, parking the row this step completes
        auto rest_of_step = [&]() {
          if (pending) { retire(); pending = false; }
          if (s == 0) {                        // rows parked from here on belong to this tile
            r_x = xt * MM_XOUT; r_b = ig * MM_IMGS; r_row = 0;
            const int col = r_x + g, b0 = r_b + 2 * warp;            // pixel g of image 2w, pixel g + 8 = image 2w + 1
            ok0 = col < W && b0 < B; ok1 = col < W && b0 + 1 < B;
            const uint32_t pix0 = (uint32_t)(b0 * H) * (uint32_t)W + (uint32_t)col;
            st_off0 = (pix0 * (uint32_t)n_chunks + (uint32_t)cc) * 8u;
            st_off1 = st_off0 + (uint32_t)H * st_row;
          }
          if (emit) {
            // pixel (g, 2t) = fragments [0] of the four channels, (g, 2t + 1) = [1], (g + 8, 2t) = [2], (g + 8, 2t + 1) = [3]
            const uint32_t p00 = pack2<T>(acc[done][0][0], acc[done][1][0]), p01 = pack2<T>(acc[done][2][0], acc[done][3][0]);
            const uint32_t p10 = pack2<T>(acc[done][0][1], acc[done][1][1]), p11 = pack2<T>(acc[done][2][1], acc[done][3][1]);
            const uint32_t p20 = pack2<T>(acc[done][0][2], acc[done][1][2]), p21 = pack2<T>(acc[done][2][2], acc[done][3][2]);
            const uint32_t p30 = pack2<T>(acc[done][0][3], acc[done][1][3]), p31 = pack2<T>(acc[done][2][3], acc[done][3][3]);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_a + st_buf), "r"(e1 ? p10 : p00), "r"(e1 ? p11 : p01) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_b + st_buf), "r"(e1 ? p00 : p10), "r"(e1 ? p01 : p11) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_a + st_buf + 64u * 64u), "r"(e1 ? p30 : p20), "r"(e1 ? p31 : p21) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(st_b + st_buf + 64u * 64u), "r"(e1 ? p20 : p30), "r"(e1 ? p21 : p31) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> visible to the bulk store
            st_buf = st_buf == 2 * M4_STAGE_BYTES ? 0u : st_buf + M4_STAGE_BYTES;
            pending = true;
          }
        };
        if (load) {
          prefetch_wait(n + 1);
          uint32_t a[4][4];                    // [channel][fragment: (g, k lo), (g+8, k lo), (g, k hi), (g+8, k hi)]
          a[0][0] = __byte_perm(wn[0].x, wn[1].x, 0x5410); a[1][0] = __byte_perm(wn[0].x, wn[1].x, 0x7632);
          a[2][0] = __byte_perm(wn[0].y, wn[1].y, 0x5410); a[3][0] = __byte_perm(wn[0].y, wn[1].y, 0x7632);
          a[0][1] = __byte_perm(wn[4].x, wn[5].x, 0x5410); a[1][1] = __byte_perm(wn[4].x, wn[5].x, 0x7632);
          a[2][1] = __byte_perm(wn[4].y, wn[5].y, 0x5410); a[3][1] = __byte_perm(wn[4].y, wn[5].y, 0x7632);
          a[0][2] = t3 ? one2 : __byte_perm(wn[2].x, wn[3].x, 0x5410); a[1][2] = t3 ? one2 : __byte_perm(wn[2].x, wn[3].x, 0x7632);
          a[2][2] = t3 ? one2 : __byte_perm(wn[2].y, wn[3].y, 0x5410); a[3][2] = t3 ? one2 : __byte_perm(wn[2].y, wn[3].y, 0x7632);
          a[0][3] = t3 ? one2 : __byte_perm(wn[6].x, wn[7].x, 0x5410); a[1][3] = t3 ? one2 : __byte_perm(wn[6].x, wn[7].x, 0x7632);
          a[2][3] = t3 ? one2 : __byte_perm(wn[6].y, wn[7].y, 0x5410); a[3][3] = t3 ? one2 : __byte_perm(wn[6].y, wn[7].y, 0x7632);
          prefetch(n + 1);
          // dy = 6 first: it completes the row that is parked in this step
#pragma unroll
          for (int dy = 6; dy >= 1; --dy) {
            const int sl = (u - dy + 7) % 7;   // output row s - dy lives in slot (s - dy) mod 7
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) mma16816_acc<T>(acc[sl][ch], a[ch], bfrag[ch][dy][0], bfrag[ch][dy][1]);
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)       // output row s, slot s mod 7
            mma16816_zero<T>(acc[u][ch], a[ch], bfrag[ch][0][0], bfrag[ch][0][1]);
          ++n;
          // every warp turned step n - 2 into fragments before the barrier that ended the previous step: its ring slot
          // is free; the refill (step n - 2 + 8) is issued here, behind this step's MMAs
          if (threadIdx.x == 0 && n >= 2) issue_next();
        }
        rest_of_step();
        m4_cta_barrier();
      }
    }
  }
  // the last tile's last row: both groups have parked it
  if (pending) retire();
  // the global writes of the bulk stores must complete before the CTA exits
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

}  // namespace

bool dwconv7_mma_supported(int dtype, int C) {
  return (dtype == GCV_BF16 || dtype == GCV_F16) && C % MM_CCH == 0 && C >= MM_CCH;
}

// y = conv_dw(x) + bias (no LayerNorm), x / y: [B,H,W,C] of `dtype`, taps: [49,C] fp32.
// stats (optional): [B*H*W][C/32] float2 = (sum, sum of squares) of y over each 32-channel chunk.
int dwconv7_mma(int dtype, const void* x, void* y, float* stats, const float* taps, const float* bias, int B, int H,
                int W, int C, cudaStream_t stream) {
  GCV_REQUIRE(dwconv7_mma_supported(dtype, C), "dwconv7_mma: needs a 16-bit dtype and C %% 32 == 0 (C=%d)", C);
  GCV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "dwconv7_mma: x, y must be 16-byte aligned");
  const int xtiles = (W + MM_XOUT - 1) / MM_XOUT, igroups = (B + MM_IMGS - 1) / MM_IMGS;
  const int n_chunks = C / MM_CCH;
  const int64_t n_tiles64 = (int64_t)xtiles * igroups;
  GCV_REQUIRE(n_tiles64 * H < 2147483647LL, "dwconv7_mma: too many pipeline steps");
  GCV_REQUIRE(((int64_t)igroups * MM_IMGS + 1) * H * W * C * 2 < 4294967295LL, "dwconv7_mma: tensor larger than 4 GiB");
  const int n_tiles = (int)n_tiles64;
  const int sms = device_sms();
  GCV_REQUIRE(n_chunks <= sms, "dwconv7_mma: C=%d has more channel chunks than the device has SMs", C);
  const int tile_slots = sms / n_chunks;                    // tiles in flight; every one gets all its channel chunks
  const int grid = (n_tiles < tile_slots ? n_tiles : tile_slots) * n_chunks;
  // GCV_DWCONV_MMA=1 keeps the first-generation kernel (16 warps x 2 channels; A/B timing)
  static int gen_env = -1;
  if (gen_env < 0) { const char* e = getenv("GCV_DWCONV_MMA"); gen_env = e ? atoi(e) : 2; }
  const bool gen2 = gen_env != 1;
  const size_t smem = gen2 ? (size_t)MM_DEPTH * M4_SLOT_BYTES + M4_STAGES * M4_STAGE_BYTES + 8 * MM_DEPTH + 1024
                           : (size_t)MM_DEPTH * MM_SLOT_BYTES + 2 * MM_STAGE_BYTES + 8 * MM_DEPTH + 32 + 1024;
  CUtensorMap tm;
  {
    EncodeTiledFn enc = mma_get_encode();
    if (!enc) {
      set_error("cuTensorMapEncodeTiled not resolvable (no CUDA driver?)");
      return GCV_ERR_NO_DRIVER;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {MM_CCH, (cuuint32_t)(gen2 ? M4_XIN : MM_XIN), 1, MM_IMGS};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapDataType tdt = dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(&tm, tdt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dwconv7_mma: cuTensorMapEncodeTiled failed: CUresult %d (B=%d H=%d W=%d C=%d)", (int)r, B, H, W, C);
      return GCV_ERR_CUDA;
    }
  }
  CUtensorMap tm_out = tm;
  if (gen2) {
    // the output leaves through bulk tensor stores of (32 channels x 8 pixels x 1 row x 16 images) boxes
    EncodeTiledFn enc = mma_get_encode();
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {MM_CCH, MM_XOUT, 1, MM_IMGS};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapDataType tdt = dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUresult r = enc(&tm_out, tdt, 4, y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("dwconv7_mma: cuTensorMapEncodeTiled (output) failed: CUresult %d (B=%d H=%d W=%d C=%d)", (int)r, B, H, W, C);
      return GCV_ERR_CUDA;
    }
  }
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
    if (gen2) {
      static unsigned long long attr_devs = 0;
      if (first_on_device(attr_devs))
        cudaFuncSetAttribute(dwconv7_mma4_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      dwconv7_mma4_kernel<T><<<grid, M4_THREADS, smem, stream>>>(tm, tm_out, reinterpret_cast<float2*>(stats), taps, bias, B, H, W,
                                                                C, xtiles, n_tiles, n_chunks);
    } else {
      static unsigned long long attr_devs = 0;
      if (first_on_device(attr_devs))
        cudaFuncSetAttribute(dwconv7_mma_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      dwconv7_mma_kernel<T><<<grid, MM_THREADS, smem, stream>>>(tm, reinterpret_cast<T*>(y), reinterpret_cast<float2*>(stats),
                                                               taps, bias, B, H, W, C, xtiles, n_tiles, n_chunks);
    }
    return check_launch("dwconv7_mma");
  };
  return dtype == GCV_BF16 ? launch(__nv_bfloat16{}) : launch(__half{});
}

}  // namespace gcv
