// Memory-bound ConvNeXt kernels (NHWC): depthwise 7x7 + LayerNorm, LayerNorm2d +
// 2x2 patchify, stem patchify, row LayerNorm, global-average-pool + LayerNorm.
// timm==0.6.5 ConvNeXt as reached from reference model/genconvit_ed.py:68,
// model/genconvit_vae.py:97 (arithmetic restated in oracle/backbones.py).
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace gcv {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn dw_get_encode() {
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

// ---------------------------------------------------------------------------------
// Depthwise 7x7 (pad 3) + bias + LayerNorm over C.
// CTA = 4 x 8 output pixels, all C channels, 128 threads.  Channels are processed in
// chunks of 64: the (4+6) x (8+6) input halo of the chunk is staged in shared
// memory; thread (row r, channel pair p) keeps 8 pixels x 2 channels of fp32
// accumulators and slides along the row so every staged value feeds 7 taps.  The
// conv outputs of the tile are parked in shared memory ([32][C] fp32) and a second
// phase normalises one pixel per warp with shuffle reductions and writes coalesced.
// ---------------------------------------------------------------------------------
constexpr int DW_TH = 4, DW_TW = 8, DW_CK = 64;
constexpr int DW_HH = DW_TH + 6, DW_HW = DW_TW + 6;

template <typename T>
__global__ void __launch_bounds__(128)
dwconv7_ln_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ taps,
                  const float* __restrict__ bias, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                  float eps, int H, int W, int C, int tiles_w, int tiles_h, int fuse_ln) {
  // fuse_ln = 0 (channel counts whose [32 pixels][C] fp32 tile does not fit in shared memory, i.e. convnext_large's
  // C = 1536 in the fp32 mode): conv + bias go straight to y, the caller runs layernorm_rows over y afterwards
  extern __shared__ __align__(16) uint8_t smem[];
  T* halo = reinterpret_cast<T*>(smem);                                     // [DW_HH][DW_HW][DW_CK]
  float* outs = reinterpret_cast<float*>(smem + DW_HH * DW_HW * DW_CK * sizeof(T));  // [32][C]

  int t = blockIdx.x;
  const int tw = t % tiles_w; t /= tiles_w;
  const int th = t % tiles_h; t /= tiles_h;
  const int b = t;
  const int h0 = th * DW_TH, w0 = tw * DW_TW;
  const T* xb = x + (int64_t)b * H * W * C;

  const int r = threadIdx.x >> 5;          // output row within the tile
  const int pr = threadIdx.x & 31;         // channel pair within the chunk

  for (int c0 = 0; c0 < C; c0 += DW_CK) {
    const int cw = min(DW_CK, C - c0);     // channels in this chunk (multiple of 32)
    // ---- stage the halo: 16-byte vectors, zero outside the image ----
    constexpr int VE = 16 / sizeof(T);
    const int vec_per_px = cw / VE;
    for (int i = threadIdx.x; i < DW_HH * DW_HW * vec_per_px; i += blockDim.x) {
      const int v = i % vec_per_px, px = i / vec_per_px;
      const int hy = px / DW_HW, hx = px - hy * DW_HW;
      const int gy = h0 + hy - 3, gx = w0 + hx - 3;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        val = *reinterpret_cast<const uint4*>(xb + ((int64_t)gy * W + gx) * C + c0 + v * VE);
      *reinterpret_cast<uint4*>(halo + px * DW_CK + v * VE) = val;
    }
    __syncthreads();
    if (2 * pr < cw) {
      const int c = c0 + 2 * pr;
      float acc0[DW_TW], acc1[DW_TW];
      const float b0 = bias[c], b1 = bias[c + 1];
#pragma unroll
      for (int i = 0; i < DW_TW; ++i) { acc0[i] = b0; acc1[i] = b1; }
#pragma unroll 1
      for (int dy = 0; dy < 7; ++dy) {
        float w0r[7], w1r[7];
#pragma unroll
        for (int dx = 0; dx < 7; ++dx) {
          const float2 wv = __ldg(reinterpret_cast<const float2*>(taps + (dy * 7 + dx) * C + c));
          w0r[dx] = wv.x; w1r[dx] = wv.y;
        }
        const T* row = halo + ((r + dy) * DW_HW) * DW_CK + 2 * pr;
#pragma unroll
        for (int ix = 0; ix < DW_HW; ++ix) {
          float v0, v1;
          if constexpr (sizeof(T) == 4) {
            const float2 f = *reinterpret_cast<const float2*>(row + ix * DW_CK);
            v0 = f.x; v1 = f.y;
          } else {
            const float2 f = unpack2<T>(*reinterpret_cast<const uint32_t*>(row + ix * DW_CK));
            v0 = f.x; v1 = f.y;
          }
#pragma unroll
          for (int dx = 0; dx < 7; ++dx) {
            const int ox = ix - dx;
            if (ox >= 0 && ox < DW_TW) {
              acc0[ox] = fmaf(v0, w0r[dx], acc0[ox]);
              acc1[ox] = fmaf(v1, w1r[dx], acc1[ox]);
            }
          }
        }
      }
      if (fuse_ln) {
#pragma unroll
        for (int i = 0; i < DW_TW; ++i)
          *reinterpret_cast<float2*>(outs + (r * DW_TW + i) * C + c) = make_float2(acc0[i], acc1[i]);
      } else if (h0 + r < H) {
#pragma unroll
        for (int i = 0; i < DW_TW; ++i) {
          if (w0 + i < W) {
            T* dst = y + (((int64_t)b * H + h0 + r) * W + w0 + i) * C + c;
            if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(dst) = make_float2(acc0[i], acc1[i]);
            else *reinterpret_cast<uint32_t*>(dst) = pack2<T>(acc0[i], acc1[i]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (!fuse_ln) return;

  // ---- LayerNorm: warp w normalises the 8 pixels of tile row w ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gy = h0 + warp;
  if (gy >= H) return;
  const float inv_c = 1.0f / (float)C;
  for (int i = 0; i < DW_TW; ++i) {
    const int gx = w0 + i;
    if (gx >= W) break;
    const float* o = outs + (warp * DW_TW + i) * C;
    float s = 0.0f;
    for (int c = 2 * lane; c < C; c += 64) { const float2 f = *reinterpret_cast<const float2*>(o + c); s += f.x + f.y; }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.0f;
    for (int c = 2 * lane; c < C; c += 64) {
      const float2 f = *reinterpret_cast<const float2*>(o + c);
      q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    T* dst = y + (((int64_t)b * H + gy) * W + gx) * C;
    for (int c = 2 * lane; c < C; c += 64) {
      const float2 f = *reinterpret_cast<const float2*>(o + c);
      const float2 g = __ldg(reinterpret_cast<const float2*>(ln_w + c));
      const float2 be = __ldg(reinterpret_cast<const float2*>(ln_b + c));
      const float a0 = (f.x - mean) * rstd * g.x + be.x, a1 = (f.y - mean) * rstd * g.y + be.y;
      if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(dst + c) = make_float2(a0, a1);
      else *reinterpret_cast<uint32_t*>(dst + c) = pack2<T>(a0, a1);
    }
  }
}

template <typename T>
__device__ __forceinline__ float2 ld_pair(const T* p) {
  if constexpr (sizeof(T) == 4) return __ldg(reinterpret_cast<const float2*>(p));
  else return unpack2<T>(__ldg(reinterpret_cast<const unsigned int*>(p)));
}

// ---------------------------------------------------------------------------------
// Depthwise 7x7 + bias + LayerNorm, column version (the one used for the ConvNeXt-T widths).
// The kernel is instruction-issue bound on CUDA cores (49 fp32 FMAs per output, no tensor-core
// form), so everything is arranged to make FMAs the majority of issued instructions:
//   * thread = (7-row group g, channel pair p) keeps a 7x7-pixel x 2-channel block of fp32
//     accumulators (98 registers); each of the 13 input rows it needs is loaded ONCE (13 32-bit
//     channel-pair loads, a warp reads 128 contiguous bytes per pixel) and feeds up to 7 output rows;
//   * the two channels of a pair are multiplied with one FFMA2 (fma.rn.f32x2): half the issue slots
//     of scalar FFMA at the same pipe throughput;
//   * taps live in shared memory ([49][C] fp32, conflict-free 8-byte reads);
//   * LayerNorm statistics: exact two-pass, half-warp shuffles -> per-half-warp partials in smem ->
//     fixed-order totals (deterministic, no atomics).
// One CTA = one 7-pixel-wide column strip of one image (all rows), so no input row is fetched from
// L2 more than ~13/7 times.  H <= 56 (blockDim = ceil(H/7) * C/2 <= 384).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d.x), "+f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

// smem tile of one pipeline step: [group][channel box][13 pixels][CBOX channels], written by TMA
template <int C> struct DwGeom {
  static constexpr int CBOX = C <= 192 ? C : 192;          // TMA box dims are limited to 256 elements
  static constexpr int NBOX = C / CBOX;
  // LayerNorm statistics through a transposed smem buffer (49 floats per thread) when it fits next to the taps
  // and the input ring; the widest stage (C = 768: 150 KB of taps) keeps the half-warp shuffle reduction
  static constexpr bool SMEM_STATS = C <= 384;
  static_assert(C % CBOX == 0, "channel boxes");
};

__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// Persistent: one CTA per SM walks over (image, 7-pixel column strip) work items.  The taps are fetched once per CTA
// with a bulk async copy; the TMA input ring runs ahead across strip boundaries, so neither the ~microsecond
// memory latency nor the per-strip prologue is exposed.
template <typename T, int C>
__global__ void __launch_bounds__(384, 1)
dwconv7_ln_col_kernel(const __grid_constant__ CUtensorMap tm_x, T* __restrict__ y, const float* __restrict__ taps,
                      const float* __restrict__ bias, const float* __restrict__ ln_w,
                      const float* __restrict__ ln_b, float eps, int H, int W, int strips_w, int n_strips, int depth) {
  using G = DwGeom<C>;
  extern __shared__ __align__(128) uint8_t dsm_raw[];
  constexpr int half_c = C >> 1, hw_per_group = half_c >> 4;
  const int groups = blockDim.x / half_c;
  const int n_hw = blockDim.x >> 4;                       // half-warps in the CTA
  // layout: [barriers 256 B][taps 49*C f32][seg partials n_hw*49 f32][totals 2*groups*49 f32]
  //         [thread partials blockDim*49 f32 (SMEM_STATS)][ring depth x step_bytes]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(dsm_raw);          // [8]
  uint64_t* empty_bar = full_bar + 8;                                  // [8]
  uint64_t* w_bar = full_bar + 16;
  float* wsm = reinterpret_cast<float*>(dsm_raw + 256);
  float* part = wsm + 49 * C;
  float* tot = part + n_hw * 49;
  float* tpart = tot + 2 * groups * 49;
  constexpr uint32_t tile_bytes = (13u * G::CBOX * (uint32_t)sizeof(T) + 127u) & ~127u;
  const uint32_t step_bytes = (uint32_t)groups * G::NBOX * tile_bytes;      // smem footprint of one step
  const uint32_t tx_bytes = (uint32_t)groups * 13 * C * (uint32_t)sizeof(T); // bytes TMA actually delivers
  const uint32_t ring_off =
      (uint32_t)((256 + (49 * C + n_hw * 49 + 2 * groups * 49 + (G::SMEM_STATS ? (int)blockDim.x * 49 : 0)) * 4 + 127) & ~127);
  const uint8_t* ring = dsm_raw + ring_off;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(dsm_raw) + ring_off;

  const int g = threadIdx.x / half_c, c = 2 * (threadIdx.x - g * half_c);
  const int hw = threadIdx.x >> 4, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool reducer = (threadIdx.x & 15) == 0;
  const int my_strips = (n_strips - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total_steps = my_strips * 13;

  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(full_bar + i)), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(empty_bar + i)), "r"(nwarps));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(w_bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // one pipeline step = the 13-pixel input row segment every row group needs next: for the strip's step j, group gg
  // reads image row 7*gg - 3 + j.  Out-of-image rows / columns are zero-filled by TMA (that IS the conv padding).
  auto issue_step = [&](int n) {
    const int si = n / 13, j = n - si * 13;
    const int strip = (int)blockIdx.x + si * (int)gridDim.x;
    const int sb = strip / strips_w, sx0 = (strip - sb * strips_w) * 7;
    const int slot = n % depth;
    const uint32_t fb = (uint32_t)__cvta_generic_to_shared(full_bar + slot);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tx_bytes) : "memory");
    for (int gg = 0; gg < groups; ++gg)
#pragma unroll
      for (int cb = 0; cb < G::NBOX; ++cb) {
        const uint32_t dst = ring_s + slot * step_bytes + (gg * G::NBOX + cb) * tile_bytes;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm_x)), "r"(fb), "r"(cb * G::CBOX), "r"(sx0 - 3),
              "r"(7 * gg - 3 + j), "r"(sb)
            : "memory");
      }
  };
  if (threadIdx.x == 0) {
    const uint32_t wb = (uint32_t)__cvta_generic_to_shared(w_bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wb), "r"(49 * C * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((uint32_t)__cvta_generic_to_shared(wsm)), "l"(reinterpret_cast<uint64_t>(taps)), "r"(49 * C * 4), "r"(wb)
                 : "memory");
    for (int n = 0; n < depth && n < total_steps; ++n) issue_step(n);
  }

  const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + c));
  const float2 gw = __ldg(reinterpret_cast<const float2*>(ln_w + c));
  const float2 gb = __ldg(reinterpret_cast<const float2*>(ln_b + c));
  const float* wbase = wsm + c;
  const uint32_t my_off = (g * G::NBOX + c / G::CBOX) * tile_bytes + (c % G::CBOX) * (uint32_t)sizeof(T);
  const int64_t row_stride = (int64_t)W * C;
  const int oy0 = g * 7;
  dw_mbar_wait((uint32_t)__cvta_generic_to_shared(w_bar), 0);          // taps have landed

  int slot = 0;
  uint32_t parity = 0;
  int n = 0;
  for (int si = 0; si < my_strips; ++si) {
    const int strip = (int)blockIdx.x + si * (int)gridDim.x;
    const int b = strip / strips_w, x0 = (strip - b * strips_w) * 7;
    float2 acc[7][7];
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
      for (int i = 0; i < 7; ++i) acc[r][i] = bv;

#pragma unroll
    for (int j = 0; j < 13; ++j, ++n) {
      dw_mbar_wait((uint32_t)__cvta_generic_to_shared(full_bar + slot), parity);
      float2 v[13];
      const uint8_t* src = ring + slot * step_bytes + my_off;
#pragma unroll
      for (int ix = 0; ix < 13; ++ix) {
        if constexpr (sizeof(T) == 4) v[ix] = *reinterpret_cast<const float2*>(src + ix * G::CBOX * 4);
        else v[ix] = unpack2<T>(*reinterpret_cast<const uint32_t*>(src + ix * G::CBOX * 2));
      }
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(empty_bar + slot)) : "memory");
      if (threadIdx.x == 0 && n >= 1 && n - 1 + depth < total_steps) {
        // refill the slot of the PREVIOUS step (every warp has long copied it into registers by now, so this
        // wait almost never blocks warp 0) with step n-1+depth
        const int pslot = slot == 0 ? depth - 1 : slot - 1;
        const uint32_t pparity = slot == 0 ? parity ^ 1 : parity;
        dw_mbar_wait((uint32_t)__cvta_generic_to_shared(empty_bar + pslot), pparity);
        issue_step(n - 1 + depth);
      }
      if (++slot == depth) { slot = 0; parity ^= 1; }
#pragma unroll
      for (int r = 0; r < 7; ++r) {
        const int dy = j - r;
        if (dy >= 0 && dy < 7) {
#pragma unroll
          for (int dx = 0; dx < 7; ++dx) {
            const float2 wv = *reinterpret_cast<const float2*>(wbase + (dy * 7 + dx) * C);
#pragma unroll
            for (int ox = 0; ox < 7; ++ox) ffma2(acc[r][ox], v[ox + dx], wv);
          }
        }
      }
    }

    // ---- LayerNorm statistics, exact two-pass, deterministic ----
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if constexpr (G::SMEM_STATS) {
        // every thread parks its 49 per-pixel partials (row of 49 floats: odd stride, conflict-free both ways);
        // 16-thread segment sums and group totals are then plain strided smem reads instead of 392 shuffles
        float* mine = tpart + threadIdx.x * 49;
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
          for (int i = 0; i < 7; ++i)
            mine[r * 7 + i] = pass == 0 ? acc[r][i].x + acc[r][i].y
                                        : acc[r][i].x * acc[r][i].x + acc[r][i].y * acc[r][i].y;
        __syncthreads();
        for (int i = threadIdx.x; i < n_hw * 49; i += blockDim.x) {
          const int seg = i / 49, px = i - seg * 49;
          const float* col = tpart + seg * 16 * 49 + px;
          float t = 0.0f;
#pragma unroll
          for (int k = 0; k < 16; ++k) t += col[k * 49];
          part[i] = t;
        }
      } else {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            float t = pass == 0 ? acc[r][i].x + acc[r][i].y
                                : acc[r][i].x * acc[r][i].x + acc[r][i].y * acc[r][i].y;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (reducer) part[hw * 49 + r * 7 + i] = t;
          }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < groups * 49; i += blockDim.x) {
        const int gg = i / 49, px = i - gg * 49;
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < hw_per_group; ++k) t += part[(gg * hw_per_group + k) * 49 + px];
        tot[pass * groups * 49 + i] = pass == 0 ? t * (1.0f / (float)C) : rsqrtf(t * (1.0f / (float)C) + eps);
      }
      __syncthreads();
      if (pass == 0) {
#pragma unroll
        for (int r = 0; r < 7; ++r)
#pragma unroll
          for (int i = 0; i < 7; ++i) {
            const float mean = tot[g * 49 + r * 7 + i];
            acc[r][i].x -= mean; acc[r][i].y -= mean;
          }
      }
    }
    T* yp = y + (((int64_t)b * H + oy0) * W + x0) * C + c;
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      if (oy0 + r < H) {
#pragma unroll
        for (int i = 0; i < 7; ++i) {
          if (x0 + i < W) {
            const float rstd = tot[groups * 49 + g * 49 + r * 7 + i];
            const float o0 = acc[r][i].x * rstd * gw.x + gb.x, o1 = acc[r][i].y * rstd * gw.y + gb.y;
            T* dst = yp + i * C;
            if constexpr (sizeof(T) == 4) *reinterpret_cast<float2*>(dst) = make_float2(o0, o1);
            else *reinterpret_cast<uint32_t*>(dst) = pack2<T>(o0, o1);
          }
        }
      }
      yp += row_stride;
    }
    // the next strip's first statistics write to tpart / part / tot happens after its 13 barrier-free steps and is
    // ordered behind this strip's last reads by the barriers above plus the barrier that follows that write
  }
}

// ---------------------------------------------------------------------------------
// Per-row LayerNorm helpers: a group of LPR lanes (16 or 32) owns one row and keeps it in
// registers as 8-element (16-byte for 16-bit T) vectors: lane l holds vectors l, l+LPR, ...
// Two-pass statistics with shuffle reductions inside the group.
// ---------------------------------------------------------------------------------
template <typename T, int LPR, int ITER>
__device__ __forceinline__ void group_ln_row(const T* __restrict__ src, T* __restrict__ dst, const float* __restrict__ w,
                                             const float* __restrict__ b, float eps, int C, int gl) {
  float v[ITER][8];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
      load8<T>(src + c, v[i]);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) q += (v[i][e] - mean) * (v[i][e] - mean);
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
      float g[8], be[8];
      load8<float>(w + c, g);
      load8<float>(b + c, be);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = (v[i][e] - mean) * rstd * g[e] + be[e];
      store8<T>(dst + c, v[i]);
    }
  }
}

template <typename T, int LPR, int ITER>
__global__ void __launch_bounds__(256)
layernorm_rows_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ b, float eps, int64_t rows, int C) {
  constexpr int RPB = 256 / LPR;
  int64_t row = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  // inactive groups clamp to the last row (they recompute it) so the group shuffles stay full-warp converged
  if (row >= rows) row = rows - 1;
  group_ln_row<T, LPR, ITER>(x + row * C, y + row * C, w, b, eps, C, threadIdx.x % LPR);
}

// Two rows per lane group (same weights): both rows' loads are in flight together and the index arithmetic is shared.
template <typename T, int LPR, int ITER>
__device__ __forceinline__ void group_ln_row2(const T* __restrict__ src0, const T* __restrict__ src1, T* __restrict__ dst0,
                                              T* __restrict__ dst1, const float* __restrict__ w,
                                              const float* __restrict__ b, float eps, int C, int gl) {
  float v0[ITER][8], v1[ITER][8];
  float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
      load8<T>(src0 + c, v0[i]);
      load8<T>(src1 + c, v1[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { s0 += v0[i][e]; s1 += v1[i][e]; }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  const float mean0 = s0 / (float)C, mean1 = s1 / (float)C;
  float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        q0 += (v0[i][e] - mean0) * (v0[i][e] - mean0);
        q1 += (v1[i][e] - mean1) * (v1[i][e] - mean1);
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  const float r0 = rsqrtf(q0 / (float)C + eps), r1 = rsqrtf(q1 / (float)C + eps);
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
      float g[8], be[8];
      load8<float>(w + c, g);
      load8<float>(b + c, be);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v0[i][e] = (v0[i][e] - mean0) * r0 * g[e] + be[e];
        v1[i][e] = (v1[i][e] - mean1) * r1 * g[e] + be[e];
      }
      store8<T>(dst0 + c, v0[i]);
      store8<T>(dst1 + c, v1[i]);
    }
  }
}

// The same for the 16-bit types with packed fp32 arithmetic (fma.rn.f32x2): sum and sum of squares in one pass (one FFMA2
// each per channel pair; fp32 sums of <= 2048 16-bit values), (v * r - mean * r) * g + b as two FFMA2 per pair -- about
// 3.5 instructions per element instead of 8; the kernel is issue-bound (ncu: 80 % issue-active at 0.52 of HBM).
template <typename T, int LPR, int ITER>
__device__ __forceinline__ void group_ln_row2_packed(const T* __restrict__ src0, const T* __restrict__ src1, T* __restrict__ dst0,
                                                     T* __restrict__ dst1, const float* __restrict__ w,
                                                     const float* __restrict__ b, float eps, int C, int gl) {
  static_assert(sizeof(T) == 2, "16-bit types only");
  uint4 raw0[ITER], raw1[ITER];
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    raw0[i] = raw1[i] = make_uint4(0u, 0u, 0u, 0u);
    if (c < C) {
      raw0[i] = *reinterpret_cast<const uint4*>(src0 + c);
      raw1[i] = *reinterpret_cast<const uint4*>(src1 + c);
    }
  }
  const float2 one = make_float2(1.0f, 1.0f);
  float2 s0 = make_float2(0.f, 0.f), s1 = s0, q0 = s0, q1 = s0;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {            // zero-filled vectors past C add nothing
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&raw0[i]);
    const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&raw1[i]);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 u = unpack2<T>(a[e]), v = unpack2<T>(c1[e]);
      s0 = fma2(u, one, s0); q0 = fma2(u, u, q0);
      s1 = fma2(v, one, s1); q1 = fma2(v, v, q1);
    }
  }
  float t0 = s0.x + s0.y, t1 = s1.x + s1.y, u0 = q0.x + q0.y, u1 = q1.x + q1.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    t0 += __shfl_xor_sync(0xffffffffu, t0, o);
    t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    u0 += __shfl_xor_sync(0xffffffffu, u0, o);
    u1 += __shfl_xor_sync(0xffffffffu, u1, o);
  }
  const float inv = 1.0f / (float)C;
  const float mean0 = t0 * inv, mean1 = t1 * inv;
  const float r0 = rsqrtf(fmaxf(fmaf(-mean0, mean0, u0 * inv), 0.0f) + eps), r1 = rsqrtf(fmaxf(fmaf(-mean1, mean1, u1 * inv), 0.0f) + eps);
  const float2 r02 = make_float2(r0, r0), m02 = make_float2(-mean0 * r0, -mean0 * r0);
  const float2 r12 = make_float2(r1, r1), m12 = make_float2(-mean1 * r1, -mean1 * r1);
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = (gl + i * LPR) * 8;
    if (c < C) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(w + c)), g1 = __ldg(reinterpret_cast<const float4*>(w + c + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c)), b1 = __ldg(reinterpret_cast<const float4*>(b + c + 4));
      const float2 g[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
      const float2 be[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
      const uint32_t* a = reinterpret_cast<const uint32_t*>(&raw0[i]);
      const uint32_t* c1 = reinterpret_cast<const uint32_t*>(&raw1[i]);
      uint32_t o0[4], o1[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 y0 = fma2(fma2(unpack2<T>(a[e]), r02, m02), g[e], be[e]);
        const float2 y1 = fma2(fma2(unpack2<T>(c1[e]), r12, m12), g[e], be[e]);
        o0[e] = pack2<T>(y0.x, y0.y);
        o1[e] = pack2<T>(y1.x, y1.y);
      }
      *reinterpret_cast<uint4*>(dst0 + c) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
      *reinterpret_cast<uint4*>(dst1 + c) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    }
  }
}

// LayerNorm2d over C of every pixel, written straight into the im2col matrix of the
// following 2x2 stride-2 conv: pixel (2ho+kh, 2wo+kw) -> row (b,ho,wo), columns
// [(kh*2+kw)*C, +C).  Pixels of an odd last row/column are dropped (floor semantics).
template <typename T, int LPR, int ITER>
__global__ void __launch_bounds__(256)
ln_patchify2_kernel(const T* __restrict__ x, T* __restrict__ a, const float* __restrict__ w,
                    const float* __restrict__ b, float eps, int B, int H, int W, int C) {
  // one lane group per horizontal pixel pair (2wo, 2wo+1) of an input row: 2C contiguous elements in, 2C out
  constexpr int RPB = 256 / LPR;
  const int Ho = H / 2, Wo = W / 2;
  int64_t idx = (int64_t)blockIdx.x * RPB + threadIdx.x / LPR;
  const int64_t total = (int64_t)B * Ho * 2 * Wo;
  if (idx >= total) idx = total - 1;
  const int wo = (int)(idx % Wo);
  const int64_t t = idx / Wo;
  const int hi = (int)(t % (2 * Ho));
  const int64_t bi = t / (2 * Ho);
  const T* src = x + ((bi * H + hi) * W + 2 * wo) * C;
  const int64_t row = (bi * Ho + (hi >> 1)) * Wo + wo;
  T* dst = a + row * (4 * (int64_t)C) + (hi & 1) * 2 * C;
  if constexpr (sizeof(T) == 2) group_ln_row2_packed<T, LPR, ITER>(src, src + C, dst, dst + C, w, b, eps, C, threadIdx.x % LPR);
  else group_ln_row2<T, LPR, ITER>(src, src + C, dst, dst + C, w, b, eps, C, threadIdx.x % LPR);
}

// pick (lanes per row, vectors per lane) for a channel count; C % 8 == 0, C <= 2048
template <typename F>
int ln_dispatch(int C, F&& f) {
  const int vecs = C / 8;
  if (vecs <= 16) return f(std::integral_constant<int, 16>{}, std::integral_constant<int, 1>{});
  if (vecs <= 32) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 1>{});
  if (vecs <= 64) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 2>{});
  if (vecs <= 96) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 3>{});
  if (vecs <= 128) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 4>{});
  if (vecs <= 192) return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 6>{});
  return f(std::integral_constant<int, 32>{}, std::integral_constant<int, 8>{});
}

// Stem im2col (4x4 stride 4, 3 channels): row (b,ho,wo), column (kh*4+kw)*3 + c.
// One thread per output row: 12 independent 16-byte loads in flight (a warp reads 512 contiguous bytes per
// (channel, kh) image row) and one contiguous 48-element row out (six 16-byte stores for 16-bit T).
template <typename T>
__global__ void __launch_bounds__(256)
stem_patchify_nchw_kernel(const float* __restrict__ x, T* __restrict__ a, int B, int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo);
  const int64_t t = idx / Wo;
  const int ho = (int)(t % Ho);
  const int64_t b = t / Ho;
  float4 px[4][3];
#pragma unroll
  for (int kh = 0; kh < 4; ++kh)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      px[kh][c] = __ldg(reinterpret_cast<const float4*>(x + ((b * 3 + c) * H + (ho * 4 + kh)) * (int64_t)W + wo * 4));
  float v[48];
#pragma unroll
  for (int kh = 0; kh < 4; ++kh) {
    float* o = v + kh * 12;
    o[0] = px[kh][0].x; o[1] = px[kh][1].x; o[2] = px[kh][2].x;
    o[3] = px[kh][0].y; o[4] = px[kh][1].y; o[5] = px[kh][2].y;
    o[6] = px[kh][0].z; o[7] = px[kh][1].z; o[8] = px[kh][2].z;
    o[9] = px[kh][0].w; o[10] = px[kh][1].w; o[11] = px[kh][2].w;
  }
  T* dst = a + idx * 48;
#pragma unroll
  for (int i = 0; i < 6; ++i) store8<T>(dst + i * 8, v + i * 8);
}

template <typename T>
__global__ void __launch_bounds__(256)
stem_patchify_nhwc_kernel(const T* __restrict__ x, T* __restrict__ a, int B, int H, int W) {
  const int Ho = H / 4, Wo = W / 4;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo);
  const int64_t t = idx / Wo;
  const int ho = (int)(t % Ho);
  const int64_t b = t / Ho;
  T* dst = a + idx * 48;
  if constexpr (sizeof(T) == 2) {
    // a 4-pixel x 3-channel run is 24 contiguous bytes (8-byte aligned): 3 x 8-byte loads per image row
    uint2 r[4][3];
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const uint2* src = reinterpret_cast<const uint2*>(x + ((b * H + ho * 4 + kh) * (int64_t)W + wo * 4) * 3);
#pragma unroll
      for (int i = 0; i < 3; ++i) r[kh][i] = __ldg(src + i);
    }
    uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
    for (int kh = 0; kh < 4; ++kh)
#pragma unroll
      for (int i = 0; i < 3; ++i) d2[kh * 3 + i] = r[kh][i];
  } else {
#pragma unroll
    for (int kh = 0; kh < 4; ++kh) {
      const T* src = x + ((b * H + ho * 4 + kh) * (int64_t)W + wo * 4) * 3;
#pragma unroll
      for (int i = 0; i < 12; ++i) dst[kh * 12 + i] = src[i];
    }
  }
}

// Global average pool over HW then LayerNorm over C: one CTA per image, thread per channel (strided).
template <typename T>
__global__ void __launch_bounds__(256)
pool_ln_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b,
               float eps, int HW, int C) {
  __shared__ float red[8];
  __shared__ float stat;
  const T* xb = x + (int64_t)blockIdx.x * HW * C;
  constexpr int MAXV = 8;  // C <= 2048
  float v[MAXV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    float acc = 0.0f;
    if (c < C) {
      for (int p = 0; p < HW; ++p) acc += to_f<T>(xb[(int64_t)p * C + c]);
      acc /= (float)HW;
    }
    v[i] = acc;
    s += acc;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float tsum = 0; for (int i = 0; i < 8; ++i) tsum += red[i]; stat = tsum / (float)C; }
  __syncthreads();
  const float mean = stat;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) { const int c = threadIdx.x + i * 256; if (c < C) q += (v[i] - mean) * (v[i] - mean); }
  q = warp_sum(q);
  __syncthreads();
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (threadIdx.x == 0) { float tsum = 0; for (int i = 0; i < 8; ++i) tsum += red[i]; stat = rsqrtf(tsum / (float)C + eps); }
  __syncthreads();
  const float rstd = stat;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < C) y[(int64_t)blockIdx.x * C + c] = from_f<T>((v[i] - mean) * rstd * w[c] + b[c]);
  }
}

template <typename F>
int dispatch(int dtype, F&& f) {
  switch (dtype) {
    case GCV_F32: return f(float{});
    case GCV_BF16: return f(__nv_bfloat16{});
    case GCV_F16: return f(__half{});
    default: set_error("bad dtype %d", dtype); return GCV_ERR_BAD_ARG;
  }
}

}  // namespace

bool dwconv7_mma_supported(int dtype, int C);
int dwconv7_mma(int dtype, const void* x, void* y, float* stats, const float* taps, const float* bias, int B, int H,
                int W, int C, cudaStream_t stream);
int layernorm_rows(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int64_t rows, int C,
                   cudaStream_t stream);

// GCV_DWCONV=ffma keeps the 16-bit modes on the CUDA-core kernels (A/B timing); fp32 always runs there
static bool dw_use_mma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GCV_DWCONV");
    v = (e && e[0] == 'f') ? 0 : 1;
  }
  return v == 1;
}

int dwconv7_ln(int dtype, const void* x, void* y, const float* taps, const float* bias, const float* ln_w,
               const float* ln_b, float eps, int B, int H, int W, int C, cudaStream_t stream) {
  GCV_REQUIRE(C % 32 == 0 && C >= 32 && B > 0 && H > 0 && W > 0, "dwconv7_ln: C must be a multiple of 32 (C=%d)", C);
  if (dw_use_mma() && dwconv7_mma_supported(dtype, C) && C % 8 == 0 && C <= 2048) {
    // 16-bit modes: tensor-core depthwise conv (dwconv_mma.cu) into y, then the row LayerNorm in place
    const int rc = dwconv7_mma(dtype, x, y, nullptr, taps, bias, B, H, W, C, stream);
    if (rc != GCV_OK) return rc;
    return layernorm_rows(dtype, y, y, ln_w, ln_b, eps, (int64_t)B * H * W, C, stream);
  }
  const int groups = (H + 6) / 7;
  const bool col_shape = (C == 96 || C == 192 || C == 384 || C == 768) && groups * (C / 2) <= 384;
  const size_t es = dtype == GCV_F32 ? 4 : 2;
  const int threads = groups * (C / 2);
  const int cbox = C <= 192 ? C : 192;
  const size_t fixed = col_shape ? ((256 + (size_t)(49 * C + (threads / 16) * 49 + 2 * groups * 49 +
                                                     (C <= 384 ? threads * 49 : 0)) * 4 + 127) & ~(size_t)127) : 0;
  const size_t step_bytes = col_shape ? (size_t)groups * (C / cbox) * ((13 * cbox * es + 127) & ~(size_t)127) : 1;
  // the TMA-fed column kernel needs the taps (+ statistics buffer) plus at least two pipeline steps in shared memory
  if (col_shape && fixed + 2 * step_bytes <= 224 * 1024) {
    const int strips_w = (W + 6) / 7;
    const int64_t n_strips = (int64_t)B * strips_w;
    GCV_REQUIRE(n_strips < 2147483647LL / 13, "dwconv7_ln: too many strips");
    GCV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "dwconv7_ln: x must be 16-byte aligned");
    int depth = (int)((224 * 1024 - fixed) / step_bytes);
    if (depth > 8) depth = 8;
    const size_t smem = fixed + depth * step_bytes;
    const int sms = device_sms();
    const int grid = (int)(n_strips < sms ? n_strips : sms);
    // 4-D view [B][H][W][C] of the NHWC activation; box = (channel box, 13 pixels, 1 row, 1 image)
    CUtensorMap tm;
    {
      EncodeTiledFn enc = dw_get_encode();
      if (!enc) {
        set_error("cuTensorMapEncodeTiled not resolvable (no CUDA driver?)");
        return GCV_ERR_NO_DRIVER;
      }
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
      cuuint32_t box[4] = {(cuuint32_t)cbox, 13, 1, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      const CUtensorMapDataType tdt = dtype == GCV_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                                          : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
      CUresult r = enc(&tm, tdt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("dwconv7_ln: cuTensorMapEncodeTiled failed: CUresult %d (B=%d H=%d W=%d C=%d)", (int)r, B, H, W, C);
        return GCV_ERR_CUDA;
      }
    }
    return dispatch(dtype, [&](auto tag) -> int {
      using T = decltype(tag);
      auto launch = [&](auto cc) -> int {
        constexpr int CC = decltype(cc)::value;
        static unsigned long long attr_devs = 0;
        if (first_on_device(attr_devs)) {
          cudaFuncSetAttribute(dwconv7_ln_col_kernel<T, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        }
        dwconv7_ln_col_kernel<T, CC><<<grid, threads, smem, stream>>>(
            tm, reinterpret_cast<T*>(y), taps, bias, ln_w, ln_b, eps, H, W, strips_w, (int)n_strips, depth);
        return check_launch("dwconv7_ln");
      };
      switch (C) {
        case 96: return launch(std::integral_constant<int, 96>{});
        case 192: return launch(std::integral_constant<int, 192>{});
        case 384: return launch(std::integral_constant<int, 384>{});
        default: return launch(std::integral_constant<int, 768>{});
      }
    });
  }
  const int tiles_w = (W + DW_TW - 1) / DW_TW, tiles_h = (H + DW_TH - 1) / DW_TH;
  const int64_t grid = (int64_t)B * tiles_w * tiles_h;
  GCV_REQUIRE(grid < 2147483647LL, "dwconv7_ln: grid too large");
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const size_t halo_bytes = DW_HH * DW_HW * DW_CK * sizeof(T);
    size_t smem = halo_bytes + (size_t)DW_TH * DW_TW * C * sizeof(float);
    const int fuse_ln = smem <= 200 * 1024 ? 1 : 0;
    if (!fuse_ln) smem = halo_bytes;
    static unsigned long long attr_devs = 0;
    if (first_on_device(attr_devs)) {
      cudaFuncSetAttribute(dwconv7_ln_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    dwconv7_ln_kernel<T><<<(unsigned)grid, 128, smem, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y),
                                                               taps, bias, ln_w, ln_b, eps, H, W, C, tiles_w, tiles_h, fuse_ln);
    const int rc = check_launch("dwconv7_ln");
    if (rc != GCV_OK || fuse_ln) return rc;
    return layernorm_rows(dtype, y, y, ln_w, ln_b, eps, (int64_t)B * H * W, C, stream);
  });
}

int ln_patchify2(int dtype, const void* x, void* a, const float* w, const float* b, float eps, int B, int H, int W, int C,
                 cudaStream_t stream) {
  GCV_REQUIRE(C % 8 == 0 && C <= 2048 && H >= 2 && W >= 2, "ln_patchify2: unsupported C=%d H=%d W=%d", C, H, W);
  const int64_t total = (int64_t)B * (H / 2) * 2 * (W / 2);          // horizontal pixel pairs
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    return ln_dispatch(C, [&](auto lpr, auto iter) -> int {
      constexpr int LPR = decltype(lpr)::value, ITER = decltype(iter)::value, RPB = 256 / LPR;
      ln_patchify2_kernel<T, LPR, ITER><<<(unsigned)((total + RPB - 1) / RPB), 256, 0, stream>>>(
          reinterpret_cast<const T*>(x), reinterpret_cast<T*>(a), w, b, eps, B, H, W, C);
      return check_launch("ln_patchify2");
    });
  });
}

int stem_patchify(int dtype, bool nchw, const void* x, void* a, int B, int H, int W, cudaStream_t stream) {
  GCV_REQUIRE(H % 4 == 0 && W % 4 == 0, "stem_patchify: H, W must be multiples of 4");
  const int64_t total = (int64_t)B * (H / 4) * (W / 4);
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (nchw)
      stem_patchify_nchw_kernel<T><<<grid, 256, 0, stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<T*>(a), B, H, W);
    else
      stem_patchify_nhwc_kernel<T><<<grid, 256, 0, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(a), B, H, W);
    return check_launch("stem_patchify");
  });
}

int layernorm_rows(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int64_t rows, int C,
                   cudaStream_t stream) {
  GCV_REQUIRE(C % 8 == 0 && C <= 3072 && rows > 0, "layernorm_rows: unsupported C=%d", C);
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if (C > 2048) {                         // swin_large patch merging: LayerNorm(4 * 768)
      constexpr int LPR = 32, ITER = 12, RPB = 256 / LPR;
      layernorm_rows_kernel<T, LPR, ITER><<<(unsigned)((rows + RPB - 1) / RPB), 256, 0, stream>>>(
          reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), w, b, eps, rows, C);
      return check_launch("layernorm_rows");
    }
    return ln_dispatch(C, [&](auto lpr, auto iter) -> int {
      constexpr int LPR = decltype(lpr)::value, ITER = decltype(iter)::value, RPB = 256 / LPR;
      layernorm_rows_kernel<T, LPR, ITER><<<(unsigned)((rows + RPB - 1) / RPB), 256, 0, stream>>>(
          reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), w, b, eps, rows, C);
      return check_launch("layernorm_rows");
    });
  });
}

int pool_ln(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int B, int HW, int C,
            cudaStream_t stream) {
  GCV_REQUIRE(C <= 2048 && B > 0 && HW > 0, "pool_ln: unsupported C=%d", C);
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    pool_ln_kernel<T><<<B, 256, 0, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), w, b, eps, HW, C);
    return check_launch("pool_ln");
  });
}


// ---------------------------------------------------------------------------------
// Reduce the LayerNorm partial sums written by dwconv7_stats ([M][chunks] x (sum, sumsq)) to one (rstd, -mean*rstd)
// pair per row, once, instead of in every epilogue thread of the consuming GEMM (4 warps per row x every n-tile).
// ---------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
ln_finalize_kernel(const float2* __restrict__ stats, float2* __restrict__ out, int64_t M, int chunks, int K, float eps) {
  const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  if ((chunks & 1) == 0) {                       // 16-byte loads: two chunks at a time (rows are 8 * chunks bytes)
    const float4* p = reinterpret_cast<const float4*>(stats + m * chunks);
    float s = 0.0f, q = 0.0f;
    for (int i = 0; i < chunks / 2; ++i) {
      const float4 v = __ldg(p + i);
      s += v.x; q += v.y;
      s += v.z; q += v.w;
    }
    const float inv = 1.0f / (float)K, mean = s * inv;
    const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, q * inv), 0.0f) + eps);
    out[m] = make_float2(rstd, -mean * rstd);
    return;
  }
  out[m] = ln_row_scale(reinterpret_cast<const float*>(stats), m, chunks, K, eps);
}
}  // namespace

int ln_finalize(const float* stats, float* out, int64_t M, int chunks, int K, float eps, cudaStream_t stream) {
  GCV_REQUIRE(M > 0 && chunks > 0 && K > 0, "ln_finalize: bad shape");
  ln_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const float2*>(stats),
                                                                      reinterpret_cast<float2*>(out), M, chunks, K, eps);
  return check_launch("ln_finalize");
}

}  // namespace gcv
