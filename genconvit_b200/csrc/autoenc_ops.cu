// Autoencoder / VAE side kernels and the scoring kernel.
//   reference model/genconvit_ed.py:13-33 (Encoder), model/genconvit_vae.py:15-31 (Encoder.features),
//   model/genconvit_vae.py:105,116 (Resize of the returned x_hat), model/pred_func.py:111-131 (scoring).
#include <stdlib.h>

#include "common.cuh"

namespace gcv {

namespace {

template <typename F>
int dispatch(int dtype, F&& f) {
  switch (dtype) {
    case GCV_F32: return f(float{});
    case GCV_BF16: return f(__nv_bfloat16{});
    case GCV_F16: return f(__half{});
    default: set_error("bad dtype %d", dtype); return GCV_ERR_BAD_ARG;
  }
}

// ---------------------------------------------------------------------------------
// First 3x3 conv, 3 -> 16 channels, pad 1, straight from NCHW fp32 frames.
// One thread per output pixel (after pooling when POOL): K = 27 is far too thin for
// the tensor cores, so this is a direct fp32 conv with the 432 weights broadcast
// from shared memory.  Output NHWC, 16 channels = two 16-byte stores (16-bit T).
// ---------------------------------------------------------------------------------
template <typename T, int STRIDE, bool POOL>
__global__ void __launch_bounds__(128, POOL ? 4 : 8)
conv3x3_first_kernel(const float* __restrict__ x, T* __restrict__ y, const float* __restrict__ w,
                     const float* __restrict__ bias, int act, int B, int H, int W, int Ho, int Wo) {
  __shared__ __align__(16) float ws[27][16];   // [ci*9 + kh*3 + kw][co]
  __shared__ float bs[16];
  for (int i = threadIdx.x; i < 432; i += blockDim.x) {
    const int co = i / 27, r = i - co * 27;      // OIHW: w[co][ci][kh][kw]
    ws[r][co] = w[i];
  }
  if (threadIdx.x < 16) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo;
  if (idx >= total) return;
  const int wo = (int)(idx % Wo);
  const int64_t t = idx / Wo;
  const int ho = (int)(t % Ho);
  const int64_t b = t / Ho;
  constexpr int NP = POOL ? 2 : 1;          // conv outputs per dim feeding this thread's output
  constexpr int IN = (NP - 1) * STRIDE + 3; // input patch edge
  const int cy0 = ho * NP * STRIDE - 1, cx0 = wo * NP * STRIDE - 1;
  // the three channel patches are loaded up front (48 independent loads in flight per thread), their addresses and
  // bounds predicates computed once
  float patch[3][IN][IN];
  {
    const float* xb = x + (b * 3) * (int64_t)H * W;
    const int64_t plane = (int64_t)H * W;
#pragma unroll
    for (int i = 0; i < IN; ++i)
#pragma unroll
      for (int j = 0; j < IN; ++j) {
        const int yy = cy0 + i, xx = cx0 + j;
        const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const float* px = xb + (int64_t)yy * W + xx;
#pragma unroll
        for (int c = 0; c < 3; ++c) patch[c][i][j] = ok ? __ldg(px + c * plane) : 0.0f;
      }
  }
  // all NP*NP conv positions advance together so each tap's 16 weights are read once
  // output channels in pairs: one FFMA2 (fma.rn.f32x2) per tap and channel pair -- the kernel is FMA-issue bound
  float2 acc[NP * NP][8];
#pragma unroll
  for (int q = 0; q < NP * NP; ++q)
#pragma unroll
    for (int co = 0; co < 8; ++co) acc[q][co] = make_float2(bs[2 * co], bs[2 * co + 1]);
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float2 wr[8];
#pragma unroll
        for (int co = 0; co < 8; co += 2) {
          const float4 w4 = *reinterpret_cast<const float4*>(&ws[c * 9 + kh * 3 + kw][2 * co]);
          wr[co] = make_float2(w4.x, w4.y); wr[co + 1] = make_float2(w4.z, w4.w);
        }
#pragma unroll
        for (int py = 0; py < NP; ++py)
#pragma unroll
          for (int px = 0; px < NP; ++px) {
            const float v = patch[c][py * STRIDE + kh][px * STRIDE + kw];
            const float2 vv = make_float2(v, v);
#pragma unroll
            for (int co = 0; co < 8; ++co) acc[py * NP + px][co] = fma2(vv, wr[co], acc[py * NP + px][co]);
          }
      }
  float out[16];
#pragma unroll
  for (int co = 0; co < 8; ++co) {
    float m0 = apply_act(acc[0][co].x, act), m1 = apply_act(acc[0][co].y, act);
#pragma unroll
    for (int q = 1; q < NP * NP; ++q) {
      m0 = fmaxf(m0, apply_act(acc[q][co].x, act));
      m1 = fmaxf(m1, apply_act(acc[q][co].y, act));
    }
    out[2 * co] = m0; out[2 * co + 1] = m1;
  }
  T* dst = y + idx * 16;
  store8<T>(dst, out);
  store8<T>(dst + 8, out + 8);
}

// im2col for the 3x3 convs after the first: x [B,H,W,C] -> a [B*Ho*Wo, 9C], pad 1.
// One thread per (output pixel, tap, 8-channel vector): a pure 16-byte gather/scatter.
template <typename T>
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const T* __restrict__ x, T* __restrict__ a, int B, int H, int W, int C, int stride, int Ho, int Wo) {
  constexpr int VE = 8;
  const int vpc = C / VE;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * 9 * vpc;
  if (idx >= total) return;
  const int v = (int)(idx % vpc);
  int64_t t = idx / vpc;
  const int tap = (int)(t % 9); t /= 9;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int64_t b = t / Ho;
  const int yy = ho * stride - 1 + tap / 3, xx = wo * stride - 1 + tap % 3;
  float vals[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  T* dst = a + ((((b * Ho + ho) * Wo + wo) * 9 + tap) * (int64_t)C) + v * VE;
  if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
    const T* src = x + ((b * H + yy) * (int64_t)W + xx) * C + v * VE;
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
      *reinterpret_cast<float4*>(dst + 4) = *reinterpret_cast<const float4*>(src + 4);
    } else {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
    }
  } else {
    store8<T>(dst, vals);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
maxpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, vpc = C / 8;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * Ho * Wo * vpc;
  if (idx >= total) return;
  const int v = (int)(idx % vpc);
  int64_t t = idx / vpc;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int64_t b = t / Ho;
  float m[8], q[8];
  const T* p = x + ((b * H + 2 * ho) * (int64_t)W + 2 * wo) * C + v * 8;
  load8<T>(p, m);
  load8<T>(p + C, q);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], q[i]);
  load8<T>(p + (int64_t)W * C, q);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], q[i]);
  load8<T>(p + (int64_t)W * C + C, q);
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], q[i]);
  store8<T>(y + ((b * Ho + ho) * (int64_t)Wo + wo) * C + v * 8, m);
}

// Bilinear x2 upscale (align_corners=False; antialias is a no-op when upscaling) of an
// NHWC image to NCHW fp32.  src coordinate = (dst + 0.5)/2 - 0.5, clamped at 0.
template <typename T>
__global__ void __launch_bounds__(256)
resize2x_kernel(const T* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int Ho = 2 * H, Wo = 2 * W;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * C * Ho * Wo;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo);
  int64_t t = idx / Wo;
  const int oy = (int)(t % Ho); t /= Ho;
  const int c = (int)(t % C);
  const int64_t b = t / C;
  const float sy = fmaxf((oy + 0.5f) * 0.5f - 0.5f, 0.0f), sx = fmaxf((ox + 0.5f) * 0.5f - 0.5f, 0.0f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float fy = sy - y0, fx = sx - x0;
  const T* xb = x + b * (int64_t)H * W * C + c;
  const float v00 = to_f<T>(xb[((int64_t)y0 * W + x0) * C]), v01 = to_f<T>(xb[((int64_t)y0 * W + x1) * C]);
  const float v10 = to_f<T>(xb[((int64_t)y1 * W + x0) * C]), v11 = to_f<T>(xb[((int64_t)y1 * W + x1) * C]);
  y[idx] = (1.0f - fy) * ((1.0f - fx) * v00 + fx * v01) + fy * ((1.0f - fx) * v10 + fx * v11);
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * C * H * W;
  if (idx >= total) return;
  const int ox = (int)(idx % W);
  int64_t t = idx / W;
  const int oy = (int)(t % H); t /= H;
  const int c = (int)(t % C);
  const int64_t b = t / C;
  y[idx] = to_f<T>(x[((b * H + oy) * (int64_t)W + ox) * C + c]);
}

// Last decoder layer: ConvTranspose2d(16 -> 3, k2 s2) + activation, NHWC.  K = 16 and N = 12 are far
// below any tensor-core tile and the layer is pure streaming (32 B in, 24 B out per input pixel), so one
// thread handles one input pixel: 16 inputs, 4 x 3 dot products, and per output row two RGB pixels
// (6 x 16-bit = three 32-bit stores; consecutive threads write consecutive 12-byte runs).
template <typename T>
__global__ void __launch_bounds__(256)
convt2x2_16to3_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ w,
                      const float* __restrict__ bias, int act, int64_t total, int H, int W) {
  __shared__ float ws[12][16];     // [(i*2+j)*3 + co][ci]
  __shared__ float bs[3];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) ws[i / 16][i % 16] = w[i];
  if (threadIdx.x < 3) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wq = (int)(idx % W);
  const int64_t t = idx / W;
  const int h = (int)(t % H);
  const int64_t b = t / H;
  float in[16];
  if constexpr (sizeof(T) == 4) {
    load8<T>(x + idx * 16, in);
    load8<T>(x + idx * 16 + 8, in + 8);
  } else {
    load8<T>(x + idx * 16, in);
    load8<T>(x + idx * 16 + 8, in + 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float o[6];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int co = 0; co < 3; ++co) {
        float acc = bs[co];
        const float* wr = ws[(i * 2 + j) * 3 + co];
#pragma unroll
        for (int ci = 0; ci < 16; ++ci) acc = fmaf(in[ci], wr[ci], acc);
        o[j * 3 + co] = apply_act(acc, act);
      }
    T* dst = y + ((b * (2 * H) + 2 * h + i) * (int64_t)(2 * W) + 2 * wq) * 3;
    if constexpr (sizeof(T) == 4) {
#pragma unroll
      for (int e = 0; e < 6; ++e) dst[e] = o[e];
    } else {
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      d32[0] = pack2<T>(o[0], o[1]);
      d32[1] = pack2<T>(o[2], o[3]);
      d32[2] = pack2<T>(o[4], o[5]);
    }
  }
}

// pred_vid / max_prediction_value, batched: one warp per video.
__global__ void __launch_bounds__(32)
score_videos_kernel(const float* __restrict__ logits, int n_nets, int n_frames, int fpv, float* __restrict__ mean_out,
                    int32_t* __restrict__ cls_out, float* __restrict__ val_out) {
  const int v = blockIdx.x, lane = threadIdx.x;
  float s0 = 0.0f, s1 = 0.0f;
  const int rows = n_nets * fpv;
  for (int i = lane; i < rows; i += 32) {
    const int net = i / fpv, f = i - net * fpv;
    const float2 l = *reinterpret_cast<const float2*>(logits + 2 * ((int64_t)net * n_frames + (int64_t)v * fpv + f));
    s0 += 1.0f / (1.0f + expf(-l.x));
    s1 += 1.0f / (1.0f + expf(-l.y));
  }
  s0 = warp_sum(s0) / (float)rows;
  s1 = warp_sum(s1) / (float)rows;
  if (lane == 0) {
    mean_out[2 * v] = s0;
    mean_out[2 * v + 1] = s1;
    cls_out[v] = s1 > s0 ? 1 : 0;                       // torch.argmax: first maximal index on ties
    val_out[v] = s0 > s1 ? s0 : fabsf(1.0f - s1);       // ties take the else branch (pred_func.py:128-130)
  }
}


// one warp per video; rows of the ED and of the VAE network are read from their own buffers (either may be null)
__global__ void __launch_bounds__(32)
score_videos_pair_kernel(const float* __restrict__ la, const float* __restrict__ lb, int n_videos, int fpv,
                         float* __restrict__ out) {
  const int v = blockIdx.x, lane = threadIdx.x;
  float s0 = 0.0f, s1 = 0.0f;
  int rows = 0;
#pragma unroll
  for (int net = 0; net < 2; ++net) {
    const float* lg = net == 0 ? la : lb;
    if (!lg) continue;
    rows += fpv;
    for (int f = lane; f < fpv; f += 32) {
      const float2 l = *reinterpret_cast<const float2*>(lg + 2 * ((int64_t)v * fpv + f));
      s0 += 1.0f / (1.0f + expf(-l.x));
      s1 += 1.0f / (1.0f + expf(-l.y));
    }
  }
  s0 = warp_sum(s0) / (float)rows;
  s1 = warp_sum(s1) / (float)rows;
  if (lane == 0) {
    out[v] = s1 > s0 ? 1.0f : 0.0f;                             // torch.argmax: first maximal index on ties
    out[n_videos + v] = s0 > s1 ? s0 : fabsf(1.0f - s1);        // ties take the else branch (pred_func.py:128-130)
  }
}

// ---- first 3x3 conv on the tensor cores (16-bit modes) -----------------------------------------------------------
// The FFMA kernel above needs 432 FMAs per conv pixel and is issue-bound at ~1 TB/s.  Here the frame tile is staged
// once in shared memory as 16-bit RGB0 pixels (8 B each, zero outside the image = the conv padding) and the conv runs
// as mma.sync m16n8k16 with K = (tap, channel) = 9 x 4 = 36 padded to 48 (three slices), N = 16 (two n-tiles): lane
// (g, t) needs for K index 16s + 8h + 2t the channel pair 2(t%2) of tap 4s + 2h + t/2 -- one 32-bit shared-memory load.
// A warp owns two conv rows of the tile (the two rows of a pooled row) and walks 7 segments of 16 pixels.
constexpr int C1_ROWS = 16, C1_SEGS = 7, C1_COLS = 16 * C1_SEGS;      // conv-output tile: 16 rows x 112 columns

template <typename T>
__device__ __forceinline__ void c1_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct C1Norm { float mean[3], sd[3]; };      // U8: Normalize(mean, std) applied to the raw crops

// U8 = false: x = fp32 [B,3,H,W] pre-processed frames.  U8 = true: x = uint8 [B,H,W,3] raw crops; the normalisation
// (model/pred_func.py:95-108 + dataset/loader.py:63-77) is a 3 x 256 table of T(fdiv(fdiv(u, 255) - mean, std)), the very
// operand the fp32 path builds from the pre-processed frame, so both paths give bit-identical outputs.
template <typename T, int STRIDE, bool POOL, bool U8>
__global__ void __launch_bounds__(256)
conv3x3_first_mma_kernel(const void* __restrict__ xin, T* __restrict__ y, const float* __restrict__ w,
                         const float* __restrict__ bias, int act, int B, int H, int W, int Hc, int Wc, int tiles_x,
                         int tiles_y, int n_tiles, const C1Norm nrm) {
  const float* x = reinterpret_cast<const float*>(xin);
  __shared__ uint16_t s_lut[U8 ? 3 * 256 : 2];
  if constexpr (U8) {
    for (int i = threadIdx.x; i < 3 * 256; i += 256) {
      const int c = i >> 8;
      const T v = from_f<T>(__fdiv_rn(__fdiv_rn((float)(i & 255), 255.0f) - nrm.mean[c], nrm.sd[c]));
      s_lut[i] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
  constexpr int IN_H = C1_ROWS * STRIDE + 2, IN_W = C1_COLS * STRIDE + 2;
  extern __shared__ __align__(16) uint8_t c1sm[];
  uint2* tile = reinterpret_cast<uint2*>(c1sm);                       // [IN_H][IN_W] pixels of 4 x 16 bit (R, G, B, 0)
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(c1sm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B fragments and this lane's tap offsets: K index k = 16s + 8h + 2t (+1): tap = k / 4, channel = k % 4
  uint32_t bf[3][2][2], aoff[3][2];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 16 * s + 8 * h + 2 * t, tap = k >> 2, c = k & 3;      // c = 0 or 2
      const bool ok = tap < 9;
      const int kh = ok ? tap / 3 : 0, kw = ok ? tap % 3 : 0;
      aoff[s][h] = (uint32_t)((kh * IN_W + kw) * 8 + c * 2);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int co = nt * 8 + g;
        const float w0 = ok ? __ldg(w + co * 27 + c * 9 + tap) : 0.0f;                       // OIHW: [co][c][kh][kw]
        const float w1 = (ok && c + 1 < 3) ? __ldg(w + co * 27 + (c + 1) * 9 + tap) : 0.0f;   // channel 3 is the zero pad
        bf[s][nt][h] = pack2<T>(w0, w1);
      }
    }
  float bv[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    bv[nt][0] = __ldg(bias + nt * 8 + 2 * t);
    bv[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
  }
  const int Ho = POOL ? Hc / 2 : Hc, Wo = POOL ? Wc / 2 : Wc;
  for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int b = tl / (tiles_x * tiles_y), r = tl - b * (tiles_x * tiles_y);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int iy0 = ty * C1_ROWS * STRIDE - 1, ix0 = tx * C1_COLS * STRIDE - 1;
    const float* xb = x + (int64_t)b * 3 * H * W;
    __syncthreads();                                   // the previous tile's readers are done (first tile: the table is built)
    if constexpr (U8) {
      const uint8_t* ub = reinterpret_cast<const uint8_t*>(xin) + (int64_t)b * H * W * 3;
      for (int i0 = threadIdx.x; i0 < IN_H * IN_W; i0 += 4 * 256) {
        uint32_t c[4][3];
        bool oks[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          const int py = i / IN_W, px = i - py * IN_W;
          const int iy = iy0 + py, ix = ix0 + px;
          oks[u] = i < IN_H * IN_W && iy >= 0 && iy < H && ix >= 0 && ix < W;
          const uint8_t* p = ub + ((int64_t)iy * W + ix) * 3;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) c[u][ch] = oks[u] ? (uint32_t)__ldg(p + ch) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 256;
          if (i < IN_H * IN_W)                        // zero padding is applied to the NORMALISED image
            tile[i] = oks[u] ? make_uint2((uint32_t)s_lut[c[u][0]] | ((uint32_t)s_lut[256 + c[u][1]] << 16), (uint32_t)s_lut[512 + c[u][2]])
                             : make_uint2(0u, 0u);
        }
      }
    } else
    // four pixels per thread and pass: all twelve loads are issued before the first one is consumed
    for (int i0 = threadIdx.x; i0 < IN_H * IN_W; i0 += 4 * 256) {
      float c[4][3];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        const int py = i / IN_W, px = i - py * IN_W;
        const int iy = iy0 + py, ix = ix0 + px;
        const bool ok = i < IN_H * IN_W && iy >= 0 && iy < H && ix >= 0 && ix < W;
        const float* p = xb + (int64_t)iy * W + ix;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) c[u][ch] = ok ? __ldg(p + ch * (int64_t)H * W) : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * 256;
        if (i < IN_H * IN_W) tile[i] = make_uint2(pack2<T>(c[u][0], c[u][1]), pack2<T>(c[u][2], 0.0f));
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int seg = 0; seg < C1_SEGS; ++seg) {
      float acc[2][2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          acc[rr][nt][0] = acc[rr][nt][2] = bv[nt][0];
          acc[rr][nt][1] = acc[rr][nt][3] = bv[nt][1];
        }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        // pixel (row 2*warp + rr, column seg*16 + g [+ 8]) of the conv tile -> top-left input pixel of its 3x3 window
        const uint32_t base = tile_s + (uint32_t)((((2 * warp + rr) * STRIDE) * IN_W + (seg * 16 + g) * STRIDE) * 8);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          uint32_t a[4];
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[0]) : "r"(base + aoff[s][0]));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[1]) : "r"(base + aoff[s][0] + 8 * STRIDE * 8));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[2]) : "r"(base + aoff[s][1]));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a[3]) : "r"(base + aoff[s][1] + 8 * STRIDE * 8));
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) c1_mma<T>(acc[rr][nt], a, bf[s][nt][0], bf[s][nt][1]);
        }
      }
      // acc[rr][nt][e]: conv row ty*16 + 2*warp + rr, column tx*112 + seg*16 + g (e < 2) / + 8 (e >= 2), channel nt*8 + 2t + (e & 1)
      if constexpr (POOL) {
        const int py = ty * (C1_ROWS / 2) + warp;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          float m[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float vmax = fmaxf(acc[0][nt][e], acc[1][nt][e]);
            m[e] = apply_act(fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, 4)), act);      // monotone act: after the max
          }
          if (((g & 1) == 0) == (nt == 0) && py < Ho) {          // even-g lanes store n-tile 0, odd-g lanes n-tile 1
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int px = tx * (C1_COLS / 2) + seg * 8 + (g >> 1) + 4 * hh;
              if (px < Wo)
                *reinterpret_cast<uint32_t*>(y + (((int64_t)b * Ho + py) * Wo + px) * 16 + nt * 8 + 2 * t) =
                    pack2<T>(m[2 * hh], m[2 * hh + 1]);
            }
          }
        }
      } else {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int oy = ty * C1_ROWS + 2 * warp + rr;
          if (oy < Ho) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int ox = tx * C1_COLS + seg * 16 + g + 8 * hh;
              if (ox < Wo) {
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                  *reinterpret_cast<uint32_t*>(y + (((int64_t)b * Ho + oy) * Wo + ox) * 16 + nt * 8 + 2 * t) =
                      pack2<T>(apply_act(acc[rr][nt][2 * hh], act), apply_act(acc[rr][nt][2 * hh + 1], act));
              }
            }
          }
        }
      }
    }
  }
}
}  // namespace

int conv3x3_first_src(int dtype, int u8, const void* x, void* y, const float* w, const float* b, int stride, int act,
                      int pool, int B, int H, int W, const float* mean3, const float* std3, cudaStream_t stream);

int conv3x3_first(int dtype, const float* x, void* y, const float* w, const float* b, int stride, int act, int pool,
                  int B, int H, int W, cudaStream_t stream) {
  return conv3x3_first_src(dtype, 0, x, y, w, b, stride, act, pool, B, H, W, nullptr, nullptr, stream);
}

// u8 != 0: x is uint8 [B,H,W,3] raw crops, normalised with (mean3, std3) (host pointers) on the fly; 16-bit dtypes only.
int conv3x3_first_src(int dtype, int u8, const void* xv, void* y, const float* w, const float* b, int stride, int act,
                      int pool, int B, int H, int W, const float* mean3, const float* std3, cudaStream_t stream) {
  const float* x = reinterpret_cast<const float*>(xv);
  GCV_REQUIRE((stride == 1 || stride == 2) && B > 0, "conv3x3_first: stride must be 1 or 2");
  GCV_REQUIRE(!u8 || (mean3 && std3 && (dtype == GCV_BF16 || dtype == GCV_F16)), "conv3x3_first: the uint8 source needs mean / std and a 16-bit dtype");
  C1Norm nrm{};
  if (u8)
    for (int c = 0; c < 3; ++c) { nrm.mean[c] = mean3[c]; nrm.sd[c] = std3[c]; }
  GCV_REQUIRE(!(stride == 2 && pool), "conv3x3_first: stride 2 with pooling is not a reference configuration");
  const int Hc = (H + 2 - 3) / stride + 1, Wc = (W + 2 - 3) / stride + 1;
  const int Ho = pool ? Hc / 2 : Hc, Wo = pool ? Wc / 2 : Wc;
  const int64_t total = (int64_t)B * Ho * Wo;
  const unsigned grid = (unsigned)((total + 127) / 128);
  static int mma_env = -1;                      // GCV_CONV1_MMA=0 keeps the fp32 FFMA kernel in the 16-bit modes (A/B timing)
  if (mma_env < 0) { const char* e = getenv("GCV_CONV1_MMA"); mma_env = e ? atoi(e) : 1; }
  GCV_REQUIRE(!u8 || !pool || (Hc % 2 == 0 && Wc % 2 == 0), "conv3x3_first: the uint8 source with pooling needs even conv output sizes");
  if ((mma_env || u8) && (dtype == GCV_BF16 || dtype == GCV_F16) && (!pool || (Hc % 2 == 0 && Wc % 2 == 0))) {
    const int sms = device_sms();
    const int tiles_x = (Wc + C1_COLS - 1) / C1_COLS, tiles_y = (Hc + C1_ROWS - 1) / C1_ROWS;
    const int64_t n_tiles64 = (int64_t)B * tiles_x * tiles_y;
    GCV_REQUIRE(n_tiles64 < 2147483647LL, "conv3x3_first: too many tiles");
    const int n_tiles = (int)n_tiles64;
    const size_t smem = (size_t)(C1_ROWS * stride + 2) * (C1_COLS * stride + 2) * 8;
    const int per_sm = stride == 1 ? 4 : 2;
    const unsigned g2 = (unsigned)(n_tiles < per_sm * sms ? n_tiles : per_sm * sms);
    return dispatch(dtype, [&](auto tag) -> int {
      using T = decltype(tag);
      if constexpr (std::is_same<T, float>::value) {
        return GCV_ERR_BAD_ARG;
      } else {
        auto go = [&](auto kernel) {
          cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          kernel<<<g2, 256, smem, stream>>>(xv, reinterpret_cast<T*>(y), w, b, act, B, H, W, Hc, Wc, tiles_x, tiles_y, n_tiles, nrm);
        };
        if (u8) {
          if (stride == 1 && pool) go(conv3x3_first_mma_kernel<T, 1, true, true>);
          else if (stride == 1) go(conv3x3_first_mma_kernel<T, 1, false, true>);
          else go(conv3x3_first_mma_kernel<T, 2, false, true>);
        } else {
          if (stride == 1 && pool) go(conv3x3_first_mma_kernel<T, 1, true, false>);
          else if (stride == 1) go(conv3x3_first_mma_kernel<T, 1, false, false>);
          else go(conv3x3_first_mma_kernel<T, 2, false, false>);
        }
        return check_launch("conv3x3_first");
      }
    });
  }
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if (stride == 1 && pool)
      conv3x3_first_kernel<T, 1, true><<<grid, 128, 0, stream>>>(x, reinterpret_cast<T*>(y), w, b, act, B, H, W, Ho, Wo);
    else if (stride == 1)
      conv3x3_first_kernel<T, 1, false><<<grid, 128, 0, stream>>>(x, reinterpret_cast<T*>(y), w, b, act, B, H, W, Ho, Wo);
    else
      conv3x3_first_kernel<T, 2, false><<<grid, 128, 0, stream>>>(x, reinterpret_cast<T*>(y), w, b, act, B, H, W, Ho, Wo);
    return check_launch("conv3x3_first");
  });
}

int im2col3x3(int dtype, const void* x, void* a, int B, int H, int W, int C, int stride, cudaStream_t stream) {
  GCV_REQUIRE(C % 8 == 0 && (stride == 1 || stride == 2), "im2col3x3: C must be a multiple of 8 (C=%d)", C);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const int64_t total = (int64_t)B * Ho * Wo * 9 * (C / 8);
  GCV_REQUIRE((total + 255) / 256 < 2147483647LL, "im2col3x3: too large");
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    im2col3x3_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const T*>(x), reinterpret_cast<T*>(a), B, H, W, C, stride, Ho, Wo);
    return check_launch("im2col3x3");
  });
}

int maxpool2(int dtype, const void* x, void* y, int B, int H, int W, int C, cudaStream_t stream) {
  GCV_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2, "maxpool2: C must be a multiple of 8");
  const int64_t total = (int64_t)B * (H / 2) * (W / 2) * (C / 8);
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    maxpool2_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const T*>(x),
                                                                             reinterpret_cast<T*>(y), B, H, W, C);
    return check_launch("maxpool2");
  });
}

int resize2x_to_nchw(int dtype, const void* x, float* y, int B, int H, int W, int C, cudaStream_t stream) {
  const int64_t total = (int64_t)B * C * 4 * H * W;
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    resize2x_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const T*>(x), y, B, H, W, C);
    return check_launch("resize2x_to_nchw");
  });
}

int nhwc_to_nchw_f32(int dtype, const void* x, float* y, int B, int H, int W, int C, cudaStream_t stream) {
  const int64_t total = (int64_t)B * C * H * W;
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    nhwc_to_nchw_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const T*>(x), y, B, H, W, C);
    return check_launch("nhwc_to_nchw_f32");
  });
}

int convt2x2_small(int dtype, const void* x, void* y, const float* w, const float* bias, int act, int B, int H, int W,
                   int CI, int CO, cudaStream_t stream) {
  GCV_REQUIRE(CI == 16 && CO == 3, "convt2x2_small: only the 16 -> 3 output layer is specialised (got %d -> %d)", CI, CO);
  const int64_t total = (int64_t)B * H * W;
  return dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    convt2x2_16to3_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), w, bias, act, total, H, W);
    return check_launch("convt2x2_small");
  });
}

int score_videos(const float* logits, int n_nets, int n_frames, int fpv, float* mean_out, int32_t* cls_out,
                 float* val_out, cudaStream_t stream) {
  GCV_REQUIRE(n_nets > 0 && fpv > 0 && n_frames >= fpv && n_frames % fpv == 0,
              "score_videos: n_frames (%d) must be a positive multiple of frames_per_video (%d)", n_frames, fpv);
  score_videos_kernel<<<n_frames / fpv, 32, 0, stream>>>(logits, n_nets, n_frames, fpv, mean_out, cls_out, val_out);
  return check_launch("score_videos");
}

int score_videos_pair(const float* la, const float* lb, int n_frames, int fpv, float* out, cudaStream_t stream) {
  GCV_REQUIRE((la || lb) && out, "score_videos_pair: null pointer");
  GCV_REQUIRE(fpv > 0 && n_frames >= fpv && n_frames % fpv == 0,
              "score_videos_pair: n_frames (%d) must be a positive multiple of frames_per_video (%d)", n_frames, fpv);
  score_videos_pair_kernel<<<n_frames / fpv, 32, 0, stream>>>(la, lb, n_frames / fpv, fpv, out);
  return check_launch("score_videos_pair");
}

}  // namespace gcv
