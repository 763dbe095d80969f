// Direct 3x3 convolution (pad 1) for the autoencoders' second layer, Cin = 16 -> Cout = 32, NHWC 16-bit, on the
// tensor cores (mma.sync m16n8k16, fp32 accumulate), with the activation and -- for Network A -- the 2x2 max-pool
// fused in:
//   stride 1 + ReLU + MaxPool2d(2)   reference model/genconvit_ed.py:18-20  (Conv2d(16,32,3,p1), ReLU, MaxPool2d)
//   stride 2 + LeakyReLU(0.01)       reference model/genconvit_vae.py:19-21 (Conv2d(16,32,3,s2,p1), BN folded, LeakyReLU)
// This layer has the largest pixel count of the encoders (112 x 112 x 256 frames, K = 9*16 = 144): as im2col + GEMM it
// wrote and re-read a 925 MB patch matrix per batch and the pool made another pass; here every input pixel is read
// once (plus the tile halo) and only the pooled output is written.
//
// CTA = 8 warps, one 16 x 16 tile of conv outputs of one image.  The input halo tile ((16 S + 2)^2 pixels x 32 B) is
// staged in shared memory with cp.async (zero fill = the conv padding), 16-byte pieces XOR-swizzled so that ldmatrix is
// conflict-free.  A tap (kh, kw) is exactly one K = 16 slice: the A fragment of an output row of 16 pixels is one
// ldmatrix.x4 of the tap-shifted pixels, the B fragments (all 9 taps x 4 n-tiles of the [32][144] weight matrix) stay in
// registers for the whole kernel.  Warp w owns output rows 2w, 2w+1 of the tile, so the vertical half of the 2x2 pool
// is register-local and the horizontal half is one shuffle.
#include "common.cuh"

namespace gcv {

namespace {

constexpr int CT = 16;          // conv-output tile edge
constexpr int CTHREADS = 256;

template <int S> struct CGeom {
  static constexpr int IN = CT * S + 2;                  // staged input pixels per edge (S = 2 needs 16*2 + 1, padded)
  static constexpr int BYTES = IN * IN * 32;
};

template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float act_apply(float v, int act) {
  return act == GCV_ACT_RELU ? fmaxf(v, 0.0f) : (act == GCV_ACT_LEAKY ? (v > 0.0f ? v : 0.01f * v) : v);
}

// smem byte offset of 16-byte piece `half` (channels 8*half ..) of staged pixel p (linear index in the halo tile)
__device__ __forceinline__ uint32_t px_off(int p, int half) { return (uint32_t)(p * 32 + ((half ^ ((p >> 2) & 1)) << 4)); }

template <typename T, int S, bool POOL>
__global__ void __launch_bounds__(CTHREADS, 2)
conv3x3_c16_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ w, const float* __restrict__ bias,
                   int act, int B, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y, int n_tiles) {
  using G = CGeom<S>;
  extern __shared__ __align__(128) uint8_t csm[];
  const uint32_t sm0 = (uint32_t)__cvta_generic_to_shared(csm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  // B fragments: bf[tap][nt] = {W[nt*8+g][tap*16 + 2t, +1], W[nt*8+g][tap*16 + 2t+8, +9]}
  uint32_t bf[9][4][2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const T* wr = w + (nt * 8 + g) * 144 + tap * 16 + 2 * t;
      bf[tap][nt][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[tap][nt][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  float bv[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    bv[nt][0] = __ldg(bias + nt * 8 + 2 * t);
    bv[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
  }

  auto stage_tile = [&](int tile, int buf) {
    const int b = tile / (tiles_x * tiles_y), r = tile - b * (tiles_x * tiles_y);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int iy0 = ty * CT * S - 1, ix0 = tx * CT * S - 1;
    const T* xb = x + (int64_t)b * H * W * 16;
    for (int i = threadIdx.x; i < G::IN * G::IN * 2; i += CTHREADS) {
      const int p = i >> 1, half = i & 1;
      const int py = p / G::IN, pxx = p - py * G::IN;
      const int iy = iy0 + py, ix = ix0 + pxx;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      const T* src = ok ? xb + ((int64_t)iy * W + ix) * 16 + half * 8 : x;
      const uint32_t dst = sm0 + (uint32_t)buf * G::BYTES + px_off(p, half);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if ((int)blockIdx.x < n_tiles) stage_tile(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < n_tiles) {
      stage_tile(next, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t sb = sm0 + (uint32_t)buf * G::BYTES;

    float acc[2][4][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[rr][nt][0] = acc[rr][nt][2] = bv[nt][0];
        acc[rr][nt][1] = acc[rr][nt][3] = bv[nt][1];
      }
    // ldmatrix.x4 lane roles: matrices (rows 0-7, k lo), (rows 8-15, k lo), (rows 0-7, k hi), (rows 8-15, k hi)
    const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_half = lane >> 4;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int oy = 2 * warp + rr;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int p = (oy * S + kh) * G::IN + a_row * S + kw;
          uint32_t a[4];
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                       : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
                       : "r"(sb + px_off(p, a_half)));
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_16816<T>(acc[rr][nt], a, bf[kh * 3 + kw][nt][0], bf[kh * 3 + kw][nt][1]);
        }
    }
    // epilogue.  acc[rr][nt][e]: conv row 2*warp + rr, pixel g (e = 0, 1) or g + 8 (e = 2, 3), channel nt*8 + 2t + (e & 1)
    const int b = tile / (tiles_x * tiles_y), r = tile - b * (tiles_x * tiles_y);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    if constexpr (POOL) {
      // max over the two conv rows (registers) and over pixel pairs g ^ 1 (lane ^ 4); activation is monotone, so it is
      // applied once after the max
      const int py = ty * (CT / 2) + warp;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float m[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = fmaxf(acc[0][nt][e], acc[1][nt][e]);
          m[e] = act_apply(fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4)), act);
        }
        // even-g lanes store n-tiles 0, 1; odd-g lanes n-tiles 2, 3 (both hold the pooled values)
        if (((g & 1) == 0) == (nt < 2) && py < Ho) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int px = tx * (CT / 2) + (g >> 1) + 4 * hh;
            if (px < Wo)
              *reinterpret_cast<uint32_t*>(y + (((int64_t)b * Ho + py) * Wo + px) * 32 + nt * 8 + 2 * t) =
                  pack2<T>(m[2 * hh], m[2 * hh + 1]);
          }
        }
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int oy = ty * CT + 2 * warp + rr;
        if (oy < Ho) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ox = tx * CT + g + 8 * hh;
            if (ox < Wo) {
              T* dst = y + (((int64_t)b * Ho + oy) * Wo + ox) * 32 + 2 * t;
#pragma unroll
              for (int nt = 0; nt < 4; ++nt)
                *reinterpret_cast<uint32_t*>(dst + nt * 8) =
                    pack2<T>(act_apply(acc[rr][nt][2 * hh], act), act_apply(acc[rr][nt][2 * hh + 1], act));
            }
          }
        }
      }
    }
    __syncthreads();            // all warps are done with this buffer before it is refilled two tiles on
    buf ^= 1;
  }
}

}  // namespace

bool conv3x3_c16_supported(int dtype, int Cin, int Cout, int stride, int pool) {
  return (dtype == GCV_BF16 || dtype == GCV_F16) && Cin == 16 && Cout == 32 &&
         ((stride == 1) || (stride == 2 && !pool));
}

// y = [pool2x2](act(conv3x3(x, w) + bias)); x: [B,H,W,16], w: [32][(kh,kw,ci)] (the GEMM B layout of the im2col path),
// y: [B,Ho,Wo,32] with Ho = (H-1)/stride + 1 (halved again by the pool; the pool needs even conv sizes).
int conv3x3_c16(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
                int H, int W, cudaStream_t stream) {
  GCV_REQUIRE(conv3x3_c16_supported(dtype, 16, 32, stride, pool), "conv3x3_c16: needs a 16-bit dtype, stride 1 or 2 (pool only with stride 1)");
  GCV_REQUIRE(B > 0 && H > 0 && W > 0, "conv3x3_c16: bad shape");
  GCV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
              "conv3x3_c16: x, y, w must be 16-byte aligned");
  const int Hc = (H - 1) / stride + 1, Wc = (W - 1) / stride + 1;           // conv output
  GCV_REQUIRE(!pool || (Hc % 2 == 0 && Wc % 2 == 0), "conv3x3_c16: the fused 2x2 max-pool needs even conv output sizes");
  const int Ho = pool ? Hc / 2 : Hc, Wo = pool ? Wc / 2 : Wc;
  const int tiles_x = (Wc + CT - 1) / CT, tiles_y = (Hc + CT - 1) / CT;
  const int64_t n_tiles64 = (int64_t)B * tiles_x * tiles_y;
  GCV_REQUIRE(n_tiles64 < 2147483647LL, "conv3x3_c16: too many tiles");
  const int n_tiles = (int)n_tiles64;
  const int sms = device_sms();
  const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;
#define GCV_CONV_LAUNCH(T, S, P)                                                                                         \
  do {                                                                                                                   \
    const size_t smem = 2 * (size_t)CGeom<S>::BYTES;                                                                     \
    static unsigned long long attr_devs = 0;                                                                                       \
    if (first_on_device(attr_devs)) {                                                                                                    \
      cudaFuncSetAttribute(conv3x3_c16_kernel<T, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
    }                                                                                                                    \
    conv3x3_c16_kernel<T, S, P><<<grid, CTHREADS, smem, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), \
                                                                  reinterpret_cast<const T*>(w), bias, act, B, H, W, Ho,  \
                                                                  Wo, tiles_x, tiles_y, n_tiles);                        \
  } while (0)
  if (dtype == GCV_BF16) {
    if (stride == 1 && pool) GCV_CONV_LAUNCH(__nv_bfloat16, 1, true);
    else if (stride == 1) GCV_CONV_LAUNCH(__nv_bfloat16, 1, false);
    else GCV_CONV_LAUNCH(__nv_bfloat16, 2, false);
  } else {
    if (stride == 1 && pool) GCV_CONV_LAUNCH(__half, 1, true);
    else if (stride == 1) GCV_CONV_LAUNCH(__half, 1, false);
    else GCV_CONV_LAUNCH(__half, 2, false);
  }
#undef GCV_CONV_LAUNCH
  return check_launch("conv3x3_c16");
}

}  // namespace gcv
