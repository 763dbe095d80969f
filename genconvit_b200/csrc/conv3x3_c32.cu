// Direct 3x3 convolution (pad 1) for the autoencoders' third layer, Cin = 32 -> Cout = 64, NHWC 16-bit, on the tensor
// cores (mma.sync m16n8k16, fp32 accumulate) with the activation and, for Network A, the 2x2 max-pool fused in:
//   stride 1 + ReLU + MaxPool2d(2)   reference model/genconvit_ed.py:22-24  (Conv2d(32,64,3,p1), ReLU, MaxPool2d)
//   stride 2 + LeakyReLU(0.01)       reference model/genconvit_vae.py:22-24 (Conv2d(32,64,3,s2,p1), BN folded, LeakyReLU)
// Same scheme as conv3x3_c16.cu (no im2col matrix: 462 MB written and re-read per batch for Network A's layer), but
// K = 9 x 32 = 288 and N = 64 no longer fit in registers as B fragments: the [64][288] weight matrix sits in shared
// memory (row pitch 592 B, so the 8 rows of an ldmatrix hit 8 different 16-byte bank groups) and is read with
// ldmatrix.x4, four loads per K = 16 slice feeding 8 (pool: 16) MMAs.  A pixel is 64 B = four 16-byte pieces,
// XOR-swizzled with the pixel index like TMA's SWIZZLE_64B so that the A ldmatrix is conflict-free as well.
#include "common.cuh"

namespace gcv {

namespace {

constexpr int C3T = 16;          // conv-output tile edge
constexpr int C3_THREADS = 256;
constexpr int C3_WPITCH = 592;   // bytes per weight row in shared memory (288 x 2 + 16)
constexpr int C3_WBYTES = 64 * C3_WPITCH;

template <int S> struct C3Geom {
  static constexpr int IN = C3T * S + 2;
  static constexpr int BYTES = IN * IN * 64;
};

template <typename T>
__device__ __forceinline__ void c3_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void c3_ldm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__device__ __forceinline__ float c3_act(float v, int act) {
  return act == GCV_ACT_RELU ? fmaxf(v, 0.0f) : (act == GCV_ACT_LEAKY ? (v > 0.0f ? v : 0.01f * v) : v);
}

// byte offset of 16-byte piece `piece` (channels 8*piece ..) of staged pixel p
__device__ __forceinline__ uint32_t c3_px(int p, int piece) { return (uint32_t)(p * 64 + ((piece ^ ((p >> 1) & 3)) << 4)); }

template <typename T, int S, bool POOL>
__global__ void __launch_bounds__(C3_THREADS, S == 1 ? 2 : 1)
conv3x3_c32_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ w, const float* __restrict__ bias,
                   int act, int B, int H, int W, int Ho, int Wo, int tiles_x, int tiles_y, int n_tiles) {
  using G = C3Geom<S>;
  extern __shared__ __align__(128) uint8_t c3sm[];
  const uint32_t wsm = (uint32_t)__cvta_generic_to_shared(c3sm);
  const uint32_t sm0 = wsm + C3_WBYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  // weights [64][288] -> shared memory rows of 592 B
  for (int i = threadIdx.x; i < 64 * 36; i += C3_THREADS) {
    const int r = i / 36, pc = i - r * 36;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wsm + r * C3_WPITCH + pc * 16), "l"(w + r * 288 + pc * 8)
                 : "memory");
  }
  float bv[8][2];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    bv[nt][0] = __ldg(bias + nt * 8 + 2 * t);
    bv[nt][1] = __ldg(bias + nt * 8 + 2 * t + 1);
  }

  auto stage_tile = [&](int tile, int buf) {
    const int b = tile / (tiles_x * tiles_y), r = tile - b * (tiles_x * tiles_y);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    const int iy0 = ty * C3T * S - 1, ix0 = tx * C3T * S - 1;
    const T* xb = x + (int64_t)b * H * W * 32;
    for (int i = threadIdx.x; i < G::IN * G::IN * 4; i += C3_THREADS) {
      const int p = i >> 2, piece = i & 3;
      const int py = p / G::IN, pxx = p - py * G::IN;
      const int iy = iy0 + py, ix = ix0 + pxx;
      const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
      const T* src = ok ? xb + ((int64_t)iy * W + ix) * 32 + piece * 8 : x;
      const uint32_t dst = sm0 + (uint32_t)buf * G::BYTES + c3_px(p, piece);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int buf = 0;
  if ((int)blockIdx.x < n_tiles) stage_tile(blockIdx.x, 0);          // (the weight copies ride in this first group)
  else asm volatile("cp.async.commit_group;" ::: "memory");
  // ldmatrix lane roles.  A: matrices (rows 0-7, k lo), (rows 8-15, k lo), (rows 0-7, k hi), (rows 8-15, k hi);
  // B: (n-tile j, k lo), (n-tile j, k hi), (n-tile j+1, k lo), (n-tile j+1, k hi)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_hi = lane >> 4;
  const uint32_t b_lane = wsm + (uint32_t)(((lane >> 4) * 8 + (lane & 7)) * C3_WPITCH + ((lane >> 3) & 1) * 16);
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

    float acc[2][8][4];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        acc[rr][nt][0] = acc[rr][nt][2] = bv[nt][0];
        acc[rr][nt][1] = acc[rr][nt][3] = bv[nt][1];
      }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw)
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          const int s = (kh * 3 + kw) * 2 + hc;          // K = 16 slice: tap (kh, kw), channels 16 hc ..
          uint32_t a[2][4];
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const int p = ((2 * warp + rr) * S + kh) * G::IN + a_row * S + kw;
            c3_ldm4(a[rr], sb + c3_px(p, 2 * hc + a_hi));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t bq[4];
            c3_ldm4(bq, b_lane + (uint32_t)(j * 16 * C3_WPITCH + s * 32));
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              c3_mma<T>(acc[rr][2 * j], a[rr], bq[0], bq[1]);
              c3_mma<T>(acc[rr][2 * j + 1], a[rr], bq[2], bq[3]);
            }
          }
        }
    // epilogue.  acc[rr][nt][e]: conv row 2*warp + rr, pixel g (e = 0, 1) or g + 8 (e = 2, 3), channel nt*8 + 2t + (e & 1)
    const int b = tile / (tiles_x * tiles_y), r = tile - b * (tiles_x * tiles_y);
    const int ty = r / tiles_x, tx = r - ty * tiles_x;
    if constexpr (POOL) {
      const int py = ty * (C3T / 2) + warp;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float m[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float v = fmaxf(acc[0][nt][e], acc[1][nt][e]);
          m[e] = c3_act(fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4)), act);     // monotone act: after the max
        }
        // even-g lanes store n-tiles 0-3, odd-g lanes n-tiles 4-7 (both hold the pooled values)
        if (((g & 1) == 0) == (nt < 4) && py < Ho) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int px = tx * (C3T / 2) + (g >> 1) + 4 * hh;
            if (px < Wo)
              *reinterpret_cast<uint32_t*>(y + (((int64_t)b * Ho + py) * Wo + px) * 64 + nt * 8 + 2 * t) =
                  pack2<T>(m[2 * hh], m[2 * hh + 1]);
          }
        }
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int oy = ty * C3T + 2 * warp + rr;
        if (oy < Ho) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int ox = tx * C3T + g + 8 * hh;
            if (ox < Wo) {
              T* dst = y + (((int64_t)b * Ho + oy) * Wo + ox) * 64 + 2 * t;
#pragma unroll
              for (int nt = 0; nt < 8; ++nt)
                *reinterpret_cast<uint32_t*>(dst + nt * 8) =
                    pack2<T>(c3_act(acc[rr][nt][2 * hh], act), c3_act(acc[rr][nt][2 * hh + 1], act));
            }
          }
        }
      }
    }
    __syncthreads();            // all warps are done with this buffer before it is refilled two tiles on
    buf ^= 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace

// y = [pool2x2](act(conv3x3(x, w) + bias)); x: [B,H,W,32], w: [64][(kh,kw,ci)] (the GEMM B layout of the im2col path),
// y: [B,Ho,Wo,64].  stride 1 (optionally + 2x2 max-pool) or stride 2; 16-bit dtypes.
int conv3x3_c32(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
                int H, int W, cudaStream_t stream) {
  GCV_REQUIRE((dtype == GCV_BF16 || dtype == GCV_F16) && (stride == 1 || (stride == 2 && !pool)),
              "conv3x3_c32: needs a 16-bit dtype, stride 1 or 2 (pool only with stride 1)");
  GCV_REQUIRE(B > 0 && H > 0 && W > 0, "conv3x3_c32: bad shape");
  GCV_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15) == 0,
              "conv3x3_c32: x, y, w must be 16-byte aligned");
  const int Hc = (H - 1) / stride + 1, Wc = (W - 1) / stride + 1;
  GCV_REQUIRE(!pool || (Hc % 2 == 0 && Wc % 2 == 0), "conv3x3_c32: the fused 2x2 max-pool needs even conv output sizes");
  const int Ho = pool ? Hc / 2 : Hc, Wo = pool ? Wc / 2 : Wc;
  const int tiles_x = (Wc + C3T - 1) / C3T, tiles_y = (Hc + C3T - 1) / C3T;
  const int64_t n_tiles64 = (int64_t)B * tiles_x * tiles_y;
  GCV_REQUIRE(n_tiles64 < 2147483647LL, "conv3x3_c32: too many tiles");
  const int n_tiles = (int)n_tiles64;
  const int sms = device_sms();
  const int per_sm = stride == 1 ? 2 : 1;
  const int grid = n_tiles < per_sm * sms ? n_tiles : per_sm * sms;
#define GCV_CONV3_LAUNCH(T, S, P)                                                                                         \
  do {                                                                                                                    \
    const size_t smem = C3_WBYTES + 2 * (size_t)C3Geom<S>::BYTES;                                                         \
    static unsigned long long attr_devs = 0;                                                                                        \
    if (first_on_device(attr_devs)) {                                                                                                     \
      cudaFuncSetAttribute(conv3x3_c32_kernel<T, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    }                                                                                                                     \
    conv3x3_c32_kernel<T, S, P><<<grid, C3_THREADS, smem, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), \
                                                                    reinterpret_cast<const T*>(w), bias, act, B, H, W, Ho, \
                                                                    Wo, tiles_x, tiles_y, n_tiles);                       \
  } while (0)
  if (dtype == GCV_BF16) {
    if (stride == 1 && pool) GCV_CONV3_LAUNCH(__nv_bfloat16, 1, true);
    else if (stride == 1) GCV_CONV3_LAUNCH(__nv_bfloat16, 1, false);
    else GCV_CONV3_LAUNCH(__nv_bfloat16, 2, false);
  } else {
    if (stride == 1 && pool) GCV_CONV3_LAUNCH(__half, 1, true);
    else if (stride == 1) GCV_CONV3_LAUNCH(__half, 1, false);
    else GCV_CONV3_LAUNCH(__half, 2, false);
  }
#undef GCV_CONV3_LAUNCH
  return check_launch("conv3x3_c32");
}

}  // namespace gcv
