// The decoders' small-channel tail: ConvTranspose2d(k2, s2) + activation on mma.sync (HMMA), NHWC, 16-bit.
// Reference: model/genconvit_ed.py:45-57 (ConvTranspose2d + ReLU), model/genconvit_vae.py:50-63 (+ LeakyReLU).
//
// A k2 s2 transposed convolution is a per-pixel GEMM: out[b, 2h+i, 2w+j, :] = W[:, :, i, j]^T x[b, h, w, :] + bias,
// i.e. [tokens, Ci] x [Ci, 4 Co] with a pixel-shuffle store.  For Ci = 64 / 32 the contraction is 1-2 tcgen05
// k-blocks deep and the 128-row tile machinery of gemm_tcgen05.cu runs at a few % of its rate (25-70 TFLOP/s);
// these layers are pure streaming (<= 200 B in + out per token), so the right shape is small: one warp owns 16
// consecutive tokens (one m16 MMA row block), keeps the whole [4 Co, Ci] weight as B fragments (registers for
// Ci = 32, shared memory for Ci = 64), and stores straight from the accumulator fragments.
//
// TAIL = true fuses the last two layers (32 -> 16 -> 3): the first layer's accumulator fragment of one tap
// (16 tokens x 16 channels, bias + activation applied, rounded to the activation type exactly like the stored
// tensor it replaces) IS the A fragment of the second layer's m16n8k16 MMA, so the [B, 2H, 2W, 16] intermediate
// (103 MB written + read per 256 ED frames) never exists.  The second layer's weight is rounded to the activation
// type like every other weight of the 16-bit modes.
#include <type_traits>

#include "common.cuh"

namespace gcv {

namespace {

template <typename T>
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float act2(float v, int act) {      // the decoders only use ReLU / LeakyReLU(0.01)
  return act == GCV_ACT_RELU ? fmaxf(v, 0.0f) : (act == GCV_ACT_LEAKY ? (v > 0.0f ? v : 0.01f * v) : v);
}

constexpr int CT_THREADS = 256;

// x [M, K] tokens (M = B*H*W, a multiple of 16 and < 2^31), w1 [(i,j,co), K],
// b1 [4*CO] (bias repeated per tap).  !TAIL: y [B, 2H, 2W, CO].  TAIL (K = 32): w2 [(i2,j2,c), 16] = [12][16],
// b2 [12], y [B, 4H, 4W, 3].
template <typename T, int K, bool TAIL>
__global__ void __launch_bounds__(CT_THREADS)
convt2x2_mma_kernel(const T* __restrict__ x, T* __restrict__ y, const T* __restrict__ w1, const float* __restrict__ b1,
                    const T* __restrict__ w2, const float* __restrict__ b2, int act, int64_t tiles, int H, int W) {
  constexpr int CO = K / 2, N = 4 * CO, NT = N / 8, KS = K / 16;
  constexpr bool B_REGS = K == 32;                 // 32 registers of B fragments; K = 64 would need 128
  constexpr int PITCH = K + 8;                     // 16-bit elements per staged weight row: conflict-free fragment reads
  __shared__ __align__(16) T ws[B_REGS ? 8 : N * PITCH];
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

  uint32_t bw[B_REGS ? NT : 1][B_REGS ? KS : 1][2];
  if constexpr (B_REGS) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(w1 + (nt * 8 + g) * K + ks * 16);
        bw[nt][ks][0] = __ldg(r + t);
        bw[nt][ks][1] = __ldg(r + 4 + t);
      }
  } else {
    for (int i = threadIdx.x; i < N * K / 8; i += CT_THREADS) {
      const int n = i / (K / 8), c = i % (K / 8);
      *reinterpret_cast<uint4*>(ws + n * PITCH + c * 8) = __ldg(reinterpret_cast<const uint4*>(w1 + n * K) + c);
    }
    __syncthreads();
  }
  // second layer: B fragment of W2^T, n2 = (i2, j2, c) in rows 0-11 (12-15 are zero), k = the 16 channels
  uint32_t bw2[2][2] = {{0u, 0u}, {0u, 0u}};
  float bias2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  if constexpr (TAIL) {
#pragma unroll
    for (int n2 = 0; n2 < 2; ++n2) {
      const int row = n2 * 8 + g;
      if (row < 12) {
        const uint32_t* r = reinterpret_cast<const uint32_t*>(w2 + row * 16);
        bw2[n2][0] = __ldg(r + t);
        bw2[n2][1] = __ldg(r + 4 + t);
      }
      const int col = n2 * 8 + 2 * t;
      if (col < 12) {
        bias2[n2][0] = __ldg(b2 + col);
        bias2[n2][1] = __ldg(b2 + col + 1);
      }
    }
  }

  const int64_t warp0 = (int64_t)blockIdx.x * (CT_THREADS / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (CT_THREADS / 32);
  uint32_t a[KS][4], an[KS][4];
  auto load_a = [&](int64_t tile, uint32_t (&f)[KS][4]) {
    const uint32_t* r0 = reinterpret_cast<const uint32_t*>(x + (tile * 16 + g) * K);
    const uint32_t* r1 = r0 + 8 * (K / 2);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      f[ks][0] = __ldg(r0 + ks * 8 + t);
      f[ks][1] = __ldg(r1 + ks * 8 + t);
      f[ks][2] = __ldg(r0 + ks * 8 + 4 + t);
      f[ks][3] = __ldg(r1 + ks * 8 + 4 + t);
    }
  };
  if (warp0 < tiles) load_a(warp0, an);
  for (int64_t tile = warp0; tile < tiles; tile += nwarps) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int q = 0; q < 4; ++q) a[ks][q] = an[ks][q];
    if (tile + nwarps < tiles) load_a(tile + nwarps, an);

    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 bv = __ldg(reinterpret_cast<const float2*>(b1 + nt * 8 + 2 * t));
      acc[nt][0] = bv.x, acc[nt][1] = bv.y, acc[nt][2] = bv.x, acc[nt][3] = bv.y;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        if constexpr (B_REGS) {
          hmma<T>(acc[nt], a[ks], bw[nt][ks][0], bw[nt][ks][1]);
        } else {
          const uint32_t* r = reinterpret_cast<const uint32_t*>(ws + (nt * 8 + g) * PITCH + ks * 16);
          hmma<T>(acc[nt], a[ks], r[t], r[4 + t]);
        }
      }
    }
    // rows g and g + 8 of the tile: tokens tile*16 + g and tile*16 + 8 + g
    int64_t obase[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const uint32_t m = (uint32_t)tile * 16u + hf * 8 + g;
      const uint32_t bh = m / (uint32_t)W, wq = m - bh * (uint32_t)W;
      const uint32_t b = bh / (uint32_t)H, h = bh - b * (uint32_t)H;
      if constexpr (TAIL) obase[hf] = (((int64_t)b * (4 * H) + 4 * h) * (int64_t)(4 * W) + 4 * wq) * 3;
      else obase[hf] = (((int64_t)b * (2 * H) + 2 * h) * (int64_t)(2 * W) + 2 * wq) * CO;
    }
    if constexpr (!TAIL) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n = nt * 8 + 2 * t, tap = n / CO, co = n % CO;
        const int64_t off = ((tap >> 1) * (int64_t)(2 * W) + (tap & 1)) * CO + co;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
          *reinterpret_cast<uint32_t*>(y + obase[hf] + off) =
              pack2<T>(act2(acc[nt][2 * hf], act), act2(acc[nt][2 * hf + 1], act));
      }
    } else {
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        uint32_t a2[4];
        a2[0] = pack2<T>(act2(acc[2 * tap][0], act), act2(acc[2 * tap][1], act));
        a2[1] = pack2<T>(act2(acc[2 * tap][2], act), act2(acc[2 * tap][3], act));
        a2[2] = pack2<T>(act2(acc[2 * tap + 1][0], act), act2(acc[2 * tap + 1][1], act));
        a2[3] = pack2<T>(act2(acc[2 * tap + 1][2], act), act2(acc[2 * tap + 1][3], act));
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2) {
          float d[4] = {bias2[n2][0], bias2[n2][1], bias2[n2][0], bias2[n2][1]};
          hmma<T>(d, a2, bw2[n2][0], bw2[n2][1]);
          const int col = n2 * 8 + 2 * t;              // (i2, j2, c) = (col / 6, col / 3 % 2, col % 3); col is even
          if (col < 12) {
            const int i2 = col / 6;
            const int64_t off = ((2 * (tap >> 1) + i2) * (int64_t)(4 * W) + 2 * (tap & 1)) * 3 + (col - 6 * i2);
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
              *reinterpret_cast<uint32_t*>(y + obase[hf] + off) =
                  pack2<T>(act2(d[2 * hf], act), act2(d[2 * hf + 1], act));
          }
        }
      }
    }
  }
}

}  // namespace

int convt2x2_mma(int dtype, const void* x, void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                 int act, int B, int H, int W, int CI, int tail, cudaStream_t stream) {
  GCV_REQUIRE(dtype == GCV_F16 || dtype == GCV_BF16, "convt2x2_mma: needs a 16-bit dtype");
  GCV_REQUIRE(CI == 32 || (CI == 64 && !tail), "convt2x2_mma: Ci must be 32 (optionally with the fused 16 -> 3 tail) or 64 (got %d)", CI);
  GCV_REQUIRE(B > 0 && H > 0 && W > 0, "convt2x2_mma: bad shape");
  GCV_REQUIRE(x && y && w1 && b1 && (!tail || (w2 && b2)), "convt2x2_mma: null pointer");
  GCV_REQUIRE(act == GCV_ACT_NONE || act == GCV_ACT_RELU || act == GCV_ACT_LEAKY, "convt2x2_mma: activation %d", act);
  GCV_REQUIRE(((uintptr_t)x | (uintptr_t)y | (uintptr_t)w1 | (uintptr_t)w2) % 16 == 0, "convt2x2_mma: 16-byte alignment");
  const int64_t tokens = (int64_t)B * H * W;
  GCV_REQUIRE(tokens % 16 == 0 && tokens < 2147483647LL, "convt2x2_mma: B*H*W must be a multiple of 16 below 2^31 (got %lld)", (long long)tokens);
  const int64_t tiles = tokens / 16;
  const int64_t want = (tiles + CT_THREADS / 32 - 1) / (CT_THREADS / 32);
  const int64_t cap = (int64_t)device_sms() * 8;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
    const T* xp = reinterpret_cast<const T*>(x);
    T* yp = reinterpret_cast<T*>(y);
    const T* w1p = reinterpret_cast<const T*>(w1);
    const T* w2p = reinterpret_cast<const T*>(w2);
    if (CI == 64)
      convt2x2_mma_kernel<T, 64, false><<<grid, CT_THREADS, 0, stream>>>(xp, yp, w1p, b1, w2p, b2, act, tiles, H, W);
    else if (tail)
      convt2x2_mma_kernel<T, 32, true><<<grid, CT_THREADS, 0, stream>>>(xp, yp, w1p, b1, w2p, b2, act, tiles, H, W);
    else
      convt2x2_mma_kernel<T, 32, false><<<grid, CT_THREADS, 0, stream>>>(xp, yp, w1p, b1, w2p, b2, act, tiles, H, W);
    return check_launch("convt2x2_mma");
  };
  return dtype == GCV_F16 ? launch(__half{}) : launch(__nv_bfloat16{});
}

}  // namespace gcv
