// ConvNeXt stem in one pass: Conv2d(3 -> 96, k4, s4) + bias + LayerNorm2d(96), frames in -> NHWC tokens out.
// timm==0.6.5 ConvNeXt.stem = Sequential(Conv2d(3, 96, 4, 4), LayerNorm2d(96, eps=1e-6)) as reached from reference
// model/genconvit_ed.py:82-83 and model/genconvit_vae.py:111-112 (arithmetic restated in oracle/backbones.py).
//
// As three kernels (4x4 im2col -> [M,48] GEMM -> row LayerNorm) the stem moved 1.9 GB per 256-frame batch for a
// K = 48 contraction; here every input pixel is read once and every token written once.  The contraction runs on
// mma.sync m16n8k16 (fp32 accumulate): a warp owns 16 consecutive tokens x all 96 channels (12 n-tiles x 3 k-slices),
// the B fragments (the whole [96][48] weight matrix) stay in registers, A fragments come straight from global memory:
//   NCHW fp32 frames (the reference's input layout): K order (c, kh, kw) = the conv weight's own OIHW flattening, so a
//     k-slice is one channel and lane (g, t) needs two adjacent pixels of patch row t/2 (+2): one 8-byte load, 8 tokens
//     of a row make a contiguous 128-byte segment;
//   NHWC 16-bit images (the autoencoders' reconstructions): K order (kh, kw, c), a patch row is 24 contiguous bytes.
//   NHWC uint8 face crops (what model/pred_func.py:95-108 starts from): the Normalize of dataset/loader.py:63-77 is
//     a 3 x 256 table of (u / 255 - mean) / std rounded to the activation type -- bit for bit the operand the fp32 path
//     builds from the pre-processed frame -- and the K order is the NCHW path's, so both paths give identical tokens.
// The LayerNorm is row-local: a token's 96 values live in the 4 lanes of a quad (24 each), exact two-pass statistics
// with two shuffles per pass.
#include "common.cuh"

namespace gcv {

namespace {

constexpr int ST_THREADS = 256;

template <typename T>
__device__ __forceinline__ void st_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct StemNorm { float mean[3], sd[3]; };        // SRC 2: Normalize(mean, std) of the uint8 source

// SRC 1: x = fp32 [B,3,H,W], w = [96][(c,kh,kw)] of T.  SRC 0: x = T [B,H,W,3], w = [96][(kh,kw,c)] of T.
// SRC 2: x = uint8 [B,H,W,3] (raw crops, normalised here), w = [96][(c,kh,kw)] of T.
template <typename T, int SRC>
__global__ void __launch_bounds__(ST_THREADS, 1)
stem_fused_kernel(const void* __restrict__ xin, T* __restrict__ y, const T* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps, int B, int H, int W,
                  int64_t M, const StemNorm nrm) {
  constexpr bool NCHW = SRC == 1;
  __shared__ float s_b[96], s_lw[96], s_lb[96];
  // SRC 2: table[c][u][lane] as 32-bit words (dynamic shared memory, 96 KB): lane l only ever reads bank l, so the 24
  // look-ups a lane makes per tile are conflict-free whatever the pixel values are
  extern __shared__ uint32_t s_lut[];
  for (int i = threadIdx.x; i < 96; i += ST_THREADS) {
    s_b[i] = bias[i];
    s_lw[i] = ln_w[i];
    s_lb[i] = ln_b[i];
  }
  if constexpr (SRC == 2) {
    // the host arithmetic of preprocess_frame, IEEE division and all (gcv_preprocess_frames), then the fp32 -> T
    // rounding the NCHW path applies to the pre-processed frame
    for (int i = threadIdx.x; i < 3 * 256 * 32; i += ST_THREADS) {
      const int c = i >> 13, u = (i >> 5) & 255;
      const T v = from_f<T>(__fdiv_rn(__fdiv_rn((float)u, 255.0f) - nrm.mean[c], nrm.sd[c]));
      s_lut[i] = *reinterpret_cast<const uint16_t*>(&v);
    }
  }
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // B fragments: bf[s][nt] = {W[nt*8+g][16s + 2t, +1], W[nt*8+g][16s + 2t+8, +9]}
  uint32_t bf[3][12][2];
#pragma unroll
  for (int s = 0; s < 3; ++s)
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const T* wr = w + (nt * 8 + g) * 48 + s * 16 + 2 * t;
      bf[s][nt][0] = *reinterpret_cast<const uint32_t*>(wr);
      bf[s][nt][1] = *reinterpret_cast<const uint32_t*>(wr + 8);
    }
  __syncthreads();
  const int Wt = W >> 2, Ht = H >> 2;
  const int64_t tiles = (M + 15) >> 4;
  const int warps = (int)(gridDim.x * (ST_THREADS / 32));
  // raw operands of a tile: loaded one tile ahead, so that their HBM latency overlaps the previous tile's MMAs,
  // LayerNorm and stores (one CTA of 8 warps per SM: the loads in flight are what feeds the memory system)
  using Raw = typename std::conditional<NCHW, float2, uint32_t>::type;      // SRC 2: look-up offsets of pixels kw, kw + 1
  Raw raw[3][4];
  auto load_raw = [&](int64_t tile) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int64_t m = tile * 16 + g + 8 * r;
      if (m >= M) m = M - 1;
      const int ox = (int)(m % Wt);
      const int64_t q = m / Wt;
      const int oy = (int)(q % Ht);
      const int64_t b = q / Ht;
      if constexpr (NCHW) {
        // slice s = channel; k = kh*4 + kw: lane needs (kh = t/2, kw = 2(t%2), +1) and kh + 2
        const float* xb = reinterpret_cast<const float*>(xin) + ((b * 3) * H + 4 * oy + (t >> 1)) * (int64_t)W + 4 * ox + 2 * (t & 1);
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          raw[s][r] = __ldg(reinterpret_cast<const float2*>(xb + (int64_t)s * H * W));
          raw[s][r + 2] = __ldg(reinterpret_cast<const float2*>(xb + (int64_t)s * H * W + 2 * (int64_t)W));
        }
      } else if constexpr (SRC == 2) {
        // same K order as NCHW: slice s = channel, lane needs pixels (kh = t/2, kw = 2(t%2), +1) and kh + 2
        const uint8_t* xb = reinterpret_cast<const uint8_t*>(xin) + ((b * H + 4 * oy + (t >> 1)) * (int64_t)W + 4 * ox + 2 * (t & 1)) * 3;
        // the two pixels are six contiguous bytes R G B R G B at an even offset: three 16-bit loads per patch row
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const uint16_t* p = reinterpret_cast<const uint16_t*>(xb + kk * 6 * (int64_t)W);
          const uint32_t rg = __ldg(p), br = __ldg(p + 1), gb = __ldg(p + 2);
          raw[0][r + 2 * kk] = (rg & 255u) | ((br >> 8) << 8);            // R0, R1
          raw[1][r + 2 * kk] = (rg >> 8) | ((gb & 255u) << 8);            // G0, G1
          raw[2][r + 2 * kk] = (br & 255u) | ((gb >> 8) << 8);            // B0, B1
        }
      } else {
        // k = kh*12 + kw*3 + c; patch row kh = 12 contiguous elements at pixel (4oy + kh, 4ox)
        const T* xb = reinterpret_cast<const T*>(xin) + ((b * H + 4 * oy) * (int64_t)W + 4 * ox) * 3;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int k0 = 16 * s + 2 * t, k1 = k0 + 8;
          raw[s][r] = __ldg(reinterpret_cast<const uint32_t*>(xb + (int64_t)(k0 / 12) * W * 3 + k0 % 12));
          raw[s][r + 2] = __ldg(reinterpret_cast<const uint32_t*>(xb + (int64_t)(k1 / 12) * W * 3 + k1 % 12));
        }
      }
    }
  };
  const int64_t tile_first = (int64_t)blockIdx.x * (ST_THREADS / 32) + (threadIdx.x >> 5);
  if (tile_first < tiles) load_raw(tile_first);
  for (int64_t tile = tile_first; tile < tiles; tile += warps) {
    // this lane's two tokens: rows g and g + 8 of the tile
    uint32_t a[3][4];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (NCHW) a[s][i] = pack2<T>(raw[s][i].x, raw[s][i].y);
        else if constexpr (SRC == 2)
          a[s][i] = s_lut[(s * 256 + (raw[s][i] & 255u)) * 32 + lane] | (s_lut[(s * 256 + (raw[s][i] >> 8)) * 32 + lane] << 16);
        else a[s][i] = raw[s][i];
      }
    if (tile + warps < tiles) load_raw(tile + warps);
    float acc[12][4];
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const float b0 = s_b[nt * 8 + 2 * t], b1 = s_b[nt * 8 + 2 * t + 1];
      acc[nt][0] = acc[nt][2] = b0;
      acc[nt][1] = acc[nt][3] = b1;
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) st_mma<T>(acc[nt], a[s], bf[s][nt][0], bf[s][nt][1]);
    // LayerNorm over the 96 channels of each token (rows g: acc[.][0..1], g+8: acc[.][2..3]); exact two-pass
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      s0 += acc[nt][0] + acc[nt][1];
      s1 += acc[nt][2] + acc[nt][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float mean0 = s0 * (1.0f / 96.0f), mean1 = s1 * (1.0f / 96.0f);
    float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 12; ++nt) {
      const float d0 = acc[nt][0] - mean0, d1 = acc[nt][1] - mean0, d2 = acc[nt][2] - mean1, d3 = acc[nt][3] - mean1;
      q0 = fmaf(d0, d0, fmaf(d1, d1, q0));
      q1 = fmaf(d2, d2, fmaf(d3, d3, q1));
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * (1.0f / 96.0f) + eps), r1 = rsqrtf(q1 * (1.0f / 96.0f) + eps);
    const int64_t m0 = tile * 16 + g, m1 = m0 + 8;
    // pairs of lanes swap halves so that each lane stores 8 contiguous bytes (4 channels) per n-tile pair:
    // even t keeps n-tile 2j (its own 2 channels + the partner's 2), odd t keeps n-tile 2j+1
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      uint32_t mine[2][2];        // [n-tile of the pair][row]: this lane's normalised channel pair
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int nt = 2 * j + h, c = nt * 8 + 2 * t;
        const float w0 = s_lw[c], w1 = s_lw[c + 1], lb0 = s_lb[c], lb1 = s_lb[c + 1];
        mine[h][0] = pack2<T>(fmaf((acc[nt][0] - mean0) * r0, w0, lb0), fmaf((acc[nt][1] - mean0) * r0, w1, lb1));
        mine[h][1] = pack2<T>(fmaf((acc[nt][2] - mean1) * r1, w0, lb0), fmaf((acc[nt][3] - mean1) * r1, w1, lb1));
      }
      const int keep = t & 1;                    // n-tile of the pair this lane ends up storing
      uint2 o0, o1;
      {
        const uint32_t give0 = mine[keep ^ 1][0], give1 = mine[keep ^ 1][1];
        const uint32_t got0 = __shfl_xor_sync(0xffffffffu, give0, 1), got1 = __shfl_xor_sync(0xffffffffu, give1, 1);
        // channels ascend with t: the even lane's pair comes first
        o0 = keep == 0 ? make_uint2(mine[0][0], got0) : make_uint2(got0, mine[1][0]);
        o1 = keep == 0 ? make_uint2(mine[0][1], got1) : make_uint2(got1, mine[1][1]);
      }
      const int col = (2 * j + keep) * 8 + 4 * (t >> 1);
      if (m0 < M) *reinterpret_cast<uint2*>(y + m0 * 96 + col) = o0;
      if (m1 < M) *reinterpret_cast<uint2*>(y + m1 * 96 + col) = o1;
    }
  }
}

}  // namespace

// y[B*(H/4)*(W/4), 96] = LayerNorm(conv4x4s4(x) + bias).  src 1 (nchw): x fp32 [B,3,H,W] and w = [96][(c,kh,kw)];
// src 0: x `dtype` [B,H,W,3] and w = [96][(kh,kw,c)]; src 2: x uint8 [B,H,W,3], normalised with (mean3, std3) (host
// pointers) on the fly, w = [96][(c,kh,kw)].  16-bit dtypes only.
int stem_fused_src(int dtype, int src, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
                   const float* ln_b, float eps, int B, int H, int W, const float* mean3, const float* std3,
                   cudaStream_t stream) {
  GCV_REQUIRE(dtype == GCV_BF16 || dtype == GCV_F16, "stem_fused: 16-bit dtypes only");
  GCV_REQUIRE(src >= 0 && src <= 2 && (src != 2 || (mean3 && std3)), "stem_fused: bad source kind %d (uint8 needs mean / std)", src);
  GCV_REQUIRE(B > 0 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0, "stem_fused: H, W must be positive multiples of 4");
  GCV_REQUIRE((src == 2 || (reinterpret_cast<uintptr_t>(x) & 7) == 0) && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(w) & 3) == 0,
              "stem_fused: misaligned pointer");
  const int64_t M = (int64_t)B * (H / 4) * (W / 4);
  const int sms = device_sms();
  const int64_t tiles = (M + 15) / 16, per_cta = ST_THREADS / 32;
  const int64_t want = (tiles + per_cta - 1) / per_cta;
  const int grid = (int)(want < sms ? want : sms);   // ~200 registers per thread: one CTA per SM
  StemNorm nrm{};
  if (src == 2)
    for (int c = 0; c < 3; ++c) { nrm.mean[c] = mean3[c]; nrm.sd[c] = std3[c]; }
  const size_t smem = src == 2 ? (size_t)3 * 256 * 32 * 4 : 0;
#define GCV_STEM_LAUNCH(T, N)                                                                                              \
  do {                                                                                                                     \
    if (N == 2) {                                                                                                          \
      static unsigned long long attr_devs = 0;                                                                             \
      if (first_on_device(attr_devs))                                                                                      \
        cudaFuncSetAttribute(stem_fused_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);             \
    }                                                                                                                      \
    stem_fused_kernel<T, N><<<grid, ST_THREADS, smem, stream>>>(x, reinterpret_cast<T*>(y), reinterpret_cast<const T*>(w), \
                                                                bias, ln_w, ln_b, eps, B, H, W, M, nrm);                   \
  } while (0)
  if (dtype == GCV_BF16) {
    if (src == 1) GCV_STEM_LAUNCH(__nv_bfloat16, 1);
    else if (src == 2) GCV_STEM_LAUNCH(__nv_bfloat16, 2);
    else GCV_STEM_LAUNCH(__nv_bfloat16, 0);
  } else {
    if (src == 1) GCV_STEM_LAUNCH(__half, 1);
    else if (src == 2) GCV_STEM_LAUNCH(__half, 2);
    else GCV_STEM_LAUNCH(__half, 0);
  }
#undef GCV_STEM_LAUNCH
  return check_launch("stem_fused");
}

int stem_fused(int dtype, int nchw, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
               const float* ln_b, float eps, int B, int H, int W, cudaStream_t stream) {
  return stem_fused_src(dtype, nchw ? 1 : 0, x, y, w, bias, ln_w, ln_b, eps, B, H, W, nullptr, nullptr, stream);
}

}  // namespace gcv
