// 3x3 convolution (pad 1, stride 1 or 2) as a tcgen05 implicit GEMM, NHWC, 16-bit, C in {64, 128}, N <= 256:
// the encoders' wide layers, Conv2d(64 -> 128) and Conv2d(128 -> 256).
// Reference: model/genconvit_ed.py:26-32 (conv + ReLU + MaxPool2d(2, 2)), model/genconvit_vae.py:25-31 (stride-2
// conv + BatchNorm (folded into w, b by the host) + LeakyReLU).
//
// No im2col matrix exists.  An M tile is 128 output pixels of a (columns x rows x images) box -- 16 x 8 x 1, or
// 8 x 8 x 2 for maps no wider than 8 -- and its A operand for tap (kh, kw) and channel block cb is ONE 4-D TMA box of
// the input tensor at (64 cb, s x0 + kw - 1, s y0 + kh - 1, b0): the hardware walks the box with element stride s,
// zero-fills everything outside the image (= the padding) and writes 128 rows x 128 B in the 128B-swizzled K-major
// layout tcgen05.mma reads.  K runs over 9 taps x C/64 blocks; B is the [N][(kh,kw,ci)] weight matrix, one 64-wide
// slice per step.  The whole N (<= 256) is one accumulator, so the input is read once per tap from L2 and the 9x
// matrix (231 MB per 256 ED frames for 64 -> 128) is never written or read.
//
// Epilogue (8 warps, thread = pixel, 32 channels per TMEM load): with pooling, a 16-column tile puts the four pixels
// of a 2x2 window in lanes l, l^1, l^16 of one warp; two exchange rounds leave each of the four lanes with the window
// maximum of 8 of the 32 channels (max first, then bias + activation: both are monotonic, the result is identical)
// and every lane stores 16 bytes.  Two TMEM accumulator stages let the epilogue overlap the next tile's MMAs.
#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace gcv {

namespace {

constexpr int XT_BM = 128, XT_BK = 64;
constexpr int XT_EPI_WARPS = 8;                       // two per TMEM lane quarter: each takes half of the N columns
constexpr int XT_THREADS = 64 + 32 * XT_EPI_WARPS;
constexpr int XT_STAGES = 4;
constexpr int XT_CTRL = 2048;                         // barriers + TMEM slot (first KB), bias [256] fp32 (second KB)
constexpr uint32_t XT_TMEM_COLS = 512;

struct XtParams {
  int B, Ho, Wo;            // conv output size (before pooling)
  int C, N;
  int stride, pool, act;
  int bx, by, bb;           // tile box in output pixels / images; bx * by * bb = 128
  int tiles_x, tiles_y, tiles_b, num_tiles;
  uint32_t idesc;
  const float* bias;
  void* y;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ float xt_act(float v, int act) {
  return act == GCV_ACT_RELU ? fmaxf(v, 0.0f) : (act == GCV_ACT_LEAKY ? (v > 0.0f ? v : 0.01f * v) : v);
}

template <typename T>
__global__ void __launch_bounds__(XT_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const XtParams p) {
  extern __shared__ uint8_t xt_dyn[];
  uint8_t* smem = xt_dyn + ((1024u - (smem_u32(xt_dyn) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty_bar = full_bar + XT_STAGES;
  uint64_t* tmem_full = empty_bar + XT_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + 1024);
  const uint32_t tiles_base = smem_u32(smem) + XT_CTRL;
  const uint32_t a_bytes = XT_BM * XT_BK * 2, b_bytes = (uint32_t)p.N * XT_BK * 2, stage_bytes = a_bytes + b_bytes;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cblocks = p.C / XT_BK, num_kb = 9 * cblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < XT_STAGES; ++s) {
      mbar_init(smem_u32(full_bar + s), 1);
      mbar_init(smem_u32(empty_bar + s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(tmem_full + s), 1);
      mbar_init(smem_u32(tmem_empty + s), XT_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(XT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.N; i += XT_THREADS) bias_s[i] = __ldg(p.bias + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x, rest = tile / p.tiles_x, ty = rest % p.tiles_y, tb = rest / p.tiles_y;
      const int x0 = tx * p.bx * p.stride - 1, y0 = ty * p.by * p.stride - 1, b0 = tb * p.bb;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / cblocks, cb = kb - tap * cblocks;
        const int kh = tap / 3, kw = tap - 3 * kh;
        mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
        if (elect_one()) {
          const uint32_t fb = smem_u32(full_bar + stage);
          const uint32_t sa = tiles_base + stage * stage_bytes;
          mbar_expect_tx(fb, stage_bytes);
          tma_load_4d(sa, &tm_x, fb, cb * XT_BK, x0 + kw, y0 + kh, b0);
          tma_load_2d(sa + a_bytes, &tm_w, fb, tap * p.C + cb * XT_BK, 0);
        }
        __syncwarp();
        if (++stage == XT_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0, as = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(tmem_empty + as), aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.N);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(full_bar + stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = tiles_base + stage * stage_bytes;
          const uint64_t ad0 = umma_desc_kmajor<128>(sa), bd0 = umma_desc_kmajor<128>(sa + a_bytes);
#pragma unroll
          for (int k = 0; k < XT_BK / 16; ++k) tc_mma(d_tmem, ad0 + 2 * k, bd0 + 2 * k, p.idesc, (kb | k) ? 1u : 0u);
          tc_commit(smem_u32(empty_bar + stage));
          if (kb == num_kb - 1) tc_commit(smem_u32(tmem_full + as));
        }
        __syncwarp();
        if (++stage == XT_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
    const int r = quarter * 32 + lane;                       // accumulator row = pixel of the tile box
    const int xx = r % p.bx, yy = (r / p.bx) % p.by, bi = r / (p.bx * p.by);
    const int col_lo = half * (p.N / 2), col_hi = col_lo + p.N / 2;
    const int Hp = p.pool ? p.Ho / 2 : p.Ho, Wp = p.pool ? p.Wo / 2 : p.Wo;
    const bool b0 = (lane & 1) != 0, b1 = (lane & 16) != 0;  // pooling: position inside the 2x2 window
    T* yout = reinterpret_cast<T*>(p.y);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x, rest = tile / p.tiles_x, ty = rest % p.tiles_y, tb = rest / p.tiles_y;
      const int ox = tx * p.bx + xx, oy = ty * p.by + yy, b = tb * p.bb + bi;
      const bool valid = ox < p.Wo && oy < p.Ho && b < p.B;
      const int64_t pix = p.pool ? ((int64_t)b * Hp + (oy >> 1)) * Wp + (ox >> 1) : ((int64_t)b * Hp + oy) * Wp + ox;
      mbar_wait(smem_u32(tmem_full + as), aphase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * p.N);
      for (int n0 = col_lo; n0 < col_hi; n0 += 32) {
        float v[32];
        {
          uint32_t q[32];
          tc_ld32(t_row + n0, q);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(q[e]);
        }
        if (n0 + 32 >= col_hi) {                             // all TMEM reads of this tile by this warp are done
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tmem_empty + as));
        }
        if (p.pool) {
          // round 1 (column neighbour): the even lane keeps channels 0-15, the odd lane 16-31
          float h[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float send = b0 ? v[e] : v[16 + e], keep = b0 ? v[16 + e] : v[e];
            h[e] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 1));
          }
          // round 2 (row neighbour): the upper row's lane keeps the first 8 of those, the lower row's lane the last 8
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float send = b1 ? h[e] : h[8 + e], keep = b1 ? h[8 + e] : h[e];
            o[e] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
          }
          const int n = n0 + (b0 ? 16 : 0) + (b1 ? 8 : 0);
          if (valid) {
            uint4 q;
            q.x = pack2<T>(xt_act(o[0] + bias_s[n], p.act), xt_act(o[1] + bias_s[n + 1], p.act));
            q.y = pack2<T>(xt_act(o[2] + bias_s[n + 2], p.act), xt_act(o[3] + bias_s[n + 3], p.act));
            q.z = pack2<T>(xt_act(o[4] + bias_s[n + 4], p.act), xt_act(o[5] + bias_s[n + 5], p.act));
            q.w = pack2<T>(xt_act(o[6] + bias_s[n + 6], p.act), xt_act(o[7] + bias_s[n + 7], p.act));
            *reinterpret_cast<uint4*>(yout + pix * p.N + n) = q;
          }
        } else if (valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int n = n0 + 8 * j;
            uint4 q;
            q.x = pack2<T>(xt_act(v[8 * j] + bias_s[n], p.act), xt_act(v[8 * j + 1] + bias_s[n + 1], p.act));
            q.y = pack2<T>(xt_act(v[8 * j + 2] + bias_s[n + 2], p.act), xt_act(v[8 * j + 3] + bias_s[n + 3], p.act));
            q.z = pack2<T>(xt_act(v[8 * j + 4] + bias_s[n + 4], p.act), xt_act(v[8 * j + 5] + bias_s[n + 5], p.act));
            q.w = pack2<T>(xt_act(v[8 * j + 6] + bias_s[n + 6], p.act), xt_act(v[8 * j + 7] + bias_s[n + 7], p.act));
            *reinterpret_cast<uint4*>(yout + pix * p.N + n) = q;
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(XT_TMEM_COLS) : "memory");
}

typedef CUresult (*XtEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

XtEncodeFn xt_get_encode() {
  static XtEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<XtEncodeFn>(ptr);
  }
  return fn;
}

}  // namespace

bool conv3x3_tc_supported(int dtype, int C, int N) {
  return (dtype == GCV_BF16 || dtype == GCV_F16) && (C == 64 || C == 128) && N % 64 == 0 && N >= 64 && N <= 256;
}

// y = act(conv3x3(x, pad 1, stride) + bias) [-> 2x2 max-pool], x [B,H,W,C], w [N][(kh,kw,ci)] of `dtype`, bias fp32 [N].
int conv3x3_tc(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
               int H, int W, int C, int N, cudaStream_t stream) {
  GCV_REQUIRE(conv3x3_tc_supported(dtype, C, N), "conv3x3_tc: needs a 16-bit dtype, C in {64,128}, N %% 64 == 0, N <= 256 (C=%d N=%d)", C, N);
  GCV_REQUIRE((stride == 1 || stride == 2) && (!pool || stride == 1), "conv3x3_tc: stride 1 or 2, pool only with stride 1");
  GCV_REQUIRE(B > 0 && H > 0 && W > 0 && x && y && w && bias, "conv3x3_tc: bad shape or null pointer");
  GCV_REQUIRE(act == GCV_ACT_NONE || act == GCV_ACT_RELU || act == GCV_ACT_LEAKY, "conv3x3_tc: activation %d", act);
  GCV_REQUIRE(((uintptr_t)x | (uintptr_t)y | (uintptr_t)w) % 16 == 0, "conv3x3_tc: x, y, w must be 16-byte aligned");
  XtParams p{};
  p.B = B; p.C = C; p.N = N; p.stride = stride; p.pool = pool; p.act = act;
  p.Ho = (H - 1) / stride + 1; p.Wo = (W - 1) / stride + 1;
  GCV_REQUIRE(!pool || (p.Ho % 2 == 0 && p.Wo % 2 == 0), "conv3x3_tc: the fused 2x2 max-pool needs even conv output sizes");
  if (p.Wo <= 8 && !pool) { p.bx = 8; p.by = 8; p.bb = 2; } else { p.bx = 16; p.by = 8; p.bb = 1; }
  p.tiles_x = (p.Wo + p.bx - 1) / p.bx; p.tiles_y = (p.Ho + p.by - 1) / p.by; p.tiles_b = (B + p.bb - 1) / p.bb;
  const int64_t tiles = (int64_t)p.tiles_x * p.tiles_y * p.tiles_b;
  GCV_REQUIRE(tiles < 2147483647LL, "conv3x3_tc: too many tiles");
  p.num_tiles = (int)tiles;
  p.idesc = umma_idesc_f16(dtype == GCV_BF16, XT_BM, N);
  p.bias = bias; p.y = y;

  XtEncodeFn enc = xt_get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not resolvable (no CUDA driver?)");
    return GCV_ERR_NO_DRIVER;
  }
  const CUtensorMapDataType tdt = dtype == GCV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap tm_x, tm_w;
  {
    // to load n elements along a dimension walked with element stride s the box must span n * s elements
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {XT_BK, (cuuint32_t)(p.bx * stride), (cuuint32_t)(p.by * stride), (cuuint32_t)p.bb};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(&tm_x, tdt, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv3x3_tc: cuTensorMapEncodeTiled (input) failed: CUresult %d (B=%d H=%d W=%d C=%d s=%d)", (int)r, B, H, W, C, stride);
      return GCV_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * C, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)9 * C * 2};
    cuuint32_t box[2] = {XT_BK, (cuuint32_t)N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm_w, tdt, 2, const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("conv3x3_tc: cuTensorMapEncodeTiled (weights) failed: CUresult %d (C=%d N=%d)", (int)r, C, N);
      return GCV_ERR_CUDA;
    }
  }
  const int sms = device_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  const size_t smem = 1024 + XT_CTRL + (size_t)XT_STAGES * (XT_BM * XT_BK * 2 + (size_t)N * XT_BK * 2);
  auto launch = [&](auto tag) -> int {
    using T = decltype(tag);
    static unsigned long long attr_devs = 0;
    if (first_on_device(attr_devs))
      cudaFuncSetAttribute(conv3x3_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    conv3x3_tc_kernel<T><<<grid, XT_THREADS, smem, stream>>>(tm_x, tm_w, p);
    return check_launch("conv3x3_tc");
  };
  return dtype == GCV_BF16 ? launch(__nv_bfloat16{}) : launch(__half{});
}

}  // namespace gcv
