// Swin-T kernels that are not plain GEMMs / row LayerNorms: (shifted-)window multi-head attention with
// relative-position bias, patch merging gather, token mean.
// timm==0.6.5 swin_tiny_patch4_window7_224 as constructed by reference model/genconvit_ed.py:69 and
// model/genconvit_vae.py:96 (`self.embedder`); arithmetic restated in oracle/backbones.py
// (swin_window_attention / swin_block / swin_patch_merging), which is pinned against torchvision.
//
// Window attention, one CTA per (image, window, head), 4 warps:
//   tokens of the window are gathered straight from the [B*res*res, 3C] qkv matrix -- the cyclic shift
//   (torch.roll by -shift), window_partition, window_reverse and the roll back are pure index arithmetic
//   on the token row: rolled (y', x') <- original ((y' + shift) mod res, (x' + shift) mod res), and the
//   result is written back to the same original row;
//   S = (q * hd^-0.5) k^T + bias_table[rel(i, j)][head] + mask(i, j), softmax over j, O = P v;
//   the shifted-window mask ({0, -100}, timm's registered `attn_mask` buffer) is a fixed function of
//   (res, window, shift) and is recomputed from the region ids of the two tokens.
// 16-bit path: warp = 16 query rows; S and O live in mma.sync m16n8k16 fragments (fp32 accumulate), P is
// re-used from the S accumulator registers as the A operand of the second MMA (flash-attention style).
// fp32 path (exact mode): thread = query row, fp32 FMAs.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"

namespace gcv {

namespace {

constexpr int WS = 7, WT = 49, HD = 32;       // window side, tokens per window, head dim (Swin-T: C / heads = 32 everywhere)

__device__ __forceinline__ int region_id(int r, int res, int shift) { return r < res - WS ? 0 : (r < res - shift ? 1 : 2); }

// token row (in [0, res*res)) of window-local token i of window (wy, wx), after undoing the cyclic shift
__device__ __forceinline__ int token_row(int i, int wy, int wx, int res, int shift) {
  const int yr = wy * WS + i / WS, xr = wx * WS + i % WS;
  int y = yr + shift, x = xr + shift;
  if (y >= res) y -= res;
  if (x >= res) x -= res;
  return y * res + x;
}

// additive logit term for (query i, key j): relative-position bias + shifted-window mask
__device__ __forceinline__ float bias_mask(const float* __restrict__ tab, int heads, int head, int i, int j, int wy, int wx,
                                           int res, int shift) {
  const int yi = i / WS, xi = i % WS, yj = j / WS, xj = j % WS;
  float v = __ldg(tab + ((yi - yj + WS - 1) * (2 * WS - 1) + (xi - xj + WS - 1)) * heads + head);
  if (shift > 0) {
    const int ri = region_id(wy * WS + yi, res, shift) * 3 + region_id(wx * WS + xi, res, shift);
    const int rj = region_id(wy * WS + yj, res, shift) * 3 + region_id(wx * WS + xj, res, shift);
    if (ri != rj) v -= 100.0f;
  }
  return v;
}

// ---------------------------------------------------------------------------------
// exact path (any T, used for fp32): 64 threads, thread = query row
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(64)
swin_attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, const float* __restrict__ tab, int res, int C,
                      int heads, int shift) {
  __shared__ float ks[WT][HD + 1], vs[WT][HD + 1];
  const int nw = res / WS;
  const int head = blockIdx.x % heads;
  const int win = (blockIdx.x / heads) % (nw * nw);
  const int b = blockIdx.x / (heads * nw * nw);
  const int wy = win / nw, wx = win % nw;
  const int64_t base = (int64_t)b * res * res;
  for (int e = threadIdx.x; e < WT * HD; e += 64) {
    const int j = e / HD, d = e % HD;
    const T* row = qkv + (base + token_row(j, wy, wx, res, shift)) * 3 * C + head * HD + d;
    ks[j][d] = to_f<T>(row[C]);
    vs[j][d] = to_f<T>(row[2 * C]);
  }
  __syncthreads();
  const int i = threadIdx.x;
  if (i >= WT) return;
  const int64_t my_row = base + token_row(i, wy, wx, res, shift);
  float q[HD];
  const float scale = rsqrtf((float)HD);
#pragma unroll
  for (int d = 0; d < HD; ++d) q[d] = to_f<T>(qkv[my_row * 3 * C + head * HD + d]) * scale;
  float s[WT];
  float mx = -3.0e38f;
#pragma unroll 7
  for (int j = 0; j < WT; ++j) {
    float a = 0.0f;
#pragma unroll
    for (int d = 0; d < HD; ++d) a = fmaf(q[d], ks[j][d], a);
    a += bias_mask(tab, heads, head, i, j, wy, wx, res, shift);
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.0f;
#pragma unroll 7
  for (int j = 0; j < WT; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.0f / sum;
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.0f;
#pragma unroll 7
  for (int j = 0; j < WT; ++j) {
    const float p = s[j] * inv;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = fmaf(p, vs[j][d], o[d]);
  }
#pragma unroll
  for (int d = 0; d < HD; ++d) out[my_row * C + head * HD + d] = from_f<T>(o[d]);
}

// ---------------------------------------------------------------------------------
// tensor-core path (bf16 / fp16): 128 threads, warp w = query rows 16w .. 16w+15 (rows >= 49 are padding)
// ---------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  if constexpr (std::is_same<T, __half>::value)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int AT_ROWS = 64;                   // 49 tokens padded to 4 MMA row tiles / 8 key tiles
constexpr int AT_LD = HD + 8;                 // padded row (80 B): conflict-free 32-bit fragment reads

template <typename T>
__global__ void __launch_bounds__(128)
swin_attn_mma_kernel(const T* __restrict__ qkv, T* __restrict__ out, const float* __restrict__ tab, int res, int C,
                     int heads, int shift) {
  __shared__ __align__(16) T qs[AT_ROWS][AT_LD], ks[AT_ROWS][AT_LD];
  __shared__ __align__(16) T vt[HD][AT_ROWS + 8];         // V transposed: [d][token], so PV's B operand pairs tokens
  __shared__ int rows[AT_ROWS];
  const int nw = res / WS;
  const int head = blockIdx.x % heads;
  const int win = (blockIdx.x / heads) % (nw * nw);
  const int b = blockIdx.x / (heads * nw * nw);
  const int wy = win / nw, wx = win % nw;
  const int64_t base = (int64_t)b * res * res;
  if (threadIdx.x < AT_ROWS) rows[threadIdx.x] = threadIdx.x < WT ? token_row(threadIdx.x, wy, wx, res, shift) : -1;
  __syncthreads();
  // gather: 64 rows x (q, k, v) x 4 pieces of 16 bytes; padding rows are zero
  for (int e = threadIdx.x; e < AT_ROWS * 12; e += 128) {
    const int j = e / 12, part = (e % 12) / 4, piece = e % 4;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (rows[j] >= 0)
      val = *reinterpret_cast<const uint4*>(qkv + (base + rows[j]) * 3 * C + part * C + head * HD + piece * 8);
    if (part == 0) *reinterpret_cast<uint4*>(&qs[j][piece * 8]) = val;
    else if (part == 1) *reinterpret_cast<uint4*>(&ks[j][piece * 8]) = val;
    else {
      const T* pv = reinterpret_cast<const T*>(&val);
#pragma unroll
      for (int d = 0; d < 8; ++d) vt[piece * 8 + d][j] = pv[d];
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int r0 = warp * 16 + g, r1 = r0 + 8;          // this lane's two query rows
  // ---- S = Q K^T: A = Q rows (k = head dim), B[k][n] = K[n][k] ----
  float sacc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.0f;
#pragma unroll
  for (int kt = 0; kt < 2; ++kt) {
    const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&qs[r0][kt * 16 + 2 * t]);
    const uint32_t a1 = *reinterpret_cast<const uint32_t*>(&qs[r1][kt * 16 + 2 * t]);
    const uint32_t a2 = *reinterpret_cast<const uint32_t*>(&qs[r0][kt * 16 + 2 * t + 8]);
    const uint32_t a3 = *reinterpret_cast<const uint32_t*>(&qs[r1][kt * 16 + 2 * t + 8]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&ks[nt * 8 + g][kt * 16 + 2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&ks[nt * 8 + g][kt * 16 + 2 * t + 8]);
      mma_16816<T>(sacc[nt], a0, a1, a2, a3, b0, b1);
    }
  }
  // ---- scale, bias, mask, softmax over the 49 valid keys (fragment: rows r0 / r1, columns nt*8 + 2t + {0,1}) ----
  const float scale = rsqrtf((float)HD);
  float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t + e;
      if (j < WT) {
        if (r0 < WT) sacc[nt][e] = fmaf(sacc[nt][e], scale, bias_mask(tab, heads, head, r0, j, wy, wx, res, shift));
        if (r1 < WT) sacc[nt][2 + e] = fmaf(sacc[nt][2 + e], scale, bias_mask(tab, heads, head, r1, j, wy, wx, res, shift));
        mx0 = fmaxf(mx0, sacc[nt][e]);
        mx1 = fmaxf(mx1, sacc[nt][2 + e]);
      }
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = nt * 8 + 2 * t + e;
      const float p0 = j < WT ? __expf(sacc[nt][e] - mx0) : 0.0f;
      const float p1 = j < WT ? __expf(sacc[nt][2 + e] - mx1) : 0.0f;
      sacc[nt][e] = p0; sacc[nt][2 + e] = p1;
      sum0 += p0; sum1 += p1;
    }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  // ---- O = P V: the S accumulator fragment of key tiles (2kt, 2kt+1) is exactly the A fragment of k-step kt ----
  float oacc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.0f;
#pragma unroll
  for (int kt = 0; kt < 4; ++kt) {
    const uint32_t a0 = pack2<T>(sacc[2 * kt][0] * inv0, sacc[2 * kt][1] * inv0);
    const uint32_t a1 = pack2<T>(sacc[2 * kt][2] * inv1, sacc[2 * kt][3] * inv1);
    const uint32_t a2 = pack2<T>(sacc[2 * kt + 1][0] * inv0, sacc[2 * kt + 1][1] * inv0);
    const uint32_t a3 = pack2<T>(sacc[2 * kt + 1][2] * inv1, sacc[2 * kt + 1][3] * inv1);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      // B[k = token][n = d] = V[token][d] = vt[d][token]: consecutive tokens are adjacent in vt
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&vt[nt * 8 + g][kt * 16 + 2 * t]);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&vt[nt * 8 + g][kt * 16 + 2 * t + 8]);
      mma_16816<T>(oacc[nt], a0, a1, a2, a3, b0, b1);
    }
  }
  // ---- store: rows r0 / r1, columns nt*8 + 2t + {0,1} of this head ----
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    if (r0 < WT)
      *reinterpret_cast<uint32_t*>(out + (base + rows[r0]) * C + head * HD + nt * 8 + 2 * t) = pack2<T>(oacc[nt][0], oacc[nt][1]);
    if (r1 < WT)
      *reinterpret_cast<uint32_t*>(out + (base + rows[r1]) * C + head * HD + nt * 8 + 2 * t) = pack2<T>(oacc[nt][2], oacc[nt][3]);
  }
}

// Patch merging gather (timm PatchMerging.forward before norm/reduction): x [B,res,res,C] ->
// out [B,res/2,res/2,4C] = cat(x[0::2,0::2], x[1::2,0::2], x[0::2,1::2], x[1::2,1::2]); 16-byte pieces.
template <typename T>
__global__ void __launch_bounds__(256)
swin_merge_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int res, int C) {
  constexpr int V = 16 / (int)sizeof(T);
  const int pieces = C / V, r2 = res / 2;
  const int64_t total = (int64_t)B * r2 * r2 * 4 * pieces;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int piece = (int)(e % pieces);
    const int quad = (int)((e / pieces) % 4);
    const int64_t tok = e / (4 * pieces);
    const int xo = (int)(tok % r2), yo = (int)((tok / r2) % r2);
    const int64_t b = tok / ((int64_t)r2 * r2);
    const int dy = quad & 1, dx = quad >> 1;                       // order (0,0), (1,0), (0,1), (1,1)
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((b * res + 2 * yo + dy) * res + 2 * xo + dx) * C + piece * V);
    *reinterpret_cast<uint4*>(out + tok * 4 * C + quad * C + piece * V) = v;
  }
}

// mean over the L tokens of each image: x [B, L, C] -> y [B, C] (fp32 accumulate)
template <typename T>
__global__ void __launch_bounds__(256)
mean_tokens_kernel(const T* __restrict__ x, T* __restrict__ y, int L, int C) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    float a = 0.0f;
    for (int l = 0; l < L; ++l) a += to_f<T>(x[((int64_t)b * L + l) * C + c]);
    y[(int64_t)b * C + c] = from_f<T>(a / (float)L);
  }
}

template <typename F>
int sw_dispatch(int dtype, F&& f) {
  switch (dtype) {
    case GCV_F32: return f(float{});
    case GCV_BF16: return f(__nv_bfloat16{});
    case GCV_F16: return f(__half{});
    default: set_error("bad dtype %d", dtype); return GCV_ERR_BAD_ARG;
  }
}

}  // namespace

int swin_window_attention(int dtype, const void* qkv, void* out, const float* bias_table, int B, int res, int C, int heads,
                          int shift, cudaStream_t stream) {
  GCV_REQUIRE(B > 0 && res > 0 && res % WS == 0 && heads > 0 && C == heads * HD && shift >= 0 && shift < WS,
              "swin_window_attention: needs res %% 7 == 0 and C == 32 * heads (res=%d C=%d heads=%d shift=%d)", res, C,
              heads, shift);
  GCV_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "swin_window_attention: pointers must be 16-byte aligned");
  const int64_t grid = (int64_t)B * (res / WS) * (res / WS) * heads;
  GCV_REQUIRE(grid < 2147483647LL, "swin_window_attention: grid too large");
  return sw_dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if constexpr (sizeof(T) == 2)
      swin_attn_mma_kernel<T><<<(unsigned)grid, 128, 0, stream>>>(reinterpret_cast<const T*>(qkv), reinterpret_cast<T*>(out),
                                                                 bias_table, res, C, heads, shift);
    else
      swin_attn_simt_kernel<T><<<(unsigned)grid, 64, 0, stream>>>(reinterpret_cast<const T*>(qkv), reinterpret_cast<T*>(out),
                                                                 bias_table, res, C, heads, shift);
    return check_launch("swin_window_attention");
  });
}

int swin_patch_merge(int dtype, const void* x, void* out, int B, int res, int C, cudaStream_t stream) {
  GCV_REQUIRE(B > 0 && res > 0 && res % 2 == 0 && C % 8 == 0, "swin_patch_merge: needs even res and C %% 8 == 0");
  return sw_dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    const int64_t total = (int64_t)B * (res / 2) * (res / 2) * 4 * (C / (16 / (int)sizeof(T)));
    const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    swin_merge_kernel<T><<<grid, 256, 0, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(out), B, res, C);
    return check_launch("swin_patch_merge");
  });
}

int mean_tokens(int dtype, const void* x, void* y, int B, int L, int C, cudaStream_t stream) {
  GCV_REQUIRE(B > 0 && L > 0 && C > 0, "mean_tokens: bad shape");
  return sw_dispatch(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    mean_tokens_kernel<T><<<B, 256, 0, stream>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), L, C);
    return check_launch("mean_tokens");
  });
}

}  // namespace gcv
