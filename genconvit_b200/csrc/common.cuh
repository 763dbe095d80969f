// Shared device/host helpers for the genconvit_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/genconvit_b200.h"

namespace gcv {

void set_error(const char* fmt, ...);
int check_launch(const char* what);
// Per-DEVICE host state (a process may drive several GPUs): the SM count of the current device, and a one-time-setup
// latch per (flag word, current device) for things like cudaFuncSetAttribute, which only applies to one device.
int device_sms();
bool first_on_device(unsigned long long& mask);

#define GCV_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      gcv::set_error(__VA_ARGS__);             \
      return GCV_ERR_BAD_ARG;                  \
    }                                          \
  } while (0)

// ---- element conversion -----------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// pack two floats into one 32-bit word of T pairs (16-bit types only)
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ float2 unpack2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack2<__nv_bfloat16>(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w));
}
template <> __device__ __forceinline__ float2 unpack2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<__half2*>(&w));
}

// load / store 8 consecutive elements (16-byte aligned for 16-bit T, 32-byte for float)
template <typename T> __device__ __forceinline__ void load8(const T* p, float* v) {
  if constexpr (sizeof(T) == 4) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 q = *reinterpret_cast<const uint4*>(p);
    float2 f;
    f = unpack2<T>(q.x); v[0] = f.x; v[1] = f.y;
    f = unpack2<T>(q.y); v[2] = f.x; v[3] = f.y;
    f = unpack2<T>(q.z); v[4] = f.x; v[5] = f.y;
    f = unpack2<T>(q.w); v[6] = f.x; v[7] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void store8(T* p, const float* v) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint4 q;
    q.x = pack2<T>(v[0], v[1]); q.y = pack2<T>(v[2], v[3]);
    q.z = pack2<T>(v[4], v[5]); q.w = pack2<T>(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = q;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- activations ---------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU for the 16-bit tensor-core epilogue, where the erf form is the instruction-issue bottleneck:
//   gelu(x) = x * Phi(x) ~= 0.5*x*(1 + tanh(x*(c1 + c3*x^2)))   with c1, c3 a minimax fit to the erf form
// (max abs error 2.7e-4 over all x, i.e. below half a bf16 ulp of the result; both coefficients are positive
// so no clamp is needed) and the hardware tanh.approx (2^-11 relative).  6 instructions per element.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = x * fmaf(x * x, 3.470094e-2f, 8.0015698e-1f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}

// The same GELU on a pair of values in packed fp16 arithmetic (HMUL2/HFMA2 + one tanh.approx.f16x2 for both):
// 6 instructions per PAIR.  fp16 carries 11 significant bits, comfortably below the 16-bit output rounding;
// |x| > 255 overflows x*x to +inf, which still yields the correct limits (x or 0) through tanh(+-inf) = +-1.
__device__ __forceinline__ __half2 gelu_fast_h2(__half2 x) {
  const __half2 c1 = __float2half2_rn(8.0015698e-1f), c3 = __float2half2_rn(3.470094e-2f), hf = __float2half2_rn(0.5f);
  const __half2 u = __hmul2(x, __hfma2(__hmul2(x, x), c3, c1));
  __half2 t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&t)) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 h = __hmul2(x, hf);
  return __hfma2(h, t, h);
}

// Row statistics of the folded LayerNorm (gcv_epilogue.ln_stats): (rstd, -mean * rstd) of row m from its per-chunk
// partial sums, so that LN-folded pre-activation = acc * rs + rm * colsum[n] + bias[n].
__device__ __forceinline__ float2 ln_row_scale(const float* stats, int64_t m, int chunks, int K, float eps) {
  if (chunks == 0) return __ldg(reinterpret_cast<const float2*>(stats) + m);   // already reduced by gcv_ln_finalize
  const float2* p = reinterpret_cast<const float2*>(stats) + m * chunks;
  float s = 0.0f, q = 0.0f;
  for (int i = 0; i < chunks; ++i) {
    const float2 v = __ldg(p + i);
    s += v.x; q += v.y;
  }
  const float inv = 1.0f / (float)K;
  const float mean = s * inv;
  const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, q * inv), 0.0f) + eps);
  return make_float2(rstd, -mean * rstd);
}

// The same in two halves, so that the loads can be issued a tile ahead of their use (their latency then overlaps work):
// ln_row_load fetches the partial sums (at most 8 chunks: K <= 256), ln_row_finish reduces them.
__device__ __forceinline__ void ln_row_load(const float* stats, int64_t m, int chunks, float2 (&raw)[8]) {
  const float2* p = reinterpret_cast<const float2*>(stats) + m * chunks;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < chunks) raw[i] = __ldg(p + i);
}
__device__ __forceinline__ float2 ln_row_finish(const float2 (&raw)[8], int chunks, int K, float eps) {
  float s = 0.0f, q = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < chunks) { s += raw[i].x; q += raw[i].y; }
  const float inv = 1.0f / (float)K;
  const float mean = s * inv;
  const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, q * inv), 0.0f) + eps);
  return make_float2(rstd, -mean * rstd);
}

// The same GELU from h = x / 2 (the callers fold the 1/2 into the scale / bias they apply anyway, which saves the
// multiply):  gelu(x) = h + h * tanh(h * (2 c1 + 8 c3 h^2)).  5 instructions per pair + the packed tanh.
__device__ __forceinline__ __half2 gelu_from_half_h2(__half2 h) {
  const __half2 c1 = __float2half2_rn(2.0f * 8.0015698e-1f), c3 = __float2half2_rn(8.0f * 3.470094e-2f);
  const __half2 u = __hmul2(h, __hfma2(__hmul2(h, h), c3, c1));
  __half2 t;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(*reinterpret_cast<uint32_t*>(&t)) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  return __hfma2(h, t, h);
}

// fp32 pair -> packed fp16 with SATURATION to +-65504 (one F2FP.SATFINITE): an overflowing pre-activation must not become
// -inf, for which gelu_from_half_h2 would evaluate (-inf) * (-1) + (-inf) = NaN instead of 0
__device__ __forceinline__ __half2 f2h2_sat(const float2 v) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
  return *reinterpret_cast<__half2*>(&r);
}

// x0..x3 hold HALF the pre-activations
template <typename T>
__device__ __forceinline__ uint4 gelu_pack8_h2(const __half2 x0, const __half2 x1, const __half2 x2, const __half2 x3) {
  const __half2 g0 = gelu_from_half_h2(x0), g1 = gelu_from_half_h2(x1), g2 = gelu_from_half_h2(x2), g3 = gelu_from_half_h2(x3);
  uint4 q;
  if constexpr (sizeof(T) == 2 && !std::is_same<T, __half>::value) {
    const float2 f0 = __half22float2(g0), f1 = __half22float2(g1), f2 = __half22float2(g2), f3 = __half22float2(g3);
    q.x = pack2<T>(f0.x, f0.y); q.y = pack2<T>(f1.x, f1.y); q.z = pack2<T>(f2.x, f2.y); q.w = pack2<T>(f3.x, f3.y);
  } else {
    q.x = *reinterpret_cast<const uint32_t*>(&g0); q.y = *reinterpret_cast<const uint32_t*>(&g1);
    q.z = *reinterpret_cast<const uint32_t*>(&g2); q.w = *reinterpret_cast<const uint32_t*>(&g3);
  }
  return q;
}

// d = a * b + c on two fp32 lanes with one FFMA2 (fma.rn.f32x2: same rounding as two fmaf, half the issue slots)
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// folded LayerNorm + bias + GELU on 8 fp32 accumulators: x = w * rs + (rm * colsum + bias), two columns per FFMA2.
// The caller passes rs = rstd / 2, rm = -mean * rstd / 2 and b = bias / 2 (GELU is evaluated from x / 2).
template <typename T>
__device__ __forceinline__ uint4 ln_bias_gelu_pack8(const float* w, const float rs, const float rm, const float4 s0,
                                                    const float4 s1, const float4 b0, const float4 b1) {
  const float2 rs2 = make_float2(rs, rs), rm2 = make_float2(rm, rm);
  const float2 x0 = fma2(make_float2(w[0], w[1]), rs2, fma2(rm2, make_float2(s0.x, s0.y), make_float2(b0.x, b0.y)));
  const float2 x1 = fma2(make_float2(w[2], w[3]), rs2, fma2(rm2, make_float2(s0.z, s0.w), make_float2(b0.z, b0.w)));
  const float2 x2 = fma2(make_float2(w[4], w[5]), rs2, fma2(rm2, make_float2(s1.x, s1.y), make_float2(b1.x, b1.y)));
  const float2 x3 = fma2(make_float2(w[6], w[7]), rs2, fma2(rm2, make_float2(s1.z, s1.w), make_float2(b1.z, b1.w)));
  return gelu_pack8_h2<T>(f2h2_sat(x0), f2h2_sat(x1), f2h2_sat(x2), f2h2_sat(x3));
}

// bias + GELU on 8 fp32 accumulators -> 8 packed 16-bit outputs (T = __half or __nv_bfloat16).  The caller passes
// b = bias / 2: x / 2 = w * 0.5 + b is one FFMA2 per column pair.
template <typename T>
__device__ __forceinline__ uint4 bias_gelu_pack8(const float* w, const float4 b0, const float4 b1) {
  const float2 hf = make_float2(0.5f, 0.5f);
  const float2 x0 = fma2(make_float2(w[0], w[1]), hf, make_float2(b0.x, b0.y));
  const float2 x1 = fma2(make_float2(w[2], w[3]), hf, make_float2(b0.z, b0.w));
  const float2 x2 = fma2(make_float2(w[4], w[5]), hf, make_float2(b1.x, b1.y));
  const float2 x3 = fma2(make_float2(w[6], w[7]), hf, make_float2(b1.z, b1.w));
  return gelu_pack8_h2<T>(f2h2_sat(x0), f2h2_sat(x1), f2h2_sat(x2), f2h2_sat(x3));
}

__device__ __forceinline__ float apply_act_fast(float v, int act) {
  switch (act) {
    case GCV_ACT_GELU: return gelu_fast(v);
    case GCV_ACT_RELU: return fmaxf(v, 0.0f);
    case GCV_ACT_LEAKY: return v > 0.0f ? v : 0.01f * v;
    default: return v;
  }
}

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case GCV_ACT_GELU: return gelu_erf(v);
    case GCV_ACT_RELU: return fmaxf(v, 0.0f);
    case GCV_ACT_LEAKY: return v > 0.0f ? v : 0.01f * v;
    default: return v;
  }
}

// ---- GEMM epilogue (shared by the tcgen05 and the SIMT back ends) -----------------
// One output element; see gcv_epilogue in the public header for the order of operations.
template <typename T>
__device__ __forceinline__ void epilogue_one(const gcv_epilogue& ep, int64_t m, int n, int N, float acc, void* D,
                                             int K = 0) {
  if (ep.ln_stats) {
    const float2 rs = ln_row_scale(ep.ln_stats, m, ep.ln_chunks, K, ep.ln_eps);
    acc = fmaf(acc, rs.x, rs.y * __ldg(ep.ln_colsum + n));
  }
  float v = acc + (ep.bias ? __ldg(ep.bias + n) : 0.0f);
  v = apply_act(v, ep.act);
  if (ep.eps) {
    const float mu = v;
    const int hw = n / ep.eps_c, c = n - hw * ep.eps_c;
    const float e = __ldg(ep.eps + m * (int64_t)N + (int64_t)c * ep.eps_hw + hw);
    if (ep.mu_out) ep.mu_out[m * (int64_t)N + n] = mu;
    v = e * expf(0.5f * mu) + mu;
  }
  if (ep.gamma) v *= __ldg(ep.gamma + n);
  if (ep.residual) v += to_f<T>(reinterpret_cast<const T*>(ep.residual)[m * ep.ldr + n]);
  int64_t off;
  if (ep.store == GCV_STORE_PIXEL_SHUFFLE2) {
    const int ij = n / ep.ps_co, co = n - ij * ep.ps_co;
    const int64_t hw = (int64_t)ep.ps_h * ep.ps_w;
    const int64_t b = m / hw;
    const int r = (int)(m - b * hw);
    const int h = r / ep.ps_w, w = r - h * ep.ps_w;
    off = ((b * (2 * ep.ps_h) + 2 * h + (ij >> 1)) * (int64_t)(2 * ep.ps_w) + 2 * w + (ij & 1)) * ep.ps_co + co;
  } else {
    off = m * ep.ldd + n;
  }
  if (ep.out_f32) reinterpret_cast<float*>(D)[off] = v;
  else reinterpret_cast<T*>(D)[off] = from_f<T>(v);
}

}  // namespace gcv
