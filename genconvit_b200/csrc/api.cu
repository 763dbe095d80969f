// extern "C" surface of libgenconvit_b200.so (declared in include/genconvit_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace gcv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return GCV_ERR_CUDA;
  }
  return GCV_OK;
}

int device_sms() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  const int slot = dev & 63;
  if (!sms[slot]) cudaDeviceGetAttribute(&sms[slot], cudaDevAttrMultiProcessorCount, dev);
  return sms[slot];
}

bool first_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) return true;                       // beyond the latch: redo the (idempotent) setup every time
  const unsigned long long bit = 1ull << dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

bool tcgen05_eligible(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, int64_t K);
int gemm_tcgen05(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t M, int64_t N,
                 int64_t K, const gcv_epilogue* ep, int force_block_n, cudaStream_t stream);
int gemm_simt(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t M, int64_t N,
              int64_t K, const gcv_epilogue* ep, cudaStream_t stream);
int dwconv7_ln(int dtype, const void* x, void* y, const float* taps, const float* bias, const float* ln_w,
               const float* ln_b, float eps, int B, int H, int W, int C, cudaStream_t stream);
int ln_patchify2(int dtype, const void* x, void* a, const float* w, const float* b, float eps, int B, int H, int W, int C,
                 cudaStream_t stream);
int stem_patchify(int dtype, bool nchw, const void* x, void* a, int B, int H, int W, cudaStream_t stream);
int layernorm_rows(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int64_t rows, int C,
                   cudaStream_t stream);
int pool_ln(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int B, int HW, int C,
            cudaStream_t stream);
int conv3x3_first(int dtype, const float* x, void* y, const float* w, const float* b, int stride, int act, int pool,
                  int B, int H, int W, cudaStream_t stream);
int im2col3x3(int dtype, const void* x, void* a, int B, int H, int W, int C, int stride, cudaStream_t stream);
int maxpool2(int dtype, const void* x, void* y, int B, int H, int W, int C, cudaStream_t stream);
int conv3x3_c32(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
                int H, int W, cudaStream_t stream);
int ln_finalize(const float* stats, float* out, int64_t M, int chunks, int K, float eps, cudaStream_t stream);
int stem_fused_src(int dtype, int src, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
                   const float* ln_b, float eps, int B, int H, int W, const float* mean3, const float* std3,
                   cudaStream_t stream);
int conv3x3_first_src(int dtype, int u8, const void* x, void* y, const float* w, const float* b, int stride, int act,
                      int pool, int B, int H, int W, const float* mean3, const float* std3, cudaStream_t stream);
int stem_fused(int dtype, int nchw, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
               const float* ln_b, float eps, int B, int H, int W, cudaStream_t stream);
int conv3x3_c16(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
                int H, int W, cudaStream_t stream);
int resize2x_to_nchw(int dtype, const void* x, float* y, int B, int H, int W, int C, cudaStream_t stream);
int nhwc_to_nchw_f32(int dtype, const void* x, float* y, int B, int H, int W, int C, cudaStream_t stream);
int score_videos(const float* logits, int n_nets, int n_frames, int fpv, float* mean_out, int32_t* cls_out,
                 float* val_out, cudaStream_t stream);
int score_videos_pair(const float* la, const float* lb, int n_frames, int fpv, float* out, cudaStream_t stream);
int convt2x2_small(int dtype, const void* x, void* y, const float* w, const float* bias, int act, int B, int H, int W,
                   int CI, int CO, cudaStream_t stream);
bool conv3x3_tc_supported(int dtype, int C, int N);
int conv3x3_tc(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool, int B,
               int H, int W, int C, int N, cudaStream_t stream);
int convt2x2_mma(int dtype, const void* x, void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                 int act, int B, int H, int W, int CI, int tail, cudaStream_t stream);
int swin_window_attention(int dtype, const void* qkv, void* out, const float* bias_table, int B, int res, int C, int heads,
                          int shift, cudaStream_t stream);
int swin_patch_merge(int dtype, const void* x, void* out, int B, int res, int C, cudaStream_t stream);
int mean_tokens(int dtype, const void* x, void* y, int B, int L, int C, cudaStream_t stream);
int preprocess_frames(const uint8_t* x, float* y, int N, int H, int W, const float* mean3, const float* std3,
                      cudaStream_t stream);
bool mlp_fused_supported(int dtype, int C);
int mlp_fused_trace(long long* out64);
int mlp_fused(int dtype, const void* y, const float* ln_stats, float ln_eps, const void* w1, const float* b1,
              const float* colsum1, const void* w2, const float* b2, const float* gamma, void* x, int64_t M, int C,
              cudaStream_t stream);
bool dwconv7_mma_supported(int dtype, int C);
int dwconv7_mma(int dtype, const void* x, void* y, float* stats, const float* taps, const float* bias, int B, int H,
                int W, int C, cudaStream_t stream);

}  // namespace gcv

using namespace gcv;
#define S(stream) reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int gcv_abi_version(void) { return GCV_ABI_VERSION; }
const char* gcv_last_error(void) { return g_err; }

int gcv_device_supported(int device) {
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return major == 10 ? 1 : 0;
}

int gcv_gemm(int backend, int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t M,
             int64_t N, int64_t K, const gcv_epilogue* ep, void* stream) {
  if (!A || !B || !D || !ep) {
    set_error("gcv_gemm: null pointer");
    return GCV_ERR_BAD_ARG;
  }
  int force_bn = 0;
  if (backend >= 1000) {              // test hook: 1000 + block_n forces the tcgen05 tile width
    force_bn = backend - 1000;
    backend = GCV_GEMM_TCGEN05;
  }
  if (backend == GCV_GEMM_AUTO) backend = tcgen05_eligible(dtype, A, lda, B, ldb, K) ? GCV_GEMM_TCGEN05 : GCV_GEMM_SIMT;
  if (backend == GCV_GEMM_TCGEN05) {
    if (dtype == GCV_F32) {
      set_error("gcv_gemm: the tcgen05 back end takes bf16/fp16 operands");
      return GCV_ERR_UNSUPPORTED;
    }
    return gemm_tcgen05(dtype, A, lda, B, ldb, D, M, N, K, ep, force_bn, S(stream));
  }
  if (backend == GCV_GEMM_SIMT) return gemm_simt(dtype, A, lda, B, ldb, D, M, N, K, ep, S(stream));
  set_error("gcv_gemm: unknown backend %d", backend);
  return GCV_ERR_BAD_ARG;
}

// debug builds only (-DGCV_FUSED_TRACE): copies 64 cycle counters of the last fused-MLP launch; returns 0 otherwise
int gcv_debug_fused_trace(long long* out64) { return mlp_fused_trace(out64); }
int gcv_mlp_fused_supported(int dtype, int C) { return mlp_fused_supported(dtype, C) ? 1 : 0; }
int gcv_mlp_fused(int dtype, const void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                  const float* gamma, void* x, int64_t M, int C, void* stream) {
  return mlp_fused(dtype, y, nullptr, 0.0f, w1, b1, nullptr, w2, b2, gamma, x, M, C, S(stream));
}
int gcv_mlp_fused_ln(int dtype, const void* y, const float* ln_stats, float ln_eps, const void* w1, const float* b1,
                     const float* colsum1, const void* w2, const float* b2, const float* gamma, void* x, int64_t M,
                     int C, void* stream) {
  if (!ln_stats) {
    set_error("gcv_mlp_fused_ln: ln_stats is NULL");
    return GCV_ERR_BAD_ARG;
  }
  return mlp_fused(dtype, y, ln_stats, ln_eps, w1, b1, colsum1, w2, b2, gamma, x, M, C, S(stream));
}

int gcv_dwconv7_ln(int dtype, const void* x, void* y, const float* taps, const float* bias, const float* ln_w,
                   const float* ln_b, float eps, int B, int H, int W, int C, void* stream) {
  return dwconv7_ln(dtype, x, y, taps, bias, ln_w, ln_b, eps, B, H, W, C, S(stream));
}
int gcv_dwconv7_stats(int dtype, const void* x, void* y, float* stats, const float* taps, const float* bias, int B,
                      int H, int W, int C, void* stream) {
  if (!dwconv7_mma_supported(dtype, C)) {
    set_error("gcv_dwconv7_stats: bf16/fp16 and C %% 32 == 0 only (dtype=%d C=%d)", dtype, C);
    return GCV_ERR_UNSUPPORTED;
  }
  if (B <= 0 || H <= 0 || W <= 0) {
    set_error("gcv_dwconv7_stats: bad shape B=%d H=%d W=%d", B, H, W);
    return GCV_ERR_BAD_ARG;
  }
  return dwconv7_mma(dtype, x, y, stats, taps, bias, B, H, W, C, S(stream));
}
int gcv_ln_patchify2(int dtype, const void* x, void* a, const float* ln_w, const float* ln_b, float eps, int B, int H,
                     int W, int C, void* stream) {
  return ln_patchify2(dtype, x, a, ln_w, ln_b, eps, B, H, W, C, S(stream));
}
int gcv_stem_patchify_nchw(int dtype, const float* x, void* a, int B, int H, int W, void* stream) {
  return stem_patchify(dtype, true, x, a, B, H, W, S(stream));
}
int gcv_stem_patchify_nhwc(int dtype, const void* x, void* a, int B, int H, int W, void* stream) {
  return stem_patchify(dtype, false, x, a, B, H, W, S(stream));
}
int gcv_layernorm_rows(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int64_t rows, int C,
                       void* stream) {
  return layernorm_rows(dtype, x, y, w, b, eps, rows, C, S(stream));
}
int gcv_pool_ln(int dtype, const void* x, void* y, const float* w, const float* b, float eps, int B, int HW, int C,
                void* stream) {
  return pool_ln(dtype, x, y, w, b, eps, B, HW, C, S(stream));
}
int gcv_conv3x3_first(int dtype, const float* x, void* y, const float* w, const float* b, int stride, int act, int pool,
                      int B, int H, int W, void* stream) {
  return conv3x3_first(dtype, x, y, w, b, stride, act, pool, B, H, W, S(stream));
}
int gcv_im2col3x3(int dtype, const void* x, void* a, int B, int H, int W, int C, int stride, void* stream) {
  return im2col3x3(dtype, x, a, B, H, W, C, stride, S(stream));
}
int gcv_conv3x3_c16(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool,
                    int B, int H, int W, void* stream) {
  return conv3x3_c16(dtype, x, y, w, bias, stride, act, pool, B, H, W, S(stream));
}
int gcv_stem_fused_u8(int dtype, const uint8_t* x, void* y, const void* w, const float* bias, const float* ln_w,
                      const float* ln_b, float eps, int B, int H, int W, const float* mean3, const float* std3, void* stream) {
  return stem_fused_src(dtype, 2, x, y, w, bias, ln_w, ln_b, eps, B, H, W, mean3, std3, S(stream));
}
int gcv_conv3x3_first_u8(int dtype, const uint8_t* x, void* y, const float* w, const float* b, int stride, int act, int pool,
                         int B, int H, int W, const float* mean3, const float* std3, void* stream) {
  return conv3x3_first_src(dtype, 1, x, y, w, b, stride, act, pool, B, H, W, mean3, std3, S(stream));
}
int gcv_stem_fused(int dtype, int nchw, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
                   const float* ln_b, float eps, int B, int H, int W, void* stream) {
  return stem_fused(dtype, nchw, x, y, w, bias, ln_w, ln_b, eps, B, H, W, S(stream));
}
int gcv_conv3x3_c32(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool,
                    int B, int H, int W, void* stream) {
  return conv3x3_c32(dtype, x, y, w, bias, stride, act, pool, B, H, W, S(stream));
}
int gcv_ln_finalize(const float* stats, float* out, int64_t M, int chunks, int K, float eps, void* stream) {
  return ln_finalize(stats, out, M, chunks, K, eps, S(stream));
}
int gcv_maxpool2(int dtype, const void* x, void* y, int B, int H, int W, int C, void* stream) {
  return maxpool2(dtype, x, y, B, H, W, C, S(stream));
}
int gcv_resize2x_to_nchw(int dtype, const void* x, float* y, int B, int H, int W, int C, void* stream) {
  return resize2x_to_nchw(dtype, x, y, B, H, W, C, S(stream));
}
int gcv_nhwc_to_nchw_f32(int dtype, const void* x, float* y, int B, int H, int W, int C, void* stream) {
  return nhwc_to_nchw_f32(dtype, x, y, B, H, W, C, S(stream));
}
int gcv_convt2x2_small(int dtype, const void* x, void* y, const float* w, const float* bias, int act, int B, int H, int W,
                       int CI, int CO, void* stream) {
  return convt2x2_small(dtype, x, y, w, bias, act, B, H, W, CI, CO, S(stream));
}
int gcv_conv3x3_tc_supported(int dtype, int C, int N) { return conv3x3_tc_supported(dtype, C, N) ? 1 : 0; }
int gcv_conv3x3_tc(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool,
                   int B, int H, int W, int C, int N, void* stream) {
  return conv3x3_tc(dtype, x, y, w, bias, stride, act, pool, B, H, W, C, N, S(stream));
}
int gcv_convt2x2_mma(int dtype, const void* x, void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                     int act, int B, int H, int W, int CI, int tail, void* stream) {
  return convt2x2_mma(dtype, x, y, w1, b1, w2, b2, act, B, H, W, CI, tail, S(stream));
}
int gcv_swin_window_attention(int dtype, const void* qkv, void* out, const float* bias_table, int B, int res, int C,
                              int heads, int shift, void* stream) {
  return swin_window_attention(dtype, qkv, out, bias_table, B, res, C, heads, shift, S(stream));
}
int gcv_swin_patch_merge(int dtype, const void* x, void* out, int B, int res, int C, void* stream) {
  return swin_patch_merge(dtype, x, out, B, res, C, S(stream));
}
int gcv_mean_tokens(int dtype, const void* x, void* y, int B, int L, int C, void* stream) {
  return mean_tokens(dtype, x, y, B, L, C, S(stream));
}
int gcv_preprocess_frames(const uint8_t* x, float* y, int N, int H, int W, const float* mean3, const float* std3,
                          void* stream) {
  return preprocess_frames(x, y, N, H, W, mean3, std3, S(stream));
}

int gcv_score_videos(const float* logits, int n_nets, int n_frames, int frames_per_video, float* mean_out,
                     int32_t* cls_out, float* val_out, void* stream) {
  return score_videos(logits, n_nets, n_frames, frames_per_video, mean_out, cls_out, val_out, S(stream));
}
int gcv_score_videos_pair(const float* logits_ed, const float* logits_vae, int n_frames, int frames_per_video, float* out,
                          void* stream) {
  return score_videos_pair(logits_ed, logits_vae, n_frames, frames_per_video, out, S(stream));
}

}  // extern "C"
