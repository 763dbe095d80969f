/*
 * genconvit_b200 -- C ABI of the sm_100a kernel library behind the GenConViT
 * frame-inference forward (libgenconvit_b200.so).
 *
 * The reference (ctxnn/GenConViT) is pure Python/PyTorch: its "FFI" for this
 * path is the ATen operator set reached from model/genconvit_ed.py,
 * model/genconvit_vae.py and timm's ConvNeXt.  Each entry point below names the
 * reference operator call site(s) it replaces.  The library has no torch
 * dependency: plain device pointers, sizes, a cudaStream_t (passed as void*),
 * int status return (0 = ok, negative = GCV_ERR_*), no exceptions.
 *
 * Conventions
 *   - activations are NHWC ("tokens x channels"), element type `dtype`
 *     (GCV_F32 / GCV_BF16 / GCV_F16); parameter vectors (bias, gamma, LN
 *     weight/bias, depthwise taps) are always fp32;
 *   - GEMMs compute D[M,N] = A[M,K] * B[N,K]^T with fp32 accumulation, A and
 *     B both K-major (row-major [rows, K] with leading dimensions lda/ldb in
 *     elements), i.e. exactly nn.Linear with B = weight;
 *   - all pointers are device pointers unless stated otherwise; all kernels are
 *     enqueued on `stream` and never synchronise.
 */
#ifndef GENCONVIT_B200_H_
#define GENCONVIT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCV_ABI_VERSION 3

enum gcv_dtype { GCV_F32 = 0, GCV_BF16 = 1, GCV_F16 = 2 };
enum gcv_act { GCV_ACT_NONE = 0, GCV_ACT_GELU = 1, GCV_ACT_RELU = 2, GCV_ACT_LEAKY = 3 };
enum gcv_store { GCV_STORE_ROWS = 0, GCV_STORE_PIXEL_SHUFFLE2 = 1 };
enum gcv_backend { GCV_GEMM_AUTO = 0, GCV_GEMM_TCGEN05 = 1, GCV_GEMM_SIMT = 2 };

enum gcv_status {
  GCV_OK = 0,
  GCV_ERR_BAD_ARG = -1,       /* shape / alignment / dtype the kernel cannot take */
  GCV_ERR_CUDA = -2,          /* a CUDA runtime call failed; see gcv_last_error() */
  GCV_ERR_UNSUPPORTED = -3,   /* e.g. tcgen05 path asked for on fp32 data        */
  GCV_ERR_NO_DRIVER = -4      /* cuTensorMapEncodeTiled could not be resolved    */
};

/*
 * Epilogue applied to the fp32 accumulator of both GEMM back ends, in this
 * order:  if (ln_stats) acc = rstd[m] * (acc - mean[m] * ln_colsum[n]);   (folded LayerNorm, see below)
 *         v = acc + bias[n];  v = act(v);
 *         if (eps)      { mu = v; v = eps[m, perm(n)] * exp(0.5*mu) + mu; }
 *         if (gamma)    v = gamma[n] * v;
 *         if (residual) v = residual[m, n] + v;
 * and stored as `dtype` (or fp32 when out_f32 != 0) at
 *   GCV_STORE_ROWS:           D[m*ldd + n]
 *   GCV_STORE_PIXEL_SHUFFLE2: NHWC [B, 2H, 2W, Co] with m = (b,h,w),
 *                             n = (i*2+j)*Co + co  ->  (b, 2h+i, 2w+j, co)
 */
typedef struct gcv_epilogue {
  const float* bias;        /* [N] or NULL */
  int32_t act;              /* gcv_act */
  const float* gamma;       /* [N] layer scale or NULL */
  const void* residual;     /* [M, ldr] of `dtype`, or NULL (may alias D) */
  int64_t ldr;
  const float* eps;         /* [M, N] fp32 VAE epsilon in the reference's latent order, or NULL */
  int32_t eps_c;            /* latent is NHWC-ordered here: n = hw*eps_c + c; reference index = c*eps_hw + hw */
  int32_t eps_hw;
  float* mu_out;            /* optional [M, N] fp32 copy of mu (same column order as D), or NULL */
  int32_t store;            /* gcv_store */
  int32_t ps_h, ps_w, ps_co;
  int64_t ldd;
  int32_t out_f32;
  /* Folded LayerNorm over the K (= channel) dimension of A (timm ConvNeXtBlock.norm feeding mlp.fc1): A holds the
   * un-normalised rows, B holds W * diag(ln_weight), bias holds b + W * ln_bias, ln_colsum[n] = sum_k B[n,k], and
   * ln_stats holds per-row partial sums [M][ln_chunks] of (sum, sum of squares) over chunks of K (written by
   * gcv_dwconv7_stats); mean = sum/K, rstd = rsqrt(sumsq/K - mean^2 + ln_eps).  NULL = no fold.
   * tcgen05 back end: only together with bias + GELU and a 16-bit row-major output. */
  const float* ln_stats;
  const float* ln_colsum;
  int32_t ln_chunks;
  float ln_eps;
} gcv_epilogue;

/* -- library ------------------------------------------------------------- */
int gcv_abi_version(void);
const char* gcv_last_error(void);          /* host string, valid until the next failing call */
int gcv_device_supported(int device);      /* 1 when `device` is compute capability 10.x */

/* -- contraction kernels --------------------------------------------------
 * Replaces: every nn.Linear / F.linear on the path (timm Mlp fc1/fc2, head.fc;
 * genconvit_ed.py:73-74,87; genconvit_vae.py:36,101-103,114), the patchify
 * convolutions (timm stem 4x4 s4, downsample 2x2 s2), the im2col'd 3x3
 * convolutions (genconvit_ed.py:14-30; genconvit_vae.py:16-28) and the k2 s2
 * transposed convolutions (genconvit_ed.py:44-56; genconvit_vae.py:68-77).
 * backend: GCV_GEMM_TCGEN05 = TMA-fed tcgen05.mma with TMEM accumulators
 * (bf16/fp16 only; K % 8 == 0, lda % 8 == 0, ldb % 8 == 0, 16-byte aligned
 * bases); GCV_GEMM_SIMT = fp32-FMA kernel (any dtype/shape); AUTO picks
 * tcgen05 whenever its constraints hold.
 */
int gcv_gemm(int backend, int dtype, const void* A, int64_t lda, const void* B, int64_t ldb,
             void* D, int64_t M, int64_t N, int64_t K, const gcv_epilogue* ep, void* stream);

/* Fused ConvNeXt MLP (timm ConvNeXtBlock: mlp.fc1 -> GELU -> mlp.fc2 -> * gamma -> + shortcut) for the wide-token
 * stages: x[M,C] += gamma * (GELU(y[M,C] * W1[4C,C]^T + b1) * W2[C,4C]^T + b2), in place on x; the [M,4C] hidden
 * activation stays in shared memory / TMEM.  bf16/fp16, C in {96,192} (gcv_mlp_fused_supported); other widths use
 * two gcv_gemm calls. */
int gcv_mlp_fused_supported(int dtype, int C);
int gcv_mlp_fused(int dtype, const void* y, const void* w1, const float* b1, const void* w2, const float* b2,
                  const float* gamma, void* x, int64_t M, int C, void* stream);
/* Same with the block's LayerNorm folded in (see gcv_epilogue.ln_stats): y holds the un-normalised depthwise-conv
 * output, w1 = W1 * diag(ln_weight), b1 = b1 + W1 * ln_bias, colsum1[n] = sum_k w1[n,k], ln_stats from
 * gcv_dwconv7_stats ([M][C/32] x (sum, sumsq)). */
int gcv_mlp_fused_ln(int dtype, const void* y, const float* ln_stats, float ln_eps, const void* w1, const float* b1,
                     const float* colsum1, const void* w2, const float* b2, const float* gamma, void* x, int64_t M,
                     int C, void* stream);

/* -- ConvNeXt memory-bound kernels ----------------------------------------
 * gcv_dwconv7_ln: timm ConvNeXtBlock.conv_dw (7x7, pad 3, groups=C, bias) fused
 *   with ConvNeXtBlock.norm (LayerNorm over C, eps).  x,y: [B,H,W,C]; taps: [49,C].
 * gcv_ln_patchify2: stage downsample = LayerNorm2d(C) then the im2col of the
 *   2x2 s2 conv: x [B,H,W,C] -> A [B*(H/2)*(W/2), 4C], column (kh*2+kw)*C + c.
 * gcv_stem_patchify_nchw / _nhwc: im2col of the 4x4 s4 stem conv from the fp32
 *   NCHW frames / from an NHWC image of `dtype`: -> A [B*(H/4)*(W/4), 48],
 *   column (kh*4+kw)*3 + c.
 * gcv_layernorm_rows: LayerNorm over the last dimension of [rows, C] (stem.1).
 * gcv_pool_ln: head.global_pool + head.norm: [B,HW,C] -> LN(mean over HW) [B,C].
 */
int gcv_dwconv7_ln(int dtype, const void* x, void* y, const float* taps, const float* bias,
                   const float* ln_w, const float* ln_b, float eps,
                   int B, int H, int W, int C, void* stream);
/* gcv_dwconv7_stats (bf16/fp16, C % 32 == 0): y = conv_dw(x) + bias WITHOUT the LayerNorm, on the tensor cores
 *   (banded-Toeplitz m16n8k16 MMAs, see csrc/dwconv_mma.cu), plus the LayerNorm partial sums of y:
 *   stats [B*H*W][C/32] x (sum, sum of squares) fp32 over each 32-channel chunk, for the folded LayerNorm of
 *   gcv_gemm (gcv_epilogue.ln_stats) / gcv_mlp_fused_ln.  stats may be NULL. */
int gcv_dwconv7_stats(int dtype, const void* x, void* y, float* stats, const float* taps, const float* bias,
                      int B, int H, int W, int C, void* stream);
/* gcv_ln_finalize: reduce gcv_dwconv7_stats' partial sums [M][chunks] x (sum, sumsq) to out [M] x (rstd, -mean*rstd)
 *   (mean = sum/K, rstd = rsqrt(sumsq/K - mean^2 + eps)); pass `out` as gcv_epilogue.ln_stats with ln_chunks = 0. */
int gcv_ln_finalize(const float* stats, float* out, int64_t M, int chunks, int K, float eps, void* stream);
int gcv_ln_patchify2(int dtype, const void* x, void* a, const float* ln_w, const float* ln_b, float eps,
                     int B, int H, int W, int C, void* stream);
/* gcv_stem_fused (bf16/fp16): the whole ConvNeXt stem -- Conv2d(3,96,k4,s4) + bias + LayerNorm2d(96) (timm ConvNeXt.stem,
 *   genconvit_ed.py:82-83 / genconvit_vae.py:111-112) -- in one pass on the tensor cores: frames in, NHWC tokens
 *   y [B*(H/4)*(W/4), 96] out.  nchw != 0: x = fp32 [B,3,H,W] and w = [96][(c,kh,kw)] (the conv weight as stored);
 *   nchw == 0: x = `dtype` [B,H,W,3] and w = [96][(kh,kw,c)] (the stem GEMM's B layout). */
int gcv_stem_fused(int dtype, int nchw, const void* x, void* y, const void* w, const float* bias, const float* ln_w,
                   const float* ln_b, float eps, int B, int H, int W, void* stream);
int gcv_stem_patchify_nchw(int dtype, const float* x, void* a, int B, int H, int W, void* stream);
int gcv_stem_patchify_nhwc(int dtype, const void* x, void* a, int B, int H, int W, void* stream);
/* gcv_stem_fused_u8 / gcv_conv3x3_first_u8: the same two first-touch kernels reading the RAW uint8 face crops
 *   x [B,H,W,3] (what model/pred_func.py:95-108 preprocess_frame starts from) and applying its
 *   (x / 255 - mean) / std (dataset/loader.py:63-77; mean3 / std3: host pointers to 3 floats) on the fly, as a
 *   3 x 256 table holding exactly the 16-bit operand the fp32 entry points build from the pre-processed frame:
 *   outputs are bit-identical to gcv_preprocess_frames followed by gcv_stem_fused(nchw = 1) / gcv_conv3x3_first.
 *   gcv_stem_fused_u8 takes w = [96][(c,kh,kw)] like the nchw form.  bf16/fp16 only. */
int gcv_stem_fused_u8(int dtype, const uint8_t* x, void* y, const void* w, const float* bias, const float* ln_w,
                      const float* ln_b, float eps, int B, int H, int W, const float* mean3, const float* std3,
                      void* stream);
int gcv_conv3x3_first_u8(int dtype, const uint8_t* x, void* y, const float* w, const float* b, int stride, int act,
                         int pool, int B, int H, int W, const float* mean3, const float* std3, void* stream);
int gcv_layernorm_rows(int dtype, const void* x, void* y, const float* w, const float* b, float eps,
                       int64_t rows, int C, void* stream);
int gcv_pool_ln(int dtype, const void* x, void* y, const float* w, const float* b, float eps,
                int B, int HW, int C, void* stream);

/* -- autoencoder kernels ----------------------------------------------------
 * gcv_conv3x3_first: first 3x3 conv (Cin=3 -> 16, pad 1) straight from the fp32
 *   NCHW frames, w: [16][3][3][3] fp32 (OIHW), followed by `act` and, when
 *   pool != 0, a 2x2 max-pool.  stride 1 + ReLU + pool = genconvit_ed.py:14-16;
 *   stride 2 + (BatchNorm folded into w,b by the host) + LeakyReLU =
 *   genconvit_vae.py:16-18.  y: NHWC.
 * gcv_im2col3x3: x [B,H,W,C] -> A [B*Ho*Wo, 9C], column (kh*3+kw)*C + c, pad 1.
 * gcv_maxpool2: NHWC 2x2 s2 max-pool (genconvit_ed.py:16,20,24,28,32).
 * gcv_conv3x3_c16: the encoders' second layer, Conv2d(16 -> 32, k3, pad 1) as a direct tensor-core
 *   convolution (no im2col matrix): x [B,H,W,16], w [32][(kh,kw,ci)] of `dtype` (the GEMM B layout),
 *   bias fp32 [32]; stride 1 + ReLU + fused 2x2 max-pool = genconvit_ed.py:18-20, stride 2 +
 *   LeakyReLU (BatchNorm folded by the host) = genconvit_vae.py:19-21.  16-bit dtypes only.
 * gcv_convt2x2_small: the decoders' output layer ConvTranspose2d(16 -> 3, k2 s2) + act
 *   (genconvit_ed.py:56-57; genconvit_vae.py:77-78) as a streaming kernel: x [B,H,W,16] ->
 *   y [B,2H,2W,3]; w: fp32 [(i*2+j)*3 + co][ci] (the GEMM B layout), bias: fp32 [3].
 * gcv_resize2x_to_nchw: the returned x_hat: bilinear (align_corners=False) 2x
 *   upscale of NHWC `dtype` [B,H,W,3] to NCHW fp32 [B,3,2H,2W]
 *   (genconvit_vae.py:105,116 -- antialias is a no-op when upscaling).
 * gcv_nhwc_to_nchw_f32: layout/dtype conversion for tensors handed back to torch.
 */
int gcv_conv3x3_first(int dtype, const float* x, void* y, const float* w, const float* b,
                      int stride, int act, int pool, int B, int H, int W, void* stream);
int gcv_im2col3x3(int dtype, const void* x, void* a, int B, int H, int W, int C, int stride, void* stream);
int gcv_maxpool2(int dtype, const void* x, void* y, int B, int H, int W, int C, void* stream);
int gcv_conv3x3_c16(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act,
                    int pool, int B, int H, int W, void* stream);
/* gcv_conv3x3_c32: the same for the third layer, Conv2d(32 -> 64, k3, pad 1): x [B,H,W,32], w [64][(kh,kw,ci)];
 *   stride 1 + ReLU + 2x2 max-pool = genconvit_ed.py:22-24, stride 2 + LeakyReLU = genconvit_vae.py:22-24. */
int gcv_conv3x3_c32(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act,
                    int pool, int B, int H, int W, void* stream);
/* gcv_conv3x3_tc: the encoders' wide layers Conv2d(64 -> 128) / Conv2d(128 -> 256), k3 pad 1, as a tcgen05 implicit
 *   GEMM (no im2col matrix: every tap's A tile is one 4-D TMA box of x, zero-filled outside the image):
 *   x [B,H,W,C], C in {64,128}; w [N][(kh,kw,ci)] of `dtype`, N % 64 == 0, N <= 256; bias fp32 [N]; stride 1 + ReLU +
 *   fused 2x2 max-pool = genconvit_ed.py:26-32, stride 2 + LeakyReLU (BatchNorm folded by the host) =
 *   genconvit_vae.py:25-27.  y [B,Ho,Wo,N] (pool: [B,Ho/2,Wo/2,N]).  16-bit dtypes only. */
int gcv_conv3x3_tc_supported(int dtype, int C, int N);
int gcv_conv3x3_tc(int dtype, const void* x, void* y, const void* w, const float* bias, int stride, int act, int pool,
                   int B, int H, int W, int C, int N, void* stream);
int gcv_convt2x2_small(int dtype, const void* x, void* y, const float* w, const float* bias, int act,
                       int B, int H, int W, int CI, int CO, void* stream);
/* gcv_convt2x2_mma: the decoders' small-channel layers ConvTranspose2d(CI -> CI/2, k2 s2) + act as a per-token
 *   tensor-core product with a pixel-shuffle store (reference model/genconvit_ed.py:51-56, genconvit_vae.py:56-62):
 *   x [B*H*W, CI] tokens, w1 [(i,j,co), CI] and b1 [4*CI/2] (bias repeated per tap) as for gcv_gemm's pixel-shuffle
 *   mode, CI = 32 or 64, B*H*W a multiple of 16, 16-bit dtypes.  tail = 0: y [B,2H,2W,CI/2].  tail = 1 (CI = 32): also applies
 *   the output layer ConvTranspose2d(16 -> 3) + act to the 16-channel result without storing it:
 *   w2 [(i,j,c), 16] of `dtype`, b2 [12], y [B,4H,4W,3]. */
int gcv_convt2x2_mma(int dtype, const void* x, void* y, const void* w1, const float* b1, const void* w2,
                     const float* b2, int act, int B, int H, int W, int CI, int tail, void* stream);
int gcv_resize2x_to_nchw(int dtype, const void* x, float* y, int B, int H, int W, int C, void* stream);
int gcv_nhwc_to_nchw_f32(int dtype, const void* x, float* y, int B, int H, int W, int C, void* stream);

/* -- Swin-T embedder kernels ------------------------------------------------
 * timm swin_tiny_patch4_window7_224 (`self.embedder`, reference model/genconvit_ed.py:69, genconvit_vae.py:96);
 * every Linear / LayerNorm of it runs on gcv_gemm / gcv_layernorm_rows, these are the rest:
 * gcv_swin_window_attention: W-MSA / SW-MSA core of one SwinTransformerBlock: qkv [B*res*res, 3C] (q | k | v, heads
 *   of 32 channels) -> out [B*res*res, C] = softmax(q k^T / sqrt(32) + rel-pos bias (+ shifted-window mask)) v per
 *   7x7 window and head; cyclic shift / window partition / reverse are index arithmetic; bias_table [169, heads].
 * gcv_swin_patch_merge: PatchMerging gather x [B,res,res,C] -> [B,res/2,res/2,4C] in timm's (0,0),(1,0),(0,1),(1,1)
 *   order (LayerNorm(4C) and the bias-free reduction follow as gcv_layernorm_rows + gcv_gemm).
 * gcv_mean_tokens: x [B,L,C] -> mean over L [B,C] (the pooling before `head`).
 */
int gcv_swin_window_attention(int dtype, const void* qkv, void* out, const float* bias_table,
                              int B, int res, int C, int heads, int shift, void* stream);
int gcv_swin_patch_merge(int dtype, const void* x, void* out, int B, int res, int C, void* stream);
int gcv_mean_tokens(int dtype, const void* x, void* y, int B, int L, int C, void* stream);

/* -- frame ingest -------------------------------------------------------------
 * model/pred_func.py:95-108 (preprocess_frame) + dataset/loader.py:63-77: uint8 NHWC frames (device) -> fp32 NCHW
 * ((x / 255) - mean) / std, bit-identical to the reference's fp32 CPU arithmetic.  mean3 / std3: HOST float[3].
 */
int gcv_preprocess_frames(const uint8_t* x, float* y, int N, int H, int W, const float* mean3, const float* std3,
                          void* stream);

/* -- scoring ----------------------------------------------------------------
 * model/pred_func.py:111-131 (pred_vid after the forward + max_prediction_value),
 * batched over videos: logits [n_nets*n_frames, 2] fp32 with each net's rows
 * contiguous (GenConViT.forward's cat(dim=0), genconvit.py:74); video v owns
 * frames [v*fpv, (v+1)*fpv) of every net.  Writes per video the mean sigmoid
 * [V,2], the argmax class and the reported score.
 */
int gcv_score_videos(const float* logits, int n_nets, int n_frames, int frames_per_video,
                     float* mean_out, int32_t* cls_out, float* val_out, void* stream);

/* The same scoring straight from the two networks' logit buffers, for the bulk runtime: no concatenation of the
 * ED and VAE rows (reference model/genconvit.py:74 torch.cat) and one packed result for a single device->host copy.
 *   logits_ed, logits_vae  fp32 [n_frames,2] each; either may be NULL (single-network model)
 *   out                    fp32 [2][n_videos]: row 0 = class (0.0 / 1.0), row 1 = score; same tie rules as above
 * (reference model/pred_func.py:111-131: sigmoid, mean over the video's rows of both networks, argmax, score). */
int gcv_score_videos_pair(const float* logits_ed, const float* logits_vae, int n_frames, int frames_per_video,
                          float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GENCONVIT_B200_H_ */
