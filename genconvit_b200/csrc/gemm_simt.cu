// fp32-FMA GEMM back end:  D[M,N] = epilogue(A[M,K] * B[N,K]^T), fp32 accumulate.
//
// Serves (a) the fp32 parity mode (max-abs logit error <= 1e-4 needs true fp32
// products, not TF32/bf16 tensor-core inputs) and (b) contractions whose shape the
// tcgen05 path cannot take (K not a multiple of 8, e.g. the 500->2 head).  Any
// dtype, any shape; 64x64x16 smem tiles, 4x4 outputs per thread.
#include "common.cuh"

namespace gcv {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, void* D, int64_t M, int N,
                 int K, const gcv_epilogue ep) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                 // 16 x 16 threads, 4x4 outputs each
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  float acc[4][4] = {};
  // loader mapping: 256 threads x 4 elements = 64 rows x 16 k
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      const int64_t am = m0 + lr;
      const int bn = n0 + lr;
      As[lk + j][lr] = (am < M && k < K) ? to_f<T>(A[am * lda + k]) : 0.0f;
      Bs[lk + j][lr] = (bn < N && k < K) ? to_f<T>(B[(int64_t)bn * ldb + k]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) epilogue_one<T>(ep, m, n, N, acc[i][j], D, (int)K);
    }
  }
}

// Skinny outputs (N <= 8, e.g. the 500 -> 2 classification head, genconvit_ed.py:75 / genconvit_vae.py:104): one warp
// per output row, lanes stride over K (coalesced A reads, B rows from cache), fixed-order shuffle reduction.  The tiled
// kernel above would run such a problem on 4 CTAs with 32 barrier-separated k-steps each.
template <typename T>
__global__ void __launch_bounds__(256)
gemm_skinny_kernel(const T* __restrict__ A, int64_t lda, const T* __restrict__ B, int64_t ldb, void* D, int64_t M, int N,
                   int K, const gcv_epilogue ep) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= M) return;
  float acc[8] = {};
  const T* a = A + m * lda;
  for (int k = lane; k < K; k += 32) {
    const float av = to_f<T>(a[k]);
#pragma unroll
    for (int n = 0; n < 8; ++n)
      if (n < N) acc[n] = fmaf(av, to_f<T>(B[(int64_t)n * ldb + k]), acc[n]);
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (n < N) {
      const float v = warp_sum(acc[n]);
      if (lane == 0) epilogue_one<T>(ep, m, n, N, v, D, K);
    }
  }
}

}  // namespace

int gemm_simt(int dtype, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t M, int64_t N,
              int64_t K, const gcv_epilogue* ep, cudaStream_t stream) {
  GCV_REQUIRE(M > 0 && N > 0 && K > 0, "bad GEMM shape");
  const int64_t gm = (M + TM - 1) / TM, gn = (N + TN - 1) / TN;
  GCV_REQUIRE(gm <= 2147483647LL && gn <= 65535, "GEMM too large for the SIMT grid");
  dim3 g((unsigned)gm, (unsigned)gn);
  if (N <= 8 && (M + 7) / 8 <= 2147483647LL) {
    const unsigned gs = (unsigned)((M + 7) / 8);
    switch (dtype) {
      case GCV_F32:
        gemm_skinny_kernel<float><<<gs, 256, 0, stream>>>(reinterpret_cast<const float*>(A), lda,
                                                          reinterpret_cast<const float*>(B), ldb, D, M, (int)N, (int)K, *ep);
        break;
      case GCV_BF16:
        gemm_skinny_kernel<__nv_bfloat16><<<gs, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A), lda,
                                                                  reinterpret_cast<const __nv_bfloat16*>(B), ldb, D, M,
                                                                  (int)N, (int)K, *ep);
        break;
      case GCV_F16:
        gemm_skinny_kernel<__half><<<gs, 256, 0, stream>>>(reinterpret_cast<const __half*>(A), lda,
                                                           reinterpret_cast<const __half*>(B), ldb, D, M, (int)N, (int)K, *ep);
        break;
      default:
        GCV_REQUIRE(false, "bad dtype %d", dtype);
    }
    return check_launch("gemm_simt");
  }
  switch (dtype) {
    case GCV_F32:
      gemm_simt_kernel<float><<<g, 256, 0, stream>>>(reinterpret_cast<const float*>(A), lda,
                                                      reinterpret_cast<const float*>(B), ldb, D, M, (int)N, (int)K, *ep);
      break;
    case GCV_BF16:
      gemm_simt_kernel<__nv_bfloat16><<<g, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(A), lda,
                                                              reinterpret_cast<const __nv_bfloat16*>(B), ldb, D, M,
                                                              (int)N, (int)K, *ep);
      break;
    case GCV_F16:
      gemm_simt_kernel<__half><<<g, 256, 0, stream>>>(reinterpret_cast<const __half*>(A), lda,
                                                       reinterpret_cast<const __half*>(B), ldb, D, M, (int)N, (int)K, *ep);
      break;
    default:
      GCV_REQUIRE(false, "bad dtype %d", dtype);
  }
  return check_launch("gemm_simt");
}

}  // namespace gcv
