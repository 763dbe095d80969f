// Frame ingest on the GPU: uint8 NHWC face crops -> ImageNet-normalised fp32 NCHW frames.
// Reference: model/pred_func.py:95-108 (preprocess_frame: torch.tensor(frame).float().permute(0,3,1,2), then per frame
// normalize_data()["vid"](x / 255.0)) with dataset/loader.py:63-77 (Normalize(mean=[0.485,0.456,0.406],
// std=[0.229,0.224,0.225])).  Same fp32 operation order (x / 255, - mean, / std with IEEE division), so the result is
// bit-identical to the reference's CPU loop; the Python loop over frames and the fp32 host->device copy (4x the
// bytes) disappear.  HBM-bound: 3 B read + 12 B written per pixel.
#include "common.cuh"

namespace gcv {

namespace {

// thread = 4 consecutive pixels of one image row: one 12-byte read (3 x 32-bit), three 16-byte channel-plane stores
__global__ void __launch_bounds__(256)
preprocess_frames_kernel(const uint8_t* __restrict__ x, float* __restrict__ y, int64_t quads, int64_t plane, float m0,
                         float m1, float m2, float s0, float s1, float s2) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= quads) return;
  const int64_t pix = i * 4;                       // pixel index inside the whole batch (H*W % 4 == 0)
  const int64_t n = pix / plane, r = pix - n * plane;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(x + pix * 3);
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  const uint8_t b[12] = {(uint8_t)w0, (uint8_t)(w0 >> 8), (uint8_t)(w0 >> 16), (uint8_t)(w0 >> 24),
                         (uint8_t)w1, (uint8_t)(w1 >> 8), (uint8_t)(w1 >> 16), (uint8_t)(w1 >> 24),
                         (uint8_t)w2, (uint8_t)(w2 >> 8), (uint8_t)(w2 >> 16), (uint8_t)(w2 >> 24)};
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  float* dst = y + n * 3 * plane + r;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 o;
    o.x = __fdiv_rn(__fdiv_rn((float)b[c], 255.0f) - mean[c], sd[c]);
    o.y = __fdiv_rn(__fdiv_rn((float)b[3 + c], 255.0f) - mean[c], sd[c]);
    o.z = __fdiv_rn(__fdiv_rn((float)b[6 + c], 255.0f) - mean[c], sd[c]);
    o.w = __fdiv_rn(__fdiv_rn((float)b[9 + c], 255.0f) - mean[c], sd[c]);
    *reinterpret_cast<float4*>(dst + c * plane) = o;
  }
}

}  // namespace

int preprocess_frames(const uint8_t* x, float* y, int N, int H, int W, const float* mean3, const float* std3,
                      cudaStream_t stream) {
  GCV_REQUIRE(N > 0 && H > 0 && W > 0 && ((int64_t)H * W) % 4 == 0, "preprocess_frames: H*W must be a multiple of 4");
  GCV_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0,
              "preprocess_frames: x must be 4-byte and y 16-byte aligned");
  GCV_REQUIRE(mean3 && std3, "preprocess_frames: mean / std (host pointers to 3 floats) are required");
  const int64_t plane = (int64_t)H * W, quads = (int64_t)N * plane / 4;
  const int64_t grid = (quads + 255) / 256;
  GCV_REQUIRE(grid < 2147483647LL, "preprocess_frames: too many pixels");
  preprocess_frames_kernel<<<(unsigned)grid, 256, 0, stream>>>(x, y, quads, plane, mean3[0], mean3[1], mean3[2], std3[0],
                                                              std3[1], std3[2]);
  return check_launch("preprocess_frames");
}

}  // namespace gcv
