// vitron_b200 — LanguageBind image / video pre-processing on the device.
// reference: vitron/model/multimodal_encoder/languagebind/image/processing_image.py:15-25
//   ToTensor -> Resize(224, BICUBIC) -> CenterCrop(224) -> Normalize(OPENAI mean / std)
// and video/processing_video.py:26-70
//   x / 255 -> NormalizeVideo -> ShortSideScale(224) (bilinear) -> CenterCropVideo(224) -> horizontal flip.
// One fused kernel: uint8 HWC frames in, normalised planar float/bf16 out; the resized image is never
// materialised — every output pixel of the crop window evaluates its own interpolation footprint.
// (Normalisation commutes with the interpolation because the weights sum to one.)
#include "common.cuh"
#include "vitron_b200.h"

namespace vb {

// ATen upsample_bicubic2d (A = -0.75): cubic convolution coefficients for fractional offset t
__device__ __forceinline__ void cubic_coeffs(float t, float A, float (&c)[4]) {
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}
// ATen _upsample_bicubic2d_aa filter (a = -0.5)
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.f) return ((a + 2.f) * x - (a + 3.f)) * x * x + 1.f;
  if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * a;
  return 0.f;
}

struct PreParams {
  const uint8_t* src;  // [n, h, w, 3]
  void* dst;
  int n, h, w;           // input frames
  int rh, rw;            // virtual resized size
  int top, left;         // crop offset inside the resized image
  int oh, ow;            // output (crop) size
  long long dst_n, dst_c;  // element strides of frame / channel in dst (row stride = ow)
  float mean[3], inv_std[3];
  int mode;              // 0 bilinear, 1 bicubic (A=-0.75), 2 bicubic antialiased
  int flip, out_bf16;
};

__device__ __forceinline__ void fetch(const uint8_t* p, float wgt, float (&acc)[3]) {
  acc[0] = fmaf(wgt, static_cast<float>(p[0]), acc[0]);
  acc[1] = fmaf(wgt, static_cast<float>(p[1]), acc[1]);
  acc[2] = fmaf(wgt, static_cast<float>(p[2]), acc[2]);
}

__global__ void preprocess_kernel(const PreParams p) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(p.n) * p.oh * p.ow;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % p.ow);
  const int oy = static_cast<int>((idx / p.ow) % p.oh);
  const int f = static_cast<int>(idx / (static_cast<long long>(p.ow) * p.oh));
  const int ry = oy + p.top;
  const int rx = (p.flip ? (p.ow - 1 - ox) : ox) + p.left;
  const float sy = static_cast<float>(p.h) / static_cast<float>(p.rh);
  const float sx = static_cast<float>(p.w) / static_cast<float>(p.rw);
  const uint8_t* img = p.src + static_cast<long long>(f) * p.h * p.w * 3;
  float acc[3] = {0.f, 0.f, 0.f};
  if (p.mode == 0) {
    // bilinear, align_corners=False: src = max(scale * (dst + 0.5) - 0.5, 0)
    const float fy = fmaxf(sy * (ry + 0.5f) - 0.5f, 0.f), fx = fmaxf(sx * (rx + 0.5f) - 0.5f, 0.f);
    const int y0 = min(static_cast<int>(fy), p.h - 1), x0 = min(static_cast<int>(fx), p.w - 1);
    const int y1 = min(y0 + 1, p.h - 1), x1 = min(x0 + 1, p.w - 1);
    const float ly = fy - y0, lx = fx - x0;
    fetch(img + (static_cast<long long>(y0) * p.w + x0) * 3, (1.f - ly) * (1.f - lx), acc);
    fetch(img + (static_cast<long long>(y0) * p.w + x1) * 3, (1.f - ly) * lx, acc);
    fetch(img + (static_cast<long long>(y1) * p.w + x0) * 3, ly * (1.f - lx), acc);
    fetch(img + (static_cast<long long>(y1) * p.w + x1) * 3, ly * lx, acc);
  } else if (p.mode == 1) {
    // bicubic, align_corners=False, no antialias: 4x4 taps, indices clamped to the image
    const float fy = sy * (ry + 0.5f) - 0.5f, fx = sx * (rx + 0.5f) - 0.5f;
    const float yf = floorf(fy), xf = floorf(fx);
    float cy[4], cx[4];
    cubic_coeffs(fy - yf, -0.75f, cy);
    cubic_coeffs(fx - xf, -0.75f, cx);
    const int iy = static_cast<int>(yf), ix = static_cast<int>(xf);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int yy = min(max(iy - 1 + a, 0), p.h - 1);
      float row[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int xx = min(max(ix - 1 + b, 0), p.w - 1);
        fetch(img + (static_cast<long long>(yy) * p.w + xx) * 3, cx[b], row);
      }
      acc[0] = fmaf(cy[a], row[0], acc[0]); acc[1] = fmaf(cy[a], row[1], acc[1]); acc[2] = fmaf(cy[a], row[2], acc[2]);
    }
  } else {
    // antialiased bicubic (ATen _compute_indices_weights_aa): support and filter stretched by the scale when
    // down-sampling, weights normalised over the taps that fall inside the image
    const float scy = fmaxf(sy, 1.f), scx = fmaxf(sx, 1.f);
    const float supy = 2.f * scy, supx = 2.f * scx;
    const float cyc = sy * (ry + 0.5f), cxc = sx * (rx + 0.5f);
    const int ymin = max(static_cast<int>(cyc - supy + 0.5f), 0), ymax = min(static_cast<int>(cyc + supy + 0.5f), p.h);
    const int xmin = max(static_cast<int>(cxc - supx + 0.5f), 0), xmax = min(static_cast<int>(cxc + supx + 0.5f), p.w);
    float wys = 0.f, wxs = 0.f;
    for (int y = ymin; y < ymax; ++y) wys += cubic_aa((y - cyc + 0.5f) / scy);
    for (int x = xmin; x < xmax; ++x) wxs += cubic_aa((x - cxc + 0.5f) / scx);
    const float iwy = wys != 0.f ? 1.f / wys : 0.f, iwx = wxs != 0.f ? 1.f / wxs : 0.f;
    for (int y = ymin; y < ymax; ++y) {
      const float wy = cubic_aa((y - cyc + 0.5f) / scy) * iwy;
      float row[3] = {0.f, 0.f, 0.f};
      for (int x = xmin; x < xmax; ++x) fetch(img + (static_cast<long long>(y) * p.w + x) * 3, cubic_aa((x - cxc + 0.5f) / scx) * iwx, row);
      acc[0] = fmaf(wy, row[0], acc[0]); acc[1] = fmaf(wy, row[1], acc[1]); acc[2] = fmaf(wy, row[2], acc[2]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = (acc[c] * (1.f / 255.f) - p.mean[c]) * p.inv_std[c];
    const long long o = f * p.dst_n + c * p.dst_c + static_cast<long long>(oy) * p.ow + ox;
    if (p.out_bf16) reinterpret_cast<bf16*>(p.dst)[o] = __float2bfloat16(v);
    else reinterpret_cast<float*>(p.dst)[o] = v;
  }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_preprocess_frames(const uint8_t* src, void* dst, int64_t n, int64_t h, int64_t w, int64_t rh,
                                       int64_t rw, int64_t top, int64_t left, int64_t oh, int64_t ow, int64_t dst_n,
                                       int64_t dst_c, const float* mean3, const float* std3, int mode, int flip,
                                       int out_bf16, cudaStream_t stream) {
  VB_CHECK_ARG(src && dst && mean3 && std3 && n > 0 && h > 0 && w > 0 && rh > 0 && rw > 0 && oh > 0 && ow > 0);
  VB_CHECK_ARG(top >= 0 && left >= 0 && top + oh <= rh && left + ow <= rw && mode >= 0 && mode <= 2);
  VB_CHECK_ARG(h < (1 << 24) && w < (1 << 24));
  PreParams p;
  p.src = src; p.dst = dst;
  p.n = (int)n; p.h = (int)h; p.w = (int)w; p.rh = (int)rh; p.rw = (int)rw; p.top = (int)top; p.left = (int)left;
  p.oh = (int)oh; p.ow = (int)ow; p.dst_n = dst_n; p.dst_c = dst_c;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.inv_std[c] = 1.0f / std3[c]; }
  p.mode = mode; p.flip = flip; p.out_bf16 = out_bf16;
  const long long total = n * oh * ow;
  preprocess_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(p);
  VB_LAUNCH_CHECK();
  return VB_OK;
}
