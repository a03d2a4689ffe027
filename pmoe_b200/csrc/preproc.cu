// Input pipeline on the GPU: what the reference's dataset does to every decoded camera frame in
// eval mode — Crop (augmenter.py:43-49: rows [top : H-bottom]) -> torchvision Resize((224, 224)) on a PIL image
// (data_loader.py:275-281; Pillow's antialiased two-pass BILINEAR resample with 22-bit fixed-point coefficients and a
// uint8 intermediate) -> ToTensor (uint8 HWC -> float CHW / 255) -> torch.stack over the frames (data_loader.py:288-300).
// The arithmetic is integer and restated bit for bit from Pillow's Resample.c (third-party, absent from /root/reference;
// pinned version 12.2 in this image): ss = 2^21 + sum pixel * k, out = clip8(ss >> 22), horizontal pass first, then the
// vertical pass over the uint8 intermediate. The coefficient tables are computed on the host (pmoe_b200/preproc.py, in
// double like precompute_coeffs) and passed in. HBM-bound: one read of the source rows that survive the crop, one
// write of the float output; the intermediate stays in L2 for typical frame sizes.
// Compiled WITHOUT --use_fast_math: the /255 is an IEEE division table.
#include "host_util.h"

namespace pmoe {

__device__ __forceinline__ int clip8_fixed(int ss) {
  const int v = ss >> 22;  // arithmetic shift, like Pillow's clip8 lookup index
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: src (N, Hs, Ws, 3) uint8, rows [top, top + Hc) -> tmp (N, Hc, OW, 3) uint8
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ tmp, int N, int Hs, int Ws,
                                                       int top, int Hc, int OW, const int* __restrict__ bounds,
                                                       const int* __restrict__ kk, int ksize) {
  const long long total = (long long)N * Hc * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    const long long r = i / OW;
    const int y = (int)(r % Hc), n = (int)(r / Hc);
    const int xmin = __ldg(bounds + 2 * ox), xcnt = __ldg(bounds + 2 * ox + 1);
    const uint8_t* row = src + (((long long)n * Hs + top + y) * Ws + xmin) * 3;
    const int* k = kk + (long long)ox * ksize;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int x = 0; x < xcnt; ++x) {
      const int w = __ldg(k + x);
      s0 += (int)row[3 * x] * w;
      s1 += (int)row[3 * x + 1] * w;
      s2 += (int)row[3 * x + 2] * w;
    }
    uint8_t* o = tmp + i * 3;
    o[0] = (uint8_t)clip8_fixed(s0);
    o[1] = (uint8_t)clip8_fixed(s1);
    o[2] = (uint8_t)clip8_fixed(s2);
  }
}

// vertical pass + ToTensor: tmp (N, Hc, OW, 3) uint8 -> dst float, element (n, c, oy, ox) at n*sn + c*sc + oy*sh + ox*sw
// (NCHW fp32 for the module API; any strided layout works). lut[v] = (float)v / 255.0f, computed on the host.
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ tmp, float* __restrict__ dst, uint8_t* __restrict__ dst_u8,
                                                       int N, int Hc, int OW, int OH, const int* __restrict__ bounds,
                                                       const int* __restrict__ kk, int ksize, long long sn, long long sc, long long sh,
                                                       long long sw, const float* __restrict__ lut) {
  const long long total = (long long)N * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % OW);
    const long long r = i / OW;
    const int oy = (int)(r % OH), n = (int)(r / OH);
    const int ymin = __ldg(bounds + 2 * oy), ycnt = __ldg(bounds + 2 * oy + 1);
    const uint8_t* col = tmp + (((long long)n * Hc + ymin) * OW + ox) * 3;
    const int* k = kk + (long long)oy * ksize;
    int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
    for (int y = 0; y < ycnt; ++y) {
      const int w = __ldg(k + y);
      const uint8_t* px = col + (long long)y * OW * 3;
      s0 += (int)px[0] * w;
      s1 += (int)px[1] * w;
      s2 += (int)px[2] * w;
    }
    const int v0 = clip8_fixed(s0), v1 = clip8_fixed(s1), v2 = clip8_fixed(s2);
    if (dst) {
      float* o = dst + n * sn + oy * sh + ox * sw;
      o[0] = __ldg(lut + v0);
      o[sc] = __ldg(lut + v1);
      o[2 * sc] = __ldg(lut + v2);
    }
    if (dst_u8) {
      uint8_t* o = dst_u8 + i * 3;
      o[0] = (uint8_t)v0;
      o[1] = (uint8_t)v1;
      o[2] = (uint8_t)v2;
    }
  }
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_preprocess_frames(const uint8_t* src, int32_t n, int32_t hs, int32_t ws, int32_t crop_top, int32_t crop_bottom,
                                      int32_t out_h, int32_t out_w, const int32_t* hbounds, const int32_t* hcoef, int32_t hksize,
                                      const int32_t* vbounds, const int32_t* vcoef, int32_t vksize, const float* lut255,
                                      uint8_t* tmp, float* dst, int64_t dst_sn, int64_t dst_sc, int64_t dst_sh, int64_t dst_sw,
                                      uint8_t* dst_u8, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int hc = hs - crop_top - crop_bottom;
  if (!src || !tmp || (!dst && !dst_u8) || !hbounds || !hcoef || !vbounds || !vcoef || !lut255 || n < 1 || hc < 1 || ws < 1 || out_h < 1 ||
      out_w < 1 || crop_top < 0 || crop_bottom < 0 || hksize < 1 || vksize < 1) {
    set_error("preprocess_frames: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long t1 = (long long)n * hc * out_w, t2 = (long long)n * out_h * out_w;
  const long long cap = (long long)num_sms() * 16;
  long long g1 = (t1 + 255) / 256, g2 = (t2 + 255) / 256;
  if (g1 > cap) g1 = cap;
  if (g2 > cap) g2 = cap;
  resize_h_kernel<<<(unsigned)g1, 256, 0, stream>>>(src, tmp, n, hs, ws, crop_top, hc, out_w, hbounds, hcoef, hksize);
  int rc = check_launch("preprocess_frames (horizontal)");
  if (rc) return rc;
  resize_v_kernel<<<(unsigned)g2, 256, 0, stream>>>(tmp, dst, dst_u8, n, hc, out_w, out_h, vbounds, vcoef, vksize, dst_sn, dst_sc, dst_sh,
                                                   dst_sw, lut255);
  return check_launch("preprocess_frames (vertical)");
}
