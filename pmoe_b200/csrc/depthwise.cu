// Depthwise k x k convolution (groups = channels) of torchvision's MobileNetV2 / V3 inverted-residual blocks — the alternative
// backbones of the reference's factory (model/blocks/backbone.py:75-104): forward, data gradient, weight gradient. NHWC, eight
// channels per thread (one 16-byte access in bf16), fp32 accumulation. These layers do k*k MACs per element read: HBM-bound
// (algorithmic bytes: x + y, and dy + x + dx for the backward), so they run on the CUDA cores with coalesced channel-contiguous
// accesses; the taps of a window hit L1/L2. Weights arrive packed as [k*k][cpad] fp32 (tap-major, channel-contiguous).
#include <cuda_bf16.h>

#include "host_util.h"

namespace pmoe {

struct DV4 {
  void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
};
static inline DV4 dv4(const PmoeView4* v) {
  DV4 o;
  o.ptr = v ? v->ptr : nullptr;
  o.n = v ? v->n : 0;
  o.h = v ? v->h : 0;
  o.w = v ? v->w : 0;
  o.c = v ? v->c : 0;
  o.sn = v ? v->sn : 0;
  o.sh = v ? v->sh : 0;
  o.sw = v ? v->sw : 0;
  return o;
}

__device__ __forceinline__ void dld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void dld8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[2 * q] = __uint_as_float(w[q] << 16);
    v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ void dst8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void dst8(__nv_bfloat16* p, const float (&v)[8]) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h);
}

// y[n,oh,ow,c] = sum_{r,s} x[n, oh*stride - pad + r, ow*stride - pad + s, c] * w[r*k+s][c]
template <typename T>
__global__ void __launch_bounds__(256) dwconv_fwd_kernel(DV4 x, DV4 y, const float* __restrict__ w, int cpad, int k, int stride, int pad) {
  const int cg = y.c / 8;
  const long long total = (long long)y.n * y.h * y.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int ow = (int)(pix % y.w);
    pix /= y.w;
    const int oh = (int)(pix % y.h), n = (int)(pix / y.h);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < k; ++r) {
      const int ih = oh * stride - pad + r;
      if (ih < 0 || ih >= x.h) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = ow * stride - pad + s;
        if (iw < 0 || iw >= x.w) continue;
        float xv[8], wv[8];
        dld8(static_cast<const T*>(x.ptr) + n * x.sn + ih * x.sh + iw * x.sw + g * 8, xv);
        dld8(w + (long long)(r * k + s) * cpad + g * 8, wv);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(xv[q], wv[q], acc[q]);
      }
    }
    dst8(static_cast<T*>(y.ptr) + n * y.sn + oh * y.sh + ow * y.sw + g * 8, acc);
  }
}

// dx[n,ih,iw,c] (+)= sum over (r,s) with (ih + pad - r) % stride == 0 etc. of dy[n,(ih+pad-r)/stride,(iw+pad-s)/stride,c] * w[r*k+s][c]
template <typename T>
__global__ void __launch_bounds__(256) dwconv_dgrad_kernel(DV4 dy, DV4 dx, const float* __restrict__ w, int cpad, int k, int stride, int pad,
                                                           int accumulate) {
  const int cg = dx.c / 8;
  const long long total = (long long)dx.n * dx.h * dx.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int iw = (int)(pix % dx.w);
    pix /= dx.w;
    const int ih = (int)(pix % dx.h), n = (int)(pix / dx.h);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    T* out = static_cast<T*>(dx.ptr) + n * dx.sn + ih * dx.sh + iw * dx.sw + g * 8;
    if (accumulate) dld8(out, acc);
    for (int r = 0; r < k; ++r) {
      const int th = ih + pad - r;
      if (th < 0 || th % stride) continue;
      const int oh = th / stride;
      if (oh >= dy.h) continue;
      for (int s = 0; s < k; ++s) {
        const int tw = iw + pad - s;
        if (tw < 0 || tw % stride) continue;
        const int ow = tw / stride;
        if (ow >= dy.w) continue;
        float dv[8], wv[8];
        dld8(static_cast<const T*>(dy.ptr) + n * dy.sn + oh * dy.sh + ow * dy.sw + g * 8, dv);
        dld8(w + (long long)(r * k + s) * cpad + g * 8, wv);
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = fmaf(dv[q], wv[q], acc[q]);
      }
    }
    dst8(out, acc);
  }
}

// dw[tap][c] += sum over output pixels of dy[n,oh,ow,c] * x[n, oh*stride - pad + r, ow*stride - pad + s, c]. grid = (pixel slabs, taps);
// block = lanes x channel groups; per-thread partial sums, shared-memory reduction over the lanes, one atomic per channel and block.
template <typename T>
__global__ void __launch_bounds__(256) dwconv_wgrad_kernel(DV4 x, DV4 dy, float* __restrict__ dw, int cpad, int k, int stride, int pad,
                                                           long long pix_per_block) {
  __shared__ float sm[256 * 8];
  const int cg = dy.c / 8;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int tap = blockIdx.y, r = tap / k, s = tap % k;
  const long long npix = (long long)dy.n * dy.h * dy.w;
  const long long p0 = blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const int ow = (int)(p % dy.w);
      const long long t = p / dy.w;
      const int oh = (int)(t % dy.h), n = (int)(t / dy.h);
      const int ih = oh * stride - pad + r, iw = ow * stride - pad + s;
      if (ih < 0 || ih >= x.h || iw < 0 || iw >= x.w) continue;
      float dv[8], xv[8];
      dld8(static_cast<const T*>(dy.ptr) + n * dy.sn + oh * dy.sh + ow * dy.sw + g * 8, dv);
      dld8(static_cast<const T*>(x.ptr) + n * x.sn + ih * x.sh + iw * x.sw + g * 8, xv);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = fmaf(dv[q], xv[q], acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) sm[threadIdx.x * 8 + q] = (lane < lanes) ? acc[q] : 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < cg * 8; c += blockDim.x) {   // channel c = group c/8, element c%8; sum over the lanes
    float tot = 0.f;
    for (int l = 0; l < lanes; ++l) tot += sm[(l * cg + c / 8) * 8 + (c % 8)];
    if (tot != 0.f) atomicAdd(dw + (long long)tap * cpad + c, tot);
  }
}

static int dw_check(const PmoeView4* v, const char* what) {
  if (!v || !v->ptr || v->c <= 0 || v->c % 8 || v->n <= 0 || v->h <= 0 || v->w <= 0) {
    set_error("%s: needs a non-empty NHWC view with a multiple of 8 channels", what);
    return PMOE_ERR_ARG;
  }
  return PMOE_OK;
}

static int dw_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace pmoe

using namespace pmoe;

#define DW_DISPATCH(dtype, ...)                      \
  if ((dtype) == PMOE_BF16) {                        \
    using T = __nv_bfloat16;                         \
    __VA_ARGS__;                                     \
  } else if ((dtype) == PMOE_F32) {                  \
    using T = float;                                 \
    __VA_ARGS__;                                     \
  } else {                                           \
    set_error("unsupported dtype %d", (int)(dtype)); \
    return PMOE_ERR_ARG;                             \
  }

static int dw_geometry(const PmoeView4* x, const PmoeView4* y, int k, int stride, int pad, const char* what) {
  if (k < 1 || k > 7 || stride < 1 || stride > 2 || pad < 0 || pad >= k) {
    set_error("%s: kernel 1..7, stride 1..2, pad < kernel", what);
    return PMOE_ERR_ARG;
  }
  const int oh = (x->h + 2 * pad - k) / stride + 1, ow = (x->w + 2 * pad - k) / stride + 1;
  if (y->n != x->n || y->c != x->c || y->h != oh || y->w != ow) {
    set_error("%s: output view (%d,%d,%d,%d) does not match the input geometry (%d,%d,%d,%d), k=%d stride=%d pad=%d", what, y->n, y->h,
              y->w, y->c, x->n, oh, ow, x->c, k, stride, pad);
    return PMOE_ERR_ARG;
  }
  return PMOE_OK;
}

extern "C" int pmoe_dwconv_fwd(const PmoeView4* x, const float* w_packed, int32_t w_cpad, const PmoeView4* y, int32_t dtype, int32_t k,
                               int32_t stride, int32_t pad, pmoe_stream_t stream_) {
  int rc;
  if ((rc = dw_check(x, "dwconv_fwd x")) || (rc = dw_check(y, "dwconv_fwd y")) || (rc = dw_geometry(x, y, k, stride, pad, "dwconv_fwd"))) return rc;
  if (!w_packed || w_cpad < y->c || ((uintptr_t)w_packed % 16)) {
    set_error("dwconv_fwd: packed weights [k*k][cpad >= channels], 16-byte aligned");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)y->n * y->h * y->w * (y->c / 8);
  DW_DISPATCH(dtype, (dwconv_fwd_kernel<T><<<dw_grid(items), 256, 0, static_cast<cudaStream_t>(stream_)>>>(dv4(x), dv4(y), w_packed, w_cpad, k,
                                                                                                        stride, pad)));
  return check_launch("dwconv_fwd");
}

extern "C" int pmoe_dwconv_dgrad(const PmoeView4* dy, const float* w_packed, int32_t w_cpad, const PmoeView4* dx, int32_t dtype, int32_t k,
                                 int32_t stride, int32_t pad, int32_t accumulate, pmoe_stream_t stream_) {
  int rc;
  if ((rc = dw_check(dy, "dwconv_dgrad dy")) || (rc = dw_check(dx, "dwconv_dgrad dx")) || (rc = dw_geometry(dx, dy, k, stride, pad, "dwconv_dgrad")))
    return rc;
  if (!w_packed || w_cpad < dx->c || ((uintptr_t)w_packed % 16)) {
    set_error("dwconv_dgrad: packed weights [k*k][cpad >= channels], 16-byte aligned");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dx->n * dx->h * dx->w * (dx->c / 8);
  DW_DISPATCH(dtype, (dwconv_dgrad_kernel<T><<<dw_grid(items), 256, 0, static_cast<cudaStream_t>(stream_)>>>(dv4(dy), dv4(dx), w_packed, w_cpad, k,
                                                                                                          stride, pad, accumulate)));
  return check_launch("dwconv_dgrad");
}

extern "C" int pmoe_dwconv_wgrad(const PmoeView4* x, const PmoeView4* dy, float* dw_packed, int32_t w_cpad, int32_t dtype, int32_t k,
                                 int32_t stride, int32_t pad, pmoe_stream_t stream_) {
  int rc;
  if ((rc = dw_check(x, "dwconv_wgrad x")) || (rc = dw_check(dy, "dwconv_wgrad dy")) || (rc = dw_geometry(x, dy, k, stride, pad, "dwconv_wgrad")))
    return rc;
  const int cg = dy->c / 8;
  if (!dw_packed || w_cpad < dy->c || cg > 256) {
    set_error("dwconv_wgrad: packed gradient [k*k][cpad >= channels], at most 2048 channels");
    return PMOE_ERR_ARG;
  }
  const long long npix = (long long)dy->n * dy->h * dy->w;
  long long blocks = (long long)num_sms() * 4;
  long long ppb = (npix + blocks - 1) / blocks;
  if (ppb < 32) ppb = 32;
  blocks = (npix + ppb - 1) / ppb;
  const int threads = (256 / cg) * cg;  // lanes x channel groups, at most 256
  dim3 grid((unsigned)blocks, (unsigned)(k * k));
  DW_DISPATCH(dtype, (dwconv_wgrad_kernel<T><<<grid, threads, 0, static_cast<cudaStream_t>(stream_)>>>(dv4(x), dv4(dy), dw_packed, w_cpad, k, stride,
                                                                                                    pad, ppb)));
  return check_launch("dwconv_wgrad");
}
