// Memory-bound (HBM roofline) kernels of the hot path: layout conversion at the module boundary,
// max-pool, ECA gate + channel scale, BatchNorm finalize/apply (+residual, +ReLU). All operate on
// strided NHWC views, 8 channels (16 B of bf16 / 32 B of fp32) per thread, grid-stride loops sized
// to a multiple of the SM count.
#include "host_util.h"
#include "act.cuh"
#include "ptx.cuh"
#include "reduce.cuh"

namespace pmoe {

// ---------------------------------------------------------------- 8-channel vector access
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

struct V4 {
  void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
};
static V4 to_v4(const PmoeView4& v) { return V4{v.ptr, v.n, v.h, v.w, v.c, v.sn, v.sh, v.sw}; }

static inline int grid_for(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static int check_view(const PmoeView4* v, int dtype, const char* what) {
  const int esz = dtype == PMOE_BF16 ? 2 : 4;
  if (!v || !v->ptr || v->c % 8 || ((uintptr_t)v->ptr % 16) || (v->sw * esz) % 16 || (v->sh * esz) % 16 || (v->sn * esz) % 16) {
    set_error("%s: view must be 16-byte aligned, channels a multiple of 8", what);
    return PMOE_ERR_ARG;
  }
  return PMOE_OK;
}

// ---------------------------------------------------------------- NCHW fp32 -> NHWC (zero-padded channels)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, long long sn, long long sc, long long sh, long long sw,
                                    int c, V4 dst) {
  const int cg = dst.c / 8;
  const long long total = (long long)dst.n * dst.h * dst.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // pixel-major over threads: consecutive threads read consecutive w of one channel plane
    const int g = (int)(i / ((long long)dst.n * dst.h * dst.w));
    long long pix = i % ((long long)dst.n * dst.h * dst.w);
    const int w = (int)(pix % dst.w);
    pix /= dst.w;
    const int h = (int)(pix % dst.h);
    const int n = (int)(pix / dst.h);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      v[k] = ch < c ? __ldg(src + n * sn + ch * sc + h * sh + w * sw) : 0.f;
    }
    store8(static_cast<T*>(dst.ptr) + n * dst.sn + h * dst.sh + w * dst.sw + g * 8, v);
  }
}

// Narrow destinations (<= 32 stored channels: the 3- and 12-channel image inputs): one thread per PIXEL writes all of its
// channel groups, so that a warp's stores are 32 consecutive 32/64-byte pixel rows (fully written sectors) instead of 16-byte
// halves of sectors completed by another warp much later; loads stay coalesced along w within each channel plane.
template <typename T, int CG>
__global__ void __launch_bounds__(256) nchw_to_nhwc_pix_kernel(const float* __restrict__ src, long long sn, long long sc, long long sh,
                                                               long long sw, int c, V4 dst) {
  const long long total = (long long)dst.n * dst.h * dst.w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long pix = i;
    const int w = (int)(pix % dst.w);
    pix /= dst.w;
    const int h = (int)(pix % dst.h);
    const int n = (int)(pix / dst.h);
    const float* sp = src + n * sn + h * sh + w * sw;
    float v[CG][8];
#pragma unroll
    for (int g = 0; g < CG; ++g)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[g][k] = (g * 8 + k < c) ? __ldg(sp + (g * 8 + k) * sc) : 0.f;
    T* dp = static_cast<T*>(dst.ptr) + n * dst.sn + h * dst.sh + w * dst.sw;
#pragma unroll
    for (int g = 0; g < CG; ++g) store8(dp + g * 8, v[g]);
  }
}

// ---------------------------------------------------------------- NHWC -> NCHW fp32 (first c channels)
template <typename T>
__global__ void nhwc_to_nchw_kernel(V4 src, int c, float* __restrict__ dst, long long dn, long long dc, long long dh,
                                    long long dw) {
  const int cg = (c + 7) / 8;
  const long long npix = (long long)src.n * src.h * src.w;
  const long long total = npix * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i / npix);
    long long pix = i % npix;
    const int w = (int)(pix % src.w);
    pix /= src.w;
    const int h = (int)(pix % src.h);
    const int n = (int)(pix / src.h);
    float v[8];
    load8(static_cast<const T*>(src.ptr) + n * src.sn + h * src.sh + w * src.sw + g * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = g * 8 + k;
      if (ch < c) dst[n * dn + ch * dc + h * dh + w * dw] = v[k];
    }
  }
}

// ---------------------------------------------------------------- max-pool (optional affine+ReLU on load)
template <typename T>
__global__ void maxpool_kernel(V4 src, V4 dst, int k, int stride, int pad, const float* __restrict__ scale,
                               const float* __restrict__ shift, int relu, uint8_t* __restrict__ idx_out) {
  const int cg = dst.c / 8;
  const long long total = (long long)dst.n * dst.h * dst.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int ow = (int)(pix % dst.w);
    pix /= dst.w;
    const int oh = (int)(pix % dst.h);
    const int n = (int)(pix / dst.h);
    float sc[8], sf[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      sc[q] = scale ? __ldg(scale + g * 8 + q) : 1.f;
      sf[q] = shift ? __ldg(shift + g * 8 + q) : 0.f;
    }
    float m[8];
    uint32_t arg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      m[q] = -INFINITY;
      arg[q] = 0;
    }
    for (int r = 0; r < k; ++r) {
      const int ih = oh * stride - pad + r;
      if (ih < 0 || ih >= src.h) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = ow * stride - pad + s;
        if (iw < 0 || iw >= src.w) continue;
        float v[8];
        load8(static_cast<const T*>(src.ptr) + n * src.sn + ih * src.sh + iw * src.sw + g * 8, v);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float x = fmaf(v[q], sc[q], sf[q]);
          if (relu) x = fmaxf(x, 0.f);
          if (x > m[q]) {  // the FIRST maximum in row-major window order wins (ATen max_pool2d tie rule)
            m[q] = x;
            arg[q] = (uint32_t)(r * k + s);
          }
        }
      }
    }
    store8(static_cast<T*>(dst.ptr) + n * dst.sn + oh * dst.sh + ow * dst.sw + g * 8, m);
    if (idx_out) {
      uint2 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      *reinterpret_cast<uint2*>(idx_out + (((long long)n * dst.h + oh) * dst.w + ow) * dst.c + g * 8) = pk;
    }
  }
}

// Contiguous bf16, compile-time window: one block row per (image, output row), 32-bit index arithmetic only, all K*K window
// loads issued before the first compare. Same tie rule and argmax codes as maxpool_kernel.
__device__ __forceinline__ void pool_bf16x8(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[2 * q] = __uint_as_float(w[q] << 16);
    v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
template <int K, int S, int P, bool IDX>
__global__ void __launch_bounds__(256) maxpool_fast_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint2* __restrict__ idx_out,
                                                           int H, int W, int OH, int OW, int cg, int cg_shift) {
  const int n = blockIdx.x / OH, oh = blockIdx.x - n * OH;
  const int row_items = OW * cg;
  const uint4* img = src + (size_t)n * H * W * cg;
  for (int item = blockIdx.y * blockDim.x + threadIdx.x; item < row_items; item += gridDim.y * blockDim.x) {
    const int ow = cg_shift >= 0 ? (item >> cg_shift) : item / cg;
    const int g = item - ow * cg;
    uint4 raw[K * K];
    bool ok[K * K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int ih = oh * S - P + r;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const int iw = ow * S - P + c;
        ok[r * K + c] = ih >= 0 && ih < H && iw >= 0 && iw < W;
        if (ok[r * K + c]) raw[r * K + c] = __ldg(img + ((size_t)ih * W + iw) * cg + g);
      }
    }
    float m[8];
    uint32_t arg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      m[q] = -INFINITY;
      arg[q] = 0;
    }
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      if (!ok[t]) continue;
      float v[8];
      pool_bf16x8(raw[t], v);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (v[q] > m[q]) {  // the FIRST maximum in row-major window order wins (ATen max_pool2d tie rule)
          m[q] = v[q];
          arg[q] = (uint32_t)t;
        }
    }
    const size_t o = ((size_t)blockIdx.x * OW + ow) * cg + g;
    uint4 pr;
    pr.x = pack_bf16x2(m[0], m[1]);
    pr.y = pack_bf16x2(m[2], m[3]);
    pr.z = pack_bf16x2(m[4], m[5]);
    pr.w = pack_bf16x2(m[6], m[7]);
    dst[o] = pr;
    if (IDX) {
      uint2 pk;
      pk.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
      pk.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
      idx_out[o] = pk;
    }
  }
}

// ---------------------------------------------------------------- ECA gate: sigmoid(conv1d_k(mean over HW))
// Channels may be stored as `groups` blocks of `group_c` logical channels every `group_stride`
// physical channels (the PU-Net mask ring stores 23 logits in 32-channel slots); the 1-D conv runs
// over the LOGICAL channel axis with zero padding k/2 (basics.py:69,73).
__global__ void eca_gate_kernel(const float* __restrict__ pool_sum, long long pool_stride, float inv_count,
                                const float* __restrict__ w, int k, int groups, int group_c, int group_stride,
                                float* __restrict__ gate, long long gate_stride) {
  const int n = blockIdx.x;
  const int L = groups * group_c;
  for (int pc = threadIdx.x; pc < groups * group_stride; pc += blockDim.x) {
    const int g = pc / group_stride, j = pc % group_stride;
    float out = 0.f;
    if (j < group_c) {
      const int l = g * group_c + j;
      float acc = 0.f;
      for (int t = 0; t < k; ++t) {
        const int ll = l + t - k / 2;
        if (ll >= 0 && ll < L) {
          const int p = (ll / group_c) * group_stride + (ll % group_c);
          acc = fmaf(__ldg(w + t), __ldg(pool_sum + n * pool_stride + p) * inv_count, acc);
        }
      }
      out = 1.f / (1.f + expf(-acc));
    }
    gate[n * gate_stride + pc] = out;
  }
}

template <typename T>
__global__ void scale_channels_kernel(V4 src, V4 dst, const float* __restrict__ gate, long long gate_stride) {
  const int cg = dst.c / 8;
  const long long total = (long long)dst.n * dst.h * dst.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int w = (int)(pix % dst.w);
    pix /= dst.w;
    const int h = (int)(pix % dst.h);
    const int n = (int)(pix / dst.h);
    float v[8], s[8];
    load8(static_cast<const T*>(src.ptr) + n * src.sn + h * src.sh + w * src.sw + g * 8, v);
    load8(gate + n * gate_stride + g * 8, s);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] *= s[q];
    store8(static_cast<T*>(dst.ptr) + n * dst.sn + h * dst.sh + w * dst.sw + g * 8, v);
  }
}

// ---------------------------------------------------------------- per-(n,c) sums over HW (ECA / global avg-pool)
template <typename T>
__global__ void __launch_bounds__(kRedThreads) channel_sums_kernel(V4 src, float* __restrict__ out, long long out_stride, int rows_per_block) {
  // grid: (ceil(h*w / rows_per_block), n); block: 256 threads = cg channel groups x (256/cg) pixel lanes
  __shared__ float sm[kRedThreads * 8];
  const int cg = src.c / 8;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const long long hw = (long long)src.h * src.w;
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  long long p1 = p0 + rows_per_block;
  if (p1 > hw) p1 = hw;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const int h = (int)(p / src.w), w = (int)(p % src.w);
      float v[8];
      load8(static_cast<const T*>(src.ptr) + n * src.sn + h * src.sh + w * src.sw + g * 8, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += v[q];
    }
  }
  float tot[kRedMaxIter];
  block_channel_sum(acc, sm, cg, lanes, tot);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8 && tot[j] != 0.f) atomicAdd(out + n * out_stride + c, tot[j]);
  }
}

// ---------------------------------------------------------------- per-channel sum / sum of squares over N,H,W
// (batch statistics for a BatchNorm that does not directly follow one of our conv epilogues: ResNet bn1)
template <typename T>
__global__ void __launch_bounds__(kRedThreads) channel_stats_kernel(V4 src, double* __restrict__ sum, double* __restrict__ sq, long long pix_per_block) {
  __shared__ float sm[kRedThreads * 8];
  const int cg = src.c / 8;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const long long npix = (long long)src.n * src.h * src.w;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const int w = (int)(p % src.w);
      const long long t = p / src.w;
      const int h = (int)(t % src.h), n = (int)(t / src.h);
      float v[8];
      load8(static_cast<const T*>(src.ptr) + n * src.sn + h * src.sh + w * src.sw + g * 8, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        a[q] += v[q];
        b[q] += v[q] * v[q];
      }
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  block_channel_sum(b, sm, cg, lanes, tb);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      atomicAdd(sum + c, (double)ta[j]);
      atomicAdd(sq + c, (double)tb[j]);
    }
  }
}

// ---------------------------------------------------------------- ECA gate folded into per-image conv weights
// out[n][co][k] = w[co][k] * gate[n][k % cphys]: x * gate followed by a conv equals the conv with per-image weights whose
// input-channel columns carry the gate, so the full-resolution "scale the tensor" pass (1 read + 1 write of a 224^2
// activation) becomes a rewrite of a few hundred KB of weights.
__global__ void gate_weights_kernel(const __nv_bfloat16* __restrict__ w, const float* __restrict__ gate, long long gate_stride,
                                    int cphys, long long per_img, int ktot, __nv_bfloat16* __restrict__ out) {
  const int n = blockIdx.y;
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 8; i < per_img; i += (long long)gridDim.x * blockDim.x * 8) {
    const int k = (int)(i % ktot);  // ktot % 8 == 0 and cphys % 8 == 0: the 8 elements share a row and a gate run
    float v[8], g[8];
    load8(w + i, v);
    load8(gate + n * gate_stride + (k % cphys), g);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] *= g[q];
    store8(out + (long long)n * per_img + i, v);
  }
}

// ---------------------------------------------------------------- BatchNorm (train): finalize + apply
// Finalize: batch mean / biased var from (sum, sumsq), running-stat update exactly as
// nn.BatchNorm2d (momentum 0.1, unbiased variance), and the fused affine (scale, shift).
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sq, float count, int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ scale,
                                   float* __restrict__ shift, int c_pad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_pad) return;
  if (i >= c) {
    if (scale) scale[i] = 0.f;
    if (shift) shift[i] = 0.f;
    if (mean_out) mean_out[i] = 0.f;
    if (rstd_out) rstd_out[i] = 0.f;
    return;
  }
  const double m = sum[i] / (double)count;
  double var = sq[i] / (double)count - m * m;
  if (var < 0) var = 0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * (float)m;
    const double unbiased = count > 1.f ? var * count / (count - 1.0) : var;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
  const float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  if (mean_out) mean_out[i] = (float)m;
  if (rstd_out) rstd_out[i] = rstd;
  if (scale) scale[i] = g * rstd;
  if (shift) shift[i] = b - (float)m * g * rstd;
}

// y = act(scale[c]*x + shift[c] (+ residual))
template <typename T, bool FLAT>
__global__ void __launch_bounds__(256) affine_act_kernel(V4 src, V4 dst, const float* __restrict__ scale, const float* __restrict__ shift,
                                  V4 res, int act) {
  const int cg = dst.c / 8;
  const long long total = (long long)dst.n * dst.h * dst.w * cg;
  const long long stride = (long long)gridDim.x * blockDim.x;  // a multiple of cg: one channel group per thread
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = (int)(first % cg);
  float sc[8], sf[8];
  load8(scale + g * 8, sc);
  load8(shift + g * 8, sf);
#pragma unroll 2
  for (long long i = first; i < total; i += stride) {
    long long o_s, o_d, o_r;
    if (FLAT) {
      o_s = o_d = o_r = i * 8;
    } else {
      long long pix = i / cg;
      const int w = (int)(pix % dst.w);
      pix /= dst.w;
      const int h = (int)(pix % dst.h);
      const int n = (int)(pix / dst.h);
      o_s = n * src.sn + h * src.sh + w * src.sw + g * 8;
      o_d = n * dst.sn + h * dst.sh + w * dst.sw + g * 8;
      o_r = n * res.sn + h * res.sh + w * res.sw + g * 8;
    }
    float v[8];
    load8(static_cast<const T*>(src.ptr) + o_s, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = fmaf(v[q], sc[q], sf[q]);
    if (res.ptr) {
      float r[8];
      load8(static_cast<const T*>(res.ptr) + o_r, r);
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] += r[q];
    }
    if (act == PMOE_ACT_RELU) {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
    } else if (act != PMOE_ACT_NONE) {
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = act_piecewise(v[q], act);
    }
    store8(static_cast<T*>(dst.ptr) + o_d, v);
  }
}


// ---------------------------------------------------------------- fast paths: dense bf16, 4 x 16-byte loads in flight per thread
// The generic reduction kernels above issue one load per loop trip behind a 64-bit div/mod and sit at 40-55 % of the HBM
// peak; these are their contiguous-bf16 forms (straight-line, no index arithmetic).
// MODE 0: per-image channel sums (fp32), MODE 1: per-channel sum and sum of squares over every pixel (fp64).
// AFFINE: y = act(scale*x + shift) is computed, stored, and the statistics are those of the STORED (bf16-rounded) y —
// the fused form of affine_act followed by channel_sums (ECA / avg-pool numerators) and/or channel_stats (a BatchNorm that
// follows directly: ResNet bn1 after the stem block), which would each re-read the tensor.
// COUNT: also the number of stored values > 0 per channel (a ReLU output feeding a BatchNorm whose backward is taken in closed
// form, see stem_tail.cu bn2_relu_maxpool_bwd_apply_kernel).
template <bool AFFINE, bool RELU, bool POOL, bool STATS, bool COUNT = false>
__global__ void __launch_bounds__(kRedThreads) act_reduce_fast_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long pix_per_img,
                                                                      int cg, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, float* __restrict__ pool,
                                                                      long long pool_stride, double* __restrict__ osum,
                                                                      double* __restrict__ osq, long long pix_per_block,
                                                                      double* __restrict__ opos) {
  __shared__ float sm[kRedThreads * 8];
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const uint4* xi = x + (size_t)n * pix_per_img * cg;
  uint4* yi = AFFINE ? y + (size_t)n * pix_per_img * cg : nullptr;
  float sc[8], sf[8];
  if (AFFINE) {
    load8(scale + g * 8, sc);
    load8(shift + g * 8, sf);
  }
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > pix_per_img) p1 = pix_per_img;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0}, cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += 4LL * lanes) {
      uint4 r[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pp = p + (long long)u * lanes;
        ok[u] = pp < p1;
        r[u] = ok[u] ? __ldg(xi + pp * cg + g) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!ok[u]) continue;
        float v[8];
        pool_bf16x8(r[u], v);
        if (AFFINE) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            v[q] = fmaf(v[q], sc[q], sf[q]);
            if (RELU) v[q] = fmaxf(v[q], 0.f);
          }
          uint4 o;
          o.x = pack_bf16x2(v[0], v[1]);
          o.y = pack_bf16x2(v[2], v[3]);
          o.z = pack_bf16x2(v[4], v[5]);
          o.w = pack_bf16x2(v[6], v[7]);
          yi[(p + (long long)u * lanes) * cg + g] = o;
          pool_bf16x8(o, v);  // statistics of what is stored
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          a[q] += v[q];
          if (STATS) b[q] = fmaf(v[q], v[q], b[q]);
          if (COUNT) cnt[q] += v[q] > 0.f ? 1.f : 0.f;   // exact in fp32: a thread sees far fewer than 2^24 pixels
        }
      }
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter], tc[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  if (STATS) block_channel_sum(b, sm, cg, lanes, tb);
  if (COUNT) block_channel_sum(cnt, sm, cg, lanes, tc);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      if (POOL && ta[j] != 0.f) atomicAdd(pool + n * pool_stride + c, ta[j]);
      if (STATS) {
        atomicAdd(osum + c, (double)ta[j]);
        atomicAdd(osq + c, (double)tb[j]);
      }
      if (COUNT) atomicAdd(opos + c, (double)tc[j]);
    }
  }
}

// y = x * gate[n, c] on dense bf16: grid (x, n), the gate of the thread's channel group in registers, four loads in flight
__global__ void __launch_bounds__(256) scale_channels_fast_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long per_img,
                                                                  int cg, const float* __restrict__ gate, long long gate_stride) {
  const int n = blockIdx.y;
  const long long stride = (long long)gridDim.x * blockDim.x;  // a multiple of cg
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = (int)(first % cg);
  float gt[8];
  load8(gate + n * gate_stride + g * 8, gt);
  const uint4* src = x + (size_t)n * per_img;
  uint4* dst = y + (size_t)n * per_img;
  for (long long i = first; i < per_img; i += 4 * stride) {
    uint4 r[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ii = i + (long long)u * stride;
      ok[u] = ii < per_img;
      r[u] = __ldg(src + (ok[u] ? ii : i));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      float v[8];
      pool_bf16x8(r[u], v);
      uint4 o;
      o.x = pack_bf16x2(v[0] * gt[0], v[1] * gt[1]);
      o.y = pack_bf16x2(v[2] * gt[2], v[3] * gt[3]);
      o.z = pack_bf16x2(v[4] * gt[4], v[5] * gt[5]);
      o.w = pack_bf16x2(v[6] * gt[6], v[7] * gt[7]);
      dst[i + (long long)u * stride] = o;
    }
  }
}

static bool dense_view(const PmoeView4* v) {
  return v && v->ptr && v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c;
}
// grid (x blocks per image, images): ~8 CTAs per SM overall, at least 64 pixels per block
static dim3 reduce_grid(long long pix_per_img, int n, long long* ppb_out) {
  long long want = ((long long)num_sms() * 8 + n - 1) / n;
  long long ppb = (pix_per_img + want - 1) / want;
  if (ppb < 64) ppb = 64;
  *ppb_out = ppb;
  return dim3((unsigned)((pix_per_img + ppb - 1) / ppb), (unsigned)n);
}


}  // namespace pmoe

using namespace pmoe;

#define DISPATCH_DTYPE(dtype, ...)                       \
  if ((dtype) == PMOE_BF16) {                            \
    using T = __nv_bfloat16;                             \
    __VA_ARGS__;                                         \
  } else if ((dtype) == PMOE_F32) {                      \
    using T = float;                                     \
    __VA_ARGS__;                                         \
  } else {                                               \
    set_error("unsupported dtype %d", (int)(dtype));     \
    return PMOE_ERR_ARG;                                 \
  }

extern "C" {

int pmoe_nchw_to_nhwc(const float* src, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int32_t c, const PmoeView4* dst,
                      int32_t dst_dtype, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(dst, dst_dtype, "nchw_to_nhwc");
  if (rc) return rc;
  if (!src || c > dst->c) {
    set_error("nchw_to_nhwc: bad source (c %d, dst.c %d)", c, dst ? dst->c : -1);
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dst->n * dst->h * dst->w * (dst->c / 8);
  const long long pixels = (long long)dst->n * dst->h * dst->w;
  if (dst->c == 16) {
    DISPATCH_DTYPE(dst_dtype, (nchw_to_nhwc_pix_kernel<T, 2><<<grid_for(pixels, 256), 256, 0, stream>>>(src, sn, sc, sh, sw, c, to_v4(*dst))));
    return check_launch("nchw_to_nhwc");
  }
  if (dst->c == 32) {
    DISPATCH_DTYPE(dst_dtype, (nchw_to_nhwc_pix_kernel<T, 4><<<grid_for(pixels, 256), 256, 0, stream>>>(src, sn, sc, sh, sw, c, to_v4(*dst))));
    return check_launch("nchw_to_nhwc");
  }
  DISPATCH_DTYPE(dst_dtype, (nchw_to_nhwc_kernel<T><<<grid_for(items, 256), 256, 0, stream>>>(src, sn, sc, sh, sw, c, to_v4(*dst))));
  return check_launch("nchw_to_nhwc");
}

int pmoe_nhwc_to_nchw(const PmoeView4* src, int32_t src_dtype, int32_t c, float* dst, int64_t dn, int64_t dc, int64_t dh,
                      int64_t dw, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, src_dtype, "nhwc_to_nchw");
  if (rc) return rc;
  if (!dst || c > src->c) {
    set_error("nhwc_to_nchw: bad destination");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)src->n * src->h * src->w * ((c + 7) / 8);
  DISPATCH_DTYPE(src_dtype, (nhwc_to_nchw_kernel<T><<<grid_for(items, 256), 256, 0, stream>>>(to_v4(*src), c, dst, dn, dc, dh, dw)));
  return check_launch("nhwc_to_nchw");
}

static int maxpool_launch(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, int32_t k, int32_t stride, int32_t pad,
                          const float* scale, const float* shift, int32_t relu, uint8_t* idx_out, cudaStream_t stream) {
  int rc = check_view(src, dtype, "maxpool src");
  if (rc) return rc;
  if ((rc = check_view(dst, dtype, "maxpool dst"))) return rc;
  if (src->c < dst->c || src->n != dst->n || k < 1 || stride < 1 || k > 15 || ((uintptr_t)idx_out & 7)) {
    set_error("maxpool: geometry mismatch");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dst->n * dst->h * dst->w * (dst->c / 8);
  const auto flat = [](const PmoeView4* v) {
    return v->sw == v->c && v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c;
  };
  const bool k3 = k == 3 && stride == 2 && pad == 1, k2 = k == 2 && stride == 2 && pad == 0;
  if (dtype == PMOE_BF16 && !scale && !shift && !relu && (k3 || k2) && flat(src) && flat(dst) && src->c == dst->c &&
      (long long)dst->n * dst->h < 2147483647LL && (long long)src->h * src->w * src->c < 2147483647LL) {
    const int cg = dst->c / 8;
    int cg_shift = -1;
    for (int sft = 0; sft < 12; ++sft)
      if ((1 << sft) == cg) cg_shift = sft;
    const dim3 grid((unsigned)(dst->n * dst->h), (unsigned)((dst->w * cg + 255) / 256));
    const uint4* ps = static_cast<const uint4*>(src->ptr);
    uint4* pd = static_cast<uint4*>(dst->ptr);
    uint2* pi = reinterpret_cast<uint2*>(idx_out);
#define PMOE_POOL_FAST(K, S, P, I) maxpool_fast_kernel<K, S, P, I><<<grid, 256, 0, stream>>>(ps, pd, pi, src->h, src->w, dst->h, dst->w, cg, cg_shift)
    if (k3 && idx_out) PMOE_POOL_FAST(3, 2, 1, true);
    else if (k3) PMOE_POOL_FAST(3, 2, 1, false);
    else if (idx_out) PMOE_POOL_FAST(2, 2, 0, true);
    else PMOE_POOL_FAST(2, 2, 0, false);
#undef PMOE_POOL_FAST
    return check_launch("maxpool");
  }
  DISPATCH_DTYPE(dtype, (maxpool_kernel<T><<<grid_for(items, 256), 256, 0, stream>>>(to_v4(*src), to_v4(*dst), k, stride, pad, scale, shift, relu, idx_out)));
  return check_launch("maxpool");
}

int pmoe_maxpool(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, int32_t k, int32_t stride, int32_t pad,
                 const float* scale, const float* shift, int32_t relu, pmoe_stream_t stream_) {
  return maxpool_launch(src, dst, dtype, k, stride, pad, scale, shift, relu, nullptr, static_cast<cudaStream_t>(stream_));
}

int pmoe_maxpool_idx(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, int32_t k, int32_t stride, int32_t pad,
                     uint8_t* idx_out, pmoe_stream_t stream_) {
  if (!idx_out) {
    set_error("maxpool_idx: index buffer missing");
    return PMOE_ERR_ARG;
  }
  return maxpool_launch(src, dst, dtype, k, stride, pad, nullptr, nullptr, 0, idx_out, static_cast<cudaStream_t>(stream_));
}

int pmoe_eca_gate(const float* pool_sum, int64_t pool_stride, int32_t n, float inv_count, const float* w, int32_t k,
                  int32_t groups, int32_t group_c, int32_t group_stride, float* gate, int64_t gate_stride,
                  pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!pool_sum || !w || !gate || n < 1 || k < 1 || groups < 1 || group_c > group_stride) {
    set_error("eca_gate: bad arguments");
    return PMOE_ERR_ARG;
  }
  eca_gate_kernel<<<n, 128, 0, stream>>>(pool_sum, pool_stride, inv_count, w, k, groups, group_c, group_stride, gate, gate_stride);
  return check_launch("eca_gate");
}

int pmoe_gate_weights(const void* wpack, const float* gate, int64_t gate_stride, int32_t n, int32_t cout_pad, int32_t ktot,
                      int32_t cphys, void* out, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!wpack || !gate || !out || n < 1 || cout_pad < 1 || ktot % 8 || cphys % 8 || ktot % cphys || gate_stride % 4 ||
      ((uintptr_t)wpack & 15) || ((uintptr_t)out & 15) || ((uintptr_t)gate & 15)) {
    set_error("gate_weights: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long per_img = (long long)cout_pad * ktot;
  dim3 grid((unsigned)((per_img / 8 + 255) / 256), (unsigned)n);
  if (grid.x > 1024) grid.x = 1024;
  gate_weights_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(wpack), gate, gate_stride, cphys, per_img, ktot,
                                               static_cast<__nv_bfloat16*>(out));
  return check_launch("gate_weights");
}

int pmoe_scale_channels(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* gate, int64_t gate_stride,
                        pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, dtype, "scale_channels src");
  if (rc) return rc;
  if ((rc = check_view(dst, dtype, "scale_channels dst"))) return rc;
  if (!gate || ((uintptr_t)gate % 16) || gate_stride % 4 || src->n != dst->n || src->h != dst->h || src->w != dst->w || src->c < dst->c) {
    set_error("scale_channels: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dst->n * dst->h * dst->w * (dst->c / 8);
  if (dtype == PMOE_BF16 && dense_view(src) && dense_view(dst) && src->c == dst->c && dst->n <= 65535) {
    const int cg = dst->c / 8;
    const long long per_img = (long long)dst->h * dst->w * cg;
    long long bx = (per_img + 4 * 256 - 1) / (4 * 256);
    const long long cap = ((long long)num_sms() * 16 + dst->n - 1) / dst->n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (256 % cg != 0) bx = (bx + cg - 1) / cg * cg;
    scale_channels_fast_kernel<<<dim3((unsigned)bx, (unsigned)dst->n), 256, 0, stream>>>(
        static_cast<const uint4*>(src->ptr), static_cast<uint4*>(dst->ptr), per_img, cg, gate, gate_stride);
    return check_launch("scale_channels");
  }
  DISPATCH_DTYPE(dtype, (scale_channels_kernel<T><<<grid_for(items, 256), 256, 0, stream>>>(to_v4(*src), to_v4(*dst), gate, gate_stride)));
  return check_launch("scale_channels");
}

int pmoe_channel_sums(const PmoeView4* src, int32_t dtype, float* out, int64_t out_stride, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, dtype, "channel_sums");
  if (rc) return rc;
  const int cg = src->c / 8;
  if (!out || cg > 256) {
    set_error("channel_sums: needs an output buffer and at most 2048 channels");
    return PMOE_ERR_ARG;
  }
  const long long hw = (long long)src->h * src->w;
  if (dtype == PMOE_BF16 && dense_view(src)) {
    long long ppb;
    const dim3 grid = reduce_grid(hw, src->n, &ppb);
    act_reduce_fast_kernel<false, false, true, false><<<grid, kRedThreads, 0, stream>>>(
        static_cast<const uint4*>(src->ptr), nullptr, hw, cg, nullptr, nullptr, out, out_stride, nullptr, nullptr, ppb, nullptr);
    return check_launch("channel_sums");
  }
  // fp32 (the parity mode): ONE block per image walks all its pixels, so the per-image sums do not depend on the order in which
  // several blocks' fp32 atomics land (they feed the ECA gates: a 1e-7 difference flips ReLU masks downstream)
  long long rows_ll = dtype == PMOE_F32 ? hw : 1024;
  if (rows_ll > 0x7fffffffLL) rows_ll = 0x7fffffffLL;
  const int rows = (int)rows_ll;
  dim3 grid((unsigned)((hw + rows - 1) / rows), (unsigned)src->n);
  DISPATCH_DTYPE(dtype, (channel_sums_kernel<T><<<grid, 256, 0, stream>>>(to_v4(*src), out, out_stride, rows)));
  return check_launch("channel_sums");
}

int pmoe_channel_stats(const PmoeView4* src, int32_t dtype, double* sum, double* sqsum, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, dtype, "channel_stats");
  if (rc) return rc;
  if (!sum || !sqsum || src->c / 8 > 256) {
    set_error("channel_stats: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long npix = (long long)src->n * src->h * src->w;
  if (dtype == PMOE_BF16 && dense_view(src)) {
    long long ppb1;
    const dim3 grid = reduce_grid(npix, 1, &ppb1);  // dense: the batch is one long image
    act_reduce_fast_kernel<false, false, false, true><<<grid, kRedThreads, 0, stream>>>(
        static_cast<const uint4*>(src->ptr), nullptr, npix, src->c / 8, nullptr, nullptr, nullptr, 0, sum, sqsum, ppb1, nullptr);
    return check_launch("channel_stats");
  }
  long long blocks = (long long)num_sms() * 8;
  long long ppb = (npix + blocks - 1) / blocks;
  if (ppb < 64) ppb = 64;
  blocks = (npix + ppb - 1) / ppb;
  DISPATCH_DTYPE(dtype, (channel_stats_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>(to_v4(*src), sum, sqsum, ppb)));
  return check_launch("channel_stats");
}

int pmoe_bn_finalize(const double* sum, const double* sqsum, float count, int32_t c, int32_t c_pad, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* mean_out,
                     float* rstd_out, float* scale, float* shift, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!sum || !sqsum || c < 1 || c_pad < c || count <= 0) {
    set_error("bn_finalize: bad arguments");
    return PMOE_ERR_ARG;
  }
  bn_finalize_kernel<<<(c_pad + 127) / 128, 128, 0, stream>>>(sum, sqsum, count, c, gamma, beta, eps, momentum, running_mean,
                                                              running_var, mean_out, rstd_out, scale, shift, c_pad);
  return check_launch("bn_finalize");
}

int pmoe_affine_act(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* scale, const float* shift,
                    const PmoeView4* residual, int32_t act, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, dtype, "affine_act src");
  if (rc) return rc;
  if ((rc = check_view(dst, dtype, "affine_act dst"))) return rc;
  if (!scale || !shift || ((uintptr_t)scale % 16) || ((uintptr_t)shift % 16)) {
    set_error("affine_act: scale/shift required, 16-byte aligned");
    return PMOE_ERR_ARG;
  }
  V4 res{nullptr, 0, 0, 0, 0, 0, 0, 0};
  if (residual && residual->ptr) {
    if ((rc = check_view(residual, dtype, "affine_act residual"))) return rc;
    res = to_v4(*residual);
  }
  const long long items = (long long)dst->n * dst->h * dst->w * (dst->c / 8);
  const int cg = dst->c / 8;
  int grid = grid_for(items, 256);
  if (256 % cg != 0) grid = (grid + cg - 1) / cg * cg;
  auto dense = [&](const PmoeView4* v) {
    return !v || !v->ptr || (v->n == dst->n && v->h == dst->h && v->w == dst->w && v->c == dst->c && v->sw == v->c &&
                             v->sh == (int64_t)v->w * v->c && v->sn == (int64_t)v->h * v->w * v->c);
  };
  if (dense(src) && dense(dst) && dense(residual)) {
    DISPATCH_DTYPE(dtype, (affine_act_kernel<T, true><<<grid, 256, 0, stream>>>(to_v4(*src), to_v4(*dst), scale, shift, res, act)));
  } else {
    DISPATCH_DTYPE(dtype, (affine_act_kernel<T, false><<<grid, 256, 0, stream>>>(to_v4(*src), to_v4(*dst), scale, shift, res, act)));
  }
  return check_launch("affine_act");
}

int pmoe_affine_act_stats(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* scale, const float* shift,
                          int32_t act, float* pool_sum, int64_t pool_stride, double* out_sum, double* out_sqsum,
                          pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, dtype, "affine_act_stats src");
  if (rc) return rc;
  if ((rc = check_view(dst, dtype, "affine_act_stats dst"))) return rc;
  if (!scale || !shift || ((uintptr_t)scale % 16) || ((uintptr_t)shift % 16) || (out_sum != nullptr) != (out_sqsum != nullptr) ||
      (!pool_sum && !out_sum) || src->c / 8 > 256) {
    set_error("affine_act_stats: scale/shift (16-byte aligned) and at least one statistics output are required");
    return PMOE_ERR_ARG;
  }
  if (dtype != PMOE_BF16 || !dense_view(src) || !dense_view(dst) || src->n != dst->n || src->h != dst->h || src->w != dst->w ||
      src->c != dst->c || (act != PMOE_ACT_NONE && act != PMOE_ACT_RELU)) {
    set_error("affine_act_stats: dense bf16 tensors of one shape, ReLU or no activation");
    return PMOE_ERR_UNSUPPORTED;
  }
  const long long hw = (long long)src->h * src->w;
  long long ppb;
  const dim3 grid = reduce_grid(hw, src->n, &ppb);
  const uint4* px = static_cast<const uint4*>(src->ptr);
  uint4* py = static_cast<uint4*>(dst->ptr);
  const int cg = src->c / 8;
#define PMOE_AAS(R, P, S) act_reduce_fast_kernel<true, R, P, S><<<grid, kRedThreads, 0, stream>>>(px, py, hw, cg, scale, shift, pool_sum, pool_stride, out_sum, out_sqsum, ppb, nullptr)
  const bool relu = act == PMOE_ACT_RELU;
  if (pool_sum && out_sum) { if (relu) PMOE_AAS(true, true, true); else PMOE_AAS(false, true, true); }
  else if (pool_sum) { if (relu) PMOE_AAS(true, true, false); else PMOE_AAS(false, true, false); }
  else { if (relu) PMOE_AAS(true, false, true); else PMOE_AAS(false, false, true); }
#undef PMOE_AAS
  return check_launch("affine_act_stats");
}

// pmoe_affine_act_stats for y = relu(scale*x + shift) that also counts the stored values > 0 per channel (out_pos, fp64, accumulated
// into a zeroed buffer): the third forward statistic a closed-form BatchNorm backward of y's producer needs.
int pmoe_affine_relu_stats_pos(const PmoeView4* src, const PmoeView4* dst, const float* scale, const float* shift, double* out_sum,
                               double* out_sqsum, double* out_pos, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc = check_view(src, PMOE_BF16, "affine_relu_stats_pos src");
  if (rc) return rc;
  if ((rc = check_view(dst, PMOE_BF16, "affine_relu_stats_pos dst"))) return rc;
  if (!scale || !shift || ((uintptr_t)scale % 16) || ((uintptr_t)shift % 16) || !out_sum || !out_sqsum || !out_pos || src->c / 8 > 256) {
    set_error("affine_relu_stats_pos: scale/shift (16-byte aligned) and the three statistics outputs are required");
    return PMOE_ERR_ARG;
  }
  if (!dense_view(src) || !dense_view(dst) || src->n != dst->n || src->h != dst->h || src->w != dst->w || src->c != dst->c) {
    set_error("affine_relu_stats_pos: dense bf16 tensors of one shape");
    return PMOE_ERR_UNSUPPORTED;
  }
  const long long hw = (long long)src->h * src->w;
  long long ppb;
  const dim3 grid = reduce_grid(hw, src->n, &ppb);
  act_reduce_fast_kernel<true, true, false, true, true><<<grid, kRedThreads, 0, stream>>>(
      static_cast<const uint4*>(src->ptr), static_cast<uint4*>(dst->ptr), hw, src->c / 8, scale, shift, nullptr, 0, out_sum, out_sqsum, ppb,
      out_pos);
  return check_launch("affine_relu_stats_pos");
}

}  // extern "C"
