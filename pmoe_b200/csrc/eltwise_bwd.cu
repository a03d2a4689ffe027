// Backward halves of the memory-bound kernels: activation / BatchNorm backward (two passes: per-channel
// reductions, then the elementwise apply), max-pool backward (first-max tie rule of ATen), ECA backward,
// gradient accumulation. Same conventions as eltwise.cu: strided NHWC views, 8 channels per thread.
#include "host_util.h"
#include "act.cuh"
#include "ptx.cuh"
#include "reduce.cuh"

#include <initializer_list>

namespace pmoe {

__device__ __forceinline__ void bload8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void bload8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void bstore8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void bstore8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = r;
}

struct BV4 {
  void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
};
static BV4 bv4(const PmoeView4* v) {
  if (!v || !v->ptr) return BV4{nullptr, 0, 0, 0, 0, 0, 0, 0};
  return BV4{v->ptr, v->n, v->h, v->w, v->c, v->sn, v->sh, v->sw};
}
template <typename T>
__device__ __forceinline__ const T* at(const BV4& v, int n, int h, int w, int c) {
  return static_cast<const T*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}
template <typename T>
__device__ __forceinline__ T* at_mut(const BV4& v, int n, int h, int w, int c) {
  return static_cast<T*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}

__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case PMOE_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case PMOE_ACT_ELU: return z > 0.f ? 1.f : z + 1.f;  // z = exp(x)-1 for x<=0  ->  dz/dx = z+1
    case PMOE_ACT_TANH: return 1.f - z * z;
    case PMOE_ACT_SIGMOID: return z * (1.f - z);
    case PMOE_ACT_RELU6: return (z > 0.f && z < 6.f) ? 1.f : 0.f;
    case PMOE_ACT_HSIGMOID: return (z > 0.f && z < 1.f) ? (1.f / 6.f) : 0.f;
    default: return 1.f;  // Hardswish is not a function of its output: callers pass the forward affine (pre-activation from x)
  }
}

static inline int bgrid(long long items, int threads) {
  long long blocks = (items + threads - 1) / threads;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------- pass 1: per-channel sums of dy and dy*xhat
// dy = dz * act'(z). Block = cg channel groups x (256/cg) pixel lanes; each block walks a slab of pixels and
// finishes with one atomicAdd per channel.
template <typename T, bool FLAT>
__global__ void __launch_bounds__(kRedThreads) bn_bwd_reduce_kernel(BV4 dz, BV4 z, BV4 x, int act, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, double* __restrict__ sum_dy,
                                     double* __restrict__ sum_dy_xhat, long long pix_per_block,
                                     const float* __restrict__ fwd_scale, const float* __restrict__ fwd_shift) {
  // fwd_scale / fwd_shift (optional): the activation's derivative is taken at the PRE-activation u = fwd_scale*x + fwd_shift,
  // recomputed from the layer input x (needed for Hardswish, whose output does not determine it; z is then not read)
  __shared__ float sm[kRedThreads * 8];
  const int cg = dz.c / 8;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const long long npix = (long long)dz.n * dz.h * dz.w;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    float m[8], r[8], fsc[8], fsh[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      m[q] = mean ? __ldg(mean + g * 8 + q) : 0.f;
      r[q] = rstd ? __ldg(rstd + g * 8 + q) : 1.f;
      fsc[q] = fwd_scale ? __ldg(fwd_scale + g * 8 + q) : 1.f;
      fsh[q] = fwd_scale ? __ldg(fwd_shift + g * 8 + q) : 0.f;
    }
#pragma unroll 4
    for (long long p = p0 + lane; p < p1; p += lanes) {
      long long o_dz, o_z, o_x;
      if (FLAT) {
        o_dz = o_z = o_x = (p * cg + g) * 8;
      } else {
        const int w = (int)(p % dz.w);
        const long long t = p / dz.w;
        const int h = (int)(t % dz.h), n = (int)(t / dz.h);
        o_dz = n * dz.sn + h * dz.sh + w * dz.sw + g * 8;
        o_z = n * z.sn + h * z.sh + w * z.sw + g * 8;
        o_x = n * x.sn + h * x.sh + w * x.sw + g * 8;
      }
      float d[8], xv[8];
      bload8(static_cast<const T*>(dz.ptr) + o_dz, d);
      if (fwd_scale) {
        bload8(static_cast<const T*>(x.ptr) + o_x, xv);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] *= act_grad_pre(fmaf(xv[q], fsc[q], fsh[q]), act);
      } else if (act != PMOE_ACT_NONE) {
        float zv[8];
        bload8(static_cast<const T*>(z.ptr) + o_z, zv);
        if (act == PMOE_ACT_RELU) {  // uniform branch hoisted out of the element loop (a per-element switch is ~10x the code)
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : 0.f;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] *= act_grad(zv[q], act);
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] += d[q];
      if (x.ptr && sum_dy_xhat) {
        if (!fwd_scale) bload8(static_cast<const T*>(x.ptr) + o_x, xv);
#pragma unroll
        for (int q = 0; q < 8; ++q) b[q] += d[q] * (xv[q] - m[q]) * r[q];
      }
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  if (sum_dy_xhat) block_channel_sum(b, sm, cg, lanes, tb);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      atomicAdd(sum_dy + c, (double)ta[j]);
      if (sum_dy_xhat) atomicAdd(sum_dy_xhat + c, (double)tb[j]);
    }
  }
}

// ---------------------------------------------------------------- pass 2: dx (and the masked dy for a residual branch)
// batch-stat BN : dx = gamma*rstd * (dy - sum_dy/N - xhat*sum_dy_xhat/N)
// eval BN / bias: dx = dy * scale (scale may be NULL = 1)
// Optional side output of the apply pass: the BatchNorm affine gradients are the two reductions themselves
// (d beta = sum dy*m, d gamma = sum dy*m*xhat), so block 0 writes them as fp32 into the parameters' gradient slots
// (a separate fp64 -> fp32 conversion launch per parameter otherwise: 2 x 20 BatchNorms x K experts per step).
struct ParamGrads {
  float* dgamma;
  float* dbeta;
  int n;
  int accumulate;
};
__device__ __forceinline__ void write_param_grads(const ParamGrads& pg, const double* __restrict__ sum_dy,
                                                  const double* __restrict__ sum_dy_xhat) {
  if (blockIdx.x != 0 || pg.n <= 0) return;
  for (int c = threadIdx.x; c < pg.n; c += blockDim.x) {
    if (pg.dbeta) pg.dbeta[c] = (pg.accumulate ? pg.dbeta[c] : 0.f) + (float)sum_dy[c];
    if (pg.dgamma && sum_dy_xhat) pg.dgamma[c] = (pg.accumulate ? pg.dgamma[c] : 0.f) + (float)sum_dy_xhat[c];
  }
}

// FLAT: every view is a dense (n,h,w,c) tensor of the same shape -> item i lives at element offset 8*i in all of
// them (no index arithmetic). The grid stride is a multiple of the channel-group count, so a thread keeps ONE channel
// group for its whole loop and the per-channel constants are loaded once.
template <typename T, bool FLAT>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(BV4 dz, BV4 z, BV4 x, int act, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma,
                                    const double* __restrict__ sum_dy, const double* __restrict__ sum_dy_xhat, float inv_n,
                                    int batch_stats, BV4 dx, BV4 dres, int accumulate_dres, ParamGrads pg,
                                    const float* __restrict__ fwd_scale, const float* __restrict__ fwd_shift) {
  if (batch_stats) write_param_grads(pg, sum_dy, sum_dy_xhat);
  const int cg = dz.c / 8;
  const long long total = (long long)dz.n * dz.h * dz.w * cg;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = (int)(first % cg);
  float ka[8], km[8], kr[8], c1[8], c2[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = g * 8 + q;
    const float gm = gamma ? __ldg(gamma + c) : 1.f;
    if (batch_stats) {
      kr[q] = __ldg(rstd + c);
      km[q] = __ldg(mean + c);
      ka[q] = gm * kr[q];
      c1[q] = (float)(sum_dy[c] * (double)inv_n);
      c2[q] = (float)(sum_dy_xhat[c] * (double)inv_n);
    } else {
      ka[q] = gm;
      kr[q] = km[q] = c1[q] = c2[q] = 0.f;
    }
  }
#pragma unroll 2
  for (long long i = first; i < total; i += stride) {
    long long o_dz, o_z, o_x, o_dx, o_dr;
    if (FLAT) {
      o_dz = o_z = o_x = o_dx = o_dr = i * 8;
    } else {
      long long pix = i / cg;
      const int w = (int)(pix % dz.w);
      pix /= dz.w;
      const int h = (int)(pix % dz.h);
      const int n = (int)(pix / dz.h);
      o_dz = n * dz.sn + h * dz.sh + w * dz.sw + g * 8;
      o_z = n * z.sn + h * z.sh + w * z.sw + g * 8;
      o_x = n * x.sn + h * x.sh + w * x.sw + g * 8;
      o_dx = n * dx.sn + h * dx.sh + w * dx.sw + g * 8;
      o_dr = n * dres.sn + h * dres.sh + w * dres.sw + g * 8;
    }
    float d[8];
    bload8(static_cast<const T*>(dz.ptr) + o_dz, d);
    if (fwd_scale) {  // derivative at the pre-activation recomputed from x (see bn_bwd_reduce_kernel)
      float xp[8];
      bload8(static_cast<const T*>(x.ptr) + o_x, xp);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        d[q] *= act_grad_pre(fmaf(xp[q], __ldg(fwd_scale + g * 8 + q), __ldg(fwd_shift + g * 8 + q)), act);
    } else if (act != PMOE_ACT_NONE) {
      float zv[8];
      bload8(static_cast<const T*>(z.ptr) + o_z, zv);
      if (act == PMOE_ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] *= act_grad(zv[q], act);
      }
    }
    if (dres.ptr) {
      float o[8];
      if (accumulate_dres) {
        bload8(static_cast<const T*>(dres.ptr) + o_dr, o);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] += d[q];
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = d[q];
      }
      bstore8(static_cast<T*>(dres.ptr) + o_dr, o);
    }
    if (dx.ptr) {
      float o[8];
      if (batch_stats) {
        float xv[8];
        bload8(static_cast<const T*>(x.ptr) + o_x, xv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float xh = (xv[q] - km[q]) * kr[q];
          o[q] = ka[q] * (d[q] - c1[q] - xh * c2[q]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = d[q] * ka[q];
      }
      bstore8(static_cast<T*>(dx.ptr) + o_dx, o);
    }
  }
}

// ---------------------------------------------------------------- max-pool backward (gather form)
// dx[i] = sum over windows w containing i of dy[w] * [i is the FIRST maximum of w in row-major scan order]
template <typename T>
__global__ void maxpool_bwd_kernel(BV4 x, BV4 dy, BV4 dx, int k, int stride, int pad, int accumulate) {
  const int cg = dx.c / 8;
  const long long total = (long long)dx.n * dx.h * dx.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int iw = (int)(pix % dx.w);
    pix /= dx.w;
    const int ih = (int)(pix % dx.h);
    const int n = (int)(pix / dx.h);
    float xv[8], o[8];
    bload8(at<T>(x, n, ih, iw, g * 8), xv);
    if (accumulate) bload8(at<T>(dx, n, ih, iw, g * 8), o);
    else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = 0.f;
    }
    // windows (oh, ow) with oh*stride - pad <= ih < oh*stride - pad + k
    const int oh_lo = max(0, (ih + pad - k + stride) / stride), oh_hi = min(dy.h - 1, (ih + pad) / stride);
    const int ow_lo = max(0, (iw + pad - k + stride) / stride), ow_hi = min(dy.w - 1, (iw + pad) / stride);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        bool mine[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) mine[q] = true;
        bool before = true;  // scanning positions that precede (ih, iw) in the window
        for (int r = 0; r < k; ++r) {
          const int yh = oh * stride - pad + r;
          if (yh < 0 || yh >= x.h) continue;
          for (int s = 0; s < k; ++s) {
            const int yw = ow * stride - pad + s;
            if (yw < 0 || yw >= x.w) continue;
            if (yh == ih && yw == iw) {
              before = false;
              continue;
            }
            float v[8];
            bload8(at<T>(x, n, yh, yw, g * 8), v);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              // an earlier element wins ties; a later one must be strictly greater (NaN ignored)
              if (before ? (v[q] >= xv[q]) : (v[q] > xv[q])) mine[q] = false;
            }
          }
        }
        float d[8];
        bload8(at<T>(dy, n, oh, ow, g * 8), d);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (mine[q]) o[q] += d[q];
      }
    }
    bstore8(at_mut<T>(dx, n, ih, iw, g * 8), o);
  }
}

// max-pool backward from the argmax codes the forward stored (pmoe_maxpool_idx): no re-scan of the windows.
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_idx_kernel(BV4 dy, const uint8_t* __restrict__ idx, BV4 dx, int k, int stride, int pad,
                                                              int accumulate, int cg_shift) {
  // grid.x = (image, input row); a thread handles (input column, 8 channels): no 64-bit divisions on the hot path
  const int cg = dx.c / 8;
  const int n = blockIdx.x / dx.h, ih = blockIdx.x % dx.h;
  const int oh_lo = max(0, (ih + pad - k + stride) / stride), oh_hi = min(dy.h - 1, (ih + pad) / stride);
  const int row_items = dx.w * cg;
  for (int item = blockIdx.y * blockDim.x + threadIdx.x; item < row_items; item += gridDim.y * blockDim.x) {
    const int iw = cg_shift >= 0 ? (item >> cg_shift) : item / cg;
    const int g = item - iw * cg;
    float o[8];
    if (accumulate) bload8(at<T>(dx, n, ih, iw, g * 8), o);
    else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = 0.f;
    }
    const int ow_lo = max(0, (iw + pad - k + stride) / stride), ow_hi = min(dy.w - 1, (iw + pad) / stride);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const uint32_t code = (uint32_t)((ih - (oh * stride - pad)) * k + (iw - (ow * stride - pad)));
        const uint2 pk = __ldg(reinterpret_cast<const uint2*>(idx + ((size_t)(n * dy.h + oh) * dy.w + ow) * dy.c + g * 8));
        // compare all 8 codes first: most windows do not select this pixel, and then dy is not needed at all
        const uint32_t c4 = code * 0x01010101u;
        const uint32_t mx = pk.x ^ c4, my = pk.y ^ c4;  // a zero byte marks a match
        const bool any = (((mx - 0x01010101u) & ~mx) | ((my - 0x01010101u) & ~my)) & 0x80808080u;
        if (!any) continue;
        float d[8];
        bload8(at<T>(dy, n, oh, ow, g * 8), d);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t a = ((q < 4 ? pk.x : pk.y) >> (8 * (q & 3))) & 0xffu;
          if (a == code) o[q] += d[q];
        }
      }
    }
    bstore8(at_mut<T>(dx, n, ih, iw, g * 8), o);
  }
}

// ---------------------------------------------------------------- ECA backward
// out = x * gate[n,c]. pass 1: dgate[n,c] = sum_hw dout*x (per-image channel sums of a product)
template <typename T>
__global__ void __launch_bounds__(kRedThreads) prod_channel_sums_kernel(BV4 a, BV4 b, double* __restrict__ out, long long out_stride, int rows_per_block) {
  __shared__ float sm[kRedThreads * 8];
  const int cg = a.c / 8;
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const long long hw = (long long)a.h * a.w;
  const long long p0 = (long long)blockIdx.x * rows_per_block;
  long long p1 = p0 + rows_per_block;
  if (p1 > hw) p1 = hw;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += lanes) {
      const int h = (int)(p / a.w), w = (int)(p % a.w);
      float u[8], v[8];
      bload8(at<T>(a, n, h, w, g * 8), u);
      bload8(at<T>(b, n, h, w, g * 8), v);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += u[q] * v[q];
    }
  }
  float tot[kRedMaxIter];
  block_channel_sum(acc, sm, cg, lanes, tot);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8 && tot[j] != 0.f) atomicAdd(out + n * out_stride + c, (double)tot[j]);  // cross-block sum in fp64
  }
}

// tiny: gate = sigmoid(pre), pre = conv1d(mean). dpre = dgate*gate*(1-gate); dmean = corr(dpre, w) / count;
// dw[t] += sum_{n,l} dpre[n,l] * mean[n, l+t-k/2]
// The ECA gate sits in front of a BatchNorm, which cancels a common channel scale: dgate and dw are small remainders of
// large cancelling sums, so this (tiny) kernel keeps them in fp64.
__global__ void eca_gate_bwd_kernel(const double* __restrict__ dgate, long long dgate_stride, const float* __restrict__ gate,
                                    long long gate_stride, const float* __restrict__ pool_sum, long long pool_stride,
                                    float inv_count, const float* __restrict__ w, int k, int groups, int group_c,
                                    int group_stride, float* __restrict__ dmean, long long dmean_stride,
                                    double* __restrict__ dw) {
  const int n = blockIdx.x;
  const int L = groups * group_c;
  extern __shared__ double s_dpre[];  // L entries
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const int p = (l / group_c) * group_stride + (l % group_c);
    const double gt = (double)gate[n * gate_stride + p];
    s_dpre[l] = dgate[n * dgate_stride + p] * gt * (1.0 - gt);
  }
  __syncthreads();
  for (int pc = threadIdx.x; pc < groups * group_stride; pc += blockDim.x) {
    const int g = pc / group_stride, j = pc % group_stride;
    double acc = 0.0;
    if (j < group_c) {
      const int l = g * group_c + j;  // d mean[l] = sum_t w[t] * dpre[l - t + k/2]
      for (int t = 0; t < k; ++t) {
        const int lo = l - t + k / 2;
        if (lo >= 0 && lo < L) acc += (double)w[t] * s_dpre[lo];
      }
    }
    dmean[n * dmean_stride + pc] = (float)(acc * (double)inv_count);
  }
  if (dw) {
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
      double acc = 0.0;
      for (int l = 0; l < L; ++l) {
        const int ll = l + t - k / 2;
        if (ll >= 0 && ll < L) {
          const int p = (ll / group_c) * group_stride + (ll % group_c);
          acc += s_dpre[l] * ((double)pool_sum[n * pool_stride + p] * (double)inv_count);
        }
      }
      atomicAdd(dw + t, acc);
    }
  }
}

// pass 2: dx = dout*gate[n,c] + dmean[n,c]   (optionally accumulated into dx)
template <typename T>
__global__ void eca_bwd_apply_kernel(BV4 dout, const float* __restrict__ gate, long long gate_stride,
                                     const float* __restrict__ dmean, long long dmean_stride, BV4 dx, int accumulate) {
  const int cg = dx.c / 8;
  const long long total = (long long)dx.n * dx.h * dx.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int w = (int)(pix % dx.w);
    pix /= dx.w;
    const int h = (int)(pix % dx.h);
    const int n = (int)(pix / dx.h);
    float d[8], gt[8], dm[8], o[8];
    bload8(at<T>(dout, n, h, w, g * 8), d);
    bload8(gate + n * gate_stride + g * 8, gt);
    bload8(dmean + n * dmean_stride + g * 8, dm);
    if (accumulate) bload8(at<T>(dx, n, h, w, g * 8), o);
    else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] += d[q] * gt[q] + dm[q];
    bstore8(at_mut<T>(dx, n, h, w, g * 8), o);
  }
}

// dst (+)= alpha * src [+ per-(n,c) broadcast term]  — gradient accumulation / avg-pool backward
template <typename T>
__global__ void axpy_kernel(BV4 src, BV4 dst, float alpha, const float* __restrict__ bcast, long long bcast_stride,
                            int accumulate) {
  const int cg = dst.c / 8;
  const long long total = (long long)dst.n * dst.h * dst.w * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % cg);
    long long pix = i / cg;
    const int w = (int)(pix % dst.w);
    pix /= dst.w;
    const int h = (int)(pix % dst.h);
    const int n = (int)(pix / dst.h);
    float o[8];
    if (accumulate) bload8(at<T>(dst, n, h, w, g * 8), o);
    else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = 0.f;
    }
    if (src.ptr) {
      float s[8];
      bload8(at<T>(src, n, h, w, g * 8), s);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] += alpha * s[q];
    }
    if (bcast) {
      float bq[8];
      bload8(bcast + n * bcast_stride + g * 8, bq);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] += bq[q];
    }
    bstore8(at_mut<T>(dst, n, h, w, g * 8), o);
  }
}


// ---------------------------------------------------------------- fast paths: dense bf16 tensors, ReLU / no activation
// Straight-line code, four independent 16-byte loads per tensor in flight per thread (the generic kernels above issue one
// and sit at ~40 % of HBM bandwidth), per-channel constants folded so the inner loop is two FMAs per element.
__device__ __forceinline__ void bf16x8_to_f32(const uint4& r, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&v)[8]) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]);
  r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]);
  r.w = pack_bf16x2(v[6], v[7]);
  return r;
}

// Contiguous bf16, compile-time window (3x3/s2/p1 and 2x2/s2): at most WN x WN windows cover an input pixel; their argmax
// codes and upstream gradients are all loaded before the first compare, indices are 32-bit, no divisions by run-time values.
template <int K, int S, int P>
__global__ void __launch_bounds__(256) maxpool_bwd_fast_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                               uint4* __restrict__ dx, int H, int W, int OH, int OW, int cg,
                                                               int cg_shift, int accumulate) {
  constexpr int WN = (K + S - 1) / S;  // windows per axis that can contain one input pixel
  const int n = blockIdx.x / H, ih = blockIdx.x - n * H;
  const int row_items = W * cg;
  const int oh_hi = (ih + P) / S;
  for (int item = blockIdx.y * blockDim.x + threadIdx.x; item < row_items; item += gridDim.y * blockDim.x) {
    const int iw = cg_shift >= 0 ? (item >> cg_shift) : item / cg;
    const int g = item - iw * cg;
    const int ow_hi = (iw + P) / S;
    uint2 code[WN * WN];
    uint4 grad[WN * WN];
    uint32_t want[WN * WN];
    bool ok[WN * WN];
#pragma unroll
    for (int a = 0; a < WN; ++a) {
      const int oh = oh_hi - a, r = ih + P - oh * S;  // row of this pixel inside window oh
#pragma unroll
      for (int b = 0; b < WN; ++b) {
        const int ow = ow_hi - b, c = iw + P - ow * S;
        const int t = a * WN + b;
        ok[t] = oh >= 0 && oh < OH && ow >= 0 && ow < OW && r < K && c < K;
        want[t] = (uint32_t)(r * K + c);
        if (ok[t]) {
          const size_t o = (((size_t)n * OH + oh) * OW + ow) * cg + g;
          code[t] = __ldg(idx + o);
          grad[t] = __ldg(dy + o);
        }
      }
    }
    const size_t od = ((size_t)blockIdx.x * W + iw) * cg + g;
    float o[8];
    if (accumulate) bf16x8_to_f32(dx[od], o);
    else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < WN * WN; ++t) {
      if (!ok[t]) continue;
      float d[8];
      bf16x8_to_f32(grad[t], d);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t a = ((q < 4 ? code[t].x : code[t].y) >> (8 * (q & 3))) & 0xffu;
        if (a == want[t]) o[q] += d[q];
      }
    }
    dx[od] = f32_to_bf16x8(o);
  }
}


// MaxPool2d(3, 2, 1) backward for even H and W (the ResNet stem pool): one thread per horizontally adjacent input pixel PAIR
// (even, odd) and channel group. The even pixel lies in column-window j only (at column 1), the odd pixel in windows j
// (column 2) and j+1 (column 0); the input row lies in one row-window (even rows) or two (odd rows) — a block-uniform
// branch. So every (code, gradient) pair that is loaded is used, the wanted argmax codes are constants, and the byte
// compares run four channels at a time (__vcmpeq4 + two byte permutes give the bf16x2 masks). The general kernel above
// spends ~4x the instructions (every thread walks 2x2 windows, a quarter of which apply) and was issue-bound at 32 % of HBM.
__device__ __forceinline__ void add_masked(float (&o)[8], const uint2& code, const uint4& grad, uint32_t want) {
  const uint32_t w4 = want * 0x01010101u;
  const uint32_t m0 = __vcmpeq4(code.x, w4), m1 = __vcmpeq4(code.y, w4);
  const uint32_t gq[4] = {grad.x & __byte_perm(m0, 0u, 0x1100u), grad.y & __byte_perm(m0, 0u, 0x3322u),
                          grad.z & __byte_perm(m1, 0u, 0x1100u), grad.w & __byte_perm(m1, 0u, 0x3322u)};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    o[2 * q] += __uint_as_float(gq[q] << 16);
    o[2 * q + 1] += __uint_as_float(gq[q] & 0xffff0000u);
  }
}
__global__ void __launch_bounds__(256) maxpool3s2_bwd_pair_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                                  uint4* __restrict__ dx, int H, int W, int OH, int OW, int cg,
                                                                  int cg_shift, int accumulate) {
  const int n = blockIdx.x / H, ih = blockIdx.x - n * H;
  const int items = (W >> 1) * cg;
  const bool odd = (ih & 1) != 0;
  const int oh0 = ih >> 1;                  // row-window that holds this row at r = 1 (even row) or r = 2 (odd row)
  const uint32_t r0 = odd ? 6u : 3u;        // r * 3
  const bool two = odd && (oh0 + 1 < OH);   // odd rows are also row 0 of the window below
  for (int item = blockIdx.y * blockDim.x + threadIdx.x; item < items; item += gridDim.y * blockDim.x) {
    const int j = cg_shift >= 0 ? (item >> cg_shift) : item / cg;
    const int g = item - j * cg;
    const bool right = j + 1 < OW;
    const size_t oa = (((size_t)n * OH + oh0) * OW + j) * cg + g;
    const size_t oc = oa + (size_t)OW * cg;
    const uint2 zc = make_uint2(0xffffffffu, 0xffffffffu);  // code 255 never matches
    const uint4 zg = make_uint4(0u, 0u, 0u, 0u);
    const uint2 cA = __ldg(idx + oa);
    const uint4 gA = __ldg(dy + oa);
    const uint2 cB = right ? __ldg(idx + oa + cg) : zc;
    const uint4 gB = right ? __ldg(dy + oa + cg) : zg;
    const uint2 cC = two ? __ldg(idx + oc) : zc;
    const uint4 gC = two ? __ldg(dy + oc) : zg;
    const uint2 cD = (two && right) ? __ldg(idx + oc + cg) : zc;
    const uint4 gD = (two && right) ? __ldg(dy + oc + cg) : zg;
    const size_t od = ((size_t)blockIdx.x * W + 2 * j) * cg + g;
    float o0[8], o1[8];
    if (accumulate) {
      bf16x8_to_f32(dx[od], o0);
      bf16x8_to_f32(dx[od + cg], o1);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) o0[q] = o1[q] = 0.f;
    }
    add_masked(o0, cA, gA, r0 + 1u);
    add_masked(o1, cA, gA, r0 + 2u);
    add_masked(o1, cB, gB, r0);
    if (two) {
      add_masked(o0, cC, gC, 1u);
      add_masked(o1, cC, gC, 2u);
      add_masked(o1, cD, gD, 0u);
    }
    dx[od] = f32_to_bf16x8(o0);
    dx[od + cg] = f32_to_bf16x8(o1);
  }
}

// dx (+)= dout * gate[n, c] + dmean[n, c] on dense bf16: grid (x, n), one channel group per thread, gate / dmean in
// registers, four 16-byte loads in flight (the generic kernel: one load behind a div/mod chain, 61 % of HBM).
// NEXT: xf is the forward input of the ECA block and at the same time the ReLU output of an upstream BatchNorm whose complete
// gradient dx is: that layer's backward sums (sum dx*[xf>0], sum dx*xf; see bn_bwd_apply_fast_kernel) are reduced here.
template <bool ACC, bool NEXT>
__global__ void __launch_bounds__(256) eca_bwd_apply_fast_kernel(const uint4* __restrict__ dout, uint4* __restrict__ dx,
                                                                 long long per_img, int cg, const float* __restrict__ gate,
                                                                 long long gate_stride, const float* __restrict__ dmean,
                                                                 long long dmean_stride, const uint4* __restrict__ xf,
                                                                 double* __restrict__ next_s1, double* __restrict__ next_s2) {
  __shared__ float sm_next[NEXT ? kRedThreads * 8 : 8];
  float na[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int n = blockIdx.y;
  const long long stride = (long long)gridDim.x * blockDim.x;  // a multiple of cg
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = (int)(first % cg);
  float gt[8], dm[8];
  bload8(gate + n * gate_stride + g * 8, gt);
  bload8(dmean + n * dmean_stride + g * 8, dm);
  const uint4* src = dout + (size_t)n * per_img;
  uint4* dst = dx + (size_t)n * per_img;
  for (long long i = first; i < per_img; i += 4 * stride) {
    uint4 r[4], ro[4], rx[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ii = i + (long long)u * stride;
      ok[u] = ii < per_img;
      const long long off = ok[u] ? ii : i;
      r[u] = __ldg(src + off);
      if (ACC) ro[u] = dst[off];
      if (NEXT) rx[u] = __ldg(xf + (size_t)n * per_img + off);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      float d[8], o[8];
      bf16x8_to_f32(r[u], d);
      if (ACC) bf16x8_to_f32(ro[u], o);
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = fmaf(d[q], gt[q], dm[q]) + (ACC ? o[q] : 0.f);
      const uint4 pk = f32_to_bf16x8(o);
      dst[i + (long long)u * stride] = pk;
      if (NEXT) {
        float ov[8], xv[8];
        bf16x8_to_f32(pk, ov);
        bf16x8_to_f32(rx[u], xv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          na[q] += xv[q] > 0.f ? ov[q] : 0.f;
          nb[q] = fmaf(ov[q], xv[q], nb[q]);
        }
      }
    }
  }
  if (NEXT) {
    float ta[kRedMaxIter], tb[kRedMaxIter];
    block_channel_sum(na, sm_next, cg, blockDim.x / cg, ta);
    block_channel_sum(nb, sm_next, cg, blockDim.x / cg, tb);
#pragma unroll
    for (int j = 0; j < kRedMaxIter; ++j) {
      const int c = threadIdx.x + j * kRedThreads;
      if (c < cg * 8) {
        atomicAdd(next_s1 + c, (double)ta[j]);
        atomicAdd(next_s2 + c, (double)tb[j]);
      }
    }
  }
}

// ACT: 0 = no activation, 1 = ReLU mask read from the saved output z, 2 = ReLU mask recomputed from x with the forward's own
// scale/shift (fmaf(x, scale, shift) > 0 is bit-for-bit what affine_act computed before its max(.,0)), which saves the read of z.
template <int ACT, bool HAS_X>
__global__ void __launch_bounds__(kRedThreads) bn_bwd_reduce_fast_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z,
                                                                         const uint4* __restrict__ x, long long npix, int cg,
                                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                         double* __restrict__ sum_dy, double* __restrict__ sum_dy_xhat,
                                                                         long long pix_per_block, const float* __restrict__ fwd_scale,
                                                                         const float* __restrict__ fwd_shift) {
  __shared__ float sm[kRedThreads * 8];
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  float msc[8], msh[8];
  if (ACT == 2) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      msc[q] = __ldg(fwd_scale + g * 8 + q);
      msh[q] = __ldg(fwd_shift + g * 8 + q);
    }
  }
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += 4LL * lanes) {
      uint4 rd[4], rz[4], rx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pp = p + (long long)u * lanes;
        const bool ok = pp < p1;
        const long long off = ok ? pp * cg + g : p * cg + g;
        rd[u] = ok ? __ldg(dz + off) : make_uint4(0u, 0u, 0u, 0u);
        if (ACT == 1 || ACT == 3) rz[u] = __ldg(z + off);
        if (HAS_X || ACT == 2) rx[u] = __ldg(x + off);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float d[8], xv[8];
        bf16x8_to_f32(rd[u], d);
        if (HAS_X || ACT == 2) bf16x8_to_f32(rx[u], xv);
        if (ACT == 1) {
          float zv[8];
          bf16x8_to_f32(rz[u], zv);
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : 0.f;
        }
        if (ACT == 3) {  // ELU from the saved output: d/dx = z + 1 for z <= 0
          float zv[8];
          bf16x8_to_f32(rz[u], zv);
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : d[q] * (zv[q] + 1.f);
        }
        if (ACT == 2) {
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = fmaf(xv[q], msc[q], msh[q]) > 0.f ? d[q] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) a[q] += d[q];
        if (HAS_X) {
#pragma unroll
          for (int q = 0; q < 8; ++q) b[q] = fmaf(d[q], xv[q], b[q]);
        }
      }
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  if (HAS_X) block_channel_sum(b, sm, cg, lanes, tb);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      atomicAdd(sum_dy + c, (double)ta[j]);
      // sum dy*xhat = rstd * (sum dy*x - mean * sum dy), combined in fp64 per block
      if (HAS_X) atomicAdd(sum_dy_xhat + c, (double)__ldg(rstd + c) * ((double)tb[j] - (double)__ldg(mean + c) * (double)ta[j]));
    }
  }
}

// NEXT: x is itself the ReLU output of an upstream BatchNorm (torchvision's bn1 directly after the stem block's conv+BN+ReLU),
// and dx is that layer's complete upstream gradient: its backward reductions  sum dx*[x>0]  and  sum dx*x  (x = 0 exactly
// where its mask is 0) are accumulated here from the registers, so the upstream layer needs no reduce pass of its own
// (block layout threadIdx.x = lane*cg + g: requires 256 % cg == 0).
template <int ACT, bool BATCH, bool NEXT>
__global__ void __launch_bounds__(256) bn_bwd_apply_fast_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z,
                                                                const uint4* __restrict__ x, uint4* __restrict__ dx, uint4* dres,
                                                                int accumulate_dres, long long total, int cg,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const double* __restrict__ sum_dy,
                                                                const double* __restrict__ sum_dy_xhat, float inv_n,
                                                                const float* __restrict__ fwd_scale, const float* __restrict__ fwd_shift,
                                                                double* __restrict__ next_s1, double* __restrict__ next_s2, ParamGrads pg) {
  if (BATCH) write_param_grads(pg, sum_dy, sum_dy_xhat);
  __shared__ float sm_next[NEXT ? kRedThreads * 8 : 8];
  float na[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long stride = (long long)gridDim.x * blockDim.x;  // a multiple of cg: one channel group per thread
  const long long first = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int g = (int)(first % cg);
  float msc[8], msh[8];  // ACT == 2: ReLU mask recomputed from x (see bn_bwd_reduce_fast_kernel)
  if (ACT == 2) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      msc[q] = __ldg(fwd_scale + g * 8 + q);
      msh[q] = __ldg(fwd_shift + g * 8 + q);
    }
  }
  // dx = gamma*rstd*(d - c1 - (x - mean)*rstd*c2) = A*d + B*x + C
  float A[8], Bc[8], Cc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int c = g * 8 + q;
    const float gm = gamma ? __ldg(gamma + c) : 1.f;
    if (BATCH) {
      const double r = (double)__ldg(rstd + c), m = (double)__ldg(mean + c);
      const double c1 = sum_dy[c] * (double)inv_n, c2 = sum_dy_xhat[c] * (double)inv_n;
      A[q] = (float)((double)gm * r);
      Bc[q] = (float)(-(double)gm * r * r * c2);
      Cc[q] = (float)((double)gm * r * (r * c2 * m - c1));
    } else {
      A[q] = gm;
      Bc[q] = Cc[q] = 0.f;
    }
  }
  for (long long i = first; i < total; i += 4 * stride) {
    uint4 rd[4], rz[4], rx[4], ro[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ii = i + (long long)u * stride;
      ok[u] = ii < total;
      const long long off = ok[u] ? ii : i;
      rd[u] = __ldg(dz + off);
      if (ACT == 1 || ACT == 3) rz[u] = __ldg(z + off);
      if (BATCH || ACT == 2) rx[u] = __ldg(x + off);
      if (dres != nullptr && accumulate_dres) ro[u] = dres[off];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const long long ii = i + (long long)u * stride;
      float d[8], xv[8];
      bf16x8_to_f32(rd[u], d);
      if (BATCH || ACT == 2) bf16x8_to_f32(rx[u], xv);
      if (ACT == 1) {
        float zv[8];
        bf16x8_to_f32(rz[u], zv);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : 0.f;
      }
      if (ACT == 3) {  // ELU from the saved output
        float zv[8];
        bf16x8_to_f32(rz[u], zv);
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = zv[q] > 0.f ? d[q] : d[q] * (zv[q] + 1.f);
      }
      if (ACT == 2) {
#pragma unroll
        for (int q = 0; q < 8; ++q) d[q] = fmaf(xv[q], msc[q], msh[q]) > 0.f ? d[q] : 0.f;
      }
      if (dres != nullptr) {
        float o[8];
        if (accumulate_dres) {
          bf16x8_to_f32(ro[u], o);
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] += d[q];
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = d[q];
        }
        dres[ii] = f32_to_bf16x8(o);
      }
      if (dx != nullptr) {
        float o[8];
        if (BATCH) {
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = fmaf(A[q], d[q], fmaf(Bc[q], xv[q], Cc[q]));
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) o[q] = A[q] * d[q];
        }
        const uint4 pk = f32_to_bf16x8(o);
        dx[ii] = pk;
        if (NEXT) {  // sums of what a separate reduce pass would read back: the bf16-rounded dx
          float ov[8];
          bf16x8_to_f32(pk, ov);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            na[q] += xv[q] > 0.f ? ov[q] : 0.f;
            nb[q] = fmaf(ov[q], xv[q], nb[q]);
          }
        }
      }
    }
  }
  if (NEXT) {
    float ta[kRedMaxIter], tb[kRedMaxIter];
    block_channel_sum(na, sm_next, cg, blockDim.x / cg, ta);
    block_channel_sum(nb, sm_next, cg, blockDim.x / cg, tb);
#pragma unroll
    for (int j = 0; j < kRedMaxIter; ++j) {
      const int c = threadIdx.x + j * kRedThreads;
      if (c < cg * 8) {
        atomicAdd(next_s1 + c, (double)ta[j]);
        atomicAdd(next_s2 + c, (double)tb[j]);
      }
    }
  }
}

// all non-null views dense (n,h,w,c) tensors of the reference view's shape?
static bool all_flat(const PmoeView4* ref, std::initializer_list<const PmoeView4*> vs) {
  for (const PmoeView4* v : vs) {
    if (!v || !v->ptr) continue;
    if (v->n != ref->n || v->h != ref->h || v->w != ref->w || v->c != ref->c || v->sw != v->c || v->sh != (int64_t)v->w * v->c ||
        v->sn != (int64_t)v->h * v->w * v->c)
      return false;
  }
  return true;
}

// grid whose total thread count is a multiple of the channel-group count (threads keep one channel group)
static int grid_cg(long long items, int cg) {
  int blocks = bgrid(items, 256);
  if (256 % cg != 0) {  // odd group counts (e.g. 48 channels): make blocks * 256 a multiple of cg
    blocks = (blocks + cg - 1) / cg * cg;
  }
  return blocks;
}

static int chk(const PmoeView4* v, int dtype, const char* what, bool optional = false) {
  if (optional && (!v || !v->ptr)) return PMOE_OK;
  const int esz = dtype == PMOE_BF16 ? 2 : 4;
  if (!v || !v->ptr || v->c % 8 || ((uintptr_t)v->ptr % 16) || (v->sw * esz) % 16 || (v->sh * esz) % 16 || (v->sn * esz) % 16) {
    set_error("%s: view must be 16-byte aligned, channels a multiple of 8", what);
    return PMOE_ERR_ARG;
  }
  return PMOE_OK;
}

}  // namespace pmoe

using namespace pmoe;

static inline bool host_piecewise(int act) {
  return act == PMOE_ACT_RELU || act == PMOE_ACT_RELU6 || act == PMOE_ACT_HSWISH || act == PMOE_ACT_HSIGMOID;
}

#define BW_DISPATCH(dtype, ...)                      \
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

extern "C" {

int pmoe_bn_bwd_reduce(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                       const float* mean, const float* rstd, double* sum_dy, double* sum_dy_xhat, const float* fwd_scale,
                       const float* fwd_shift, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  // pre_x: no saved output given; the activation's derivative is taken at the pre-activation recomputed from x with the forward's
  // scale/shift (ReLU / ReLU6 / Hardswish / Hardsigmoid). mask_x: its dense-bf16 ReLU form (fast kernels)
  const bool pre_x = host_piecewise(act) && (!z || !z->ptr) && x && x->ptr && fwd_scale && fwd_shift;
  const bool mask_x = pre_x && act == PMOE_ACT_RELU;
  if ((rc = chk(dz, dtype, "bn_bwd_reduce dz"))) return rc;
  if ((rc = chk(z, dtype, "bn_bwd_reduce z", act == PMOE_ACT_NONE || pre_x))) return rc;
  if ((rc = chk(x, dtype, "bn_bwd_reduce x", true))) return rc;
  const int cg = dz->c / 8;
  if (!sum_dy || cg > 256 || (act != PMOE_ACT_NONE && !pre_x && (!z || !z->ptr)) || (act == PMOE_ACT_HSWISH && !pre_x)) {
    set_error("bn_bwd_reduce: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long npix = (long long)dz->n * dz->h * dz->w;
  long long blocks = (long long)num_sms() * 8;
  long long ppb = (npix + blocks - 1) / blocks;
  if (ppb < 64) ppb = 64;
  blocks = (npix + ppb - 1) / ppb;
  const bool flat = all_flat(dz, {dz, z, x});
  if (flat && dtype == PMOE_BF16 && (act == PMOE_ACT_NONE || act == PMOE_ACT_RELU || act == PMOE_ACT_ELU) && (!sum_dy_xhat || (x && x->ptr && mean && rstd))) {
    const uint4* pdz = static_cast<const uint4*>(dz->ptr);
    const uint4* pz = (act && !mask_x) ? static_cast<const uint4*>(z->ptr) : nullptr;
    const bool has_x = sum_dy_xhat != nullptr;
    const uint4* px = (has_x || mask_x) ? static_cast<const uint4*>(x->ptr) : nullptr;
#define PMOE_RED_FAST(A, X) bn_bwd_reduce_fast_kernel<A, X><<<(unsigned)blocks, 256, 0, stream>>>(pdz, pz, px, npix, cg, mean, rstd, sum_dy, sum_dy_xhat, ppb, fwd_scale, fwd_shift)
    if (mask_x && has_x) PMOE_RED_FAST(2, true);
    else if (mask_x) PMOE_RED_FAST(2, false);
    else if (act == PMOE_ACT_ELU && has_x) PMOE_RED_FAST(3, true);
    else if (act == PMOE_ACT_ELU) PMOE_RED_FAST(3, false);
    else if (act && has_x) PMOE_RED_FAST(1, true);
    else if (act) PMOE_RED_FAST(1, false);
    else if (has_x) PMOE_RED_FAST(0, true);
    else PMOE_RED_FAST(0, false);
#undef PMOE_RED_FAST
    return check_launch("bn_bwd_reduce");
  }
  const float* gsc = pre_x ? fwd_scale : nullptr;
  const float* gsh = pre_x ? fwd_shift : nullptr;
  if (flat) {
    BW_DISPATCH(dtype, (bn_bwd_reduce_kernel<T, true><<<(unsigned)blocks, 256, 0, stream>>>(bv4(dz), bv4(z), bv4(x), act, mean, rstd, sum_dy, sum_dy_xhat, ppb, gsc, gsh)));
  } else {
    BW_DISPATCH(dtype, (bn_bwd_reduce_kernel<T, false><<<(unsigned)blocks, 256, 0, stream>>>(bv4(dz), bv4(z), bv4(x), act, mean, rstd, sum_dy, sum_dy_xhat, ppb, gsc, gsh)));
  }
  return check_launch("bn_bwd_reduce");
}

static int bn_bwd_apply_impl(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                             const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                             const double* sum_dy_xhat, float inv_n, int32_t batch_stats, const PmoeView4* dx,
                             const PmoeView4* dres, int32_t accumulate_dres, const float* fwd_scale, const float* fwd_shift,
                             double* next_s1, double* next_s2, const PmoeBnParamGrads* pgrads, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  ParamGrads pg = {nullptr, nullptr, 0, 0};
  if (pgrads && batch_stats) {
    if (pgrads->n < 0 || pgrads->n > dz->c) {
      set_error("bn_bwd_apply: parameter-gradient channel count out of range");
      return PMOE_ERR_ARG;
    }
    pg.dgamma = pgrads->dgamma;
    pg.dbeta = pgrads->dbeta;
    pg.n = pgrads->n;
    pg.accumulate = pgrads->accumulate;
  }
  const bool pre_x = host_piecewise(act) && (!z || !z->ptr) && x && x->ptr && fwd_scale && fwd_shift;
  const bool mask_x = pre_x && act == PMOE_ACT_RELU;
  if ((rc = chk(dz, dtype, "bn_bwd_apply dz"))) return rc;
  if ((rc = chk(z, dtype, "bn_bwd_apply z", act == PMOE_ACT_NONE || pre_x))) return rc;
  if (act == PMOE_ACT_HSWISH && !pre_x) {
    set_error("bn_bwd_apply: Hardswish needs the layer input x and the forward affine (its output does not determine the pre-activation)");
    return PMOE_ERR_ARG;
  }
  if ((rc = chk(x, dtype, "bn_bwd_apply x", !batch_stats))) return rc;
  if ((rc = chk(dx, dtype, "bn_bwd_apply dx", true))) return rc;
  if ((rc = chk(dres, dtype, "bn_bwd_apply dres", true))) return rc;
  if (batch_stats && (!mean || !rstd || !sum_dy || !sum_dy_xhat)) {
    set_error("bn_bwd_apply: batch statistics mode needs mean/rstd/sums");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dz->n * dz->h * dz->w * (dz->c / 8);
  const int grid = grid_cg(items, dz->c / 8);
  const bool flat_all = all_flat(dz, {dz, z, x, dx, dres});
  if (flat_all && dtype == PMOE_BF16 && (act == PMOE_ACT_NONE || act == PMOE_ACT_RELU || (act == PMOE_ACT_ELU && !next_s1))) {
    const uint4* pdz = static_cast<const uint4*>(dz->ptr);
    const uint4* pz = (act && !mask_x) ? static_cast<const uint4*>(z->ptr) : nullptr;
    const uint4* px = (batch_stats || mask_x) ? static_cast<const uint4*>(x->ptr) : nullptr;
    uint4* pdx = (dx && dx->ptr) ? static_cast<uint4*>(dx->ptr) : nullptr;
    uint4* pdr = (dres && dres->ptr) ? static_cast<uint4*>(dres->ptr) : nullptr;
    const int cg = dz->c / 8;
    int g4 = (grid + 3) / 4;  // four items per thread and iteration
    if (256 % cg != 0) g4 = (g4 + cg - 1) / cg * cg;
#define PMOE_APPLY_FAST(A, B) bn_bwd_apply_fast_kernel<A, B, false><<<g4, 256, 0, stream>>>(pdz, pz, px, pdx, pdr, accumulate_dres, items, cg, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, fwd_scale, fwd_shift, nullptr, nullptr, pg)
    if (next_s1) {
      if (!next_s2 || !batch_stats || !pdx || 256 % cg != 0 || !(mask_x || act == PMOE_ACT_RELU)) {
        set_error("bn_bwd_apply_sums: needs batch statistics, a dx output, ReLU and a channel-group count that divides 256");
        return PMOE_ERR_UNSUPPORTED;
      }
      if (mask_x)
        bn_bwd_apply_fast_kernel<2, true, true><<<g4, 256, 0, stream>>>(pdz, pz, px, pdx, pdr, accumulate_dres, items, cg, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, fwd_scale, fwd_shift, next_s1, next_s2, pg);
      else
        bn_bwd_apply_fast_kernel<1, true, true><<<g4, 256, 0, stream>>>(pdz, pz, px, pdx, pdr, accumulate_dres, items, cg, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, fwd_scale, fwd_shift, next_s1, next_s2, pg);
    } else if (mask_x && batch_stats) PMOE_APPLY_FAST(2, true);
    else if (mask_x) PMOE_APPLY_FAST(2, false);
    else if (act == PMOE_ACT_ELU && batch_stats) PMOE_APPLY_FAST(3, true);
    else if (act == PMOE_ACT_ELU) PMOE_APPLY_FAST(3, false);
    else if (act && batch_stats) PMOE_APPLY_FAST(1, true);
    else if (act) PMOE_APPLY_FAST(1, false);
    else if (batch_stats) PMOE_APPLY_FAST(0, true);
    else PMOE_APPLY_FAST(0, false);
#undef PMOE_APPLY_FAST
    return check_launch("bn_bwd_apply");
  }
  if (next_s1) {
    set_error("bn_bwd_apply: the fused-sums form needs contiguous bf16 tensors");
    return PMOE_ERR_UNSUPPORTED;
  }
  const float* gsc = pre_x ? fwd_scale : nullptr;
  const float* gsh = pre_x ? fwd_shift : nullptr;
  if (flat_all) {
    BW_DISPATCH(dtype, (bn_bwd_apply_kernel<T, true><<<grid, 256, 0, stream>>>(bv4(dz), bv4(z), bv4(x), act, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, batch_stats, bv4(dx), bv4(dres), accumulate_dres, pg, gsc, gsh)));
  } else {
    BW_DISPATCH(dtype, (bn_bwd_apply_kernel<T, false><<<grid, 256, 0, stream>>>(bv4(dz), bv4(z), bv4(x), act, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, batch_stats, bv4(dx), bv4(dres), accumulate_dres, pg, gsc, gsh)));
  }
  return check_launch("bn_bwd_apply");
}

int pmoe_bn_bwd_apply(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                      const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                      const double* sum_dy_xhat, float inv_n, int32_t batch_stats, const PmoeView4* dx,
                      const PmoeView4* dres, int32_t accumulate_dres, const float* fwd_scale, const float* fwd_shift,
                      const PmoeBnParamGrads* param_grads, pmoe_stream_t stream_) {
  return bn_bwd_apply_impl(dz, z, x, dtype, act, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, batch_stats, dx, dres, accumulate_dres,
                           fwd_scale, fwd_shift, nullptr, nullptr, param_grads, stream_);
}

int pmoe_bn_bwd_apply_sums(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                           const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                           const double* sum_dy_xhat, float inv_n, const PmoeView4* dx, const float* fwd_scale,
                           const float* fwd_shift, double* next_sum_dx, double* next_sum_dx_x, const PmoeBnParamGrads* param_grads,
                           pmoe_stream_t stream_) {
  if (!next_sum_dx || !next_sum_dx_x) {
    set_error("bn_bwd_apply_sums: both output sums are required");
    return PMOE_ERR_ARG;
  }
  return bn_bwd_apply_impl(dz, z, x, dtype, act, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, 1, dx, nullptr, 0, fwd_scale, fwd_shift,
                           next_sum_dx, next_sum_dx_x, param_grads, stream_);
}

int pmoe_maxpool_bwd(const PmoeView4* x, const PmoeView4* dy, const PmoeView4* dx, int32_t dtype, int32_t k, int32_t stride,
                     int32_t pad, int32_t accumulate, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = chk(x, dtype, "maxpool_bwd x"))) return rc;
  if ((rc = chk(dy, dtype, "maxpool_bwd dy"))) return rc;
  if ((rc = chk(dx, dtype, "maxpool_bwd dx"))) return rc;
  const long long items = (long long)dx->n * dx->h * dx->w * (dx->c / 8);
  BW_DISPATCH(dtype, (maxpool_bwd_kernel<T><<<bgrid(items, 256), 256, 0, stream>>>(bv4(x), bv4(dy), bv4(dx), k, stride, pad, accumulate)));
  return check_launch("maxpool_bwd");
}

int pmoe_maxpool_bwd_idx(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* dx, int32_t dtype, int32_t k,
                         int32_t stride, int32_t pad, int32_t accumulate, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = chk(dy, dtype, "maxpool_bwd_idx dy"))) return rc;
  if ((rc = chk(dx, dtype, "maxpool_bwd_idx dx"))) return rc;
  if (!idx || ((uintptr_t)idx & 7) || dy->n != dx->n || dy->c != dx->c || k < 1 || stride < 1) {
    set_error("maxpool_bwd_idx: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dx->n * dx->h * dx->w * (dx->c / 8);
  if (items >= (1LL << 31)) {
    set_error("maxpool_bwd_idx: more than 2^31 items");
    return PMOE_ERR_UNSUPPORTED;
  }
  const int cg = dx->c / 8;
  int cg_shift = -1;
  for (int sft = 0; sft < 12; ++sft)
    if ((1 << sft) == cg) cg_shift = sft;
  const long long rows = (long long)dx->n * dx->h;
  if (rows > 2147483647LL) {
    set_error("maxpool_bwd_idx: too many rows");
    return PMOE_ERR_UNSUPPORTED;
  }
  dim3 grid((unsigned)rows, (unsigned)((dx->w * cg + 255) / 256));
  const bool k3 = k == 3 && stride == 2 && pad == 1, k2 = k == 2 && stride == 2 && pad == 0;
  if (dtype == PMOE_BF16 && (k3 || k2) && all_flat(dx, {dx}) && all_flat(dy, {dy})) {
    const uint4* pdy = static_cast<const uint4*>(dy->ptr);
    const uint2* pidx = reinterpret_cast<const uint2*>(idx);
    uint4* pdx = static_cast<uint4*>(dx->ptr);
    if (k3 && dx->h % 2 == 0 && dx->w % 2 == 0 && dy->h == dx->h / 2 && dy->w == dx->w / 2) {
      dim3 gp((unsigned)rows, (unsigned)(((dx->w / 2) * cg + 255) / 256));
      maxpool3s2_bwd_pair_kernel<<<gp, 256, 0, stream>>>(pdy, pidx, pdx, dx->h, dx->w, dy->h, dy->w, cg, cg_shift, accumulate);
    } else if (k3) maxpool_bwd_fast_kernel<3, 2, 1><<<grid, 256, 0, stream>>>(pdy, pidx, pdx, dx->h, dx->w, dy->h, dy->w, cg, cg_shift, accumulate);
    else maxpool_bwd_fast_kernel<2, 2, 0><<<grid, 256, 0, stream>>>(pdy, pidx, pdx, dx->h, dx->w, dy->h, dy->w, cg, cg_shift, accumulate);
    return check_launch("maxpool_bwd_idx");
  }
  BW_DISPATCH(dtype, (maxpool_bwd_idx_kernel<T><<<grid, 256, 0, stream>>>(bv4(dy), idx, bv4(dx), k, stride, pad, accumulate, cg_shift)));
  return check_launch("maxpool_bwd_idx");
}

int pmoe_prod_channel_sums(const PmoeView4* a, const PmoeView4* b, int32_t dtype, double* out, int64_t out_stride,
                           pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = chk(a, dtype, "prod_channel_sums a"))) return rc;
  if ((rc = chk(b, dtype, "prod_channel_sums b"))) return rc;
  if (!out || a->c / 8 > 256) {
    set_error("prod_channel_sums: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long hw = (long long)a->h * a->w;
  const int rows = 1024;
  dim3 grid((unsigned)((hw + rows - 1) / rows), (unsigned)a->n);
  BW_DISPATCH(dtype, (prod_channel_sums_kernel<T><<<grid, 256, 0, stream>>>(bv4(a), bv4(b), out, out_stride, rows)));
  return check_launch("prod_channel_sums");
}

int pmoe_eca_gate_bwd(const double* dgate, int64_t dgate_stride, const float* gate, int64_t gate_stride, const float* pool_sum,
                      int64_t pool_stride, int32_t n, float inv_count, const float* w, int32_t k, int32_t groups,
                      int32_t group_c, int32_t group_stride, float* dmean, int64_t dmean_stride, double* dw,
                      pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!dgate || !gate || !pool_sum || !w || !dmean || n < 1) {
    set_error("eca_gate_bwd: bad arguments");
    return PMOE_ERR_ARG;
  }
  eca_gate_bwd_kernel<<<n, 128, groups * group_c * sizeof(double), stream>>>(dgate, dgate_stride, gate, gate_stride, pool_sum,
                                                                            pool_stride, inv_count, w, k, groups, group_c,
                                                                            group_stride, dmean, dmean_stride, dw);
  return check_launch("eca_gate_bwd");
}

static int eca_bwd_apply_impl(const PmoeView4* dout, int32_t dtype, const float* gate, int64_t gate_stride, const float* dmean,
                              int64_t dmean_stride, const PmoeView4* dx, int32_t accumulate, const PmoeView4* xf, double* next_s1,
                              double* next_s2, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = chk(dout, dtype, "eca_bwd_apply dout"))) return rc;
  if ((rc = chk(dx, dtype, "eca_bwd_apply dx"))) return rc;
  if (!gate || !dmean || ((uintptr_t)gate % 16) || ((uintptr_t)dmean % 16) || gate_stride % 4 || dmean_stride % 4) {
    set_error("eca_bwd_apply: gate/dmean must be 16-byte aligned rows");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dx->n * dx->h * dx->w * (dx->c / 8);
  if (dtype == PMOE_BF16 && all_flat(dx, {dx, dout}) && dx->n <= 65535) {
    const int cg = dx->c / 8;
    const long long per_img = (long long)dx->h * dx->w * cg;
    long long bx = (per_img + 4 * 256 - 1) / (4 * 256);
    const long long cap = ((long long)num_sms() * 16 + dx->n - 1) / dx->n;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (256 % cg != 0) bx = (bx + cg - 1) / cg * cg;
    dim3 grid((unsigned)bx, (unsigned)dx->n);
    const uint4* pd = static_cast<const uint4*>(dout->ptr);
    uint4* px = static_cast<uint4*>(dx->ptr);
    if (next_s1) {
      if (accumulate || 256 % cg != 0 || !xf || !xf->ptr || !all_flat(dx, {xf})) {
        set_error("eca_bwd_apply_sums: needs a fresh dx, a dense forward input of dx's shape and a channel-group count dividing 256");
        return PMOE_ERR_UNSUPPORTED;
      }
      eca_bwd_apply_fast_kernel<false, true><<<grid, 256, 0, stream>>>(pd, px, per_img, cg, gate, gate_stride, dmean, dmean_stride,
                                                                        static_cast<const uint4*>(xf->ptr), next_s1, next_s2);
    } else if (accumulate) {
      eca_bwd_apply_fast_kernel<true, false><<<grid, 256, 0, stream>>>(pd, px, per_img, cg, gate, gate_stride, dmean, dmean_stride, nullptr, nullptr, nullptr);
    } else {
      eca_bwd_apply_fast_kernel<false, false><<<grid, 256, 0, stream>>>(pd, px, per_img, cg, gate, gate_stride, dmean, dmean_stride, nullptr, nullptr, nullptr);
    }
    return check_launch("eca_bwd_apply");
  }
  if (next_s1) {
    set_error("eca_bwd_apply_sums: dense bf16 tensors only");
    return PMOE_ERR_UNSUPPORTED;
  }
  BW_DISPATCH(dtype, (eca_bwd_apply_kernel<T><<<bgrid(items, 256), 256, 0, stream>>>(bv4(dout), gate, gate_stride, dmean, dmean_stride, bv4(dx), accumulate)));
  return check_launch("eca_bwd_apply");
}

int pmoe_eca_bwd_apply(const PmoeView4* dout, int32_t dtype, const float* gate, int64_t gate_stride, const float* dmean,
                       int64_t dmean_stride, const PmoeView4* dx, int32_t accumulate, pmoe_stream_t stream_) {
  return eca_bwd_apply_impl(dout, dtype, gate, gate_stride, dmean, dmean_stride, dx, accumulate, nullptr, nullptr, nullptr, stream_);
}

int pmoe_eca_bwd_apply_sums(const PmoeView4* dout, int32_t dtype, const float* gate, int64_t gate_stride, const float* dmean,
                            int64_t dmean_stride, const PmoeView4* dx, const PmoeView4* x_fwd, double* next_sum_dx,
                            double* next_sum_dx_x, pmoe_stream_t stream_) {
  if (!next_sum_dx || !next_sum_dx_x) {
    set_error("eca_bwd_apply_sums: both output sums are required");
    return PMOE_ERR_ARG;
  }
  return eca_bwd_apply_impl(dout, dtype, gate, gate_stride, dmean, dmean_stride, dx, 0, x_fwd, next_sum_dx, next_sum_dx_x, stream_);
}

int pmoe_axpy(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, float alpha, const float* bcast, int64_t bcast_stride,
              int32_t accumulate, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = chk(src, dtype, "axpy src", true))) return rc;
  if ((rc = chk(dst, dtype, "axpy dst"))) return rc;
  if (bcast && (((uintptr_t)bcast % 16) || bcast_stride % 4)) {
    set_error("axpy: broadcast rows must be 16-byte aligned");
    return PMOE_ERR_ARG;
  }
  const long long items = (long long)dst->n * dst->h * dst->w * (dst->c / 8);
  BW_DISPATCH(dtype, (axpy_kernel<T><<<bgrid(items, 256), 256, 0, stream>>>(bv4(src), bv4(dst), alpha, bcast, bcast_stride, accumulate)));
  return check_launch("axpy");
}

}  // extern "C"
