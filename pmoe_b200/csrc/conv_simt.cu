// CUDA-core implicit-GEMM convolution kernels that share the segment-list descriptor of conv_tc:
//   * pmoe_conv_simt   — same contract as pmoe_conv_tc for fp32 (or bf16) activations/weights with fp32
//                        FMA accumulation: the fp32-precision mode of the library (<= 1e-4 parity target),
//                        and the fallback-free path for shapes the tensor-core kernel does not take.
//   * pmoe_conv_wgrad_simt — weight gradient in the PACKED layout: dW[co][k] += sum_pixels dy[p][co] * x_k[p],
//                        k enumerating (segment, chunk, channel) exactly as the forward's wpack.
// Tiles: 64 pixels x 64 output channels x 16 K per step, 256 threads, 4x4 register block each.
#include "host_util.h"
#include "act.cuh"
#include "ptx.cuh"

#include <string.h>

namespace pmoe {

struct SimtSeg {
  int src, dh, dw, c0, nchunks;
};

struct SimtView {
  const void* ptr;
  int n, h, w, c;
  long long sn, sh, sw;
};

struct SimtParams {
  SimtView src[PMOE_MAX_SRC];
  SimtSeg seg[PMOE_MAX_SEG];
  int n_seg, ck, ktot, cout_pad;
  const void* wpack;
  SimtView out;  // forward: output; wgrad: dy
  const float* scale;
  const float* shift;
  int act;
  SimtView res;
  double* stat_sum;
  double* stat_sq;
  float* pool_sum;
  int pool_stride;
  int ph, pw, tiles_h, tiles_w;  // pixel patch of a tile
  float* dw;                     // wgrad output [cout_pad][ktot]
  long long pix_per_split;
};

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) {
  return __ldg(p);
}
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

template <typename T>
__device__ __forceinline__ float round_as(float v);
template <>
__device__ __forceinline__ float round_as<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float round_as<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

__device__ __forceinline__ float simt_act(float x, int act) {
  switch (act) {
    case PMOE_ACT_RELU: return fmaxf(x, 0.f);
    case PMOE_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case PMOE_ACT_TANH: return tanhf(x);
    case PMOE_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return act_piecewise(x, act);
  }
}

// locate the 16-channel sub-chunk `kc` (index into the packed K axis / 16)
__device__ __forceinline__ void locate_subchunk(const SimtParams& p, int kc, int& src, int& dh, int& dw, int& c) {
  const int per = p.ck / 16;
  int chunk = kc / per, sub = kc % per;
  int s = 0;
  while (s < p.n_seg - 1 && chunk >= p.seg[s].nchunks) {
    chunk -= p.seg[s].nchunks;
    ++s;
  }
  src = p.seg[s].src;
  dh = p.seg[s].dh;
  dw = p.seg[s].dw;
  c = p.seg[s].c0 + chunk * p.ck + sub * 16;
}

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ SimtParams p) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  __shared__ float s_red[2][64];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int tiles_per_img = p.tiles_h * p.tiles_w;
  const int img = blockIdx.x / tiles_per_img;
  const int rem = blockIdx.x % tiles_per_img;
  const int h0 = (rem / p.tiles_w) * p.ph, w0 = (rem % p.tiles_w) * p.pw;
  const int n0 = blockIdx.y * 64;

  // A-load role: pixel lp = tid / 4, channels (tid % 4) * 4 .. +3 of the 16-channel sub-chunk
  const int lp = tid / 4, lc = (tid % 4) * 4;
  const int lph = h0 + lp / p.pw, lpw = w0 + lp % p.pw;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = p.ktot / 16;
  for (int kc = 0; kc < nk; ++kc) {
    int src, dh, dw, c;
    locate_subchunk(p, kc, src, dh, dw, c);
    const SimtView& sv = p.src[src];
    const int ih = lph + dh, iw = lpw + dw;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    if (lph < p.out.h && lpw < p.out.w && ih >= 0 && ih < sv.h && iw >= 0 && iw < sv.w) {
      const T* ap = static_cast<const T*>(sv.ptr) + img * sv.sn + ih * sv.sh + iw * sv.sw + c + lc;
#pragma unroll
      for (int q = 0; q < 4; ++q) a4[q] = ldf<T>(ap + q);
    }
    float b4[4] = {0.f, 0.f, 0.f, 0.f};
    if (n0 + lp < p.cout_pad) {
      const T* bp = static_cast<const T*>(p.wpack) + (long long)(n0 + lp) * p.ktot + kc * 16 + lc;
#pragma unroll
      for (int q = 0; q < 4; ++q) b4[q] = ldf<T>(bp + q);
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      As[lc + q][lp] = a4[q];
      Bs[lc + q][lp] = b4[q];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }

  // epilogue
  if (tid < 64) {
    s_red[0][tid] = 0.f;
    s_red[1][tid] = 0.f;
  }
  __syncthreads();
  float st_sum[4] = {0, 0, 0, 0}, st_sq[4] = {0, 0, 0, 0}, pl[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pix = ty * 4 + i;
    const int oh = h0 + pix / p.pw, ow = w0 + pix % p.pw;
    const bool valid = pix < p.ph * p.pw && oh < p.out.h && ow < p.out.w;
    if (!valid) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= p.cout_pad) continue;
      const float raw = acc[i][j];
      st_sum[j] += raw;
      st_sq[j] += raw * raw;
      float y = raw * (p.scale ? __ldg(p.scale + co) : 1.f) + (p.shift ? __ldg(p.shift + co) : 0.f);
      if (p.res.ptr && co < p.res.c)
        y += ldf<T>(static_cast<const T*>(p.res.ptr) + img * p.res.sn + oh * p.res.sh + ow * p.res.sw + co);
      y = simt_act(y, p.act);
      if (co < p.out.c) {
        T* op = static_cast<T*>(const_cast<void*>(p.out.ptr)) + img * p.out.sn + oh * p.out.sh + ow * p.out.sw + co;
        stf<T>(op, y);
        pl[j] += round_as<T>(y);  // pool what is stored
      }
    }
  }
  if (p.stat_sum) {
    // fixed summation order inside the block (the fp32 parity mode must not depend on the order shared-memory atomics land in:
    // a 1e-7 difference in a batch statistic flips ReLU masks downstream); the operand tiles are free by now
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[ty][tx * 4 + j] = st_sum[j];
      Bs[ty][tx * 4 + j] = st_sq[j];
    }
    __syncthreads();
    if (tid < 64 && n0 + tid < p.cout_pad) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        a += (double)As[r][tid];
        b += (double)Bs[r][tid];
      }
      atomicAdd(p.stat_sum + n0 + tid, a);   // fp64 across blocks: order effects are 1e-16
      atomicAdd(p.stat_sq + n0 + tid, b);
    }
    __syncthreads();
  }
  if (p.pool_sum) {
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(&s_red[0][tx * 4 + j], pl[j]);
    __syncthreads();
    if (tid < 64 && n0 + tid < p.cout_pad) atomicAdd(p.pool_sum + (long long)img * p.pool_stride + n0 + tid, s_red[0][tid]);
  }
}

// dW[co][k] += sum_p dy[p][co] * x_k[p]; grid = (ktot/64 rounded up, cout_pad/64 rounded up, splits)
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const __grid_constant__ SimtParams p) {
  __shared__ float Ds[16][64 + 4];  // dy  [pixel][cout]
  __shared__ float Xs[16][64 + 4];  // x   [pixel][k]
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // tx -> 4 k's, ty -> 4 couts
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const long long npix = (long long)p.out.n * p.out.h * p.out.w;
  const long long p_begin = (long long)blockIdx.z * p.pix_per_split;
  long long p_end = p_begin + p.pix_per_split;
  if (p_end > npix) p_end = npix;

  // load roles: pixel lp = tid / 16 (0..15); dy: couts (tid%16)*4..+3 ; x: k's (tid%16)*4..+3
  const int lp = tid / 16, lq = (tid % 16) * 4;
  int xsrc = 0, xdh = 0, xdw = 0, xc = 0;
  const bool k_ok = k0 + lq < p.ktot;
  if (k_ok) {
    locate_subchunk(p, (k0 + lq) / 16, xsrc, xdh, xdw, xc);
    xc += (k0 + lq) % 16;
  }
  const SimtView sv = p.src[xsrc];
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pb = p_begin; pb < p_end; pb += 16) {
    const long long pp = pb + lp;
    float d4[4] = {0, 0, 0, 0}, x4[4] = {0, 0, 0, 0};
    if (pp < p_end) {
      const int w = (int)(pp % p.out.w);
      const long long t = pp / p.out.w;
      const int h = (int)(t % p.out.h), n = (int)(t / p.out.h);
      if (n0 + lq < p.out.c) {
        const T* dp = static_cast<const T*>(p.out.ptr) + n * p.out.sn + h * p.out.sh + w * p.out.sw + n0 + lq;
#pragma unroll
        for (int q = 0; q < 4; ++q) d4[q] = ldf<T>(dp + q);
      }
      const int ih = h + xdh, iw = w + xdw;
      if (k_ok && ih >= 0 && ih < sv.h && iw >= 0 && iw < sv.w) {
        const T* xp = static_cast<const T*>(sv.ptr) + n * sv.sn + ih * sv.sh + iw * sv.sw + xc;
#pragma unroll
        for (int q = 0; q < 4; ++q) x4[q] = ldf<T>(xp + q);
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Ds[lp][lq + q] = d4[q];
      Xs[lp][lq + q] = x4[q];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float dv[4], xv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) dv[i] = Ds[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = Xs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], xv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = n0 + ty * 4 + i;
    if (co >= p.cout_pad) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < p.ktot) atomicAdd(p.dw + (long long)co * p.ktot + k, acc[i][j]);
    }
  }
}

static SimtView to_sv(const PmoeView4& v) { return SimtView{v.ptr, v.n, v.h, v.w, v.c, v.sn, v.sh, v.sw}; }

static int fill_common(SimtParams& p, const PmoeConvTc* d, const char* what) {
  if (!d || d->n_src < 1 || d->n_src > PMOE_MAX_SRC || d->n_seg < 1 || d->n_seg > PMOE_MAX_SEG || !d->out.ptr) {
    set_error("%s: bad descriptor", what);
    return PMOE_ERR_ARG;
  }
  if (d->ck != 16 && d->ck != 32 && d->ck != 64) {
    set_error("%s: ck must be 16, 32 or 64", what);
    return PMOE_ERR_ARG;
  }
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < d->n_src; ++i) p.src[i] = to_sv(d->src[i]);
  int kiters = 0;
  for (int i = 0; i < d->n_seg; ++i) {
    const PmoeSeg& s = d->seg[i];
    if (s.src < 0 || s.src >= d->n_src || s.nchunks == 0 || s.c0 + s.nchunks * d->ck > d->src[s.src].c) {
      set_error("%s: segment %d out of range", what, i);
      return PMOE_ERR_ARG;
    }
    p.seg[i] = SimtSeg{s.src, s.dh, s.dw, s.c0, s.nchunks};
    kiters += s.nchunks;
  }
  if (kiters * d->ck != d->ktot) {
    set_error("%s: ktot does not match the segment list", what);
    return PMOE_ERR_ARG;
  }
  p.n_seg = d->n_seg;
  p.ck = d->ck;
  p.ktot = d->ktot;
  p.cout_pad = d->cout_pad;
  p.out = to_sv(d->out);
  return PMOE_OK;
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_conv_simt(const PmoeConvTc* d, int32_t dtype, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SimtParams p;
  int rc = fill_common(p, d, "conv_simt");
  if (rc) return rc;
  if (!d->wpack) {
    set_error("conv_simt: wpack missing");
    return PMOE_ERR_ARG;
  }
  p.wpack = d->wpack;
  p.scale = d->scale;
  p.shift = d->shift;
  p.act = d->act;
  if (d->residual.ptr) p.res = to_sv(d->residual);
  p.stat_sum = d->stat_sum;
  p.stat_sq = d->stat_sqsum;
  p.pool_sum = d->pool_sum;
  p.pool_stride = d->pool_stride > 0 ? d->pool_stride : d->cout_pad;
  const int H = d->out.h;
  if (H >= 8) { p.ph = 8; p.pw = 8; }
  else if (H >= 4) { p.ph = 4; p.pw = 16; }
  else if (H >= 2) { p.ph = 2; p.pw = 32; }
  else { p.ph = 1; p.pw = 64; }
  p.tiles_h = (d->out.h + p.ph - 1) / p.ph;
  p.tiles_w = (d->out.w + p.pw - 1) / p.pw;
  const long long tiles = (long long)p.tiles_h * p.tiles_w * d->out.n;
  if (tiles <= 0 || tiles > 0x7fffffffLL) {
    set_error("conv_simt: bad tile count");
    return PMOE_ERR_ARG;
  }
  dim3 grid((unsigned)tiles, (unsigned)((d->cout_pad + 63) / 64));
  if (dtype == PMOE_F32) conv_simt_kernel<float><<<grid, 256, 0, stream>>>(p);
  else if (dtype == PMOE_BF16) conv_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  else {
    set_error("conv_simt: unsupported dtype");
    return PMOE_ERR_ARG;
  }
  return check_launch("conv_simt");
}

extern "C" int pmoe_conv_wgrad_simt(const PmoeConvTc* d, int32_t dtype, float* dwpack, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SimtParams p;
  int rc = fill_common(p, d, "conv_wgrad_simt");
  if (rc) return rc;
  if (!dwpack) {
    set_error("conv_wgrad_simt: output missing");
    return PMOE_ERR_ARG;
  }
  p.dw = dwpack;
  const long long npix = (long long)d->out.n * d->out.h * d->out.w;
  const int gx = (d->ktot + 63) / 64, gy = (d->cout_pad + 63) / 64;
  long long want = ((long long)num_sms() * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
  long long max_splits = (npix + 255) / 256;
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 65535) want = 65535;
  p.pix_per_split = ((npix + want - 1) / want + 15) / 16 * 16;
  const int splits = (int)((npix + p.pix_per_split - 1) / p.pix_per_split);
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)splits);
  if (dtype == PMOE_F32) conv_wgrad_simt_kernel<float><<<grid, 256, 0, stream>>>(p);
  else if (dtype == PMOE_BF16) conv_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(p);
  else {
    set_error("conv_wgrad_simt: unsupported dtype");
    return PMOE_ERR_ARG;
  }
  return check_launch("conv_wgrad_simt");
}
