// Optimizer-side multi-tensor kernels: global gradient norm, gradient clipping and Adam(amsgrad) over a table of
// (parameter, gradient, state) chunks — one launch each instead of ~5 elementwise launches and one host sync per
// parameter tensor (reference: trainer/train_2.py:160-165, utils/nn.py:10-19 -> torch.nn.utils.clip_grad_norm_,
// torch.optim.Adam). HBM-bound: Adam(amsgrad) moves 5 reads + 4 writes of 4 B per parameter.
#include "host_util.h"

namespace pmoe {

struct MtChunk {  // must match PmoeMtChunk
  float* p;
  float* g;
  float* m;
  float* v;
  float* vmax;
  int32_t n;
  int32_t pad;
};

__device__ __forceinline__ double block_sum(double x) {
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  if (lane == 0) sh[w] = x;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  x = threadIdx.x < nw ? sh[threadIdx.x] : 0.0;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  }
  return x;
}

__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const MtChunk* __restrict__ chunks, int n_chunks, double* __restrict__ out) {
  double acc = 0.0;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const MtChunk ch = chunks[c];
    const float* g = ch.g;
    float part = 0.f;
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const int n4 = ch.n >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(g);
      for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        const float4 x = __ldg(g4 + i);
        part += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < ch.n; i += blockDim.x) part += g[i] * g[i];
    } else {
      for (int i = threadIdx.x; i < ch.n; i += blockDim.x) part += g[i] * g[i];
    }
    acc += (double)part;
  }
  const double tot = block_sum(acc);
  if (threadIdx.x == 0 && tot != 0.0) atomicAdd(out, tot);
}

// clip_coef = min(1, max_norm / (sqrt(sqnorm) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__device__ __forceinline__ float clip_coef_of(const double* sqnorm, float max_norm) {
  if (sqnorm == nullptr || max_norm <= 0.f) return 1.f;
  const float total = (float)sqrt(*sqnorm);
  const float c = max_norm / (total + 1e-6f);
  return c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(256) mt_scale_kernel(const MtChunk* __restrict__ chunks, int n_chunks, const double* sqnorm,
                                                       float max_norm) {
  const float coef = clip_coef_of(sqnorm, max_norm);
  if (coef == 1.f) return;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const MtChunk ch = chunks[c];
    for (int i = threadIdx.x; i < ch.n; i += blockDim.x) ch.g[i] *= coef;
  }
}

struct AdamArgs {
  float step_size, beta1, beta2, om_beta1, om_beta2, eps, weight_decay, bias_c2_sqrt, max_norm;
  int amsgrad;
};

template <bool VEC>
__device__ __forceinline__ void adam_chunk(const MtChunk& ch, const AdamArgs& a, float coef) {
  const float step_size = a.step_size;
  auto upd = [&](float& p, float g, float& m, float& v, float& vm) {
    g *= coef;
    if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
    m = m + (g - m) * a.om_beta1;                   // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(g * g, a.om_beta2, v * a.beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    float vv = v;
    if (a.amsgrad) {
      vm = fmaxf(vm, v);
      vv = vm;
    }
    const float denom = sqrtf(vv) / a.bias_c2_sqrt + a.eps;
    p = p - step_size * (m / denom);
  };
  if constexpr (VEC) {
    const int n4 = ch.n >> 2;
    float4* p4 = reinterpret_cast<float4*>(ch.p);
    const float4* g4 = reinterpret_cast<const float4*>(ch.g);
    float4* m4 = reinterpret_cast<float4*>(ch.m);
    float4* v4 = reinterpret_cast<float4*>(ch.v);
    float4* x4 = reinterpret_cast<float4*>(ch.vmax);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 p = p4[i], m = m4[i], v = v4[i];
      const float4 g = __ldg(g4 + i);
      float4 x = a.amsgrad ? x4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      upd(p.x, g.x, m.x, v.x, x.x);
      upd(p.y, g.y, m.y, v.y, x.y);
      upd(p.z, g.z, m.z, v.z, x.z);
      upd(p.w, g.w, m.w, v.w, x.w);
      p4[i] = p;
      m4[i] = m;
      v4[i] = v;
      if (a.amsgrad) x4[i] = x;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < ch.n; i += blockDim.x) {
      float dummy = 0.f;
      upd(ch.p[i], ch.g[i], ch.m[i], ch.v[i], a.amsgrad ? ch.vmax[i] : dummy);
    }
  } else {
    for (int i = threadIdx.x; i < ch.n; i += blockDim.x) {
      float dummy = 0.f;
      upd(ch.p[i], ch.g[i], ch.m[i], ch.v[i], a.amsgrad ? ch.vmax[i] : dummy);
    }
  }
}

__global__ void __launch_bounds__(256) mt_adam_kernel(const MtChunk* __restrict__ chunks, int n_chunks, const AdamArgs a,
                                                      const double* sqnorm) {
  const float coef = clip_coef_of(sqnorm, a.max_norm);
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const MtChunk ch = chunks[c];
    const uintptr_t al = reinterpret_cast<uintptr_t>(ch.p) | reinterpret_cast<uintptr_t>(ch.g) | reinterpret_cast<uintptr_t>(ch.m) |
                         reinterpret_cast<uintptr_t>(ch.v) | reinterpret_cast<uintptr_t>(ch.vmax);
    if ((al & 15) == 0) adam_chunk<true>(ch, a, coef);
    else adam_chunk<false>(ch, a, coef);
  }
}


// torch.optim.RMSprop (conf/stage_2.yaml:147-153: centered, alpha 0.99, momentum 0) over all chunks. Chunk fields:
// p = parameter, g = gradient, m = square_avg, v = momentum_buffer (or null), vmax = grad_avg (centered, or null).
struct RmsArgs {
  float lr, alpha, om_alpha, eps, weight_decay, momentum, max_norm;
};
__global__ void __launch_bounds__(256) mt_rmsprop_kernel(const MtChunk* __restrict__ chunks, int n_chunks, const RmsArgs a,
                                                         const double* sqnorm) {
  const float coef = clip_coef_of(sqnorm, a.max_norm);
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const MtChunk ch = chunks[c];
    for (int i = threadIdx.x; i < ch.n; i += blockDim.x) {
      float p = ch.p[i];
      float g = ch.g[i] * coef;
      if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
      const float sq = fmaf(a.om_alpha * g, g, ch.m[i] * a.alpha);  // square_avg.mul_(alpha).addcmul_(g, g, value=1-alpha)
      ch.m[i] = sq;
      float avg;
      if (ch.vmax) {  // centered: grad_avg.lerp_(g, 1-alpha); avg = sqrt(square_avg - grad_avg^2)
        const float ga0 = ch.vmax[i];
        const float ga = fmaf(a.om_alpha, g - ga0, ga0);
        ch.vmax[i] = ga;
        avg = sqrtf(fmaf(-ga, ga, sq));
      } else {
        avg = sqrtf(sq);
      }
      avg += a.eps;
      if (ch.v) {  // momentum: buf.mul_(momentum).addcdiv_(g, avg); p.add_(buf, alpha=-lr)
        const float buf = fmaf(ch.v[i], a.momentum, g / avg);
        ch.v[i] = buf;
        p = fmaf(-a.lr, buf, p);
      } else {
        p = fmaf(-a.lr, g / avg, p);  // p.addcdiv_(g, avg, value=-lr)
      }
      ch.p[i] = p;
    }
  }
}

// torch.optim.swa_utils.AveragedModel.update_parameters with the default avg_fn (train_2.py:119-121, 179-187):
// p_swa += (p_model - p_swa) / (n_averaged + 1). Chunk fields: p = averaged parameter, g = current model parameter.
__global__ void __launch_bounds__(256) mt_swa_kernel(const MtChunk* __restrict__ chunks, int n_chunks, float denom) {
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const MtChunk ch = chunks[c];
    for (int i = threadIdx.x; i < ch.n; i += blockDim.x) {
      const float avg = ch.p[i];
      ch.p[i] = avg + (ch.g[i] - avg) / denom;
    }
  }
}

static int mt_grid(int n_chunks) {
  const int cap = num_sms() * 8;
  return n_chunks < cap ? n_chunks : cap;
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_mt_sqnorm(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double* sqnorm_accum, pmoe_stream_t stream_) {
  if (n_chunks <= 0) return PMOE_OK;
  if (!chunks_dev || !sqnorm_accum) {
    set_error("mt_sqnorm: null argument");
    return PMOE_ERR_ARG;
  }
  mt_sqnorm_kernel<<<mt_grid(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const MtChunk*>(chunks_dev),
                                                                                      n_chunks, sqnorm_accum);
  return check_launch("mt_sqnorm");
}

extern "C" int pmoe_mt_clip(const PmoeMtChunk* chunks_dev, int32_t n_chunks, const double* sqnorm, float max_norm,
                            pmoe_stream_t stream_) {
  if (n_chunks <= 0) return PMOE_OK;
  if (!chunks_dev || !sqnorm) {
    set_error("mt_clip: null argument");
    return PMOE_ERR_ARG;
  }
  mt_scale_kernel<<<mt_grid(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const MtChunk*>(chunks_dev),
                                                                                     n_chunks, sqnorm, max_norm);
  return check_launch("mt_clip");
}

extern "C" int pmoe_mt_adam(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double lr, double beta1, double beta2, double eps,
                            double weight_decay, int32_t step, int32_t amsgrad, const double* sqnorm, float max_norm,
                            pmoe_stream_t stream_) {
  if (n_chunks <= 0) return PMOE_OK;
  if (!chunks_dev || step < 1) {
    set_error("mt_adam: null chunk table or step < 1");
    return PMOE_ERR_ARG;
  }
  AdamArgs a;
  // hyper-parameters arrive as doubles (Python floats) and are rounded once, like torch's scalar arguments:
  // float(1 - 0.999) is 0.001f, while 1.f - 0.999f is 1.3e-5 off it
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.om_beta1 = (float)(1.0 - beta1);
  a.om_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.weight_decay = (float)weight_decay;
  a.step_size = (float)(lr / (1.0 - pow(beta1, (double)step)));
  a.bias_c2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.max_norm = sqnorm ? max_norm : 0.f;
  a.amsgrad = amsgrad;
  mt_adam_kernel<<<mt_grid(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const MtChunk*>(chunks_dev),
                                                                                    n_chunks, a, sqnorm);
  return check_launch("mt_adam");
}

extern "C" int pmoe_mt_rmsprop(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double lr, double alpha, double eps, double weight_decay,
                               double momentum, const double* sqnorm, float max_norm, pmoe_stream_t stream_) {
  if (n_chunks <= 0) return PMOE_OK;
  if (!chunks_dev) {
    set_error("mt_rmsprop: null chunk table");
    return PMOE_ERR_ARG;
  }
  RmsArgs a;
  a.lr = (float)lr;
  a.alpha = (float)alpha;
  a.om_alpha = (float)(1.0 - alpha);
  a.eps = (float)eps;
  a.weight_decay = (float)weight_decay;
  a.momentum = (float)momentum;
  a.max_norm = max_norm;
  mt_rmsprop_kernel<<<mt_grid(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const MtChunk*>(chunks_dev),
                                                                                       n_chunks, a, sqnorm);
  return check_launch("mt_rmsprop");
}

extern "C" int pmoe_mt_swa_update(const PmoeMtChunk* chunks_dev, int32_t n_chunks, int64_t n_averaged, pmoe_stream_t stream_) {
  if (n_chunks <= 0) return PMOE_OK;
  if (!chunks_dev || n_averaged < 0) {
    set_error("mt_swa_update: bad arguments");
    return PMOE_ERR_ARG;
  }
  mt_swa_kernel<<<mt_grid(n_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const MtChunk*>(chunks_dev), n_chunks,
                                                                                   (float)(n_averaged + 1));
  return check_launch("mt_swa_update");
}
