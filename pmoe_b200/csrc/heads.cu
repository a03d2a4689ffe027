// MoE gating / mixture head and the small losses (k10, k12 of SURVEY.md §8a):
//   * gate_mixture_fwd/bwd : softmax over the K expert logits (warp-shuffle free: K <= 32 values per row live in
//                            registers), sigma = ELU(raw)+1 evaluated as x>0 ? x+1 : exp(x) in fp32, routing index =
//                            argmax_k (lowest index wins ties, like torch.argmax)
//   * moe_loss_fwd_bwd     : mixture-of-Gaussians NLL (+ speed MSE /K) and its analytic gradient in ONE pass
//   * dropout              : counter-based (stateless) Bernoulli mask, regenerated in backward
//   * l1 / mse             : mean reductions with gradients
// All small tensors are fp32 except the strided head outputs, which may be bf16 (they come out of the GEMM epilogue).
#include "host_util.h"
#include "ptx.cuh"

namespace pmoe {

constexpr int kMaxExperts = 32;

template <typename T>
__device__ __forceinline__ float hload(const T* p);
template <>
__device__ __forceinline__ float hload<float>(const float* p) {
  return *p;
}
template <>
__device__ __forceinline__ float hload<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void hstore(T* p, float v);
template <>
__device__ __forceinline__ void hstore<float>(float* p, float v) {
  *p = v;
}
template <>
__device__ __forceinline__ void hstore<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// alpha: element (b,k) at alpha[b*a_sb + k*a_sk]; ap: (b,k,j) j<4 at ap[b*p_sb + k*p_sk + j] = [mu0, mu1, raw_s0, raw_s1]
template <typename T>
__global__ void gate_mixture_fwd_kernel(const T* __restrict__ alpha, long long a_sb, long long a_sk, const T* __restrict__ ap,
                                        long long p_sb, long long p_sk, int B, int K, int relu_alpha,
                                        float* __restrict__ probs, float* __restrict__ mean, float* __restrict__ std,
                                        long long* __restrict__ route) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float a[kMaxExperts];
  float mx = -INFINITY;
  int arg = 0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float v = hload<T>(alpha + b * a_sb + k * a_sk);
    if (relu_alpha) v = fmaxf(v, 0.f);
    a[k] = v;
    if (v > mx) {
      mx = v;
      arg = k;
    }
  }
  float s = 0.f;
  for (int k = 0; k < K; ++k) {
    a[k] = expf(a[k] - mx);
    s += a[k];
  }
  const float inv = 1.f / s;
  for (int k = 0; k < K; ++k) {
    probs[(long long)b * K + k] = a[k] * inv;
    const T* q = ap + b * p_sb + k * p_sk;
    mean[((long long)b * K + k) * 2 + 0] = hload<T>(q + 0);
    mean[((long long)b * K + k) * 2 + 1] = hload<T>(q + 1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float r = hload<T>(q + 2 + j);
      std[((long long)b * K + k) * 2 + j] = r > 0.f ? r + 1.f : expf(r);  // ELU(r)+1 without the bf16/1-cancellation
    }
  }
  if (route) route[b] = arg;
}

// gradients w.r.t. the raw head outputs, written in the same strided layout (dtype T)
template <typename T>
__global__ void gate_mixture_bwd_kernel(const float* __restrict__ dprobs, const float* __restrict__ dmean,
                                        const float* __restrict__ dstd, const float* __restrict__ probs,
                                        const float* __restrict__ std, const T* __restrict__ alpha, long long a_sb,
                                        long long a_sk, int B, int K, int relu_alpha, T* __restrict__ dalpha,
                                        T* __restrict__ dap, long long p_sb, long long p_sk) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float dot = 0.f;
  for (int k = 0; k < K; ++k) dot += (dprobs ? dprobs[(long long)b * K + k] : 0.f) * probs[(long long)b * K + k];
  for (int k = 0; k < K; ++k) {
    const long long i = (long long)b * K + k;
    float da = dprobs ? probs[i] * (dprobs[i] - dot) : 0.f;
    if (relu_alpha && hload<T>(alpha + b * a_sb + k * a_sk) <= 0.f) da = 0.f;
    hstore<T>(dalpha + b * a_sb + k * a_sk, da);
    T* q = dap + b * p_sb + k * p_sk;
    hstore<T>(q + 0, dmean ? dmean[i * 2 + 0] : 0.f);
    hstore<T>(q + 1, dmean ? dmean[i * 2 + 1] : 0.f);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float sv = std[i * 2 + j];  // sv = r+1 (r>0) or exp(r) (r<=0, so sv<=1): d sv / d r = sv>1 ? 1 : sv
      hstore<T>(q + 2 + j, dstd ? dstd[i * 2 + j] * (sv > 1.f ? 1.f : sv) : 0.f);
    }
  }
}

// loss = c0 * NLL + c1 * speed_loss (trainer/loss.py:121-132) with d(loss)/d(probs, mean, std, speed_pred).
// log-mixture weights follow torch.distributions.Categorical(probs): p/sum(p) clamped to [eps, 1-eps], log, log_softmax.
__global__ void moe_loss_kernel(const float* __restrict__ probs, const float* __restrict__ mean, const float* __restrict__ std,
                                const float* __restrict__ speed_pred, int speed_k, const float* __restrict__ act_gt,
                                const float* __restrict__ speed_gt, int B, int K, float c0, float c1,
                                float* __restrict__ loss_out /*[3]: total, nll, speed*/, float* __restrict__ dprobs,
                                float* __restrict__ dmean, float* __restrict__ dstd, float* __restrict__ dspeed,
                                float* __restrict__ logp_out) {
  const float eps = 1.1920928955078125e-07f;
  const float half_log_2pi = 0.9189385332046727f;
  float nll_acc = 0.f, sp_acc = 0.f;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    float lw[kMaxExperts], comp[kMaxExperts], pc[kMaxExperts];
    float psum = 0.f;
    for (int k = 0; k < K; ++k) psum += probs[(long long)b * K + k];
    float lse_w_m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const float pn = probs[(long long)b * K + k] / psum;
      pc[k] = fminf(fmaxf(pn, eps), 1.f - eps);
      lw[k] = logf(pc[k]);
      lse_w_m = fmaxf(lse_w_m, lw[k]);
    }
    float lse_w = 0.f;
    for (int k = 0; k < K; ++k) lse_w += expf(lw[k] - lse_w_m);
    lse_w = lse_w_m + logf(lse_w);
    const float a0 = act_gt[b * 2 + 0], a1 = act_gt[b * 2 + 1];
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) {
      const long long i = ((long long)b * K + k) * 2;
      const float z0 = (a0 - mean[i]) / std[i], z1 = (a1 - mean[i + 1]) / std[i + 1];
      comp[k] = -0.5f * (z0 * z0 + z1 * z1) - logf(std[i]) - logf(std[i + 1]) - 2.f * half_log_2pi + (lw[k] - lse_w);
      m = fmaxf(m, comp[k]);
    }
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(comp[k] - m);
    const float logp = m + logf(se);
    if (logp_out) logp_out[b] = logp;
    nll_acc += -logp;
    // gradients: r_k = responsibility; d(-logp)/d comp_k = -r_k
    const float gscale = c0 / (float)B;
    float wsum = 0.f;  // sum_j softmax(lw)_j is 1; d lse_w / d lw_j = softmax_j
    float rk[kMaxExperts];
    for (int k = 0; k < K; ++k) rk[k] = expf(comp[k] - logp);
    // d(-logp)/d lw_k = -(r_k - softmax(lw)_k)
    float dpn[kMaxExperts];
    float dot = 0.f;
    for (int k = 0; k < K; ++k) {
      const float sm = expf(lw[k] - lse_w);
      const float dlw = -(rk[k] - sm);
      const float pn = probs[(long long)b * K + k] / psum;
      const bool inside = pn > eps && pn < 1.f - eps;  // clamp passes gradient only strictly inside
      dpn[k] = inside ? dlw / pc[k] : 0.f;
      dot += dpn[k] * pn;
      wsum += sm;
    }
    for (int k = 0; k < K; ++k) {
      const long long i = ((long long)b * K + k);
      if (dprobs) dprobs[i] = gscale * (dpn[k] - dot) / psum;  // through p/sum(p)
      const long long j = i * 2;
      const float s0 = std[j], s1 = std[j + 1];
      const float z0 = (a0 - mean[j]) / s0, z1 = (a1 - mean[j + 1]) / s1;
      if (dmean) {
        dmean[j] = gscale * (-rk[k]) * (z0 / s0);
        dmean[j + 1] = gscale * (-rk[k]) * (z1 / s1);
      }
      if (dstd) {
        dstd[j] = gscale * (-rk[k]) * ((z0 * z0 - 1.f) / s0);
        dstd[j + 1] = gscale * (-rk[k]) * ((z1 * z1 - 1.f) / s1);
      }
    }
    // speed: mean over (B*speed_k) of (pred-gt)^2, divided by speed_k again when 3-D (loss.py:126-128)
    const float denom = (float)B * (float)speed_k * (speed_k > 1 || false ? (float)speed_k : 1.f);
    for (int k = 0; k < speed_k; ++k) {
      const float d = speed_pred[(long long)b * speed_k + k] - speed_gt[b];
      sp_acc += d * d;
      if (dspeed) dspeed[(long long)b * speed_k + k] = c1 * 2.f * d / denom;
    }
  }
  // block reduce -> atomics
  __shared__ float red[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 16; o > 0; o >>= 1) {
    nll_acc += __shfl_xor_sync(0xffffffffu, nll_acc, o);
    sp_acc += __shfl_xor_sync(0xffffffffu, sp_acc, o);
  }
  if (lane == 0) {
    red[0][warp] = nll_acc;
    red[1][warp] = sp_acc;
  }
  __syncthreads();
  if (warp == 0) {
    float a = lane < (blockDim.x >> 5) ? red[0][lane] : 0.f, s = lane < (blockDim.x >> 5) ? red[1][lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) {
      const float nll = a / (float)B;
      const float denom = (float)B * (float)speed_k * (speed_k > 1 ? (float)speed_k : 1.f);
      const float sp = s / denom;
      atomicAdd(loss_out + 1, nll);
      atomicAdd(loss_out + 2, sp);
      atomicAdd(loss_out + 0, c0 * nll + c1 * sp);
    }
  }
}

// stateless dropout: keep = hash(seed, index) >= p * 2^32; y = x * keep / (1-p). The same call with the
// incoming gradient as x is the backward.
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return (uint32_t)x;
}
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n, float p, unsigned long long seed,
                               const unsigned long long* __restrict__ seed_dev) {
  if (seed_dev) seed += *seed_dev * 0xD1B54A32D192ED03ULL;  // step counter in device memory: fresh masks under CUDA-graph replay
  const uint32_t thresh = (uint32_t)fminf(p * 4294967296.f, 4294967295.f);
  const float scale = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const bool keep = mix32(seed * 0x9E3779B97F4A7C15ULL + (unsigned long long)i) >= thresh;
    hstore<T>(y + i, keep ? hload<T>(x + i) * scale : 0.f);
  }
}

// mean |a-b| (l1) or (a-b)^2 (mse) with gradient w.r.t. a scaled by `coef`
__global__ void l1_mse_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int is_mse, float coef,
                              float* __restrict__ loss, float* __restrict__ da) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += is_mse ? d * d : fabsf(d);
    if (da) da[i] = coef * (is_mse ? 2.f * d : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f))) / (float)n;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, coef * acc / (float)n);
}

}  // namespace pmoe

using namespace pmoe;

extern "C" {

int pmoe_gate_mixture_fwd(const void* alpha, int64_t a_sb, int64_t a_sk, const void* ap, int64_t p_sb, int64_t p_sk,
                          int32_t dtype, int32_t B, int32_t K, int32_t relu_alpha, float* probs, float* mean, float* std,
                          int64_t* route, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!alpha || !ap || !probs || !mean || !std || B < 1 || K < 1 || K > kMaxExperts) {
    set_error("gate_mixture_fwd: bad arguments (K must be 1..%d)", kMaxExperts);
    return PMOE_ERR_ARG;
  }
  const int th = 128, bl = (B + th - 1) / th;
  if (dtype == PMOE_BF16)
    gate_mixture_fwd_kernel<__nv_bfloat16><<<bl, th, 0, stream>>>(static_cast<const __nv_bfloat16*>(alpha), a_sb, a_sk,
                                                                  static_cast<const __nv_bfloat16*>(ap), p_sb, p_sk, B, K,
                                                                  relu_alpha, probs, mean, std, (long long*)route);
  else
    gate_mixture_fwd_kernel<float><<<bl, th, 0, stream>>>(static_cast<const float*>(alpha), a_sb, a_sk,
                                                          static_cast<const float*>(ap), p_sb, p_sk, B, K, relu_alpha, probs,
                                                          mean, std, (long long*)route);
  return check_launch("gate_mixture_fwd");
}

int pmoe_gate_mixture_bwd(const float* dprobs, const float* dmean, const float* dstd, const float* probs, const float* std,
                          const void* alpha, int64_t a_sb, int64_t a_sk, int32_t dtype, int32_t B, int32_t K,
                          int32_t relu_alpha, void* dalpha, void* dap, int64_t p_sb, int64_t p_sk, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!probs || !std || !alpha || !dalpha || !dap || B < 1 || K < 1 || K > kMaxExperts) {
    set_error("gate_mixture_bwd: bad arguments");
    return PMOE_ERR_ARG;
  }
  const int th = 128, bl = (B + th - 1) / th;
  if (dtype == PMOE_BF16)
    gate_mixture_bwd_kernel<__nv_bfloat16><<<bl, th, 0, stream>>>(dprobs, dmean, dstd, probs, std,
                                                                  static_cast<const __nv_bfloat16*>(alpha), a_sb, a_sk, B, K,
                                                                  relu_alpha, static_cast<__nv_bfloat16*>(dalpha),
                                                                  static_cast<__nv_bfloat16*>(dap), p_sb, p_sk);
  else
    gate_mixture_bwd_kernel<float><<<bl, th, 0, stream>>>(dprobs, dmean, dstd, probs, std, static_cast<const float*>(alpha),
                                                          a_sb, a_sk, B, K, relu_alpha, static_cast<float*>(dalpha),
                                                          static_cast<float*>(dap), p_sb, p_sk);
  return check_launch("gate_mixture_bwd");
}

int pmoe_moe_loss(const float* probs, const float* mean, const float* std, const float* speed_pred, int32_t speed_k,
                  const float* act_gt, const float* speed_gt, int32_t B, int32_t K, float c0, float c1, float* loss_out,
                  float* dprobs, float* dmean, float* dstd, float* dspeed, float* logp_out, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!probs || !mean || !std || !speed_pred || !act_gt || !speed_gt || !loss_out || B < 1 || K < 1 || K > kMaxExperts ||
      speed_k < 1) {
    set_error("moe_loss: bad arguments");
    return PMOE_ERR_ARG;
  }
  const int th = 128;
  int bl = (B + th - 1) / th;
  if (bl > num_sms() * 4) bl = num_sms() * 4;
  moe_loss_kernel<<<bl, th, 0, stream>>>(probs, mean, std, speed_pred, speed_k, act_gt, speed_gt, B, K, c0, c1, loss_out,
                                         dprobs, dmean, dstd, dspeed, logp_out);
  return check_launch("moe_loss");
}

static int dropout_launch(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                          pmoe_stream_t stream_);

int pmoe_dropout(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t seed, pmoe_stream_t stream_) {
  return dropout_launch(x, y, dtype, n, p, seed, nullptr, stream_);
}

int pmoe_dropout_dev(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t salt, const uint64_t* seed_dev,
                     pmoe_stream_t stream_) {
  if (!seed_dev) {
    set_error("dropout_dev: device seed missing");
    return PMOE_ERR_ARG;
  }
  return dropout_launch(x, y, dtype, n, p, salt, seed_dev, stream_);
}

static int dropout_launch(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev,
                          pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!x || !y || n < 1 || p < 0.f || p >= 1.f) {
    set_error("dropout: bad arguments");
    return PMOE_ERR_ARG;
  }
  long long bl = (n + 255) / 256;
  if (bl > (long long)num_sms() * 16) bl = (long long)num_sms() * 16;
  if (dtype == PMOE_BF16)
    dropout_kernel<__nv_bfloat16><<<(unsigned)bl, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x),
                                                                    static_cast<__nv_bfloat16*>(y), n, p, seed,
                                                                    reinterpret_cast<const unsigned long long*>(seed_dev));
  else
    dropout_kernel<float><<<(unsigned)bl, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(y), n, p, seed,
                                                            reinterpret_cast<const unsigned long long*>(seed_dev));
  return check_launch("dropout");
}

int pmoe_l1_mse(const float* a, const float* b, int64_t n, int32_t is_mse, float coef, float* loss, float* da,
                pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!a || !b || !loss || n < 1) {
    set_error("l1_mse: bad arguments");
    return PMOE_ERR_ARG;
  }
  long long bl = (n + 255) / 256;
  if (bl > (long long)num_sms() * 4) bl = (long long)num_sms() * 4;
  l1_mse_kernel<<<(unsigned)bl, 256, 0, stream>>>(a, b, n, is_mse, coef, loss, da);
  return check_launch("l1_mse");
}

}  // extern "C"
