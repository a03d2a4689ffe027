// Backward of "Linear -> activation" stacks without BatchNorm (the experts' head MLPs, reference PMoE/model/blocks/basics.py:11-45 as
// used at model/moe.py:88-101): dy = dz * act'(z) from the SAVED output z, fused with the bias gradient, the per-image channel sum of
// dy (image = expert in the grouped heads). The separate launches (pmoe_bn_bwd_apply without statistics + pmoe_channel_sums) read dy
// a second time: at 65536 rows x 512 channels x K experts that second read is 6 % of the whole forward + backward step.
// Dense bf16 NHWC, at most 2048 channels, channel-group count dividing 256.
#include <cuda_bf16.h>

#include "host_util.h"
#include "reduce.cuh"

namespace pmoe {

__device__ __forceinline__ void ab_unpack(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[2 * q] = __uint_as_float(w[q] << 16);
    v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t ab_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ACT: 0 none, 1 ReLU, 2 ELU (alpha = 1), all from the saved output: relu' = [z > 0], elu' = z > 0 ? 1 : z + 1
template <int ACT>
__global__ void __launch_bounds__(kRedThreads) act_bwd_bias_kernel(const uint4* __restrict__ dz, const uint4* __restrict__ z, uint4* __restrict__ dy,
                                                                   long long pix_per_img, int cg, long long pix_per_block,
                                                                   float* __restrict__ bias_sum, long long bias_stride) {
  __shared__ float sm[kRedThreads * 8];
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const size_t img = (size_t)n * pix_per_img * cg;
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > pix_per_img) p1 = pix_per_img;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long p = p0 + lane; p < p1; p += 4LL * lanes) {   // four 16-byte loads per tensor in flight
    uint4 rd[4], rz[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long pp = p + (long long)u * lanes;
      ok[u] = pp < p1;
      const size_t off = img + (size_t)(ok[u] ? pp : p) * cg + g;
      rd[u] = __ldg(dz + off);
      if (ACT != 0) rz[u] = __ldg(z + off);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      float d[8];
      ab_unpack(rd[u], d);
      if (ACT != 0) {
        float zv[8];
        ab_unpack(rz[u], zv);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (ACT == 1) d[q] = zv[q] > 0.f ? d[q] : 0.f;
          else d[q] = zv[q] > 0.f ? d[q] : d[q] * (zv[q] + 1.f);
        }
      }
      dy[img + (size_t)(p + (long long)u * lanes) * cg + g] =
          make_uint4(ab_pack2(d[0], d[1]), ab_pack2(d[2], d[3]), ab_pack2(d[4], d[5]), ab_pack2(d[6], d[7]));
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] += d[q];
    }
  }
  float ta[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8 && ta[j] != 0.f) atomicAdd(bias_sum + n * bias_stride + c, ta[j]);
  }
}

static bool ab_dense(const PmoeView4* v) {
  return v && v->ptr && v->c % 8 == 0 && ((uintptr_t)v->ptr % 16) == 0 && v->sw == v->c && v->sh == (int64_t)v->w * v->c &&
         v->sn == (int64_t)v->h * v->w * v->c;
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_act_bwd_bias(const PmoeView4* dz, const PmoeView4* z, int32_t act, const PmoeView4* dy, float* bias_sum,
                                 int64_t bias_stride, pmoe_stream_t stream_) {
  const bool need_z = act != PMOE_ACT_NONE;
  if (!ab_dense(dz) || !ab_dense(dy) || (need_z && !ab_dense(z)) || !bias_sum || bias_stride < dz->c || dy->n != dz->n || dy->h != dz->h ||
      dy->w != dz->w || dy->c != dz->c || (need_z && (z->n != dz->n || z->h != dz->h || z->w != dz->w || z->c != dz->c))) {
    set_error("act_bwd_bias: dense bf16 NHWC tensors of one shape and an (n, >= c) fp32 sum buffer are required");
    return PMOE_ERR_UNSUPPORTED;
  }
  const int cg = dz->c / 8;
  if (cg > 256 || 256 % cg != 0 || dz->n > 65535 || (act != PMOE_ACT_NONE && act != PMOE_ACT_RELU && act != PMOE_ACT_ELU)) {
    set_error("act_bwd_bias: channel-group count dividing 256, at most 65535 images, activation none / ReLU / ELU");
    return PMOE_ERR_UNSUPPORTED;
  }
  const long long hw = (long long)dz->h * dz->w;
  long long want = ((long long)num_sms() * 8 + dz->n - 1) / dz->n;
  long long ppb = (hw + want - 1) / want;
  if (ppb < 64) ppb = 64;
  dim3 grid((unsigned)((hw + ppb - 1) / ppb), (unsigned)dz->n);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const uint4* pd = static_cast<const uint4*>(dz->ptr);
  const uint4* pz = need_z ? static_cast<const uint4*>(z->ptr) : nullptr;
  uint4* py = static_cast<uint4*>(dy->ptr);
  if (act == PMOE_ACT_NONE) act_bwd_bias_kernel<0><<<grid, kRedThreads, 0, stream>>>(pd, pz, py, hw, cg, ppb, bias_sum, bias_stride);
  else if (act == PMOE_ACT_RELU) act_bwd_bias_kernel<1><<<grid, kRedThreads, 0, stream>>>(pd, pz, py, hw, cg, ppb, bias_sum, bias_stride);
  else act_bwd_bias_kernel<2><<<grid, kRedThreads, 0, stream>>>(pd, pz, py, hw, cg, ppb, bias_sum, bias_stride);
  return check_launch("act_bwd_bias");
}
