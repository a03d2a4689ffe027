// Weight layout kernels: fp32 (out, in, kh, kw) nn.Parameter <-> the packed [cout_pad][K] operand layout of the conv /
// linear kernels, driven by a static int32 index map (packed element i comes from parameter element idx[i]; -1 = zero
// padding). The maps are built once per (weight, layer geometry) on the host; the kernels run on every optimizer step
// (and inside captured CUDA graphs), so the packed operands always follow the live parameters — reference: the ATen
// convolution reads `conv.weight` directly on every call (model/blocks/basics.py:51,54), there is no derived copy there.
// HBM-bound: 4 B index + 4 B gathered read + 2 B (bf16) / 4 B write per packed element.
#include <cuda_bf16.h>

#include "host_util.h"

namespace pmoe {

struct PackJob {  // must match PmoePackJob
  const float* w;
  const int32_t* idx;
  void* out;
  int64_t n;
  int32_t dtype;
  int32_t chunk0;  // first chunk of this job in the launch's flat chunk numbering
};

constexpr int PACK_CHUNK = 8192;  // packed elements per CTA-iteration (256 threads x 4 groups of 8)

template <typename OutT>
__device__ __forceinline__ void gather_range(const float* __restrict__ w, const int32_t* __restrict__ idx, OutT* __restrict__ out,
                                             int64_t begin, int64_t end) {
  // groups of 8 consecutive packed elements per thread: two 16-byte index loads, one 16-byte (bf16) / two (fp32) stores
  const bool vec = ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && (begin % 8 == 0);
  if (vec) {
    const int64_t g_end = begin + ((end - begin) / 8) * 8;
    for (int64_t i = begin + (int64_t)threadIdx.x * 8; i < g_end; i += (int64_t)blockDim.x * 8) {
      const int4 a = __ldg(reinterpret_cast<const int4*>(idx + i));
      const int4 b = __ldg(reinterpret_cast<const int4*>(idx + i + 4));
      const int32_t id[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = id[j] >= 0 ? __ldg(w + id[j]) : 0.f;
      if constexpr (sizeof(OutT) == 2) {
        __nv_bfloat162 p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(p);
      } else {
        *reinterpret_cast<float4*>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(out + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    begin = g_end;
  }
  for (int64_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    const int32_t id = __ldg(idx + i);
    const float v = id >= 0 ? __ldg(w + id) : 0.f;
    if constexpr (sizeof(OutT) == 2) out[i] = __float2bfloat16_rn(v);
    else out[i] = v;
  }
}

__global__ void __launch_bounds__(256) pack_gather_kernel(const float* __restrict__ w, const int32_t* __restrict__ idx, void* out, int64_t n,
                                                          int dtype) {
  for (int64_t c = blockIdx.x; c * PACK_CHUNK < n; c += gridDim.x) {
    const int64_t b = c * PACK_CHUNK, e = (b + PACK_CHUNK < n) ? b + PACK_CHUNK : n;
    if (dtype == PMOE_BF16) gather_range(w, idx, static_cast<__nv_bfloat16*>(out), b, e);
    else gather_range(w, idx, static_cast<float*>(out), b, e);
  }
}

// Multi-tensor form: one launch refreshes every packed operand of a model. `jobs` is sorted by chunk0; a CTA finds the
// job of its chunk by binary search.
__global__ void __launch_bounds__(256) pack_gather_mt_kernel(const PackJob* __restrict__ jobs, int n_jobs, int total_chunks) {
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].chunk0 <= c) lo = mid;
      else hi = mid - 1;
    }
    const PackJob j = jobs[lo];
    const int64_t b = (int64_t)(c - j.chunk0) * PACK_CHUNK, e = (b + PACK_CHUNK < j.n) ? b + PACK_CHUNK : j.n;
    if (j.dtype == PMOE_BF16) gather_range(j.w, j.idx, static_cast<__nv_bfloat16*>(j.out), b, e);
    else gather_range(j.w, j.idx, static_cast<float*>(j.out), b, e);
  }
}

// dst[idx[i]] (+)= alpha * packed[i] for idx[i] >= 0. Every parameter element is the image of exactly one packed element
// (the forward packing is a bijection onto the non-padding entries), so plain stores suffice and dst needs no zero-fill.
__global__ void __launch_bounds__(256) unpack_scatter_kernel(const float* __restrict__ packed, const int32_t* __restrict__ idx,
                                                             float* __restrict__ dst, int64_t n, float alpha, int accumulate) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  const bool vec = ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(packed) & 15) == 0);
  if (vec) {
    const int64_t n4 = n / 4 * 4;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n4; i += stride) {
      const int4 a = __ldg(reinterpret_cast<const int4*>(idx + i));
      const float4 p = __ldg(reinterpret_cast<const float4*>(packed + i));
      const int32_t id[4] = {a.x, a.y, a.z, a.w};
      const float v[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (id[j] >= 0) dst[id[j]] = accumulate ? dst[id[j]] + alpha * v[j] : alpha * v[j];
    }
    for (int64_t i = n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int32_t id = idx[i];
      if (id >= 0) dst[id] = accumulate ? dst[id] + alpha * packed[i] : alpha * packed[i];
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const int32_t id = idx[i];
      if (id >= 0) dst[id] = accumulate ? dst[id] + alpha * packed[i] : alpha * packed[i];
    }
  }
}

// The same for a GROUP of up to 16 equally shaped parameters in one launch: packed is [K][n] (the K experts' stacked gradients of
// one layer), group g scatters into dst[g]. grid.y = group.
struct ScatterGroup {
  float* dst[16];
  int accumulate[16];
};
__global__ void __launch_bounds__(256) unpack_scatter_group_kernel(const float* __restrict__ packed, const int32_t* __restrict__ idx,
                                                                   ScatterGroup grp, int64_t n, float alpha) {
  const int g = blockIdx.y;
  float* __restrict__ dst = grp.dst[g];
  if (dst == nullptr) return;
  const int acc = grp.accumulate[g];
  const float* __restrict__ src = packed + (int64_t)g * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t id = __ldg(idx + i);
    if (id >= 0) dst[id] = acc ? dst[id] + alpha * src[i] : alpha * src[i];
  }
}

// The same result through the INVERSE map (parameter element j <- packed element inv[j]): the writes into the gradient slot are
// coalesced and only the fp32 reads are gathered (the scatter form writes 4-byte words nine floats apart for a 3x3 weight: every
// store a partial sector). grid.y = group (1 for a single parameter).
__global__ void __launch_bounds__(256) unpack_gather_group_kernel(const float* __restrict__ packed, const int32_t* __restrict__ inv,
                                                                  ScatterGroup grp, int64_t n_param, int64_t n_packed, float alpha) {
  const int g = blockIdx.y;
  float* __restrict__ dst = grp.dst[g];
  if (dst == nullptr) return;
  const int acc = grp.accumulate[g];
  const float* __restrict__ src = packed + (int64_t)g * n_packed;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_param; j += (int64_t)gridDim.x * blockDim.x) {
    const float v = alpha * __ldg(src + __ldg(inv + j));
    dst[j] = acc ? dst[j] + v : v;
  }
}

// dst[i] (+)= (float) src[i]: fp64 per-channel reductions (BatchNorm / bias gradients) into fp32 gradient slots.
__global__ void __launch_bounds__(256) cvt_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, int n, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = accumulate ? dst[i] + (float)src[i] : (float)src[i];
}

static int grid_for(int64_t chunks) {
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(chunks < cap ? (chunks > 0 ? chunks : 1) : cap);
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_pack_gather(const float* w, const int32_t* idx, void* out, int32_t out_dtype, int64_t n, pmoe_stream_t stream_) {
  if (n <= 0) return PMOE_OK;
  if (!w || !idx || !out || (out_dtype != PMOE_F32 && out_dtype != PMOE_BF16)) {
    set_error("pack_gather: null pointer or bad dtype");
    return PMOE_ERR_ARG;
  }
  const int64_t chunks = (n + PACK_CHUNK - 1) / PACK_CHUNK;
  pack_gather_kernel<<<grid_for(chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(w, idx, out, n, out_dtype);
  return check_launch("pack_gather");
}

extern "C" int pmoe_pack_gather_mt(const PmoePackJob* jobs_dev, int32_t n_jobs, int32_t total_chunks, pmoe_stream_t stream_) {
  if (n_jobs <= 0 || total_chunks <= 0) return PMOE_OK;
  if (!jobs_dev) {
    set_error("pack_gather_mt: null job table");
    return PMOE_ERR_ARG;
  }
  pack_gather_mt_kernel<<<grid_for(total_chunks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const PackJob*>(jobs_dev),
                                                                                               n_jobs, total_chunks);
  return check_launch("pack_gather_mt");
}

extern "C" int pmoe_pack_chunk_elems(void) { return PACK_CHUNK; }

extern "C" int pmoe_unpack_scatter(const float* packed, const int32_t* idx, float* dst, int64_t n, float alpha, int32_t accumulate,
                                   pmoe_stream_t stream_) {
  if (n <= 0) return PMOE_OK;
  if (!packed || !idx || !dst) {
    set_error("unpack_scatter: null pointer");
    return PMOE_ERR_ARG;
  }
  const int64_t blocks = (n + 1023) / 1024;
  unpack_scatter_kernel<<<grid_for(blocks), 256, 0, static_cast<cudaStream_t>(stream_)>>>(packed, idx, dst, n, alpha, accumulate);
  return check_launch("unpack_scatter");
}

extern "C" int pmoe_cvt_f64_f32(const double* src, float* dst, int32_t n, int32_t accumulate, pmoe_stream_t stream_) {
  if (n <= 0) return PMOE_OK;
  if (!src || !dst) {
    set_error("cvt_f64_f32: null pointer");
    return PMOE_ERR_ARG;
  }
  cvt_f64_f32_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, dst, n, accumulate);
  return check_launch("cvt_f64_f32");
}

extern "C" int pmoe_unpack_scatter_group(const float* packed, const int32_t* idx, float* const* dst, const int32_t* accumulate,
                                         int32_t n_groups, int64_t n, float alpha, pmoe_stream_t stream_) {
  if (n <= 0 || n_groups <= 0) return PMOE_OK;
  if (!packed || !idx || !dst || !accumulate || n_groups > 16) {
    set_error("unpack_scatter_group: null pointer or more than 16 groups");
    return PMOE_ERR_ARG;
  }
  ScatterGroup grp;
  for (int g = 0; g < 16; ++g) {
    grp.dst[g] = g < n_groups ? dst[g] : nullptr;      // host arrays: copied into the launch parameters
    grp.accumulate[g] = g < n_groups ? accumulate[g] : 0;
  }
  const int64_t blocks = (n + 1023) / 1024;
  int gx = grid_for(blocks);
  if (gx > 64) gx = 64;
  dim3 grid((unsigned)gx, (unsigned)n_groups);
  unpack_scatter_group_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(packed, idx, grp, n, alpha);
  return check_launch("unpack_scatter_group");
}

extern "C" int pmoe_unpack_gather_group(const float* packed, const int32_t* inv, float* const* dst, const int32_t* accumulate,
                                        int32_t n_groups, int64_t n_param, int64_t n_packed, float alpha, pmoe_stream_t stream_) {
  if (n_param <= 0 || n_groups <= 0) return PMOE_OK;
  if (!packed || !inv || !dst || !accumulate || n_groups > 16) {
    set_error("unpack_gather_group: null pointer or more than 16 groups");
    return PMOE_ERR_ARG;
  }
  ScatterGroup grp;
  for (int g = 0; g < 16; ++g) {
    grp.dst[g] = g < n_groups ? dst[g] : nullptr;
    grp.accumulate[g] = g < n_groups ? accumulate[g] : 0;
  }
  const int64_t blocks = (n_param + 1023) / 1024;
  int gx = grid_for(blocks);
  if (n_groups > 1 && gx > 128) gx = 128;
  dim3 grid((unsigned)gx, (unsigned)n_groups);
  unpack_gather_group_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(packed, inv, grp, n_param, n_packed, alpha);
  return check_launch("unpack_gather_group");
}
