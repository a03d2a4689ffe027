// Tensor-core weight gradient of a 3x3 / stride-1 / pad-1 convolution (bf16 activations, fp32 accumulation in TMEM).
//
//   dW[co][tap][ci] = sum over pixels p of dy[p][co] * x[p + tap][ci]
//
// GEMM view per CTA: D[M = (tap pair) x 64 input channels][N = output channels] += A^T B over K = pixels, where
//   A = the (16+2) x (8+2) input HALO of a 16x8 pixel patch for one 64-channel chunk, brought by ONE TMA; the two taps
//       of a pair are the two 64-element MN-blocks of an MN-major UMMA descriptor (leading byte offset = the row shift
//       between the taps), and each K=16 step is two 8-pixel row groups of the halo (stride byte offset = 10 rows);
//   B = the dy patch (MN-major, N = 64 / 128 / 256 output channels as 64-channel TMA boxes).
// So the nine taps re-use one staged halo tile, like the forward halo kernel, and x is never re-read from L2 per tap.
// The pixel reduction is split across CTAs; each CTA keeps its accumulators in TMEM for its whole pixel range and
// finishes with fp32 atomics into the packed weight-gradient buffer [cout_pad][ktot] (same K order as the forward).
#include "host_util.h"
#include "ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace pmoe {

constexpr int kWgThreads = 192;
constexpr int kWgHaloBytes = 23 * 1024;  // 180 rows x 128 B, padded
constexpr int kWgHaloTx = 180 * 128;
constexpr int kWgDyBlock = 128 * 128;    // 128 pixels x 64 channels bf16
constexpr int kWgMaxStages = 6;

struct alignas(64) WgradParams {
  CUtensorMap tm_x[PMOE_MAX_SRC];
  CUtensorMap tm_dy;
  int8_t chunk_src[PMOE_MAX_SEG];
  uint16_t chunk_c0[PMOE_MAX_SEG];
  int n_chunks;        // 64-channel chunks over all sources
  int N;               // output channels per CTA (64, 128 or 256)
  int n_tiles;         // cout_pad / N
  int n_groups;        // tap-pair groups
  int group_first[4];  // first pair of each group
  int group_pairs[4];  // pairs in each group
  int tiles_w, tiles_h, n_img;
  long long m_tiles;
  int splits;
  int stages;
  float* dw;
  int ktot, cout_pad;
};

// MN-major operand descriptor (SWIZZLE_128B): 64-element MN blocks `lbo_bytes` apart, 8-row K groups `sbo_bytes` apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1u) << 46;
  d |= static_cast<uint64_t>(kLayoutSW128) << 61;
  return d;
}

__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = kWgHaloBytes + (p.N / 64) * kWgDyBlock;
  uint8_t* pipe = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(pipe + (size_t)p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + kWgMaxStages;
  uint64_t* done_bar = empty_bar + kWgMaxStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_dy);
    tma_prefetch_desc(&p.tm_x[0]);
  }
  if (warp == 2) {
    tmem_alloc(s_tmem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // decode the work unit of this CTA
  int u = blockIdx.x;
  const int split = u % p.splits;
  u /= p.splits;
  const int grp = u % p.n_groups;
  u /= p.n_groups;
  const int g = u % p.n_chunks;
  const int nt = u / p.n_chunks;
  const long long t_begin = (p.m_tiles * split) / p.splits;
  const long long t_end = (p.m_tiles * (split + 1)) / p.splits;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int npairs = p.group_pairs[grp];
  const int pair0 = p.group_first[grp];
  const int src = p.chunk_src[g];
  const int c0 = p.chunk_c0[g];

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = t_begin; t < t_end; ++t) {
        const int img = (int)(t / tiles_per_img);
        const int rem = (int)(t % tiles_per_img);
        const int h0 = (rem / p.tiles_w) * 16, w0 = (rem % p.tiles_w) * 8;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* st = pipe + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(kWgHaloTx + (p.N / 64) * kWgDyBlock));
        tma_load_4d(st, &p.tm_x[src], &full_bar[stage], c0, w0 - 1, h0 - 1, img);
        for (int b = 0; b < p.N / 64; ++b)
          tma_load_4d(st + kWgHaloBytes + b * kWgDyBlock, &p.tm_dy, &full_bar[stage], nt * p.N + b * 64, w0, h0, img);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.N, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (long long t = t_begin; t < t_end; ++t) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t halo = smem_u32(pipe + (size_t)stage * stage_bytes);
        const uint32_t dyb = halo + kWgHaloBytes;
        for (int pi = 0; pi < npairs; ++pi) {
          const int t0 = 2 * (pair0 + pi);
          const int t1 = t0 + 1 > 8 ? 8 : t0 + 1;
          const int row0 = (t0 / 3) * 10 + (t0 % 3);
          const int row1 = (t1 / 3) * 10 + (t1 % 3);
          const uint32_t lbo = (uint32_t)(row1 - row0) * 128u;
          const uint32_t d_tmem = tmem_base + (uint32_t)(pi * p.N);
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // K = 16 pixels = patch rows 2j, 2j+1
            const uint64_t adesc = umma_desc_mnmajor(halo + (uint32_t)(row0 + 20 * j) * 128u, lbo, 1280u);
            const uint64_t bdesc = umma_desc_mnmajor(dyb + (uint32_t)j * 2048u, (uint32_t)kWgDyBlock, 1024u);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (first && j == 0) ? 0u : 1u);
          }
        }
        first = false;
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: accumulator row m = (tap of the pair) * 64 + input channel; column = output channel
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int m = quad * 32 + lane;
    const int half = m >> 6, ch = m & 63;
    if (t_end > t_begin) {
      for (int pi = 0; pi < npairs; ++pi) {
        const int t0 = 2 * (pair0 + pi);
        const int tap = t0 + half;
        const bool live = tap <= 8;  // the second half of the last pair (tap 9) duplicates tap 8: dropped
        float* dst = p.dw + (long long)((tap > 8 ? 8 : tap) * p.n_chunks + g) * 64 + ch;
        for (int cb = 0; cb < p.N; cb += 32) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(pi * p.N + cb), raw);
          tmem_ld_wait();
          if (live) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int co = nt * p.N + cb + k;
              if (co < p.cout_pad) atomicAdd(dst + (long long)co * p.ktot, __uint_as_float(raw[k]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int wg_view_tmap(CUtensorMap* tm, const PmoeView4& v, int bw, int bh, const char* what) {
  if (((uintptr_t)v.ptr & 15) || (v.sw % 8) || (v.sh % 8) || (v.sn % 8) || (v.c % 8)) {
    set_error("%s: view must be 16-byte aligned with strides/channels in multiples of 8 elements", what);
    return PMOE_ERR_ARG;
  }
  const uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  const uint64_t strides[3] = {(uint64_t)v.sw * 2, (uint64_t)v.sh * 2, (uint64_t)v.sn * 2};
  const uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace pmoe

using namespace pmoe;

// Returns PMOE_ERR_UNSUPPORTED (without setting a launch) when the descriptor is not a canonical 3x3/s1/p1 conv over
// whole 64-channel-chunked sources; the caller then uses pmoe_conv_wgrad_simt.
extern "C" int pmoe_conv_wgrad_tc(const PmoeConvTc* d, float* dwpack, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || !dwpack || d->n_src < 1 || d->n_src > PMOE_MAX_SRC || !d->out.ptr) {
    set_error("conv_wgrad_tc: bad descriptor");
    return PMOE_ERR_ARG;
  }
  bool ok = d->ck == 64 && d->n_seg == 9 * d->n_src && d->out.h >= 18 && d->out.w >= 10 && d->cout_pad % 64 == 0;
  for (int t = 0; ok && t < 9; ++t)
    for (int i = 0; ok && i < d->n_src; ++i) {
      const PmoeSeg& s = d->seg[t * d->n_src + i];
      ok = s.src == i && s.dh == t / 3 - 1 && s.dw == t % 3 - 1 && s.c0 == 0 && s.nchunks * 64 == d->src[i].c;
    }
  int total_chunks = 0;
  for (int i = 0; i < d->n_src; ++i) total_chunks += d->src[i].c / 64;
  if (!ok || total_chunks > PMOE_MAX_SEG || 9 * total_chunks * 64 != d->ktot) {
    set_error("conv_wgrad_tc: not a canonical 3x3/s1/p1 convolution over 64-channel chunks");
    return PMOE_ERR_UNSUPPORTED;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int g = 0;
  for (int i = 0; i < d->n_src; ++i)
    for (int c = 0; c < d->src[i].c / 64; ++c, ++g) {
      p.chunk_src[g] = (int8_t)i;
      p.chunk_c0[g] = (uint16_t)(c * 64);
    }
  p.n_chunks = total_chunks;
  p.N = d->cout_pad % 256 == 0 ? 256 : (d->cout_pad % 128 == 0 ? 128 : 64);
  p.n_tiles = d->cout_pad / p.N;
  const int max_pairs = 512 / p.N;  // TMEM columns
  int first = 0;
  while (first < 5) {
    int n = 5 - first < max_pairs ? 5 - first : max_pairs;
    if (p.N == 128 && first == 0) n = 3;  // 3 + 2 rather than 4 + 1
    p.group_first[p.n_groups] = first;
    p.group_pairs[p.n_groups] = n;
    ++p.n_groups;
    first += n;
  }
  p.tiles_w = (d->out.w + 7) / 8;
  p.tiles_h = (d->out.h + 15) / 16;
  p.n_img = d->out.n;
  p.m_tiles = (long long)p.tiles_w * p.tiles_h * p.n_img;
  const long long units = (long long)p.n_tiles * p.n_chunks * p.n_groups;
  long long splits = (long long)num_sms() / units;  // one resident CTA per SM (TMEM + smem): a single wave
  if (splits > p.m_tiles) splits = p.m_tiles;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  const int stage_bytes = kWgHaloBytes + (p.N / 64) * kWgDyBlock;
  int stages = (220 * 1024) / stage_bytes;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  p.stages = stages;
  const int smem_bytes = 1024 + stages * stage_bytes + (2 * kWgMaxStages + 1) * 8 + 16;
  p.dw = dwpack;
  p.ktot = d->ktot;
  p.cout_pad = d->cout_pad;
  int rc;
  for (int i = 0; i < d->n_src; ++i)
    if ((rc = wg_view_tmap(&p.tm_x[i], d->src[i], 10, 18, "conv_wgrad_tc source")) != PMOE_OK) return rc;
  if ((rc = wg_view_tmap(&p.tm_dy, d->out, 8, 16, "conv_wgrad_tc dy")) != PMOE_OK) return rc;
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      set_error("conv_wgrad_tc: cannot reserve %d bytes of shared memory: %s", smem_bytes, cudaGetErrorString(e));
      return PMOE_ERR_LAUNCH;
    }
    configured = smem_bytes;
  }
  const long long grid = units * p.splits;
  if (grid > 0x7fffffffLL) {
    set_error("conv_wgrad_tc: grid too large");
    return PMOE_ERR_ARG;
  }
  conv_wgrad_tc_kernel<<<(unsigned)grid, kWgThreads, smem_bytes, stream>>>(p);
  return check_launch("conv_wgrad_tc");
}
