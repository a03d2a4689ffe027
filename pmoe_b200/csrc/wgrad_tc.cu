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

  // warp index through a shuffle: provably warp-uniform, so the role branches below stay convergent for ptxas
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
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
  // one CTA per SM, one allocation of all 512 columns: the MMA issuer relies on base 0. The condition is warp-uniform on
  // purpose: a thread-dependent trap is a possible partial exit after which ptxas cannot prove the MMA warp converged.
  if (__any_sync(0xffffffffu, tmem_base != 0u)) {
    if (threadIdx.x == 0) printf("pmoe conv_wgrad_tc: unexpected TMEM base %u\n", tmem_base);
    __trap();
  }

  // decode the work unit of this CTA
  int u = blockIdx.x;
  const int split = u % p.splits;
  u /= p.splits;
  const int grp = u % p.n_groups;
  u /= p.n_groups;
  const int g = u % p.n_chunks;
  const int nt = u / p.n_chunks;
  // 32-bit tile arithmetic below: a 64-bit division is a subroutine call, and a call inside the lane-0 producer branch
  // makes ptxas give up on proving the MMA warp converged
  const int t_begin = (int)((p.m_tiles * split) / p.splits);
  const int t_end = (int)((p.m_tiles * (split + 1)) / p.splits);
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int npairs = p.group_pairs[grp];
  const int pair0 = p.group_first[grp];
  const int src = p.chunk_src[g];
  const int c0 = p.chunk_c0[g];

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int img = t / tiles_per_img;
        const int rem = t % tiles_per_img;
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
    {  // converged warp: every lane runs the loop, the elected lane issues (see umma_bf16_elect)
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.N, 1, 1);
      constexpr uint32_t kHiA = umma_desc_hi(1280u, kLayoutSW128);  // 8-pixel K groups of a halo view: 10 rows apart
      constexpr uint32_t kHiB = umma_desc_hi(1024u, kLayoutSW128);
      int stage = 0;
      uint32_t phase = 0;
      bool first = true;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait_warp(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t halo = smem_u32(pipe + (size_t)stage * stage_bytes);
        const uint32_t dyb = halo + kWgHaloBytes;
        for (int pi = 0; pi < npairs; ++pi) {
          const int t0 = 2 * (pair0 + pi);
          const int t1 = t0 + 1 > 8 ? 8 : t0 + 1;
          const int row0 = (t0 / 3) * 10 + (t0 % 3);
          const int row1 = (t1 / 3) * 10 + (t1 % 3);
          const uint32_t lbo = (uint32_t)(row1 - row0) * 128u;
          const uint32_t d_tmem = (uint32_t)(pi * p.N);  // TMEM base is 0 (checked after the allocation)
          const uint32_t a_lo = umma_desc_lo(halo + (uint32_t)row0 * 128u, lbo);
          const uint32_t b_lo = umma_desc_lo(dyb, (uint32_t)kWgDyBlock);
#pragma unroll
          for (int j = 0; j < 8; ++j)  // K = 16 pixels = patch rows 2j, 2j+1
            umma_bf16_elect(d_tmem, a_lo + (uint32_t)j * (20u * 128u >> 4), kHiA, b_lo + (uint32_t)j * (2048u >> 4), kHiB, idesc,
                            (first && j == 0) ? 0u : 1u);
        }
        first = false;
        umma_commit_elect(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      umma_commit_elect(done_bar);
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


// ------------------------------------------------------------------------------------------------------------------
// Generic ("streaming") tensor-core weight gradient for everything the halo kernel does not cover: stride-2 convs over
// parity views, 1x1 convs / linear layers, ConvTranspose2d pixel-shuffle views, small images (14x14), 16/32-channel
// sources. Same GEMM orientation: D[M = 128 rows = (128/CK) K-units x CK input channels][N = output channels] over
// K = the pixels of one tile, with A = the units' input tiles at their tap offsets (one TMA box each, MN-major,
// leading-dimension stride = one unit tile) and B = the dy tile. A CTA owns (N tile, group of units, pixel split); the
// dy tile is staged once per pixel tile and re-used by all accumulators of the group.
struct alignas(64) WgStreamParams {
  CUtensorMap tm_x[PMOE_MAX_SRC];
  CUtensorMap tm_dy;
  int8_t unit_src[PMOE_MAX_SEG * 4];
  int8_t unit_dh[PMOE_MAX_SEG * 4];
  int8_t unit_dw[PMOE_MAX_SEG * 4];
  uint16_t unit_c0[PMOE_MAX_SEG * 4];
  int n_units;
  int N, n_tiles, n_acc, n_groups;
  int bw, bh, tiles_w, tiles_h, n_img;
  long long m_tiles;
  int splits, a_stages;
  float* dw;
  int ktot, cout_pad;
  // grouped mode (dw_img_stride > 0): every image is one GROUP with its own weight gradient, dw_img_stride floats apart — the K
  // experts' equally shaped Linear layers as one launch (expert = image axis of the stacked activations, model/moe.py:53-72);
  // a CTA's pixel range then lies inside one image
  long long dw_img_stride;
};
constexpr int kWgMaxUnits = PMOE_MAX_SEG * 4;
constexpr int kWgAStage = 32 * 1024;  // 128 pixel rows x 128 input channels (bf16)

template <int CK>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_tc_stream_kernel(const __grid_constant__ WgStreamParams p) {
  constexpr int UPM = 128 / CK;                // K-units per 128-row accumulator
  constexpr int ROWB = CK * 2;                 // bytes per pixel row of one unit tile
  constexpr int UNIT_BYTES = 128 * ROWB;
  constexpr uint32_t LAYOUT = CK == 64 ? kLayoutSW128 : (CK == 32 ? kLayoutSW64 : kLayoutSW32);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int dy_stage_bytes = (p.N / 64) * kWgDyBlock;
  uint8_t* dy_pipe = smem;
  uint8_t* a_pipe = smem + 2 * dy_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(a_pipe + (size_t)p.a_stages * kWgAStage);
  uint64_t* a_empty = a_full + kWgMaxStages;
  uint64_t* dy_full = a_empty + kWgMaxStages;
  uint64_t* dy_empty = dy_full + 2;
  uint64_t* done_bar = dy_empty + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(done_bar + 1);

  // warp index through a shuffle: provably warp-uniform, so the role branches below stay convergent for ptxas
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgMaxStages; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&dy_full[s], 1);
      mbar_init(&dy_empty[s], 1);
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
  // Pixel rows >= bw*bh of every stage are never written by TMA; they are reduction terms, so they must be zero.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (2 * dy_stage_bytes + p.a_stages * kWgAStage) / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (__any_sync(0xffffffffu, tmem_base != 0u)) {  // warp-uniform condition (see conv_wgrad_tc_kernel)
    if (threadIdx.x == 0) printf("pmoe conv_wgrad_tc_stream: unexpected TMEM base %u\n", tmem_base);
    __trap();
  }

  int u = blockIdx.x;
  const int split = u % p.splits;
  u /= p.splits;
  const int grp = u % p.n_groups;
  u /= p.n_groups;
  const int nt = u % p.n_tiles;
  const int gimg = u / p.n_tiles;  // 0 unless grouped
  // 32-bit tile arithmetic below: a 64-bit division is a subroutine call, and a call inside the lane-0 producer branch
  // makes ptxas give up on proving the MMA warp converged
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const bool grouped = p.dw_img_stride > 0;
  const int t_base = grouped ? gimg * tiles_per_img : 0;
  const int t_count = grouped ? tiles_per_img : (int)p.m_tiles;
  const int t_begin = t_base + (int)(((long long)t_count * split) / p.splits);
  const int t_end = t_base + (int)(((long long)t_count * (split + 1)) / p.splits);
  const int unit0 = grp * p.n_acc * UPM;
  int n_acc = (p.n_units - unit0 + UPM - 1) / UPM;  // accumulators with at least one live unit
  if (n_acc > p.n_acc) n_acc = p.n_acc;
  const uint32_t box_bytes = (uint32_t)(p.bw * p.bh * ROWB);
  const uint32_t dy_box_bytes = (uint32_t)(p.bw * p.bh * 128);

  if (warp == 0) {
    if (lane == 0) {
      int as = 0;
      uint32_t aphase = 0;
      uint32_t it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int img = t / tiles_per_img;
        const int rem = t % tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.bh, w0 = (rem % p.tiles_w) * p.bw;
        const uint32_t ds = it & 1u, dphase = (it >> 1) & 1u;
        mbar_wait(&dy_empty[ds], dphase ^ 1u);
        mbar_arrive_expect_tx(&dy_full[ds], dy_box_bytes * (uint32_t)(p.N / 64));
        for (int b = 0; b < p.N / 64; ++b)
          tma_load_4d(dy_pipe + ds * dy_stage_bytes + b * kWgDyBlock, &p.tm_dy, &dy_full[ds], nt * p.N + b * 64, w0, h0, img);
        for (int a = 0; a < n_acc; ++a) {
          mbar_wait(&a_empty[as], aphase ^ 1u);
          mbar_arrive_expect_tx(&a_full[as], box_bytes * UPM);
#pragma unroll
          for (int q = 0; q < UPM; ++q) {
            int un = unit0 + a * UPM + q;
            if (un >= p.n_units) un = p.n_units - 1;  // duplicate of the last unit: dropped by the epilogue
            tma_load_4d(a_pipe + (size_t)as * kWgAStage + q * UNIT_BYTES, &p.tm_x[p.unit_src[un]], &a_full[as], p.unit_c0[un],
                        w0 + p.unit_dw[un], h0 + p.unit_dh[un], img);
          }
          if (++as == p.a_stages) {
            as = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // converged warp: every lane runs the loop, the elected lane issues (see umma_bf16_elect)
      const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.N, 1, 1);
      constexpr uint32_t kHiA = umma_desc_hi(8u * ROWB, LAYOUT);
      constexpr uint32_t kHiB = umma_desc_hi(1024u, kLayoutSW128);
      int as = 0;
      uint32_t aphase = 0;
      uint32_t it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const uint32_t ds = it & 1u, dphase = (it >> 1) & 1u;
        mbar_wait_warp(&dy_full[ds], dphase);
        tc_fence_after();
        const uint32_t dyb = smem_u32(dy_pipe + ds * dy_stage_bytes);
        for (int a = 0; a < n_acc; ++a) {
          mbar_wait_warp(&a_full[as], aphase);
          tc_fence_after();
          const uint32_t ab = smem_u32(a_pipe + (size_t)as * kWgAStage);
          const uint32_t d_tmem = (uint32_t)(a * p.N);  // TMEM base is 0 (checked after the allocation)
          const uint32_t a_lo = umma_desc_lo(ab, (uint32_t)UNIT_BYTES);
          const uint32_t b_lo = umma_desc_lo(dyb, (uint32_t)kWgDyBlock);
#pragma unroll
          for (int j = 0; j < 8; ++j)  // K = 16 pixels per MMA
            umma_bf16_elect(d_tmem, a_lo + (uint32_t)j * (16u * ROWB >> 4), kHiA, b_lo + (uint32_t)j * (2048u >> 4), kHiB, idesc,
                            (it | (uint32_t)j) != 0u ? 1u : 0u);
          umma_commit_elect(&a_empty[as]);
          if (++as == p.a_stages) {
            as = 0;
            aphase ^= 1u;
          }
        }
        umma_commit_elect(&dy_empty[ds]);
      }
      umma_commit_elect(done_bar);
    }
  } else {
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // accumulator row = (unit within the accumulator) * CK + input channel
    const int uq = m / CK, ch = m % CK;
    if (t_end > t_begin) {
      for (int a = 0; a < n_acc; ++a) {
        const int un = unit0 + a * UPM + uq;
        const bool live = un < p.n_units;
        float* dst = p.dw + (long long)gimg * p.dw_img_stride + (long long)(live ? un : 0) * CK + ch;
        for (int cb = 0; cb < p.N; cb += 32) {
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * p.N + cb), raw);
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

static void wg_choose_tile(int H, int W, int* bh_out, int* bw_out) {
  long long best_tiles = -1;
  int best_bw = 1, best_bh = 1;
  const int wmax = W < 128 ? W : 128;
  for (int bw = 1; bw <= wmax; ++bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    if (bh < 1) continue;
    const long long tiles = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && bw * bh > best_bw * best_bh)) {
      best_tiles = tiles;
      best_bw = bw;
      best_bh = bh;
    }
  }
  *bh_out = best_bh;
  *bw_out = best_bw;
}

static int wg_tmap(CUtensorMap* tm, const PmoeView4& v, int box_c, int bw, int bh, CUtensorMapSwizzle swz, const char* what) {
  if (((uintptr_t)v.ptr & 15) || (v.sw % 8) || (v.sh % 8) || (v.sn % 8) || (v.c % 8)) {
    set_error("%s: view must be 16-byte aligned with strides/channels in multiples of 8 elements", what);
    return PMOE_ERR_ARG;
  }
  const uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  const uint64_t strides[3] = {(uint64_t)v.sw * 2, (uint64_t)v.sh * 2, (uint64_t)v.sn * 2};
  const uint32_t box[4] = {(uint32_t)box_c, (uint32_t)bw, (uint32_t)bh, 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, dims, strides, box, swz);
}

template <int CK>
static int launch_wg_stream(const WgStreamParams& p, int smem_bytes, long long grid, cudaStream_t stream) {
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_tc_stream_kernel<CK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      set_error("conv_wgrad_tc_stream<%d>: cannot reserve %d bytes of shared memory: %s", CK, smem_bytes, cudaGetErrorString(e));
      return PMOE_ERR_LAUNCH;
    }
    configured = smem_bytes;
  }
  conv_wgrad_tc_stream_kernel<CK><<<(unsigned)grid, kWgThreads, smem_bytes, stream>>>(p);
  return check_launch("conv_wgrad_tc_stream");
}

static int wgrad_stream(const PmoeConvTc* d, float* dwpack, cudaStream_t stream) {
  const PmoeView4& o = d->out;
  if ((d->ck != 16 && d->ck != 32 && d->ck != 64) || o.c % 64 != 0 || d->cout_pad % 64 != 0 || d->cout_pad < o.c) {
    set_error("conv_wgrad_tc: needs ck in {16,32,64} and a dy view whose channel count is a multiple of 64");
    return PMOE_ERR_UNSUPPORTED;
  }
  WgStreamParams p;
  memset(&p, 0, sizeof(p));
  int nu = 0;
  for (int i = 0; i < d->n_seg; ++i) {
    const PmoeSeg& s = d->seg[i];
    if (s.src < 0 || s.src >= d->n_src || s.nchunks == 0 || s.c0 + s.nchunks * d->ck > d->src[s.src].c) {
      set_error("conv_wgrad_tc: segment %d out of range", i);
      return PMOE_ERR_ARG;
    }
    for (int c = 0; c < s.nchunks; ++c, ++nu) {
      if (nu >= kWgMaxUnits) {
        set_error("conv_wgrad_tc: more than %d K units", kWgMaxUnits);
        return PMOE_ERR_UNSUPPORTED;
      }
      p.unit_src[nu] = s.src;
      p.unit_dh[nu] = s.dh;
      p.unit_dw[nu] = s.dw;
      p.unit_c0[nu] = (uint16_t)(s.c0 + c * d->ck);
    }
  }
  if (nu * d->ck != d->ktot) {
    set_error("conv_wgrad_tc: ktot %d does not match the segment list", d->ktot);
    return PMOE_ERR_ARG;
  }
  p.n_units = nu;
  p.N = o.c % 128 == 0 ? 128 : 64;
  p.n_tiles = o.c / p.N;
  p.n_acc = 512 / p.N;
  const int upm = 128 / d->ck;
  const int group_units = p.n_acc * upm;
  p.n_groups = (nu + group_units - 1) / group_units;
  wg_choose_tile(o.h, o.w, &p.bh, &p.bw);
  p.tiles_w = (o.w + p.bw - 1) / p.bw;
  p.tiles_h = (o.h + p.bh - 1) / p.bh;
  p.n_img = o.n;
  p.m_tiles = (long long)p.tiles_w * p.tiles_h * p.n_img;
  p.dw_img_stride = d->wpack_img_stride > 0 ? d->wpack_img_stride : 0;
  const long long n_grp_img = p.dw_img_stride > 0 ? p.n_img : 1;
  const long long units = (long long)p.n_tiles * p.n_groups * n_grp_img;
  const long long tiles_each = p.dw_img_stride > 0 ? (long long)p.tiles_w * p.tiles_h : p.m_tiles;
  long long splits = (long long)num_sms() / units;
  if (splits < 1) {
    // more (N tile, unit group, image) work items than SMs: split each into a few pixel ranges so that the last wave is full
    // (256 images on 148 SMs: 1 split = 2 waves of whole images, 4 splits = 7 waves of quarter images = 1.75)
    double best = 1e30;
    splits = 1;
    for (long long sp = 1; sp <= 8 && sp <= tiles_each; ++sp) {
      const double waves = (double)((units * sp + num_sms() - 1) / num_sms()) / (double)sp;
      if (waves < best - 1e-9) {
        best = waves;
        splits = sp;
      }
    }
  }
  if (splits > tiles_each) splits = tiles_each;
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  const int dy_stage = (p.N / 64) * kWgDyBlock;
  int a_stages = (214 * 1024 - 2 * dy_stage) / kWgAStage;
  if (a_stages > kWgMaxStages) a_stages = kWgMaxStages;
  p.a_stages = a_stages;
  const int smem_bytes = 1024 + 2 * dy_stage + a_stages * kWgAStage + (2 * kWgMaxStages + 5) * 8 + 16;
  p.dw = dwpack;
  p.ktot = d->ktot;
  p.cout_pad = d->cout_pad;
  const CUtensorMapSwizzle swz = d->ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (d->ck == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int rc;
  for (int i = 0; i < d->n_src; ++i)
    if ((rc = wg_tmap(&p.tm_x[i], d->src[i], d->ck, p.bw, p.bh, swz, "conv_wgrad_tc source")) != PMOE_OK) return rc;
  if ((rc = wg_tmap(&p.tm_dy, o, 64, p.bw, p.bh, CU_TENSOR_MAP_SWIZZLE_128B, "conv_wgrad_tc dy")) != PMOE_OK) return rc;
  const long long grid = units * p.splits;
  if (grid > 0x7fffffffLL) {
    set_error("conv_wgrad_tc: grid too large");
    return PMOE_ERR_ARG;
  }
  switch (d->ck) {
    case 64: return launch_wg_stream<64>(p, smem_bytes, grid, stream);
    case 32: return launch_wg_stream<32>(p, smem_bytes, grid, stream);
    default: return launch_wg_stream<16>(p, smem_bytes, grid, stream);
  }
}

}  // namespace pmoe

using namespace pmoe;

// Canonical 3x3/s1/p1 convs over whole 64-channel-chunked sources take the halo kernel; everything else the streaming
// kernel. Returns PMOE_ERR_UNSUPPORTED (nothing launched) only when dy has fewer than 64 stored channels or the K-unit
// list is too long; the caller then uses pmoe_conv_wgrad_simt.
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
  static const bool halo_off = getenv("PMOE_NO_HALO") != nullptr;
  if (halo_off || !ok || total_chunks > PMOE_MAX_SEG || 9 * total_chunks * 64 != d->ktot || d->wpack_img_stride > 0)
    return wgrad_stream(d, dwpack, stream);
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
