// Tensor-core implicit-GEMM convolution for sm_100a: TMA-staged NHWC tiles -> tcgen05.mma with the
// fp32 accumulator in TMEM -> fused epilogue (per-channel affine = folded BN / bias, residual add,
// activation, BN batch statistics, per-image channel sums for ECA / global avg-pool) -> bf16 NHWC
// via TMA store.
//
// GEMM view: M = output pixels (one 128-row tile = a BH x BW spatial patch of one image),
// N = output channels, K = list of "segments" (source view, spatial tap, channel range). The
// segment list is what makes one kernel serve 3x3 / 1x1 convs, virtual concat of several sources
// (U-Net skip connections, the PU-Net mask ring), stride-2 convs (parity views), ConvTranspose2d
// k2s2 (pixel-shuffle output views), data-gradients (flipped taps) and linear layers (1x1 conv).
// Zero padding comes from TMA out-of-bounds fill; partial tiles are clipped by the TMA store.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue.
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
#include "host_util.h"
#include "act.cuh"
#include "ptx.cuh"

#include <initializer_list>
#include <stdlib.h>
#include <string.h>

namespace pmoe {

struct TcSeg {
  int8_t src, dh, dw, pad;
  uint16_t c0, nchunks;
};

struct alignas(64) ConvTcParams {
  CUtensorMap tm_src[PMOE_MAX_SRC];
  CUtensorMap tm_w;
  // Output views: [0] = out; [1..3] = the extra views of a multi-view launch (ConvTranspose2d k2s2). The epilogue
  // stores straight from registers (each thread owns one pixel row: 32-byte st.global.v8), so no staging tile.
  CUtensorMap tm_out[4];
  int out_c;              // channels stored per view
  int out_bufs;           // staging buffers for the TMA store (2, or 1 when shared memory is tight)
  int out_cols;           // GEMM columns per output view (0 = single view)
  int tiles_n_varies;     // the N tile changes between a CTA's tiles (streaming kernel)
  // fused 2x2/stride-2 max-pool of the stored output (eval-mode U-Net encoder, unet.py:29)
  __nv_bfloat16* pool2_ptr;
  long long pool2_sn, pool2_sh, pool2_sw;
  int pool2_align32;
  // optional fp32 NCHW copy of the first nchw_c output channels (the module boundary, punet.py:118-120)
  float* nchw_ptr;
  long long nchw_sn, nchw_sc, nchw_sh, nchw_sw;
  int nchw_c;
  TcSeg seg[PMOE_MAX_SEG];
  int n_seg, kiters;
  int tiles_w, tiles_h, tiles_n, n_img;
  int bw, bh, H, W;
  long long total_tiles;
  const float* scale;
  const float* shift;
  int act;
  int res_c;
  const __nv_bfloat16* res;
  long long res_sn, res_sh, res_sw;
  double* stat_sum;
  double* stat_sq;
  float* pool_sum;
  int cout_pad;
  int pool_stride;
  // resident-weight ("halo") variant
  int halo_stages, w_slots, n_chunks, w_bytes;
  long long shift_img_stride;  // > 0: shift (bias) has one row per image (grouped expert layers: image = expert)
  int w_per_img;  // one weight set per image (tm_w is 3-D): resident halo kernel (ECA gate folded in) or streaming kernel (grouped experts) (ECA gate folded into the weights), tm_w is 3-D
  int stream_w;  // halo kernel with the weight tiles streamed through a ring of w_slots stages instead of resident
  long long m_tiles;
  int feat;                 // compile-time epilogue variant to use (-1 = generic)
  int pair;                 // streaming kernel launched as clusters of 2 with weight-tile multicast
  unsigned long long* dbg;  // optional per-CTA cycle counters (pmoe_conv_tc_set_debug): who waits for whom
};

// cycle accounting for tuning (8 counters per CTA): 0 producer waits for a free stage, 1 MMA waits for TMA data,
// 2 MMA waits for a free accumulator (= epilogue too slow), 3 MMA thread total, 4 epilogue waits for the accumulator
// (= main loop too slow), 5 epilogue total, 6 epilogue waits for its staging buffer (TMA store drain), 7 tiles.
constexpr int kDbgSlots = 16;  // counters per CTA; 8..12: epilogue phases (barrier 1, TMEM load, math+staging, fence+barrier 2, store issue)
struct DbgClock {
  unsigned long long* base;
  long long acc[8];
  __device__ __forceinline__ DbgClock(unsigned long long* b) : base(b), acc{0, 0, 0, 0, 0, 0, 0, 0} {}
  __device__ __forceinline__ long long now() const { return base ? clock64() : 0; }
  __device__ __forceinline__ void add(int i, long long t0) {
    if (base) acc[i] += clock64() - t0;
  }
  __device__ __forceinline__ void flush(int slot0, int n) {
    if (base)
      for (int i = 0; i < n; ++i) base[blockIdx.x * kDbgSlots + slot0 + i] = (unsigned long long)acc[i];
  }
};

constexpr int kMaxStatC = 512;
constexpr int kEpiWarps = 8;               // two warps per TMEM lane quadrant: they split the columns of a tile
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kNumThreads = 64 + kEpiThreads;

template <int BN, int CK>
struct TcCfg {
  static constexpr int OCW = BN < 64 ? BN : 64;   // channels per TMA store box
  static constexpr int SUB = OCW == 32 ? 16 : (OCW < 32 ? OCW : 32);  // columns per tcgen05.ld (32-wide chunks: one half per warp)
  static constexpr int A_BYTES = 128 * CK * 2;
  static constexpr int B_BYTES = BN * CK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = 128 * OCW * 2;
  static constexpr int OUT_BUFS = (BN >= 256 && CK == 64) ? 1 : 2;  // one staging tile buys the 4th pipeline stage at BN = 256
  static constexpr int BUDGET = 210 * 1024;
  static constexpr int STAGES_RAW = (BUDGET - OUT_BUFS * OUT_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int PIPE_BYTES = ((STAGES * STAGE_BYTES + 1023) / 1024) * 1024;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr uint32_t LAYOUT = CK == 64 ? kLayoutSW128 : (CK == 32 ? kLayoutSW64 : kLayoutSW32);
  static constexpr uint32_t SBO = 8 * CK * 2;  // 8 rows of one swizzle span
  static constexpr int AUX_FLOATS = 3 * BN + 2 * kMaxStatC;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + PIPE_BYTES + OUT_BUFS * OUT_BYTES + AUX_FLOATS * 4 +
                                    (2 * STAGES + 4) * 8 + 16;
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case PMOE_ACT_RELU: return fmaxf(x, 0.f);
    case PMOE_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case PMOE_ACT_TANH: return tanhf(x);
    case PMOE_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    default: return act_piecewise(x, act);
  }
}

// Column sums over the 32 rows held by a warp (row = lane, NC columns per lane) with a
// transpose-reduce butterfly: 31 (NC=32) shuffles instead of 32*5. Lane c ends with column c.
template <int NC>
__device__ __forceinline__ float warp_col_sums(float (&v)[NC], int lane) {
  if constexpr (NC == 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
#pragma unroll
  for (int off = (NC == 32 ? 16 : 8); off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Tile enumerators. StreamTiles: contiguous chunk of the (m, n) tile list per CTA (per-tap streaming kernel).
// ResidentTiles: the CTA owns ONE n-tile (its weights stay in shared memory) and strides over the m-tiles.
// All tile arithmetic is 32-bit on purpose: a 64-bit division is a subroutine CALL, and a call inside the lane-0 producer
// branch makes ptxas give up on proving the MMA warp converged (it then emits an ELECT + R2UR waterfall per MMA).
struct StreamTiles {
  int t_begin, t_end;
  int tiles_per_img;
  int mt_mul, mt_add;  // cluster pairs: the two CTAs of a pair take m-tiles 2i and 2i+1 of the same n-tile sequence
  __device__ __forceinline__ int nt_last(const ConvTcParams& p) const { return (t_end - 1) % p.tiles_n; }
  __device__ __forceinline__ bool get(const ConvTcParams& p, uint32_t iter, int& img, int& h0, int& w0, int& nt) const {
    const int t = t_begin + (int)iter;
    if (t >= t_end) return false;
    nt = t % p.tiles_n;
    const int mt = (t / p.tiles_n) * mt_mul + mt_add;
    img = mt / tiles_per_img;  // >= n_img for the padding tile of an odd pair: loads read zeros, stores are clipped
    const int rem = mt % tiles_per_img;
    h0 = (rem / p.tiles_w) * p.bh;
    w0 = (rem % p.tiles_w) * p.bw;
    return true;
  }
};
struct ResidentTiles {  // a contiguous run of m-tiles per CTA: an image boundary is crossed at most a few times
  int nt_fixed, tiles_per_img;
  int m_first, m_end;
  __device__ __forceinline__ int nt_last(const ConvTcParams&) const { return nt_fixed; }
  __device__ __forceinline__ bool get(const ConvTcParams& p, uint32_t iter, int& img, int& h0, int& w0, int& nt) const {
    const int mt = m_first + (int)iter;
    if (mt >= m_end) return false;
    nt = nt_fixed;
    img = mt / tiles_per_img;
    const int rem = mt % tiles_per_img;
    h0 = (rem / p.tiles_w) * p.bh;
    w0 = (rem % p.tiles_w) * p.bw;
    return true;
  }
};

__device__ __forceinline__ void stg256(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void stg128(void* ptr, const uint32_t* v) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
// 16 consecutive bf16 channels (8 packed words) to a 16-byte aligned address
__device__ __forceinline__ void store16(__nv_bfloat16* ptr, const uint32_t* v, bool align32) {
  if (align32) {
    stg256(ptr, v);
  } else {
    stg128(ptr, v);
    stg128(ptr + 8, v + 4);
  }
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Epilogue role (8 warps, threads 64..319): TMEM -> registers -> affine / residual / activation / statistics -> bf16 ->
// swizzled staging tile -> TMA store. Two warps share each TMEM lane quadrant and split every 64-column chunk between
// them (32 columns each), so the dependent LDTM -> math -> pack chain of a tile runs on two warps per scheduler. The
// main store goes through shared memory + TMA because 32 rows x 32 B register stores are one L1 transaction per row:
// measured 2-2.5x slower than the bulk store on the write-heavy 224^2 layers. Optional extras ride on the registers: a
// fused 2x2 max-pool (two warp shuffles: the window partners are lanes l^1 and l^bw), an fp32 NCHW copy at the module
// boundary, BN batch statistics and per-image channel sums. Shared by both main-loop variants.
// FEAT < 0: every optional feature is a run-time flag (generic). FEAT >= 0: the feature set is fixed at compile time
// (kFeatShift | kFeatRelu | kFeatStats | kFeatPool2 | kFeatPoolSum | kFeatNchw | kFeatRes; no scale / other activations), which turns the per-element path into straight-line code (~3 instead of ~11 instructions per element: the
// epilogue, not the tensor core, was the limit of the 64-channel layers).
constexpr int kFeatShift = 1, kFeatRelu = 2, kFeatStats = 4, kFeatPool2 = 8, kFeatPoolSum = 16, kFeatNchw = 32, kFeatRes = 64,
              kFeatElu = 128;
// ELU for the compile-time epilogue of the expert heads (make_mlp act='elu', conf/stage_2.yaml:83-90): expm1f is a ~25-instruction
// libdevice routine and made the 512-wide ELU layers 2-4x slower than their ReLU twins at 64k rows. exp(x) - 1 through the SFU, with a
// cubic series near zero where that difference cancels (|error| < 1e-5 relative, far inside the bf16 rounding of the stored value).
__device__ __forceinline__ float elu_fast(float x) {
  const float series = x * fmaf(x, fmaf(x, 0.16666667f, 0.5f), 1.f);
  const float ex = __expf(x) - 1.f;
  return x > 0.f ? x : (x > -0.0625f ? series : ex);
}
template <int BN, int OCW, int SUB, int OUT_BYTES, int FEAT, class Iter>
__device__ __forceinline__ void run_epilogue(const ConvTcParams& p, const Iter& it, uint8_t* out_stage, float* s_scale, float* s_shift,
                                             float* s_pool, float* s_sum, float* s_sq, uint64_t* tfull_bar, uint64_t* tempty_bar,
                                             int warp, int lane) {
  const int e = threadIdx.x - 64;         // 0..kEpiThreads-1
  const int quad = warp & 3;              // TMEM lane quadrant this warp may read (warp id % 4)
  const int half = (warp - 2) >> 2;       // which of the quadrant's two warps
  const int row = quad * 32 + lane;       // accumulator row == pixel index inside the tile
  // Two store issuers (one per staging buffer, in different warps): issuing a bulk tensor store costs its thread ~200
  // cycles, which would otherwise serialise consecutive chunks.
  const bool issuer = (e == 0);
  const bool issuer_b = (e == 128);
  const int ti = row / p.bw, tj = row % p.bw;
  const bool row_in_box = row < p.bw * p.bh;
  constexpr int ROWB = OCW * 2;
  constexpr uint32_t SWMASK = ROWB == 128 ? 7u : (ROWB == 64 ? 3u : 1u);
  constexpr int NSB = OCW / SUB;          // 32-column blocks per chunk: 2 (one per warp of the pair) or 1
  const bool worker = half < NSB;         // with a single block per chunk the second warp only keeps the barriers
  constexpr bool kGen = FEAT < 0;
  const int act = kGen ? p.act : ((FEAT & kFeatRelu) ? PMOE_ACT_RELU : ((FEAT & kFeatElu) ? PMOE_ACT_ELU : PMOE_ACT_NONE));
  const bool has_scale = kGen && p.scale != nullptr;
  const bool has_shift = kGen ? (p.shift != nullptr) : (FEAT & kFeatShift) != 0;
  const bool has_stats = kGen ? (p.stat_sum != nullptr) : (FEAT & kFeatStats) != 0;
  const bool has_pool2 = kGen ? (p.pool2_ptr != nullptr) : (FEAT & kFeatPool2) != 0;
  const bool has_res = kGen ? (p.res != nullptr) : (FEAT & kFeatRes) != 0;
  const bool has_nchw = kGen ? (p.nchw_ptr != nullptr) : (FEAT & kFeatNchw) != 0;
  const bool has_poolsum = kGen ? (p.pool_sum != nullptr) : (FEAT & kFeatPoolSum) != 0;
  const uint32_t sc_addr = smem_u32(s_scale), sh_addr = smem_u32(s_shift);
  const bool pool_owner = ((ti | tj) & 1) == 0;  // this lane holds the top-left pixel of a 2x2 pooling window
  // Per-image channel sums (ECA / avg-pool numerators): with one chunk per tile each thread keeps running sums of its own
  // row in registers and reduces them across the warp only when the image changes (tiles of a CTA are contiguous).
  const bool defer_pool = (BN / OCW) == 1 && !p.tiles_n_varies;
  float psum[SUB];
#pragma unroll
  for (int k = 0; k < SUB; ++k) psum[k] = 0.f;
  int pimg = -1;
  auto flush_pool = [&](int image, int n0_) {
    const float sa = warp_col_sums<SUB>(psum, lane);
    if (lane < SUB && sa != 0.f) atomicAdd(p.pool_sum + (long long)image * p.pool_stride + n0_ + half * SUB + lane, sa);
#pragma unroll
    for (int k = 0; k < SUB; ++k) psum[k] = 0.f;
  };
  // BN batch statistics under the same condition: per-thread running sums of the thread's own row (raw accumulator and
  // its square), one transpose-reduce per CTA instead of two per tile (62 shuffles + 64 shared-memory atomics per warp and
  // tile were ~25 % of the epilogue of the 224^2 training layers).
  float ssa[SUB], ssb[SUB];
#pragma unroll
  for (int k = 0; k < SUB; ++k) ssa[k] = ssb[k] = 0.f;
  uint32_t nstore = 0;
  int shift_img = -1;
  DbgClock dc(issuer ? p.dbg : nullptr);
  const long long t_start = dc.now();
  uint32_t ntiles = 0;
  int img, h0, w0, nt;
  for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
    ++ntiles;
    const int n0 = nt * BN;
    if (defer_pool && has_poolsum && worker && img != pimg) {
      if (pimg >= 0) flush_pool(pimg, n0);
      pimg = img;
    }
    const int oh = h0 + ti, ow = w0 + tj;
    const bool valid = row_in_box && oh < p.H && ow < p.W && img < p.n_img;
    const uint32_t acc = titer & 1u;
    const uint32_t acc_phase = (titer >> 1) & 1u;

    if (titer == 0 || p.tiles_n_varies || (p.shift_img_stride > 0 && img != shift_img)) {
      // resident kernels keep one N tile: the per-channel affine is loaded once
      // (all reads of the previous tile's values happened before its last named barrier 2)
      shift_img = img;
      const long long so = p.shift_img_stride > 0 && img < p.n_img ? (long long)img * p.shift_img_stride : 0;
      for (int i = e; i < BN; i += kEpiThreads) {
        if (has_scale) s_scale[i] = __ldg(p.scale + n0 + i);
        if (has_shift) s_shift[i] = __ldg(p.shift + so + n0 + i);
      }
    }
    const long long wt = dc.now();
    mbar_wait(&tfull_bar[acc], acc_phase);
    dc.add(0, wt);
    tc_fence_after();

    const __nv_bfloat16* res_row = nullptr;
    if (has_res && valid) res_row = p.res + (long long)img * p.res_sn + (long long)oh * p.res_sh + (long long)ow * p.res_sw;

#pragma unroll 1
    for (int ch = 0; ch < BN / OCW; ++ch) {
      const uint32_t buf = p.out_bufs == 2 ? (nstore & 1u) : 0u;
      uint8_t* obuf = out_stage + buf * OUT_BYTES;
      const bool my_store = buf == 0 ? issuer : issuer_b;
      if (my_store) {  // the store that last used this buffer (always issued by this thread) has drained
        const long long ws = dc.now();
        tma_store_wait_read<0>();
        dc.add(2, ws);
      }
      const long long tb1 = dc.now();
      named_bar_sync(1, kEpiThreads);
      dc.add(3, tb1);
      const long long tld = dc.now();
      if (worker) {
        const int cb = ch * OCW + half * SUB;  // column offset inside the N tile
        uint32_t raw[SUB];
        const uint32_t taddr = ((uint32_t)(quad * 32) << 16) + acc * BN + cb;  // TMEM base is 0
        if constexpr (SUB == 32) tmem_ld_32x32(taddr, raw);
        else tmem_ld_32x16(taddr, raw);
        tmem_ld_wait();
        dc.add(4, tld);
        float y[SUB];
#pragma unroll
        for (int k = 0; k < SUB; ++k) y[k] = __uint_as_float(raw[k]);
        // per-channel affine: 128-bit broadcast LDS of scale / shift, each only when present (eval-mode BN is folded
        // into the packed weights by the host, so inference needs the shift alone)
        if (has_scale) {
#pragma unroll
          for (int q = 0; q < SUB / 4; ++q) {
            const float4 sc4 = lds128(sc_addr + (uint32_t)(cb + 4 * q) * 4u);
            y[4 * q + 0] *= sc4.x;
            y[4 * q + 1] *= sc4.y;
            y[4 * q + 2] *= sc4.z;
            y[4 * q + 3] *= sc4.w;
          }
        }
        if (has_shift) {
#pragma unroll
          for (int q = 0; q < SUB / 4; ++q) {
            const float4 sh4 = lds128(sh_addr + (uint32_t)(cb + 4 * q) * 4u);
            y[4 * q + 0] += sh4.x;
            y[4 * q + 1] += sh4.y;
            y[4 * q + 2] += sh4.z;
            y[4 * q + 3] += sh4.w;
          }
        }
        if (has_stats) {
          if (defer_pool) {
            if (valid) {
#pragma unroll
              for (int k = 0; k < SUB; ++k) {
                const float f = __uint_as_float(raw[k]);
                ssa[k] += f;
                ssb[k] = fmaf(f, f, ssb[k]);
              }
            }
          } else {
            float a[SUB], b[SUB];
#pragma unroll
            for (int k = 0; k < SUB; ++k) {
              const float f = valid ? __uint_as_float(raw[k]) : 0.f;
              a[k] = f;
              b[k] = f * f;
            }
            const float sa = warp_col_sums<SUB>(a, lane);
            const float sq = warp_col_sums<SUB>(b, lane);
            if (lane < SUB && n0 + cb + lane < kMaxStatC) {
              atomicAdd(&s_sum[n0 + cb + lane], sa);
              atomicAdd(&s_sq[n0 + cb + lane], sq);
            }
          }
        }
        if (has_res && res_row != nullptr) {
#pragma unroll
          for (int q = 0; q < SUB / 8; ++q) {
            if (n0 + cb + q * 8 < p.res_c) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res_row + n0 + cb + q * 8));
              const __nv_bfloat162* r2 = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f2 = __bfloat1622float2(r2[u]);
                y[q * 8 + 2 * u] += f2.x;
                y[q * 8 + 2 * u + 1] += f2.y;
              }
            }
          }
        }
        if (act == PMOE_ACT_RELU) {  // the branch is uniform and hoisted out of the element loop
#pragma unroll
          for (int k = 0; k < SUB; ++k) y[k] = fmaxf(y[k], 0.f);
        } else if (!kGen && act == PMOE_ACT_ELU) {
#pragma unroll
          for (int k = 0; k < SUB; ++k) y[k] = elu_fast(y[k]);
        } else if (act != PMOE_ACT_NONE) {
#pragma unroll
          for (int k = 0; k < SUB; ++k) y[k] = apply_act(y[k], act);
        }
        // bf16 pack -> swizzled staging tile (row = pixel, ROWB bytes per row)
        uint32_t pk[SUB / 2];
#pragma unroll
        for (int k = 0; k < SUB / 2; ++k) pk[k] = pack_bf16x2(y[2 * k], y[2 * k + 1]);
#pragma unroll
        for (int q = 0; q < SUB / 8; ++q) {
          uint32_t off = (uint32_t)row * ROWB + (uint32_t)(half * SUB + q * 8) * 2u;
          off ^= ((off >> 7) & SWMASK) << 4;
          *reinterpret_cast<uint4*>(obuf + off) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        const int col = n0 + cb;
        if (has_pool2) {
          // 2x2 max over pixels (ti, tj), (ti, tj^1), (ti^1, tj), (ti^1, tj^1) = lanes l, l^1, l^bw, l^bw^1 (bw is a
          // power of two <= 16 and the tile origin is even, both guaranteed by the host)
          uint32_t m[SUB / 2];
#pragma unroll
          for (int k = 0; k < SUB / 2; ++k) {
            const uint32_t v = bf16x2_max(pk[k], __shfl_xor_sync(0xffffffffu, pk[k], 1));
            m[k] = bf16x2_max(v, __shfl_xor_sync(0xffffffffu, v, p.bw));
          }
          if (valid && pool_owner) {
            __nv_bfloat16* prow = p.pool2_ptr + (long long)img * p.pool2_sn + (long long)(oh >> 1) * p.pool2_sh +
                                  (long long)(ow >> 1) * p.pool2_sw + col;
#pragma unroll
            for (int q = 0; q < SUB / 16; ++q)
              if (col + q * 16 < p.out_c) store16(prow + q * 16, m + q * 8, p.pool2_align32 != 0);
          }
        }
        if (has_nchw && valid) {
          float* nrow = p.nchw_ptr + (long long)img * p.nchw_sn + (long long)oh * p.nchw_sh + (long long)ow * p.nchw_sw;
#pragma unroll
          for (int k = 0; k < SUB; ++k)
            if (col + k < p.nchw_c) nrow[(long long)(col + k) * p.nchw_sc] = y[k];
        }
        if (has_poolsum) {
          // pool what is actually stored (bf16-rounded), masked to valid pixels
          if (defer_pool) {
            if (valid) {
#pragma unroll
              for (int k = 0; k < SUB; ++k) psum[k] += __bfloat162float(__float2bfloat16_rn(y[k]));
            }
          } else {
            float a[SUB];
#pragma unroll
            for (int k = 0; k < SUB; ++k) a[k] = valid ? __bfloat162float(__float2bfloat16_rn(y[k])) : 0.f;
            const float sa = warp_col_sums<SUB>(a, lane);
            if (lane < SUB) atomicAdd(&s_pool[cb + lane], sa);
          }
        }
      }
      dc.add(5, tld);  // TMEM load + math + staging (+ extras)
      const long long tf = dc.now();
      fence_proxy_async_smem();
      named_bar_sync(2, kEpiThreads);
      dc.add(6, tf);
      const long long ti_ = dc.now();
      if (my_store) {
        const int col = n0 + ch * OCW;
        if (p.out_cols > 0) tma_store_4d(&p.tm_out[col / p.out_cols], obuf, col % p.out_cols, w0, h0, img);
        else tma_store_4d(&p.tm_out[0], obuf, col, w0, h0, img);
        tma_store_commit();
      }
      dc.add(7, ti_);
      ++nstore;
    }
    // accumulator stage fully read -> hand it back to the MMA warp
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    if (has_poolsum && !defer_pool) {
      named_bar_sync(3, kEpiThreads);
      for (int i = e; i < BN; i += kEpiThreads) {
        const float v = s_pool[i];
        if (v != 0.f) atomicAdd(p.pool_sum + (long long)img * p.pool_stride + n0 + i, v);
        s_pool[i] = 0.f;
      }
    }
  }
  if (defer_pool && has_poolsum && worker && pimg >= 0) flush_pool(pimg, it.nt_last(p) * BN);
  if (issuer || issuer_b) tma_store_wait_read<0>();
  if (issuer && p.dbg) {
    dc.add(1, t_start);
    dc.flush(4, 3);
    p.dbg[blockIdx.x * kDbgSlots + 7] = ntiles;
    for (int i = 3; i < 8; ++i) p.dbg[blockIdx.x * kDbgSlots + 5 + i] = (unsigned long long)dc.acc[i];
  }
  if (has_stats && defer_pool && worker && ntiles > 0) {
    const int col = it.nt_last(p) * BN + half * SUB + lane;  // one chunk per tile and one n-tile per CTA
    const float sa = warp_col_sums<SUB>(ssa, lane);
    const float sq = warp_col_sums<SUB>(ssb, lane);
    if (lane < SUB && col < kMaxStatC) {
      atomicAdd(&s_sum[col], sa);
      atomicAdd(&s_sq[col], sq);
    }
  }
  if (has_stats) {
    named_bar_sync(3, kEpiThreads);
    const int nc = p.cout_pad < kMaxStatC ? p.cout_pad : kMaxStatC;
    for (int i = e; i < nc; i += kEpiThreads) {
      atomicAdd(p.stat_sum + i, (double)s_sum[i]);
      atomicAdd(p.stat_sq + i, (double)s_sq[i]);
    }
  }
}

template <int BN, int CK, int FEAT>
__global__ void __launch_bounds__(kNumThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using C = TcCfg<BN, CK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* pipe = smem;
  uint8_t* out_stage = smem + C::PIPE_BYTES;
  float* s_scale = reinterpret_cast<float*>(out_stage + C::OUT_BUFS * C::OUT_BYTES);
  float* s_shift = s_scale + BN;
  float* s_pool = s_shift + BN;
  float* s_sum = s_pool + BN;
  float* s_sq = s_sum + kMaxStatC;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_sq + kMaxStatC);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform: role branches stay convergent
  const int lane = threadIdx.x & 31;

  // Cluster pairs (p.pair): the two CTAs run the same (n-tile, K) sequence on neighbouring m-tiles; each loads HALF of
  // every weight tile and multicasts it to both, halving the L2 -> SM weight traffic that bounds the K >= 1152 layers.
  const bool pair = p.pair != 0;
  const uint32_t crank = pair ? cluster_ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], pair ? 2 : 1);  // both consumers release a stage that the peer's multicast writes
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_src[0]);
  }
  if (warp == 2) {
    tmem_alloc(s_tmem, C::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += kNumThreads) s_pool[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * kMaxStatC; i += kNumThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // the peer's barriers exist before anything remote can reach them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (__any_sync(0xffffffffu, tmem_base != 0u)) {  // one CTA per SM, one allocation per CTA: the MMA issuer relies on base 0
    if (threadIdx.x == 0) printf("pmoe conv_tc: unexpected TMEM base %u\n", tmem_base);  // (warp-uniform condition, see the halo kernel)
    __trap();
  }
  const uint32_t pipe_addr = smem_u32(pipe);

  StreamTiles it;
  {
    const long long G = pair ? gridDim.x / 2 : gridDim.x;      // work units are split over CTAs, or over pairs
    const long long b = pair ? blockIdx.x / 2 : blockIdx.x;
    it.t_begin = (int)((p.total_tiles * b) / G);
    it.t_end = (int)((p.total_tiles * (b + 1)) / G);
    it.tiles_per_img = p.tiles_w * p.tiles_h;
    it.mt_mul = pair ? 2 : 1;
    it.mt_add = (int)crank;
  }
  int img, h0, w0, nt;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      DbgClock dc(p.dbg);
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        int kofs = 0;
        for (int sidx = 0; sidx < p.n_seg; ++sidx) {
          const TcSeg sg = p.seg[sidx];
          for (int c = 0; c < sg.nchunks; ++c) {
            const long long w0c = dc.now();
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            dc.add(0, w0c);
            uint8_t* a_dst = pipe + stage * C::STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(p.bw * p.bh * CK * 2 + C::B_BYTES));
            tma_load_4d(a_dst, &p.tm_src[sg.src], &full_bar[stage], sg.c0 + c * CK, w0 + sg.dw, h0 + sg.dh, img);
            if (pair)  // my half of the weight tile, to both CTAs (the other half arrives from the peer)
              tma_load_2d_mc(a_dst + C::A_BYTES + crank * (C::B_BYTES / 2), &p.tm_w, &full_bar[stage], kofs,
                             nt * BN + (int)crank * (BN / 2), (uint16_t)3);
            else if (p.w_per_img)  // grouped layers: the "image" index selects the expert's weights
              tma_load_3d(a_dst + C::A_BYTES, &p.tm_w, &full_bar[stage], kofs, nt * BN, img);
            else
              tma_load_2d(a_dst + C::A_BYTES, &p.tm_w, &full_bar[stage], kofs, nt * BN);
            kofs += CK;
            if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
      dc.flush(0, 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      constexpr uint32_t desc_hi = umma_desc_hi(C::SBO, C::LAYOUT);
      DbgClock dc(lane == 0 ? p.dbg : nullptr);
      const long long t_start = dc.now();
      int stage = 0;
      uint32_t phase = 0;
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        const uint32_t acc = titer & 1u;
        const uint32_t acc_phase = (titer >> 1) & 1u;
        const long long wa = dc.now();
        mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u);
        dc.add(1, wa);
        tc_fence_after();
        const uint32_t d_tmem = acc * BN;  // TMEM base is 0 (checked after the allocation)
        for (int ki = 0; ki < p.kiters; ++ki) {
          const long long wf = dc.now();
          mbar_wait_warp(&full_bar[stage], phase);
          dc.add(0, wf);
          tc_fence_after();
          const uint32_t a_lo = umma_desc_lo(pipe_addr + (uint32_t)stage * C::STAGE_BYTES, 16u);
          const uint32_t b_lo = a_lo + (C::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < CK / 16; ++k)
            umma_bf16_elect(d_tmem, a_lo + 2 * k, desc_hi, b_lo + 2 * k, desc_hi, idesc, (ki | k) != 0 ? 1u : 0u);
          if (pair) umma_commit_mc_elect(&empty_bar[stage], (uint16_t)3);
          else umma_commit_elect(&empty_bar[stage]);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_elect(&tfull_bar[acc]);
      }
      dc.add(2, t_start);
      dc.flush(1, 3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    run_epilogue<BN, C::OCW, C::SUB, C::OUT_BYTES, FEAT>(p, it, out_stage, s_scale, s_shift, s_pool, s_sum, s_sq, tfull_bar,
                                                         tempty_bar, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (pair) cluster_sync_all();  // the peer may still arrive on this CTA's barriers until it is done as well
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Resident-weight / halo variant for 3x3 stride-1 convolutions whose weight slice fits in shared memory.
// One TMA brings the (16+2) x (8+2) input halo of a 16x8 output patch for a 64-channel chunk; the nine taps are nine
// UMMA descriptor VIEWS of that halo tile (start row r*10+s, 8-row groups 10 rows apart — the swizzle is a function of
// the absolute shared-memory address, so row-shifted views need no re-staging: profiles/r01_conv_bringup.json probe).
// L2->SMEM traffic per tile drops from 9 x 16 KB to 23 KB per chunk, and the weights are fetched once per CTA.
constexpr int kMaxWStages = 8;
template <int BN, int CK>
struct HaloCfg {
  static constexpr int OCW = BN < 64 ? BN : 64;
  static constexpr int SUB = OCW == 32 ? 16 : (OCW < 32 ? OCW : 32);
  static constexpr int OUT_BYTES = 128 * OCW * 2;
  static constexpr int ROWB = CK * 2;                                       // bytes per halo pixel row (one swizzle span)
  static constexpr int HALO_TX = 18 * 10 * ROWB;
  static constexpr int HALO_BYTES = ((HALO_TX + 1023) / 1024) * 1024;       // 23 KB at CK = 64
  static constexpr int WSLOT_BYTES = BN * ROWB;
  static constexpr uint32_t LAYOUT = CK == 64 ? kLayoutSW128 : (CK == 32 ? kLayoutSW64 : kLayoutSW32);
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int AUX_FLOATS = 3 * BN + 2 * kMaxStatC;
  static constexpr int MAX_STAGES = 6;
};

template <int BN, int CK, int FEAT>
__global__ void __launch_bounds__(kNumThreads, 1) conv_tc_halo_kernel(const __grid_constant__ ConvTcParams p) {
  using C = HaloCfg<BN, CK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* halo = wsm + (size_t)p.w_bytes;
  uint8_t* out_stage = halo + (size_t)p.halo_stages * C::HALO_BYTES;
  float* s_scale = reinterpret_cast<float*>(out_stage + p.out_bufs * C::OUT_BYTES);
  float* s_shift = s_scale + BN;
  float* s_pool = s_shift + BN;
  float* s_sum = s_pool + BN;
  float* s_sq = s_sum + kMaxStatC;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_sq + kMaxStatC);
  uint64_t* empty_bar = full_bar + C::MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + C::MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* wfull_bar = tempty_bar + 2;
  uint64_t* wfree_bar = wfull_bar + 1;          // per-image weights: the MMA warp is done with the resident set
  uint64_t* wb_full = wfree_bar + 1;            // streamed-weight ring (p.stream_w): one barrier pair per weight stage
  uint64_t* wb_empty = wb_full + kMaxWStages;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(wb_empty + kMaxWStages);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // provably warp-uniform: role branches stay convergent
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::MAX_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < kMaxWStages; ++s) {
      mbar_init(&wb_full[s], 1);
      mbar_init(&wb_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], kEpiWarps);
    }
    mbar_init(wfull_bar, 1);
    mbar_init(wfree_bar, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_src[0]);
  }
  if (warp == 2) {
    tmem_alloc(s_tmem, C::TMEM_COLS);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < BN; i += kNumThreads) s_pool[i] = 0.f;
  for (int i = threadIdx.x; i < 2 * kMaxStatC; i += kNumThreads) s_sum[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (__any_sync(0xffffffffu, tmem_base != 0u)) {  // warp-uniform on purpose: a thread-dependent trap is a possible partial exit,
    if (threadIdx.x == 0) printf("pmoe conv_tc_halo: unexpected TMEM base %u\n", tmem_base);  // after which ptxas cannot prove
    __trap();                                                                                // the MMA warp converged
  }

  ResidentTiles it;
  it.nt_fixed = (int)(blockIdx.x % p.tiles_n);
  {
    const long long c = blockIdx.x / p.tiles_n, per_n = gridDim.x / p.tiles_n;
    it.m_first = (int)((p.m_tiles * c) / per_n);
    it.m_end = (int)((p.m_tiles * (c + 1)) / per_n);
  }
  it.tiles_per_img = p.tiles_w * p.tiles_h;
  int img, h0, w0, nt;

  if (warp == 0) {
    if (lane == 0 && p.stream_w) {
      // Weights too large to stay resident: the (tap, chunk) weight tiles stream through a ring of p.w_slots stages while
      // the halo tiles keep their own ring, one chunk ahead. Per 64-channel chunk the SM receives 23 KB of input instead of
      // 9 x 16 KB: the K >= 1152 layers are bound by the bytes an SM can take in per cycle, not by L2 or the tensor core.
      DbgClock dc(p.dbg);
      int hs = 0, ws = 0;
      uint32_t hphase = 0, wphase = 0;
      auto issue_halo = [&](uint32_t titer_, int g_) {
        int img_, h0_, w0_, nt_;
        if (!it.get(p, titer_, img_, h0_, w0_, nt_)) return;
        const TcSeg sg = p.seg[g_];
        const long long w0c = dc.now();
        mbar_wait(&empty_bar[hs], hphase ^ 1u);
        dc.add(0, w0c);
        mbar_arrive_expect_tx(&full_bar[hs], (uint32_t)C::HALO_TX);
        tma_load_4d(halo + (size_t)hs * C::HALO_BYTES, &p.tm_src[sg.src], &full_bar[hs], sg.c0, w0_ - 1, h0_ - 1, img_);
        if (++hs == p.halo_stages) {
          hs = 0;
          hphase ^= 1u;
        }
      };
      issue_halo(0, 0);
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        for (int g = 0; g < p.n_chunks; ++g) {
          if (g + 1 < p.n_chunks) issue_halo(titer, g + 1);  // keep the input one chunk ahead of the weights
          else issue_halo(titer + 1, 0);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(&wb_empty[ws], wphase ^ 1u);
            mbar_arrive_expect_tx(&wb_full[ws], (uint32_t)C::WSLOT_BYTES);
            tma_load_2d(wsm + (size_t)ws * C::WSLOT_BYTES, &p.tm_w, &wb_full[ws], (t * p.n_chunks + g) * CK, it.nt_fixed * BN);
            if (++ws == p.w_slots) {
              ws = 0;
              wphase ^= 1u;
            }
          }
        }
      }
      dc.flush(0, 1);
    } else if (lane == 0) {
      if (!p.w_per_img) {
        mbar_arrive_expect_tx(wfull_bar, (uint32_t)(p.w_slots * C::WSLOT_BYTES));
        for (int j = 0; j < p.w_slots; ++j) tma_load_2d(wsm + (size_t)j * C::WSLOT_BYTES, &p.tm_w, wfull_bar, j * CK, it.nt_fixed * BN);
      }
      DbgClock dc(p.dbg);
      int stage = 0;
      uint32_t phase = 0;
      int w_img = -1;
      uint32_t w_loads = 0;
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        if (p.w_per_img && img != w_img) {
          // a new image: its own (ECA-gated) weight set replaces the resident one once the MMA warp has released it
          if (w_loads > 0) mbar_wait(wfree_bar, (w_loads - 1) & 1u);
          mbar_arrive_expect_tx(wfull_bar, (uint32_t)(p.w_slots * C::WSLOT_BYTES));
          for (int j = 0; j < p.w_slots; ++j)
            tma_load_3d(wsm + (size_t)j * C::WSLOT_BYTES, &p.tm_w, wfull_bar, j * CK, it.nt_fixed * BN, img);
          w_img = img;
          ++w_loads;
        }
        for (int g = 0; g < p.n_chunks; ++g) {
          const TcSeg sg = p.seg[g];
          const long long w0c = dc.now();
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          dc.add(0, w0c);
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)C::HALO_TX);
          tma_load_4d(halo + (size_t)stage * C::HALO_BYTES, &p.tm_src[sg.src], &full_bar[stage], sg.c0, w0 - 1, h0 - 1, img);
          if (++stage == p.halo_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      dc.flush(0, 1);
    }
  } else if (warp == 1) {
    if (p.stream_w) {  // converged warp; weights arrive through the ring
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      constexpr uint32_t a_hi = umma_desc_hi(10u * C::ROWB, C::LAYOUT);
      constexpr uint32_t b_hi = umma_desc_hi(8u * C::ROWB, C::LAYOUT);
      const uint32_t w_lo = umma_desc_lo(smem_u32(wsm), 16u);
      const uint32_t halo_lo = umma_desc_lo(smem_u32(halo), 16u);
      DbgClock dc(lane == 0 ? p.dbg : nullptr);
      const long long t_start = dc.now();
      int hs = 0, ws = 0;
      uint32_t hphase = 0, wphase = 0;
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        const uint32_t acc = titer & 1u;
        const uint32_t acc_phase = (titer >> 1) & 1u;
        const long long wa = dc.now();
        mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u);
        dc.add(1, wa);
        tc_fence_after();
        const uint32_t d_tmem = acc * BN;
        for (int g = 0; g < p.n_chunks; ++g) {
          const long long wf = dc.now();
          mbar_wait_warp(&full_bar[hs], hphase);
          dc.add(0, wf);
          tc_fence_after();
          const uint32_t h_lo = halo_lo + (uint32_t)hs * (uint32_t)(C::HALO_BYTES >> 4);
#pragma unroll 1
          for (int t = 0; t < 9; ++t) {
            const long long wf2 = dc.now();
            mbar_wait_warp(&wb_full[ws], wphase);
            dc.add(0, wf2);
            tc_fence_after();
            const uint32_t a_lo = h_lo + (uint32_t)(((t / 3) * 10 + (t % 3)) * C::ROWB >> 4);
            const uint32_t b_lo = w_lo + (uint32_t)ws * (uint32_t)(C::WSLOT_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < CK / 16; ++k)
              umma_bf16_elect(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (g | t | k) != 0 ? 1u : 0u);
            umma_commit_elect(&wb_empty[ws]);
            if (++ws == p.w_slots) {
              ws = 0;
              wphase ^= 1u;
            }
          }
          umma_commit_elect(&empty_bar[hs]);
          if (++hs == p.halo_stages) {
            hs = 0;
            hphase ^= 1u;
          }
        }
        umma_commit_elect(&tfull_bar[acc]);
      }
      dc.add(2, t_start);
      dc.flush(1, 3);
    } else {  // converged warp: every lane runs the loop, the elected lane issues (see umma_bf16_elect)
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      constexpr uint32_t a_hi = umma_desc_hi(10u * C::ROWB, C::LAYOUT);  // 8-row groups of a halo view are 10 rows apart
      constexpr uint32_t b_hi = umma_desc_hi(8u * C::ROWB, C::LAYOUT);
      if (!p.w_per_img) {
        mbar_wait_warp(wfull_bar, 0);
        tc_fence_after();
      }
      const uint32_t w_lo = umma_desc_lo(smem_u32(wsm), 16u);
      const uint32_t w_step = (uint32_t)p.n_chunks * (uint32_t)(C::WSLOT_BYTES >> 4);  // descriptor units between taps
      const uint32_t halo_lo = umma_desc_lo(smem_u32(halo), 16u);
      DbgClock dc(lane == 0 ? p.dbg : nullptr);
      const long long t_start = dc.now();
      int stage = 0;
      uint32_t phase = 0;
      int w_img = -1;
      uint32_t w_loads = 0;
      for (uint32_t titer = 0; it.get(p, titer, img, h0, w0, nt); ++titer) {
        if (p.w_per_img && img != w_img) {
          mbar_wait_warp(wfull_bar, w_loads & 1u);
          tc_fence_after();
          w_img = img;
          ++w_loads;
        }
        const uint32_t acc = titer & 1u;
        const uint32_t acc_phase = (titer >> 1) & 1u;
        const long long wa = dc.now();
        mbar_wait_warp(&tempty_bar[acc], acc_phase ^ 1u);
        dc.add(1, wa);
        tc_fence_after();
        const uint32_t d_tmem = acc * BN;  // TMEM base is 0 (checked after the allocation)
        for (int g = 0; g < p.n_chunks; ++g) {
          const long long wf = dc.now();
          mbar_wait_warp(&full_bar[stage], phase);
          dc.add(0, wf);
          tc_fence_after();
          const uint32_t h_lo = halo_lo + (uint32_t)stage * (uint32_t)(C::HALO_BYTES >> 4);
          uint32_t b_lo = w_lo + (uint32_t)g * (uint32_t)(C::WSLOT_BYTES >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t a_lo = h_lo + (uint32_t)(((t / 3) * 10 + (t % 3)) * C::ROWB >> 4);
#pragma unroll
            for (int k = 0; k < CK / 16; ++k)
              umma_bf16_elect(d_tmem, a_lo + 2 * k, a_hi, b_lo + 2 * k, b_hi, idesc, (g | t | k) != 0 ? 1u : 0u);
            b_lo += w_step;
          }
          umma_commit_elect(&empty_bar[stage]);
          if (++stage == p.halo_stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit_elect(&tfull_bar[acc]);
        if (p.w_per_img) {  // last tile of this image for this CTA: the producer may bring the next image's weights
          int img2, h2, w2, n2;
          if (it.get(p, titer + 1, img2, h2, w2, n2) && img2 != img) umma_commit_elect(wfree_bar);
        }
      }
      dc.add(2, t_start);
      dc.flush(1, 3);
    }
  } else {
    run_epilogue<BN, C::OCW, C::SUB, C::OUT_BYTES, FEAT>(p, it, out_stage, s_scale, s_shift, s_pool, s_sum, s_sq, tfull_bar,
                                                         tempty_bar, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

#ifdef PMOE_KERNELS_ONLY  // codegen experiments instantiate single kernels from another translation unit
}  // namespace pmoe
#else
// ------------------------------------------------------------------------------------------ host
static void choose_tile(int H, int W, int* bh_out, int* bw_out, bool prefer_rows = false) {
  // fewest 128-row tiles first, then the smallest halo (squarer patches re-use more of L2). prefer_rows: the epilogue also
  // writes fp32 NCHW straight from registers, one pixel per lane, so a 32-pixel-wide patch makes every warp store one
  // contiguous 128-byte row per channel instead of four 32-byte pieces.
  long long best_tiles = -1, best_halo = 0;
  int best_bw = 1, best_bh = 1;
  const int wmax = W < 128 ? W : 128;
  for (int bw = 1; bw <= wmax; ++bw) {
    int bh = 128 / bw;
    if (bh > H) bh = H;
    if (bh < 1) continue;
    const long long tiles = (long long)((W + bw - 1) / bw) * ((H + bh - 1) / bh);
    const long long halo = prefer_rows ? (bw == 32 ? 0 : 1000 + (bw > 32 ? bw - 32 : 32 - bw)) : (long long)(bw + 2) * (bh + 2);
    if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && halo < best_halo)) {
      best_tiles = tiles;
      best_halo = halo;
      best_bw = bw;
      best_bh = bh;
    }
  }
  *bh_out = best_bh;
  *bw_out = best_bw;
}

static int make_view_tmap(CUtensorMap* tm, const PmoeView4& v, int box_c, int bw, int bh, CUtensorMapSwizzle swz,
                          const char* what) {
  if (((uintptr_t)v.ptr & 15) || (v.sw % 8) || (v.sh % 8) || (v.sn % 8) || (v.c % 8)) {
    set_error("%s: view must be 16-byte aligned with strides/channels in multiples of 8 elements (ptr %p c %d sw %lld sh "
              "%lld sn %lld)",
              what, v.ptr, v.c, (long long)v.sw, (long long)v.sh, (long long)v.sn);
    return PMOE_ERR_ARG;
  }
  const uint64_t dims[4] = {(uint64_t)v.c, (uint64_t)v.w, (uint64_t)v.h, (uint64_t)v.n};
  const uint64_t strides[3] = {(uint64_t)v.sw * 2, (uint64_t)v.sh * 2, (uint64_t)v.sn * 2};
  const uint32_t box[4] = {(uint32_t)box_c, (uint32_t)bw, (uint32_t)bh, 1u};
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, v.ptr, dims, strides, box, swz);
}

static CUtensorMapSwizzle swizzle_for_bytes(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

template <int BN, int CK, int FEAT>
static int launch_tc_feat(const ConvTcParams& p, cudaStream_t stream) {
  using C = TcCfg<BN, CK>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, CK, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_tc<%d,%d>: cannot reserve %d bytes of shared memory: %s", BN, CK, C::SMEM_BYTES, cudaGetErrorString(e));
      return PMOE_ERR_LAUNCH;
    }
    configured = true;
  }
  long long grid = p.total_tiles < (long long)num_sms() ? p.total_tiles : (long long)num_sms();
  ConvTcParams q = p;
  q.out_bufs = C::OUT_BUFS;
  if (p.pair) {
    // clusters of two CTAs; total_tiles counts pair units here
    long long pairs = p.total_tiles < (long long)(num_sms() / 2) ? p.total_tiles : (long long)(num_sms() / 2);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
    cfg.blockDim = dim3(kNumThreads, 1, 1);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, CK, FEAT>, q);
    if (e != cudaSuccess) {
      set_error("conv_tc<%d,%d>: cluster launch failed: %s", BN, CK, cudaGetErrorString(e));
      return PMOE_ERR_LAUNCH;
    }
    return check_launch("conv_tc");
  }
  conv_tc_kernel<BN, CK, FEAT><<<(unsigned)grid, kNumThreads, C::SMEM_BYTES, stream>>>(q);
  return check_launch("conv_tc");
}

// Specialised epilogues exist for the hot shapes (64-channel chunks, N tile >= 64); everything else is generic.
template <int BN, int CK>
static int launch_tc(const ConvTcParams& p, cudaStream_t stream) {
  if constexpr (CK == 64 && BN == 32) {  // the U-Net's 1x1 output conv: bias + per-image sums (+ fp32 NCHW copy)
    switch (p.feat) {
      case kFeatShift | kFeatPoolSum: return launch_tc_feat<BN, CK, kFeatShift | kFeatPoolSum>(p, stream);
      case kFeatShift | kFeatPoolSum | kFeatNchw: return launch_tc_feat<BN, CK, kFeatShift | kFeatPoolSum | kFeatNchw>(p, stream);
      default: break;
    }
  }
  if constexpr (CK == 64 && BN >= 64) {
    switch (p.feat) {
      case 0: return launch_tc_feat<BN, CK, 0>(p, stream);
      case kFeatShift: return launch_tc_feat<BN, CK, kFeatShift>(p, stream);
      case kFeatShift | kFeatRelu: return launch_tc_feat<BN, CK, kFeatShift | kFeatRelu>(p, stream);
      case kFeatShift | kFeatRelu | kFeatPool2: return launch_tc_feat<BN, CK, kFeatShift | kFeatRelu | kFeatPool2>(p, stream);
      case kFeatStats: return launch_tc_feat<BN, CK, kFeatStats>(p, stream);
      case kFeatShift | kFeatElu: return launch_tc_feat<BN, CK, kFeatShift | kFeatElu>(p, stream);  // expert-head ELU layers
      case kFeatRes: return launch_tc_feat<BN, CK, kFeatRes>(p, stream);  // data gradient accumulated into an existing buffer
      case kFeatShift | kFeatRelu | kFeatRes: return launch_tc_feat<BN, CK, kFeatShift | kFeatRelu | kFeatRes>(p, stream);
      default: break;
    }
  }
  if constexpr (CK == 16 && BN >= 64) {  // K = 16: data gradients of the 512->1 / 512->4 head Linears (one MMA per tile: pure epilogue)
    switch (p.feat) {
      case 0: return launch_tc_feat<BN, CK, 0>(p, stream);
      case kFeatRes: return launch_tc_feat<BN, CK, kFeatRes>(p, stream);
      default: break;
    }
  }
  return launch_tc_feat<BN, CK, -1>(p, stream);
}

template <int BN>
static int launch_tc_ck(const ConvTcParams& p, int ck, cudaStream_t stream) {
  switch (ck) {
    case 64: return launch_tc<BN, 64>(p, stream);
    case 32: return launch_tc<BN, 32>(p, stream);
    case 16: return launch_tc<BN, 16>(p, stream);
  }
  set_error("conv_tc: ck must be 16, 32 or 64 (got %d)", ck);
  return PMOE_ERR_ARG;
}


template <int BN, int CK, int FEAT>
static int launch_halo_feat(const ConvTcParams& p, int smem_bytes, cudaStream_t stream) {
  static int configured = 0;
  if (configured < smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel<BN, CK, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) {
      set_error("conv_tc_halo<%d>: cannot reserve %d bytes of shared memory: %s", BN, smem_bytes, cudaGetErrorString(e));
      return PMOE_ERR_LAUNCH;
    }
    configured = smem_bytes;
  }
  long long per_n = num_sms() / p.tiles_n;
  if (per_n > p.m_tiles) per_n = p.m_tiles;
  if (per_n < 1) per_n = 1;
  const unsigned grid = (unsigned)(per_n * p.tiles_n);
  conv_tc_halo_kernel<BN, CK, FEAT><<<grid, kNumThreads, smem_bytes, stream>>>(p);
  return check_launch("conv_tc_halo");
}

template <int BN, int CK>
static int launch_halo(const ConvTcParams& p, int smem_bytes, cudaStream_t stream) {
  if constexpr (BN >= 64 && (CK == 64 || CK == 16)) {
    switch (p.feat) {
      case 0: return launch_halo_feat<BN, CK, 0>(p, smem_bytes, stream);
      case kFeatShift | kFeatRelu: return launch_halo_feat<BN, CK, kFeatShift | kFeatRelu>(p, smem_bytes, stream);
      case kFeatShift | kFeatRelu | kFeatPool2: return launch_halo_feat<BN, CK, kFeatShift | kFeatRelu | kFeatPool2>(p, smem_bytes, stream);
      case kFeatShift | kFeatRelu | kFeatPoolSum: return launch_halo_feat<BN, CK, kFeatShift | kFeatRelu | kFeatPoolSum>(p, smem_bytes, stream);
      case kFeatStats: return launch_halo_feat<BN, CK, kFeatStats>(p, smem_bytes, stream);
      case kFeatRes: return launch_halo_feat<BN, CK, kFeatRes>(p, smem_bytes, stream);  // dgrad into a fan-out point (ResNet skip)
      case kFeatShift | kFeatRelu | kFeatRes: return launch_halo_feat<BN, CK, kFeatShift | kFeatRelu | kFeatRes>(p, smem_bytes, stream);
      default: break;
    }
  }
  return launch_halo_feat<BN, CK, -1>(p, smem_bytes, stream);
}

template <int BN>
static int launch_halo_ck(const ConvTcParams& p, int ck, int smem_bytes, cudaStream_t stream) {
  switch (ck) {
    case 64: return launch_halo<BN, 64>(p, smem_bytes, stream);
    case 32: return launch_halo<BN, 32>(p, smem_bytes, stream);
    default: return launch_halo<BN, 16>(p, smem_bytes, stream);
  }
}

// 3x3 / stride 1 / pad 1 over whole sources in the canonical (tap, source) segment order?
static bool is_canonical_3x3(const PmoeConvTc* d) {
  if ((d->ck != 64 && d->ck != 32 && d->ck != 16) || d->n_seg != 9 * d->n_src) return false;
  for (int t = 0; t < 9; ++t)
    for (int i = 0; i < d->n_src; ++i) {
      const PmoeSeg& s = d->seg[t * d->n_src + i];
      if (s.src != i || s.dh != t / 3 - 1 || s.dw != t % 3 - 1 || s.c0 != 0 || s.nchunks * d->ck != d->src[i].c) return false;
    }
  return true;
}
}  // namespace pmoe

using namespace pmoe;

static unsigned long long* g_conv_dbg = nullptr;
extern "C" int pmoe_conv_tc_set_debug(unsigned long long* counters_dev) {
  g_conv_dbg = counters_dev;  // NULL switches the accounting off; otherwise >= 16 * #CTAs counters, zeroed by the caller
  return PMOE_OK;
}

extern "C" int pmoe_conv_tc(const PmoeConvTc* d, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!d || d->n_src < 1 || d->n_src > PMOE_MAX_SRC || d->n_seg < 1 || d->n_seg > PMOE_MAX_SEG || !d->wpack || !d->out.ptr) {
    set_error("conv_tc: bad descriptor (n_src %d n_seg %d)", d ? d->n_src : -1, d ? d->n_seg : -1);
    return PMOE_ERR_ARG;
  }
  if (d->ck != 16 && d->ck != 32 && d->ck != 64) {
    set_error("conv_tc: ck must be 16, 32 or 64 (got %d)", d->ck);
    return PMOE_ERR_ARG;
  }
  int bn = 0;
  for (int cand : {256, 128, 64, 32, 16})
    if (d->cout_pad % cand == 0) {
      bn = cand;
      break;
    }
  if (bn == 0 || d->cout_pad <= 0) {
    set_error("conv_tc: cout_pad (%d) must be a positive multiple of 16", d->cout_pad);
    return PMOE_ERR_ARG;
  }
  if ((d->stat_sum != nullptr) != (d->stat_sqsum != nullptr) || (d->stat_sum && d->cout_pad > kMaxStatC)) {
    set_error("conv_tc: batch statistics need both buffers and cout_pad <= %d", kMaxStatC);
    return PMOE_ERR_ARG;
  }
  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  const PmoeView4& o = d->out;
  // ---- resident-weight / halo variant when it applies
  int halo_bn = 0, total_chunks = 0;
  static const bool halo_off = getenv("PMOE_NO_HALO") != nullptr;
  if (!halo_off && is_canonical_3x3(d) && o.h >= 18 && o.w >= 10) {
    for (int i = 0; i < d->n_src; ++i) total_chunks += d->src[i].c / d->ck;
    if (total_chunks <= PMOE_MAX_SEG) {
      for (int cand : {128, 64, 32, 16})
        if (d->cout_pad % cand == 0 && 9 * total_chunks * cand * d->ck * 2 <= 144 * 1024) {
          halo_bn = cand;
          break;
        }
      if (halo_bn < 64 && d->cout_pad >= 64) halo_bn = 0;  // would starve the MMA: stream the taps instead
      static const bool wstream_off = getenv("PMOE_NO_WSTREAM") != nullptr;
      // An N = 64 tile is bound by shared-memory operand reads (A 4 KB + B 2 KB per K=16 step: ~60 clk measured against 32 on
      // the tensor pipe), so when cout allows N >= 128 the weights are streamed through a ring rather than kept resident at
      // N = 64 (measured: 112x112 256->128 at N = 128 streams at 1.4 PFLOP/s). Per-image weights need the resident variant.
      if (halo_bn < 128 && !wstream_off && d->ck == 64 && d->cout_pad % 128 == 0 && d->wpack_img_stride <= 0) {
        halo_bn = d->cout_pad % 256 == 0 ? 256 : 128;  // halo input tiles + weight tiles streamed through a ring
        p.stream_w = 1;
      }
    }
  }
  if (halo_bn) {
    bn = halo_bn;
    p.bh = 16;
    p.bw = 8;
  } else if (d->pool2_out.ptr) {
    p.bw = 8;  // the fused max-pool finds its window partners with lane shuffles: power-of-two tile width
    p.bh = 16;
  } else {
    choose_tile(o.h, o.w, &p.bh, &p.bw, d->nchw_out != nullptr);
  }
  p.H = o.h;
  p.W = o.w;
  p.n_img = o.n;
  p.tiles_w = (o.w + p.bw - 1) / p.bw;
  p.tiles_h = (o.h + p.bh - 1) / p.bh;
  p.tiles_n = d->cout_pad / bn;
  p.total_tiles = (long long)p.tiles_w * p.tiles_h * p.n_img * p.tiles_n;
  if (p.total_tiles <= 0) {
    set_error("conv_tc: empty output");
    return PMOE_ERR_ARG;
  }
  if (p.total_tiles > 0x3fffffffLL) {  // the device-side tile arithmetic is 32-bit
    set_error("conv_tc: too many tiles");
    return PMOE_ERR_ARG;
  }
  int kiters = 0;
  for (int i = 0; i < d->n_seg; ++i) {
    const PmoeSeg& s = d->seg[i];
    if (s.src < 0 || s.src >= d->n_src || s.nchunks == 0 || s.c0 + s.nchunks * d->ck > d->src[s.src].c) {
      set_error("conv_tc: segment %d out of range (src %d c0 %d nchunks %d ck %d, source has %d channels)", i, s.src, s.c0,
                s.nchunks, d->ck, (s.src >= 0 && s.src < d->n_src) ? d->src[s.src].c : -1);
      return PMOE_ERR_ARG;
    }
    p.seg[i].src = s.src;
    p.seg[i].dh = s.dh;
    p.seg[i].dw = s.dw;
    p.seg[i].c0 = s.c0;
    p.seg[i].nchunks = s.nchunks;
    kiters += s.nchunks;
  }
  if (kiters * d->ck != d->ktot) {
    set_error("conv_tc: ktot %d does not match the segment list (%d chunks of %d)", d->ktot, kiters, d->ck);
    return PMOE_ERR_ARG;
  }
  p.n_seg = d->n_seg;
  p.kiters = kiters;
  const CUtensorMapSwizzle swz_in = swizzle_for_bytes(d->ck * 2);
  int rc;
  for (int i = 0; i < d->n_src; ++i) {
    if ((rc = make_view_tmap(&p.tm_src[i], d->src[i], d->ck, halo_bn ? 10 : p.bw, halo_bn ? 18 : p.bh, swz_in,
                             "conv_tc source")) != PMOE_OK)
      return rc;
  }
  {
    if ((uintptr_t)d->wpack & 15) {
      set_error("conv_tc: wpack must be 16-byte aligned");
      return PMOE_ERR_ARG;
    }
    if (d->wpack_img_stride > 0) {
      // one weight set per image (ECA gate folded in): only the resident halo kernel reloads weights per image
      if (p.stream_w || d->wpack_img_stride < (int64_t)d->cout_pad * d->ktot || (d->wpack_img_stride % 8)) {
        set_error("conv_tc: per-image weights are not supported by the streamed-weights 3x3 kernel (or bad stride)");
        return PMOE_ERR_UNSUPPORTED;
      }
      const uint64_t dims[3] = {(uint64_t)d->ktot, (uint64_t)d->cout_pad, (uint64_t)o.n};
      const uint64_t strides[2] = {(uint64_t)d->ktot * 2, (uint64_t)d->wpack_img_stride * 2};
      const uint32_t box[3] = {(uint32_t)d->ck, (uint32_t)bn, 1u};
      if ((rc = encode_tmap(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d->wpack), dims, strides, box,
                            swz_in)) != PMOE_OK)
        return rc;
      p.w_per_img = 1;
    } else {
      const uint64_t dims[2] = {(uint64_t)d->ktot, (uint64_t)d->cout_pad};
      const uint64_t strides[1] = {(uint64_t)d->ktot * 2};
      static const bool pair_off = getenv("PMOE_PAIR") == nullptr;
      // measured: the multicast halves the L2 reads but not the bytes each SM takes in, which is what bounds these layers
      // (the halo + streamed-weights kernel fixes that instead) -> opt-in only
      const long long m_tiles_all = (long long)p.tiles_w * p.tiles_h * p.n_img;
      p.pair = (!pair_off && !halo_bn && d->ck == 64 && (bn == 128 || bn == 256) && kiters >= 9 && m_tiles_all >= 4 &&
                num_sms() % 2 == 0) ? 1 : 0;
      if (p.pair) p.total_tiles = ((m_tiles_all + 1) / 2) * p.tiles_n;
      const uint32_t box[2] = {(uint32_t)d->ck, (uint32_t)(p.pair ? bn / 2 : bn)};
      if ((rc = encode_tmap(&p.tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(d->wpack), dims, strides, box,
                            swz_in)) != PMOE_OK)
        return rc;
    }
  }
  auto view_ok = [&](const PmoeView4& v, int esz) {
    return v.ptr && !((uintptr_t)v.ptr & 15) && (v.sw * esz) % 16 == 0 && (v.sh * esz) % 16 == 0 && (v.sn * esz) % 16 == 0;
  };
  auto view_a32 = [&](const PmoeView4& v) { return !((uintptr_t)v.ptr & 31) && v.sw % 16 == 0 && v.sh % 16 == 0 && v.sn % 16 == 0; };
  const int ocw = bn < 64 ? bn : 64;
  if ((rc = make_view_tmap(&p.tm_out[0], o, ocw, p.bw, p.bh, swizzle_for_bytes(ocw * 2), "conv_tc output")) != PMOE_OK) return rc;
  p.out_c = o.c;
  p.out_bufs = 2;
  if (d->n_out_extra > 0) {
    // multi-view launch: GEMM column n lands in view n / out_cols at channel n % out_cols
    if (d->n_out_extra > 3 || d->out_cols <= 0 || d->out_cols % ocw != 0 || d->cout_pad != d->out_cols * (d->n_out_extra + 1) ||
        d->residual.ptr || d->stat_sum || d->pool_sum || d->pool2_out.ptr || d->nchw_out || halo_bn) {
      set_error("conv_tc: bad multi-view output (n_out_extra %d out_cols %d cout_pad %d)", d->n_out_extra, d->out_cols, d->cout_pad);
      return PMOE_ERR_ARG;
    }
    for (int i = 0; i < d->n_out_extra; ++i) {
      const PmoeView4& e = d->out_extra[i];
      if (e.n != o.n || e.h != o.h || e.w != o.w || e.c != o.c) {
        set_error("conv_tc: extra output view %d does not match out", i);
        return PMOE_ERR_ARG;
      }
      if ((rc = make_view_tmap(&p.tm_out[i + 1], e, ocw, p.bw, p.bh, swizzle_for_bytes(ocw * 2), "conv_tc extra output")) != PMOE_OK) return rc;
    }
    p.out_cols = d->out_cols;
  }
  if (d->pool2_out.ptr) {
    const PmoeView4& q = d->pool2_out;
    const bool pow2 = p.bw == 2 || p.bw == 4 || p.bw == 8 || p.bw == 16;
    if (!pow2 || (p.bh & 1) || (o.h & 1) || (o.w & 1) || q.n != o.n || q.h != o.h / 2 || q.w != o.w / 2 || q.c < o.c || !view_ok(q, 2)) {
      set_error("conv_tc: fused 2x2 max-pool needs even H/W, a (n, H/2, W/2, >=c) view and a power-of-two tile width (bw %d bh %d)", p.bw, p.bh);
      return PMOE_ERR_UNSUPPORTED;
    }
    p.pool2_ptr = static_cast<__nv_bfloat16*>(q.ptr);
    p.pool2_sn = q.sn;
    p.pool2_sh = q.sh;
    p.pool2_sw = q.sw;
    p.pool2_align32 = view_a32(q) && o.c % 16 == 0;
    (void)view_ok;
  }
  if (d->nchw_out) {
    p.nchw_ptr = d->nchw_out;
    p.nchw_sn = d->nchw_sn;
    p.nchw_sc = d->nchw_sc;
    p.nchw_sh = d->nchw_sh;
    p.nchw_sw = d->nchw_sw;
    p.nchw_c = d->nchw_c;
  }
  p.dbg = g_conv_dbg;
  p.shift_img_stride = d->shift_img_stride;
  p.scale = d->scale;
  p.shift = d->shift;
  p.act = d->act;
  if (d->residual.ptr) {
    const PmoeView4& r = d->residual;
    if (r.n != o.n || r.h != o.h || r.w != o.w || ((uintptr_t)r.ptr & 15) || (r.sw % 8) || (r.sh % 8) || (r.sn % 8)) {
      set_error("conv_tc: residual view must match the output geometry and be 16-byte aligned");
      return PMOE_ERR_ARG;
    }
    p.res = static_cast<const __nv_bfloat16*>(r.ptr);
    p.res_sn = r.sn;
    p.res_sh = r.sh;
    p.res_sw = r.sw;
    p.res_c = r.c;
  }
  p.stat_sum = d->stat_sum;
  p.stat_sq = d->stat_sqsum;
  p.pool_sum = d->pool_sum;
  p.cout_pad = d->cout_pad;
  p.pool_stride = d->pool_stride > 0 ? d->pool_stride : d->cout_pad;
  p.feat = -1;
  if (!d->scale && (d->act == PMOE_ACT_NONE || d->act == PMOE_ACT_RELU || d->act == PMOE_ACT_ELU))
    p.feat = (d->shift ? kFeatShift : 0) | (d->act == PMOE_ACT_RELU ? kFeatRelu : 0) | (d->act == PMOE_ACT_ELU ? kFeatElu : 0) | (d->stat_sum ? kFeatStats : 0) |
             (d->pool2_out.ptr ? kFeatPool2 : 0) | (d->pool_sum ? kFeatPoolSum : 0) | (d->nchw_out ? kFeatNchw : 0) |
             (d->residual.ptr ? kFeatRes : 0);
  if (halo_bn) {
    int g = 0;
    for (int i = 0; i < d->n_src; ++i)
      for (int c = 0; c < d->src[i].c / d->ck; ++c, ++g) {
        p.seg[g].src = (int8_t)i;
        p.seg[g].dh = 0;
        p.seg[g].dw = 0;
        p.seg[g].c0 = (uint16_t)(c * d->ck);
        p.seg[g].nchunks = 1;
      }
    p.n_chunks = total_chunks;
    p.w_slots = 9 * total_chunks;
    p.m_tiles = (long long)p.tiles_w * p.tiles_h * p.n_img;
    const int halo_bytes = ((18 * 10 * d->ck * 2 + 1023) / 1024) * 1024;
    const int ocw_h = halo_bn < 64 ? halo_bn : 64;
    const int bar_bytes = (2 * 6 + 6 + 2 * kMaxWStages) * 8 + 16;
    int wbytes = ((p.w_slots * halo_bn * d->ck * 2 + 1023) / 1024) * 1024;  // halo stages stay 1 KB aligned
    int fixed = 0, stages = 0;
    if (p.stream_w) {
      // three halo stages, one (N = 256) or two staging tiles, the rest of shared memory for the weight ring
      p.out_bufs = halo_bn >= 256 ? 1 : 2;
      fixed = 1024 + p.out_bufs * 128 * ocw_h * 2 + (3 * halo_bn + 2 * kMaxStatC) * 4 + bar_bytes;
      stages = 3;
      int ws = (227 * 1024 - fixed - stages * halo_bytes) / (halo_bn * 128);
      if (ws > kMaxWStages) ws = kMaxWStages;
      if (ws < 3) {
        set_error("conv_tc: internal error: no room for the weight ring");
        return PMOE_ERR_ARG;
      }
      p.w_slots = ws;
      wbytes = ws * halo_bn * 128;
    } else {
      for (p.out_bufs = 2; p.out_bufs >= 1; --p.out_bufs) {
        fixed = 1024 + p.out_bufs * 128 * ocw_h * 2 + (3 * halo_bn + 2 * kMaxStatC) * 4 + bar_bytes;
        stages = (227 * 1024 - fixed - wbytes) / halo_bytes;
        if (stages >= 3 || p.out_bufs == 1) break;
      }
    }
    if (stages > 6) stages = 6;
    if (stages >= 2) {
      p.halo_stages = stages;
      p.w_bytes = wbytes;
      const int smem_bytes = fixed + wbytes + stages * halo_bytes;
      switch (halo_bn) {
        case 256: return launch_halo_ck<256>(p, d->ck, smem_bytes, stream);
        case 128: return launch_halo_ck<128>(p, d->ck, smem_bytes, stream);
        case 64: return launch_halo_ck<64>(p, d->ck, smem_bytes, stream);
        case 32: return launch_halo_ck<32>(p, d->ck, smem_bytes, stream);
        default: return launch_halo_ck<16>(p, d->ck, smem_bytes, stream);
      }
    }
    set_error("conv_tc: internal error: halo variant selected without room for two stages");
    return PMOE_ERR_ARG;
  }
  p.tiles_n_varies = p.tiles_n > 1 ? 1 : 0;
  switch (bn) {
    case 256: return launch_tc_ck<256>(p, d->ck, stream);
    case 128: return launch_tc_ck<128>(p, d->ck, stream);
    case 64: return launch_tc_ck<64>(p, d->ck, stream);
    case 32: return launch_tc_ck<32>(p, d->ck, stream);
    default: return launch_tc_ck<16>(p, d->ck, stream);
  }
}

// ------------------------------------------------------------------------------------------ probe
namespace pmoe {
struct alignas(64) DbgParams {
  CUtensorMap tm_a, tm_b;
  int rows, start_row, group_rows, bo_mode;
  float* d_out;
};

__global__ void __launch_bounds__(128, 1) dbg_umma_view_kernel(const __grid_constant__ DbgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                // up to 384 rows * 128 B
  uint8_t* sb = smem + 384 * 128;    // 64 rows * 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(s_tmem, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], (uint32_t)(p.rows * 128 + 64 * 128));
    // A is loaded in slabs of <=128 rows (box height limit is 256; keep it simple)
    for (int r = 0; r < p.rows; r += 128) tma_load_2d(sa + r * 128, &p.tm_a, &bars[0], 0, r);
    tma_load_2d(sb, &p.tm_b, &bars[0], 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sa) + (uint32_t)p.start_row * 128u;
    const uint32_t b_addr = smem_u32(sb);
    const uint32_t bo = p.bo_mode == 1 ? ((a_addr >> 7) & 7u) : 0u;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) {
      const uint64_t adesc = umma_desc_kmajor(a_addr + k * 32, (uint32_t)p.group_rows * 128u, kLayoutSW128, bo);
      const uint64_t bdesc = umma_desc_kmajor(b_addr + k * 32, 1024u, kLayoutSW128, 0);
      umma_bf16(tmem_base, adesc, bdesc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  for (int sbk = 0; sbk < 2; ++sbk) {
    uint32_t raw[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + sbk * 32, raw);
    tmem_ld_wait();
    float* o = p.d_out + (warp * 32 + lane) * 64 + sbk * 32;
    for (int k = 0; k < 32; ++k) o[k] = __uint_as_float(raw[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}
}  // namespace pmoe

extern "C" int pmoe_dbg_umma_view(const void* a_bf16, int rows, const void* b_bf16, float* d_out, int start_row,
                                  int group_rows, int base_offset_mode, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (rows < 128 || rows > 384 || rows % 128 != 0) {
    set_error("dbg_umma_view: rows must be 128, 256 or 384");
    return PMOE_ERR_ARG;
  }
  if (start_row + 15 * group_rows + 8 > rows) {
    set_error("dbg_umma_view: view exceeds the loaded tile");
    return PMOE_ERR_ARG;
  }
  DbgParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    const uint64_t dims[2] = {64, (uint64_t)rows};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 128};
    if ((rc = encode_tmap(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a_bf16), dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != PMOE_OK)
      return rc;
  }
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t strides[1] = {128};
    const uint32_t box[2] = {64, 64};
    if ((rc = encode_tmap(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(b_bf16), dims, strides, box,
                          CU_TENSOR_MAP_SWIZZLE_128B)) != PMOE_OK)
      return rc;
  }
  p.rows = rows;
  p.start_row = start_row;
  p.group_rows = group_rows;
  p.bo_mode = base_offset_mode;
  p.d_out = d_out;
  const int smem_bytes = 1024 + 384 * 128 + 64 * 128 + 64;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(dbg_umma_view_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    configured = true;
  }
  dbg_umma_view_kernel<<<1, 128, smem_bytes, stream>>>(p);
  return check_launch("dbg_umma_view");
}
#endif  // PMOE_KERNELS_ONLY
