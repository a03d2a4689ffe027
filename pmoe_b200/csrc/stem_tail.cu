// The tail of the ResNet stem in training: torchvision's bn1 -> ReLU -> MaxPool2d(3, 2, 1) directly behind the stem block
// (model/blocks/backbone.py:57-61 keeps torchvision's `bn1`, `relu`, `maxpool` after `conv1 := EfficientConvBlock`). At 224^2 x 64
// channels these are the largest memory-bound tensors of an expert (1.6 GB at 256 samples), and as separate launches the chain
// reads / writes them 5 + 6.4 times (apply, pool | pool backward, BatchNorm reduce, BatchNorm apply). Here:
//
//  forward   p = maxpool(relu(scale * x + shift)) WITHOUT materialising the normalised tensor: relu(scale*x + shift) is monotone in
//            x, so the window maximum is taken on x itself (sign-flipped for channels with scale < 0) and the affine + ReLU is
//            applied once per OUTPUT. Stores p, the argmax codes and x at the argmax (exact bf16 bits, 1/4 of a tensor).
//  reduce    the two BatchNorm-backward sums need dz = route(dp) * mask only where dz != 0, i.e. at the argmax positions: a pass
//            over the POOLED grid (dp and x-at-argmax: half a tensor instead of three).
//  apply     dx = gamma * rstd * (dz*m - c1 - xhat * c2) for every input pixel with dz gathered from the (<= 4) windows that
//            hold the pixel (the pair form of maxpool3s2_bwd_pair_kernel) — the pooled gradient is never scattered to a full
//            tensor — plus the backward sums of the UPSTREAM BatchNorm (x is its ReLU output) and the affine gradients.
//
// Dense bf16 NHWC, even H and W, 256 % (C/8) == 0. Ties between equal inputs go to the first window element like ATen's; ties
// between equal OUTPUTS of different inputs go to the larger input (the fp32 reference has no such ties; windows whose maximum is
// <= 0 pass no gradient whichever position is recorded; a channel whose gamma is exactly 0 is constant, and only its own d gamma
// depends on the position).
#include <cuda_bf16.h>

#include <cstdlib>

#include "host_util.h"
#include "reduce.cuh"

namespace pmoe {

__device__ __forceinline__ void st_bf16x8_to_f32(const uint4& r, float (&v)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    v[2 * q] = __uint_as_float(w[q] << 16);
    v[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t st_pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint4 st_f32_to_bf16x8(const float (&v)[8]) {
  return make_uint4(st_pack2(v[0], v[1]), st_pack2(v[2], v[3]), st_pack2(v[4], v[5]), st_pack2(v[6], v[7]));
}
__device__ __forceinline__ void st_load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// ---------------------------------------------------------------------------------------------------------------- forward
// Two bf16 channels per 32-bit lane all the way: `set.gt.bf16x2` gives a 0xffff / 0 mask per channel, and the running maximum
// and its 16-bit argmax code are merged with one LOP3 each (3 instructions per tap and channel PAIR instead of ~10 on unpacked
// fp32 values: the unpacked form was ALU-pipe bound at 3.0 TB/s, ncu 65 % alu / 37 % DRAM).
__device__ __forceinline__ uint32_t st_gt2_mask(uint32_t a, uint32_t b) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) bn_relu_maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint2* __restrict__ idx,
                                                                  uint4* __restrict__ xmax, int H, int W, int OH, int OW, int cg,
                                                                  int cg_shift, const float* __restrict__ scale, const float* __restrict__ shift) {
  const int n = blockIdx.x / OH, oh = blockIdx.x - n * OH;
  const int row_items = OW * cg;
  const uint4* img = x + (size_t)n * H * W * cg;
  const int g = threadIdx.x & (cg - 1);   // cg divides 256 (a power of two) and the item stride is a multiple of 256: one channel group per thread
  float sc[8], sh[8];
  st_load8(scale + g * 8, sc);
  st_load8(shift + g * 8, sh);
  uint32_t flip[4];                  // sign-bit mask of the channels whose scale is negative (their window MINIMUM wins)
#pragma unroll
  for (int q = 0; q < 4; ++q) flip[q] = (sc[2 * q] < 0.f ? 0x00008000u : 0u) | (sc[2 * q + 1] < 0.f ? 0x80000000u : 0u);
  // (grid.y = 256-item chunk of the output row: blocks that run together hold the same columns of neighbouring rows and share
  // their input rows through L1 / L2; a 1-D grid with the chunks of a row adjacent measured 12 % slower)
  for (int item = blockIdx.y * blockDim.x + threadIdx.x; item < row_items; item += gridDim.y * blockDim.x) {
    const int ow = item >> cg_shift;
    // even H and W: only the top row (oh == 0, r == 0) and the left column (ow == 0, c == 0) of a window can fall outside
    const uint4* base = img + ((long long)(oh * 2 - 1) * W + (ow * 2 - 1)) * cg + g;
    uint32_t m[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u};   // -inf
    uint32_t arg[4] = {0u, 0u, 0u, 0u};                                      // 16-bit code per channel
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      if (r == 0 && oh == 0) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (c == 0 && ow == 0) continue;
        const uint4 raw = __ldg(base + ((long long)r * W + c) * cg);
        const uint32_t v[4] = {raw.x ^ flip[0], raw.y ^ flip[1], raw.z ^ flip[2], raw.w ^ flip[3]};
        const uint32_t code2 = (uint32_t)(r * 3 + c) * 0x00010001u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t gt = st_gt2_mask(v[q], m[q]);   // strictly greater: the first maximum in row-major window order wins
          m[q] = (v[q] & gt) | (m[q] & ~gt);
          arg[q] = (code2 & gt) | (arg[q] & ~gt);
        }
      }
    }
    const uint4 xm = make_uint4(m[0] ^ flip[0], m[1] ^ flip[1], m[2] ^ flip[2], m[3] ^ flip[3]);   // exact input bits
    float xv[8], o[8];
    st_bf16x8_to_f32(xm, xv);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = fmaxf(fmaf(xv[q], sc[q], sh[q]), 0.f);
    const size_t off = ((size_t)blockIdx.x * OW + ow) * cg + g;
    y[off] = st_f32_to_bf16x8(o);
    xmax[off] = xm;
    idx[off] = make_uint2(__byte_perm(arg[0], arg[1], 0x6420u), __byte_perm(arg[2], arg[3], 0x6420u));
  }
}

// ---------------------------------------------------------------------------------------------------------------- reduce
__global__ void __launch_bounds__(kRedThreads) bn_relu_maxpool_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ xmax,
                                                                                 long long npix, int cg, const float* __restrict__ fsc,
                                                                                 const float* __restrict__ fsh, const float* __restrict__ mean,
                                                                                 const float* __restrict__ rstd, double* __restrict__ sum_dy,
                                                                                 double* __restrict__ sum_dy_xhat, long long pix_per_block,
                                                                                 double* __restrict__ sum_dy_xpos) {
  __shared__ float sm[kRedThreads * 8];
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x % cg, lane = threadIdx.x / cg;
  float sc[8], sh[8];
  st_load8(fsc + g * 8, sc);
  st_load8(fsh + g * 8, sh);
  const long long p0 = (long long)blockIdx.x * pix_per_block;
  long long p1 = p0 + pix_per_block;
  if (p1 > npix) p1 = npix;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0}, e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (lane < lanes) {
    for (long long p = p0 + lane; p < p1; p += lanes) {
      float d[8], xv[8];
      st_bf16x8_to_f32(__ldg(dy + p * cg + g), d);
      st_bf16x8_to_f32(__ldg(xmax + p * cg + g), xv);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float dm = fmaf(xv[q], sc[q], sh[q]) > 0.f ? d[q] : 0.f;
        a[q] += dm;
        b[q] = fmaf(dm, xv[q], b[q]);
        e[q] += xv[q] > 0.f ? dm : 0.f;   // the part routed to positions where the input itself is > 0 (sum_dy_xpos)
      }
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter], te[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  block_channel_sum(b, sm, cg, lanes, tb);
  if (sum_dy_xpos) block_channel_sum(e, sm, cg, lanes, te);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      atomicAdd(sum_dy + c, (double)ta[j]);
      atomicAdd(sum_dy_xhat + c, (double)__ldg(rstd + c) * ((double)tb[j] - (double)__ldg(mean + c) * (double)ta[j]));
      if (sum_dy_xpos) atomicAdd(sum_dy_xpos + c, (double)te[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- apply
struct StParamGrads {
  float* dgamma;
  float* dbeta;
  int n;
  int accumulate;
};

// volatile shared-memory read of 8 floats: per-channel constants are re-read where they are used instead of being hoisted out
// of the item loop into registers.
// Layout of one per-channel array of C = cg*8 floats: [half][g][4], so that the 8 threads of a quarter warp (g = 0..7) read 8
// CONSECUTIVE 16-byte chunks per instruction. The natural [g][8] layout puts g and g + 4 on the same banks: ncu showed 8 shared
// wavefronts per LDS.128 and the L1 pipe at 90 %.
__device__ __forceinline__ int st_cst_index(int c, int cg) { return ((((c >> 2) & 1) * cg + (c >> 3)) << 2) + (c & 3); }
__device__ __forceinline__ void st_lds8(uint32_t addr, uint32_t half_stride, float (&v)[8]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(addr + half_stride));
}
// The ReLU mask of the fused layer, fmaf(x, scale, shift) > 0, as a compare of the bf16 input itself: with x' = x for scale > 0
// and -x for scale < 0 the predicate fmaf(x', |scale|, shift) > 0 is monotone in x' (rounding is monotone), so it equals
// x' > T for the largest bf16 T at which it is still false. T is found by bisection over the ordered bf16 bit patterns between
// -inf and +inf (16 evaluations of the very expression the forward kernel used), once per channel and block: the mask is
// bit-identical to the forward's, and the item loop needs neither scale / shift nor the 3 instructions per element.
__device__ __forceinline__ float st_key_to_float(uint32_t k) {   // ordered key -> bf16 value (0x007f = -inf ... 0xff80 = +inf)
  const uint32_t bits = (k & 0x8000u) ? (k ^ 0x8000u) : (~k & 0xffffu);
  return __uint_as_float(bits << 16);
}
__device__ __forceinline__ uint32_t st_relu_threshold(float scale, float shift) {
  if (scale == 0.f) return shift > 0.f ? 0xff80u : 0x7f80u;   // constant channel: every (finite) input passes / none does
  const float s = fabsf(scale);
  uint32_t lo = 0x007fu, hi = 0xff80u;                        // predicate false at -inf, true at +inf
  while (hi - lo > 1u) {
    const uint32_t mid = (lo + hi) >> 1;
    if (fmaf(st_key_to_float(mid), s, shift) > 0.f) hi = mid; else lo = mid;
  }
  return (lo & 0x8000u) ? (lo ^ 0x8000u) : (~lo & 0xffffu);
}
// prmt.b32 with the selector's sign-replicate bit (bit 3 of a nibble: the selected byte's msb fills the target byte);
// __byte_perm() masks the selector to 3 bits per nibble, so the instruction is written out.
__device__ __forceinline__ uint32_t st_prmt(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
// Pooled gradient of the windows whose recorded argmax is `want`, two bf16 channels per lane. Codes are < 16, so per byte
// 0x80 - (code ^ want) keeps bit 7 only where they are equal (no borrow crosses a byte); PRMT's sign-replicate mode widens
// that bit to the 16-bit lane of the channel. The sum over the (<= 4) windows that hold a pixel is taken with bf16 adds: one or
// two terms round exactly like the separate pool-backward launch (fp32 sum stored as bf16); three or four terms (a pixel that is
// the maximum of every window around it) round once more.
__device__ __forceinline__ void st_add_masked(uint32_t (&o)[4], const uint2& code, const uint4& grad, uint32_t want) {
  const uint32_t w4 = want * 0x01010101u;
  const uint32_t f0 = 0x80808080u - (code.x ^ w4), f1 = 0x80808080u - (code.y ^ w4);
  const uint32_t gq[4] = {grad.x & st_prmt(f0, 0x9988u), grad.y & st_prmt(f0, 0xbbaau),
                          grad.z & st_prmt(f1, 0x9988u), grad.w & st_prmt(f1, 0xbbaau)};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&o[q]), *reinterpret_cast<const __nv_bfloat162*>(&gq[q]));
    o[q] = *reinterpret_cast<const uint32_t*>(&r);
  }
}

// block_channel_sum (reduce.cuh) for a block of BT threads: thread t holds the totals of channels t, t + BT, ... in out[].
template <int BT>
__device__ __forceinline__ void st_block_channel_sum(const float (&v)[8], float* sm, int cg, float (&out)[2048 / BT]) {
  float4* s4 = reinterpret_cast<float4*>(sm) + threadIdx.x * 2;
  s4[0] = make_float4(v[0], v[1], v[2], v[3]);
  s4[1] = make_float4(v[4], v[5], v[6], v[7]);
  __syncthreads();
  const int nch = cg * 8, lanes = BT / cg;
#pragma unroll
  for (int j = 0; j < 2048 / BT; ++j) {
    const int c = threadIdx.x + j * BT;
    float t = 0.f;
    if (c < nch)
      for (int l = 0; l < lanes; ++l) t += sm[l * nch + c];
    out[j] = t;
  }
  __syncthreads();
}

// BT = 128 threads wherever the channel-group count allows: 112 pixel pairs x 8 groups of a 224-wide 64-channel row are exactly 7
// rounds of 128 items (3.5 rounds of 256 left every eighth thread idle), and registers are allocated in finer steps.
template <bool NEXT, int BT>
__global__ void __launch_bounds__(BT, (NEXT ? 896 : 1024) / BT) bn_relu_maxpool_bwd_apply_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                                        const uint4* __restrict__ x, uint4* __restrict__ dx, int rows, int H,
                                                                        int W, int OH, int OW, int cg, int cg_shift, const float* __restrict__ fsc,
                                                                        const float* __restrict__ fsh, const float* __restrict__ mean,
                                                                        const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                        const double* __restrict__ sum_dy,
                                                                        const double* __restrict__ sum_dy_xhat, float inv_n,
                                                                        double* __restrict__ next_s1, double* __restrict__ next_s2,
                                                                        StParamGrads pg) {
  __shared__ float sm_next[NEXT ? BT * 8 : 8];
  extern __shared__ float st_cst[];   // [3][C]: dx = A*d + B*x + C per channel, re-read per item (24 registers less: one more block per SM)
  __shared__ uint32_t sm_thr[256 * 4], sm_flip[256 * 4];   // per channel PAIR: packed ReLU thresholds and sign flips (see st_relu_threshold)
  if (blockIdx.x == 0 && pg.n > 0) {   // the BatchNorm affine gradients are the two reductions themselves
    for (int c = threadIdx.x; c < pg.n; c += blockDim.x) {
      if (pg.dbeta) pg.dbeta[c] = (pg.accumulate ? pg.dbeta[c] : 0.f) + (float)sum_dy[c];
      if (pg.dgamma) pg.dgamma[c] = (pg.accumulate ? pg.dgamma[c] : 0.f) + (float)sum_dy_xhat[c];
    }
  }
  for (int c = threadIdx.x; c < cg * 8; c += blockDim.x) {  // dx = gamma*rstd*(d - c1 - (x - mean)*rstd*c2) = A*d + B*x + C
    const double r = (double)__ldg(rstd + c), m = (double)__ldg(mean + c), gm = gamma ? (double)__ldg(gamma + c) : 1.0;
    const double c1 = sum_dy[c] * (double)inv_n, c2 = sum_dy_xhat[c] * (double)inv_n;
    const int ci = st_cst_index(c, cg);
    st_cst[ci] = (float)(gm * r);
    st_cst[cg * 8 + ci] = (float)(-gm * r * r * c2);
    st_cst[cg * 16 + ci] = (float)(gm * r * (r * c2 * m - c1));
  }
  for (int c2 = threadIdx.x; c2 < cg * 4; c2 += blockDim.x) {
    const float s0 = __ldg(fsc + 2 * c2), s1 = __ldg(fsc + 2 * c2 + 1);
    sm_thr[c2] = st_relu_threshold(s0, __ldg(fsh + 2 * c2)) | (st_relu_threshold(s1, __ldg(fsh + 2 * c2 + 1)) << 16);
    sm_flip[c2] = (s0 < 0.f ? 0x00008000u : 0u) | (s1 < 0.f ? 0x80000000u : 0u);
  }
  __syncthreads();
  const int items = (W >> 1) * cg;
  const int g = threadIdx.x & (cg - 1);
  uint32_t thr[4], flip[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    thr[q] = sm_thr[g * 4 + q];
    flip[q] = sm_flip[g * 4 + q];
  }
  const uint32_t cst = (uint32_t)__cvta_generic_to_shared(st_cst) + (uint32_t)g * 16u;
  const uint32_t cst_stride = (uint32_t)cg * 32u, cst_half = (uint32_t)cg * 16u;
  float na[8] = {0, 0, 0, 0, 0, 0, 0, 0}, nb[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // a block walks whole input rows (few blocks, so the upstream sums cost a handful of atomics per channel)
  // (whole rows per block, rows round-robin over the resident blocks; (row, part) jobs measured 10 % slower)
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
  const int n = row / H, ih = row - n * H;
  const bool odd = (ih & 1) != 0;
  const int oh0 = ih >> 1;                  // row-window that holds this row at r = 1 (even row) or r = 2 (odd row)
  const uint32_t r0 = odd ? 6u : 3u;        // r * 3
  const bool two = odd && (oh0 + 1 < OH);   // odd rows are also row 0 of the window below
  for (int item = threadIdx.x; item < items; item += blockDim.x) {
    const int j = item >> cg_shift;
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
    const size_t od = ((size_t)row * W + 2 * j) * cg + g;
    const uint4 x0r = __ldg(x + od), x1r = __ldg(x + od + cg);
    uint32_t q0[4] = {0u, 0u, 0u, 0u}, q1[4] = {0u, 0u, 0u, 0u};
    st_add_masked(q0, cA, gA, r0 + 1u);
    st_add_masked(q1, cA, gA, r0 + 2u);
    st_add_masked(q1, cB, gB, r0);
    if (two) {
      st_add_masked(q0, cC, gC, 1u);
      st_add_masked(q1, cC, gC, 2u);
      st_add_masked(q1, cD, gD, 0u);
    }
    {  // ReLU mask of the fused layer on the packed values: fmaf(x, scale, shift) > 0  <=>  (+-x) > threshold, exactly
      const uint32_t xa[4] = {x0r.x, x0r.y, x0r.z, x0r.w}, xb[4] = {x1r.x, x1r.y, x1r.z, x1r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        q0[q] &= st_gt2_mask(xa[q] ^ flip[q], thr[q]);
        q1[q] &= st_gt2_mask(xb[q] ^ flip[q], thr[q]);
      }
    }
    float o0[8], o1[8], x0[8], x1[8];
    st_bf16x8_to_f32(make_uint4(q0[0], q0[1], q0[2], q0[3]), o0);
    st_bf16x8_to_f32(make_uint4(q1[0], q1[1], q1[2], q1[3]), o1);
    st_bf16x8_to_f32(x0r, x0);
    st_bf16x8_to_f32(x1r, x1);
    {
      float A[8], Bc[8];
      st_lds8(cst, cst_half, A);
      st_lds8(cst + cst_stride, cst_half, Bc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        o0[q] = fmaf(A[q], o0[q], Bc[q] * x0[q]);
        o1[q] = fmaf(A[q], o1[q], Bc[q] * x1[q]);
      }
    }
    {
      float Cc[8];
      st_lds8(cst + 2u * cst_stride, cst_half, Cc);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        o0[q] += Cc[q];
        o1[q] += Cc[q];
      }
    }
    const uint4 p0 = st_f32_to_bf16x8(o0), p1 = st_f32_to_bf16x8(o1);
    dx[od] = p0;
    dx[od + cg] = p1;
    if (NEXT) {  // backward sums of the upstream BatchNorm whose ReLU output x is, from the fp32 values before dx is rounded for
                 // storage (a reduce pass would read the rounded ones back: the difference is a zero-mean sum of rounding errors)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        na[q] += (x0[q] > 0.f ? o0[q] : 0.f) + (x1[q] > 0.f ? o1[q] : 0.f);
        nb[q] = fmaf(o0[q], x0[q], fmaf(o1[q], x1[q], nb[q]));
      }
    }
  }
  }
  if (NEXT) {
    float ta[2048 / BT], tb[2048 / BT];
    st_block_channel_sum<BT>(na, sm_next, cg, ta);
    st_block_channel_sum<BT>(nb, sm_next, cg, tb);
#pragma unroll
    for (int j = 0; j < 2048 / BT; ++j) {
      const int c = threadIdx.x + j * BT;
      if (c < cg * 8) {
        atomicAdd(next_s1 + c, (double)ta[j]);
        atomicAdd(next_s2 + c, (double)tb[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- two BatchNorms
// The same backward continued through the conv + BatchNorm + ReLU in FRONT of bn1 (the stem block's conv2, basics.py:122-125): x =
// relu(BN_up(raw)) is recomputed from the raw conv output exactly as the forward stored it (fma, max, round to bf16), the gradient
// dx = A*d + B*x + C of the kernel above stays in registers, and what is written is d raw = A'*[x > 0]*dx + B'*raw + C' of the
// upstream BatchNorm. Its two backward sums are not reduced here: they follow in closed form from the pooled-grid sums of the reduce
// kernel and the forward statistics of x (sum x, sum x^2, count of x > 0), see train.bn_relu_maxpool_op. One pass of
// dy/4 + codes/8 + raw + d raw instead of (dy/4 + codes/8 + x + dx) + (dx + raw + d raw).
struct StUpstream {
  const float* scale;          // forward affine of the upstream BatchNorm: x = relu(scale*raw + shift)
  const float* shift;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const double* sum_dy;        // sum dx*[x > 0]
  const double* sum_dy_xhat;   // sum dx*[x > 0]*xhat_up
  float inv_n;
  StParamGrads pg;
};

template <int BT>
__global__ void __launch_bounds__(BT, 768 / BT) bn2_relu_maxpool_bwd_apply_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                                                 const uint4* __restrict__ raw, uint4* __restrict__ draw, int rows,
                                                                                 int H, int W, int OH, int OW, int cg, int cg_shift,
                                                                                 const float* __restrict__ fsc, const float* __restrict__ fsh,
                                                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                                 const float* __restrict__ gamma, const double* __restrict__ sum_dy,
                                                                                 const double* __restrict__ sum_dy_xhat, float inv_n,
                                                                                 StParamGrads pg, StUpstream up) {
  extern __shared__ float st_cst[];   // [7][C]: A'A, A'B, A'C of this BatchNorm; scale, shift, B', C' of the upstream one
  __shared__ uint32_t sm_thr[BT * 4], sm_flip[BT * 4];
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < pg.n; c += BT) {
      if (pg.dbeta) pg.dbeta[c] = (pg.accumulate ? pg.dbeta[c] : 0.f) + (float)sum_dy[c];
      if (pg.dgamma) pg.dgamma[c] = (pg.accumulate ? pg.dgamma[c] : 0.f) + (float)sum_dy_xhat[c];
    }
    for (int c = threadIdx.x; c < up.pg.n; c += BT) {
      if (up.pg.dbeta) up.pg.dbeta[c] = (up.pg.accumulate ? up.pg.dbeta[c] : 0.f) + (float)up.sum_dy[c];
      if (up.pg.dgamma) up.pg.dgamma[c] = (up.pg.accumulate ? up.pg.dgamma[c] : 0.f) + (float)up.sum_dy_xhat[c];
    }
  }
  const int C8 = cg * 8;
  for (int c = threadIdx.x; c < C8; c += BT) {
    const int ci = st_cst_index(c, cg);
    // dx = A*d + B*x + C of this BatchNorm and d raw = A'*[x > 0]*dx + B'*raw + C' of the upstream one: A' is folded into (A, B, C)
    const double ur = (double)__ldg(up.rstd + c), um = (double)__ldg(up.mean + c), ugm = up.gamma ? (double)__ldg(up.gamma + c) : 1.0;
    const double uc1 = up.sum_dy[c] * (double)up.inv_n, uc2 = up.sum_dy_xhat[c] * (double)up.inv_n;
    const double a2 = ugm * ur;
    const double r = (double)__ldg(rstd + c), m = (double)__ldg(mean + c), gm = gamma ? (double)__ldg(gamma + c) : 1.0;
    const double c1 = sum_dy[c] * (double)inv_n, c2 = sum_dy_xhat[c] * (double)inv_n;
    st_cst[ci] = (float)(a2 * gm * r);
    st_cst[C8 + ci] = (float)(-a2 * gm * r * r * c2);
    st_cst[2 * C8 + ci] = (float)(a2 * gm * r * (r * c2 * m - c1));
    st_cst[3 * C8 + ci] = __ldg(up.scale + c);
    st_cst[4 * C8 + ci] = __ldg(up.shift + c);
    st_cst[5 * C8 + ci] = (float)(-ugm * ur * ur * uc2);
    st_cst[6 * C8 + ci] = (float)(ugm * ur * (ur * uc2 * um - uc1));
  }
  for (int c2 = threadIdx.x; c2 < cg * 4; c2 += BT) {
    const float s0 = __ldg(fsc + 2 * c2), s1 = __ldg(fsc + 2 * c2 + 1);
    sm_thr[c2] = st_relu_threshold(s0, __ldg(fsh + 2 * c2)) | (st_relu_threshold(s1, __ldg(fsh + 2 * c2 + 1)) << 16);
    sm_flip[c2] = (s0 < 0.f ? 0x00008000u : 0u) | (s1 < 0.f ? 0x80000000u : 0u);
  }
  __syncthreads();
  const int items = (W >> 1) * cg;
  const int g = threadIdx.x & (cg - 1);
  uint32_t thr[4], flip[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    thr[q] = sm_thr[g * 4 + q];
    flip[q] = sm_flip[g * 4 + q];
  }
  const uint32_t cst = (uint32_t)__cvta_generic_to_shared(st_cst) + (uint32_t)g * 16u;
  const uint32_t cst_stride = (uint32_t)cg * 32u, cst_half = (uint32_t)cg * 16u;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, ih = row - n * H;
    const bool odd = (ih & 1) != 0;
    const int oh0 = ih >> 1;
    const uint32_t r0 = odd ? 6u : 3u;
    const bool two = odd && (oh0 + 1 < OH);
    for (int item = threadIdx.x; item < items; item += BT) {
      const int j = item >> cg_shift;
      const bool right = j + 1 < OW;
      const size_t oa = (((size_t)n * OH + oh0) * OW + j) * cg + g;
      const size_t oc = oa + (size_t)OW * cg;
      const uint2 zc = make_uint2(0xffffffffu, 0xffffffffu);
      const uint4 zg = make_uint4(0u, 0u, 0u, 0u);
      const uint2 cA = __ldg(idx + oa);
      const uint4 gA = __ldg(dy + oa);
      const uint2 cB = right ? __ldg(idx + oa + cg) : zc;
      const uint4 gB = right ? __ldg(dy + oa + cg) : zg;
      const uint2 cC = two ? __ldg(idx + oc) : zc;
      const uint4 gC = two ? __ldg(dy + oc) : zg;
      const uint2 cD = (two && right) ? __ldg(idx + oc + cg) : zc;
      const uint4 gD = (two && right) ? __ldg(dy + oc + cg) : zg;
      const size_t od = ((size_t)row * W + 2 * j) * cg + g;
      const uint4 w0r = __ldg(raw + od), w1r = __ldg(raw + od + cg);
      uint32_t q0[4] = {0u, 0u, 0u, 0u}, q1[4] = {0u, 0u, 0u, 0u};
      st_add_masked(q0, cA, gA, r0 + 1u);
      st_add_masked(q1, cA, gA, r0 + 2u);
      st_add_masked(q1, cB, gB, r0);
      if (two) {
        st_add_masked(q0, cC, gC, 1u);
        st_add_masked(q1, cC, gC, 2u);
        st_add_masked(q1, cD, gD, 0u);
      }
      float w0[8], w1[8], x0[8], x1[8];
      st_bf16x8_to_f32(w0r, w0);
      st_bf16x8_to_f32(w1r, w1);
      uint4 x0r, x1r;
      {  // x as the forward stored it: fma, ReLU, round to bf16
        float usc[8], ush[8];
        st_lds8(cst + 3u * cst_stride, cst_half, usc);
        st_lds8(cst + 4u * cst_stride, cst_half, ush);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          x0[q] = fmaxf(fmaf(w0[q], usc[q], ush[q]), 0.f);
          x1[q] = fmaxf(fmaf(w1[q], usc[q], ush[q]), 0.f);
        }
        x0r = st_f32_to_bf16x8(x0);
        x1r = st_f32_to_bf16x8(x1);
        st_bf16x8_to_f32(x0r, x0);
        st_bf16x8_to_f32(x1r, x1);
      }
      {
        const uint32_t xa[4] = {x0r.x, x0r.y, x0r.z, x0r.w}, xb[4] = {x1r.x, x1r.y, x1r.z, x1r.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          q0[q] &= st_gt2_mask(xa[q] ^ flip[q], thr[q]);
          q1[q] &= st_gt2_mask(xb[q] ^ flip[q], thr[q]);
        }
      }
      float o0[8], o1[8];
      st_bf16x8_to_f32(make_uint4(q0[0], q0[1], q0[2], q0[3]), o0);
      st_bf16x8_to_f32(make_uint4(q1[0], q1[1], q1[2], q1[3]), o1);
      {
        float A[8], Bc[8];
        st_lds8(cst, cst_half, A);
        st_lds8(cst + cst_stride, cst_half, Bc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          o0[q] = fmaf(A[q], o0[q], Bc[q] * x0[q]);
          o1[q] = fmaf(A[q], o1[q], Bc[q] * x1[q]);
        }
      }
      {
        float Cc[8];
        st_lds8(cst + 2u * cst_stride, cst_half, Cc);
#pragma unroll
        for (int q = 0; q < 8; ++q) {   // A' * dx complete; through the upstream ReLU
          o0[q] = x0[q] > 0.f ? o0[q] + Cc[q] : 0.f;
          o1[q] = x1[q] > 0.f ? o1[q] + Cc[q] : 0.f;
        }
      }
      {
        float B2[8], C2[8];
        st_lds8(cst + 5u * cst_stride, cst_half, B2);
        st_lds8(cst + 6u * cst_stride, cst_half, C2);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          o0[q] += fmaf(B2[q], w0[q], C2[q]);
          o1[q] += fmaf(B2[q], w1[q], C2[q]);
        }
      }
      draw[od] = st_f32_to_bf16x8(o0);
      draw[od + cg] = st_f32_to_bf16x8(o1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------- ECA + BatchNorm
// EfficientConvBlock's second gate in training (basics.py:118-121): c1 = relu(BN(raw)) -> c1s = c1 * gate[n, c] -> conv2. The data
// gradient of conv2 arrives as dy = d c1s. The separate launches spend three passes of 2 + 3 + 3 tensors on it (gate gradient
// sum dy*c1; dc1 = dy*gate + dmean stored; BatchNorm apply reading dc1 and raw). dc1 is an affine function of dy per (image,
// channel), so it never has to exist in memory:
//   sums   per (image, channel): P1 = sum dy*[c1 > 0], P2 = sum dy*c1 (= d gate), M0 = sum [c1 > 0]. With the gate's own backward
//          (dmean[n, c], a tiny kernel) the BatchNorm's two backward sums follow on the host side of the launch:
//          sum dc1*m = sum_n gate*P1 + dmean*M0,  sum dc1*c1 = sum_n gate*P2 + dmean*sum c1 (the forward's pooled sums).
//   apply  draw = A * m * (dy*gate[n, c] + dmean[n, c]) + B * raw + C with (A, B, C) as in bn_relu_maxpool_bwd_apply: 2 reads + 1 write.
__global__ void __launch_bounds__(kRedThreads) eca_bn_bwd_sums_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ c1, int hw,
                                                                      int cg, int pix_per_block, double* __restrict__ P1,
                                                                      double* __restrict__ P2, double* __restrict__ M0,
                                                                      long long out_stride) {
  __shared__ float sm[kRedThreads * 8];
  const int lanes = blockDim.x / cg;
  const int g = threadIdx.x & (cg - 1), lane = threadIdx.x / cg;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, hw);
  const uint4* dyi = dy + (size_t)n * hw * cg + g;
  const uint4* xi = c1 + (size_t)n * hw * cg + g;
  float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, b[8] = {0, 0, 0, 0, 0, 0, 0, 0}, m0[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = p0 + lane; p < p1; p += 2 * lanes) {   // two pixels per round: four 16-byte loads in flight per thread
    const bool two = p + lanes < p1;
    const uint4 d0 = __ldg(dyi + (size_t)p * cg), x0 = __ldg(xi + (size_t)p * cg);
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4 d1 = two ? __ldg(dyi + (size_t)(p + lanes) * cg) : zero, x1 = two ? __ldg(xi + (size_t)(p + lanes) * cg) : zero;
    float dv0[8], xv0[8], dv1[8], xv1[8];
    st_bf16x8_to_f32(d0, dv0);
    st_bf16x8_to_f32(x0, xv0);
    st_bf16x8_to_f32(d1, dv1);
    st_bf16x8_to_f32(x1, xv1);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const bool k0 = xv0[q] > 0.f, k1 = xv1[q] > 0.f;
      a[q] += (k0 ? dv0[q] : 0.f) + (k1 ? dv1[q] : 0.f);
      b[q] = fmaf(dv0[q], xv0[q], fmaf(dv1[q], xv1[q], b[q]));
      m0[q] += (k0 ? 1.f : 0.f) + (k1 ? 1.f : 0.f);
    }
  }
  float ta[kRedMaxIter], tb[kRedMaxIter], tm[kRedMaxIter];
  block_channel_sum(a, sm, cg, lanes, ta);
  block_channel_sum(b, sm, cg, lanes, tb);
  block_channel_sum(m0, sm, cg, lanes, tm);
#pragma unroll
  for (int j = 0; j < kRedMaxIter; ++j) {
    const int c = threadIdx.x + j * kRedThreads;
    if (c < cg * 8) {
      atomicAdd(P1 + n * out_stride + c, (double)ta[j]);
      atomicAdd(P2 + n * out_stride + c, (double)tb[j]);
      atomicAdd(M0 + n * out_stride + c, (double)tm[j]);
    }
  }
}

template <int BT>
__global__ void __launch_bounds__(BT, 640 / BT) eca_bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ raw,
                                                                         uint4* __restrict__ dx, int n_img, int hw, int cg,
                                                                         const float* __restrict__ gate,
                                                                         long long gate_stride, const float* __restrict__ dmean,
                                                                         long long dmean_stride, const float* __restrict__ fsc,
                                                                         const float* __restrict__ fsh, const float* __restrict__ mean,
                                                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                         const double* __restrict__ sum_dy,
                                                                         const double* __restrict__ sum_dy_xhat, float inv_n,
                                                                         StParamGrads pg) {
  __shared__ uint32_t sm_thr[BT * 4], sm_flip[BT * 4];
  __shared__ float sm_a[BT * 8], sm_b[BT * 8], sm_c[BT * 8];
  if (blockIdx.x == 0 && pg.n > 0) {   // the BatchNorm affine gradients are the two reductions themselves
    for (int c = threadIdx.x; c < pg.n; c += BT) {
      if (pg.dbeta) pg.dbeta[c] = (pg.accumulate ? pg.dbeta[c] : 0.f) + (float)sum_dy[c];
      if (pg.dgamma) pg.dgamma[c] = (pg.accumulate ? pg.dgamma[c] : 0.f) + (float)sum_dy_xhat[c];
    }
  }
  for (int c = threadIdx.x; c < cg * 8; c += BT) {  // draw = gamma*rstd*(d - c1 - (raw - mean)*rstd*c2) = A*d + B*raw + C
    const double r = (double)__ldg(rstd + c), m = (double)__ldg(mean + c), gm = gamma ? (double)__ldg(gamma + c) : 1.0;
    const double c1 = sum_dy[c] * (double)inv_n, c2 = sum_dy_xhat[c] * (double)inv_n;
    sm_a[c] = (float)(gm * r);
    sm_b[c] = (float)(-gm * r * r * c2);
    sm_c[c] = (float)(gm * r * (r * c2 * m - c1));
  }
  for (int c2 = threadIdx.x; c2 < cg * 4; c2 += BT) {
    const float s0 = __ldg(fsc + 2 * c2), s1 = __ldg(fsc + 2 * c2 + 1);
    sm_thr[c2] = st_relu_threshold(s0, __ldg(fsh + 2 * c2)) | (st_relu_threshold(s1, __ldg(fsh + 2 * c2 + 1)) << 16);
    sm_flip[c2] = (s0 < 0.f ? 0x00008000u : 0u) | (s1 < 0.f ? 0x80000000u : 0u);
  }
  __syncthreads();
  const int g = threadIdx.x & (cg - 1);
  uint32_t thr[4], flip[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    thr[q] = sm_thr[g * 4 + q];
    flip[q] = sm_flip[g * 4 + q];
  }
  float Bc[8], Cc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    Bc[q] = sm_b[g * 8 + q];
    Cc[q] = sm_c[g * 8 + q];
  }
  const unsigned img_items = (unsigned)(hw * cg);
  const unsigned total = img_items * (unsigned)n_img;
  // the whole grid sweeps the tensor front to back (all SMs inside the same few MB at any time: a block per image region ran at
  // 4.3 instead of 6+ TB/s); the (image, channel) factors are reloaded when a thread's item crosses into another image
  const unsigned stride = gridDim.x * BT;   // a multiple of cg: one channel group per thread
  int cur_n = -1;
  float Ag[8], Ad[8];
  auto one = [&](const uint4& dr, const uint4& xr, unsigned item) -> uint4 {
    const int n = (int)(item / img_items);
    if (n != cur_n) {
      cur_n = n;
      float gt[8], dm[8];
      st_load8(gate + n * gate_stride + g * 8, gt);
      st_load8(dmean + n * dmean_stride + g * 8, dm);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float A = sm_a[g * 8 + q];
        Ag[q] = A * gt[q];
        Ad[q] = A * dm[q];
      }
    }
    const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
    uint32_t dw[4] = {dr.x, dr.y, dr.z, dr.w};
    uint32_t mk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mk[q] = st_gt2_mask(xw[q] ^ flip[q], thr[q]);   // ReLU mask of the forward, exactly (see st_relu_threshold)
      dw[q] &= mk[q];
    }
    float d[8], x[8], o[8];
    st_bf16x8_to_f32(make_uint4(dw[0], dw[1], dw[2], dw[3]), d);
    st_bf16x8_to_f32(xr, x);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const bool on = (mk[q >> 1] >> ((q & 1) * 16)) & 1u;
      o[q] = fmaf(Ag[q], d[q], fmaf(Bc[q], x[q], Cc[q] + (on ? Ad[q] : 0.f)));
    }
    return st_f32_to_bf16x8(o);
  };
  unsigned item = blockIdx.x * BT + threadIdx.x;
  for (; item < total && total - item > 3u * stride; item += 4u * stride) {   // four independent 16-byte loads per tensor in flight
    uint4 dr[4], xr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dr[u] = __ldg(dy + item + u * stride);
      xr[u] = __ldg(raw + item + u * stride);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) dx[item + u * stride] = one(dr[u], xr[u], item + u * stride);
  }
  for (; item < total; item += stride) dx[item] = one(__ldg(dy + item), __ldg(raw + item), item);
}

static bool st_dense(const PmoeView4* v) {
  return v && v->ptr && v->c % 8 == 0 && ((uintptr_t)v->ptr % 16) == 0 && v->sw == v->c && v->sh == (int64_t)v->w * v->c &&
         v->sn == (int64_t)v->h * v->w * v->c;
}

static int st_geometry(const PmoeView4* x, const PmoeView4* y, const char* what) {
  if (!st_dense(x) || !st_dense(y)) {
    set_error("%s: dense 16-byte aligned NHWC bf16 tensors with a multiple of 8 channels", what);
    return PMOE_ERR_UNSUPPORTED;
  }
  const int cg = x->c / 8;
  if (x->h % 2 || x->w % 2 || y->h != x->h / 2 || y->w != x->w / 2 || y->n != x->n || y->c != x->c || cg > 256 || 256 % cg != 0) {
    set_error("%s: MaxPool2d(3, 2, 1) of an even-sized input, channel-group count dividing 256", what);
    return PMOE_ERR_UNSUPPORTED;
  }
  if ((long long)x->n * x->h > 2147483647LL) {
    set_error("%s: too many rows", what);
    return PMOE_ERR_UNSUPPORTED;
  }
  return PMOE_OK;
}

}  // namespace pmoe

using namespace pmoe;

extern "C" int pmoe_bn_relu_maxpool_fwd(const PmoeView4* x, const float* scale, const float* shift, const PmoeView4* y, uint8_t* idx,
                                        void* x_at_max, pmoe_stream_t stream_) {
  int rc = st_geometry(x, y, "bn_relu_maxpool_fwd");
  if (rc) return rc;
  if (!scale || !shift || !idx || !x_at_max || ((uintptr_t)scale % 16) || ((uintptr_t)shift % 16) || ((uintptr_t)idx % 8) || ((uintptr_t)x_at_max % 16)) {
    set_error("bn_relu_maxpool_fwd: scale / shift (16-byte aligned), argmax codes and x-at-argmax buffers are required");
    return PMOE_ERR_ARG;
  }
  const int cg = x->c / 8;
  dim3 grid((unsigned)(y->n * y->h), (unsigned)((y->w * cg + 255) / 256));
  bn_relu_maxpool_fwd_kernel<4><<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(x->ptr), static_cast<uint4*>(y->ptr), reinterpret_cast<uint2*>(idx), static_cast<uint4*>(x_at_max), x->h, x->w,
      y->h, y->w, cg, __builtin_ctz((unsigned)cg), scale, shift);   // (5 or 6 resident blocks per SM spill and measured 3-7 % slower)
  return check_launch("bn_relu_maxpool_fwd");
}

extern "C" int pmoe_bn_relu_maxpool_bwd_reduce(const PmoeView4* dy, const void* x_at_max, const float* fwd_scale, const float* fwd_shift,
                                               const float* mean, const float* rstd, double* sum_dy, double* sum_dy_xhat,
                                               double* sum_dy_xpos, pmoe_stream_t stream_) {
  if (!st_dense(dy) || !x_at_max || !fwd_scale || !fwd_shift || !mean || !rstd || !sum_dy || !sum_dy_xhat || ((uintptr_t)x_at_max % 16) ||
      ((uintptr_t)fwd_scale % 16) || ((uintptr_t)fwd_shift % 16)) {
    set_error("bn_relu_maxpool_bwd_reduce: dense bf16 pooled gradient and all statistics are required");
    return PMOE_ERR_ARG;
  }
  const int cg = dy->c / 8;
  if (cg > 256 || 256 % cg != 0) {
    set_error("bn_relu_maxpool_bwd_reduce: channel-group count must divide 256");
    return PMOE_ERR_UNSUPPORTED;
  }
  const long long npix = (long long)dy->n * dy->h * dy->w;
  long long blocks = (long long)num_sms() * 8;
  long long ppb = (npix + blocks - 1) / blocks;
  if (ppb < 64) ppb = 64;
  blocks = (npix + ppb - 1) / ppb;
  bn_relu_maxpool_bwd_reduce_kernel<<<(unsigned)blocks, kRedThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(dy->ptr), static_cast<const uint4*>(x_at_max), npix, cg, fwd_scale, fwd_shift, mean, rstd, sum_dy, sum_dy_xhat,
      ppb, sum_dy_xpos);
  return check_launch("bn_relu_maxpool_bwd_reduce");
}

extern "C" int pmoe_bn_relu_maxpool_bwd_apply(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* x, const float* fwd_scale,
                                              const float* fwd_shift, const float* mean, const float* rstd, const float* gamma,
                                              const double* sum_dy, const double* sum_dy_xhat, float inv_n, const PmoeView4* dx,
                                              double* next_sum_dx, double* next_sum_dx_x, const PmoeBnParamGrads* param_grads,
                                              pmoe_stream_t stream_) {
  int rc = st_geometry(x, dy, "bn_relu_maxpool_bwd_apply");
  if (rc) return rc;
  if (!st_dense(dx) || dx->n != x->n || dx->h != x->h || dx->w != x->w || dx->c != x->c || !idx || ((uintptr_t)idx % 8) || !fwd_scale ||
      !fwd_shift || !mean || !rstd || !sum_dy || !sum_dy_xhat || ((uintptr_t)fwd_scale % 16) || ((uintptr_t)fwd_shift % 16) ||
      (next_sum_dx != nullptr) != (next_sum_dx_x != nullptr)) {
    set_error("bn_relu_maxpool_bwd_apply: bad arguments");
    return PMOE_ERR_ARG;
  }
  StParamGrads pg = {nullptr, nullptr, 0, 0};
  if (param_grads) {
    if (param_grads->n < 0 || param_grads->n > x->c) {
      set_error("bn_relu_maxpool_bwd_apply: parameter-gradient channel count out of range");
      return PMOE_ERR_ARG;
    }
    pg.dgamma = param_grads->dgamma;
    pg.dbeta = param_grads->dbeta;
    pg.n = param_grads->n;
    pg.accumulate = param_grads->accumulate;
  }
  const int cg = x->c / 8;
  const int rows = x->n * x->h;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t cst_bytes = (size_t)3 * x->c * sizeof(float);   // <= 24 KB (cg <= 256)
  // ONE resident wave: blocks walk rows round-robin, so every SM holds its full complement of blocks until the end (8 blocks per
  // SM at 3 resident ran as 3 + 3 + 2), and the upstream sums cost one atomic per channel and resident block.
  const uint4* dyp = static_cast<const uint4*>(dy->ptr);
  const uint2* idp = reinterpret_cast<const uint2*>(idx);
  const uint4* xp = static_cast<const uint4*>(x->ptr);
  uint4* dxp = static_cast<uint4*>(dx->ptr);
  const int sh_ = __builtin_ctz((unsigned)cg);
#define ST_APPLY(NEXT_, BT_)                                                                                                       \
  do {                                                                                                                             \
    static int per_sm_cached = 0; /* per instantiation; the 3 * C floats of dynamic shared memory never change the answer here */ \
    if (per_sm_cached == 0) {                                                                                                      \
      int q_ = 0;                                                                                                                  \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q_, bn_relu_maxpool_bwd_apply_kernel<NEXT_, BT_>, BT_, 24 * 1024);            \
      per_sm_cached = q_ < 1 ? 1 : q_;                                                                                             \
    }                                                                                                                              \
    const int per_sm = per_sm_cached;                                                                                              \
    int grid = num_sms() * per_sm;                                                                                                 \
    if (grid > rows) grid = rows;                                                                                                  \
    bn_relu_maxpool_bwd_apply_kernel<NEXT_, BT_><<<grid, BT_, cst_bytes, stream>>>(                                                \
        dyp, idp, xp, dxp, rows, x->h, x->w, dy->h, dy->w, cg, sh_, fwd_scale, fwd_shift, mean, rstd, gamma, sum_dy, sum_dy_xhat,  \
        inv_n, next_sum_dx, next_sum_dx_x, pg);                                                                                    \
  } while (0)
  if (cg <= 128) {
    if (next_sum_dx) ST_APPLY(true, 128); else ST_APPLY(false, 128);
  } else {
    if (next_sum_dx) ST_APPLY(true, 256); else ST_APPLY(false, 256);
  }
#undef ST_APPLY
  return check_launch("bn_relu_maxpool_bwd_apply");
}

extern "C" int pmoe_eca_bn_bwd_sums(const PmoeView4* dy, const PmoeView4* c1, double* sum_dy_m, double* sum_dy_c1, double* sum_m,
                                    int64_t out_stride, pmoe_stream_t stream_) {
  if (!st_dense(dy) || !st_dense(c1) || dy->n != c1->n || dy->h != c1->h || dy->w != c1->w || dy->c != c1->c || !sum_dy_m || !sum_dy_c1 ||
      !sum_m || out_stride < dy->c) {
    set_error("eca_bn_bwd_sums: two dense bf16 NHWC tensors of one shape and three (n, >= c) fp64 outputs are required");
    return PMOE_ERR_UNSUPPORTED;
  }
  const int cg = dy->c / 8;
  const long long hw = (long long)dy->h * dy->w;
  if (cg > 256 || 256 % cg != 0 || hw * cg > 2147483647LL || dy->n > 65535) {
    set_error("eca_bn_bwd_sums: channel-group count must divide 256, image size and count within the grid limits");
    return PMOE_ERR_UNSUPPORTED;
  }
  long long blocks = ((long long)num_sms() * 8 + dy->n - 1) / dy->n;   // per image
  long long ppb = (hw + blocks - 1) / blocks;
  if (ppb < 64) ppb = 64;
  blocks = (hw + ppb - 1) / ppb;
  dim3 grid((unsigned)blocks, (unsigned)dy->n);
  eca_bn_bwd_sums_kernel<<<grid, kRedThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(dy->ptr), static_cast<const uint4*>(c1->ptr), (int)hw, cg, (int)ppb, sum_dy_m, sum_dy_c1, sum_m, out_stride);
  return check_launch("eca_bn_bwd_sums");
}

extern "C" int pmoe_eca_bn_bwd_apply(const PmoeView4* dy, const PmoeView4* raw, const float* gate, int64_t gate_stride, const float* dmean,
                                     int64_t dmean_stride, const float* fwd_scale, const float* fwd_shift, const float* mean,
                                     const float* rstd, const float* gamma, const double* sum_dy, const double* sum_dy_xhat, float inv_n,
                                     const PmoeView4* dx, const PmoeBnParamGrads* param_grads, pmoe_stream_t stream_) {
  if (!st_dense(dy) || !st_dense(raw) || !st_dense(dx) || dy->n != raw->n || dy->h != raw->h || dy->w != raw->w || dy->c != raw->c ||
      dx->n != raw->n || dx->h != raw->h || dx->w != raw->w || dx->c != raw->c) {
    set_error("eca_bn_bwd_apply: three dense bf16 NHWC tensors of one shape are required");
    return PMOE_ERR_UNSUPPORTED;
  }
  if (!gate || !dmean || !fwd_scale || !fwd_shift || !mean || !rstd || !sum_dy || !sum_dy_xhat || gate_stride < dy->c || dmean_stride < dy->c ||
      ((uintptr_t)gate % 16) || ((uintptr_t)dmean % 16) || gate_stride % 4 || dmean_stride % 4) {
    set_error("eca_bn_bwd_apply: bad arguments (gate / dmean rows 16-byte aligned, all statistics required)");
    return PMOE_ERR_ARG;
  }
  const int cg = dy->c / 8;
  const long long hw = (long long)dy->h * dy->w;
  if (cg > 128 || 128 % cg != 0 || hw * cg * dy->n > 2147483647LL) {
    set_error("eca_bn_bwd_apply: channel-group count must divide 128, at most 2^31 16-byte items");
    return PMOE_ERR_UNSUPPORTED;
  }
  StParamGrads pg = {nullptr, nullptr, 0, 0};
  if (param_grads) {
    if (param_grads->n < 0 || param_grads->n > dy->c) {
      set_error("eca_bn_bwd_apply: parameter-gradient channel count out of range");
      return PMOE_ERR_ARG;
    }
    pg.dgamma = param_grads->dgamma;
    pg.dbeta = param_grads->dbeta;
    pg.n = param_grads->n;
    pg.accumulate = param_grads->accumulate;
  }
  constexpr int BT = 128;
  static int per_sm = 0;
  if (per_sm == 0) {
    int q = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, eca_bn_bwd_apply_kernel<BT>, BT, 0);
    per_sm = q < 1 ? 1 : q;
  }
  long long grid = (long long)num_sms() * per_sm;                 // one resident wave sweeping the tensor front to back
  const long long need = (hw * cg * dy->n + BT - 1) / BT;
  if (grid > need) grid = need;
  eca_bn_bwd_apply_kernel<BT><<<(unsigned)grid, BT, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(dy->ptr), static_cast<const uint4*>(raw->ptr), static_cast<uint4*>(dx->ptr), dy->n, (int)hw, cg, gate,
      gate_stride, dmean, dmean_stride, fwd_scale, fwd_shift, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, pg);
  return check_launch("eca_bn_bwd_apply");
}

extern "C" int pmoe_bn2_relu_maxpool_bwd_apply(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* raw, const float* fwd_scale,
                                               const float* fwd_shift, const float* mean, const float* rstd, const float* gamma,
                                               const double* sum_dy, const double* sum_dy_xhat, float inv_n,
                                               const PmoeBnParamGrads* param_grads, const float* up_scale, const float* up_shift,
                                               const float* up_mean, const float* up_rstd, const float* up_gamma, const double* up_sum_dy,
                                               const double* up_sum_dy_xhat, const PmoeBnParamGrads* up_param_grads, const PmoeView4* draw,
                                               pmoe_stream_t stream_) {
  int rc = st_geometry(raw, dy, "bn2_relu_maxpool_bwd_apply");
  if (rc) return rc;
  if (!st_dense(draw) || draw->n != raw->n || draw->h != raw->h || draw->w != raw->w || draw->c != raw->c || !idx || ((uintptr_t)idx % 8) ||
      !fwd_scale || !fwd_shift || !mean || !rstd || !sum_dy || !sum_dy_xhat || !up_scale || !up_shift || !up_mean || !up_rstd || !up_sum_dy ||
      !up_sum_dy_xhat) {
    set_error("bn2_relu_maxpool_bwd_apply: bad arguments");
    return PMOE_ERR_ARG;
  }
  const int cg = raw->c / 8;
  if (cg > 128 || 128 % cg != 0 || (size_t)7 * raw->c * sizeof(float) > 40 * 1024) {
    set_error("bn2_relu_maxpool_bwd_apply: channel-group count must divide 128, at most 1280 channels");
    return PMOE_ERR_UNSUPPORTED;
  }
  auto to_pg = [&](const PmoeBnParamGrads* g, StParamGrads* o) {
    *o = StParamGrads{nullptr, nullptr, 0, 0};
    if (!g) return true;
    if (g->n < 0 || g->n > raw->c) return false;
    o->dgamma = g->dgamma;
    o->dbeta = g->dbeta;
    o->n = g->n;
    o->accumulate = g->accumulate;
    return true;
  };
  StParamGrads pg;
  StUpstream up;
  if (!to_pg(param_grads, &pg) || !to_pg(up_param_grads, &up.pg)) {
    set_error("bn2_relu_maxpool_bwd_apply: parameter-gradient channel count out of range");
    return PMOE_ERR_ARG;
  }
  up.scale = up_scale;
  up.shift = up_shift;
  up.mean = up_mean;
  up.rstd = up_rstd;
  up.gamma = up_gamma;
  up.sum_dy = up_sum_dy;
  up.sum_dy_xhat = up_sum_dy_xhat;
  up.inv_n = inv_n;
  constexpr int BT = 128;
  const size_t cst_bytes = (size_t)7 * raw->c * sizeof(float);
  static int per_sm = 0;
  if (per_sm == 0) {
    int q = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, bn2_relu_maxpool_bwd_apply_kernel<BT>, BT, 40 * 1024);
    per_sm = q < 1 ? 1 : q;
  }
  const int rows = raw->n * raw->h;
  int grid = num_sms() * per_sm;
  if (grid > rows) grid = rows;
  bn2_relu_maxpool_bwd_apply_kernel<BT><<<grid, BT, cst_bytes, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const uint4*>(dy->ptr), reinterpret_cast<const uint2*>(idx), static_cast<const uint4*>(raw->ptr), static_cast<uint4*>(draw->ptr),
      rows, raw->h, raw->w, dy->h, dy->w, cg, __builtin_ctz((unsigned)cg), fwd_scale, fwd_shift, mean, rstd, gamma, sum_dy, sum_dy_xhat, inv_n, pg,
      up);
  return check_launch("bn2_relu_maxpool_bwd_apply");
}
