// Fused segmentation loss (k11): cross_entropy_tversky_weighted_loss of trainer/loss.py:6-55 on fp32 NCHW logits.
//   pass A (one read of the logits): online softmax per pixel -> arg-max class, per-class counts for the dice
//           weights, per-class NLL sums, and the (class, column) Tversky accumulators. NB the reference reduces the
//           Tversky terms over (batch, HEIGHT) only (loss.py:40), leaving a (C, W) map that is then averaged.
//   finalize (1 block): dice weights w_c = 1 - 2(|P∩T|+eps)/(|P|+|T|+eps), CE = sum_c w_c nll_c / sum_c w_c n_c,
//           TV = 1 - mean_{c,w} 2 tp/(sum_p + cnt), loss = wc*CE + wt*TV; also the per-(c,w) backward coefficients.
//   pass B (one read, one write): d loss / d logits.
// The reference needs ~12 passes over the (B,23,H,W) tensor plus a 23-iteration Python loop for the same result.
#include "host_util.h"
#include "ptx.cuh"

namespace pmoe {

constexpr int kMaxClasses = 32;

// workspace layout (floats): [0,C) pred_cnt | [C,2C) tgt_cnt | [2C,3C) inter_cnt | [3C,4C) nll_sum |
// [4C,5C) weight | 5C: wsum | then three (C*W) maps: tp, sum_p, cnt ; then two (C*W) maps: coefA, coefB
struct SegWs {
  float* pred_cnt;
  float* tgt_cnt;
  float* inter_cnt;
  float* nll_sum;
  float* weight;
  float* scalars;  // [0]=sum_c w_c*n_c, [1]=ce, [2]=tv
  float* tp;
  float* sum_p;
  float* cnt;
  float* coefA;
  float* coefB;
};
__host__ __device__ inline SegWs seg_ws(float* ws, int C, int W) {
  SegWs s;
  s.pred_cnt = ws;
  s.tgt_cnt = ws + C;
  s.inter_cnt = ws + 2 * C;
  s.nll_sum = ws + 3 * C;
  s.weight = ws + 4 * C;
  s.scalars = ws + 5 * C;
  float* m = ws + 5 * C + 8;
  s.tp = m;
  s.sum_p = m + (size_t)C * W;
  s.cnt = m + 2 * (size_t)C * W;
  s.coefA = m + 3 * (size_t)C * W;
  s.coefB = m + 4 * (size_t)C * W;
  return s;
}

// block = 256 threads = columns of one row chunk; each thread owns one pixel at a time with a fixed column w, so the
// (c, w) accumulators stay in registers across the rows the block walks, and are flushed with one atomic per (c,w).
__global__ void segloss_fwd_kernel(const float* __restrict__ logits, long long sb, long long sc, long long sh, long long sw,
                                   const long long* __restrict__ target, long long tb, long long th, long long tw, int B, int C,
                                   int H, int W, int rows_per_block, float* __restrict__ ws) {
  const SegWs s = seg_ws(ws, C, W);
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const long long nrows = (long long)B * H;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > nrows) r1 = nrows;
  __shared__ float sh_cls[4][kMaxClasses];  // pred_cnt, tgt_cnt, inter_cnt, nll_sum
  for (int i = threadIdx.x; i < 4 * kMaxClasses; i += blockDim.x) (&sh_cls[0][0])[i] = 0.f;
  __syncthreads();
  if (w < W) {
    float tp[kMaxClasses], sp[kMaxClasses], cn[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c) tp[c] = sp[c] = cn[c] = 0.f;
    for (long long r = r0; r < r1; ++r) {
      const int b = (int)(r / H), h = (int)(r % H);
      const float* px = logits + b * sb + h * sh + w * sw;
      float v[kMaxClasses];
      float mx = -INFINITY;
      int arg = 0;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c) {
        if (c < C) {
          v[c] = __ldg(px + c * sc);
          if (v[c] > mx) {
            mx = v[c];
            arg = c;
          }
        }
      }
      float se = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) {
          v[c] = expf(v[c] - mx);
          se += v[c];
        }
      const float inv = 1.f / se;
      const int y = (int)target[b * tb + h * th + w * tw];
#pragma unroll
      for (int c = 0; c < kMaxClasses; ++c)
        if (c < C) {
          const float p = v[c] * inv;
          sp[c] += p;
          if (c == y) {
            tp[c] += p;
            cn[c] += 1.f;
            atomicAdd(&sh_cls[3][c], -logf(fmaxf(p, 1e-45f)));
          }
        }
      atomicAdd(&sh_cls[0][arg], 1.f);
      if (y >= 0 && y < C) {
        atomicAdd(&sh_cls[1][y], 1.f);
        if (arg == y) atomicAdd(&sh_cls[2][y], 1.f);
      }
    }
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        if (tp[c] != 0.f) atomicAdd(s.tp + (size_t)c * W + w, tp[c]);
        atomicAdd(s.sum_p + (size_t)c * W + w, sp[c]);
        if (cn[c] != 0.f) atomicAdd(s.cnt + (size_t)c * W + w, cn[c]);
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (sh_cls[0][i] != 0.f) atomicAdd(s.pred_cnt + i, sh_cls[0][i]);
    if (sh_cls[1][i] != 0.f) atomicAdd(s.tgt_cnt + i, sh_cls[1][i]);
    if (sh_cls[2][i] != 0.f) atomicAdd(s.inter_cnt + i, sh_cls[2][i]);
    if (sh_cls[3][i] != 0.f) atomicAdd(s.nll_sum + i, sh_cls[3][i]);
  }
}

__global__ void segloss_finalize_kernel(float* __restrict__ ws, int C, int W, float wce, float wtv, float dice_eps,
                                        float* __restrict__ loss_out /*[3]: total, ce, tv*/) {
  const SegWs s = seg_ws(ws, C, W);
  __shared__ float red[32];
  __shared__ float sh_w[kMaxClasses];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const float inter = s.inter_cnt[c] + dice_eps;
    const float uni = s.pred_cnt[c] + s.tgt_cnt[c] + dice_eps;
    sh_w[c] = 1.f - 2.f * inter / uni;  // loss.py:13-16 (classes absent from both get 1 - 2 = -1: reproduced)
    s.weight[c] = sh_w[c];
  }
  __syncthreads();
  float num = 0.f, den = 0.f;
  if (threadIdx.x == 0) {
    for (int c = 0; c < C; ++c) {
      num += sh_w[c] * s.nll_sum[c];
      den += sh_w[c] * s.tgt_cnt[c];
    }
    s.scalars[0] = den;
    s.scalars[1] = num / den;
  }
  // tversky mean over (C, W) of 2 tp / (sum_p + cnt); coefficients for backward
  float acc = 0.f;
  const float inv_cw = 1.f / ((float)C * (float)W);
  for (int i = threadIdx.x; i < C * W; i += blockDim.x) {
    const float S = s.sum_p[i] + s.cnt[i];
    const float t = 2.f * s.tp[i] / S;
    acc += t;
    // d TV / d p[b,c,h,w] = -(1/(CW)) * ( onehot * 2/S - 2 tp / S^2 )
    s.coefA[i] = -wtv * inv_cw * 2.f / S;
    s.coefB[i] = wtv * inv_cw * 2.f * s.tp[i] / (S * S);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    const float tv = 1.f - t * inv_cw;
    s.scalars[2] = tv;
    loss_out[1] = s.scalars[1];
    loss_out[2] = tv;
    loss_out[0] = (wce != 0.f ? wce * s.scalars[1] : 0.f) + wtv * tv;  // tversky_loss alone: CE weight 0
  }
}

// dlogits = gscale * softmax-backward( wce * w_y/den * (-onehot/p) + coefA*onehot + coefB )
__global__ void segloss_bwd_kernel(const float* __restrict__ logits, long long sb, long long sc, long long sh, long long sw,
                                   const long long* __restrict__ target, long long tb, long long th, long long tw, int B, int C,
                                   int H, int W, const float* __restrict__ ws, float wce, const float* __restrict__ gscale_ptr,
                                   float gscale_const, float* __restrict__ dlogits, long long db, long long dc, long long dh,
                                   long long dw, int accumulate) {
  const SegWs s = seg_ws(const_cast<float*>(ws), C, W);
  const long long total = (long long)B * H * W;
  const float gscale = gscale_ptr ? *gscale_ptr * gscale_const : gscale_const;
  const float inv_den = 1.f / s.scalars[0];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long t = i / W;
    const int h = (int)(t % H), b = (int)(t / H);
    const float* px = logits + b * sb + h * sh + w * sw;
    float v[kMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        v[c] = __ldg(px + c * sc);
        mx = fmaxf(mx, v[c]);
      }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        v[c] = expf(v[c] - mx);
        se += v[c];
      }
    const float inv = 1.f / se;
    const int y = (int)target[b * tb + h * th + w * tw];
    const float wy = (y >= 0 && y < C && wce != 0.f) ? s.weight[y] * inv_den * wce : 0.f;
    // g_c = dL/dp_c (Tversky part); CE part handled in closed form: wy * (p - onehot)
    float dot = 0.f;
    float g[kMaxClasses];
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        const float p = v[c] * inv;
        v[c] = p;
        g[c] = s.coefB[(size_t)c * W + w] + (c == y ? s.coefA[(size_t)c * W + w] : 0.f);
        dot += p * g[c];
      }
    float* dp = dlogits + b * db + h * dh + w * dw;
#pragma unroll
    for (int c = 0; c < kMaxClasses; ++c)
      if (c < C) {
        const float p = v[c];
        float d = p * (g[c] - dot) + wy * (p - (c == y ? 1.f : 0.f));
        d *= gscale;
        if (accumulate) d += dp[c * dc];
        dp[c * dc] = d;
      }
  }
}

}  // namespace pmoe

using namespace pmoe;

extern "C" {

size_t pmoe_segloss_workspace_floats(int32_t C, int32_t W) { return (size_t)5 * C + 8 + (size_t)5 * C * W; }

int pmoe_segloss_fwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                     int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, float wce, float wtv, float* workspace,
                     float* loss_out, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !workspace || !loss_out || C < 1 || C > kMaxClasses || B < 1 || H < 1 || W < 1) {
    set_error("segloss_fwd: bad arguments (classes must be 1..%d)", kMaxClasses);
    return PMOE_ERR_ARG;
  }
  cudaError_t e = cudaMemsetAsync(workspace, 0, pmoe_segloss_workspace_floats(C, W) * sizeof(float), stream);
  if (e != cudaSuccess) {
    set_error("segloss_fwd: memset failed: %s", cudaGetErrorString(e));
    return PMOE_ERR_LAUNCH;
  }
  const long long nrows = (long long)B * H;
  const int gx = (W + 255) / 256;
  long long want_y = ((long long)num_sms() * 4 + gx - 1) / gx;
  int rows = (int)((nrows + want_y - 1) / want_y);
  if (rows < 4) rows = 4;
  dim3 grid((unsigned)gx, (unsigned)((nrows + rows - 1) / rows));
  segloss_fwd_kernel<<<grid, 256, 0, stream>>>(logits, sb, sc, sh, sw, (const long long*)target, tb, th, tw, B, C, H, W, rows,
                                               workspace);
  int rc = check_launch("segloss_fwd");
  if (rc) return rc;
  segloss_finalize_kernel<<<1, 256, 0, stream>>>(workspace, C, W, wce, wtv, 1e-6f, loss_out);
  return check_launch("segloss_finalize");
}

int pmoe_segloss_bwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                     int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, float wce, const float* workspace,
                     const float* grad_scale_dev, float grad_scale, float* dlogits, int64_t db, int64_t dc, int64_t dh,
                     int64_t dw, int32_t accumulate, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !workspace || !dlogits || C < 1 || C > kMaxClasses) {
    set_error("segloss_bwd: bad arguments");
    return PMOE_ERR_ARG;
  }
  const long long total = (long long)B * H * W;
  long long bl = (total + 255) / 256;
  if (bl > (long long)num_sms() * 8) bl = (long long)num_sms() * 8;
  segloss_bwd_kernel<<<(unsigned)bl, 256, 0, stream>>>(logits, sb, sc, sh, sw, (const long long*)target, tb, th, tw, B, C, H, W,
                                                       workspace, wce, grad_scale_dev, grad_scale, dlogits, db, dc, dh, dw,
                                                       accumulate);
  return check_launch("segloss_bwd");
}

}  // extern "C"
