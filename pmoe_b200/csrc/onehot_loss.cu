// Losses of fp32 NCHW logits against the ONE-HOT of an int64 class map (trainer/loss.py):
//   mode 0  nn.L1Loss(x, onehot)   -> AutoregressiveCriterion(loss_type='l1')   loss.py:93-94,109-116
//   mode 1  nn.MSELoss(x, onehot)  -> AutoregressiveCriterion(loss_type='l2')   loss.py:95-96,109-116
//   mode 2  l1_gdl's last-frame terms (loss.py:58-83): L1 mean plus the gradient-difference sum
//           | |oh[h+1]-oh[h]| - |x[h+1]-x[h]| | + | |oh[w]-oh[w+1]| - |x[w]-x[w+1]| |   with a zero row / column appended at
//           the bottom / right (ZeroPad2d, :67-68), summed over (H,W) and averaged over (B,C) (:79).
// The reference materialises the one-hot tensor (scatter_), two padded copies and ~10 elementwise temporaries; here the
// forward is one read of the logits (neighbours come from L1/L2) and the backward one read + one write.
#include "host_util.h"

namespace pmoe {

__device__ __forceinline__ float sgn(float v) { return (float)((v > 0.f) - (v < 0.f)); }

struct OhArgs {
  const float* x;
  long long sb, sc, sh, sw;
  const long long* tgt;
  long long tb, th, tw;
  int B, C, H, W;
};

// out[0] += sum of the pointwise term (|d| or d^2), out[1] += gradient-difference sum (mode 2)
__global__ void onehot_loss_fwd_kernel(OhArgs a, int mode, double* __restrict__ out) {
  const long long total = (long long)a.B * a.C * a.H * a.W;
  double s0 = 0.0, s1 = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % a.W);
    long long t = i / a.W;
    const int h = (int)(t % a.H);
    t /= a.H;
    const int c = (int)(t % a.C);
    const int b = (int)(t / a.C);
    const float* px = a.x + b * a.sb + c * a.sc + h * a.sh + w * a.sw;
    const long long* pt = a.tgt + b * a.tb + h * a.th + w * a.tw;
    const float x = __ldg(px);
    const float oh = (*pt == c) ? 1.f : 0.f;
    const float d = x - oh;
    s0 += (mode == 1) ? d * d : fabsf(d);
    if (mode == 2) {
      const bool hb = h + 1 < a.H, wb = w + 1 < a.W;
      const float xd = hb ? __ldg(px + a.sh) : 0.f, xr = wb ? __ldg(px + a.sw) : 0.f;
      const float od = (hb && pt[a.th] == c) ? 1.f : 0.f, orr = (wb && pt[a.tw] == c) ? 1.f : 0.f;
      s1 += fabsf(fabsf(od - oh) - fabsf(xd - x)) + fabsf(fabsf(oh - orr) - fabsf(x - xr));
    }
  }
  __shared__ double red[2][32];
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s0;
    red[1][threadIdx.x >> 5] = s1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) {
      t0 += red[0][k];
      t1 += red[1][k];
    }
    atomicAdd(out, t0);
    if (mode == 2) atomicAdd(out + 1, t1);
  }
}

// loss = out[0]/(B*C*H*W) [+ out[1]/(B*C)]
__global__ void onehot_loss_finalize_kernel(const double* __restrict__ sums, int mode, double inv_n, double inv_bc,
                                            float* __restrict__ loss) {
  double v = sums[0] * inv_n;
  if (mode == 2) v += sums[1] * inv_bc;
  *loss = (float)v;
}

// dx = g * ( pointwise'(d)/N  [+ gdl'/ (B*C)] ); torch's abs backward is sign() with sign(0) = 0.
__global__ void onehot_loss_bwd_kernel(OhArgs a, int mode, const float* __restrict__ gptr, float inv_n, float inv_bc,
                                       float* __restrict__ dx, long long db, long long dc, long long dh, long long dw) {
  const long long total = (long long)a.B * a.C * a.H * a.W;
  const float g = *gptr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % a.W);
    long long t = i / a.W;
    const int h = (int)(t % a.H);
    t /= a.H;
    const int c = (int)(t % a.C);
    const int b = (int)(t / a.C);
    const float* px = a.x + b * a.sb + c * a.sc + h * a.sh + w * a.sw;
    const long long* pt = a.tgt + b * a.tb + h * a.th + w * a.tw;
    const float x = __ldg(px);
    const float oh = (*pt == c) ? 1.f : 0.f;
    const float d = x - oh;
    float r = ((mode == 1) ? 2.f * d : sgn(d)) * inv_n;
    if (mode == 2) {
      const bool hb = h + 1 < a.H, wb = w + 1 < a.W;
      const float xd = hb ? __ldg(px + a.sh) : 0.f, xr = wb ? __ldg(px + a.sw) : 0.f;
      const float od = (hb && pt[a.th] == c) ? 1.f : 0.f, orr = (wb && pt[a.tw] == c) ? 1.f : 0.f;
      // vertical term of this row: e = x[h+1]-x[h];  d/dx[h] = sign(a-|e|) * sign(e)
      float ev = xd - x, eh = x - xr;
      float q = sgn(fabsf(od - oh) - fabsf(ev)) * sgn(ev) - sgn(fabsf(oh - orr) - fabsf(eh)) * sgn(eh);
      if (h > 0) {  // vertical term of the row above: e = x[h]-x[h-1];  d/dx[h] = -sign(a-|e|) * sign(e)
        const float xu = __ldg(px - a.sh);
        const float ou = (pt[-a.th] == c) ? 1.f : 0.f;
        ev = x - xu;
        q -= sgn(fabsf(oh - ou) - fabsf(ev)) * sgn(ev);
      }
      if (w > 0) {  // horizontal term of the column to the left: e = x[w-1]-x[w];  d/dx[w] = +sign(a-|e|) * sign(e)
        const float xl = __ldg(px - a.sw);
        const float ol = (pt[-a.tw] == c) ? 1.f : 0.f;
        eh = xl - x;
        q += sgn(fabsf(ol - oh) - fabsf(eh)) * sgn(eh);
      }
      r += q * inv_bc;
    }
    dx[b * db + c * dc + h * dh + w * dw] = g * r;
  }
}

}  // namespace pmoe

using namespace pmoe;

extern "C" {

int pmoe_onehot_loss_fwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                         int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode, double* sums2,
                         float* loss_out, pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !sums2 || !loss_out || B < 1 || C < 1 || H < 1 || W < 1 || mode < 0 || mode > 2) {
    set_error("onehot_loss_fwd: bad arguments (mode 0 = L1, 1 = MSE, 2 = L1 + gradient difference)");
    return PMOE_ERR_ARG;
  }
  cudaError_t e = cudaMemsetAsync(sums2, 0, 2 * sizeof(double), stream);
  if (e != cudaSuccess) {
    set_error("onehot_loss_fwd: memset failed: %s", cudaGetErrorString(e));
    return PMOE_ERR_LAUNCH;
  }
  OhArgs a{logits, sb, sc, sh, sw, (const long long*)target, tb, th, tw, B, C, H, W};
  const long long total = (long long)B * C * H * W;
  long long bl = (total + 255) / 256;
  if (bl > (long long)num_sms() * 8) bl = (long long)num_sms() * 8;
  onehot_loss_fwd_kernel<<<(unsigned)bl, 256, 0, stream>>>(a, mode, sums2);
  int rc = check_launch("onehot_loss_fwd");
  if (rc) return rc;
  onehot_loss_finalize_kernel<<<1, 1, 0, stream>>>(sums2, mode, 1.0 / (double)total, 1.0 / ((double)B * C), loss_out);
  return check_launch("onehot_loss_finalize");
}

int pmoe_onehot_loss_bwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                         int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode,
                         const float* grad_scale_dev, float* dlogits, int64_t db, int64_t dc, int64_t dh, int64_t dw,
                         pmoe_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!logits || !target || !grad_scale_dev || !dlogits || mode < 0 || mode > 2) {
    set_error("onehot_loss_bwd: bad arguments");
    return PMOE_ERR_ARG;
  }
  OhArgs a{logits, sb, sc, sh, sw, (const long long*)target, tb, th, tw, B, C, H, W};
  const long long total = (long long)B * C * H * W;
  long long bl = (total + 255) / 256;
  if (bl > (long long)num_sms() * 8) bl = (long long)num_sms() * 8;
  onehot_loss_bwd_kernel<<<(unsigned)bl, 256, 0, stream>>>(a, mode, grad_scale_dev, (float)(1.0 / (double)total),
                                                           (float)(1.0 / ((double)B * C)), dlogits, db, dc, dh, dw);
  return check_launch("onehot_loss_bwd");
}

}  // extern "C"
