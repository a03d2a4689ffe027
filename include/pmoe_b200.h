/*
 * pmoe_b200 — C-ABI of the B200-native (sm_100a) hot path of mhnazeri/PMoE.
 *
 * The reference (pure PyTorch) has no FFI of its own: the interface each entry point replaces is
 * the ATen operator the reference module dispatches to. Citations are `PMoE/<file>:<line>` in
 * the reference tree. All pointers are DEVICE pointers owned by the caller (PyTorch's caching
 * allocator on the Python side); the library never allocates or frees device memory and keeps
 * no pointer past return. All launches go to the caller's stream; nothing synchronises.
 * Return value: 0 = ok, negative = error (see pmoe_last_error()). No C++ exceptions cross here.
 *
 * Layout convention: activations are NHWC ("pixel-major") with the channel axis contiguous,
 * described by a strided 4-D view so that channel slices (virtual concat, the PU-Net mask ring),
 * spatial parity views (stride-2 convs) and pixel-shuffle views (ConvTranspose2d k2s2) need no
 * copies.
 */
#ifndef PMOE_B200_H_
#define PMOE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pmoe_stream_t; /* cudaStream_t */

enum { PMOE_OK = 0, PMOE_ERR_ARG = -1, PMOE_ERR_UNSUPPORTED = -2, PMOE_ERR_LAUNCH = -3, PMOE_ERR_DRIVER = -4 };
enum { PMOE_F32 = 0, PMOE_BF16 = 1 };
enum { PMOE_ACT_NONE = 0, PMOE_ACT_RELU = 1, PMOE_ACT_ELU = 2, PMOE_ACT_TANH = 3, PMOE_ACT_SIGMOID = 4,
       PMOE_ACT_RELU6 = 5, PMOE_ACT_HSWISH = 6, PMOE_ACT_HSIGMOID = 7 }; /* 5-7: torchvision MobileNetV2/V3 (backbone.py:75-104) */

/* Strided NHWC view; strides in ELEMENTS, channel stride is 1. */
typedef struct PmoeView4 {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
} PmoeView4;

#define PMOE_MAX_SRC 6
#define PMOE_MAX_SEG 64

/* One K-segment of the implicit GEMM: `nchunks` channel chunks of `ck` channels starting at
 * channel c0 of source `src`, read at spatial offset (dh, dw) from the output pixel. */
typedef struct PmoeSeg {
  int8_t src, dh, dw, reserved;
  uint16_t c0, nchunks;
} PmoeSeg;

/* Tensor-core implicit-GEMM convolution / linear layer (bf16 in, fp32 accumulate in TMEM, bf16 out).
 * Replaces aten::convolution (+ folded/eval BatchNorm + ReLU + residual add) as used by
 * conv3 (model/blocks/basics.py:48-59), EfficientConvBlock (:80-135), UNet (model/blocks/unet.py:50-95),
 * torchvision BasicBlock (model/blocks/backbone.py:57-61), ConvTranspose2d k2s2 (unet.py:35-45) via
 * four pixel-shuffle output views, and aten::linear of make_mlp (basics.py:11-45) as a 1x1 conv.
 * out[n,h,w,co] = act( scale[co] * sum_seg sum_c src[seg.src][n,h+dh,w+dw,c0+c] * wpack[co,k] + shift[co] (+ residual) )
 * with zero padding outside the source view. wpack is [cout_pad][ktot] bf16, k enumerating
 * (segment, chunk, channel) in order. */
typedef struct PmoeConvTc {
  int32_t n_src;
  PmoeView4 src[PMOE_MAX_SRC];
  int32_t n_seg;
  PmoeSeg seg[PMOE_MAX_SEG];
  int32_t ck;       /* channel chunk: 16, 32 or 64 */
  int32_t ktot;     /* = ck * sum(nchunks) */
  int32_t cout_pad; /* rows of wpack; multiple of the N tile (16/32/64/128/256) */
  const void* wpack;
  PmoeView4 out; /* bf16; out.c = channels stored (<= cout_pad, multiple of 8) */
  const float* scale; /* [cout_pad] or NULL (=1) */
  const float* shift; /* [cout_pad] or NULL (=0) */
  int32_t act;
  PmoeView4 residual; /* bf16, ptr NULL = none; added before the activation */
  double* stat_sum;   /* optional [cout_pad]: += sum over valid pixels of the raw accumulator (fp64 accumulation) */
  double* stat_sqsum; /* optional [cout_pad]: += sum of squares (train-mode BN batch statistics)                 */
  float* pool_sum;    /* optional [n][pool_stride]: += per-image sum of the stored output (ECA / avgpool) */
  int32_t pool_stride; /* row stride of pool_sum in floats; 0 = cout_pad */
  /* Multi-view output (ConvTranspose2d k2s2 as ONE GEMM with N = 4*Cout): GEMM column n is stored to view n / out_cols
   * (view 0 = out, view i = out_extra[i-1]) at channel n % out_cols. n_out_extra = 0: single view. No residual /
   * statistics / pooling in this mode; scale/shift are indexed by the GEMM column. */
  int32_t n_out_extra;
  int32_t out_cols;
  PmoeView4 out_extra[3];
  /* Optional fused nn.MaxPool2d(2,2) (unet.py:29) of the stored output: (n, H/2, W/2, >= out.c) bf16 view, H and W
   * even. PMOE_ERR_UNSUPPORTED if the geometry does not allow it (the caller then runs pmoe_maxpool). */
  PmoeView4 pool2_out;
  /* Optional fp32 NCHW copy of output channels [0, nchw_c) (the module boundary hands fp32 NCHW logits back,
   * punet.py:118-120); strides in elements. */
  float* nchw_out;
  int64_t nchw_sn, nchw_sc, nchw_sh, nchw_sw;
  int32_t nchw_c;
  /* > 0: wpack holds one [cout_pad][ktot] weight set per image, this many elements apart (EfficientBlock's gate folded
   * into the next conv: pmoe_gate_weights; grouped expert layers). Not supported by the streamed-weights 3x3 kernel
   * (PMOE_ERR_UNSUPPORTED). */
  int64_t wpack_img_stride;
  /* > 0: shift has one [cout_pad] row per image, this many floats apart. With wpack_img_stride this makes the call a
   * GROUPED GEMM: source/out views carry one expert per "image", each with its own weights and bias (the K experts'
   * equally-shaped Linear layers of model/moe.py:53-72 in one launch). */
  int64_t shift_img_stride;
} PmoeConvTc;

int pmoe_conv_tc(const PmoeConvTc* desc, pmoe_stream_t stream);
/* Tuning aid: when set (device array of 16 x #CTAs uint64, zeroed by the caller), every pmoe_conv_tc launch records per-CTA
 * cycle counters: [0] producer waits for a free stage, [1] MMA waits for TMA data, [2] MMA waits for a free accumulator,
 * [3] MMA thread total, [4] epilogue waits for the accumulator, [5] epilogue total, [6] epilogue waits for its staging
 * buffer, [7] tiles. NULL switches it off (the default). */
int pmoe_conv_tc_set_debug(unsigned long long* counters_dev);

/* Same contract as pmoe_conv_tc on CUDA cores with fp32 FMA accumulation; dtype selects fp32 (the
 * <=1e-4 parity mode: activations, wpack and out are float) or bf16 storage. */
int pmoe_conv_simt(const PmoeConvTc* desc, int32_t dtype, pmoe_stream_t stream);
/* Weight gradient of the same descriptor (aten::convolution_backward, weight half): desc->out is read as
 * dy; dwpack[cout_pad][ktot] (fp32, packed K order of wpack) is ACCUMULATED into. */
int pmoe_conv_wgrad_simt(const PmoeConvTc* desc, int32_t dtype, float* dwpack, pmoe_stream_t stream);
/* Tensor-core weight gradient (bf16 dy / activations, fp32 accumulation in TMEM, fp32 atomics into dwpack). Same contract
 * as pmoe_conv_wgrad_simt with dtype = PMOE_BF16. Returns PMOE_ERR_UNSUPPORTED, launching nothing, when the descriptor is
 * outside what the tensor-core kernels cover (the caller then uses pmoe_conv_wgrad_simt). */
/* Grouped form: desc->wpack_img_stride > 0 makes every image of the dy / source views one GROUP whose gradient goes to
 * dwpack + image * wpack_img_stride (floats): the K experts' equally shaped Linear layers (model/moe.py:53-72) in ONE launch. */
int pmoe_conv_wgrad_tc(const PmoeConvTc* desc, float* dwpack, pmoe_stream_t stream);

/* ---- memory-bound kernels (eltwise.cu); dtype = PMOE_F32 | PMOE_BF16 of the NHWC views ------------- */
/* Module boundary: the reference hands fp32 NCHW tensors to forward() (model/moe.py:90-93, punet.py:88). */
int pmoe_nchw_to_nhwc(const float* src, int64_t sn, int64_t sc, int64_t sh, int64_t sw, int32_t c, const PmoeView4* dst,
                      int32_t dst_dtype, pmoe_stream_t stream);
int pmoe_nhwc_to_nchw(const PmoeView4* src, int32_t src_dtype, int32_t c, float* dst, int64_t dn, int64_t dc, int64_t dh,
                      int64_t dw, pmoe_stream_t stream);
/* nn.MaxPool2d(2,2) (unet.py:29) / MaxPool2d(3,2,1) (torchvision ResNet), optionally applying
 * relu(scale*x+shift) on load (eval-mode bn1+relu of the ResNet stem, backbone.py:63). */
int pmoe_maxpool(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, int32_t k, int32_t stride, int32_t pad,
                 const float* scale, const float* shift, int32_t relu, pmoe_stream_t stream);
/* Training-mode max-pool: also stores, per output element, the row-major window position (r*k+s) of the FIRST maximum
 * (ATen's tie rule) in idx_out, a dense (n, oh, ow, dst.c) uint8 tensor consumed by pmoe_maxpool_bwd_idx. */
int pmoe_maxpool_idx(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, int32_t k, int32_t stride, int32_t pad,
                     uint8_t* idx_out, pmoe_stream_t stream);
/* EfficientBlock (basics.py:62-77): gate = sigmoid(conv1d_k(mean_hw(x))) from per-image channel sums. */
int pmoe_eca_gate(const float* pool_sum, int64_t pool_stride, int32_t n, float inv_count, const float* w, int32_t k,
                  int32_t groups, int32_t group_c, int32_t group_stride, float* gate, int64_t gate_stride,
                  pmoe_stream_t stream);
/* out[n][co][k] = wpack[co][k] * gate[n][k % cphys] (bf16): the conv of x * gate[n, c] as a conv of x with per-image weights
 * (cphys = physical channels per tap of the packed K axis). Replaces the scale pass of EfficientBlock (basics.py:76). */
int pmoe_gate_weights(const void* wpack, const float* gate, int64_t gate_stride, int32_t n, int32_t cout_pad, int32_t ktot,
                      int32_t cphys, void* out, pmoe_stream_t stream);
int pmoe_scale_channels(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* gate, int64_t gate_stride,
                        pmoe_stream_t stream);
/* out[n][c] += sum over h,w (adaptive_avg_pool2d numerator, basics.py:72, unet.py:90). */
int pmoe_channel_sums(const PmoeView4* src, int32_t dtype, float* out, int64_t out_stride, pmoe_stream_t stream);
/* sum[c], sqsum[c] += over n,h,w (batch statistics of a BatchNorm not fed by a conv epilogue: ResNet bn1). */
int pmoe_channel_stats(const PmoeView4* src, int32_t dtype, double* sum, double* sqsum, pmoe_stream_t stream);
/* nn.BatchNorm2d training step (basics.py:52,55): batch stats from (sum, sumsq), running-stat update, fused affine. */
int pmoe_bn_finalize(const double* sum, const double* sqsum, float count, int32_t c, int32_t c_pad, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean, float* running_var, float* mean_out,
                     float* rstd_out, float* scale, float* shift, pmoe_stream_t stream);
/* y = act(scale[c]*x + shift[c] (+ residual)): BN apply + ReLU (+ BasicBlock residual add). */
int pmoe_affine_act(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* scale, const float* shift,
                    const PmoeView4* residual, int32_t act, pmoe_stream_t stream);
/* The same affine + activation on dense bf16 tensors, fused with the statistics of the STORED output that a following layer
 * needs (each would otherwise re-read the tensor): pool_sum (N, pool_stride) fp32 per-image channel sums (ECA gate /
 * avg-pool numerators, basics.py:71) and/or out_sum / out_sqsum per-channel fp64 sums for a BatchNorm that follows directly
 * (torchvision ResNet bn1 after the stem, backbone.py:57-61). Accumulates into zero-initialised buffers. Returns
 * PMOE_ERR_UNSUPPORTED for strided / fp32 views (callers then use pmoe_affine_act + pmoe_channel_sums / _stats). */
int pmoe_affine_act_stats(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, const float* scale, const float* shift,
                          int32_t act, float* pool_sum, int64_t pool_stride, double* out_sum, double* out_sqsum,
                          pmoe_stream_t stream);

/* ---- backward halves (eltwise_bwd.cu) ------------------------------------------------------------- */
/* dy = dz * act'(z); sum_dy[c] += sum dy, sum_dy_xhat[c] += sum dy*(x-mean)*rstd  (native_batch_norm_backward reductions).
 * fwd_scale/fwd_shift (optional): with act = RELU and a null z view the mask is recomputed as fmaf(x, fwd_scale[c],
 * fwd_shift[c]) > 0 — exactly what pmoe_affine_act evaluated in the forward — so the saved output is not read at all
 * (contiguous bf16 only; PMOE_ERR_UNSUPPORTED otherwise). */
int pmoe_bn_bwd_reduce(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                       const float* mean, const float* rstd, double* sum_dy, double* sum_dy_xhat, const float* fwd_scale,
                       const float* fwd_shift, pmoe_stream_t stream);
/* dx = gamma*rstd*(dy - sum_dy/N - xhat*sum_dy_xhat/N) (batch_stats) or dy*gamma (eval BN / plain);
 * dres (optional) receives the masked dy for a residual branch. fwd_scale/fwd_shift: as in pmoe_bn_bwd_reduce. */
/* param_grads (optional, batch-statistics mode): the two reductions ARE the BatchNorm affine gradients (d bias = sum_dy,
 * d weight = sum_dy_xhat); the kernel also stores them as fp32 into the parameters' gradient slots (first n channels;
 * accumulate: add to what is there), which saves a conversion launch per parameter. */
typedef struct PmoeBnParamGrads {
  float* dgamma;
  float* dbeta;
  int32_t n;
  int32_t accumulate;
} PmoeBnParamGrads;
int pmoe_bn_bwd_apply(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                      const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                      const double* sum_dy_xhat, float inv_n, int32_t batch_stats, const PmoeView4* dx,
                      const PmoeView4* dres, int32_t accumulate_dres, const float* fwd_scale, const float* fwd_shift,
                      const PmoeBnParamGrads* param_grads, pmoe_stream_t stream);
/* The same batch-statistics apply when x is itself the ReLU output of an UPSTREAM BatchNorm and dx is that layer's complete
 * gradient (torchvision's bn1 directly after the stem block, backbone.py:57-61): also accumulates next_sum_dx[c] += sum
 * dx*[x>0] and next_sum_dx_x[c] += sum dx*x (zero-initialised fp64), from which the upstream layer's two backward
 * reductions follow without a pass of its own: sum dy*m = next_sum_dx, sum dy*m*raw = (next_sum_dx_x - shift*next_sum_dx)
 * / scale with the upstream forward's (scale, shift). Dense bf16, ReLU, channel-group count dividing 256; otherwise
 * PMOE_ERR_UNSUPPORTED (nothing launched). */
int pmoe_bn_bwd_apply_sums(const PmoeView4* dz, const PmoeView4* z, const PmoeView4* x, int32_t dtype, int32_t act,
                           const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                           const double* sum_dy_xhat, float inv_n, const PmoeView4* dx, const float* fwd_scale,
                           const float* fwd_shift, double* next_sum_dx, double* next_sum_dx_x, const PmoeBnParamGrads* param_grads,
                           pmoe_stream_t stream);
/* The tail of the ResNet stem in training, torchvision's bn1 -> relu -> MaxPool2d(3, 2, 1) behind conv1 := EfficientConvBlock
 * (reference PMoE/model/blocks/backbone.py:57-61; replaces pmoe_affine_act + pmoe_maxpool_idx | pmoe_maxpool_bwd_idx +
 * pmoe_bn_bwd_reduce + pmoe_bn_bwd_apply_sums on that path). Dense bf16 NHWC, even H and W, channel-group count dividing 256;
 * otherwise PMOE_ERR_UNSUPPORTED and the caller runs the separate entry points.
 *   fwd:    y = maxpool(relu(scale*x + shift)) with (scale, shift) of pmoe_bn_finalize; idx: argmax codes r*3+c as pmoe_maxpool_idx
 *           writes them, x_at_max: x at the argmax, dense bf16 (n, oh, ow, c) — the normalised tensor is never stored.
 *   reduce: sum_dy / sum_dy_xhat of pmoe_bn_bwd_reduce for the gradient routed through pool and ReLU, from the POOLED grid
 *           (both must be zeroed by the caller); sum_dy_xpos (optional, zeroed): the part of sum_dy routed to positions whose
 *           INPUT x is > 0 (x being a ReLU output itself: what the upstream BatchNorm's closed-form backward needs).
 *   apply:  dx of pmoe_bn_bwd_apply for every input pixel (the routed gradient is gathered from the windows that hold the pixel,
 *           never stored); next_sum_dx / next_sum_dx_x (optional, zeroed by the caller) and param_grads as in
 *           pmoe_bn_bwd_apply_sums. inv_n = 1 / (n*h*w) of x. */
int pmoe_bn_relu_maxpool_fwd(const PmoeView4* x, const float* scale, const float* shift, const PmoeView4* y, uint8_t* idx,
                             void* x_at_max, pmoe_stream_t stream);
int pmoe_bn_relu_maxpool_bwd_reduce(const PmoeView4* dy, const void* x_at_max, const float* fwd_scale, const float* fwd_shift,
                                    const float* mean, const float* rstd, double* sum_dy, double* sum_dy_xhat, double* sum_dy_xpos,
                                    pmoe_stream_t stream);
int pmoe_bn_relu_maxpool_bwd_apply(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* x, const float* fwd_scale,
                                   const float* fwd_shift, const float* mean, const float* rstd, const float* gamma,
                                   const double* sum_dy, const double* sum_dy_xhat, float inv_n, const PmoeView4* dx,
                                   double* next_sum_dx, double* next_sum_dx_x, const PmoeBnParamGrads* param_grads,
                                   pmoe_stream_t stream);
/* pmoe_bn_relu_maxpool_bwd_apply continued through the conv + BatchNorm + ReLU in front of it (the stem block's conv2 + BN + ReLU
 * followed by torchvision's bn1 -> relu -> maxpool, reference backbone.py:57-61 / basics.py:122-125): x = relu(up_scale*raw + up_shift)
 * is recomputed from the raw conv output exactly as the forward stored it, dx stays in registers and what is written is
 * draw = up_gamma*up_rstd*([x > 0]*dx - up_sum_dy/N - xhat_up*up_sum_dy_xhat/N): the gradient of the RAW conv output. The upstream
 * sums are inputs: the caller forms them in closed form from pmoe_bn_relu_maxpool_bwd_reduce's three sums and the forward statistics
 * of x (pmoe_affine_relu_stats_pos): sum dx*[x>0] = A*sum_dy_xpos + B*sum x + C*count(x>0), sum dx*x = A*sum dy*x + B*sum x^2 + C*sum x
 * with dx = A*d + B*x + C per channel. Replaces pmoe_bn_relu_maxpool_bwd_apply + pmoe_bn_bwd_apply (5.4 tensor passes) by 2.4.
 * param_grads / up_param_grads: the affine gradients of the two BatchNorms (optional). Dense bf16, channel groups dividing 128. */
int pmoe_bn2_relu_maxpool_bwd_apply(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* raw, const float* fwd_scale,
                                    const float* fwd_shift, const float* mean, const float* rstd, const float* gamma, const double* sum_dy,
                                    const double* sum_dy_xhat, float inv_n, const PmoeBnParamGrads* param_grads, const float* up_scale,
                                    const float* up_shift, const float* up_mean, const float* up_rstd, const float* up_gamma,
                                    const double* up_sum_dy, const double* up_sum_dy_xhat, const PmoeBnParamGrads* up_param_grads,
                                    const PmoeView4* draw, pmoe_stream_t stream);
/* pmoe_affine_act_stats for y = relu(scale*x + shift) that also counts the stored values > 0 per channel (out_pos, fp64, zeroed by
 * the caller). Dense bf16. */
int pmoe_affine_relu_stats_pos(const PmoeView4* src, const PmoeView4* dst, const float* scale, const float* shift, double* out_sum,
                               double* out_sqsum, double* out_pos, pmoe_stream_t stream);
int pmoe_maxpool_bwd(const PmoeView4* x, const PmoeView4* dy, const PmoeView4* dx, int32_t dtype, int32_t k, int32_t stride,
                     int32_t pad, int32_t accumulate, pmoe_stream_t stream);
int pmoe_maxpool_bwd_idx(const PmoeView4* dy, const uint8_t* idx, const PmoeView4* dx, int32_t dtype, int32_t k,
                         int32_t stride, int32_t pad, int32_t accumulate, pmoe_stream_t stream);
int pmoe_prod_channel_sums(const PmoeView4* a, const PmoeView4* b, int32_t dtype, double* out, int64_t out_stride,
                           pmoe_stream_t stream);
int pmoe_eca_gate_bwd(const double* dgate, int64_t dgate_stride, const float* gate, int64_t gate_stride, const float* pool_sum,
                      int64_t pool_stride, int32_t n, float inv_count, const float* w, int32_t k, int32_t groups,
                      int32_t group_c, int32_t group_stride, float* dmean, int64_t dmean_stride, double* dw,
                      pmoe_stream_t stream);
int pmoe_eca_bwd_apply(const PmoeView4* dout, int32_t dtype, const float* gate, int64_t gate_stride, const float* dmean,
                       int64_t dmean_stride, const PmoeView4* dx, int32_t accumulate, pmoe_stream_t stream);
/* The same pass when the ECA block's forward input x_fwd is the ReLU output of an upstream BatchNorm and dx its complete
 * gradient (EfficientConvBlock's second gate, basics.py:118-121): also accumulates that layer's backward sums, as
 * pmoe_bn_bwd_apply_sums does. Dense bf16, fresh dx, channel-group count dividing 256; otherwise PMOE_ERR_UNSUPPORTED. */
int pmoe_eca_bwd_apply_sums(const PmoeView4* dout, int32_t dtype, const float* gate, int64_t gate_stride, const float* dmean,
                            int64_t dmean_stride, const PmoeView4* dx, const PmoeView4* x_fwd, double* next_sum_dx,
                            double* next_sum_dx_x, pmoe_stream_t stream);
/* EfficientConvBlock's second gate in training (reference PMoE/model/blocks/basics.py:118-121: conv1 -> BN -> ReLU -> ECA gate ->
 * conv2) without ever storing the gradient of the gated tensor's input: d c1 = dy * gate[n, c] + dmean[n, c] is affine in the data
 * gradient dy of conv2 per (image, channel). Replaces pmoe_prod_channel_sums + pmoe_eca_bwd_apply_sums + pmoe_bn_bwd_apply (8 tensor
 * passes) by 5. Dense bf16 NHWC; otherwise PMOE_ERR_UNSUPPORTED and the caller runs the separate entry points.
 *   sums:  per (image, channel), accumulated into zeroed fp64 (n, out_stride) arrays: sum_dy_m = sum dy * [c1 > 0],
 *          sum_dy_c1 = sum dy * c1 (the gate's gradient, input of pmoe_eca_gate_bwd), sum_m = sum [c1 > 0]. The BatchNorm's two
 *          backward sums follow as sum_n gate*sum_dy_m + dmean*sum_m and sum_n gate*sum_dy_c1 + dmean*pool_sum (then through the
 *          forward affine as for pmoe_bn_bwd_apply_sums' outputs).
 *   apply: draw = gamma*rstd*(dz - sum_dy/N - xhat*sum_dy_xhat/N) with dz = [relu mask of raw] * (dy*gate + dmean), the ReLU mask
 *          recomputed from raw with the forward's (fwd_scale, fwd_shift); param_grads as in pmoe_bn_bwd_apply_sums.
 *          inv_n = 1 / (n*h*w). */
int pmoe_eca_bn_bwd_sums(const PmoeView4* dy, const PmoeView4* c1, double* sum_dy_m, double* sum_dy_c1, double* sum_m,
                         int64_t out_stride, pmoe_stream_t stream);
int pmoe_eca_bn_bwd_apply(const PmoeView4* dy, const PmoeView4* raw, const float* gate, int64_t gate_stride, const float* dmean,
                          int64_t dmean_stride, const float* fwd_scale, const float* fwd_shift, const float* mean, const float* rstd,
                          const float* gamma, const double* sum_dy, const double* sum_dy_xhat, float inv_n, const PmoeView4* dx,
                          const PmoeBnParamGrads* param_grads, pmoe_stream_t stream);
/* Backward of Linear -> activation without BatchNorm (make_mlp stacks of the expert heads, reference PMoE/model/blocks/basics.py:11-45,
 * model/moe.py:88-101): dy = dz * act'(z) from the saved OUTPUT z (act: none / ReLU / ELU) fused with the bias gradient, the per-image
 * channel sum of dy accumulated into the zeroed fp32 (n, bias_stride) buffer (image = expert in the grouped heads). Replaces
 * pmoe_bn_bwd_apply (no statistics) + pmoe_channel_sums. Dense bf16; otherwise PMOE_ERR_UNSUPPORTED. */
int pmoe_act_bwd_bias(const PmoeView4* dz, const PmoeView4* z, int32_t act, const PmoeView4* dy, float* bias_sum, int64_t bias_stride,
                      pmoe_stream_t stream);
/* dst (+)= alpha*src + bcast[n][c]: gradient accumulation and global-avg-pool backward. */
int pmoe_axpy(const PmoeView4* src, const PmoeView4* dst, int32_t dtype, float alpha, const float* bcast, int64_t bcast_stride,
              int32_t accumulate, pmoe_stream_t stream);

/* ---- MoE gating / mixture head and losses (heads.cu, segloss.cu) ---------------------------------- */
/* softmax_K(alpha) (after ReLU for BaseExpert, model/moe.py:100,150-151), sigma = ELU(raw)+1 (moe.py:99), routing
 * index = argmax_k. alpha/ap are the strided raw head outputs (element (b,k) at b*sb + k*sk). */
int pmoe_gate_mixture_fwd(const void* alpha, int64_t a_sb, int64_t a_sk, const void* ap, int64_t p_sb, int64_t p_sk,
                          int32_t dtype, int32_t B, int32_t K, int32_t relu_alpha, float* probs, float* mean, float* std,
                          int64_t* route, pmoe_stream_t stream);
int pmoe_gate_mixture_bwd(const float* dprobs, const float* dmean, const float* dstd, const float* probs, const float* std,
                          const void* alpha, int64_t a_sb, int64_t a_sk, int32_t dtype, int32_t B, int32_t K,
                          int32_t relu_alpha, void* dalpha, void* dap, int64_t p_sb, int64_t p_sk, pmoe_stream_t stream);
/* moe_loss (trainer/loss.py:121-132): c0*NLL of the Gaussian mixture + c1*speed MSE(/K), with gradients. loss_out[3]
 * (total, nll, speed) must be zeroed by the caller. */
int pmoe_moe_loss(const float* probs, const float* mean, const float* std, const float* speed_pred, int32_t speed_k,
                  const float* act_gt, const float* speed_gt, int32_t B, int32_t K, float c0, float c1, float* loss_out,
                  float* dprobs, float* dmean, float* dstd, float* dspeed, float* logp_out, pmoe_stream_t stream);
/* nn.Dropout (basics.py:39-40) with a stateless counter-based mask; call again on the gradient for backward. */
int pmoe_dropout(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t seed, pmoe_stream_t stream);
/* Same mask generator seeded from DEVICE memory (*seed_dev mixed with the per-layer salt): a training step captured in a
 * CUDA graph draws fresh masks on every replay when the caller bumps the counter between replays. */
int pmoe_dropout_dev(const void* x, void* y, int32_t dtype, int64_t n, float p, uint64_t salt, const uint64_t* seed_dev,
                     pmoe_stream_t stream);
/* nn.L1Loss / nn.MSELoss (loss.py:135-151): *loss += coef*mean(...), da = gradient. */
int pmoe_l1_mse(const float* a, const float* b, int64_t n, int32_t is_mse, float coef, float* loss, float* da,
                pmoe_stream_t stream);
/* cross_entropy_tversky_weighted_loss (loss.py:47-55) fused: one read forward, one read + one write backward. */
size_t pmoe_segloss_workspace_floats(int32_t C, int32_t W);
int pmoe_segloss_fwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                     int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, float wce, float wtv, float* workspace,
                     float* loss_out, pmoe_stream_t stream);
int pmoe_segloss_bwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                     int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, float wce, const float* workspace,
                     const float* grad_scale_dev, float grad_scale, float* dlogits, int64_t db, int64_t dc, int64_t dh,
                     int64_t dw, int32_t accumulate, pmoe_stream_t stream);

/* Losses against the one-hot of an int64 class map (onehot_loss.cu). mode 0: nn.L1Loss, 1: nn.MSELoss — the 'l1' / 'l2'
 * variants of AutoregressiveCriterion (trainer/loss.py:93-96,109-116, which scatter_ a one-hot tensor first); mode 2: the
 * last-frame L1 + gradient-difference terms of l1_gdl (loss.py:58-83). sums2: two doubles of scratch; *loss_out is written.
 * The backward writes dlogits = *grad_scale_dev * d loss / d logits (any strides). */
int pmoe_onehot_loss_fwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                         int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode, double* sums2,
                         float* loss_out, pmoe_stream_t stream);
int pmoe_onehot_loss_bwd(const float* logits, int64_t sb, int64_t sc, int64_t sh, int64_t sw, const int64_t* target, int64_t tb,
                         int64_t th, int64_t tw, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode,
                         const float* grad_scale_dev, float* dlogits, int64_t db, int64_t dc, int64_t dh, int64_t dw,
                         pmoe_stream_t stream);

/* ---- optimizer side (optim.cu): multi-tensor kernels over a DEVICE table of chunks ------------------------- */
/* One chunk = up to 2^31-1 consecutive fp32 elements of one parameter with its gradient and Adam state. */
typedef struct PmoeMtChunk {
  float* p;    /* parameter            (mt_adam) */
  float* g;    /* gradient             (all)     */
  float* m;    /* exp_avg              (mt_adam) */
  float* v;    /* exp_avg_sq           (mt_adam) */
  float* vmax; /* max_exp_avg_sq       (mt_adam with amsgrad) */
  int32_t n;
  int32_t pad;
} PmoeMtChunk;
/* *sqnorm_accum += sum g^2 over all chunks: the global gradient norm of check_grad_norm (utils/nn.py:10-19) and of
 * torch.nn.utils.clip_grad_norm_ (trainer/train_2.py:160-161) in one launch and no host sync. */
int pmoe_mt_sqnorm(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double* sqnorm_accum, pmoe_stream_t stream);
/* g *= min(1, max_norm / (sqrt(*sqnorm) + 1e-6)) — clip_grad_norm_'s scaling, coefficient computed on the device. */
int pmoe_mt_clip(const PmoeMtChunk* chunks_dev, int32_t n_chunks, const double* sqnorm, float max_norm, pmoe_stream_t stream);
/* torch.optim.Adam(amsgrad) step (conf/stage_2.yaml:137-144, train_2.py:165) over all chunks; when sqnorm != NULL the
 * clip coefficient is applied to the gradient on the fly (gradients themselves are left unscaled). */
int pmoe_mt_adam(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double lr, double beta1, double beta2, double eps,
                 double weight_decay, int32_t step, int32_t amsgrad, const double* sqnorm, float max_norm, pmoe_stream_t stream);

/* torch.optim.RMSprop step (conf/stage_2.yaml:147-153, train_2.py:67-71) over all chunks. Chunk fields: m = square_avg,
 * v = momentum_buffer (NULL when momentum == 0), vmax = grad_avg (NULL unless centered); sqnorm/max_norm as in pmoe_mt_adam. */
int pmoe_mt_rmsprop(const PmoeMtChunk* chunks_dev, int32_t n_chunks, double lr, double alpha, double eps, double weight_decay,
                    double momentum, const double* sqnorm, float max_norm, pmoe_stream_t stream);
/* torch.optim.swa_utils.AveragedModel.update_parameters, default avg_fn (train_2.py:119-121,179-187):
 * chunk.p (averaged) += (chunk.g (current model parameter) - chunk.p) / (n_averaged + 1). */
int pmoe_mt_swa_update(const PmoeMtChunk* chunks_dev, int32_t n_chunks, int64_t n_averaged, pmoe_stream_t stream);

/* ---- depthwise convolutions (depthwise.cu): torchvision MobileNetV2 / V3 inverted-residual blocks, the alternative backbones
 * of the reference's factory (model/blocks/backbone.py:75-104 -> nn.Conv2d(c, c, k, stride, (k-1)/2, groups=c, bias=False)).
 * Views are NHWC (dtype PMOE_F32 | PMOE_BF16, channels a multiple of 8); w_packed is [k*k][w_cpad] fp32, tap-major
 * (w_packed[(r*k+s)*w_cpad + c] = weight[c, 0, r, s]), produced by pmoe_pack_gather. */
int pmoe_dwconv_fwd(const PmoeView4* x, const float* w_packed, int32_t w_cpad, const PmoeView4* y, int32_t dtype, int32_t k,
                    int32_t stride, int32_t pad, pmoe_stream_t stream);
/* dx (+)= conv_transpose of dy with the same weights (aten::convolution_backward, input half). */
int pmoe_dwconv_dgrad(const PmoeView4* dy, const float* w_packed, int32_t w_cpad, const PmoeView4* dx, int32_t dtype, int32_t k,
                      int32_t stride, int32_t pad, int32_t accumulate, pmoe_stream_t stream);
/* dw_packed[k*k][w_cpad] (fp32, ACCUMULATED into) += sum over pixels of dy * shifted x (weight half). */
int pmoe_dwconv_wgrad(const PmoeView4* x, const PmoeView4* dy, float* dw_packed, int32_t w_cpad, int32_t dtype, int32_t k,
                      int32_t stride, int32_t pad, pmoe_stream_t stream);

/* ---- weight layout (pack.cu) ------------------------------------------------------------------------------------------
 * The conv / linear kernels read weights in a packed [cout_pad][K] operand layout; the nn.Parameter stays fp32 in the
 * reference's (out, in, kh, kw) layout (model/blocks/basics.py:51,54). `idx` is a static int32 map of the
 * packed layout: packed element i = parameter element idx[i], -1 = zero padding. */
/* out[i] = idx[i] >= 0 ? w[idx[i]] : 0, stored as out_dtype (PMOE_F32 | PMOE_BF16). Runs after every optimizer step. */
int pmoe_pack_gather(const float* w, const int32_t* idx, void* out, int32_t out_dtype, int64_t n, pmoe_stream_t stream);
typedef struct PmoePackJob {
  const float* w;
  const int32_t* idx;
  void* out;
  int64_t n;
  int32_t dtype;
  int32_t chunk0; /* first chunk (of pmoe_pack_chunk_elems() elements) of this job; jobs sorted by chunk0 */
} PmoePackJob;
/* The same for a device table of jobs in ONE launch (all packed operands of a model). */
int pmoe_pack_gather_mt(const PmoePackJob* jobs_dev, int32_t n_jobs, int32_t total_chunks, pmoe_stream_t stream);
int pmoe_pack_chunk_elems(void);
/* Weight gradient back to the parameter layout (aten::convolution_backward's grad_weight): dst[idx[i]] = alpha * packed[i]
 * (+ dst[idx[i]] when accumulate) for idx[i] >= 0; dst may be a slot of a flat all-reduce bucket. */
int pmoe_unpack_scatter(const float* packed, const int32_t* idx, float* dst, int64_t n, float alpha, int32_t accumulate,
                        pmoe_stream_t stream);
/* The same for up to 16 equally shaped parameters at once: packed is [n_groups][n] (the K experts' stacked gradients of one
 * layer), `dst` / `accumulate` are HOST arrays of n_groups device pointers / flags (a NULL pointer skips its group). */
int pmoe_unpack_scatter_group(const float* packed, const int32_t* idx, float* const* dst, const int32_t* accumulate,
                              int32_t n_groups, int64_t n, float alpha, pmoe_stream_t stream);
/* The same gradients through the INVERSE map (parameter element j = packed element inv[j], n_param entries): coalesced writes
 * into the gradient slots, gathered reads of the packed gradient ([n_groups][n_packed]). The form the training path uses. */
int pmoe_unpack_gather_group(const float* packed, const int32_t* inv, float* const* dst, const int32_t* accumulate,
                             int32_t n_groups, int64_t n_param, int64_t n_packed, float alpha, pmoe_stream_t stream);
/* dst[i] (+)= (float) src[i]: fp64 per-channel sums (BatchNorm weight/bias gradients, Linear bias gradients) into fp32 slots. */
int pmoe_cvt_f64_f32(const double* src, float* dst, int32_t n, int32_t accumulate, pmoe_stream_t stream);

/* ---- input pipeline (preproc.cu) — SURVEY.md §8f rank 1 ------------------------------------------------------- */
/* The reference dataset's eval-mode transform for every decoded frame, on the device and bit for bit:
 * Crop rows [crop_top, hs - crop_bottom) (augmenter.py:43-49) -> torchvision Resize((out_h, out_w)) on a PIL image = Pillow's
 * antialiased two-pass BILINEAR resample in 22-bit fixed point with a uint8 intermediate (data_loader.py:275-281) ->
 * ToTensor (/255, data_loader.py:281) -> frames stacked (data_loader.py:288-300).
 * src: (n, hs, ws, 3) uint8 RGB, contiguous. hbounds/hcoef (out_w x 2, out_w x hksize) and vbounds/vcoef are Pillow's
 * precompute_coeffs + normalize_coeffs_8bpc tables (built by pmoe_b200/preproc.py). lut255[v] = (float)v / 255.
 * tmp: (n, hs - crop_top - crop_bottom, out_w, 3) uint8 scratch. dst (optional): fp32, element (i, c, y, x) at
 * i*dst_sn + c*dst_sc + y*dst_sh + x*dst_sw (NCHW for the module API). dst_u8 (optional): (n, out_h, out_w, 3) uint8. */
int pmoe_preprocess_frames(const uint8_t* src, int32_t n, int32_t hs, int32_t ws, int32_t crop_top, int32_t crop_bottom,
                           int32_t out_h, int32_t out_w, const int32_t* hbounds, const int32_t* hcoef, int32_t hksize,
                           const int32_t* vbounds, const int32_t* vcoef, int32_t vksize, const float* lut255, uint8_t* tmp,
                           float* dst, int64_t dst_sn, int64_t dst_sc, int64_t dst_sh, int64_t dst_sw, uint8_t* dst_u8,
                           pmoe_stream_t stream);

/* Library info / errors. */
int pmoe_version(void);
const char* pmoe_last_error(void);
int pmoe_device_check(void); /* 0 iff device 0..current is sm_100 */

/* Debug probe used by tests/ and by the round-1 descriptor study (profiles/): one CTA computes
 * D[128,64] = A_view[128,64] * B[64,64]^T where A_view row m reads smem row
 * start_row + (m/8)*group_rows + (m%8) of a SWIZZLE_128B tile of `rows` rows loaded by TMA. */
int pmoe_dbg_umma_view(const void* a_bf16, int rows, const void* b_bf16, float* d_out, int start_row,
                       int group_rows, int base_offset_mode, pmoe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PMOE_B200_H_ */
