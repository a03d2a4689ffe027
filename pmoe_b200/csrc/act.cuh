// Activation functions of the reference modules in one place: ReLU / ELU / Tanh / Sigmoid (make_mlp, basics.py:11-45; conv3,
// basics.py:48-59) and the MobileNet family's ReLU6 / Hardswish / Hardsigmoid (backbone.py:75-104 -> torchvision MobileNetV2/V3).
#pragma once
#include "../../include/pmoe_b200.h"

namespace pmoe {

__device__ __forceinline__ float relu6_f(float x) { return fminf(fmaxf(x, 0.f), 6.f); }
// torch: hardswish(x) = x * relu6(x + 3) / 6, hardsigmoid(x) = relu6(x + 3) / 6
__device__ __forceinline__ float hswish_f(float x) { return x * relu6_f(x + 3.f) * (1.f / 6.f); }
__device__ __forceinline__ float hsigmoid_f(float x) { return relu6_f(x + 3.f) * (1.f / 6.f); }

// the piecewise-linear activations (the smooth ones stay where their callers chose the exp flavour)
__device__ __forceinline__ bool act_is_piecewise(int act) {
  return act == PMOE_ACT_RELU || act == PMOE_ACT_RELU6 || act == PMOE_ACT_HSWISH || act == PMOE_ACT_HSIGMOID;
}
__device__ __forceinline__ float act_piecewise(float x, int act) {
  switch (act) {
    case PMOE_ACT_RELU: return fmaxf(x, 0.f);
    case PMOE_ACT_RELU6: return relu6_f(x);
    case PMOE_ACT_HSWISH: return hswish_f(x);
    case PMOE_ACT_HSIGMOID: return hsigmoid_f(x);
    default: return x;
  }
}
// d act(u) / du from the PRE-activation u (ATen's backward formulas: hardswish_backward uses (2u+3)/6 on (-3, 3), 0 below,
// 1 above; hardsigmoid_backward 1/6 on (-3, 3); hardtanh_backward (ReLU6) 1 on (0, 6))
__device__ __forceinline__ float act_grad_pre(float u, int act) {
  switch (act) {
    case PMOE_ACT_RELU: return u > 0.f ? 1.f : 0.f;
    case PMOE_ACT_RELU6: return (u > 0.f && u < 6.f) ? 1.f : 0.f;
    case PMOE_ACT_HSWISH: return u <= -3.f ? 0.f : (u < 3.f ? (2.f * u + 3.f) * (1.f / 6.f) : 1.f);  // open interval, as ATen's CUDA kernel
    case PMOE_ACT_HSIGMOID: return (u > -3.f && u < 3.f) ? (1.f / 6.f) : 0.f;
    default: return 1.f;
  }
}

}  // namespace pmoe
