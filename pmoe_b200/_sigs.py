"""argtypes of the C-ABI entry points other than pmoe_conv_tc (see include/pmoe_b200.h)."""
import ctypes as C

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

SIGS = {
    "pmoe_conv_tc_set_debug": [vp],
    "pmoe_nchw_to_nhwc": [vp, i64, i64, i64, i64, i32, vp, i32, vp],
    "pmoe_nhwc_to_nchw": [vp, i32, i32, vp, i64, i64, i64, i64, vp],
    "pmoe_maxpool": [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp],
    "pmoe_maxpool_idx": [vp, vp, i32, i32, i32, i32, vp, vp],
    "pmoe_maxpool_bwd_idx": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "pmoe_eca_gate": [vp, i64, i32, f32, vp, i32, i32, i32, i32, vp, i64, vp],
    "pmoe_scale_channels": [vp, vp, i32, vp, i64, vp],
    "pmoe_channel_sums": [vp, i32, vp, i64, vp],
    "pmoe_channel_stats": [vp, i32, vp, vp, vp],
    "pmoe_bn_finalize": [vp, vp, f32, i32, i32, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp],
    "pmoe_affine_act": [vp, vp, i32, vp, vp, vp, i32, vp],
    "pmoe_conv_simt": [vp, i32, vp],
    "pmoe_conv_wgrad_simt": [vp, i32, vp, vp],
    "pmoe_conv_wgrad_tc": [vp, vp, vp],
    "pmoe_bn_bwd_reduce": [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp],
    "pmoe_bn_bwd_apply": [vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, f32, i32, vp, vp, i32, vp],
    "pmoe_maxpool_bwd": [vp, vp, vp, i32, i32, i32, i32, i32, vp],
    "pmoe_prod_channel_sums": [vp, vp, i32, vp, i64, vp],
    "pmoe_eca_gate_bwd": [vp, i64, vp, i64, vp, i64, i32, f32, vp, i32, i32, i32, i32, vp, i64, vp, vp],
    "pmoe_eca_bwd_apply": [vp, i32, vp, i64, vp, i64, vp, i32, vp],
    "pmoe_axpy": [vp, vp, i32, f32, vp, i64, i32, vp],
    "pmoe_gate_mixture_fwd": [vp, i64, i64, vp, i64, i64, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "pmoe_gate_mixture_bwd": [vp, vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, vp, vp, i64, i64, vp],
    "pmoe_moe_loss": [vp, vp, vp, vp, i32, vp, vp, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp],
    "pmoe_dropout": [vp, vp, i32, i64, f32, C.c_uint64, vp],
    "pmoe_l1_mse": [vp, vp, i64, i32, f32, vp, vp, vp],
    "pmoe_mt_sqnorm": [vp, i32, vp, vp],
    "pmoe_mt_clip": [vp, i32, vp, f32, vp],
    "pmoe_mt_adam": [vp, i32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, i32, i32, vp, f32, vp],
    "pmoe_segloss_fwd": [vp, i64, i64, i64, i64, vp, i64, i64, i64, i32, i32, i32, i32, f32, f32, vp, vp, vp],
    "pmoe_segloss_bwd": [vp, i64, i64, i64, i64, vp, i64, i64, i64, i32, i32, i32, i32, f32, vp, vp, f32, vp, i64, i64, i64, i64, i32, vp],
}
