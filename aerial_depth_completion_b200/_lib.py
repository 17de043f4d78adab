"""ctypes binding of the C ABI declared in include/adc_b200.h (libadc_b200.so, built by csrc/Makefile).

The library is the only compute path: if it is missing this module raises at import, loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libadc_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "aerial_depth_completion_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C aerial_depth_completion_b200/csrc`). There is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

ADC_PYR_CALLS = 11
ADC_PYR_MODULES = 7
ADC_WSLOTS = 4

c_float_p = ctypes.c_void_p   # device pointers travel as integers
c_ll = ctypes.c_longlong
c_int = ctypes.c_int


class Src(ctypes.Structure):
    """adc_src_t"""
    _fields_ = [("x", ctypes.c_void_p), ("c", ctypes.c_void_p), ("batch_stride", c_ll),
                ("channels", c_int), ("H", c_int), ("W", c_int)]


class PyramidFwd(ctypes.Structure):
    """adc_pyramid_fwd_t"""
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("training", c_int), ("slot", c_int),
                ("x0", ctypes.c_void_p), ("c0", ctypes.c_void_p), ("x0_batch_stride", c_ll), ("c0_batch_stride", c_ll),
                ("weight", ctypes.c_void_p * ADC_PYR_MODULES), ("bias", ctypes.c_void_p * ADC_PYR_MODULES),
                ("weight_out", ctypes.c_void_p), ("saved", ctypes.c_void_p), ("idx", ctypes.c_void_p),
                ("xout", ctypes.c_void_p), ("cout", ctypes.c_void_p)]


class PyramidBwd(ctypes.Structure):
    """adc_pyramid_bwd_t"""
    _fields_ = [("N", c_int), ("H", c_int), ("W", c_int), ("slot", c_int), ("need_input_grad", c_int),
                ("need_weight_grad", c_int),
                ("x0", ctypes.c_void_p), ("c0", ctypes.c_void_p), ("x0_batch_stride", c_ll), ("c0_batch_stride", c_ll),
                ("weight_bwd", ctypes.c_void_p * ADC_PYR_MODULES), ("bias", ctypes.c_void_p * ADC_PYR_MODULES),
                ("saved", ctypes.c_void_p), ("idx", ctypes.c_void_p), ("xout", ctypes.c_void_p), ("cout", ctypes.c_void_p),
                ("g_xout", ctypes.c_void_p), ("g_cout", ctypes.c_void_p), ("scratch", ctypes.c_void_p),
                ("gx0", ctypes.c_void_p), ("gc0", ctypes.c_void_p), ("gweight", ctypes.c_void_p), ("gbias", ctypes.c_void_p)]


class ConvDesc(ctypes.Structure):
    """adc_conv_t"""
    _fields_ = [("kind", c_int),
                ("x", ctypes.c_void_p), ("x_lo", ctypes.c_void_p), ("x_stride", c_int), ("x_off", c_int), ("x_pad", c_int),
                ("w", ctypes.c_void_p), ("w_lo", ctypes.c_void_p), ("bias", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("out_lo", ctypes.c_void_p), ("out_f32", c_int), ("out_stride", c_int), ("out_off", c_int),
                ("out_pad", c_int), ("out_zero_to", c_int), ("accumulate", c_int),
                ("N", c_int), ("H", c_int), ("W", c_int), ("cin", c_int), ("cout", c_int), ("stride", c_int), ("act", c_int),
                ("tap_mask", c_int)]


ADC_WPREP_MAX = 40


class WPrepItem(ctypes.Structure):
    """adc_wprep_item_t"""
    _fields_ = [("w", ctypes.c_void_p), ("out", ctypes.c_void_p), ("out_lo", ctypes.c_void_p), ("cout", c_int), ("cin", c_int),
                ("ci_off", c_int), ("ci_cnt", c_int), ("cin_pad", c_int), ("mode", c_int)]


# every symbol include/adc_b200.h declares: name -> (restype, argtypes)
_vp, _sp = ctypes.c_void_p, ctypes.POINTER(Src)
SIGNATURES = {
    "adc_abi_version": (c_int, []),
    "adc_last_error": (ctypes.c_char_p, []),
    "adc_launch_count": (c_ll, []),
    "adc_profile_enable": (c_int, [c_int]),
    "adc_set_side_stream": (c_int, [c_int]),
    "adc_profile_read": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)]),
    "adc_nconv2d_fwd_scratch_floats": (ctypes.c_size_t, [_sp, _sp] + [c_int] * 4),
    "adc_nconv2d_fwd": (c_int, [_sp, _sp, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_nconv2d_bwd_scratch_floats": (ctypes.c_size_t, [_sp, _sp] + [c_int] * 4),
    "adc_nconv2d_bwd": (c_int, [_sp, _sp] + [_vp] * 14 + [c_int] * 6 + [_vp]),
    "adc_conf_pool_fwd": (c_int, [_vp] * 5 + [c_int] * 4 + [_vp]),
    "adc_conf_pool_bwd": (c_int, [_vp] * 7 + [c_int] * 4 + [_vp]),
    "adc_softplus_pow": (c_int, [_vp, _vp, c_int, c_int, _vp]),
    "adc_pyramid_weight_floats": (ctypes.c_size_t, []),
    "adc_pyramid_saved_floats": (ctypes.c_size_t, [c_int] * 3),
    "adc_pyramid_idx_count": (ctypes.c_size_t, [c_int] * 3),
    "adc_pyramid_idx_layout": (c_int, [c_int] * 3 + [ctypes.POINTER(c_ll), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "adc_pyramid_fwd": (c_int, [ctypes.POINTER(PyramidFwd), _vp]),
    "adc_pyramid_bwd_scratch_floats": (ctypes.c_size_t, [c_int] * 3),
    "adc_pyramid_bwd": (c_int, [ctypes.POINTER(PyramidBwd), _vp]),
    "adc_nchw_to_nhwc_bf16": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_ll, c_int, c_int, c_int, _vp]),
    "adc_nhwc_to_nchw_f32": (c_int, [_vp, c_int, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_conv3x3_weight_elems": (ctypes.c_size_t, [c_int, c_int]),
    "adc_conv3x3_prep_weights": (c_int, [_vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_conv3x3_prep_weights_batch": (c_int, [_vp, c_int, _vp]),
    "adc_conv3x3_fprop_tc": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_conv3x3_s1_tc_v2": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_conv3x3_dgrad_s2_tc": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int, _vp]),
    "adc_conv3x3_wgrad_box_tc": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, _vp]),
    "adc_conv1x1_wgrad_box_tc": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_int, _vp]),
    "adc_sum_rows": (c_int, [_vp, c_int, c_int, _vp, _vp]),
    "adc_nhwc_upsample_nearest": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_nhwc_grad_assemble": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, _vp, _vp, c_int, c_int, _vp, c_int, c_int, ctypes.c_float,
                                       _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp, c_int, _vp]),
    "adc_nhwc_grad_assemble_parts": (c_int, []),
    "adc_nhwc_colsum_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "adc_nhwc_colsum": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp, _vp, _vp]),
    "adc_ged_head_fwd": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_ged_head_bwd_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "adc_ged_head_bwd": (c_int, [_vp, _vp, c_int, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_int, _vp, _vp, c_int, _vp, _vp, _vp, _vp,
                                 c_int, c_int, c_int, c_int, _vp]),
    "adc_ged_head_bwd_x0d": (c_int, [_vp, _vp, _vp, c_int, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_ged_pack_im2col": (c_int, [_vp, _vp, _vp, c_ll, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_conv1x1_tc_v2": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, _vp]),
    "adc_conv1x1_wgrad_tc": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                     _vp]),
    "adc_ged_pack_input": (c_int, [_vp, _vp, _vp, c_ll, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_head_finish": (c_int, [_vp, c_int, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_head_bwd_prep_parts": (c_int, []),
    "adc_quad_head_bwd_prep": (c_int, [_vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_head_wgrad": (c_int, [_vp, _vp, c_int, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_head_dgrad": (c_int, [_vp, c_int, _vp, _vp, _vp, _vp, c_int, _vp, _vp, c_int, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_unshuffle": (c_int, [_vp, c_int, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_tap_mix": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_quad_tap_mix_adjoint": (c_int, [_vp, _vp, c_int, _vp]),
    "adc_conv3x3_wgrad_tc_workspace_bytes": (ctypes.c_size_t, [c_int, c_int]),
    "adc_conv3x3_wgrad_tc": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int,
                                     _vp]),
    "adc_conf_head_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "adc_conf_head_coef_floats": (c_int, []),
    "adc_conf_bn_stats": (c_int, [_vp, c_int, c_int, c_int, c_int, c_int, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, c_int, _vp,
                                  _vp, _vp]),
    "adc_conf_head_apply": (c_int, [_vp, c_int, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_conf_head_bwd": (c_int, [_vp, c_int, _vp, _vp, _vp, _vp, c_int, _vp, _vp, c_int, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_int, c_int,
                                  _vp]),
    "adc_loss_metrics_workspace_bytes": (ctypes.c_size_t, []),
    "adc_masked_loss_fwd": (c_int, [_vp, _vp, c_ll, c_int, _vp, _vp, _vp]),
    "adc_masked_loss_bwd": (c_int, [_vp, _vp, _vp, _vp, c_ll, c_int, _vp, _vp]),
    "adc_depth_metrics": (c_int, [_vp, _vp, _vp, ctypes.c_float, c_ll, _vp, _vp, _vp]),
    "adc_adam_step": (c_int, [_vp, _vp, _vp, _vp, _vp, c_ll] + [ctypes.c_double] * 5 + [_vp]),
    "adc_conv_tc": (c_int, [ctypes.POINTER(ConvDesc), _vp]),
    "adc_bn_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "adc_bn_coef_floats": (c_int, [c_int]),
    "adc_bn_stats": (c_int, [_vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, c_int, _vp, _vp,
                             _vp]),
    "adc_bn_apply": (c_int, [_vp, c_int, c_int, _vp, _vp, _vp, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_bn_bwd": (c_int, [_vp, c_int, c_int, _vp, c_int, c_int, _vp, c_int, c_int, _vp, c_int, c_int, _vp, _vp, c_int, c_int, _vp, c_int,
                           c_int, c_int, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_dot_head_fwd": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, c_int, c_int, c_int, c_int, c_int, _vp, _vp, c_int, c_int, _vp]),
    "adc_dot_head_bwd": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, _vp, c_int, c_int, c_int, _vp, _vp, _vp, c_int, c_int, c_int, c_int, _vp,
                                 c_int, _vp]),
    "adc_im2col3x3_pack": (c_int, [_vp, c_int, c_ll, _vp, c_int, c_ll, _vp, c_int, c_ll, _vp, _vp, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_nchw_to_nhwc_f32": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_adam_step_dev": (c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, c_ll, _vp]),
    "adc_im2col_kxk_pack": (c_int, [_vp, c_ll, c_int, c_int, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_col2im_kxk": (c_int, [_vp, c_int, c_int, c_int, c_int, c_int, _vp, c_int, c_int, c_int, _vp]),
    "adc_maxpool3x3s2_fwd": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, _vp, c_int, c_int, c_int, c_int, _vp]),
    "adc_maxpool3x3s2_bwd": (c_int, [_vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_pixel_shuffle_f32": (c_int, [_vp, c_int, c_int, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_pixel_unshuffle_pair": (c_int, [_vp, _vp, c_int, c_int, _vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_bilinear_ac_fwd": (c_int, [_vp, _vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _vp, c_int, c_int, _vp]),
    "adc_bilinear_ac_bwd": (c_int, [_vp, c_int, c_int, c_int, c_int, _vp, c_int, c_int, c_int, c_int, c_int, _vp]),
    "adc_uniform_sampling": (c_int, [_vp, _vp, c_int, ctypes.c_float, c_int, _vp, c_int, c_ll, _vp]),
    "adc_sample_assemble": (c_int, [_vp, _vp, c_int, ctypes.c_float, c_int, _vp, c_int, c_ll, _vp]),
    "adc_conf_pr_histogram": (c_int, [_vp, _vp, _vp, c_ll, _vp, c_int, _vp, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here = the library does not export what the header declares
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc, what):
    """Convert a non-zero C-ABI status into the RuntimeError the reference's callers expect."""
    if rc != 0:
        msg = lib.adc_last_error()
        raise RuntimeError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))
