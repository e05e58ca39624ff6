"""ctypes binding of libt2i_b200.so (include/t2i_b200.h) -- the stub a maintainer of the reference
would add to reach the CUDA path.  There is no fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("T2I_B200_LIB") or os.path.join(_HERE, "libt2i_b200.so")     # the override is a development aid

CONV_S1, CONV_K4S2, DECONV_K4S2 = 0, 1, 2
ACT_NONE, ACT_LRELU, ACT_RELU = 0, 1, 2
MASK_NONE, MASK_LRELU, MASK_RELU = 0, 1, 2
W_NK, W_KN = 0, 1
SCALARS = ["D_loss", "D_loss_real", "D_loss_fake", "D_loss_mismatch", "wdist", "wdist2", "reg_loss",
           "balance_loss", "real_gp", "real_gp2", "kt", "kt_grad", "G_loss", "G_kl_loss"]
S_COUNT = 16


class T2IError(RuntimeError):
    pass


class Act(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("plane_stride", C.c_longlong), ("n", C.c_int), ("h", C.c_int),
                ("w", C.c_int), ("c", C.c_int), ("pitch", C.c_int), ("coff", C.c_int)]


class ConvGemmDesc(C.Structure):
    _fields_ = [("mode", C.c_int), ("k", C.c_int), ("flip", C.c_int), ("np", C.c_int), ("x", Act),
                ("w", C.c_void_p), ("w_plane_stride", C.c_longlong), ("w_rows", C.c_int), ("w_cols", C.c_int),
                ("w_layout", C.c_int),
                ("y", Act), ("bias", C.c_void_p), ("add", Act), ("mask", Act), ("act", C.c_int),
                ("mask_kind", C.c_int), ("stat_sum", C.c_void_p), ("stat_sq", C.c_void_p), ("stat_dot", C.c_void_p),
                ("stat_x", Act), ("stat_n", C.c_int), ("stat_c", C.c_int), ("w_n0", C.c_int),
                ("x_img", C.c_void_p)]


class WgradDesc(C.Structure):
    _fields_ = [("mode", C.c_int), ("k", C.c_int), ("np", C.c_int), ("x", Act), ("dy", Act),
                ("dw", C.c_void_p), ("cout", C.c_int), ("cin", C.c_int), ("split_k", C.c_int)]


_P, _LL, _I, _F = C.c_void_p, C.c_longlong, C.c_int, C.c_float
# name -> argtypes (all return int status); mirrors include/t2i_b200.h one to one
SIGNATURES = {
    "t2i_conv_gemm": [C.POINTER(ConvGemmDesc), _P],
    "t2i_wgrad_gemm": [C.POINTER(WgradDesc), _P],
    "t2i_debug_timeline": [_P, _I],
    "t2i_to_planes": [_P, _P, _LL, _I, _LL, _I, _P, _P],
    "t2i_from_planes": [_P, _LL, _I, _P, _LL, _P],
    "t2i_im2col_k4s2_c3": [_P, _I, _I, _I, _P, _P, _LL, _I, _P],
    "t2i_col2im_k4s2_c3": [_P, _LL, _I, _I, _I, _I, _P, _P, _P],
    "t2i_im2col_k3s1_c3": [_P, _I, _I, _I, _P, _LL, _I, _P],
    "t2i_tanh_c3_fwd": [_P, _LL, _I, _P, _LL, _P],
    "t2i_tanh_c3_bwd": [_P, _P, _P, _LL, _I, _LL, _P],
    "t2i_conv3x3_c3_tanh_fwd": [_P, _P, _P, _P, _I, _I, _I, _P],
    "t2i_conv3x3_c3_tanh_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    "t2i_colsum": [_P, _LL, _I, _LL, _I, _I, _I, _P, _P],
    "t2i_bn_stats": [_P, _LL, _I, _LL, _I, _P, _P, _P, _P, _F, _P],
    "t2i_bn_apply": [_P, _LL, _P, _P, _P, _P, _P, _LL, _P, _LL, _I, _LL, _I, _I, _P],
    "t2i_bn_bwd_reduce": [_P, _LL, _P, _LL, _P, _P, _I, _LL, _I, _P, _P, _P],
    "t2i_bn_bwd_apply": [_P, _LL, _P, _LL, _P, _P, _P, _P, _P, _P, _LL, _I, _LL, _I, _P],
    "t2i_bn_apply_train": [_P, _LL, _P, _F, _P, _P, _P, _LL, _P, _LL, _I, _LL, _I, _I, _P, _P, _P, _P, _P, _F, _LL, _I, _F, _P],
    "t2i_bn_bwd_fused": [_P, _LL, _P, _LL, _P, _P, _P, _P, _P, _P, _P, _F, _I, _P, _LL, _P, _I, _LL, _I, _LL, _I, _F, _P],
    "t2i_ce_seeds": [_P, _I, _F, _F, _F, _P, _P, _P],
    "t2i_s1_scalars": [_P, _P, _I, _I, _F, _F, _I, _P],
    "t2i_bn_update_moving": [_P, _P, _P, _P, _LL, _I, _F, _P],
    "t2i_act_bwd": [_P, _LL, _P, _LL, _P, _LL, _I, _LL, _I, _P],
    "t2i_embed_tile": [_P, _LL, _P, _LL, _I, _I, _I, _I, _I, _I, _P],
    "t2i_embed_reduce": [_P, _LL, _P, _LL, _I, _I, _I, _I, _I, _I, _P],
    "t2i_dout_fwd": [_P, _LL, _I, _P, _P, _P, _I, _I, _P],
    "t2i_dout_bwd_data": [_P, _LL, _I, _P, _P, _P, _LL, _I, _I, _P],
    "t2i_dout_bwd_weight": [_P, _LL, _I, _P, _P, _P, _I, _I, _I, _P],
    "t2i_gp_interp": [_P, _P, _P, _P, _I, _I, _P],
    "t2i_gp_penalty": [_P, _I, _I, _F, _F, _P, _P, _P, _P],
    "t2i_ca_fwd": [_P, _P, _P, _P, _LL, _I, _I, _I, _I, _P, _P],
    "t2i_ca_bwd": [_P, _P, _LL, _P, _P, _LL, _I, _I, _I, _I, _F, _P],
    "t2i_deconv_img": [C.POINTER(Act), _P, _LL, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "t2i_wgrad_img": [_P, _LL, _I, _I, _I, C.POINTER(Act), _I, _I, _P, _I, _I, _P],
    "t2i_img_to_rows": [_P, _I, _I, _I, _P, _P, _LL, _I, _P],
    "t2i_dense_f32": [_P, _I, _I, _P, _P, _I, _I, _P, _P],
    "t2i_scale_rows": [_P, _P, _P, _LL, _I, _P],
    "t2i_d_seeds": [_P, _P, _I, _F, _P],
    "t2i_d_sums": [_P, _I, _P, _P],
    "t2i_d_scalars": [_P, _P, _P, _I, _F, _F, _P],
    "t2i_g_sums": [_P, _I, _P, _P],
    "t2i_g_scalars": [_P, _P, _I, _I, _F, _P],
    "t2i_pack_weight": [_P, _I, _I, _I, _P, _LL, _P, _LL, _I, _P],
    "t2i_adam_tf": [_P, _P, _P, _P, _LL, _P, _F, _F, _F, _F, _P, _LL, _I, _P],
    "t2i_ln_stats": [_P, _LL, _I, _I, _LL, _P, _P],
    "t2i_ln_apply": [_P, _LL, _P, _F, _P, _P, _P, _LL, _I, _I, _LL, _I, _I, _P],
    "t2i_ln_bwd_reduce": [_P, _LL, _P, _LL, _P, _F, _P, _P, _P, _P, _I, _I, _LL, _I, _P],
    "t2i_ln_bwd_apply": [_P, _LL, _P, _LL, _P, _F, _P, _P, _P, _LL, _P, _I, _I, _LL, _I, _P],
    "t2i_upscale2x": [_P, _LL, _P, _LL, _I, _I, _I, _I, _I, _F, _P, _LL, _I, _P],
    "t2i_copy_window": [_P, _LL, _I, _I, _P, _LL, _I, _I, _I, _LL, _I, _P],
    "t2i_pool2x": [_P, _LL, _P, _LL, _I, _I, _I, _I, _I, _F, _P],
    "t2i_axpby": [_P, _LL, _P, _LL, _P, _LL, _I, _LL, _P, _P],
    "t2i_img_to_c8": [_P, _I, _LL, _P, _P, _LL, _I, _P],
    "t2i_c8_to_img": [_P, _LL, _I, _P, _LL, _P],
}
OTHER_SYMBOLS = ["t2i_last_error", "t2i_version", "t2i_launch_count"]

_lib = None


def load():
    """Load the shared library (built by ``make -C text-to-image_b200/csrc`` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise T2IError("libt2i_b200.so is not built (%s); run __graft_entry__.build(). "
                       "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        if os.environ.get("T2I_B200_LIB") and not hasattr(lib, name):
            continue        # development aid: an older build without the newest entry points
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.t2i_last_error.restype = C.c_char_p
    lib.t2i_version.restype = C.c_int
    lib.t2i_launch_count.restype = C.c_longlong
    _lib = lib
    return lib


TIMELINE = None     # bench.py: when a list, every C-ABI call appends [entry point, start event, end event]


def call(name, *args):
    lib = load()
    if TIMELINE is not None:
        import torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = getattr(lib, name)(*args)
        ev1.record()
        TIMELINE.append([name, ev0, ev1])
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise T2IError("%s failed (%d): %s" % (name, rc, lib.t2i_last_error().decode()))


def launch_count():
    return int(load().t2i_launch_count())
