"""ctypes binding of ``libtcvom_b200.so`` (the C ABI declared in ``include/tcvom_b200.h``).

The library is built in-tree by ``tcvom_b200/csrc/build.sh`` (or ``__graft_entry__.build()``).
Loading is lazy and does NOT touch CUDA, so that ``pred_test.py``-style ``fork`` after import
stays legal.  There is no fallback: if the library is missing or a call fails, a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

# TCV_LIB: kernel A/B experiments only (a second build of the same sources with other tile constants)
LIB_PATH = os.environ.get("TCV_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtcvom_b200.so")

ACT_NONE, ACT_RELU, ACT_LEAKY02, ACT_TANH01, ACT_LEAKY001, ACT_CLAMP01, ACT_RELU6 = 0, 1, 2, 3, 4, 5, 6
PAD_ZERO, PAD_REFLECT = 0, 1
MAX_TAPS = 16

c_void_p, c_int, c_ll, c_float = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class ConvDesc(C.Structure):
    """Mirror of ``tcv_conv_desc``."""
    _fields_ = [
        ("x", c_void_p), ("x_plane", c_ll), ("x_img_stride", c_ll),
        ("n", c_int), ("ih", c_int), ("iw", c_int), ("cin", c_int),
        ("w", c_void_p), ("w_tc", c_void_p), ("w_tc_taps", c_int), ("w_tc_fold", c_void_p),
        ("ntaps", c_int), ("dy", c_int * MAX_TAPS), ("dx", c_int * MAX_TAPS),
        ("wtap", c_int * MAX_TAPS),
        ("stride", c_int), ("pad_mode", c_int),
        ("y", c_void_p), ("y_f32", c_void_p),
        ("oh", c_int), ("ow", c_int), ("cout", c_int), ("gh", c_int), ("gw", c_int),
        ("oy_mul", c_int), ("oy_off", c_int), ("ox_mul", c_int), ("ox_off", c_int),
        ("s1", c_void_p), ("b1", c_void_p), ("res1", c_void_p), ("res1_plane", c_ll),
        ("res1_shift", c_int), ("act", c_int),
        ("s2", c_void_p), ("b2", c_void_p), ("res2", c_void_p), ("res2_plane", c_ll),
        ("stats", c_void_p), ("stats_groups", c_int), ("stats_copies", c_int),
    ]


class BnDesc(C.Structure):
    """Mirror of ``tcv_bn_desc`` (train-mode BatchNorm fused with 1/sigma, activation and residuals)."""
    _fields_ = [
        ("z", c_void_p), ("z_plane", c_ll),
        ("n", c_int), ("h", c_int), ("w", c_int), ("c", c_int),
        ("groups", c_int), ("inv_sigma", c_void_p),
        ("mode", c_int), ("act", c_int),
        ("gamma", c_void_p), ("beta", c_void_p), ("mean", c_void_p), ("invstd", c_void_p),
        ("res1", c_void_p), ("res1_plane", c_ll), ("res1_shift", c_int),
        ("res2", c_void_p), ("res2_plane", c_ll),
        ("y", c_void_p), ("y_plane", c_ll),
    ]


class SnDesc(C.Structure):
    """Mirror of ``tcv_sn_desc`` (one spectral-norm layer of the batched power iteration)."""
    _fields_ = [
        ("w_bar", c_void_p), ("rows", c_int), ("cols", c_int), ("u", c_void_p), ("v", c_void_p),
        ("calls", c_int), ("u_hist", c_void_p), ("v_hist", c_void_p), ("sigma", c_void_p),
        ("inv_sigma", c_void_p),
    ]


c_double = C.c_double

# name -> (restype, argtypes); every symbol include/tcvom_b200.h declares
SIGNATURES = {
    "tcv_version": (c_int, []),
    "tcv_last_error": (C.c_char_p, []),
    "tcv_launch_count": (c_ll, []),
    "tcv_conv2d": (c_int, [C.POINTER(ConvDesc), c_void_p]),
    "tcv_conv2d_path": (c_int, [C.POINTER(ConvDesc)]),
    "tcv_set_conv_tc_version": (c_int, [c_int]),
    "tcv_set_debug_flags": (c_int, [c_int]),
    "tcv_sn_fold_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p]),
    "tcv_bn_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "tcv_preprocess_eval": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "tcv_postprocess_eval": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "tcv_preprocess_eval_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "tcv_postprocess_eval_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                        c_void_p]),
    "tcv_preprocess_train": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p]),
    "tcv_losses_vmd": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float] * 3 + [c_void_p] * 6),
    "tcv_avgpool2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_pad_reflect1": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_unknown_os8": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_prep": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                             c_int, c_void_p]),
    "tcv_gca_values": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tcv_gca_softmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tcv_pack_weight_tc": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_pack_weight_fold": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "tcv_gemm_tn_tc": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_ll, c_ll, c_int,
                               c_int, c_int, c_int, c_void_p]),
    "tcv_gca_fold": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_prep_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "tcv_gca_values_parity": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_rowstats": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "tcv_gca_shift_add": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_softmax_shift": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_unfold_parity": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gemm_tn_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_ll, c_ll, c_ll, c_int, c_void_p]),
    "tcv_tam_attend": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcv_nchw_to_split": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "tcv_split_to_nchw": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_ll, c_void_p, c_void_p]),
    # ---- training path
    "tcv_copy_images": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p]),
    "tcv_add_split": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_ll, c_void_p]),
    "tcv_channel_sum": (c_int, [c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_pool2_scaled": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "tcv_upsample2_scaled": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "tcv_pad_reflect1_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_tanh01_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_void_p]),
    "tcv_head_tanh01": (c_int, [c_void_p, c_ll, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_gn_finalize_acc": (c_int, [c_void_p, c_int, c_int, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p,
                                    c_void_p, c_void_p]),
    "tcv_gemm_tc_ex": (c_int, [c_void_p, c_ll, c_ll, c_ll, c_int, c_void_p, c_ll, c_ll, c_ll, c_int, c_void_p, c_int, c_int,
                               c_int, c_ll, c_ll, c_int, c_void_p]),
    "tcv_zero_bytes": (c_int, [c_void_p, c_ll, c_void_p]),
    "tcv_frame_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                  c_void_p]),
    "tcv_dwconv3x3": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                              c_void_p, c_void_p]),
    "tcv_index_finish": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    "tcv_index_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tcv_index_upcat": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_ll, c_void_p, c_int, c_ll, c_int, c_int,
                                c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_maxpool2_idx": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tcv_maxunpool2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_head_conv5_clamp01": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcv_dim_fix_inputs": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_unfold_parity_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_values_parity_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_softmax_bwd_grid": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_prep_bwd_grid": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                      c_void_p]),
    "tcv_peer_allreduce_f64": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, C.c_ulonglong, c_ll, c_void_p]),
    "tcv_peer_buffer_bytes": (c_int, [c_ll, C.POINTER(c_ll)]),
    "tcv_head_conv_tanh01": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcv_head_tanh01_bwd": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_f32_to_split": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p]),
    "tcv_split_to_f32": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_void_p]),
    "tcv_transpose_packed": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_sn_power_iter": (c_int, [c_void_p, c_int, c_void_p]),
    "tcv_bn_stats": (c_int, [C.POINTER(BnDesc), c_void_p, c_void_p]),
    "tcv_bn_finalize": (c_int, [c_void_p, c_double, c_double, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "tcv_bn_apply": (c_int, [C.POINTER(BnDesc), c_void_p]),
    "tcv_bn_bwd_reduce": (c_int, [C.POINTER(BnDesc), c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p]),
    "tcv_bn_param_grads": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tcv_bn_bwd_apply": (c_int, [C.POINTER(BnDesc), c_void_p, c_ll, c_void_p, c_double, c_void_p, c_ll, c_void_p,
                                 c_int, c_void_p]),
    "tcv_group_dot": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_conv2d_wgrad": (c_int, [C.POINTER(ConvDesc), c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_transpose_pad": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_ll, c_ll, c_void_p]),
    "tcv_wgrad_tc": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_ll, c_int, c_int, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "tcv_conv2d_wgrad_nhwc_tc": (c_int, [C.POINTER(ConvDesc), c_void_p, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_weight_grad_unpack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "tcv_gemm_f32_strided": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_int,
                                     c_ll, c_ll, c_ll, c_int, c_int, c_void_p]),
    "tcv_gca_fold_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcv_gca_softmax_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_transpose_planes": (c_int, [c_void_p, c_ll, c_int, c_int, c_ll, c_ll, c_void_p, c_ll, c_ll, c_ll, c_int,
                                     c_void_p]),
    "tcv_gca_values_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gca_prep_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                 c_void_p]),
    "tcv_tam_attend_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tcv_losses_vmd_bwd": (c_int, [c_void_p] * 9 + [c_int] * 5 + [c_float] * 3 + [c_void_p] * 4),
    # ---- FBA base network
    "tcv_ws_pack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_gn_stats": (c_int, [c_void_p, c_ll, c_int, c_ll, c_int, c_void_p, c_void_p]),
    "tcv_gn_finalize": (c_int, [c_void_p, c_int, c_ll, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                c_void_p]),
    "tcv_gn_apply": (c_int, [c_void_p, c_ll, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p,
                             c_ll, c_int, c_int, c_void_p]),
    "tcv_maxpool3s2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_adaptive_avgpool": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p]),
    "tcv_bilinear": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int,
                             c_void_p]),
    "tcv_copy_channels": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_ll, c_void_p]),
    "tcv_space_to_depth2": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_s2d_pack_stem": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "tcv_s2d_pack": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_fba_encode_inputs": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_fba_edt_cols": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_fba_edt_rows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_fba_cat_inputs": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "tcv_fba_fusion": (c_int, [c_void_p, c_void_p, c_ll, c_ll, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tcv_postprocess_eval_fba": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    """Loads the shared library once (thread-safe: nn.DataParallel calls from worker threads)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"tcvom_b200: native library not built ({LIB_PATH}); run "
                        "`python -c 'import __graft_entry__ as g; g.build()'` or tcvom_b200/csrc/build.sh. "
                        "There is no CPU/PyTorch fallback.")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().tcv_last_error().decode(errors="replace")
        raise RuntimeError(f"tcvom_b200.{what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().tcv_launch_count())
