"""ctypes binding of the C ABI declared in ``include/bpreveal_b200.h``.

The shared library is built in-tree by ``bpreveal_b200/csrc/Makefile`` (``__graft_entry__.build``).
There is deliberately no fallback: if the library is missing, or a call fails, an exception is
raised.  PyTorch is used by the callers only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BPB_LIB_PATH: another build of the SAME library (the role-profile build of csrc/Makefile)
LIB_PATH = os.environ.get("BPB_LIB_PATH") or os.path.join(_HERE, "libbpreveal_b200.so")

c_void_p = C.c_void_p
c_int = C.c_int
c_ll = C.c_longlong
c_float = C.c_float


class Conv3Args(C.Structure):
    _fields_ = [
        ("B", c_int), ("L_in", c_int), ("L_out", c_int), ("F", c_int),
        ("tap_off", c_int * 3),
        ("mode", c_int),
        ("in_hi", c_void_p), ("in_lo", c_void_p),
        ("w_hi", c_void_p), ("w_lo", c_void_p),
        ("bias", c_void_p),
        ("resid_hi", c_void_p), ("resid_lo", c_void_p),
        ("L_res", c_int), ("resid_off", c_int),
        ("out_hi", c_void_p), ("out_lo", c_void_p),
        ("out2_hi", c_void_p), ("out2_lo", c_void_p),
        ("mask_out", c_void_p),
        ("mask_in", c_void_p),
        ("mult_hi", c_void_p), ("mult_lo", c_void_p),
        ("zx_in", c_void_p),
        ("zx_div", c_int),
        ("z_out", c_void_p),
        ("mask_a", c_void_p),
    ]


class Wgrad3Args(C.Structure):
    _fields_ = [
        ("B", c_int), ("L_in", c_int), ("L_out", c_int), ("F", c_int), ("dil", c_int),
        ("act_hi", c_void_p), ("act_lo", c_void_p),
        ("gz_hi", c_void_p), ("gz_lo", c_void_p),
        ("dw", c_void_p), ("db", c_void_p),
        ("ws", c_void_p), ("ws_bytes", c_ll),
        ("gz_mask", c_void_p),
    ]


class InputConvArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("L_in", c_int), ("L_out", c_int), ("W", c_int), ("Fw", c_int), ("F", c_int),
        ("x8", c_void_p), ("w_op", c_void_p), ("w_planes", c_int), ("bias", c_void_p),
        ("out_hi", c_void_p), ("out_lo", c_void_p),
        ("mask_out", c_void_p), ("z_out", c_void_p),
        ("zx_in", c_void_p), ("zx_div", c_int),
        ("mult_hi", c_void_p), ("mult_lo", c_void_p),
    ]


class HeadsDgradArgs(C.Structure):
    _fields_ = [
        ("B", c_int), ("L_in", c_int), ("L_out", c_int), ("W", c_int), ("F", c_int), ("Cp", c_int),
        ("d8_hi", c_void_p), ("d8_lo", c_void_p), ("w_op", c_void_p),
        ("g_hi", c_void_p), ("g_lo", c_void_p), ("gz_hi", c_void_p), ("gz_lo", c_void_p),
        ("mask_in", c_void_p), ("mult_hi", c_void_p), ("mult_lo", c_void_p),
    ]


class PrepTable(C.Structure):
    _fields_ = [
        ("off_dil_w", c_ll), ("off_conv1_w", c_ll), ("off_prof_w", c_ll), ("off_cnt_w", c_ll),
        ("n_layers", c_int), ("F", c_int), ("Fw", c_int), ("W_in", c_int), ("W_out", c_int),
        ("T", c_int), ("H", c_int), ("Cp", c_int), ("planes", c_int), ("in_planes", c_int),
        ("N_heads", c_int),
        ("w_fwd", c_void_p), ("w_bwd", c_void_p), ("w_in_op", c_void_p), ("w_hd_op", c_void_p),
        ("w_hf_op", c_void_p),
        ("hf_kcg", c_int),
    ]


class PeerArgs(C.Structure):
    _fields_ = [("world", c_int), ("rank", c_int), ("grads", c_void_p * 8), ("flags", c_void_p * 8)]


# name -> (restype, argtypes); mirrors include/bpreveal_b200.h one to one.
SIGNATURES = {
    "bpb_version": (c_int, []),
    "bpb_last_error": (C.c_char_p, []),
    "bpb_device_info": (c_int, [C.POINTER(c_int)] * 3),
    "bpb_abi_struct_size": (c_int, [c_int]),
    "bpb_debug_dump": (c_int, [C.POINTER(C.c_uint), c_int]),
    "bpb_conv3_tc": (c_int, [C.POINTER(Conv3Args), c_void_p]),
    "bpb_wgrad3_ws_bytes": (c_ll, [c_int, c_int, c_int, c_int]),
    "bpb_wgrad3_tc": (c_int, [C.POINTER(Wgrad3Args), c_void_p]),
    "bpb_wgrad3_multi_ws_bytes": (c_ll, [c_int, c_int, C.POINTER(c_int), c_int]),
    "bpb_wgrad3_multi": (c_int, [C.POINTER(Wgrad3Args), c_int, c_void_p, c_ll, c_void_p]),
    "bpb_x8_elems": (c_ll, [c_int, c_int]),
    "bpb_onehot_expand": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "bpb_input_conv_wop_elems": (c_ll, [c_int, c_int, c_int]),
    "bpb_prep_input_conv_weights": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                            c_void_p]),
    "bpb_input_conv_fwd": (c_int, [C.POINTER(InputConvArgs), c_void_p]),
    "bpb_heads_d8_elems": (c_ll, [c_int] * 5),
    "bpb_heads_dgrad_wop_elems": (c_ll, [c_int] * 4),
    "bpb_heads_pack_grad": (c_int, [c_int] * 8 + [c_void_p] * 6),
    "bpb_prep_heads_dgrad_weights": (c_int, [c_int] * 6 + [c_void_p] * 3 + [c_int, c_void_p]),
    "bpb_heads_dgrad_tc": (c_int, [C.POINTER(HeadsDgradArgs), c_void_p]),
    "bpb_wgrad_tower_input_ws_bytes": (c_ll, [c_int, c_int, c_void_p] + [c_int] * 5),
    "bpb_wgrad_tower_input": (c_int, [C.POINTER(Wgrad3Args), c_int, c_int, c_int, c_int, c_int] +
                              [c_void_p] * 5 + [c_int] * 4 + [c_void_p] * 5 + [c_ll, c_void_p]),
    "bpb_input_conv_wgrad_ws_bytes": (c_ll, [c_int, c_int, c_int, c_int]),
    "bpb_input_conv_wgrad": (c_int, [c_int] * 6 + [c_void_p] * 6 + [c_ll, c_void_p]),
    "bpb_heads_wgrad_ws_bytes": (c_ll, [c_int] * 5),
    "bpb_heads_wgrad_tc": (c_int, [c_int] * 8 + [c_void_p] * 7 + [c_ll, c_void_p]),
    "bpb_input_conv_dgrad": (c_int, [c_int] * 6 + [c_void_p] * 5),
    "bpb_heads_fwd": (c_int, [c_int] * 8 + [c_void_p] * 10),
    "bpb_input_dgrad_wop_elems": (c_ll, [c_int] * 3),
    "bpb_input_conv_dgrad_tc_supported": (c_int, [c_int] * 3),
    "bpb_prep_input_dgrad_weights": (c_int, [c_int] * 3 + [c_void_p] * 2 + [c_int, c_void_p]),
    "bpb_input_conv_dgrad_tc": (c_int, [c_int] * 5 + [c_void_p] * 5),
    "bpb_heads_fwd_wop_elems": (c_ll, [c_int] * 5),
    "bpb_prep_heads_fwd_weights": (c_int, [c_int] * 5 + [c_void_p] * 3 + [c_int, c_void_p]),
    "bpb_heads_fwd_ws_bytes": (c_ll, [c_int] * 3),
    "bpb_heads_fwd_groups": (c_int, [c_int] * 5),
    "bpb_heads_fwd_tc_supported": (c_int, [c_int] * 5),
    "bpb_heads_fwd_tc": (c_int, [c_int] * 6 + [c_void_p] * 8 + [c_ll, c_void_p]),
    "bpb_loss_fwd_bwd": (c_int, [c_int] * 4 + [c_void_p] * 11),
    "bpb_loss_pack_ws_elems": (c_ll, [c_int] * 3),
    "bpb_loss_fwd_bwd_pack": (c_int, [c_int] * 4 + [c_void_p] * 10 + [c_int] * 4 + [c_void_p] * 6 +
                              [c_int, c_void_p, c_void_p]),  # ..., hto, counts_partials | tiles, cb, stream
    "bpb_ushuffle_ohe": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, C.c_uint]),
    "bpb_adam_step": (c_int, [c_ll] + [c_void_p] * 5 + [c_float] * 4 + [c_void_p]),
    "bpb_peer_alloc": (c_int, [c_ll, C.POINTER(c_void_p), c_void_p]),
    "bpb_peer_open": (c_int, [c_void_p, C.POINTER(c_void_p)]),
    "bpb_peer_close": (c_int, [c_void_p]),
    "bpb_peer_free": (c_int, [c_void_p]),
    "bpb_peer_wait": (c_int, [C.POINTER(PeerArgs), c_void_p]),
    "bpb_adam_prep_step_peer": (c_int, [c_ll] + [c_void_p] * 4 + [c_float] * 3 +
                                [C.POINTER(PrepTable), C.POINTER(PeerArgs), c_void_p]),
    "bpb_adam_prep_step": (c_int, [c_ll] + [c_void_p] * 5 + [c_float] * 4 +
                           [C.POINTER(PrepTable), c_void_p]),
    "bpb_prep_conv3_weights": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                       c_void_p]),
    "bpb_onehot_encode": (c_int, [c_ll] + [c_void_p] * 4),
    "bpb_logits_to_profile": (c_int, [c_int] * 4 + [c_void_p] * 5),
    "bpb_slide_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                             c_void_p]),
    "bpb_slide_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p]),
    "bpb_log_row_sums": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "bpb_transform_fwd": (c_int, [c_ll, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "bpb_transform_bwd": (c_int, [c_ll, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "bpb_add_f32": (c_int, [c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bpb_mul_f32": (c_int, [c_ll, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bpb_shuffle_rows": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "bpb_flat_profile_seed": (c_int, [c_int] * 4 + [c_void_p] * 7),
    "bpb_shap_combine": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                 c_void_p]),
    "bpb_counts_lse": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p]),
    "bpb_combined_shap_seed": (c_int, [c_int] * 5 + [c_void_p] * 7 + [c_int, c_void_p, c_int] +
                               [c_void_p] * 10),
    "bpb_combine_bias": (c_int, [c_int] * 4 + [c_void_p] * 3 + [c_int, c_void_p, c_int] +
                         [c_void_p] * 11),
}

_lib = None


class BprevealB200Error(RuntimeError):
    """Raised when the native library is missing or a native call fails."""


def load():
    """Load (once) and return the native library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BprevealB200Error(
            f"{LIB_PATH} is missing: build it with `make -C bpreveal_b200/csrc` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def debug_dump():
    """Watchdog records of timed-out pipeline waits: (tag, block, thread, parity, payload)."""
    buf = (C.c_uint * 2048)()
    n = load().bpb_debug_dump(buf, 2048)
    if n == 0 or buf[0] == 0:
        return []
    return [tuple(buf[8 + 5 * i:13 + 5 * i]) for i in range(min(int(buf[0]), 400))]


def check(rc, what):
    if rc != 0:
        msg = load().bpb_last_error().decode("utf-8", "replace")
        raise BprevealB200Error(f"{what} failed (status {rc}): {msg}")
