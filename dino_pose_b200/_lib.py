"""ctypes binding of ``libdinopose_sm100a.so`` (C ABI declared in ``include/dinopose.h``).

The library is the product: there is no CPU or PyTorch fallback.  ``lib()`` raises if the shared
object is missing (build it with ``python -m dino_pose_b200.build`` or ``__graft_entry__.build()``),
and every wrapper raises ``DinoPoseError`` on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdinopose_sm100a.so")

ABI_VERSION = 7

c_ll = C.c_longlong
c_vp = C.c_void_p


class DinoPoseError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    """Mirror of ``dp_gemm_args`` (include/dinopose.h)."""
    _fields_ = [
        ("A", c_vp), ("W", c_vp),
        ("lda", c_ll), ("ldw", c_ll),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("block_n", C.c_int),
        ("a_mode", C.c_int),
        ("C", C.c_int), ("IW", C.c_int), ("IH", C.c_int), ("NB", C.c_int),
        ("a_stride_w", c_ll), ("a_stride_h", c_ll), ("a_stride_b", c_ll),
        ("KH", C.c_int), ("KW", C.c_int), ("pad_y", C.c_int), ("pad_x", C.c_int), ("OH", C.c_int), ("OW", C.c_int),
        ("out", c_vp),
        ("ldo", c_ll),
        ("out_dtype", C.c_int),
        ("act", C.c_int),
        ("bias", c_vp), ("scale", c_vp), ("ls", c_vp), ("residual", c_vp),
        ("ldr", c_ll),
        ("res_is_bf16", C.c_int),
        ("aux_out", c_vp), ("aux_in", c_vp),
        ("ld_aux", c_ll),
        ("row_map", C.c_int), ("n_valid", C.c_int), ("map_a", C.c_int), ("map_b", C.c_int),
        ("stats", c_vp), ("stats_c", C.c_int), ("cta_pair", C.c_int),
        ("ln_gamma", c_vp), ("ln_beta", c_vp), ("ln_out", c_vp), ("ld_ln", c_ll), ("ln_eps", C.c_float),
        ("lora_A", c_vp), ("lora_B", c_vp), ("lora_y_out", c_vp), ("ld_lora_y", c_ll), ("lora_u_out", c_vp),
        ("lora_seed", c_vp), ("lora_scaling", C.c_float), ("lora_p_drop", C.c_float), ("lora_rank", C.c_int),
    ]


class WgradArgs(C.Structure):
    """Mirror of ``dp_wgrad_args`` (include/dinopose.h)."""
    _fields_ = [
        ("A", c_vp), ("B", c_vp),
        ("mode", C.c_int), ("P", C.c_int),
        ("lda", c_ll), ("ldb", c_ll),
        ("Mc", C.c_int), ("Nc", C.c_int),
        ("NB", C.c_int), ("OH", C.c_int), ("OW", C.c_int), ("IH", C.c_int), ("IW", C.c_int),
        ("a_sw", c_ll), ("a_sh", c_ll), ("a_sb", c_ll), ("b_sw", c_ll), ("b_sh", c_ll), ("b_sb", c_ll),
        ("KH", C.c_int), ("KW", C.c_int), ("pad_y", C.c_int), ("pad_x", C.c_int),
        ("out", c_vp),
        ("so_m", c_ll), ("so_mo", c_ll), ("so_n", c_ll), ("so_no", c_ll), ("so_t", c_ll),
        ("m_inner", C.c_int), ("n_inner", C.c_int),
        ("block_n", C.c_int), ("splits", C.c_int),
        ("workspace", c_vp), ("workspace_bytes", c_ll),
    ]


i, f, d, ull = C.c_int, C.c_float, C.c_double, C.c_ulonglong

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "dp_last_error": [],
    "dp_abi_version": [],
    "dp_sizeof_gemm_args": [],
    "dp_sizeof_wgrad_args": [],
    "dp_gemm_bf16": [C.POINTER(GemmArgs), c_vp],
    "dp_wgrad_bf16": [C.POINTER(WgradArgs), c_vp],
    "dp_debug_read_trace": [c_vp, i],
    "dp_set_reserved_sms": [i],
    "dp_layernorm_fwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, f, c_vp],
    "dp_layernorm_bwd": [c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, f, c_vp],
    "dp_patch_im2col": [c_vp, c_vp, i, i, i, i, c_vp],
    "dp_fill_cls": [c_vp, c_vp, i, i, i, c_vp],
    "dp_lora_fwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, f, f, c_vp, c_vp],
    "dp_lora_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, f, f, c_vp, c_vp],
    "dp_attention_fwd": [c_vp, c_vp, i, i, i, f, c_vp],
    "dp_attention_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, i, i, i, f, c_vp],
    "dp_layernorm_bwd_params": [c_vp, i, c_vp, c_vp, c_vp, c_ll, i, f, c_vp],
    "dp_colsum_prod": [c_vp, c_vp, c_vp, c_ll, i, c_vp],
    "dp_pred1x1_fwd": [c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, c_vp],
    "dp_pred1x1_bwd": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, c_vp],
    "dp_decode": [c_vp, i, i, i, d, d, c_vp, c_vp, c_vp, c_vp],
    "dp_im2col": [c_vp, c_vp, i, i, i, i, i, i, i, i, i, i, c_vp],
    "dp_col2im": [c_vp, c_vp, c_vp, i, i, i, i, i, i, i, i, i, i, i, c_vp],
    "dp_dwconv3x3": [c_vp, c_vp, c_vp, c_vp, c_vp, i, i, i, i, i, i, c_vp],
    "dp_dwconv3x3_wgrad": [c_vp, c_vp, c_vp, i, i, i, i, c_vp],
    "dp_bn_stats": [c_vp, i, c_vp, c_ll, i, c_vp],
    "dp_bn_finalize": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, i, d, f, f, c_vp],
    "dp_bn_fold_eval": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, i, f, c_vp],
    "dp_bn_apply": [c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, c_vp],
    "dp_preprocess_workspace_bytes": [i, i, i, i, i],
    "dp_preprocess_u8": [c_vp, i, i, i, i, i, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp],
    "dp_bn_finalize_apply": [c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i,
                             f, f, c_vp],
    "dp_bn_bwd_reduce": [c_vp, c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i, c_vp],
    "dp_bn_bwd_apply": [c_vp, c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_ll, i, i, i,
                        i, i, i, c_vp],
    "dp_avgpool2": [c_vp, c_vp, c_ll, i, i, c_vp],
    "dp_hm_grad_to_nhwc": [c_vp, c_vp, i, i, i, i, i, i, c_vp],
    "dp_mean_tokens": [c_vp, c_vp, i, i, i, c_vp],
    "dp_mean_tokens_bwd": [c_vp, c_vp, i, i, i, c_vp],
    "dp_sgemm_small": [c_vp, c_ll, c_ll, c_vp, c_ll, c_ll, c_vp, c_ll, i, i, i, c_vp, i, c_vp, c_ll, f, c_vp, i, c_vp],
    "dp_colsum": [c_vp, i, c_vp, c_ll, i, c_ll, c_vp],
    "dp_relu_mask": [c_vp, c_vp, c_vp, c_ll, f, c_vp],
    "dp_pose_loss": [c_vp, c_vp, c_vp, i, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, i, i, i, f, f, c_vp],
    "dp_adamw": [c_vp, c_vp, c_vp, c_vp, c_ll, f, f, f, f, f, f, c_vp, c_vp],
    "dp_adamw_dev": [c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, f, f, f, f, c_vp, i, c_vp],
    "dp_pack_weights_bf16": [c_vp, i, c_ll, c_vp],
    "dp_add_i64": [c_vp, i, c_ll, c_vp],
}
_RESTYPES = {"dp_last_error": C.c_char_p, "dp_preprocess_workspace_bytes": C.c_longlong}

_lib = None


def lib():
    """Load (once) and return the shared library; raise loudly if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DinoPoseError(
            f"{LIB_PATH} not found: the CUDA extension is the only execution path of dino_pose_b200 "
            "(no CPU / PyTorch fallback). Build it with `python -m dino_pose_b200.build`.")
    h = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(h, name)
        except AttributeError as e:
            raise DinoPoseError(f"{LIB_PATH} does not export {name}; rebuild the extension") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    if h.dp_abi_version() != ABI_VERSION:
        raise DinoPoseError(f"ABI version mismatch: library {h.dp_abi_version()} vs python {ABI_VERSION}")
    if h.dp_sizeof_gemm_args() != C.sizeof(GemmArgs) or h.dp_sizeof_wgrad_args() != C.sizeof(WgradArgs):
        raise DinoPoseError("argument struct layout mismatch between include/dinopose.h and _lib.py")
    _lib = h
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dp_last_error()
        raise DinoPoseError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
