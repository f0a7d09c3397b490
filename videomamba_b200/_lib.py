"""ctypes binding of libvmb200.so (C ABI declared in include/vmb200.h).

The library is the ONLY compute backend of this package: if it is missing, or a call fails,
the caller gets an exception -- there is no CPU or eager-PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvmb200.so")

VMB_F32, VMB_BF16 = 0, 1
ABI_VERSION = 5

c_void_p, c_int, c_int32, c_int64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float


class ScanArgs(C.Structure):
    """struct vmb_scan_args"""
    _fields_ = [
        ("u", c_void_p), ("u_bstride", c_int64), ("u_tstride", c_int64),
        ("delta", c_void_p), ("d_bstride", c_int64), ("d_tstride", c_int64),
        ("z", c_void_p), ("z_bstride", c_int64), ("z_tstride", c_int64),
        ("bc", c_void_p), ("bc_bstride", c_int64), ("bc_tstride", c_int64),
        ("b_off", c_int32), ("c_off", c_int32),
        ("A2", c_void_p), ("D", c_void_p), ("dt_bias", c_void_p),
        ("h0", c_void_p), ("h0_dtype", c_int32),
        ("y", c_void_p), ("y_bstride", c_int64), ("y_tstride", c_int64),
        ("h_last", c_void_p),
        ("B", c_int32), ("L", c_int32), ("Di", c_int32), ("N", c_int32),
        ("dtype", c_int32), ("softplus", c_int32), ("reverse", c_int32), ("frame_len", c_int32),
    ]


class FusedScanArgs(C.Structure):
    """struct vmb_fused_scan_args"""
    _fields_ = [
        ("u", c_void_p), ("u_bstride", c_int64), ("u_tstride", c_int64),
        ("z", c_void_p), ("z_bstride", c_int64), ("z_tstride", c_int64),
        ("xdbl", c_void_p), ("x_bstride", c_int64), ("x_tstride", c_int64),
        ("w_dt", c_void_p), ("A2", c_void_p), ("D", c_void_p), ("dt_bias", c_void_p),
        ("h0", c_void_p), ("h0_dtype", c_int32),
        ("y", c_void_p), ("y_bstride", c_int64), ("y_tstride", c_int64),
        ("h_last", c_void_p),
        ("B", c_int32), ("L", c_int32), ("Di", c_int32), ("N", c_int32), ("R", c_int32),
        ("Rp", c_int32), ("Xp", c_int32), ("reverse", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("a_geometric", c_int32), ("tune", c_int32), ("frame_len", c_int32),
        ("bwd_ckpt", c_void_p), ("z_gate", c_int32),
    ]


class MixerArgs(C.Structure):
    """struct vmb_mixer_args"""
    _fields_ = [
        ("hidden", c_void_p), ("h_bstride", c_int64), ("h_tstride", c_int64),
        ("out", c_void_p), ("o_bstride", c_int64), ("o_tstride", c_int64),
        ("w_in", c_void_p), ("b_in", c_void_p), ("w_conv", c_void_p), ("b_conv", c_void_p),
        ("w_x", c_void_p), ("w_dt", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p),
        ("A2", c_void_p), ("Dskip", c_void_p), ("dt_bias", c_void_p),
        ("w_x_pad", c_void_p), ("w_dt_pad", c_void_p), ("Xp", c_int32), ("Rp", c_int32),
        ("conv_state_in", c_void_p), ("cs_in_dtype", c_int32),
        ("conv_state_out", c_void_p), ("cs_out_dtype", c_int32),
        ("ssm_state_in", c_void_p), ("ss_in_dtype", c_int32),
        ("ssm_state_out", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("B", c_int32), ("L", c_int32), ("D", c_int32), ("Di", c_int32), ("N", c_int32),
        ("R", c_int32), ("W", c_int32),
        ("dtype", c_int32), ("reverse", c_int32), ("path", c_int32),
        ("a_geometric", c_int32), ("scan_tune", c_int32), ("frame_len", c_int32),
        ("fuse_conv_xproj", c_int32), ("gate_in_proj", c_int32),
    ]


class StepArgs(C.Structure):
    """struct vmb_step_args"""
    _fields_ = [
        ("xz", c_void_p), ("xz_bstride", c_int64),
        ("conv_state", c_void_p), ("cs_dtype", c_int32),
        ("ssm_state", c_void_p), ("ss_dtype", c_int32),
        ("w_conv", c_void_p), ("b_conv", c_void_p), ("w_x", c_void_p), ("w_dt", c_void_p),
        ("A2", c_void_p), ("Dskip", c_void_p), ("dt_bias", c_void_p),
        ("y", c_void_p), ("y_bstride", c_int64),
        ("B", c_int32), ("Di", c_int32), ("N", c_int32), ("R", c_int32), ("W", c_int32),
        ("dtype", c_int32),
    ]


class ScanBwdArgs(C.Structure):
    """struct vmb_scan_bwd_args"""
    _fields_ = [
        ("u", c_void_p), ("u_bstride", c_int64), ("u_tstride", c_int64),
        ("delta", c_void_p), ("d_bstride", c_int64), ("d_tstride", c_int64),
        ("z", c_void_p), ("z_bstride", c_int64), ("z_tstride", c_int64),
        ("bc", c_void_p), ("bc_bstride", c_int64), ("bc_tstride", c_int64),
        ("b_off", c_int32), ("c_off", c_int32),
        ("A2", c_void_p), ("D", c_void_p), ("dt_bias", c_void_p),
        ("h0", c_void_p), ("h0_dtype", c_int32),
        ("dout", c_void_p), ("dout_bstride", c_int64), ("dout_tstride", c_int64),
        ("dh_last", c_void_p),
        ("du", c_void_p), ("ddelta", c_void_p), ("dz", c_void_p),
        ("dz_bstride", c_int64), ("dz_tstride", c_int64),
        ("dbc", c_void_p), ("dbc_tstride", c_int64),
        ("dA", c_void_p), ("dD", c_void_p), ("ddt_bias", c_void_p), ("dh0", c_void_p),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("B", c_int32), ("L", c_int32), ("Di", c_int32), ("N", c_int32),
        ("dtype", c_int32), ("softplus", c_int32),
        ("fwd_ckpt", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol include/vmb200.h declares
SIGNATURES = {
    "vmb_abi_version": (c_int, []),
    "vmb_last_error": (C.c_char_p, []),
    "vmb_device_info": (c_int, [C.POINTER(c_int)] * 3),
    "vmb_add_norm_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                 c_int, c_void_p, c_void_p, c_int, c_int64, c_int, c_float, c_int,
                                 c_void_p]),
    "vmb_conv_xproj_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                   c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "vmb_gate_blend_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_void_p]),
    "vmb_linear_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                               c_int64, c_int, c_int, c_int, c_void_p]),
    "vmb_linear_fwd_act": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                   c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "vmb_causal_conv1d_fwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                      c_int, c_void_p, c_int64, c_int64, c_void_p, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vmb_causal_conv1d_update": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "vmb_selective_scan_fwd": (c_int, [C.POINTER(ScanArgs), c_void_p]),
    "vmb_fused_scan_workspace_bytes": (c_int64, [c_int] * 4),
    "vmb_selective_scan_fused_fwd": (c_int, [C.POINTER(FusedScanArgs), c_void_p]),
    "vmb_selective_state_update": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_void_p, c_int64,
                                           c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                           c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int64,
                                           c_int, c_int, c_int, c_int, c_void_p]),
    "vmb_mixer_step_fwd": (c_int, [C.POINTER(StepArgs), c_void_p]),
    "vmb_mixer_workspace_bytes": (c_int64, [c_int] * 7),
    "vmb_mixer_fwd": (c_int, [C.POINTER(MixerArgs), c_void_p]),
    "vmb_patchify": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_int, c_void_p]),
    "vmb_embed_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "vmb_pool_norm_workspace_bytes": (c_int64, [c_int] * 4),
    "vmb_pool_norm_fwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_float, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vmb_gather_rows": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                c_void_p]),
    "vmb_state_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "vmb_state_scatter": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int64, c_int, c_void_p]),
    "vmb_add_norm_bwd_workspace_bytes": (c_int64, [c_int64, c_int]),
    "vmb_add_norm_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_float, c_int, c_void_p, c_int64, c_void_p]),
    "vmb_causal_conv1d_bwd_workspace_bytes": (c_int64, [c_int] * 4),
    "vmb_causal_conv1d_bwd": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                      c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int64,
                                      c_void_p]),
    "vmb_selective_scan_bwd_workspace_bytes": (c_int64, [c_int] * 4),
    "vmb_selective_scan_bwd": (c_int, [C.POINTER(ScanBwdArgs), c_void_p]),
    "vmb_scan_bwd_ckpt_bytes": (c_int64, [c_int] * 3),
    "vmb_linear_wgrad_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "vmb_linear_wgrad": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_int,
                                 c_void_p, c_int64, c_void_p]),
    "vmb_transpose_2d": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    "vmb_colsum_workspace_bytes": (c_int64, [c_int64, c_int]),
    "vmb_colsum": (c_int, [c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_int64,
                           c_void_p]),
    "vmb_launch_count": (c_int64, []),
    "vmb_prof_enable": (c_int, [c_int]),
    "vmb_prof_read": (c_int, [C.POINTER(C.c_double), C.POINTER(c_int64), c_int]),
}
PROF_KINDS = ("add_norm", "in_proj", "conv", "x_proj", "dt_proj", "scan", "out_proj", "other")

_lib = None
_lock = threading.Lock()


class ExtensionMissing(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libvmb200.so (once).  Raises ExtensionMissing when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise ExtensionMissing(
                f"{LIB_PATH} not found: build the sm_100a extension first "
                "(python -m videomamba_b200.build). videomamba_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.vmb_abi_version() != ABI_VERSION:
            raise RuntimeError("libvmb200.so ABI version mismatch; rebuild the extension")
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().vmb_last_error().decode("utf-8", "replace")
        kind = {-1: "invalid argument", -2: "unsupported", -3: "CUDA error"}.get(rc, "error")
        raise RuntimeError(f"{what} failed ({kind}): {msg}")
