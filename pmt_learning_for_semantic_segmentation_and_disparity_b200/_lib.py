"""ctypes binding of libpmt_ops.so (the C ABI in include/pmt_ops.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the op raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libpmt_ops.so")

_lib = None
_lock = threading.Lock()

_I = ctypes.c_int
_F = ctypes.c_float
_P = ctypes.c_void_p
_L = ctypes.c_int64

# name -> argtypes (every function returns int unless listed in _SPECIAL)
_SIGNATURES = {
    "pmt_corr1d_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_bwd_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_fwd_simt_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_bwd_simt_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_fwd_tc_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_bwd_tc_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr_bwd_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_uses_fast_path": [_P, _P, _P, _I, _I, _I, _I, _I],
    "pmt_concat_volume_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_concat_volume_bwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_dispreg_fwd_f32": [_P, _P, _I, _I, _I, _I, _P],
    "pmt_dispreg_bwd_f32": [_P, _P, _I, _I, _I, _I, _P],
    "pmt_softargmin_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_softargmin_bwd_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_upsample_softargmin_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_upsample_softargmin_bwd_supported": [_I, _I, _I, _I, _I, _I, _I],
    "pmt_upsample_softargmin_bwd_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "pmt_bn_pair_stats_f32": [_P, _P, _I, _I, _I, _P],
    "pmt_bn_pair_apply_f32": [_P, _P, _I, _P, _P, _P, _P, _F, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_bn_pair_bwd_reduce_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _I, _P],
    "pmt_bn_pair_bwd_apply_f32": [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _I, _P],
    "pmt_bn_pair_stats_peer_f32": [_P, _P, _P, _I, _I, _L, _L, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_bn_pair_apply_peer_f32": [_P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _F, _F, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_bn_pair_bwd_reduce_peer_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _L, _L, _P, _P, _P, _I, _P, _P, _I, _I, _I, _P, _P, _I, _P],
    "pmt_bn_pair_bwd_apply_peer_f32": [_P, _P, _P, _P, _P, _P, _I, _L, _L, _P, _P, _P, _I, _I, _I, _P, _I, _P],
    "pmt_warp1d_fwd_f32": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pmt_warp1d_bwd_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_conv_relu_supported": [_I, _I, _I, _I, _I],
    "pmt_corr1d_conv_relu_fwd_f32": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_corr1d_conv_relu_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P],
    "pmt_warp1d_rows_supported": [_I, _I, _I],
    "pmt_warp1d_blend_fwd_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_warp1d_blend_bwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_warp1d_mse_workspace": [],
    "pmt_warp1d_mse_fwd_f32": [_P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P],
    "pmt_warp1d_mse_bwd_f32": [_P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P],
    "pmt_corr1d_fwd_bwd_host_f32": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I],
    "pmt_probe_fp32_fma": [_I, ctypes.POINTER(ctypes.c_double), _P],
    "pmt_probe_copy": [_P, _P, ctypes.c_int64, ctypes.POINTER(ctypes.c_double), _P],
    "pmt_device_supported": [_I],
    "pmt_version": [],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["pmt_last_error"])


class PmtOpsError(RuntimeError):
    """A libpmt_ops call returned a non-zero status (message from pmt_last_error())."""


def load() -> ctypes.CDLL:
    """Load libpmt_ops.so from the package directory; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). This package has no CPU or PyTorch fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, argtypes in _SIGNATURES.items():
                fn = getattr(lib, name)
                fn.argtypes = argtypes
                fn.restype = _I
            lib.pmt_last_error.argtypes = []
            lib.pmt_last_error.restype = ctypes.c_char_p
            _lib = lib
    return _lib


def last_error() -> str:
    return (load().pmt_last_error() or b"").decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != 0:
        raise PmtOpsError(f"{what} failed (status {status}): {last_error()}")
