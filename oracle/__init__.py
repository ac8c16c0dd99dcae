"""CPU oracle for the stereo cost-volume hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package never does: it fails
loudly when its CUDA library is missing instead of falling back to anything in here.

What is in here
---------------
* ``pmt_oracle.c`` -> ``_build/libpmt_oracle.so`` (``make -C oracle``): plain-C restatement of the
  four reference ops (see the header of that file for the reference file:line each follows and
  for which of them are pinned by golden fixtures).
* ``torch_ref.py``: the same ops restated in pure PyTorch (any dtype, autograd-differentiable),
  used to separate oracle rounding from kernel rounding (fp64) and to cross-check the C port.
* ``make_golden.py``: imports the reference's own Python ops from ``/root/reference`` (build
  container only) and writes ``tests/golden/*.npz``.

Parity status: correlation = **parity unpinned** (third-party ``spatial-correlation-sampler``,
unpinned, absent from /root/reference); concat volume, disparityregression/soft-argmin and
apply_disparity = pinned by fixtures generated from the reference's own code.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpmt_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    """Compile the C oracle (gcc). Returns the path of the shared object."""
    src = os.path.join(_HERE, "pmt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.pmt_oracle_num_threads.restype = ctypes.c_int
    return _lib


def num_threads() -> int:
    return int(lib().pmt_oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().pmt_oracle_set_num_threads(ctypes.c_int(int(n)))


def _c(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _p(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _pair(v):
    return (int(v), int(v)) if np.isscalar(v) else (int(v[0]), int(v[1]))


def corr_params(kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    """Pack the sampler's constructor arguments in the order of the upstream backend call."""
    kH, kW = _pair(kernel_size)
    pH, pW = _pair(patch_size)
    dH, dW = _pair(stride)
    padH, padW = _pair(padding)
    dilH, dilW = _pair(dilation)
    dpH, dpW = _pair(dilation_patch)
    return np.array([kH, kW, pH, pW, padH, padW, dilH, dilW, dpH, dpW, dH, dW], dtype=np.int32)


def corr_fwd(in1, in2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    in1, in2 = _c(in1), _c(in2)
    B, C, H, W = in1.shape
    prm = corr_params(kernel_size, patch_size, stride, padding, dilation, dilation_patch)
    oH, oW = ctypes.c_int(), ctypes.c_int()
    lib().pmt_oracle_corr_out_size(H, W, prm.ctypes.data_as(_i32p), ctypes.byref(oH), ctypes.byref(oW))
    out = np.empty((B, int(prm[2]), int(prm[3]), oH.value, oW.value), dtype=np.float32)
    lib().pmt_oracle_corr_fwd(_p(in1), _p(in2), _p(out), B, C, H, W, prm.ctypes.data_as(_i32p))
    return out


def corr_bwd(in1, in2, gout, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
    in1, in2, gout = _c(in1), _c(in2), _c(gout)
    B, C, H, W = in1.shape
    prm = corr_params(kernel_size, patch_size, stride, padding, dilation, dilation_patch)
    g1, g2 = np.empty_like(in1), np.empty_like(in2)
    lib().pmt_oracle_corr_bwd(_p(in1), _p(in2), _p(gout), _p(g1), _p(g2), B, C, H, W,
                              prm.ctypes.data_as(_i32p))
    return g1, g2


def concat_fwd(ref, tgt, ndisp):
    ref, tgt = _c(ref), _c(tgt)
    B, C, H, W = ref.shape
    cost = np.empty((B, 2 * C, int(ndisp), H, W), dtype=np.float32)
    lib().pmt_oracle_concat_fwd(_p(ref), _p(tgt), _p(cost), B, C, int(ndisp), H, W)
    return cost


def concat_bwd(gcost):
    gcost = _c(gcost)
    B, C2, D, H, W = gcost.shape
    C = C2 // 2
    gref = np.empty((B, C, H, W), dtype=np.float32)
    gtgt = np.empty((B, C, H, W), dtype=np.float32)
    lib().pmt_oracle_concat_bwd(_p(gcost), _p(gref), _p(gtgt), B, C, D, H, W)
    return gref, gtgt


def dispreg_fwd(x):
    x = _c(x)
    B, D, H, W = x.shape
    out = np.empty((B, H, W), dtype=np.float32)
    lib().pmt_oracle_dispreg_fwd(_p(x), _p(out), B, D, H, W)
    return out


def dispreg_bwd(gout, D):
    gout = _c(gout)
    B, H, W = gout.shape
    gx = np.empty((B, int(D), H, W), dtype=np.float32)
    lib().pmt_oracle_dispreg_bwd(_p(gout), _p(gx), B, int(D), H, W)
    return gx


def softargmin_fwd(cost):
    cost = _c(cost)
    B, D, H, W = cost.shape
    out = np.empty((B, H, W), dtype=np.float32)
    lib().pmt_oracle_softargmin_fwd(_p(cost), _p(out), B, D, H, W)
    return out


def softargmin_bwd(cost, gout):
    cost, gout = _c(cost), _c(gout)
    B, D, H, W = cost.shape
    gcost = np.empty_like(cost)
    lib().pmt_oracle_softargmin_bwd(_p(cost), _p(gout), _p(gcost), B, D, H, W)
    return gcost


def warp_fwd(img, off):
    img, off = _c(img), _c(off)
    N, C, H, W = img.shape
    out = np.empty_like(img)
    lib().pmt_oracle_warp_fwd(_p(img), _p(off), _p(out), N, C, H, W)
    return out


def warp_bwd(img, off, gout):
    img, off, gout = _c(img), _c(off), _c(gout)
    N, C, H, W = img.shape
    gimg = np.empty_like(img)
    goff = np.empty((N, 1, H, W), dtype=np.float32)
    lib().pmt_oracle_warp_bwd(_p(img), _p(off), _p(gout), _p(gimg), _p(goff), N, C, H, W)
    return gimg, goff


def upsample_softargmin_fwd(lowres, maxdisp, size):
    lowres = _c(lowres)
    if lowres.ndim == 5:
        lowres = np.ascontiguousarray(lowres[:, 0])
    B, Dq, Hq, Wq = lowres.shape
    H, W = int(size[0]), int(size[1])
    out = np.empty((B, H, W), dtype=np.float32)
    lib().pmt_oracle_upsample_softargmin_fwd(_p(lowres), _p(out), B, Dq, Hq, Wq, int(maxdisp), H, W)
    return out


def upsample_softargmin_bwd(lowres, gout, maxdisp, size):
    """Gradient of sum(gout * pred) w.r.t. the low-res logits (same shape as `lowres`)."""
    lowres = _c(lowres)
    shape = lowres.shape
    if lowres.ndim == 5:
        lowres = np.ascontiguousarray(lowres[:, 0])
    gout = _c(gout)
    B, Dq, Hq, Wq = lowres.shape
    H, W = int(size[0]), int(size[1])
    glow = np.empty_like(lowres)
    lib().pmt_oracle_upsample_softargmin_bwd(_p(lowres), _p(gout), _p(glow), B, Dq, Hq, Wq, int(maxdisp), H, W)
    return glow.reshape(shape)
