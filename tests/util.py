import ctypes

import numpy as np
import torch

FP32_TOL = 1e-5  # north_star: "within 1e-5 relative (fp32)": max|a-b| / max|b| over the tensor


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)) if a.size else 0.0


def npy(t):
    return t.detach().contiguous().cpu().numpy()


def vp(t):
    return ctypes.c_void_p(t.data_ptr())
