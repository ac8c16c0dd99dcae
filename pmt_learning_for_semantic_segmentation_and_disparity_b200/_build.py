"""In-tree build of libpmt_ops.so (nvcc, sm_100a only).  Called by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), "include")
LIB_PATH = os.path.join(PKG_DIR, "libpmt_ops.so")
SOURCES = ["abi.cu", "corr_generic.cu", "corr1d_fwd.cu", "corr1d_bwd.cu", "corr1d_fwd_tc.cu", "corr1d_bwd_tc.cu", "corr1d_bwd_tca.cu", "corr_conv_fused.cu", "corr2d_rows.cu", "psmnet_ops.cu", "warp1d.cu", "bn_pair.cu"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpmt_ops.so cannot be built")


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "pmt_ops.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every kernel for sm_100a into pmt_.../libpmt_ops.so (kept in-tree, git-ignored)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-I", INCLUDE, "-I", CSRC, "-Xptxas", "-v", *(["-DPMT_BWD_PROFILE", "-DPMT_DEV_KNOBS"] if os.environ.get("PMT_BWD_PROFILE") else []),
           *(["-DPMT_DEV_KNOBS"] if os.environ.get("PMT_DEV_KNOBS") and not os.environ.get("PMT_BWD_PROFILE") else []),
           "-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG_DIR, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH
