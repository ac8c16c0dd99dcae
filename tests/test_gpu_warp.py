"""-m gpu parity of apply_disparity: bit-exact forward vs the reference fixtures and the C oracle, gradients
within 1e-5, the reference's quirks (last column 0, output strides, closed clamp interval)."""
import os

import numpy as np
import pytest
import torch

import oracle
from tests.util import FP32_TOL, npy, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pmt():
    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as m
    m.load_library()
    return m


@pytest.mark.parametrize("name", ["warp_small.npz", "warp_probe.npz"])
def test_warp_golden(pmt, golden_dir, name):
    d = np.load(os.path.join(golden_dir, name))
    img = torch.from_numpy(d["img"]).to(DEV).requires_grad_(True)
    off = torch.from_numpy(d["off"]).to(DEV).requires_grad_(True)
    out = pmt.apply_disparity(img, off)
    assert np.array_equal(npy(out), d["out"])          # bit-exact vs the reference's own output
    if "out_stride" in d:
        assert tuple(out.stride()) == tuple(int(s) for s in d["out_stride"])  # permuted [C,N,H,W] view
    gout = torch.from_numpy(d["gout"]).to(DEV) if "gout" in d else torch.ones_like(out)
    out.backward(gout)
    assert rel_err(npy(img.grad), d["gimg"]) <= FP32_TOL
    assert rel_err(npy(off.grad), d["goff"]) <= FP32_TOL


@pytest.mark.parametrize("N,C,H,W", [(2, 3, 540, 960), (4, 2, 256, 512), (1, 128, 20, 64), (1, 1, 1, 2)])
def test_warp_vs_oracle(pmt, N, C, H, W):
    rng = np.random.default_rng(C + W)
    img = rng.standard_normal((N, C, H, W), dtype=np.float32)
    off = (-64.0 * rng.random((N, 1, H, W))).astype(np.float32)   # offset = -U(0,64)
    off[..., -1] = float(W)                                      # a column saturating on the right
    off[:, :, 0, : W // 2] = np.round(off[:, :, 0, : W // 2])    # integer offsets
    gout = rng.standard_normal((N, C, H, W), dtype=np.float32)
    i = torch.from_numpy(img).to(DEV).requires_grad_(True)
    o = torch.from_numpy(off).to(DEV).requires_grad_(True)
    out = pmt.apply_disparity(i, -(-o))       # callers pass -disp (dsnet_t2_warp.py:697)
    out.backward(torch.from_numpy(gout).to(DEV))
    assert np.array_equal(npy(out), oracle.warp_fwd(img, off))
    gi, go = oracle.warp_bwd(img, off, gout)
    assert rel_err(npy(i.grad), gi) <= FP32_TOL and rel_err(npy(o.grad), go) <= FP32_TOL


def test_warp_quirks_and_modes(pmt):
    N, C, H, W = 2, 3, 4, 16
    img = torch.randn(N, C, H, W, device=DEV)
    zero = torch.zeros(N, 1, H, W, device=DEV)
    out = pmt.apply_disparity(img, zero)
    assert torch.equal(out[..., :-1], img[..., :-1]) and torch.count_nonzero(out[..., -1]) == 0
    assert pmt.apply_disparity(img, zero, wrap_mode="nonsense") is None
    left = pmt.apply_disparity(img, zero - 100.0)
    assert torch.equal(left, img[..., :1].expand_as(img))
    # only the offset needs a gradient (minidsnetDivideDisp2: image is an input batch, dsnet_t2_warp.py:946)
    off = (torch.rand(N, 1, H, W, device=DEV) * 6 - 3).requires_grad_(True)
    pmt.apply_disparity(img, off).sum().backward()
    assert off.grad is not None and off.grad.shape == off.shape
    # border mode == zero-pad by one pixel, sample in padded coordinates, crop
    o_b = pmt.apply_disparity(img, off.detach(), wrap_mode="border")
    imgp = torch.nn.functional.pad(img, (1, 1, 1, 1))
    offp = torch.nn.functional.pad(off.detach(), (1, 1, 1, 1))
    ref = oracle.warp_fwd(npy(imgp), npy(offp))[:, :, 1:-1, 1:-1]
    assert np.array_equal(npy(o_b), ref)
    with pytest.raises(RuntimeError):
        pmt.apply_disparity(img.cpu(), zero.cpu())
