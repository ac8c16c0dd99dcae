"""-m gpu: deterministic warp backward and the fused consumers of the warp: blend
`(1-a)*seg + a*warp(seg_r, -d)` (models/dsnet_t2_warp.py:697-698) and the photo-consistency MSE
(torch_implementation.py:314-317).  Golden vectors come from the reference's own apply_disparity + expressions
(tests/golden/warp_fused_small.npz); larger shapes are checked against the oracle's differentiable restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import torch_ref
from tests.util import FP32_TOL, npy, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pmt():
    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as m
    m.load_library()
    return m


def _rand_case(N, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(N, C, H, W, generator=g)
    off = -(W / 6.0) * torch.rand(N, 1, H, W, generator=g) + 2.0
    off[0, 0, 0, :] = float(W)                      # saturates on the right: every tap hits the last column
    off[0, 0, H - 1, :] = -torch.arange(W).float()  # every pixel lands exactly on x = 0: one bucket holds the whole row
    return img, off, g


@pytest.mark.parametrize("N,C,H,W", [(1, 128, 20, 960), (4, 3, 64, 512), (2, 2, 7, 100), (1, 5, 3, 1024)])
def test_warp_backward_is_deterministic_and_matches_autograd(pmt, N, C, H, W):
    """gimg is a per-row gather in fixed order: two runs are torch.equal; values follow autograd through the oracle's
    restatement of apply_disparity; every element of gimg is written (the buffer starts as NaN)."""
    img, off, g = _rand_case(N, C, H, W, 7 + W)
    gout = torch.randn(N, C, H, W, generator=g)
    runs = []
    for _ in range(2):
        a = img.to(DEV).requires_grad_(True)
        o = off.to(DEV).requires_grad_(True)
        out = pmt.apply_disparity(a, o)
        out.backward(gout.to(DEV))
        runs.append((a.grad.clone(), o.grad.clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
    assert not torch.isnan(runs[0][0]).any()
    # fp32 oracle on purpose: x = w + off is rounded to fp32 BEFORE the floor in the reference, and at a pixel where that
    # rounding crosses an integer a float64 oracle picks other taps (a genuine discontinuity, not an error of either side)
    ar, orf = img.clone().requires_grad_(True), off.clone().requires_grad_(True)
    torch_ref.warp_ref(ar, orf).backward(gout)
    assert rel_err(npy(runs[0][0]), ar.grad.numpy()) <= FP32_TOL
    assert rel_err(npy(runs[0][1]), orf.grad.numpy()) <= FP32_TOL


def test_warp_backward_c_abi_writes_all_of_gimg(pmt):
    import ctypes

    from tests.util import vp

    lib = pmt.load_library()
    N, C, H, W = 2, 3, 5, 64
    img, off, g = _rand_case(N, C, H, W, 3)
    gout = torch.randn(C, N, H, W, generator=g)              # the forward's [C,N,H,W] storage order
    di, do, dg = img.to(DEV), off.to(DEV), gout.to(DEV)
    gimg = torch.full_like(di, float("nan"))
    goff = torch.full_like(do, float("nan"))
    assert lib.pmt_warp1d_rows_supported(N, H, W) == 1
    assert lib.pmt_warp1d_bwd_f32(vp(di), vp(do), vp(dg), vp(gimg), vp(goff), N, C, H, W, 1, None) == 0
    torch.cuda.synchronize()
    ar, orf = img.clone().requires_grad_(True), off.clone().requires_grad_(True)
    torch_ref.warp_ref(ar, orf).backward(gout.permute(1, 0, 2, 3))
    assert rel_err(npy(gimg), ar.grad.numpy()) <= FP32_TOL and rel_err(npy(goff), orf.grad.numpy()) <= FP32_TOL
    assert lib.pmt_warp1d_rows_supported(64, 540, 960) == 0   # N*H*W >= 2^24: atomic scatter path (indices inexact)


def test_warp_blend_golden(pmt, golden_dir):
    d = np.load(os.path.join(golden_dir, "warp_fused_small.npz"))
    t = lambda k: torch.from_numpy(d[k]).to(DEV).requires_grad_(True)
    seg, seg_r, disp, att = t("seg"), t("seg_r"), t("disp"), t("att")
    both, warped = pmt.warp_blend(seg, seg_r, -disp, att)
    assert np.array_equal(npy(both), d["both"]) and np.array_equal(npy(warped), d["warped"])   # bit-exact
    torch.autograd.backward((both, warped), (torch.from_numpy(d["gboth"]).to(DEV), torch.from_numpy(d["gwarped"]).to(DEV)))
    for got, key in ((seg.grad, "gseg"), (seg_r.grad, "gseg_r"), (disp.grad, "gdisp"), (att.grad, "gatt")):
        assert rel_err(npy(got), d[key]) <= FP32_TOL, key


@pytest.mark.parametrize("mask", [False, True])
def test_photo_consistency_golden(pmt, golden_dir, mask):
    d = np.load(os.path.join(golden_dir, "warp_fused_small.npz"))
    name = "masked" if mask else "plain"
    t = lambda k: torch.from_numpy(d[k]).to(DEV).requires_grad_(True)
    left, right, disp = t("left"), t("right"), t("disp")
    loss = pmt.photo_consistency_mse(right, -disp, left, mask_positive_disparity=mask)
    assert abs(float(loss) - float(d[f"mse_{name}_loss"])) <= 1e-6 * float(d[f"mse_{name}_loss"])
    (3.0 * loss).backward()
    for got, key in ((left.grad, "gleft"), (right.grad, "gright"), (disp.grad, "gdisp")):
        assert rel_err(npy(got), d[f"mse_{name}_{key}"]) <= FP32_TOL, (name, key)


@pytest.mark.parametrize("N,C,H,W", [(4, 2, 256, 512), (1, 3, 540, 960)])
def test_fused_ops_at_production_sizes(pmt, N, C, H, W):
    """Production shapes (seg logits C=2 at 256x512; RGB at 540x960) against the differentiable oracle in float64, plus
    run-to-run bit reproducibility of every output of the fused backward kernels."""
    img, off, g = _rand_case(N, C, H, W, 11)
    seg = torch.randn(N, C, H, W, generator=g)
    att = torch.rand(N, 1, H, W, generator=g)
    gb, gw = torch.randn(N, C, H, W, generator=g), torch.randn(N, C, H, W, generator=g)
    res = []
    for _ in range(2):
        ts = [x.to(DEV).requires_grad_(True) for x in (seg, img, off, att)]
        both, warped = pmt.warp_blend(*ts)
        torch.autograd.backward((both, warped), (gb.to(DEV), gw.to(DEV)))
        loss = pmt.photo_consistency_mse(ts[1], ts[2], ts[0].detach(), True)
        res.append([both.detach(), warped.detach(), loss.detach()] + [x.grad.clone() for x in ts])
    for a, b in zip(*res):
        assert torch.equal(a, b)
    # fp32 oracle (the reference's arithmetic; see test_warp_backward_is_deterministic_and_matches_autograd)
    tsr = [x.clone().requires_grad_(True) for x in (seg, img, off, att)]
    both_r, warped_r = torch_ref.warp_blend_ref(*tsr)
    torch.autograd.backward((both_r, warped_r), (gb, gw))
    assert np.array_equal(npy(res[0][0]), both_r.detach().numpy())          # forward: bit-exact
    assert np.array_equal(npy(res[0][1]), warped_r.detach().numpy())
    for got, want in zip(res[0][3:], tsr):
        assert rel_err(npy(got), want.grad.numpy()) <= FP32_TOL
    loss_r = torch_ref.photo_mse_ref(tsr[1].detach(), tsr[2].detach(), tsr[0].detach(), True)
    assert abs(float(res[0][2]) - float(loss_r)) <= 1e-5 * abs(float(loss_r))
