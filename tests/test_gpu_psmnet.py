"""-m gpu parity of the PSMNet ops: concat volume (bit-exact), matchshifted, disparityregression, soft-argmin."""
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


def test_concat_golden_bit_exact(pmt, golden_dir):
    d = np.load(os.path.join(golden_dir, "concat_small.npz"))
    ref = torch.from_numpy(d["ref"]).to(DEV).requires_grad_(True)
    tgt = torch.from_numpy(d["tgt"]).to(DEV).requires_grad_(True)
    cost = pmt.build_concat_volume(ref, tgt, int(d["ndisp"]))
    assert cost.is_contiguous() and np.array_equal(npy(cost), d["cost"])
    cost.backward(torch.from_numpy(d["gcost"]).to(DEV))
    assert rel_err(npy(ref.grad), d["gref"]) <= 1e-6 and rel_err(npy(tgt.grad), d["gtgt"]) <= 1e-6


def test_matchshifted_golden(pmt, golden_dir):
    d = np.load(os.path.join(golden_dir, "matchshifted_small.npz"))
    out = pmt.matchshifted()(torch.from_numpy(d["left"]).to(DEV), torch.from_numpy(d["right"]).to(DEV), int(d["shift"]))
    assert out.shape == d["out"].shape and np.array_equal(npy(out), d["out"])


@pytest.mark.parametrize("B,C,D,H,W", [(2, 32, 48, 64, 128),   # BASELINE config 3 geometry (2 of the 4 pairs)
                                       (1, 3, 5, 7, 30),       # W%4 != 0 -> scalar path
                                       (1, 2, 70, 3, 64),      # more planes than columns: fully-zero planes
                                       (1, 1, 1, 1, 4)])
def test_concat_vs_oracle_bit_exact(pmt, B, C, D, H, W):
    rng = np.random.default_rng(D + W)
    ref = rng.standard_normal((B, C, H, W), dtype=np.float32)
    tgt = rng.standard_normal((B, C, H, W), dtype=np.float32)
    r = torch.from_numpy(ref).to(DEV).requires_grad_(True)
    t = torch.from_numpy(tgt).to(DEV).requires_grad_(True)
    cost = pmt.build_concat_volume(r, t, D)
    assert np.array_equal(npy(cost), oracle.concat_fwd(ref, tgt, D))
    g = rng.standard_normal(cost.shape, dtype=np.float32)
    cost.backward(torch.from_numpy(g).to(DEV))
    gr, gt = oracle.concat_bwd(g)
    assert rel_err(npy(r.grad), gr) <= 1e-6 and rel_err(npy(t.grad), gt) <= 1e-6
    for s in {0, min(3, W), min(D - 1, W)}:
        ms = pmt.matchshifted()(r.detach(), t.detach(), s)
        assert np.array_equal(npy(ms)[:, :, 0], oracle.concat_fwd(ref, tgt, s + 1)[:, :, s])


def test_concat_full_size_properties(pmt):
    """config 3: B=4, (32,64,128) -> (64,48,64,128): checksum-of-planes property on the full volume."""
    B, C, D, H, W = 4, 32, 48, 64, 128
    g = torch.Generator(device=DEV).manual_seed(1)
    ref = torch.randn(B, C, H, W, device=DEV, generator=g)
    tgt = torch.randn(B, C, H, W, device=DEV, generator=g)
    cost = pmt.build_concat_volume(ref, tgt, D)
    assert cost.shape == (B, 2 * C, D, H, W)
    for i in (0, 1, 17, 47):
        assert torch.equal(cost[:, :C, i, :, i:], ref[:, :, :, i:])
        assert torch.equal(cost[:, C:, i, :, i:], tgt[:, :, :, :W - i])
        assert torch.count_nonzero(cost[:, :, i, :, :i]) == 0


def test_softargmin_and_dispreg_golden(pmt, golden_dir):
    d = np.load(os.path.join(golden_dir, "softargmin_small.npz"))
    cost = torch.from_numpy(d["cost"]).to(DEV).requires_grad_(True)
    out = pmt.softargmin(cost)
    assert rel_err(npy(out), d["out"]) <= FP32_TOL
    assert abs(float(out[0, 0, 0]) - 7.0) < 1e-5
    out.backward(torch.from_numpy(d["gout"]).to(DEV))
    assert rel_err(npy(cost.grad), d["gcost"]) <= FP32_TOL
    x = torch.from_numpy(d["x"]).to(DEV).requires_grad_(True)
    o = pmt.disparityregression(x.shape[1])(x)
    assert rel_err(npy(o), d["dispreg_out"]) <= FP32_TOL
    o.backward(torch.from_numpy(d["gout"]).to(DEV))
    assert np.array_equal(npy(x.grad), d["gx"])
    # composition the reference uses: softmax -> disparityregression
    o2 = pmt.disparityregression(cost.shape[1])(torch.softmax(cost.detach(), dim=1))
    assert rel_err(npy(o2), d["out"]) <= FP32_TOL
    with pytest.raises(RuntimeError):
        pmt.disparityregression(5)(x)


@pytest.mark.parametrize("B,D,H,W", [(1, 192, 256, 512),   # one full config-3 pair
                                     (2, 192, 16, 33),     # H*W % 4 != 0 -> scalar path
                                     (1, 7, 4, 8),         # D < unroll batch
                                     (1, 1, 2, 4)])
def test_softargmin_vs_oracle(pmt, B, D, H, W):
    rng = np.random.default_rng(D)
    cost = (4.0 * rng.standard_normal((B, D, H, W))).astype(np.float32)
    gout = rng.standard_normal((B, H, W), dtype=np.float32)
    c = torch.from_numpy(cost).to(DEV).requires_grad_(True)
    out = pmt.softargmin(c)
    out.backward(torch.from_numpy(gout).to(DEV))
    assert rel_err(npy(out), oracle.softargmin_fwd(cost)) <= FP32_TOL
    assert rel_err(npy(c.grad), oracle.softargmin_bwd(cost, gout)) <= FP32_TOL
    x = np.abs(cost) / np.abs(cost).sum(1, keepdims=True)
    o = pmt.disparityregression(D)(torch.from_numpy(x).to(DEV))
    assert rel_err(npy(o), oracle.dispreg_fwd(x)) <= FP32_TOL


def test_softargmin_properties_full_batch(pmt):
    """config 3 batch (4,192,256,512): one-hot -> d; shift invariance; bounds 0 <= out <= D-1."""
    B, D, H, W = 4, 192, 256, 512
    g = torch.Generator(device=DEV).manual_seed(2)
    cost = 4.0 * torch.randn(B, D, H, W, device=DEV, generator=g)
    out = pmt.softargmin(cost)
    assert float(out.min()) >= 0.0 and float(out.max()) <= D - 1
    out_shift = pmt.softargmin(cost + 3.25)
    assert float((out - out_shift).abs().max()) / float(out.abs().max()) <= FP32_TOL
    ref = (torch.softmax(cost[:1], dim=1) * torch.arange(D, device=DEV).view(1, D, 1, 1)).sum(1)
    assert float((out[:1] - ref).abs().max()) / float(ref.abs().max()) <= FP32_TOL
    onehot = torch.full((1, D, 4, 8), -40.0, device=DEV)
    idx = torch.randint(0, D, (1, 1, 4, 8), device=DEV)
    onehot.scatter_(1, idx, 40.0)
    assert torch.allclose(pmt.softargmin(onehot), idx[:, 0].float(), atol=1e-4)


# ---- f1: F.upsample(trilinear) + softmax + disparityregression in one kernel ---------------------------------
def test_upsample_softargmin_golden(pmt, golden_dir):
    d = np.load(os.path.join(golden_dir, "upsoftargmin_small.npz"))
    c = torch.from_numpy(d["cost3"]).to(DEV).requires_grad_(True)
    pred = pmt.upsample_softargmin(c, int(d["maxdisp"]), tuple(int(v) for v in d["size"]))
    assert rel_err(npy(pred), d["pred"]) <= FP32_TOL
    pred.backward(torch.from_numpy(d["gpred"]).to(DEV))
    assert rel_err(npy(c.grad), d["gcost3"]) <= FP32_TOL


@pytest.mark.parametrize("B,Dq,Hq,Wq,scale", [(2, 48, 64, 128, 4),      # config 3: (48,64,128) -> (192,256,512)
                                              (1, 5, 7, 9, 3), (1, 1, 2, 3, 4)])
def test_upsample_softargmin_vs_oracle_and_unfused(pmt, B, Dq, Hq, Wq, scale):
    rng = np.random.default_rng(Dq)
    low = (3.0 * rng.standard_normal((B, 1, Dq, Hq, Wq))).astype(np.float32)
    D, H, W = Dq * scale, Hq * scale, Wq * scale
    c = torch.from_numpy(low).to(DEV)
    pred = pmt.upsample_softargmin(c, D, (H, W))
    assert pred.shape == (B, H, W)
    if B * H * W <= 300000:
        assert rel_err(npy(pred), oracle.upsample_softargmin_fwd(low, D, (H, W))) <= FP32_TOL
    # the unfused sequence the reference runs (ATen upsample -> softmax -> regression), on the GPU
    cr = c.clone().requires_grad_(True)
    up = torch.nn.functional.interpolate(cr, size=[D, H, W], mode="trilinear", align_corners=False)[:, 0]
    ref = (torch.softmax(up, 1) * torch.arange(D, device=DEV).view(1, D, 1, 1)).sum(1)
    assert float((pred - ref).abs().max()) / float(ref.abs().max()) <= FP32_TOL
    # backward: fused kernels (per-pixel plane gradients + adjoint of the spatial interpolation) vs autograd through
    # the unfused sequence
    g = torch.from_numpy(rng.standard_normal((B, H, W)).astype(np.float32)).to(DEV)
    cf = c.clone().requires_grad_(True)
    pmt.upsample_softargmin(cf, D, (H, W)).backward(g)
    ref.backward(g)
    assert cf.grad.shape == cr.grad.shape
    # (absolute floor: with a single source plane the gradient is exactly 0 and autograd returns the round-off of a
    # few hundred atomically added O(1) terms, ~1e-6 and different from run to run)
    err = float((cf.grad - cr.grad).abs().max())
    assert err <= 1e-4 * float(cr.grad.abs().max()) + 5e-5, (err, float(cr.grad.abs().max()))
