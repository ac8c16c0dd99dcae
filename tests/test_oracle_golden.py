"""CPU oracle (oracle/pmt_oracle.c + oracle/torch_ref.py) against the fixtures generated from the
reference's own Python ops (oracle/make_golden.py) and against analytic known-answer tests."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ref

FP32_TOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---- a4 warp: pinned, the C port keeps every fp32 step of the reference => bit-exact forward -----
def test_warp_fwd_bit_exact_vs_reference(golden_dir):
    for name in ("warp_small.npz", "warp_probe.npz"):
        d = load(golden_dir, name)
        out = oracle.warp_fwd(d["img"], d["off"])
        assert np.array_equal(out, d["out"]), name


def test_warp_bwd_vs_reference_autograd(golden_dir):
    d = load(golden_dir, "warp_small.npz")
    gimg, goff = oracle.warp_bwd(d["img"], d["off"], d["gout"])
    assert rel_err(gimg, d["gimg"]) <= 1e-5
    assert rel_err(goff, d["goff"]) <= 1e-5
    p = load(golden_dir, "warp_probe.npz")
    gimg, goff = oracle.warp_bwd(p["img"], p["off"], np.ones_like(p["out"]))
    assert np.array_equal(gimg, p["gimg"]) and np.array_equal(goff, p["goff"])


def test_warp_probe_known_answers(golden_dir):
    p = load(golden_dir, "warp_probe.npz")
    assert p["out"].ravel().tolist() == [10.0, 11.0, 12.25, 13.5, 0.0, 0.0]
    assert p["goff"].ravel().tolist() == [0.0, 1.0, 1.0, 1.0, 0.0, 0.0]
    assert p["gimg"].ravel().tolist() == [1.0, 1.0, 0.75, 0.75, 0.5, 0.0]


def test_warp_torch_ref_matches_reference(golden_dir):
    d = load(golden_dir, "warp_small.npz")
    img = torch.from_numpy(d["img"]).requires_grad_(True)
    off = torch.from_numpy(d["off"]).requires_grad_(True)
    out = torch_ref.warp_ref(img, off)
    assert torch.equal(out.detach(), torch.from_numpy(d["out"]))
    gimg, goff = torch.autograd.grad(out, (img, off), torch.from_numpy(d["gout"]))
    assert rel_err(gimg.numpy(), d["gimg"]) <= 1e-5 and rel_err(goff.numpy(), d["goff"]) <= 1e-5


# ---- a2 concat volume: pinned, bit-exact ----------------------------------------------------------
def test_concat_fwd_bit_exact(golden_dir):
    d = load(golden_dir, "concat_small.npz")
    cost = oracle.concat_fwd(d["ref"], d["tgt"], int(d["ndisp"]))
    assert np.array_equal(cost, d["cost"])
    t = torch_ref.concat_ref(torch.from_numpy(d["ref"]), torch.from_numpy(d["tgt"]), int(d["ndisp"]))
    assert np.array_equal(t.numpy(), d["cost"])


def test_matchshifted_slice(golden_dir):
    d = load(golden_dir, "matchshifted_small.npz")
    s = int(d["shift"])
    cost = oracle.concat_fwd(d["left"], d["right"], s + 1)
    assert np.array_equal(cost[:, :, s:s + 1], d["out"])


def test_concat_bwd(golden_dir):
    d = load(golden_dir, "concat_small.npz")
    gref, gtgt = oracle.concat_bwd(d["gcost"])
    assert rel_err(gref, d["gref"]) <= 1e-6 and rel_err(gtgt, d["gtgt"]) <= 1e-6


# ---- a3 disparityregression / soft-argmin: pinned ---------------------------------------------------
def test_softargmin_and_dispreg(golden_dir):
    d = load(golden_dir, "softargmin_small.npz")
    assert rel_err(oracle.softargmin_fwd(d["cost"]), d["out"]) <= 1e-5
    assert abs(float(oracle.softargmin_fwd(d["cost"])[0, 0, 0]) - 7.0) < 1e-5  # one-hot KAT
    assert rel_err(oracle.softargmin_bwd(d["cost"], d["gout"]), d["gcost"]) <= 1e-5
    assert rel_err(oracle.dispreg_fwd(d["x"]), d["dispreg_out"]) <= 1e-6
    assert np.array_equal(oracle.dispreg_bwd(d["gout"], d["x"].shape[1]), d["gx"])
    t = torch_ref.softargmin_ref(torch.from_numpy(d["cost"]))
    assert rel_err(t.numpy(), d["out"]) <= 1e-6


# ---- a1 correlation: UNPINNED by the reference; C port vs fp64 restatement + analytic KATs --------
@pytest.mark.parametrize("case", ["p1x8", "p1x7", "p3x5", "p1x5d2"])
def test_corr_c_port_vs_fp64_restatement(golden_dir, case):
    d = load(golden_dir, "corr_small_unpinned.npz")
    assert not bool(d["pinned"])
    patch = tuple(int(v) for v in d[f"{case}_patch"])
    dil = int(d[f"{case}_dil"])
    out = oracle.corr_fwd(d["in1"], d["in2"], patch_size=patch, dilation_patch=dil)
    assert out.shape == d[f"{case}_out"].shape
    assert rel_err(out, d[f"{case}_out"]) <= 1e-6
    g1, g2 = oracle.corr_bwd(d["in1"], d["in2"], d[f"{case}_gout"], patch_size=patch, dilation_patch=dil)
    assert rel_err(g1, d[f"{case}_g1"]) <= 1e-6 and rel_err(g2, d[f"{case}_g2"]) <= 1e-6


@pytest.mark.parametrize("P", [17, 40, 192])
def test_corr_all_ones_kat(P):
    """in1 = in2 = 1  =>  out = C inside the image, 0 outside, window -(P-1)//2 .. P-1-(P-1)//2."""
    B, C, H, W = 1, 3, 2, 64
    ones = np.ones((B, C, H, W), np.float32)
    out = oracle.corr_fwd(ones, ones, patch_size=(1, P))
    r = (P - 1) // 2
    for p in range(P):
        s = p - r
        w = np.arange(W)
        expect = np.where((w + s >= 0) & (w + s < W), float(C), 0.0)
        assert np.array_equal(out[0, 0, p, 0], expect), (P, p)
    assert int(out.sum()) == C * H * sum(max(0, W - abs(p - r)) for p in range(P))


def test_corr_impulse_kat():
    B, C, H, W, P = 1, 4, 3, 32, 8
    r = (P - 1) // 2
    for s in (-3, 0, 4):
        a = np.zeros((B, C, H, W), np.float32)
        b = np.zeros((B, C, H, W), np.float32)
        a[0, 2, 1, 10] = 1.0
        b[0, 2, 1, 10 + s] = 1.0
        out = oracle.corr_fwd(a, b, patch_size=(1, P))
        nz = np.argwhere(out != 0)
        assert nz.tolist() == [[0, 0, s + r, 1, 10]]


def test_corr_general_kernel_matches_conv_definition():
    """kernel_size/stride/padding path of the C port against an explicit unfold-based definition."""
    torch.manual_seed(0)
    a = torch.randn(1, 2, 7, 9)
    b = torch.randn(1, 2, 7, 9)
    out = oracle.corr_fwd(a.numpy(), b.numpy(), kernel_size=3, patch_size=3, stride=2, padding=1)
    assert out.shape == (1, 3, 3, 4, 5)
    ap = torch.nn.functional.pad(a, (1 + 1, 1 + 1, 1 + 1, 1 + 1))
    bp = torch.nn.functional.pad(b, (1 + 1, 1 + 1, 1 + 1, 1 + 1))
    for ph in range(3):
        for pw in range(3):
            for h in range(4):
                for w in range(5):
                    u, v = h * 2 + 1, w * 2 + 1  # (-pad + h*stride) shifted by the 2-px zero border
                    x = ap[0, :, u:u + 3, v:v + 3]
                    y = bp[0, :, u + ph - 1:u + ph - 1 + 3, v + pw - 1:v + pw - 1 + 3]
                    assert abs(float((x * y).sum()) - float(out[0, ph, pw, h, w])) < 1e-4


# ---- f1 fused trilinear upsample + soft-argmin: pinned by the reference's own call sequence -------
def test_upsample_softargmin_oracle_vs_reference(golden_dir):
    d = load(golden_dir, "upsoftargmin_small.npz")
    out = oracle.upsample_softargmin_fwd(d["cost3"], int(d["maxdisp"]), d["size"])
    assert rel_err(out, d["pred"]) <= 1e-5


def test_upsample_softargmin_bwd_oracle_vs_reference_autograd(golden_dir):
    d = load(golden_dir, "upsoftargmin_small.npz")
    g = oracle.upsample_softargmin_bwd(d["cost3"], d["gpred"], int(d["maxdisp"]), d["size"])
    assert g.shape == d["gcost3"].shape
    assert rel_err(g, d["gcost3"]) <= 1e-5



def test_warp_fused_torch_ref_matches_reference(golden_dir):
    """f4: the oracle's blend / photo-consistency restatements against fixtures produced by the reference's own
    apply_disparity + its blend expression / nn.MSELoss, gradients by autograd (oracle/make_golden.py::make_warp_fused)."""
    import torch

    from oracle import torch_ref

    d = np.load(os.path.join(golden_dir, "warp_fused_small.npz"))
    t = lambda k: torch.from_numpy(d[k]).requires_grad_(True)
    seg, seg_r, disp, att = t("seg"), t("seg_r"), t("disp"), t("att")
    both, warped = torch_ref.warp_blend_ref(seg, seg_r, -disp, att)
    assert np.array_equal(both.detach().numpy(), d["both"]) and np.array_equal(warped.detach().numpy(), d["warped"])
    gs, gr, gd, ga = torch.autograd.grad((both, warped), (seg, seg_r, disp, att),
                                         (torch.from_numpy(d["gboth"]), torch.from_numpy(d["gwarped"])))
    for got, key in ((gs, "gseg"), (gr, "gseg_r"), (gd, "gdisp"), (ga, "gatt")):
        assert rel_err(got.numpy(), d[key]) <= FP32_TOL, key
    for name, mask in (("plain", False), ("masked", True)):
        left, right, dd = t("left"), t("right"), t("disp")
        loss = torch_ref.photo_mse_ref(right, -dd, left, mask)
        assert abs(float(loss) - float(d[f"mse_{name}_loss"])) <= 1e-6 * float(d[f"mse_{name}_loss"])
        gl, grr, gdd = torch.autograd.grad(3.0 * loss, (left, right, dd))
        for got, key in ((gl, "gleft"), (grr, "gright"), (gdd, "gdisp")):
            assert rel_err(got.numpy(), d[f"mse_{name}_{key}"]) <= FP32_TOL, (name, key)
