"""-m gpu parity of the correlation kernels (through the C ABI) against the CPU oracle, the golden fixtures
and size-independent properties at the BASELINE.json sizes."""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_ref
from tests.util import FP32_TOL, npy, rel_err, vp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pmt():
    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as m
    m.load_library()
    return m


def run_corr(pmt, L, R, G, patch, dil=1):
    dev = torch.device("cuda:0")
    Ld = torch.from_numpy(L).to(dev).requires_grad_(True)
    Rd = torch.from_numpy(R).to(dev).requires_grad_(True)
    s = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=patch, stride=1, padding=0, dilation=1,
                                      dilation_patch=dil)
    out = s(Ld, Rd)
    out.backward(torch.from_numpy(G).to(dev))
    torch.cuda.synchronize()
    return npy(out), npy(Ld.grad), npy(Rd.grad)


# (B, C, H, W, P, engine the default entry points pick: 2 tensor core, 1 CUDA-core tiled, 0 generic)
CORR1D_CASES = [
    (2, 64, 64, 128, 40, 2),     # BASELINE config 1
    (2, 352, 32, 64, 17, 2),     # production call inside minidsnetExt (C > 128: three channel blocks on the tensor cores)
    (1, 130, 2, 260, 33, 2),     # two channel blocks, the second with 2 live channels; W%128 != 0
    (1, 64, 5, 512, 192, 2),     # headline row shape
    (1, 16, 3, 512, 193, 2),     # largest P of the fast paths
    (1, 5, 3, 100, 8, 2),        # ragged: W%64 != 0, C%16 != 0, even P
    (2, 33, 2, 68, 7, 2),        # C%32 == 1
    (1, 3, 1, 4, 1, 2),          # degenerate patch
    (1, 7, 4, 260, 2, 2),
    (1, 40, 3, 132, 100, 2),
    (1, 128, 3, 960, 192, 2),    # config-4 row geometry (C=128, W=960)
    (1, 4, 3, 62, 9, 0),         # W%4 != 0 -> generic kernels
    (1, 4, 3, 64, 250, 0),       # P beyond the tiled band -> generic kernels
]
TF32_TOL = 5e-3  # plain-TF32 tensor-core variant (north_star allows 2e-2 for the reduced-precision variant)


@pytest.fixture(params=["auto", "simt"])
def engine(request, pmt):
    prev = pmt.set_correlation_engine(request.param)
    yield request.param
    pmt.set_correlation_engine(prev)


@pytest.mark.parametrize("B,C,H,W,P,fast", CORR1D_CASES)
def test_corr1d_vs_oracle(pmt, engine, B, C, H, W, P, fast):
    rng = np.random.default_rng(B * 1000 + C * 7 + W + P)
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    G = rng.standard_normal((B, 1, P, H, W), dtype=np.float32)
    lib = pmt.load_library()
    t = torch.from_numpy(L).cuda()
    assert lib.pmt_corr1d_uses_fast_path(vp(t), vp(t), vp(t), C, H, W, P, 1) == fast
    out, g1, g2 = run_corr(pmt, L, R, G, (1, P))
    ref = oracle.corr_fwd(L, R, patch_size=(1, P))
    r1, r2 = oracle.corr_bwd(L, R, G, patch_size=(1, P))
    assert out.shape == (B, 1, P, H, W)
    assert rel_err(out, ref) <= FP32_TOL
    assert rel_err(g1, r1) <= FP32_TOL
    assert rel_err(g2, r2) <= FP32_TOL
    # exact zeros where the shifted column leaves the image (sampler skips those terms)
    rW = (P - 1) // 2
    for p in (0, P - 1):
        s = p - rW
        w = np.arange(W)
        oob = (w + s < 0) | (w + s >= W)
        assert np.all(out[:, 0, p][..., oob] == 0.0)


@pytest.mark.parametrize("B,C,H,W,P", [(2, 64, 64, 128, 40), (1, 64, 5, 512, 192), (1, 128, 3, 960, 192), (1, 5, 3, 100, 8)])
def test_corr1d_tf32_variant(pmt, B, C, H, W, P):
    """Reduced-precision tensor-core variant (plain TF32 inputs, fp32 accumulate) within its looser tolerance."""
    rng = np.random.default_rng(P)
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    G = rng.standard_normal((B, 1, P, H, W), dtype=np.float32)
    prev = pmt.set_correlation_engine("tf32")
    try:
        out, g1, g2 = run_corr(pmt, L, R, G, (1, P))
    finally:
        pmt.set_correlation_engine(prev)
    r1, r2 = oracle.corr_bwd(L, R, G, patch_size=(1, P))
    e = (rel_err(out, oracle.corr_fwd(L, R, patch_size=(1, P))), rel_err(g1, r1), rel_err(g2, r2))
    assert max(e) <= TF32_TOL and min(e) > 1e-6   # really the TF32 engine, not the fp32 one
    with pytest.raises(Exception):
        prev = pmt.set_correlation_engine("tf32")
        try:
            bad = torch.zeros(1, 4, 3, 62, device="cuda:0")     # W % 4 != 0: no silent fallback
            pmt.spatial_correlation_sample(bad, bad, patch_size=(1, 9))
        finally:
            pmt.set_correlation_engine(prev)


@pytest.mark.parametrize("patch,dil", [((3, 5), 1), ((1, 5), 2), ((17, 17), 1), ((1, 21), 4), ((5, 1), (2, 1))])
def test_corr_general_patch_vs_oracle(pmt, patch, dil):
    rng = np.random.default_rng(11)
    B, C, H, W = 2, 6, 12, 40
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    G = rng.standard_normal((B, patch[0], patch[1], H, W), dtype=np.float32)
    out, g1, g2 = run_corr(pmt, L, R, G, patch, dil)
    ref = oracle.corr_fwd(L, R, patch_size=patch, dilation_patch=dil)
    r1, r2 = oracle.corr_bwd(L, R, G, patch_size=patch, dilation_patch=dil)
    assert rel_err(out, ref) <= FP32_TOL and rel_err(g1, r1) <= FP32_TOL and rel_err(g2, r2) <= FP32_TOL


@pytest.mark.parametrize("case", ["p1x8", "p1x7", "p3x5", "p1x5d2"])
def test_corr_golden_fp64(pmt, golden_dir, case):
    d = np.load(os.path.join(golden_dir, "corr_small_unpinned.npz"))
    patch = tuple(int(v) for v in d[f"{case}_patch"])
    dil = int(d[f"{case}_dil"])
    out, g1, g2 = run_corr(pmt, d["in1"], d["in2"], d[f"{case}_gout"].astype(np.float32), patch, dil)
    assert rel_err(out, d[f"{case}_out"]) <= FP32_TOL
    assert rel_err(g1, d[f"{case}_g1"]) <= FP32_TOL and rel_err(g2, d[f"{case}_g2"]) <= FP32_TOL


def test_corr_kats(pmt):
    dev = torch.device("cuda:0")
    for P in (17, 40, 192):
        B, C, H, W = 1, 3, 2, 256
        ones = torch.ones(B, C, H, W, device=dev)
        out = npy(pmt.spatial_correlation_sample(ones, ones, patch_size=(1, P)))
        r = (P - 1) // 2
        w = np.arange(W)
        for p in range(P):
            expect = np.where((w + p - r >= 0) & (w + p - r < W), float(C), 0.0)
            assert np.array_equal(out[0, 0, p, 0], expect), (P, p)
    # impulse: single non-zero at the plane of the relative shift
    a = torch.zeros(1, 4, 3, 64, device=dev)
    b = torch.zeros(1, 4, 3, 64, device=dev)
    a[0, 2, 1, 10] = 1.0
    b[0, 2, 1, 14] = 1.0
    out = npy(pmt.spatial_correlation_sample(a, b, patch_size=(1, 40)))
    assert np.argwhere(out != 0).tolist() == [[0, 0, 4 + 19, 1, 10]]


def test_corr_headline_shape_one_pair_vs_oracle(pmt, engine):
    """One full 256x512, C=64, P=192 pair against the C oracle (a few seconds of CPU)."""
    rng = np.random.default_rng(5)
    B, C, H, W, P = 1, 64, 256, 512, 192
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    G = rng.standard_normal((B, 1, P, H, W), dtype=np.float32)
    out, g1, g2 = run_corr(pmt, L, R, G, (1, P))
    assert rel_err(out, oracle.corr_fwd(L, R, patch_size=(1, P))) <= FP32_TOL
    r1, r2 = oracle.corr_bwd(L, R, G, patch_size=(1, P))
    assert rel_err(g1, r1) <= FP32_TOL and rel_err(g2, r2) <= FP32_TOL


def test_corr_full_size_properties(pmt, engine):
    """BASELINE sizes (B=4, C=64, 256x512, D=192): adjoint identity, linearity, determinism, all-ones count."""
    dev = torch.device("cuda:0")
    B, C, H, W, P = 4, 64, 256, 512, 192
    g = torch.Generator(device=dev).manual_seed(0)
    L = torch.randn(B, C, H, W, device=dev, generator=g)
    R = torch.randn(B, C, H, W, device=dev, generator=g)
    G = torch.randn(B, 1, P, H, W, device=dev, generator=g)
    Lr, Rr = L.clone().requires_grad_(True), R.clone().requires_grad_(True)
    out = pmt.spatial_correlation_sample(Lr, Rr, patch_size=(1, P))
    out.backward(G)
    # <corr(L,R), G> == <L, gL> == <R, gR>  (bilinear form; double accumulation of the dot products)
    lhs = (out.detach().double() * G.double()).sum().item()
    d1 = (L.double() * Lr.grad.double()).sum().item()
    d2 = (R.double() * Rr.grad.double()).sum().item()
    scale = (out.detach().double().abs() * G.double().abs()).sum().item()
    assert abs(lhs - d1) / scale < 1e-6 and abs(lhs - d2) / scale < 1e-6
    # linearity in the first argument
    L2 = torch.randn(B, C, H, W, device=dev, generator=g)
    o2 = pmt.spatial_correlation_sample(L2, R, patch_size=(1, P))
    o12 = pmt.spatial_correlation_sample(0.5 * L + 2.0 * L2, R, patch_size=(1, P))
    num = (o12 - (0.5 * out.detach() + 2.0 * o2)).abs().max().item()
    assert num / o12.abs().max().item() < 1e-5
    # run-to-run bit reproducibility of the gather backward
    g1a, g2a = Lr.grad.clone(), Rr.grad.clone()
    Lr.grad = None
    Rr.grad = None
    pmt.spatial_correlation_sample(Lr, Rr, patch_size=(1, P)).backward(G)
    assert torch.equal(g1a, Lr.grad) and torch.equal(g2a, Rr.grad)
    # all-ones: number of in-bounds terms
    ones = torch.ones(1, C, H, W, device=dev)
    cnt = pmt.spatial_correlation_sample(ones, ones, patch_size=(1, P)).double().sum().item()
    r = (P - 1) // 2
    assert cnt == C * H * sum(max(0, W - abs(p - r)) for p in range(P))


def test_corr_c_abi_direct_and_errors(pmt):
    """Call the C ABI without the autograd shim; check argument validation surfaces as non-zero status."""
    lib = pmt.load_library()
    dev = torch.device("cuda:0")
    B, C, H, W, P = 1, 8, 4, 64, 9
    L = torch.randn(B, C, H, W, device=dev)
    R = torch.randn(B, C, H, W, device=dev)
    out = torch.empty(B, 1, P, H, W, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pmt_corr1d_fwd_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, st) == 0
    torch.cuda.synchronize()
    assert rel_err(npy(out), oracle.corr_fwd(npy(L), npy(R), patch_size=(1, P))) <= FP32_TOL
    assert lib.pmt_corr1d_fwd_f32(vp(L), vp(R), vp(out), B, C, H, W, 0, 1, st) != 0
    assert b"patch" in lib.pmt_last_error()
    assert lib.pmt_corr1d_fwd_f32(None, vp(R), vp(out), B, C, H, W, P, 1, st) != 0


def test_corr_host_entry_point(pmt):
    lib = pmt.load_library()
    B, C, H, W, P = 3, 16, 8, 64, 24
    g = torch.Generator().manual_seed(3)
    L = torch.randn(B, C, H, W, generator=g).pin_memory()
    R = torch.randn(B, C, H, W, generator=g).pin_memory()
    G = torch.randn(B, 1, P, H, W, generator=g).pin_memory()
    out = torch.empty(B, 1, P, H, W).pin_memory()
    g1 = torch.empty_like(L).pin_memory()
    g2 = torch.empty_like(R).pin_memory()
    rc = lib.pmt_corr1d_fwd_bwd_host_f32(vp(L), vp(R), vp(G), vp(out), vp(g1), vp(g2), B, C, H, W, P, 1)
    assert rc == 0, lib.pmt_last_error()
    assert rel_err(out.numpy(), oracle.corr_fwd(L.numpy(), R.numpy(), patch_size=(1, P))) <= FP32_TOL
    r1, r2 = oracle.corr_bwd(L.numpy(), R.numpy(), G.numpy(), patch_size=(1, P))
    assert rel_err(g1.numpy(), r1) <= FP32_TOL and rel_err(g2.numpy(), r2) <= FP32_TOL


def test_corr_module_surface(pmt):
    dev = torch.device("cuda:0")
    s = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 17), stride=1, padding=0, dilation_patch=1)
    assert list(s.parameters()) == [] and list(s.buffers()) == []
    a = torch.randn(2, 8, 4, 32, device=dev, requires_grad=True)
    b = torch.randn(2, 8, 4, 32, device=dev)  # no grad on the second input
    y = s(a, b)
    assert y.shape == (2, 1, 17, 4, 32) and y.is_contiguous()
    y = torch.squeeze(y, dim=1)  # as models/dsnet_t2.py:881
    y.sum().backward()
    assert a.grad is not None and b.grad is None
    with pytest.raises(NotImplementedError):
        pmt.SpatialCorrelationSampler(kernel_size=3, patch_size=1)(a, b)
    with pytest.raises(NotImplementedError):
        s(a.half(), b.half())
    # non-contiguous inputs are accepted like upstream (made contiguous)
    at = torch.randn(2, 4, 32, 8, device=dev).permute(0, 3, 1, 2)
    y2 = s(at, at)
    assert rel_err(npy(y2), oracle.corr_fwd(npy(at), npy(at), patch_size=(1, 17))) <= FP32_TOL


def test_corr_fp64_restatement_bound(pmt):
    """Separate oracle rounding from ours: both fp32 results must sit within 1e-5 of the fp64 restatement."""
    rng = np.random.default_rng(2)
    B, C, H, W, P = 1, 64, 4, 128, 40
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    ref64 = torch_ref.corr_ref(torch.from_numpy(L).double(), torch.from_numpy(R).double(), (1, P)).numpy()
    out = npy(pmt.spatial_correlation_sample(torch.from_numpy(L).cuda(), torch.from_numpy(R).cuda(),
                                             patch_size=(1, P)))
    assert rel_err(out, ref64) <= FP32_TOL
    assert rel_err(oracle.corr_fwd(L, R, patch_size=(1, P)), ref64) <= FP32_TOL


@pytest.mark.parametrize("B,C,H,W,pH", [(2, 352, 32, 64, 17), (1, 40, 5, 120, 17), (1, 8, 20, 16, 5)])
def test_corr2d_row_pass_kernels_vs_oracle(pmt, B, C, H, W, pH):
    """f3: (pH,17) patches run as row-pair passes through shared memory (csrc/corr2d_rows.cu), incl. the full
    `-corrType 2dcorr` call of the reference (patch (17,17), C=352, 32x64: models/dsnet_t2.py:129-133,221-223)."""
    rng = np.random.default_rng(C + W + pH)
    L = rng.standard_normal((B, C, H, W), dtype=np.float32)
    R = rng.standard_normal((B, C, H, W), dtype=np.float32)
    G = rng.standard_normal((B, pH, 17, H, W), dtype=np.float32)
    out, g1, g2 = run_corr(pmt, L, R, G, (pH, 17))
    ref = oracle.corr_fwd(L, R, patch_size=(pH, 17))
    r1, r2 = oracle.corr_bwd(L, R, G, patch_size=(pH, 17))
    assert rel_err(out, ref) <= FP32_TOL and rel_err(g1, r1) <= FP32_TOL and rel_err(g2, r2) <= FP32_TOL
