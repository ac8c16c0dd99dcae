"""-m gpu: f2, the 1 x 17 correlation fused with corrConv2d (bias-free 1x1 convolution) + ReLU
(models/dsnet_t2.py:1187-1197) against oracle o F.conv2d o ReLU, forward and all three gradients."""
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


# (B, C, H, W, O): production call of minidsnetExt; the 540x960 input of config 4 at 1/8 (68 x 120); ragged channel chunk
CASES = [(4, 352, 32, 64, 128), (1, 352, 68, 120, 128), (2, 37, 3, 16, 8), (1, 64, 5, 128, 128), (1, 40, 2, 32, 256)]


@pytest.mark.parametrize("B,C,H,W,O", CASES)
def test_corr_conv_relu_vs_oracle(pmt, B, C, H, W, O):
    g = torch.Generator().manual_seed(C + W)
    a = torch.randn(B, C, H, W, generator=g)
    b = torch.randn(B, C, H, W, generator=g)
    wt = torch.randn(O, 17, 1, 1, generator=g) * (2.0 / O) ** 0.5
    gz = torch.randn(B, O, H, W, generator=g)
    outs = []
    for _ in range(2):
        ta, tb, tw = (x.to(DEV).requires_grad_(True) for x in (a, b, wt))
        z = pmt.correlation_conv1x1_relu(ta, tb, tw, patch_size=(1, 17))
        z.backward(gz.to(DEV))
        outs.append((z.detach(), ta.grad, tb.grad, tw.grad))
    for x, y in zip(*outs):
        assert torch.equal(x, y)                                    # deterministic, run to run
    a64, b64, w64 = (x.double().requires_grad_(True) for x in (a, b, wt))
    z64 = torch_ref.corr_conv_relu_ref(a64, b64, w64, (1, 17))
    z64.backward(gz.double())
    z, ga, gb, gw = outs[0]
    assert z.shape == (B, O, H, W) and float(z.min()) >= 0.0
    assert rel_err(npy(z), z64.detach().numpy()) <= FP32_TOL
    # the ReLU mask of pixels within rounding of 0 may differ between fp32 and fp64: compare gradients with the fp32 mask
    mask = (z > 0).cpu()
    a64.grad = b64.grad = w64.grad = None
    y64 = torch.nn.functional.conv2d(torch.squeeze(torch_ref.corr_ref(a64, b64, (1, 17)), 1), w64)
    (y64 * mask).backward(gz.double())
    assert rel_err(npy(ga), a64.grad.numpy()) <= FP32_TOL
    assert rel_err(npy(gb), b64.grad.numpy()) <= FP32_TOL
    assert rel_err(npy(gw), w64.grad.numpy()) <= FP32_TOL


def test_module_shares_the_reference_conv_weight_and_matches_the_unfused_ops(pmt):
    """CorrelationConvReLU.from_reference(sampler, corrConv2d): same Parameter object; output equals this package's
    sampler followed by torch's conv2d + ReLU."""
    from torch import nn

    sampler = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 17), stride=1, padding=0, dilation_patch=1)
    conv = nn.Sequential(nn.Conv2d(17, 128, 1, bias=False), nn.ReLU(inplace=True)).to(DEV)
    fused = pmt.CorrelationConvReLU.from_reference(sampler, conv)
    assert fused.weight is conv[0].weight
    a = torch.randn(2, 352, 8, 64, device=DEV)
    b = torch.randn(2, 352, 8, 64, device=DEV)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # cuDNN would otherwise run the reference convolution in TF32 (~1e-3)
    try:
        want = conv(torch.squeeze(sampler(a, b), dim=1))
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    got = fused(a, b)
    assert rel_err(npy(got), npy(want)) <= FP32_TOL


def test_unsupported_shapes_raise(pmt):
    a = torch.randn(1, 8, 2, 64, device=DEV)
    with pytest.raises(NotImplementedError):
        pmt.correlation_conv1x1_relu(a, a, torch.randn(4, 9, device=DEV), patch_size=(1, 9))
    with pytest.raises(NotImplementedError):
        pmt.correlation_conv1x1_relu(a[..., :12].contiguous(), a[..., :12].contiguous(), torch.randn(4, 17, device=DEV))
    big = torch.randn(1, 8, 2, 128, device=DEV)
    with pytest.raises(NotImplementedError):                       # (O, W) gradient tile of the backward exceeds shared memory
        pmt.correlation_conv1x1_relu(big, big, torch.randn(256, 17, device=DEV))
