"""Pure-PyTorch restatements of the hot-path ops -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Differentiable with autograd and dtype-generic, so tests can run them in fp64 to separate the
oracle's own rounding from the CUDA kernels' rounding, and can obtain reference gradients without
a hand-written backward.  Each function cites the reference lines it follows.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def corr_ref(in1: torch.Tensor, in2: torch.Tensor, patch_size=1, dilation_patch=1) -> torch.Tensor:
    """SpatialCorrelationSampler with kernel_size=1, stride=1, padding=0, dilation=1.

    Restates the upstream definition as used at models/dsnet_t2.py:129-133,221-223 and :841-851,879:
    out[n,ph,pw,h,w] = sum_c in1[n,c,h,w] * in2[n,c,h+sh,w+sw], OOB terms skipped, not normalised;
    sh = (ph-(pH-1)//2)*dpH, sw = (pw-(pW-1)//2)*dpW.  PARITY UNPINNED (third-party package absent).
    """
    pH, pW = _pair(patch_size)
    dpH, dpW = _pair(dilation_patch)
    B, C, H, W = in1.shape
    rH, rW = (pH - 1) // 2, (pW - 1) // 2
    out = in1.new_zeros((B, pH, pW, H, W))
    for ph in range(pH):
        sh = (ph - rH) * dpH
        h_lo, h_hi = max(0, -sh), min(H, H - sh)
        if h_lo >= h_hi:
            continue
        for pw in range(pW):
            sw = (pw - rW) * dpW
            w_lo, w_hi = max(0, -sw), min(W, W - sw)
            if w_lo >= w_hi:
                continue
            a = in1[:, :, h_lo:h_hi, w_lo:w_hi]
            b = in2[:, :, h_lo + sh:h_hi + sh, w_lo + sw:w_hi + sw]
            out[:, ph, pw, h_lo:h_hi, w_lo:w_hi] = (a * b).sum(dim=1)
    return out


def concat_ref(ref: torch.Tensor, tgt: torch.Tensor, ndisp: int) -> torch.Tensor:
    """The slice-assign loop of models_psmnet/stackhourglass.py:110-119 (device-agnostic)."""
    B, C, H, W = ref.shape
    cost = ref.new_zeros((B, 2 * C, ndisp, H, W))
    for i in range(ndisp):
        if i > 0:
            cost[:, :C, i, :, i:] = ref[:, :, :, i:]
            cost[:, C:, i, :, i:] = tgt[:, :, :, :-i]
        else:
            cost[:, :C, i, :, :] = ref
            cost[:, C:, i, :, :] = tgt
    return cost.contiguous()


def dispreg_ref(x: torch.Tensor) -> torch.Tensor:
    """disparityregression.forward, models_psmnet/submodule.py:61-64 (without the .cuda())."""
    D = x.shape[1]
    disp = torch.arange(D, dtype=x.dtype, device=x.device).view(1, D, 1, 1)
    disp = disp.repeat(x.size(0), 1, x.size(2), x.size(3))
    return torch.sum(x * disp, 1)


def softargmin_ref(cost: torch.Tensor) -> torch.Tensor:
    """F.softmax(cost, dim=1) -> disparityregression, models_psmnet/stackhourglass.py:151,155."""
    return dispreg_ref(F.softmax(cost, dim=1))


def warp_ref(img: torch.Tensor, off: torch.Tensor) -> torch.Tensor:
    """Closed form of apply_disparity(wrap_mode='edge'), models/torch_dsnet.py:25-84.

    Valid while N*H*W < 2**24 (the reference builds its gather indices in float32); differentiable
    w.r.t. both arguments exactly like the reference graph (floor/clamp sub-gradients included).
    """
    N, C, H, W = img.shape
    x = torch.arange(W, dtype=img.dtype, device=img.device).view(1, 1, 1, W) + off
    x = torch.clamp(x, 0.0, W - 1)
    x0 = torch.floor(x)
    x1 = torch.clamp(x0 + 1, max=W - 1)
    il = x0.long().expand(N, C, H, W)
    ir = x1.long().expand(N, C, H, W)
    pl = torch.gather(img, 3, il)
    pr = torch.gather(img, 3, ir)
    return (x1 - x) * pl + (x - x0) * pr


def warp_blend_ref(seg_left, seg_right, off, att):
    """models/dsnet_t2_warp.py:697-698: (1 - at_d) * seg_branch + at_d * apply_disparity(seg_branch_right, off)."""
    warped = warp_ref(seg_right, off)
    return (1 - att) * seg_left + att * warped, warped


def photo_mse_ref(right, off, left, mask_positive_disparity=False):
    """torch_implementation.py:314-317 with warped_right = apply_disparity(right, off) [* (disp > 0), disp = -off,
    models/dsnet_t2_warp.py:811]."""
    wr = warp_ref(right, off)
    if mask_positive_disparity:
        wr = wr * (off < 0)
    return F.mse_loss(wr, left)


def corr_conv_relu_ref(in1, in2, weight, patch_size=(1, 17)):
    """models/dsnet_t2.py:1187-1197 for `1dcorr`: squeeze(correlation_sampler(a, b), 1) -> corrConv2d (bias-free 1x1
    convolution, models/torch_model.py:236-272 with kernel 1 => no padding) -> ReLU."""
    y = torch.squeeze(corr_ref(in1, in2, patch_size), dim=1)
    return F.relu(F.conv2d(y, weight.view(weight.size(0), -1, 1, 1)))
