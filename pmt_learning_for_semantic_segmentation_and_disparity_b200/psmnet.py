"""PSMNet hot-path ops on B200: concat cost volume, disparityregression, fused soft-argmin.

Reference surfaces kept: ``disparityregression(maxdisp)(x)`` and ``matchshifted()(left, right, shift)`` of
models_psmnet/submodule.py:45-64; ``build_concat_volume(ref, tgt, ndisp)`` replaces the slice-assign loop of
models_psmnet/stackhourglass.py:110-119; ``softargmin(cost)`` replaces the ``F.softmax(cost, dim=1)`` +
``disparityregression`` pair of stackhourglass.py:142-155.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _util as U


class _ConcatVolume(Function):
    @staticmethod
    def forward(ctx, ref, tgt, ndisp, first_disp):
        ref = U.require_cuda_f32(ref, "ref")
        tgt = U.require_cuda_f32(tgt, "tgt")
        if ref.dim() != 4 or ref.shape != tgt.shape:
            raise ValueError(f"ref/tgt must be 4-D (B,C,H,W) of equal shape, got {tuple(ref.shape)}, {tuple(tgt.shape)}")
        dev = U.same_device(ref, tgt)
        B, C, H, W = ref.shape
        ctx.dims = (B, C, int(ndisp), H, W, int(first_disp))
        cost = torch.empty((B, 2 * C, int(ndisp), H, W), device=dev, dtype=torch.float32)
        U.call("pmt_concat_volume_fwd_f32", dev, U.ptr(ref), U.ptr(tgt), U.ptr(cost), B, C, int(ndisp), H, W,
               int(first_disp))
        return cost

    @staticmethod
    @once_differentiable
    def backward(ctx, gcost):
        B, C, D, H, W, d0 = ctx.dims
        g = U.require_cuda_f32(gcost, "grad_cost")
        gref = torch.empty((B, C, H, W), device=g.device, dtype=torch.float32)
        gtgt = torch.empty_like(gref)
        U.call("pmt_concat_volume_bwd_f32", g.device, U.ptr(g), U.ptr(gref), U.ptr(gtgt), B, C, D, H, W, d0)
        return gref, gtgt, None, None


def build_concat_volume(ref: torch.Tensor, tgt: torch.Tensor, ndisp: int) -> torch.Tensor:
    """(B,C,H,W) x 2 -> contiguous (B,2C,ndisp,H,W); ndisp = maxdisp//4 in PSMNet. Bit-exact vs the loop."""
    if int(ndisp) < 0:
        raise ValueError("ndisp must be >= 0")
    return _ConcatVolume.apply(ref, tgt, int(ndisp), 0)


class matchshifted(nn.Module):
    """models_psmnet/submodule.py:45-54 -- one shifted left||right slice, (B,2C,1,H,W)."""

    def forward(self, left, right, shift):
        shift = int(shift)
        if shift < 0 or shift > left.size(3):
            raise ValueError(f"shift {shift} outside [0, width]")
        return _ConcatVolume.apply(left, right, 1, shift)


class _DispReg(Function):
    @staticmethod
    def forward(ctx, x):
        x = U.require_cuda_f32(x, "x")
        B, D, H, W = x.shape
        ctx.dims = (B, D, H, W)
        out = torch.empty((B, H, W), device=x.device, dtype=torch.float32)
        U.call("pmt_dispreg_fwd_f32", x.device, U.ptr(x), U.ptr(out), B, D, H, W)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        B, D, H, W = ctx.dims
        g = U.require_cuda_f32(gout, "grad_output")
        gx = torch.empty((B, D, H, W), device=g.device, dtype=torch.float32)
        U.call("pmt_dispreg_bwd_f32", g.device, U.ptr(g), U.ptr(gx), B, D, H, W)
        return gx


class disparityregression(nn.Module):
    """models_psmnet/submodule.py:56-64: out[b,h,w] = sum_d x[b,d,h,w]*d over an already soft-maxed x."""

    def __init__(self, maxdisp):
        super().__init__()
        self.maxdisp = int(maxdisp)

    def forward(self, x):
        if x.dim() != 4:
            raise ValueError(f"expected (B,D,H,W), got {tuple(x.shape)}")
        if x.size(1) != self.maxdisp:
            # the reference multiplies by a (B,maxdisp,H,W) ramp, which fails to broadcast the same way
            raise RuntimeError(f"The size of tensor a ({x.size(1)}) must match the size of tensor b "
                               f"({self.maxdisp}) at non-singleton dimension 1")
        return _DispReg.apply(x)


class _SoftArgmin(Function):
    @staticmethod
    def forward(ctx, cost):
        cost = U.require_cuda_f32(cost, "cost")
        B, D, H, W = cost.shape
        out = torch.empty((B, H, W), device=cost.device, dtype=torch.float32)
        lse = torch.empty_like(out)
        U.call("pmt_softargmin_fwd_f32", cost.device, U.ptr(cost), U.ptr(out), U.ptr(lse), B, D, H, W)
        ctx.save_for_backward(cost, out, lse)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        cost, out, lse = ctx.saved_tensors
        B, D, H, W = cost.shape
        g = U.require_cuda_f32(gout, "grad_output")
        gcost = torch.empty_like(cost)
        U.call("pmt_softargmin_bwd_f32", cost.device, U.ptr(cost), U.ptr(out), U.ptr(lse), U.ptr(g), U.ptr(gcost),
               B, D, H, W)
        return gcost


def softargmin(cost: torch.Tensor) -> torch.Tensor:
    """Fused softmax(dim=1) + disparity expectation: (B,D,H,W) -> (B,H,W), one pass over `cost`."""
    if cost.dim() != 4 or cost.size(1) < 1:
        raise ValueError(f"expected (B,D,H,W) with D>=1, got {tuple(cost.shape)}")
    return _SoftArgmin.apply(cost)


class _UpsampleSoftArgmin(Function):
    """Forward: fused kernel (the (B,D,H,W) volume is never written).  Backward: fused as well: a kernel
    re-interpolates the logits per pixel and writes the gradient w.r.t. the Dq sampled source planes into a (B,Dq,H,W)
    workspace -- 4x smaller than the volume for PSMNet -- and a second one applies the adjoint of the spatial
    interpolation.  Shapes the fused backward does not cover (pmt_upsample_softargmin_bwd_supported() == 0: shared-memory
    footprint, interpolation windows > 40) raise NotImplementedError when a gradient is requested -- there is no ATen
    fallback; such a caller keeps the reference's unfused sequence (F.interpolate -> softargmin)."""

    @staticmethod
    def forward(ctx, cost_lowres, maxdisp, height, width):
        c = U.require_cuda_f32(cost_lowres, "cost")
        if c.dim() == 5:
            if c.size(1) != 1:
                raise ValueError(f"expected (B,1,Dq,Hq,Wq), got {tuple(c.shape)}")
            c4 = c[:, 0]
        elif c.dim() == 4:
            c4 = c
        else:
            raise ValueError(f"expected (B,1,Dq,Hq,Wq) or (B,Dq,Hq,Wq), got {tuple(c.shape)}")
        c4 = c4.contiguous()
        B, Dq, Hq, Wq = c4.shape
        D, H, W = int(maxdisp), int(height), int(width)
        out = torch.empty((B, H, W), device=c.device, dtype=torch.float32)
        lse = torch.empty((B, H, W), device=c.device, dtype=torch.float32)
        if ctx.needs_input_grad[0] and U._lib.load().pmt_upsample_softargmin_bwd_supported(B, Dq, Hq, Wq, D, H, W) != 1:
            raise NotImplementedError(
                f"upsample_softargmin: the fused backward does not cover (Dq,Hq,Wq)=({Dq},{Hq},{Wq}) -> (D,H,W)=({D},{H},{W}); "
                "use F.interpolate(..., mode='trilinear') followed by softargmin() for this shape (no silent fallback)")
        U.call("pmt_upsample_softargmin_fwd_f32", c.device, U.ptr(c4), U.ptr(out), U.ptr(lse), B, Dq, Hq, Wq, D, H, W)
        ctx.save_for_backward(cost_lowres, c4, out, lse)
        ctx.size = (D, H, W)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        cost_lowres, c4, out, lse = ctx.saved_tensors
        D, H, W = ctx.size
        B, Dq, Hq, Wq = c4.shape
        g = U.require_cuda_f32(gout, "grad_output")
        work = torch.empty((B, Dq, H, W), device=c4.device, dtype=torch.float32)
        glow = torch.empty_like(c4)
        U.call("pmt_upsample_softargmin_bwd_f32", c4.device, U.ptr(c4), U.ptr(out), U.ptr(lse), U.ptr(g), U.ptr(work),
               U.ptr(glow), B, Dq, Hq, Wq, D, H, W)
        return glow.view(cost_lowres.shape), None, None, None


def upsample_softargmin(cost_lowres: torch.Tensor, maxdisp: int, size) -> torch.Tensor:
    """pred = disparityregression(maxdisp)(softmax(squeeze(F.upsample(cost, [maxdisp, H, W], mode='trilinear'), 1), 1))
    -- models_psmnet/stackhourglass.py:149-155 -- in one kernel.  `size` = (H, W) of the full-resolution image."""
    H, W = int(size[0]), int(size[1])
    return _UpsampleSoftArgmin.apply(cost_lowres, int(maxdisp), H, W)
