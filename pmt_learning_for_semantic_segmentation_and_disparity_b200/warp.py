"""apply_disparity on B200 -- drop-in for models/torch_dsnet.py:10-86.

Same signature and the same quirks: output is 0 where x >= W-1 (both lerp weights
vanish), img[...,0] where x <= 0, gather indices formed in float32, and the result is returned as a permuted
view of a [C,N,H,W] buffer exactly like the reference (strides (HW, NHW, W, 1)).  `wrap_mode='border'` pads by
one pixel like the reference; any other wrap_mode returns None like the reference (torch_dsnet.py:21-22).
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.functional import pad

from . import _util as U


class _Warp1D(Function):
    @staticmethod
    def forward(ctx, img, off):
        img = U.require_cuda_f32(img, "input_images")
        off = U.require_cuda_f32(off, "x_offset")
        dev = U.same_device(img, off)
        N, C, H, W = img.shape
        ctx.save_for_backward(img, off)
        buf = torch.empty((C, N, H, W), device=dev, dtype=torch.float32)
        U.call("pmt_warp1d_fwd_f32", dev, U.ptr(img), U.ptr(off), U.ptr(buf), N, C, H, W, 1)
        return buf.permute(1, 0, 2, 3)

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        img, off = ctx.saved_tensors
        N, C, H, W = img.shape
        if not gout.is_cuda or gout.dtype != torch.float32:
            raise RuntimeError("grad_output must be a CUDA float32 tensor")
        if gout.permute(1, 0, 2, 3).is_contiguous():
            g, cnhw = gout, 1          # the layout the forward produced
        else:
            g, cnhw = gout.contiguous(), 0
        need_img, need_off = ctx.needs_input_grad
        gimg = torch.zeros_like(img) if need_img else None   # scatter target
        goff = torch.empty_like(off) if need_off else None
        if need_img or need_off:
            U.call("pmt_warp1d_bwd_f32", img.device, U.ptr(img), U.ptr(off), U.ptr(g), U.ptr(gimg), U.ptr(goff),
                   N, C, H, W, cnhw)
        return gimg, goff


def apply_disparity(input_images, x_offset, wrap_mode='edge', tensor_type='torch.cuda.FloatTensor'):
    """out[n,c,h,w] = lerp of input_images[n,c,h,:] at x = clamp(w + x_offset[n,0,h,w], 0, W-1).

    `tensor_type` is accepted for signature compatibility; tensors stay on the inputs' CUDA device."""
    del tensor_type
    if wrap_mode == 'border':
        edge = 1
        input_images = pad(input_images, (1, 1, 1, 1))
    elif wrap_mode == 'edge':
        edge = 0
    else:
        return None
    if input_images.dim() != 4:
        raise ValueError(f"input_images must be (N,C,H,W), got {tuple(input_images.shape)}")
    N, C, Hp, Wp = input_images.shape
    H, W = Hp - 2 * edge, Wp - 2 * edge
    if x_offset.numel() != N * H * W:
        raise RuntimeError(f"x_offset has {x_offset.numel()} elements, expected N*H*W = {N * H * W}")
    off = x_offset.contiguous().view(N, 1, H, W)
    if edge:
        off = pad(off, (1, 1, 1, 1))
    out = _Warp1D.apply(input_images, off)
    if edge:
        out = out[:, :, 1:-1, 1:-1]
    return out
