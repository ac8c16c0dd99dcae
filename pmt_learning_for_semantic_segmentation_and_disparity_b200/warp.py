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
        gimg = torch.empty_like(img) if need_img else None   # fully written by the kernel (deterministic row gather)
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


class _WarpBlend(Function):
    """(1 - att) * seg_left + att * apply_disparity(seg_right, off) in one kernel per direction."""

    @staticmethod
    def forward(ctx, seg_left, seg_right, off, att):
        seg_left = U.require_cuda_f32(seg_left, "seg_left")
        seg_right = U.require_cuda_f32(seg_right, "seg_right")
        off = U.require_cuda_f32(off, "x_offset")
        att = U.require_cuda_f32(att, "att")
        dev = U.same_device(seg_left, seg_right, off, att)
        N, C, H, W = seg_right.shape
        if seg_left.shape != seg_right.shape or off.numel() != N * H * W or att.numel() != N * H * W:
            raise ValueError(f"warp_blend: shapes {tuple(seg_left.shape)}, {tuple(seg_right.shape)}, {tuple(off.shape)}, "
                             f"{tuple(att.shape)} do not match (N,C,H,W)/(N,C,H,W)/(N,1,H,W)/(N,1,H,W)")
        if any(ctx.needs_input_grad) and U._lib.load().pmt_warp1d_rows_supported(N, H, W) != 1:
            raise NotImplementedError("warp_blend backward needs N*H*W < 2**24 and W <= 1024; use apply_disparity + the "
                                      "blend expression for this shape (no silent fallback)")
        out = torch.empty_like(seg_right)
        warped = torch.empty_like(seg_right)
        U.call("pmt_warp1d_blend_fwd_f32", dev, U.ptr(seg_right), U.ptr(off), U.ptr(att), U.ptr(seg_left), U.ptr(out),
               U.ptr(warped), N, C, H, W)
        ctx.save_for_backward(seg_left, seg_right, off, att)
        return out, warped

    @staticmethod
    @once_differentiable
    def backward(ctx, gout, gwarped):
        seg_left, seg_right, off, att = ctx.saved_tensors
        N, C, H, W = seg_right.shape
        gout = U.require_cuda_f32(gout, "grad_output")
        gw = U.require_cuda_f32(gwarped, "grad_warped") if gwarped is not None else None
        gimg, gseg = torch.empty_like(seg_right), torch.empty_like(seg_left)
        goff, gatt = torch.empty_like(off), torch.empty_like(att)
        U.call("pmt_warp1d_blend_bwd_f32", seg_right.device, U.ptr(seg_right), U.ptr(off), U.ptr(att), U.ptr(seg_left),
               U.ptr(gout), U.ptr(gw), U.ptr(gimg), U.ptr(goff), U.ptr(gatt), U.ptr(gseg), N, C, H, W)
        return gseg, gimg, goff, gatt


def warp_blend(seg_left, seg_right, x_offset, att):
    """Fused form of models/dsnet_t2_warp.py:697-698::

        seg_right_w = apply_disparity(seg_right, x_offset)          # callers pass x_offset = -disp_out
        seg_both    = (1 - att) * seg_left + att * seg_right_w

    Returns (seg_both, seg_right_w), both dense (N,C,H,W); values are bit-identical to the reference expression.  att and
    x_offset are (N,1,H,W).  Differentiable in all four arguments (deterministic backward)."""
    N, C, H, W = seg_right.shape
    return _WarpBlend.apply(seg_left, seg_right, x_offset.contiguous().view(N, 1, H, W), att.contiguous().view(N, 1, H, W))


class _WarpMSE(Function):
    @staticmethod
    def forward(ctx, right, off, left, mask_positive):
        right = U.require_cuda_f32(right, "right")
        off = U.require_cuda_f32(off, "x_offset")
        left = U.require_cuda_f32(left, "left")
        dev = U.same_device(right, off, left)
        N, C, H, W = right.shape
        if left.shape != right.shape or off.numel() != N * H * W:
            raise ValueError("photo_consistency_mse: right/left must be (N,C,H,W) and x_offset (N,1,H,W)")
        if right.numel() == 0:
            raise ValueError("photo_consistency_mse of an empty tensor is undefined")
        lib = U._lib.load()
        if any(ctx.needs_input_grad) and lib.pmt_warp1d_rows_supported(N, H, W) != 1:
            raise NotImplementedError("photo_consistency_mse backward needs N*H*W < 2**24 and W <= 1024 (no silent fallback)")
        work = torch.empty(lib.pmt_warp1d_mse_workspace(), device=dev, dtype=torch.float64)
        loss = torch.empty((), device=dev, dtype=torch.float32)
        U.call("pmt_warp1d_mse_fwd_f32", dev, U.ptr(right), U.ptr(off), U.ptr(left), int(bool(mask_positive)), U.ptr(work),
               U.ptr(loss), N, C, H, W)
        ctx.save_for_backward(right, off, left)
        ctx.mask = int(bool(mask_positive))
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, gloss):
        right, off, left = ctx.saved_tensors
        N, C, H, W = right.shape
        need_r, need_o, need_l, _ = ctx.needs_input_grad
        g = U.require_cuda_f32(gloss, "grad_loss")
        gimg = torch.empty_like(right) if need_r else None
        goff = torch.empty_like(off) if need_o else None
        gleft = torch.empty_like(left) if need_l else None
        if need_r or need_o or need_l:
            U.call("pmt_warp1d_mse_bwd_f32", right.device, U.ptr(right), U.ptr(off), U.ptr(left), ctx.mask, U.ptr(g),
                   U.ptr(gimg), U.ptr(goff), U.ptr(gleft), N, C, H, W)
        return gimg, goff, gleft, None


def photo_consistency_mse(right, x_offset, left, mask_positive_disparity=False):
    """Fused form of torch_implementation.py:314-317::

        warped_right = apply_disparity(right, x_offset)             # x_offset = -disp
        [warped_right = warped_right * (disp > 0)]                  # mask_positive_disparity, dsnet_t2_warp.py:811
        loss = nn.MSELoss()(warped_right, left)

    One kernel + a fixed-order reduction: bit-reproducible scalar; differentiable in right, x_offset and left."""
    N, C, H, W = right.shape
    return _WarpMSE.apply(right, x_offset.contiguous().view(N, 1, H, W), left, mask_positive_disparity)
