"""Drop-in for the third-party ``spatial_correlation_sampler`` package on B200.

Mirrors the surface the reference constructs (models/dsnet_t2.py:129-133,425,847-851,1078-1087;
models/dsnet_t2_warp.py:197,506,615,742,877; models/torch_dsnet.py:133-138;
models_deeplab_mod/net.py:99-103): ``SpatialCorrelationSampler(kernel_size, patch_size, stride, padding,
dilation, dilation_patch)(input1, input2) -> (B, pH, pW, H, W)``, the functional
``spatial_correlation_sample`` and the autograd ``SpatialCorrelationSamplerFunction`` whose backward
returns ``(grad_input1, grad_input2, None x 6)``.

Every call site of the reference uses kernel_size=1, stride=1, padding=0, dilation=1; that is what the
CUDA kernels implement (any patch_size, any dilation_patch).  Other values raise NotImplementedError --
there is no silent fallback.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _util as U

# Engine used for the 1 x P horizontal patch (the hot path):
#   "auto" : fp32-accurate, fastest engine that fits (tensor cores with the 3xTF32 split -> CUDA-core
#            tiled -> generic); this is the drop-in default.
#   "simt" : force the CUDA-core (fp32 FFMA) kernels.
#   "tf32" : plain-TF32 tensor-core variant (~1e-3 relative, the reduced-precision option that plays the
#            role of north_star's bf16 variant); raises if the shape does not fit -- never a silent fallback.
_ENGINES = ("auto", "simt", "tf32")
_engine = "auto"


def set_correlation_engine(name: str) -> str:
    """Select the engine for 1 x P correlations; returns the previous setting."""
    global _engine
    if name not in _ENGINES:
        raise ValueError(f"engine must be one of {_ENGINES}, got {name!r}")
    prev, _engine = _engine, name
    return prev


def get_correlation_engine() -> str:
    return _engine


def _check_supported(kernel_size, stride, padding, dilation):
    kH, kW = U.pair(kernel_size, "kernel_size")
    dH, dW = U.pair(stride, "stride")
    padH, padW = U.pair(padding, "padding")
    dilH, dilW = U.pair(dilation, "dilation")
    if (kH, kW) != (1, 1) or (dH, dW) != (1, 1) or (padH, padW) != (0, 0):
        raise NotImplementedError(
            "B200 correlation implements kernel_size=1, stride=1, padding=0 (every reference call site); got "
            f"kernel_size={kernel_size}, stride={stride}, padding={padding}")
    del dilH, dilW  # dilation is irrelevant for a 1x1 kernel


class SpatialCorrelationSamplerFunction(Function):
    @staticmethod
    def forward(ctx, input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                dilation_patch=1):
        _check_supported(kernel_size, stride, padding, dilation)
        pH, pW = U.pair(patch_size, "patch_size")
        dpH, dpW = U.pair(dilation_patch, "dilation_patch")
        if pH < 1 or pW < 1 or dpH < 1 or dpW < 1:
            raise ValueError("patch_size and dilation_patch must be >= 1")
        if (pH % 2 == 0 and dpH > 1) or (pW % 2 == 0 and dpW > 1):
            raise NotImplementedError(
                "an even patch_size combined with dilation_patch > 1 is not implemented: upstream's CPU and CUDA builds "
                "centre that window differently and no reference call site uses it")
        in1 = U.require_cuda_f32(input1, "input1")
        in2 = U.require_cuda_f32(input2, "input2")
        if in1.dim() != 4 or in1.shape != in2.shape:
            raise ValueError(f"input1/input2 must be 4-D (B,C,H,W) of equal shape, got {tuple(input1.shape)} and "
                             f"{tuple(input2.shape)}")
        dev = U.same_device(in1, in2)
        B, C, H, W = in1.shape
        ctx.save_for_backward(in1, in2)
        ctx.patch = (pH, pW, dpH, dpW)
        engine = _engine if pH == 1 else "auto"
        ctx.engine = engine
        out = torch.empty((B, pH, pW, H, W), device=dev, dtype=torch.float32)
        if engine == "simt":
            U.call("pmt_corr1d_fwd_simt_f32", dev, U.ptr(in1), U.ptr(in2), U.ptr(out), B, C, H, W, pW, dpW)
        elif engine == "tf32":
            U.call("pmt_corr1d_fwd_tc_f32", dev, U.ptr(in1), U.ptr(in2), U.ptr(out), B, C, H, W, pW, dpW, 1)
        else:
            U.call("pmt_corr_fwd_f32", dev, U.ptr(in1), U.ptr(in2), U.ptr(out), B, C, H, W, pH, pW, dpH, dpW)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        in1, in2 = ctx.saved_tensors
        pH, pW, dpH, dpW = ctx.patch
        B, C, H, W = in1.shape
        g = U.require_cuda_f32(grad_output, "grad_output")
        g1 = torch.empty_like(in1)
        g2 = torch.empty_like(in2)
        args = (U.ptr(in1), U.ptr(in2), U.ptr(g), U.ptr(g1), U.ptr(g2), B, C, H, W)
        if ctx.engine == "simt":
            U.call("pmt_corr1d_bwd_simt_f32", in1.device, *args, pW, dpW)
        elif ctx.engine == "tf32":
            U.call("pmt_corr1d_bwd_tc_f32", in1.device, *args, pW, dpW, 1)
        else:
            U.call("pmt_corr_bwd_f32", in1.device, *args, pH, pW, dpH, dpW)
        return g1, g2, None, None, None, None, None, None


def spatial_correlation_sample(input1, input2, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1,
                               dilation_patch=1):
    """Functional form; returns (B, pH, pW, H, W), un-normalised (callers divide by C themselves for the
    2-D patch, models/dsnet_t2.py:223,884)."""
    return SpatialCorrelationSamplerFunction.apply(input1, input2, kernel_size, patch_size, stride, padding,
                                                   dilation, dilation_patch)


class SpatialCorrelationSampler(nn.Module):
    """Parameter-free module, same constructor keywords as the upstream class."""

    def __init__(self, kernel_size=1, patch_size=1, stride=1, padding=0, dilation=1, dilation_patch=1):
        super().__init__()
        self.kernel_size = kernel_size
        self.patch_size = patch_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.dilation_patch = dilation_patch

    def forward(self, input1, input2):
        return SpatialCorrelationSamplerFunction.apply(input1, input2, self.kernel_size, self.patch_size,
                                                       self.stride, self.padding, self.dilation,
                                                       self.dilation_patch)

    def extra_repr(self):
        return (f"kernel_size={self.kernel_size}, patch_size={self.patch_size}, stride={self.stride}, "
                f"padding={self.padding}, dilation={self.dilation}, dilation_patch={self.dilation_patch}")


class _CorrConvReLU(Function):
    @staticmethod
    def forward(ctx, input1, input2, weight, patch):
        in1 = U.require_cuda_f32(input1, "input1")
        in2 = U.require_cuda_f32(input2, "input2")
        wt = U.require_cuda_f32(weight, "weight")
        if in1.dim() != 4 or in1.shape != in2.shape:
            raise ValueError(f"input1/input2 must be 4-D (B,C,H,W) of equal shape, got {tuple(input1.shape)} and "
                             f"{tuple(input2.shape)}")
        B, C, H, W = in1.shape
        O = wt.size(0)
        if wt.numel() != O * patch:
            raise ValueError(f"weight must be (O, {patch}) or (O, {patch}, 1, 1), got {tuple(weight.shape)}")
        dev = U.same_device(in1, in2, wt)
        if U._lib.load().pmt_corr1d_conv_relu_supported(C, H, W, patch, O) != 1:
            raise NotImplementedError(
                f"correlation_conv1x1_relu covers patch (1,17), 16 <= W <= 128, W % 4 == 0, O <= 256; got patch (1,{patch}), "
                f"W={W}, O={O} -- use SpatialCorrelationSampler + Conv2d + ReLU for this shape (no silent fallback)")
        z = torch.empty((B, O, H, W), device=dev, dtype=torch.float32)
        corr = torch.empty((B, patch, H, W), device=dev, dtype=torch.float32)
        U.call("pmt_corr1d_conv_relu_fwd_f32", dev, U.ptr(in1), U.ptr(in2), U.ptr(wt), U.ptr(z), U.ptr(corr), B, C, H, W,
               patch, O)
        ctx.save_for_backward(in1, in2, wt, z, corr)
        ctx.wshape = weight.shape
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, gz):
        in1, in2, wt, z, corr = ctx.saved_tensors
        B, C, H, W = in1.shape
        O, P = wt.size(0), corr.size(1)
        g = U.require_cuda_f32(gz, "grad_output")
        g1, g2 = torch.empty_like(in1), torch.empty_like(in2)
        gw = torch.empty((O, P), device=in1.device, dtype=torch.float32)
        work = torch.empty((B * H, O * P), device=in1.device, dtype=torch.float32)
        U.call("pmt_corr1d_conv_relu_bwd_f32", in1.device, U.ptr(in1), U.ptr(in2), U.ptr(wt), U.ptr(z), U.ptr(corr), U.ptr(g),
               U.ptr(g1), U.ptr(g2), U.ptr(gw), U.ptr(work), B, C, H, W, P, O)
        return g1, g2, gw.view(ctx.wshape), None


def correlation_conv1x1_relu(input1, input2, weight, patch_size=(1, 17)):
    """Fused form of models/dsnet_t2.py:1187-1197 (and :879-888, models/dsnet_t2_warp.py:664-671) for `-corrType 1dcorr`::

        y = torch.squeeze(correlation_sampler(a, b), dim=1)        # (B, P, H, W), not divided by C
        y = corrConv2d(y)                                          # conv2dSame(P, O, 1) without bias, then ReLU

    `weight` is the 1x1 convolution's weight, (O, P, 1, 1) or (O, P).  Returns (B, O, H, W).  The (B,P,H,W) slab stays in
    shared memory; differentiable in input1, input2 and weight (deterministic)."""
    pH, pW = U.pair(patch_size, "patch_size")
    if pH != 1:
        raise NotImplementedError("correlation_conv1x1_relu implements the 1 x P horizontal patch (`-corrType 1dcorr`)")
    return _CorrConvReLU.apply(input1, input2, weight, pW)


class CorrelationConvReLU(nn.Module):
    """`correlation_sampler` + `corrConv2d` of the reference's 1dcorr models as one module.  Build it from the two
    reference sub-modules with `CorrelationConvReLU.from_reference(model.correlation_sampler, model.corrConv2d)`: the
    convolution's weight Parameter is shared, not copied, so optimisers and checkpoints see the same tensor."""

    def __init__(self, patch_size=(1, 17), out_channels=128):
        super().__init__()
        self.patch_size = U.pair(patch_size, "patch_size")
        self.weight = nn.Parameter(torch.empty(out_channels, self.patch_size[1], 1, 1))
        # conv2dSame's init (models/torch_model.py:261-264): N(0, sqrt(2 / (k*k*out_channels)))
        nn.init.normal_(self.weight, 0.0, (2.0 / out_channels) ** 0.5)

    @classmethod
    def from_reference(cls, correlation_sampler, corr_conv2d):
        conv = None
        for m in corr_conv2d.modules():
            if isinstance(m, nn.Conv2d):
                conv = m
                break
        if conv is None or conv.kernel_size != (1, 1) or conv.bias is not None:
            raise ValueError("corrConv2d must hold a bias-free 1x1 nn.Conv2d followed by ReLU (conv2dSame(P, O, 1) + ReLU)")
        self = cls(correlation_sampler.patch_size, conv.out_channels)
        self.weight = conv.weight
        return self

    def forward(self, input1, input2):
        return correlation_conv1x1_relu(input1, input2, self.weight, self.patch_size)
