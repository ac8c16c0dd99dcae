"""SyncBatchNorm for a siamese tower that is fed [left; right] in ONE pass.

The reference runs its feature tower twice per step (models/dsnet_t2.py:1159-1160) under nn.SyncBatchNorm
(torch_implementation.py:739): every BN layer is invoked twice, each time with its own batch statistics and its own
collective.  `PairedSyncBatchNorm` keeps those semantics (per-half statistics, running statistics updated for left then
right) but needs one collective and two kernel launches per layer and direction (csrc/bn_pair.cu through the C ABI).
`pair_batchnorms(module)` converts the BatchNorm2d / SyncBatchNorm layers below a module; the caller then feeds
`torch.cat([left, right])` and splits the result with `.chunk(2)`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class _PairedSyncBNFn(torch.autograd.Function):
    """Batch norm of a concatenated [left; right] batch with the statistics of each half kept separate -- what two
    consecutive calls of one nn.SyncBatchNorm on `left` and on `right` compute (the reference runs its siamese tower
    twice, models/dsnet_t2.py:1159-1160) -- but with ONE collective per layer instead of two in the forward
    (all_gather of both halves' mean / invstd / count) and one instead of two in the backward (all_reduce of both
    halves' sum_dy / sum_dy_xmu).  SURVEY.md section 8 f4: the step's scaling is bound by the latency of these tiny
    collectives, so halving their number is worth more than any bandwidth."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, world_size):
        x = x.contiguous()
        half = x.size(0) // 2
        halves = (x[:half], x[half:])
        C = x.size(1)
        local = []
        for xi in halves:
            mean, invstd = torch.batch_norm_stats(xi, eps)
            local += [mean, invstd]
        count = torch.full((1,), halves[0].numel() // C, dtype=local[0].dtype, device=x.device)
        if world_size > 1:
            combined = torch.cat(local + [count])                                    # (4C + 1,)
            gathered = torch.empty(world_size, combined.numel(), dtype=combined.dtype, device=x.device)
            torch.distributed.all_gather_into_tensor(gathered, combined, group=group)
            counts = gathered[:, 4 * C]
        else:
            gathered = torch.cat(local + [count]).unsqueeze(0)
            counts = count
        outs, saved = [], []
        for i, xi in enumerate(halves):                                               # left first, like two calls
            mean_all = gathered[:, 2 * i * C:(2 * i + 1) * C]
            invstd_all = gathered[:, (2 * i + 1) * C:(2 * i + 2) * C]
            mom = momentum[i] if isinstance(momentum, (tuple, list)) else momentum
            mean, invstd = torch.batch_norm_gather_stats_with_counts(xi, mean_all, invstd_all, running_mean,
                                                                     running_var, mom, eps, counts.view(-1))
            outs.append(torch.batch_norm_elemt(xi, weight, bias, mean, invstd, eps))
            saved += [mean, invstd]
        ctx.save_for_backward(x, weight, *saved, counts.to(torch.int32))
        ctx.group, ctx.world_size = group, world_size
        return torch.cat(outs)

    @staticmethod
    def backward(ctx, grad):
        x, weight, mean_l, invstd_l, mean_r, invstd_r, counts = ctx.saved_tensors
        grad = grad.contiguous()
        half = x.size(0) // 2
        C = x.size(1)
        parts = ((x[:half], grad[:half], mean_l, invstd_l), (x[half:], grad[half:], mean_r, invstd_r))
        red, gw, gb = [], None, None
        for xi, gi, mean, invstd in parts:
            sum_dy, sum_dy_xmu, gwi, gbi = torch.batch_norm_backward_reduce(gi, xi, mean, invstd, weight, True, True, True)
            red += [sum_dy, sum_dy_xmu]
            gw = gwi if gw is None else gw + gwi
            gb = gbi if gb is None else gb + gbi
        combined = torch.cat(red)                                                    # (4C,)
        if ctx.world_size > 1:
            torch.distributed.all_reduce(combined, group=ctx.group)
        gins = []
        for i, (xi, gi, mean, invstd) in enumerate(parts):
            sum_dy = combined[2 * i * C:(2 * i + 1) * C]
            sum_dy_xmu = combined[(2 * i + 1) * C:(2 * i + 2) * C]
            gins.append(torch.batch_norm_backward_elemt(gi, xi, mean, invstd, weight, sum_dy, sum_dy_xmu, counts))
        return torch.cat(gins), gw, gb, None, None, None, None, None, None


class _PairedSyncBNFusedFn(torch.autograd.Function):
    """Same operator as _PairedSyncBNFn on this package's own kernels (csrc/bn_pair.cu, include/pmt_ops.h section f4):
    two launches + one collective per direction instead of ~9 ATen launches + one collective.  fp32 CUDA NCHW only."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, world_size, relu=False):
        import ctypes

        from . import _util as U

        x = x.contiguous()
        B2, C = x.size(0), x.size(1)
        B, HW = B2 // 2, x[0, 0].numel()
        dev = x.device
        payload = torch.empty(4 * C + 1, device=dev, dtype=torch.float32)   # [half][c][mean, M2], count
        U.call("pmt_bn_pair_stats_f32", dev, U.ptr(x), U.ptr(payload), B, C, HW)
        if world_size > 1:
            gathered = torch.empty(world_size, 4 * C + 1, device=dev, dtype=torch.float32)
            torch.distributed.all_gather_into_tensor(gathered, payload, group=group)
        else:
            gathered = payload
        out = torch.empty_like(x)
        save_mean = torch.empty(2 * C, device=dev, dtype=torch.float32)
        save_invstd = torch.empty(2 * C + 1, device=dev, dtype=torch.float32)   # [2C] = total count
        lib = U._lib.load()
        with torch.cuda.device(dev):
            st = lib.pmt_bn_pair_apply_f32(U.ptr(x), U.ptr(gathered), int(world_size), U.ptr(weight), U.ptr(bias),
                                           U.ptr(running_mean), U.ptr(running_var), ctypes.c_float(momentum),
                                           ctypes.c_float(eps), U.ptr(out), U.ptr(save_mean), U.ptr(save_invstd), B, C, HW,
                                           int(bool(relu)), U.stream_ptr(dev))
        U._lib.check(st, "pmt_bn_pair_apply_f32")
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
        ctx.group, ctx.world_size, ctx.relu = group, world_size, int(bool(relu))
        return out

    @staticmethod
    def backward(ctx, grad):
        from . import _util as U

        x, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        grad = grad.contiguous()
        B2, C = x.size(0), x.size(1)
        B, HW = B2 // 2, x[0, 0].numel()
        dev = x.device
        sums = torch.empty(4 * C, device=dev, dtype=torch.float32)
        gwb = torch.zeros(2, C, device=dev, dtype=torch.float32)
        U.call("pmt_bn_pair_bwd_reduce_f32", dev, U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd), U.ptr(sums),
               U.ptr(gwb[0]), U.ptr(gwb[1]), B, C, HW, U.ptr(weight), U.ptr(bias), ctx.relu)
        if ctx.world_size > 1:
            torch.distributed.all_reduce(sums, group=ctx.group)
        dx = torch.empty_like(x)
        U.call("pmt_bn_pair_bwd_apply_f32", dev, U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd), U.ptr(weight),
               U.ptr(sums), U.ptr(dx), B, C, HW, U.ptr(bias), ctx.relu)
        gw = gwb[0] if weight is not None else None
        gb = gwb[1] if weight is not None else None
        return dx, gw, gb, None, None, None, None, None, None, None


class PairedSyncBatchNorm(nn.BatchNorm2d):
    """Drop-in for the BatchNorm2d / SyncBatchNorm layers of a siamese tower that is fed [left; right] in one pass (see
    _PairedSyncBNFn).  Single process: equals calling the BatchNorm2d on each half in turn.

    `fused=True` (default) runs this package's kernels and is strict: fp32 CUDA input and a numeric `momentum`, anything
    else raises (no silent fallback).  `fused=False` is an explicit opt-in to the composition of ATen ops (any dtype; it
    is the float64 reference of the tests and supports `momentum=None`, the cumulative moving average).
    `process_group` restricts the statistics exchange to a sub-group like nn.SyncBatchNorm's argument of that name."""

    fused = True
    relu = False   # True: the ReLU that follows this BN is computed by the same kernels (pair_batchnorms sets it)
    process_group = None

    def _world(self):
        dist = torch.distributed
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.process_group)

    def forward(self, x):
        use_batch_stats = self.training or self.running_mean is None   # stock BN: no running stats -> batch stats in eval
        if not use_batch_stats:
            y = F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0, self.eps)
            return F.relu(y) if self.relu else y
        if x.size(0) % 2:
            raise ValueError("PairedSyncBatchNorm expects an even batch: [left; right]")
        if not x.is_cuda:
            raise RuntimeError(f"PairedSyncBatchNorm got a tensor on {x.device}: batch statistics need CUDA tensors (no CPU path)")
        track = self.training and self.running_mean is not None
        ws = self._world()
        if self.fused:
            if x.dtype != torch.float32:
                raise NotImplementedError(f"PairedSyncBatchNorm(fused=True) implements float32 only, got {x.dtype}; set "
                                          ".fused = False for the ATen composition")
            if self.momentum is None and track:
                raise NotImplementedError("PairedSyncBatchNorm(fused=True) needs a numeric momentum; momentum=None "
                                          "(cumulative average) is implemented by .fused = False")
            if track and self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(2)
            if x.data_ptr() % 16:
                x = x.clone(memory_format=torch.contiguous_format)   # an offset view: the kernels read 16-byte vectors
            return _PairedSyncBNFusedFn.apply(x, self.weight, self.bias, self.running_mean if track else None,
                                              self.running_var if track else None, self.eps,
                                              self.momentum if self.momentum is not None else 0.0, self.process_group, ws,
                                              self.relu)
        if self.momentum is None and track:
            # cumulative moving average: the left call sees num_batches_tracked+1, the right call +2 (two calls of one BN)
            n = int(self.num_batches_tracked) if self.num_batches_tracked is not None else 0
            factors = (1.0 / (n + 1), 1.0 / (n + 2))
        else:
            factors = (self.momentum if self.momentum is not None else 0.0,) * 2
        if track and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(2)
        y = _PairedSyncBNFn.apply(x, self.weight, self.bias, self.running_mean if track else None,
                                  self.running_var if track else None, self.eps, factors, self.process_group, ws)
        return F.relu(y) if self.relu else y


# Parents whose registration order is their call order, so "the next registered sibling is an nn.ReLU" really means
# "this ReLU is applied to this BN's output and to nothing else": nn.Sequential and torchvision's DenseNet blocks
# (norm1 -> relu1 -> conv1 -> norm2 -> relu2 -> conv2).  ResNet-style blocks register ONE relu after bn1 and apply it
# again after the residual add -- replacing it would silently drop that second use -- so nothing else is ever fused.
_RELU_FUSION_PARENTS = ("_DenseLayer", "_Transition")


def _may_fuse_relu(parent: nn.Module) -> bool:
    return isinstance(parent, nn.Sequential) or type(parent).__name__ in _RELU_FUSION_PARENTS


def pair_batchnorms(module: nn.Module, fuse_relu: bool = False) -> nn.Module:
    """Replace every BatchNorm2d / SyncBatchNorm below `module` by a PairedSyncBatchNorm that shares its parameters and
    buffers (and keeps a converted SyncBatchNorm's process_group).

    fuse_relu (opt-in): a BN whose NEXT registered sibling is an nn.ReLU takes that ReLU over (the sibling becomes
    nn.Identity) -- only inside nn.Sequential containers and torchvision's _DenseLayer / _Transition, where registration
    order is call order and the ReLU has no other use.  A ReLU that is a shared attribute of a hand-written block
    (ResNet BasicBlock/Bottleneck: `self.relu` after bn1 AND after the residual add) is never touched."""
    names = [n for n, _ in module.named_children()]
    for i, name in enumerate(names):
        child = getattr(module, name)
        if isinstance(child, (nn.BatchNorm2d, nn.SyncBatchNorm)) and not isinstance(child, PairedSyncBatchNorm):
            new = PairedSyncBatchNorm(child.num_features, child.eps, child.momentum, child.affine, child.track_running_stats)
            new.weight, new.bias = child.weight, child.bias
            new.training = child.training
            new.running_mean, new.running_var, new.num_batches_tracked = (child.running_mean, child.running_var,
                                                                          child.num_batches_tracked)
            new.process_group = getattr(child, "process_group", None)
            if (fuse_relu and _may_fuse_relu(module) and i + 1 < len(names)
                    and isinstance(getattr(module, names[i + 1]), nn.ReLU)):
                new.relu = True
                setattr(module, names[i + 1], nn.Identity())
            setattr(module, name, new)
        else:
            pair_batchnorms(child, fuse_relu)
    return module
