"""Data-parallel training-step harness around the hot path (BASELINE configs 2 and 5).

The reference's models (models/dsnet_t2.py::minidsnetExt etc.) cannot travel to the GPU box and are out of scope as
build targets; what matters for row (e) is the SHAPE of the step they run: a siamese DenseNet-121 tower on
left/right 256x512 images, the 1x17 correlation of the two 352-channel 1/8-resolution feature maps
(models/dsnet_t2.py:1159-1160,1188), a 1x1 `corrConv2d` 17->128 + ReLU (:1197), decoders for disparity and
segmentation, the disparity-guided warp of the right branch (models/dsnet_t2_warp.py:697-698), CE + L1 losses and
Adam(lr=1.5e-3, eps=1e-7) (torch_implementation.py:279-305,724), under DistributedDataParallel with
nn.SyncBatchNorm (torch_implementation.py:739-741).  `SDNetLite` reproduces that call pattern with this package's
ops on the hot path; everything else is stock torch/cuDNN (out of scope).  The only collectives are DDP's gradient
all-reduce and SyncBatchNorm's statistics exchange -- the hot-path ops never communicate.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .correlation import SpatialCorrelationSampler
from .warp import apply_disparity


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
            mean, invstd = torch.batch_norm_gather_stats_with_counts(xi, mean_all, invstd_all, running_mean,
                                                                     running_var, momentum, eps, counts.view(-1))
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
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, group, world_size):
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
                                           U.stream_ptr(dev))
        U._lib.check(st, "pmt_bn_pair_apply_f32")
        ctx.save_for_backward(x, weight, save_mean, save_invstd)
        ctx.group, ctx.world_size = group, world_size
        return out

    @staticmethod
    def backward(ctx, grad):
        from . import _util as U

        x, weight, save_mean, save_invstd = ctx.saved_tensors
        grad = grad.contiguous()
        B2, C = x.size(0), x.size(1)
        B, HW = B2 // 2, x[0, 0].numel()
        dev = x.device
        sums = torch.empty(4 * C, device=dev, dtype=torch.float32)
        gwb = torch.zeros(2, C, device=dev, dtype=torch.float32)
        U.call("pmt_bn_pair_bwd_reduce_f32", dev, U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd), U.ptr(sums),
               U.ptr(gwb[0]), U.ptr(gwb[1]), B, C, HW)
        if ctx.world_size > 1:
            torch.distributed.all_reduce(sums, group=ctx.group)
        dx = torch.empty_like(x)
        U.call("pmt_bn_pair_bwd_apply_f32", dev, U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd), U.ptr(weight),
               U.ptr(sums), U.ptr(dx), B, C, HW)
        gw = gwb[0] if weight is not None else None
        gb = gwb[1] if weight is not None else None
        return dx, gw, gb, None, None, None, None, None, None


class PairedSyncBatchNorm(nn.BatchNorm2d):
    """Drop-in for the BatchNorm2d layers of a siamese tower that is fed [left; right] in one pass (see
    _PairedSyncBNFn).  Single process: equals calling the BatchNorm2d on each half in turn.  `fused` selects this
    package's kernels (fp32 CUDA) over the composition of ATen ops (any dtype; the float64 reference of the tests)."""

    fused = True

    def forward(self, x):
        if not self.training:
            return F.batch_norm(x, self.running_mean, self.running_var, self.weight, self.bias, False, 0.0, self.eps)
        if x.size(0) % 2:
            raise ValueError("PairedSyncBatchNorm expects an even batch: [left; right]")
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(2)
        dist = torch.distributed
        ws = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        fused = (self.fused and x.is_cuda and x.dtype == torch.float32 and self.momentum is not None
                 and x.data_ptr() % 16 == 0)
        fn = _PairedSyncBNFusedFn if fused else _PairedSyncBNFn
        return fn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum, None, ws)


def pair_batchnorms(module: nn.Module) -> nn.Module:
    """Replace every BatchNorm2d below `module` by a PairedSyncBatchNorm that shares its parameters and buffers."""
    for name, child in module.named_children():
        if isinstance(child, (nn.BatchNorm2d, nn.SyncBatchNorm)) and not isinstance(child, PairedSyncBatchNorm):
            new = PairedSyncBatchNorm(child.num_features, child.eps, child.momentum, child.affine, child.track_running_stats)
            new.weight, new.bias = child.weight, child.bias
            new.running_mean, new.running_var, new.num_batches_tracked = (child.running_mean, child.running_var,
                                                                          child.num_batches_tracked)
            setattr(module, name, new)
        else:
            pair_batchnorms(child)
    return module


def _cbr(cin, cout, k=3, s=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class SDNetLite(nn.Module):
    """Joint segmentation + disparity net with the reference's hot-path call pattern (1dcorr, max_disp 8 at 1/8)."""

    def __init__(self, n_labels: int = 2, feat_ch: int = 352, max_disp: int = 8, backbone: str = "densenet121",
                 paired_tower: bool = False):
        super().__init__()
        self.paired_tower = paired_tower
        if backbone == "densenet121":
            import torchvision

            f = torchvision.models.densenet121(weights=None).features
            self.tower = nn.Sequential(f.conv0, f.norm0, f.relu0, f.pool0, f.denseblock1, f.transition1, f.denseblock2)
            tower_ch = 512
        else:  # small tower for CPU/unit tests
            self.tower = nn.Sequential(_cbr(3, 32, 3, 2), _cbr(32, 64, 3, 2), _cbr(64, 96, 3, 2))
            tower_ch = 96
        self.reduce = _cbr(tower_ch, feat_ch, 1)
        if paired_tower:
            self.pair_tower()
        self.patch = 2 * max_disp + 1                                   # patch_corr = (1, max_disp*2+1)
        self.correlation_sampler = SpatialCorrelationSampler(kernel_size=1, patch_size=(1, self.patch), stride=1,
                                                             padding=0, dilation_patch=1)
        self.corrConv2d = nn.Sequential(nn.Conv2d(self.patch, 128, 1), nn.ReLU(inplace=True))
        self.skip = _cbr(feat_ch, 64, 1)
        self.dec = nn.Sequential(_cbr(128 + 64, 128), _cbr(128, 64))
        self.disp_head = nn.Conv2d(64, 1, 3, 1, 1)
        self.seg_l = nn.Sequential(_cbr(feat_ch, 64), nn.Conv2d(64, n_labels, 3, 1, 1))
        self.seg_r = nn.Sequential(_cbr(feat_ch, 64), nn.Conv2d(64, n_labels, 3, 1, 1))
        self.att = nn.Sequential(nn.Conv2d(1, 1, 3, 1, 1), nn.Sigmoid())

    def pair_tower(self):
        """Run the siamese tower once over [left; right] with per-half BN statistics and one collective per BN layer
        (call again after nn.SyncBatchNorm.convert_sync_batchnorm, which replaces every _BatchNorm it finds)."""
        self.paired_tower = True
        pair_batchnorms(self.tower)
        pair_batchnorms(self.reduce)
        return self

    def forward(self, left, right):
        H, W = left.shape[-2:]
        if self.paired_tower:
            a, b = self.reduce(self.tower(torch.cat([left, right]))).chunk(2)
        else:
            a = self.reduce(self.tower(left))
            b = self.reduce(self.tower(right))
        y = self.correlation_sampler(a, b)                              # (B,1,17,h,w)
        y = torch.squeeze(y, dim=1)                                     # 1dcorr: not divided by C
        y = self.corrConv2d(y)
        x = self.dec(torch.cat([y, self.skip(a)], 1))
        disp = F.interpolate(self.disp_head(x), size=(H, W), mode="bilinear", align_corners=False) * 8.0
        seg_left = F.interpolate(self.seg_l(a), size=(H, W), mode="bilinear", align_corners=False)
        seg_right = F.interpolate(self.seg_r(b), size=(H, W), mode="bilinear", align_corners=False)
        warped = apply_disparity(seg_right, -disp)                      # sample the right branch at w - d
        at_d = self.att(disp)
        seg_both = (1 - at_d) * seg_left + at_d * warped
        return seg_left, disp, seg_both, disp


def sdnet_loss(outputs, seg_target, disp_target):
    seg1, disp1, seg2, _ = outputs
    return (F.cross_entropy(seg1, seg_target) + F.cross_entropy(seg2, seg_target)
            + F.l1_loss(disp1.squeeze(1), disp_target))


def synthetic_batch(batch: int, h: int, w: int, n_labels: int, device, generator=None):
    left = torch.rand(batch, 3, h, w, device=device, generator=generator)
    right = torch.rand(batch, 3, h, w, device=device, generator=generator)
    disp = 64.0 * torch.rand(batch, h, w, device=device, generator=generator)
    seg = torch.randint(0, n_labels, (batch, h, w), device=device, generator=generator)
    return left, right, seg, disp


def build_training_step(world, batch_per_gpu: int = 4, h: int = 256, w: int = 512, n_labels: int = 2,
                        backbone: str = "densenet121", sync_bn: bool = True, cuda_graph: bool = False,
                        paired_tower: bool = False):
    """Returns (step_fn, model): step_fn() runs one fwd + loss + bwd + Adam step on a fixed synthetic batch.

    cuda_graph=True captures the whole step -- forward, our hot-path kernels, backward, DDP's bucketed gradient
    all-reduce, SyncBatchNorm's statistics collectives and the Adam update -- into ONE CUDA graph.  The eager step
    is launch-bound (hundreds of tiny BN/NCCL launches per step, section 5.8 of SURVEY.md); replaying a graph
    removes the host from the loop, which is what lets the step scale across GPUs."""
    dev = torch.device("cuda", world.local_rank)
    torch.manual_seed(1234)  # identical initial weights on every rank
    model = SDNetLite(n_labels=n_labels, backbone=backbone, paired_tower=paired_tower and sync_bn).to(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):  # DDP must be built and warmed up on the stream family the graph is captured from
        if world.distributed:
            if sync_bn:
                model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
                if paired_tower:
                    model.pair_tower()
            model = nn.parallel.DistributedDataParallel(model, device_ids=[world.local_rank])
        opt = torch.optim.Adam(model.parameters(), lr=1.5e-3, eps=1e-7, capturable=cuda_graph)
        g = torch.Generator(device=dev).manual_seed(world.rank)
        left, right, seg, disp = synthetic_batch(batch_per_gpu, h, w, n_labels, dev, g)

        def eager_step():
            opt.zero_grad(set_to_none=True)
            loss = sdnet_loss(model(left, right), seg, disp)
            loss.backward()
            opt.step()
            return loss

        if cuda_graph:
            for _ in range(11):  # DDP needs >= 11 eager iterations before its collectives can be captured
                eager_step()
    torch.cuda.current_stream(dev).wait_stream(side)
    if not cuda_graph:
        return eager_step, model

    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    # thread_local: the NCCL watchdog thread may touch the CUDA runtime while this thread captures
    with torch.cuda.graph(graph, capture_error_mode="thread_local" if world.distributed else "global"):
        static_loss = sdnet_loss(model(left, right), seg, disp)
        static_loss.backward()
        opt.step()

    def graph_step():
        graph.replay()
        return static_loss

    graph_step.graph = graph
    return graph_step, model
