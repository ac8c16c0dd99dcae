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

from .correlation import CorrelationConvReLU, SpatialCorrelationSampler
from .syncbn import (PairedSyncBatchNorm, PeerExchange, pair_batchnorms,  # noqa: F401  (re-exported: the harness API)
                     sync_batchnorms_to_peer)
from .warp import apply_disparity, warp_blend


def _cbr(cin, cout, k=3, s=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class SDNetLite(nn.Module):
    """Joint segmentation + disparity net with the reference's hot-path call pattern (1dcorr, max_disp 8 at 1/8)."""

    def __init__(self, n_labels: int = 2, feat_ch: int = 352, max_disp: int = 8, backbone: str = "densenet121",
                 paired_tower: bool = False, full_depth: bool = False, fused_ops: bool = False):
        """full_depth: run the WHOLE DenseNet-121 feature extractor per image like the reference's `densenet` backbone
        (models/dsnet_t2.py:953-964: segnet_input = 1024*2): 121 BatchNorm layers per tower pass; the 1/32 features feed
        the segmentation heads.  Default (False) stops after denseblock2 (39 BN layers), the part the correlation needs.
        fused_ops: correlation + corrConv2d + ReLU as one kernel (f2) and warp + attention blend as one kernel (f4)."""
        super().__init__()
        self.paired_tower = paired_tower
        self.full_depth = full_depth and backbone == "densenet121"
        self.fused_ops = fused_ops
        if backbone == "densenet121":
            import torchvision

            f = torchvision.models.densenet121(weights=None).features
            self.tower = nn.Sequential(f.conv0, f.norm0, f.relu0, f.pool0, f.denseblock1, f.transition1, f.denseblock2)
            tower_ch = 512
            if self.full_depth:
                self.tower_deep = nn.Sequential(f.transition2, f.denseblock3, f.transition3, f.denseblock4, f.norm5,
                                                nn.ReLU(inplace=True))
                self.deep_seg = _cbr(1024, feat_ch, 1)
        else:  # small tower for CPU/unit tests
            self.tower = nn.Sequential(_cbr(3, 32, 3, 2), _cbr(32, 64, 3, 2), _cbr(64, 96, 3, 2))
            tower_ch = 96
        self.reduce = _cbr(tower_ch, feat_ch, 1)
        if paired_tower:
            self.pair_tower()
        self.patch = 2 * max_disp + 1                                   # patch_corr = (1, max_disp*2+1)
        self.correlation_sampler = SpatialCorrelationSampler(kernel_size=1, patch_size=(1, self.patch), stride=1,
                                                             padding=0, dilation_patch=1)
        self.corrConv2d = nn.Sequential(nn.Conv2d(self.patch, 128, 1, bias=False), nn.ReLU(inplace=True))   # conv2dSame: bias=False
        self.skip = _cbr(feat_ch, 64, 1)
        self.dec = nn.Sequential(_cbr(128 + 64, 128), _cbr(128, 64))
        self.disp_head = nn.Conv2d(64, 1, 3, 1, 1)
        self.seg_l = nn.Sequential(_cbr(feat_ch, 64), nn.Conv2d(64, n_labels, 3, 1, 1))
        self.seg_r = nn.Sequential(_cbr(feat_ch, 64), nn.Conv2d(64, n_labels, 3, 1, 1))
        self.att = nn.Sequential(nn.Conv2d(1, 1, 3, 1, 1), nn.Sigmoid())

    def pair_tower(self):
        """Run the siamese tower once over [left; right] with per-half BN statistics and one collective per BN layer.
        Call it AFTER nn.SyncBatchNorm.convert_sync_batchnorm (which replaces every _BatchNorm it finds, paired ones
        included, and would drop the ReLUs they have taken over), and only once."""
        self.paired_tower = True
        pair_batchnorms(self.tower, fuse_relu=True)   # torchvision _DenseLayer/_Transition + Sequential: safe to fuse
        pair_batchnorms(self.reduce, fuse_relu=True)
        if self.full_depth:
            pair_batchnorms(self.tower_deep, fuse_relu=True)
            pair_batchnorms(self.deep_seg, fuse_relu=True)
        return self

    def forward(self, left, right):
        H, W = left.shape[-2:]
        if self.paired_tower:
            mid = self.tower(torch.cat([left, right]))
            a, b = self.reduce(mid).chunk(2)
            if self.full_depth:
                da, db = self.deep_seg(self.tower_deep(mid)).chunk(2)
        else:
            ma, mb = self.tower(left), self.tower(right)
            a, b = self.reduce(ma), self.reduce(mb)
            if self.full_depth:
                da, db = self.deep_seg(self.tower_deep(ma)), self.deep_seg(self.tower_deep(mb))
        if self.fused_ops:
            if not hasattr(self, "_corr_conv"):                         # shares corrConv2d's weight Parameter
                self._corr_conv = CorrelationConvReLU.from_reference(self.correlation_sampler, self.corrConv2d)
            y = self._corr_conv(a, b)
        else:
            y = self.correlation_sampler(a, b)                          # (B,1,17,h,w)
            y = torch.squeeze(y, dim=1)                                 # 1dcorr: not divided by C
            y = self.corrConv2d(y)
        x = self.dec(torch.cat([y, self.skip(a)], 1))
        disp = F.interpolate(self.disp_head(x), size=(H, W), mode="bilinear", align_corners=False) * 8.0
        sa, sb = a, b
        if self.full_depth:                                             # 1/32 context added to the 1/8 features
            sa = a + F.interpolate(da, size=a.shape[-2:], mode="bilinear", align_corners=False)
            sb = b + F.interpolate(db, size=b.shape[-2:], mode="bilinear", align_corners=False)
        seg_left = F.interpolate(self.seg_l(sa), size=(H, W), mode="bilinear", align_corners=False)
        seg_right = F.interpolate(self.seg_r(sb), size=(H, W), mode="bilinear", align_corners=False)
        at_d = self.att(disp)
        if self.fused_ops:
            seg_both, warped = warp_blend(seg_left, seg_right, -disp, at_d)
        else:
            warped = apply_disparity(seg_right, -disp)                  # sample the right branch at w - d
            seg_both = (1 - at_d) * seg_left + at_d * warped
        return seg_left, disp, seg_both, disp


def lovasz_softmax(probas, labels):
    """Lovasz-Softmax (Berman et al. 2018) over all classes of the batch -- the term the reference adds to the
    cross-entropy of seg2 (`-loss cross_entropy lovasz`, torch_implementation.py:292 -> util/lovasz_losses.py:153-199).
    Written from the published algorithm: per class, sort the absolute errors in decreasing order and dot them with the
    discrete gradient of the Jaccard index.  classes='all' (no data-dependent control flow, so the step stays
    capturable in a CUDA graph; the reference's default 'present' skips absent classes with a host-side test)."""
    B, K, H, W = probas.shape
    p = probas.permute(0, 2, 3, 1).reshape(-1, K)
    lab = labels.reshape(-1)
    losses = []
    for k in range(K):
        fg = (lab == k).to(p.dtype)
        err = (fg - p[:, k]).abs()
        err_sorted, perm = torch.sort(err, 0, descending=True)
        fg_sorted = fg[perm]
        gts = fg_sorted.sum()
        inter = gts - fg_sorted.cumsum(0)
        union = gts + (1.0 - fg_sorted).cumsum(0)
        jac = 1.0 - inter / union
        jac = torch.cat([jac[:1], jac[1:] - jac[:-1]])
        losses.append(torch.dot(err_sorted, jac))
    return torch.stack(losses).mean()


def sdnet_loss(outputs, seg_target, disp_target, lovasz: bool = False):
    """CE on seg1, CE (+ Lovasz-Softmax) on seg2, L1 on the disparity -- torch_implementation.py:279-305."""
    seg1, disp1, seg2, _ = outputs
    loss = (F.cross_entropy(seg1, seg_target) + F.cross_entropy(seg2, seg_target)
            + F.l1_loss(disp1.squeeze(1), disp_target))
    if lovasz:
        loss = loss + lovasz_softmax(F.softmax(seg2, dim=1), seg_target)
    return loss


def synthetic_batch(batch: int, h: int, w: int, n_labels: int, device, generator=None):
    left = torch.rand(batch, 3, h, w, device=device, generator=generator)
    right = torch.rand(batch, 3, h, w, device=device, generator=generator)
    disp = 64.0 * torch.rand(batch, h, w, device=device, generator=generator)
    seg = torch.randint(0, n_labels, (batch, h, w), device=device, generator=generator)
    return left, right, seg, disp


def build_training_step(world, batch_per_gpu: int = 4, h: int = 256, w: int = 512, n_labels: int = 2,
                        backbone: str = "densenet121", sync_bn: bool = True, cuda_graph: bool = False,
                        paired_tower: bool = False, peer_bn: bool = False, full_depth: bool = False,
                        fused_ops: bool = False, lovasz: bool = False):
    """Returns (step_fn, model): step_fn() runs one fwd + loss + bwd + Adam step on a fixed synthetic batch.

    cuda_graph=True captures the whole step -- forward, our hot-path kernels, backward, DDP's bucketed gradient
    all-reduce, SyncBatchNorm's statistics collectives and the Adam update -- into ONE CUDA graph.  The eager step
    is launch-bound (hundreds of tiny BN/NCCL launches per step, section 5.8 of SURVEY.md); replaying a graph
    removes the host from the loop, which is what lets the step scale across GPUs."""
    dev = torch.device("cuda", world.local_rank)
    torch.manual_seed(1234)  # identical initial weights on every rank
    # the tower is paired AFTER the SyncBatchNorm conversion (which would otherwise replace the paired layers -- and
    # lose the ReLUs they have taken over)
    model = SDNetLite(n_labels=n_labels, backbone=backbone, paired_tower=False, full_depth=full_depth,
                      fused_ops=fused_ops).to(dev)
    exchange = None
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):  # DDP must be built and warmed up on the stream family the graph is captured from
        if world.distributed:
            if sync_bn:
                model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
        if paired_tower and sync_bn:
            model.pair_tower()
            if peer_bn and world.distributed:
                # statistics of EVERY BN layer travel over NVLink peer memory inside the kernels: the paired tower layers
                # and (merged=True: whole-batch statistics) the decoders' -- no collective launch per layer is left
                sync_batchnorms_to_peer(model)
                exchange = PeerExchange.attach(model)
        if world.distributed:
            model = nn.parallel.DistributedDataParallel(model, device_ids=[world.local_rank])
        opt = torch.optim.Adam(model.parameters(), lr=1.5e-3, eps=1e-7, capturable=cuda_graph)
        g = torch.Generator(device=dev).manual_seed(world.rank)
        left, right, seg, disp = synthetic_batch(batch_per_gpu, h, w, n_labels, dev, g)

        def eager_step():
            opt.zero_grad(set_to_none=True)
            loss = sdnet_loss(model(left, right), seg, disp, lovasz)
            loss.backward()
            opt.step()
            return loss

        if cuda_graph:
            for _ in range(11):  # DDP needs >= 11 eager iterations before its collectives can be captured
                eager_step()
    torch.cuda.current_stream(dev).wait_stream(side)
    if not cuda_graph:
        eager_step.exchange = exchange
        return eager_step, model

    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    # thread_local: the NCCL watchdog thread may touch the CUDA runtime while this thread captures
    with torch.cuda.graph(graph, capture_error_mode="thread_local" if world.distributed else "global"):
        static_loss = sdnet_loss(model(left, right), seg, disp, lovasz)
        static_loss.backward()
        opt.step()

    def graph_step():
        graph.replay()
        return static_loss

    graph_step.graph = graph
    graph_step.exchange = exchange
    return graph_step, model
