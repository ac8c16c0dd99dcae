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
                                           int(relu), U.stream_ptr(dev))
        U._lib.check(st, "pmt_bn_pair_apply_f32")
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
        ctx.group, ctx.world_size, ctx.relu = group, world_size, int(relu)   # flags: bit 0 ReLU, bit 1 merged halves
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


class PeerExchange:
    """NVLink peer-memory exchange of the BN statistics of every PairedSyncBatchNorm layer of a model (csrc/bn_pair.cu,
    include/pmt_ops.h `pmt_bn_pair_*_peer_f32`): one symmetric buffer per rank, mapped into every peer, holds for each
    layer and direction a payload region [2 parity][world][n] and a flag region [2][world]; the kernels push their
    payload into every peer's copy and spin on their own flags, so no collective is launched on the dependency chain
    (DDP's gradient all-reduce stays with NCCL).

    `PeerExchange.attach(module, group)` allocates the buffer with torch.distributed._symmetric_memory, assigns every
    PairedSyncBatchNorm below `module` its offsets and switches those layers to the peer kernels.
    `PeerExchange.emulated(world, ...)` builds `world` plain buffers on ONE device (tests: the ranks run one after the
    other, so the consumers never have to wait)."""

    def __init__(self, world, rank, local, ptr_table, keepalive=()):
        self.world, self.rank = int(world), int(rank)
        self.local = local                      # this rank's buffer (float32, flat)
        self.ptr_table = ptr_table              # int64 device tensor [world]: base pointers of all ranks' buffers
        self._keepalive = keepalive
        self.cursor = 0                         # floats handed out so far (identical on all ranks)
        self.wait = 1                           # producers wait for their peers inside the kernel (0: emulated ranks)
        dev = local.device
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self._counters = []

    # ---- layout ----
    @staticmethod
    def floats_needed(module: nn.Module, world: int) -> int:
        total = 0
        for m in module.modules():
            if isinstance(m, PairedSyncBatchNorm):
                c = m.num_features
                for n in (4 * c + 1, 4 * c):
                    total += _round4(2 * world * n) + _round4(2 * world)
        return total

    def reserve(self, n: int):
        """(payload_off, flag_off, epoch, done) for one layer+direction with `n` payload floats per rank."""
        payload_off = self.cursor
        self.cursor += _round4(2 * self.world * n)
        flag_off = self.cursor
        self.cursor += _round4(2 * self.world)
        if self.cursor > self.local.numel():
            raise RuntimeError("PeerExchange buffer too small")
        counters = torch.zeros(4, dtype=torch.int32, device=self.local.device)   # [epoch, done, ready, -]
        self._counters.append(counters)
        return payload_off, flag_off, counters

    # ---- construction ----
    @classmethod
    def attach(cls, module: nn.Module, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        group = group if group is not None else dist.group.WORLD
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        n = max(cls.floats_needed(module, world), 4)
        buf = symm.empty(n, dtype=torch.float32, device=dev)
        buf.zero_()
        hdl = symm.rendezvous(buf, group)
        torch.cuda.synchronize(dev)
        hdl.barrier()                            # every rank's flags are zero before anyone publishes
        table = torch.tensor([int(p) for p in hdl.buffer_ptrs], dtype=torch.int64, device=dev)
        self = cls(world, rank, buf, table, keepalive=(hdl,))
        self._bind(module, group)
        return self

    @classmethod
    def emulated(cls, world: int, floats: int, device):
        bufs = [torch.zeros(max(floats, 4), dtype=torch.float32, device=device) for _ in range(world)]
        table = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=device)
        ranks = [cls(world, r, bufs[r], table, keepalive=tuple(bufs)) for r in range(world)]
        for x in ranks:
            x.wait = 0                          # one device runs the ranks one after the other: nobody may wait in a kernel
        return ranks

    def _bind(self, module: nn.Module, group):
        for m in module.modules():
            if isinstance(m, PairedSyncBatchNorm):
                c = m.num_features
                m.peer = (self, self.reserve(4 * c + 1), self.reserve(4 * c))
                m.process_group = group

    def check(self):
        """Raise if a kernel gave up waiting for a peer (reads one device word: call outside the hot loop)."""
        if int(self.err.item()) != 0:
            raise RuntimeError("PeerExchange: a batch-norm kernel waited > 2 s for the statistics of a peer rank")


def _round4(n: int) -> int:
    return (int(n) + 3) & ~3


class _PairedSyncBNPeerFn(torch.autograd.Function):
    """_PairedSyncBNFusedFn with the exchange done inside the kernels over NVLink peer memory (no collective launch)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, peer, relu=False):
        import ctypes

        from . import _util as U

        xch, (f_pay, f_flag, f_cnt), bwd = peer
        x = x.contiguous()
        B2, C = x.size(0), x.size(1)
        B, HW = B2 // 2, x[0, 0].numel()
        dev = x.device
        lib = U._lib.load()
        vp = ctypes.c_void_p
        with torch.cuda.device(dev):
            st = lib.pmt_bn_pair_stats_peer_f32(U.ptr(x), vp(xch.ptr_table.data_ptr()), vp(xch.local.data_ptr()), xch.world,
                                                xch.rank, f_pay, f_flag, vp(f_cnt.data_ptr()), vp(f_cnt.data_ptr() + 4),
                                                vp(xch.err.data_ptr()), xch.wait, B, C, HW, U.stream_ptr(dev))
            U._lib.check(st, "pmt_bn_pair_stats_peer_f32")
            out = torch.empty_like(x)
            save_mean = torch.empty(2 * C, device=dev, dtype=torch.float32)
            save_invstd = torch.empty(2 * C + 1, device=dev, dtype=torch.float32)   # [2C] = total count
            st = lib.pmt_bn_pair_apply_peer_f32(U.ptr(x), vp(xch.local.data_ptr()), xch.world, f_pay, f_flag,
                                                vp(f_cnt.data_ptr()), vp(xch.err.data_ptr()), U.ptr(weight), U.ptr(bias),
                                                U.ptr(running_mean), U.ptr(running_var), ctypes.c_float(momentum),
                                                ctypes.c_float(eps), U.ptr(out), U.ptr(save_mean), U.ptr(save_invstd), B, C,
                                                HW, int(relu), U.stream_ptr(dev))
            U._lib.check(st, "pmt_bn_pair_apply_peer_f32")
        ctx.save_for_backward(x, weight, bias, save_mean, save_invstd)
        ctx.peer, ctx.relu = (xch, bwd), int(relu)
        return out

    @staticmethod
    def backward(ctx, grad):
        import ctypes

        from . import _util as U

        x, weight, bias, save_mean, save_invstd = ctx.saved_tensors
        xch, (b_pay, b_flag, b_cnt) = ctx.peer
        grad = grad.contiguous()
        B2, C = x.size(0), x.size(1)
        B, HW = B2 // 2, x[0, 0].numel()
        dev = x.device
        lib = U._lib.load()
        vp = ctypes.c_void_p
        gwb = torch.zeros(2, C, device=dev, dtype=torch.float32)
        dx = torch.empty_like(x)
        with torch.cuda.device(dev):
            st = lib.pmt_bn_pair_bwd_reduce_peer_f32(U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd),
                                                     vp(xch.ptr_table.data_ptr()), vp(xch.local.data_ptr()), xch.world, xch.rank,
                                                     b_pay, b_flag, vp(b_cnt.data_ptr()), vp(b_cnt.data_ptr() + 4),
                                                     vp(xch.err.data_ptr()), xch.wait, U.ptr(gwb[0]), U.ptr(gwb[1]), B, C, HW,
                                                     U.ptr(weight), U.ptr(bias), ctx.relu, U.stream_ptr(dev))
            U._lib.check(st, "pmt_bn_pair_bwd_reduce_peer_f32")
            st = lib.pmt_bn_pair_bwd_apply_peer_f32(U.ptr(grad), U.ptr(x), U.ptr(save_mean), U.ptr(save_invstd), U.ptr(weight),
                                                    vp(xch.local.data_ptr()), xch.world, b_pay, b_flag, vp(b_cnt.data_ptr()),
                                                    vp(xch.err.data_ptr()), U.ptr(dx), B, C, HW, U.ptr(bias), ctx.relu,
                                                    U.stream_ptr(dev))
            U._lib.check(st, "pmt_bn_pair_bwd_apply_peer_f32")
        gw = gwb[0] if weight is not None else None
        gb = gwb[1] if weight is not None else None
        return dx, gw, gb, None, None, None, None, None, None


class PairedSyncBatchNorm(nn.BatchNorm2d):
    """Drop-in for the BatchNorm2d / SyncBatchNorm layers of a siamese tower that is fed [left; right] in one pass (see
    _PairedSyncBNFn).  Single process: equals calling the BatchNorm2d on each half in turn.

    `fused=True` (default) runs this package's kernels and is strict: fp32 CUDA input and a numeric `momentum`, anything
    else raises (no silent fallback).  `fused=False` is an explicit opt-in to the composition of ATen ops (any dtype; it
    is the float64 reference of the tests and supports `momentum=None`, the cumulative moving average).
    `process_group` restricts the statistics exchange to a sub-group like nn.SyncBatchNorm's argument of that name."""

    fused = True
    relu = False   # True: the ReLU that follows this BN is computed by the same kernels (pair_batchnorms sets it)
    merged = False  # True: the two halves are ONE batch -- plain (Sync)BatchNorm semantics over the whole (even) batch, on
                    # the same kernels and the same exchange (sync_batchnorms_to_peer converts the non-siamese layers)
    process_group = None
    peer = None    # (PeerExchange, forward slot, backward slot): statistics travel over NVLink peer memory, no collective

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
                self.num_batches_tracked.add_(1 if self.merged else 2)
            if x.data_ptr() % 16:
                x = x.clone(memory_format=torch.contiguous_format)   # an offset view: the kernels read 16-byte vectors
            flags = int(bool(self.relu)) | (2 if self.merged else 0)
            if self.peer is not None:
                return _PairedSyncBNPeerFn.apply(x, self.weight, self.bias, self.running_mean if track else None,
                                                 self.running_var if track else None, self.eps,
                                                 self.momentum if self.momentum is not None else 0.0, self.peer, flags)
            return _PairedSyncBNFusedFn.apply(x, self.weight, self.bias, self.running_mean if track else None,
                                              self.running_var if track else None, self.eps,
                                              self.momentum if self.momentum is not None else 0.0, self.process_group, ws,
                                              flags)
        if self.merged:
            raise NotImplementedError("PairedSyncBatchNorm(merged=True) is implemented by the fused kernels only")
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


def sync_batchnorms_to_peer(module: nn.Module) -> nn.Module:
    """Convert the remaining nn.SyncBatchNorm / BatchNorm2d layers below `module` (the non-siamese ones: decoders, heads)
    into PairedSyncBatchNorm(merged=True): whole-batch statistics, i.e. what nn.SyncBatchNorm computes, but on the bn_pair
    kernels -- so that `PeerExchange.attach` can route their statistics over NVLink peer memory as well and no BN layer of
    the model launches a collective.  The per-GPU batch must be even (the kernels split it in two halves internally)."""
    for name, child in list(module.named_children()):
        if isinstance(child, PairedSyncBatchNorm):
            continue
        if isinstance(child, (nn.BatchNorm2d, nn.SyncBatchNorm)):
            new = PairedSyncBatchNorm(child.num_features, child.eps, child.momentum, child.affine, child.track_running_stats)
            new.weight, new.bias = child.weight, child.bias
            new.training = child.training
            new.running_mean, new.running_var, new.num_batches_tracked = (child.running_mean, child.running_var,
                                                                          child.num_batches_tracked)
            new.process_group = getattr(child, "process_group", None)
            new.merged = True
            setattr(module, name, new)
        else:
            sync_batchnorms_to_peer(child)
    return module
