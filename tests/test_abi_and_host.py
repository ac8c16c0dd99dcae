"""CPU-side checks: the C-ABI library loads without a GPU and exports every symbol include/pmt_ops.h declares;
the host shims validate arguments like the reference surface; nothing in the product package imports the oracle."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "pmt_learning_for_semantic_segmentation_and_disparity_b200")


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    ge.build()
    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as m
    return m


def declared_symbols():
    with open(os.path.join(ROOT, "include", "pmt_ops.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(pmt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(os.path.join(PKG, "libpmt_ops.so"))
    names = declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pmt_ops.h but not exported"
    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == names       # the ctypes table binds exactly the header
    assert built.load_library().pmt_version() >= 100


def test_argument_validation_without_gpu(built):
    lib = built.load_library()
    # null pointers / bad patch are rejected before anything touches a device
    assert lib.pmt_corr1d_fwd_f32(None, None, None, 1, 1, 1, 4, 3, 1, None) != 0
    assert b"null" in lib.pmt_last_error()
    assert lib.pmt_softargmin_fwd_f32(None, None, None, 1, 4, 1, 4, None) != 0
    assert lib.pmt_warp1d_bwd_f32(None, None, None, None, None, 1, 1, 1, 1, 0, None) != 0


def test_shims_fail_loudly_on_cpu_tensors(built):
    a = torch.zeros(1, 2, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        built.SpatialCorrelationSampler(1, (1, 5), 1, 0, 1, 1)(a, a)
    with pytest.raises(RuntimeError, match="no CPU path"):
        built.build_concat_volume(a, a, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        built.softargmin(a)
    with pytest.raises(RuntimeError, match="no CPU path"):
        built.apply_disparity(a, torch.zeros(1, 1, 3, 8))
    with pytest.raises(NotImplementedError):
        built.SpatialCorrelationSampler(kernel_size=3, patch_size=1)(a, a)
    assert built.apply_disparity(a, a, wrap_mode="other") is None   # reference returns None (torch_dsnet.py:21-22)


def test_missing_library_is_an_error(built, monkeypatch):
    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(PKG, "does_not_exist.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_compat_module_names(built):
    built.install_reference_shims()
    import spatial_correlation_sampler as scs
    assert scs.SpatialCorrelationSampler is built.SpatialCorrelationSampler
    s = scs.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 17), stride=1, padding=0, dilation_patch=1)
    assert list(s.parameters()) == []
    assert callable(scs.spatial_correlation_sample) and hasattr(scs, "SpatialCorrelationSamplerFunction")
    del sys.modules["spatial_correlation_sampler"]


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert "libpmt_oracle" not in src, f


def test_pair_batchnorms_conversion_host_logic(built):
    """pair_batchnorms: parameters/buffers are shared (not copied), a ReLU that follows a BN is taken over, and in eval
    mode (no kernel involved) the paired module computes what BN + ReLU computed."""
    import torch
    from torch import nn

    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt

    torch.manual_seed(0)
    net = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1), nn.BatchNorm2d(4), nn.ReLU(inplace=True),
                        nn.Sequential(nn.Conv2d(4, 4, 1), nn.BatchNorm2d(4)), nn.ReLU())
    with torch.no_grad():
        net[1].running_mean.uniform_(-1, 1), net[1].running_var.uniform_(0.5, 2), net[1].weight.uniform_(0.5, 1.5)
    x = torch.randn(4, 3, 6, 8)
    net.eval()
    want = net(x)
    bn_w, bn_rm = net[1].weight, net[1].running_mean
    pmt.pair_batchnorms(net, fuse_relu=True)
    assert isinstance(net[1], pmt.PairedSyncBatchNorm) and net[1].relu and isinstance(net[2], nn.Identity)
    assert isinstance(net[3][1], pmt.PairedSyncBatchNorm) and not net[3][1].relu   # its ReLU is not a sibling: left alone
    assert isinstance(net[4], nn.ReLU)
    assert net[1].weight is bn_w and net[1].running_mean is bn_rm
    assert torch.allclose(net(x), want, atol=1e-6)
    net.train()
    with pytest.raises(ValueError):
        net[1](torch.randn(3, 4, 6, 8))                                            # odd batch cannot be [left; right]


def test_pair_batchnorms_never_drops_a_shared_relu(built):
    """ResNet-style blocks register ONE nn.ReLU right after bn1 and apply it again after the residual add
    (reference: models/Resnet.py BasicBlock/Bottleneck).  The conversion must leave that ReLU alone -- with the default
    AND with fuse_relu=True -- and the default must not fuse anything."""
    import torch
    from torch import nn

    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt

    class BasicBlock(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(c)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c)

        def forward(self, x):
            out = self.relu(self.bn1(self.conv1(x)))
            out = self.bn2(self.conv2(out))
            return self.relu(out + x)

    torch.manual_seed(1)
    for fuse in (False, True):
        blk = BasicBlock(4).eval()
        with torch.no_grad():
            blk.bn1.running_mean.uniform_(-1, 1), blk.bn2.running_mean.uniform_(-1, 1)
        x = torch.randn(2, 4, 6, 8)
        want = blk(x)
        pmt.pair_batchnorms(blk, fuse_relu=fuse)
        assert isinstance(blk.bn1, pmt.PairedSyncBatchNorm) and isinstance(blk.relu, nn.ReLU) and not blk.bn1.relu
        got = blk(x)
        assert torch.equal(got, want) and float(got.min()) >= 0.0
    seq = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4), nn.ReLU())
    pmt.pair_batchnorms(seq)                                       # default: no fusion even where it would be safe
    assert isinstance(seq[2], nn.ReLU) and not seq[1].relu


def test_paired_batchnorm_host_contract(built):
    """Drop-in details of the converted layer: process_group of a SyncBatchNorm is kept; eval without running
    statistics needs batch statistics (so it raises on CPU instead of calling F.batch_norm(training=False) on None);
    fused=True refuses what it does not implement instead of falling back."""
    import torch
    from torch import nn

    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt

    sbn = nn.SyncBatchNorm(4)
    sbn.process_group = "sentinel-group"
    holder = nn.Sequential(sbn)
    pmt.pair_batchnorms(holder)
    assert holder[0].process_group == "sentinel-group"
    bn = pmt.PairedSyncBatchNorm(4, track_running_stats=False).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        bn(torch.randn(2, 4, 3, 3))                               # batch statistics even in eval: needs the CUDA kernels
