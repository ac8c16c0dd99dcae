"""-m gpu: the training-step harness (hot-path ops inside a real fwd+bwd+optimizer step) runs and learns."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sdnet_lite_step_small_tower():
    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import harness, sharding

    world = sharding.World(0, 0, 1, None)
    step, model = harness.build_training_step(world, batch_per_gpu=2, h=64, w=128, backbone="small")
    losses = [float(step()) for _ in range(8)]
    assert all(l == l for l in losses)            # finite
    assert losses[-1] < losses[0]                 # Adam on a fixed batch makes progress through our backward kernels
    # every parameter that feeds the hot path received a gradient
    m = model
    assert m.corrConv2d[0].weight.grad is not None and m.reduce[0].weight.grad is not None
    assert m.seg_r[1].weight.grad is not None     # reached only through apply_disparity's image gradient
    assert m.disp_head.weight.grad.abs().sum() > 0


def test_sdnet_lite_densenet_shapes():
    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import harness

    net = harness.SDNetLite().cuda()
    left = torch.rand(1, 3, 256, 512, device="cuda")
    seg1, disp, seg2, _ = net(left, left)
    assert seg1.shape == (1, 2, 256, 512) and disp.shape == (1, 1, 256, 512) and seg2.shape == seg1.shape


def test_paired_tower_matches_two_tower_calls(monkeypatch):
    """PairedSyncBatchNorm over [left; right] == the same BatchNorm2d applied to left, then to right
    (per-call batch statistics, running stats updated twice), forward and backward."""
    from pmt_learning_for_semantic_segmentation_and_disparity_b200 import harness

    DEV = torch.device("cuda:0")
    # cuDNN picks different (TF32) algorithms for batch B and batch 2B; compare the BN maths in full fp32
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    torch.manual_seed(0)
    ref = harness.SDNetLite(backbone="small").to(DEV).train()
    par = harness.SDNetLite(backbone="small").to(DEV).train()
    par.load_state_dict(ref.state_dict())
    par.pair_tower()
    left, right, seg, disp = harness.synthetic_batch(2, 64, 128, 2, DEV)
    lr, lp = harness.sdnet_loss(ref(left, right), seg, disp), harness.sdnet_loss(par(left, right), seg, disp)
    assert abs(float(lr) - float(lp)) <= 1e-4 * abs(float(lr)), (float(lr), float(lp))
    lr.backward()
    lp.backward()

    def close(a, b, tol, what):
        err = float((a.double() - b.double()).abs().max())
        scale = float(a.double().abs().max())
        assert err <= tol * scale + 1e-7, f"{what}: max err {err:.3e} vs scale {scale:.3e}"

    # different BN kernels (cuDNN fused vs stats/elemt) and different conv batch sizes: fp32 round-off only
    for (n, a), (_, b) in zip(ref.named_parameters(), par.named_parameters()):
        close(a.grad, b.grad, 5e-3, n)
    for (n, a), (_, b) in zip(ref.named_buffers(), par.named_buffers()):
        close(a.float(), b.float(), 1e-4, n)
