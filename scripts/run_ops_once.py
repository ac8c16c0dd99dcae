"""Launch every non-correlation kernel of the library a few times at its BASELINE configuration (target of
`ncu --set full -k regex:...`; config 3: PSMNet volume / soft-argmin at 256x512 maxdisp 192 batch 4; warp at 540x960)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
R = 3
B, C, D, H, W = 4, 32, 48, 64, 128
ref, tgt = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
cost = torch.empty(B, 2 * C, D, H, W, device=dev); gref, gtgt = torch.empty_like(ref), torch.empty_like(tgt)
for _ in range(R):
    lib.pmt_concat_volume_fwd_f32(vp(ref), vp(tgt), vp(cost), B, C, D, H, W, 0, sp)
    lib.pmt_concat_volume_bwd_f32(vp(cost), vp(gref), vp(gtgt), B, C, D, H, W, 0, sp)
del cost
B, D, H, W = 4, 192, 256, 512
c = 4 * torch.randn(B, D, H, W, device=dev); gc = torch.empty_like(c)
out, lse, go = (torch.randn(B, H, W, device=dev) for _ in range(3))
for _ in range(R):
    lib.pmt_softargmin_fwd_f32(vp(c), vp(out), vp(lse), B, D, H, W, sp)
    lib.pmt_softargmin_bwd_f32(vp(c), vp(out), vp(lse), vp(go), vp(gc), B, D, H, W, sp)
    lib.pmt_dispreg_fwd_f32(vp(c), vp(out), B, D, H, W, sp)
    lib.pmt_dispreg_bwd_f32(vp(go), vp(gc), B, D, H, W, sp)
low = 3.0 * torch.randn(B, 1, D // 4, H // 4, W // 4, device=dev); glow = torch.empty_like(low)
work = torch.empty(B, D // 4, H, W, device=dev)
for _ in range(R):
    lib.pmt_upsample_softargmin_fwd_f32(vp(low), vp(out), vp(lse), B, D // 4, H // 4, W // 4, D, H, W, sp)
    lib.pmt_upsample_softargmin_bwd_f32(vp(low), vp(out), vp(lse), vp(go), vp(work), vp(glow), B, D // 4, H // 4, W // 4, D, H, W, sp)
del c, gc, work
for (N, C, H, W) in [(4, 3, 540, 960), (1, 128, 540, 960)]:
    img = torch.randn(N, C, H, W, device=dev); off = -64.0 * torch.rand(N, 1, H, W, device=dev)
    o = torch.empty(C, N, H, W, device=dev); g = torch.randn(C, N, H, W, device=dev)
    gi, gof = torch.empty_like(img), torch.empty_like(off)
    for _ in range(R):
        lib.pmt_warp1d_fwd_f32(vp(img), vp(off), vp(o), N, C, H, W, 1, sp)
        lib.pmt_warp1d_bwd_f32(vp(img), vp(off), vp(g), vp(gi), vp(gof), N, C, H, W, 1, sp)
# bn_pair at a DenseNet-121 layer shape of the harness: (2*4, 128, 64, 128)
x = torch.randn(8, 128, 64, 128, device=dev).requires_grad_(True)
bn = pmt.PairedSyncBatchNorm(128).to(dev).train(); bn.relu = True
for _ in range(R):
    bn(x).sum().backward()
# f2 / f3 at the production feature shape
a, b = torch.randn(4, 352, 32, 64, device=dev, requires_grad=True), torch.randn(4, 352, 32, 64, device=dev, requires_grad=True)
wt = (0.1 * torch.randn(128, 17, device=dev)).requires_grad_(True)
for _ in range(R):
    pmt.correlation_conv1x1_relu(a, b, wt).sum().backward()
    pmt.spatial_correlation_sample(a, b, patch_size=(17, 17)).sum().backward()
torch.cuda.synchronize()
print("run_ops_once: done")
