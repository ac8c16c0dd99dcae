"""Every kernel of the library once at small shapes (target of compute-sanitizer memcheck / racecheck runs)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)


def corr(B, C, H, W, patch, engine="auto", dil=1):
    prev = pmt.set_correlation_engine(engine)
    try:
        a, b = rn(B, C, H, W).requires_grad_(True), rn(B, C, H, W).requires_grad_(True)
        out = pmt.spatial_correlation_sample(a, b, patch_size=patch, dilation_patch=dil)
        out.backward(rn(*out.shape))
    finally:
        pmt.set_correlation_engine(prev)


corr(1, 64, 2, 256, (1, 192))            # tcgen05 forward (TMEM-A) + first-generation backward, 3xTF32
corr(1, 64, 2, 256, (1, 192), "tf32")    # plain TF32
corr(1, 130, 2, 132, (1, 17))            # second-generation backward (two channel blocks)
corr(1, 16, 2, 128, (1, 40), "simt")     # CUDA-core tiled
corr(1, 4, 3, 62, (1, 9))                # generic kernels (W % 4 != 0)
corr(1, 8, 6, 16, (5, 17))               # 2-D row-pass kernels
corr(1, 3, 5, 20, (3, 5), dil=2)         # generic 2-D, dilated
a, b = rn(2, 37, 3, 16).requires_grad_(True), rn(2, 37, 3, 16).requires_grad_(True)
w = rn(8, 17).requires_grad_(True)
pmt.correlation_conv1x1_relu(a, b, w).backward(rn(2, 8, 3, 16))          # f2
r, t = rn(1, 4, 6, 32).requires_grad_(True), rn(1, 4, 6, 32).requires_grad_(True)
pmt.build_concat_volume(r, t, 7).backward(rn(1, 8, 7, 6, 32))
pmt.matchshifted()(r, t, 3)
c = (4 * rn(1, 24, 6, 32)).requires_grad_(True)
pmt.softargmin(c).backward(rn(1, 6, 32))
x = torch.rand(1, 24, 6, 32, device=dev).requires_grad_(True)
pmt.disparityregression(24)(x).backward(rn(1, 6, 32))
low = (3 * rn(1, 1, 6, 3, 8)).requires_grad_(True)
pmt.upsample_softargmin(low, 24, (12, 32)).backward(rn(1, 12, 32))
img, off = rn(2, 3, 6, 32).requires_grad_(True), (torch.rand(2, 1, 6, 32, device=dev) * 10 - 7).requires_grad_(True)
pmt.apply_disparity(img, off).backward(rn(2, 3, 6, 32))
seg, att = rn(2, 3, 6, 32).requires_grad_(True), torch.rand(2, 1, 6, 32, device=dev).requires_grad_(True)
both, warped = pmt.warp_blend(seg, img, off, att)
(both.sum() + warped.sum()).backward()
pmt.photo_consistency_mse(img, off, seg, True).backward()
bn = pmt.PairedSyncBatchNorm(8).to(dev).train()
bn.relu = True
xb = rn(4, 8, 6, 10).requires_grad_(True)
bn(xb).sum().backward()
bn.merged = True
bn(xb).sum().backward()
torch.cuda.synchronize()
print("sanitize_small: all ops ran")
