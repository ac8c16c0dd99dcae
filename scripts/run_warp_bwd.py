"""Run the warp backward at the config-4 feature shape a few times (target of an ncu capture)."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
N, C, H, W = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (1, 128, 540, 960)))
img = torch.randn(N, C, H, W, device=dev); off = -64.0 * torch.rand(N, 1, H, W, device=dev)
g = torch.randn(C, N, H, W, device=dev); gi = torch.empty_like(img); go = torch.empty_like(off)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    assert lib.pmt_warp1d_bwd_f32(vp(img), vp(off), vp(g), vp(gi), vp(go), N, C, H, W, 1, st) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.pmt_warp1d_bwd_f32(vp(img), vp(off), vp(g), vp(gi), vp(go), N, C, H, W, 1, st)
e1.record(); torch.cuda.synchronize()
print(f"warp bwd N={N} C={C} {H}x{W}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
