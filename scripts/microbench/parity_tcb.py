import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
def run(B, C, H, W, P, passes, iters=0):
    g = torch.Generator(device=dev).manual_seed(1)
    L = torch.randn(B, C, H, W, device=dev, generator=g); R = torch.randn(B, C, H, W, device=dev, generator=g)
    G = torch.randn(B, 1, P, H, W, device=dev, generator=g)
    r1 = torch.empty_like(L); r2 = torch.empty_like(L)
    g1 = torch.full_like(L, float('nan')); g2 = torch.full_like(L, float('nan'))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pmt_corr1d_bwd_simt_f32(vp(L), vp(R), vp(G), vp(r1), vp(r2), B, C, H, W, P, 1, st) == 0
    rc = lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
    if rc != 0:
        print("rc", rc, lib.pmt_last_error()); return
    torch.cuda.synchronize()
    e1 = ((g1 - r1).abs().max() / r1.abs().max()).item(); e2 = ((g2 - r2).abs().max() / r2.abs().max()).item()
    msg = f"B{B} C{C} H{H} W{W} P{P} passes={passes}: err g1={e1:.3e} g2={e2:.3e} nans={torch.isnan(g1).sum().item()+torch.isnan(g2).sum().item()}"
    if iters:
        e0, ee = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
        e0.record()
        for _ in range(iters): lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
        ee.record(); torch.cuda.synchronize()
        msg += f"  {e0.elapsed_time(ee)/iters*1e3:.1f} us/launch"
    print(msg, flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    run(1, 16, 1, 128, 192, 1)
    run(1, 64, 2, 256, 192, 1)
    run(1, 64, 2, 256, 192, 3)
    run(2, 64, 8, 512, 192, 3)
    run(1, 40, 3, 132, 40, 3)
    run(1, 128, 4, 64, 17, 3)
    run(1, 5, 3, 100, 8, 3)
    run(2, 352, 32, 64, 17, 3)      # production call of minidsnetExt: three channel blocks
    run(1, 130, 5, 260, 193, 3)     # 2 channel blocks, P > 192, ragged W
    run(1, 96, 4, 256, 100, 1)
else:
    run(4, 64, 256, 512, 192, 1, iters=20)
    run(4, 64, 256, 512, 192, 3, iters=20)
