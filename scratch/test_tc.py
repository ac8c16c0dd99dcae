import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library()
dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
def run(B, C, H, W, P, passes, check=True, iters=0):
    g = torch.Generator(device=dev).manual_seed(1)
    L = torch.randn(B, C, H, W, device=dev, generator=g); R = torch.randn(B, C, H, W, device=dev, generator=g)
    ref = torch.empty(B, 1, P, H, W, device=dev); out = torch.full((B, 1, P, H, W), float('nan'), device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert lib.pmt_corr1d_fwd_simt_f32(vp(L), vp(R), vp(ref), B, C, H, W, P, 1, st) == 0
    rc = lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, passes, st)
    if rc != 0:
        print("rc", rc, lib.pmt_last_error()); return
    torch.cuda.synchronize()
    err = ((out - ref).abs().max() / ref.abs().max()).item()
    nan = torch.isnan(out).sum().item()
    msg = f"B{B} C{C} H{H} W{W} P{P} passes={passes}: rel_err={err:.3e} nans={nan}"
    if iters:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, passes, st)
        e0.record()
        for _ in range(iters): lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, passes, st)
        e1.record(); torch.cuda.synchronize()
        msg += f"  {e0.elapsed_time(e1)/iters*1e3:.1f} us/launch"
    print(msg, flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "small"
if mode == "small":
    run(1, 16, 1, 128, 192, 1)
    run(1, 64, 2, 256, 192, 1)
    run(1, 64, 2, 256, 192, 3)
    run(2, 64, 8, 512, 192, 3)
    run(1, 40, 3, 132, 40, 3)
    run(1, 352, 4, 64, 17, 3)
else:
    run(4, 64, 256, 512, 192, 1, iters=20)
    run(4, 64, 256, 512, 192, 3, iters=20)
