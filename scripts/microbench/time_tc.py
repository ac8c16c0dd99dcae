import ctypes, sys, os
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
B, C, H, W, P = 4, 64, 256, 512, 192
L = torch.randn(B, C, H, W, device=dev); R = torch.randn(B, C, H, W, device=dev); G = torch.randn(B,1,P,H,W, device=dev)
out = torch.empty(B, 1, P, H, W, device=dev); g1=torch.empty_like(L); g2=torch.empty_like(L)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
which = sys.argv[1]
def f(passes):
    if which == "fwd": return lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, passes, st)
    return lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
for passes in (1, 3):
    for _ in range(3): assert f(passes) == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f(passes)
    e1.record(); torch.cuda.synchronize()
    print(f"{which} debug={os.environ.get('PMT_TC_DEBUG','0')} passes={passes}: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
