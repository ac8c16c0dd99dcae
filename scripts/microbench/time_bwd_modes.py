"""Time the 3xTF32 backward of one library build at the headline shape (burst, 20 launches).  Knobs come from the
environment of a development build (PMT_BWD_GEN2, PMT_TC_DEBUG 2048/4096 = skip gin2/gin1 in gen 1, PMT_TCA_ONLY in gen 2,
PMT_BWD_SPLIT).  usage: time_bwd_modes.py lib.so [label]"""
import ctypes, os, sys
import torch
dev = torch.device("cuda:0"); vp = lambda t: ctypes.c_void_p(t.data_ptr()); I, P_ = ctypes.c_int, ctypes.c_void_p
lib = ctypes.CDLL(sys.argv[1])
lib.pmt_corr1d_bwd_tc_f32.argtypes = [P_, P_, P_, P_, P_, I, I, I, I, I, I, I, P_]
B, C, H, W, P = 4, 64, 256, 512, 192
L = torch.randn(B, C, H, W, device=dev); R = torch.randn(B, C, H, W, device=dev); G = torch.randn(B, 1, P, H, W, device=dev)
g1 = torch.empty_like(L); g2 = torch.empty_like(L)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
passes = int(os.environ.get("PASSES", "3"))
f = lambda: lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
for _ in range(3): assert f() == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): f()
e1.record(); torch.cuda.synchronize()
print(f"{sys.argv[2] if len(sys.argv) > 2 else ''}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us", flush=True)
