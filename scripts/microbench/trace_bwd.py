import ctypes, sys, os
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = ctypes.CDLL(os.environ["PMT_PROF_LIB"]) if os.environ.get("PMT_PROF_LIB") else pmt.load_library()  # a -DPMT_BWD_PROFILE build
dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
B, C, H, W, P = 4, 64, 256, 512, 192
L = torch.randn(B, C, H, W, device=dev); R = torch.randn(B, C, H, W, device=dev); G = torch.randn(B,1,P,H,W, device=dev)
g1=torch.empty_like(L); g2=torch.empty_like(L)
prof = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.pmt_debug_set_ptr(0, vp(prof))
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2): lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, 3, st)
torch.cuda.synchronize()
p = prof.cpu()
print("debug", os.environ.get("PMT_TC_DEBUG", "0"))
for mode in (0, 1):
    t = p[1024 + mode*1024: 2048 + mode*1024].view(8, 16, 8)
    base = int(t[t > 0].min()) if (t > 0).any() else 0
    print(f"mode {mode}: chunk | bandprod(wake) | MMA(enter wait, pass wait, after commit) | builder grp0/1 (enter, boxes ok, gd_empty ok, built, enter band wait, band ok, arrived)")
    for g in range(16):
        f = lambda role, idx: (int(t[role, g, idx]) - base) if int(t[role, g, idx]) > 0 else -1
        grp = g & 1
        print(f"  {g:2d} | {f(0,0):6d} | {f(1,0):6d} {f(1,1):6d} {f(1,2):6d} | " + " ".join(f"{f(2+grp,i):6d}" for i in range(7)))
