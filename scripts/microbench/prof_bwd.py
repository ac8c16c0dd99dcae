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
prof = torch.zeros(2*32*10, dtype=torch.int64, device=dev)
lib.pmt_debug_set_ptr(0, vp(prof))
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(3): lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
torch.cuda.synchronize()
p = prof.cpu().view(2, 32, 10)
names = {0:"band-producer",1:"MMA",2:"epi",3:"epi",4:"epi",5:"epi"}
print("debug", os.environ.get("PMT_TC_DEBUG", "0"))
for mode in (0,1):
    print(f"mode {mode}: warp: [wait0 wait1 wait2 -] / total | sec0 sec1 sec2 sec3 (kcycles)")
    for w in range(23):
        row = p[mode, w].tolist()
        if row[4] == 0: continue
        if w in (3,4,5,8,9,10,11,12,13,16,17,18,19,20): continue
        role = names.get(w, "builder" if w < 22 else "raw-producer")
        print(f"  w{w:2d} {role:14s} " + " ".join(f"{v/1e3:8.1f}" for v in row[:3]) + f"  / {row[4]/1e3:8.1f} | " + " ".join(f"{v/1e3:8.1f}" for v in row[5:9]))
