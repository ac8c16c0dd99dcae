"""Per-role wait counters of corr1d_bwd_tca_kernel (needs a library built with PMT_BWD_PROFILE=1)."""
import ctypes, sys, os
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
B, C, H, W, P = 4, 64, 256, 512, 192
L = torch.randn(B, C, H, W, device=dev); R = torch.randn(B, C, H, W, device=dev); G = torch.randn(B,1,P,H,W, device=dev)
g1=torch.empty_like(L); g2=torch.empty_like(L)
prof = torch.zeros(2*17*8, dtype=torch.int64, device=dev)
lib.pmt_debug_set_ptr.argtypes = [ctypes.c_int, ctypes.c_void_p]
lib.pmt_debug_set_ptr(0, vp(prof))
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(3): assert lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st) == 0
torch.cuda.synchronize()
p = prof.cpu().view(2, 17, 8)
names = {0: "band-prod [w0 band_empty]", 1: "MMA [w0 tmem_empty w1 a_built w2 band_ready]", 2: "epi [w0 tmem_full]", 6: "builder g0 [w0 raw_full w1 a_empty s5 wait::st]",
         10: "builder g1", 14: "split [w0 band_full]", 16: "raw-prod [w1 raw_empty]"}
for mode in (0, 1):
    print(f"mode {mode} (kcycles): w0 w1 w2 w3 | total | s5 s6 s7")
    for w, nm in names.items():
        row = p[mode, w].tolist()
        print(f"  w{w:2d} " + " ".join(f"{v/1e3:8.1f}" for v in row[:4]) + f" | {row[4]/1e3:8.1f} | " + " ".join(f"{v/1e3:8.1f}" for v in row[5:8]) + f"   {nm}")
