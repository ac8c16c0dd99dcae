import ctypes, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, C, H, W, P = 1, 16, 1, 128, 192
for name, L, R in [("ones", torch.ones(B,C,H,W,device=dev), torch.ones(B,C,H,W,device=dev)),
                   ("ramp", torch.ones(B,C,H,W,device=dev), torch.arange(W,device=dev).float().view(1,1,1,W).expand(B,C,H,W).contiguous())]:
    out = torch.full((B,1,P,H,W), float('nan'), device=dev); ref = torch.empty_like(out)
    lib.pmt_corr1d_fwd_f32(vp(L), vp(R), vp(ref), B, C, H, W, P, 1, st)
    rc = lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, 1, st)
    torch.cuda.synchronize()
    o = out[0,0,:,0,:].cpu().numpy(); r = ref[0,0,:,0,:].cpu().numpy()
    print(name, "rc", rc, "out absmax", np.abs(o).max(), "ref absmax", np.abs(r).max(), "nonzero", (o!=0).sum(), "of", o.size)
    np.set_printoptions(linewidth=250, precision=0, suppress=True)
    for p in (0, 95, 96, 191): print(" p", p, "out", o[p, ::8]); print("      ref", r[p, ::8])
