import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
lib = pmt.load_library(); dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, C, H, W, P = 1, 16, 1, 128, 192
L = torch.ones(B,C,H,W,device=dev); R = torch.ones(B,C,H,W,device=dev)
out = torch.full((B,1,P,H,W), float('nan'), device=dev)
rc = lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, 1, st)
torch.cuda.synchronize()
o = out[0,0,:,0,:].cpu().numpy()
np.set_printoptions(linewidth=250, precision=0, suppress=True)
print("debug", os.environ.get("PMT_TC_DEBUG"), "rc", rc, "absmax", np.nanmax(np.abs(o)), "nans", np.isnan(o).sum())
# expected pattern: out[p][wl] = wl*1000 + (p + delta + wl), delta=1
exp = np.arange(128)[None,:]*1000 + (np.arange(192)[:,None] + 1 + np.arange(128)[None,:])
print("pattern match:", np.array_equal(o, exp.astype(np.float32)))
print(o[0,:6], o[5,:6], o[191,120:])
