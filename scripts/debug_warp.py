"""Locate mismatches of the row-CSR warp backward against the differentiable oracle (development helper)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt
from oracle import torch_ref
DEV = "cuda:0"

def case(N, C, H, W, special=True):
    g = torch.Generator().manual_seed(7 + W)
    img = torch.randn(N, C, H, W, generator=g)
    off = -(W / 6.0) * torch.rand(N, 1, H, W, generator=g) + 2.0
    if special:
        off[0, 0, 0, :] = float(W)
        off[0, 0, H - 1, :] = -torch.arange(W).float()
    gout = torch.randn(N, C, H, W, generator=g)
    a = img.to(DEV).requires_grad_(True); o = off.to(DEV).requires_grad_(True)
    pmt.apply_disparity(a, o).backward(gout.to(DEV))
    a64, o64 = img.clone().requires_grad_(True), off.clone().requires_grad_(True)
    torch_ref.warp_ref(a64, o64).backward(gout)
    for name, got, want in (("gimg", a.grad.cpu().double(), a64.grad), ("goff", o.grad.cpu().double(), o64.grad)):
        d = (got - want).abs()
        bad = (d > 1e-4 * want.abs().max()).nonzero()
        print(f"{(N,C,H,W)} special={special} {name}: max err {float(d.max()):.3e} / scale {float(want.abs().max()):.3e}; bad {bad.shape[0]}",
              [tuple(int(v) for v in b) for b in bad[:6]], flush=True)
        for b in bad[:3]:
            b = tuple(int(v) for v in b)
            print("    got", float(got[b]), "want", float(want[b]), "off", float(off[b[0], 0, b[2], b[3]]))

for cfg in [(4, 3, 64, 512), (1, 3, 64, 512), (4, 3, 8, 512), (2, 2, 7, 100), (4, 3, 64, 256)]:
    case(*cfg)
case(4, 3, 64, 512, special=False)
