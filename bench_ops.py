#!/usr/bin/env python
"""bench_ops.py -- per-op device timings and roofline fractions for every row of SURVEY.md section 8(a):
correlation engines (incl. BASELINE config 4, C=128 540x960), concat volume, soft-argmin / disparityregression,
apply_disparity.  One JSON object per line.  Inputs resident in HBM, CUDA events, rotating buffers > L2.

    python bench_ops.py [--iters 50] [--out profiles/ops.jsonl]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run_ops(iters: int = 50, device_index: int = 0, emit=None, quick: bool = False):
    """Time every op; returns the list of result dicts (emit(d) is called per line).  quick=True skips the
    CUDA-core / plain-TF32 engine rows and the ATen comparison rows (what bench.py folds into its `ops` record)."""
    import torch

    import pmt_learning_for_semantic_segmentation_and_disparity_b200 as pmt

    class _A:
        pass

    args = _A()
    args.iters = iters
    dev = torch.device("cuda", device_index)
    lib = pmt.load_library()
    hbm = 6525.2
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm = float(json.load(open(mp))["hbm_gbs"])
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    lines = []

    def timed_eager(fn, n_sets):
        for i in range(5):
            fn(i % n_sets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.iters):
            fn(i % n_sets)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    def timed(fn, n_sets):
        """Device time per launch.  Launches shorter than ~100 us are host-bound when issued from Python (a ctypes call
        costs 5-10 us), so they are re-measured as a CUDA graph of `iters` launches: back-to-back on the device."""
        ms = timed_eager(fn, n_sets)
        if ms >= 0.1:
            return ms
        main = sp.value
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph):
                sp.value = torch.cuda.current_stream(dev).cuda_stream   # the capture stream
                for i in range(args.iters):
                    fn(i % n_sets)
        except Exception:
            sp.value = main
            torch.cuda.synchronize()
            return ms            # not capturable (autograd / allocator activity): keep the eager number
        finally:
            sp.value = main
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    def report(op, cfg, ms, alg_bytes, pairs, extra=None):
        gbs = alg_bytes / (ms * 1e-3) * 1e-9
        d = {"op": op, "config": cfg, "ms_per_launch": ms, "pairs_per_s": pairs / (ms * 1e-3),
             "algorithmic_bytes": alg_bytes, "achieved_gbs": gbs, "hbm_peak_gbs": hbm, "hbm_frac": gbs / hbm}
        if extra:
            d.update(extra)
        lines.append(d)
        if emit is not None:
            emit(d)

    # ---- correlation engines ------------------------------------------------------------------------
    for (B, C, H, W, P, tag) in [(4, 64, 256, 512, 192, "headline C=64 256x512 D=192"),
                                 (1, 128, 540, 960, 192, "config4 C=128 540x960 D=192"),
                                 (4, 352, 32, 64, 17, "production minidsnetExt call C=352 32x64 P=17"),
                                 (2, 64, 64, 128, 40, "config1 C=64 64x128 P=40")]:
        n_sets = 2
        S = [dict(L=torch.randn(B, C, H, W, device=dev), R=torch.randn(B, C, H, W, device=dev),
                  G=torch.randn(B, 1, P, H, W, device=dev), out=torch.empty(B, 1, P, H, W, device=dev),
                  gL=torch.empty(B, C, H, W, device=dev), gR=torch.empty(B, C, H, W, device=dev)) for _ in range(n_sets)]
        feat, vol = 4 * C * H * W, 4 * P * H * W
        fb, bb = B * (2 * feat + vol), B * (vol + 4 * feat)
        r = (P - 1) // 2
        macs = C * H * sum(max(0, W - abs(q - r)) for q in range(P))
        engines = {"auto": ("pmt_corr1d_fwd_f32", "pmt_corr1d_bwd_f32", ()),
                   "simt": ("pmt_corr1d_fwd_simt_f32", "pmt_corr1d_bwd_simt_f32", ()),
                   "tf32": ("pmt_corr1d_fwd_tc_f32", "pmt_corr1d_bwd_tc_f32", (1,))}
        for name, (ff, bf, ex) in engines.items():
            if quick and name != "auto":
                continue

            def fwd(i, ff=ff, ex=ex):
                s = S[i]
                return getattr(lib, ff)(vp(s["L"]), vp(s["R"]), vp(s["out"]), B, C, H, W, P, 1, *ex, sp)

            def bwd(i, bf=bf, ex=ex):
                s = S[i]
                return getattr(lib, bf)(vp(s["L"]), vp(s["R"]), vp(s["G"]), vp(s["gL"]), vp(s["gR"]), B, C, H, W, P, 1, *ex, sp)

            if fwd(0) != 0 or bwd(0) != 0:
                continue  # engine does not support the shape (e.g. tensor-core backward needs C <= 128)
            engine_id = lib.pmt_corr1d_uses_fast_path(vp(S[0]["L"]), vp(S[0]["R"]), vp(S[0]["G"]), C, H, W, P, 1)
            ms_f, ms_b = timed(fwd, n_sets), timed(bwd, n_sets)
            report(f"corr1d_fwd[{name}]", tag, ms_f, fb, B, {"useful_tflops": B * 2 * macs / (ms_f * 1e-3) * 1e-12,
                                                               "default_engine_id": engine_id})
            report(f"corr1d_bwd[{name}]", tag, ms_b, bb, B, {"useful_tflops": B * 4 * macs / (ms_b * 1e-3) * 1e-12})
        del S
        torch.cuda.empty_cache()

    # ---- f3: 2-D (17,17) patch of `-corrType 2dcorr` at the production feature shape; f2: corr + corrConv2d + ReLU ----
    B, C, H, W = 4, 352, 32, 64
    a2, b2 = torch.randn(B, C, H, W, device=dev), torch.randn(B, C, H, W, device=dev)
    o2 = torch.empty(B, 17, 17, H, W, device=dev)
    g2 = torch.randn(B, 17, 17, H, W, device=dev)
    ga2, gb2 = torch.empty_like(a2), torch.empty_like(b2)
    ms = timed(lambda i: lib.pmt_corr_fwd_f32(vp(a2), vp(b2), vp(o2), B, C, H, W, 17, 17, 1, 1, sp), 1)
    report("corr2d_fwd (17,17)", "2dcorr C=352 32x64 B=4", ms, 4 * (2 * B * C * H * W + o2.numel()), B,
           {"useful_tflops": 2.0 * B * C * H * W * 289 / (ms * 1e-3) * 1e-12})
    ms = timed(lambda i: lib.pmt_corr_bwd_f32(vp(a2), vp(b2), vp(g2), vp(ga2), vp(gb2), B, C, H, W, 17, 17, 1, 1, sp), 1)
    report("corr2d_bwd (17,17)", "2dcorr C=352 32x64 B=4", ms, 4 * (4 * B * C * H * W + o2.numel()), B,
           {"useful_tflops": 4.0 * B * C * H * W * 289 / (ms * 1e-3) * 1e-12})
    wt = torch.randn(128, 17, device=dev) * 0.1
    z = torch.empty(B, 128, H, W, device=dev)
    cs = torch.empty(B, 17, H, W, device=dev)
    gz = torch.randn(B, 128, H, W, device=dev)
    gw, work = torch.empty(128, 17, device=dev), torch.empty(B * H, 128 * 17, device=dev)
    ms = timed(lambda i: lib.pmt_corr1d_conv_relu_fwd_f32(vp(a2), vp(b2), vp(wt), vp(z), vp(cs), B, C, H, W, 17, 128, sp), 1)
    report("corr1d+conv1x1+relu fwd (fused f2)", "production C=352 32x64 P=17 O=128 B=4", ms, 4 * (2 * B * C * H * W + z.numel()), B)
    ms = timed(lambda i: lib.pmt_corr1d_conv_relu_bwd_f32(vp(a2), vp(b2), vp(wt), vp(z), vp(cs), vp(gz), vp(ga2), vp(gb2), vp(gw),
                                                          vp(work), B, C, H, W, 17, 128, sp), 1)
    report("corr1d+conv1x1+relu bwd (fused f2)", "production", ms, 4 * (4 * B * C * H * W + 2 * z.numel()), B)
    if not quick:
        conv = torch.nn.Sequential(torch.nn.Conv2d(17, 128, 1, bias=False), torch.nn.ReLU(inplace=True)).to(dev)
        smp = pmt.SpatialCorrelationSampler(kernel_size=1, patch_size=(1, 17), stride=1, padding=0, dilation_patch=1)
        ar, br = a2.clone().requires_grad_(True), b2.clone().requires_grad_(True)

        def unfused_f2(i):
            ar.grad = br.grad = None
            conv(torch.squeeze(smp(ar, br), 1)).backward(gz)

        ms = timed(unfused_f2, 1)
        report("sampler -> conv2d -> relu fwd+bwd [unfused: our sampler + cuDNN]", "production", ms, 0, B)
    del a2, b2, o2, g2, ga2, gb2

    # ---- PSMNet ops (config 3: B=4, 32ch 64x128 -> (64,48,64,128); cost (4,192,256,512)) ---------------
    B, C, D, H, W = 4, 32, 48, 64, 128
    ref = [torch.randn(B, C, H, W, device=dev) for _ in range(2)]
    tgt = [torch.randn(B, C, H, W, device=dev) for _ in range(2)]
    cost = [torch.empty(B, 2 * C, D, H, W, device=dev) for _ in range(2)]
    gref, gtgt = torch.empty(B, C, H, W, device=dev), torch.empty(B, C, H, W, device=dev)
    vb, fbts = 4 * B * 2 * C * D * H * W, 4 * B * 2 * C * H * W
    ms = timed(lambda i: lib.pmt_concat_volume_fwd_f32(vp(ref[i]), vp(tgt[i]), vp(cost[i]), B, C, D, H, W, 0, sp), 2)
    report("concat_volume_fwd", "config3 B=4 (32,64,128)->(64,48,64,128)", ms, vb + fbts, B)
    ms = timed(lambda i: lib.pmt_concat_volume_bwd_f32(vp(cost[i]), vp(gref), vp(gtgt), B, C, D, H, W, 0, sp), 2)
    report("concat_volume_bwd", "config3", ms, vb + fbts, B)
    del cost
    B, D, H, W = 4, 192, 256, 512
    c = [4 * torch.randn(B, D, H, W, device=dev) for _ in range(2)]
    gc = torch.empty(B, D, H, W, device=dev)
    out, lse, go = (torch.empty(B, H, W, device=dev) for _ in range(3))
    go.normal_()
    cb, ob = 4 * B * D * H * W, 4 * B * H * W
    ms = timed(lambda i: lib.pmt_softargmin_fwd_f32(vp(c[i]), vp(out), vp(lse), B, D, H, W, sp), 2)
    report("softargmin_fwd", "config3 cost (4,192,256,512)", ms, cb + 2 * ob, B)
    ms = timed(lambda i: lib.pmt_softargmin_bwd_f32(vp(c[i]), vp(out), vp(lse), vp(go), vp(gc), B, D, H, W, sp), 2)
    report("softargmin_bwd", "config3", ms, 2 * cb + 3 * ob, B)
    ms = timed(lambda i: lib.pmt_dispreg_fwd_f32(vp(c[i]), vp(out), B, D, H, W, sp), 2)
    report("dispreg_fwd", "config3", ms, cb + ob, B)
    ms = timed(lambda i: lib.pmt_dispreg_bwd_f32(vp(go), vp(gc), B, D, H, W, sp), 2)
    report("dispreg_bwd", "config3", ms, cb + ob, B)
    # the reference's own op sequence on the same GPU (softmax -> repeat ramp -> mul -> sum), for context
    ramp = torch.arange(D, device=dev, dtype=torch.float32).view(1, D, 1, 1)
    if not quick:
        ms = timed(lambda i: torch.sum(torch.softmax(c[i], dim=1) * ramp.repeat(B, 1, H, W), 1), 2)
        report("softargmin_fwd[ATen sequence of the reference]", "config3", ms, cb + 2 * ob, B)
    # f1: trilinear x4 upsample fused into the soft-argmin vs the reference sequence (upsample -> softmax -> regression)
    low = [3.0 * torch.randn(B, 1, D // 4, H // 4, W // 4, device=dev) for _ in range(2)]
    lb = 4 * B * (D // 4) * (H // 4) * (W // 4)
    ms = timed(lambda i: lib.pmt_upsample_softargmin_fwd_f32(vp(low[i]), vp(out), None, B, D // 4, H // 4, W // 4, D, H, W, sp), 2)
    report("upsample_softargmin_fwd (fused f1)", "config3 logits (4,1,48,64,128) -> pred (4,256,512)", ms, lb + ob, B)

    lse_u = torch.empty(B, H, W, device=dev)
    lib.pmt_upsample_softargmin_fwd_f32(vp(low[0]), vp(out), vp(lse_u), B, D // 4, H // 4, W // 4, D, H, W, sp)
    work = torch.empty(B, D // 4, H, W, device=dev)
    glow = torch.empty_like(low[0])
    ms = timed(lambda i: lib.pmt_upsample_softargmin_bwd_f32(vp(low[0]), vp(out), vp(lse_u), vp(go), vp(work), vp(glow), B, D // 4,
                                                             H // 4, W // 4, D, H, W, sp), 1)
    report("upsample_softargmin_bwd (fused f1)", "config3", ms, 2 * lb + 3 * ob, B)

    def unfused(i):
        up = torch.nn.functional.interpolate(low[i], size=[D, H, W], mode="trilinear", align_corners=False)[:, 0]
        return torch.sum(torch.softmax(up, dim=1) * ramp.repeat(B, 1, H, W), 1)

    if not quick:
        ms = timed(unfused, 2)
        report("upsample+softmax+regression [ATen sequence of the reference]", "config3", ms, lb + ob, B)
    lowg = low[0].clone().requires_grad_(True)

    def unfused_fb(i):
        lowg.grad = None
        up = torch.nn.functional.interpolate(lowg, size=[D, H, W], mode="trilinear", align_corners=False)[:, 0]
        torch.sum(torch.softmax(up, dim=1) * ramp.repeat(B, 1, H, W), 1).backward(go)

    if not quick:
        ms = timed(unfused_fb, 1)
        report("upsample+softmax+regression fwd+bwd [ATen autograd]", "config3", ms, 2 * lb + 3 * ob, B)
    del c, gc

    # ---- warp (config 4: 540x960, C=3; production 256x512 C=2) --------------------------------------
    for (N, C, H, W) in [(4, 3, 540, 960), (4, 2, 256, 512), (1, 128, 540, 960)]:
        img = torch.randn(N, C, H, W, device=dev)
        off = -64.0 * torch.rand(N, 1, H, W, device=dev)
        o = torch.empty(C, N, H, W, device=dev)
        g = torch.randn(C, N, H, W, device=dev)
        gi, gof = torch.zeros_like(img), torch.empty_like(off)
        ms = timed(lambda i: lib.pmt_warp1d_fwd_f32(vp(img), vp(off), vp(o), N, C, H, W, 1, sp), 1)
        report("warp1d_fwd", f"N={N} C={C} {H}x{W}", ms, 4 * N * H * W * (2 * C + 1), N)

        def wb(i):
            return lib.pmt_warp1d_bwd_f32(vp(img), vp(off), vp(g), vp(gi), vp(gof), N, C, H, W, 1, sp)

        ms = timed(wb, 1)
        report("warp1d_bwd (deterministic row gather)", f"N={N} C={C} {H}x{W}", ms, 4 * N * H * W * (3 * C + 2), N)
        if C <= 3:
            seg, att = torch.randn(N, C, H, W, device=dev), torch.rand(N, 1, H, W, device=dev)
            ob, ow = torch.empty(N, C, H, W, device=dev), torch.empty(N, C, H, W, device=dev)
            gb_, gs_, ga_ = torch.randn(N, C, H, W, device=dev), torch.empty(N, C, H, W, device=dev), torch.empty_like(att)
            ms = timed(lambda i: lib.pmt_warp1d_blend_fwd_f32(vp(img), vp(off), vp(att), vp(seg), vp(ob), vp(ow), N, C, H, W, sp), 1)
            report("warp+blend fwd (fused f4)", f"N={N} C={C} {H}x{W}", ms, 4 * N * H * W * (4 * C + 2), N)
            ms = timed(lambda i: lib.pmt_warp1d_blend_bwd_f32(vp(img), vp(off), vp(att), vp(seg), vp(gb_), None, vp(gi), vp(gof),
                                                              vp(ga_), vp(gs_), N, C, H, W, sp), 1)
            report("warp+blend bwd (fused f4)", f"N={N} C={C} {H}x{W}", ms, 4 * N * H * W * (5 * C + 4), N)

    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--out", default="")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lines = run_ops(args.iters, 0, lambda d: print(json.dumps(d), flush=True), args.quick)
    if args.out:
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
