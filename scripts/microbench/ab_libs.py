"""A/B of two builds of libpmt_ops.so on one GPU box: same inputs, results compared bit for bit, burst and sustained timing
of the headline correlation (and config 4).  usage: ab_libs.py base.so new.so"""
import ctypes, sys, time
import torch

dev = torch.device("cuda:0")
vp = lambda t: ctypes.c_void_p(t.data_ptr())
I, P_ = ctypes.c_int, ctypes.c_void_p


def load(path):
    lib = ctypes.CDLL(path)
    lib.pmt_corr1d_fwd_f32.argtypes = [P_, P_, P_, I, I, I, I, I, I, P_]
    lib.pmt_corr1d_bwd_f32.argtypes = [P_, P_, P_, P_, P_, I, I, I, I, I, I, P_]
    lib.pmt_corr1d_fwd_tc_f32.argtypes = [P_, P_, P_, I, I, I, I, I, I, I, P_]
    lib.pmt_corr1d_bwd_tc_f32.argtypes = [P_, P_, P_, P_, P_, I, I, I, I, I, I, I, P_]
    lib.pmt_last_error.restype = ctypes.c_char_p
    return lib


libs = [(p, load(p)) for p in sys.argv[1:]]
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for (B, C, H, W, P) in [(4, 64, 256, 512, 192), (2, 128, 540, 960, 192), (4, 64, 64, 128, 40)]:
    g = torch.Generator(device=dev).manual_seed(1)
    L = torch.randn(B, C, H, W, device=dev, generator=g); R = torch.randn(B, C, H, W, device=dev, generator=g)
    G = torch.randn(B, 1, P, H, W, device=dev, generator=g)
    res = {}
    for passes in (3, 1):
        for name, lib in libs:
            out = torch.empty(B, 1, P, H, W, device=dev); g1 = torch.empty_like(L); g2 = torch.empty_like(L)
            f = lambda: lib.pmt_corr1d_fwd_tc_f32(vp(L), vp(R), vp(out), B, C, H, W, P, 1, passes, st)
            b = lambda: lib.pmt_corr1d_bwd_tc_f32(vp(L), vp(R), vp(G), vp(g1), vp(g2), B, C, H, W, P, 1, passes, st)
            assert f() == 0, lib.pmt_last_error()
            assert b() == 0, lib.pmt_last_error()
            torch.cuda.synchronize()
            key = passes
            if key in res:
                same = [bool(torch.equal(a_, b_)) for a_, b_ in zip(res[key], (out, g1, g2))]
                md = [float((a_ - b_).abs().max()) for a_, b_ in zip(res[key], (out, g1, g2))]
            else:
                res[key] = (out.clone(), g1.clone(), g2.clone()); same, md = None, None
            time.sleep(0.5)
            for _ in range(3): f(); b()
            tf, tb = timed(f, 20), timed(b, 20)
            # sustained: 1.5 s of fwd+bwd first, then time each kernel inside an alternating loop
            t_end = time.time() + 1.5
            while time.time() < t_end:
                for _ in range(50): f(); b()
                torch.cuda.synchronize()
            n = 100
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * n + 1)]
            ev[0].record()
            for i in range(n):
                f(); ev[2 * i + 1].record(); b(); ev[2 * i + 2].record()
            torch.cuda.synchronize()
            sf = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(n)) / n * 1e3
            sb = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(n)) / n * 1e3
            print(f"B{B} C{C} {H}x{W} P{P} passes={passes} {name.split('/')[-1]:24s} burst fwd {tf:7.1f} bwd {tb:7.1f} us | sustained fwd {sf:7.1f} bwd {sb:7.1f} us | same_as_base={same} maxdiff={md}", flush=True)
    del L, R, G
