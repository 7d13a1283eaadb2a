"""CUDA-event timing of the stem / transition / head / optimiser kernels at the cfg3 shapes (batch 128).  Not a pytest.
usage: python tools/gpu_glue_bench.py [once]      ('once': a single launch of each, for ncu)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi

lib = _abi.lib()
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
once = len(sys.argv) > 1 and sys.argv[1] == "once"
N = 128
dev = "cuda"


def timeit(name, fn, nbytes, flops):
    if once:
        fn(); torch.cuda.synchronize(); return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # rotate nothing: inputs are 4-17 MB, but successive launches of the step interleave >1 GB of chain traffic; report both
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print("%-34s %8.1f us   %7.1f GB/s (algorithmic bytes)   %6.2f TFLOP/s" % (name, us, nbytes / us * 1e-3, flops / us * 1e-6), flush=True)


g = torch.Generator(device=dev).manual_seed(0)
img = torch.randint(0, 256, (N, 32, 32, 3), dtype=torch.uint8, device=dev, generator=g)
K1 = torch.randn((3, 3, 3, 16), device=dev, generator=g) * 0.2; b1 = torch.zeros(16, device=dev)
out1 = torch.empty((N, 32, 32, 16), device=dev); d1 = torch.randn((N, 32, 32, 16), device=dev, generator=g)
dp1 = torch.empty(27 * 16 + 16, device=dev)
timeit("stem_fwd 3->16 @32x32", lambda: _abi.check(lib.b200ode_stem_fwd(P(img), 1, 127.5, 127.5, 1, P(K1), P(b1), P(out1), N, 32, 32, 3, 16, None)),
       img.numel() + out1.numel() * 4, 2.0 * N * 1024 * 27 * 16)
timeit("stem_wgrad", lambda: _abi.check(lib.b200ode_stem_wgrad(P(img), 1, 127.5, 127.5, 1, P(out1), P(d1), P(dp1), N, 32, 32, 3, 16, None, 0, None)),
       img.numel() + 2 * out1.numel() * 4, 2.0 * N * 1024 * 27 * 16)
for (Ci, Co, H) in ((16, 32, 32), (32, 64, 16)):
    x = torch.randn((N, H, H, Ci), device=dev, generator=g)
    Km = torch.randn((3, 3, Ci, Co), device=dev, generator=g) * 0.1; bm = torch.zeros(Co, device=dev)
    Ks = torch.randn((1, 1, Ci, Co), device=dev, generator=g) * 0.1; bs = torch.zeros(Co, device=dev)
    Ho = H // 2
    out = torch.empty((N, Ho, Ho, Co), device=dev); mask = torch.empty((N, Ho, Ho, Co // 8), dtype=torch.uint8, device=dev)
    dout = torch.randn((N, Ho, Ho, Co), device=dev, generator=g); dx = torch.empty_like(x)
    dp = torch.empty(9 * Ci * Co + Co + Ci * Co + Co, device=dev)
    fl = 2.0 * N * Ho * Ho * Co * 10 * Ci
    tag = "%d->%d @%dx%d" % (Ci, Co, H, H)
    timeit("transition_fwd " + tag, lambda: _abi.check(lib.b200ode_transition_fwd(P(x), P(Km), P(bm), P(Ks), P(bs), P(out), P(mask), N, H, H, Ci, Co, 2, 2, None)),
           (x.numel() + out.numel()) * 4, fl)
    timeit("transition_dgrad " + tag, lambda: _abi.check(lib.b200ode_transition_dgrad(P(dout), P(mask), P(Km), P(Ks), P(dx), N, H, H, Ci, Co, 2, 2, None)),
           (x.numel() + out.numel()) * 4, fl)
    timeit("transition_wgrad " + tag, lambda: _abi.check(lib.b200ode_transition_wgrad(P(x), P(dout), P(mask), P(dp), N, H, H, Ci, Co, 2, 2, None, 0, None)),
           (x.numel() + out.numel()) * 4, fl)
xh = torch.randn((N, 8, 8, 64), device=dev, generator=g); Wf = torch.randn((64, 10), device=dev, generator=g) * 0.1; bf = torch.zeros(10, device=dev)
oh = torch.nn.functional.one_hot(torch.randint(0, 10, (N,), device=dev), 10).float()
loss = torch.zeros(1, device=dev); dxh = torch.empty_like(xh); dph = torch.empty(64 * 10 + 10, device=dev)
timeit("head fwd+bwd (GAP, FC, CE)", lambda: _abi.check(lib.b200ode_head_fwd_bwd(P(xh), P(Wf), P(bf), P(oh), 1e-7, None, P(loss), P(dxh), P(dph), N, 64, 64, 10, None, 0, None)),
       2 * xh.numel() * 4, 0)
n = 899_866
th, gr, m, v = (torch.randn(n, device=dev) for _ in range(4)); v.abs_()
cnt = torch.ones(1, dtype=torch.int32, device=dev)
timeit("adam (%d params)" % n, lambda: _abi.check(lib.b200ode_adam_step(P(th), P(gr), P(m), P(v), n, P(cnt), 1e-3, 0.9, 0.999, 1e-7, 1.0, None)), 7 * n * 4, 0)
