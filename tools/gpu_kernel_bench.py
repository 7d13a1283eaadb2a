"""Micro-benchmark / profiling driver for the three tensor-core kernels (not a pytest).
usage: python tools/gpu_kernel_bench.py N H W C precision [iters]"""
import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr

def main():
    N, H, W, C = [int(v) for v in sys.argv[1:5]]
    prec = sys.argv[5] if len(sys.argv) > 5 else "fast_tf32"
    iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
    lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
    hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
    dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
    eb = 2 if prec == "fast_bf16" else 4
    params = torch.randn(hd.num_params, device="cuda") * 0.05
    _abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(params), None, st))
    per = N * H * W * C * eb
    nbuf = max(3, int(300e6 // per) + 1)
    xs = [torch.randn((N, H, W, C), device="cuda").to(dt) for _ in range(nbuf)]
    ys = [torch.empty((N, H, W, C), device="cuda", dtype=dt) for _ in range(nbuf)]
    ms = [torch.empty((N, H, W, C // 8), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    g = torch.empty(hd.num_params, device="cuda"); G = torch.empty((3, 3, C, C), device="cuda")
    def f_fwd(i): _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(xs[i]), _ptr(ys[i]), _ptr(ms[i]), None, N, H, W, 0.1, 15, st))
    def f_dgrad(i): _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(xs[i]), _ptr(ys[(i + 1) % nbuf]), _ptr(ys[i]), N, H, W, st))
    def f_wgrad(i):
        rc = lib.b200ode_euler_wgrad(hd._h, _ptr(xs[i]), _ptr(ys[i]), _ptr(g), _ptr(G), N, H, W, 0, st)
        if rc not in (0, -2): _abi.check(rc)
    Mpix = N * H * W
    for name, fn, nbytes in (("fwd", f_fwd, 2 * per + Mpix * C // 8), ("dgrad", f_dgrad, 3 * per), ("wgrad", f_wgrad, 2 * per)):
        for i in range(min(nbuf, 5)): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): fn(i % nbuf)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        fl = 2.0 * Mpix * 9 * C * C
        print("%-6s %s %-9s %8.1f us  %7.1f GB/s (alg)  %7.1f TFLOP/s (alg)" % (name, (N, H, W, C), prec, us, nbytes / us * 1e-3, fl / us * 1e-6), flush=True)

if __name__ == "__main__":
    main()
