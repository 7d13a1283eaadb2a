"""cfg2 sweep (BASELINE.json configs[1]): one Euler-step layer, channels 16-256, fwd / dgrad / wgrad in
every tensor-core precision mode.  Not a pytest; prints one line per (kernel, shape, mode) and writes
JSON lines to the path given by --out.

usage: python tools/gpu_sweep.py [--batch 256] [--hw 32] [--channels 16,32,64,128,256]
                                 [--modes fast_tf32,fast_bf16,strict] [--out gpurun_out/sweep.jsonl]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from differential_equations_resnet_b200 import _abi  # noqa: E402
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr  # noqa: E402


def bench_layer(N, H, W, C, prec, min_bytes=400e6, reps=3, quiet=False):
    lib = _abi.lib()
    st = torch.cuda.current_stream().cuda_stream
    hd = LayerHandle(C, 3, -0.1, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
    dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
    eb = 2 if prec == "fast_bf16" else 4
    params = torch.randn(hd.num_params, device="cuda") * 0.05
    _abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(params), None, st))
    per = N * H * W * C * eb
    nbuf = max(3, int(min_bytes // (2 * per)) + 1)
    xs = [torch.randn((N, H, W, C), device="cuda").to(dt) for _ in range(nbuf)]
    ys = [torch.empty((N, H, W, C), device="cuda", dtype=dt) for _ in range(nbuf)]
    ms = [torch.empty((N, H, W, C // 8), dtype=torch.uint8, device="cuda") for _ in range(nbuf)]
    g = torch.empty(hd.num_params, device="cuda")
    Mpix = N * H * W

    def f_fwd(i):
        _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(xs[i]), _ptr(ys[i]), _ptr(ms[i]), None, N, H, W, 0.1, 15, st))

    def f_dgrad(i):
        _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(xs[i]), _ptr(ys[(i + 1) % nbuf]), _ptr(ys[i]), N, H, W, st))

    def f_wgrad(i):
        _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(xs[i]), _ptr(ys[i]), _ptr(g), None, N, H, W, 0, st))

    out = []
    iters = max(nbuf, 10)
    for name, fn, nbytes in (("fwd", f_fwd, 2 * per + Mpix * C // 8), ("dgrad", f_dgrad, 3 * per),
                             ("wgrad", f_wgrad, 2 * per)):
        if name == "wgrad" and prec == "fast_bf16":
            # bias gradient of the bf16 path is a separate column sum (include/b200ode.h)
            pass
        try:
            for i in range(min(nbuf, 4)):
                fn(i)
        except Exception as e:  # unsupported combination: report, keep going
            if not quiet:
                print("%-6s %s %-9s unsupported: %s" % (name, (N, H, W, C), prec, e), flush=True)
            continue
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(iters):
                fn(i % nbuf)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
        fl = 2.0 * Mpix * 9 * C * C
        rec = {"kernel": name, "shape": [N, H, W, C], "mode": prec, "us": best, "alg_bytes": nbytes,
               "GBps": nbytes / best * 1e-3, "alg_TFLOPs": fl / best * 1e-6}
        out.append(rec)
        if not quiet:
            print("%-6s %s %-9s %9.1f us  %7.1f GB/s (alg)  %7.1f TFLOP/s (alg)" %
                  (name, (N, H, W, C), prec, best, rec["GBps"], rec["alg_TFLOPs"]), flush=True)
    del xs, ys, ms
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--channels", default="16,32,64,128,256")
    ap.add_argument("--modes", default="fast_tf32,fast_bf16,strict")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    recs = []
    for C in [int(c) for c in a.channels.split(",")]:
        for m in a.modes.split(","):
            recs += bench_layer(a.batch, a.hw, a.hw, C, m)
    if a.out:
        with open(a.out, "w") as f:
            for r in recs:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
