"""HBM-roofline numbers of the memory-bound kernels (K1, K5, K6 of SURVEY.md section 7 / north_star item 3): achieved
algorithmic GB/s against the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs), CUDA events, tensors larger than the
126 MB L2 (M = 262144 pixels x 256 channels fp32 = 268 MB each) and at a cfg1 size for comparison.  Not a pytest.
usage: python tools/gpu_tails_bench.py > profiles/rNN_tails_hbm.log"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from differential_equations_resnet_b200 import _abi

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6456.5
P = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())


def timeit(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    gbs = nbytes / us / 1e3
    print("%-44s %9.1f us  %8.1f MB  %7.1f GB/s  %5.1f %% of %.0f" % (name, us, nbytes / 1e6, gbs, 100 * gbs / PEAK, PEAK), flush=True)


def run(M, C, tag):
    lib, st = _abi.lib(), None
    n = M * C
    z, x, y, dy, dz = (torch.randn(n, device="cuda") for _ in range(5))
    mask = torch.empty(M * C // 8, dtype=torch.uint8, device="cuda")
    scale, shift, mean, inv, gam, bet, dg, db, s0, s1 = (torch.rand(C, device="cuda") + 0.5 for _ in range(10))
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    ws = torch.empty(2 * _abi.COLSUM_PARTS * C, device="cuda")
    print("--- %s: M = %d pixels, C = %d (%.1f MB per fp32 tensor)" % (tag, M, C, n * 4 / 1e6))
    timeit("colsum (sum z, sum z^2)", lambda: _abi.check(lib.b200ode_colsum(P(z), None, P(s0), P(s1), P(ws), M, C, st)), n * 4)
    timeit("euler_tail BN (y = x + h relu(z s + t))", lambda: _abi.check(lib.b200ode_euler_tail(P(z), P(scale), P(shift), P(x), P(y), None, M, C, 0.1, 2 | 4 | 8, st)), 3 * n * 4)
    timeit("euler_tail + relu mask", lambda: _abi.check(lib.b200ode_euler_tail(P(z), None, None, P(x), P(y), P(mask), M, C, 0.1, 2 | 4 | 8, st)), 3 * n * 4 + n // 8)
    timeit("bn_bwd_reduce (sum du, sum du zhat)", lambda: _abi.check(lib.b200ode_bn_bwd_reduce(P(dy), P(z), P(scale), P(shift), P(mean), P(inv), P(dg), P(db), P(ws), M, C, 0.1, st)), 2 * n * 4)
    timeit("bn_bwd_apply (dz)", lambda: _abi.check(lib.b200ode_bn_bwd_apply(P(dy), P(z), P(scale), P(shift), P(mean), P(inv), P(gam), P(dg), P(db), P(dz), M, C, 0.1, st)), 3 * n * 4)
    timeit("relu_scale_bwd (dz = h dy mask)", lambda: _abi.check(lib.b200ode_relu_scale_bwd(P(dy), P(mask), P(dz), M, C, 0.1, 0, st)), 2 * n * 4 + n // 8)
    cnt = torch.ones(1, dtype=torch.int32, device="cuda")
    timeit("adam_step (theta, g, m, v)", lambda: _abi.check(lib.b200ode_adam_step(P(x), P(dy), P(y), P(dz.abs_()), n, P(cnt), 1e-3, 0.9, 0.999, 1e-7, 1.0, st)), 7 * n * 4)
    timeit("torch copy_ (reference point)", lambda: y.copy_(x), 2 * n * 4)


if __name__ == "__main__":
    torch.manual_seed(0)
    run(262144, 256, "cfg2 C=256 (N=256, 32x32)")
    run(262144, 64, "cfg2 C=64")
    run(131072, 16, "cfg1 stage 1 (N=128, 32x32x16): L2 resident")
    from differential_equations_resnet_b200.layers._base import LayerHandle
    # pack kernel K1: reads the packed parameters, writes dense + staged copies
    for C in (64, 256):
        hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PREC_FAST_BF16, _abi.LAYOUT_3BY3)
        th = torch.randn(hd.num_params, device="cuda")
        timeit("pack_kernel C=%d (bf16 staging)" % C, lambda: _abi.check(_abi.lib().b200ode_pack_kernel(hd._h, P(th), None, None)),
               hd.num_params * 4 + 9 * C * C * (4 + 2))
