"""Per-CTA timeline of the tensor-core kernels (debug hook b200ode_debug_set_trace).  Not a pytest.
usage: python tools/gpu_trace.py N H W C precision"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr


def main():
    N, H, W, C = [int(v) for v in sys.argv[1:5]]
    prec = sys.argv[5] if len(sys.argv) > 5 else "fast_tf32"
    lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
    hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
    dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
    params = torch.randn(hd.num_params, device="cuda") * 0.05
    _abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(params), None, st))
    x = torch.randn((N, H, W, C), device="cuda").to(dt)
    y = torch.empty_like(x); y2 = torch.empty_like(x)
    m = torch.empty((N, H, W, C // 8), dtype=torch.uint8, device="cuda")
    g = torch.empty(hd.num_params, device="cuda")
    tr = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
    names = {"fwd": ["wall0", "setup", "mma:a_full", "mma:w_full", "mma:tile0 issued", "mma:all issued", "epi:acc_full",
                     "epi:tile0 done", "epi:all done", "end"],
             "wgrad": ["wall0", "setup", "mma:full0", "mma:tile0 issued", "mma:all issued", "epi:bias done", "epi:acc_full",
                       "epi:done", "-", "end"]}
    def run(kind):
        for rep in range(3):
            tr.zero_()
            lib.b200ode_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
            if kind == "fwd":
                _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(x), _ptr(y), _ptr(m), None, N, H, W, 0.1, 15, st))
            elif kind == "dgrad":
                _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(x), _ptr(y), _ptr(y2), N, H, W, st))
            else:
                _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(x), _ptr(y), _ptr(g), None, N, H, W, 0, st))
            torch.cuda.synchronize()
            lib.b200ode_debug_set_trace(None)
        t = tr.cpu().view(-1, 16)
        used = t[:, 15] != 0
        t = t[used]
        w0, w1 = t[:, 0].min().item(), t[:, 15].max().item()
        print("%s %s %s: %d CTAs, wall span %.2f us (first CTA start -> last CTA end); CTA start spread %.2f us" %
              (kind, (N, H, W, C), prec, t.shape[0], (w1 - w0) / 1e3, (t[:, 0].max().item() - w0) / 1e3))
        nm = names["wgrad" if kind == "wgrad" else "fwd"]
        for cta in (0, t.shape[0] // 2, t.shape[0] - 1):
            row = t[cta]
            print("  cta %3d: " % cta + "  ".join("%s=%d" % (nm[i], row[i].item()) for i in range(1, 10) if nm[i] != "-")
                  + "  wall=%.2fus" % ((row[15].item() - row[0].item()) / 1e3))
        med = t[:, 1:10].float().median(dim=0).values
        print("  median : " + "  ".join("%s=%d" % (nm[i], med[i - 1].item()) for i in range(1, 10) if nm[i] != "-"))
        if kind != "wgrad":
            w = t[:, 10:13].float().median(dim=0).values
            print("  MMA warp waits (cycles, median): strips %d  weights %d  free accumulators %d" % (w[0].item(), w[1].item(), w[2].item()))
    for kind in ("fwd", "dgrad", "wgrad"):
        run(kind)


if __name__ == "__main__":
    main()
