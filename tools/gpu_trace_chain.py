"""Per-CTA timeline of the persistent chain kernels (debug hook b200ode_debug_set_trace).  Not a pytest.
usage: python tools/gpu_trace_chain.py N H W C [L] [fast_tf32|fast_f16]"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import ChainHandle

NAMES = {1: "setup", 2: "mma:L4 start", 3: "mma:L4 w_full|wstall", 4: "mma:L4 issued", 5: "mma:L5 start", 6: "epi:L4 seg0 rdy",
         7: "epi:L4 seg1 rdy", 8: "epi:L4 done", 9: "epi:L5 seg0 rdy", 10: "mma:end", 11: "epi:end", 12: "end", 13: "tma:L5 first issue", 14: "tma:L6 first issue"}


def main():
    N, H, W, C = [int(v) for v in sys.argv[1:5]]
    L = int(sys.argv[5]) if len(sys.argv) > 5 else 36
    prec = _abi.CHAIN_PRECISIONS[sys.argv[6] if len(sys.argv) > 6 else "fast_tf32"]
    lib = _abi.lib()
    ch = ChainHandle(C, L, 0.0, precision=prec)
    sdt = ch.saved_dtype
    params = torch.randn(L * ch.num_params, device="cuda") * 0.05
    ch.pack(params)
    x0 = torch.relu(torch.randn((N, H, W, C), device="cuda")); dy = torch.randn((N, H, W, C), device="cuda")
    acts = torch.empty((L, N, H, W, C), device="cuda", dtype=sdt); dz = torch.empty((L, N, H, W, C), device="cuda", dtype=sdt)
    yfin = torch.empty_like(x0)
    masks = torch.empty((L, N, H, W, C // 8), dtype=torch.uint8, device="cuda"); dx = torch.empty_like(x0)
    grad = torch.empty(L * ch.num_params, device="cuda")
    tr = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
    for kind in ("fwd", "dgrad", "wgrad"):
        for rep in range(3):
            tr.zero_()
            lib.b200ode_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
            if kind == "fwd": ch.forward(x0, 0.07, acts=acts, masks=masks, y_final=yfin if ch.f16 else None)
            elif kind == "dgrad": ch.dgrad(dy, masks, dz, dx, 0.07)
            else: ch.wgrad(x0, acts, dz, grad)
            torch.cuda.synchronize()
            lib.b200ode_debug_set_trace(None)
        t = tr.cpu().view(-1, 16)
        t = t[t[:, 15] != 0]
        w0, w1 = t[:, 0].min().item(), t[:, 15].max().item()
        print("%s %s L=%d: %d CTAs, wall span %.2f us" % (kind, (N, H, W, C), L, t.shape[0], (w1 - w0) / 1e3))
        if kind == "wgrad":
            nm = {1: "setup", 2: "mma:full0", 3: "mma:tile0 issued", 4: "mma:all issued", 5: "epi:bias done", 6: "epi:acc_full", 7: "epi:done", 9: "end", 10: "mma wait on stages"}
        elif ch.f16:
            nm = {1: "setup", 2: "mma:L4 start", 4: "mma:L4 issued", 5: "mma:L5 start", 6: "epi(w2):L4 first acc ready",
                  8: "epi(w2):L4 done", 9: "epi(w2):L5 first acc ready", 10: "mma:end", 11: "epi:end", 12: "end"}
        else:
            nm = NAMES
        for cta in (0, t.shape[0] // 2, t.shape[0] - 1):
            row = t[cta]
            print("  cta %3d: " % cta + "  ".join("%s=%d" % (nm[i], row[i].item()) for i in sorted(nm)) + "  wall=%.2fus" % ((row[15].item() - row[0].item()) / 1e3))
        med = t.float().median(dim=0).values
        print("  median : " + "  ".join("%s=%d" % (nm[i], med[i].item()) for i in sorted(nm)))


if __name__ == "__main__":
    main()
