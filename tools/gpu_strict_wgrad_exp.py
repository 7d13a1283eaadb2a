"""Timing of the strict (3xTF32) layer-batched weight gradient at the cfg3 stage shapes, for A/B runs with the debug switches
B200ODE_WGRAD_DBG=1 (no MMAs: staging + converter warps + epilogue only), B200ODE_WGRAD_NOSTACK=1 (three MMAs per entry).  Not a pytest.
usage: python tools/gpu_strict_wgrad_exp.py [precision=strict]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import ChainHandle

prec = sys.argv[1] if len(sys.argv) > 1 else "strict"
N, L = 128, 36
tag = " ".join("%s=%s" % (k, os.environ[k]) for k in ("B200ODE_WGRAD_DBG", "B200ODE_WGRAD_NOSTACK") if k in os.environ) or "default"
for (C, H) in ((16, 32), (32, 16), (64, 8)):
    ch = ChainHandle(C, L, 0.0, precision=_abi.CHAIN_PRECISIONS[prec])
    dt = ch.saved_dtype
    x0 = torch.randn((N, H, H, C), device="cuda")
    acts = torch.randn((L, N, H, H, C), device="cuda").to(dt)
    dz = (torch.randn((L, N, H, H, C), device="cuda") * 1e-2).to(dt)
    grad = torch.empty((L, ch.num_params), device="cuda")
    ws = torch.empty(ch.workspace_bytes(N, H, H), dtype=torch.uint8, device="cuda")
    ch.bind_workspace(ws)
    for _ in range(3):
        ch.wgrad(x0, acts, dz, grad.view(-1))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ch.wgrad(x0, acts, dz, grad.view(-1))
    e1.record(); torch.cuda.synchronize()
    print("%-8s [%s] wgrad+fold 36 layers (%d,%d,%d,%d): %8.1f us" % (prec, tag, N, H, H, C, e0.elapsed_time(e1) * 100), flush=True)
