"""The persistent chain kernels (forward sweep, backward sweep) and the layer-batched weight gradient at the three cfg3
stage shapes, one warm and one measured launch each (for `ncu --set full -k regex:chain_f16|wgrad_tc`).  Not a pytest.
usage: python tools/gpu_ncu_chain.py [precision=fast_f16] [batch=128]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import ChainHandle

prec = sys.argv[1] if len(sys.argv) > 1 else "fast_f16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
L, h = 36, 2.0 / 108
for C, HW in ((16, 32), (32, 16), (64, 8)):
    ch = ChainHandle(C, L, 0.0, precision=_abi.CHAIN_PRECISIONS[prec])
    ch.pack(torch.randn(L * ch.num_params, device="cuda") * 0.05)
    shape = (N, HW, HW, C)
    x0, dy = torch.relu(torch.randn(shape, device="cuda")), torch.randn(shape, device="cuda")
    acts = torch.empty((L,) + shape, device="cuda", dtype=ch.saved_dtype)
    masks = torch.empty((L, N, HW, HW, C // 8), dtype=torch.uint8, device="cuda")
    dz = torch.empty((L,) + shape, device="cuda", dtype=ch.saved_dtype)
    dx, y = torch.empty(shape, device="cuda"), torch.empty(shape, device="cuda")
    grad = torch.empty(L * ch.num_params, device="cuda")
    for _ in range(2):
        ch.forward(x0, h, acts=acts, masks=masks, y_final=y if ch.f16 else None)
        ch.dgrad(dy, masks, dz, dx, h)
        ch.wgrad(x0, acts, dz, grad)
    torch.cuda.synchronize()
    print("C=%d %dx%d done" % (C, HW, HW), flush=True)
