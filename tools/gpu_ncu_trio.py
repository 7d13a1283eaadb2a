"""One forward, data-gradient and weight-gradient launch of a single Euler layer (for `ncu --set full`).  Not a pytest.
usage: python tools/gpu_ncu_trio.py N H W C precision"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr
N, H, W, C = [int(v) for v in sys.argv[1:5]]; prec = sys.argv[5]
lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
hd = LayerHandle(C, 3, -0.1, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
_abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(torch.randn(hd.num_params, device="cuda") * 0.05), None, st))
x = torch.randn((N, H, W, C), device="cuda").to(dt); y = torch.empty_like(x); y2 = torch.empty_like(x)
m = torch.empty((N, H, W, C // 8), dtype=torch.uint8, device="cuda"); g = torch.empty(hd.num_params, device="cuda")
for i in range(2):
    _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(x), _ptr(y), _ptr(m), None, N, H, W, 0.1, 15, st))
    _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(x), _ptr(y), _ptr(y2), N, H, W, st))
    _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(x), _ptr(y), _ptr(g), None, N, H, W, 0, st))
torch.cuda.synchronize()
