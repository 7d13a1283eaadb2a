"""Debug: time the per-layer wgrad kernel for a given shape/mode (tiling overrides via B200ODE_WGRAD_TG/NT). Not a pytest."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr
N, H, W, C = [int(v) for v in sys.argv[1:5]]
prec = sys.argv[5]
lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
hd = LayerHandle(C, 3, -0.1, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
_abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(torch.randn(hd.num_params, device="cuda") * 0.05), None, st))
nb = 4
xs = [torch.randn((N, H, W, C), device="cuda").to(dt) for _ in range(nb)]
ys = [torch.randn((N, H, W, C), device="cuda").to(dt) for _ in range(nb)]
g = torch.empty(hd.num_params, device="cuda")
def f(i): _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(xs[i]), _ptr(ys[i]), _ptr(g), None, N, H, W, 0, st))
for i in range(nb): f(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(3 * nb): f(i % nb)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (3 * nb)
print("wgrad %s %s TG=%s NT=%s: %.1f us  %.1f TFLOP/s" % ((N, H, W, C), prec, os.environ.get("B200ODE_WGRAD_TG", "auto"), os.environ.get("B200ODE_WGRAD_NT", "auto"), us, 2.0 * N * H * W * 9 * C * C / us * 1e-6))
