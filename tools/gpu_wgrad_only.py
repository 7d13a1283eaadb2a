import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr
N, H, W, C = [int(v) for v in sys.argv[1:5]]; prec = sys.argv[5]
lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
params = torch.randn(hd.num_params, device="cuda") * 0.05
_abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(params), None, st))
x = torch.randn((N, H, W, C), device="cuda"); dz = torch.randn((N, H, W, C), device="cuda")
g = torch.empty(hd.num_params, device="cuda")
for i in range(3):
    _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(x), _ptr(dz), _ptr(g), None, N, H, W, 0, st))
torch.cuda.synchronize()
