"""Per-role timeline of the tensor-core weight-gradient kernel (debug hook b200ode_debug_set_trace).  Not a pytest.
usage: python tools/gpu_trace_wgrad.py N H W C precision"""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle, _ptr

N, H, W, C = [int(v) for v in sys.argv[1:5]]
prec = sys.argv[5] if len(sys.argv) > 5 else "fast_bf16"
lib = _abi.lib(); st = torch.cuda.current_stream().cuda_stream
hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PRECISIONS[prec], _abi.LAYOUT_3BY3)
dt = torch.bfloat16 if prec == "fast_bf16" else torch.float32
_abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(torch.randn(hd.num_params, device="cuda") * 0.05), None, st))
x = torch.randn((N, H, W, C), device="cuda").to(dt); dz = torch.randn((N, H, W, C), device="cuda").to(dt)
g = torch.empty(hd.num_params, device="cuda")
tr = torch.zeros(1024 * 16, dtype=torch.int64, device="cuda")
for rep in range(3):
    tr.zero_()
    lib.b200ode_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
    _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(x), _ptr(dz), _ptr(g), None, N, H, W, 0, st))
    torch.cuda.synchronize()
    lib.b200ode_debug_set_trace(None)
t = tr.cpu().view(-1, 16)
t = t[t[:, 15] != 0]
n = t.shape[0]
w0 = t[:, 0].min().item()
print("wgrad %s %s: %d CTAs, wall span %.1f us, CTA start spread %.1f us" % ((N, H, W, C), prec, n, (t[:, 15].max().item() - w0) / 1e3, (t[:, 0].max().item() - w0) / 1e3))
names = {1: "setup", 2: "mma:first stage", 3: "mma:tile0 issued", 4: "mma:all issued", 5: "epi:bias done", 6: "epi:acc_full", 7: "epi:done", 9: "end", 10: "mma wait on stages"}
wall = (t[:, 15] - t[:, 0]).float() / 1e3
bias = (t[:, 5] - t[:, 1]) * 4 > t[:, 9]   # the other CTAs pass mark 5 right after setup
for label, sel in (("bias CTAs", bias), ("other CTAs", ~bias)):
    if sel.sum() == 0: continue
    tt = t[sel].float()
    print("  %s (%d): wall med %.1f max %.1f us; " % (label, int(sel.sum()), wall[sel].median().item(), wall[sel].max().item())
          + "  ".join("%s=%d" % (names[i], tt[:, i].median().item()) for i in sorted(names)))
