"""Smallest invocation of every hand-rolled mbarrier / TMEM / TMA protocol (per-layer conv fwd, dgrad, fused dgrad, wgrad in
every precision; tf32 and fp16 chains forward, backward sweep, layer-batched wgrad; glue kernels; Adam), meant to run under
compute-sanitizer, ONE tool per run (SURVEY.md section 5 "race detection"):
    compute-sanitizer --tool memcheck  python tools/gpu_sanitize.py
    compute-sanitizer --tool racecheck python tools/gpu_sanitize.py
    compute-sanitizer --tool synccheck python tools/gpu_sanitize.py
Not a pytest.  Prints one line per case; results are checked against each other (chain = per-layer) only loosely: the point
here is the sanitizer's report, the numerics are covered by tests/."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import differential_equations_resnet_b200 as pkg
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import ChainHandle
from differential_equations_resnet_b200.training import EulerNet, NetSpec


def layer_case(precision, shape):
    N, H, W, C = shape
    layer = pkg.Conv2DAntisymmetric3By3(gamma=-0.1, precision=precision, seed=0)
    dt = torch.bfloat16 if precision == "fast_bf16" else torch.float32
    x = torch.relu(torch.randn(shape)).to(dt).cuda().requires_grad_(True)
    y = layer.euler_step(x, 0.125)
    y.backward(torch.randn(shape).to(dt).cuda())
    torch.cuda.synchronize()
    print("layer %-9s %-18s |y| %.4f |dx| %.4f |g| %.4f" % (precision, shape, float(y.float().norm()), float(x.grad.float().norm()),
                                                           float(layer.packed.grad.norm())), flush=True)


def chain_case(prec_name, C, HW, L=2, N=2):
    prec = _abi.CHAIN_PRECISIONS[prec_name]
    ch = ChainHandle(C, L, -0.1, precision=prec)
    theta = (torch.randn(L, ch.num_params) * 0.1).cuda()
    ch.pack(theta.view(-1))
    shape = (N, HW, HW, C)
    x = torch.relu(torch.randn(shape)).cuda()
    acts = torch.empty((L,) + shape, device="cuda", dtype=ch.saved_dtype)
    masks = torch.empty((L, N, HW, HW, C // 8), dtype=torch.uint8, device="cuda")
    y = torch.empty(shape, device="cuda")
    ch.forward(x, 0.125, acts=acts, masks=masks, y_final=y)
    dz = torch.empty((L,) + shape, device="cuda", dtype=ch.saved_dtype)
    dx = torch.empty(shape, device="cuda")
    ch.dgrad(torch.randn(shape).cuda(), masks, dz, dx, 0.125)
    grad = torch.empty((L, ch.num_params), device="cuda")
    ch.wgrad(x, acts, dz, grad.view(-1))
    torch.cuda.synchronize()
    print("chain %-9s C=%-3d %dx%d      |y| %.4f |dx| %.4f |g| %.4f" % (prec_name, C, HW, HW, float(y.norm()), float(dx.norm()), float(grad.norm())), flush=True)


def net_case():
    net = EulerNet(NetSpec(blocks_per_stage=(2, 2, 2), h=0.1), precision="fast_f16", seed=0)
    img = torch.randint(0, 256, (4, 32, 32, 3), dtype=torch.uint8).cuda()
    oh = torch.nn.functional.one_hot(torch.randint(0, 10, (4,)), 10).float().cuda()
    loss = [float(net.train_step(img, oh)) for _ in range(2)]
    print("net   fast_f16 cfg3-shaped (2,2,2) batch 4: loss %s" % loss, flush=True)


if __name__ == "__main__":
    torch.manual_seed(0)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "layer"):
        for precision, shape in (("strict", (2, 8, 8, 16)), ("fast_tf32", (2, 8, 8, 32)), ("fast_bf16", (1, 8, 8, 128)),
                                 ("fast_tf32", (1, 6, 5, 64)), ("strict", (1, 4, 4, 128)), ("fast_bf16", (1, 4, 8, 256))):
            layer_case(precision, shape)
    if which in ("all", "chain"):
        for prec_name in ("fast_tf32", "fast_f16"):
            for C, HW in ((16, 32), (32, 16), (64, 8)):
                chain_case(prec_name, C, HW)
    if which in ("all", "net"):
        net_case()
    print("done", flush=True)
