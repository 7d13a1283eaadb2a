"""BASELINE.json configs[3]: wide antisymmetric ResNet -- stem 3->256 at 64x64, 8 Euler steps with 256 channels,
GAP + FC, bf16 fast mode; one full train step (forward, loss, backward, Adam).  Not a pytest.
usage: python tools/gpu_cfg4.py [batch=512] [precision=fast_bf16] [steps=5]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.training import EulerNet, NetSpec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "fast_bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
C, L, HW = 256, 8, 64
spec = NetSpec(num_stages=2, blocks_per_stage=(L,), filters_per_block=(C,), strides=((1, 1),), h=1.0 / L)
net = EulerNet(spec, precision=prec, seed=0)
g = torch.Generator().manual_seed(0)
img = torch.randint(0, 256, (B, HW, HW, 3), generator=g, dtype=torch.uint8).cuda()
lab = torch.nn.functional.one_hot(torch.randint(0, 10, (B,), generator=g), 10).float().cuda()
l0 = _abi.launch_count()
losses = [float(net.train_step(img, lab)) for _ in range(2)]
per_step = (_abi.launch_count() - l0) // 2
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    losses.append(float(net.train_step(img, lab)))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flops = 3 * L * 2.0 * B * HW * HW * 9 * C * C          # Euler blocks only (fwd + dgrad + wgrad)
print("cfg4 %s batch %d: %.2f ms/step = %.0f img/s; Euler-block algorithmic %.0f TFLOP/s (%.1f%% of the measured bf16 peak 1648.7); "
      "%d libb200ode launches/step; losses %s; peak memory %.1f GB" % (
          prec, B, ms, B / (ms * 1e-3), flops / (ms * 1e-3) * 1e-12, 100 * flops / (ms * 1e-3) * 1e-12 / 1648.7, per_step,
          " ".join("%.3f" % l for l in losses), torch.cuda.max_memory_allocated() / 2**30))
