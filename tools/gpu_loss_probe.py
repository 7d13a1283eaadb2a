"""Debug: loss trajectory of the cfg3 net for a few step sizes h (is the synthetic workload numerically alive?). Not a pytest."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200.training import EulerNet, NetSpec
g = torch.Generator().manual_seed(1236)
img = torch.randint(0, 256, (128, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
lab = torch.nn.functional.one_hot(torch.randint(0, 10, (128,), generator=g), 10).float().cuda()
for T in (8.0, 4.0, 2.0, 1.0, 0.5):
    spec = NetSpec(blocks_per_stage=(36, 37, 37), filters_per_block=(16, 32, 64), h=T / 108.0)
    net = EulerNet(spec, precision="fast_tf32", seed=1236)
    losses = [float(net.train_step(img, lab)) for _ in range(12)]
    print("T=%.1f h=%.5f losses %s" % (T, T / 108.0, " ".join("%.3f" % l for l in losses)), flush=True)
