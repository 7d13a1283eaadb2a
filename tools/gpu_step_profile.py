"""Run a few EAGER cfg3 train steps (no CUDA graph) so that `ncu --metrics gpu__time_duration.sum`
lists every launch of one step.  Not a pytest.

usage: python tools/gpu_step_profile.py [steps] [precision] [blocks]
Prints the number of libb200ode launches per step (use it for ncu's -s / -c).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from differential_equations_resnet_b200 import _abi  # noqa: E402
from differential_equations_resnet_b200.training import EulerNet, NetSpec  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    prec = sys.argv[2] if len(sys.argv) > 2 else "fast_tf32"
    blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 36
    spec = NetSpec(blocks_per_stage=(blocks, blocks + 1, blocks + 1), filters_per_block=(16, 32, 64), h=2.0 / 108.0)
    net = EulerNet(spec, precision=prec, seed=1236)
    g = torch.Generator().manual_seed(1236)
    img = torch.randint(0, 256, (128, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (128,), generator=g), 10).float().cuda()
    for s in range(steps):
        n0 = _abi.launch_count()
        loss = net.train_step(img, lab)
        torch.cuda.synchronize()
        print("step %d loss %.5f  libb200ode launches %d" % (s, float(loss), _abi.launch_count() - n0), flush=True)


if __name__ == "__main__":
    main()
