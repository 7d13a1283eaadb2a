"""The two configurations the reference's notebooks publish numbers for, on this implementation:
  * train throughput, antisymmetric net: stem + 64 single-layer Euler blocks, 16 channels, 32x32, h = 8/64, no BN,
    batch 32, Adam(1e-3, eps 1e-7)  -- experiments_antisymmetric_resnet_v6.ipynb:362 (1.46 it/s = 46.7 img/s)
  * inference latency at batch 1 of the same net -- experiments_antisymmetric_resnet_v7.ipynb:650-651 (199.3 ms, 5.02 FPS)
Not a pytest.  usage: python tools/gpu_reference_configs.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200.training import EulerNet, NetSpec

spec = NetSpec(num_stages=2, blocks_per_stage=(64,), filters_per_block=(16,), strides=((1, 1),), h=8.0 / 64)
g = torch.Generator().manual_seed(0)
net = EulerNet(spec, precision="fast_tf32", seed=0)
img = torch.randint(0, 256, (32, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
lab = torch.nn.functional.one_hot(torch.randint(0, 10, (32,), generator=g), 10).float().cuda()
net.train_step(img, lab)
net.capture(img, lab)
for _ in range(5): net.train_step_graph()
torch.cuda.synchronize(); t0 = time.time()
for _ in range(200): net.train_step_graph()
torch.cuda.synchronize(); dt = (time.time() - t0) / 200
print("train, batch 32, 64 blocks x 16 ch: %.3f ms/step = %.1f it/s = %.0f img/s   (reference notebook: 1.46 it/s = 46.7 img/s)" % (dt * 1e3, 1 / dt, 32 / dt))
one = img[:1].contiguous()
for _ in range(5): net.predict(one)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(300): net.predict(one)
torch.cuda.synchronize(); dt = (time.time() - t0) / 300
print("inference, batch 1: %.3f ms = %.0f FPS   (reference notebook: 199.3 ms = 5.02 FPS antisymmetric, 4.4 ms = 229 FPS regular)" % (dt * 1e3, 1 / dt))
