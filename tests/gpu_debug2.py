import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import antisym_numpy as O0
import differential_equations_resnet_b200 as pkg
def rel(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return float(np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30))
for prec in ("strict","fast_tf32"):
  for shape in ((9,8,8,64),(2,4,4,64),(2,8,8,16),(4,4,4,16)):
    N,H,W,C = shape
    layer = pkg.Conv2DAntisymmetric3By3(gamma=-0.1, precision=prec, seed=0); layer.build(shape)
    flat = layer.packed.detach().cpu().numpy().astype(np.float64)
    K = O0.assemble_kernel_3by3_closed(flat, C, -0.1)
    g = torch.Generator().manual_seed(11)
    x = torch.relu(torch.randn(shape, generator=g)); x64 = x.numpy().astype(np.float64); x = x.cuda()
    for rep in range(3):
        with torch.no_grad():
            z = layer(x).cpu().numpy()
        zr = O0.layer_call(x64, K, flat[-C:])
        print(prec, shape, "rep", rep, "total %.2e"%rel(z,zr), "per-image", ["%.1e"%rel(z[i],zr[i]) for i in range(N)], flush=True)
