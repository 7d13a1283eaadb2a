"""Debug: dense weight gradient of one C=16 layer through the pixel-pair wgrad path vs torch.  Not a pytest."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import LayerHandle

N, H, W, C = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (1, 8, 8, 16))]
hd = LayerHandle(C, 3, 0.0, (1, 1), True, True, _abi.PREC_FAST_TF32, _abi.LAYOUT_3BY3)
hd.pack(torch.zeros(hd.num_params, device="cuda"))
g = torch.Generator().manual_seed(0)
x = torch.randn((N, H, W, C), generator=g).cuda(); dz = torch.randn((N, H, W, C), generator=g).cuda()
gp, G = hd.wgrad(x, dz, want_dense=True)
torch.cuda.synchronize()
xp = torch.nn.functional.pad(x.double(), (0, 0, 3, 3, 3, 3))   # pad H, W by 3
def corr(da, db):   # sum_q x[y+da, x+db, ci] dz[y, x, o]
    xs = xp[:, 3 + da:3 + da + H, 3 + db:3 + db + W, :]
    return torch.einsum("nyxc,nyxo->co", xs, dz.double())
ref = {(a, b): corr(a - 1, b - 1) for a in range(3) for b in range(3)}
for a in range(3):
    for b in range(3):
        got = G[a, b].double()
        err = float((got - ref[(a, b)]).norm() / ref[(a, b)].norm())
        best = None
        if err > 1e-2:
            cands = []
            for da in range(-3, 4):
                for db in range(-3, 4):
                    r = corr(da, db)
                    cands.append((float((got - r).norm() / r.norm()), da + 1, db + 1, "x"))
                    cands.append((float((got - r.t()).norm() / r.norm()), da + 1, db + 1, "xT"))
            best = sorted(cands)[:2]
        print("tap (%d,%d) rel err %.3e" % (a, b, err), best or "")
