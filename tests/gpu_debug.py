"""Verbose GPU bring-up script (not a pytest): prints the relative error of every kernel family
against the oracle for a grid of shapes/modes without stopping at the first failure.
Each configuration runs in a subprocess so a trapping kernel does not hide the others."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASE = r'''
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from oracle import antisym_numpy as O0
import differential_equations_resnet_b200 as pkg
from differential_equations_resnet_b200.layers._base import relu_scale_bwd
N,H,W,C = %(shape)s
prec = %(prec)r
def rel(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64); return float(np.linalg.norm(a-b)/max(np.linalg.norm(b),1e-30))
gamma,h=-0.1,0.125
layer = pkg.Conv2DAntisymmetric3By3(gamma=gamma, precision=prec, seed=0); layer.build((N,H,W,C))
with torch.no_grad(): layer.packed[-C:] = torch.randn(C).cuda()*0.1
flat = layer.packed.detach().cpu().numpy().astype(np.float64)
K = O0.assemble_kernel_3by3_closed(flat, C, gamma)
Kg = layer.get_kernel()
print("  pack bit-exact:", np.array_equal(Kg, K.astype(np.float32)))
g = torch.Generator().manual_seed(1)
cast = (lambda t: t.to(torch.bfloat16)) if prec=="fast_bf16" else (lambda t: t)
x = cast(torch.relu(torch.randn((N,H,W,C),generator=g))); dy = cast(torch.randn((N,H,W,C),generator=g))
x64, dy64 = x.float().numpy().astype(np.float64), dy.float().numpy().astype(np.float64)
x, dy = x.cuda(), dy.cuda()
hd = layer._handle
hd.pack(layer.packed.detach())
z,_,_ = hd.forward(x, 1.0, 1); torch.cuda.synchronize()
print("  conv+bias rel err: %%.3e" %% rel(z.float().cpu().numpy(), O0.layer_call(x64,K,flat[-C:])))
y,mask,_ = hd.forward(x, h, 15, want_mask=True); torch.cuda.synchronize()
yref, cache = O0.euler_step_fwd(x64,K,flat[-C:],h)
print("  euler fwd rel err: %%.3e" %% rel(y.float().cpu().numpy(), yref))
mref = (cache["z"]>0)
mb = np.unpackbits(mask.cpu().numpy().reshape(-1,(C+7)//8), axis=1, bitorder="little")[:, :C].reshape(N,H,W,C).astype(bool)
print("  mask mismatches: %%d / %%d" %% (int((mb!=mref).sum()), mref.size))
dz = relu_scale_bwd(dy, mask, h)
dX,G,dbias,_,dZ = O0.euler_step_bwd(dy64, cache, K, h)
print("  dz rel err: %%.3e" %% rel(dz.float().cpu().numpy(), dZ))
dx = hd.dgrad(dz, dy, (H,W)); torch.cuda.synchronize()
print("  dgrad rel err: %%.3e" %% rel(dx.float().cpu().numpy(), dX))
if prec != "fast_bf16":
    gp, Gd = hd.wgrad(x, dz, want_dense=True); torch.cuda.synchronize()
    Gref = O0.conv_kernel_grad_stride1(x64, dz.float().cpu().numpy().astype(np.float64))
    print("  wgrad dense rel err: %%.3e" %% rel(Gd.cpu().numpy(), Gref))
    print("  wgrad folded rel err: %%.3e" %% rel(gp.cpu().numpy(), O0.fold_grad_3by3(Gref, C, dz.float().cpu().numpy().astype(np.float64).sum(axis=(0,1,2)))))
else:
    import ctypes
    from differential_equations_resnet_b200 import _abi
    Gd = torch.empty((3,3,C,C),device="cuda"); gp = torch.zeros(hd.num_params,device="cuda")
    rc = _abi.lib().b200ode_euler_wgrad(hd._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(dz.data_ptr()), ctypes.c_void_p(gp.data_ptr()), ctypes.c_void_p(Gd.data_ptr()), N,H,W,0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    Gref = O0.conv_kernel_grad_stride1(x64, dz.float().cpu().numpy().astype(np.float64))
    print("  wgrad dense rel err: %%.3e (rc=%%d)" %% (rel(Gd.cpu().numpy(), Gref), rc))
'''

def main():
    shapes = [(2, 8, 8, 16), (3, 12, 10, 32), (2, 32, 32, 64), (5, 16, 16, 32), (9, 8, 8, 64), (2, 16, 16, 128),
              (2, 8, 8, 256), (1, 33, 17, 16), (4, 32, 32, 16)]
    precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["simt", "fast_tf32", "strict", "fast_bf16"]
    for prec in precs:
        for shape in shapes:
            if prec == "simt" and shape[0] * shape[1] * shape[2] * shape[3] > 70000:
                continue
            print("== %s %s" % (prec, shape), flush=True)
            code = CASE % {"root": ROOT, "shape": repr(shape), "prec": prec}
            try:
                r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
                print(r.stdout.rstrip())
                if r.returncode != 0:
                    print("  FAILED rc=%d: %s" % (r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ""))
                    tail = [l for l in r.stderr.splitlines() if "b200ode" in l or "Error" in l]
                    for l in tail[-4:]:
                        print("   |", l)
            except subprocess.TimeoutExpired:
                print("  TIMEOUT")
            sys.stdout.flush()

if __name__ == "__main__":
    main()
