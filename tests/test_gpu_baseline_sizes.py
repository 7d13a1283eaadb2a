"""GPU parity at the REAL BASELINE.json sizes (run with `-m gpu` on a B200).

  cfg2  single Euler layer, N=256, 32x32, C in {16,32,64,128,256}, strict / fast_tf32 / fast_bf16:
        forward, data gradient, dense + folded weight gradient against a float64 restatement
        (oracle/antisym_torch.py ops on float64 CPU tensors + torch autograd = SURVEY.md App. A.3/A.4).
  cfg3  the 108-Euler-step net (36/37/37 blocks, 16/32/64 channels) at batch 128 through EulerNet:
        loss, EVERY layer's gradient and one Adam update against O1, strict and fast mode, eager and CUDA graph.
  cfg4  one Euler block at (512,64,64,256) in fast_bf16: forward / data gradient on sampled pixels against
        float64 patch GEMMs, weight gradient through rank-one probes u^T G v evaluated independently in float64.

Tolerances (north star): strict <= 1e-5 relative; fast modes stated per assertion.
"""
import functools

import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0
from oracle import antisym_torch as O1

pytestmark = pytest.mark.gpu

FWD_TOL = {"strict": 1e-5, "fast_tf32": 1e-3, "fast_bf16": 1.5e-2}
# data gradient dX = dY - conv_K(dZ) + 2*gamma*dZ and folded weight gradient S = G - rot180(G)^T
DGRAD_TOL = {"strict": 1e-5, "fast_tf32": 1e-3, "fast_bf16": 1.5e-2}
WGRAD_TOL = {"strict": 1e-5, "fast_tf32": 5e-3, "fast_bf16": 3e-2}


def rel(a, b):
    a = a.double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a, np.float64))
    b = b.double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b, np.float64))
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def _unpack_mask(mask, C):
    """[..., C/8] uint8 bit mask (bit j of byte g = channel 8g+j) -> [..., C] bool"""
    sh = torch.arange(8, device=mask.device, dtype=torch.uint8)
    return ((mask[..., None] >> sh) & 1).reshape(mask.shape[:-1] + (C,)).bool()


def _cfg2_inputs(C):
    rng = np.random.default_rng(1234 + 1)
    flat = O0.init_params_3by3(rng, C, bias_std=0.1).astype(np.float32)
    g = torch.Generator().manual_seed(1234 + 1)
    x = torch.relu(torch.randn((256, 32, 32, C), generator=g))
    dy = torch.randn((256, 32, 32, C), generator=g)
    return flat, x, dy


def _euler_reference(flat, x, dy, mask, C, gamma, h, device):
    """float64 restatement (oracle/antisym_torch.py ops + torch autograd) of one Euler step and its backward on
    `device`.  The backward is evaluated with the relu mask the GPU path took (`mask`, bool): relu's derivative is
    discontinuous, so an element whose pre-activation sits at rounding level legitimately takes either branch; the
    test bounds the number of such elements separately instead of letting a handful of them dominate the norm."""
    K = torch.from_numpy(O0.assemble_kernel_3by3_closed(flat.astype(np.float64), C, gamma)).to(device).requires_grad_(True)
    b = torch.from_numpy(flat[-C:].astype(np.float64)).to(device).requires_grad_(True)
    x64 = x.double().to(device).requires_grad_(True)
    z = O1.conv2d_same_nhwc(x64, K) + b
    y = x64 + h * torch.relu(z)
    y_forced = x64 + h * (z * mask.to(device).double())
    dX, dK, db = torch.autograd.grad(y_forced, (x64, K, b), dy.double().to(device))
    gflat = O0.fold_grad_3by3(dK.cpu().numpy(), C, db.cpu().numpy())
    return dict(y=y.detach(), z=z.detach(), dX=dX, G=dK, g=torch.from_numpy(gflat))


def test_gpu_float64_reference_equals_cpu_oracle():
    """The full-size references below are the oracle's float64 ops executed by torch on the GPU (cuDNN float64);
    this pins that evaluation to the CPU oracle at a size the CPU finishes in seconds."""
    C, gamma, h = 32, -0.1, 0.125
    flat, x, dy = _cfg2_inputs(C)
    x, dy = x[:16], dy[:16]
    mask = torch.randn(x.shape, generator=torch.Generator().manual_seed(2)) > 0
    a = _euler_reference(flat, x, dy, mask, C, gamma, h, "cpu")
    b = _euler_reference(flat, x, dy, mask, C, gamma, h, "cuda")
    for k in ("y", "dX", "G", "g"):
        assert rel(b[k], a[k]) <= 1e-12, k
    # and the CPU float64 torch restatement equals the NumPy literal oracle (O0)
    K = O0.assemble_kernel_3by3_closed(flat.astype(np.float64), C, gamma)
    y0, _ = O0.euler_step_fwd(x[:2].double().numpy(), K, flat[-C:].astype(np.float64), h)
    assert rel(a["y"][:2], y0) <= 1e-12


@pytest.mark.parametrize("precision", ["strict", "fast_tf32", "fast_bf16"])
@pytest.mark.parametrize("C", [16, 32, 64, 128, 256])
def test_cfg2_layer_at_full_size(C, precision):
    import ctypes
    import differential_equations_resnet_b200 as pkg
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import relu_scale_bwd
    gamma, h = -0.1, 0.125
    flat, x, dy = _cfg2_inputs(C)
    layer = pkg.Conv2DAntisymmetric3By3(gamma=gamma, precision=precision, seed=0)
    layer.build((None, 32, 32, C))
    with torch.no_grad():
        layer.packed.copy_(torch.from_numpy(flat))
    hd = layer._handle
    dt = hd.io_dtype
    x, dy = x.to(dt), dy.to(dt)            # bf16 mode consumes bf16-rounded inputs: its reference sees exactly those values
    xd, dyd = x.cuda(), dy.cuda()
    hd.pack(layer.packed.detach())
    y, mask, _ = hd.forward(xd, h, 15, want_mask=True)
    dz = relu_scale_bwd(dyd, mask, h)
    dx = hd.dgrad(dz, dyd, (32, 32))
    if precision == "fast_bf16":
        G = torch.empty((3, 3, C, C), device="cuda")
        gpk = torch.zeros(hd.num_params, device="cuda")
        _abi.check(_abi.lib().b200ode_euler_wgrad(hd._h, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(dz.data_ptr()),
                                                  ctypes.c_void_p(gpk.data_ptr()), ctypes.c_void_p(G.data_ptr()), 256, 32, 32, 0,
                                                  torch.cuda.current_stream().cuda_stream))
    else:
        gpk, G = hd.wgrad(xd, dz, want_dense=True)
    torch.cuda.synchronize()
    mb = _unpack_mask(mask, C)
    ref = _euler_reference(flat, x.float(), dy.float(), mb, C, gamma, h, "cuda")
    e_fwd, e_dx, e_G = rel(y, ref["y"]), rel(dx, ref["dX"]), rel(G, ref["G"])
    nb = C if precision == "fast_bf16" else 0       # bf16 mode: no bias gradient inside the wgrad launch
    e_g = rel(gpk[:gpk.numel() - nb], ref["g"][:gpk.numel() - nb])
    # relu branch disagreements: only where the reference pre-activation is at the rounding level of the mode
    flip = mb != (ref["z"] > 0)
    nflip = int(flip.sum())
    zmax = float(ref["z"][flip].abs().max()) if nflip else 0.0
    zscale = float(ref["z"].abs().mean())
    print("cfg2 C=%d %s: fwd %.2e dgrad %.2e dense wgrad %.2e folded wgrad %.2e | relu flips %d of %d, max |z| there %.1e (mean |z| %.2f)"
          % (C, precision, e_fwd, e_dx, e_G, e_g, nflip, mb.numel(), zmax, zscale))
    assert e_fwd <= FWD_TOL[precision]
    assert e_dx <= DGRAD_TOL[precision]
    assert e_G <= WGRAD_TOL[precision]
    assert e_g <= WGRAD_TOL[precision]
    assert zmax <= 20 * FWD_TOL[precision] * zscale
    assert nflip <= 40 * FWD_TOL[precision] * mb.numel()


# --------------------------------------------------------------------------------------------- cfg3 ---
CFG3 = dict(blocks_per_stage=(36, 37, 37), filters_per_block=(16, 32, 64), h=2.0 / 108.0, gamma=0.0)


def _cfg3_inputs():
    spec = O1.NetSpec(**CFG3)
    P = O1.init_net_params(spec, seed=1236)
    gen = torch.Generator().manual_seed(1236)
    img = torch.randint(0, 256, (128, 32, 32, 3), generator=gen, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (128,), generator=gen), 10).float()
    return spec, P, img, lab


def _euler_masks(net, C_of):
    """name -> bool [N,H,W,C] relu branch taken by the GPU path in every Euler layer of the last step"""
    out = {}
    names = [n for n, _, _, _ in net.layer_param_slices()]
    i = 0
    for seg in net.segments:
        if seg[0] != "chain":
            continue
        ch = seg[1]
        for l in range(ch.n):
            m = ch.saved_mask(l)
            out[names[i]] = _unpack_mask(m, ch.C).cpu()
            i += 1
    return out


@pytest.mark.parametrize("precision,graph", [("strict", False), ("fast_f16", False), ("fast_f16", True), ("fast_tf32", False), ("fast_tf32", True),
                                             ("strict", True)])
def test_cfg3_full_depth_batch128(precision, graph):
    """108 Euler steps + 2 transitions, batch 128: the grid the bench runs (128 CTAs, clusters, 36-step chains).
    Loss, every layer's gradient and one Adam update against O1.  O1's backward is evaluated with the relu branches the
    GPU took in the Euler layers (226 M relu decisions per step; the few hundred whose pre-activation is at rounding
    level may legitimately differ and each would move its layer's dZ by ~1e-3); their number is bounded separately."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    spec, P, img, lab = _cfg3_inputs()
    P0 = {k: v.clone() for k, v in P.items()}
    net = EulerNet(NetSpec(**CFG3), precision=precision, seed=0)
    net.import_params(P0)
    imgd, labd = img.cuda(), lab.cuda()
    if graph:
        net.capture(imgd, labd, warmup=1)
        loss = float(net.train_step_graph())
    else:
        loss = float(net.train_step(imgd, labd))
    torch.cuda.synchronize()
    masks = _euler_masks(net, None)
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    rec = {}
    loss_ref, grads = O1.train_step(spec, P, M, V, 1, img, lab, forced_masks=masks, record_z=rec)
    nflip = sum(int((masks[k] != rec[k][0]).sum()) for k in masks)
    ntot = sum(m.numel() for m in masks.values())
    zflip = max([float(rec[k][1][masks[k] != rec[k][0]].max()) for k in masks if bool((masks[k] != rec[k][0]).any())] or [0.0])
    tol_loss, tol_g, tol_upd, tol_flip = {"strict": (1e-5, 1e-4, 3e-2, 2e-6), "fast_tf32": (1e-3, 1e-2, 0.35, 2e-3), "fast_f16": (1e-3, 1e-2, 0.35, 2e-3)}[precision]
    # (Adam's first update is ~lr*sign(g): its relative error is sqrt(4 * fraction of sign disagreements) ~ 2*sqrt(gradient error))
    g = net.export_grads()
    worst, worst_u = ("", 0.0), ("", 0.0)
    errs = {}
    for k, gr in grads.items():
        errs[k] = rel(g[k], gr)
        if errs[k] > worst[1]:
            worst = (k, errs[k])
    th = net.export_params()
    for k in P0:       # one Adam update, compared RELATIVELY on the update itself (theta1 - theta0 ~ 1e-3 per element)
        du, du_ref = th[k] - P0[k], P[k] - P0[k]
        if float(du_ref.abs().max()) == 0.0:
            assert float(du.abs().max()) == 0.0, k
            continue
        e = rel(du, du_ref)
        if e > worst_u[1]:
            worst_u = (k, e)
    print("cfg3 %s graph=%s: loss %.6f (ref %.6f), worst gradient %s %.2e, worst Adam update %s %.2e, relu flips %d of %d (max |z| %.1e)"
          % (precision, graph, loss, loss_ref, worst[0], worst[1], worst_u[0], worst_u[1], nflip, ntot, zflip))
    assert abs(loss - loss_ref) <= tol_loss * max(1.0, abs(loss_ref)), (loss, loss_ref)
    assert worst[1] <= tol_g, worst
    assert worst_u[1] <= tol_upd, worst_u
    assert nflip <= tol_flip * ntot


# --------------------------------------------------------------------------------------------- cfg4 ---
def test_cfg4_block_at_full_size_bf16():
    """One Euler block of BASELINE configs[3]: (512,64,64,256), bf16 activations/operands, fp32 accumulate."""
    import ctypes
    import differential_equations_resnet_b200 as pkg
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import relu_scale_bwd
    N, H, W, C, gamma, h = 512, 64, 64, 256, -0.1, 0.125
    layer = pkg.Conv2DAntisymmetric3By3(gamma=gamma, precision="fast_bf16", seed=3)
    layer.build((None, H, W, C))
    with torch.no_grad():
        layer.packed[-C:] = torch.randn(C, generator=torch.Generator().manual_seed(4)).cuda() * 0.1
    hd = layer._handle
    hd.pack(layer.packed.detach())
    flat = layer.packed.detach().cpu().numpy().astype(np.float64)
    K = torch.from_numpy(O0.assemble_kernel_3by3_closed(flat, C, gamma)).cuda()          # float64 on the GPU (torch ops: checker)
    bias = torch.from_numpy(flat[-C:]).cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.relu(torch.randn((N, H, W, C), generator=gen, device="cuda")).to(torch.bfloat16)
    dy = torch.randn((N, H, W, C), generator=gen, device="cuda").to(torch.bfloat16)
    y, mask, _ = hd.forward(x, h, 15, want_mask=True)
    dz = relu_scale_bwd(dy, mask, h)
    dx = hd.dgrad(dz, dy, (H, W))
    # ---- sampled pixels: float64 patch GEMM ----
    S = 4096
    cpu_gen = torch.Generator().manual_seed(6)
    n_i = torch.randint(0, N, (S,), generator=cpu_gen).cuda()
    y_i = torch.randint(0, H, (S,), generator=cpu_gen).cuda()
    x_i = torch.randint(0, W, (S,), generator=cpu_gen).cuda()
    y_i[:64], x_i[:64] = 0, 0                      # corners and edges are in the sample
    y_i[64:128], x_i[64:128] = H - 1, W - 1
    y_i[128:192] = 0
    x_i[192:256] = W - 1

    def patches(t):
        out = torch.zeros((S, 3, 3, C), dtype=torch.float64, device="cuda")
        for a in range(3):
            for b in range(3):
                yy, xx = y_i + a - 1, x_i + b - 1
                ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                v = t[n_i, yy.clamp(0, H - 1), xx.clamp(0, W - 1)].double()
                out[:, a, b] = v * ok[:, None]
        return out

    z_ref = patches(x).reshape(S, -1) @ K.reshape(9 * C, C) + bias
    y_ref = x[n_i, y_i, x_i].double() + h * torch.relu(z_ref)
    e_fwd = rel(y[n_i, y_i, x_i], y_ref)
    conv_dz = patches(dz).reshape(S, -1) @ K.reshape(9 * C, C)
    dx_ref = dy[n_i, y_i, x_i].double() - conv_dz + 2.0 * gamma * dz[n_i, y_i, x_i].double()
    e_dx = rel(dx[n_i, y_i, x_i], dx_ref)
    # relu mask of the sampled pixels (bits flip only where |z| is at rounding level)
    bits = mask[n_i, y_i, x_i]                                    # [S, C/8]
    got = ((bits[:, :, None] >> torch.arange(8, device="cuda", dtype=torch.uint8)) & 1).reshape(S, C).bool()
    flips = float((got != (z_ref > 0)).double().mean())
    # ---- weight gradient: rank-one probes u^T G[tap] v = sum_q (x[q+shift].u)(dz[q].v), float64, all 2.1 M pixels ----
    G = torch.empty((3, 3, C, C), device="cuda")
    gpk = torch.zeros(hd.num_params, device="cuda")
    _abi.check(_abi.lib().b200ode_euler_wgrad(hd._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(dz.data_ptr()),
                                              ctypes.c_void_p(gpk.data_ptr()), ctypes.c_void_p(G.data_ptr()), N, H, W, 0,
                                              torch.cuda.current_stream().cuda_stream))
    worst = 0.0
    pg = torch.Generator().manual_seed(7)
    for _ in range(3):
        u = torch.randn(C, generator=pg, dtype=torch.float64).cuda()
        v = torch.randn(C, generator=pg, dtype=torch.float64).cuda()
        xu = torch.zeros((N, H + 2, W + 2), dtype=torch.float64, device="cuda")
        for n0 in range(0, N, 64):                                # chunked: keeps the float64 temporaries small
            xu[n0:n0 + 64, 1:-1, 1:-1] = x[n0:n0 + 64].double() @ u
        dv = torch.empty((N, H, W), dtype=torch.float64, device="cuda")
        for n0 in range(0, N, 64):
            dv[n0:n0 + 64] = dz[n0:n0 + 64].double() @ v
        want = torch.stack([torch.stack([(xu[:, a:a + H, b:b + W] * dv).sum() for b in range(3)]) for a in range(3)])
        gotp = torch.einsum("abio,i,o->ab", G.double(), u, v)
        worst = max(worst, rel(gotp, want))
    print("cfg4 block (512,64,64,256) bf16: fwd %.2e dgrad %.2e relu-mask flips %.2e wgrad probes %.2e" % (e_fwd, e_dx, flips, worst))
    assert e_fwd <= 1.5e-2 and e_dx <= 1.5e-2
    assert flips <= 2e-3
    assert worst <= 1.5e-2
