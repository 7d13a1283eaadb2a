"""GPU parity of the persistent Euler-step chain kernels (b200ode_chain_*): n stacked Euler steps
(models/tfkeras_resnets.py:28-94 / :575-593 of the reference) forward, backward sweep and the
layer-batched weight gradient, against the NumPy float64 oracle and against the per-layer kernels.

Tolerance: fast_tf32 mode (tf32-truncated operands, fp32 accumulate): 2e-3 relative per step
output, 1e-2 for the data gradient, 5e-2 for the folded weight gradient (same as the per-layer
fast_tf32 tests); forward results must equal the per-layer fast_tf32 kernels bit for bit."""
import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _setup(C, L, gamma, seed):
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    rng = np.random.default_rng(seed)
    flats = np.stack([O0.init_params_3by3(rng, C, bias_std=0.1) for _ in range(L)]).astype(np.float32)
    ch = ChainHandle(C, L, gamma)
    assert ch.num_params == flats.shape[1]
    theta = torch.from_numpy(flats).cuda().contiguous()
    ch.pack(theta.view(-1))
    return ch, flats, theta


def _oracle_chain(x, flats, C, gamma, h, dy):
    L = flats.shape[0]
    Ks = [O0.assemble_kernel_3by3_closed(flats[l].astype(np.float64), C, gamma) for l in range(L)]
    xs, caches = [x], []
    for l in range(L):
        y, cache = O0.euler_step_fwd(xs[-1], Ks[l], flats[l, -C:].astype(np.float64), h)
        xs.append(y); caches.append(cache)
    grads, dzs = [None] * L, [None] * L
    d = dy
    for l in range(L - 1, -1, -1):
        dX, G, dbias, _, _ = O0.euler_step_bwd(d, caches[l], Ks[l], h)
        grads[l] = O0.fold_grad_3by3(G, C, dbias)
        d = dX
    return xs, d, grads


@pytest.mark.parametrize("C,H,W,N,L,gamma", [
    (16, 32, 32, 5, 3, -0.1), (32, 16, 16, 4, 4, 0.0), (64, 8, 8, 6, 3, -0.1),
    (16, 8, 8, 3, 2, 0.0), (32, 9, 12, 2, 3, -0.1), (16, 6, 5, 151, 2, -0.1), (64, 4, 4, 2, 1, 0.0),
])
def test_chain_matches_oracle(C, H, W, N, L, gamma):
    from differential_equations_resnet_b200.layers._base import ChainHandle
    assert ChainHandle.supported(C, H, W)
    h = 0.125
    ch, flats, theta = _setup(C, L, gamma, seed=C + H + L)
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn((N, H, W, C), generator=g))
    dy = torch.randn((N, H, W, C), generator=g)
    xd, dyd = x.cuda(), dy.cuda()
    acts = torch.empty((L, N, H, W, C), device="cuda")
    masks = torch.empty((L, N, H, W, C // 8), dtype=torch.uint8, device="cuda")
    yfin = torch.empty((N, H, W, C), device="cuda")
    ch.forward(xd, h, acts=acts, masks=masks, y_final=None)
    ch.forward(xd, h, acts=None, masks=None, y_final=yfin)       # inference form: only the last step leaves the SM
    dz = torch.empty((L, N, H, W, C), device="cuda")
    dx = torch.empty((N, H, W, C), device="cuda")
    ch.dgrad(dyd, masks, dz, dx, h)
    grad = torch.empty((L, ch.num_params), device="cuda")
    ch.wgrad(xd, acts, dz, grad.view(-1))
    torch.cuda.synchronize()
    xs, dX, grads = _oracle_chain(x.numpy().astype(np.float64), flats, C, gamma, h, dy.numpy().astype(np.float64))
    for l in range(L):
        assert rel(acts[l].cpu().numpy(), xs[l + 1]) <= 2e-3, ("step", l)
    assert torch.equal(yfin, acts[L - 1])
    assert rel(dx.cpu().numpy(), dX) <= 1e-2
    for l in range(L):
        assert rel(grad[l].cpu().numpy(), grads[l]) <= 5e-2, ("wgrad", l)


@pytest.mark.parametrize("C,H,W", [(16, 32, 32), (32, 16, 16), (64, 8, 8)])
def test_chain_equals_per_layer_kernels(C, H, W):
    """Same MMA accumulation order and epilogue arithmetic as conv_tc_kernel -> identical bits."""
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import LayerHandle, relu_scale_bwd
    N, L, h, gamma = 3, 3, 8.0 / 108.0, -0.05
    ch, flats, theta = _setup(C, L, gamma, seed=11)
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn((N, H, W, C), generator=g)).cuda()
    dy = torch.randn((N, H, W, C), generator=g).cuda()
    acts = torch.empty((L, N, H, W, C), device="cuda")
    masks = torch.empty((L, N, H, W, C // 8), dtype=torch.uint8, device="cuda")
    ch.forward(x, h, acts=acts, masks=masks)
    dz = torch.empty((L, N, H, W, C), device="cuda")
    dx = torch.empty((N, H, W, C), device="cuda")
    ch.dgrad(dy, masks, dz, dx, h)
    grad = torch.empty((L, ch.num_params), device="cuda")
    ch.wgrad(x, acts, dz, grad.view(-1))
    # per-layer path
    hds = [LayerHandle(C, 3, gamma, (1, 1), True, True, _abi.PREC_FAST_TF32, _abi.LAYOUT_3BY3) for _ in range(L)]
    cur, ys, ms = x, [], []
    for l in range(L):
        hds[l].pack(theta[l])
        y, m, _ = hds[l].forward(cur, h, flags=_abi.F_EULER, want_mask=True)
        ys.append(y); ms.append(m); cur = y
    d = dy
    for l in range(L - 1, -1, -1):
        assert torch.equal(acts[l], ys[l]) and torch.equal(masks[l], ms[l])
        dzl = relu_scale_bwd(d, ms[l], h)
        assert torch.equal(dz[l], dzl)
        gl = hds[l].wgrad(x if l == 0 else ys[l - 1], dzl)
        assert rel(grad[l].cpu().numpy(), gl.cpu().numpy()) <= 1e-5    # split-K partition differs -> summation order
        d = hds[l].dgrad(dzl, d, (H, W))
    assert torch.equal(dx, d)


@pytest.mark.parametrize("C,H,W,h,gamma,n", [(16, 32, 32, 0.01, -0.1, 200), (16, 32, 32, 0.125, 0.0, 1000),
                                             (64, 8, 8, 0.01, -0.1, 1000)])
def test_chain_long_horizon_shared_weights(C, H, W, h, gamma, n):
    """BASELINE cfg5: one block applied n times with shared weights (n_layers == 1, n_steps = n), ONE launch.
    Reports the free-running norm drift ratio against the float64 oracle."""
    N = 8
    ch, flats, theta = _setup(C, 1, gamma, seed=5)
    g = torch.Generator().manual_seed(1)
    x = torch.randn((N, H, W, C), generator=g)
    y = torch.empty((N, H, W, C), device="cuda")
    ch.forward(x.cuda(), h, n_steps=n, y_final=y)
    torch.cuda.synchronize()
    K = O0.assemble_kernel_3by3_closed(flats[0].astype(np.float64), C, gamma)
    cur = x.numpy().astype(np.float64)
    for _ in range(n):
        cur, _ = O0.euler_step_fwd(cur, K, flats[0, -C:].astype(np.float64), h)
    err = rel(y.cpu().numpy(), cur)
    print("cfg5 C=%d h=%g gamma=%g n=%d: |x_n| oracle %.4e gpu %.4e rel err %.2e" % (
        C, h, gamma, n, np.linalg.norm(cur), float(y.double().norm()), err))
    # free-running norm drift ratio (BASELINE cfg5): always; state error: where the dynamics are contractive
    # (gamma < 0, small h).  With gamma = 0 and h = 0.125 the state grows 1e3-fold over 1000 steps and the
    # per-step tf32 error (~3e-4 of the conv term) is amplified with it: the states decorrelate (~1e-1)
    # while the norms still agree to < 1 %.
    drift = float(y.double().norm()) / np.linalg.norm(cur)
    assert abs(drift - 1.0) <= 2e-2, drift
    if gamma < 0 and h <= 0.01:
        assert err <= 2e-3


def test_chain_refuses_what_does_not_fit():
    from differential_equations_resnet_b200.layers._base import ChainHandle
    assert not ChainHandle.supported(64, 32, 32)
    assert not ChainHandle.supported(16, 64, 64)
    ch = ChainHandle(64, 1, 0.0)
    ch.pack(torch.zeros(ch.num_params, device="cuda"))
    x = torch.zeros((1, 32, 32, 64), device="cuda")
    with pytest.raises(ValueError):
        ch.forward(x, 0.1, y_final=torch.empty_like(x))
