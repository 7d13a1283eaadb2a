"""GPU parity of the fp16-operand persistent chains (b200ode_chain_* with B200ODE_PREC_FAST_F16; csrc/kernels_chain_f16.cuh):
n stacked Euler steps (models/tfkeras_resnets.py:28-94 / :575-593 of the reference) forward, backward sweep and the
layer-batched weight gradient against the NumPy float64 oracle.

Tolerances: operands carry an 11-bit significand rounded to nearest (fp16 = the tf32 grade), accumulation and the
residual stream are fp32: <= 1e-3 relative per step output, <= 1e-3 for the data gradient, <= 5e-3 for the folded
weight gradient.  The oracle's backward takes the relu branches the GPU took (relu' is discontinuous; the number of
disagreeing branch decisions is bounded separately: only where |z| is at the rounding level)."""
import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def unpack(mask, C):
    m = mask.cpu().numpy()
    return np.unpackbits(m, axis=-1, bitorder="little").reshape(m.shape[:-1] + (C,)).astype(bool)


def _setup(C, L, gamma, seed):
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    rng = np.random.default_rng(seed)
    flats = np.stack([O0.init_params_3by3(rng, C, bias_std=0.1) for _ in range(L)]).astype(np.float32)
    ch = ChainHandle(C, L, gamma, precision=_abi.PREC_FAST_F16)
    assert ch.num_params == flats.shape[1] and ch.f16
    theta = torch.from_numpy(flats).cuda().contiguous()
    ch.pack(theta.view(-1))
    return ch, flats, theta


def _run_gpu(ch, x, dy, h, L):
    N, H, W, C = x.shape
    xd, dyd = x.cuda(), dy.cuda()
    acts = torch.empty((L, N, H, W, C), device="cuda", dtype=torch.float16)
    masks = torch.empty((L, N, H, W, C // 8), dtype=torch.uint8, device="cuda")
    y = torch.empty((N, H, W, C), device="cuda")
    y2 = torch.empty((N, H, W, C), device="cuda")
    ch.forward(xd, h, acts=acts, masks=masks, y_final=y)
    ch.forward(xd, h, acts=None, masks=None, y_final=y2)         # inference form: nothing but the last step leaves the SM
    dz = torch.empty((L, N, H, W, C), device="cuda", dtype=torch.float16)
    dx = torch.empty((N, H, W, C), device="cuda")
    ch.dgrad(dyd, masks, dz, dx, h)
    grad = torch.empty((L, ch.num_params), device="cuda")
    ch.wgrad(xd, acts, dz, grad.view(-1))
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    return acts, masks, y, dz, dx, grad


def _oracle(x, flats, C, gamma, h, dy, masks):
    L = flats.shape[0]
    Ks = [O0.assemble_kernel_3by3_closed(flats[l].astype(np.float64), C, gamma) for l in range(L)]
    xs, caches = [x], []
    for l in range(L):
        y, cache = O0.euler_step_fwd(xs[-1], Ks[l], flats[l, -C:].astype(np.float64), h)
        xs.append(y); caches.append(cache)
    grads, dzs = [None] * L, [None] * L
    d = dy
    for l in range(L - 1, -1, -1):
        dX, G, dbias, _, dZ = O0.euler_step_bwd(d, caches[l], Ks[l], h, mask=masks[l])
        grads[l] = O0.fold_grad_3by3(G, C, dbias)
        dzs[l] = dZ
        d = dX
    flips = sum(int((masks[l] != (caches[l]["u"] > 0)).sum()) for l in range(L))
    return xs, d, grads, dzs, flips


@pytest.mark.parametrize("C,H,W,N,L,gamma,h", [
    (16, 32, 32, 5, 3, -0.1, 0.125), (32, 16, 16, 4, 4, 0.0, 0.125), (64, 8, 8, 6, 3, -0.1, 0.125),
    (16, 8, 8, 3, 2, 0.0, 1.0), (32, 9, 12, 2, 3, -0.1, 0.5), (16, 6, 5, 151, 2, -0.1, 0.125), (64, 4, 4, 2, 1, 0.0, 0.125),
    (64, 8, 8, 130, 2, -0.05, 2.0 / 108), (32, 16, 16, 9, 5, 0.0, 2.0 / 108), (16, 32, 32, 3, 6, 0.0, 2.0 / 108),
])
def test_chain_f16_matches_oracle(C, H, W, N, L, gamma, h):
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    assert ChainHandle.supported(C, H, W, _abi.PREC_FAST_F16)
    ch, flats, theta = _setup(C, L, gamma, seed=C + H + L)
    g = torch.Generator().manual_seed(7)
    x = torch.relu(torch.randn((N, H, W, C), generator=g))
    dy = torch.randn((N, H, W, C), generator=g) * 1e-3           # realistic gradient magnitudes (mean loss over a batch)
    acts, masks, y, dz, dx, grad = _run_gpu(ch, x, dy, h, L)
    mb = [unpack(masks[l], C) for l in range(L)]
    xs, dX, grads, dzs, flips = _oracle(x.numpy().astype(np.float64), flats, C, gamma, h, dy.numpy().astype(np.float64), mb)
    for l in range(L):
        assert rel(acts[l].float().cpu().numpy(), xs[l]) <= 1e-3, ("saved operand of step", l)
    assert rel(y.cpu().numpy(), xs[L]) <= 1e-3
    assert flips <= 2e-3 * N * H * W * C * L
    e_dx = rel(dx.cpu().numpy(), dX)
    e_g = max(rel(grad[l].cpu().numpy(), grads[l]) for l in range(L))
    print("chain f16 C=%d %dx%d N=%d L=%d h=%g: y %.2e dx %.2e wgrad %.2e flips %d" % (C, H, W, N, L, h, rel(y.cpu().numpy(), xs[L]), e_dx, e_g, flips))
    assert e_dx <= 1e-3
    assert e_g <= 5e-3


@pytest.mark.parametrize("scale", [2.0 ** -60, 2.0 ** -20, 1.0, 2.0 ** 20, 2.0 ** 40, 0.0])
def test_chain_f16_gradient_scale_invariance(scale):
    """The fp16 backward strips carry one power-of-two scale per launch derived from max|dy|: for power-of-two
    factors the gradients must be EXACTLY linear in dy over 30 decades (every operation commutes with the factor),
    and zero for dy == 0."""
    C, H, W, N, L, gamma, h = 32, 16, 16, 3, 4, -0.1, 0.25
    ch, flats, theta = _setup(C, L, gamma, seed=3)
    g = torch.Generator().manual_seed(9)
    x = torch.relu(torch.randn((N, H, W, C), generator=g))
    dy = torch.randn((N, H, W, C), generator=g)
    _, _, _, _, dx1, g1 = _run_gpu(ch, x, dy * 2.0 ** -3, h, L)       # power-of-two reference scale
    _, _, _, _, dxs, gs = _run_gpu(ch, x, dy * (2.0 ** -3 * scale), h, L)
    if scale == 0.0:
        assert float(dxs.abs().max()) == 0.0 and float(gs.abs().max()) == 0.0
        return
    assert torch.equal(dxs / scale, dx1)
    assert torch.equal(gs / scale, g1)


@pytest.mark.parametrize("C,H,W,h,gamma,n", [(16, 32, 32, 0.01, -0.1, 200), (16, 32, 32, 0.125, 0.0, 1000),
                                             (64, 8, 8, 0.01, -0.1, 1000)])
def test_chain_f16_long_horizon_shared_weights(C, H, W, h, gamma, n):
    """BASELINE cfg5: one block applied n times with shared weights (n_layers == 1, n_steps = n), ONE launch."""
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
    drift = float(y.double().norm()) / np.linalg.norm(cur)
    print("cfg5 f16 C=%d h=%g gamma=%g n=%d: |x_n| oracle %.4e gpu %.4e rel err %.2e drift %.5f" % (
        C, h, gamma, n, np.linalg.norm(cur), float(y.double().norm()), err, drift))
    assert abs(drift - 1.0) <= 2e-2, drift
    if gamma < 0 and h <= 0.01:
        assert err <= 2e-3


def test_chain_f16_refuses_what_does_not_fit():
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    assert not ChainHandle.supported(64, 32, 32, _abi.PREC_FAST_F16)
    assert not ChainHandle.supported(16, 64, 64, _abi.PREC_FAST_F16)
    ch = ChainHandle(64, 1, 0.0, precision=_abi.PREC_FAST_F16)
    ch.pack(torch.zeros(ch.num_params, device="cuda"))
    x = torch.zeros((1, 32, 32, 64), device="cuda")
    with pytest.raises(ValueError):
        ch.forward(x, 0.1, y_final=torch.empty_like(x))
    with pytest.raises(ValueError):                                # FAST_F16 needs y_final
        ch.forward(torch.zeros((1, 8, 8, 64), device="cuda"), 0.1, acts=torch.empty((1, 1, 8, 8, 64), device="cuda", dtype=torch.float16))


def test_chain_f16_dgrad_with_supplied_amax_is_bit_identical():
    """b200ode_chain_dgrad_amax with max|dy| supplied by the caller (as the *_amax data-gradient entries of the head / transitions
    leave it): same dz_all, dx and weight gradients, bit for bit, as the entry that reduces dy itself."""
    C, H, W, N, L, h = 32, 16, 16, 5, 3, 0.125
    ch, flats, theta = _setup(C, L, -0.1, seed=21)
    g = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn((N, H, W, C), generator=g))
    dy = torch.randn((N, H, W, C), generator=g) * 3e-3
    acts, masks, y, dz, dx, grad = _run_gpu(ch, x, dy, h, L)
    dyd = dy.cuda()
    am = dyd.abs().max().reshape(1).clone()
    dz2, dx2, grad2 = torch.empty_like(dz), torch.empty_like(dx), torch.empty_like(grad)
    ch.dgrad(dyd, masks, dz2, dx2, h, dy_amax=am)
    ch.wgrad(x.cuda(), acts, dz2, grad2.view(-1))
    torch.cuda.synchronize()
    assert torch.equal(dz, dz2) and torch.equal(dx, dx2) and torch.equal(grad, grad2)
