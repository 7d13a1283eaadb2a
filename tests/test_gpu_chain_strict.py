"""GPU parity of the STRICT persistent Euler-step chains (chain_tc_kernel<C, DIR, ST = true>: 3xTF32, fp32-grade): n stacked
Euler steps (models/tfkeras_resnets.py:28-94 / :575-593 of the reference) forward, backward sweep and the layer-batched strict
weight gradient in ONE launch each, against the NumPy float64 oracle.

Tolerance: the fp32-accumulate mode of BASELINE.json's north_star -- 1e-5 relative for every step output, the data gradient
and every layer's folded weight gradient.  The oracle's backward runs along the relu branches the GPU took (its saved masks):
a pre-activation at rounding level may legitimately land on either side of 0; the number of such flips is bounded separately."""
import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _unpack(mask, C):
    return np.unpackbits(mask.cpu().numpy(), axis=-1, bitorder="little").astype(bool).reshape(mask.shape[:-1] + (C,))


def _run(C, H, W, N, L, gamma, h, seed):
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    assert ChainHandle.supported(C, H, W, _abi.PREC_STRICT)
    rng = np.random.default_rng(seed)
    flats = np.stack([O0.init_params_3by3(rng, C, bias_std=0.1) for _ in range(L)]).astype(np.float32)
    ch = ChainHandle(C, L, gamma, precision=_abi.PREC_STRICT)
    assert ch.num_params == flats.shape[1] and ch.saved_dtype == torch.float32
    theta = torch.from_numpy(flats).cuda().contiguous()
    ch.pack(theta.view(-1))
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
    return flats, x.numpy().astype(np.float64), dy.numpy().astype(np.float64), acts, masks, yfin, dz, dx, grad


@pytest.mark.parametrize("C,H,W,N,L,gamma", [
    (16, 32, 32, 5, 3, -0.1), (32, 16, 16, 4, 4, 0.0), (64, 8, 8, 6, 3, -0.1),
    (16, 8, 8, 3, 2, 0.0), (32, 9, 12, 2, 3, -0.1), (16, 6, 5, 151, 2, -0.1), (64, 4, 4, 2, 1, 0.0),
    (32, 16, 16, 130, 2, -0.05), (64, 8, 8, 129, 7, 0.0),
])
def test_strict_chain_matches_oracle(C, H, W, N, L, gamma):
    h = 0.125
    flats, x, dy, acts, masks, yfin, dz, dx, grad = _run(C, H, W, N, L, gamma, h, seed=C + H + L)
    Ks = [O0.assemble_kernel_3by3_closed(flats[l].astype(np.float64), C, gamma) for l in range(L)]
    cur, caches, flips = x, [], 0
    for l in range(L):
        y, cache = O0.euler_step_fwd(cur, Ks[l], flats[l, -C:].astype(np.float64), h)
        assert rel(acts[l].cpu().numpy(), y) <= TOL, ("step", l)
        taken = _unpack(masks[l], C)
        differs = taken != (cache["u"] > 0)
        assert not (differs & (np.abs(cache["u"]) > 1e-4)).any()
        flips += int(differs.sum())
        caches.append(cache); cur = y
    assert flips <= max(2, int(1e-4 * L * x.size)), flips
    assert torch.equal(yfin, acts[L - 1])
    d = dy
    for l in range(L - 1, -1, -1):
        dX, G, dbias, _, dZ = O0.euler_step_bwd(d, caches[l], Ks[l], h, mask=_unpack(masks[l], C))
        assert rel(dz[l].cpu().numpy(), dZ) <= TOL, ("dz", l)
        assert rel(grad[l].cpu().numpy(), O0.fold_grad_3by3(G, C, dbias)) <= TOL, ("wgrad", l)
        d = dX
    assert rel(dx.cpu().numpy(), d) <= TOL


def test_strict_chain_close_to_per_layer_strict_kernels():
    """The chain and the per-layer strict kernels split their operands differently (truncate + remainder in the chain's
    epilogue vs the converter warps of conv_tc_kernel) but both are fp32-grade: they agree to 1e-5."""
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import LayerHandle
    C, H, W, N, L, gamma, h = 32, 16, 16, 3, 3, -0.05, 8.0 / 108.0
    flats, x, dy, acts, masks, yfin, dz, dx, grad = _run(C, H, W, N, L, gamma, h, seed=11)
    cur = torch.from_numpy(x).float().cuda()
    for l in range(L):
        hd = LayerHandle(C, 3, gamma, (1, 1), True, True, _abi.PREC_STRICT, _abi.LAYOUT_3BY3)
        hd.pack(torch.from_numpy(flats[l]).cuda())
        y, m, _ = hd.forward(cur, h, flags=_abi.F_EULER, want_mask=True)
        assert rel(acts[l].cpu().numpy(), y.cpu().numpy()) <= TOL
        cur = y


def test_strict_chain_long_horizon():
    """BASELINE cfg5 in the strict mode: 1000 Euler steps of one block (shared weights) in ONE launch, state error against
    the float64 oracle (contractive dynamics: gamma < 0, small h)."""
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    C, H, W, N, h, gamma, n = 16, 32, 32, 4, 0.01, -0.1, 1000
    rng = np.random.default_rng(5)
    flat = O0.init_params_3by3(rng, C, bias_std=0.1).astype(np.float32)
    ch = ChainHandle(C, 1, gamma, precision=_abi.PREC_STRICT)
    ch.pack(torch.from_numpy(flat).cuda())
    x = torch.randn((N, H, W, C), generator=torch.Generator().manual_seed(1))
    y = torch.empty((N, H, W, C), device="cuda")
    ch.forward(x.cuda(), h, n_steps=n, y_final=y)
    torch.cuda.synchronize()
    K = O0.assemble_kernel_3by3_closed(flat.astype(np.float64), C, gamma)
    cur = x.numpy().astype(np.float64)
    for _ in range(n):
        cur, _ = O0.euler_step_fwd(cur, K, flat[-C:].astype(np.float64), h)
    err = rel(y.cpu().numpy(), cur)
    print("cfg5 strict C=%d n=%d: rel err %.2e" % (C, n, err))
    assert err <= 1e-4           # 1000 steps of fp32 rounding of the residual stream itself
