"""GPU parity of the stem / transition / head kernels (b200ode_stem_*, b200ode_transition_*,
b200ode_head_fwd_bwd; reference: models/tfkeras_resnets.py:204-269, 555-572, 595-597 and
training/training.py:295) against the O1 oracle's torch-CPU fp32 ops with autograd.
Tolerance: fp32 FMA kernels vs fp32 oneDNN: 1e-5 relative (summation order only)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import antisym_torch as O1

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def _lib():
    from differential_equations_resnet_b200 import _abi
    return _abi, _abi.lib()


def P(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("u8", [True, False])
def test_stem_matches_oracle(u8):
    _abi, lib = _lib()
    N, H, W, Ci, Co = 5, 12, 10, 3, 16
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (N, H, W, Ci), generator=g, dtype=torch.uint8)
    if not u8:
        img = img.float() + 0.25
    K = (torch.randn((3, 3, Ci, Co), generator=g) * 0.2).requires_grad_(True)
    b = (torch.randn(Co, generator=g) * 0.1).requires_grad_(True)
    dout = torch.randn((N, H, W, Co), generator=g)
    x = (img.float() - 127.5) / 127.5
    ref = torch.relu(O1.conv2d_same_nhwc(x, K, (1, 1)) + b)
    ref.backward(dout)
    out = torch.empty((N, H, W, Co), device="cuda")
    imgd, Kd, bd, dd = img.cuda(), K.detach().cuda(), b.detach().cuda(), dout.cuda()   # keep the device copies alive
    _abi.check(lib.b200ode_stem_fwd(P(imgd), int(u8), 127.5, 127.5, 1, P(Kd), P(bd), P(out), N, H, W, Ci, Co, None))
    dp = torch.empty(27 * Co + Co, device="cuda")
    _abi.check(lib.b200ode_stem_wgrad(P(imgd), int(u8), 127.5, 127.5, 1, P(out), P(dd), P(dp), N, H, W, Ci, Co, None, 0, None))
    torch.cuda.synchronize()
    assert rel(out, ref) <= 1e-5
    assert rel(dp[:27 * Co].view(3, 3, Ci, Co), K.grad) <= 1e-5
    assert rel(dp[27 * Co:], b.grad) <= 1e-5


@pytest.mark.parametrize("Ci,Co,H,W,st,N", [(16, 32, 32, 32, (2, 2), 3), (32, 64, 16, 16, (2, 2), 3), (16, 32, 9, 11, (2, 2), 3),
                                            (32, 64, 8, 8, (1, 1), 3), (32, 64, 7, 10, (2, 2), 2), (16, 32, 64, 64, (2, 2), 1),
                                            (32, 64, 16, 16, (2, 2), 150), (16, 32, 12, 12, (2, 1), 2)])
def test_transition_matches_oracle(Ci, Co, H, W, st, N):
    """Stride-2 transitions 16 -> 32 / 32 -> 64 run on the tensor-core kernels (kernels_glue_mma.cuh: 3xTF32 warp MMAs, fp32-grade),
    every other shape on the CUDA-core kernels; N = 150 makes a weight-gradient block walk two images, odd sizes exercise
    pad_before = 1 and ragged tiles."""
    _abi, lib = _lib()
    g = torch.Generator().manual_seed(1)
    x = torch.randn((N, H, W, Ci), generator=g).requires_grad_(True)
    Wm = (torch.randn((3, 3, Ci, Co), generator=g) * 0.1).requires_grad_(True)
    bm = (torch.randn(Co, generator=g) * 0.1).requires_grad_(True)
    Ws = (torch.randn((1, 1, Ci, Co), generator=g) * 0.1).requires_grad_(True)
    bs = (torch.randn(Co, generator=g) * 0.1).requires_grad_(True)
    main = O1.conv2d_same_nhwc(x, Wm, st) + bm
    short = O1.conv2d_same_nhwc(x, Ws, st) + bs
    ref = torch.relu(main) + short
    dout = torch.randn(ref.shape, generator=g)
    Ho, Wo = ref.shape[1], ref.shape[2]
    out = torch.empty((N, Ho, Wo, Co), device="cuda")
    mask = torch.empty((N, Ho, Wo, Co // 8), dtype=torch.uint8, device="cuda")
    xd, Wmd, Wsd, bmd, bsd = x.detach().cuda(), Wm.detach().cuda(), Ws.detach().cuda(), bm.detach().cuda(), bs.detach().cuda()
    _abi.check(lib.b200ode_transition_fwd(P(xd), P(Wmd), P(bmd), P(Wsd), P(bsd), P(out), P(mask),
                                          N, H, W, Ci, Co, st[0], st[1], None))
    dx = torch.empty((N, H, W, Ci), device="cuda")
    dd = dout.cuda()
    _abi.check(lib.b200ode_transition_dgrad(P(dd), P(mask), P(Wmd), P(Wsd), P(dx), N, H, W, Ci, Co, st[0], st[1], None))
    nm = 9 * Ci * Co
    dp = torch.empty(nm + Co + Ci * Co + Co, device="cuda")
    _abi.check(lib.b200ode_transition_wgrad(P(xd), P(dd), P(mask), P(dp), N, H, W, Ci, Co, st[0], st[1], None, 0, None))
    torch.cuda.synchronize()
    assert rel(out, ref) <= 1e-5
    # backward reference along the relu branches the GPU took (its mask): a pre-activation within rounding of 0 may land on
    # either side, and one flipped bit in 10^5 would otherwise dominate a 1e-5 gradient comparison
    taken = torch.from_numpy(np.unpackbits(mask.cpu().numpy(), axis=-1, bitorder="little").astype(bool))
    (torch.where(taken, main, torch.zeros_like(main)) + short).backward(dout)
    bits = np.unpackbits(mask.cpu().numpy(), axis=-1, bitorder="little").astype(bool)
    mref = main.detach().numpy()
    differs = bits != (mref > 0)
    assert not (differs & (np.abs(mref) > 1e-5)).any()      # a relu bit may only differ where the pre-activation is ~0
    assert differs.sum() <= 2
    assert rel(dx, x.grad) <= 1e-5
    assert rel(dp[:nm].view(3, 3, Ci, Co), Wm.grad) <= 1e-5
    assert rel(dp[nm:nm + Co], bm.grad) <= 1e-5
    assert rel(dp[nm + Co:nm + Co + Ci * Co].view(1, 1, Ci, Co), Ws.grad) <= 1e-5
    assert rel(dp[nm + Co + Ci * Co:], bs.grad) <= 1e-5


def test_head_matches_oracle():
    _abi, lib = _lib()
    N, H, W, C, K = 7, 8, 8, 64, 10
    g = torch.Generator().manual_seed(2)
    x = torch.randn((N, H, W, C), generator=g).requires_grad_(True)
    Wfc = (torch.randn((C, K), generator=g) * 0.5).requires_grad_(True)
    bfc = (torch.randn(K, generator=g) * 0.1).requires_grad_(True)
    onehot = torch.nn.functional.one_hot(torch.randint(0, K, (N,), generator=g), K).float()
    probs_ref = torch.softmax(x.mean(dim=(1, 2)) @ Wfc + bfc, dim=-1)
    loss_ref = O1.loss_fn(probs_ref, onehot)
    loss_ref.backward()
    probs = torch.empty((N, K), device="cuda"); loss = torch.zeros(1, device="cuda")
    dx = torch.empty((N, H, W, C), device="cuda"); dp = torch.empty(C * K + K, device="cuda")
    xd, Wd, bd, od = x.detach().cuda(), Wfc.detach().cuda(), bfc.detach().cuda(), onehot.cuda()   # keep alive
    _abi.check(lib.b200ode_head_fwd_bwd(P(xd), P(Wd), P(bd), P(od), 1e-7, P(probs), P(loss), P(dx), P(dp), N, H * W, C, K, None, 0, None))
    torch.cuda.synchronize()
    assert rel(probs, probs_ref) <= 1e-5
    assert abs(float(loss) - float(loss_ref)) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert rel(dx, x.grad) <= 1e-4
    assert rel(dp[:C * K].view(C, K), Wfc.grad) <= 1e-4
    assert rel(dp[C * K:], bfc.grad) <= 1e-4


@pytest.mark.parametrize("Ci,Co,H,st", [(16, 32, 32, (2, 2)), (32, 64, 16, (2, 2)), (32, 64, 8, (1, 1))])
def test_amax_entries_leave_the_maximum_of_dx(Ci, Co, H, st):
    """b200ode_transition_dgrad_amax / b200ode_head_fwd_bwd_amax: same results as the plain entries, and the caller's scalar
    holds max|dx| afterwards (tensor-core kernels: from their epilogue; other shapes: one more launch).  Feeds
    b200ode_chain_dgrad_amax (tests/test_gpu_chain_f16.py)."""
    _abi, lib = _lib()
    N = 4
    g = torch.Generator().manual_seed(9)
    Ho = (H + st[0] - 1) // st[0]
    dout = torch.randn((N, Ho, Ho, Co), generator=g).cuda()
    mask = torch.randint(0, 256, (N, Ho, Ho, Co // 8), generator=g, dtype=torch.uint8).cuda()
    Wm = (torch.randn((3, 3, Ci, Co), generator=g) * 0.1).cuda(); Ws = (torch.randn((Ci, Co), generator=g) * 0.1).cuda()
    dx, dx2 = torch.empty((N, H, H, Ci), device="cuda"), torch.empty((N, H, H, Ci), device="cuda")
    am = torch.zeros(1, device="cuda")
    _abi.check(lib.b200ode_transition_dgrad(P(dout), P(mask), P(Wm), P(Ws), P(dx), N, H, H, Ci, Co, st[0], st[1], None))
    _abi.check(lib.b200ode_transition_dgrad_amax(P(dout), P(mask), P(Wm), P(Ws), P(dx2), N, H, H, Ci, Co, st[0], st[1], P(am), None))
    torch.cuda.synchronize()
    assert torch.equal(dx, dx2)
    assert float(am) == float(dx.abs().max())
    # head
    C, K = 64, 10
    x = torch.randn((N, 8, 8, C), generator=g).cuda()
    Wfc = (torch.randn((C, K), generator=g) * 0.5).cuda(); bfc = torch.zeros(K, device="cuda")
    oh = torch.nn.functional.one_hot(torch.randint(0, K, (N,), generator=g), K).float().cuda()
    loss, loss2 = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    hx, hx2 = torch.empty_like(x), torch.empty_like(x)
    dp, dp2 = torch.empty(C * K + K, device="cuda"), torch.empty(C * K + K, device="cuda")
    am.zero_()
    _abi.check(lib.b200ode_head_fwd_bwd(P(x), P(Wfc), P(bfc), P(oh), 1e-7, None, P(loss), P(hx), P(dp), N, 64, C, K, None, 0, None))
    _abi.check(lib.b200ode_head_fwd_bwd_amax(P(x), P(Wfc), P(bfc), P(oh), 1e-7, None, P(loss2), P(hx2), P(dp2), N, 64, C, K, None, 0, P(am), None))
    torch.cuda.synchronize()
    assert torch.equal(hx, hx2) and torch.equal(dp, dp2) and torch.equal(loss, loss2)
    assert float(am) == float(hx.abs().max())


@pytest.mark.parametrize("N,H,W,Ci,Co,st", [(4, 32, 32, 16, 32, (2, 2)), (5, 16, 16, 32, 64, (2, 2)), (2, 9, 11, 16, 32, (2, 2)),
                                            (150, 7, 10, 32, 64, (2, 2)), (3, 8, 8, 32, 64, (1, 1))])
def test_transition_fast_entries_within_tf32_tolerance(N, H, W, Ci, Co, st):
    """b200ode_transition_{fwd,dgrad,wgrad}_fast: one tf32 MMA on round-to-nearest operands (the fast modes' grade).
    Stated tolerance 2e-3 relative (measured ~4e-4: two 2^-11 roundings per product, K = 9*Cin + Cin terms); shapes off the
    tensor-core path (stride 1 here) run the fp32 kernels and stay at 1e-5; dx_amax = max|dx| exactly."""
    _abi, lib = _lib()
    g = torch.Generator().manual_seed(3)
    x = torch.randn((N, H, W, Ci), generator=g).requires_grad_(True)
    Wm = (torch.randn((3, 3, Ci, Co), generator=g) * 0.1).requires_grad_(True)
    bm = (torch.randn(Co, generator=g) * 0.1).requires_grad_(True)
    Ws = (torch.randn((1, 1, Ci, Co), generator=g) * 0.1).requires_grad_(True)
    bs = (torch.randn(Co, generator=g) * 0.1).requires_grad_(True)
    main = O1.conv2d_same_nhwc(x, Wm, st) + bm
    short = O1.conv2d_same_nhwc(x, Ws, st) + bs
    ref = torch.relu(main) + short
    dout = torch.randn(ref.shape, generator=g)
    Ho, Wo = ref.shape[1], ref.shape[2]
    out = torch.empty((N, Ho, Wo, Co), device="cuda")
    mask = torch.empty((N, Ho, Wo, Co // 8), dtype=torch.uint8, device="cuda")
    xd, Wmd, Wsd, bmd, bsd = x.detach().cuda(), Wm.detach().cuda(), Ws.detach().cuda(), bm.detach().cuda(), bs.detach().cuda()
    _abi.check(lib.b200ode_transition_fwd_fast(P(xd), P(Wmd), P(bmd), P(Wsd), P(bsd), P(out), P(mask), N, H, W, Ci, Co, st[0], st[1], None))
    dx = torch.empty((N, H, W, Ci), device="cuda")
    dd = dout.cuda()
    am = torch.zeros(1, device="cuda")
    _abi.check(lib.b200ode_transition_dgrad_fast(P(dd), P(mask), P(Wmd), P(Wsd), P(dx), N, H, W, Ci, Co, st[0], st[1], P(am), None))
    dx_plain = torch.empty_like(dx)
    _abi.check(lib.b200ode_transition_dgrad_fast(P(dd), P(mask), P(Wmd), P(Wsd), P(dx_plain), N, H, W, Ci, Co, st[0], st[1], None, None))
    nm = 9 * Ci * Co
    dp = torch.empty(nm + Co + Ci * Co + Co, device="cuda")
    _abi.check(lib.b200ode_transition_wgrad_fast(P(xd), P(dd), P(mask), P(dp), N, H, W, Ci, Co, st[0], st[1], None, 0, None))
    torch.cuda.synchronize()
    tol = 2e-3 if st == (2, 2) else 1e-5
    e_out = rel(out, ref)
    bits = np.unpackbits(mask.cpu().numpy(), axis=-1, bitorder="little").astype(bool)
    mref = main.detach().numpy()
    differs = bits != (mref > 0)
    assert not (differs & (np.abs(mref) > 5e-3)).any()          # relu bits may flip only where the pre-activation is ~0 at this grade
    taken = torch.from_numpy(bits)
    (torch.where(taken, main, torch.zeros_like(main)) + short).backward(dout)
    errs = dict(out=e_out, dx=rel(dx, x.grad), dWm=rel(dp[:nm].view(3, 3, Ci, Co), Wm.grad), dbm=rel(dp[nm:nm + Co], bm.grad),
                dWs=rel(dp[nm + Co:nm + Co + Ci * Co].view(1, 1, Ci, Co), Ws.grad), dbs=rel(dp[nm + Co + Ci * Co:], bs.grad))
    print("transition fast", (N, H, W, Ci, Co, st), {k: "%.1e" % v for k, v in errs.items()})
    # the mask-dependent output can differ by a whole relu where a bit flipped; everything else within the stated tolerance
    assert errs["dx"] <= tol and errs["dWm"] <= tol and errs["dWs"] <= tol and errs["dbm"] <= tol and errs["dbs"] <= tol, errs
    assert e_out <= max(tol, 1e-2 * differs.mean() ** 0.5 + tol), errs
    assert torch.equal(dx, dx_plain) and float(am) == float(dx.abs().max())
