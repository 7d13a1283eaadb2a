"""Out-of-bounds write check with guard bands (compute-sanitizer is closed on this GPU pool --
profiles/r02_compute_sanitizer_closed_on_pool.txt -- so memcheck's job is done by hand for the hand-rolled TMA / TMEM /
mbarrier kernels): every output of a compute call is carved out of the MIDDLE of a larger buffer pre-filled with a
sentinel; after the call the bands on both sides must be untouched and the result must equal the run on tight buffers
(a write through a stale pointer or past a tile's last row would show up in one of the two).  Inputs get bands of NaNs:
reading them (instead of the TMA zero fill / the in-bounds data) poisons the output."""
import numpy as np
import pytest
import torch

from test_gpu_parity import make_layer, rand_x

pytestmark = pytest.mark.gpu
BAND = 4096          # elements on each side


def banded(shape, dtype, fill):
    n = int(np.prod(shape))
    buf = torch.full((n + 2 * BAND,), fill, dtype=dtype, device="cuda")
    return buf, buf[BAND:BAND + n].view(shape)


def bands_intact(buf, n, fill):
    lo, hi = buf[:BAND], buf[BAND + n:]
    if fill != fill:       # NaN sentinel
        return bool(torch.isnan(lo.float()).all()) and bool(torch.isnan(hi.float()).all())
    return bool((lo == fill).all()) and bool((hi == fill).all())


@pytest.mark.parametrize("precision,shape", [("strict", (3, 9, 7, 16)), ("fast_tf32", (2, 16, 16, 32)), ("fast_bf16", (2, 12, 10, 64)),
                                             ("strict", (1, 8, 8, 128)), ("fast_bf16", (1, 6, 9, 256)), ("fast_tf32", (5, 5, 33, 64))])
def test_layer_calls_stay_inside_their_buffers(precision, shape):
    import ctypes
    from differential_equations_resnet_b200 import _abi
    N, H, W, C = shape
    layer = make_layer(C, precision)
    hd = layer._handle
    hd.pack(layer.packed.detach())
    dt = torch.bfloat16 if precision == "fast_bf16" else torch.float32
    n = N * H * W * C
    nan = float("nan")
    xb, x = banded(shape, dt, nan)
    dyb, dy = banded(shape, dt, nan)
    x.copy_(rand_x(shape, 1, precision, relu_like=True)[0]); dy.copy_(rand_x(shape, 2, precision)[0])
    yb, y = banded(shape, dt, 7.0)
    mb, m = banded((N, H, W, C // 8), torch.uint8, 0x5A)
    dzb, dz = banded(shape, dt, 7.0)
    dxb, dx = banded(shape, dt, 7.0)
    gb, g = banded((hd.num_params,), torch.float32, 7.0)
    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _abi.check(lib.b200ode_euler_fwd(hd._h, P(x), P(y), P(m), None, N, H, W, 0.25, _abi.F_EULER, st))
    _abi.check(lib.b200ode_relu_scale_bwd(P(dy), P(m), P(dz), N * H * W, C, 0.25, int(dt == torch.bfloat16), st))
    _abi.check(lib.b200ode_euler_dgrad(hd._h, P(dz), P(dy), P(dx), N, H, W, st))
    _abi.check(lib.b200ode_euler_wgrad(hd._h, P(x), P(dz), P(g), None, N, H, W, 0, st))
    torch.cuda.synchronize()
    assert bands_intact(yb, n, 7.0) and bands_intact(mb, n // 8, 0x5A) and bands_intact(dzb, n, 7.0)
    assert bands_intact(dxb, n, 7.0) and bands_intact(gb, hd.num_params, 7.0)
    assert bands_intact(xb, n, nan) and bands_intact(dyb, n, nan)
    for t in (y, dz, dx, g):
        assert bool(torch.isfinite(t.float()).all()), "a NaN guard band of an input leaked into the result"
    # same calls on tight buffers give the same bits
    y2, m2, dz2, dx2 = torch.empty_like(y), torch.empty_like(m), torch.empty_like(dz), torch.empty_like(dx)
    g2 = torch.empty_like(g)
    xt, dyt = x.clone(), dy.clone()
    _abi.check(lib.b200ode_euler_fwd(hd._h, P(xt), P(y2), P(m2), None, N, H, W, 0.25, _abi.F_EULER, st))
    _abi.check(lib.b200ode_relu_scale_bwd(P(dyt), P(m2), P(dz2), N * H * W, C, 0.25, int(dt == torch.bfloat16), st))
    _abi.check(lib.b200ode_euler_dgrad(hd._h, P(dz2), P(dyt), P(dx2), N, H, W, st))
    _abi.check(lib.b200ode_euler_wgrad(hd._h, P(xt), P(dz2), P(g2), None, N, H, W, 0, st))
    torch.cuda.synchronize()
    assert torch.equal(y, y2) and torch.equal(m, m2) and torch.equal(dx, dx2) and torch.equal(g, g2)


@pytest.mark.parametrize("prec_name,C,HW,N", [("fast_f16", 16, 32, 3), ("fast_f16", 32, 16, 5), ("fast_f16", 64, 8, 6),
                                              ("fast_tf32", 16, 32, 2), ("fast_tf32", 64, 8, 9),
                                              ("strict", 16, 32, 2), ("strict", 32, 16, 5), ("strict", 64, 8, 9)])
def test_chain_calls_stay_inside_their_buffers(prec_name, C, HW, N):
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    L = 3
    ch = ChainHandle(C, L, -0.1, precision=_abi.CHAIN_PRECISIONS[prec_name])
    theta = (torch.randn(L, ch.num_params, generator=torch.Generator().manual_seed(0)) * 0.1).cuda()
    ch.pack(theta.view(-1))
    shape = (N, HW, HW, C)
    n = N * HW * HW * C
    nan = float("nan")
    sdt = ch.saved_dtype
    xb, x = banded(shape, torch.float32, nan)
    dyb, dy = banded(shape, torch.float32, nan)
    x.copy_(torch.relu(torch.randn(shape, generator=torch.Generator().manual_seed(1))).cuda())
    dy.copy_(torch.randn(shape, generator=torch.Generator().manual_seed(2)).cuda())
    ab, acts = banded((L,) + shape, sdt, 7.0)
    mb, masks = banded((L, N, HW, HW, C // 8), torch.uint8, 0x5A)
    yb, y = banded(shape, torch.float32, 7.0)
    zb, dz = banded((L,) + shape, sdt, 7.0)
    dxb, dx = banded(shape, torch.float32, 7.0)
    gb, grad = banded((L * ch.num_params,), torch.float32, 7.0)
    ch.forward(x, 0.125, acts=acts, masks=masks, y_final=y if ch.f16 else None)
    ch.dgrad(dy, masks, dz, dx, 0.125)
    ch.wgrad(x, acts, dz, grad)
    torch.cuda.synchronize()
    assert bands_intact(ab, L * n, 7.0) and bands_intact(mb, L * n // 8, 0x5A) and bands_intact(zb, L * n, 7.0)
    assert bands_intact(dxb, n, 7.0) and bands_intact(gb, L * ch.num_params, 7.0)
    assert (not ch.f16) or bands_intact(yb, n, 7.0)
    assert bands_intact(xb, n, nan) and bands_intact(dyb, n, nan)
    for t in (acts, dz, dx, grad):
        assert bool(torch.isfinite(t.float()).all())


@pytest.mark.parametrize("Ci,Co,H,W,N", [(16, 32, 32, 32, 3), (32, 64, 16, 16, 5), (16, 32, 9, 11, 2), (32, 64, 7, 10, 150)])
def test_transition_calls_stay_inside_their_buffers(Ci, Co, H, W, N):
    """The tensor-core transition kernels (kernels_glue_mma.cuh: cp.async staging, shared-memory gathers with a zero pixel for
    out-of-image taps, float2 stores per MMA fragment): outputs between sentinel bands, inputs between NaN bands."""
    import ctypes
    from differential_equations_resnet_b200 import _abi
    lib = _abi.lib()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    g = torch.Generator().manual_seed(4)
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    nan = float("nan")
    xb, x = banded((N, H, W, Ci), torch.float32, nan)
    x.copy_(torch.randn((N, H, W, Ci), generator=g).cuda())
    db, dout = banded((N, Ho, Wo, Co), torch.float32, nan)
    dout.copy_(torch.randn((N, Ho, Wo, Co), generator=g).cuda())
    wb, Wm = banded((3, 3, Ci, Co), torch.float32, nan)
    Wm.copy_((torch.randn((3, 3, Ci, Co), generator=g) * 0.1).cuda())
    sb, Ws = banded((Ci, Co), torch.float32, nan)
    Ws.copy_((torch.randn((Ci, Co), generator=g) * 0.1).cuda())
    bm = (torch.randn(Co, generator=g) * 0.1).cuda(); bs = (torch.randn(Co, generator=g) * 0.1).cuda()
    ob, out = banded((N, Ho, Wo, Co), torch.float32, 7.0)
    mb, mask = banded((N, Ho, Wo, Co // 8), torch.uint8, 0x5A)
    dxb, dx = banded((N, H, W, Ci), torch.float32, 7.0)
    npar = 9 * Ci * Co + Co + Ci * Co + Co
    pb, dp = banded((npar,), torch.float32, 7.0)
    _abi.check(lib.b200ode_transition_fwd(P(x), P(Wm), P(bm), P(Ws), P(bs), P(out), P(mask), N, H, W, Ci, Co, 2, 2, None))
    _abi.check(lib.b200ode_transition_dgrad(P(dout), P(mask), P(Wm), P(Ws), P(dx), N, H, W, Ci, Co, 2, 2, None))
    _abi.check(lib.b200ode_transition_wgrad(P(x), P(dout), P(mask), P(dp), N, H, W, Ci, Co, 2, 2, None, 0, None))
    torch.cuda.synchronize()
    assert bands_intact(ob, out.numel(), 7.0) and bands_intact(mb, mask.numel(), 0x5A)
    assert bands_intact(dxb, dx.numel(), 7.0) and bands_intact(pb, npar, 7.0)
    assert bands_intact(xb, x.numel(), nan) and bands_intact(db, dout.numel(), nan)
    assert bands_intact(wb, Wm.numel(), nan) and bands_intact(sb, Ws.numel(), nan)
    for t in (out, dx, dp):
        assert bool(torch.isfinite(t).all()), "a NaN guard band of an input leaked into the result"
    # tight buffers: same bits
    out2, mask2, dx2, dp2 = torch.empty_like(out), torch.empty_like(mask), torch.empty_like(dx), torch.empty_like(dp)
    xt, dt_, Wt, St = x.clone(), dout.clone(), Wm.clone(), Ws.clone()
    _abi.check(lib.b200ode_transition_fwd(P(xt), P(Wt), P(bm), P(St), P(bs), P(out2), P(mask2), N, H, W, Ci, Co, 2, 2, None))
    _abi.check(lib.b200ode_transition_dgrad(P(dt_), P(mask2), P(Wt), P(St), P(dx2), N, H, W, Ci, Co, 2, 2, None))
    _abi.check(lib.b200ode_transition_wgrad(P(xt), P(dt_), P(mask2), P(dp2), N, H, W, Ci, Co, 2, 2, None, 0, None))
    torch.cuda.synchronize()
    assert torch.equal(out, out2) and torch.equal(mask, mask2) and torch.equal(dx, dx2) and torch.equal(dp, dp2)
