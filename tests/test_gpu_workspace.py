"""Workspace contract of the C ABI (include/b200ode.h "Conventions"): the caller owns the scratch of a
compute call (bound per handle or passed with the call, sized by the *_workspace_bytes queries); without a bound block
every call draws its own block from the stream-ordered pool, so concurrent calls on different streams -- also on ONE
packed handle -- do not interfere; nothing synchronises the device."""
import ctypes

import numpy as np
import pytest
import torch

from test_gpu_parity import make_layer, rand_x

pytestmark = pytest.mark.gpu


def _wgrad(hd, x, dz, stream=None):
    from differential_equations_resnet_b200 import _abi
    N, H, W, C = x.shape
    g = torch.empty(hd.num_params, dtype=torch.float32, device=x.device)
    st = (stream or torch.cuda.current_stream()).cuda_stream
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    _abi.check(_abi.lib().b200ode_euler_wgrad(hd._h, P(x), P(dz), P(g), None, N, H, W, 0, st))
    return g


@pytest.mark.parametrize("precision,C", [("strict", 32), ("fast_tf32", 64), ("fast_bf16", 128), ("simt", 8)])
def test_bound_workspace_equals_pool_and_small_block_is_refused(precision, C):
    from differential_equations_resnet_b200.layers._base import _alloc_workspace
    shape = (4, 16, 16, C)
    layer = make_layer(C, precision)
    hd = layer._handle
    hd.pack(layer.packed.detach())
    x, _ = rand_x(shape, 1, precision, relu_like=True)
    dz, _ = rand_x(shape, 2, precision)
    g_pool = _wgrad(hd, x, dz)                                  # no block bound: stream-ordered pool
    need = hd.workspace_bytes(*shape[:3])
    assert need > 0
    ws = _alloc_workspace(need, x.device)
    ws.fill_(0xFF)
    hd.bind_workspace(ws)
    g_bound = _wgrad(hd, x, dz)
    torch.cuda.synchronize()
    assert torch.equal(g_pool, g_bound)                         # same kernels, same reduction order
    assert int((ws != 0xFF).sum()) > 0                          # the bound block really was the scratch
    hd.bind_workspace(_alloc_workspace(256, x.device)[:256])
    with pytest.raises(ValueError, match="workspace too small"):
        _wgrad(hd, x, dz)
    hd.bind_workspace(None)
    assert torch.equal(_wgrad(hd, x, dz), g_pool)


@pytest.mark.parametrize("precision,C", [("fast_tf32", 32), ("strict", 64), ("fast_bf16", 128)])
def test_concurrent_streams_on_one_handle(precision, C):
    """Two weight gradients (different inputs) in flight on two streams through ONE packed handle: with the old
    process-wide scratch their split-K partials overwrote each other."""
    shape = (32, 32, 32, C)
    layer = make_layer(C, precision)
    hd = layer._handle
    hd.pack(layer.packed.detach())
    xa, _ = rand_x(shape, 3, precision, relu_like=True)
    xb, _ = rand_x(shape, 4, precision, relu_like=True)
    dza, _ = rand_x(shape, 5, precision)
    dzb, _ = rand_x(shape, 6, precision)
    ref_a, ref_b = _wgrad(hd, xa, dza), _wgrad(hd, xb, dzb)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(10):
        ga = _wgrad(hd, xa, dza, s1)
        gb = _wgrad(hd, xb, dzb, s2)
        ga2 = _wgrad(hd, xa, dza, s1)
        torch.cuda.synchronize()
        assert torch.equal(ga, ref_a) and torch.equal(gb, ref_b) and torch.equal(ga2, ref_a)


def test_chain_and_glue_workspace_queries():
    from differential_equations_resnet_b200 import _abi
    from differential_equations_resnet_b200.layers._base import ChainHandle
    ch = ChainHandle(16, 4, 0.0, precision=_abi.PREC_FAST_F16)
    n = ch.workspace_bytes(8, 32, 32)
    assert n > 0 and ch.workspace_bytes(0, 32, 32) == 0
    out = ctypes.c_size_t()
    lib = _abi.lib()
    _abi.check(lib.b200ode_glue_workspace_bytes(_abi.GLUE_STEM_WGRAD, 8, 32, 32, 3, 16, 1, 1, ctypes.byref(out)))
    assert out.value >= 8 * (27 * 16 + 16) * 4
    _abi.check(lib.b200ode_glue_workspace_bytes(_abi.GLUE_TRANSITION_WGRAD, 8, 32, 32, 16, 32, 2, 2, ctypes.byref(out)))
    assert out.value >= 8 * (10 * 16 * 32 + 64) * 4
    _abi.check(lib.b200ode_glue_workspace_bytes(_abi.GLUE_HEAD, 8, 1, 1, 64, 10, 1, 1, ctypes.byref(out)))
    assert out.value >= 8 * (64 * 10 + 11) * 4
    with pytest.raises(ValueError):
        _abi.check(lib.b200ode_glue_workspace_bytes(7, 8, 1, 1, 64, 10, 1, 1, ctypes.byref(out)))


def test_train_step_graph_uses_bound_workspaces():
    """EulerNet binds caller-owned workspaces everywhere: a captured step contains no allocation and replays bit-identically."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    spec = NetSpec(blocks_per_stage=(2, 2, 2), h=0.1)
    net = EulerNet(spec, precision="fast_f16", seed=0)
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=g, dtype=torch.uint8).cuda()
    oh = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=g), 10).float().cuda()
    net.capture(img, oh)
    th0 = net.theta.clone()
    l1 = float(net.train_step_graph())
    th1 = net.theta.clone()
    net.theta.copy_(th0); net.adam_m.zero_(); net.adam_v.zero_(); net.step_counter.fill_(1)
    l2 = float(net.train_step_graph())
    assert l1 == l2 and torch.equal(net.theta, th1) and np.isfinite(l1)
