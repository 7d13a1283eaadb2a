"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the
reference-shaped Python layers -> ctypes C ABI -> libb200ode.so, against the CPU oracle.

Tolerances (stated per north star):
  strict (3xTF32, fp32 accumulate) and simt (fp32 FMA):  ||y - O0|| / ||O0|| <= 1e-5
  fast_tf32 (1xTF32 operands):                            <= 2e-3   (2^-11 operand truncation, x2 operands)
  fast_bf16 (bf16 operands and I/O):                      <= 1.5e-2 (2^-8 operand rounding + bf16 output rounding)
Kernel assembly (get_kernel) and its antisymmetry: bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0

pytestmark = pytest.mark.gpu

TOL = {"strict": 1e-5, "simt": 1e-5, "fast_tf32": 2e-3, "fast_bf16": 1.5e-2}


def _pkg():
    import differential_equations_resnet_b200 as pkg
    return pkg


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_layer(C, precision, gamma=-0.1, seed=0, bias_std=0.1):
    pkg = _pkg()
    layer = pkg.Conv2DAntisymmetric3By3(gamma=gamma, precision=precision, seed=seed)
    layer.build((None, None, None, C))
    if bias_std:
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            layer.packed[-C:] = (torch.randn(C, generator=g) * bias_std).cuda()
    return layer


def io_cast(t, precision):
    return t.to(torch.bfloat16) if precision == "fast_bf16" else t


def rand_x(shape, seed, precision, relu_like=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    if relu_like:
        x = torch.relu(x)
    x = io_cast(x, precision)
    return x.cuda(), x.float().numpy().astype(np.float64)


# ------------------------------------------------------------------------------------------ K1 ---
@pytest.mark.parametrize("C,precision", [(1, "simt"), (5, "simt"), (16, "strict"), (32, "fast_tf32"),
                                         (64, "strict"), (64, "fast_bf16"), (128, "strict")])
def test_pack_bit_exact_and_antisymmetric(C, precision):
    gamma = np.float32(-0.1)
    layer = make_layer(C, precision, gamma=float(gamma))
    flat = layer.packed.detach().cpu().numpy()
    K = layer.get_kernel()
    assert K.shape == (3, 3, C, C)
    assert np.array_equal(K, O0.assemble_kernel_3by3_closed(flat, C, gamma))
    if C <= 16:
        assert np.array_equal(K, O0.assemble_kernel_3by3_literal(O0.split_params_3by3(flat, C), C, gamma))
    S = K + np.transpose(K[::-1, ::-1], (0, 1, 3, 2))
    want = np.zeros_like(S)
    want[1, 1, np.arange(C), np.arange(C)] = gamma + gamma
    assert np.array_equal(S, want)
    w = layer.get_weights()
    assert len(w) == C + 4 and w[0].shape == (1, 1, 1, C) and w[-1].shape == (C,)
    layer.set_weights(w)
    assert np.array_equal(layer.get_kernel(), K)


def test_kernel_structure_golden_on_gpu(golden_dir):
    """v6 cell 26 golden through the CUDA pack kernel."""
    g = json.load(open(os.path.join(golden_dir, "kernel_structure_v6_cell26.json")))
    layer = make_layer(64, "strict", gamma=0.0, bias_std=0)
    w = layer.get_weights()
    w[4 + 10][:, :, 20] = np.array(g["K_31_10"], np.float32).reshape(3, 3)
    d = np.array(g["K_4_4"], np.float32).reshape(3, 3)
    w[0][0, 0, 0, 4], w[1][0, 0, 0, 4], w[2][0, 0, 0, 4], w[3][0, 0, 0, 4] = d[0, 0], d[0, 1], d[0, 2], d[1, 0]
    layer.set_weights(w)
    K = layer.get_kernel()
    assert np.array_equal(K[:, :, 10, 31], np.array(g["K_10_31"], np.float32).reshape(3, 3))
    assert np.array_equal(K[:, :, 4, 4], d)


def test_diagonal_blocks_golden_on_gpu(golden_dir):
    """Printed K[:,:,1,1] of trained antisymmetric layers (antisymmetric_conv_kernel.ipynb cell 15, v2.0 cell 16)
    through the CUDA pack kernel, for both layer classes (3By3: d at (1,0); general k=3: free scalar at (1,2))."""
    g = json.load(open(os.path.join(golden_dir, "diagonal_blocks.json")))
    pkg = _pkg()
    for cell in g["cells"]:
        for key in ("res2a_K_1_1", "res2d_K_1_1"):
            B = np.array(cell[key], np.float32).reshape(3, 3)
            layer = make_layer(16, "strict", gamma=0.0, bias_std=0)
            w = layer.get_weights()
            w[0][0, 0, 0, 1], w[1][0, 0, 0, 1], w[2][0, 0, 0, 1], w[3][0, 0, 0, 1] = B[0, 0], B[0, 1], B[0, 2], B[1, 0]
            layer.set_weights(w)
            assert np.array_equal(layer.get_kernel()[:, :, 1, 1], B), (cell["source"], key)
            lg = pkg.Conv2DAntisymmetric(3, gamma=0.0, antisymmetric=True, precision="strict", seed=1)
            lg.build((None, 8, 8, 16))
            wg = lg.get_weights()
            scal = [i for i, a in enumerate(wg) if a.size == 1][4:8]      # the four diagonal scalars of output channel 1
            for slot, val in zip(scal, (B[0, 0], B[0, 1], B[0, 2], B[1, 2])):
                wg[slot][...] = val
            lg.set_weights(wg)
            assert np.array_equal(lg.get_kernel()[:, :, 1, 1], B), (cell["source"], key, "general class")


@pytest.mark.parametrize("k,anti,C", [(3, True, 5), (3, False, 4), (5, True, 3), (3, True, 16)])
def test_general_layer_pack(k, anti, C):
    pkg = _pkg()
    layer = pkg.Conv2DAntisymmetric(k, gamma=0.2, antisymmetric=anti, precision="strict", seed=3)
    layer.build((None, 8, 8, C))
    flat = layer.packed.detach().cpu().numpy()
    want = O0.assemble_kernel_general_literal(O0.split_params_general(flat, C, k, anti), C, k, np.float32(0.2), anti)
    assert np.array_equal(layer.get_kernel(), want)
    x, x64 = rand_x((2, 8, 8, C), 5, "strict")
    y = layer(x).detach().cpu().numpy()
    assert rel(y, O0.layer_call(x64, want.astype(np.float64), flat[-C:].astype(np.float64))) <= 1e-5


def test_conv_known_answer_on_gpu(golden_dir):
    """antisymmetric_conv_kernel.ipynb cells 1-3 through the CUDA conv (1 channel -> CUDA-core path)."""
    g = json.load(open(os.path.join(golden_dir, "conv2d_known_answer.json")))
    pkg = _pkg()
    layer = pkg.Conv2DAntisymmetric(3, antisymmetric=False, use_bias=False, precision="simt")
    layer.build((1, 7, 7, 1))
    k = np.array(g["kernel_3x3"], np.float32).reshape(3, 3)
    # non-antisymmetric diag block is centrosymmetric, so feed the golden kernel through two layers:
    # K = Ks (centrosymmetric part) + Ka (anti part) and conv is linear in K.
    ks, ka = (k + k[::-1, ::-1]) / 2, (k - k[::-1, ::-1]) / 2
    x = torch.tensor(g["image_7x7"], dtype=torch.float32).reshape(1, 7, 7, 1).cuda()
    layer.set_weights([np.float32(ks[0, 0]).reshape(1, 1, 1, 1), np.float32(ks[0, 1]).reshape(1, 1, 1, 1),
                       np.float32(ks[0, 2]).reshape(1, 1, 1, 1), np.float32(ks[1, 1]).reshape(1, 1, 1, 1),
                       np.float32(ks[1, 2]).reshape(1, 1, 1, 1)])
    y = layer(x).detach()
    la = pkg.Conv2DAntisymmetric(3, antisymmetric=True, use_bias=False, precision="simt")
    la.build((1, 7, 7, 1))
    la.set_weights([np.float32(ka[0, 0]).reshape(1, 1, 1, 1), np.float32(ka[0, 1]).reshape(1, 1, 1, 1),
                    np.float32(ka[0, 2]).reshape(1, 1, 1, 1), np.float32(ka[1, 2]).reshape(1, 1, 1, 1)])
    y = (y + la(x).detach()).cpu().numpy().reshape(-1)
    assert np.abs(y - np.array(g["output_7x7"], np.float32)).max() < 2e-6


# --------------------------------------------------------------------------------------- K2 / K3 / K4
SHAPES = [
    (2, 8, 8, 16), (3, 12, 10, 32), (2, 32, 32, 64), (5, 16, 16, 32), (9, 8, 8, 64),
    (2, 16, 16, 128), (2, 8, 8, 256), (1, 33, 17, 16), (3, 7, 5, 5), (2, 4, 4, 64),
    (1, 32, 32, 256),   # BASELINE cfg2's widest layer at its real 32x32 extent (strict: two weight stages, M-block wgrad)
]


def _supported(C, precision):
    return precision == "simt" or C in (16, 32, 64, 128, 256)


@pytest.mark.parametrize("precision", ["simt", "strict", "fast_tf32", "fast_bf16"])
@pytest.mark.parametrize("shape", SHAPES)
def test_euler_step_forward_backward(shape, precision):
    N, H, W, C = shape
    if not _supported(C, precision):
        pytest.skip("tensor path needs C in {16,32,64,128,256}")
    if precision == "simt" and N * H * W * C > 200000:
        pytest.skip("CUDA-core reference path: small shapes only")
    gamma, h = -0.1, 0.125
    layer = make_layer(C, precision, gamma=gamma)
    flat = layer.packed.detach().cpu().numpy().astype(np.float64)
    K = O0.assemble_kernel_3by3_closed(flat, C, gamma)
    x, x64 = rand_x(shape, 11, precision, relu_like=True)
    dy, dy64 = rand_x(shape, 12, precision)
    tol = TOL[precision]

    # plain layer call (reference call(): conv + bias)
    z = layer(x).detach().float().cpu().numpy()
    z_ref = O0.layer_call(x64, K, flat[-C:])
    assert rel(z, z_ref) <= tol, "conv+bias"

    # fused Euler step and its backward
    xr = x.clone().requires_grad_(True)
    y = layer.euler_step(xr, h)
    y_ref, cache = O0.euler_step_fwd(x64, K, flat[-C:], h)
    assert rel(y.detach().float().cpu().numpy(), y_ref) <= tol, "euler fwd"
    if precision == "fast_bf16":
        hd = layer._handle
        from differential_equations_resnet_b200.layers._base import relu_scale_bwd
        _, mask, _ = hd.forward(x, h, 15, want_mask=True)
        dz = relu_scale_bwd(dy, mask, h)
        dx = hd.dgrad(dz, dy, (H, W)).float().cpu().numpy()
        dX, G, dbias, _, dZ = O0.euler_step_bwd(dy64, cache, K, h)
        assert rel(dx, dX) <= tol, "dgrad"
        return
    y.backward(dy)
    dX, G, dbias, _, dZ = O0.euler_step_bwd(dy64, cache, K, h)
    gflat = O0.fold_grad_3by3(G, C, dbias)
    # relu mask flips on |z| ~ 0 are measure-zero for strict; loose modes may flip a few
    # fast modes: the tensor core TRUNCATES fp32 operands to tf32 (biased, probe-verified) and a few
    # relu-mask bits flip where |z| ~ 1e-3; the folded gradient S = G - rot180(G)^T additionally cancels.
    gtol = {"strict": (1e-5, 1e-5), "simt": (1e-5, 1e-5), "fast_tf32": (1e-2, 5e-2)}[precision]
    assert rel(xr.grad.float().cpu().numpy(), dX) <= gtol[0], "dgrad"
    assert rel(layer.packed.grad.cpu().numpy(), gflat) <= gtol[1], "wgrad+fold"


@pytest.mark.parametrize("precision", ["strict", "fast_tf32", "fast_bf16"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 16), (3, 12, 10, 32), (2, 32, 32, 64), (2, 16, 16, 128), (2, 8, 8, 256),
                                   (3, 10, 12, 128), (2, 32, 32, 256), (1, 9, 64, 256)])
def test_dense_wgrad(shape, precision):
    N, H, W, C = shape
    if shape in [(3, 10, 12, 128), (2, 32, 32, 256), (1, 9, 64, 256)] and precision == "strict":
        pytest.skip("row-aligned wgrad tiles (ragged last tile, W=32 and W=64 pitches) exist in the bf16 / tf32 kernels only")
    layer = make_layer(C, precision)
    x, x64 = rand_x(shape, 21, precision)
    dz, dz64 = rand_x(shape, 22, precision)
    hd = layer._handle
    hd.pack(layer.packed.detach())
    if precision == "fast_bf16":
        import ctypes
        from differential_equations_resnet_b200 import _abi
        G = torch.empty((3, 3, C, C), device="cuda")
        g = torch.zeros(hd.num_params, device="cuda")
        rc = _abi.lib().b200ode_euler_wgrad(hd._h, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(dz.data_ptr()),
                                            ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(G.data_ptr()), N, H, W, 0,
                                            torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    else:
        g, G = hd.wgrad(x, dz, want_dense=True)
    G_ref = O0.conv_kernel_grad_stride1(x64, dz64)
    assert rel(G.cpu().numpy(), G_ref) <= TOL[precision]
    # folded gradient (off-diagonal tiles, diagonal scalars, bias column sums) against the oracle's fold of ITS dense gradient
    g_ref = O0.fold_grad_3by3(G_ref, C, dz64.sum(axis=(0, 1, 2)))
    assert rel(g.cpu().numpy(), g_ref) <= 4 * TOL[precision]
    assert rel(g[-C:].cpu().numpy(), g_ref[-C:]) <= TOL[precision], "bias gradient"


def test_strided_layer_simt():
    pkg = _pkg()
    layer = pkg.Conv2DAntisymmetric3By3(gamma=0.0, strides=(2, 2), precision="strict", seed=4)
    x, x64 = rand_x((2, 9, 8, 6), 3, "strict")
    xr = x.clone().requires_grad_(True)
    y = layer(xr)
    flat = layer.packed.detach().cpu().numpy().astype(np.float64)
    K = O0.assemble_kernel_3by3_closed(flat, 6, 0.0)
    y_ref = O0.layer_call(x64, K, flat[-6:], (2, 2))
    assert y.shape == y_ref.shape and rel(y.detach().cpu().numpy(), y_ref) <= 1e-5
    # backward against torch autograd of the O1 restatement
    from oracle import antisym_torch as O1
    tx = torch.from_numpy(x64).requires_grad_(True)
    tf_ = torch.from_numpy(flat).requires_grad_(True)
    ty = O1.conv2d_same_nhwc(tx, O1.assemble_closed(tf_, 6, 0.0), (2, 2)) + tf_[-6:]
    dy = torch.randn(ty.shape, generator=torch.Generator().manual_seed(9), dtype=torch.float64)
    ty.backward(dy)
    y.backward(dy.float().cuda())
    assert rel(xr.grad.cpu().numpy(), tx.grad.numpy()) <= 1e-5
    assert rel(layer.packed.grad.cpu().numpy(), tf_.grad.numpy()) <= 1e-5


def test_linearity_and_antisymmetry_at_full_size():
    """Size-independent properties at BASELINE cfg2 scale (N=256, 32x32, C=64): <x, Kx> = gamma<x,x>
    (skew part contributes nothing) and conv(ax+by) = a conv(x) + b conv(y)."""
    gamma = -0.1
    layer = make_layer(64, "strict", gamma=gamma, bias_std=0)
    x, _ = rand_x((256, 32, 32, 64), 1, "strict")
    y, _ = rand_x((256, 32, 32, 64), 2, "strict")
    with torch.no_grad():
        kx, ky = layer(x), layer(y)
        kxy = layer(2.0 * x - 3.0 * y)
    lin = (kxy - (2.0 * kx - 3.0 * ky)).double().norm() / kxy.double().norm()
    assert float(lin) <= 2e-5
    quad = (x.double() * kx.double()).sum() / (x.double() * x.double()).sum()
    assert abs(float(quad) - gamma) <= 1e-5


@pytest.mark.parametrize("precision", ["strict", "fast_tf32", "fast_bf16", "simt"])
@pytest.mark.parametrize("shape", [(2, 8, 8, 16), (3, 12, 10, 32), (2, 16, 16, 128), (1, 9, 64, 256)])
def test_dgrad_fused_tail_bit_identical(shape, precision):
    """b200ode_euler_dgrad_fused == b200ode_euler_dgrad followed by b200ode_relu_scale_bwd, bit for bit."""
    import ctypes
    from differential_equations_resnet_b200 import _abi
    N, H, W, C = shape
    if precision == "simt" and C > 32:
        pytest.skip("CUDA-core path: small shapes only")
    layer = make_layer(C, precision, gamma=-0.1)
    hd = layer._handle
    hd.pack(layer.packed.detach())
    dz, _ = rand_x(shape, 31, precision)
    dy, _ = rand_x(shape, 32, precision)
    mask = torch.randint(0, 256, (N, H, W, C // 8), dtype=torch.uint8, device="cuda")
    lib, st = _abi.lib(), torch.cuda.current_stream().cuda_stream
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    dx_a, dx_b, dzp_a, dzp_b = (torch.empty_like(dz) for _ in range(4))
    h = 0.37
    _abi.check(lib.b200ode_euler_dgrad(hd._h, P(dz), P(dy), P(dx_a), N, H, W, st))
    _abi.check(lib.b200ode_relu_scale_bwd(P(dx_a), P(mask), P(dzp_a), N * H * W, C, h, int(dz.dtype == torch.bfloat16), st))
    _abi.check(lib.b200ode_euler_dgrad_fused(hd._h, P(dz), P(dy), P(dx_b), P(mask), P(dzp_b), h, N, H, W, st))
    torch.cuda.synchronize()
    assert torch.equal(dx_a, dx_b)
    assert torch.equal(dzp_a.view(torch.int16 if dz.dtype == torch.bfloat16 else torch.int32),
                       dzp_b.view(torch.int16 if dz.dtype == torch.bfloat16 else torch.int32))
    assert float(dzp_a.float().abs().sum()) > 0


# --------------------------------------------------------------------------------------- general k on tcgen05
@pytest.mark.parametrize("k,shape,precision", [(5, (2, 8, 8, 16), "strict"), (5, (3, 12, 10, 32), "fast_tf32"), (7, (2, 9, 11, 16), "strict"),
                                               (5, (2, 16, 16, 64), "strict"), (5, (1, 32, 32, 128), "fast_tf32"), (7, (2, 8, 8, 32), "fast_tf32"),
                                               (5, (64, 32, 32, 16), "strict")])
def test_general_k_on_tensor_path(k, shape, precision):
    """Conv2DAntisymmetric with k = 5 / 7 (layers/tfkeras_layer_Conv2DAntisymmetric.py:90-175): k*k taps through the SAME
    tcgen05 kernel (halo pitch W + k/2, one tap per weight stage) forward and data gradient, weight gradient on the CUDA-core
    kernel; fused Euler step and its backward against O0 (float64)."""
    from differential_equations_resnet_b200 import _abi
    pkg = _pkg()
    N, H, W, C = shape
    gamma, h = -0.1, 0.25
    layer = pkg.Conv2DAntisymmetric(k, gamma=gamma, precision=precision, seed=7)
    layer.build((None, H, W, C))
    assert layer._handle.effective_mode == _abi.PRECISIONS[precision], "k = %d must stay on the tensor path" % k
    with torch.no_grad():
        layer.packed[-C:] = (torch.randn(C, generator=torch.Generator().manual_seed(1)) * 0.1).cuda()
    flat = layer.packed.detach().cpu().numpy()
    K32 = O0.assemble_kernel_general_literal(O0.split_params_general(flat, C, k, True), C, k, np.float32(gamma), True)
    assert np.array_equal(layer.get_kernel(), K32)
    K = K32.astype(np.float64)
    x, x64 = rand_x(shape, 21, precision, relu_like=True)
    dy, dy64 = rand_x(shape, 22, precision)
    xg = x.clone().requires_grad_(True)
    y = layer.euler_step(xg, h)
    y.backward(dy)
    torch.cuda.synchronize()
    tol = TOL[precision]
    y_ref, cache = O0.euler_step_fwd(x64, K, flat[-C:].astype(np.float64), h)
    assert rel(y.detach().cpu().numpy(), y_ref) <= tol, rel(y.detach().cpu().numpy(), y_ref)
    # plain call() too (conv + bias, no tail)
    assert rel(layer(x).detach().cpu().numpy(), O0.layer_call(x64, K, flat[-C:].astype(np.float64))) <= tol
    # backward with the relu branches the GPU took (y - x = h * relu(z) > 0): relu' is discontinuous, a tf32 pre-activation
    # at rounding level may legitimately sit on the other side of zero
    mask = (y.detach().cpu().numpy().astype(np.float64) - x64) > 0
    dX, G, dbias, _, _ = O0.euler_step_bwd(dy64, cache, K, h, mask=mask)
    gtol = 1e-5 if precision == "strict" else 1e-2
    assert rel(xg.grad.cpu().numpy(), dX) <= gtol, rel(xg.grad.cpu().numpy(), dX)
    # folded weight gradient: oracle = autograd-free fold of the dense gradient through the general layout
    S = G - np.transpose(G[::-1, ::-1, :, :], (0, 1, 3, 2))
    got = layer.packed.grad.cpu().numpy()
    assert rel(got[-C:], dbias) <= max(gtol, 2e-5)
    # spot-check the off-diagonal block of output channel 0: W_0[a,b,j] <-> S[a,b,1+j,0]
    from differential_equations_resnet_b200.layers.tfkeras_layer_Conv2DAntisymmetric import diag_slots
    nd = len(diag_slots(k, True))
    w0 = got[nd:nd + k * k * (C - 1)].reshape(k, k, C - 1)
    assert rel(w0, S[:, :, 1:, 0]) <= max(gtol, 2e-5), rel(w0, S[:, :, 1:, 0])
