"""CPU tests: pin the oracle against the reference's golden vectors and against
itself (literal loops == closed form, closed-form backward == autograd)."""
import json
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import antisym_numpy as O0
from oracle import antisym_torch as O1


def _load(golden_dir, name):
    return json.load(open(os.path.join(golden_dir, name)))


# ---------------------------------------------------------------- goldens ---

def test_conv2d_known_answer(golden_dir):
    """antisymmetric_conv_kernel.ipynb cells 1-3: tf.nn.conv2d is cross-correlation,
    SAME zero padding, NHWC/HWIO."""
    g = _load(golden_dir, "conv2d_known_answer.json")
    x = np.array(g["image_7x7"], np.float32).reshape(1, 7, 7, 1)
    k = np.array(g["kernel_3x3"], np.float32).reshape(3, 3, 1, 1)
    want = np.array(g["output_7x7"], np.float32).reshape(1, 7, 7, 1)
    got0 = O0.conv2d_same(x.astype(np.float64), k.astype(np.float64))
    got1 = O1.conv2d_same_nhwc(torch.from_numpy(x), torch.from_numpy(k)).numpy()
    assert np.abs(got0 - want).max() < 5e-7   # printed to 7-8 significant digits
    assert np.abs(got1 - want).max() < 5e-7


def test_kernel_structure_v6_cell26(golden_dir):
    """v6 cell 26: K[:,:,10,31] == -rot180(K[:,:,31,10]); diag block skew-centro, centre gamma=0."""
    g = _load(golden_dir, "kernel_structure_v6_cell26.json")
    C = 64
    rng = np.random.default_rng(0)
    flat = O0.init_params_3by3(rng, C)
    off = O0.offsets_3by3(C)
    free = np.array(g["K_31_10"], np.float32).reshape(3, 3)   # ci=31 > o=10: W_10[:,:,20]
    Wo = flat[off["W"][10]:off["W"][10] + 9 * 53].reshape(3, 3, 53)
    Wo[:, :, 20] = free
    d = np.array(g["K_4_4"], np.float32).reshape(3, 3)
    for name, (i, j) in {"a": (0, 0), "b": (0, 1), "c": (0, 2), "d": (1, 0)}.items():
        flat[off[name] + 4] = d[i, j]
    for K in (O0.assemble_kernel_3by3_literal(O0.split_params_3by3(flat, C), C, 0.0),
              O0.assemble_kernel_3by3_closed(flat, C, 0.0)):
        assert K.shape == tuple(g["shape"])
        assert np.array_equal(K[:, :, 31, 10], free)
        assert np.array_equal(K[:, :, 10, 31], np.array(g["K_10_31"], np.float32).reshape(3, 3))
        assert np.array_equal(K[:, :, 4, 4], d)


def test_kernel_structure_v6_cell41(golden_dir):
    """v6 cell 41: integer prototype, exact negate + rot180 for (1,0)/(0,1) and (234,14)/(14,234)."""
    g = _load(golden_dir, "kernel_structure_v6_cell41.json")
    for free, dep in (("K_1_0", "K_0_1"), ("K_234_14", "K_14_234")):
        f = np.array(g[free]).reshape(3, 3, 1)
        t = O0._anti_centrosymmetric_transpose(f)[:, :, 0]
        assert np.array_equal(t, np.array(g[dep]).reshape(3, 3))
    for dname in ("K_0_0", "K_100_100"):
        m = np.array(g[dname]).reshape(3, 3)
        off_centre = m + m[::-1, ::-1]
        off_centre[1, 1] = 0
        assert not off_centre.any()


def test_centrosymmetric_7x7(golden_dir):
    """v6 cell 35: off-centre entries of the k=7 prototype obey m[i,j] = -m[6-i,6-j];
    our general-k slot enumeration reproduces it."""
    g = _load(golden_dir, "centrosymmetric_7x7_v6_cell35.json")
    m = np.array(g["matrix_7x7"]).reshape(7, 7)
    slots = O0.diag_slots_general(7, True)
    assert len(slots) == 24          # 21 (j>i) + 3 (j==i, i<=2)
    scal = [m[i, j] for (i, j) in slots]
    rebuilt = O0._centrosymmetric_matrix(scal, 7, gamma=m[3, 3], antisymmetric=True, dtype=np.int64)
    assert np.array_equal(rebuilt, m)


def test_variable_count_and_order():
    """training/training.py:397-398 hard-codes 20 variables per layer at C=16."""
    C = 16
    flat = np.arange(O0.num_params_3by3(C), dtype=np.float32)
    v = O0.split_params_3by3(flat, C)
    assert len(v) == 20
    assert [tuple(x.shape) for x in v[:4]] == [(1, 1, 1, C)] * 4
    assert [tuple(x.shape) for x in v[4:19]] == [(3, 3, C - o - 1) for o in range(15)]
    assert v[19].shape == (C,)
    assert np.array_equal(O0.join_params(v), flat)
    for C, n in ((16, 1160), (32, 4624), (64, 18464), (128, 73792), (256, 295040)):
        assert O0.num_params_3by3(C) == n


# ------------------------------------------------------- self-consistency ---

@settings(max_examples=25, deadline=None)
@given(C=st.integers(1, 12), gamma=st.sampled_from([0.0, -0.1, 0.25]), seed=st.integers(0, 2**16))
def test_literal_equals_closed_and_antisymmetric(C, gamma, seed):
    rng = np.random.default_rng(seed)
    flat = rng.standard_normal(O0.num_params_3by3(C)).astype(np.float32)
    Kl = O0.assemble_kernel_3by3_literal(O0.split_params_3by3(flat, C), C, np.float32(gamma))
    Kc = O0.assemble_kernel_3by3_closed(flat, C, np.float32(gamma))
    assert np.array_equal(Kl, Kc)
    # K[a,b,ci,o] + K[2-a,2-b,o,ci] == 2*gamma*[centre][ci==o]  bit-exactly
    S = Kc + np.transpose(Kc[::-1, ::-1], (0, 1, 3, 2))
    want = np.zeros_like(S)
    want[1, 1, np.arange(C), np.arange(C)] = np.float32(gamma) + np.float32(gamma)
    assert np.array_equal(S, want)
    Kt = O1.assemble_closed(torch.from_numpy(flat), C, gamma).numpy()
    assert np.array_equal(Kt, Kc)
    Ktl = O1.assemble_literal(O1.split_params(torch.from_numpy(flat), C), C, gamma).numpy()
    assert np.array_equal(Ktl, Kc)


@pytest.mark.parametrize("k,anti", [(3, True), (3, False), (5, True)])
def test_general_layer_matches_3by3_law(k, anti):
    C = 5
    rng = np.random.default_rng(3)
    flat = rng.standard_normal(O0.num_params_general(C, k, anti)).astype(np.float64)
    K = O0.assemble_kernel_general_literal(O0.split_params_general(flat, C, k, anti), C, k, -0.2, anti)
    assert K.shape == (k, k, C, C)
    for ci in range(C):
        for o in range(C):
            if ci != o:   # off-diagonal: always negated rot180 (:139)
                assert np.array_equal(K[:, :, ci, o], -K[::-1, ::-1, o, ci])
    d = K[:, :, 2, 2]
    if anti:
        e = d + d[::-1, ::-1]
        assert e[k // 2, k // 2] == -0.4
        e[k // 2, k // 2] = 0
        assert not e.any()
    else:
        assert np.array_equal(d, d[::-1, ::-1])
    if k == 3 and anti:   # same diagonal law as the 3By3 class with d at (1,0) = -v12
        v = O0.split_params_general(flat, C, k, anti)[2 * 5:]   # o=2: (4 scalars + W) per o
        assert d[0, 0] == v[0].item() and d[1, 2] == v[3].item() and d[1, 0] == -v[3].item()


def test_convolution_matrix_antisymmetry():
    """A + A^T = 2 gamma I for the doubly block Toeplitz matrix."""
    C, H, W, gamma = 3, 4, 5, -0.3
    rng = np.random.default_rng(1)
    flat = rng.standard_normal(O0.num_params_3by3(C))
    K = O0.assemble_kernel_3by3_closed(flat, C, gamma)
    n = H * W * C
    A = np.zeros((n, n))
    for i in range(n):
        e = np.zeros(n); e[i] = 1
        A[:, i] = O0.conv2d_same(e.reshape(1, H, W, C), K).reshape(-1)
    assert np.abs(A + A.T - 2 * gamma * np.eye(n)).max() < 1e-12


@pytest.mark.parametrize("strides", [(1, 1), (2, 2), (2, 1)])
def test_conv_o0_vs_o1_strided(strides):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 9, 8, 4)).astype(np.float32)
    K = rng.standard_normal((3, 3, 4, 4)).astype(np.float32)
    y0 = O0.conv2d_same(x.astype(np.float64), K.astype(np.float64), strides)
    y1 = O1.conv2d_same_nhwc(torch.from_numpy(x), torch.from_numpy(K), strides).numpy()
    assert y0.shape == y1.shape
    assert np.abs(y0 - y1).max() < 1e-4


@pytest.mark.parametrize("use_bn", [False, True])
def test_closed_form_backward_matches_autograd(use_bn):
    """SURVEY App. A.3/A.4 closed forms vs torch autograd through the literal assembly."""
    C, gamma, h = 6, -0.3, 0.125
    rng = np.random.default_rng(7)
    flat = rng.standard_normal(O0.num_params_3by3(C)) * 0.3
    x = rng.standard_normal((2, 5, 4, C))
    dY = rng.standard_normal((2, 5, 4, C))
    bn_g, bn_b = rng.standard_normal(C) + 1.5, rng.standard_normal(C) * 0.1
    K = O0.assemble_kernel_3by3_closed(flat, C, gamma)
    bias = flat[-C:]
    y, cache = O0.euler_step_fwd(x, K, bias, h, (bn_g, bn_b) if use_bn else None)
    dX, G, dbias, bn_grads, dZ = O0.euler_step_bwd(dY, cache, K, h, bn_g if use_bn else None)
    gflat = O0.fold_grad_3by3(G, C, dbias)

    tf_ = torch.from_numpy(flat).requires_grad_(True)
    tx = torch.from_numpy(x).requires_grad_(True)
    tg, tb = torch.from_numpy(bn_g).requires_grad_(True), torch.from_numpy(bn_b).requires_grad_(True)
    Kt = O1.assemble_literal(O1.split_params(tf_, C), C, gamma)
    ty = O1.euler_step(tx, Kt, tf_[-C:], h, (tg, tb) if use_bn else None)
    assert np.abs(ty.detach().numpy() - y).max() < 1e-12
    ty.backward(torch.from_numpy(dY))
    assert np.abs(tx.grad.numpy() - dX).max() < 1e-11
    assert np.abs(tf_.grad.numpy() - gflat).max() < 1e-11
    if use_bn:
        assert np.abs(tg.grad.numpy() - bn_grads[0]).max() < 1e-11
        assert np.abs(tb.grad.numpy() - bn_grads[1]).max() < 1e-11
    else:
        # App. A.4: dX = dY - conv_K(dZ) + 2 gamma dZ (dgrad reuses the forward weights)
        alt = dY - O0.conv2d_same(dZ, K) + 2 * gamma * dZ
        assert np.abs(alt - dX).max() < 1e-12


def test_net_plan_and_train_step_runs():
    spec = O1.NetSpec(blocks_per_stage=(2, 2, 2), h=0.5, gamma=-0.1)
    kinds = [k for k, *_ in spec.plan()]
    assert kinds == ["stem", "euler", "euler", "transition", "euler", "transition", "euler"]
    P = O1.init_net_params(spec, seed=1)
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (4, 32, 32, 3), generator=g, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (4,), generator=g), 10).float()
    l1, _ = O1.train_step(spec, P, M, V, 1, img, lab)
    l2, _ = O1.train_step(spec, P, M, V, 2, img, lab, assembly="literal")
    assert np.isfinite(l1) and np.isfinite(l2) and l2 < l1 + 0.5


def test_adam_tf1_matches_torch_formula():
    th, g = np.array([1.0, -2.0]), np.array([0.5, 0.25])
    m, v = np.zeros(2), np.zeros(2)
    th1, m1, v1 = O0.adam_step_tf1(th, g, m, v, 1)
    # first step of Adam moves each coordinate by ~lr*sign(g)
    assert np.allclose(th - th1, 1e-3 * np.sign(g), rtol=1e-5)


def test_diagonal_blocks_golden(golden_dir):
    """Printed K[:,:,1,1] of trained antisymmetric layers (antisymmetric_conv_kernel.ipynb cell 15,
    experiments_antisymmetric_resnet_v2.0.ipynb cell 16): both layer classes must rebuild every printed block exactly
    from its four free scalars -- pins the position and sign of every dependent entry of the diagonal law."""
    g = _load(golden_dir, "diagonal_blocks.json")
    assert len(g["cells"]) >= 2
    rng = np.random.default_rng(0)
    for cell in g["cells"]:
        for key in ("res2a_K_1_1", "res2d_K_1_1"):
            B = np.array(cell[key], np.float64).reshape(3, 3)
            C = 3
            # 3By3 class: a,b,c at (0,0),(0,1),(0,2), d at (1,0)   (tfkeras_layer_Conv2DAntisymmetric3By3.py:262-275)
            flat = rng.standard_normal(O0.num_params_3by3(C)).astype(np.float64)
            v = O0.split_params_3by3(flat, C)
            v[0][..., 1], v[1][..., 1], v[2][..., 1], v[3][..., 1] = B[0, 0], B[0, 1], B[0, 2], B[1, 0]
            flat = O0.join_params(v)
            for K in (O0.assemble_kernel_3by3_literal(O0.split_params_3by3(flat, C), C, 0.0),
                      O0.assemble_kernel_3by3_closed(flat, C, 0.0)):
                assert np.array_equal(K[:, :, 1, 1], B), (cell["source"], key)
            # general class at k = 3: free scalars at (0,0),(0,1),(0,2),(1,2); (1,0) = -v   (tfkeras_layer_Conv2DAntisymmetric.py:231-249)
            flat = rng.standard_normal(O0.num_params_general(C, 3)).astype(np.float64)
            vg = O0.split_params_general(flat, C, 3)
            idx = [i for i, a in enumerate(vg) if a.size == 1]      # diagonal scalars, 4 per output channel in creation order
            o1 = idx[4:8]
            for slot, val in zip(o1, (B[0, 0], B[0, 1], B[0, 2], B[1, 2])):
                vg[slot][...] = val
            Kg = O0.assemble_kernel_general_literal(vg, C, 3, 0.0)
            assert np.array_equal(Kg[:, :, 1, 1], B), (cell["source"], key)
