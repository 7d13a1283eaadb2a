"""Checkpoint interchange with the reference's variable layout."""
import json
import os

import numpy as np
import pytest
import torch

from differential_equations_resnet_b200 import checkpoint as ck

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_variable_order_matches_reference_golden():
    # training/training.py:397-398 assumes 20 variables per 16-channel layer; names from notebook v6 raw :1718
    names = [n for n, _ in ck.variable_shapes_3by3(16)]
    assert len(names) == 16 + 4
    assert names[:4] == ["a", "b", "c", "d"] and names[-1] == "bias"
    assert names[4] == "input_kernels_for_output_kernel_0" and names[-2] == "input_kernels_for_output_kernel_14"
    shapes = dict(ck.variable_shapes_3by3(16))
    assert shapes["a"] == (1, 1, 1, 16) and shapes["input_kernels_for_output_kernel_0"] == (3, 3, 15)
    assert shapes["input_kernels_for_output_kernel_14"] == (3, 3, 1) and shapes["bias"] == (16,)
    gpath = os.path.join(GOLDEN, "variable_order.json")
    if os.path.exists(gpath):
        g = json.load(open(gpath))
        ref = [n.split("/")[-1].split(":")[0] for n in g.get("names", [])][:20]
        if ref:
            assert ref == names


@pytest.mark.parametrize("C", [2, 5, 16, 64])
def test_split_join_roundtrip(C):
    n = 4 * C + 9 * C * (C - 1) // 2 + C
    flat = np.random.default_rng(C).standard_normal(n).astype(np.float32)
    v = ck.split_packed_3by3(flat, C)
    assert len(v) == C + 4
    # the oracle's literal assembly consumes exactly these variables: same kernel as from the flat vector
    from oracle import antisym_numpy as O0
    K1 = O0.assemble_kernel_3by3_closed(flat.astype(np.float64), C, -0.1)
    back = ck.join_packed_3by3({k + ":0": a for k, a in v.items()}, C)
    assert np.array_equal(back, flat)
    K2 = O0.assemble_kernel_3by3_closed(back.astype(np.float64), C, -0.1)
    assert np.array_equal(K1, K2)
    with pytest.raises(ValueError):
        ck.split_packed_3by3(flat[:-1], C)
    v["a"] = v["a"].reshape(C)
    with pytest.raises(ValueError):
        ck.join_packed_3by3(v, C)


@pytest.mark.gpu
def test_save_load_resume_bit_identical(tmp_path):
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(2, 3, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
    gen = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8).cuda()
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float().cuda()
    a = EulerNet(NetSpec(**kw), seed=3)
    a.train_step(img, lab); a.train_step(img, lab)
    path = str(tmp_path / "variables.npz")
    ck.save_variables(a, path)
    names = list(ck.export_reference_variables(a).keys())
    assert names[0] == "conv1/kernel" and "res2_0_branch2/a" in names and "res3_0_branch2/kernel" in names
    assert "res3_1_branch2/input_kernels_for_output_kernel_30" in names and names[-1] == "fc/bias"
    b = EulerNet(NetSpec(**kw), seed=99)
    ck.load_variables(b, path)
    assert torch.equal(a.theta, b.theta) and torch.equal(a.adam_m, b.adam_m) and int(b.step_counter) == int(a.step_counter)
    la, lb = float(a.train_step(img, lab)), float(b.train_step(img, lab))
    assert la == lb and torch.equal(a.theta, b.theta)


@pytest.mark.gpu
def test_dense_weights_and_double_load(tmp_path):
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    from oracle import antisym_numpy as O0
    small = EulerNet(NetSpec(blocks_per_stage=(2, 2, 2), filters_per_block=(16, 32, 64), h=0.5, gamma=-0.1), seed=7)
    dense = ck.dense_layer_weights(small)
    # conv1, 2 Euler, (transition main + shortcut, 1 Euler) x 2, fc
    assert len(dense) == 1 + 2 + 3 + 3 + 1
    flat = small.export_params()["res2_1_branch2/packed"].numpy()
    K = O0.assemble_kernel_3by3_closed(flat.astype(np.float64), 16, -0.1).astype(np.float32)
    assert dense[2]["kernel"].shape == (3, 3, 16, 16) and np.array_equal(dense[2]["kernel"], K)
    # antisymmetry of the exported dense kernel is bit-exact: K[a,b,ci,o] + K[2-a,2-b,o,ci] = 2 gamma [centre][ci=o]
    S = dense[2]["kernel"] + dense[2]["kernel"][::-1, ::-1].transpose(0, 1, 3, 2)
    S[1, 1][np.arange(16), np.arange(16)] -= np.float32(2 * -0.1)
    assert not S.any()
    ck.pickle_model_weights(small, str(tmp_path / "w.pkl"))
    import pickle
    assert len(pickle.load(open(str(tmp_path / "w.pkl"), "rb"))) == len(dense)
    big = EulerNet(NetSpec(blocks_per_stage=(4, 3, 3), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.1), seed=8)
    ck.double_load_variables(big, ck.export_reference_variables(small))
    ps, pb = small.export_params(), big.export_params()
    assert torch.equal(pb["res2_2_branch2/packed"], ps["res2_1_branch2/packed"]) and torch.equal(pb["res2_3_branch2/packed"], ps["res2_1_branch2/packed"])
    assert torch.equal(pb["res3_1_branch2/packed"], ps["res3_1_branch2/packed"]) and torch.equal(pb["res3_2_branch2/packed"], ps["res3_1_branch2/packed"])
    assert torch.equal(pb["fc/kernel"], ps["fc/kernel"]) and torch.equal(pb["res4_0_branch2/kernel"], ps["res4_0_branch2/kernel"])
    # halving h while doubling the blocks integrates the same ODE: predictions stay close (same final time)
    gen = torch.Generator().manual_seed(2)
    img = torch.randint(0, 256, (4, 32, 32, 3), generator=gen, dtype=torch.uint8).cuda()
    pa, pb_ = small.predict(img), big.predict(img)
    assert float((pa - pb_).abs().max()) < 0.2
