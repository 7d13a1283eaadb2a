"""CPU tests: the C-ABI library loads, exports every symbol include/b200ode.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200ode.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ode_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from differential_equations_resnet_b200 import _abi
    lib = _abi.lib()
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_abi.EXPORTED_SYMBOLS) == syms
    assert lib.b200ode_version() == 200


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine without a GPU")
def test_no_cpu_fallback():
    import differential_equations_resnet_b200 as pkg
    from differential_equations_resnet_b200 import _abi
    assert _abi.lib().b200ode_device_ok() == 0
    layer = pkg.Conv2DAntisymmetric3By3(gamma=0.0)
    with pytest.raises(_abi.B200OdeError, match="no CPU fallback"):
        layer(torch.zeros(1, 4, 4, 16))


def test_layer_constructor_signatures_match_reference():
    """Same keyword names and defaults as the reference constructors
    (layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:60-66, tfkeras_layer_Conv2DAntisymmetric.py:60-68)."""
    import inspect
    import differential_equations_resnet_b200 as pkg
    s3 = inspect.signature(pkg.Conv2DAntisymmetric3By3.__init__)
    assert list(s3.parameters)[:6] == ["self", "gamma", "strides", "use_bias", "kernel_initializer", "kernel_regularizer"]
    assert s3.parameters["gamma"].default == 0.0 and s3.parameters["strides"].default == (1, 1)
    assert s3.parameters["use_bias"].default is True and s3.parameters["kernel_initializer"].default == "he_normal"
    sg = inspect.signature(pkg.Conv2DAntisymmetric.__init__)
    assert list(sg.parameters)[:8] == ["self", "kernel_size", "gamma", "strides", "use_bias", "kernel_initializer",
                                       "kernel_regularizer", "antisymmetric"]
    assert sg.parameters["antisymmetric"].default is True
    layer = pkg.Conv2DAntisymmetric3By3(gamma=-0.1, name="res2_0_branch2")
    assert layer.name == "res2_0_branch2" and layer.compute_output_shape((1, 2, 3, 4)) == (1, 2, 3, 4)
    cfg = layer.get_config()
    assert cfg["strides"] == (1, 1) and cfg["use_bias"] is True and cfg["gamma"] == -0.1


def test_variable_shapes_follow_reference_order():
    import differential_equations_resnet_b200 as pkg
    from oracle import antisym_numpy as O0
    layer = pkg.Conv2DAntisymmetric3By3()
    layer.num_channels, layer.use_bias = 16, True
    shapes = layer._variable_shapes()
    assert len(shapes) == 20 and shapes[:4] == [(1, 1, 1, 16)] * 4 and shapes[4] == (3, 3, 15) and shapes[-1] == (16,)
    gl = pkg.Conv2DAntisymmetric(5, antisymmetric=True)
    gl.num_channels, gl.use_bias = 3, True
    import numpy as np
    assert sum(int(np.prod(s)) for s in gl._variable_shapes()) == O0.num_params_general(3, 5, True)


def test_netspec_plan_matches_oracle_plan():
    from differential_equations_resnet_b200.training import NetSpec
    from oracle import antisym_torch as O1
    kw = dict(blocks_per_stage=(36, 37, 37), filters_per_block=(16, 32, 64), h=8 / 108)
    assert NetSpec(**kw).plan() == O1.NetSpec(**kw).plan()
    kinds = [p[0] for p in NetSpec(**kw).plan()]
    assert kinds.count("euler") == 108 and kinds.count("transition") == 2
    # MaxPooling2D in front of a stage (models/tfkeras_resnets.py:577-578): that stage starts with a conv block (:589-593)
    kp = dict(blocks_per_stage=(2, 3, 2), filters_per_block=(16, 16, 32), strides=((1, 1), (1, 1), (2, 2)), use_max_pooling=[False, True, False])
    plan = NetSpec(**kp).plan()
    assert plan == O1.NetSpec(**kp).plan()
    assert [p[0] for p in plan] == ["stem", "euler", "euler", "maxpool", "transition", "euler", "euler", "transition", "euler"]
    assert plan[3][4] == "stage3_pooling" and plan[4][1:4] == (16, 16, (1, 1))


def test_conv_tile_planner_decisions():
    """The tile planner runs on the host: pin the decisions DESIGN.md's measurements rest on (no GPU needed).
    C >= 128 (and bf16 C = 64) at cfg2's 256x32x32 takes 128-position tiles with DOUBLE-buffered accumulators (the drain of a tile
    overlaps the next tile's MMAs: bf16 C=256 forward 316 -> 279 us); every plan covers all positions."""
    import ctypes
    from differential_equations_resnet_b200 import _abi
    lib = _abi.lib()
    out = (ctypes.c_int * 8)()
    for mode, name in ((1, "fast_tf32"), (2, "fast_bf16")):
        for C in (64, 128, 256):
            assert lib.b200ode_debug_conv_plan(mode, C, 256, 32, 32, out) == 0
            nimg, spi, tpi, tiles, grid, sa, sw, acc = list(out)
            assert nimg == 1 and acc == 2, (name, C, list(out))
            assert spi == 1 or (name == "fast_tf32" and C == 64 and spi <= 3), (name, C, list(out))
            assert tiles == 256 * tpi and tpi * spi * 128 >= 32 * 33 and 1 <= grid <= 148 and sa >= 2 and sw >= 2
    # small images: whole images per tile
    assert lib.b200ode_debug_conv_plan(1, 64, 128, 8, 8, out) == 0
    assert out[2] == 1 and out[0] >= 1
    # strict mode at C = 256 fits (two weight stages) instead of being refused
    assert lib.b200ode_debug_conv_plan(0, 256, 256, 32, 32, out) == 0 and out[6] == 2
    # unsupported requests fail with a message instead of planning nonsense
    assert lib.b200ode_debug_conv_plan(1, 48, 8, 8, 8, out) < 0
    assert b"tensor-core plans" in lib.b200ode_last_error()


@pytest.mark.parametrize("size", [3, 5, 7])
@pytest.mark.parametrize("anti", [True, False])
def test_get_centrosymmetric_matrix_package_function(size, anti):
    """The package's free function (reference layers/antisymmetric_conv2d_utils.py:23-75): free scalars at (i,j) for
    j > i or (j == i, i <= size//2 - 1) in creation order, the mirror (size-1-i, size-1-j) gets -v (anti) / +v, the
    centre of an anti-centrosymmetric block is a constant zero; rank reshapes with trailing singleton axes.
    Checked against the oracle's literal builder fed with the scalars the function drew."""
    import math
    import numpy as np
    from oracle import antisym_numpy as O0
    from differential_equations_resnet_b200.layers.antisymmetric_conv2d_utils import get_centrosymmetric_matrix
    C = 4
    g = torch.Generator().manual_seed(size * 2 + int(anti))
    m4 = get_centrosymmetric_matrix(size, C, rank=4, anti=anti, generator=g)
    assert tuple(m4.shape) == (size, size, 1, 1)
    m = m4.reshape(size, size).numpy()
    slots = O0.diag_slots_general(size, anti)
    want = O0._centrosymmetric_matrix([m[i, j] for i, j in slots], size, 0.0, anti, np.float32)
    assert np.array_equal(m, want)
    sign = -1.0 if anti else 1.0
    assert np.array_equal(m, sign * m[::-1, ::-1])
    if anti:
        assert m[size // 2, size // 2] == 0.0
    assert len(slots) == (size * size - 1) // 2 + (0 if anti else 1)
    assert float(np.abs(m).max()) <= 2.0 * math.sqrt(2.0 / (size * size * C)) + 1e-7     # truncated normal, +-2 sigma
    assert float(np.abs(m).max()) > 0.0
    m2 = get_centrosymmetric_matrix(size, C, rank=2, anti=anti, generator=torch.Generator().manual_seed(size * 2 + int(anti)))
    assert tuple(m2.shape) == (size, size) and np.array_equal(m2.numpy(), m)


def test_glue_workspace_query_is_host_only():
    """b200ode_glue_workspace_bytes plans on the host (no device needed): sizes of the caller-owned scratch of the stem /
    transition / head gradient calls (include/b200ode.h, SURVEY.md 8b workspace contract)."""
    import ctypes
    from differential_equations_resnet_b200 import _abi
    lib = _abi.lib()
    n = ctypes.c_size_t()
    assert lib.b200ode_glue_workspace_bytes(_abi.GLUE_STEM_WGRAD, 128, 32, 32, 3, 16, 1, 1, ctypes.byref(n)) == 0
    assert n.value == 128 * 4 * (27 * 16 + 16) * 4                      # images x bands of 8 rows x (kernel + bias) floats
    assert lib.b200ode_glue_workspace_bytes(_abi.GLUE_TRANSITION_WGRAD, 128, 32, 32, 16, 32, 2, 2, ctypes.byref(n)) == 0
    assert n.value % (10 * 16 * 32 + 64) == 0 and n.value >= 128 * (10 * 16 * 32 + 64) * 4
    assert lib.b200ode_glue_workspace_bytes(_abi.GLUE_HEAD, 128, 1, 1, 64, 10, 1, 1, ctypes.byref(n)) == 0
    assert n.value == 128 * (64 * 10 + 10 + 1) * 4
    assert lib.b200ode_glue_workspace_bytes(_abi.GLUE_HEAD, 0, 1, 1, 64, 10, 1, 1, ctypes.byref(n)) == 0 and n.value == 0
    assert lib.b200ode_glue_workspace_bytes(9, 1, 1, 1, 1, 1, 1, 1, ctypes.byref(n)) == -1
    assert "unknown glue op" in _abi.last_error()
    assert lib.b200ode_glue_workspace_bytes(_abi.GLUE_TRANSITION_WGRAD, 8, 32, 32, 12, 20, 2, 2, ctypes.byref(n)) == -2   # unsupported pair: no fallback


def test_bottleneck_builder_signatures_match_reference():
    """Keyword names, order and defaults of the bottleneck builders (reference models/tfkeras_resnets.py:96-105, :271-282,
    :698-713); `precision` / `seed` are additions at the end."""
    import inspect
    from differential_equations_resnet_b200 import models as M
    p = inspect.signature(M.bottleneck_identity_block).parameters
    assert list(p) == ['input_tensor', 'kernel_size', 'num_filters', 'antisymmetric', 'use_batch_norm', 'stage', 'block',
                       'gamma', 'kernel_regularizer', 'bias_regularizer'] and p['gamma'].default == 0.0
    p = inspect.signature(M.bottleneck_conv_block).parameters
    assert list(p) == ['input_tensor', 'kernel_size', 'num_filters', 'antisymmetric', 'use_batch_norm', 'stage', 'block',
                       'version', 'strides', 'gamma', 'kernel_regularizer', 'bias_regularizer']
    assert p['version'].default == 1 and p['strides'].default == (1, 1)
    p = inspect.signature(M.get_resnet_build_function).parameters
    assert list(p)[:12] == ['kernel_type', 'include_top', 'fc_activation', 'num_classes', 'l2_regularization', 'subtract_mean',
                            'divide_by_stddev', 'version', 'preset', 'blocks_per_stage', 'filters_per_block', 'use_batch_norm']
    assert p['blocks_per_stage'].default == [3, 4, 6, 3] and p['filters_per_block'].default[3] == [512, 512, 2048]
    assert p['use_batch_norm'].default is True and p['kernel_type'].default == 'antisymmetric'
    with pytest.raises(ValueError, match="num_classes"):
        M.get_resnet_build_function()
    with pytest.raises(ValueError, match="preset"):
        M.get_resnet_build_function(num_classes=10, preset='resnet18')
    assert callable(M.get_resnet_build_function(num_classes=10, preset='resnet152')) and callable(M.build_resnet)
