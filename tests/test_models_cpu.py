"""CPU tests of the reference-shaped model builders with REGULAR kernels (torch ops only, no device code): the graph
wiring of the bottleneck ResNet (reference models/tfkeras_resnets.py:96-202, :271-425, :698-818) against the oracle's
restatement, and the Keras `.h5` round trip through `Model.save_weights / load_weights`.  The antisymmetric variants
need the CUDA library and live in tests/test_gpu_models.py."""
import pytest
import torch

from oracle import antisym_torch as O1

FPB = [[16, 16, 32], [32, 32, 64], [32, 32, 64], [64, 64, 128]]
BPS = [2, 1, 2, 1]


def _params(model, dtype=torch.float64):
    from differential_equations_resnet_b200.models import tfkeras_resnets as M
    P = {}
    for name in model.scope.order:
        l = model.scope.layers[name]
        if isinstance(l, M._BatchNorm):
            P[name + "/gamma"], P[name + "/beta"] = l.gamma.detach().to(dtype), l.beta.detach().to(dtype)
        else:
            P[name + "/kernel"], P[name + "/bias"] = l.kernel.detach().to(dtype), l.bias.detach().to(dtype)
    return P


@pytest.mark.parametrize("version,use_bn", [(1, True), (1.5, False), (1.5, True)])
def test_regular_bottleneck_resnet_matches_oracle_and_h5_roundtrip(version, use_bn, tmp_path):
    from differential_equations_resnet_b200.models import get_resnet_build_function
    kw = dict(kernel_type='regular', num_classes=10, version=version, blocks_per_stage=BPS, filters_per_block=FPB,
              use_batch_norm=use_bn, subtract_mean=127.5, divide_by_stddev=127.5)
    img = torch.randint(0, 256, (4, 64, 64, 3), generator=torch.Generator().manual_seed(2), dtype=torch.uint8)
    m = get_resnet_build_function(seed=3, **kw)(None)
    probs = m(img, training=True)
    assert m.name == 'resnet_regular' and probs.shape == (4, 10)
    order = m.scope.order
    assert order[0] == 'conv1' and order[-1] == 'fc' and ('bn_conv1' in order) == use_bn
    assert [n for n in order if n.startswith('res2_0')] == ['res2_0_branch2a', 'res2_0_branch2b', 'res2_0_branch2c', 'res2_0_branch1']
    assert m.get_layer('res3_0_branch2a').strides == ((2, 2) if version == 1 else (1, 1))
    assert m.get_layer('res3_0_branch2b').strides == ((1, 1) if version == 1 else (2, 2))
    assert m.get_layer('res3_0_branch1').strides == (2, 2) and m.get_layer('res2_0_branch1').strides == (1, 1)
    ref = O1.bottleneck_resnet_forward(_params(m), img.double(), BPS, FPB, False, use_bn, version, 0.0, 127.5, 127.5)
    assert float((probs.detach().double() - ref).abs().max()) <= 2e-5

    m2 = get_resnet_build_function(seed=4, **kw)(None)
    m2(img, training=True)
    assert not torch.equal(m(img, training=False), m2(img, training=False))
    path = str(tmp_path / "resnet.h5")
    m.save_weights(path)
    m2.load_weights(path)
    assert torch.equal(m(img, training=False), m2(img, training=False))       # moving statistics travel too
    other = get_resnet_build_function(seed=4, **dict(kw, filters_per_block=[[8, 8, 32]] + FPB[1:]))(None)
    other(img, training=True)
    with pytest.raises(ValueError, match="shape"):
        other.load_weights(path)


def test_unbuilt_model_refuses_to_save(tmp_path):
    from differential_equations_resnet_b200.models import get_resnet_build_function
    m = get_resnet_build_function(kernel_type='regular', num_classes=3, blocks_per_stage=[1, 1, 1, 1], filters_per_block=FPB)(None)
    m.save_weights(str(tmp_path / "empty.h5"))            # no layers yet: an empty (valid) file
    from differential_equations_resnet_b200.keras_h5 import load_keras_weights
    assert load_keras_weights(str(tmp_path / "empty.h5")) == {}


def _single_block(kernel_type, blocks, seed, **extra):
    from differential_equations_resnet_b200.models import get_single_block_resnet_build_function
    kw = dict(kernel_type=kernel_type, h=0.5, gamma=-0.1, num_stages=2, blocks_per_stage=[blocks], filters_per_block=[16],
              strides=[(1, 1)], num_classes=10, subtract_mean=127.5, divide_by_stddev=127.5, seed=seed)
    kw.update(extra)
    return get_single_block_resnet_build_function(**kw)(None)


def test_pickle_and_double_load_weights_regular(tmp_path):
    """`model_utils/weight_utils.py:23-80` on regular single-block nets: an (l+2)-layer net's blocks land twice in the
    (2l+2)-layer net, stem and dense layer once."""
    import pickle
    import numpy as np
    from differential_equations_resnet_b200.model_utils import double_load_weights, load_pickled_weights, pickle_model_weights
    img = torch.randint(0, 256, (2, 8, 8, 3), generator=torch.Generator().manual_seed(0), dtype=torch.uint8)
    small, big = _single_block('regular', 3, 1), _single_block('regular', 6, 2)
    small(img), big(img)
    path = str(tmp_path / "w.pkl")
    pickle_model_weights(small, path)
    saved = pickle.load(open(path, 'rb'))
    assert len(saved) == 5 and set(saved[0]) == {'kernel', 'bias'} and saved[0]['kernel'].shape == (3, 3, 3, 16)
    double_load_weights(big, path)
    lay = [l for l in big.layers if len(l.get_weights()) > 0]
    assert len(lay) == 8
    assert np.array_equal(lay[0].get_weights()[0], saved[0]['kernel']) and np.array_equal(lay[-1].get_weights()[0], saved[-1]['kernel'])
    for l in range(1, 4):
        for tgt in (lay[2 * (l - 1) + 1], lay[2 * l]):
            assert np.array_equal(tgt.get_weights()[0], saved[l]['kernel']) and np.array_equal(tgt.get_weights()[1], saved[l]['bias'])
    with pytest.raises(ValueError, match="weighted layers"):
        double_load_weights(small, path)
    twin = _single_block('regular', 3, 9)
    twin(img)
    load_pickled_weights(twin, path)
    assert torch.equal(twin(img, training=False), small(img, training=False))
    bn = _single_block('regular', 1, 1, use_batch_norm=True)
    bn(img)
    with pytest.raises(ValueError, match="weight arrays"):
        pickle_model_weights(bn, path)               # like the reference: (kernel, bias) layers only


@pytest.mark.parametrize("C,gamma", [(5, 0.0), (16, -0.1)])
def test_unpack_dense_is_the_inverse_of_the_assembly(C, gamma):
    import numpy as np
    from oracle import antisym_numpy as O0
    from differential_equations_resnet_b200.model_utils import unpack_dense_3by3
    from differential_equations_resnet_b200.checkpoint import variable_shapes_3by3
    n = 4 * C + 9 * C * (C - 1) // 2 + C
    flat = np.random.default_rng(C).standard_normal(n).astype(np.float32)
    K = O0.assemble_kernel_3by3_closed(flat.astype(np.float64), C, gamma).astype(np.float32)
    vs = unpack_dense_3by3(K, gamma, flat[-C:])
    assert [tuple(v.shape) for v in vs] == [tuple(s) for _, s in variable_shapes_3by3(C)]
    assert np.array_equal(np.concatenate([v.reshape(-1) for v in vs]), flat)
    bad = K.copy()
    bad[0, 1, 2, 3] += 1e-3
    with pytest.raises(ValueError, match="anti-centrosymmetric"):
        unpack_dense_3by3(bad, gamma)
    with pytest.raises(ValueError, match="anti-centrosymmetric"):
        unpack_dense_3by3(K, gamma + 0.5)
