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
