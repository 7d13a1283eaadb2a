"""GPU tests of the reference-shaped model builders: BatchNorm Euler step and whole-model forward
against the oracle."""
import numpy as np
import pytest
import torch

from oracle import antisym_numpy as O0
from oracle import antisym_torch as O1

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("C,shape", [(16, (4, 8, 8)), (64, (3, 6, 5)), (5, (2, 7, 7))])
def test_batchnorm_euler_step_forward_backward(C, shape):
    from differential_equations_resnet_b200.models import tfkeras_resnets as M
    N, H, W = shape
    gamma, h = -0.2, 0.25
    M._Scope.current = M._Scope(precision="strict", seed=1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn((N, H, W, C), generator=g)
    dy = torch.randn((N, H, W, C), generator=g)
    xr = x.cuda().requires_grad_(True)
    y = M.single_layer_identity_block(xr, 3, True, True, stage=2, block=0, h=h, gamma=gamma)
    sc = M._Scope.current
    layer, bn = sc.layers["res2_0_branch2"], sc.layers["bn2_0_branch2"]
    with torch.no_grad():
        bn.gamma.copy_(torch.rand(C, generator=g) + 0.5); bn.beta.copy_(torch.randn(C, generator=g) * 0.1)
        layer.packed[-C:] = torch.randn(C, generator=g).cuda() * 0.1
    y = M.single_layer_identity_block(xr, 3, True, True, stage=2, block=0, h=h, gamma=gamma)
    y.backward(dy.cuda())
    flat = layer.packed.detach().cpu().numpy().astype(np.float64)
    K = O0.assemble_kernel_3by3_closed(flat, C, gamma)
    bg, bb = bn.gamma.detach().cpu().numpy().astype(np.float64), bn.beta.detach().cpu().numpy().astype(np.float64)
    y_ref, cache = O0.euler_step_fwd(x.numpy().astype(np.float64), K, flat[-C:], h, (bg, bb))
    dX, G, dbias, bn_grads, _ = O0.euler_step_bwd(dy.numpy().astype(np.float64), cache, K, h, bg)
    assert rel(y.detach().cpu().numpy(), y_ref) <= 1e-5
    assert rel(xr.grad.cpu().numpy(), dX) <= 2e-5
    gref = O0.fold_grad_3by3(G, C, dbias)
    assert rel(layer.packed.grad.cpu().numpy()[:-C], gref[:-C]) <= 2e-5
    assert np.abs(layer.packed.grad.cpu().numpy()[-C:]).max() <= 1e-4      # bias grad vanishes under BN
    assert rel(bn.gamma.grad.cpu().numpy(), bn_grads[0]) <= 2e-5 and rel(bn.beta.grad.cpu().numpy(), bn_grads[1]) <= 2e-5
    M._Scope.current = None


def test_model_forward_matches_oracle_and_predict_runs():
    from differential_equations_resnet_b200.models import get_single_block_resnet_build_function
    kw = dict(h=0.5, gamma=-0.1, num_stages=4, blocks_per_stage=[2, 2, 2], filters_per_block=[16, 32, 64],
              strides=[(1, 1), (2, 2), (2, 2)], num_classes=10, subtract_mean=127.5, divide_by_stddev=127.5)
    img = torch.randint(0, 256, (4, 32, 32, 3), generator=torch.Generator().manual_seed(0), dtype=torch.uint8)
    model = get_single_block_resnet_build_function(kernel_type='antisymmetric', precision='strict', seed=7, **kw)(img.cuda())
    assert model.name == 'single_block_resnet_antisymmetric'
    names = [l.name for l in model.layers if hasattr(l, "name")]
    assert names[:3] == ['conv1', 'res2_0_branch2', 'res2_1_branch2'] and 'res3_0_branch2' in names and 'res3_0_branch1' in names
    # copy the parameters into the oracle's layout and compare probabilities
    ospec = O1.NetSpec(blocks_per_stage=(2, 2, 2), h=0.5, gamma=-0.1)
    P = {}
    for l in model.layers:
        n = getattr(l, "name", None)
        if isinstance(l, type(model.get_layer('conv1'))):
            P[n + "/kernel"], P[n + "/bias"] = l.kernel.detach().cpu(), l.bias.detach().cpu()
        elif n is not None and n.startswith("res"):
            P[n + "/packed"] = l.packed.detach().cpu()
    fc = model.get_layer('fc')
    P["fc/kernel"], P["fc/bias"] = fc.kernel.detach().cpu(), fc.bias.detach().cpu()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        probs = model(img.cuda()).detach().cpu()
    ref = O1.net_forward(ospec, P, img)
    assert float((probs - ref).abs().max()) <= 2e-5
    out = model.predict(img.numpy(), batch_size=2)
    assert out.shape == (4, 10) and np.allclose(out.sum(axis=1), 1.0, atol=1e-5)


def _bottleneck_params(model, dtype=torch.float64):
    """Model layers -> the oracle's parameter dict (float64 CPU leaves)."""
    from differential_equations_resnet_b200.models import tfkeras_resnets as M
    P = {}
    for name in model.scope.order:
        l = model.scope.layers[name]
        if isinstance(l, (M._RegularConv, M._Dense)):
            P[name + "/kernel"], P[name + "/bias"] = l.kernel.detach().cpu().to(dtype), l.bias.detach().cpu().to(dtype)
        elif isinstance(l, M._BatchNorm):
            P[name + "/gamma"], P[name + "/beta"] = l.gamma.detach().cpu().to(dtype), l.beta.detach().cpu().to(dtype)
        else:
            P[name + "/packed"] = l.packed.detach().cpu().to(dtype)
    return P


@pytest.mark.parametrize("version,use_bn", [(1, True), (1.5, False)])
def test_bottleneck_resnet_matches_oracle(version, use_bn):
    """`get_resnet_build_function` (reference models/tfkeras_resnets.py:698-818) with antisymmetric middle convolutions
    (`None` as the middle filter count; :163-169, :370-376): forward against the float64 restatement; version 1.5 puts
    stride 2 on the antisymmetric layer.  Without BN also the gradient of every antisymmetric layer's packed parameters."""
    from differential_equations_resnet_b200.models import get_resnet_build_function
    from differential_equations_resnet_b200.layers import Conv2DAntisymmetric3By3
    fpb = [[16, None, 32], [32, None, 64], [32, None, 64], [64, None, 128]]
    bps = [2, 1, 2, 1]
    kw = dict(kernel_type='antisymmetric', num_classes=10, version=version, blocks_per_stage=bps, filters_per_block=fpb,
              use_batch_norm=use_bn, subtract_mean=127.5, divide_by_stddev=127.5)
    img = torch.randint(0, 256, (4, 64, 64, 3), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)
    model = get_resnet_build_function(precision='strict', seed=11, **kw)(img.cuda())
    assert model.name == 'resnet_antisymmetric'
    names = model.scope.order
    assert names[0] == 'conv1' and 'res2_0_branch2a' in names and 'res2_0_branch1' in names and 'res2_1_branch2b' in names
    assert isinstance(model.get_layer('res3_0_branch2b'), Conv2DAntisymmetric3By3)
    assert model.get_layer('res3_0_branch2b').strides == ((2, 2) if version == 1.5 else (1, 1))
    assert ('bn2_0_branch2c' in names) == use_bn
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        torch.backends.cuda.matmul.allow_tf32 = False
        probs = model(img.cuda(), training=True)
        w = torch.randn(probs.shape, generator=torch.Generator().manual_seed(6), dtype=torch.float64)
        if not use_bn:
            (probs.double() * w.cuda()).sum().backward()
    P = _bottleneck_params(model)
    for v in P.values():
        v.requires_grad_(True)
    ref = O1.bottleneck_resnet_forward(P, img.double(), bps, fpb, True, use_bn, version, 0.0, 127.5, 127.5)
    assert probs.shape == (4, 10)
    assert float((probs.detach().cpu().double() - ref.detach()).abs().max()) <= 2e-5
    if not use_bn:
        (ref * w).sum().backward()
        for name in names:
            l = model.scope.layers[name]
            if isinstance(l, Conv2DAntisymmetric3By3):
                assert rel(l.packed.grad.cpu().numpy(), P[name + "/packed"].grad.numpy()) <= 1e-4, name


def test_bottleneck_regular_middle_layer_and_errors():
    from differential_equations_resnet_b200.models import bottleneck_conv_block, get_resnet_build_function
    from differential_equations_resnet_b200.models import tfkeras_resnets as M
    with pytest.raises(ValueError, match="num_classes"):
        get_resnet_build_function()
    with pytest.raises(ValueError, match="preset"):
        get_resnet_build_function(num_classes=10, preset='resnet18')
    M._Scope.current = M._Scope("strict", 1)
    try:
        x = torch.rand(2, 8, 8, 16, device="cuda")
        y = bottleneck_conv_block(x, 3, (8, 8, 32), True, False, stage=2, block=0, version=1.5, strides=(2, 2))   # middle count given: regular conv
        assert y.shape == (2, 4, 4, 32) and isinstance(M._Scope.current.layers['res2_0_branch2b'], M._RegularConv)
        with pytest.raises(ValueError, match="version"):
            bottleneck_conv_block(x, 3, (8, None, 32), True, False, stage=3, block=0, version=2)
    finally:
        M._Scope.current = None


def test_antisymmetric_weights_into_regular_model_and_double_load(tmp_path):
    """experiments_antisymmetric_resnet_v7.ipynb 'Antisymmetric 16 Weights Loaded into Regular 16 Model': the dense kernels
    the pack kernel assembles, pickled in the reference's format (`model_utils/weight_utils.py:23-39`), make a regular net
    of the same shape compute the same function; `double_load_weights` (:41-80) into a net twice as deep, antisymmetric
    layers included (free parameters read back out of the dense kernels, exact)."""
    import pickle
    from differential_equations_resnet_b200.models import get_single_block_resnet_build_function
    from differential_equations_resnet_b200.model_utils import double_load_weights, load_pickled_weights, pickle_model_weights

    def build(kernel_type, blocks, seed):
        return get_single_block_resnet_build_function(
            kernel_type=kernel_type, h=0.5, gamma=-0.1, num_stages=2, blocks_per_stage=[blocks], filters_per_block=[16],
            strides=[(1, 1)], num_classes=10, subtract_mean=127.5, divide_by_stddev=127.5, precision='strict', seed=seed)

    img = torch.randint(0, 256, (4, 16, 16, 3), generator=torch.Generator().manual_seed(3), dtype=torch.uint8).cuda()
    anti, regular, deep = build('antisymmetric', 3, 1)(img), build('regular', 3, 2)(img), build('antisymmetric', 6, 3)(img)
    path = str(tmp_path / "anti.pkl")
    pickle_model_weights(anti, path)
    saved = pickle.load(open(path, 'rb'))
    assert len(saved) == 5 and saved[1]['kernel'].shape == (3, 3, 16, 16) and saved[1]['bias'].shape == (16,)
    K = saved[2]['kernel']
    S = K + K[::-1, ::-1].transpose(0, 1, 3, 2)
    S[1, 1, np.arange(16), np.arange(16)] -= np.float32(-0.2)
    assert not S.any()                                                   # bit-exact antisymmetry survives the file
    load_pickled_weights(regular, path)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        pa, pr = anti(img, training=False), regular(img, training=False)
    assert float((pa - pr).detach().abs().max()) <= 2e-5
    double_load_weights(deep, path)
    src = [l for l in anti.layers if l.name.startswith('res')]
    dst = [l for l in deep.layers if l.name.startswith('res')]
    assert len(src) == 3 and len(dst) == 6
    for i, l in enumerate(dst):
        assert torch.equal(l.packed.detach(), src[i // 2].packed.detach())
    rnd = build('regular', 3, 7)(img)
    with pytest.raises(ValueError, match="anti-centrosymmetric"):
        rnd_path = str(tmp_path / "reg.pkl")
        pickle_model_weights(rnd, rnd_path)
        load_pickled_weights(anti, rnd_path)                             # a generic kernel is not antisymmetric: refused
