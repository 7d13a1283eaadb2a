"""GPU parity of the whole train step (forward, Keras CE loss, backward, TF1 Adam) against the O1
oracle with identical parameters and inputs."""
import numpy as np
import pytest
import torch

from oracle import antisym_torch as O1

pytestmark = pytest.mark.gpu


def _run(precision, tol_loss, tol_grad, graph=False):
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(3, 3, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
    ospec = O1.NetSpec(**kw)
    P = O1.init_net_params(ospec, seed=5)
    gen = torch.Generator().manual_seed(3)
    for k in P:                      # non-zero biases everywhere
        if k.endswith("/bias"):
            P[k] = torch.randn(P[k].shape, generator=gen) * 0.05
        if k.endswith("/packed"):
            C = int(k and [c for c in (16, 32, 64) if P[k].numel() == 4 * c + 9 * c * (c - 1) // 2 + c][0])
            P[k][-C:] = torch.randn(C, generator=gen) * 0.05
    net = EulerNet(NetSpec(**kw), precision=precision, seed=0)
    net.import_params(P)
    P0 = {k: v.clone() for k, v in P.items()}
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float()
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    losses_ref, losses = [], []
    if graph:
        net.capture(img.cuda(), lab.cuda(), warmup=1)
        net.import_params(P)
        net.adam_m.zero_(); net.adam_v.zero_(); net.step_counter.fill_(1)
    for t in (1, 2, 3):
        lr, gr = O1.train_step(ospec, P, M, V, t, img, lab)
        losses_ref.append(lr)
        l = net.train_step_graph() if graph else net.train_step(img.cuda(), lab.cuda())
        losses.append(float(l))
        if t == 1:
            g = net.export_grads()
            for k in gr:
                err = float((g[k].double() - gr[k].double()).norm() / max(float(gr[k].double().norm()), 1e-30))
                assert err <= tol_grad, (k, err)
    # Step 1's loss is pure forward parity.  Steps 2 and 3 sit behind Adam updates: at t <= 3 Adam moves every element by
    # ~lr * sign(g), so elements whose gradient is at rounding level move the OTHER way by 2e-3 (see the update check below)
    # and the later losses inherit that: measured 1.3e-5 on step 3 in strict mode with the fp32-grade native glue kernels.
    for t, (a, b) in enumerate(zip(losses, losses_ref)):
        assert abs(a - b) <= (tol_loss if t == 0 else max(tol_loss, 1e-4)) * max(1.0, abs(b)), (losses, losses_ref)
    th = net.export_params()
    # After 3 Adam steps compare the UPDATES relatively.  At step t <= 3 Adam moves every element by ~lr*sign(g)
    # (|g| >> eps), so the update's relative error is sqrt(4 * fraction of elements whose gradient sign differs):
    # a gradient relative error e flips about a fraction e of the signs -> strict (1e-5) ~ 6e-3, fast (1e-2) ~ 0.2;
    # an unrelated update scores sqrt(2).
    tol_upd = 2e-2 if precision == "strict" else 0.35
    num = den = 0.0
    for k in P:
        du, du_ref = (th[k] - P0[k]).double(), (P[k] - P0[k]).double()
        num += float((du - du_ref).pow(2).sum()); den += float(du_ref.pow(2).sum())
        if k.endswith("/packed") or k.endswith("/kernel"):
            e = float((du - du_ref).norm() / du_ref.norm())
            assert e <= 2 * tol_upd, (k, e)
    assert (num / den) ** 0.5 <= tol_upd, (num / den) ** 0.5


def test_train_step_strict_matches_oracle():
    _run("strict", 1e-5, 2e-4)


def test_train_step_fast_tf32_matches_oracle():
    _run("fast_tf32", 2e-3, 3e-2)


def test_train_step_fast_bf16_matches_oracle():
    """fast_bf16 (bf16 operands and activations, fp32 accumulate; per-layer kernels): loss within 2e-3 relative,
    gradients within 6e-2 (measured 2e-2 worst, 4e-3 median on this net)."""
    _run("fast_bf16", 2e-3, 6e-2)


def test_train_step_cuda_graph_matches_oracle():
    _run("strict", 1e-5, 2e-4, graph=True)


def test_train_step_fast_tf32_cuda_graph_matches_oracle():
    """fast_tf32 takes the persistent chain kernels (one launch per stage and direction)."""
    _run("fast_tf32", 2e-3, 3e-2, graph=True)


def test_train_step_fast_f16_matches_oracle():
    """fast_f16: persistent chains with fp16 operands (the tf32 significand, rounded to nearest) around an fp32 residual stream."""
    _run("fast_f16", 2e-3, 3e-2)


def test_train_step_fast_f16_cuda_graph_matches_oracle():
    _run("fast_f16", 2e-3, 3e-2, graph=True)


@pytest.mark.parametrize("precision", ["fast_tf32", "fast_f16"])
def test_native_predict_matches_oracle(precision):
    """Inference through stem + one chain launch per stage (activations never leave shared memory inside a
    stage) + transitions + head against the O1 forward pass; fast_tf32 tolerance on probabilities."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(4, 5, 3), filters_per_block=(16, 32, 64), h=0.125, gamma=-0.05)
    ospec = O1.NetSpec(**kw)
    P = O1.init_net_params(ospec, seed=9)
    net = EulerNet(NetSpec(**kw), precision=precision, seed=0)
    net.import_params(P)
    gen = torch.Generator().manual_seed(4)
    for N in (1, 9):
        img = torch.randint(0, 256, (N, 32, 32, 3), generator=gen, dtype=torch.uint8)
        ref = O1.net_forward(ospec, P, img)
        got = net.predict(img.cuda()).cpu()
        assert float((got - ref).abs().max()) <= 5e-3, float((got - ref).abs().max())
        assert torch.equal(got.argmax(-1), ref.argmax(-1))


def test_gradient_history_metrics():
    """Reference trainer metrics (training/training.py:385-407): per-layer ||g||_2/size in one native launch, and the
    column layout of numerical_results/csv/*_gradient_history.csv."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    net = EulerNet(NetSpec(blocks_per_stage=(3, 2, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05), seed=2)
    gen = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8).cuda()
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float().cuda()
    loss = net.train_step(img, lab)
    norms = net.gradient_mean_norms()
    names = list(norms.keys())
    assert names[:3] == ["conv1_kernel_gradient_mean_norm", "res2_0_branch2_kernel_gradient_mean_norm",
                         "res2_1_branch2_kernel_gradient_mean_norm"]
    assert net.gradient_history_header().split(" ")[:4] == ["global_step", "mean_loss", "accuracy", "conv1_kernel_gradient_mean_norm"]
    g = net.grad.detach().double().cpu()
    a, shape = net.torch_params["conv1/kernel"]
    n = int(np.prod(shape))
    assert abs(norms[names[0]] - float(g[a:a + n].norm() / n)) <= 1e-6 * float(g[a:a + n].norm() / n)
    for (name, off, size, C), key in zip(net.layer_param_slices(), names[1:]):
        ref = float(g[off:off + size - C].norm() / (size - C))     # 19 kernel variables of a 16-channel layer, bias excluded
        assert abs(norms[key] - ref) <= 1e-5 * ref, key
    row = net.history_row(loss, net.predict(img), lab).split(" ")
    assert len(row) == 3 + len(names) and int(row[0]) == 1 and abs(float(row[1]) - float(loss)) < 1e-6


def test_prefetched_graph_step_equals_direct_step():
    """EulerNet.prefetch + train_step_graph_prefetched (copy stream, staging buffers, device-to-device hand-over) runs the
    same captured step on the same batches as train_step_graph(images, onehot): identical losses and parameters."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(2, 2, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
    gen = torch.Generator().manual_seed(11)
    batches = [(torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8).pin_memory(),
                torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float().pin_memory()) for _ in range(4)]
    nets = [EulerNet(NetSpec(**kw), precision="fast_f16", seed=4) for _ in range(2)]
    for n in nets:
        n.capture(batches[0][0].cuda(), batches[0][1].cuda(), warmup=1)
    la = [float(nets[0].train_step_graph(*b)) for b in batches]
    lb = []
    nets[1].prefetch(*batches[0])
    for i in range(len(batches)):
        out = nets[1].train_step_graph_prefetched()
        if i + 1 < len(batches):
            nets[1].prefetch(*batches[i + 1])          # overlaps the replay that was just queued
        lb.append(float(out))
    torch.cuda.synchronize()
    assert la == lb and torch.equal(nets[0].theta, nets[1].theta)


def test_train_step_with_max_pooling_stage_matches_oracle():
    """`use_max_pooling` (models/tfkeras_resnets.py:577-578): MaxPooling2D(2,2) in front of a stage, which then starts with a conv
    block; the Euler chains keep their CUDA kernels, the pooling itself is a torch op.  Loss and gradients vs O1."""
    from differential_equations_resnet_b200.training import EulerNet, NetSpec
    kw = dict(blocks_per_stage=(2, 3, 2), filters_per_block=(16, 16, 32), strides=((1, 1), (1, 1), (2, 2)), h=0.25, gamma=-0.05,
              use_max_pooling=[False, True, False])
    ospec = O1.NetSpec(**kw)
    P = O1.init_net_params(ospec, seed=5)
    net = EulerNet(NetSpec(**kw), precision="strict", seed=0)
    net.import_params(P)
    gen = torch.Generator().manual_seed(3)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=gen, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=gen), 10).float()
    M = {k: torch.zeros_like(v) for k, v in P.items()}
    V = {k: torch.zeros_like(v) for k, v in P.items()}
    lr, gr = O1.train_step(ospec, P, M, V, 1, img, lab)
    l = float(net.train_step(img.cuda(), lab.cuda()))
    assert abs(l - lr) <= 1e-5 * max(1.0, abs(lr))
    g = net.export_grads()
    for k in gr:
        err = float((g[k].double() - gr[k].double()).norm() / max(float(gr[k].double().norm()), 1e-30))
        assert err <= 1e-3, (k, err)     # max-pool argmax near-ties and relu branches at rounding level: measured 4e-4 at conv1
    p = net.predict(img.cuda()).cpu()
    assert float((p - O1.net_forward(ospec, {k: v for k, v in net.export_params().items()}, img)).abs().max()) <= 1e-4
