"""World-size-2 gloo test (CPU) of the data-parallel logic: batch sharding + one flat all-reduced
gradient bucket + 1/world scaling reproduce the single-process full-batch step.  The per-rank
gradient is produced by the CPU oracle (the GPU kernels cannot run here); what is under test is
the host-side sharding / bucket / optimiser-scaling path that EulerNet uses on GPUs."""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _flat_grads(spec, P, img, lab):
    from oracle import antisym_torch as O1
    names = sorted(P)
    leaves = [P[n].clone().requires_grad_(True) for n in names]
    loss = O1.loss_fn(O1.net_forward(spec, dict(zip(names, leaves)), img), lab)
    grads = torch.autograd.grad(loss, leaves)
    return torch.cat([g.reshape(-1) for g in grads]), float(loss)


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    from oracle import antisym_torch as O1
    from differential_equations_resnet_b200.parallel import adam_reference_step, allreduce_bucket, shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = O1.NetSpec(blocks_per_stage=(1, 2, 1), h=0.5, gamma=-0.1)
    P = O1.init_net_params(spec, seed=3)
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=g, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=g), 10).float()
    lo, hi = shard_bounds(8, rank, world)
    flat, _ = _flat_grads(spec, P, img[lo:hi], lab[lo:hi])
    allreduce_bucket(flat, world)
    theta = torch.cat([P[n].reshape(-1) for n in sorted(P)])
    m, v = torch.zeros_like(theta), torch.zeros_like(theta)
    adam_reference_step(theta, flat, m, v, 1, world)
    if rank == 0:
        torch.save({"grad_mean": flat / world, "theta": theta}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_matches_full_batch(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import antisym_torch as O1
    from differential_equations_resnet_b200.parallel import adam_reference_step, shard_bounds
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out), nprocs=2, join=True)
    got = torch.load(out)
    spec = O1.NetSpec(blocks_per_stage=(1, 2, 1), h=0.5, gamma=-0.1)
    P = O1.init_net_params(spec, seed=3)
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (8, 32, 32, 3), generator=g, dtype=torch.uint8)
    lab = torch.nn.functional.one_hot(torch.randint(0, 10, (8,), generator=g), 10).float()
    full, _ = _flat_grads(spec, P, img, lab)
    # mean CE over the global batch == mean of the per-shard means for equal shards
    assert torch.allclose(got["grad_mean"], full, rtol=1e-4, atol=1e-6)
    theta = torch.cat([P[n].reshape(-1) for n in sorted(P)])
    adam_reference_step(theta, full, torch.zeros_like(theta), torch.zeros_like(theta), 1, 1)
    assert torch.allclose(got["theta"], theta, rtol=0, atol=2e-6)
    assert shard_bounds(256, 3, 8) == (96, 128)


def _syncbn_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch.distributed as dist
    from differential_equations_resnet_b200.training import _SyncStats
    from differential_equations_resnet_b200.parallel import shard_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)

    class Net:                      # the two things _SyncStats needs from EulerNet
        world_size = world

        @staticmethod
        def _ar(t):
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return t
    g = torch.Generator().manual_seed(11)
    z_all = torch.randn(8, 5, 4, 6, generator=g) * 2.0 + 0.5
    w_all = torch.randn(8, 5, 4, 6, generator=g)
    lo, hi = shard_bounds(8, rank, world)
    z = z_all[lo:hi].clone().requires_grad_(True)
    st = _SyncStats.apply(z, Net)
    y = (z - st[0]) / torch.sqrt(st[1] + 1e-3)
    # every rank differentiates its LOCAL loss (mean over its shard); the trainer later averages the gradients over ranks
    loss = (y * w_all[lo:hi]).sum() / (hi - lo)
    loss.backward()
    torch.save({"stats": st.detach(), "dz": z.grad, "lo": lo, "hi": hi}, out + ".%d" % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_syncbn_statistics_match_full_batch(tmp_path):
    """SyncBN protocol of the BN path: local sums -> all-reduce of 2C floats -> global statistics, backward
    all-reduces (d mean, d var); with the trainer's convention (per-rank loss = mean over the local shard, gradients
    averaged over ranks) the input gradient equals the full-batch one."""
    sys.path.insert(0, ROOT)
    out = str(tmp_path / "sb")
    mp.spawn(_syncbn_worker, args=(2, 31500 + os.getpid() % 2000, out), nprocs=2, join=True)
    g = torch.Generator().manual_seed(11)
    z_all = (torch.randn(8, 5, 4, 6, generator=g) * 2.0 + 0.5).requires_grad_(True)
    w_all = torch.randn(8, 5, 4, 6, generator=g)
    mu, var = z_all.mean(dim=(0, 1, 2)), z_all.var(dim=(0, 1, 2), unbiased=False)
    y = (z_all - mu) / torch.sqrt(var + 1e-3)
    ((y * w_all).sum() / 8).backward()
    for r in range(2):
        got = torch.load(out + ".%d" % r)
        assert torch.allclose(got["stats"][0], mu.detach(), rtol=1e-5, atol=1e-6)
        assert torch.allclose(got["stats"][1], var.detach(), rtol=1e-4, atol=1e-6)
        # per-rank gradient is world x the global-loss gradient (local mean vs global mean); the 1/world of the optimiser undoes it
        assert torch.allclose(got["dz"] / 2, z_all.grad[got["lo"]:got["hi"]], rtol=1e-4, atol=1e-6)
