"""Batch-sharded data parallelism: one process per GPU, parameters and Adam
state replicated, the global batch split evenly over ranks, ONE flat fp32 bucket of *packed
free-parameter* gradients summed with an all-reduce (NCCL over NVLink on GPUs, gloo in CPU tests)
and scaled by 1/world_size inside the Adam kernel.  The reference has no counterpart (single
tf.Session, training/training.py:132 of the reference)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world_size: int):
    """[lo, hi) of this rank's images; the global batch must divide evenly (weak scaling keeps the
    per-rank batch fixed, so the train step launches identical kernels on every rank)."""
    if global_batch % world_size:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world_size))
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def allreduce_bucket(flat_grad: torch.Tensor, world_size: int):
    """Sum the flat gradient bucket over ranks in place (no-op for a single rank).  The mean is
    taken later by the optimiser's grad_scale = 1/world_size."""
    if world_size > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return flat_grad


def allreduce_async(flat_slice: torch.Tensor):
    """Start the sum all-reduce of one contiguous slice of the gradient bucket and return the work handle;
    the collective runs on the backend's own stream, so kernels launched afterwards on the caller's stream
    overlap it (the exchange is hidden behind the remaining backward pass).  Call
    `.wait()` on the handle before the optimiser reads the slice."""
    return dist.all_reduce(flat_slice, op=dist.ReduceOp.SUM, async_op=True)


def adam_reference_step(theta, grad_sum, m, v, t, world_size, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """Host restatement of b200ode_adam_step (tf.train.AdamOptimizer form) used by the CPU tests."""
    g = grad_sum / world_size
    lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    theta.sub_(lr_t * m / (v.sqrt() + eps))
    return theta
