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


class _AbiWork:
    """Handle of an all-reduce queued on the communicator's side stream (same `.wait()` as a torch work handle)."""

    def __init__(self, done):
        self._done = done

    def wait(self):
        torch.cuda.current_stream().wait_event(self._done)


class AbiComm:
    """Gradient exchange through the C ABI (`b200ode_comm_*`, include/b200ode.h): libb200ode binds NCCL itself, the
    all-reduce runs on a dedicated stream ordered against the caller's stream with events (capturable in a CUDA
    graph).  The 128-byte unique id travels from rank 0 through the already initialised `torch.distributed` group
    (any backend; only a broadcast of 128 bytes) or, without one, through `id_bytes` supplied by the launcher."""

    def __init__(self, rank: int, world_size: int, id_bytes: bytes = None, p2p: bool = False):
        """p2p=True: the trainer keeps its gradient bucket in peer-mapped memory (`shared_bucket`) and the all-reduce
        happens INSIDE the Adam kernel (`adam_step`: b200ode_comm_adam_step reads the N gradient replicas over NVLink in
        rank order); p2p=False: NCCL all-reduce of bucket slices on a side stream (`allreduce_async`)."""
        import ctypes
        from . import _abi
        lib = _abi.lib()
        self._lib, self._abi = lib, _abi
        self.rank, self.world_size = rank, world_size
        if id_bytes is None:
            buf = ctypes.create_string_buffer(128)
            if rank == 0:
                _abi.check(lib.b200ode_comm_unique_id(buf))
            box = [bytes(buf.raw)]
            if world_size > 1:
                dist.broadcast_object_list(box, src=0)
            id_bytes = box[0]
        self._id = ctypes.create_string_buffer(id_bytes, 128)
        h = ctypes.c_void_p()
        _abi.check(lib.b200ode_comm_init(world_size, rank, self._id, ctypes.byref(h)))
        self._h = h
        self.stream = torch.cuda.Stream()
        self.p2p = bool(p2p) and world_size > 1
        self._bucket = None

    def shared_bucket(self, n_floats: int) -> torch.Tensor:
        """This rank's gradient bucket inside the library's peer-mapped region (collective call: every rank, same size),
        as a zero-copy torch tensor."""
        import ctypes
        ptr, pptr = ctypes.c_void_p(), ctypes.c_void_p()
        self._abi.check(self._lib.b200ode_comm_shared_alloc(self._h, int(n_floats), ctypes.byref(ptr), ctypes.byref(pptr)))

        class _Dev:   # __cuda_array_interface__ view of the library-owned allocation (kept alive by the communicator)
            def __init__(self, p):
                self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(p), False), "version": 3,
                                                 "strides": None}
        self._bucket = torch.as_tensor(_Dev(ptr.value), device="cuda")
        # parameter replica in the same peer-mapped region: adam_step on it takes the two-shot form (sharded update)
        self.shared_params = torch.as_tensor(_Dev(pptr.value), device="cuda")
        return self._bucket

    def adam_step(self, theta, grad_slice, m, v, step_counter, lr, eps, stream=None):
        """Adam over one slice with the gradient summed over ranks from peer memory inside the kernel (1/world folded in)."""
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        self._abi.check(self._lib.b200ode_comm_adam_step(self._h, theta.data_ptr(), grad_slice.data_ptr(), m.data_ptr(), v.data_ptr(),
                                                         grad_slice.numel(), float(lr), 0.9, 0.999, float(eps),
                                                         step_counter.data_ptr(), st))

    def allreduce_async(self, flat_slice: torch.Tensor):
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.stream.wait_event(ready)
        self._abi.check(self._lib.b200ode_comm_allreduce_bucket(self._h, flat_slice.data_ptr(), flat_slice.numel(),
                                                                self.stream.cuda_stream))
        done = torch.cuda.Event()
        done.record(self.stream)
        return _AbiWork(done)

    def allreduce_bucket(self, flat_grad: torch.Tensor):
        if self.world_size > 1:
            self.allreduce_async(flat_grad).wait()
        return flat_grad

    def close(self):
        if self._h:
            self._lib.b200ode_comm_destroy(self._h)
            self._h = None


def adam_reference_step(theta, grad_sum, m, v, t, world_size, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """Host restatement of b200ode_adam_step (tf.train.AdamOptimizer form) used by the CPU tests."""
    g = grad_sum / world_size
    lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    theta.sub_(lr_t * m / (v.sqrt() + eps))
    return theta
