"""Train-step engine for the antisymmetric single-block ResNet.

Mirrors the *step semantics* of the reference trainer (training/training.py:283-304 of the
reference: softmax output -> mean Keras categorical cross-entropy -> gradients -> Adam with
epsilon 1e-7) and the model assembly of models/tfkeras_resnets.py:511-604 for
kernel_type='antisymmetric'.  It is not a port of the TF1 session plumbing.

Hot path (hand-written CUDA, libb200ode.so): every Euler step
x_{n+1} = x_n + h*relu(conv_K(x_n)+b) forward and backward, weight packing, gradient fold,
Adam.  Stem / transition / head (<= 5 layers per net) run as torch ops.

Data parallel: one process per GPU; parameters replicated, batch sharded, ONE flat fp32 gradient
bucket (packed free parameters, not dense kernels) all-reduced with NCCL via torch.distributed.
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.nn.functional as F

from . import _abi
from .layers._base import (BN_EPS, BN_MOMENTUM, BNEulerStep, ChainHandle, LayerHandle, truncated_normal_, _alloc_workspace, _ptr,
                           _stream_ptr)
from .parallel import allreduce_async, allreduce_bucket


class NetSpec:
    """Keyword mirror of get_single_block_resnet_build_function (models/tfkeras_resnets.py:511-527)."""

    def __init__(self, num_stages=4, blocks_per_stage=(3, 3, 3), filters_per_block=(16, 32, 64),
                 strides=((1, 1), (2, 2), (2, 2)), h=1.0, gamma=0.0, num_classes=10, use_batch_norm=False,
                 subtract_mean=127.5, divide_by_stddev=127.5, kernel_size=3, in_channels=3, use_max_pooling=None):
        self.use_batch_norm = bool(use_batch_norm)
        # MaxPooling2D(2,2) in front of stage s+2 (models/tfkeras_resnets.py:577-578); such a stage starts with a conv block
        self.use_max_pooling = list(use_max_pooling) if use_max_pooling is not None else [False] * (num_stages - 1)
        if kernel_size != 3:
            raise ValueError("antisymmetric Euler blocks are 3x3")
        self.num_stages, self.blocks_per_stage = num_stages, list(blocks_per_stage)
        self.filters_per_block, self.strides = list(filters_per_block), [tuple(s) for s in strides]
        self.h, self.gamma, self.num_classes = float(h), float(gamma), num_classes
        self.subtract_mean, self.divide_by_stddev = subtract_mean, divide_by_stddev
        self.kernel_size, self.in_channels = kernel_size, in_channels

    def plan(self):
        """Graph order of (kind, C_in, C_out, stride, name); stage loop models/tfkeras_resnets.py:575-593."""
        fp, st = self.filters_per_block, self.strides
        ops = [("stem", self.in_channels, fp[0], st[0], "conv1")]
        for s in range(self.num_stages - 1):
            pool = self.use_max_pooling[s]
            if pool:
                c = fp[s - 1] if s > 0 else fp[0]
                ops.append(("maxpool", c, c, (2, 2), "stage%d_pooling" % (s + 2)))
            if not pool and (s == 0 or (fp[s] == fp[s - 1] and st[s] == (1, 1))):
                for b in range(self.blocks_per_stage[s]):
                    ops.append(("euler", fp[s], fp[s], (1, 1), "res%d_%d_branch2" % (s + 2, b)))
            else:
                ops.append(("transition", fp[s - 1] if s > 0 else fp[0], fp[s], st[s], "res%d_0_branch" % (s + 2)))
                for b in range(1, self.blocks_per_stage[s]):
                    ops.append(("euler", fp[s], fp[s], (1, 1), "res%d_%d_branch2" % (s + 2, b)))
        return ops


def _same_pad(in_size, k, s):
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv2d_same_nhwc(x, w_hwio, bias, strides):
    """Regular Keras Conv2D(padding='same') on NHWC tensors with TF's asymmetric SAME padding."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = _same_pad(x.shape[1], kh, strides[0])
    pl, pr = _same_pad(x.shape[2], kw, strides[1])
    xn = x.permute(0, 3, 1, 2)
    if pt != pb or pl != pr:
        xn = F.pad(xn, (pl, pr, pt, pb))
        pad = (0, 0)
    else:
        pad = (pt, pl)
    y = F.conv2d(xn, w_hwio.permute(3, 2, 0, 1), bias, stride=strides, padding=pad)
    return y.permute(0, 2, 3, 1).contiguous()


class _Chain:
    """One run of consecutive Euler steps with equal shape: handles + saved activations."""

    def __init__(self, C, n_layers, gamma, precision, offset, persistent=True, bn=False):
        self.C, self.n, self.gamma, self.precision = C, n_layers, gamma, precision
        # BatchNorm needs the batch statistics of every step before its activation: per-layer kernels (the conv kernel
        # emits the partial sums from its epilogue), not the one-image-per-CTA persistent chains
        self.bn = bn
        self.bn_offset = None                     # offset of [gamma_0, beta_0, gamma_1, ...] (C floats each) in the flat bucket
        self.bn_moving = None                     # [n, 2, C] moving mean / variance (not trained)
        if bn and precision == "fast_bf16":
            raise ValueError("use_batch_norm=True needs fp32 activations (precision strict / fast_tf32 / fast_f16 / simt)")
        self.persistent = persistent and precision in _abi.CHAIN_PRECISIONS and not bn
        self.chain_prec = _abi.CHAIN_PRECISIONS.get(precision, _abi.PREC_FAST_TF32)
        # per-layer fallback of the fp16 chains (images too large for shared memory): the FAST_TF32 kernels
        self.layer_prec = _abi.PRECISIONS["fast_tf32" if precision == "fast_f16" else precision]
        self._handles = None
        self.fused = None                         # ChainHandle once a shape that fits shared memory is seen
        self.np_layer = 4 * C + 9 * C * (C - 1) // 2 + C
        self.offset = offset                      # offset of layer 0 in the flat Euler bucket
        self.acts = None
        self.masks = None
        self._fshape = None
        self._fbufs = {}                          # shape -> buffer set (kept: a captured CUDA graph holds their addresses)
        self._last_fused = False
        self.generation = 0                       # bumped by every forward: a stale backward must not read newer state

    @property
    def handles(self):
        """Per-layer handles (fallback path: strict mode, or images too large for shared memory)."""
        if self._handles is None:
            self._handles = [LayerHandle(self.C, 3, self.gamma, (1, 1), True, True, self.layer_prec,
                                         _abi.LAYOUT_3BY3) for _ in range(self.n)]
            assert self._handles[0].num_params == self.np_layer
        return self._handles

    def use_fused(self, shape):
        if not self.persistent or not ChainHandle.supported(self.C, shape[1], shape[2], self.chain_prec):
            return False
        if self.fused is None:
            self.fused = ChainHandle(self.C, self.n, self.gamma, precision=self.chain_prec)
            assert self.fused.num_params == self.np_layer
        return True

    def ensure_fused_buffers(self, shape, device):
        """Buffer sets are keyed by shape and never replaced: a CUDA graph captured at one batch size keeps valid
        addresses when an eager call with another shape (a partial eval batch) comes in between replays."""
        if self._fshape == shape:
            return
        if shape not in self._fbufs:
            N, H, W, C = shape
            sdt = self.fused.saved_dtype
            self._fbufs[shape] = dict(
                f_acts=torch.empty((self.n,) + shape, dtype=sdt, device=device),
                f_masks=torch.empty((self.n, N, H, W, (C + 7) // 8), dtype=torch.uint8, device=device),
                f_dz=torch.empty((self.n,) + shape, dtype=sdt, device=device),
                f_dx=torch.empty(shape, dtype=torch.float32, device=device),
                f_y=torch.empty(shape, dtype=torch.float32, device=device) if self.fused.f16 else None,
                f_ws=_alloc_workspace(self.fused.workspace_bytes(N, H, W), device))   # split-K partials of the chain wgrad
        for k, v in self._fbufs[shape].items():
            setattr(self, k, v)
        self._fshape = shape

    def saved_mask(self, l):
        """relu bit mask [N,H,W,C/8] of Euler step l as saved by the last forward pass"""
        return self.f_masks[l] if self._last_fused else self.masks[l]

    def fused_forward(self, x, h):
        """All Euler steps of the chain in ONE launch (training form: saves the weight-gradient operands and relu
        masks); returns the stage output."""
        self.x0 = x
        self._last_fused = True
        if self.fused.f16:          # acts[l] = fp16 input of step l; the fp32 output of the last step goes to f_y
            self.fused.forward(x, h, acts=self.f_acts, masks=self.f_masks, y_final=self.f_y)
            return self.f_y
        self.fused.forward(x, h, acts=self.f_acts, masks=self.f_masks)
        return self.f_acts[self.n - 1]

    def fused_dgrad(self, dy, h, dy_amax=None):
        """Backward sweep over all steps of the chain (one launch); returns dL/dx0."""
        self.fused.dgrad(dy, self.f_masks, self.f_dz, self.f_dx, h, dy_amax=dy_amax)
        return self.f_dx

    def fused_wgrad(self, grad_euler):
        """Weight + bias gradients of all layers (one launch + fold) into the flat bucket."""
        if getattr(self.fused, "_ws", None) is not self.f_ws:
            self.fused.bind_workspace(self.f_ws)
        self.fused.wgrad(self.x0, self.f_acts, self.f_dz, grad_euler[self.offset:], self.np_layer)

    def fused_backward(self, dy, h, grad_euler):
        dx = self.fused_dgrad(dy, h)
        self.fused_wgrad(grad_euler)
        return dx

    def ensure_buffers(self, shape, device):
        if self.acts is not None and self.acts[1].shape == shape:
            return
        N, H, W, C = shape
        dt = self.handles[0].io_dtype                 # fp32, or bf16 in fast_bf16 mode
        self.acts = [None] + [torch.empty(shape, dtype=dt, device=device) for _ in range(self.n)]
        self.masks = [torch.empty((N, H, W, (C + 7) // 8), dtype=torch.uint8, device=device) for _ in range(self.n)]
        self.dz = [torch.empty(shape, dtype=dt, device=device) for _ in range(2)]   # dZ_l lives in dz[l & 1]
        self.dx = [torch.empty(shape, dtype=dt, device=device) for _ in range(2)]
        # one workspace for all per-layer handles of the chain (their calls are serialised on one stream)
        self.ws = _alloc_workspace(self.handles[0].workspace_bytes(N, H, W), device)
        for hd in self.handles:
            hd.bind_workspace(self.ws)
        if self.bn:   # pre-activations and batch statistics of every step (saved for the backward pass)
            self.z = [torch.empty(shape, dtype=torch.float32, device=device) for _ in range(self.n)]
            self.stat = torch.empty((self.n, 6, C), dtype=torch.float32, device=device)
            self.stats_ws = BNEulerStep.stats_workspace(C, device)


class _ChainFn(torch.autograd.Function):
    """All Euler steps of a chain as ONE autograd node; parameter gradients are written by the
    wgrad/fold kernels straight into the flat gradient bucket (side effect), not returned."""

    @staticmethod
    def forward(ctx, x, chain, net):
        lib, st = _abi.lib(), _stream_ptr()
        x = x.contiguous()
        N, H, W, C = x.shape
        ctx.chain, ctx.net, ctx.shape = chain, net, (N, H, W, C)
        chain.generation += 1                     # saved state lives in per-chain buffers: ONE forward in flight
        ctx.generation = chain.generation
        ctx.fused = chain._last_fused = chain.use_fused(tuple(x.shape))
        if ctx.fused:
            # persistent path: pack all layers, then ONE launch runs every Euler step of the chain
            chain.ensure_fused_buffers(tuple(x.shape), x.device)
            chain.fused.pack(net.theta_euler[chain.offset:], chain.np_layer)
            return chain.fused_forward(x.detach(), net.spec.h).view(N, H, W, C)
        chain.ensure_buffers(tuple(x.shape), x.device)
        dt = chain.handles[0].io_dtype
        ctx.in_dtype = x.dtype
        chain.acts[0] = x.detach().to(dt)         # layer 0 reads the caller's tensor (no copy when dtypes agree)
        for l, hd in enumerate(chain.handles):
            off = chain.offset + l * chain.np_layer
            _abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(net.theta_euler[off:]), None, st))
            if chain.bn:
                bo = chain.bn_offset + 2 * C * l
                BNEulerStep.forward(hd, chain.acts[l], net.theta[bo:bo + C], net.theta[bo + C:bo + 2 * C], chain.bn_moving[l, 0],
                                    chain.bn_moving[l, 1], net.spec.h, chain.z[l], chain.acts[l + 1], chain.stat[l], chain.stats_ws,
                                    net._bn_allreduce, net.world_size)
                continue
            _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(chain.acts[l]), _ptr(chain.acts[l + 1]), _ptr(chain.masks[l]),
                                             None, N, H, W, net.spec.h, _abi.F_EULER, st))
        return chain.acts[chain.n].view(N, H, W, C).to(ctx.in_dtype)

    @staticmethod
    def backward(ctx, dy):
        chain, net = ctx.chain, ctx.net
        if ctx.generation != chain.generation:
            raise RuntimeError("EulerNet chains keep the saved activations of ONE forward pass: another forward ran "
                               "before this backward (gradient accumulation / interleaved eval needs a second EulerNet)")
        N, H, W, C = ctx.shape
        lib, st = _abi.lib(), _stream_ptr()
        dy = dy.contiguous()
        if ctx.fused:
            return chain.fused_backward(dy, net.spec.h, net.grad_euler).view(N, H, W, C), None, None
        dt = chain.handles[0].io_dtype
        cur = dy.to(dt)
        is_bf16 = int(dt == torch.bfloat16)
        top = chain.n - 1
        if chain.bn:
            for l in range(top, -1, -1):
                off = chain.offset + l * chain.np_layer
                bo = chain.bn_offset + 2 * C * l
                nxt = chain.dx[l & 1]
                BNEulerStep.backward(chain.handles[l], chain.acts[l], cur, chain.z[l], chain.stat[l], net.theta[bo:bo + C], net.spec.h,
                                     chain.dz[0], nxt, net.grad_euler[off:], net.grad[bo:bo + 2 * C].view(2, C), chain.stats_ws,
                                     net._bn_allreduce, net.world_size)
                cur = nxt
            return cur.view(N, H, W, C).to(ctx.in_dtype), None, None
        _abi.check(lib.b200ode_relu_scale_bwd(_ptr(cur), _ptr(chain.masks[top]), _ptr(chain.dz[top & 1]), N * H * W, C,
                                              net.spec.h, is_bf16, st))
        for l in range(top, -1, -1):
            hd = chain.handles[l]
            off = chain.offset + l * chain.np_layer
            nxt = chain.dx[l & 1]
            if l > 0:
                # dY_{l-1} and dZ_{l-1} = h * dY_{l-1} * mask_{l-1} in ONE pass (the epilogue of the data-gradient kernel)
                _abi.check(lib.b200ode_euler_dgrad_fused(hd._h, _ptr(chain.dz[l & 1]), _ptr(cur), _ptr(nxt), _ptr(chain.masks[l - 1]),
                                                         _ptr(chain.dz[(l - 1) & 1]), net.spec.h, N, H, W, st))
            else:
                _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(chain.dz[0]), _ptr(cur), _ptr(nxt), N, H, W, st))
            _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(chain.acts[l]), _ptr(chain.dz[l & 1]), _ptr(net.grad_euler[off:]),
                                               None, N, H, W, 0, st))
            cur = nxt
        return cur.view(N, H, W, C).to(ctx.in_dtype), None, None


class _SyncStats(torch.autograd.Function):
    """Per-channel (mean, biased variance) of an NHWC tensor over the GLOBAL batch: local sums -> all-reduce of 2C
    floats -> statistics; backward all-reduces (d mean, d var) the same way (every rank's loss sees the shared statistics)."""

    @staticmethod
    def forward(ctx, z, net):
        C = z.shape[-1]
        Mg = (z.numel() // C) * net.world_size
        s = torch.stack([z.sum(dim=(0, 1, 2)), (z * z).sum(dim=(0, 1, 2))])
        net._ar(s)
        mu = s[0] / Mg
        var = (s[1] / Mg - mu * mu).clamp_min(0.0)
        ctx.save_for_backward(z, mu)
        ctx.net, ctx.Mg = net, Mg
        return torch.stack([mu, var])

    @staticmethod
    def backward(ctx, g):
        z, mu = ctx.saved_tensors
        g = g.contiguous().clone()
        ctx.net._ar(g)
        return (g[0] + 2.0 * g[1] * (z - mu)) / ctx.Mg, None


class EulerNet:
    """Antisymmetric single-block ResNet with a fused train step.

    precision: 'strict' (3xTF32, fp32-accurate), 'fast_f16' (persistent chains with fp16 operands -- the tf32
    significand -- around an fp32 residual stream), 'fast_tf32' (persistent chains with tf32 operands),
    'fast_bf16' (bf16 operands and activations, fp32 accumulate; per-layer kernels), or 'simt'."""

    def __init__(self, spec: NetSpec, precision="fast_tf32", device="cuda", seed=0, lr=1e-3, adam_eps=1e-7,
                 world_size=1, persistent=True, native_glue=True, comm=None, sync_bn=True, split_first_chain=None,
                 glue_fast=None):
        _abi.require_device()
        # transition blocks: one tf32 MMA on round-to-nearest operands in the fast modes (the grade of their chains' fp16 / tf32 /
        # bf16 operands), the 3xTF32 split (fp32 grade) in strict mode; glue_fast=False keeps the fp32-grade kernels everywhere
        self.glue_fast = (precision in ("fast_f16", "fast_tf32", "fast_bf16")) if glue_fast is None else bool(glue_fast)
        if split_first_chain is None:
            split_first_chain = int(os.environ.get("B200ODE_SPLIT_FIRST_CHAIN", "1"))
        self.spec, self.precision, self.device = spec, precision, torch.device(device)
        self.lr, self.adam_eps, self.world_size = lr, adam_eps, world_size
        # gradient exchange: torch.distributed (default) or a parallel.AbiComm (NCCL bound by libb200ode itself)
        self.comm = comm
        gen = torch.Generator().manual_seed(seed)
        plan = spec.plan()
        # ---- flat Euler bucket -------------------------------------------------------------------
        self.segments = []          # graph order: ('torch', name, ...) or ('chain', _Chain)
        off = 0
        i = 0
        while i < len(plan):
            kind, ci, co, st, name = plan[i]
            if kind == "euler":
                j = i
                while j < len(plan) and plan[j][0] == "euler" and plan[j][2] == co:
                    j += 1
                # The FIRST stage may run as `split_first_chain` back-to-back chains (opt-in, B200ODE_SPLIT_FIRST_CHAIN): its weight
                # gradient is the last big kernel of the step with nothing left to hide behind, and as two halves the upper half's
                # weight gradient runs on the side stream under the lower half's backward sweep.  MEASURED SLOWER on cfg3 (1.011 ->
                # 1.030 ms with 2 parts, 1.050 ms with 3: the side-stream CTAs find no free SM under the chain kernels, and every
                # part costs six more launches), so the default stays one chain per stage.
                parts = split_first_chain if (split_first_chain > 1 and not self.segments_have_chain() and j - i >= 2 * split_first_chain
                                             and not spec.use_batch_norm) else 1
                n_left = j - i
                for part in range(parts):
                    n_part = n_left // (parts - part)
                    ch = _Chain(co, n_part, spec.gamma, precision, off, persistent, bn=spec.use_batch_norm)
                    off += ch.np_layer * ch.n
                    self.segments.append(("chain", ch))
                    n_left -= n_part
                i = j
            else:
                self.segments.append((kind, ci, co, st, name))
                i += 1
        self.n_euler_params = off
        # ---- regular (torch-op) parameters, appended to the same flat buffers --------------------
        self.torch_shapes = []
        k = spec.kernel_size
        for seg in self.segments:
            if seg[0] == "stem":
                _, ci, co, st, name = seg
                self.torch_shapes += [(name + "/kernel", (k, k, ci, co), k * k * ci), (name + "/bias", (co,), 0)]
                if spec.use_batch_norm:      # bn_conv1 (models/tfkeras_resnets.py:570-571); fan_in -1 = ones
                    self.torch_shapes += [("bn_" + name + "/gamma", (co,), -1), ("bn_" + name + "/beta", (co,), 0)]
            elif seg[0] == "maxpool":
                continue
            elif seg[0] == "transition":
                _, ci, co, st, name = seg
                self.torch_shapes += [(name + "2/kernel", (k, k, ci, co), k * k * ci), (name + "2/bias", (co,), 0),
                                      (name + "1/kernel", (1, 1, ci, co), ci), (name + "1/bias", (co,), 0)]
                if spec.use_batch_norm:      # bn{s}_0_branch2 / branch1 (models/tfkeras_resnets.py:258-263)
                    bn = name.replace("res", "bn")
                    for br in ("2", "1"):
                        self.torch_shapes += [(bn + br + "/gamma", (co,), -1), (bn + br + "/beta", (co,), 0)]
        c_last = spec.filters_per_block[spec.num_stages - 2]
        self.torch_shapes += [("fc/kernel", (c_last, spec.num_classes), c_last), ("fc/bias", (spec.num_classes,), 0)]
        n_torch = sum(math.prod(s) for _, s, _ in self.torch_shapes)
        # BatchNorm scale / offset of the Euler steps (bn{s}_{b}_branch2): per chain [layer][gamma (C) | beta (C)] behind the
        # regular parameters, in the same flat bucket (same Adam launch, same all-reduce)
        n_bn = 0
        for seg in self.segments:
            if seg[0] == "chain" and seg[1].bn:
                if (off + n_torch + n_bn) % 4:     # 16-byte aligned slices: the BN kernels read gamma / beta as float4
                    n_bn += 4 - (off + n_torch + n_bn) % 4
                seg[1].bn_offset = off + n_torch + n_bn
                n_bn += seg[1].n * 2 * seg[1].C
        self.n_params = (off + n_torch + n_bn + 3) // 4 * 4      # whole 16-byte vectors (the pad is never read by a layer)
        theta = torch.zeros(self.n_params, dtype=torch.float32)
        for seg in self.segments:
            if seg[0] == "chain" and seg[1].bn:
                ch = seg[1]
                theta[ch.bn_offset:ch.bn_offset + ch.n * 2 * ch.C].view(ch.n, 2, ch.C)[:, 0, :] = 1.0
        # Euler layers: truncated normal sigma = sqrt(2/(9C)), zero bias (reference 3By3.py:95-98,148-153)
        for seg in self.segments:
            if seg[0] != "chain":
                continue
            ch = seg[1]
            for l in range(ch.n):
                a = ch.offset + l * ch.np_layer
                truncated_normal_(theta[a:a + ch.np_layer - ch.C], math.sqrt(2.0 / (9 * ch.C)), gen)
        cur = off
        self.torch_params = {}
        for name, shape, fan_in in self.torch_shapes:
            n = math.prod(shape)
            if fan_in > 0:  # Keras he_normal (VarianceScaling(2, fan_in, truncated normal))
                std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
                truncated_normal_(theta[cur:cur + n], std, gen)
            elif fan_in < 0:
                theta[cur:cur + n] = 1.0
            self.torch_params[name] = (cur, shape)
            cur += n
        self.theta = theta.to(self.device)
        self.p2p = bool(getattr(comm, "p2p", False)) and world_size > 1
        # peer-memory exchange: the gradient bucket lives in the communicator's peer-mapped region, the all-reduce happens
        # inside the Adam kernel (parallel.AbiComm.adam_step)
        self.grad = comm.shared_bucket(self.n_params) if self.p2p else torch.zeros_like(self.theta)
        if self.p2p and not os.environ.get("B200ODE_P2P_ONESHOT"):
            # parameters in the peer-mapped region too: two-shot exchange (this rank reduces and updates 1/world of the bucket
            # and writes the new parameters into every replica; its Adam moments are the only ones it touches)
            comm.shared_params.copy_(self.theta)
            self.theta = comm.shared_params
        self.adam_m = torch.zeros_like(self.theta)
        self.adam_v = torch.zeros_like(self.theta)
        self.theta_euler = self.theta[:off]
        self.grad_euler = self.grad[:off]
        self.step_counter = torch.ones(1, dtype=torch.int32, device=self.device)
        # leaf views of the flat parameter buffer for the torch ops
        self.leaves = {}
        for name, (a, shape) in self.torch_params.items():
            n = math.prod(shape)
            self.leaves[name] = self.theta[a:a + n].view(shape).detach().requires_grad_(True)
        self._leaf_list = list(self.leaves.values())
        self._leaf_grad_views = [self.grad[a:a + math.prod(shape)].view(shape) for a, shape in self.torch_params.values()]
        self._graph = None
        self._static_in = None
        self.native_glue = native_glue
        self._nb = None
        self._nb_cache = {}
        self._pending, self._reduced_upto = [], self.n_euler_params     # overlapped all-reduces of this step
        # BatchNorm state: moving statistics (not trained) and the SyncBN hook (statistics over the GLOBAL batch)
        self.training = True
        self.bn_moving = {}
        for seg in self.segments:
            if seg[0] == "chain" and seg[1].bn:
                ch = seg[1]
                ch.bn_moving = torch.zeros((ch.n, 2, ch.C), dtype=torch.float32, device=self.device)
                ch.bn_moving[:, 1, :] = 1.0
        for name, shape, _ in self.torch_shapes:
            if name.endswith("/gamma"):
                self.bn_moving[name[:-6]] = [torch.zeros(shape, device=self.device), torch.ones(shape, device=self.device)]
        self.sync_bn = bool(sync_bn) and world_size > 1 and spec.use_batch_norm
        self._bn_allreduce = (lambda t: self._ar(t)) if self.sync_bn else None

    # ----------------------------------------------------------------------------------------------
    def forward(self, images):
        """images: uint8 or float [N,H,W,3] -> softmax probabilities."""
        spec = self.spec
        x = images.to(torch.float32)
        if spec.subtract_mean is not None:
            x = x - spec.subtract_mean
        if spec.divide_by_stddev is not None:
            x = x / spec.divide_by_stddev
        L = self.leaves
        bn = self._glue_bn if spec.use_batch_norm else (lambda t, name: t)
        for seg in self.segments:
            if seg[0] == "stem":
                _, ci, co, st, name = seg
                x = torch.relu(bn(conv2d_same_nhwc(x, L[name + "/kernel"], L[name + "/bias"], st), "bn_" + name))
            elif seg[0] == "maxpool":       # Keras MaxPooling2D((2,2)): torch op (the native step has no pooling kernel)
                x = F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()
            elif seg[0] == "transition":
                _, ci, co, st, name = seg
                bname = name.replace("res", "bn")
                main = bn(conv2d_same_nhwc(x, L[name + "2/kernel"], L[name + "2/bias"], st), bname + "2")
                short = bn(conv2d_same_nhwc(x, L[name + "1/kernel"], L[name + "1/bias"], st), bname + "1")
                x = torch.relu(main) + short          # models/tfkeras_resnets.py:266-267
            elif seg[1].bn and not self.training:
                x = self._chain_bn_inference(seg[1], x)
            else:
                x = _ChainFn.apply(x, seg[1], self)
        x = x.mean(dim=(1, 2))
        logits = x @ L["fc/kernel"] + L["fc/bias"]
        return torch.softmax(logits, dim=-1)

    def _chain_bn_inference(self, chain, x):
        """Euler steps with BatchNorm in inference mode (moving statistics): conv kernel + fused affine/relu/h/residual tail."""
        x = x.detach().contiguous().float()
        chain.ensure_buffers(tuple(x.shape), x.device)
        C, lib, st = chain.C, _abi.lib(), _stream_ptr()
        cur = x
        for l, hd in enumerate(chain.handles):
            off, bo = chain.offset + l * chain.np_layer, chain.bn_offset + 2 * C * l
            _abi.check(lib.b200ode_pack_kernel(hd._h, _ptr(self.theta_euler[off:]), None, st))
            BNEulerStep.inference(hd, cur, self.theta[bo:bo + C], self.theta[bo + C:bo + 2 * C], chain.bn_moving[l, 0],
                                  chain.bn_moving[l, 1], self.spec.h, chain.z[l], chain.acts[l + 1])
            cur = chain.acts[l + 1]
        return cur.clone()

    def _glue_bn(self, z, name):
        """BatchNormalization(axis=3) of a stem / transition branch as torch ops (SURVEY 8f-1 glue): batch statistics
        (over the global batch under SyncBN) in training mode, moving statistics otherwise."""
        g, b = self.leaves[name + "/gamma"], self.leaves[name + "/beta"]
        mm, mv = self.bn_moving[name]
        if not self.training:
            return g * (z - mm) / torch.sqrt(mv + BN_EPS) + b
        M = z.numel() // z.shape[-1]
        if self.sync_bn:
            stats = _SyncStats.apply(z, self)
            mu, var = stats[0], stats[1]
            M = M * self.world_size
        else:
            mu, var = z.mean(dim=(0, 1, 2)), z.var(dim=(0, 1, 2), unbiased=False)
        with torch.no_grad():
            mm.mul_(BN_MOMENTUM).add_(mu.detach(), alpha=1 - BN_MOMENTUM)
            mv.mul_(BN_MOMENTUM).add_(var.detach() * (M / max(M - 1, 1)), alpha=1 - BN_MOMENTUM)
        return g * (z - mu) / torch.sqrt(var + BN_EPS) + b

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    @staticmethod
    def loss_fn(probs, onehot, eps=1e-7):
        """mean K.categorical_crossentropy(from_logits=False) (training/training.py:295)."""
        p = probs / probs.sum(dim=-1, keepdim=True)
        p = torch.clamp(p, eps, 1.0 - eps)
        return -(onehot * torch.log(p)).sum(dim=-1).mean()

    # ----------------------------------------------------------------------------------------------
    # all-native step: stem / transition / head kernels of libb200ode (include/b200ode.h) around the
    # persistent Euler chains; no torch autograd, ~25 launches per step
    # ----------------------------------------------------------------------------------------------
    def _native_plan(self, shape, device):
        """Per-input-shape buffers of the native path, or None when some layer cannot take it."""
        if self._nb is not None and self._nb["shape"] == tuple(shape):
            return self._nb
        if tuple(shape) in self._nb_cache:         # kept per shape (CUDA graphs hold the addresses of their set)
            self._nb = self._nb_cache[tuple(shape)]
            for e in self._nb["plan"] or []:
                if e["kind"] == "chain":
                    e["chain"].ensure_fused_buffers((shape[0], e["h"], e["w"], e["c"]), device)
            return self._nb
        spec = self.spec
        N, H, W, Cin = shape
        ok = self.native_glue and spec.kernel_size == 3 and spec.num_classes <= 32 and not spec.use_batch_norm and not any(spec.use_max_pooling)
        plan, h, w, c = [], H, W, Cin
        for seg in self.segments:
            if not ok:
                break
            if seg[0] == "stem":
                _, ci, co, st, name = seg
                ok = ok and tuple(st) == (1, 1) and co % 4 == 0 and ci == c
                plan.append(dict(kind="stem", name=name, ci=ci, co=co, h=h, w=w,
                                 out=torch.empty((N, h, w, co), dtype=torch.float32, device=device),
                                 ws=self._glue_ws(ok, _abi.GLUE_STEM_WGRAD, N, h, w, ci, co, 1, 1, device)))
                c = co
            elif seg[0] == "transition":
                _, ci, co, st, name = seg
                ho, wo = -(-h // st[0]), -(-w // st[1])
                lanes = 256 // co if co and 256 % co == 0 else 0
                ok = ok and ci % 4 == 0 and co % 8 == 0 and lanes > 0 and ci % lanes == 0 and ci // lanes <= 16
                plan.append(dict(kind="transition", name=name, ci=ci, co=co, h=h, w=w, st=tuple(st),
                                 out=torch.empty((N, ho, wo, co), dtype=torch.float32, device=device),
                                 mask=torch.empty((N, ho, wo, co // 8), dtype=torch.uint8, device=device),
                                 dx=torch.empty((N, h, w, ci), dtype=torch.float32, device=device),
                                 ws=self._glue_ws(ok, _abi.GLUE_TRANSITION_WGRAD, N, h, w, ci, co, st[0], st[1], device)))
                h, w, c = ho, wo, co
            else:
                ch = seg[1]
                ok = ok and ch.use_fused((N, h, w, c))
                if ok:
                    ch.ensure_fused_buffers((N, h, w, c), device)
                plan.append(dict(kind="chain", chain=ch, h=h, w=w, c=c))
        ok = ok and c % 32 == 0 and c <= 1024 and spec.num_classes <= c
        if not ok:
            self._nb = dict(shape=tuple(shape), plan=None)
        else:
            self._nb = dict(shape=tuple(shape), plan=plan, hw=h * w, c=c,
                            head_dx=torch.empty((N, h, w, c), dtype=torch.float32, device=device),
                            head_ws=self._glue_ws(True, _abi.GLUE_HEAD, N, 1, 1, c, spec.num_classes, 1, 1, device),
                            loss=torch.zeros(1, dtype=torch.float32, device=device),
                            # max|dY| of every fp16 chain's backward input, left there by the kernel that produces dY
                            # (head / transition data gradient); zeroed once per step
                            amax=torch.zeros(max(4, len(plan)), dtype=torch.float32, device=device))
        self._nb_cache[tuple(shape)] = self._nb
        return self._nb

    def segments_have_chain(self):
        return any(seg[0] == "chain" for seg in self.segments)

    def _off(self, name):
        return self.torch_params[name][0]

    @staticmethod
    def _glue_ws(ok, op, N, H, W, ci, co, sh, sw, device):
        """Caller-owned workspace of a stem / transition / head gradient call (b200ode_glue_workspace_bytes)."""
        if not ok:
            return None
        n = ctypes.c_size_t()
        _abi.check(_abi.lib().b200ode_glue_workspace_bytes(op, N, H, W, ci, co, sh, sw, ctypes.byref(n)))
        return _alloc_workspace(n.value, device)

    def _fwd_bwd_native(self, images, onehot, nb):
        lib, st, spec = _abi.lib(), _stream_ptr(), self.spec
        N, H, W, _ = images.shape
        is_u8 = images.dtype == torch.uint8
        images = images.contiguous() if is_u8 else images.contiguous().float()
        onehot = onehot.contiguous().float()
        norm = spec.subtract_mean is not None or spec.divide_by_stddev is not None
        sub = float(spec.subtract_mean or 0.0)
        div = float(spec.divide_by_stddev if spec.divide_by_stddev is not None else 1.0)
        th, gr = self.theta, self.grad
        cur = images
        # The chains' weight staging (K1, one small launch per stage) only depends on the parameters: it runs on a side
        # stream under the stem / first stages instead of in front of every chain launch (fork / join, graph capturable).
        if getattr(self, "_pack_stream", None) is None:
            self._pack_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream()
        overlap_wgrad = not os.environ.get("B200ODE_NO_WGRAD_OVERLAP")
        overlap_pack = not os.environ.get("B200ODE_NO_PACK_OVERLAP")
        if overlap_pack:
            self._pack_stream.wait_stream(main)
        with torch.cuda.stream(self._pack_stream if overlap_pack else main):
            for e in nb["plan"]:
                if e["kind"] == "chain":
                    e["chain"].fused.pack(self.theta_euler[e["chain"].offset:], e["chain"].np_layer)
            packed = torch.cuda.Event()
            packed.record(self._pack_stream if overlap_pack else main)
        joined = False
        for e in nb["plan"]:
            if e["kind"] == "stem":
                ko = self._off(e["name"] + "/kernel")
                _abi.check(lib.b200ode_stem_fwd(_ptr(cur), int(is_u8), sub, div, int(norm), _ptr(th[ko:]),
                                                _ptr(th[self._off(e["name"] + "/bias"):]), _ptr(e["out"]), N, e["h"], e["w"],
                                                e["ci"], e["co"], st))
                cur = e["out"]
            elif e["kind"] == "transition":
                nm = e["name"]
                e["x"] = cur
                _abi.check((lib.b200ode_transition_fwd_fast if self.glue_fast else lib.b200ode_transition_fwd)(_ptr(cur), _ptr(th[self._off(nm + "2/kernel"):]),
                                                      _ptr(th[self._off(nm + "2/bias"):]), _ptr(th[self._off(nm + "1/kernel"):]),
                                                      _ptr(th[self._off(nm + "1/bias"):]), _ptr(e["out"]), _ptr(e["mask"]),
                                                      N, e["h"], e["w"], e["ci"], e["co"], e["st"][0], e["st"][1], st))
                cur = e["out"]
            else:
                ch = e["chain"]
                if not joined:
                    main.wait_event(packed)
                    joined = True
                cur = ch.fused_forward(cur, spec.h)
        if not joined:
            main.wait_event(packed)
        fo = self._off("fc/kernel")
        # fp16 chains scale their backward strips by max|dY|: the kernels that PRODUCE dY (head, tensor-core transition data
        # gradients) leave it in nb["amax"][i] (i = index of the chain in the plan) instead of three extra reductions + memsets
        plan_ = nb["plan"]
        fuse_amax = os.environ.get("B200ODE_FUSED_AMAX", "1") != "0" and all(
            e["kind"] != "chain" or (e["chain"].fused is not None and e["chain"].fused.f16) for e in plan_)
        amax = nb["amax"]
        if fuse_amax:
            amax.zero_()
        if fuse_amax and plan_[-1]["kind"] == "chain":
            _abi.check(lib.b200ode_head_fwd_bwd_amax(_ptr(cur), _ptr(th[fo:]), _ptr(th[self._off("fc/bias"):]), _ptr(onehot), 1e-7, None,
                                                     _ptr(nb["loss"]), _ptr(nb["head_dx"]), _ptr(gr[fo:]), N, nb["hw"], nb["c"],
                                                     spec.num_classes, _ptr(nb["head_ws"]), nb["head_ws"].numel(),
                                                     _ptr(amax[len(plan_) - 1:]), st))
            d_amax = amax[len(plan_) - 1:len(plan_)]
        else:
            _abi.check(lib.b200ode_head_fwd_bwd(_ptr(cur), _ptr(th[fo:]), _ptr(th[self._off("fc/bias"):]), _ptr(onehot), 1e-7, None,
                                                _ptr(nb["loss"]), _ptr(nb["head_dx"]), _ptr(gr[fo:]), N, nb["hw"], nb["c"],
                                                spec.num_classes, _ptr(nb["head_ws"]), nb["head_ws"].numel(), st))
            d_amax = None
        # Backward.  Critical path (main stream): head -> chain dgrad -> transition dgrad -> chain dgrad -> ...; every
        # weight gradient (chain wgrad + fold, transition wgrad, stem wgrad: ~1/4 of the step) only feeds the optimiser,
        # so it runs on a side stream, ordered by events after the data gradient that produces its dZ, and is joined
        # before Adam.  The fork / join is captured in the CUDA graph like everything else.
        d = nb["head_dx"]
        side = self._side_stream if overlap_wgrad else main
        # Opt-in (B200ODE_EARLY_ADAM / B200ODE_AR_OVERLAP): exchanging slices under the backward pass was measured SLOWER than
        # one exchange at the end (N=2: 1.239 vs 1.198 ms/step) -- NCCL's CTAs cannot co-reside with the one-CTA-per-SM chain
        # kernels (registers), so a chain launch that meets them runs a second wave.  Default: peer-memory exchange inside the
        # Adam kernel (p2p), else ONE all-reduce of the whole bucket before Adam.
        early = self._early_slices(nb) if overlap_wgrad and self.world_size > 1 and not self.p2p and os.environ.get("B200ODE_EARLY_ADAM") else None
        for ei in range(len(plan_) - 1, -1, -1):
            e = plan_[ei]
            if e["kind"] == "chain":
                ch = e["chain"]
                dnext = ch.fused_dgrad(d, spec.h, dy_amax=d_amax)
                d_amax = None
                if overlap_wgrad:
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                with torch.cuda.stream(side):
                    ch.fused_wgrad(self.grad_euler)
                    if self.world_size > 1 and not self.p2p and os.environ.get("B200ODE_AR_OVERLAP"):
                        # the stage's packed gradients are final: start their all-reduce now (NCCL stream); joined in
                        # _optimizer before Adam (opt-in, see above)
                        lo, hi = ch.offset, ch.offset + ch.n * ch.np_layer
                        self._pending.append(self._ar_async(self.grad_euler[lo:hi]))
                        self._reduced_upto = min(self._reduced_upto, lo)
                d = dnext
            elif e["kind"] == "transition":
                nm = e["name"]
                with torch.cuda.stream(side):      # needs d (ready: the side stream already waits for the chain's dgrad)
                    _abi.check((lib.b200ode_transition_wgrad_fast if self.glue_fast else lib.b200ode_transition_wgrad)(_ptr(e["x"]), _ptr(d), _ptr(e["mask"]), _ptr(gr[self._off(nm + "2/kernel"):]),
                                                            N, e["h"], e["w"], e["ci"], e["co"], e["st"][0], e["st"][1],
                                                            _ptr(e["ws"]), e["ws"].numel(), side.cuda_stream))
                if self.glue_fast:
                    want_amax = fuse_amax and ei > 0 and plan_[ei - 1]["kind"] == "chain"
                    _abi.check(lib.b200ode_transition_dgrad_fast(_ptr(d), _ptr(e["mask"]), _ptr(th[self._off(nm + "2/kernel"):]),
                                                                 _ptr(th[self._off(nm + "1/kernel"):]), _ptr(e["dx"]), N, e["h"], e["w"],
                                                                 e["ci"], e["co"], e["st"][0], e["st"][1],
                                                                 _ptr(amax[ei - 1:]) if want_amax else None, st))
                    d_amax = amax[ei - 1:ei] if want_amax else None
                elif fuse_amax and ei > 0 and plan_[ei - 1]["kind"] == "chain":
                    _abi.check(lib.b200ode_transition_dgrad_amax(_ptr(d), _ptr(e["mask"]), _ptr(th[self._off(nm + "2/kernel"):]),
                                                                 _ptr(th[self._off(nm + "1/kernel"):]), _ptr(e["dx"]), N, e["h"], e["w"],
                                                                 e["ci"], e["co"], e["st"][0], e["st"][1], _ptr(amax[ei - 1:]), st))
                    d_amax = amax[ei - 1:ei]
                else:
                    _abi.check(lib.b200ode_transition_dgrad(_ptr(d), _ptr(e["mask"]), _ptr(th[self._off(nm + "2/kernel"):]),
                                                            _ptr(th[self._off(nm + "1/kernel"):]), _ptr(e["dx"]), N, e["h"], e["w"],
                                                            e["ci"], e["co"], e["st"][0], e["st"][1], st))
                    d_amax = None
                d = e["dx"]
                if early is not None and e is early["after"]:
                    # Everything behind the first stage is final now (gradients written, parameters no longer read by
                    # this step): exchange and update those slices on the side stream while the first stage's backward
                    # sweep and weight gradient run; only the first chain + stem slices are left for _optimizer.
                    ev = torch.cuda.Event()
                    ev.record(main)
                    side.wait_event(ev)
                    with torch.cuda.stream(side):
                        if self.world_size > 1:
                            lo, hi = early["slices"][-1]
                            self._pending.append(self._ar_async(self.grad[lo:hi]))
                            for w in self._pending:
                                w.wait()
                            self._pending, self._reduced_upto = [], self.n_euler_params
                        for lo, hi in early["slices"]:
                            self._adam(lo, hi, side.cuda_stream)
                    self._adam_done = early["slices"]
            else:
                # the stem's weight gradient needs only d (made on the main stream, which has nothing else left to do): on the
                # main stream it runs BESIDE the first stage's weight gradient (side stream) instead of behind it
                _abi.check(lib.b200ode_stem_wgrad(_ptr(images), int(is_u8), sub, div, int(norm), _ptr(e["out"]), _ptr(d),
                                                  _ptr(gr[self._off(e["name"] + "/kernel"):]), N, e["h"], e["w"], e["ci"],
                                                  e["co"], _ptr(e["ws"]), e["ws"].numel(), st))
        if overlap_wgrad:
            main.wait_stream(side)
        return nb["loss"].view(())

    def predict(self, images):
        """Inference (softmax probabilities [N, classes]) through the native kernels: stem, one persistent
        chain launch per stage that keeps every intermediate activation in shared memory (nothing but the
        stage output is written), transitions, head.  Falls back to `forward` for shapes / modes the native
        path does not take."""
        nb = self._native_plan(images.shape, images.device)
        if nb["plan"] is None:
            was = self.training
            self.training = False                  # BatchNorm: moving statistics
            try:   # strict mode: the torch-op layers (stem / transitions / head) must not use TF32 either
                with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=self.precision != "strict"):
                    return self.forward(images)
            finally:
                self.training = was
        lib, st, spec = _abi.lib(), _stream_ptr(), self.spec
        N = images.shape[0]
        is_u8 = images.dtype == torch.uint8
        images = images.contiguous() if is_u8 else images.contiguous().float()
        norm = spec.subtract_mean is not None or spec.divide_by_stddev is not None
        sub = float(spec.subtract_mean or 0.0)
        div = float(spec.divide_by_stddev if spec.divide_by_stddev is not None else 1.0)
        th = self.theta
        cur = images
        for e in nb["plan"]:
            if e["kind"] == "stem":
                _abi.check(lib.b200ode_stem_fwd(_ptr(cur), int(is_u8), sub, div, int(norm), _ptr(th[self._off(e["name"] + "/kernel"):]),
                                                _ptr(th[self._off(e["name"] + "/bias"):]), _ptr(e["out"]), N, e["h"], e["w"],
                                                e["ci"], e["co"], st))
                cur = e["out"]
            elif e["kind"] == "transition":
                nm = e["name"]
                _abi.check((lib.b200ode_transition_fwd_fast if self.glue_fast else lib.b200ode_transition_fwd)(_ptr(cur), _ptr(th[self._off(nm + "2/kernel"):]),
                                                      _ptr(th[self._off(nm + "2/bias"):]), _ptr(th[self._off(nm + "1/kernel"):]),
                                                      _ptr(th[self._off(nm + "1/bias"):]), _ptr(e["out"]), _ptr(e["mask"]),
                                                      N, e["h"], e["w"], e["ci"], e["co"], e["st"][0], e["st"][1], st))
                cur = e["out"]
            else:
                ch = e["chain"]
                ch.fused.pack(self.theta_euler[ch.offset:], ch.np_layer)
                ch.fused.forward(cur, spec.h, y_final=ch.f_dx)      # f_dx doubles as the stage output buffer
                cur = ch.f_dx
        probs = torch.empty((N, spec.num_classes), dtype=torch.float32, device=images.device)
        zeros = torch.zeros((N, spec.num_classes), dtype=torch.float32, device=images.device)
        fo = self._off("fc/kernel")
        _abi.check(lib.b200ode_head_fwd_bwd(_ptr(cur), _ptr(th[fo:]), _ptr(th[self._off("fc/bias"):]), _ptr(zeros), 1e-7,
                                            _ptr(probs), _ptr(nb["loss"]), None, None, N, nb["hw"], nb["c"],
                                            spec.num_classes, _ptr(nb["head_ws"]), nb["head_ws"].numel(), st))
        return probs

    def _fwd_bwd(self, images, onehot):
        nb = self._native_plan(images.shape, images.device)
        if nb["plan"] is not None:
            return self._fwd_bwd_native(images, onehot, nb)
        # strict mode: the few torch-op layers (stem / transitions / head) must not use TF32 either
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=self.precision != "strict"):
            return self._fwd_bwd_inner(images, onehot)

    def _fwd_bwd_inner(self, images, onehot):
        loss = self.loss_fn(self.forward(images), onehot)
        # torch.autograd.grad (no AccumulateGrad nodes -> CUDA-graph capturable); the Euler chains write
        # their parameter gradients into the flat bucket as a side effect of their backward node.
        grads = torch.autograd.grad(loss, self._leaf_list)
        torch._foreach_copy_(self._leaf_grad_views, list(grads))
        return loss

    def _ar_async(self, t):
        return self.comm.allreduce_async(t) if self.comm is not None else allreduce_async(t)

    def _ar(self, t):
        return self.comm.allreduce_bucket(t) if self.comm is not None else allreduce_bucket(t, self.world_size)

    def _early_slices(self, nb):
        """Slices of the flat bucket that are final once the FIRST transition's data gradient is queued: the chains behind
        the first one and the glue parameters behind the stem ([first transition .. fc]).  None for nets without that shape."""
        plan = nb["plan"]
        if len(plan) < 3 or plan[0]["kind"] != "stem" or plan[1]["kind"] != "chain" or plan[2]["kind"] != "transition":
            return None
        first = plan[1]["chain"]
        lo_e = first.offset + first.n * first.np_layer
        name = plan[0]["name"]
        a, shape = self.torch_params[name + "/bias"]
        lo_g = a + math.prod(shape)
        if self.torch_params[name + "/kernel"][0] != self.n_euler_params or lo_e >= self.n_euler_params:
            return None
        return dict(after=plan[2], slices=[(lo_e, self.n_euler_params), (lo_g, self.n_params)])

    def _adam(self, lo, hi, stream):
        _abi.check(_abi.lib().b200ode_adam_step(_ptr(self.theta[lo:]), _ptr(self.grad[lo:]), _ptr(self.adam_m[lo:]), _ptr(self.adam_v[lo:]),
                                                hi - lo, _ptr(self.step_counter), self.lr, 0.9, 0.999, self.adam_eps,
                                                1.0 / self.world_size, stream))

    def _optimizer(self):
        """All-reduce (data parallel) and Adam for every slice of the flat bucket the step has not updated yet
        (`_adam_done`: slices exchanged and updated early, under the tail of the backward pass)."""
        if self.p2p:
            self.comm.adam_step(self.theta, self.grad, self.adam_m, self.adam_v, self.step_counter, self.lr, self.adam_eps)
            _abi.check(_abi.lib().b200ode_increment(_ptr(self.step_counter), _stream_ptr()))
            return
        done = sorted(getattr(self, "_adam_done", None) or [])
        self._adam_done = []
        rest, cur = [], 0
        for lo, hi in done:
            if lo > cur:
                rest.append((cur, lo))
            cur = max(cur, hi)
        if cur < self.n_params:
            rest.append((cur, self.n_params))
        if self.world_size > 1:
            if self._pending or done:
                # chains were reduced stage by stage as their weight gradients were queued: [reduced_upto, n_euler) is in
                # flight or done; reduce what is left of the remaining slices, then join the overlapped collectives
                for lo, hi in rest:
                    if lo < self.n_euler_params:
                        hi_e = min(hi, self._reduced_upto)
                        if hi_e > lo:
                            self._pending.append(self._ar_async(self.grad[lo:hi_e]))
                        lo = max(lo, self.n_euler_params)
                    if hi > lo and hi > self.n_euler_params:
                        self._pending.append(self._ar_async(self.grad[lo:hi]))
                for w in self._pending:
                    w.wait()
                self._pending, self._reduced_upto = [], self.n_euler_params
            else:
                self._ar(self.grad)
        st = _stream_ptr()
        for lo, hi in rest:
            self._adam(lo, hi, st)
        _abi.check(_abi.lib().b200ode_increment(_ptr(self.step_counter), st))

    def train_step(self, images, onehot):
        """Eager step; returns the loss tensor (device)."""
        loss = self._fwd_bwd(images, onehot)
        self._optimizer()
        return loss.detach()

    # ----------------------------------------------------------------------------------------------
    def capture(self, images, onehot, warmup=3):
        """Capture the whole train step in a CUDA graph (inputs are copied into static buffers).  The warm-up
        steps run on the static batch but leave no trace: parameters, Adam moments and the step counter are
        snapshotted before and restored after them (capture itself executes nothing)."""
        self._static_in = (images.clone(), onehot.clone())
        snap = (self.theta.clone(), self.adam_m.clone(), self.adam_v.clone(), self.step_counter.clone())
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.train_step(*self._static_in)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for dst, src in zip((self.theta, self.adam_m, self.adam_v, self.step_counter), snap):
            dst.copy_(src)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self.train_step(*self._static_in)
        return self

    def release(self):
        """Drop the captured CUDA graph and its static inputs (before a communicator whose collectives it captured is
        destroyed: NCCL teardown blocks while such a graph is alive)."""
        self._graph = None
        self._static_in = None
        self._static_loss = None
        self._stage = None

    # ---- input pipeline of the captured step: the next batch travels host -> device while the current step runs ----------
    def prefetch(self, images_host, onehot_host):
        """Start the host -> device copy of the NEXT step's batch (pinned host tensors) on a copy stream, into staging
        buffers; `train_step_graph_prefetched` moves them into the graph's static inputs (device -> device, ~2 us) right
        before its replay.  The copy overlaps the step that is running; the static inputs are never written while a replay
        may still read them (the stem's weight gradient reads the images at the END of the step)."""
        if getattr(self, "_stage", None) is None:
            self._stage = (torch.empty_like(self._static_in[0]), torch.empty_like(self._static_in[1]))
            self._copy_stream = torch.cuda.Stream()
            self._copied, self._stage_free = torch.cuda.Event(), None
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free is not None:
                self._copy_stream.wait_event(self._stage_free)      # the previous batch has left the staging buffers
            self._stage[0].copy_(images_host, non_blocking=True)
            self._stage[1].copy_(onehot_host, non_blocking=True)
            self._copied.record(self._copy_stream)

    def train_step_graph_prefetched(self):
        """Replay the captured step on the batch delivered by the last `prefetch` call."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copied)
        self._static_in[0].copy_(self._stage[0], non_blocking=True)
        self._static_in[1].copy_(self._stage[1], non_blocking=True)
        self._stage_free = torch.cuda.Event()
        self._stage_free.record(cur)
        self._graph.replay()
        return self._static_loss

    def train_step_graph(self, images=None, onehot=None):
        if images is not None:
            self._static_in[0].copy_(images, non_blocking=True)
            self._static_in[1].copy_(onehot, non_blocking=True)
        self._graph.replay()
        return self._static_loss

    # ----------------------------------------------------------------------------------------------
    def bn_param_slices(self):
        """(bn layer name, offset of gamma, C, chain, index) of every Euler-step BatchNorm (`bn{s}_{b}_branch2`,
        models/tfkeras_resnets.py:66, 85-87); beta follows gamma."""
        out, idx = [], 0
        plan = [p for p in self.spec.plan() if p[0] == "euler"]
        for seg in self.segments:
            if seg[0] != "chain":
                continue
            ch = seg[1]
            for l in range(ch.n):
                if ch.bn:
                    out.append((plan[idx][4].replace("res", "bn"), ch.bn_offset + 2 * ch.C * l, ch.C, ch, l))
                idx += 1
        return out

    def export_params(self):
        """dict name -> CPU tensor: '<layer>/packed' for Euler layers (reference variable order,
        flattened), '<layer>/kernel' (HWIO) and '<layer>/bias' for regular layers, '<bn>/gamma|beta|moving_mean|
        moving_variance' for BatchNorm layers (the Keras variable names)."""
        out = {}
        th = self.theta.detach().cpu()
        for name, off, n, C in self.layer_param_slices():
            out[name + "/packed"] = th[off:off + n].clone()
        for name, (a, shape) in self.torch_params.items():
            out[name] = th[a:a + math.prod(shape)].view(shape).clone()
        for name, off, C, ch, l in self.bn_param_slices():
            out[name + "/gamma"], out[name + "/beta"] = th[off:off + C].clone(), th[off + C:off + 2 * C].clone()
            out[name + "/moving_mean"], out[name + "/moving_variance"] = ch.bn_moving[l, 0].cpu().clone(), ch.bn_moving[l, 1].cpu().clone()
        for name, (mm, mv) in self.bn_moving.items():
            out[name + "/moving_mean"], out[name + "/moving_variance"] = mm.cpu().clone(), mv.cpu().clone()
        return out

    def import_params(self, params):
        with torch.no_grad():
            for name, off, n, C in self.layer_param_slices():
                self.theta[off:off + n].copy_(params[name + "/packed"].reshape(-1))
            for name, (a, shape) in self.torch_params.items():
                self.theta[a:a + math.prod(shape)].copy_(params[name].reshape(-1))
            for name, off, C, ch, l in self.bn_param_slices():
                self.theta[off:off + C].copy_(params[name + "/gamma"])
                self.theta[off + C:off + 2 * C].copy_(params[name + "/beta"])
                if name + "/moving_mean" in params:
                    ch.bn_moving[l, 0].copy_(params[name + "/moving_mean"])
                    ch.bn_moving[l, 1].copy_(params[name + "/moving_variance"])
            for name, (mm, mv) in self.bn_moving.items():
                if name + "/moving_mean" in params:
                    mm.copy_(params[name + "/moving_mean"])
                    mv.copy_(params[name + "/moving_variance"])

    def save_weights(self, filepath):
        """`model.save_weights(path + '.h5')` (reference notebooks v6 / v7): Keras HDF5 weights file, `keras_h5.py`."""
        from .keras_h5 import save_weights_h5
        save_weights_h5(self, filepath)

    def load_weights(self, filepath):
        """`model.load_weights(path)`: parameters only; optimizer state is untouched (Keras weights files hold none)."""
        from .keras_h5 import load_weights_h5
        load_weights_h5(self, filepath)

    def export_grads(self):
        out = {}
        g = self.grad.detach().cpu()
        for name, off, n, C in self.layer_param_slices():
            out[name + "/packed"] = g[off:off + n].clone()
        for name, (a, shape) in self.torch_params.items():
            out[name] = g[a:a + math.prod(shape)].view(shape).clone()
        for name, off, C, ch, l in self.bn_param_slices():
            out[name + "/gamma"], out[name + "/beta"] = g[off:off + C].clone(), g[off + C:off + 2 * C].clone()
        return out

    def layer_param_slices(self):
        """(name, offset, size, C) of every Euler layer in the flat bucket, graph order."""
        out, idx = [], 0
        plan = [p for p in self.spec.plan() if p[0] == "euler"]
        for seg in self.segments:
            if seg[0] != "chain":
                continue
            ch = seg[1]
            for l in range(ch.n):
                out.append((plan[idx][4], ch.offset + l * ch.np_layer, ch.np_layer, ch.C))
                idx += 1
        return out

    def gradient_metric_names(self):
        """Column names of the reference's gradient history (`training/training.py:385-407`; header of
        numerical_results/csv/*_gradient_history.csv): conv1 first, then every Euler layer in graph order."""
        return ["conv1_kernel_gradient_mean_norm"] + [name + "_kernel_gradient_mean_norm" for name, _, _, _ in self.layer_param_slices()]

    def gradient_mean_norms(self):
        """Per-layer ||g||_2 / size over the kernel variables (training/training.py:385-407): the stem kernel and, per
        antisymmetric layer, its C+3 kernel variables a,b,c,d,W_o merged (bias excluded) -- ONE launch over the flat
        gradient bucket (`b200ode_gradient_mean_norms`).  Returns an OrderedDict name -> float."""
        from collections import OrderedDict
        if getattr(self, "_gm_tables", None) is None:
            a, shape = self.torch_params["conv1/kernel"]
            offs, sizes = [a], [math.prod(shape)]
            for _, off, n, C in self.layer_param_slices():
                offs.append(off); sizes.append(n - C)
            self._gm_tables = (torch.tensor(offs, dtype=torch.int64, device=self.device),
                               torch.tensor(sizes, dtype=torch.int64, device=self.device),
                               torch.empty(len(offs), dtype=torch.float32, device=self.device))
        offs, sizes, out = self._gm_tables
        _abi.check(_abi.lib().b200ode_gradient_mean_norms(_ptr(self.grad), _ptr(offs), _ptr(sizes), offs.numel(),
                                                          1.0 / self.world_size, _ptr(out), _stream_ptr()))
        return OrderedDict(zip(self.gradient_metric_names(), out.cpu().tolist()))

    def history_row(self, loss, probs=None, onehot=None):
        """One row of the reference's gradient-history CSV: `global_step mean_loss accuracy <layer>_kernel_gradient_mean_norm ...`
        (space separated; `header=True` rows via gradient_history_header())."""
        acc = float("nan")
        if probs is not None and onehot is not None:
            acc = float((probs.argmax(dim=1) == onehot.argmax(dim=1)).float().mean())
        step = int(self.step_counter) - 1
        return " ".join([str(step), repr(float(loss)), repr(acc)] + [repr(v) for v in self.gradient_mean_norms().values()])

    def gradient_history_header(self):
        return " ".join(["global_step", "mean_loss", "accuracy"] + self.gradient_metric_names())
