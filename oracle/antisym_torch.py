"""O1 oracle: torch-CPU float32 restatement (oneDNN conv) of the reference path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is the stand-in for
the reference's TF-1.12 CPU path that bench.py times as `cpu_baseline` /
`--impl reference` (TensorFlow cannot be installed here).
Paths cited are relative to /root/reference.

Two kernel-assembly variants are provided:
  * `assemble_literal`  -- the reference's O(C^2) slice/negate/concat loops
    (layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:113-141), re-executed every
    step exactly like the reference graph does;
  * `assemble_closed`   -- a single gather with a precomputed index/sign table.
Both are differentiable so torch autograd supplies the reference backward
(training/training.py:300 `optimizer.compute_gradients`).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import antisym_numpy as O0

# ---------------------------------------------------------------------------
# Assembly
# ---------------------------------------------------------------------------


def assemble_literal(variables, C: int, gamma: float) -> torch.Tensor:
    """Reference build() loops on torch tensors (…3By3.py:113-141, 210-293)."""
    a, b, c, d = variables[:4]
    W = variables[4:4 + C - 1]
    e = torch.full((1, 1, 1, C), float(gamma), dtype=a.dtype)
    row1 = torch.cat([a, b, c], dim=1)
    row2 = torch.cat([d, e, -d], dim=1)
    row3 = torch.cat([-c, -b, -a], dim=1)
    diag = torch.cat([row1, row2, row3], dim=0)  # [3,3,1,C]
    outs, transposes = [], []
    for o in range(C):
        n = C - o - 1
        if n > 0:
            w = W[o]
            t = torch.stack([
                torch.stack([-w[2, 2, :], -w[2, 1, :], -w[2, 0, :]], dim=0),
                torch.stack([-w[1, 2, :], -w[1, 1, :], -w[1, 0, :]], dim=0),
                torch.stack([-w[0, 2, :], -w[0, 1, :], -w[0, 0, :]], dim=0)], dim=0)
            single = torch.cat([diag[:, :, :, o], w], dim=-1)
        else:
            single = diag[:, :, :, o]
        for i in range(o):
            ct = transposes[-(i + 1)][:, :, i].unsqueeze(-1)
            single = torch.cat([ct, single], dim=-1)
        outs.append(single)
        if n > 0:
            transposes.append(t)
    return torch.stack(outs, dim=-1)


_TABLE_CACHE = {}


def closed_form_table(C: int):
    """index/sign/gamma-mask tables so that K = sign * flat[index] + gamma * centre."""
    if C in _TABLE_CACHE:
        return _TABLE_CACHE[C]
    off = O0.offsets_3by3(C)
    idx = np.zeros((3, 3, C, C), dtype=np.int64)
    sgn = np.zeros((3, 3, C, C), dtype=np.float32)
    cen = np.zeros((3, 3, C, C), dtype=np.float32)
    diag_src = {(0, 0): ("a", 1), (0, 1): ("b", 1), (0, 2): ("c", 1), (1, 0): ("d", 1),
                (1, 2): ("d", -1), (2, 0): ("c", -1), (2, 1): ("b", -1), (2, 2): ("a", -1)}
    for o in range(C):
        for (al, be), (name, s) in diag_src.items():
            idx[al, be, o, o] = off[name] + o
            sgn[al, be, o, o] = s
        cen[1, 1, o, o] = 1.0
        n = C - o - 1
        for j in range(n):
            ci = o + 1 + j
            for al in range(3):
                for be in range(3):
                    p = off["W"][o] + (al * 3 + be) * n + j
                    idx[al, be, ci, o] = p
                    sgn[al, be, ci, o] = 1.0
                    idx[2 - al, 2 - be, o, ci] = p
                    sgn[2 - al, 2 - be, o, ci] = -1.0
    tabs = (torch.from_numpy(idx), torch.from_numpy(sgn), torch.from_numpy(cen))
    _TABLE_CACHE[C] = tabs
    return tabs


def assemble_closed(flat: torch.Tensor, C: int, gamma: float) -> torch.Tensor:
    idx, sgn, cen = closed_form_table(C)
    return flat[idx.reshape(-1)].reshape(3, 3, C, C) * sgn.to(flat.dtype) + float(gamma) * cen.to(flat.dtype)


def split_params(flat: torch.Tensor, C: int):
    off = O0.offsets_3by3(C)
    out = [flat[off[k]:off[k] + C].reshape(1, 1, 1, C) for k in "abcd"]
    for o in range(C - 1):
        n = C - o - 1
        out.append(flat[off["W"][o]:off["W"][o] + 9 * n].reshape(3, 3, n))
    out.append(flat[off["bias"]:off["bias"] + C])
    return out


# ---------------------------------------------------------------------------
# Ops (NHWC at the boundary; torch conv runs NCHW internally via channels_last)
# ---------------------------------------------------------------------------


def conv2d_same_nhwc(x: torch.Tensor, K: torch.Tensor, strides=(1, 1)) -> torch.Tensor:
    """tf.nn.conv2d(NHWC, HWIO, SAME) semantics incl. TF's asymmetric padding."""
    kh, kw = K.shape[0], K.shape[1]
    _, pt, pb = O0.same_padding(x.shape[1], kh, strides[0])
    _, pl, pr = O0.same_padding(x.shape[2], kw, strides[1])
    xn = x.permute(0, 3, 1, 2)
    if pt != pb or pl != pr:
        xn = F.pad(xn, (pl, pr, pt, pb))
        pad = (0, 0)
    else:
        pad = (pt, pl)
    y = F.conv2d(xn, K.permute(3, 2, 0, 1), stride=strides, padding=pad)
    return y.permute(0, 2, 3, 1)


def batch_norm_train(z, bn_gamma, bn_beta, eps=O0.BN_EPS):
    mu = z.mean(dim=(0, 1, 2))
    var = z.var(dim=(0, 1, 2), unbiased=False)
    return bn_gamma * (z - mu) / torch.sqrt(var + eps) + bn_beta


def euler_step(x, K, bias, h=1.0, bn=None, forced_mask=None):
    """models/tfkeras_resnets.py:69-92.  `forced_mask` (bool, optional; parity tests only) replaces relu's own
    branch decision: relu(z) -> z*mask, so that a backward pass can be compared element by element with an
    implementation whose pre-activations differ from this one at rounding level (relu' is discontinuous at 0)."""
    z = conv2d_same_nhwc(x, K) + bias
    if bn is not None:
        z = batch_norm_train(z, bn[0], bn[1])
    r = torch.relu(z) if forced_mask is None else z * forced_mask.to(z.dtype)
    if h != 1.0:
        r = h * r
    return r + x


# ---------------------------------------------------------------------------
# Single-block antisymmetric ResNet (models/tfkeras_resnets.py:511-604)
# ---------------------------------------------------------------------------


class NetSpec:
    """Configuration mirror of get_single_block_resnet_build_function kwargs."""

    def __init__(self, num_stages=4, blocks_per_stage=(3, 3, 3), filters_per_block=(16, 32, 64),
                 strides=((1, 1), (2, 2), (2, 2)), h=1.0, gamma=0.0, num_classes=10,
                 use_batch_norm=False, subtract_mean=127.5, divide_by_stddev=127.5, kernel_size=3,
                 in_channels=3, use_max_pooling=None):
        self.use_max_pooling = list(use_max_pooling) if use_max_pooling is not None else [False] * (num_stages - 1)
        self.num_stages = num_stages
        self.blocks_per_stage = list(blocks_per_stage)
        self.filters_per_block = list(filters_per_block)
        self.strides = [tuple(s) for s in strides]
        self.h, self.gamma = h, gamma
        self.num_classes = num_classes
        self.use_batch_norm = use_batch_norm
        self.subtract_mean, self.divide_by_stddev = subtract_mean, divide_by_stddev
        self.kernel_size = kernel_size
        self.in_channels = in_channels

    def plan(self):
        """List of ('stem'|'euler'|'transition'|'maxpool', C_in, C_out, stride, name) in graph order
        (stage loop models/tfkeras_resnets.py:575-593; MaxPooling2D(2,2) in front of a stage when use_max_pooling[s], :577-578,
        after which the stage starts with a conv block, :589-593)."""
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


def he_normal_keras(gen, shape, fan_in):
    """Keras he_normal = VarianceScaling(2, fan_in, truncated normal): TF1.12 divides
    stddev by .87962566 to correct the truncation."""
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    t = torch.empty(shape)
    torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=gen)
    return t


def init_net_params(spec: NetSpec, seed=0):
    """dict name -> fp32 tensor.  Antisymmetric layers hold one flat packed vector
    in reference variable order; regular Keras layers hold kernel(HWIO)+bias."""
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.default_rng(seed)
    P = {}
    k = spec.kernel_size
    for kind, ci, co, st, name in spec.plan():
        if kind == "stem":
            P[name + "/kernel"] = he_normal_keras(gen, (k, k, ci, co), k * k * ci)
            P[name + "/bias"] = torch.zeros(co)
            if spec.use_batch_norm:
                P["bn_conv1/gamma"], P["bn_conv1/beta"] = torch.ones(co), torch.zeros(co)
        elif kind == "euler":
            P[name + "/packed"] = torch.from_numpy(O0.init_params_3by3(rng, co))
            if spec.use_batch_norm:
                bn = name.replace("res", "bn")
                P[bn + "/gamma"], P[bn + "/beta"] = torch.ones(co), torch.zeros(co)
        elif kind == "maxpool":
            pass
        else:
            P[name + "2/kernel"] = he_normal_keras(gen, (k, k, ci, co), k * k * ci)
            P[name + "2/bias"] = torch.zeros(co)
            P[name + "1/kernel"] = he_normal_keras(gen, (1, 1, ci, co), ci)
            P[name + "1/bias"] = torch.zeros(co)
            if spec.use_batch_norm:
                for br in ("2", "1"):
                    bn = name.replace("res", "bn") + br
                    P[bn + "/gamma"], P[bn + "/beta"] = torch.ones(co), torch.zeros(co)
    c_last = spec.filters_per_block[spec.num_stages - 2]
    P["fc/kernel"] = he_normal_keras(gen, (c_last, spec.num_classes), c_last)
    P["fc/bias"] = torch.zeros(spec.num_classes)
    return P


def net_forward(spec: NetSpec, P, images_u8_or_f32, assembly="closed", forced_masks=None, record_z=None):
    """images [N,H,W,3] (uint8 or float) -> softmax probabilities [N,num_classes].
    forced_masks: optional dict layer name -> bool mask for Euler layers (see euler_step);
    record_z: optional dict filled with the sign of every Euler layer's pre-activation (z > 0)."""
    x = images_u8_or_f32.to(torch.float32)
    if spec.subtract_mean is not None:
        x = x - spec.subtract_mean
    if spec.divide_by_stddev is not None:
        x = x / spec.divide_by_stddev
    for kind, ci, co, st, name in spec.plan():
        if kind == "stem":
            x = conv2d_same_nhwc(x, P[name + "/kernel"], st) + P[name + "/bias"]
            if spec.use_batch_norm:
                x = batch_norm_train(x, P["bn_conv1/gamma"], P["bn_conv1/beta"])
            x = torch.relu(x)
        elif kind == "maxpool":
            x = F.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)     # Keras MaxPooling2D((2,2)): 'valid', floor
        elif kind == "euler":
            flat = P[name + "/packed"]
            if assembly == "closed":
                K = assemble_closed(flat, co, spec.gamma)
            else:
                K = assemble_literal(split_params(flat, co), co, spec.gamma)
            bias = flat[-co:]
            bn = None
            if spec.use_batch_norm:
                b = name.replace("res", "bn")
                bn = (P[b + "/gamma"], P[b + "/beta"])
            if record_z is not None:
                with torch.no_grad():
                    zz = conv2d_same_nhwc(x, K) + bias
                    record_z[name] = (zz > 0, zz.abs())
            x = euler_step(x, K, bias, spec.h, bn, None if forced_masks is None else forced_masks.get(name))
        else:
            main = conv2d_same_nhwc(x, P[name + "2/kernel"], st) + P[name + "2/bias"]
            short = conv2d_same_nhwc(x, P[name + "1/kernel"], st) + P[name + "1/bias"]
            if spec.use_batch_norm:
                b = name.replace("res", "bn")
                main = batch_norm_train(main, P[b + "2/gamma"], P[b + "2/beta"])
                short = batch_norm_train(short, P[b + "1/gamma"], P[b + "1/beta"])
            x = torch.relu(main) + short      # models/tfkeras_resnets.py:266-267 (no h)
    x = x.mean(dim=(1, 2))                    # GlobalAveragePooling2D
    logits = x @ P["fc/kernel"] + P["fc/bias"]
    return torch.softmax(logits, dim=-1)



def bottleneck_resnet_forward(P, images, blocks_per_stage, filters_per_block, antisymmetric=True, use_batch_norm=True,
                              version=1, gamma=0.0, subtract_mean=None, divide_by_stddev=None, include_top=True):
    """Restatement of the reference's bottleneck ResNet graph (`models/tfkeras_resnets.py:761-813` = `_build_function` of
    `get_resnet_build_function`, blocks :96-202 and :271-425) on CPU tensors.  P: '<layer>/kernel' (HWIO), '<layer>/bias',
    '<layer>/packed' for antisymmetric layers, '<bn>/gamma|beta'; BatchNorm in training mode (batch statistics)."""
    def conv(x, name, st=(1, 1), valid=False):
        K = P[name + "/kernel"]
        if valid:
            y = F.conv2d(x.permute(0, 3, 1, 2), K.permute(3, 2, 0, 1), None, stride=st).permute(0, 2, 3, 1)
        else:
            y = conv2d_same_nhwc(x, K, st)
        return y + P[name + "/bias"]

    def bn(y, name):
        return batch_norm_train(y, P[name + "/gamma"], P[name + "/beta"]) if use_batch_norm else y

    x = images.to(P["conv1/kernel"].dtype)
    if subtract_mean is not None:
        x = x - subtract_mean
    if divide_by_stddev is not None:
        x = x / divide_by_stddev
    x = F.pad(x, (0, 0, 3, 3, 3, 3))                                            # ZeroPadding2D((3,3)) :775
    x = torch.relu(bn(conv(x, "conv1", (2, 2), valid=True), "bn_conv1"))         # :776-785
    x = F.pad(x, (0, 0, 1, 1, 1, 1))                                            # :786
    x = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1)            # :787
    for s in range(4):
        nf = filters_per_block[s]
        for b in range(blocks_per_stage[s]):
            base, bnb = "res%d_%d_branch" % (s + 2, b), "bn%d_%d_branch" % (s + 2, b)
            st = ((1, 1) if s == 0 else (2, 2)) if b == 0 else (1, 1)
            s1, sk = (st, (1, 1)) if version == 1 else ((1, 1), st)                # :341-348
            y = torch.relu(bn(conv(x, base + "2a", s1), bnb + "2a"))
            if antisymmetric and nf[1] is None:
                flat = P[base + "2b/packed"]
                C = y.shape[-1]
                y = conv2d_same_nhwc(y, assemble_closed(flat, C, gamma), sk) + flat[-C:]
            else:
                y = conv(y, base + "2b", sk)
            y = torch.relu(bn(y, bnb + "2b"))
            y = bn(conv(y, base + "2c"), bnb + "2c")
            short = bn(conv(x, base + "1", st), bnb + "1") if b == 0 else x
            x = torch.relu(y + short)
    if not include_top:
        return x
    x = x.mean(dim=(1, 2))
    return torch.softmax(x @ P["fc/kernel"] + P["fc/bias"], dim=-1)


def loss_fn(probs, onehot, eps=1e-7):
    """training/training.py:295: mean(K.categorical_crossentropy(from_logits=False))."""
    p = probs / probs.sum(dim=-1, keepdim=True)
    p = torch.clamp(p, eps, 1.0 - eps)
    return -(onehot * torch.log(p)).sum(dim=-1).mean()


def train_step(spec, P, M, V, t, images, onehot, lr=1e-3, assembly="closed", forced_masks=None, record_z=None):
    """One reference train step: fwd, loss, autodiff, tf.train.AdamOptimizer(eps=1e-7).
    P, M, V are dicts updated in place; returns (loss, grads)."""
    names = list(P.keys())
    leaves = [P[n].detach().requires_grad_(True) for n in names]
    Pl = dict(zip(names, leaves))
    loss = loss_fn(net_forward(spec, Pl, images, assembly, forced_masks, record_z), onehot)
    grads = torch.autograd.grad(loss, leaves)
    lr_t = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
    out_g = {}
    with torch.no_grad():
        for n, g in zip(names, grads):
            M[n].mul_(0.9).add_(g, alpha=0.1)
            V[n].mul_(0.999).addcmul_(g, g, value=0.001)
            P[n].sub_(lr_t * M[n] / (V[n].sqrt() + 1e-7))
            out_g[n] = g
    return float(loss.detach()), out_g
