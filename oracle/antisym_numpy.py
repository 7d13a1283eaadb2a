"""O0 oracle: NumPy restatement of the antisymmetric conv layer + Euler step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it restates; paths are relative to /root/reference.

Layer restated:
  layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:85-171, 210-293
  layers/tfkeras_layer_Conv2DAntisymmetric.py:90-175, 216-270
  layers/antisymmetric_conv2d_utils.py:23-75
  models/tfkeras_resnets.py:28-94 (Euler step), 204-269, 511-604
  training/training.py:283-304 (loss / optimizer step), 385-409 (grad metric)

All functions work in whatever dtype they are handed (float64 for ground
truth, float32 to mimic the reference's fp32 CPU arithmetic).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# Variable bookkeeping (reference variable order = flat "packed" order)
# ---------------------------------------------------------------------------


def num_params_3by3(C: int, use_bias: bool = True) -> int:
    """Free scalars of Conv2DAntisymmetric3By3: a,b,c,d [C] each, W_o [3,3,C-o-1],
    bias [C]  (tfkeras_layer_Conv2DAntisymmetric3By3.py:119-124, 148-153, 219-245)."""
    return 4 * C + 9 * C * (C - 1) // 2 + (C if use_bias else 0)


def offsets_3by3(C: int):
    """Flat offsets of the reference variables in creation order
    [a, b, c, d, W_0 .. W_{C-2}, bias] (training/training.py:397-398 pins the
    count: 4 + 15 + 1 = 20 variables at C=16)."""
    off = {"a": 0, "b": C, "c": 2 * C, "d": 3 * C}
    w = []
    cur = 4 * C
    for o in range(C - 1):
        w.append(cur)
        cur += 9 * (C - o - 1)
    off["W"] = w
    off["bias"] = cur
    return off


def split_params_3by3(flat: np.ndarray, C: int, use_bias: bool = True):
    """flat -> list of arrays with the reference variable shapes, in order."""
    off = offsets_3by3(C)
    out = [flat[off[k]:off[k] + C].reshape(1, 1, 1, C) for k in "abcd"]
    for o in range(C - 1):
        n = C - o - 1
        out.append(flat[off["W"][o]:off["W"][o] + 9 * n].reshape(3, 3, n))
    if use_bias:
        out.append(flat[off["bias"]:off["bias"] + C].reshape(C))
    return out


def join_params(variables) -> np.ndarray:
    return np.concatenate([np.asarray(v).reshape(-1) for v in variables])


# ---------------------------------------------------------------------------
# Kernel assembly, literal loops
# ---------------------------------------------------------------------------


def _anti_centrosymmetric_transpose(w: np.ndarray) -> np.ndarray:
    """tfkeras_layer_Conv2DAntisymmetric3By3.py:277-293: negate every tap and
    rotate the 3x3 spatial grid by 180 degrees.  w: [3,3,n]."""
    a, b, c = -w[0, 0, :], -w[0, 1, :], -w[0, 2, :]
    d, e, f = -w[1, 0, :], -w[1, 1, :], -w[1, 2, :]
    g, h, i = -w[2, 0, :], -w[2, 1, :], -w[2, 2, :]
    row1 = np.stack([i, h, g], axis=0)
    row2 = np.stack([f, e, d], axis=0)
    row3 = np.stack([c, b, a], axis=0)
    return np.stack([row1, row2, row3], axis=0)


def _anti_centrosymmetric_kernel(a, b, c, d, gamma, dtype):
    """tfkeras_layer_Conv2DAntisymmetric3By3.py:210-275.  a..d: [1,1,1,C]."""
    C = a.shape[-1]
    e = np.full((1, 1, 1, C), gamma, dtype=dtype)  # :248-250 tf.fill
    f, g, h, i = -d, -c, -b, -a  # :262-265
    row1 = np.concatenate([a, b, c], axis=1)
    row2 = np.concatenate([d, e, f], axis=1)
    row3 = np.concatenate([g, h, i], axis=1)
    return np.concatenate([row1, row2, row3], axis=0)  # [3,3,1,C]


def assemble_kernel_3by3_literal(variables, C: int, gamma: float) -> np.ndarray:
    """Literal restatement of build() (tfkeras_layer_Conv2DAntisymmetric3By3.py:
    113-141).  `variables` = [a,b,c,d,W_0..W_{C-2},(bias)] in reference shapes.
    Returns K [3,3,C_in,C_out]."""
    a, b, c, d = variables[:4]
    dtype = a.dtype
    W = variables[4:4 + C - 1]
    diag = _anti_centrosymmetric_kernel(a, b, c, d, gamma, dtype)
    single_output_kernels = []
    transposes = []
    for o in range(C):
        num_independent = C - o - 1
        if num_independent > 0:
            independent = W[o]
            t = _anti_centrosymmetric_transpose(independent)
            single = np.concatenate([diag[:, :, :, o], independent], axis=-1)
        else:
            single = diag[:, :, :, o]
        for i in range(o):
            ct = np.expand_dims(transposes[-(i + 1)][:, :, i], axis=-1)
            single = np.concatenate([ct, single], axis=-1)
        single_output_kernels.append(single)
        if num_independent > 0:
            transposes.append(t)
    return np.stack(single_output_kernels, axis=-1)


def assemble_kernel_3by3_closed(flat: np.ndarray, C: int, gamma: float) -> np.ndarray:
    """Closed form of the same kernel."""
    off = offsets_3by3(C)
    K = np.zeros((3, 3, C, C), dtype=flat.dtype)
    a, b, c, d = (flat[off[k]:off[k] + C] for k in "abcd")
    for o in range(C):
        K[:, :, o, o] = np.array([[a[o], b[o], c[o]], [d[o], gamma, -d[o]],
                                  [-c[o], -b[o], -a[o]]], dtype=flat.dtype)
        n = C - o - 1
        if n > 0:
            Wo = flat[off["W"][o]:off["W"][o] + 9 * n].reshape(3, 3, n)
            K[:, :, o + 1:, o] = Wo                    # ci > o
            K[:, :, o, o + 1:] = -Wo[::-1, ::-1, :]    # ci' = o < o' = o+1+j
    return K


# ---- general-k layer -------------------------------------------------------


def diag_slots_general(k: int, antisymmetric: bool = True):
    """Free-scalar positions of one diagonal block in creation order
    (tfkeras_layer_Conv2DAntisymmetric.py:231-264)."""
    slots = []
    for i in range(k):
        for j in range(i, k):
            if j > i or (j == i and i <= k // 2 - 1):
                slots.append((i, j))
            elif j == i and i == k // 2 and k % 2 == 1 and not antisymmetric:
                slots.append((i, j))  # trainable centre only when not anti
    return slots


def num_params_general(C: int, k: int, antisymmetric: bool = True, use_bias: bool = True) -> int:
    nd = len(diag_slots_general(k, antisymmetric))
    return nd * C + k * k * C * (C - 1) // 2 + (C if use_bias else 0)


def split_params_general(flat, C, k, antisymmetric=True, use_bias=True):
    """Reference variable order of Conv2DAntisymmetric: for each o, the diagonal
    scalars [1,1,1,1] then W_o [k,k,C-o-1,1]; bias last (:117-143, :150-157)."""
    nd = len(diag_slots_general(k, antisymmetric))
    out, cur = [], 0
    for o in range(C):
        for _ in range(nd):
            out.append(flat[cur:cur + 1].reshape(1, 1, 1, 1)); cur += 1
        n = C - o - 1
        if n > 0:
            out.append(flat[cur:cur + k * k * n].reshape(k, k, n, 1)); cur += k * k * n
    if use_bias:
        out.append(flat[cur:cur + C].reshape(C)); cur += C
    return out


def _centrosymmetric_matrix(scalars, k, gamma, antisymmetric, dtype):
    """tfkeras_layer_Conv2DAntisymmetric.py:216-270 /
    antisymmetric_conv2d_utils.py:23-75 (there gamma == 0, centre non-trainable)."""
    m = np.zeros((k, k), dtype=dtype)
    it = iter(scalars)
    for i in range(k):
        for j in range(i, k):
            if j > i or (j == i and i <= k // 2 - 1):
                v = next(it)
                m[i, j] = v
                m[k - 1 - i, k - 1 - j] = -v if antisymmetric else v
            elif j == i and i == k // 2 and k % 2 == 1:
                m[i, j] = gamma if antisymmetric else next(it)
    return m


def assemble_kernel_general_literal(variables, C, k, gamma, antisymmetric=True):
    """Literal restatement of Conv2DAntisymmetric.build() (:107-145): dependent
    kernels are -E.W.E (:139) regardless of `antisymmetric`."""
    nd = len(diag_slots_general(k, antisymmetric))
    dtype = np.asarray(variables[0]).dtype
    E = np.eye(k, dtype=dtype)[::-1]
    single_output_kernels, independent_kernels = [], []
    cur = 0
    for o in range(C):
        scal = [np.asarray(variables[cur + t]).reshape(()) for t in range(nd)]
        cur += nd
        centro = _centrosymmetric_matrix(scal, k, gamma, antisymmetric, dtype).reshape(k, k, 1, 1)
        n = C - o - 1
        if n > 0:
            independent = np.asarray(variables[cur]); cur += 1
            single = np.concatenate([centro, independent], axis=2)
        else:
            single = centro
        for i in range(o):
            ct = -(E @ (independent_kernels[-(i + 1)][:, :, i, 0] @ E))
            single = np.concatenate([ct[:, :, None, None], single], axis=2)
        single_output_kernels.append(single)
        if n > 0:
            independent_kernels.append(independent)
    return np.concatenate(single_output_kernels, axis=3)


# ---------------------------------------------------------------------------
# Convolution (tf.nn.conv2d NHWC / HWIO / SAME, cross-correlation)
# ---------------------------------------------------------------------------


def same_padding(in_size: int, k: int, s: int):
    """TF SAME: out=ceil(in/s), pad_total=max((out-1)s+k-in,0), before=total//2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x: np.ndarray, K: np.ndarray, strides=(1, 1)) -> np.ndarray:
    """tf.nn.conv2d(x, K, [1,s0,s1,1], 'SAME', NHWC) as called at
    tfkeras_layer_Conv2DAntisymmetric3By3.py:159-166.  x [N,H,W,Ci], K [kh,kw,Ci,Co]."""
    N, H, W, Ci = x.shape
    kh, kw, Ci2, Co = K.shape
    assert Ci == Ci2
    oh, pt, pb = same_padding(H, kh, strides[0])
    ow, pl, pr = same_padding(W, kw, strides[1])
    xp = np.zeros((N, H + pt + pb, W + pl + pr, Ci), dtype=x.dtype)
    xp[:, pt:pt + H, pl:pl + W, :] = x
    y = np.zeros((N, oh, ow, Co), dtype=np.result_type(x.dtype, K.dtype))
    for a in range(kh):
        for b in range(kw):
            win = xp[:, a:a + (oh - 1) * strides[0] + 1:strides[0],
                     b:b + (ow - 1) * strides[1] + 1:strides[1], :]
            y += win @ K[a, b]
    return y


def layer_call(x, K, bias=None, strides=(1, 1)):
    """Conv2DAntisymmetric3By3.call (:157-171)."""
    y = conv2d_same(x, K, strides)
    if bias is not None:
        y = y + bias
    return y


# ---------------------------------------------------------------------------
# Batch normalisation (tf.keras.layers.BatchNormalization(axis=3), Keras defaults)
# ---------------------------------------------------------------------------

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def bn_train_fwd(z, bn_gamma, bn_beta, eps=BN_EPS):
    """Training-mode BN over (N,H,W): biased variance for normalisation."""
    mu = z.mean(axis=(0, 1, 2))
    var = z.var(axis=(0, 1, 2))
    inv = 1.0 / np.sqrt(var + eps)
    zhat = (z - mu) * inv
    return bn_gamma * zhat + bn_beta, (zhat, inv, mu, var)


def bn_train_bwd(dout, zhat, inv, bn_gamma):
    """Standard BN backward.  Returns dz, dgamma, dbeta."""
    M = dout.shape[0] * dout.shape[1] * dout.shape[2]
    dbeta = dout.sum(axis=(0, 1, 2))
    dgamma = (dout * zhat).sum(axis=(0, 1, 2))
    dz = (bn_gamma * inv) * (dout - dbeta / M - zhat * (dgamma / M))
    return dz, dgamma, dbeta


def bn_infer(z, bn_gamma, bn_beta, moving_mean, moving_var, eps=BN_EPS):
    return bn_gamma * (z - moving_mean) / np.sqrt(moving_var + eps) + bn_beta


def bn_update_moving(moving_mean, moving_var, mu, var, M, momentum=BN_MOMENTUM):
    """Keras fused BN: moving variance is updated with the unbiased estimate."""
    unbiased = var * (M / max(M - 1, 1))
    return (moving_mean * momentum + mu * (1 - momentum),
            moving_var * momentum + unbiased * (1 - momentum))


# ---------------------------------------------------------------------------
# Euler step (models/tfkeras_resnets.py:69-92) and its closed-form backward
# ---------------------------------------------------------------------------


def euler_step_fwd(x, K, bias, h=1.0, bn=None):
    """single_layer_identity_block: conv -> BN? -> relu -> h* (if h != 1) -> + x.
    bn = (bn_gamma, bn_beta) for training-mode BN, or None.  Returns y, cache."""
    z = layer_call(x, K, bias)
    if bn is not None:
        u, bn_cache = bn_train_fwd(z, bn[0], bn[1])
    else:
        u, bn_cache = z, None
    r = np.maximum(u, 0)
    if h != 1.0:                       # :90-91 Lambda only when h != 1.0
        r = (np.asarray(h, dtype=r.dtype) * r).astype(r.dtype)
    y = r + x                          # :92 add([x, input_tensor])
    return y, {"x": x, "z": z, "u": u, "bn": bn_cache}


def conv_input_grad_stride1(dZ, K):
    """dX of a stride-1 SAME cross-correlation: correlate dZ with the spatially
    flipped, channel-transposed kernel."""
    Kt = np.transpose(K[::-1, ::-1, :, :], (0, 1, 3, 2))
    return conv2d_same(dZ, Kt)


def conv_kernel_grad_stride1(x, dZ, k=3):
    """G[a,b,ci,o] = sum_{n,y,x} x_pad[n,y+a,x+b,ci] * dZ[n,y,x,o]."""
    N, H, W, Ci = x.shape
    Co = dZ.shape[-1]
    p = k // 2
    xp = np.zeros((N, H + 2 * p, W + 2 * p, Ci), dtype=x.dtype)
    xp[:, p:p + H, p:p + W, :] = x
    G = np.zeros((k, k, Ci, Co), dtype=np.result_type(x.dtype, dZ.dtype))
    for a in range(k):
        for b in range(k):
            win = xp[:, a:a + H, b:b + W, :].reshape(-1, Ci)
            G[a, b] = win.T @ dZ.reshape(-1, Co)
    return G


def fold_grad_3by3(G, C, dbias=None):
    """Gradient of the loss wrt the packed free parameters given dense dL/dK
    (SURVEY.md App. A.3): S = G - rot180(G)^T; dW_o[a,b,j] = S[a,b,o+1+j,o];
    da_o = S[0,0,o,o], db_o = S[0,1,o,o], dc_o = S[0,2,o,o], dd_o = S[1,0,o,o]."""
    S = G - np.transpose(G[::-1, ::-1, :, :], (0, 1, 3, 2))
    off = offsets_3by3(C)
    flat = np.zeros(num_params_3by3(C, dbias is not None), dtype=G.dtype)
    idx = np.arange(C)
    flat[off["a"]:off["a"] + C] = S[0, 0, idx, idx]
    flat[off["b"]:off["b"] + C] = S[0, 1, idx, idx]
    flat[off["c"]:off["c"] + C] = S[0, 2, idx, idx]
    flat[off["d"]:off["d"] + C] = S[1, 0, idx, idx]
    for o in range(C - 1):
        n = C - o - 1
        flat[off["W"][o]:off["W"][o] + 9 * n] = S[:, :, o + 1:, o].reshape(-1)
    if dbias is not None:
        flat[off["bias"]:off["bias"] + C] = dbias
    return flat


def euler_step_bwd(dY, cache, K, h=1.0, bn_gamma=None, mask=None):
    """Backward of euler_step_fwd.  Returns dX, G (dense dL/dK), dbias,
    and (dgamma_bn, dbeta_bn) or None.  `mask` (bool, parity tests only) overrides the relu branch decision
    (u > 0): relu' is discontinuous, so an implementation whose pre-activation differs at rounding level may
    legitimately take the other branch for a few elements; tests bound their number separately."""
    u = cache["u"]
    dU = (np.asarray(h, dtype=dY.dtype) * dY) * ((u > 0) if mask is None else mask)
    if cache["bn"] is not None:
        zhat, inv, _, _ = cache["bn"]
        dZ, dgam, dbet = bn_train_bwd(dU, zhat, inv, bn_gamma)
        bn_grads = (dgam, dbet)
    else:
        dZ, bn_grads = dU, None
    dX = dY + conv_input_grad_stride1(dZ, K)
    G = conv_kernel_grad_stride1(cache["x"], dZ, K.shape[0])
    dbias = dZ.sum(axis=(0, 1, 2))
    return dX, G, dbias, bn_grads, dZ


# ---------------------------------------------------------------------------
# Loss / optimiser step semantics (training/training.py:295-301)
# ---------------------------------------------------------------------------


def softmax(logits):
    e = np.exp(logits - logits.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def keras_categorical_crossentropy(target, output, eps=1e-7):
    """K.categorical_crossentropy(from_logits=False): renormalise, clip, -sum t log p."""
    output = output / output.sum(axis=-1, keepdims=True)
    output = np.clip(output, eps, 1.0 - eps)
    return -(target * np.log(output)).sum(axis=-1)


def adam_step_tf1(theta, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-7):
    """tf.train.AdamOptimizer update (epsilon-hat formulation), t = 1-based step."""
    lr_t = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    theta = theta - lr_t * m / (np.sqrt(v) + eps)
    return theta, m, v


def gradient_mean_norm(g):
    """training/training.py:385-407: ||g||_2 / size(g)."""
    g = np.asarray(g).reshape(-1)
    return np.sqrt((g.astype(np.float64) ** 2).sum()) / g.size


# ---------------------------------------------------------------------------
# Initialiser (he_normal as redefined by the layer: truncated normal, sigma=sqrt(2/(k*k*C)))
# ---------------------------------------------------------------------------


def truncated_normal(rng: np.random.Generator, shape, stddev, dtype=np.float32):
    """tf.initializers.truncated_normal: resample outside +-2 sigma, no rescale."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(dtype)


def init_params_3by3(rng, C, dtype=np.float32, bias_std=0.0):
    """Fresh packed parameter vector (:95-98 initializer, :148-153 zero bias)."""
    n = num_params_3by3(C)
    flat = truncated_normal(rng, (n,), np.sqrt(2.0 / (9 * C)), dtype)
    off = offsets_3by3(C)
    if bias_std == 0.0:
        flat[off["bias"]:] = 0
    else:
        flat[off["bias"]:] = (rng.standard_normal(C) * bias_std).astype(dtype)
    return flat
