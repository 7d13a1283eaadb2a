"""Shared host-side machinery of the two antisymmetric convolution layers.

Mirrors the tf.keras `Layer` protocol the reference layers rely on
(`__call__` -> lazy `build(input_shape)` -> `call(input_tensor)`;
layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:60-208 of the reference) on top
of the C ABI in include/b200ode.h.  Tensors are NHWC torch CUDA tensors, or any
object exporting `__dlpack__` (consumed zero-copy).  PyTorch is only plumbing here
(device memory, streams, autograd tape); all arithmetic happens in libb200ode.so.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from .. import _abi


def as_torch(x) -> torch.Tensor:
    """Zero-copy view of `x` as a torch tensor (DLPack for foreign tensors)."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "__dlpack__"):
        return torch.from_dlpack(x)
    raise TypeError("expected a torch tensor or an object exporting __dlpack__, got %r" % type(x))


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _alloc_workspace(nbytes, device):
    """Caller-owned workspace block for the *_set_workspace / glue entry points (torch allocations are 512-byte aligned)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def truncated_normal_(t: torch.Tensor, stddev: float, generator=None):
    """tf.initializers.truncated_normal: resample outside +-2 sigma, no variance rescale
    (layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:95-98)."""
    torch.nn.init.trunc_normal_(t, mean=0.0, std=stddev, a=-2.0 * stddev, b=2.0 * stddev, generator=generator)
    return t


class LayerHandle:
    """Owns one b200ode_layer_t (staged tensor-core weights, TMA descriptors)."""

    def __init__(self, channels, ksize, gamma, strides, use_bias, antisymmetric, precision, layout):
        _abi.require_device()
        h = ctypes.c_void_p()
        _abi.check(_abi.lib().b200ode_layer_create(int(channels), int(ksize), float(gamma), int(strides[0]),
                                                   int(strides[1]), int(bool(use_bias)), int(bool(antisymmetric)),
                                                   int(precision), int(layout), ctypes.byref(h)))
        self._h = h
        self.channels = int(channels)
        self.ksize = int(ksize)
        self.strides = (int(strides[0]), int(strides[1]))
        self.num_params = int(_abi.lib().b200ode_layer_num_params(h))
        self.effective_mode = int(_abi.lib().b200ode_layer_effective_mode(h))
        self.io_dtype = torch.bfloat16 if self.effective_mode == _abi.PREC_FAST_BF16 else torch.float32
        self._packed_key = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _abi.lib().b200ode_layer_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- workspace (include/b200ode.h: the caller owns it; unbound handles draw from the stream-ordered pool) ----------
    def workspace_bytes(self, N, H, W):
        n = ctypes.c_size_t()
        _abi.check(_abi.lib().b200ode_layer_workspace_bytes(self._h, int(N), int(H), int(W), ctypes.byref(n)))
        return int(n.value)

    def bind_workspace(self, ws):
        """ws: uint8 CUDA tensor (kept alive by the handle) or None to unbind."""
        _abi.check(_abi.lib().b200ode_layer_set_workspace(self._h, _ptr(ws), 0 if ws is None else ws.numel()))
        self._ws = ws

    # -- K1 ------------------------------------------------------------------------------------
    def pack(self, params: torch.Tensor, dense_out: torch.Tensor = None, force=False):
        """Stage the weights if `params` changed since the last call (torch version counter)."""
        key = (params.data_ptr(), params._version)
        if not force and dense_out is None and key == self._packed_key:
            return
        _abi.check(_abi.lib().b200ode_pack_kernel(self._h, _ptr(params), _ptr(dense_out), _stream_ptr()))
        self._packed_key = key

    def out_hw(self, H, W):
        return -(-H // self.strides[0]), -(-W // self.strides[1])

    # -- K2 ------------------------------------------------------------------------------------
    def forward(self, x, h=1.0, flags=_abi.F_BIAS, want_mask=False, want_z=False, want_y=True):
        N, H, W, C = x.shape
        Ho, Wo = self.out_hw(H, W)
        y = torch.empty((N, Ho, Wo, C), dtype=self.io_dtype, device=x.device) if want_y else None
        mask = torch.empty((N, Ho, Wo, (C + 7) // 8), dtype=torch.uint8, device=x.device) if want_mask else None
        z = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device=x.device) if want_z else None
        _abi.check(_abi.lib().b200ode_euler_fwd(self._h, _ptr(x), _ptr(y), _ptr(mask), _ptr(z), N, H, W, float(h),
                                                int(flags), _stream_ptr()))
        return y, mask, z

    # -- K2 + BatchNorm statistics from the epilogue ----------------------------------------------
    def forward_bn_stats(self, x, z, stats_ws):
        """z = conv_K(x) + b (fp32, into `z`) and the per-channel partial sums of z, z*z (rows of `stats_ws`) in one
        kernel; returns the number of rows written (b200ode_euler_fwd_bn_stats)."""
        N, H, W, C = x.shape
        rows = ctypes.c_int()
        _abi.check(_abi.lib().b200ode_euler_fwd_bn_stats(self._h, _ptr(x), _ptr(z), _ptr(stats_ws), ctypes.byref(rows), N, H, W,
                                                         _stream_ptr()))
        return rows.value

    # -- K3 ------------------------------------------------------------------------------------
    def dgrad(self, dz, skip, in_hw):
        N, _, _, C = dz.shape
        H, W = in_hw
        dx = torch.empty((N, H, W, C), dtype=self.io_dtype, device=dz.device)
        _abi.check(_abi.lib().b200ode_euler_dgrad(self._h, _ptr(dz), _ptr(skip), _ptr(dx), N, H, W, _stream_ptr()))
        return dx

    # -- K4 ------------------------------------------------------------------------------------
    def wgrad(self, x, dz, want_dense=False):
        N, H, W, C = x.shape
        g = torch.empty(self.num_params, dtype=torch.float32, device=x.device)
        G = torch.empty((self.ksize, self.ksize, C, C), dtype=torch.float32, device=x.device) if want_dense else None
        _abi.check(_abi.lib().b200ode_euler_wgrad(self._h, _ptr(x), _ptr(dz), _ptr(g), _ptr(G), N, H, W, 0,
                                                  _stream_ptr()))
        return (g, G) if want_dense else g


class ChainHandle:
    """Owns one b200ode_chain_t: `n_layers` antisymmetric 3x3 Euler steps of equal shape run by the
    persistent per-image kernels (one launch per direction).  Mirrors n stacked
    single_layer_identity_block calls (models/tfkeras_resnets.py:28-94, stage loop :575-593)."""

    def __init__(self, channels, n_layers, gamma, use_bias=True, precision=_abi.PREC_FAST_TF32):
        _abi.require_device()
        h = ctypes.c_void_p()
        _abi.check(_abi.lib().b200ode_chain_create(int(channels), int(n_layers), float(gamma), int(bool(use_bias)),
                                                   int(precision), ctypes.byref(h)))
        self._h = h
        self.channels, self.n_layers, self.precision = int(channels), int(n_layers), int(precision)
        self.f16 = self.precision == _abi.PREC_FAST_F16
        # dtype of the saved weight-gradient operands (acts, dz_all): fp16 in FAST_F16 mode (acts[l] = INPUT of step l)
        self.saved_dtype = torch.float16 if self.f16 else torch.float32
        self.num_params = int(_abi.lib().b200ode_chain_layer_params(h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _abi.lib().b200ode_chain_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @staticmethod
    def supported(channels, H, W, precision=_abi.PREC_FAST_TF32):
        return bool(_abi.lib().b200ode_chain_supported(int(channels), int(H), int(W), int(precision)))

    def workspace_bytes(self, N, H, W):
        n = ctypes.c_size_t()
        _abi.check(_abi.lib().b200ode_chain_workspace_bytes(self._h, int(N), int(H), int(W), ctypes.byref(n)))
        return int(n.value)

    def bind_workspace(self, ws):
        _abi.check(_abi.lib().b200ode_chain_set_workspace(self._h, _ptr(ws), 0 if ws is None else ws.numel()))
        self._ws = ws

    def pack(self, params, layer_stride=None):
        _abi.check(_abi.lib().b200ode_chain_pack(self._h, _ptr(params), int(layer_stride or self.num_params),
                                                 _stream_ptr()))

    def forward(self, x0, h, n_steps=None, acts=None, masks=None, y_final=None):
        N, H, W, C = x0.shape
        n_steps = self.n_layers if n_steps is None else int(n_steps)
        _abi.check(_abi.lib().b200ode_chain_fwd(self._h, _ptr(x0), _ptr(acts), _ptr(masks), _ptr(y_final), N, H, W,
                                                float(h), n_steps, _stream_ptr()))

    def dgrad(self, dy, masks, dz_all, dx, h, dy_amax=None):
        """dy_amax: optional 1-element fp32 device tensor already holding max|dy| (written by b200ode_head_fwd_bwd_amax /
        b200ode_transition_dgrad_amax): FAST_F16 chains then skip their own reduction over dy."""
        N, H, W, C = dy.shape
        if dy_amax is not None:
            _abi.check(_abi.lib().b200ode_chain_dgrad_amax(self._h, _ptr(dy), _ptr(masks), _ptr(dz_all), _ptr(dx), N, H, W,
                                                           float(h), _ptr(dy_amax), _stream_ptr()))
            return
        _abi.check(_abi.lib().b200ode_chain_dgrad(self._h, _ptr(dy), _ptr(masks), _ptr(dz_all), _ptr(dx), N, H, W,
                                                  float(h), _stream_ptr()))

    def wgrad(self, x0, acts, dz_all, grad, layer_stride=None):
        N, H, W, C = x0.shape
        _abi.check(_abi.lib().b200ode_chain_wgrad(self._h, _ptr(x0), _ptr(acts), _ptr(dz_all), _ptr(grad),
                                                  int(layer_stride or self.num_params), N, H, W, _stream_ptr()))


BN_EPS, BN_MOMENTUM = 1e-3, 0.99        # tf.keras.layers.BatchNormalization defaults (models/tfkeras_resnets.py:85-87)


class BNEulerStep:
    """x + h*relu(BN(conv_K(x)+b)) in training mode on the CUDA kernels (single_layer_identity_block with
    use_batch_norm=True, models/tfkeras_resnets.py:70-92), shared by the Keras-shaped model mirror and the EulerNet trainer.

    forward : conv kernel (z + partial sums from its epilogue) -> bn_stats_finalize (mean, inv_std, scale, shift, moving
              statistics) -> euler_tail (y = x + h*relu(z*scale+shift))                                   3 launches
    backward: bn_bwd_reduce (2) -> bn_bwd_apply -> dgrad with the skip add fused -> wgrad + fold
    `allreduce`: optional callable summing a small fp32 tensor over the data-parallel ranks in place (SyncBN: the
    statistics and the two backward sums are taken over the GLOBAL batch, `world` ranks of equal local size)."""

    @staticmethod
    def stats_workspace(C, device):
        return torch.empty(2 * _abi.COLSUM_PARTS * C, dtype=torch.float32, device=device)

    @staticmethod
    def forward(hd, x, bn_gamma, bn_beta, moving_mean, moving_var, h, z, y, stat, stats_ws, allreduce=None, world=1):
        """stat: fp32 [6, C] = (mean, inv_std, scale, shift, sum, sumsq) written here; z, y: [N,H,W,C] fp32 outputs."""
        lib, st = _abi.lib(), _stream_ptr()
        N, H, W, C = x.shape
        M = N * H * W
        rows = hd.forward_bn_stats(x, z, stats_ws)
        if allreduce is None or world == 1:
            _abi.check(lib.b200ode_bn_stats_finalize(_ptr(stats_ws), rows, None, None, _ptr(bn_gamma), _ptr(bn_beta), _ptr(stat[0]),
                                                     _ptr(stat[1]), _ptr(stat[2]), _ptr(stat[3]), _ptr(moving_mean), _ptr(moving_var),
                                                     M, C, BN_EPS, BN_MOMENTUM, st))
        else:
            _abi.check(lib.b200ode_bn_stats_finalize(_ptr(stats_ws), rows, _ptr(stat[4]), _ptr(stat[5]), None, None, None, None, None,
                                                     None, None, None, M, C, BN_EPS, BN_MOMENTUM, st))
            allreduce(stat[4:6])
            _abi.check(lib.b200ode_bn_finalize(_ptr(stat[4]), _ptr(stat[5]), _ptr(bn_gamma), _ptr(bn_beta), _ptr(stat[0]), _ptr(stat[1]),
                                               _ptr(stat[2]), _ptr(stat[3]), _ptr(moving_mean), _ptr(moving_var), M * world, C,
                                               BN_EPS, BN_MOMENTUM, st))
        flags = _abi.F_RELU | _abi.F_RESIDUAL | (_abi.F_SCALE if h != 1.0 else 0)
        _abi.check(lib.b200ode_euler_tail(_ptr(z), _ptr(stat[2]), _ptr(stat[3]), _ptr(x), _ptr(y), None, M, C, float(h), flags, st))

    @staticmethod
    def backward(hd, x, dy, z, stat, bn_gamma, h, dz, dx, grad_params, dbn, stats_ws, allreduce=None, world=1, want_dx=True):
        """dbn: fp32 [2, C] receives (dgamma, dbeta) of THIS rank's batch; dz / dx: [N,H,W,C] fp32 outputs;
        grad_params: packed conv gradient (fold + bias) written by the wgrad kernels."""
        lib, st = _abi.lib(), _stream_ptr()
        N, H, W, C = x.shape
        M = N * H * W
        _abi.check(lib.b200ode_bn_bwd_reduce(_ptr(dy), _ptr(z), _ptr(stat[2]), _ptr(stat[3]), _ptr(stat[0]), _ptr(stat[1]),
                                             _ptr(dbn[0]), _ptr(dbn[1]), _ptr(stats_ws), M, C, float(h), st))
        red = dbn
        if allreduce is not None and world > 1:
            red = stat[4:6]                      # global sums for the data gradient; the bucket keeps the local ones
            red.copy_(dbn)
            allreduce(red)
        _abi.check(lib.b200ode_bn_bwd_apply(_ptr(dy), _ptr(z), _ptr(stat[2]), _ptr(stat[3]), _ptr(stat[0]), _ptr(stat[1]),
                                            _ptr(bn_gamma), _ptr(red[0]), _ptr(red[1]), _ptr(dz), M * world, C, float(h), st))
        if want_dx:
            _abi.check(lib.b200ode_euler_dgrad(hd._h, _ptr(dz), _ptr(dy), _ptr(dx), N, H, W, st))
        _abi.check(lib.b200ode_euler_wgrad(hd._h, _ptr(x), _ptr(dz), _ptr(grad_params), None, N, H, W, 0, st))

    @staticmethod
    def inference(hd, x, bn_gamma, bn_beta, moving_mean, moving_var, h, z, y):
        """Inference mode: BN with the moving statistics (an affine per channel) fused into the tail."""
        lib, st = _abi.lib(), _stream_ptr()
        N, H, W, C = x.shape
        scale = bn_gamma / torch.sqrt(moving_var + BN_EPS)
        shift = bn_beta - moving_mean * scale
        _abi.check(lib.b200ode_euler_fwd(hd._h, _ptr(x), None, None, _ptr(z), N, H, W, 1.0, _abi.F_BIAS, st))
        flags = _abi.F_RELU | _abi.F_RESIDUAL | (_abi.F_SCALE if h != 1.0 else 0)
        _abi.check(lib.b200ode_euler_tail(_ptr(z), _ptr(scale), _ptr(shift), _ptr(x), _ptr(y), None, N * H * W, C, float(h), flags, st))


def relu_scale_bwd(dy, mask, h):
    dz = torch.empty_like(dy)
    C = dy.shape[-1]
    _abi.check(_abi.lib().b200ode_relu_scale_bwd(_ptr(dy), _ptr(mask), _ptr(dz), dy.numel() // C, C, float(h),
                                                 int(dy.dtype == torch.bfloat16), _stream_ptr()))
    return dz


class _ConvFn(torch.autograd.Function):
    """y = conv_K(x) + b   (Conv2DAntisymmetric3By3.call, reference :157-171)."""

    @staticmethod
    def forward(ctx, x, params, handle):
        handle.pack(params)
        y, _, _ = handle.forward(x, 1.0, _abi.F_BIAS)
        ctx.handle = handle
        ctx.save_for_backward(x, params)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, params = ctx.saved_tensors
        hd = ctx.handle
        dy = dy.contiguous()
        hd.pack(params)
        dx = hd.dgrad(dy, None, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[0] else None
        gp = None
        if ctx.needs_input_grad[1]:
            gp = _wgrad_any(hd, x, dy)
        return dx, gp, None


def _wgrad_any(hd, x, dz):
    """Weight + bias gradient (the bias column sums are accumulated inside the wgrad kernel)."""
    return hd.wgrad(x, dz)


class _EulerFn(torch.autograd.Function):
    """y = x + h*relu(conv_K(x)+b)   (single_layer_identity_block without BN, reference
    models/tfkeras_resnets.py:69-92), one fused kernel forward; backward = relu/scale mask kernel,
    dgrad with fused skip add, wgrad + fold."""

    @staticmethod
    def forward(ctx, x, params, handle, h):
        handle.pack(params)
        y, mask, _ = handle.forward(x, h, _abi.F_EULER, want_mask=True)
        ctx.handle, ctx.h = handle, h
        ctx.save_for_backward(x, params, mask)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, params, mask = ctx.saved_tensors
        hd = ctx.handle
        dy = dy.contiguous()
        hd.pack(params)
        dz = relu_scale_bwd(dy, mask, ctx.h)
        dx = hd.dgrad(dz, dy, (x.shape[1], x.shape[2])) if ctx.needs_input_grad[0] else None
        gp = _wgrad_any(hd, x, dz) if ctx.needs_input_grad[1] else None
        return dx, gp, None, None


class AntisymmetricConvBase:
    """Keras-shaped base class (constructor kwargs name/trainable/dtype like tf.keras.layers.Layer)."""

    _layout = _abi.LAYOUT_3BY3
    _uid = 0

    def __init__(self, name=None, trainable=True, dtype="float32", precision="strict", device=None, seed=None, **kwargs):
        if kwargs:
            raise TypeError("unexpected keyword arguments: %s" % sorted(kwargs))
        if dtype not in ("float32", torch.float32, None):
            raise ValueError("the reference layer computes in float32; use precision='fast_bf16' for bf16 I/O")
        if precision not in _abi.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_abi.PRECISIONS))
        if name is None:
            AntisymmetricConvBase._uid += 1
            name = "%s_%d" % (type(self).__name__.lower(), AntisymmetricConvBase._uid)
        self.name = name
        self.trainable = trainable
        self.dtype = "float32"
        self.precision = precision
        self.built = False
        self._device = device
        self._seed = seed
        self._handle = None
        self.packed = None      # flat fp32 parameter vector in the reference's variable order

    # ---- Keras protocol -------------------------------------------------------------------------
    def __call__(self, input_tensor):
        x = as_torch(input_tensor)
        if not self.built:
            self.build(tuple(x.shape))
        return self.call(x)

    def _build_common(self, input_shape, ksize, antisymmetric):
        self.num_channels = int(input_shape[-1])   # channels-last; out channels == in channels
        C = self.num_channels
        dev = torch.device(self._device if self._device is not None else "cuda")
        self._handle = LayerHandle(C, ksize, self.gamma, self.strides, self.use_bias, antisymmetric,
                                   _abi.PRECISIONS[self.precision], self._layout)
        n = self._handle.num_params
        flat = torch.empty(n, dtype=torch.float32)
        init = self.kernel_initializer
        gen = torch.Generator().manual_seed(self._seed) if self._seed is not None else None
        nk = n - (C if self.use_bias else 0)
        if init == "he_normal":
            # the layer redefines he_normal as truncated normal with sigma = sqrt(2/(k*k*C))
            truncated_normal_(flat[:nk], math.sqrt(2.0 / (ksize * ksize * C)), gen)
        elif callable(init):
            flat[:nk] = torch.as_tensor(init((nk,)), dtype=torch.float32).reshape(-1)
        else:
            raise ValueError("unsupported kernel_initializer %r" % (init,))
        if self.use_bias:
            flat[nk:] = 0.0
        self.packed = flat.to(dev).requires_grad_(bool(self.trainable))
        self.built = True

    def call(self, input_tensor):
        x = self._check_input(as_torch(input_tensor))
        return _ConvFn.apply(x, self.packed, self._handle)

    def euler_step(self, input_tensor, h=1.0):
        """Fused x + h*relu(conv(x)+b): the non-BN body of single_layer_identity_block."""
        x = as_torch(input_tensor)
        if not self.built:
            self.build(tuple(x.shape))
        x = self._check_input(x)
        return _EulerFn.apply(x, self.packed, self._handle, float(h))

    def _check_input(self, x):
        if x.dim() != 4 or x.shape[-1] != self.num_channels:
            raise ValueError("expected NHWC input with %d channels, got shape %s" % (self.num_channels, tuple(x.shape)))
        if not x.is_cuda:
            raise _abi.B200OdeError("b200ode has no CPU path: input tensor must live on a CUDA device")
        want = self._handle.io_dtype
        if x.dtype != want:
            raise ValueError("precision=%r expects %s inputs, got %s" % (self.precision, want, x.dtype))
        return x.contiguous()

    def compute_output_shape(self, input_shape):
        return input_shape                       # reference :173-175

    # ---- weights --------------------------------------------------------------------------------
    @property
    def kernel(self):
        """Assembled [k,k,C,C] kernel as a device tensor (the reference's `self.kernel`)."""
        C, k = self.num_channels, self._handle.ksize
        K = torch.empty((k, k, C, C), dtype=torch.float32, device=self.packed.device)
        self._handle.pack(self.packed.detach(), K)
        return K

    @property
    def bias(self):
        return self.packed.detach()[-self.num_channels:] if self.use_bias else None

    def get_kernel(self):
        return self.kernel.cpu().numpy()          # reference :188-199 returns an ndarray

    def get_bias(self):
        return self.bias.cpu().numpy()            # reference :201-208

    @property
    def trainable_weights(self):
        return [self.packed] if self.trainable else []

    @property
    def weights(self):
        return [self.packed]

    def _variable_shapes(self):
        raise NotImplementedError

    def get_weights(self):
        """List of numpy arrays in the reference's variable order and shapes."""
        flat = self.packed.detach().cpu().numpy()
        out, cur = [], 0
        for shp in self._variable_shapes():
            n = int(np.prod(shp))
            out.append(flat[cur:cur + n].reshape(shp).copy())
            cur += n
        return out

    def set_weights(self, weights):
        shapes = self._variable_shapes()
        if len(weights) != len(shapes):
            raise ValueError("expected %d weight arrays, got %d" % (len(shapes), len(weights)))
        parts = []
        for w, shp in zip(weights, shapes):
            w = np.asarray(w, dtype=np.float32)
            if tuple(w.shape) != tuple(shp):
                raise ValueError("weight shape %s does not match %s" % (w.shape, shp))
            parts.append(w.reshape(-1))
        with torch.no_grad():
            self.packed.copy_(torch.from_numpy(np.concatenate(parts)))
