"""Host-side mirror of the reference's model builders for the single-block (Euler) ResNets
(reference models/tfkeras_resnets.py:28-94 `single_layer_identity_block`, :204-269
`single_layer_conv_block`, :427-509 `build_single_block_resnet`, :511-604
`get_single_block_resnet_build_function`): same function names, argument names and defaults.

The reference builds a static tf.keras graph; here the same calls build an eager `Model` whose
layers are created on first use and reused afterwards.  The Euler step
x_{n+1} = x_n + h*relu(BN?(conv_K(x_n)+b)) runs on the hand-written CUDA kernels
(fused single kernel without BN; conv with the batch statistics from its epilogue + BN finalize + fused tail with BN).
Stem / transition / head layers are regular Keras layers in the reference and run as torch ops here
(SURVEY.md section 8f-1).  The bottleneck ResNet-50/101/152 builders are out of scope (SURVEY.md
section 2, row 6).
"""
from __future__ import annotations

import math
from collections import OrderedDict
import zlib

import numpy as np
import torch

from ..layers._base import BNEulerStep, as_torch, truncated_normal_
from ..layers.tfkeras_layer_Conv2DAntisymmetric3By3 import Conv2DAntisymmetric3By3
from ..training import conv2d_same_nhwc

BN_EPS, BN_MOMENTUM = 1e-3, 0.99     # Keras BatchNormalization defaults


# ---------------------------------------------------------------------------------------------------
# build scope: layers are registered by name so that repeated calls reuse them
# ---------------------------------------------------------------------------------------------------
class _Scope:
    current = None

    def __init__(self, precision="strict", seed=None, training=True):
        self.layers, self.order = {}, []
        self.precision, self.seed, self.training = precision, seed, training

    def get(self, name, factory):
        if name not in self.layers:
            self.layers[name] = factory()
            self.order.append(name)
        return self.layers[name]

    def seed_for(self, name):
        """Per-layer seed derived from (model seed, layer name): same-shape layers must not start as copies of each
        other (the reference draws every he_normal variable independently)."""
        if self.seed is None:
            return None
        return (int(self.seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 31 - 1)


def _scope():
    if _Scope.current is None:
        _Scope.current = _Scope()
    return _Scope.current


class _KernelBiasWeights:
    """Keras `get_weights()` / `set_weights()` of a (kernel, bias) layer."""

    def get_weights(self):
        if self.kernel is None:
            return []
        return [self.kernel.detach().cpu().numpy(), self.bias.detach().cpu().numpy()]

    def set_weights(self, weights):
        if self.kernel is None:
            raise RuntimeError("layer %r has no weights yet: call the model once first" % getattr(self, "name", self))
        if len(weights) != 2:
            raise ValueError("expected [kernel, bias], got %d arrays" % len(weights))
        with torch.no_grad():
            for t, w in zip((self.kernel, self.bias), weights):
                w = np.asarray(w, dtype=np.float32)
                if tuple(w.shape) != tuple(t.shape):
                    raise ValueError("weight shape %s does not match %s" % (w.shape, tuple(t.shape)))
                t.copy_(torch.from_numpy(w))


class _RegularConv(_KernelBiasWeights):
    """Keras Conv2D(padding='same', kernel_initializer='he_normal') as torch ops."""

    def __init__(self, filters, kernel_size, strides, name, seed=None, padding='same'):
        self.filters, self.kernel_size, self.strides, self.name = filters, kernel_size, tuple(strides), name
        self.kernel = self.bias = None
        self.seed = seed
        self.padding = padding

    def __call__(self, x):
        if self.kernel is None:
            k, ci = self.kernel_size, x.shape[-1]
            g = torch.Generator().manual_seed(self.seed) if self.seed is not None else None
            w = truncated_normal_(torch.empty(k, k, ci, self.filters), math.sqrt(2.0 / (k * k * ci)) / 0.87962566103423978, g)
            self.kernel = w.to(x.device).requires_grad_(True)
            self.bias = torch.zeros(self.filters, device=x.device, requires_grad=True)
        if self.padding == 'valid':
            y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), self.kernel.permute(3, 2, 0, 1), self.bias, stride=self.strides)
            return y.permute(0, 2, 3, 1).contiguous()
        return conv2d_same_nhwc(x, self.kernel, self.bias, self.strides)

    @property
    def trainable_weights(self):
        return [self.kernel, self.bias]


class _BatchNorm:
    """tf.keras.layers.BatchNormalization(axis=3) state (gamma, beta, moving statistics)."""

    def __init__(self, name):
        self.name, self.gamma = name, None

    def ensure(self, C, device):
        if self.gamma is None:
            self.gamma = torch.ones(C, device=device, requires_grad=True)
            self.beta = torch.zeros(C, device=device, requires_grad=True)
            self.moving_mean = torch.zeros(C, device=device)
            self.moving_var = torch.ones(C, device=device)

    def torch_apply(self, x, training):
        self.ensure(x.shape[-1], x.device)
        if training:
            mu, var = x.mean(dim=(0, 1, 2)), x.var(dim=(0, 1, 2), unbiased=False)
            with torch.no_grad():
                M = x.numel() // x.shape[-1]
                self.moving_mean.mul_(BN_MOMENTUM).add_(mu.detach(), alpha=1 - BN_MOMENTUM)
                self.moving_var.mul_(BN_MOMENTUM).add_(var.detach() * (M / max(M - 1, 1)), alpha=1 - BN_MOMENTUM)
        else:
            mu, var = self.moving_mean, self.moving_var
        return self.gamma * (x - mu) / torch.sqrt(var + BN_EPS) + self.beta

    @property
    def trainable_weights(self):
        return [self.gamma, self.beta]

    def get_weights(self):
        """Keras order: gamma, beta, moving_mean, moving_variance."""
        if self.gamma is None:
            return []
        return [t.detach().cpu().numpy() for t in (self.gamma, self.beta, self.moving_mean, self.moving_var)]

    def set_weights(self, weights):
        if self.gamma is None or len(weights) != 4:
            raise ValueError("BatchNormalization layer %r takes [gamma, beta, moving_mean, moving_variance] once built" % self.name)
        with torch.no_grad():
            for t, w in zip((self.gamma, self.beta, self.moving_mean, self.moving_var), weights):
                t.copy_(torch.from_numpy(np.asarray(w, dtype=np.float32)))


class _EulerBNFn(torch.autograd.Function):
    """x + h*relu(BN(conv_K(x)+b)) in training mode on the CUDA kernels (layers._base.BNEulerStep): conv with the batch
    statistics taken from its epilogue -> finalize -> fused tail; backward: BN reductions, BN apply (dz), dgrad with the
    skip add, wgrad + fold."""

    @staticmethod
    def forward(ctx, x, params, bn_gamma, bn_beta, layer, bn, h):
        hd = layer._handle
        hd.pack(params)
        C = x.shape[-1]
        z, y = torch.empty_like(x), torch.empty_like(x)
        stat = torch.empty((6, C), dtype=torch.float32, device=x.device)
        ws = BNEulerStep.stats_workspace(C, x.device)
        BNEulerStep.forward(hd, x, bn_gamma.detach(), bn_beta.detach(), bn.moving_mean, bn.moving_var, h, z, y, stat, ws)
        ctx.save_for_backward(x, params, bn_gamma, z, stat)
        ctx.hd, ctx.h, ctx.ws = hd, h, ws
        return y

    @staticmethod
    def backward(ctx, dy):
        x, params, bn_gamma, z, stat = ctx.saved_tensors
        hd, h = ctx.hd, ctx.h
        dy = dy.contiguous()
        C = x.shape[-1]
        dbn = torch.empty((2, C), dtype=torch.float32, device=x.device)
        dz, dx = torch.empty_like(x), torch.empty_like(x)
        gp = torch.empty(hd.num_params, dtype=torch.float32, device=x.device)
        hd.pack(params)
        BNEulerStep.backward(hd, x, dy, z, stat, bn_gamma.detach(), h, dz, dx, gp, dbn, ctx.ws)
        return dx, gp, dbn[0], dbn[1], None, None, None


def single_layer_identity_block(input_tensor,
                                kernel_size,
                                antisymmetric,
                                use_batch_norm,
                                stage,
                                block,
                                h=1.0,
                                gamma=0.0,
                                kernel_regularizer=None,
                                bias_regularizer=None):
    """Euler step: conv -> BN? -> relu -> h* (only if h != 1.0) -> + input
    (reference models/tfkeras_resnets.py:28-94)."""
    sc = _scope()
    x = as_torch(input_tensor)
    conv_name_base = 'res' + str(stage) + '_' + str(block) + '_branch'
    bn_name_base = 'bn' + str(stage) + '_' + str(block) + '_branch'
    if antisymmetric:
        layer = sc.get(conv_name_base + '2', lambda: Conv2DAntisymmetric3By3(
            gamma=gamma, strides=(1, 1), use_bias=True, kernel_initializer='he_normal',
            kernel_regularizer=kernel_regularizer, name=conv_name_base + '2', precision=sc.precision, seed=sc.seed_for(conv_name_base + '2')))
        if not use_batch_norm:
            return layer.euler_step(x, h)                       # one fused kernel
        if not layer.built:
            layer.build(tuple(x.shape))
        bn = sc.get(bn_name_base + '2', lambda: _BatchNorm(bn_name_base + '2'))
        bn.ensure(x.shape[-1], x.device)
        if layer._handle.io_dtype != torch.float32:
            raise ValueError("use_batch_norm=True needs fp32 activations: the BatchNorm tail kernels are fp32-only "
                             "(precision=%r computes with %s I/O)" % (sc.precision, layer._handle.io_dtype))
        if sc.training:
            return _EulerBNFn.apply(layer._check_input(x), layer.packed, bn.gamma, bn.beta, layer, bn, float(h))
        y = bn.torch_apply(layer(x), False)
    else:
        conv = sc.get(conv_name_base + '2', lambda: _RegularConv(int(x.shape[-1]), kernel_size, (1, 1), conv_name_base + '2', sc.seed_for(conv_name_base + '2')))
        y = conv(x)
        if use_batch_norm:
            y = sc.get(bn_name_base + '2', lambda: _BatchNorm(bn_name_base + '2')).torch_apply(y, sc.training)
    y = torch.relu(y)
    if h != 1.0:
        y = h * y
    return y + x


def single_layer_conv_block(input_tensor,
                            kernel_size,
                            num_filters,
                            strides,
                            use_batch_norm,
                            stage,
                            block,
                            kernel_regularizer=None,
                            bias_regularizer=None):
    """Transition block: strided regular conv + 1x1 strided shortcut, no h
    (reference models/tfkeras_resnets.py:204-269)."""
    sc = _scope()
    x = as_torch(input_tensor)
    conv_name_base = 'res' + str(stage) + '_' + str(block) + '_branch'
    bn_name_base = 'bn' + str(stage) + '_' + str(block) + '_branch'
    main = sc.get(conv_name_base + '2', lambda: _RegularConv(num_filters, kernel_size, strides, conv_name_base + '2', sc.seed_for(conv_name_base + '2')))(x)
    short = sc.get(conv_name_base + '1', lambda: _RegularConv(num_filters, 1, strides, conv_name_base + '1', sc.seed_for(conv_name_base + '1')))(x)
    if use_batch_norm:
        main = sc.get(bn_name_base + '2', lambda: _BatchNorm(bn_name_base + '2')).torch_apply(main, sc.training)
        short = sc.get(bn_name_base + '1', lambda: _BatchNorm(bn_name_base + '1')).torch_apply(short, sc.training)
    return torch.relu(main) + short


def _bottleneck_main(x, kernel_size, num_filters, antisymmetric, use_batch_norm, conv_name_base, bn_name_base,
                     strides_1_by_1, strides_k_by_k, gamma, kernel_regularizer):
    """1x1 -> kxk -> 1x1 main branch shared by the two bottleneck blocks (reference :148-195 and :352-400): the middle
    convolution is a Conv2DAntisymmetric3By3 when `antisymmetric` and `num_filters[1] is None` (square matrix)."""
    sc = _scope()

    def bn(y, tag):
        if use_batch_norm:
            y = sc.get(bn_name_base + tag, lambda: _BatchNorm(bn_name_base + tag)).torch_apply(y, sc.training)
        return y

    y = sc.get(conv_name_base + '2a', lambda: _RegularConv(num_filters[0], 1, strides_1_by_1, conv_name_base + '2a',
                                                           sc.seed_for(conv_name_base + '2a')))(x)
    y = torch.relu(bn(y, '2a'))
    if antisymmetric and (num_filters[1] is None):
        layer = sc.get(conv_name_base + '2b', lambda: Conv2DAntisymmetric3By3(
            gamma=gamma, strides=tuple(strides_k_by_k), use_bias=True, kernel_initializer='he_normal',
            kernel_regularizer=kernel_regularizer, name=conv_name_base + '2b', precision=sc.precision,
            seed=sc.seed_for(conv_name_base + '2b')))
        y = layer(y)                                     # the hot-path layer (stride 2 in version 1.5: CUDA-core kernels)
        if y.dtype != torch.float32:
            y = y.float()
    else:
        if num_filters[1] is None:
            raise ValueError("num_filters[1] may only be None for an antisymmetric block")
        y = sc.get(conv_name_base + '2b', lambda: _RegularConv(num_filters[1], kernel_size, strides_k_by_k, conv_name_base + '2b',
                                                               sc.seed_for(conv_name_base + '2b')))(y)
    y = torch.relu(bn(y, '2b'))
    y = sc.get(conv_name_base + '2c', lambda: _RegularConv(num_filters[2], 1, (1, 1), conv_name_base + '2c',
                                                           sc.seed_for(conv_name_base + '2c')))(y)
    return bn(y, '2c')


def bottleneck_identity_block(input_tensor,
                              kernel_size,
                              num_filters,
                              antisymmetric,
                              use_batch_norm,
                              stage,
                              block,
                              gamma=0.0,
                              kernel_regularizer=None,
                              bias_regularizer=None):
    """Bottleneck identity block 1x1 -> kxk -> 1x1, + input, relu (reference models/tfkeras_resnets.py:96-202; the
    antisymmetric layer is its middle convolution, :163-169)."""
    x = as_torch(input_tensor)
    conv_name_base = 'res' + str(stage) + '_' + str(block) + '_branch'
    bn_name_base = 'bn' + str(stage) + '_' + str(block) + '_branch'
    y = _bottleneck_main(x, kernel_size, num_filters, antisymmetric, use_batch_norm, conv_name_base, bn_name_base,
                         (1, 1), (1, 1), gamma, kernel_regularizer)
    return torch.relu(y + x)


def bottleneck_conv_block(input_tensor,
                          kernel_size,
                          num_filters,
                          antisymmetric,
                          use_batch_norm,
                          stage,
                          block,
                          version=1,
                          strides=(1, 1),
                          gamma=0.0,
                          kernel_regularizer=None,
                          bias_regularizer=None):
    """Bottleneck block with a strided 1x1 shortcut convolution (reference models/tfkeras_resnets.py:271-425): version 1
    strides the first 1x1 convolution, version 1.5 the kxk one (:341-348) -- there the antisymmetric layer (:370-376)
    runs with stride 2."""
    if version == 1:
        strides_1_by_1, strides_k_by_k = tuple(strides), (1, 1)
    elif version == 1.5:
        strides_1_by_1, strides_k_by_k = (1, 1), tuple(strides)
    else:
        raise ValueError("Supported values for `version` are 1 and 1.5.")
    sc = _scope()
    x = as_torch(input_tensor)
    conv_name_base = 'res' + str(stage) + '_' + str(block) + '_branch'
    bn_name_base = 'bn' + str(stage) + '_' + str(block) + '_branch'
    y = _bottleneck_main(x, kernel_size, num_filters, antisymmetric, use_batch_norm, conv_name_base, bn_name_base,
                         strides_1_by_1, strides_k_by_k, gamma, kernel_regularizer)
    short = sc.get(conv_name_base + '1', lambda: _RegularConv(num_filters[2], 1, tuple(strides), conv_name_base + '1',
                                                              sc.seed_for(conv_name_base + '1')))(x)
    if use_batch_norm:
        short = sc.get(bn_name_base + '1', lambda: _BatchNorm(bn_name_base + '1')).torch_apply(short, sc.training)
    return torch.relu(y + short)


class Model:
    """Eager stand-in for tf.keras.models.Model: callable, `predict`, `layers`, `trainable_weights`."""

    def __init__(self, fn, name, precision, seed):
        self._fn, self.name = fn, name
        self.scope = _Scope(precision, seed)

    def __call__(self, images, training=True):
        prev, _Scope.current = _Scope.current, self.scope
        self.scope.training = training
        try:
            return self._fn(as_torch(images))
        finally:
            _Scope.current = prev

    def predict(self, images, batch_size=32):
        x = torch.as_tensor(np.asarray(images)) if not isinstance(images, torch.Tensor) else images
        outs = []
        with torch.no_grad():
            for i in range(0, x.shape[0], batch_size):
                outs.append(self(x[i:i + batch_size].cuda(), training=False).cpu())
        return torch.cat(outs).numpy()

    @property
    def layers(self):
        return [self.scope.layers[n] for n in self.scope.order]

    def get_layer(self, name):
        return self.scope.layers[name]

    @property
    def trainable_weights(self):
        out = []
        for l in self.layers:
            out += list(l.trainable_weights)
        return out

    def _named_weights(self, name, layer):
        """Ordered {Keras weight name: tensor} of one regular layer of this model (None for antisymmetric layers)."""
        if isinstance(layer, (_RegularConv, _Dense)):
            if layer.kernel is None:
                raise RuntimeError("layer %r has no weights yet: call the model once before saving / loading" % name)
            return OrderedDict((name + '/' + v + ':0', getattr(layer, v)) for v in ('kernel', 'bias'))
        if isinstance(layer, _BatchNorm):
            if layer.gamma is None:
                raise RuntimeError("layer %r has no weights yet: call the model once before saving / loading" % name)
            return OrderedDict((name + '/' + v + ':0', getattr(layer, a)) for v, a in
                               (('gamma', 'gamma'), ('beta', 'beta'), ('moving_mean', 'moving_mean'), ('moving_variance', 'moving_var')))
        if not getattr(layer, "built", True):
            raise RuntimeError("layer %r has no weights yet: call the model once before saving / loading" % name)
        return None                                  # antisymmetric layers go through get_weights / set_weights

    def save_weights(self, filepath):
        """`model.save_weights(path + '.h5')` of the reference's notebooks (v6 cells 8 / 11): a Keras HDF5 weights file,
        one group per layer, the C+4 variables of every antisymmetric layer under the reference's names."""
        from ..keras_h5 import save_keras_weights
        out = OrderedDict()
        for lname in self.scope.order:
            layer = self.scope.layers[lname]
            named = self._named_weights(lname, layer)
            if named is None:
                named = OrderedDict((lname + '/' + v + ':0', w) for v, w in zip(layer._variable_names(), layer.get_weights()))
            else:
                named = OrderedDict((k, t.detach().cpu().numpy()) for k, t in named.items())
            out[lname] = named
        save_keras_weights(filepath, out)

    def load_weights(self, filepath):
        """`model.load_weights(path)`: every layer of this model must be in the file with the same weight shapes."""
        from ..keras_h5 import load_keras_weights
        saved = load_keras_weights(filepath)
        for lname in self.scope.order:
            layer = self.scope.layers[lname]
            if lname not in saved:
                raise ValueError("weights file has no layer %r" % lname)
            ws = saved[lname]
            named = self._named_weights(lname, layer)
            wanted = list(named) if named is not None else [lname + '/' + v + ':0' for v in layer._variable_names()]
            missing = [k for k in wanted if k not in ws]
            if missing:
                raise ValueError("weights file has no %r (layer %r holds %d weights in the file)" % (missing[0], lname, len(ws)))
            if named is None:
                layer.set_weights([ws[k] for k in wanted])
                continue
            with torch.no_grad():
                for k, t in named.items():
                    if tuple(ws[k].shape) != tuple(t.shape):
                        raise ValueError("weight %s has shape %s in the file, %s in the model" % (k, ws[k].shape, tuple(t.shape)))
                    t.copy_(torch.from_numpy(np.asarray(ws[k], dtype=np.float32)))


def get_single_block_resnet_build_function(kernel_type='antisymmetric',
                                           kernel_size=3,
                                           h=1.0,
                                           gamma=0.0,
                                           num_stages=5,
                                           blocks_per_stage=[3, 4, 6, 3],
                                           filters_per_block=[64, 128, 256, 512],
                                           strides=[(2, 2), (2, 2), (2, 2), (2, 2)],
                                           include_top=True,
                                           fc_activation='softmax',
                                           num_classes=None,
                                           use_batch_norm=False,
                                           use_max_pooling=[False, False, False, False],
                                           l2_regularization=0.0,
                                           subtract_mean=None,
                                           divide_by_stddev=None,
                                           verbose=False,
                                           precision='strict',
                                           seed=None):
    """Reference models/tfkeras_resnets.py:511-604 (same keywords; `precision`/`seed` are additions)."""
    if include_top and (num_classes is None):
        raise ValueError("You must pass a positive integer for `num_classes` if `include_top` is `True`.")
    antisymmetric = kernel_type == 'antisymmetric'
    name = 'single_block_resnet' + ('_antisymmetric' if antisymmetric else '_regular')
    strides = [tuple(s) for s in strides]

    def _forward(x):
        sc = _scope()
        x = x.to(torch.float32)
        if subtract_mean is not None:
            x = x - torch.as_tensor(np.array(subtract_mean), dtype=torch.float32, device=x.device)
        if divide_by_stddev is not None:
            x = x / torch.as_tensor(np.array(divide_by_stddev), dtype=torch.float32, device=x.device)
        x = sc.get('conv1', lambda: _RegularConv(filters_per_block[0], kernel_size, strides[0], 'conv1', sc.seed_for('conv1')))(x)
        if use_batch_norm:
            x = sc.get('bn_conv1', lambda: _BatchNorm('bn_conv1')).torch_apply(x, sc.training)
        x = torch.relu(x)
        for s in range(num_stages - 1):
            if use_max_pooling[s]:
                x = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous()
            ident = (s == 0 and not use_max_pooling[s]) or (s > 0 and not use_max_pooling[s] and
                                                             filters_per_block[s] == filters_per_block[s - 1] and strides[s] == (1, 1))
            if ident:
                for b in range(0, blocks_per_stage[s]):
                    x = single_layer_identity_block(x, kernel_size, antisymmetric, use_batch_norm, stage=s + 2, block=b, h=h, gamma=gamma)
            else:
                x = single_layer_conv_block(x, kernel_size, filters_per_block[s], strides[s], use_batch_norm, stage=s + 2, block=0)
                for b in range(1, blocks_per_stage[s]):
                    x = single_layer_identity_block(x, kernel_size, antisymmetric, use_batch_norm, stage=s + 2, block=b, h=h, gamma=gamma)
        if include_top:
            x = x.mean(dim=(1, 2))
            fc = sc.get('fc', lambda: _Dense(num_classes, fc_activation, sc.seed_for('fc')))
            x = fc(x)
        return x

    def _build_function(input_tensor):
        model = Model(_forward, name, precision, seed)
        if input_tensor is not None and isinstance(input_tensor, torch.Tensor) and input_tensor.is_cuda:
            with torch.no_grad():
                model(input_tensor)                      # instantiate the layers, like Keras graph building
        return model

    return _build_function


class _Dense(_KernelBiasWeights):
    def __init__(self, units, activation, seed=None):
        self.units, self.activation, self.kernel, self.seed = units, activation, None, seed
        self.name = 'fc'

    def __call__(self, x):
        if self.kernel is None:
            g = torch.Generator().manual_seed(self.seed) if self.seed is not None else None
            w = truncated_normal_(torch.empty(x.shape[-1], self.units), math.sqrt(2.0 / x.shape[-1]) / 0.87962566103423978, g)
            self.kernel = w.to(x.device).requires_grad_(True)
            self.bias = torch.zeros(self.units, device=x.device, requires_grad=True)
        y = x @ self.kernel + self.bias
        return torch.softmax(y, dim=-1) if self.activation == 'softmax' else y

    @property
    def trainable_weights(self):
        return [self.kernel, self.bias]


def build_single_block_resnet(image_shape, sample_input=None, **kwargs):
    """Reference models/tfkeras_resnets.py:427-509: build function applied to an Input of `image_shape`.
    Layers are created lazily on the first call unless a CUDA `sample_input` is given."""
    return get_single_block_resnet_build_function(**kwargs)(sample_input)


_RESNET_PRESETS = {'resnet50': ([3, 4, 6, 3], '50'), 'resnet101': ([3, 4, 23, 3], '101'), 'resnet152': ([3, 8, 36, 3], '152')}


def get_resnet_build_function(kernel_type='antisymmetric',
                              include_top=True,
                              fc_activation='softmax',
                              num_classes=None,
                              l2_regularization=0.0,
                              subtract_mean=None,
                              divide_by_stddev=None,
                              version=1,
                              preset=None,
                              blocks_per_stage=[3, 4, 6, 3],
                              filters_per_block=[[64, 64, 256],
                                                 [128, 128, 512],
                                                 [256, 256, 1024],
                                                 [512, 512, 2048]],
                              use_batch_norm=True,
                              precision='strict',
                              seed=None):
    """Five-stage bottleneck ResNet (reference models/tfkeras_resnets.py:698-818; same keywords, `precision` / `seed` are
    additions): 7x7/2 stem on a 3-pixel zero border, 3x3/2 max pooling on a 1-pixel border, four stages of one
    `bottleneck_conv_block` (strides (1,1), then (2,2)) and `blocks_per_stage[s] - 1` identity blocks, GAP + dense.  A
    `None` as the middle entry of a stage's filters makes that stage's 3x3 convolutions antisymmetric."""
    if include_top and (num_classes is None):
        raise ValueError("You must pass a positive integer for `num_classes` if `include_top` is `True`.")
    name = 'resnet'
    if preset is not None:
        if preset not in _RESNET_PRESETS:
            raise ValueError("`preset` must be either `None` or one of 'resnet50', 'resnet101', and 'resnet152', "
                             "but you passed `preset={}`.".format(preset))
        blocks_per_stage, tag = _RESNET_PRESETS[preset]
        filters_per_block = [[64, 64, 256], [128, 128, 512], [256, 256, 1024], [512, 512, 2048]]
        use_batch_norm = True
        name += tag
    antisymmetric = kernel_type == 'antisymmetric'
    name += '_antisymmetric' if antisymmetric else '_regular'

    def _forward(x):
        sc = _scope()
        x = x.to(torch.float32)
        if subtract_mean is not None:
            x = x - torch.as_tensor(np.array(subtract_mean), dtype=torch.float32, device=x.device)
        if divide_by_stddev is not None:
            x = x / torch.as_tensor(np.array(divide_by_stddev), dtype=torch.float32, device=x.device)
        x = torch.nn.functional.pad(x, (0, 0, 3, 3, 3, 3))                                   # ZeroPadding2D((3, 3)), NHWC
        x = sc.get('conv1', lambda: _RegularConv(64, 7, (2, 2), 'conv1', sc.seed_for('conv1'), padding='valid'))(x)
        if use_batch_norm:
            x = sc.get('bn_conv1', lambda: _BatchNorm('bn_conv1')).torch_apply(x, sc.training)
        x = torch.relu(x)
        x = torch.nn.functional.pad(x, (0, 0, 1, 1, 1, 1))                                   # zero border (values are >= 0)
        x = torch.nn.functional.max_pool2d(x.permute(0, 3, 1, 2), 3, 2).permute(0, 2, 3, 1).contiguous()
        for s in range(4):
            x = bottleneck_conv_block(x, 3, filters_per_block[s], antisymmetric, use_batch_norm, stage=s + 2, block=0,
                                      version=version, strides=(1, 1) if s == 0 else (2, 2))
            for i in range(1, blocks_per_stage[s]):
                x = bottleneck_identity_block(x, 3, filters_per_block[s], antisymmetric, use_batch_norm, stage=s + 2, block=i)
        if include_top:
            x = x.mean(dim=(1, 2))
            x = sc.get('fc', lambda: _Dense(num_classes, fc_activation, sc.seed_for('fc')))(x)
        return x

    def _build_function(input_tensor):
        model = Model(_forward, name, precision, seed)
        if input_tensor is not None and isinstance(input_tensor, torch.Tensor) and input_tensor.is_cuda:
            with torch.no_grad():
                model(input_tensor)
        return model

    return _build_function


def build_resnet(image_shape, sample_input=None, **kwargs):
    """Reference models/tfkeras_resnets.py:606-696: the build function applied to an Input of `image_shape`."""
    return get_resnet_build_function(**kwargs)(sample_input)
