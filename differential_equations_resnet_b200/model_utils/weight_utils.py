"""`model_utils/weight_utils.py` of the reference for the models of this package: same function
names and arguments, on `models.Model` objects (layers with `get_weights()` / `set_weights()`).  Host-side only.

The reference unpacks `kernel, bias = layer.get_weights()` (:35, :59-77), i.e. it handles layers with exactly two weight
arrays.  An antisymmetric layer holds C+4 variables, so here its DENSE assembled kernel [3,3,C,C] (K1 pack kernel on the
GPU, `layer.get_kernel()`) and its bias are pickled -- the form a regular ResNet of the same shape loads ("Antisymmetric 16
Weights Loaded into Regular 16 Model", experiments_antisymmetric_resnet_v7.ipynb) -- and a dense kernel is loaded INTO an
antisymmetric layer only if it satisfies K[a,b,ci,o] = -K[2-a,2-b,o,ci] (+ 2*gamma on the centre diagonal) exactly, by
reading the free parameters back out of it (`unpack_dense_3by3`)."""
import pickle

import numpy as np


def _is_antisymmetric_layer(layer):
    return hasattr(layer, "get_kernel") and hasattr(layer, "_variable_shapes")


def _weighted_layers(model):
    """Layers with weights, in graph order (reference :32-34, :52-56: `len(layer.get_weights()) > 0`)."""
    out = []
    for layer in model.layers:
        if not hasattr(layer, "get_weights"):
            continue
        if _is_antisymmetric_layer(layer) and not layer.built:
            raise RuntimeError("layer %r is not built: call the model once first" % layer.name)
        if len(layer.get_weights()) > 0:
            out.append(layer)
    return out


def unpack_dense_3by3(kernel, gamma=0.0, bias=None):
    """Dense [3,3,C,C] kernel -> list of the C+4 (or C+3 without bias) variables of a Conv2DAntisymmetric3By3 layer in
    creation order (`layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:119-124, 219-245`): the inverse of the assembly
    (:113-141).  Raises ValueError unless the kernel is anti-centrosymmetric with centre gamma bit-exactly."""
    K = np.asarray(kernel, dtype=np.float32)
    if K.ndim != 4 or K.shape[:2] != (3, 3) or K.shape[2] != K.shape[3]:
        raise ValueError("expected a [3,3,C,C] kernel, got %s" % (K.shape,))
    C = K.shape[2]
    S = K + K[::-1, ::-1].transpose(0, 1, 3, 2)
    expect = np.zeros_like(K)
    expect[1, 1, np.arange(C), np.arange(C)] = np.float32(2.0) * np.float32(gamma)
    if not np.array_equal(S, expect):
        raise ValueError("kernel is not anti-centrosymmetric with centre gamma=%g (max deviation %.3g): it cannot be "
                         "loaded into an antisymmetric layer" % (gamma, float(np.abs(S - expect).max())))
    idx = np.arange(C)
    out = [K[0, 0, idx, idx].reshape(1, 1, 1, C).copy(), K[0, 1, idx, idx].reshape(1, 1, 1, C).copy(),
           K[0, 2, idx, idx].reshape(1, 1, 1, C).copy(), K[1, 0, idx, idx].reshape(1, 1, 1, C).copy()]
    out += [K[:, :, o + 1:, o].copy() for o in range(C - 1)]
    if bias is not None:
        out.append(np.asarray(bias, dtype=np.float32).reshape(C).copy())
    return out


def _get_kernel_bias(layer):
    if _is_antisymmetric_layer(layer):
        return layer.get_kernel(), layer.get_bias()
    ws = layer.get_weights()
    if len(ws) != 2:
        raise ValueError("layer %r has %d weight arrays; like the reference, only (kernel, bias) layers are handled"
                         % (getattr(layer, "name", layer), len(ws)))
    return ws[0], ws[1]


def _set_kernel_bias(layer, kernel, bias):
    if _is_antisymmetric_layer(layer):
        layer.set_weights(unpack_dense_3by3(kernel, layer.gamma, bias if layer.use_bias else None))
    else:
        layer.set_weights([kernel, bias])


def pickle_model_weights(model, save_filename):
    """Reference :23-39: a pickled list with one {'kernel', 'bias'} dict per layer that has weights."""
    weights = []
    for layer in _weighted_layers(model):
        kernel, bias = _get_kernel_bias(layer)
        weights.append({'kernel': np.asarray(kernel), 'bias': np.asarray(bias)})
    with open(save_filename, 'wb') as f:
        pickle.dump(weights, f, protocol=pickle.HIGHEST_PROTOCOL)


def load_pickled_weights(model, weights_pickle_file):
    """Addition (the reference's notebook cell for it is empty): load a `pickle_model_weights` file layer by layer into
    a model of the same shape -- e.g. the dense kernels of an antisymmetric net into the regular net of the same
    architecture (experiments_antisymmetric_resnet_v7.ipynb, 'Antisymmetric 16 Weights Loaded into Regular 16 Model')."""
    with open(weights_pickle_file, 'rb') as f:
        saved = pickle.load(f)
    new = _weighted_layers(model)
    if len(new) != len(saved):
        raise ValueError("the model has %d weighted layers, the file %d" % (len(new), len(saved)))
    for layer, w in zip(new, saved):
        _set_kernel_bias(layer, w['kernel'], w['bias'])


def double_load_weights(model, weights_pickle_file):
    """Reference :41-80: the l weighted blocks of a saved (l+2)-layer single-block ResNet are each loaded into two
    consecutive layers of a (2l+2)-layer model; first convolution and final dense layer once."""
    with open(weights_pickle_file, 'rb') as f:
        saved = pickle.load(f)
    new = _weighted_layers(model)
    if len(new) != 2 * (len(saved) - 2) + 2:
        raise ValueError("the model has %d weighted layers, a double load of %d saved layers needs %d"
                         % (len(new), len(saved), 2 * (len(saved) - 2) + 2))
    _set_kernel_bias(new[0], saved[0]['kernel'], saved[0]['bias'])
    for l in range(1, len(saved) - 1):
        _set_kernel_bias(new[2 * (l - 1) + 1], saved[l]['kernel'], saved[l]['bias'])
        _set_kernel_bias(new[2 * l], saved[l]['kernel'], saved[l]['bias'])
    _set_kernel_bias(new[-1], saved[-1]['kernel'], saved[-1]['bias'])
