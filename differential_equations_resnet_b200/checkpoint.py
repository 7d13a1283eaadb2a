"""Checkpoint interchange with the reference's variable layout.  Host-side only: reshapes and
copies of the flat parameter bucket, no arithmetic.

The reference stores every antisymmetric layer as C+4 TF variables in creation order
(`layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:119-124, 148-153, 219-245`):
    <layer>/a, /b, /c, /d                         each [1,1,1,C]
    <layer>/input_kernels_for_output_kernel_{o}   [3,3,C-o-1]   for o = 0 .. C-2
    <layer>/bias                                  [C]
`tf.train.Saver` (`training/training.py:848-865`) writes them under exactly these names (+ ':0' in the graph);
regular layers are `<layer>/kernel` (HWIO) and `<layer>/bias`.  `model_utils/weight_utils.py:23-39` pickles
a list of {'kernel', 'bias'} per weighted layer, and `double_load_weights` (:41-80) copies every block of an
l-block net into two consecutive blocks of a 2l-block net.
"""
import pickle
from collections import OrderedDict

import numpy as np
import torch

ANTISYM_VARIABLES = ("a", "b", "c", "d")


def variable_shapes_3by3(C):
    """(name, shape) of the C+4 variables of one Conv2DAntisymmetric3By3 layer, creation order."""
    out = [(v, (1, 1, 1, C)) for v in ANTISYM_VARIABLES]
    out += [("input_kernels_for_output_kernel_%d" % o, (3, 3, C - o - 1)) for o in range(C - 1)]
    out.append(("bias", (C,)))
    return out


def split_packed_3by3(flat, C):
    """Flat packed parameter vector of one layer -> OrderedDict variable name -> ndarray (reference shapes)."""
    flat = np.asarray(flat, dtype=np.float32).reshape(-1)
    out, cur = OrderedDict(), 0
    for name, shp in variable_shapes_3by3(C):
        n = int(np.prod(shp))
        out[name] = flat[cur:cur + n].reshape(shp).copy()
        cur += n
    if cur != flat.size:
        raise ValueError("packed vector has %d scalars, a %d-channel layer has %d" % (flat.size, C, cur))
    return out


def join_packed_3by3(variables, C):
    """Inverse of split_packed_3by3; accepts names with or without a ':0' suffix."""
    parts = []
    for name, shp in variable_shapes_3by3(C):
        v = variables[name] if name in variables else variables[name + ":0"]
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(shp):
            raise ValueError("variable %s has shape %s, expected %s" % (name, tuple(v.shape), tuple(shp)))
        parts.append(v.reshape(-1))
    return np.concatenate(parts)


def export_reference_variables(net):
    """EulerNet -> OrderedDict '<layer>/<variable>' -> ndarray, graph order, reference names and shapes."""
    out = OrderedDict()
    packed = net.export_params()
    euler = {name: (n, C) for name, _, n, C in net.layer_param_slices()}
    for op in net.spec.plan():
        kind, name = op[0], op[4]
        if kind == "euler":
            n, C = euler[name]
            for v, arr in split_packed_3by3(packed[name + "/packed"].numpy(), C).items():
                out[name + "/" + v] = arr
        elif kind == "stem":
            out[name + "/kernel"] = packed[name + "/kernel"].numpy().copy()
            out[name + "/bias"] = packed[name + "/bias"].numpy().copy()
        elif kind == "maxpool":
            continue
        else:   # transition block: main 3x3 branch '...branch2', 1x1 shortcut '...branch1'
            for br in ("2", "1"):
                out[name + br + "/kernel"] = packed[name + br + "/kernel"].numpy().copy()
                out[name + br + "/bias"] = packed[name + br + "/bias"].numpy().copy()
    out["fc/kernel"] = packed["fc/kernel"].numpy().copy()
    out["fc/bias"] = packed["fc/bias"].numpy().copy()
    for k, v in packed.items():     # BatchNormalization layers: the Keras variable names (gamma, beta, moving_mean, moving_variance)
        if k.rsplit("/", 1)[-1] in ("gamma", "beta", "moving_mean", "moving_variance"):
            out[k] = v.numpy().copy()
    return out


def import_reference_variables(net, variables):
    """Load a dict produced by export_reference_variables (or read from a reference checkpoint with the same names)."""
    def get(key):
        if key in variables:
            return variables[key]
        if key + ":0" in variables:
            return variables[key + ":0"]
        raise KeyError("checkpoint has no variable %r" % key)

    params = {}
    for name, _, n, C in net.layer_param_slices():
        layer_vars = {v: get(name + "/" + v) for v, _ in variable_shapes_3by3(C)}
        params[name + "/packed"] = torch.from_numpy(join_packed_3by3(layer_vars, C))
    for name, (_, shape) in net.torch_params.items():
        arr = np.asarray(get(name), dtype=np.float32)
        if tuple(arr.shape) != tuple(shape):
            raise ValueError("variable %s has shape %s, expected %s" % (name, tuple(arr.shape), tuple(shape)))
        params[name] = torch.from_numpy(arr.copy())
    for name, _, C, _, _ in net.bn_param_slices():      # Euler-step BatchNorm layers (use_batch_norm=True)
        for v in ("gamma", "beta"):
            params[name + "/" + v] = torch.from_numpy(np.asarray(get(name + "/" + v), dtype=np.float32).copy())
    for key in variables:
        k = key[:-2] if key.endswith(":0") else key
        if k.rsplit("/", 1)[-1] in ("moving_mean", "moving_variance"):
            params[k] = torch.from_numpy(np.asarray(variables[key], dtype=np.float32).copy())
    net.import_params(params)


def save_variables(net, path):
    """Counterpart of Training.save(saver='train_saver') (`training/training.py:848-865`): one .npz with the
    reference's variable names, plus the Adam state and the step counter so training resumes bit-identically."""
    arrays = {k.replace("/", "__"): v for k, v in export_reference_variables(net).items()}
    arrays["__adam_m"] = net.adam_m.detach().cpu().numpy()
    arrays["__adam_v"] = net.adam_v.detach().cpu().numpy()
    arrays["__global_step"] = net.step_counter.detach().cpu().numpy()
    np.savez(path, **arrays)


def load_variables(net, path, restore_optimizer=True):
    """Counterpart of Training.load_variables (`training/training.py:867-872`)."""
    with np.load(path) as z:
        variables = {k.replace("__", "/"): z[k] for k in z.files if not k.startswith("__")}
        import_reference_variables(net, variables)
        if restore_optimizer and "__adam_m" in z.files:
            with torch.no_grad():
                net.adam_m.copy_(torch.from_numpy(z["__adam_m"]))
                net.adam_v.copy_(torch.from_numpy(z["__adam_v"]))
                net.step_counter.copy_(torch.from_numpy(z["__global_step"]))


def dense_layer_weights(net):
    """List of {'kernel', 'bias'} per weighted layer in graph order -- the structure `pickle_model_weights`
    (`model_utils/weight_utils.py:23-39`) writes -- with every antisymmetric layer's kernel ASSEMBLED to its
    dense [3,3,C,C] form on the GPU (K1 pack kernel).  This is what the reference loads into a regular ResNet
    of the same shape (experiments v7, 'Antisymmetric 16 Weights Loaded into Regular 16 Model')."""
    from .layers._base import LayerHandle
    from . import _abi
    packed = net.export_params()
    handles = {}
    out = []
    for op in net.spec.plan():
        kind, name = op[0], op[4]
        if kind == "euler":
            C = op[2]
            if C not in handles:
                handles[C] = LayerHandle(C, 3, net.spec.gamma, (1, 1), True, True, _abi.PRECISIONS["simt"], _abi.LAYOUT_3BY3)
            flat = packed[name + "/packed"].to(net.device)
            K = torch.empty((3, 3, C, C), device=net.device)
            handles[C].pack(flat, K, force=True)
            out.append({"kernel": K.cpu().numpy(), "bias": flat[-C:].cpu().numpy()})
        elif kind == "stem":
            out.append({"kernel": packed[name + "/kernel"].numpy().copy(), "bias": packed[name + "/bias"].numpy().copy()})
        elif kind == "maxpool":
            continue
        else:
            for br in ("2", "1"):
                out.append({"kernel": packed[name + br + "/kernel"].numpy().copy(),
                            "bias": packed[name + br + "/bias"].numpy().copy()})
    out.append({"kernel": packed["fc/kernel"].numpy().copy(), "bias": packed["fc/bias"].numpy().copy()})
    return out


def pickle_model_weights(net, save_filename):
    """`model_utils/weight_utils.py:23-39` for an EulerNet (dense kernels, see dense_layer_weights)."""
    with open(save_filename, "wb") as f:
        pickle.dump(dense_layer_weights(net), f, protocol=pickle.HIGHEST_PROTOCOL)


def double_load_variables(net, saved_variables):
    """`double_load_weights` (`model_utils/weight_utils.py:41-80`) on the reference variable layout: a single-block
    net with l Euler blocks per stage is loaded into one with 2l blocks per stage, block b going to blocks 2b and
    2b+1 (stem, transition blocks and the dense layer are copied once).  `saved_variables` is a dict as produced by
    export_reference_variables of the smaller net."""
    def get(key):
        return saved_variables[key] if key in saved_variables else saved_variables[key + ":0"]

    def has(key):
        return key in saved_variables or key + ":0" in saved_variables

    params = {}
    for name, _, n, C in net.layer_param_slices():
        stage, block = name[len("res"):].split("_")[:2]
        first = 0 if has("res%s_0_branch2/a" % stage) else 1   # stages behind a transition block start at block 1
        src = "res%s_%d_branch2" % (stage, first + (int(block) - first) // 2)
        layer_vars = {v: get(src + "/" + v) for v, _ in variable_shapes_3by3(C)}
        params[name + "/packed"] = torch.from_numpy(join_packed_3by3(layer_vars, C))
    for name, (_, shape) in net.torch_params.items():
        params[name] = torch.from_numpy(np.asarray(get(name), dtype=np.float32).copy())
    net.import_params(params)
