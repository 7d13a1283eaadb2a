"""ctypes binding of libb200ode.so (C ABI declared in include/b200ode.h).

There is no CPU fallback: if the shared library is missing or no sm_100 device
is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200ode.so")

PREC_STRICT, PREC_FAST_TF32, PREC_FAST_BF16, PREC_SIMT_FP32, PREC_FAST_F16 = 0, 1, 2, 3, 4
PRECISIONS = {"strict": PREC_STRICT, "fast_tf32": PREC_FAST_TF32, "fast": PREC_FAST_TF32,
              "fast_bf16": PREC_FAST_BF16, "simt": PREC_SIMT_FP32}
# "fast_f16" exists for the persistent chains only (EulerNet); layers created with it use the FAST_TF32 per-layer kernels
# "strict": the 3xTF32 variant of the tf32 chain kernels (fp32-grade, <= 1e-5)
CHAIN_PRECISIONS = {"fast_tf32": PREC_FAST_TF32, "fast": PREC_FAST_TF32, "fast_f16": PREC_FAST_F16, "strict": PREC_STRICT}
LAYOUT_3BY3, LAYOUT_GENERAL = 0, 1
F_BIAS, F_RELU, F_SCALE, F_RESIDUAL = 1, 2, 4, 8
F_EULER = 15
COLSUM_PARTS = 2048

_lib = None

c_void_p, c_int, c_float, c_int64, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64, ctypes.c_size_t
GLUE_STEM_WGRAD, GLUE_TRANSITION_WGRAD, GLUE_HEAD = 0, 1, 2

_SIGNATURES = {
    "b200ode_version": (c_int, []),
    "b200ode_last_error": (ctypes.c_char_p, []),
    "b200ode_device_ok": (c_int, []),
    "b200ode_launch_count": (c_int64, []),
    "b200ode_comm_unique_id": (c_int, [c_void_p]),
    "b200ode_comm_init": (c_int, [c_int, c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    "b200ode_comm_allreduce_bucket": (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    "b200ode_comm_shared_alloc": (c_int, [c_void_p, c_size_t, ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p)]),
    "b200ode_comm_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                       c_void_p, c_void_p]),
    "b200ode_comm_destroy": (c_int, [c_void_p]),
    "b200ode_debug_set_trace": (c_int, [c_void_p]),
    "b200ode_debug_conv_plan": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(c_int)]),
    "b200ode_layer_create": (c_int, [c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_int,
                                     ctypes.POINTER(c_void_p)]),
    "b200ode_layer_destroy": (c_int, [c_void_p]),
    "b200ode_layer_num_params": (c_int64, [c_void_p]),
    "b200ode_layer_effective_mode": (c_int, [c_void_p]),
    "b200ode_pack_kernel": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200ode_euler_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                  c_int, c_void_p]),
    "b200ode_euler_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "b200ode_euler_dgrad_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int,
                                          c_int, c_void_p]),
    "b200ode_euler_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "b200ode_relu_scale_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_int, c_void_p]),
    "b200ode_euler_tail": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                   c_float, c_int, c_void_p]),
    "b200ode_colsum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "b200ode_euler_fwd_bn_stats": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_int), c_int, c_int, c_int, c_void_p]),
    "b200ode_bn_stats_finalize": (c_int, [c_void_p, c_int] + [c_void_p] * 10 + [c_int64, c_int, c_float, c_float, c_void_p]),
    "b200ode_bn_finalize": (c_int, [c_void_p] * 10 + [c_int64, c_int, c_float, c_float, c_void_p]),
    "b200ode_bn_bwd_reduce": (c_int, [c_void_p] * 9 + [c_int64, c_int, c_float, c_void_p]),
    "b200ode_bn_bwd_apply": (c_int, [c_void_p] * 10 + [c_int64, c_int, c_float, c_void_p]),
    "b200ode_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_float, c_float,
                                  c_float, c_float, c_float, c_void_p]),
    "b200ode_increment": (c_int, [c_void_p, c_void_p]),
    "b200ode_gradient_mean_norms": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "b200ode_stem_fwd": (c_int, [c_void_p, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p]),
    "b200ode_stem_wgrad": (c_int, [c_void_p, c_int, c_float, c_float, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                   c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200ode_transition_fwd": (c_int, [c_void_p] * 7 + [c_int] * 7 + [c_void_p]),
    "b200ode_transition_dgrad": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p]),
    "b200ode_transition_dgrad_amax": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p, c_void_p]),
    "b200ode_transition_wgrad": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p, c_size_t, c_void_p]),
    "b200ode_transition_fwd_fast": (c_int, [c_void_p] * 7 + [c_int] * 7 + [c_void_p]),
    "b200ode_transition_dgrad_fast": (c_int, [c_void_p] * 5 + [c_int] * 7 + [c_void_p, c_void_p]),
    "b200ode_transition_wgrad_fast": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p, c_size_t, c_void_p]),
    "b200ode_head_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "b200ode_head_fwd_bwd_amax": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "b200ode_layer_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "b200ode_layer_set_workspace": (c_int, [c_void_p, c_void_p, c_size_t]),
    "b200ode_chain_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, ctypes.POINTER(c_size_t)]),
    "b200ode_chain_set_workspace": (c_int, [c_void_p, c_void_p, c_size_t]),
    "b200ode_glue_workspace_bytes": (c_int, [c_int] * 8 + [ctypes.POINTER(c_size_t)]),
    "b200ode_chain_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "b200ode_chain_create": (c_int, [c_int, c_int, c_float, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "b200ode_chain_destroy": (c_int, [c_void_p]),
    "b200ode_chain_layer_params": (c_int64, [c_void_p]),
    "b200ode_chain_pack": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "b200ode_chain_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                  c_int, c_void_p]),
    "b200ode_chain_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                    c_void_p]),
    "b200ode_chain_dgrad_amax": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                         c_void_p, c_void_p]),
    "b200ode_chain_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                    c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class B200OdeError(RuntimeError):
    pass


def lib():
    """Load (once) and return the shared library; raise loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200OdeError(
                "native library %s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(differential_equations_resnet_b200/csrc/build.sh).  There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def last_error() -> str:
    return lib().b200ode_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == 0:
        return
    msg = last_error()
    if rc in (-1, -2):
        raise ValueError("b200ode: %s" % msg)
    raise B200OdeError("b200ode (status %d): %s" % (rc, msg))


def require_device():
    if not lib().b200ode_device_ok():
        raise B200OdeError("b200ode: %s" % last_error())


def launch_count() -> int:
    return int(lib().b200ode_launch_count())
