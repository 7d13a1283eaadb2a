"""Drop-in for the reference's `layers.tfkeras_layer_Conv2DAntisymmetric.Conv2DAntisymmetric`
(reference file layers/tfkeras_layer_Conv2DAntisymmetric.py:26-270): general odd kernel size,
per-output-channel diagonal scalars, dependent kernels -E.W.E, `antisymmetric` flag that only
changes the diagonal blocks.  Variable order: for each output channel o the diagonal scalars
[1,1,1,1] in creation order, then input_kernels_for_output_kernel_{o} [k,k,C-o-1,1]; bias last.

k = 3, 5 and 7 with strides (1,1) and `antisymmetric=True` run forward and data gradient on the tcgen05
path (k*k taps through the same halo-strip kernel, halo pitch W + k//2; fp32-I/O precisions for k > 3, whose
weight gradient stays on the CUDA-core kernel); other strides, or `antisymmetric=False`, run on the CUDA-core
kernels (still GPU only).
"""
from __future__ import annotations

from .. import _abi
from ._base import AntisymmetricConvBase


def diag_slots(kernel_size, antisymmetric=True):
    """(i, j) positions of the free scalars of one diagonal block, in creation order
    (reference :231-264)."""
    k, slots = kernel_size, []
    for i in range(k):
        for j in range(i, k):
            if j > i or (j == i and i <= k // 2 - 1):
                slots.append((i, j))
            elif j == i and i == k // 2 and k % 2 == 1 and not antisymmetric:
                slots.append((i, j))
    return slots


class Conv2DAntisymmetric(AntisymmetricConvBase):
    _layout = _abi.LAYOUT_GENERAL

    def __init__(self,
                 kernel_size,
                 gamma=0.0,
                 strides=(1, 1),
                 use_bias=True,
                 kernel_initializer='he_normal',
                 kernel_regularizer=None,
                 antisymmetric=True,
                 **kwargs):
        super(Conv2DAntisymmetric, self).__init__(**kwargs)
        self.kernel_size = kernel_size
        self.gamma = gamma
        self.strides = tuple(strides)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel_regularizer = kernel_regularizer
        self.antisymmetric = antisymmetric

    def build(self, input_shape):
        if self.kernel_size % 2 == 0:
            raise ValueError("even kernel sizes have no centre tap under TF SAME padding; odd sizes only")
        self._build_common(input_shape, self.kernel_size, self.antisymmetric)

    def _variable_shapes(self):
        C, k = self.num_channels, self.kernel_size
        nd = len(diag_slots(k, self.antisymmetric))
        shapes = []
        for o in range(C):
            shapes += [(1, 1, 1, 1)] * nd
            if C - o - 1 > 0:
                shapes.append((k, k, C - o - 1, 1))
        if self.use_bias:
            shapes.append((C,))
        return shapes

    def _variable_names(self):
        """Variable names in creation order (reference :117-128, :231-264, :152).  Every output channel creates the
        scalars `centro_sym_{i}_{j}` again; the TF1 graph makes repeated names unique with `_1`, `_2`, ... ."""
        C = self.num_channels
        slots = diag_slots(self.kernel_size, self.antisymmetric)
        names = []
        for o in range(C):
            names += ['centro_sym_%d_%d' % ij + ('_%d' % o if o else '') for ij in slots]
            if C - o - 1 > 0:
                names.append('input_kernels_for_output_kernel_%d' % o)
        if self.use_bias:
            names.append('bias')
        return names

    def get_config(self):
        return {'name': self.name, 'trainable': self.trainable, 'dtype': self.dtype,
                'kernel_size': self.kernel_size, 'gamma': self.gamma, 'strides': self.strides,
                'use_bias': self.use_bias, 'kernel_initializer': self.kernel_initializer,
                'kernel_regularizer': self.kernel_regularizer, 'antisymmetric': self.antisymmetric}
