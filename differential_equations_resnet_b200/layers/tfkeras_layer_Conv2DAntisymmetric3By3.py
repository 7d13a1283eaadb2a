"""Drop-in for the reference's `layers.tfkeras_layer_Conv2DAntisymmetric3By3.Conv2DAntisymmetric3By3`
(reference file layers/tfkeras_layer_Conv2DAntisymmetric3By3.py:26-293): same constructor
keywords and defaults, same `build` / `call` / `compute_output_shape` / `get_config` /
`get_kernel` / `get_bias` surface, same variable order for `get_weights` / `set_weights`.

The 3x3xCxC kernel obeys K[a,b,ci,o] = -K[2-a,2-b,o,ci] off the diagonal and
[[a,b,c],[d,gamma,-d],[-c,-b,-a]] on it, so the convolution matrix A satisfies
A + A^T = 2*gamma*I bit-exactly.  Unlike the reference, the kernel is never
materialised by O(C^2) slice/concat ops: one CUDA kernel (K1 antisym_pack) writes the
tensor-core operand layouts straight from the packed free parameters.
"""
from __future__ import annotations

from .. import _abi
from ._base import AntisymmetricConvBase


class Conv2DAntisymmetric3By3(AntisymmetricConvBase):
    _layout = _abi.LAYOUT_3BY3

    def __init__(self,
                 gamma=0.0,
                 strides=(1, 1),
                 use_bias=True,
                 kernel_initializer='he_normal',
                 kernel_regularizer=None,
                 **kwargs):
        super(Conv2DAntisymmetric3By3, self).__init__(**kwargs)
        self.gamma = gamma
        self.strides = tuple(strides)
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.kernel_regularizer = kernel_regularizer   # kept for API parity; the reference trainer ignores it

    def build(self, input_shape):
        self._build_common(input_shape, 3, True)

    def _variable_shapes(self):
        # a, b, c, d [1,1,1,C]; input_kernels_for_output_kernel_{o} [3,3,C-o-1]; bias [C]
        C = self.num_channels
        shapes = [(1, 1, 1, C)] * 4 + [(3, 3, C - o - 1) for o in range(C - 1)]
        if self.use_bias:
            shapes.append((C,))
        return shapes

    def _variable_names(self):
        # creation order, reference :119-124, :148-153, :219-245
        names = ['a', 'b', 'c', 'd'] + ['input_kernels_for_output_kernel_%d' % o for o in range(self.num_channels - 1)]
        if self.use_bias:
            names.append('bias')
        return names

    def get_config(self):
        # superset of the reference's config (it omits gamma, reference :177-186)
        return {'name': self.name, 'trainable': self.trainable, 'dtype': self.dtype,
                'gamma': self.gamma, 'strides': self.strides, 'use_bias': self.use_bias,
                'kernel_initializer': self.kernel_initializer, 'kernel_regularizer': self.kernel_regularizer}
