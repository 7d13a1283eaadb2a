"""Free-function twin of the diagonal-block builder (reference
layers/antisymmetric_conv2d_utils.py:23-75, imported by nothing there): returns the
(anti-)centrosymmetric `size x size` block as a tensor of rank `rank`, built from freshly
initialised free scalars; the centre of an anti-centrosymmetric block is a constant zero.
"""
from __future__ import annotations

import math

import torch

from ._base import truncated_normal_


def get_centrosymmetric_matrix(size, in_channels, rank=4, anti=True, regularizer=None, trainable=True, prefix=None,
                               generator=None, device=None):
    stddev = math.sqrt(2.0 / (size * size * in_channels))
    m = torch.zeros(size, size)
    for i in range(size):
        for j in range(i, size):
            if j > i or (j == i and i <= size // 2 - 1):
                v = truncated_normal_(torch.empty(1), stddev, generator)[0]
                m[i, j] = v
                m[size - 1 - i, size - 1 - j] = -v if anti else v
            elif j == i and i == size // 2 and size % 2 == 1:
                m[i, j] = 0.0 if anti else truncated_normal_(torch.empty(1), stddev, generator)[0]
    out = m.reshape([size, size] + [1] * (rank - 2))
    return out.to(device) if device is not None else out
