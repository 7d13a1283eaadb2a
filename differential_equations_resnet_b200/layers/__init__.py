from .tfkeras_layer_Conv2DAntisymmetric3By3 import Conv2DAntisymmetric3By3
from .tfkeras_layer_Conv2DAntisymmetric import Conv2DAntisymmetric
from .antisymmetric_conv2d_utils import get_centrosymmetric_matrix

__all__ = ["Conv2DAntisymmetric3By3", "Conv2DAntisymmetric", "get_centrosymmetric_matrix"]
