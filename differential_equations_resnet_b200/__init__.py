"""B200-native antisymmetric-ResNet Euler blocks (drop-in for the hot path of
pierluigiferrari/differential_equations_resnet).  See DESIGN.md / INTEGRATION.md."""
from . import _abi
from .layers import Conv2DAntisymmetric, Conv2DAntisymmetric3By3, get_centrosymmetric_matrix

__all__ = ["Conv2DAntisymmetric", "Conv2DAntisymmetric3By3", "get_centrosymmetric_matrix", "_abi"]
__version__ = "0.1.0"
