from .tfkeras_resnets import (Model, build_single_block_resnet, get_single_block_resnet_build_function,
                              single_layer_conv_block, single_layer_identity_block)

__all__ = ["Model", "build_single_block_resnet", "get_single_block_resnet_build_function",
           "single_layer_conv_block", "single_layer_identity_block"]
