from .weight_utils import double_load_weights, load_pickled_weights, pickle_model_weights, unpack_dense_3by3

__all__ = ["double_load_weights", "load_pickled_weights", "pickle_model_weights", "unpack_dense_3by3"]
