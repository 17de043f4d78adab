"""B200-native (sm_100a) hot path of aerial-depth-completion behind the reference's own module API.

Drop-in for the reference's ``model_zoo`` / ``metrics`` / ``dataloaders.dense_to_sparse`` modules on the
NConv depth-completion path: same class names, constructor arguments, ``state_dict`` keys and error behaviour.
The compute is hand-written CUDA reached through the C ABI in ``include/adc_b200.h`` (``libadc_b200.so``);
there is no CPU fallback -- importing ``_lib`` without the built library, or calling an op on a non-CUDA
tensor, raises.
"""
__version__ = "0.1.0"
