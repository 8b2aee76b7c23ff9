"""B200-native drop-in for the CTR hot path of TIXhjq/ML_Function (``kon.model.ctr_model``).

Host side (this package) mirrors the reference's Keras layer / model-builder interface on
torch tensors; every FLOP of the hot path runs in ``libkon_b200.so`` (hand-written sm_100a
CUDA behind the C-ABI of ``include/kon_b200.h``).  No CPU fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops", "layers", "models", "data_prepare", "dist", "optim"]
