"""``tensorflow.keras.utils`` of the shim.  TEST INFRASTRUCTURE ONLY."""
import numpy as _np


def to_categorical(y, num_classes=None, dtype="float32"):
    """One-hot rows (DP:359 turns the 0/1 click label into ``[1-y, y]``)."""
    y = _np.array(y, dtype="int64").ravel()
    n = int(num_classes or (y.max() + 1))
    out = _np.zeros((y.shape[0], n), dtype=dtype)
    out[_np.arange(y.shape[0]), y] = 1
    return out
