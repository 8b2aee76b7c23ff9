"""``tensorflow.keras.backend`` of the shim.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import torch as _torch


def ndim(x):
    return x.dim()


def dtype(x):
    return str(x.dtype).replace("torch.", "")


def int_shape(x):
    return tuple(int(s) for s in x.shape)


def floatx():
    import tensorflow as tf
    return "float64" if tf._floatx() is _torch.float64 else "float32"


def dot(x, y):
    """``K.dot``: for rank > 2 operands, ``x`` is flattened to ``[-1, x.shape[-1]]``, ``y`` is
    permuted so its second-to-last axis leads and flattened to ``[y.shape[-2], -1]``; one MatMul;
    the result is reshaped to ``x.shape[:-1] + y.shape[:-2] + y.shape[-1:]``."""
    if x.dim() > 2 or y.dim() > 2:
        xs, ys = list(x.shape), list(y.shape)
        yperm = list(range(y.dim()))
        yperm = [yperm.pop(-2)] + yperm
        xt = x.reshape(-1, xs[-1])
        yt = y.permute(*yperm).reshape(ys[-2], -1)
        return _torch.mm(xt, yt).reshape(xs[:-1] + ys[:-2] + ys[-1:])
    return _torch.matmul(x, y)


def batch_dot(x, y, axes=None):
    """``K.batch_dot`` (TF 2.1): contracts ``x`` axis ``a0`` with ``y`` axis ``a1`` per batch
    element; default axes ``[x.ndim-1, y.ndim-2]`` (``y.ndim-1`` when ``y`` is 2-D)."""
    xn, yn = x.dim(), y.dim()
    if xn < 2 or yn < 2:
        raise ValueError("Cannot do batch_dot on inputs with rank < 2.")
    if x.shape[0] != y.shape[0]:
        raise ValueError("Cannot do batch_dot on inputs with different batch sizes.")
    if isinstance(axes, int):
        axes = [axes, axes]
    if axes is None:
        axes = [xn - 1, yn - 1] if yn == 2 else [xn - 1, yn - 2]
    a0, a1 = [a if a >= 0 else a + n for a, n in zip(axes, (xn, yn))]
    if a0 == 0 or a1 == 0:
        raise ValueError("Cannot perform batch_dot over axis 0.")
    if x.shape[a0] != y.shape[a1]:
        raise ValueError("Cannot do batch_dot on inputs with shapes %s and %s with axes=%s."
                         % (tuple(x.shape), tuple(y.shape), axes))
    x_free = [i for i in range(1, xn) if i != a0]
    y_free = [i for i in range(1, yn) if i != a1]
    xm = x.permute(0, *x_free, a0).reshape(x.shape[0], -1, x.shape[a0])
    ym = y.permute(0, a1, *y_free).reshape(y.shape[0], y.shape[a1], -1)
    out = _torch.bmm(xm, ym).reshape([x.shape[0]] + [x.shape[i] for i in x_free] + [y.shape[i] for i in y_free])
    if out.dim() == 1:
        out = out.unsqueeze(1)
    return out


def softmax(x, axis=-1):
    return _torch.softmax(x, dim=axis)


def sigmoid(x):
    return _torch.sigmoid(x)


def relu(x):
    return _torch.relu(x)


def zeros_like(x, dtype=None, name=None):
    return _torch.zeros_like(x)


def expand_dims(x, axis=-1):
    return x.unsqueeze(axis)


def reverse(x, axes):
    return _torch.flip(x, [axes] if isinstance(axes, int) else list(axes))


def sum(x, axis=None, keepdims=False):  # noqa: A001
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def mean(x, axis=None, keepdims=False):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)
