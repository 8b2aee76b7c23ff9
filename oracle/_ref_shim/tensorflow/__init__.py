"""Minimal eager ``tensorflow`` stand-in backed by torch-CPU.  TEST INFRASTRUCTURE ONLY.

Exists so that the UNMODIFIED reference layer classes under /root/reference can be imported
and executed here (TensorFlow itself is not installable in this image); see README.md in this
directory for exactly which TF/Keras semantics are supplied by the shim.  Tensors are plain
``torch.Tensor`` (fp32, CPU); python lists of tensors are packed with ``stack`` wherever TF's
``convert_to_tensor`` would do so.
"""
from __future__ import annotations

import numpy as _np
import torch as _torch

__version__ = "2.1.0-shim"

float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool  # noqa: A001  (tf.bool)
Tensor = _torch.Tensor

_DTYPES = {"float": _torch.float32, "float32": _torch.float32, "float64": _torch.float64,
           "int32": _torch.int32, "int64": _torch.int64, "bool": _torch.bool}

# dtype every python float / numpy array is converted to (tf.keras.backend.floatx()); the
# fixture generator flips it to float64 to obtain "truth" runs of the same reference code.
_FLOATX = [_torch.float32]


def _floatx():
    return _FLOATX[0]


def _dtype(d):
    if d is None:
        return None
    if isinstance(d, str):
        return _floatx() if d in ("float", "float32", "floatx") else _DTYPES[d]
    return _floatx() if d is _torch.float32 else d


def convert_to_tensor(value, dtype=None):
    """``ops.convert_to_tensor``: tensors pass through, (nested) lists of tensors are packed
    along a new leading axis, python/numpy numbers become floatx / int32 tensors."""
    if isinstance(value, _torch.Tensor):
        return value if dtype is None else value.to(_dtype(dtype))
    if isinstance(value, (list, tuple)) and len(value) and any(
            isinstance(v, (_torch.Tensor, list, tuple)) for v in value):
        if all(isinstance(v, (list, tuple)) and not any(isinstance(u, _torch.Tensor) for u in v) for v in value):
            t = _torch.as_tensor(_np.asarray(value))
        else:
            t = _torch.stack([convert_to_tensor(v) for v in value], dim=0)
    else:
        t = _torch.as_tensor(_np.asarray(value))
    if dtype is not None:
        return t.to(_dtype(dtype))
    if t.dtype in (_torch.float64, _torch.float32, _torch.float16):
        t = t.to(_floatx())
    return t


constant = convert_to_tensor
_c = convert_to_tensor


# ---- elementwise ----------------------------------------------------------------------------
def multiply(x, y, name=None):
    return _c(x) * _c(y)


def add(x, y, name=None):
    return _c(x) + _c(y)


def subtract(x, y, name=None):
    return _c(x) - _c(y)


def equal(x, y, name=None):
    return _c(x) == _c(y)


def not_equal(x, y, name=None):
    return _c(x) != _c(y)


def cast(x, dtype, name=None):
    return _c(x).to(_dtype(dtype))


def sigmoid(x, name=None):
    return _torch.sigmoid(x)


def exp(x, name=None):
    return _torch.exp(x)


def sqrt(x, name=None):
    return _torch.sqrt(x)


def square(x, name=None):
    return x * x


def abs(x, name=None):  # noqa: A001
    return _torch.abs(x)


def zeros_like(x, dtype=None, name=None):
    return _torch.zeros_like(_c(x), dtype=_dtype(dtype))


def ones_like(x, dtype=None, name=None):
    return _torch.ones_like(_c(x), dtype=_dtype(dtype))


def zeros(shape, dtype=None, name=None):
    return _torch.zeros(tuple(shape), dtype=_dtype(dtype) or _floatx())


def ones(shape, dtype=None, name=None):
    return _torch.ones(tuple(shape), dtype=_dtype(dtype) or _floatx())


def where(cond, x=None, y=None, name=None):
    return _torch.where(cond, x, y)


def stop_gradient(x, name=None):
    return x.detach()


# ---- shapes -----------------------------------------------------------------------------------
def shape(x, name=None):
    return _torch.tensor(list(_c(x).shape), dtype=_torch.int32)


def reshape(tensor, shape, name=None):  # noqa: A002
    return _c(tensor).reshape([int(s) for s in shape])


def transpose(a, perm=None, name=None):
    a = _c(a)
    if perm is None:
        perm = list(range(a.dim()))[::-1]
    return a.permute(*[int(p) for p in perm])


def expand_dims(input, axis=None, name=None, dim=None):  # noqa: A002
    return _c(input).unsqueeze(axis if axis is not None else dim)


def squeeze(input, axis=None, name=None):  # noqa: A002
    t = _c(input)
    if axis is None:
        return t.squeeze()
    if isinstance(axis, (list, tuple)):
        for a in sorted([a % t.dim() for a in axis], reverse=True):
            if t.shape[a] != 1:
                raise ValueError("Can not squeeze dim[%d], expected a dimension of 1, got %d" % (a, t.shape[a]))
            t = t.squeeze(a)
        return t
    if t.shape[axis] != 1:
        raise ValueError("Can not squeeze dim[%d], expected a dimension of 1, got %d" % (axis, t.shape[axis]))
    return t.squeeze(axis)


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    value = _c(value)
    if isinstance(num_or_size_splits, int):
        if value.shape[axis] % num_or_size_splits:
            raise ValueError("Dimension size must be evenly divisible")
        return list(_torch.split(value, value.shape[axis] // num_or_size_splits, dim=axis))
    sizes = [int(s) for s in num_or_size_splits]
    if sum(sizes) != value.shape[axis]:
        raise ValueError("Sum of split sizes %d != dimension %d" % (sum(sizes), value.shape[axis]))
    return list(_torch.split(value, sizes, dim=axis))


def concat(values, axis, name=None):
    return _torch.cat([_c(v) for v in values], dim=axis)


def stack(values, axis=0, name=None):
    return _torch.stack([_c(v) for v in values], dim=axis)


def tile(input, multiples, name=None):  # noqa: A002
    return _c(input).repeat(*[int(m) for m in multiples])


def gather(params, indices, axis=0, name=None):
    return _torch.index_select(_c(params), axis, _c(indices).reshape(-1).long()).reshape(
        tuple(params.shape[:axis]) + tuple(indices.shape) + tuple(params.shape[axis + 1:]))


# ---- reductions ---------------------------------------------------------------------------------
def _axis(axis):
    return tuple(axis) if isinstance(axis, (list, tuple)) else axis


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None):
    t = _c(input_tensor)
    return t.sum() if axis is None else t.sum(dim=_axis(axis), keepdim=keepdims)


def reduce_mean(input_tensor, axis=None, keepdims=False, name=None):
    t = _c(input_tensor)
    return t.mean() if axis is None else t.mean(dim=_axis(axis), keepdim=keepdims)


def reduce_max(input_tensor, axis=None, keepdims=False, name=None):
    t = _c(input_tensor)
    return t.max() if axis is None else t.amax(dim=_axis(axis), keepdim=keepdims)


# ---- contractions -------------------------------------------------------------------------------
def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    """``tf.matmul``: (batched) matrix product on the two innermost axes; the reference hands
    in python lists of ``[B,m,1]`` slices (IL:311,316), which TF packs to ``[D,B,m,1]``."""
    a, b = _c(a), _c(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return _torch.matmul(a, b)


def tensordot(a, b, axes, name=None):
    """``tf.tensordot``: both operands are reshaped to matrices (free axes x contracted axes),
    multiplied with one MatMul and reshaped back."""
    a, b = _c(a), _c(b)
    if isinstance(axes, int):
        return _torch.tensordot(a, b, dims=axes)
    ax_a, ax_b = axes
    ax_a = [ax_a] if isinstance(ax_a, int) else list(ax_a)
    ax_b = [ax_b] if isinstance(ax_b, int) else list(ax_b)
    return _torch.tensordot(a, b, dims=(ax_a, ax_b))


class _Linalg:
    matmul = staticmethod(matmul)


linalg = _Linalg()


class _Math:
    multiply = staticmethod(multiply)
    add = staticmethod(add)
    sigmoid = staticmethod(sigmoid)
    reduce_sum = staticmethod(reduce_sum)
    reduce_mean = staticmethod(reduce_mean)
    equal = staticmethod(equal)
    not_equal = staticmethod(not_equal)


math = _Math()


class _NN:
    @staticmethod
    def sigmoid(x, name=None):
        return _torch.sigmoid(x)

    @staticmethod
    def relu(x, name=None):
        return _torch.relu(x)

    @staticmethod
    def softmax(x, axis=-1, name=None):
        return _torch.softmax(x, dim=axis)

    @staticmethod
    def embedding_lookup(params, ids, name=None):
        return params[ids.long()]


nn = _NN()


class _Random:
    @staticmethod
    def set_seed(seed):
        _torch.manual_seed(int(seed))


random = _Random()


class _Dataset:
    """``tf.data.Dataset.from_tensor_slices(...).shuffle(n).repeat(r).batch(b).prefetch(p)``
    (DP:335-337) as a lazy description; ``__iter__`` materialises it with numpy."""

    def __init__(self, data, ops=()):
        self._data, self._ops = data, tuple(ops)

    @classmethod
    def from_tensor_slices(cls, data):
        return cls(data)

    def _with(self, op):
        return _Dataset(self._data, self._ops + (op,))

    def shuffle(self, buffer_size, seed=None, reshuffle_each_iteration=None):
        return self._with(("shuffle", int(buffer_size), seed))

    def repeat(self, count=None):
        return self._with(("repeat", count))

    def batch(self, batch_size, drop_remainder=False):
        return self._with(("batch", batch_size, drop_remainder))

    def prefetch(self, buffer_size):
        return self._with(("prefetch", buffer_size))


class _Data:
    Dataset = _Dataset


data = _Data()

from . import keras  # noqa: E402,F401
