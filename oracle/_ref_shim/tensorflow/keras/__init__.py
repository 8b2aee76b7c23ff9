"""``tensorflow.keras`` of the shim: eager layers, ``Input``/``Model`` and the small helper
namespaces the reference imports.  TEST INFRASTRUCTURE ONLY (see ../../README.md)."""
from __future__ import annotations

import numpy as _np
import torch as _torch

from . import backend, initializers, layers, regularizers, utils  # noqa: F401
from .layers import Layer, _registry

# --- eager functional API ----------------------------------------------------------------------
_FEED = {}


def feed(values: dict):
    """Arrays that ``Input(name=...)`` hands out (the eager replacement for placeholders)."""
    _FEED.clear()
    _FEED.update(values)


class _InputTensor(_torch.Tensor):
    name = None


def Input(shape=None, batch_size=None, name=None, dtype=None, sparse=False, tensor=None,
          ragged=False, batch_shape=None, **kwargs):
    """``tf.keras.Input``: the reference declares every feature as a float32 ``[batch, L]`` input
    (DP:290-292, 303-304).  Eagerly: return the fed array for ``name`` as floatx -- ids
    included, exactly like the reference's float32 id inputs -- after checking the shape."""
    import tensorflow as tf
    if name not in _FEED:
        raise KeyError("shim: no array fed for Input(name=%r); call tensorflow.keras.feed first" % (name,))
    t = tf.convert_to_tensor(_np.asarray(_FEED[name]), dtype=dtype or "float32")
    want = tuple(batch_shape) if batch_shape is not None else (batch_size,) + tuple(shape)
    if len(want) != t.dim() or any(w is not None and w != s for w, s in zip(want, t.shape)):
        raise ValueError("Input %r: fed shape %s does not match declared %s" % (name, tuple(t.shape), want))
    t = t.as_subclass(_InputTensor)
    t.name = "%s:0" % name
    return t


class Model:
    """``tf.keras.Model(inputs, outputs)`` after an eager build: keeps the tensors and every
    layer constructed since the last ``reset_layers()``."""

    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs, self.name = inputs, outputs, name
        self.layers = list(_registry)
        self.losses = []

    def add_loss(self, loss):
        self.losses.append(loss)

    @property
    def weights(self):
        out = []
        for layer in self.layers:
            out.extend(layer._own_weights)
        return out

    trainable_weights = weights

    def summary(self):
        return "\n".join("%-40s %s" % (w.kon_name, tuple(w.shape)) for w in self.weights)


def reset_layers():
    del _registry[:]


def created_layers():
    return list(_registry)


class _Activations:
    @staticmethod
    def get(name):
        return layers._activation(name)


activations = _Activations()


class _Losses:
    @staticmethod
    def binary_crossentropy(y_true, y_pred, from_logits=False):
        """``tf.keras.losses.binary_crossentropy`` on probabilities: clip to [eps, 1-eps],
        ``-(y log(p+eps) + (1-y) log(1-p+eps))``, mean over the last axis."""
        eps = 1e-7
        p = _torch.clamp(y_pred, eps, 1 - eps)
        bce = y_true * _torch.log(p + eps) + (1 - y_true) * _torch.log(1 - p + eps)
        return (-bce).mean(dim=-1)


losses = _Losses()
