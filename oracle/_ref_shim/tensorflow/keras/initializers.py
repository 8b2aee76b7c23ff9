"""``tensorflow.keras.initializers`` of the shim.  TEST INFRASTRUCTURE ONLY.

Distributions and fan rules follow Keras (``glorot_uniform`` = U(+-sqrt(6/(fan_in+fan_out)));
``'uniform'`` = U(+-0.05); ``'random_normal'`` = N(0, 0.05)), but the random STREAM is numpy's,
one global ``RandomState`` advanced by every draw, with the initializer's ``seed`` folded in --
TensorFlow's stream is not reproducible here.  Fixtures store the drawn weights, so nothing
depends on the stream.  ``set_hook`` lets the fixture generator rescale/replace a weight as it
is created (e.g. N(0,1) stress tables)."""
from __future__ import annotations

import numpy as _np
import torch as _torch

_STATE = {"rng": _np.random.RandomState(2020), "hook": None}


def reseed(seed=2020):
    _STATE["rng"] = _np.random.RandomState(seed)


def set_hook(fn):
    """``fn(weight) -> weight | None``; ``weight.kon_name`` / ``.kon_layer`` identify it."""
    _STATE["hook"] = fn


def _apply_hook(w):
    fn = _STATE["hook"]
    if fn is None:
        return w
    r = fn(w)
    if r is None or r is w:
        return w
    r = r.detach().to(w.dtype).clone().requires_grad_(True)
    r.kon_name, r.kon_regularizer, r.kon_layer = w.kon_name, w.kon_regularizer, w.kon_layer
    return r


def _floatx():
    import tensorflow as tf
    return tf._floatx()


def _fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for d in shape[:-2]:
        rf *= d
    return shape[-2] * rf, shape[-1] * rf


class Initializer:
    seed = None

    def _rng(self):
        rng = _STATE["rng"]
        if self.seed is not None:
            rng.randint(0, 2 ** 31 - 1)      # advance, so equal seeds still give distinct draws
        return rng

    def __call__(self, shape, dtype=None):
        raise NotImplementedError


class GlorotUniform(Initializer):
    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(tuple(shape))
        lim = (6.0 / (fi + fo)) ** 0.5
        return _torch.as_tensor(self._rng().uniform(-lim, lim, size=tuple(shape))).to(_floatx())


class GlorotNormal(Initializer):
    def __init__(self, seed=None):
        self.seed = seed

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(tuple(shape))
        std = (2.0 / (fi + fo)) ** 0.5
        return _torch.as_tensor(self._rng().normal(0, std, size=tuple(shape))).to(_floatx())


class RandomUniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval, self.seed = minval, maxval, seed

    def __call__(self, shape, dtype=None):
        return _torch.as_tensor(self._rng().uniform(self.minval, self.maxval, size=tuple(shape))).to(_floatx())


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed

    def __call__(self, shape, dtype=None):
        return _torch.as_tensor(self._rng().normal(self.mean, self.stddev, size=tuple(shape))).to(_floatx())


class Zeros(Initializer):
    def __call__(self, shape, dtype=None):
        return _torch.zeros(tuple(shape), dtype=_floatx())


class Ones(Initializer):
    def __call__(self, shape, dtype=None):
        return _torch.ones(tuple(shape), dtype=_floatx())


class Constant(Initializer):
    def __init__(self, value=0):
        self.value = value

    def __call__(self, shape, dtype=None):
        return _torch.full(tuple(shape), float(self.value), dtype=_floatx())


glorot_uniform = GlorotUniform
glorot_normal = GlorotNormal
zeros = Zeros
ones = Ones
constant = Constant
random_uniform = RandomUniform
random_normal = RandomNormal

_BY_NAME = {"glorot_uniform": GlorotUniform, "glorot_normal": GlorotNormal, "zeros": Zeros, "ones": Ones,
            "uniform": RandomUniform, "random_uniform": RandomUniform, "random_normal": RandomNormal,
            "normal": RandomNormal}


def get(identifier):
    if identifier is None:
        return GlorotUniform()
    if isinstance(identifier, str):
        return _BY_NAME[identifier]()
    if isinstance(identifier, type):
        return identifier()
    if callable(identifier):
        return identifier
    raise ValueError("Could not interpret initializer identifier: %r" % (identifier,))
