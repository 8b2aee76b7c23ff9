"""``tensorflow.keras.layers`` of the shim (eager, torch-CPU).  TEST INFRASTRUCTURE ONLY.

Only the Keras behaviour the reference's hot-path classes lean on is implemented; each class
says which TF-2.1 rule it follows.  Unknown layer names resolve to a stub that can be
constructed (the reference builds e.g. ``Dot``/``BatchNormalization`` objects it never calls)
but raises when called.
"""
from __future__ import annotations

import inspect

import torch as _torch

from . import initializers as _init

_registry = []          # every Layer constructed, in construction order (see keras.Model.layers)
_name_counts = {}


class TensorShape(tuple):
    """What ``build(input_shape)`` receives for one tensor: indexable, ``len`` = rank."""

    def as_list(self):
        return list(self)


def _shapes(x):
    if isinstance(x, _torch.Tensor):
        return TensorShape(int(s) for s in x.shape)
    if isinstance(x, (list, tuple)):
        return [_shapes(v) for v in x]
    return None


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for v in x:
            out.extend(_flatten(v))
        return out
    return [x]


def _masks_of(x):
    if isinstance(x, (list, tuple)):
        return [_masks_of(v) for v in x]
    return getattr(x, "_keras_mask", None)


def _activation(act):
    if act is None or act == "linear":
        return lambda x: x
    if callable(act):
        return act
    table = {"sigmoid": _torch.sigmoid, "relu": _torch.relu, "tanh": _torch.tanh,
             "softmax": lambda x: _torch.softmax(x, dim=-1)}
    if act not in table:
        raise ValueError("shim: activation %r not provided" % (act,))
    return table[act]


class Layer:
    """``tf.keras.layers.Layer``: ``__call__`` builds once from the input shapes, forwards
    ``mask=`` when ``call`` takes it and some input carries ``_keras_mask``, then attaches
    ``compute_mask``'s result to the outputs (base_layer.py ``__call__`` / ``_set_mask_metadata``)."""

    def __init__(self, trainable=True, name=None, dtype=None, dynamic=False, **kwargs):
        self._initial_weights = kwargs.pop("weights", None)
        kwargs.pop("input_shape", None)
        kwargs.pop("batch_input_shape", None)
        if kwargs:
            raise TypeError("Keyword argument not understood: %s" % sorted(kwargs))
        cls = type(self).__name__
        if name is None:
            n = _name_counts.get(cls, 0)
            _name_counts[cls] = n + 1
            name = cls.lower() if n == 0 else "%s_%d" % (cls.lower(), n)
        self.name = name
        self.trainable = trainable
        self.built = False
        self._own_weights = []
        if not hasattr(self, "supports_masking"):
            self.supports_masking = False
        _registry.append(self)

    # -- weights --
    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, regularizer=None,
                   trainable=None, constraint=None, **kwargs):
        init = _init.get(initializer if initializer is not None else "glorot_uniform")
        w = init(tuple(int(s) for s in shape)).clone().requires_grad_(True)
        w.kon_name = "%s/%s" % (self.name, name or "weight_%d" % len(self._own_weights))
        w.kon_regularizer = regularizer
        w.kon_layer = self
        w = _init._apply_hook(w)
        self._own_weights.append(w)
        return w

    @property
    def weights(self):
        return list(self._own_weights)

    trainable_weights = weights

    def get_weights(self):
        return [w.detach().numpy() for w in self._own_weights]

    def set_weights(self, values):
        if len(values) != len(self._own_weights):
            raise ValueError("set_weights: expected %d arrays, got %d" % (len(self._own_weights), len(values)))
        with _torch.no_grad():
            for w, v in zip(self._own_weights, values):
                v = _torch.as_tensor(v, dtype=w.dtype)
                if tuple(v.shape) != tuple(w.shape):
                    raise ValueError("Layer weight shape %s not compatible with provided weight shape %s"
                                     % (tuple(w.shape), tuple(v.shape)))
                w.copy_(v)

    # -- protocol --
    def build(self, input_shape):
        self.built = True

    def call(self, inputs, **kwargs):
        return inputs

    def compute_mask(self, inputs, mask=None):
        if not self.supports_masking:
            if any(m is not None for m in _flatten(mask)):
                raise TypeError("Layer %s does not support masking, but was passed an input_mask" % self.name)
            return None
        return mask

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(_shapes(inputs))
            self.built = True
            if self._initial_weights is not None:
                self.set_weights(self._initial_weights)
        in_masks = _masks_of(inputs)
        any_mask = any(m is not None for m in _flatten(in_masks))
        params = inspect.signature(self.call).parameters
        if "mask" in params and "mask" not in kwargs and not args and any_mask:
            kwargs["mask"] = in_masks
        outputs = self.call(inputs, *args, **kwargs)
        flat_out = [o for o in _flatten(outputs) if isinstance(o, _torch.Tensor)]
        if flat_out and not all(getattr(o, "_keras_mask", None) is not None for o in flat_out):
            out_masks = self.compute_mask(inputs, kwargs.get("mask", in_masks if any_mask else None))
            if out_masks is not None:
                for o, m in zip(flat_out, _flatten(out_masks)):
                    if m is not None:
                        try:
                            o._keras_mask = m
                        except AttributeError:
                            pass
        return outputs


# ----------------------------------------------------------------------------------------------
class Embedding(Layer):
    """embeddings.py: weight ``[input_dim, output_dim]`` (init ``'uniform'`` = U(-0.05, 0.05));
    ``call`` casts non-int32/int64 ids to int32 and gathers; ``compute_mask`` is ``ids != 0``
    when ``mask_zero``."""

    def __init__(self, input_dim, output_dim, embeddings_initializer="uniform",
                 embeddings_regularizer=None, activity_regularizer=None, embeddings_constraint=None,
                 mask_zero=False, input_length=None, **kwargs):
        super().__init__(**kwargs)
        self.input_dim, self.output_dim = int(input_dim), int(output_dim)
        self.embeddings_initializer = embeddings_initializer
        self.embeddings_regularizer = embeddings_regularizer
        self.mask_zero = mask_zero
        self.supports_masking = mask_zero
        self.input_length = input_length

    def build(self, input_shape):
        self.embeddings = self.add_weight(shape=(self.input_dim, self.output_dim),
                                          initializer=self.embeddings_initializer, name="embeddings",
                                          regularizer=self.embeddings_regularizer)
        self.built = True

    def compute_mask(self, inputs, mask=None):
        if not self.mask_zero:
            return None
        return inputs != 0

    def call(self, inputs):
        if inputs.dtype not in (_torch.int32, _torch.int64):
            inputs = inputs.to(_torch.int32)
        return self.embeddings[inputs.long()]


class _Merge(Layer):
    """merge.py ``_Merge``: ``build`` validates (>= 2 inputs, one batch size, trailing dims
    broadcastable -> else ``ValueError``) and notes whether ranks differ; ``call`` then expands
    lower-rank inputs at axis 1 before ``_merge_function``."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.supports_masking = True

    @staticmethod
    def _elemwise_shape(shape1, shape2):
        if len(shape1) < len(shape2):
            return _Merge._elemwise_shape(shape2, shape1)
        if not shape2:
            return shape1
        out = list(shape1[:-len(shape2)])
        for i, j in zip(shape1[-len(shape2):], shape2):
            if i == 1:
                out.append(j)
            elif j == 1:
                out.append(i)
            else:
                if i != j:
                    raise ValueError("Operands could not be broadcast together with shapes "
                                     + str(shape1) + " " + str(shape2))
                out.append(i)
        return tuple(out)

    def build(self, input_shape):
        if not isinstance(input_shape, (list, tuple)) or not isinstance(input_shape[0], tuple):
            raise ValueError("A merge layer should be called on a list of inputs.")
        if len(input_shape) < 2:
            raise ValueError("A merge layer should be called on a list of at least 2 inputs. "
                             "Got " + str(len(input_shape)) + " inputs.")
        if len({s[0] for s in input_shape if len(s)}) > 1:
            raise ValueError("Can not merge tensors with different batch sizes. Got tensors with shapes : "
                             + str(input_shape))
        out = tuple(input_shape[0][1:])
        for s in input_shape[1:]:
            out = self._elemwise_shape(out, tuple(s[1:]))
        self._reshape_required = len({len(s) for s in input_shape}) != 1
        self.built = True

    def call(self, inputs):
        if not isinstance(inputs, (list, tuple)):
            raise ValueError("A merge layer should be called on a list of inputs.")
        inputs = list(inputs)
        if self._reshape_required:
            nd = max(t.dim() for t in inputs)
            reshaped = []
            for t in inputs:
                for _ in range(nd - t.dim()):
                    t = t.unsqueeze(1)
                reshaped.append(t)
            inputs = reshaped
        try:
            return self._merge_function(inputs)
        except RuntimeError as e:
            # The reference always runs under the functional API (symbolic Keras Inputs), where an
            # elementwise op on statically incompatible shapes fails shape inference with ValueError
            # -- also on an already-built shared Add (DnnLayer re-uses one, CL:179,212-214).
            raise ValueError("Dimensions must be equal: %s" % e) from None

    def compute_mask(self, inputs, mask=None):
        if mask is None or all(m is None for m in _flatten(mask)):
            return None
        ms = [m.unsqueeze(0) for m in _flatten(mask) if m is not None]
        return _torch.cat(ms, 0).all(dim=0)


class Add(_Merge):
    """``output = inputs[0]; for i in 1..: output += inputs[i]`` -- strictly left to right."""

    def _merge_function(self, inputs):
        output = inputs[0]
        for i in range(1, len(inputs)):
            output = output + inputs[i]
        return output


class Multiply(_Merge):
    def _merge_function(self, inputs):
        output = inputs[0]
        for i in range(1, len(inputs)):
            output = output * inputs[i]
        return output


class Concatenate(_Merge):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def build(self, input_shape):
        if not isinstance(input_shape, (list, tuple)) or len(input_shape) < 2:
            raise ValueError("A `Concatenate` layer should be called on a list of at least 2 inputs")
        ranks = {len(s) for s in input_shape}
        if len(ranks) != 1:
            raise ValueError("A `Concatenate` layer requires inputs with matching shapes except for the "
                             "concat axis. Got inputs shapes: %s" % (input_shape,))
        ax = self.axis % len(input_shape[0])
        rest = {tuple(d for i, d in enumerate(s) if i != ax) for s in input_shape}
        if len(rest) != 1:
            raise ValueError("A `Concatenate` layer requires inputs with matching shapes except for the "
                             "concat axis. Got inputs shapes: %s" % (input_shape,))
        self.built = True

    def call(self, inputs):
        return _torch.cat(list(inputs), dim=self.axis)

    def compute_mask(self, inputs, mask=None):
        return None


class Flatten(Layer):
    """``[B, ...] -> [B, prod(...)]`` (row-major)."""

    def __init__(self, data_format=None, **kwargs):
        super().__init__(**kwargs)

    def call(self, inputs):
        return inputs.reshape(inputs.shape[0], -1)


class Dense(Layer):
    """core.py ``Dense``: kernel ``[last_dim, units]`` glorot-uniform, bias zeros; rank-2 inputs
    use MatMul, higher ranks contract the last axis (tensordot); ``activation(x@W + b)``."""

    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.units, self.use_bias = int(units), use_bias
        self.activation = _activation(activation)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer = kernel_regularizer
        self.supports_masking = True
        self.bias = None

    def build(self, input_shape):
        last = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", shape=[last, self.units], initializer=self.kernel_initializer,
                                      regularizer=self.kernel_regularizer)
        if self.use_bias:
            self.bias = self.add_weight("bias", shape=[self.units], initializer=self.bias_initializer)
        self.built = True

    def call(self, inputs):
        if inputs.dim() > 2:
            outputs = _torch.tensordot(inputs, self.kernel, dims=([inputs.dim() - 1], [0]))
        else:
            outputs = _torch.mm(inputs, self.kernel)
        if self.use_bias:
            outputs = outputs + self.bias
        return self.activation(outputs)


class Conv1D(Layer):
    """convolutional.py ``Conv1D`` (channels-last, ``valid``, stride 1): kernel
    ``[kernel_size, C_in, filters]`` glorot-uniform, bias zeros.  Kernel size 1 -- the only
    size the reference uses (IL:308) -- is a per-position ``x @ kernel[0] + bias`` (TF's CPU
    Conv2D also lowers 1x1/stride-1 convolutions to one matrix product)."""

    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format="channels_last",
                 dilation_rate=1, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", **kwargs):
        super().__init__(**kwargs)
        ks = kernel_size[0] if isinstance(kernel_size, (list, tuple)) else kernel_size
        if ks != 1 or strides != 1 or padding != "valid" or data_format != "channels_last":
            raise NotImplementedError("shim Conv1D: only kernel_size=1, stride 1, valid, channels_last")
        self.filters, self.use_bias = int(filters), use_bias
        self.activation = _activation(activation)
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer

    def build(self, input_shape):
        c_in = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", shape=[1, c_in, self.filters], initializer=self.kernel_initializer)
        self.bias = self.add_weight("bias", shape=[self.filters], initializer=self.bias_initializer) \
            if self.use_bias else None
        self.built = True

    def call(self, inputs):
        outputs = _torch.matmul(inputs, self.kernel[0])
        if self.use_bias:
            outputs = outputs + self.bias
        return self.activation(outputs)


class LayerNormalization(Layer):
    """normalization.py ``LayerNormalization()`` defaults: axis -1, epsilon 1e-3, gamma ones,
    beta zeros; ``nn.moments`` (biased variance) then ``nn.batch_normalization``:
    ``inv = rsqrt(var + eps) * gamma;  y = x * inv + (beta - mean * inv)``."""

    def __init__(self, axis=-1, epsilon=1e-3, center=True, scale=True, **kwargs):
        super().__init__(**kwargs)
        if axis != -1:
            raise NotImplementedError("shim LayerNormalization: axis=-1 only")
        self.epsilon, self.center, self.scale = epsilon, center, scale
        self.supports_masking = True

    def build(self, input_shape):
        n = int(input_shape[-1])
        self.gamma = self.add_weight("gamma", shape=[n], initializer="ones") if self.scale else None
        self.beta = self.add_weight("beta", shape=[n], initializer="zeros") if self.center else None
        self.built = True

    def call(self, inputs):
        mean = inputs.mean(dim=-1, keepdim=True)
        var = ((inputs - mean) ** 2).mean(dim=-1, keepdim=True)
        inv = _torch.rsqrt(var + self.epsilon)
        if self.scale:
            inv = inv * self.gamma
        off = -mean * inv
        if self.center:
            off = self.beta - mean * inv
        return inputs * inv + off


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super().__init__(**kwargs)
        self.activation = _activation(activation)
        self.supports_masking = True

    def call(self, inputs):
        return self.activation(inputs)


class ReLU(Layer):
    def __init__(self, max_value=None, negative_slope=0, threshold=0, **kwargs):
        super().__init__(**kwargs)
        if max_value is not None or negative_slope != 0 or threshold != 0:
            raise NotImplementedError("shim ReLU: plain max(x, 0) only")
        self.supports_masking = True

    def call(self, inputs):
        return _torch.relu(inputs)


class Softmax(Layer):
    def __init__(self, axis=-1, **kwargs):
        super().__init__(**kwargs)
        self.axis = axis

    def call(self, inputs):
        return _torch.softmax(inputs, dim=self.axis)


class Dot(Layer):
    """merge.py ``Dot(axes)`` on two inputs = ``K.batch_dot(x1, x2, axes)``."""

    def __init__(self, axes, normalize=False, **kwargs):
        super().__init__(**kwargs)
        self.axes = axes

    def call(self, inputs):
        from . import backend as K
        return K.batch_dot(inputs[0], inputs[1], self.axes)


class _Stub(Layer):
    """A layer the reference constructs somewhere but the hot path never calls."""

    def __init__(self, *args, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if k in ("name", "trainable", "dtype")}
        super().__init__(**kwargs)

    def call(self, inputs, **kwargs):
        raise NotImplementedError("shim: tf.keras.layers.%s is constructible but not executable "
                                  "(outside the hot path)" % type(self).__name__)


def __getattr__(name):
    if name.startswith("_"):
        raise AttributeError(name)
    cls = type(name, (_Stub,), {})
    globals()[name] = cls
    return cls
