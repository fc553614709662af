"""A NumPy stand-in for exactly the TensorFlow symbols the reference's model code touches  (TEST INFRASTRUCTURE ONLY).

TensorFlow cannot be installed in the build container, so the reference cannot run as shipped.  Its *own source files*
(positionalencoding.py, Attention.py, point_ffn.py, Encoder.py, Decoder.py, Transformer.py and, from train.py, the
CustomSchedule and Train.loss_function bodies) can, however, be imported and executed unmodified on top of this module
installed as ``tensorflow``: every function below implements the documented semantics of the TF / Keras primitive of the
same name in float32 NumPy (SURVEY.md Appendix A lists them).  What such a run pins is everything the reference itself
wrote -- the graph wiring: which mask goes where, residual order, sqrt(d_model) embedding scale, positional-encoding
slice, head split / merge, the -1e9 additive mask, loss masking and the 1/batch_size normalisation, the schedule formula
-- and ``oracle/make_golden.py`` stores its outputs as fixtures for ``tests/test_oracle.py``.  What it cannot pin is
TensorFlow's own kernels (their numerics are restated here from the documentation).

Only inference-mode calls are supported (``Dropout`` is the identity unless ``training`` is truthy, in which case it
raises: TF's random stream is not reproducible here).
"""
from __future__ import annotations

import sys
import types

import numpy as np

F32 = np.float32


def _arr(x, dtype=None):
    a = np.asarray(x)
    return a.astype(dtype) if dtype is not None else a


# ---------------------------------------------------------------- keras layers (documented semantics) ----------------
class Layer:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.call(*a, **k)


class Model(Layer):
    pass


class Dense(Layer):
    """keras.layers.Dense: y = activation(x @ kernel + bias), kernel [in, units] built lazily (zeros here: the golden
    generator overwrites every variable explicitly)."""

    def __init__(self, units, activation=None):
        self.units, self.activation = int(units), activation
        self.kernel = self.bias = None

    def call(self, x):
        x = _arr(x, F32)
        if self.kernel is None:
            self.kernel = np.zeros((x.shape[-1], self.units), F32)
            self.bias = np.zeros((self.units,), F32)
        y = (x @ self.kernel + self.bias).astype(F32)
        if self.activation == "relu":
            y = np.maximum(y, F32(0))
        elif self.activation is not None:
            raise NotImplementedError(self.activation)
        return y


class Embedding(Layer):
    def __init__(self, input_dim, output_dim):
        self.embeddings = np.zeros((int(input_dim), int(output_dim)), F32)

    def call(self, ids):
        return self.embeddings[_arr(ids).astype(np.int64)]


class LayerNormalization(Layer):
    """keras LayerNormalization over the last axis, non-fused formula: (x - mean) * rsqrt(var + eps) * gamma + beta with
    the biased variance."""

    def __init__(self, epsilon=1e-3):
        self.epsilon = F32(epsilon)
        self.gamma = self.beta = None

    def call(self, x):
        x = _arr(x, F32)
        if self.gamma is None:
            self.gamma = np.ones((x.shape[-1],), F32)
            self.beta = np.zeros((x.shape[-1],), F32)
        mean = x.mean(-1, keepdims=True, dtype=F32)
        var = ((x - mean) ** 2).mean(-1, keepdims=True, dtype=F32)
        return ((x - mean) / np.sqrt(var + self.epsilon) * self.gamma + self.beta).astype(F32)


class Dropout(Layer):
    def __init__(self, rate):
        self.rate = rate

    def call(self, x, training=False):
        if training and self.rate > 0:
            raise NotImplementedError("tf_numpy_stub runs inference-mode graphs only")
        return x


class Sequential(Layer):
    def __init__(self, layers):
        self.layers = list(layers)

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class SparseCategoricalCrossentropy:
    """from_logits=True, reduction='none': lse(pred) - pred[label] per position."""

    def __init__(self, from_logits=False, reduction="auto"):
        assert from_logits and reduction == "none"

    def __call__(self, real, pred):
        pred = _arr(pred, F32)
        m = pred.max(-1, keepdims=True)
        lse = (m + np.log(np.exp(pred - m).sum(-1, keepdims=True, dtype=F32))).squeeze(-1)
        picked = np.take_along_axis(pred, _arr(real).astype(np.int64)[..., None], -1).squeeze(-1)
        return (lse - picked).astype(F32)


class LearningRateSchedule:
    pass


# ---------------------------------------------------------------- functional ops -------------------------------------
def _softmax(x, axis=-1):
    x = _arr(x, F32)
    m = x.max(axis, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis, keepdims=True, dtype=F32)).astype(F32)


def _matmul(a, b, transpose_b=False):
    b = _arr(b, F32)
    return (_arr(a, F32) @ (np.swapaxes(b, -1, -2) if transpose_b else b)).astype(F32)


def _band_part(x, num_lower, num_upper):
    """tf.linalg.band_part: keep element (m, n) iff (num_lower < 0 or m - n <= num_lower) and (num_upper < 0 or n - m <= num_upper)"""
    x = _arr(x)
    m, n = np.indices(x.shape[-2:])
    keep = ((num_lower < 0) | (m - n <= num_lower)) & ((num_upper < 0) | (n - m <= num_upper))
    return x * keep.astype(x.dtype)


def _function(f=None, **_):
    return f if f is not None else (lambda g: g)


def build():
    """-> a module object to be installed as sys.modules['tensorflow']"""
    tf = types.ModuleType("tensorflow")
    tf.__version__ = "2.numpy-stub"
    tf.float32, tf.int64, tf.newaxis = F32, np.int64, np.newaxis
    tf.cast = lambda x, dtype: _arr(x).astype(dtype)
    tf.shape = lambda x: np.asarray(_arr(x).shape)
    tf.ones = lambda shape: np.ones(tuple(int(s) for s in shape), F32)
    tf.maximum = lambda a, b: np.maximum(_arr(a), _arr(b))
    tf.matmul = _matmul
    tf.reshape = lambda x, shape: _arr(x).reshape(tuple(int(s) for s in shape))
    tf.transpose = lambda x, perm: np.transpose(_arr(x), perm)
    tf.reduce_sum = lambda x: _arr(x).sum(dtype=F32)
    tf.function = _function
    tf.math = types.SimpleNamespace(equal=lambda a, b: _arr(a) == b, logical_not=lambda a: ~_arr(a),
                                    sqrt=lambda x: np.sqrt(_arr(x, F32)), rsqrt=lambda x: F32(1) / np.sqrt(_arr(x, F32)),
                                    minimum=lambda a, b: np.minimum(_arr(a, F32), _arr(b, F32)))
    tf.nn = types.SimpleNamespace(softmax=_softmax)
    tf.linalg = types.SimpleNamespace(band_part=_band_part)
    layers = types.SimpleNamespace(Layer=Layer, Dense=Dense, Embedding=Embedding, LayerNormalization=LayerNormalization,
                                   Dropout=Dropout)
    tf.keras = types.SimpleNamespace(
        layers=layers, Model=Model, Sequential=Sequential,
        losses=types.SimpleNamespace(SparseCategoricalCrossentropy=SparseCategoricalCrossentropy),
        optimizers=types.SimpleNamespace(schedules=types.SimpleNamespace(LearningRateSchedule=LearningRateSchedule)))
    return tf


def install():
    """Install the stand-in as ``tensorflow`` (and an empty ``tensorflow_datasets``, which the reference's utils.py imports
    at module level but the functions pinned here never touch)."""
    tf = build()
    sys.modules["tensorflow"] = tf
    sys.modules.setdefault("tensorflow_datasets", types.ModuleType("tensorflow_datasets"))
    return tf
