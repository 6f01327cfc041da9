"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.utilities [3P-memory, gpflow 2.7.0]."""
import tensorflow as tf
import tensorflow_probability as tfp

from . import parameter_or_function  # noqa: F401


def positive(lower=None, base=None):
    """default_positive_bijector = softplus, default_positive_minimum = 0."""
    return tfp.bijectors.Softplus()


def triangular():
    return tfp.bijectors.FillTriangular()


def to_default_float(x):
    return tf.cast(x, tf.float64)


def to_default_int(x):
    return tf.cast(x, tf.int64)


def parameter_dict(module):
    """gpflow.utilities.parameter_dict: {'.path.to.parameter': Parameter} (keys carry GPflow's leading dot)."""
    return {"." + k: p for k, p in module.parameters_dict.items()}


def print_summary(module, fmt=None):
    for k, p in module.parameters_dict.items():
        print(f"{k:50s} {type(p.transform).__name__:18s} trainable={p.trainable} shape={tuple(p.shape)}")


def square_distance(X, X2):
    """gpflow/utilities/ops.py::square_distance — ||x||^2 + ||x2||^2 - 2 x.x2, NO clamp at zero."""
    if X2 is None:
        Xs = tf.reduce_sum(tf.square(X), axis=-1, keepdims=True)
        dist = -2 * tf.matmul(X, X, transpose_b=True)
        dist = dist + (Xs + tf.linalg.adjoint(Xs))
        return dist
    Xs = tf.reduce_sum(tf.square(X), axis=-1)
    X2s = tf.reduce_sum(tf.square(X2), axis=-1)
    dist = -2 * tf.tensordot(X, X2, [[-1], [-1]])
    # broadcasting_elementwise(tf.add, Xs, X2s): outer sum over all index pairs
    Xs_, X2s_ = tf._t(Xs), tf._t(X2s)
    outer = Xs_.reshape(list(Xs_.shape) + [1] * X2s_.dim()) + X2s_
    dist = dist + outer
    return dist
