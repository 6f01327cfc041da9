"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.base — Parameter with bijector, Module attribute walk.
[3P-memory, gpflow 2.7.0 gpflow/base.py]; SURVEY.md Appendix A.7."""
from typing import Any, Tuple, Union

import numpy as np
import torch

import tensorflow as tf

TensorType = Any
MeanAndVariance = Tuple[Any, Any]


class _Identity:
    def forward(self, x):
        return x

    def inverse(self, y):
        return torch.as_tensor(y, dtype=torch.float64)


class Parameter:
    """Constrained view of an unconstrained leaf tensor.  value = transform.forward(unconstrained)."""

    def __init__(self, value, *, transform=None, prior=None, trainable=True, dtype=None, name=None):
        if isinstance(value, Parameter):
            value = value._shim_value().detach()
        self.transform = transform if transform is not None else _Identity()
        v = torch.as_tensor(np.asarray(value, dtype=np.float64) if not isinstance(value, torch.Tensor) else value,
                            dtype=torch.float64)
        self.unconstrained_variable = self.transform.inverse(v).detach().clone().requires_grad_(bool(trainable))
        self.trainable = bool(trainable)
        self.prior = prior
        self.name = name

    def _shim_value(self):
        return self.transform.forward(self.unconstrained_variable)

    def assign(self, value):
        v = torch.as_tensor(np.asarray(value, dtype=np.float64), dtype=torch.float64)
        with torch.no_grad():
            self.unconstrained_variable.copy_(self.transform.inverse(v))

    def numpy(self):
        return self._shim_value().detach().numpy()

    @property
    def shape(self):
        return self._shim_value().shape

    @property
    def dtype(self):
        return torch.float64

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def unwrap(a):
            if isinstance(a, Parameter):
                return a._shim_value()
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(e) for e in a)
            return a
        return func(*unwrap(args), **{k: unwrap(v) for k, v in (kwargs or {}).items()})

    def __array__(self, dtype=None):
        return self.numpy()

    def __getitem__(self, i):
        return self._shim_value()[i]

    def __neg__(self):
        return -self._shim_value()

    def __add__(self, o):
        return self._shim_value() + tf._t(o)

    __radd__ = __add__

    def __sub__(self, o):
        return self._shim_value() - tf._t(o)

    def __rsub__(self, o):
        return tf._t(o) - self._shim_value()

    def __mul__(self, o):
        return self._shim_value() * tf._t(o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self._shim_value() / tf._t(o)

    def __rtruediv__(self, o):
        return tf._t(o) / self._shim_value()

    def __pow__(self, o):
        return self._shim_value() ** o


class Module(tf.Module):
    """gpflow.base.Module: tf.Module attribute walk (Modules, lists, tuples, dicts only — plain
    Python objects such as the reference's BroadcastingLikelihood are NOT descended into)."""

    def _walk(self, prefix, seen, out):
        for name in sorted(vars(self)):
            _walk_value(vars(self)[name], f"{prefix}.{name}" if prefix else name, seen, out)

    @property
    def parameters_dict(self):
        out, seen = {}, set()
        self._walk("", seen, out)
        return out

    @property
    def trainable_parameters(self):
        return tuple(p for p in self.parameters_dict.values() if p.trainable)

    @property
    def trainable_variables(self):
        return tuple(p.unconstrained_variable for p in self.trainable_parameters)


def _walk_value(v, path, seen, out):
    if isinstance(v, Parameter):
        if id(v) not in seen:
            seen.add(id(v))
            out[path] = v
    elif isinstance(v, Module):
        if id(v) not in seen:
            seen.add(id(v))
            v._walk(path, seen, out)
    elif isinstance(v, (list, tuple)):
        for i, e in enumerate(v):
            _walk_value(e, f"{path}[{i}]", seen, out)
    elif isinstance(v, dict):
        for k, e in v.items():
            _walk_value(e, f"{path}[{k}]", seen, out)
