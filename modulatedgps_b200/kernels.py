"""Covariance functions (parameter holders; the arithmetic lives in libmgp).

SquaredExponential mirrors gpflow.kernels.SquaredExponential as constructed by the reference demos
(demos/demo_tf2.py:37-38): `variance` and `lengthscales` are softplus-constrained parameters; a scalar
lengthscale is isotropic, a length-D vector is ARD.  K(x, x') = variance * exp(-0.5 * |(x - x')/lengthscales|^2).
"""
from __future__ import annotations

import numpy as np

from .parameter import Module, Parameter, Softplus


class SquaredExponential(Module):
    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        if kwargs.get("active_dims") is not None:
            raise NotImplementedError("active_dims is not used anywhere on the reference's path")
        self.variance = Parameter(np.float64(variance), transform=Softplus())
        self.lengthscales = Parameter(np.asarray(lengthscales, dtype=np.float64), transform=Softplus())

    @property
    def ard(self) -> bool:
        return self.lengthscales.unconstrained_variable.dim() > 0


RBF = SquaredExponential
