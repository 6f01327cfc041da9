"""Likelihoods of the mixture model — parameter holders mirroring MixtureGPs/likelihoods.py and the gpflow
classes the demos construct.  Their expectations are evaluated inside the fused Monte-Carlo kernel
(csrc/mc_pass.cu); nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Any

import numpy as np

from .parameter import Module, Parameter, Softplus


class Likelihood(Module):
    pass


class GaussianModified(Likelihood):
    """MixtureGPs/likelihoods.py:12-41.  Gaussian likelihood with one variance per component
    (`variance * np.ones((1, D))`, likelihoods.py:16-17); its variational expectation is NOT summed over
    components (likelihoods.py:39-41)."""

    def __init__(self, variance=1e-0, D: int = None, **kwargs: Any) -> None:
        if D is not None:
            variance = variance * np.ones((1, D))
        self.variance = Parameter(np.asarray(variance, dtype=np.float64), transform=Softplus())

    def component_variances(self, K: int):
        v = self.variance.value().reshape(-1)
        if v.numel() == 1 and K > 1:
            v = v.expand(K)
        if v.numel() != K:
            raise ValueError(f"GaussianModified has {v.numel()} variances but the model has K={K} components")
        return v.contiguous()


class RobustMax(Module):
    """gpflow.likelihoods.RobustMax(num_classes): epsilon = 1e-3 is sigmoid-transformed and NOT trainable; its
    Beta prior never enters the reference's loss (SMGP._training_loss ignores priors, models.py:81-83)."""

    def __init__(self, num_classes, epsilon=1e-3, **kwargs):
        if epsilon != 1e-3:
            raise NotImplementedError("libmgp fixes RobustMax.epsilon at GPflow's default 1e-3")
        self.num_classes = int(num_classes)
        self.epsilon = float(epsilon)


class MultiClass(Likelihood):
    """gpflow.likelihoods.MultiClass(num_classes, invlink=RobustMax(num_classes)) as built at
    demos/demo_tf2_2d_modified_multiclass.py:41-42; 20-point Gauss-Hermite expectation."""

    def __init__(self, num_classes, invlink=None, **kwargs):
        self.num_classes = int(num_classes)
        self.num_gauss_hermite_points = 20
        self.invlink = invlink if invlink is not None else RobustMax(self.num_classes)
        if not isinstance(self.invlink, RobustMax):
            raise NotImplementedError("only RobustMax is supported (as in gpflow)")
        if self.invlink.num_classes != self.num_classes:
            raise ValueError("RobustMax.num_classes != MultiClass.num_classes")
