"""Likelihoods of the mixture model, mirroring MixtureGPs/likelihoods.py and the gpflow classes the demos construct.
On the ELBO path their expectations are evaluated inside the fused Monte-Carlo kernel (csrc/mc_pass.cu); the methods
the reference also exposes for direct calls (`_variational_expectations`, `_predict_mean_and_var`, `_scalar_log_prob`,
...) run as small device kernels behind `mgp_lik_*` (include/mgp.h).  Nothing here computes on the CPU.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from . import _lib
from .parameter import F64, Module, Parameter, Softplus, to_device_f64


def _wrap(t):
    from .models import _wrap as w
    return w(t)


def _rows(F, what):
    """[.., K] -> ([R, K] contiguous device tensor, leading shape)."""
    Ft = to_device_f64(F)
    if Ft.dim() < 1:
        raise ValueError(f"{what} must have a trailing component axis")
    return Ft.reshape(-1, Ft.shape[-1]).contiguous(), tuple(Ft.shape[:-1])


def _y_for(Y, lead, device):
    """Y as the reference hands it to a likelihood — [N, 1], [1, N, 1] (BroadcastingLikelihood's expand_dims) or already
    tiled [S*N, 1] — -> (flat [n] device tensor, n) such that row r of the flattened F reads Y[r % n]."""
    Yt = to_device_f64(Y, device)
    if Yt.dim() >= 1 and Yt.shape[-1] == 1:
        Yt = Yt.reshape(Yt.shape[:-1])
    Yt = Yt.reshape(-1).contiguous()
    R = int(np.prod(lead)) if lead else 1
    n = Yt.numel()
    if n == 0 or R % n != 0:
        raise ValueError(f"Y with {n} rows does not broadcast against F with leading shape {lead}")
    return Yt, n


class Likelihood(Module):
    kind = None

    def component_variances(self, K):
        return None

    # gpflow.likelihoods.Likelihood's public wrappers around the underscored methods
    def variational_expectations(self, X, Fmu, Fvar, Y):
        return self._variational_expectations(X, Fmu, Fvar, Y)

    def predict_mean_and_var(self, X, Fmu, Fvar):
        return self._predict_mean_and_var(X, Fmu, Fvar)

    def _predict_mean_and_var(self, X, Fmu, Fvar):
        Fm, lead = _rows(Fmu, "Fmu")
        Fv, lead_v = _rows(Fvar, "Fvar")
        if Fm.shape != Fv.shape:
            raise ValueError(f"Fmu {tuple(Fm.shape)} and Fvar {tuple(Fv.shape)} differ")
        K = Fm.shape[1]
        ctx = _lib.get_context(Fm.device)
        mean, var = torch.empty_like(Fm), torch.empty_like(Fm)
        lv = self.component_variances(K)
        ctx.check(ctx.lib.mgp_lik_predict_mean_and_var(ctx.handle, self.kind, _lib.ptr(lv), _lib.ptr(Fm), _lib.ptr(Fv),
                                                       Fm.shape[0], K, _lib.ptr(mean), _lib.ptr(var)))
        return _wrap(mean.reshape(*lead, K)), _wrap(var.reshape(*lead, K))

    def _ve(self, Fmu, Fvar, Y):
        Fm, lead = _rows(Fmu, "Fmu")
        Fv, _ = _rows(Fvar, "Fvar")
        if Fm.shape != Fv.shape:
            raise ValueError(f"Fmu {tuple(Fm.shape)} and Fvar {tuple(Fv.shape)} differ")
        K = Fm.shape[1]
        Yt, n = _y_for(Y, lead, Fm.device)
        ctx = _lib.get_context(Fm.device)
        gaussian = self.kind == _lib.LIK_GAUSSIAN
        out = torch.empty(Fm.shape if gaussian else (Fm.shape[0],), dtype=F64, device=Fm.device)
        lv = self.component_variances(K)
        ctx.check(ctx.lib.mgp_lik_variational_expectations(ctx.handle, self.kind, _lib.ptr(lv), _lib.ptr(Fm), _lib.ptr(Fv),
                                                           _lib.ptr(Yt), Fm.shape[0] // n, n, K, _lib.ptr(out)))
        return out, lead, K


class GaussianModified(Likelihood):
    """MixtureGPs/likelihoods.py:12-41.  Gaussian likelihood with one variance per component
    (`variance * np.ones((1, D))`, likelihoods.py:16-17); its variational expectation is NOT summed over
    components (likelihoods.py:39-41)."""

    def __init__(self, variance=1e-0, D: int = None, **kwargs: Any) -> None:
        if D is not None:
            variance = variance * np.ones((1, D))
        self.variance = Parameter(np.asarray(variance, dtype=np.float64), transform=Softplus())

    kind = _lib.LIK_GAUSSIAN

    def component_variances(self, K: int):
        v = self.variance.value().reshape(-1)
        if v.numel() == 1 and K > 1:
            v = v.expand(K)
        if v.numel() != K:
            raise ValueError(f"GaussianModified has {v.numel()} variances but the model has K={K} components")
        return v.contiguous()

    def _variational_expectations(self, X, Fmu, Fvar, Y):
        """likelihoods.py:39-41: -0.5 log 2 pi - 0.5 log variance - 0.5 ((Y - Fmu)^2 + Fvar) / variance, per component
        (NOT reduced over the component axis).  Fmu, Fvar [.., K]; Y [N, 1] broadcast over the leading axes."""
        out, lead, K = self._ve(Fmu, Fvar, Y)
        return _wrap(out.reshape(*lead, K))

    def _scalar_log_prob(self, X, F, Y):
        """likelihoods.py:21-22: logdensities.gaussian(Y, F, variance) -> [.., K]."""
        Fm, lead = _rows(F, "F")
        K = Fm.shape[1]
        Yt, n = _y_for(Y, lead, Fm.device)
        ctx = _lib.get_context(Fm.device)
        out = torch.empty_like(Fm)
        ctx.check(ctx.lib.mgp_lik_log_prob(ctx.handle, _lib.ptr(self.component_variances(K)), _lib.ptr(Fm), _lib.ptr(Yt),
                                           Fm.shape[0] // n, n, K, _lib.ptr(out)))
        return _wrap(out.reshape(*lead, K))

    def _predict_log_density(self, X, Fmu, Fvar, Y):
        """likelihoods.py:34-35: sum_k gaussian(Y, Fmu, Fvar + variance) -> [..]."""
        Fm, lead = _rows(Fmu, "Fmu")
        Fv, _ = _rows(Fvar, "Fvar")
        K = Fm.shape[1]
        Yt, n = _y_for(Y, lead, Fm.device)
        ctx = _lib.get_context(Fm.device)
        out = torch.empty(Fm.shape[0], dtype=F64, device=Fm.device)
        ctx.check(ctx.lib.mgp_lik_predict_log_density(ctx.handle, _lib.ptr(self.component_variances(K)), _lib.ptr(Fm),
                                                      _lib.ptr(Fv), _lib.ptr(Yt), Fm.shape[0] // n, n, K, _lib.ptr(out)))
        return _wrap(out.reshape(lead))

    def _conditional_mean(self, X, F):
        """likelihoods.py:24-25: tf.identity(F)."""
        return _wrap(to_device_f64(F).clone())

    def _conditional_variance(self, X, F):
        """likelihoods.py:27-29: the variance broadcast to F's shape."""
        Ft = to_device_f64(F)
        return _wrap(self.component_variances(Ft.shape[-1]).expand(Ft.shape))


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

    kind = _lib.LIK_MULTICLASS

    def _variational_expectations(self, X, Fmu, Fvar, Y):
        """gpflow MultiClass._variational_expectations: p log(1 - eps) + (1 - p) log(eps / (K - 1)) with p the 20-point
        Gauss-Hermite estimate of P(f_y is largest) -> [R] for Fmu, Fvar [R, K], Y [R, 1] class indices."""
        out, lead, _ = self._ve(Fmu, Fvar, Y)
        return _wrap(out.reshape(lead))
