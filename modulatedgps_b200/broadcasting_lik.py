"""BroadcastingLikelihood — MixtureGPs/broadcasting_lik.py:5-46.

In the reference this wrapper makes a likelihood accept [S, N, K] F with [N, 1] Y: pass-through with `Y[None]` for
GaussianModified (:22-24), tile + flatten to [S*N, .] + call + reshape back to [S, N, -1] for the rest (:26-37).  Here
the tiling never happens: the device kernels behind `mgp_lik_*` index Y by `row % N`, so the same [S, N, K] -> [S, N, .]
contract holds without the S-fold copy of Y.  On the ELBO path the S axis never exists as data at all (the SVGP
conditional is evaluated once per point and the fused Monte-Carlo kernel consumes it); `kind` tells that kernel which
expectation to evaluate.
"""
from __future__ import annotations

from . import _lib
from .likelihoods import GaussianModified, MultiClass, _wrap
from .parameter import to_device_f64


class BroadcastingLikelihood:
    def __init__(self, likelihood):
        self.likelihood = likelihood
        self.needs_broadcasting = not isinstance(likelihood, GaussianModified)   # broadcasting_lik.py:17-20
        if isinstance(likelihood, GaussianModified):
            self.kind = _lib.LIK_GAUSSIAN
        elif isinstance(likelihood, MultiClass):
            self.kind = _lib.LIK_MULTICLASS
        else:
            raise NotImplementedError(
                f"{type(likelihood).__name__}: libmgp implements the likelihoods the reference's demos use "
                "(GaussianModified, MultiClass(RobustMax))")

    def component_variances(self, K):
        return self.likelihood.component_variances(K) if self.kind == _lib.LIK_GAUSSIAN else None

    @staticmethod
    def _snk(F, what):
        Ft = to_device_f64(F)
        if Ft.dim() != 3:
            raise ValueError(f"{what} must be [S, N, K]; got shape {tuple(Ft.shape)}")
        return Ft

    def variational_expectations(self, X, Fmu, Fvar, Y):
        """broadcasting_lik.py:39-42.  Fmu, Fvar [S, N, K], Y [N, 1] -> [S, N, K] (GaussianModified, per component) or
        [S, N, 1] (MultiClass).  X is ignored, as in the reference (it passes [])."""
        Fm, Fv = self._snk(Fmu, "Fmu"), self._snk(Fvar, "Fvar")
        S, N, K = Fm.shape
        Yt = to_device_f64(Y, Fm.device).reshape(-1)
        if Yt.numel() != N:
            raise ValueError(f"Y has {Yt.numel()} rows but F has N={N}")
        out = self.likelihood._variational_expectations([], Fm, Fv, Yt.reshape(N, 1))
        return _wrap(out.reshape(S, N, -1))

    def predict_mean_and_var(self, X, Fmu, Fvar):
        """broadcasting_lik.py:44-46 -> (mean [S, N, K], var [S, N, K])."""
        Fm, Fv = self._snk(Fmu, "Fmu"), self._snk(Fvar, "Fvar")
        S, N, K = Fm.shape
        mean, var = self.likelihood._predict_mean_and_var([], Fm, Fv)
        return _wrap(mean.reshape(S, N, -1)), _wrap(var.reshape(S, N, -1))
