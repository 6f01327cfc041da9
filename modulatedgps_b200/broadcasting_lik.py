"""BroadcastingLikelihood — MixtureGPs/broadcasting_lik.py:5-46.

In the reference this wrapper makes a likelihood accept [S, N, K] F with [N, 1] Y (pass-through for
GaussianModified, tile+flatten+reshape for the rest).  Here the S axis never exists as data (the SVGP
conditional is evaluated once per point), so the wrapper only records which path the fused kernel must take.
"""
from __future__ import annotations

from . import _lib
from .likelihoods import GaussianModified, MultiClass


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
