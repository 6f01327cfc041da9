"""TEST INFRASTRUCTURE ONLY (oracle shim): the slice of gpflow.posteriors the reference subclasses
(MixtureGPs/models.py:126-160) [3P-memory, gpflow 2.7.0 gpflow/posteriors.py]."""
import enum

from .base import Module, Parameter


class PrecomputeCacheType(enum.Enum):
    TENSOR = "tensor"
    VARIABLE = "variable"
    NOCACHE = "nocache"


class AbstractPosterior(Module):
    def __init__(self, kernel, X_data, cache=None, mean_function=None):
        self.kernel = kernel
        self.X_data = X_data
        self.cache = cache
        self.mean_function = mean_function

    def _add_mean_function(self, Xnew, mean):
        if self.mean_function is None:
            return mean
        return mean + self.mean_function(Xnew)

    def fused_predict_f(self, Xnew, full_cov=False, full_output_cov=False):
        mean, cov = self._conditional_fused(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
        return self._add_mean_function(Xnew, mean), cov


class BasePosterior(AbstractPosterior):
    def __init__(self, kernel, inducing_variable, q_mu, q_sqrt, whiten=True, mean_function=None, *,
                 precompute_cache=None):
        super().__init__(kernel, inducing_variable, mean_function=mean_function)
        self.whiten = whiten
        self._q_mu = q_mu
        self._q_sqrt = q_sqrt
        # caches (alpha, Qinv) are only used by predict_f on a cached posterior; SVGP.predict_f
        # asks for NOCACHE and the reference never calls the cached path, so none is built here.

    @property
    def q_mu(self):
        return self._q_mu

    @property
    def q_sqrt(self):
        return self._q_sqrt


class IndependentPosterior(BasePosterior):
    def _post_process_mean_and_cov(self, mean, cov, full_cov, full_output_cov):
        # expand_independent_outputs: identity for full_cov=False, full_output_cov=False ([N, P])
        if full_cov or full_output_cov:
            raise NotImplementedError("the reference only calls full_cov=False (models.py:40,56,64,87,96,113,118)")
        return mean, cov
