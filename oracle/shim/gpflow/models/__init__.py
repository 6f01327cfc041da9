"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.models.SVGP / GPModel
[3P-memory, gpflow 2.7.0 gpflow/models/{model,svgp}.py]; SURVEY.md Appendix A.5."""
import numpy as np

from .. import kullback_leiblers, posteriors
from ..base import Module, Parameter
from ..config import default_float
from ..inducing_variables import inducingpoint_wrapper
from ..utilities import triangular
from . import training_mixins  # noqa: F401
from .training_mixins import ExternalDataTrainingLossMixin


class GPModel(Module):
    def __init__(self, kernel, likelihood, mean_function=None, num_latent_gps=None):
        assert num_latent_gps is not None
        self.num_latent_gps = num_latent_gps
        self.mean_function = mean_function  # None == Zero()
        self.kernel = kernel
        self.likelihood = likelihood


class SVGP(GPModel, ExternalDataTrainingLossMixin):
    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps=1,
                 q_diag=False, q_mu=None, q_sqrt=None, whiten=True, num_data=None):
        super().__init__(kernel, likelihood, mean_function, num_latent_gps)
        self.num_data = num_data
        self.q_diag = q_diag
        self.whiten = whiten
        self.inducing_variable = inducingpoint_wrapper(inducing_variable)
        num_inducing = self.inducing_variable.num_inducing
        self._init_variational_parameters(num_inducing, q_mu, q_sqrt, q_diag)

    def _init_variational_parameters(self, num_inducing, q_mu, q_sqrt, q_diag):
        q_mu = np.zeros((num_inducing, self.num_latent_gps)) if q_mu is None else q_mu
        self.q_mu = Parameter(q_mu, dtype=default_float())
        if q_diag:
            raise NotImplementedError("q_diag=True is not on the reference's path")
        if q_sqrt is None:
            q_sqrt = np.array([np.eye(num_inducing, dtype=np.float64) for _ in range(self.num_latent_gps)])
        self.q_sqrt = Parameter(q_sqrt, transform=triangular())

    def prior_kl(self):
        return kullback_leiblers.prior_kl(self.inducing_variable, self.kernel, self.q_mu, self.q_sqrt,
                                          whiten=self.whiten)

    def posterior(self, precompute_cache=posteriors.PrecomputeCacheType.TENSOR):
        raise NotImplementedError("the reference overrides posterior() (MixtureGPs/models.py:147-160)")

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False):
        return self.posterior(posteriors.PrecomputeCacheType.NOCACHE).fused_predict_f(
            Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
