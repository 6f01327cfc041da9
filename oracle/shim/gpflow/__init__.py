"""TEST INFRASTRUCTURE ONLY (oracle).  Torch-fp64 restatement of the slice of GPflow 2.7.0
(reference environment.yml:95) that /root/reference/MixtureGPs executes, laid out under GPflow's own
module names so the reference's files import unmodified.  Everything here is restated from memory of
the pinned version ([3P-memory]) — GPflow itself is not installable in this image.
Never imported by the product package."""
from . import base, config, covariances, conditionals, kernels, kullback_leiblers, likelihoods  # noqa: F401
from . import logdensities, models, posteriors, utilities, inducing_variables  # noqa: F401
from .base import Module, Parameter  # noqa: F401
from .config import default_float, default_int, default_jitter  # noqa: F401
