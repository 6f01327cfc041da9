"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.kernels.SquaredExponential
[3P-memory, gpflow 2.7.0 gpflow/kernels/stationaries.py]; SURVEY.md Appendix A.1."""
import tensorflow as tf

from .base import Module, Parameter
from .utilities import positive, square_distance


class Kernel(Module):
    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):
        if (not full_cov) and (X2 is not None):
            raise ValueError("Ambiguous inputs: `not full_cov` and `X2` are not compatible.")
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)


class Stationary(Kernel):
    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())

    @property
    def ard(self):
        return len(self.lengthscales.shape) > 0

    def scale(self, X):
        return tf._t(X) / self.lengthscales if X is not None else X

    def K_diag(self, X):
        return tf.fill(tf.shape(X)[:-1], tf.squeeze(self.variance))


class SquaredExponential(Stationary):
    def K(self, X, X2=None):
        r2 = self.scaled_squared_euclid_dist(X, X2)
        return self.K_r2(r2)

    def scaled_squared_euclid_dist(self, X, X2=None):
        return square_distance(self.scale(X), self.scale(X2))

    def K_r2(self, r2):
        return self.variance * tf.exp(-0.5 * r2)


RBF = SquaredExponential
