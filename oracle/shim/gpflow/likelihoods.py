"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.likelihoods base classes + MultiClass/RobustMax,
transcribed from memory of gpflow 2.7.0 gpflow/likelihoods/{base,multiclass}.py [3P-memory];
SURVEY.md Appendix A.5.  Constructed by the reference at demos/demo_tf2_2d_modified_multiclass.py:41-42."""
import numpy as np
import tensorflow as tf
import tensorflow_probability as tfp

from .base import Module, Parameter
from .utilities import to_default_int

# gpflow's literal in RobustMax.prob_is_largest: `cdfs = cdfs * (1 - 2e-4) + 1e-4` (the attribute _squash = 1e-6 below
# exists in gpflow too and is never read there).  Same named constant as oracle/svgp_mixture.py and include/mgp.h.
ROBUSTMAX_CDF_SQUASH = 1e-4


def hermgauss(n):
    x, w = np.polynomial.hermite.hermgauss(n)
    return x.astype(np.float64), w.astype(np.float64)


class Likelihood(Module):
    def __init__(self, input_dim=None, latent_dim=None, observation_dim=None, **kwargs):
        self.input_dim, self.latent_dim, self.observation_dim = input_dim, latent_dim, observation_dim

    def variational_expectations(self, X, Fmu, Fvar, Y):
        return self._variational_expectations(X, Fmu, Fvar, Y)

    def predict_mean_and_var(self, X, Fmu, Fvar):
        return self._predict_mean_and_var(X, Fmu, Fvar)


class ScalarLikelihood(Likelihood):
    def __init__(self, **kwargs):
        super().__init__(input_dim=None, latent_dim=None, observation_dim=None, **kwargs)
        self.num_gauss_hermite_points = 20


class RobustMax(Module):
    def __init__(self, num_classes, epsilon=1e-3, **kwargs):
        # Beta(0.2, 5) prior never enters: SMGP._training_loss ignores log_prior_density (models.py:81-83)
        self.epsilon = Parameter(epsilon, transform=tfp.bijectors.Sigmoid(), trainable=False)
        self.num_classes = num_classes
        self._squash = 1e-6

    @property
    def eps_k1(self):
        return self.epsilon / (self.num_classes - 1.0)

    def safe_sqrt(self, val):
        return tf.sqrt(tf.clip_by_value(val, 1e-10, np.inf))

    def prob_is_largest(self, Y, mu, var, gh_x, gh_w):
        Y = to_default_int(Y)
        mu, var = tf._t(mu), tf._t(var)
        gh_x, gh_w = tf._t(gh_x), tf._t(gh_w)
        oh_on = tf.cast(tf.one_hot(tf.reshape(Y, (-1,)), self.num_classes, 1.0, 0.0), dtype=mu.dtype)
        mu_selected = tf.reduce_sum(oh_on * mu, 1)
        var_selected = tf.reduce_sum(oh_on * var, 1)
        # Gauss-Hermite grid
        X = tf.reshape(mu_selected, (-1, 1)) + gh_x * tf.reshape(self.safe_sqrt(2.0 * var_selected), (-1, 1))
        # CDF of the Gaussian between the latent functions and the grid (including the selected function)
        dist = (tf.expand_dims(X, 1) - tf.expand_dims(mu, 2)) / tf.expand_dims(self.safe_sqrt(var), 2)
        cdfs = 0.5 * (1.0 + tf.math.erf(dist / np.sqrt(2.0)))
        cdfs = cdfs * (1 - 2 * ROBUSTMAX_CDF_SQUASH) + ROBUSTMAX_CDF_SQUASH
        # blank out all the distances on the selected latent function
        oh_off = tf.cast(tf.one_hot(tf.reshape(Y, (-1,)), self.num_classes, 0.0, 1.0), dtype=mu.dtype)
        cdfs = cdfs * tf.expand_dims(oh_off, 2) + tf.expand_dims(oh_on, 2)
        # product over the latent functions, sum over the GH grid
        gh_w = tf.reshape(gh_w, (-1, 1))
        return tf.linalg.matmul(tf.reduce_prod(cdfs, axis=[1]), gh_w / np.sqrt(np.pi))


class MultiClass(Likelihood):
    def __init__(self, num_classes, invlink=None, **kwargs):
        super().__init__(input_dim=None, latent_dim=num_classes, observation_dim=None, **kwargs)
        self.num_classes = num_classes
        self.num_gauss_hermite_points = 20
        if invlink is None:
            invlink = RobustMax(self.num_classes)
        if not isinstance(invlink, RobustMax):
            raise NotImplementedError
        self.invlink = invlink

    def _variational_expectations(self, X, Fmu, Fvar, Y):
        gh_x, gh_w = hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        ve = p * tf.math.log(1.0 - self.invlink.epsilon) + (1.0 - p) * tf.math.log(self.invlink.eps_k1)
        return tf.reduce_sum(ve, axis=-1)

    def _predict_non_logged_density(self, X, Fmu, Fvar, Y):
        gh_x, gh_w = hermgauss(self.num_gauss_hermite_points)
        p = self.invlink.prob_is_largest(Y, Fmu, Fvar, gh_x, gh_w)
        den = p * (1.0 - self.invlink.epsilon) + (1.0 - p) * (self.invlink.eps_k1)
        return den

    def _predict_mean_and_var(self, X, Fmu, Fvar):
        n = tf.shape(Fmu)[0]
        possible_outputs = [np.full((n, 1), i, dtype=np.int64) for i in range(self.num_classes)]
        ps = [self._predict_non_logged_density(X, Fmu, Fvar, po) for po in possible_outputs]
        ps = tf.transpose(tf.stack([tf.reshape(p, (-1,)) for p in ps]))
        return ps, ps - tf.square(ps)
