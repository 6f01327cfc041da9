"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.kullback_leiblers.prior_kl / gauss_kl
[3P-memory, gpflow 2.7.0]; SURVEY.md Appendix A.4.  Call site: MixtureGPs/models.py:79."""
import tensorflow as tf

from .covariances import Kuu
from .config import default_jitter


def prior_kl(inducing_variable, kernel, q_mu, q_sqrt, whiten=False):
    if whiten:
        return gauss_kl(q_mu, q_sqrt, None)
    K = Kuu(inducing_variable, kernel, jitter=default_jitter())
    return gauss_kl(q_mu, q_sqrt, K)


def gauss_kl(q_mu, q_sqrt, K=None):
    q_mu, q_sqrt = tf._t(q_mu), tf._t(q_sqrt)
    is_white = K is None
    is_diag = q_sqrt.dim() == 2
    if is_white:
        alpha = q_mu  # [M, L]
    else:
        Lp = tf.linalg.cholesky(K)
        alpha = tf.linalg.triangular_solve(Lp, q_mu, lower=True)
    if is_diag:
        Lq = Lq_diag = q_sqrt
    else:
        Lq = tf.linalg.band_part(q_sqrt, -1, 0)  # [L, M, M]
        Lq_diag = tf.linalg.diag_part(Lq)  # [L, M]
    mahalanobis = tf.reduce_sum(tf.square(alpha))
    constant = -float(tf.size(q_mu))
    logdet_qcov = tf.reduce_sum(tf.math.log(tf.square(Lq_diag)))
    if is_white:
        trace = tf.reduce_sum(tf.square(Lq))
    else:
        raise NotImplementedError("unwhitened KL is not on the reference's path (whiten=True everywhere)")
    twoKL = mahalanobis + constant - logdet_qcov + trace
    return 0.5 * twoKL
