"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.conditionals.base_conditional, transcribed op-for-op
from memory of gpflow 2.7.0 gpflow/conditionals/util.py [3P-memory]; SURVEY.md Appendix A.3.
Call site in the reference: MixtureGPs/models.py:141-143."""
import tensorflow as tf


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    Lm = tf.linalg.cholesky(Kmm)
    return base_conditional_with_lm(Kmn=Kmn, Lm=Lm, Knn=Knn, f=f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


def base_conditional_with_lm(Kmn, Lm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    f = tf._t(f)
    num_func = tf.shape(f)[-1]  # R
    N = tf.shape(Kmn)[-1]
    M = tf.shape(f)[-2]

    # get the leading dims in Kmn to the front of the tensor: [M, ..., N] -> [..., M, N]
    K = tf.rank(Kmn)
    perm = list(range(1, K - 1)) + [0, K - 1]
    Kmn = tf.transpose(Kmn, perm)
    leading_dims = list(tf.shape(Kmn)[:-2])

    # projection matrix A
    Lm = tf.broadcast_to(Lm, leading_dims + list(tf.shape(Lm)))  # [..., M, M]
    A = tf.linalg.triangular_solve(Lm, Kmn, lower=True)  # [..., M, N]

    # covariance due to the conditioning
    if full_cov:
        fvar = Knn - tf.linalg.matmul(A, A, transpose_a=True)
        fvar = tf.broadcast_to(tf.expand_dims(fvar, -3), leading_dims + [num_func, N, N])
    else:
        fvar = Knn - tf.reduce_sum(tf.square(A), -2)  # [..., N]
        fvar = tf.broadcast_to(tf.expand_dims(fvar, -2), leading_dims + [num_func, N])  # [..., R, N]

    # another backsubstitution in the unwhitened case
    if not white:
        A = tf.linalg.triangular_solve(tf.linalg.adjoint(Lm), A, lower=False)

    # conditional mean
    f = tf.broadcast_to(f, leading_dims + [M, num_func])  # [..., M, R]
    fmean = tf.linalg.matmul(A, f, transpose_a=True)  # [..., N, R]

    if q_sqrt is not None:
        q_sqrt = tf._t(q_sqrt)
        q_sqrt_dims = q_sqrt.dim()
        if q_sqrt_dims == 2:
            LTA = A * tf.expand_dims(tf.transpose(q_sqrt), 2)  # [R, M, N]
        elif q_sqrt_dims == 3:
            L = tf.linalg.band_part(q_sqrt, -1, 0)  # force lower triangle [R, M, M]
            L = tf.broadcast_to(L, leading_dims + list(tf.shape(L)))
            A_tiled = tf.broadcast_to(tf.expand_dims(A, -3), leading_dims + [num_func, M, N])
            LTA = tf.linalg.matmul(L, A_tiled, transpose_a=True)  # [..., R, M, N]
        else:
            raise ValueError("Bad dimension for q_sqrt: %s" % str(q_sqrt_dims))
        if full_cov:
            fvar = fvar + tf.linalg.matmul(LTA, LTA, transpose_a=True)
        else:
            fvar = fvar + tf.reduce_sum(tf.square(LTA), -2)  # [..., R, N]

    if not full_cov:
        fvar = tf.linalg.adjoint(fvar)  # [..., N, R]
    return fmean, fvar
