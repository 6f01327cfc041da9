"""reparameterize — MixtureGPs/utils.py:8-36 (diagonal branch; the full_cov branch is dead code in the
reference: it calls TF1's tf.cholesky and is never reached, models.py:58,99,101).  Elementwise torch op on
device tensors, offered for API parity; the ELBO / predict_samples kernels fuse it."""
from __future__ import annotations

JITTER = 1e-6


def reparameterize(mean, var, z, full_cov=False):
    if var is None:
        return mean
    if full_cov:
        raise NotImplementedError("full_cov=True is dead code in the reference (tf.cholesky, TF1 API)")
    return mean + z * (var + JITTER) ** 0.5
