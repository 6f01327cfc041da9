"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.covariances.Kuu for InducingPoints x Kernel
[3P-memory, gpflow 2.7.0 gpflow/covariances/kuus.py]; SURVEY.md Appendix A.2."""
import tensorflow as tf


def Kuu(inducing_variable, kernel, *, jitter=0.0):
    Kzz = kernel(inducing_variable.Z)
    Kzz = Kzz + jitter * tf.eye(inducing_variable.num_inducing, dtype=Kzz.dtype)
    return Kzz


def Kuf(inducing_variable, kernel, Xnew):
    return kernel(inducing_variable.Z, Xnew)
