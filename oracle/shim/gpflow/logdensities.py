"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.logdensities.gaussian [3P-memory]."""
import numpy as np
import tensorflow as tf


def gaussian(x, mu, var):
    return -0.5 * (np.log(2 * np.pi) + tf.math.log(var) + tf.square(tf._t(mu) - tf._t(x)) / tf._t(var))
