"""TEST INFRASTRUCTURE ONLY (oracle shim): InducingPoints [3P-memory]."""
import numpy as np

from .base import Module, Parameter


class InducingPoints(Module):
    def __init__(self, Z, name=None):
        self.Z = Z if isinstance(Z, Parameter) else Parameter(np.asarray(Z, dtype=np.float64))

    @property
    def num_inducing(self):
        return int(self.Z.shape[0])


def inducingpoint_wrapper(iv):
    return iv if isinstance(iv, InducingPoints) else InducingPoints(iv)
