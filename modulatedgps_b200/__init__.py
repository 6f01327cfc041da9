"""modulatedgps_b200 — B200-native (sm_100a) SVGP-mixture / data-association GP hot path with the class API
of LouieMiddle/ModulatedGPs' `MixtureGPs` package.  All arithmetic is hand-written CUDA behind the C-ABI in
include/mgp.h (libmgp.so, bound with ctypes); there is no CPU fallback.

    from modulatedgps_b200 import SquaredExponential, GaussianModified, SVGPModified, SMGP, run_adam
"""
from . import _lib
from ._lib import MgpError, NotPositiveDefiniteError
from .broadcasting_lik import BroadcastingLikelihood
from .kernels import RBF, SquaredExponential
from .likelihoods import GaussianModified, MultiClass, RobustMax
from .models import SGP, SMGP, DeviceArray, InducingPoints, SMGPModified, SVGPModified
from .parameter import Module, Parameter, print_summary
from .training import DeviceMinibatches, FusedAdam, HostBatchStream, kmeans, make_adam, predict_samples_batched, run_adam
from .utils import reparameterize

__all__ = ["SquaredExponential", "RBF", "GaussianModified", "MultiClass", "RobustMax", "BroadcastingLikelihood",
           "SVGPModified", "SGP", "SMGP", "SMGPModified", "InducingPoints", "Parameter", "Module", "print_summary",
           "run_adam", "make_adam", "FusedAdam", "DeviceMinibatches", "HostBatchStream", "kmeans", "predict_samples_batched", "reparameterize", "MgpError", "NotPositiveDefiniteError", "DeviceArray"]
