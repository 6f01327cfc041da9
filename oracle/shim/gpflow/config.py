"""TEST INFRASTRUCTURE ONLY (oracle shim): gpflow.config defaults [3P-memory, gpflow 2.7.0]."""
import torch


def default_float():
    return torch.float64


def default_int():
    return torch.int64


def default_jitter():
    return 1e-6
