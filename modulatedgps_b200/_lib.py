"""ctypes binding of libmgp.so (include/mgp.h).  The only way the Python side reaches the GPU arithmetic.

There is no CPU fallback: a missing library or a machine without a CUDA device raises immediately.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmgp.so")

(MGP_OK, MGP_ERR_BAD_ARG, MGP_ERR_CUDA, MGP_ERR_NOT_PD, MGP_ERR_NOMEM, MGP_ERR_STALE_PRECOMPUTE,
 MGP_ERR_NCCL) = 0, 1, 2, 3, 4, 5, 6
ROBUSTMAX_CDF_SQUASH = 1e-4      # MGP_ROBUSTMAX_CDF_SQUASH of include/mgp.h (tests/test_host_logic.py keeps them equal)
MODEL_SMGP, MODEL_SMGP_MODIFIED = 0, 1
LIK_GAUSSIAN, LIK_MULTICLASS = 0, 1
MAX_K, MAX_D = 8, 32

_dp = C.POINTER(C.c_double)


class MgpLayer(C.Structure):
    _fields_ = [("M", C.c_int32), ("D", C.c_int32), ("K", C.c_int32), ("n_lengthscales", C.c_int32),
                ("Z", C.c_void_p), ("q_mu", C.c_void_p), ("q_sqrt", C.c_void_p), ("variance", C.c_void_p),
                ("lengthscales", C.c_void_p)]


class MgpLayerGrad(C.Structure):
    _fields_ = [("Z", C.c_void_p), ("q_mu", C.c_void_p), ("q_sqrt", C.c_void_p), ("variance", C.c_void_p),
                ("lengthscales", C.c_void_p)]


class MgpNoise(C.Structure):
    _fields_ = [("z", C.c_void_p), ("u", C.c_void_p), ("seed", C.c_uint64), ("point_offset", C.c_int64)]


class MgpElboCfg(C.Structure):
    _fields_ = [("model", C.c_int32), ("lik", C.c_int32), ("S", C.c_int32), ("reserved", C.c_int32),
                ("temperature", C.c_double), ("num_data", C.c_double), ("n_global", C.c_int64)]


class MgpAdamSlot(C.Structure):
    _fields_ = [("theta", C.c_void_p), ("grad", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("gather", C.c_void_p),
                ("n", C.c_int64), ("transform", C.c_int32), ("reserved", C.c_int32)]


TRANSFORM_IDENTITY, TRANSFORM_SOFTPLUS = 0, 1
ADAM_MAX_SLOTS = 16


class MgpError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmgp error {code}: {message}")
        self.code = code


class NotPositiveDefiniteError(MgpError):
    """Cholesky of Kuu failed (the reference surfaces this as a TF InvalidArgumentError)."""


class StalePrecomputeError(MgpError):
    """Parameter values were changed in place between mgp_elbo_local and mgp_elbo_finish."""


_lib = None

# every exported symbol of include/mgp.h: (restype, argtypes)
_PROTOTYPES = {
    "mgp_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "mgp_ctx_destroy": (None, [C.c_void_p]),
    "mgp_last_error": (C.c_char_p, [C.c_void_p]),
    "mgp_launch_count": (C.c_int64, [C.c_void_p]),
    "mgp_check_status": (C.c_int, [C.c_void_p]),
    "mgp_set_chunk_points": (C.c_int, [C.c_void_p, C.c_int64]),
    "mgp_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mgp_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "mgp_num_stages": (C.c_int, []),
    "mgp_stage_name": (C.c_char_p, [C.c_int]),
    "mgp_svgp_predict_f": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mgp_prior_kl": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_void_p]),
    "mgp_predict_y": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_void_p, C.c_void_p]),
    "mgp_predict_assign": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "mgp_predict_samples": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.POINTER(MgpLayer), C.c_int32, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.POINTER(MgpNoise), C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "mgp_w_sample": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_void_p, C.c_int64, C.c_int32, C.c_double,
                               C.POINTER(MgpNoise), C.c_void_p]),
    "mgp_e_log_p_y": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_int32, C.c_void_p, C.c_void_p]),
    "mgp_reduce_buffer_len": (C.c_int64, [C.POINTER(MgpLayer), C.POINTER(MgpLayer)]),
    "mgp_elbo_local": (C.c_int, [C.c_void_p, C.POINTER(MgpElboCfg), C.POINTER(MgpLayer), C.POINTER(MgpLayer),
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(MgpNoise),
                                 C.c_void_p]),
    "mgp_elbo_finish": (C.c_int, [C.c_void_p, C.POINTER(MgpElboCfg), C.POINTER(MgpLayer), C.POINTER(MgpLayer),
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(MgpLayerGrad),
                                  C.POINTER(MgpLayerGrad), C.c_void_p, C.c_void_p]),
    "mgp_elbo_fwd_bwd": (C.c_int, [C.c_void_p, C.POINTER(MgpElboCfg), C.POINTER(MgpLayer), C.POINTER(MgpLayer),
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(MgpNoise),
                                   C.c_void_p, C.POINTER(MgpLayerGrad), C.POINTER(MgpLayerGrad), C.c_void_p,
                                   C.c_void_p]),
    "mgp_lik_variational_expectations": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "mgp_lik_predict_mean_and_var": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                               C.c_int32, C.c_void_p, C.c_void_p]),
    "mgp_lik_log_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                   C.c_void_p]),
    "mgp_lik_predict_log_density": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                              C.c_int64, C.c_int32, C.c_void_p]),
    "mgp_debug_philox": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "mgp_debug_noise": (C.c_int, [C.c_void_p, C.POINTER(MgpNoise), C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p]),
    "mgp_debug_kuu_chol": (C.c_int, [C.c_void_p, C.POINTER(MgpLayer), C.c_void_p, C.c_void_p, C.c_void_p]),
    "mgp_adam_step": (C.c_int, [C.c_void_p, C.POINTER(MgpAdamSlot), C.c_int32, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "mgp_set_robustmax_squash": (C.c_int, [C.c_void_p, C.c_double]),
    "mgp_ctx_set_comm": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgp_all_reduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "mgp_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                  C.c_void_p]),
    "mgp_fill_triangular": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]),
    "mgp_kmeans_iterate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def exported_symbols():
    return sorted(_PROTOTYPES)


def load_library():
    """dlopen libmgp.so and attach prototypes.  Does not need a GPU (used by the CPU test-suite to check the
    exported symbol table); any compute call does."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m modulatedgps_b200.build` "
            "(nvcc, sm_100a).  modulatedgps_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


_contexts: Dict[Tuple[int, int], "Context"] = {}
MAX_CONTEXTS_PER_DEVICE = 2


class Context:
    """One mgp_ctx per (device, stream)."""

    def __init__(self, device: int, stream_ptr: int):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("modulatedgps_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        h = C.c_void_p()
        rc = self.lib.mgp_ctx_create(int(device), C.c_void_p(stream_ptr), C.byref(h))
        if rc != MGP_OK or not h.value:
            raise MgpError(rc, "mgp_ctx_create failed (no usable CUDA device?)")
        self.handle = h
        self.device = device
        self.comm = None

    def check(self, rc: int):
        if rc == MGP_OK:
            return
        msg = (self.lib.mgp_last_error(self.handle) or b"").decode()
        if rc == MGP_ERR_NOT_PD:
            raise NotPositiveDefiniteError(rc, msg)
        if rc == MGP_ERR_STALE_PRECOMPUTE:
            raise StalePrecomputeError(rc, msg)
        raise MgpError(rc, msg)

    def check_status(self):
        self.check(self.lib.mgp_check_status(self.handle))

    @property
    def launch_count(self) -> int:
        return int(self.lib.mgp_launch_count(self.handle))

    def timing_enable(self, on: bool):
        self.check(self.lib.mgp_timing_enable(self.handle, 1 if on else 0))

    def timing_read(self, reset: bool = True):
        """{stage name: (milliseconds, launches-groups)} accumulated since the last reset (synchronises)."""
        n = int(self.lib.mgp_num_stages())
        ms = (C.c_double * n)()
        calls = (C.c_int64 * n)()
        self.check(self.lib.mgp_timing_read(self.handle, ms, calls, 1 if reset else 0))
        return {self.lib.mgp_stage_name(i).decode(): (float(ms[i]), int(calls[i])) for i in range(n)}

    def set_robustmax_squash(self, squash: float):
        self.check(self.lib.mgp_set_robustmax_squash(self.handle, float(squash)))

    def set_comm(self, nccl_comm):
        """Attach (or, with None, detach) an ncclComm_t: mgp_elbo_local / mgp_elbo_fwd_bwd then all-reduce by themselves."""
        self.check(self.lib.mgp_ctx_set_comm(self.handle, C.c_void_p(nccl_comm) if nccl_comm else None))
        self.comm = nccl_comm

    def set_chunk_points(self, n: int):
        self.check(self.lib.mgp_set_chunk_points(self.handle, int(n)))

    def __del__(self):
        try:
            if getattr(self, "handle", None) is not None and self.handle.value:
                self.lib.mgp_ctx_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass


def get_context(device: torch.device | int | None = None) -> Context:
    if not torch.cuda.is_available():
        load_library()
        raise RuntimeError("modulatedgps_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        dev = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        dev = device.index if device.index is not None else torch.cuda.current_device()
    else:
        dev = int(device)
    stream_ptr = int(torch.cuda.current_stream(dev).cuda_stream)
    key = (dev, stream_ptr)
    ctx = _contexts.pop(key, None)
    if ctx is None:
        # a context owns a workspace of up to 40 % of HBM: keep at most MAX_CONTEXTS_PER_DEVICE alive per device and
        # drop the least recently used one (a recycled stream handle then gets a fresh context, never a stale one)
        mine = [k for k in _contexts if k[0] == dev]
        while len(mine) >= MAX_CONTEXTS_PER_DEVICE:
            _contexts.pop(mine.pop(0))
        with torch.cuda.device(dev):
            ctx = Context(dev, stream_ptr)
    _contexts[key] = ctx                     # (re)inserted last: dict order is the LRU order
    return ctx


def total_launches() -> int:
    return sum(c.launch_count for c in _contexts.values())


def ptr(t) -> C.c_void_p:
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())
