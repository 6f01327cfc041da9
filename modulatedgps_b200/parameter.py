"""Parameters with GPflow's bijectors, held as torch CUDA float64 leaves.

Mirrors gpflow.Parameter as the reference uses it (MixtureGPs/likelihoods.py:19; gpflow kernel / SVGP
parameters): positive() = softplus, triangular() = TFP FillTriangular (a pure permutation).
The optimiser sees the UNCONSTRAINED leaves (`model.trainable_variables`), exactly as TF does.
"""
from __future__ import annotations

import numpy as np
import torch

F64 = torch.float64


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("modulatedgps_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_device_f64(x, device=None) -> torch.Tensor:
    """numpy / torch / anything exporting __dlpack__ (e.g. a TF tensor via tf.experimental.dlpack) ->
    contiguous float64 tensor on the device."""
    device = device or default_device()
    if isinstance(x, Parameter):
        x = x.value()
    if isinstance(x, torch.Tensor):
        t = x
    elif hasattr(x, "__dlpack__") and not isinstance(x, np.ndarray):
        t = torch.from_dlpack(x)
    else:
        t = torch.as_tensor(np.asarray(x))
    return t.to(device=device, dtype=F64, non_blocking=True).contiguous()


def fill_triangular_index(m: int) -> np.ndarray:
    """TFP fill_triangular (lower): vector x of length m(m+1)/2 -> concat([x[m:], reversed(x)]) reshaped [m, m],
    lower band.  Returns the [m, m] map of source indices on the lower triangle (-1 above it).
    fill_triangular([1,2,3,4,5,6]) == [[4,0,0],[6,5,0],[3,2,1]]."""
    n = m * (m + 1) // 2
    x = np.arange(n)
    mat = np.concatenate([x[m:], x[::-1]]).reshape(m, m)
    return np.where(np.tril(np.ones((m, m), dtype=bool)), mat, -1)


class Identity:
    name = "identity"

    def forward(self, x):
        return x

    def inverse(self, y):
        return y

    def grad_to_unconstrained(self, g_constrained, x_unconstrained):
        return g_constrained

    def forward_shape(self, shape):
        return tuple(shape)


class Softplus:
    """gpflow.utilities.positive() with the default softplus bijector and zero lower bound."""
    name = "softplus"

    def forward(self, x):
        return torch.nn.functional.softplus(x)

    def inverse(self, y):
        return y + torch.log(-torch.expm1(-y))

    def grad_to_unconstrained(self, g_constrained, x_unconstrained):
        return g_constrained * torch.sigmoid(x_unconstrained)

    def forward_shape(self, shape):
        return tuple(shape)


class FillTriangular:
    """gpflow.utilities.triangular(): [.., m(m+1)/2] <-> lower-triangular [.., m, m]."""
    name = "fill_triangular"

    def __init__(self):
        self._cache = {}

    def _maps(self, m, device):
        key = (m, str(device))
        if key not in self._cache:
            idx = fill_triangular_index(m)
            ii, jj = np.nonzero(idx >= 0)
            flat_pos = torch.as_tensor(ii * m + jj, device=device)            # positions in the flattened matrix
            src = torch.as_tensor(idx[ii, jj], device=device)                 # which vector entry lands there
            self._cache[key] = (flat_pos, src)
        return self._cache[key]

    @staticmethod
    def _kernel(src, m, out, inverse):
        """One libmgp launch (mgp_fill_triangular) for device tensors outside autograd: the index / index_put route
        below costs three torch kernels per call, and the step evaluates this bijector four times."""
        import ctypes as C
        from . import _lib
        lib = _lib.load_library()
        batch = src.numel() // (m * m if inverse else m * (m + 1) // 2)
        rc = lib.mgp_fill_triangular(C.c_void_p(torch.cuda.current_stream(src.device).cuda_stream), _lib.ptr(src), batch, m,
                                     _lib.ptr(out), 1 if inverse else 0)
        if rc != 0:
            raise _lib.MgpError(rc, "mgp_fill_triangular failed")
        return out

    @staticmethod
    def _use_kernel(t):
        return t.is_cuda and t.dtype == F64 and not (torch.is_grad_enabled() and t.requires_grad)

    def forward(self, x):
        n = x.shape[-1]
        m = int(round((np.sqrt(8 * n + 1) - 1) / 2))
        if self._use_kernel(x):
            out = torch.empty(*x.shape[:-1], m, m, dtype=x.dtype, device=x.device)
            return self._kernel(x.contiguous(), m, out, False)
        flat_pos, src = self._maps(m, x.device)
        out = torch.zeros(*x.shape[:-1], m * m, dtype=x.dtype, device=x.device)
        out[..., flat_pos] = x[..., src]
        return out.reshape(*x.shape[:-1], m, m)

    def inverse(self, y):
        m = y.shape[-1]
        if self._use_kernel(y):
            out = torch.empty(*y.shape[:-2], m * (m + 1) // 2, dtype=y.dtype, device=y.device)
            return self._kernel(y.contiguous(), m, out, True)
        flat_pos, src = self._maps(m, y.device)
        out = torch.zeros(*y.shape[:-2], m * (m + 1) // 2, dtype=y.dtype, device=y.device)
        out[..., src] = y.reshape(*y.shape[:-2], m * m)[..., flat_pos]
        return out

    def grad_to_unconstrained(self, g_constrained, x_unconstrained):
        return self.inverse(g_constrained)

    def forward_shape(self, shape):
        n = shape[-1]
        m = int(round((np.sqrt(8 * n + 1) - 1) / 2))
        return tuple(shape[:-1]) + (m, m)


class Parameter:
    def __init__(self, value, transform=None, trainable=True, name=None, device=None):
        self.transform = transform or Identity()
        self.trainable = bool(trainable)
        self.name = name
        v = to_device_f64(value, device)
        self.unconstrained_variable = self.transform.inverse(v).detach().clone().requires_grad_(self.trainable)

    def value(self) -> torch.Tensor:
        with torch.no_grad():
            return self.transform.forward(self.unconstrained_variable).contiguous()

    def assign(self, value):
        v = to_device_f64(value, self.unconstrained_variable.device)
        with torch.no_grad():
            self.unconstrained_variable.copy_(self.transform.inverse(v))

    def numpy(self):
        return self.value().cpu().numpy()

    def __array__(self, dtype=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    @property
    def shape(self):
        """Shape of the constrained value (no kernel is launched to find it)."""
        return torch.Size(self.transform.forward_shape(self.unconstrained_variable.shape))


class Module:
    """Attribute walk in the manner of tf.Module: descends into Modules, lists, tuples and dicts only (so the
    reference's plain-class BroadcastingLikelihood is not walked: SURVEY.md §3.1)."""

    @property
    def parameters_dict(self):
        out, seen = {}, set()
        _walk(self, "", seen, out)
        return out

    @property
    def trainable_parameters(self):
        return tuple(p for p in self.parameters_dict.values() if p.trainable)

    @property
    def trainable_variables(self):
        return tuple(p.unconstrained_variable for p in self.trainable_parameters)


def _walk(obj, prefix, seen, out):
    if isinstance(obj, Parameter):
        if id(obj) not in seen:
            seen.add(id(obj))
            out[prefix] = obj
    elif isinstance(obj, Module):
        if id(obj) in seen:
            return
        seen.add(id(obj))
        for name in sorted(vars(obj)):
            _walk(vars(obj)[name], f"{prefix}.{name}" if prefix else name, seen, out)
    elif isinstance(obj, (list, tuple)):
        for i, e in enumerate(obj):
            _walk(e, f"{prefix}[{i}]", seen, out)
    elif isinstance(obj, dict):
        for k, e in obj.items():
            _walk(e, f"{prefix}[{k}]", seen, out)


def print_summary(module, fmt=None):
    """gpflow.utilities.print_summary look-alike (demos/demo_tf2.py:51,60)."""
    print(f"{'name':52s} {'transform':16s} {'trainable':9s} {'shape':14s} value")
    for k, p in module.parameters_dict.items():
        v = p.numpy()
        s = np.array2string(v.reshape(-1)[:4], precision=5)
        print(f"{k:52s} {p.transform.name:16s} {str(p.trainable):9s} {str(tuple(v.shape)):14s} {s}{'...' if v.size > 4 else ''}")
