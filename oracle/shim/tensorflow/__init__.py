"""TEST INFRASTRUCTURE ONLY (oracle).  A tiny torch-fp64-backed stand-in for the subset of the
TensorFlow 2.10 API that /root/reference/MixtureGPs and the GPflow-2.7 restatement under
oracle/shim/gpflow call.  It exists so that the reference's OWN model files
(MixtureGPs/models.py, likelihoods.py, broadcasting_lik.py, utils.py) can be imported and executed
UNMODIFIED in a container where TensorFlow is not installable; autograd comes from torch.

Never imported by the product package (modulatedgps_b200).  Semantics restated from TF 2.10.1
(environment.yml:130 of the reference) — [3P-memory], see SURVEY.md Appendix A.
"""
import numpy as _np
import torch as _torch

float64 = _torch.float64
float32 = _torch.float32
int64 = _torch.int64
int32 = _torch.int32
Tensor = _torch.Tensor
newaxis = None


class _TFTensor(_torch.Tensor):
    """tf.Tensor is immutable: `x *= y` in the reference (models.py:66,115,120) REBINDS x to a new,
    broadcast tensor.  torch's in-place dunders would mutate (and refuse to broadcast), so map them to
    the out-of-place ops.  torch ops on this subclass return the subclass, so the property propagates."""
    def __imul__(self, o):
        return self * o

    def __iadd__(self, o):
        return self + o

    def __isub__(self, o):
        return self - o

    def __itruediv__(self, o):
        return self / o


def _wrap(x):
    return x if isinstance(x, _TFTensor) else x.as_subclass(_TFTensor)


def _t(x, dtype=None):
    """Anything (Parameter, ndarray, scalar, list) -> torch tensor (immutable-style subclass)."""
    if hasattr(x, "_shim_value"):
        x = x._shim_value()
    if isinstance(x, _torch.Tensor):
        return _wrap(x if dtype is None else x.to(dtype))
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, _torch.Tensor) for e in x):
        return _wrap(_torch.stack([_t(e) for e in x]))
    a = _np.asarray(x)
    if a.dtype == _np.float32 or a.dtype == _np.float64 or a.dtype.kind == "f":
        return _wrap(_torch.as_tensor(a, dtype=dtype or _torch.float64))
    if a.dtype.kind == "b":
        return _wrap(_torch.as_tensor(a))
    return _wrap(_torch.as_tensor(a, dtype=dtype or _torch.int64))


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def constant(x, dtype=None):
    return _t(x, dtype)


class _Shape(tuple):
    """tf.shape(x) result: indexable ints (eager mode), sliceable."""
    def __getitem__(self, i):
        r = tuple.__getitem__(self, i)
        return _Shape(r) if isinstance(i, slice) else r


def shape(x):
    return _Shape(int(s) for s in _t(x).shape)


def rank(x):
    return _t(x).dim()


def size(x, out_type=None):
    return _t(x).numel()


def _ints(shp):
    return [int(s) for s in shp]


def reshape(x, shp):
    return _t(x).reshape(_ints(shp))


def tile(x, multiples):
    return _t(x).repeat(*_ints(multiples))


def cast(x, dtype):
    return _t(x).to(dtype)


def identity(x):
    return _t(x).clone()


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def squeeze(x, axis=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(axis)


def broadcast_to(x, shp):
    return _t(x).broadcast_to(_ints(shp))


def fill(dims, value):
    v = _t(value)
    return v.reshape(()).expand(_ints(dims)) if v.numel() == 1 else v.expand(_ints(dims))


def transpose(x, perm=None):
    x = _t(x)
    if perm is None:
        perm = list(range(x.dim()))[::-1]
    return x.permute(*_ints(perm))


def concat(values, axis):
    return _torch.cat([_t(v).reshape(-1) if _t(v).dim() == 0 else _t(v) for v in values], dim=axis)


def stack(values, axis=0):
    return _torch.stack([_t(v) for v in values], dim=axis)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_prod(x, axis=None, keepdims=False):
    x = _t(x)
    if axis is None:
        return x.prod()
    if isinstance(axis, (list, tuple)):
        for a in sorted(axis, reverse=True):
            x = x.prod(dim=a, keepdim=keepdims)
        return x
    return x.prod(dim=axis, keepdim=keepdims)


def reduce_logsumexp(x, axis=None, keepdims=False):
    # TF: max-subtracted, max under stop_gradient; torch.logsumexp is the same function and gradient.
    x = _t(x)
    if axis is None:
        return _torch.logsumexp(x.reshape(-1), 0)
    return _torch.logsumexp(x, dim=axis, keepdim=keepdims)


def square(x):
    x = _t(x)
    return x * x


def sqrt(x):
    return _torch.sqrt(_t(x))


def exp(x):
    return _torch.exp(_t(x))


def add(a, b):
    return _t(a) + _t(b)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def tensordot(a, b, axes):
    return _torch.tensordot(_t(a), _t(b), dims=(list(axes[0]), list(axes[1])))


def eye(n, dtype=float64):
    return _torch.eye(int(n), dtype=dtype)


def ones(shp, dtype=float64):
    return _torch.ones(_ints(shp), dtype=dtype)


def zeros(shp, dtype=float64):
    return _torch.zeros(_ints(shp), dtype=dtype)


def one_hot(indices, depth, on_value=1.0, off_value=0.0, dtype=None):
    idx = _t(indices).to(_torch.int64)
    oh = _torch.nn.functional.one_hot(idx, int(depth)).to(_torch.float64)
    return oh * float(on_value) + (1.0 - oh) * float(off_value)


def clip_by_value(x, lo, hi):
    return _torch.clamp(_t(x), min=lo, max=hi)


def argmax(x, axis=None):
    return _torch.argmax(_t(x), dim=axis)


def equal(a, b):
    return _t(a) == _t(b)


def where(c, a, b):
    return _torch.where(_t(c), _t(a), _t(b))


def range(*a):  # noqa: A001  (tf.range)
    return _torch.arange(*a)


class _NoiseSource:
    """Noise for tf.random.normal / tfp uniform: either queued arrays (explicit z, u — parity mode)
    or a seeded numpy Generator.  TF's own Philox stream is not reproducible outside TF."""
    def __init__(self):
        self.queue = []
        self.rng = _np.random.default_rng(0)
        self.log = []

    def push(self, *arrays):
        self.queue.extend(arrays)

    def clear(self):
        self.queue.clear()
        self.log.clear()

    def draw(self, kind, shp):
        shp = tuple(_ints(shp))
        if self.queue:
            a = _np.asarray(self.queue.pop(0), dtype=_np.float64)
            assert a.size == int(_np.prod(shp)), f"queued {kind} noise has shape {a.shape}, wanted {shp}"
            a = a.reshape(shp)
        elif kind == "normal":
            a = self.rng.standard_normal(shp)
        else:
            tiny = _np.finfo(_np.float64).tiny
            a = self.rng.uniform(tiny, 1.0, shp)
        self.log.append((kind, shp))
        return _wrap(_torch.as_tensor(a, dtype=_torch.float64))


_noise = _NoiseSource()


class random:  # noqa: N801  (tf.random namespace)
    @staticmethod
    def normal(shp, mean=0.0, stddev=1.0, dtype=float64, seed=None):
        return _noise.draw("normal", shp) * stddev + mean

    @staticmethod
    def uniform(shp, minval=0.0, maxval=1.0, dtype=float64, seed=None):
        return _noise.draw("uniform", shp)

    @staticmethod
    def set_seed(seed):
        _noise.rng = _np.random.default_rng(seed)


class math:  # noqa: N801  (tf.math namespace)
    @staticmethod
    def log(x):
        return _torch.log(_t(x))

    @staticmethod
    def exp(x):
        return _torch.exp(_t(x))

    @staticmethod
    def erf(x):
        return _torch.erf(_t(x))

    @staticmethod
    def log_softmax(x, axis=-1):
        return _torch.log_softmax(_t(x), dim=axis)

    @staticmethod
    def softplus(x):
        return _torch.nn.functional.softplus(_t(x))

    @staticmethod
    def sigmoid(x):
        return _torch.sigmoid(_t(x))

    square = staticmethod(square)
    sqrt = staticmethod(sqrt)
    reduce_sum = staticmethod(reduce_sum)
    reduce_logsumexp = staticmethod(reduce_logsumexp)


class nn:  # noqa: N801  (tf.nn namespace)
    @staticmethod
    def softmax(x, axis=-1):
        return _torch.softmax(_t(x), dim=axis)

    @staticmethod
    def log_softmax(x, axis=-1):
        return _torch.log_softmax(_t(x), dim=axis)

    @staticmethod
    def softplus(x):
        return _torch.nn.functional.softplus(_t(x))


class linalg:  # noqa: N801  (tf.linalg namespace)
    @staticmethod
    def cholesky(a):
        return _torch.linalg.cholesky(_t(a))

    @staticmethod
    def triangular_solve(matrix, rhs, lower=True, adjoint=False):
        m = _t(matrix)
        if adjoint:
            m, lower = m.transpose(-1, -2), not lower
        return _torch.linalg.solve_triangular(m, _t(rhs), upper=not lower)

    @staticmethod
    def band_part(x, num_lower, num_upper):
        x = _t(x)
        if num_lower == -1 and num_upper == 0:
            return _torch.tril(x)
        if num_lower == 0 and num_upper == -1:
            return _torch.triu(x)
        raise NotImplementedError

    @staticmethod
    def adjoint(x):
        return _t(x).transpose(-1, -2)

    @staticmethod
    def diag_part(x):
        return _torch.diagonal(_t(x), dim1=-2, dim2=-1)

    matmul = staticmethod(matmul)
    eye = staticmethod(eye)


def zeros_like(x):
    return _wrap(_torch.zeros_like(_t(x)))


def stop_gradient(x):
    return _wrap(_t(x).detach())


def py_function(func, inp, Tout):
    """tf.py_function: run `func` eagerly on the inputs (the shim is always eager)."""
    out = func(*[_t(a) for a in inp])
    return [_t(o) for o in out] if isinstance(out, (list, tuple)) else _t(out)


class GradientTape:
    """The slice of tf.GradientTape modulatedgps_b200.tf_adapter uses: watch + gradient(target, source,
    output_gradients=) as a vector-Jacobian product (torch.autograd.grad underneath)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        self._prev = _torch.is_grad_enabled()
        _torch.set_grad_enabled(True)
        return self

    def __exit__(self, *exc):
        _torch.set_grad_enabled(self._prev)
        return False

    def watch(self, x):
        assert _t(x).requires_grad, "the shim can only watch tensors that already require grad"

    def gradient(self, target, sources, output_gradients=None):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        tgt = _t(target)
        if not tgt.requires_grad:                    # identity bijector: target IS the source
            g = [_t(output_gradients) if tgt is _t(s) or tgt.data_ptr() == _t(s).data_ptr() else None for s in srcs]
        else:
            go = None if output_gradients is None else _t(output_gradients)
            # (a source that already is a torch tensor is passed AS IS: re-wrapping it makes an alias that is not the
            #  leaf the graph was recorded on)
            src_t = [s if isinstance(s, _torch.Tensor) else _t(s) for s in srcs]
            g = _torch.autograd.grad(tgt, src_t, grad_outputs=go, retain_graph=True, allow_unused=True)
        g = [None if e is None else _wrap(e) for e in g]
        return g[0] if single else g


def custom_gradient(f):
    """tf.custom_gradient: f(*args) -> (result, grad_fn); grad_fn(upstream) -> one gradient (or None) per arg."""
    def wrapped(*args):
        targs = tuple(_t(a) for a in args)
        holder = {}

        class _Fn(_torch.autograd.Function):
            @staticmethod
            def forward(ctx, *xs):
                with _torch.enable_grad():           # inner tapes differentiate pieces of the forward
                    y, grad_fn = f(*xs)
                holder["grad_fn"] = grad_fn
                return _t(y).detach().clone()

            @staticmethod
            def backward(ctx, upstream):
                g = holder["grad_fn"](_wrap(upstream))
                return tuple(None if e is None else _t(e) for e in g)

        return _wrap(_Fn.apply(*targs))
    return wrapped


class experimental:  # noqa: N801  (tf.experimental namespace)
    class dlpack:  # noqa: N801
        @staticmethod
        def to_dlpack(x):
            return _torch.utils.dlpack.to_dlpack(_t(x).detach().as_subclass(_torch.Tensor).contiguous())

        @staticmethod
        def from_dlpack(capsule):
            return _wrap(_torch.utils.dlpack.from_dlpack(capsule))


def function(fn=None, **kw):
    """tf.function: eager pass-through."""
    if fn is None:
        return lambda f: f
    return fn


class Module:
    """tf.Module: attribute-walk for trainable variables is done by gpflow.base.Module in the shim."""
    def __init__(self, name=None):
        pass


class test:  # noqa: N801
    @staticmethod
    def is_built_with_cuda():
        return False


class config:  # noqa: N801
    @staticmethod
    def list_physical_devices(kind=None):
        return []
