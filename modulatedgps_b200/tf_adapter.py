"""TF2 / GPflow adapter: run the reference's OWN `MixtureGPs` model objects on libmgp.

This is the binding a maintainer of LouieMiddle/ModulatedGPs adds to keep GPflow `Parameter`s, `tf.function`,
`tf.optimizers.Adam` and `run_adam` (utils/training_utils.py:4-28) exactly as they are, while the arithmetic behind
`SMGP._build_likelihood` (MixtureGPs/models.py:69-79, and SMGPModified.E_log_p_Y :112-123) moves to the B200 kernels:

    import modulatedgps_b200.tf_adapter as mgp_tf
    model = SMGP(likelihood=lik, pred_layer=pred_layer, assign_layer=assign_layer, K=K, ...)   # reference classes
    mgp_tf.attach(model)                                  # model._training_loss / _build_likelihood now call libmgp
    iters, elbos = run_adam(model, num_iter, train_iter, lr, compile=False)                     # unchanged

How it is wired (SURVEY.md §8b "Python side"):
  * `tf.custom_gradient` node `neg_elbo(X, Y, *unconstrained_variables)`: the forward makes ONE `mgp_elbo_fwd_bwd` call
    (ELBO and every gradient w.r.t. CONSTRAINED values), wrapped in `tf.py_function` so that it also runs inside the
    `@tf.function` of the reference's `optimization_step`; the backward is the bijector chain rule (GPflow `positive()`
    = softplus, `triangular()` = FillTriangular, identity), taken as a vector-Jacobian product through
    `Parameter.transform.forward` with a `tf.GradientTape`, returned in the order of the variables passed in;
  * tensors cross the boundary as DLPack capsules (`tf.experimental.dlpack.to_dlpack` -> `torch.from_dlpack` ->
    `data_ptr()`; zero-copy when TF holds them on the GPU), torch being only the owner of device memory here;
  * the C structs are the ctypes mirrors of include/mgp.h in `_lib.py`.

TensorFlow / GPflow are imported lazily and only through the names `tensorflow` / `gpflow` resolve to at call time:
this image cannot install them, so the CPU test-suite runs this module against the torch-backed stand-ins under
the test tree (struct marshalling, gradient ordering, chain rule) and the GPU suite runs it end to end through
torch <-> DLPack against the golden vectors.  Importing this module needs neither TF nor a GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Sequence, Tuple

from . import _lib

LAYER_KEYS = ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")
TEMPERATURE = 1e-2      # literal at MixtureGPs/models.py:60


def _tf():
    import tensorflow as tf
    return tf


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------------------------
# tensor hand-off
# ---------------------------------------------------------------------------------------------------------------
def to_torch(t, device=None, with_origin=False):
    """TF tensor -> torch tensor through a DLPack capsule (no copy when it already lives on `device`).
    with_origin=True also returns the device the TF tensor lived on, so that results can be handed back there."""
    torch = _torch()
    tf = _tf()
    x = torch.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(t)))
    origin = x.device
    if x.dtype != torch.float64:
        x = x.to(torch.float64)
    if device is not None and x.device != device:
        x = x.to(device)
    x = x.contiguous()
    return (x, origin) if with_origin else x


def to_tf(x):
    """torch tensor -> TF tensor through a DLPack capsule."""
    torch = _torch()
    return _tf().experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(x.contiguous()))


# ---------------------------------------------------------------------------------------------------------------
# marshalling into the C structs of include/mgp.h
# ---------------------------------------------------------------------------------------------------------------
class Marshalled:
    """Everything one mgp_elbo_fwd_bwd call points at, kept alive together."""

    def __init__(self):
        self.keep: List = []
        self.cfg = self.pred = self.assign = self.pred_grad = self.assign_grad = self.noise = None
        self.lik_var = self.assign_lik_var = self.X = self.Y = None
        self.elbo = self.glik = self.galik = None
        self.grads: Dict[str, object] = {}
        self.N = 0


def layer_struct(t: Dict[str, object], m: Marshalled) -> Tuple[_lib.MgpLayer, _lib.MgpLayerGrad, Dict[str, object]]:
    """t: constrained torch tensors of one SVGPModified layer (variance [] or [1], lengthscales [] or [D], Z [M, D],
    q_mu [M, K], q_sqrt [K, M, M]) -> (mgp_layer, mgp_layer_grad, gradient buffers)."""
    torch = _torch()
    var, ls = t["variance"].reshape(1).contiguous(), t["lengthscales"].reshape(-1).contiguous()
    Z, q_mu, q_sqrt = t["Z"].contiguous(), t["q_mu"].contiguous(), t["q_sqrt"].contiguous()
    if Z.dim() != 2 or q_mu.dim() != 2 or q_sqrt.dim() != 3:
        raise ValueError("layer tensors must be Z [M, D], q_mu [M, K], q_sqrt [K, M, M]")
    Mn, D = Z.shape
    K = q_mu.shape[1]
    if q_mu.shape[0] != Mn or tuple(q_sqrt.shape) != (K, Mn, Mn):
        raise ValueError(f"q_mu {tuple(q_mu.shape)} / q_sqrt {tuple(q_sqrt.shape)} do not match Z {tuple(Z.shape)}")
    if ls.numel() not in (1, D):
        raise ValueError(f"lengthscales has {ls.numel()} entries; expected 1 or D = {D}")
    g = {"variance": torch.empty_like(var), "lengthscales": torch.empty_like(ls), "Z": torch.empty_like(Z),
         "q_mu": torch.empty_like(q_mu), "q_sqrt": torch.empty_like(q_sqrt)}
    m.keep += [var, ls, Z, q_mu, q_sqrt]
    layer = _lib.MgpLayer(Mn, D, K, ls.numel(), Z.data_ptr(), q_mu.data_ptr(), q_sqrt.data_ptr(), var.data_ptr(),
                          ls.data_ptr())
    grad = _lib.MgpLayerGrad(g["Z"].data_ptr(), g["q_mu"].data_ptr(), g["q_sqrt"].data_ptr(), g["variance"].data_ptr(),
                             g["lengthscales"].data_ptr())
    return layer, grad, g


def marshal(model_kind: int, lik_kind: int, S: int, num_data: float, n_global: int, pred: Dict[str, object],
            assign: Dict[str, object], lik_var, assign_lik_var, X, Y, noise=None, seed: int = 0, point_offset: int = 0,
            temperature: float = TEMPERATURE) -> Marshalled:
    """Build the argument block of mgp_elbo_fwd_bwd from torch tensors (any device: only pointers and shapes are read)."""
    torch = _torch()
    m = Marshalled()
    m.pred, m.pred_grad, gp = layer_struct(pred, m)
    m.assign, m.assign_grad, ga = layer_struct(assign, m)
    K = m.pred.K
    if m.assign.K != K or m.assign.D != m.pred.D:
        raise ValueError("pred / assign layers disagree on K or D")
    m.X = X.contiguous()
    m.Y = Y.reshape(-1).contiguous()
    if m.X.dim() != 2 or m.X.shape[1] != m.pred.D or m.Y.numel() != m.X.shape[0]:
        raise ValueError(f"X {tuple(X.shape)} / Y {tuple(Y.shape)} do not match D = {m.pred.D}")
    m.N = int(m.X.shape[0])

    def per_component(v):
        if v is None:
            return None
        v = v.reshape(-1)
        if v.numel() == 1 and K > 1:
            v = v.expand(K)
        if v.numel() != K:
            raise ValueError(f"likelihood variance has {v.numel()} entries but K = {K}")
        return v.contiguous()

    m.lik_var, m.assign_lik_var = per_component(lik_var), per_component(assign_lik_var)
    dev = m.X.device
    m.elbo = torch.empty(1, dtype=torch.float64, device=dev)
    m.glik = torch.zeros(K, dtype=torch.float64, device=dev)
    m.galik = torch.zeros(K, dtype=torch.float64, device=dev)
    if noise is not None:
        z, u = (a.contiguous() for a in noise)
        for a in (z, u):
            if tuple(a.shape) != (S, m.N, K):
                raise ValueError(f"noise arrays must be [S, N, K] = {(S, m.N, K)}; got {tuple(a.shape)}")
        m.keep += [z, u]
        m.noise = _lib.MgpNoise(z.data_ptr(), u.data_ptr(), 0, int(point_offset))
    else:
        m.noise = _lib.MgpNoise(None, None, int(seed), int(point_offset))
    m.cfg = _lib.MgpElboCfg(int(model_kind), int(lik_kind), int(S), 0, float(temperature), float(num_data), int(n_global))
    m.grads = {f"pred.{k}": v for k, v in gp.items()}
    m.grads.update({f"assign.{k}": v for k, v in ga.items()})
    m.grads["lik_var"], m.grads["assign_lik_var"] = m.glik, m.galik
    return m


class LibmgpBackend:
    """Launches the marshalled call on the CUDA device.  No CPU fallback: constructing it without a GPU raises."""

    def __init__(self, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            _lib.load_library()
            raise RuntimeError("modulatedgps_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def launch(self, m: Marshalled):
        ctx = _lib.get_context(self.device)
        ctx.check(ctx.lib.mgp_elbo_fwd_bwd(ctx.handle, C.byref(m.cfg), C.byref(m.pred), C.byref(m.assign),
                                           _lib.ptr(m.lik_var), _lib.ptr(m.assign_lik_var), _lib.ptr(m.X), _lib.ptr(m.Y),
                                           m.N, C.byref(m.noise), _lib.ptr(m.elbo), C.byref(m.pred_grad),
                                           C.byref(m.assign_grad), _lib.ptr(m.glik), _lib.ptr(m.galik)))


# ---------------------------------------------------------------------------------------------------------------
# the model side
# ---------------------------------------------------------------------------------------------------------------
def parameter_slots(model) -> Dict[str, object]:
    """{libmgp gradient slot: GPflow Parameter} for an SMGP / SMGPModified built as the demos do
    (demos/demo_tf2.py:36-49, demo_tf2_2d_modified_multiclass.py:36-53)."""
    slots = {}
    for lname, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
        if not getattr(layer, "whiten", True) or getattr(layer, "q_diag", False):
            raise NotImplementedError("libmgp implements the reference's whitened, full-q_sqrt SVGP layers")
        slots[f"{lname}.variance"] = layer.kernel.variance
        slots[f"{lname}.lengthscales"] = layer.kernel.lengthscales
        slots[f"{lname}.Z"] = layer.inducing_variable.Z
        slots[f"{lname}.q_mu"] = layer.q_mu
        slots[f"{lname}.q_sqrt"] = layer.q_sqrt
    lik = model.likelihood.likelihood                     # BroadcastingLikelihood wraps it (models.py:32)
    if hasattr(lik, "variance"):
        slots["lik_var"] = lik.variance
    if hasattr(model, "assign_likelihood"):
        slots["assign_lik_var"] = model.assign_likelihood.likelihood.variance
    return slots


def model_kinds(model) -> Tuple[int, int]:
    model_kind = _lib.MODEL_SMGP_MODIFIED if hasattr(model, "assign_likelihood") else _lib.MODEL_SMGP
    name = type(model.likelihood.likelihood).__name__
    if name == "GaussianModified":
        return model_kind, _lib.LIK_GAUSSIAN
    if name == "MultiClass":
        if type(model.likelihood.likelihood.invlink).__name__ != "RobustMax":
            raise NotImplementedError("MultiClass is implemented with the RobustMax inverse link (as gpflow)")
        return model_kind, _lib.LIK_MULTICLASS
    raise NotImplementedError(f"{name}: libmgp implements GaussianModified and MultiClass(RobustMax) experts")


def build_training_loss(model, *, backend=None, noise: Optional[Callable] = None, seed: int = 0,
                        n_global: Optional[int] = None, point_offset: int = 0) -> Callable:
    """Returns `training_loss(data) -> -ELBO` (a TF scalar) for the reference model object `model`, differentiable
    w.r.t. `model.trainable_variables` through a `tf.custom_gradient` whose forward is one libmgp call.

    noise: None -> the on-device Philox stream (fresh per call, keyed by `seed`); or a callable
           `noise(S, N, K) -> (z, u)` returning the arrays tf.random.normal (models.py:57) and TFP's uniform draw
           (models.py:73) would have produced (bit-for-bit comparable runs)."""
    tf = _tf()
    backend = backend if backend is not None else LibmgpBackend()
    slots = parameter_slots(model)
    slot_names = list(slots)
    model_kind, lik_kind = model_kinds(model)
    trainable = list(model.trainable_parameters)
    variables = [p.unconstrained_variable for p in trainable]
    index_of = {id(p): i for i, p in enumerate(trainable)}
    S = int(model.num_samples)
    if model.num_data is None:
        raise ValueError("num_data must be set (the reference divides the KL term by it, models.py:79)")
    state = {"calls": 0}

    def eager_fwd_bwd(X, Y, *constrained):
        """EagerTensors in, [elbo] + gradients w.r.t. the constrained values (slot order) out."""
        dev = getattr(backend, "device", None)
        moved = {name: to_torch(v, dev, with_origin=True) for name, v in zip(slot_names, constrained)}
        vals = {name: t for name, (t, _) in moved.items()}
        Xt, Yt = to_torch(X, dev), to_torch(Y, dev)
        pred = {k: vals[f"pred.{k}"] for k in LAYER_KEYS}
        assign = {k: vals[f"assign.{k}"] for k in LAYER_KEYS}
        K = pred["q_mu"].shape[1]
        state["calls"] += 1
        nz = None
        if noise is not None:
            nz = tuple(to_torch(a, dev) for a in noise(S, Xt.shape[0], K))
        m = marshal(model_kind, lik_kind, S, float(model.num_data), int(n_global or Xt.shape[0]), pred, assign,
                    vals.get("lik_var"), vals.get("assign_lik_var"), Xt, Yt, noise=nz,
                    seed=(int(seed) << 20) + state["calls"], point_offset=point_offset)
        backend.launch(m)
        out = [to_tf(m.elbo.reshape(()).to(moved[slot_names[0]][1]))]
        for name, v in zip(slot_names, constrained):
            g = m.grads[name].to(moved[name][1])       # a gradient goes back to the device its parameter came from
            shape = tuple(int(s) for s in v.shape)
            if g.numel() != _numel(shape):
                g = g.sum().reshape(1)         # a scalar variance broadcast over the K components: sum of the parts
            out.append(to_tf(g.reshape(shape)))
        return out

    @tf.custom_gradient
    def neg_elbo(X, Y, *unconstrained):
        with tf.GradientTape(persistent=True, watch_accessed_variables=False) as tape:
            for u in unconstrained:
                tape.watch(u)
            forward = [p.transform.forward(u) if getattr(p, "transform", None) is not None else u
                       for p, u in zip(trainable, unconstrained)]
        constrained = []
        for name in slot_names:
            p = slots[name]
            i = index_of.get(id(p))
            constrained.append(forward[i] if i is not None else tf.convert_to_tensor(p))   # non-trainable: its value
        outs = tf.py_function(eager_fwd_bwd, [X, Y] + [tf.stop_gradient(c) for c in constrained],
                              Tout=[tf.float64] * (1 + len(slot_names)))
        elbo, grads_c = outs[0], outs[1:]

        def grad(upstream):
            per_param: List = [None] * len(trainable)
            for name, gc in zip(slot_names, grads_c):
                i = index_of.get(id(slots[name]))
                if i is None:
                    continue
                gc = tf.reshape(gc, tf.shape(forward[i]))
                per_param[i] = gc if per_param[i] is None else per_param[i] + gc     # shared Parameters accumulate
            result = []
            for i, u in enumerate(unconstrained):
                if per_param[i] is None:                   # trainable, but not on the ELBO's path
                    result.append(tf.zeros_like(u))
                else:                                      # d(-ELBO)/du = -upstream * J_bijector^T dELBO/dc
                    result.append(tape.gradient(forward[i], u, output_gradients=-upstream * per_param[i]))
            return (None, None) + tuple(result)

        return -elbo, grad

    def training_loss(data):
        X, Y = data
        return neg_elbo(tf.convert_to_tensor(X), tf.cast(tf.convert_to_tensor(Y), tf.float64), *variables)

    training_loss.variables = variables
    training_loss.slots = slots
    return training_loss


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def attach(model, **kw):
    """Route `model._training_loss`, `training_loss` and `_build_likelihood` (MixtureGPs/models.py:69-83) through libmgp,
    in place; everything else on the object (GPflow Parameters, predict_*, print_summary) is untouched."""
    loss = build_training_loss(model, **kw)
    model._mgp_training_loss = loss
    model._training_loss = lambda data: loss(data)
    model.training_loss = lambda data: loss(data)
    model._build_likelihood = lambda X, Y: -loss((X, Y))
    return model


def predict_f(layer, Xnew, backend=None):
    """SVGPModified.predict_f(Xnew, full_cov=False) (models.py:129-144) of a GPflow layer object on libmgp:
    Xnew [N, D] -> (fmean [N, K], fvar [N, K]) as TF tensors."""
    torch = _torch()
    backend = backend if backend is not None else LibmgpBackend()
    dev = backend.device
    m = Marshalled()
    t = {"variance": to_torch(layer.kernel.variance, dev), "lengthscales": to_torch(layer.kernel.lengthscales, dev),
         "Z": to_torch(layer.inducing_variable.Z, dev), "q_mu": to_torch(layer.q_mu, dev),
         "q_sqrt": to_torch(layer.q_sqrt, dev)}
    struct, _, _ = layer_struct(t, m)
    X = to_torch(Xnew, dev)
    N, K = X.shape[0], struct.K
    fmean = torch.empty(N, K, dtype=torch.float64, device=dev)
    fvar = torch.empty(N, K, dtype=torch.float64, device=dev)
    ctx = _lib.get_context(dev)
    ctx.check(ctx.lib.mgp_svgp_predict_f(ctx.handle, C.byref(struct), _lib.ptr(X), N, _lib.ptr(fmean), _lib.ptr(fvar)))
    return to_tf(fmean), to_tf(fvar)
