"""Model classes with the API surface of MixtureGPs/models.py, backed by libmgp (hand-written sm_100a CUDA).

    SVGPModified   MixtureGPs/models.py:147-160 (+ IndependentPosteriorSingleOutputModified, :126-144)
    SGP            models.py:23-41
    SMGP           models.py:44-103
    SMGPModified   models.py:106-123

Same constructor arguments, attribute names and method names as the reference, so the demos' model blocks
(e.g. demos/demo_tf2.py:36-49) carry over with `gpflow.kernels` / `MixtureGPs` imports swapped for this
package.  Host tensors are framework-neutral: numpy arrays, torch tensors or anything exporting `__dlpack__`
go in; device tensors that behave like arrays (`np.mean(x, 0)`, `np.hstack`, `np.argmax`) come out.

Differences a user can observe, all deliberate:
  * the S tiled copies of X (`integrate`, models.py:35-36) are never materialised; predict_f / predict_y
    return broadcast views of the one copy that is computed;
  * noise: the reference draws from TF's global Philox stream; here `noise=(z, u)` arrays may be passed for
    bit-for-bit comparable runs, otherwise an on-device Philox stream keyed by (`seed`, global point index,
    sample, component) is used, which makes results independent of how the points are sharded;
  * `loss.backward()` (torch autograd) delivers d loss / d unconstrained variable, as TF's GradientTape does
    for `model.trainable_variables` (utils/training_utils.py:8-10); the backward pass is hand-written CUDA,
    not autodiff.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, parallel
from .broadcasting_lik import BroadcastingLikelihood
from .likelihoods import GaussianModified
from .parameter import F64, FillTriangular, Module, Parameter, to_device_f64

float_type = F64
jitter_level = 1e-6
TEMPERATURE = 1e-2          # literal at MixtureGPs/models.py:60


class DeviceArray(torch.Tensor):
    """A CUDA tensor that converts to numpy on demand, so `np.mean(x, 0)`, `np.hstack`, `np.argmax(x, 1)` work on
    the values the model returns, as they do on TF eager tensors in the reference demos (demo_tf2.py:68-69,99,101)."""

    def __array__(self, dtype=None, copy=None):
        a = self.detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def numpy(self):
        return self.detach().cpu().as_subclass(torch.Tensor).numpy()


def _wrap(t: torch.Tensor) -> torch.Tensor:
    return t.as_subclass(DeviceArray)


class InducingPoints(Module):
    def __init__(self, Z):
        self.Z = Z if isinstance(Z, Parameter) else Parameter(np.asarray(Z, dtype=np.float64))

    @property
    def num_inducing(self) -> int:
        return int(self.Z.unconstrained_variable.shape[0])


class _LayerView:
    """Constrained parameter tensors of one layer plus the C struct that points at them (kept alive together)."""

    def __init__(self, layer: "SVGPModified"):
        self.variance = layer.kernel.variance.value().reshape(1)
        ls = layer.kernel.lengthscales.value()
        self.lengthscales = ls.reshape(-1)
        self.Z = layer.inducing_variable.Z.value()
        self.q_mu = layer.q_mu.value()
        self.q_sqrt = layer.q_sqrt.value()
        M, D = self.Z.shape
        K = self.q_mu.shape[1]
        if self.lengthscales.numel() not in (1, D):
            raise ValueError(f"lengthscales has {self.lengthscales.numel()} entries; expected 1 or D={D}")
        if tuple(self.q_sqrt.shape) != (K, M, M):
            raise ValueError(f"q_sqrt has shape {tuple(self.q_sqrt.shape)}; expected {(K, M, M)}")
        self.M, self.D, self.K = M, D, K
        self.struct = _lib.MgpLayer(M, D, K, self.lengthscales.numel(), self.Z.data_ptr(), self.q_mu.data_ptr(),
                                    self.q_sqrt.data_ptr(), self.variance.data_ptr(), self.lengthscales.data_ptr())

    def grad_buffers(self):
        g = {"Z": torch.empty_like(self.Z), "q_mu": torch.empty_like(self.q_mu), "q_sqrt": torch.empty_like(self.q_sqrt),
             "variance": torch.empty_like(self.variance), "lengthscales": torch.empty_like(self.lengthscales)}
        s = _lib.MgpLayerGrad(g["Z"].data_ptr(), g["q_mu"].data_ptr(), g["q_sqrt"].data_ptr(),
                              g["variance"].data_ptr(), g["lengthscales"].data_ptr())
        return g, s


def _points(X, D_expected=None) -> Tuple[torch.Tensor, Optional[int]]:
    """[N, D] or tiled [S, N, D] (the output of `integrate`) -> ([N, D] contiguous device tensor, S or None)."""
    Xt = to_device_f64(X)
    S = None
    if Xt.dim() == 3:            # S identical copies (models.py:36): use the first
        S = Xt.shape[0]
        Xt = Xt[0].contiguous()
    if Xt.dim() != 2:
        raise ValueError(f"X must be [N, D] or [S, N, D]; got shape {tuple(Xt.shape)}")
    if D_expected is not None and Xt.shape[1] != D_expected:
        raise ValueError(f"X has {Xt.shape[1]} input dimensions; the layer's inducing points have {D_expected}")
    return Xt, S


class SVGPModified(Module):
    """Whitened SVGP layer with K latent GPs sharing one kernel and one set of inducing inputs.

    Signature of gpflow.models.SVGP as the reference uses it (demos/demo_tf2.py:43-45):
    SVGPModified(kernel=, likelihood=, inducing_variable=, num_latent_gps=, whiten=True).
    GPflow defaults: q_mu = 0 [M, K], q_sqrt = I [K, M, M] with the triangular() bijector."""

    def __init__(self, kernel, likelihood, inducing_variable, *, mean_function=None, num_latent_gps: int = 1,
                 q_diag: bool = False, q_mu=None, q_sqrt=None, whiten: bool = True, num_data=None):
        if not whiten:
            raise NotImplementedError("the reference builds every layer with whiten=True; only that path exists here")
        if q_diag:
            raise NotImplementedError("q_diag=True is not on the reference's path")
        if mean_function is not None:
            raise NotImplementedError("the reference uses the default Zero mean function")
        self.kernel = kernel
        self.likelihood = likelihood
        self.num_latent_gps = int(num_latent_gps)
        self.num_data = num_data
        self.whiten = True
        self.q_diag = False
        self.mean_function = None
        self.inducing_variable = inducing_variable if isinstance(inducing_variable, InducingPoints) \
            else InducingPoints(inducing_variable)
        M, K = self.inducing_variable.num_inducing, self.num_latent_gps
        if K > _lib.MAX_K:
            raise ValueError(f"num_latent_gps={K} exceeds libmgp's MGP_MAX_K={_lib.MAX_K}")
        self.q_mu = Parameter(np.zeros((M, K)) if q_mu is None else q_mu)
        self.q_sqrt = Parameter(np.stack([np.eye(M)] * K) if q_sqrt is None else q_sqrt, transform=FillTriangular())

    # -- reference API ------------------------------------------------------------------------------
    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        """SVGP.predict_f -> posterior().fused_predict_f -> _conditional_fused (models.py:129-144)."""
        if full_cov or full_output_cov:
            raise NotImplementedError("the reference only calls predict_f(full_cov=False) (models.py:40,56,64,87,96,113,118)")
        view = _LayerView(self)
        X, S = _points(Xnew, view.D)
        ctx = _lib.get_context(X.device)
        N = X.shape[0]
        fmean = torch.empty(N, view.K, dtype=F64, device=X.device)
        fvar = torch.empty(N, view.K, dtype=F64, device=X.device)
        ctx.check(ctx.lib.mgp_svgp_predict_f(ctx.handle, C.byref(view.struct), _lib.ptr(X), N, _lib.ptr(fmean), _lib.ptr(fvar)))
        if S is not None:
            fmean, fvar = fmean.unsqueeze(0).expand(S, N, view.K), fvar.unsqueeze(0).expand(S, N, view.K)
        return _wrap(fmean), _wrap(fvar)

    def prior_kl(self):
        view = _LayerView(self)
        ctx = _lib.get_context(view.Z.device)
        out = torch.empty(1, dtype=F64, device=view.Z.device)
        ctx.check(ctx.lib.mgp_prior_kl(ctx.handle, C.byref(view.struct), _lib.ptr(out)))
        return _wrap(out[0])

    def posterior(self, precompute_cache=None):
        """Kept for API parity (models.py:148-160): the posterior object IS the layer here."""
        return self

    def fused_predict_f(self, Xnew, full_cov=False, full_output_cov=False):
        return self.predict_f(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)


class _ElboFunction(torch.autograd.Function):
    """ELBO as a torch autograd node: forward runs libmgp's forward AND backward kernels once and caches
    d ELBO / d unconstrained variable; backward scales them by the incoming gradient."""

    @staticmethod
    def forward(ctx, model, X, Y, noise, n_global, point_offset, *variables):
        elbo, grads = model._elbo_and_unconstrained_grads(X, Y, noise, n_global, point_offset, variables)
        ctx.grads = grads
        return elbo

    @staticmethod
    def backward(ctx, gout):
        live = [g for g in ctx.grads if g is not None]
        scaled = iter(torch._foreach_mul(live, gout))        # one multi-tensor kernel instead of one launch per variable
        return (None, None, None, None, None, None) + tuple(None if g is None else next(scaled) for g in ctx.grads)


class _RelaxedAssignments:
    """What SMGP.W_dist returns: `.sample(n)` draws relaxed one-hot assignment weights [n, S*N, K] (models.py:60,73)."""

    def __init__(self, model, view, X, S, noise):
        self.model, self.view, self.X, self.S, self.noise = model, view, X, int(S), noise
        self.temperature = model.temperature

    def sample(self, n=1):
        X, S, K = self.X, self.S, self.view.K
        N = X.shape[0]
        ctx = _lib.get_context(X.device)
        out = torch.empty(int(n), S, N, K, dtype=F64, device=X.device)
        for r in range(int(n)):
            keep = []
            if self.noise is not None and r == 0:
                z, u = (to_device_f64(a, X.device) for a in self.noise)
                for a in (z, u):
                    if tuple(a.shape) != (S, N, K):
                        raise ValueError(f"noise arrays must have shape {(S, N, K)}; got {tuple(a.shape)}")
                keep = [z, u]
                nz = _lib.MgpNoise(z.data_ptr(), u.data_ptr(), 0, 0)
            else:
                self.model._step += 1
                nz = _lib.MgpNoise(None, None, (self.model.seed << 20) + self.model._step, 0)
            ctx.check(ctx.lib.mgp_w_sample(ctx.handle, C.byref(self.view.struct), _lib.ptr(X), N, S, float(self.temperature),
                                           C.byref(nz), _lib.ptr(out[r])))
            del keep
        return _wrap(out.reshape(int(n), S * N, K))


class SGP(Module):
    """Scalable GP: X -> Xt = integrate(X) -> GP -> Y   (MixtureGPs/models.py:23-41)."""

    def __init__(self, likelihood, pred_layer, num_samples=1, num_data=None):
        self.num_samples = num_samples
        self.num_data = num_data
        self.likelihood = BroadcastingLikelihood(likelihood)
        self.pred_layer = pred_layer

    def integrate(self, X, S=1):
        """models.py:35-36: S tiled copies of X — returned as a broadcast VIEW (no S-fold memory)."""
        Xt = to_device_f64(X)
        return _wrap(Xt.unsqueeze(0).expand(S, *Xt.shape)), None

    def predict_y(self, Xnew, S=1):
        """models.py:38-41 -> ([S, N, K], [S, N, K]) (broadcast views of one computed copy)."""
        view = _LayerView(self.pred_layer)
        X, _ = _points(Xnew, view.D)
        ctx = _lib.get_context(X.device)
        N = X.shape[0]
        mean = torch.empty(N, view.K, dtype=F64, device=X.device)
        var = torch.empty(N, view.K, dtype=F64, device=X.device)
        lik_var = self.likelihood.component_variances(view.K)
        ctx.check(ctx.lib.mgp_predict_y(ctx.handle, C.byref(view.struct), self.likelihood.kind, _lib.ptr(lik_var),
                                        _lib.ptr(X), N, _lib.ptr(mean), _lib.ptr(var)))
        return _wrap(mean.unsqueeze(0).expand(S, N, view.K)), _wrap(var.unsqueeze(0).expand(S, N, view.K))


class SMGP(SGP):
    """Mixture of sparse GPs for data association (MixtureGPs/models.py:44-103)."""

    _model_kind = _lib.MODEL_SMGP

    def __init__(self, likelihood, pred_layer, assign_layer, K=3, num_samples=1, num_data=None):
        SGP.__init__(self, likelihood, pred_layer, num_samples, num_data)
        self.assign_layer = assign_layer
        self.K = K
        self.temperature = TEMPERATURE
        self.seed = 0                 # Philox key of the on-device noise stream
        self._step = 0                # advanced every stochastic evaluation (fresh noise per call, as TF)
        self.process_group = None     # set by enable_data_parallel()
        self._dp = False
        for layer in (pred_layer, assign_layer):
            if layer.num_latent_gps != K:
                raise ValueError(f"layer has num_latent_gps={layer.num_latent_gps} but the model has K={K}")

    # -- data parallelism over the points -------------------------------------------
    def enable_data_parallel(self, process_group=None, collective="torch"):
        """Each rank passes ITS shard of (X, Y) to the loss; per-shard sums are combined by an all-reduce of the flat
        reduce buffer (NCCL over NVLink when the group's backend is nccl).  collective = "torch": torch.distributed
        reduces between mgp_elbo_local and mgp_elbo_finish; "nccl": a communicator of this model's own is attached to
        the libmgp context (mgp_ctx_set_comm) and the C library issues ncclAllReduce itself — no torch.distributed call
        on the step's path."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if collective not in ("torch", "nccl"):
            raise ValueError("collective must be 'torch' or 'nccl'")
        self.process_group = process_group
        self._dp = True
        self._own_comm = None
        if collective == "nccl":
            dev = torch.device("cuda", torch.cuda.current_device())
            self._own_comm = parallel.create_nccl_comm(process_group, dev)
            _lib.get_context(dev).set_comm(self._own_comm)
        return self

    # -- reference API ------------------------------------------------------------------------------
    def _build_likelihood(self, X, Y, noise=None, n_global=None, point_offset=None):
        """ELBO (models.py:69-79).  noise = (z, u), each [S, N, K]: what tf.random.normal (models.py:57) and
        TFP's uniform draw (models.py:73) would have returned; None -> on-device Philox."""
        variables = self.trainable_variables
        return _wrap(_ElboFunction.apply(self, X, Y, noise, n_global, point_offset, *variables))

    def _training_loss(self, data, **kw):
        X, Y = data
        return -self._build_likelihood(X, Y, **kw)

    def training_loss(self, data, **kw):
        return self._training_loss(data, **kw)

    def training_loss_closure(self, data, *, compile=True):
        """gpflow ExternalDataTrainingLossMixin.training_loss_closure (utils/training_utils.py:5): `data` is a
        tuple or an iterator of minibatches.  `compile` is accepted and ignored (nothing is traced)."""
        if hasattr(data, "__next__"):
            def closure():
                return self.training_loss(next(data))
        else:
            def closure():
                return self.training_loss(data)
        return closure

    def W_dist(self, X, noise=None):
        """models.py:55-61: the relaxed one-hot (Gumbel-softmax, T = 1e-2) distribution over assignments of the points
        in X ([S, N, D] as `integrate` returns it, or [N, D] with S = num_samples).  Like the TFP object the reference
        gets back, it is consumed through `.sample(n)` -> [n, S*N, K].  noise = (z, u), each [S, N, K], fixes the
        draw of `.sample(1)`; otherwise the device Philox stream is used."""
        view = _LayerView(self.assign_layer)
        Xd, S = _points(X, view.D)
        return _RelaxedAssignments(self, view, Xd, int(self.num_samples) if S is None else S, noise)

    def E_log_p_Y(self, X, Y, W_SND):
        """models.py:63-67: logsumexp_S(sum_k W ve) - log S per point -> [N].  W_SND [S, N, K]."""
        return _wrap(self._e_log_p_y(self.pred_layer, self.likelihood, X, Y, W_SND))

    def _e_log_p_y(self, layer, likelihood, X, Y, W_SND):
        view = _LayerView(layer)
        Xd, _ = _points(X, view.D)
        W = to_device_f64(W_SND, Xd.device)
        N = Xd.shape[0]
        if W.dim() != 3 or W.shape[1] != N or W.shape[2] != view.K:
            raise ValueError(f"W_SND must be [S, {N}, {view.K}]; got {tuple(W.shape)}")
        Yd = to_device_f64(Y, Xd.device).reshape(-1)
        if Yd.numel() != N:
            raise ValueError(f"Y has {Yd.numel()} rows but X has {N}")
        ctx = _lib.get_context(Xd.device)
        out = torch.empty(N, dtype=F64, device=Xd.device)
        lik_var = likelihood.component_variances(view.K)
        ctx.check(ctx.lib.mgp_e_log_p_y(ctx.handle, C.byref(view.struct), likelihood.kind, _lib.ptr(lik_var), _lib.ptr(Xd),
                                        _lib.ptr(Yd), N, int(W.shape[0]), _lib.ptr(W), _lib.ptr(out)))
        return out

    def predict_assign(self, Xnew, S=1):
        """models.py:85-89 -> softmax assignment probabilities [N, K]."""
        return self.predict_assign_with_argmax(Xnew)[0]

    def predict_assign_with_argmax(self, Xnew):
        """Probabilities [N, K] and the int64 argmax [N] the demos take with np.argmax (demo_tf2.py:101)."""
        view = _LayerView(self.assign_layer)
        X, _ = _points(Xnew, view.D)
        ctx = _lib.get_context(X.device)
        N = X.shape[0]
        probs = torch.empty(N, view.K, dtype=F64, device=X.device)
        arg = torch.empty(N, dtype=torch.int64, device=X.device)
        ctx.check(ctx.lib.mgp_predict_assign(ctx.handle, C.byref(view.struct), _lib.ptr(X), N, _lib.ptr(probs), _lib.ptr(arg)))
        return _wrap(probs), _wrap(arg)

    def predict_samples(self, Xnew, S=1, noise=None):
        """models.py:91-103 -> (samples_y [S, N, 1], samples_f [S, N, 1]).  noise = (z_assign, u, z_pred)."""
        pv, av = _LayerView(self.pred_layer), _LayerView(self.assign_layer)
        X, _ = _points(Xnew, pv.D)
        ctx = _lib.get_context(X.device)
        N = X.shape[0]
        lik_var = self.likelihood.component_variances(pv.K)
        keep = []
        if noise is not None:
            z, u, zp = (to_device_f64(a, X.device) for a in noise)
            for a in (z, u, zp):
                if tuple(a.shape) != (S, N, pv.K):
                    raise ValueError(f"noise arrays must have shape {(S, N, pv.K)}; got {tuple(a.shape)}")
            keep = [z, u, zp]
            nz = _lib.MgpNoise(z.data_ptr(), u.data_ptr(), 0, 0)
            zp_ptr = _lib.ptr(zp)
        else:
            self._step += 1
            nz = _lib.MgpNoise(None, None, (self.seed << 20) + self._step, 0)
            zp_ptr = _lib.ptr(None)
        sy = torch.empty(S, N, dtype=F64, device=X.device)
        sf = torch.empty(S, N, dtype=F64, device=X.device)
        ctx.check(ctx.lib.mgp_predict_samples(ctx.handle, C.byref(pv.struct), C.byref(av.struct), self.likelihood.kind,
                                              _lib.ptr(lik_var), _lib.ptr(X), N, S, self.temperature, C.byref(nz), zp_ptr,
                                              _lib.ptr(sy), _lib.ptr(sf)))
        del keep
        return _wrap(sy.unsqueeze(-1)), _wrap(sf.unsqueeze(-1))

    # -- native path ----------------------------------------------------------------------------------
    def _assign_lik_variances(self, K):
        return None

    def elbo_and_grads(self, X, Y, noise=None, n_global=None, point_offset=None):
        """(ELBO, {parameter path: d ELBO / d CONSTRAINED value}) — what the C-ABI returns, before the
        bijector chain rule.  Used by the parity tests and by callers that own their optimiser."""
        return self._run(X, Y, noise, n_global, point_offset)[:2]

    def _run(self, X, Y, noise, n_global, point_offset):
        pv, av = _LayerView(self.pred_layer), _LayerView(self.assign_layer)
        Xd, _ = _points(X, pv.D)
        dev = Xd.device
        Yd = to_device_f64(Y, dev).reshape(-1)
        N = Xd.shape[0]
        if Yd.numel() != N:
            raise ValueError(f"Y has {Yd.numel()} rows but X has {N}")
        K, S = pv.K, int(self.num_samples)
        if self.num_data is None:
            raise ValueError("num_data must be set (the reference divides the KL term by it, models.py:79)")
        ctx = _lib.get_context(dev)
        lik_var = self.likelihood.component_variances(K)
        alik_var = self._assign_lik_variances(K)
        # global batch size / Philox offset of this shard
        if self._dp:
            if n_global is None or point_offset is None:
                n_global, point_offset = parallel.global_count_and_offset(N, self.process_group, dev)
        else:
            n_global = N if n_global is None else int(n_global)
            point_offset = 0 if point_offset is None else int(point_offset)
        keep = []
        if noise is not None:
            z, u = (to_device_f64(a, dev) for a in noise)
            for a in (z, u):
                if tuple(a.shape) != (S, N, K):
                    raise ValueError(f"noise arrays must have shape {(S, N, K)}; got {tuple(a.shape)}")
            keep = [z, u]
            nz = _lib.MgpNoise(z.data_ptr(), u.data_ptr(), 0, point_offset)
        else:
            self._step += 1
            nz = _lib.MgpNoise(None, None, (self.seed << 20) + self._step, point_offset)
        cfg = _lib.MgpElboCfg(self._model_kind, self.likelihood.kind, S, 0, float(self.temperature),
                              float(self.num_data), int(n_global))
        pg, pgs = pv.grad_buffers()
        ag, ags = av.grad_buffers()
        elbo = torch.empty(1, dtype=F64, device=dev)
        glik = torch.zeros(K, dtype=F64, device=dev)
        galik = torch.zeros(K, dtype=F64, device=dev)
        lib, h = ctx.lib, ctx.handle
        if self._dp and getattr(ctx, "comm", None) is None:
            rb = torch.empty(int(lib.mgp_reduce_buffer_len(C.byref(pv.struct), C.byref(av.struct))), dtype=F64, device=dev)
            ctx.check(lib.mgp_elbo_local(h, C.byref(cfg), C.byref(pv.struct), C.byref(av.struct), _lib.ptr(lik_var),
                                         _lib.ptr(alik_var), _lib.ptr(Xd), _lib.ptr(Yd), N, C.byref(nz), _lib.ptr(rb)))
            parallel.all_reduce_sum_(rb, self.process_group)        # the single fused collective of the step
            ctx.check(lib.mgp_elbo_finish(h, C.byref(cfg), C.byref(pv.struct), C.byref(av.struct), _lib.ptr(lik_var),
                                          _lib.ptr(alik_var), _lib.ptr(rb), _lib.ptr(elbo), C.byref(pgs), C.byref(ags),
                                          _lib.ptr(glik), _lib.ptr(galik)))
        else:
            ctx.check(lib.mgp_elbo_fwd_bwd(h, C.byref(cfg), C.byref(pv.struct), C.byref(av.struct), _lib.ptr(lik_var),
                                           _lib.ptr(alik_var), _lib.ptr(Xd), _lib.ptr(Yd), N, C.byref(nz), _lib.ptr(elbo),
                                           C.byref(pgs), C.byref(ags), _lib.ptr(glik), _lib.ptr(galik)))
        del keep
        grads = {}
        for lname, g in (("pred", pg), ("assign", ag)):
            for k, v in g.items():
                grads[f"{lname}.{k}"] = v
        grads["lik_var"] = glik
        grads["assign_lik_var"] = galik
        views = {"pred": pv, "assign": av}
        return elbo[0], grads, views

    def _param_grad_map(self, grads):
        """(Parameter -> gradient w.r.t. its constrained value), accumulating shared parameters."""
        out = {}

        def add(p, g):
            if p is None or not isinstance(p, Parameter):
                return
            shape = p.shape                      # (of the constrained value; evaluating p.value() here would re-run
            if g.numel() == shape.numel():       # the bijector kernels twice per parameter)
                g = g.reshape(shape)
            elif shape.numel() == 1:
                # a scalar parameter broadcast over the K components (GaussianModified(variance=v) with D=None is the
                # reference's constructor default, likelihoods.py:13-19): the gradient of a broadcast is the sum
                g = g.sum().reshape(shape)
            else:
                raise ValueError(f"gradient of {g.numel()} entries for a parameter of shape {tuple(shape)}")
            out[id(p)] = (p, g if id(p) not in out else out[id(p)][1] + g)

        for lname, layer in (("pred", self.pred_layer), ("assign", self.assign_layer)):
            add(layer.kernel.variance, grads[f"{lname}.variance"])
            add(layer.kernel.lengthscales, grads[f"{lname}.lengthscales"])
            add(layer.inducing_variable.Z, grads[f"{lname}.Z"])
            add(layer.q_mu, grads[f"{lname}.q_mu"])
            add(layer.q_sqrt, grads[f"{lname}.q_sqrt"])
        if isinstance(self.likelihood.likelihood, GaussianModified):
            add(self.likelihood.likelihood.variance, grads["lik_var"])
        self._add_extra_param_grads(add, grads)
        return out

    def _add_extra_param_grads(self, add, grads):
        pass

    def _elbo_and_unconstrained_grads(self, X, Y, noise, n_global, point_offset, variables: Sequence[torch.Tensor]):
        elbo, grads, _ = self._run(X, Y, noise, n_global, point_offset)
        pmap = self._param_grad_map(grads)
        by_var = {id(p.unconstrained_variable): (p, g) for p, g in pmap.values()}
        out = []
        for v in variables:
            hit = by_var.get(id(v))
            if hit is None:
                out.append(torch.zeros_like(v))      # trainable but not on the ELBO's path (e.g. an unused likelihood)
            else:
                p, g = hit
                out.append(p.transform.grad_to_unconstrained(g, v.detach()).reshape(v.shape))
        return elbo, out


class SMGPModified(SMGP):
    """MixtureGPs/models.py:106-123: a second weighted term scores the ASSIGN layer's outputs against Y under
    `assign_likelihood`; the data term is logsumexp_S(E_log_p_A) + logsumexp_S(E_log_p_y)."""

    _model_kind = _lib.MODEL_SMGP_MODIFIED

    def __init__(self, likelihood, assign_likelihood, pred_layer, assign_layer, K=3, num_samples=1, num_data=None):
        SMGP.__init__(self, likelihood, pred_layer, assign_layer, K, num_samples, num_data)
        self.assign_likelihood = BroadcastingLikelihood(assign_likelihood)
        if self.assign_likelihood.kind != _lib.LIK_GAUSSIAN:
            raise NotImplementedError("assign_likelihood must be GaussianModified (as in every reference demo)")

    def _assign_lik_variances(self, K):
        return self.assign_likelihood.component_variances(K)

    def E_log_p_Y(self, X, Y, W_SND):
        """models.py:112-123: logsumexp_S of the assign layer's outputs scored against Y under `assign_likelihood`
        plus logsumexp_S of the experts' term, each sum_k W ve - log S."""
        term_a = self._e_log_p_y(self.assign_layer, self.assign_likelihood, X, Y, W_SND)
        term_y = self._e_log_p_y(self.pred_layer, self.likelihood, X, Y, W_SND)
        return _wrap(term_a + term_y)

    def _add_extra_param_grads(self, add, grads):
        add(self.assign_likelihood.likelihood.variance, grads["assign_lik_var"])
