"""The callers either side of the ELBO step, on the device.

    run_adam              utils/training_utils.py:4-28 with the same signature, printing and return value
    FusedAdam             tf.optimizers.Adam(lr).minimize(training_loss, model.trainable_variables) (:6-10) as ONE
                          libmgp kernel over every unconstrained variable, bijector chain rule included
    DeviceMinibatches     tf.data ... .shuffle(N, seed).batch(B).repeat() (demos/demo_tf2.py:53-56) without the host
    HostBatchStream       host-resident batches staged to the device one step ahead on a copy stream
    kmeans                scipy.cluster.vq.kmeans as the demos call it for the inducing inputs (demo_tf2.py:39)
    predict_samples_batched   the demos' chunked predict_samples loop (demo_tf2.py:62-68)

TF's Adam defaults are kept (beta1 0.9, beta2 0.999, epsilon 1e-7, bias correction folded into the step size).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .parameter import F64, FillTriangular, Softplus, to_device_f64


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(rc, what):
    if rc != 0:
        raise _lib.MgpError(rc, f"{what} failed")


class FusedAdam:
    """Adam on the model's unconstrained variables, fed with the gradients w.r.t. CONSTRAINED values that
    mgp_elbo_fwd_bwd returns; softplus / fill-triangular chain rule and the update run in one kernel launch."""

    def __init__(self, model, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.model, self.lr, self.beta_1, self.beta_2, self.epsilon = model, float(lr), beta_1, beta_2, epsilon
        self.iterations = 0
        self.params = list(model.trainable_parameters)
        if len(self.params) > _lib.ADAM_MAX_SLOTS:
            raise ValueError(f"{len(self.params)} trainable parameters; libmgp updates at most {_lib.ADAM_MAX_SLOTS} per launch")
        self.state = {}
        for p in self.params:
            th = p.unconstrained_variable
            gather, kind = None, _lib.TRANSFORM_IDENTITY
            if isinstance(p.transform, Softplus):
                kind = _lib.TRANSFORM_SOFTPLUS
            elif isinstance(p.transform, FillTriangular):
                n = th.shape[-1]
                m = int(round((np.sqrt(8 * n + 1) - 1) / 2))
                flat_pos, src = p.transform._maps(m, th.device)
                pos_of = torch.empty(n, dtype=torch.int64, device=th.device)
                pos_of[src] = flat_pos                                    # vector entry i sits at tril position pos_of[i]
                lead = int(th.numel() // n)
                gather = (torch.arange(lead, device=th.device).unsqueeze(1) * (m * m) + pos_of.unsqueeze(0)).reshape(-1).contiguous()
            self.state[id(p)] = {"m": torch.zeros_like(th), "v": torch.zeros_like(th), "gather": gather, "kind": kind}

    def apply_gradients(self, param_grads, grad_scale=-1.0, guard=None):
        """param_grads: {id(Parameter): (Parameter, d ELBO / d constrained value)}; grad_scale -1 minimises -ELBO.
        guard: device scalar (the step's ELBO); a non-finite value makes the kernel skip the whole update."""
        self.iterations += 1
        slots = (_lib.MgpAdamSlot * _lib.ADAM_MAX_SLOTS)()
        keep, ns = [], 0
        for p in self.params:
            hit = param_grads.get(id(p))
            if hit is None:
                continue                                                   # no gradient: skipped, as TF's minimize does
            g = hit[1].contiguous()
            st = self.state[id(p)]
            th = p.unconstrained_variable.data
            if st["gather"] is None and g.numel() != th.numel():
                raise ValueError(f"gradient of {g.numel()} entries for a parameter of {th.numel()}")
            keep.append(g)
            slots[ns] = _lib.MgpAdamSlot(th.data_ptr(), g.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(),
                                         None if st["gather"] is None else st["gather"].data_ptr(), th.numel(), st["kind"], 0)
            ns += 1
        lib = _lib.load_library()
        _check(lib.mgp_adam_step(_stream_ptr(), slots, ns, float(grad_scale), self.lr, self.beta_1, self.beta_2,
                                 self.epsilon, self.iterations, _lib.ptr(guard)), "mgp_adam_step")

    def minimize(self, data, **kw):
        """One optimisation step on the minibatch `data` = (X, Y); returns the training loss (-ELBO) tensor."""
        X, Y = data
        elbo, grads, _ = self.model._run(X, Y, kw.get("noise"), kw.get("n_global"), kw.get("point_offset"))
        self.apply_gradients(self.model._param_grad_map(grads), guard=elbo)
        return -elbo


def make_adam(model, lr):
    return FusedAdam(model, lr)


def run_adam(model, num_iter, train_iter, lr, compile=True):
    """utils/training_utils.py:4-28.  `train_iter`: an iterator of (X, Y) minibatches or one (X, Y) tuple.

    The loop below deliberately keeps the reference's control flow and output format line for line (header, an ELBO
    line every 5 iterations from a SECOND forward on a new minibatch, KeyboardInterrupt -> "stopping training",
    (iters, elbos) returned) so that logs and plots made from it are interchangeable with the reference's; what differs
    is the optimiser behind `optimization_step` (one fused libmgp launch instead of TF's minimize) and that a failed
    Cholesky is raised at the logging points, where the loop synchronises anyway (TF raises inside the step)."""
    training_loss = model.training_loss_closure(train_iter, compile=compile)
    optimizer = make_adam(model, lr)
    batches = train_iter if hasattr(train_iter, "__next__") else None

    def optimization_step():
        optimizer.minimize(next(batches) if batches is not None else train_iter)

    print('{:>5s}'.format("iter") + '{:>24s}'.format("ELBO:"))
    iters = []
    elbos = []
    for i in range(1, num_iter + 1):
        try:
            optimization_step()

            if i % 5 == 0 or i == 0:
                elbo = -float(training_loss().detach())     # a second forward on a NEW minibatch, as the reference
                _lib.get_context().check_status()           # not-PD Kuu since the last log: raise (updates were skipped)
                print('{:>5d}'.format(i) + '{:>24.6f}'.format(elbo))
                iters.append(i)
                elbos.append(elbo)
        except KeyboardInterrupt:
            print("stopping training")
            break

    return iters, elbos


class DeviceMinibatches:
    """Infinite iterator of shuffled minibatches held on the device: a fresh permutation per epoch (tf.data's
    reshuffle_each_iteration), batches of `batch_size` rows with a short last batch, repeated for ever."""

    def __init__(self, X, Y, batch_size, seed=0, shuffle=True):
        self.X = to_device_f64(X)
        self.Y = to_device_f64(Y, self.X.device).reshape(-1)
        if self.X.dim() != 2 or self.Y.numel() != self.X.shape[0]:
            raise ValueError("X must be [N, D] and Y must have N rows")
        self.N, self.D = self.X.shape
        self.batch_size, self.shuffle = int(batch_size), bool(shuffle)
        self.gen = torch.Generator(device=self.X.device)
        self.gen.manual_seed(int(seed))
        self._perm, self._pos = None, 0

    def __iter__(self):
        return self

    def __next__(self):
        if self._perm is None or self._pos >= self.N:
            self._perm = (torch.randperm(self.N, generator=self.gen, device=self.X.device) if self.shuffle
                          else torch.arange(self.N, device=self.X.device))
            self._pos = 0
        idx = self._perm[self._pos:self._pos + self.batch_size].contiguous()
        self._pos += self.batch_size
        B = idx.numel()
        Xb = torch.empty(B, self.D, dtype=F64, device=self.X.device)
        Yb = torch.empty(B, dtype=F64, device=self.X.device)
        lib = _lib.load_library()
        _check(lib.mgp_gather_rows(_stream_ptr(), _lib.ptr(self.X), _lib.ptr(self.Y), idx.data_ptr(), B, self.D,
                                   _lib.ptr(Xb), _lib.ptr(Yb)), "mgp_gather_rows")
        return Xb, Yb.unsqueeze(1)


class HostBatchStream:
    """(X, Y) batches that live in HOST memory, handed to the training step as device tensors with the copy of batch
    i + 1 in flight on a copy stream while step i computes (what `tf.data`'s `.prefetch()` / device staging does for
    the reference's input pipeline, demos/demo_tf2.py:53-56, when the data do not fit on or do not start on the GPU).

        for Xd, Yd in HostBatchStream(batches, device): loss = model._training_loss((Xd, Yd)); ...

    `batches`: any iterator of (X, Y) host arrays / CPU tensors; pinned float64 tensors are copied as they are, anything
    else is staged through a pinned buffer first.  Two device buffer pairs: the copy into a pair starts only after the
    step that last read it (two batches back) has been enqueued in full on the compute stream."""

    def __init__(self, batches, device=None):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.batches = iter(batches)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._dev = [None, None]          # device buffer pairs
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._i = 0
        self._pending = None              # index of the pair whose copy is in flight
        self._done = False
        self._keep = [None, None]
        self._enqueue()

    @staticmethod
    def _pinned(a):
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
        t = t.to(F64).contiguous()
        return t if t.is_pinned() else t.pin_memory()

    def _enqueue(self):
        try:
            X, Y = next(self.batches)
        except StopIteration:
            self._done, self._pending = True, None
            return
        Xh, Yh = self._pinned(X), self._pinned(Y)
        k = self._i & 1
        self._i += 1
        buf = self._dev[k]
        if buf is None or buf[0].shape != Xh.shape or buf[1].shape != Yh.shape:
            buf = self._dev[k] = (torch.empty(Xh.shape, dtype=F64, device=self.device),
                                  torch.empty(Yh.shape, dtype=F64, device=self.device))
        # everything enqueued on the compute stream so far (the step that last read this pair included) comes first
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.copy_stream):
            buf[0].copy_(Xh, non_blocking=True)
            buf[1].copy_(Yh, non_blocking=True)
            self._ready[k].record(self.copy_stream)
        self._keep[k] = (Xh, Yh)          # the pinned source outlives the asynchronous copy
        self._pending = k

    def __iter__(self):
        return self

    def __next__(self):
        if self._pending is None:
            raise StopIteration
        k = self._pending
        torch.cuda.current_stream(self.device).wait_event(self._ready[k])
        out = self._dev[k]
        self._enqueue()                   # the next batch's copy overlaps the step the caller is about to launch
        return out


def kmeans(obs, k_or_guess, iter=20, thresh=1e-5, seed=None, max_lloyd=200):
    """scipy.cluster.vq.kmeans(obs, k, iter, thresh, seed=) on the device: `iter` runs of Lloyd's algorithm from k
    random observations, each until the distortion (mean Euclidean distance to the nearest code) changes by at most
    `thresh`; the codebook of the best run with empty clusters dropped, and its distortion, come back as numpy.
    Not bit-comparable with scipy (different random streams); tested for equal-or-better distortion."""
    X = to_device_f64(obs)
    if X.dim() == 1:
        X = X.unsqueeze(1)
    N, D = X.shape
    lib = _lib.load_library()
    rng = np.random.default_rng(seed)
    guess = None if np.isscalar(k_or_guess) else to_device_f64(k_or_guess, X.device).reshape(-1, D)
    M = int(k_or_guess) if guess is None else guess.shape[0]
    label = torch.empty(N, dtype=torch.int32, device=X.device)
    count = torch.empty(M, dtype=torch.int32, device=X.device)
    scratch = torch.empty(1024, dtype=F64, device=X.device)
    dist = torch.empty(1, dtype=F64, device=X.device)
    best = (None, np.inf)
    for _ in range(1 if guess is not None else int(iter)):
        cent = guess.clone() if guess is not None else X[torch.as_tensor(rng.choice(N, M, replace=False), device=X.device)].clone()
        prev, done_it = np.inf, 0
        while done_it < max_lloyd:
            _check(lib.mgp_kmeans_iterate(_stream_ptr(), _lib.ptr(X), N, D, _lib.ptr(cent), M, 4, label.data_ptr(),
                                          count.data_ptr(), _lib.ptr(scratch), _lib.ptr(dist)), "mgp_kmeans_iterate")
            done_it += 4
            cur = float(dist)
            if abs(prev - cur) <= thresh:
                break
            prev = cur
        if cur < best[1]:
            keep = count > 0
            best = (cent[keep].cpu().numpy(), cur)
    return best


def predict_samples_batched(model, Xnew, S=1, batch=500, noise=None):
    """demos/demo_tf2.py:62-68: predict_samples over a long test set in chunks, stacked along the point axis.
    noise = (z_assign, u, z_pred), each [S, N, K], is cut along the point axis with the chunks."""
    X = to_device_f64(Xnew)
    if noise is not None:
        noise = tuple(to_device_f64(a, X.device) for a in noise)
    ys, fs = [], []
    for i in range(0, X.shape[0], batch):
        nz = None if noise is None else tuple(a[:, i:i + batch].contiguous() for a in noise)
        y, f = model.predict_samples(X[i:i + batch], S=S, noise=nz)
        ys.append(y)
        fs.append(f)
    return torch.cat(ys, 1), torch.cat(fs, 1)
