"""modulatedgps_b200.tf_adapter — the TF2/GPflow binding (SURVEY.md §8b "Python side") — exercised without TensorFlow:

  CPU  the adapter runs on the torch-backed tensorflow / gpflow stand-ins of oracle/shim with a backend that reads the
       marshalled C structs back through their raw pointers (as libmgp does), evaluates the CPU oracle on what it finds
       there and writes the gradients through the gradient-struct pointers.  What comes out of `loss.backward()` on the
       GPflow-style Parameters must be the golden d(-ELBO)/d(unconstrained variable): this pins struct marshalling,
       gradient slot ordering, shared-parameter accumulation and the bijector chain rule of the custom gradient.
       When /root/reference is mounted the same is done on the reference's OWN model objects (`attach(model)`).
  GPU  the same model objects through the real backend (torch <-> DLPack hand-off, one mgp_elbo_fwd_bwd per loss) against
       the golden vectors at 1e-9.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from tests.helpers import RTOL, grad_keys, load_golden, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "shim")
CASES = ["demo_tf2.pert", "demo_tf2_2d_modified_multiclass.pert", "synth4_small.pert", "smgpmod_gauss_small.pert", "shape_d5_k1.pert"]


def _activate_shim():
    if SHIM not in sys.path:
        sys.path.insert(0, SHIM)
    import tensorflow as tf
    assert "oracle/shim" in tf.__file__.replace("\\", "/"), "a real TensorFlow shadows the shim"
    return tf


def _shim_model(case):
    """An SMGP / SMGPModified-shaped object made of the shim's GPflow classes only (no /root/reference needed): the
    attributes the adapter reads are the ones the reference's constructors set (MixtureGPs/models.py:29-33,49-53,107-110)."""
    _activate_shim()
    import gpflow
    from gpflow.base import Module, Parameter
    from gpflow.models import SVGP
    from gpflow.utilities import positive

    class GaussianModified(Module):                       # MixtureGPs/likelihoods.py:12-19
        def __init__(self, variance):
            self.variance = Parameter(np.asarray(variance, dtype=np.float64).reshape(1, -1), transform=positive())

    class Broadcasting:                                   # MixtureGPs/broadcasting_lik.py:14-20 (a plain object)
        def __init__(self, likelihood):
            self.likelihood = likelihood

    K = int(case["K"])

    def layer(p, lik):
        ls = np.asarray(p["lengthscales"], dtype=np.float64)
        kern = gpflow.kernels.SquaredExponential(variance=float(p["variance"]), lengthscales=float(ls) if ls.ndim == 0 else ls)
        svgp = SVGP(kernel=kern, likelihood=lik, inducing_variable=np.asarray(p["Z"], dtype=np.float64), num_latent_gps=K,
                    whiten=True)
        svgp.q_mu.assign(p["q_mu"])
        svgp.q_sqrt.assign(np.tril(p["q_sqrt"]))
        return svgp

    class Model(Module):
        pass

    model = Model()
    if case["lik"] == "multiclass":
        lik = gpflow.likelihoods.MultiClass(K, invlink=gpflow.likelihoods.RobustMax(K))
    else:
        lik = GaussianModified(case["lik_var"])
    model.likelihood = Broadcasting(lik)
    if case["model"] == "SMGP":
        model.pred_layer, model.assign_layer = layer(case["pred"], lik), layer(case["assign"], lik)
    else:
        alik = GaussianModified(case["assign_lik_var"])
        model.assign_likelihood = Broadcasting(alik)
        model.pred_layer, model.assign_layer = layer(case["pred"], lik), layer(case["assign"], alik)
    model.K, model.num_samples, model.num_data = K, int(case["S"]), case["num_data"]
    return model


def _read(ptr, shape):
    n = int(np.prod(shape)) if len(shape) else 1
    return np.ctypeslib.as_array((C.c_double * n).from_address(ptr)).reshape(shape).copy()


def _write(ptr, a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    C.memmove(ptr, a.ctypes.data, a.nbytes)


class OracleBackend:
    """Stands where libmgp stands, on the CPU: everything is read from / written to the raw pointers of the structs."""
    device = None

    def __init__(self):
        self.calls = []

    def launch(self, m):
        from oracle import svgp_mixture as O
        self.calls.append(m)

        def layer(s):
            return {"variance": _read(s.variance, ())[()], "lengthscales": _read(s.lengthscales, (s.n_lengthscales,)),
                    "Z": _read(s.Z, (s.M, s.D)), "q_mu": _read(s.q_mu, (s.M, s.K)), "q_sqrt": _read(s.q_sqrt, (s.K, s.M, s.M))}

        K, S, N, D = m.pred.K, m.cfg.S, m.N, m.pred.D
        pred, assign = layer(m.pred), layer(m.assign)
        X, Y = _read(m.X.data_ptr(), (N, D)), _read(m.Y.data_ptr(), (N, 1))
        z, u = _read(m.noise.z, (S, N, K)), _read(m.noise.u, (S, N, K))
        lv = None if m.lik_var is None else O.as_t(_read(m.lik_var.data_ptr(), (K,)))
        alv = None if m.assign_lik_var is None else O.as_t(_read(m.assign_lik_var.data_ptr(), (K,)))
        elbo, g = O.elbo_and_grads("SMGP" if m.cfg.model == 0 else "SMGPModified", "gaussian" if m.cfg.lik == 0 else "multiclass",
                                   O.layer_from_numpy(pred), O.layer_from_numpy(assign), lv, alv, X, Y, z, u, m.cfg.num_data,
                                   temperature=m.cfg.temperature, n_total=m.cfg.n_global)
        _write(m.elbo.data_ptr(), [elbo])
        for lname, gs in (("pred", m.pred_grad), ("assign", m.assign_grad)):
            for key in ("Z", "q_mu", "q_sqrt", "variance", "lengthscales"):
                _write(getattr(gs, key), g[f"{lname}.{key}"])
        if "lik_var" in g:
            _write(m.glik.data_ptr(), g["lik_var"])
        if "assign_lik_var" in g:
            _write(m.galik.data_ptr(), g["assign_lik_var"])


def _golden_gradu(model, g):
    """{golden key: Parameter} for the shim / reference model objects."""
    out = {}
    for lname, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
        out[f"{lname}.variance"], out[f"{lname}.lengthscales"] = layer.kernel.variance, layer.kernel.lengthscales
        out[f"{lname}.Z"], out[f"{lname}.q_mu"], out[f"{lname}.q_sqrt"] = layer.inducing_variable.Z, layer.q_mu, layer.q_sqrt
    if hasattr(model.likelihood.likelihood, "variance"):
        out["lik_var"] = model.likelihood.likelihood.variance
    if hasattr(model, "assign_likelihood"):
        out["assign_lik_var"] = model.assign_likelihood.likelihood.variance
    return out


def _check_against_golden(model, g, loss_fn, tol):
    for p in model.trainable_parameters:
        p.unconstrained_variable.grad = None
    loss = loss_fn((g["X"], g["Y"]))
    assert abs(float(loss) + float(g["out.elbo"])) <= tol * abs(float(g["out.elbo"]))
    (2.0 * loss).backward()                                # upstream gradient 2: the custom gradient must scale by it
    params = _golden_gradu(model, g)
    for k in grad_keys(g):
        ref = -2.0 * g["out.gradu." + k]
        mine = params[k].unconstrained_variable.grad.detach().cpu().numpy().reshape(ref.shape)
        if np.max(np.abs(ref)) < 1e-12:
            assert np.max(np.abs(mine)) < 1e-11, k
        else:
            assert relerr(mine, ref) <= tol, (k, relerr(mine, ref))


@pytest.mark.parametrize("name", CASES)
def test_adapter_marshals_and_chains_on_the_shim(name):
    from modulatedgps_b200 import _lib, tf_adapter
    case, g = load_golden(name)
    model = _shim_model(case)
    backend = OracleBackend()
    loss_fn = tf_adapter.build_training_loss(model, backend=backend, noise=lambda S, N, K: (g["z"], g["u"]))
    _check_against_golden(model, g, loss_fn, 1e-10)         # oracle vs reference-on-shim: same arithmetic, ~1e-13
    m = backend.calls[-1]
    # the structs carry what include/mgp.h says they carry
    M, D = case["pred"]["Z"].shape
    assert (m.pred.M, m.pred.D, m.pred.K) == (M, D, case["K"])
    assert m.pred.n_lengthscales == np.asarray(case["pred"]["lengthscales"]).size
    assert (m.cfg.model, m.cfg.lik, m.cfg.S) == (0 if case["model"] == "SMGP" else 1, 0 if case["lik"] == "gaussian" else 1, case["S"])
    assert m.cfg.temperature == 1e-2 and m.cfg.num_data == case["num_data"] and m.cfg.n_global == g["X"].shape[0]
    assert np.array_equal(_read(m.pred.q_sqrt, (case["K"], M, M)), np.tril(case["pred"]["q_sqrt"]))
    assert m.noise.z and m.noise.u
    assert sorted(loss_fn.slots) == sorted(k for k in m.grads if m.grads[k] is not None and (k in loss_fn.slots))


def test_adapter_philox_mode_and_frozen_parameters():
    """No explicit noise -> NULL z/u and a fresh Philox seed per call; a non-trainable Parameter still reaches the struct
    (its constrained value) but gets no gradient slot in the custom gradient's outputs."""
    from modulatedgps_b200 import tf_adapter
    case, g = load_golden("demo_tf2.pert")
    model = _shim_model(case)
    frozen = model.assign_layer.kernel.lengthscales
    frozen.trainable = False
    frozen.unconstrained_variable.requires_grad_(False)

    class Recording(OracleBackend):
        def launch(self, m):
            self.calls.append(m)
            _write(m.elbo.data_ptr(), [-1.5])
            for t in m.grads.values():
                t.fill_(1.0)

    backend = Recording()
    loss_fn = tf_adapter.build_training_loss(model, backend=backend, seed=7)
    assert all(v is not frozen.unconstrained_variable for v in loss_fn.variables)
    l1 = loss_fn((g["X"], g["Y"]))
    l2 = loss_fn((g["X"], g["Y"]))
    assert float(l1) == 1.5 and float(l2) == 1.5
    a, b = backend.calls
    assert not a.noise.z and not a.noise.u and a.noise.seed != b.noise.seed and (a.noise.seed >> 20) == 7
    assert _read(a.assign.lengthscales, (1,))[0] == pytest.approx(float(np.asarray(case["assign"]["lengthscales"]).reshape(-1)[0]), rel=1e-14)
    l2.backward()
    v = model.pred_layer.kernel.variance
    # softplus chain rule on a unit constrained gradient: d(-ELBO)/du = -sigmoid(u)
    assert float(v.unconstrained_variable.grad) == pytest.approx(-float(torch.sigmoid(v.unconstrained_variable.detach())), rel=1e-14)
    assert frozen.unconstrained_variable.grad is None


def test_adapter_on_the_reference_model_objects():
    """`attach(model)` on an object built by the reference's own constructors (needs /root/reference)."""
    from oracle import run_reference as ref
    if not ref.available():
        pytest.skip("/root/reference is not mounted")
    from modulatedgps_b200 import tf_adapter
    for name in ("demo_tf2.pert", "demo_tf2_2d_modified_multiclass.pert"):
        case, g = load_golden(name)
        model = ref.build_model(case)
        tf_adapter.attach(model, backend=OracleBackend(), noise=lambda S, N, K: (g["z"], g["u"]))
        _check_against_golden(model, g, model._training_loss, 1e-10)
        assert float(model._build_likelihood(g["X"], g["Y"])) == pytest.approx(float(g["out.elbo"]), rel=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_adapter_end_to_end_on_the_gpu(name):
    from modulatedgps_b200 import _lib, tf_adapter
    case, g = load_golden(name)
    model = _shim_model(case)
    before = _lib.total_launches()
    loss_fn = tf_adapter.build_training_loss(model, noise=lambda S, N, K: (g["z"], g["u"]))
    _check_against_golden(model, g, loss_fn, RTOL)
    assert _lib.total_launches() > before
    fm, fv = tf_adapter.predict_f(model.pred_layer, g["Xtest"])
    assert relerr(fm.cpu().numpy(), g["out.predict_f.pred.mean"]) <= RTOL
    assert relerr(fv.cpu().numpy(), g["out.predict_f.pred.var"]) <= RTOL
