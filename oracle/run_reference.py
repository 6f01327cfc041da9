"""TEST INFRASTRUCTURE ONLY — run the reference's OWN MixtureGPs package, unmodified, from
/root/reference on the torch-backed TF/GPflow shim in oracle/shim.

This only works in the authoring container (where /root/reference is mounted).  It is used by
tests/golden/make_golden.py to produce the committed golden vectors, and by the CPU test-suite (when
/root/reference is present) to cross-check oracle/svgp_mixture.py.  Nothing on the GPU box may import it.

What is "the reference" here and what is restated:
  * executed as shipped: MixtureGPs/models.py, likelihoods.py, broadcasting_lik.py, utils.py
    (SGP/SMGP/SMGPModified/SVGPModified/IndependentPosteriorSingleOutputModified, GaussianModified,
    BroadcastingLikelihood, reparameterize) and utils/dataset_utils.py;
  * restated ([3P-memory]): the GPflow/TF/TFP functions those files call (oracle/shim/*).
Noise: tf.random.normal and TFP's uniform draw are served from a queue of explicit arrays (z, then u —
the order in SMGP._build_likelihood, models.py:72-73), because TF's Philox stream cannot be reproduced.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REFERENCE_ROOT = "/root/reference"
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "MixtureGPs"))


def _activate():
    if not available():
        raise RuntimeError("/root/reference is not mounted: the reference can only be run in the authoring container")
    for p in (REFERENCE_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import tensorflow as tf  # noqa: F401  (the shim)
    assert "oracle/shim" in tf.__file__.replace("\\", "/"), "a real TensorFlow is shadowing the shim"


def build_model(case: dict):
    """case: dict with keys model, lik, K, S, num_data, pred{...}, assign{...}, lik_var, assign_lik_var
    (constrained values as numpy arrays).  Returns the reference model object."""
    _activate()
    import gpflow
    from MixtureGPs.likelihoods import GaussianModified
    from MixtureGPs.models import SMGP, SMGPModified, SVGPModified

    K = int(case["K"])

    def make_layer(p, likelihood):
        ls = np.asarray(p["lengthscales"], dtype=np.float64)
        kern = gpflow.kernels.SquaredExponential(variance=float(p["variance"]),
                                                 lengthscales=float(ls) if ls.ndim == 0 else ls)
        layer = SVGPModified(kernel=kern, likelihood=likelihood, inducing_variable=np.asarray(p["Z"], dtype=np.float64),
                             num_latent_gps=K, whiten=True)
        layer.q_mu.assign(p["q_mu"])
        layer.q_sqrt.assign(np.tril(p["q_sqrt"]))
        return layer

    if case["model"] == "SMGP":
        lik = GaussianModified(variance=1.0, D=K)
        lik.variance.assign(np.asarray(case["lik_var"]).reshape(1, K))
        pred, assign = make_layer(case["pred"], lik), make_layer(case["assign"], lik)
        model = SMGP(likelihood=lik, pred_layer=pred, assign_layer=assign, K=K, num_samples=int(case["S"]),
                     num_data=case["num_data"])
    else:
        if case["lik"] == "multiclass":
            lik = gpflow.likelihoods.MultiClass(num_classes=K, invlink=gpflow.likelihoods.RobustMax(num_classes=K))
        else:
            lik = GaussianModified(variance=1.0, D=K)
            lik.variance.assign(np.asarray(case["lik_var"]).reshape(1, K))
        assign_lik = GaussianModified(variance=1.0, D=K)
        assign_lik.variance.assign(np.asarray(case["assign_lik_var"]).reshape(1, K))
        pred, assign = make_layer(case["pred"], lik), make_layer(case["assign"], assign_lik)
        model = SMGPModified(likelihood=lik, assign_likelihood=assign_lik, pred_layer=pred, assign_layer=assign,
                             K=K, num_samples=int(case["S"]), num_data=case["num_data"])
    return model


def _constrained_grad(param):
    """d/d(constrained value) from the gradient autograd left on the unconstrained leaf."""
    import torch
    g = param.unconstrained_variable.grad
    if g is None:
        g = torch.zeros_like(param.unconstrained_variable)
    name = type(param.transform).__name__
    if name == "_Softplus":       # y = softplus(x): dy/dx = sigmoid(x)
        return (g / torch.sigmoid(param.unconstrained_variable)).detach().numpy().copy()
    if name == "_FillTriangular":  # pure permutation into the lower triangle
        return param.transform.forward(g).detach().numpy().copy()
    return g.detach().numpy().copy()


def evaluate(case: dict, X, Y, z, u, Xtest=None, sample_noise=None) -> dict:
    """ELBO (= -_training_loss), gradients (constrained and unconstrained) and the predict_* outputs of the
    reference model for explicit noise z,u [S,N,K]."""
    _activate()
    import tensorflow as tf
    model = build_model(case)
    S, N, K = z.shape
    tf._noise.clear()
    tf._noise.push(z, u.reshape(1, S * N, K))
    loss = model._training_loss((np.asarray(X, dtype=np.float64), np.asarray(Y)))
    assert not tf._noise.queue, "the reference consumed less noise than supplied"
    (-loss).backward()
    out = {"elbo": float(-loss.detach())}
    for path, p in model.parameters_dict.items():
        if not p.trainable:
            continue
        key = (path.replace("pred_layer.", "pred.").replace("assign_layer.", "assign.")
               .replace("inducing_variable.Z", "Z").replace("kernel.", ""))
        if key.endswith("likelihood.variance"):
            # tf.Module walk reaches the shared GaussianModified through a layer
            key = "lik_var" if case["model"] == "SMGP" or key.startswith("pred.") else "assign_lik_var"
        out["grad." + key] = _constrained_grad(p)
        g = p.unconstrained_variable.grad
        out["gradu." + key] = (np.zeros(tuple(p.unconstrained_variable.shape)) if g is None else g.detach().numpy().copy())
    if Xtest is not None:
        import torch
        with torch.no_grad():
            Xtest = np.asarray(Xtest, dtype=np.float64)
            my, vy = model.predict_y(Xtest, S=2)
            assert torch.equal(my[0], my[1]) and torch.equal(vy[0], vy[1])   # S-invariance (models.py:36)
            out["predict_y.mean"], out["predict_y.var"] = my[0].numpy().copy(), vy[0].numpy().copy()
            Xt1 = model.integrate(Xtest, 1)[0]
            for name, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
                fm, fv = layer.predict_f(Xt1, full_cov=False)
                out[f"predict_f.{name}.mean"], out[f"predict_f.{name}.var"] = fm[0].numpy().copy(), fv[0].numpy().copy()
            pa = model.predict_assign(Xtest, S=3)
            out["predict_assign.probs"] = pa.numpy().copy()
            out["predict_assign.argmax"] = np.argmax(pa.numpy(), 1).astype(np.int64)   # demos: np.argmax(assign_, 1)
            if sample_noise is not None:
                za, us, zp = sample_noise
                S2, Nt, _ = za.shape
                tf._noise.clear()
                tf._noise.push(za, us.reshape(1, S2 * Nt, K), zp)        # order: models.py:57(W_dist), :95, :98
                sy, sf = model.predict_samples(Xtest, S=S2)
                assert not tf._noise.queue
                out["predict_samples.y"], out["predict_samples.f"] = sy.numpy().copy(), sf.numpy().copy()
    return out


def load_reference_dataset(name: str, seed: int = 0):
    """Call the reference's own loaders (utils/dataset_utils.py) unmodified."""
    _activate()
    from utils import dataset_utils
    rng = np.random.default_rng(seed)
    if name == "toy_multimodal":
        return dataset_utils.load_toy_multimodal_data(rng)           # dataset_utils.py:100-114
    if name == "toy_2d_categorical":
        return dataset_utils.load_toy_2d_data_categorical(rng)       # dataset_utils.py:149-165
    if name == "toy_2d":
        return dataset_utils.load_toy_2d_data(rng)                   # dataset_utils.py:128-146 (demo_tf2_2d.py:22)
    if name == "toy_categorical":
        return dataset_utils.load_toy_data_categorical(rng)          # dataset_utils.py:83-97 (demo_tf2_modified_multiclass.py:22)
    if name == "john_doe_boundary":
        cwd = os.getcwd()
        os.chdir(os.path.join(REFERENCE_ROOT, "demos"))              # loader reads "../data/..." (dataset_utils.py:43)
        try:
            np.random.seed(seed)   # train_test_split has no random_state (dataset_utils.py:76): pin the global RNG
            n, Xtr, Ytr, Xte, _ = dataset_utils.load_john_doe()      # demo_john_doe_multi_class.py:23
        finally:
            os.chdir(cwd)
        return n, Xtr.astype(np.float64), Ytr.astype(np.float64), Xte.astype(np.float64)
    if name == "john_doe_runs":
        cwd = os.getcwd()
        os.chdir(os.path.join(REFERENCE_ROOT, "demos"))              # loader reads "../data/..." (dataset_utils.py:10)
        try:
            np.random.seed(seed)   # train_test_split has no random_state (dataset_utils.py:33): pin the global RNG
            n, Xtr, Ytr, Xte, _ = dataset_utils.load_john_doe_runs()
        finally:
            os.chdir(cwd)
        return n, Xtr.astype(np.float64), Ytr.astype(np.float64), Xte.astype(np.float64)
    raise KeyError(name)
