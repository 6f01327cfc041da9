#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — NOT RUNNABLE IN THIS IMAGE (TensorFlow / GPflow / TFP are not installable here).

Check the committed golden vectors (tests/golden/*.npz) against the TRUE reference: the unmodified MixtureGPs package of
LouieMiddle/ModulatedGPs on the REAL pinned stack (tensorflow 2.10.1, gpflow 2.7.0, tensorflow-probability 0.18.0,
environment.yml:95,130,133).  The goldens were produced by running the reference's own four files on a torch-backed
stand-in for that stack (oracle/shim, oracle/run_reference.py), so the third-party arithmetic behind them is a
restatement (SURVEY.md §8c, Appendix A "[3P-memory]").  On a machine that has the real stack this script closes that gap:

    python oracle/verify_goldens_real_tf.py /path/to/ModulatedGPs [golden-name ...]
    python oracle/verify_goldens_real_tf.py --shim /root/reference [golden-name ...]     # self-test of this script on the stand-ins

For every fixture it rebuilds the reference model from the fixture's constrained parameter values, serves the fixture's
explicit noise to the two places the reference draws randomness — tf.random.normal (models.py:57,98) and TFP's uniform
draw inside ExpRelaxedOneHotCategorical._sample_n (reached from models.py:60,73,94) — and compares ELBO, every gradient
w.r.t. the unconstrained variables, predict_f / predict_y / predict_assign (+ argmax) / predict_samples with the stored
outputs at 1e-9 relative (norm-wise per tensor, argmax exact: BASELINE.md §4).

It also prints the source line of gpflow's RobustMax.prob_is_largest that squashes the CDFs: the ONE third-party
constant this repo could not read from source (include/mgp.h: MGP_ROBUSTMAX_CDF_SQUASH = 1e-4).
"""
from __future__ import annotations

import glob
import inspect
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
RTOL = 1e-9


class NoiseQueue:
    """Explicit arrays handed out, in call order, instead of TF's / TFP's random draws."""

    def __init__(self):
        self.queue = []

    def push(self, *arrays):
        self.queue.extend(np.asarray(a, dtype=np.float64) for a in arrays)

    def pop(self, shape, what):
        if not self.queue:
            raise RuntimeError(f"the reference asked for more noise than supplied ({what}, shape {tuple(shape)})")
        a = self.queue.pop(0)
        want = tuple(int(s) for s in shape)
        if a.size != int(np.prod(want)):
            raise RuntimeError(f"{what}: the reference asked for shape {want}, the fixture's next array has {a.shape}")
        return a.reshape(want)


def patch_noise(tf, tfp, noise):
    """tf.random.normal -> queued z ; the uniform sampler TFP's relaxed one-hot uses -> queued u."""
    def normal(shape, mean=0.0, stddev=1.0, dtype=tf.float32, seed=None, name=None):
        shape = [int(s) for s in (shape.numpy() if hasattr(shape, "numpy") else shape)]
        return tf.constant(noise.pop(shape, "tf.random.normal"), dtype=dtype) * stddev + mean
    tf.random.normal = normal
    from tensorflow_probability.python.distributions import relaxed_onehot_categorical as roc

    class _Samplers:
        def __init__(self, real):
            self._real = real

        def __getattr__(self, name):
            return getattr(self._real, name)

        def uniform(self, shape, minval=0, maxval=None, dtype=tf.float32, seed=None, name=None):
            shape = [int(s) for s in (shape.numpy() if hasattr(shape, "numpy") else shape)]
            return tf.constant(noise.pop(shape, "tfp samplers.uniform"), dtype=dtype)
    roc.samplers = _Samplers(roc.samplers)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: d[k] for k in d.files}
    case = {"model": str(g["meta.model"]), "lik": str(g["meta.lik"]), "K": int(g["meta.K"]), "S": int(g["meta.S"]),
            "num_data": float(g["meta.num_data"])}
    for lname in ("pred", "assign"):
        case[lname] = {k: g[f"{lname}.{k}"] for k in ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")}
    case["lik_var"], case["assign_lik_var"] = g.get("lik_var"), g.get("assign_lik_var")
    return case, g


def build_model(case):
    """The same construction oracle/run_reference.py::build_model performs on the stand-ins, on the real classes."""
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

    def gaussian(var):
        lik = GaussianModified(variance=1.0, D=K)
        lik.variance.assign(np.asarray(var).reshape(1, K))
        return lik

    if case["model"] == "SMGP":
        lik = gaussian(case["lik_var"])
        return SMGP(likelihood=lik, pred_layer=make_layer(case["pred"], lik), assign_layer=make_layer(case["assign"], lik),
                    K=K, num_samples=int(case["S"]), num_data=case["num_data"])
    if case["lik"] == "multiclass":
        lik = gpflow.likelihoods.MultiClass(num_classes=K, invlink=gpflow.likelihoods.RobustMax(num_classes=K))
    else:
        lik = gaussian(case["lik_var"])
    assign_lik = gaussian(case["assign_lik_var"])
    return SMGPModified(likelihood=lik, assign_likelihood=assign_lik, pred_layer=make_layer(case["pred"], lik),
                        assign_layer=make_layer(case["assign"], assign_lik), K=K, num_samples=int(case["S"]),
                        num_data=case["num_data"])


def golden_key(path, case):
    key = (path.lstrip(".").replace("pred_layer.", "pred.").replace("assign_layer.", "assign.")
           .replace("inducing_variable.Z", "Z").replace("kernel.", ""))
    if key.endswith("likelihood.variance"):
        key = "lik_var" if case["model"] == "SMGP" or key.startswith("pred.") else "assign_lik_var"
    return key


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        a = a.reshape(b.shape)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), np.finfo(np.float64).tiny)) if b.size else 0.0


def to_np(x):
    """EagerTensor -> numpy (the stand-ins' tensors are torch tensors that may carry a graph)."""
    return x.detach().numpy() if hasattr(x, "detach") else x.numpy()


def evaluate(tf, gpflow, noise, case, g):
    model = build_model(case)
    z, u = g["z"], g["u"]
    S, N, K = z.shape
    noise.queue.clear()
    noise.push(z, u.reshape(1, S * N, K))                      # order in SMGP._build_likelihood, models.py:72-73
    variables = model.trainable_variables
    with tf.GradientTape() as tape:
        loss = model._training_loss((g["X"], g["Y"]))
    assert not noise.queue, "the reference consumed less noise than supplied"
    grads = tape.gradient(loss, variables)
    out = {"elbo": -float(to_np(loss))}
    by_var = {id(p.unconstrained_variable): golden_key(path, case)
              for path, p in gpflow.utilities.parameter_dict(model).items() if p.trainable}
    # the likelihood variances are shared objects reached through several module paths: name them by
    # identity — pred-side likelihood -> lik_var, SMGPModified.assign_likelihood -> assign_lik_var
    def lik_variance(wrapper):
        inner = getattr(wrapper, "likelihood", None)
        return getattr(inner, "variance", None)
    pv = lik_variance(model.likelihood)
    if pv is not None:
        by_var[id(pv.unconstrained_variable)] = "lik_var"
    av = lik_variance(getattr(model, "assign_likelihood", None)) if hasattr(model, "assign_likelihood") else None
    if av is not None and av is not pv:
        by_var[id(av.unconstrained_variable)] = "assign_lik_var"
    for v, gr in zip(variables, grads):
        out["gradu." + by_var[id(v)]] = np.zeros(tuple(v.shape)) if gr is None else -to_np(gr)
    Xtest = g["Xtest"]
    my, vy = model.predict_y(Xtest, S=2)
    out["predict_y.mean"], out["predict_y.var"] = to_np(my[0]), to_np(vy[0])
    Xt1 = model.integrate(Xtest, 1)[0]
    for name, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
        fm, fv = layer.predict_f(Xt1, full_cov=False)
        out[f"predict_f.{name}.mean"], out[f"predict_f.{name}.var"] = to_np(fm[0]), to_np(fv[0])
    pa = to_np(model.predict_assign(Xtest, S=3))
    out["predict_assign.probs"], out["predict_assign.argmax"] = pa, np.argmax(pa, 1).astype(np.int64)
    if "sample.z_assign" in g:
        za, us, zp = g["sample.z_assign"], g["sample.u"], g["sample.z_pred"]
        S2, Nt, _ = za.shape
        noise.queue.clear()
        noise.push(za, us.reshape(1, S2 * Nt, K), zp)          # order: models.py:57 (W_dist), :95, :98
        sy, sf = model.predict_samples(Xtest, S=S2)
        assert not noise.queue
        out["predict_samples.y"], out["predict_samples.f"] = to_np(sy), to_np(sf)
    return out


class ShimNoise:
    """--shim self-test: the stand-ins draw from their own queue (tf._noise); same interface as NoiseQueue."""

    def __init__(self, tf):
        self.tf = tf

    @property
    def queue(self):
        return self.tf._noise.queue

    def push(self, *arrays):
        self.tf._noise.push(*arrays)


def main(argv):
    shim = "--shim" in argv
    argv = [a for a in argv if a != "--shim"]
    if len(argv) < 2:
        print(__doc__)
        return 2
    ref_root = os.path.abspath(argv[1])
    if not os.path.isdir(os.path.join(ref_root, "MixtureGPs")):
        raise SystemExit(f"{ref_root} is not a checkout of LouieMiddle/ModulatedGPs (no MixtureGPs/)")
    sys.path.insert(0, ref_root)
    if shim:
        # self-test of THIS script's logic (model construction, noise order, gradient naming, comparisons) on the
        # stand-ins that produced the goldens: must report agreement to the last digit.  Proves nothing about GPflow.
        sys.path.insert(0, os.path.join(HERE, "shim"))
    import tensorflow as tf
    import tensorflow_probability as tfp
    import gpflow
    on_shim = any("oracle/shim" in (mod.__file__ or "").replace("\\", "/") for mod in (tf, gpflow))
    if on_shim != shim:
        raise SystemExit("the torch-backed stand-in is on the path: this script needs the REAL TensorFlow / GPflow"
                         if on_shim else "--shim given but a real TensorFlow / GPflow was imported")
    if shim:
        print("SELF-TEST on the torch-backed stand-ins (oracle/shim): checks this script, not GPflow")
        noise = ShimNoise(tf)
    else:
        print(f"tensorflow {tf.__version__}, gpflow {gpflow.__version__}, tensorflow-probability {tfp.__version__} "
              "(pinned: 2.10.1 / 2.7.0 / 0.18.0)")
        src = inspect.getsource(gpflow.likelihoods.RobustMax.prob_is_largest)
        print("gpflow RobustMax.prob_is_largest squashes its CDFs with:")
        for line in src.splitlines():
            if "cdfs" in line and "*" in line and "+" in line:
                print("    " + line.strip() + "        <- include/mgp.h ships MGP_ROBUSTMAX_CDF_SQUASH = 1e-4")
        noise = NoiseQueue()
        patch_noise(tf, tfp, noise)
    names = argv[2:] or sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                               if not os.path.basename(p).startswith("hp_"))
    worst_all, failed = 0.0, []
    for name in names:
        case, g = load_golden(name)
        out = evaluate(tf, gpflow, noise, case, g)
        worst, where = 0.0, ""
        for k, v in out.items():
            ref = g.get("out." + k)
            if ref is None:
                continue
            if k.endswith("argmax"):
                e = 0.0 if np.array_equal(np.asarray(v).reshape(-1), ref.reshape(-1)) else np.inf
            elif k == "elbo":
                e = abs(v - float(ref)) / max(abs(float(ref)), 1e-300)
            else:
                e = relerr(v, ref)
            if e > worst:
                worst, where = e, k
        ok = worst <= RTOL
        worst_all = max(worst_all, worst)
        if not ok:
            failed.append(name)
        print(f"{name:48s} worst rel. err. {worst:9.2e} ({where})  {'ok' if ok else 'MISMATCH'}")
    print(f"{len(names) - len(failed)} / {len(names)} fixtures agree with the {'stand-ins (self-test)' if shim else 'real reference'} "
          f"at {RTOL:g}; worst {worst_all:.2e}")
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
