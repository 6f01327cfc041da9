"""Generate tests/golden/*.npz by running the reference's own MixtureGPs code (unmodified, from
/root/reference) on the torch-backed TF/GPflow shim (oracle/run_reference.py).

Run in the authoring container only:   python tests/golden/make_golden.py
Each file holds the inputs (data batch, constrained parameters, explicit noise) and the reference's
outputs (ELBO, every gradient w.r.t. constrained and unconstrained variables, predict_f / predict_y /
predict_assign (+argmax) / predict_samples) plus cond(Kuu) per layer.

Cases mirror BASELINE.json `configs`:
  demo_tf2                          config #1  (demos/demo_tf2.py:24-49)            SMGP + GaussianModified, D=1, M=25, K=3, S=25
  demo_tf2_2d_modified_multiclass   config #2  (demos/demo_tf2_2d_modified_multiclass.py:25-53)  SMGPModified + MultiClass(RobustMax), D=2, K=2
  demo_john_doe                     config #3  (demos/demo_john_doe.py:29-56)       SMGP, D=2, M=25, K=4 on data/john_doe_dataset.csv
  synth4_small / synth5_small       configs #4/#5 scaled down (ARD lengthscales, jittered-grid / sampled Z)
  smgpmod_gauss_small               SMGPModified with Gaussian experts (demos/demo_tf2_modified.py)
each at GPflow-default initialisation ("init") and at a perturbed parameter state ("pert"), and (round 2)
  *_fullbatch.pert                  configs #1-#3 at the demos' OWN batch sizes (500 / 500 / 445 rows), perturbed state
  shape_d5_k1 / shape_d6_k7 / shape_d10_k2   shapes the kernels special-case and round 1 never exercised: D = 5, 6, 10
                                    put the Kuf exponent's row/column terms into the padding of the SECOND / THIRD k4 block
                                    of the z.x contraction (kuf_fold); K = 1 (softmax over one component) and K = 7 (odd K:
                                    unpaired Box-Muller branch, KP padding)
All cases now carry predict_samples (MultiClass experts included: models.py:91-103 through
MultiClass._predict_mean_and_var).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import run_reference as ref  # noqa: E402

BATCH = 192      # rows per golden batch (kept small so the fixtures stay a few hundred KB each)
NTEST = 40


def default_layer(Z, K, variance, lengthscales):
    M = Z.shape[0]
    return {"variance": np.float64(variance), "lengthscales": np.asarray(lengthscales, dtype=np.float64),
            "Z": np.asarray(Z, dtype=np.float64), "q_mu": np.zeros((M, K)),
            "q_sqrt": np.stack([np.eye(M)] * K)}


def perturb_layer(layer, rng, ard):
    M, D = layer["Z"].shape
    K = layer["q_mu"].shape[1]
    out = dict(layer)
    out["variance"] = np.float64(layer["variance"] * rng.uniform(0.7, 1.4))
    ls = np.asarray(layer["lengthscales"], dtype=np.float64)
    if ard and ls.ndim == 0:
        ls = ls * np.ones(D)
    out["lengthscales"] = ls * rng.uniform(0.8, 1.25, size=ls.shape)
    out["Z"] = layer["Z"] + 0.05 * rng.standard_normal((M, D)) * np.maximum(np.std(layer["Z"], 0), 1e-3)
    out["q_mu"] = 0.3 * rng.standard_normal((M, K))
    q = np.stack([np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))) for _ in range(K)])
    idx = np.arange(M)
    q[:, idx, idx] = np.abs(q[:, idx, idx]) + 0.05
    out["q_sqrt"] = q
    return out


def cond_kuu(layer):
    sys.path.insert(0, ROOT)
    from oracle import svgp_mixture as O
    return float(np.linalg.cond(O.kuu(O.layer_from_numpy(layer)).numpy()))


def emit(name, case, X, Y, Xtest, seed):
    rng = np.random.default_rng(seed)
    N, K, S = X.shape[0], case["K"], case["S"]
    tiny = np.finfo(np.float64).tiny
    z = rng.standard_normal((S, N, K))
    u = rng.uniform(tiny, 1.0, (S, N, K))
    S2, Nt = 7, Xtest.shape[0]
    sample_noise = (rng.standard_normal((S2, Nt, K)), rng.uniform(tiny, 1.0, (S2, Nt, K)), rng.standard_normal((S2, Nt, K)))
    out = ref.evaluate(case, X, Y, z, u, Xtest=Xtest, sample_noise=sample_noise)
    rec = {"meta.model": case["model"], "meta.lik": case["lik"], "meta.K": K, "meta.S": S,
           "meta.num_data": float(case["num_data"]), "X": X, "Y": np.asarray(Y, dtype=np.float64), "Xtest": Xtest,
           "z": z, "u": u, "cond.pred": cond_kuu(case["pred"]), "cond.assign": cond_kuu(case["assign"])}
    rec["sample.z_assign"], rec["sample.u"], rec["sample.z_pred"] = sample_noise
    for lname in ("pred", "assign"):
        for k, v in case[lname].items():
            rec[f"{lname}.{k}"] = np.asarray(v, dtype=np.float64)
    for k in ("lik_var", "assign_lik_var"):
        if case.get(k) is not None:
            rec[k] = np.asarray(case[k], dtype=np.float64).reshape(-1)
    for k, v in out.items():
        rec["out." + k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    if os.path.exists(path):          # leave byte-identical fixtures alone (np.savez stamps the current time)
        old = np.load(path)
        if sorted(old.files) == sorted(rec) and all(np.array_equal(np.asarray(old[k]), np.asarray(rec[k])) for k in rec):
            print(f"{name:44s} unchanged")
            return
    np.savez(path, **rec)
    print(f"{name:44s} elbo={out['elbo']:+.15e}  cond(Kuu)=({rec['cond.pred']:.2e},{rec['cond.assign']:.2e})  "
          f"{os.path.getsize(path) / 1024:.0f} KB")


def both_states(name, model, lik, K, S, num_data, pred0, assign0, lik_var0, assign_lik_var0, X, Y, Xtest, seed, ard,
                init=True):
    base = {"model": model, "lik": lik, "K": K, "S": S, "num_data": num_data}
    if init:
        emit(name + ".init", dict(base, pred=pred0, assign=assign0, lik_var=lik_var0, assign_lik_var=assign_lik_var0),
             X, Y, Xtest, seed)
    rng = np.random.default_rng(seed + 100)
    lv = None if lik_var0 is None else np.asarray(lik_var0) * rng.uniform(0.6, 1.5, size=K)
    alv = None if assign_lik_var0 is None else np.asarray(assign_lik_var0) * rng.uniform(0.6, 1.5, size=K)
    emit(name + ".pert", dict(base, pred=perturb_layer(pred0, rng, ard), assign=perturb_layer(assign0, rng, ard),
                              lik_var=lv, assign_lik_var=alv), X, Y, Xtest, seed + 1)


def main():
    from scipy.cluster.vq import kmeans

    # ---- config #1: demos/demo_tf2.py ------------------------------------------------------
    N, Xtr, Ytr, Xte = ref.load_reference_dataset("toy_multimodal", 0)
    Z, Za = kmeans(Xtr, 25, seed=0)[0], kmeans(Xtr, 25, seed=1)[0]           # demo_tf2.py:39
    sel = np.random.default_rng(10).choice(N, BATCH, replace=False)
    K = 3
    both_states("demo_tf2", "SMGP", "gaussian", K, 25, N,
                default_layer(Z, K, 0.5, 0.5), default_layer(Za, K, 0.1, 1.0), 0.5 * np.ones(K), None,
                Xtr[sel], Ytr[sel], Xte[:NTEST], seed=11, ard=False)
    full = np.random.default_rng(13).choice(N, 500, replace=False)          # batch_size 500, demo_tf2.py:27
    both_states("demo_tf2_fullbatch", "SMGP", "gaussian", K, 25, N,
                default_layer(Z, K, 0.5, 0.5), default_layer(Za, K, 0.1, 1.0), 0.5 * np.ones(K), None,
                Xtr[full], Ytr[full], Xte[:NTEST], seed=14, ard=False, init=False)
    # the analytic known-answer needs the full 1500 points at init: keep it as its own file
    case = {"model": "SMGP", "lik": "gaussian", "K": K, "S": 2, "num_data": N,
            "pred": default_layer(Z, K, 0.5, 0.5), "assign": default_layer(Za, K, 0.1, 1.0),
            "lik_var": 0.5 * np.ones(K), "assign_lik_var": None}
    emit("demo_tf2_full.init", case, Xtr, Ytr, Xte[:4], seed=12)

    # ---- config #2: demos/demo_tf2_2d_modified_multiclass.py -------------------------------
    N, Xtr, Ytr, Xte = ref.load_reference_dataset("toy_2d_categorical", 0)
    Z, Za = kmeans(Xtr, 25, seed=0)[0], kmeans(Xtr, 25, seed=1)[0]
    sel = np.random.default_rng(20).choice(N, BATCH, replace=False)
    K = 2
    both_states("demo_tf2_2d_modified_multiclass", "SMGPModified", "multiclass", K, 25, N,
                default_layer(Z, K, 0.1, 1.0), default_layer(Za, K, 0.1, 1.0), None, 0.5 * np.ones(K),
                Xtr[sel], Ytr[sel], Xte[:NTEST], seed=21, ard=True)
    both_states("demo_tf2_2d_modified_multiclass_fullbatch", "SMGPModified", "multiclass", K, 25, N,   # the whole set is
                default_layer(Z, K, 0.1, 1.0), default_layer(Za, K, 0.1, 1.0), None, 0.5 * np.ones(K),  # one batch (:28)
                Xtr, Ytr, Xte[:NTEST], seed=23, ard=True, init=False)

    # ---- config #3: demos/demo_john_doe.py --------------------------------------------------
    N, Xtr, Ytr, Xte = ref.load_reference_dataset("john_doe_runs", 0)
    Z, Za = kmeans(Xtr, 25, seed=0)[0], kmeans(Xtr, 25, seed=1)[0]
    sel = np.random.default_rng(30).choice(N, BATCH, replace=False)
    K = 4
    both_states("demo_john_doe", "SMGP", "gaussian", K, 25, N,
                default_layer(Z, K, 0.1, 1.0), default_layer(Za, K, 0.1, 1.0), 0.5 * np.ones(K), None,
                Xtr[sel], Ytr[sel], Xte[:NTEST], seed=31, ard=True)
    both_states("demo_john_doe_fullbatch", "SMGP", "gaussian", K, 25, N,                              # 445 train rows
                default_layer(Z, K, 0.1, 1.0), default_layer(Za, K, 0.1, 1.0), 0.5 * np.ones(K), None,
                Xtr, Ytr, Xte[:NTEST], seed=33, ard=True, init=False)

    # ---- config #4 scaled down: D=2, M=64 (8x8 jittered grid), K=4, S=16, ARD ---------------
    rng = np.random.default_rng(40)
    N, D, K, S = BATCH, 2, 4, 16
    X = rng.uniform(0, 8, (N, D))
    comp = rng.integers(0, K, N)
    om, ph = np.random.default_rng(41).uniform(0.5, 1.5, (K, D)), np.random.default_rng(41).uniform(0, 6, K)
    Y = (np.sin((X * om[comp]).sum(1) + ph[comp]) + 1.5 * comp + 0.1 * rng.standard_normal(N))[:, None]
    g = np.linspace(0.5, 7.5, 8)
    grid = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, 2)
    Zp, Zq = grid + rng.uniform(-0.2, 0.2, grid.shape), grid + rng.uniform(-0.2, 0.2, grid.shape)
    both_states("synth4_small", "SMGP", "gaussian", K, S, 4096,
                default_layer(Zp, K, 1.0, [1.0, 1.0]), default_layer(Zq, K, 0.5, [1.5, 1.5]),
                0.1 + 0.05 * np.arange(K), None, X, Y, rng.uniform(0, 8, (NTEST, D)), seed=42, ard=True)

    # ---- config #5 scaled down: D=8, M=48 (rows of X), K=8, S=8 ------------------------------
    rng = np.random.default_rng(50)
    N, D, K, S = 128, 8, 8, 8
    X = rng.standard_normal((N, D))
    comp = rng.integers(0, K, N)
    om, ph = np.random.default_rng(51).uniform(0.5, 1.5, (K, D)), np.random.default_rng(51).uniform(0, 6, K)
    Y = (np.sin((X * om[comp]).sum(1) + ph[comp]) + 1.5 * comp + 0.1 * rng.standard_normal(N))[:, None]
    Xpool = rng.standard_normal((400, D))
    Zp, Zq = Xpool[:48], Xpool[48:96]
    both_states("synth5_small", "SMGP", "gaussian", K, S, 8192,
                default_layer(Zp, K, 1.0, 2.5 * np.ones(D)), default_layer(Zq, K, 0.5, 3.0 * np.ones(D)),
                0.1 + 0.05 * np.arange(K), None, X, Y, rng.standard_normal((NTEST, D)), seed=52, ard=True)

    # ---- SMGPModified with Gaussian experts (demos/demo_tf2_modified.py:41-51) ---------------
    N, Xtr, Ytr, Xte = ref.load_reference_dataset("toy_multimodal", 0)
    Z, Za = kmeans(Xtr, 20, seed=2)[0], kmeans(Xtr, 20, seed=3)[0]
    sel = np.random.default_rng(60).choice(N, 96, replace=False)
    K = 3
    both_states("smgpmod_gauss_small", "SMGPModified", "gaussian", K, 10, N,
                default_layer(Z, K, 0.5, 0.5), default_layer(Za, K, 0.1, 1.0), 0.5 * np.ones(K), 0.5 * np.ones(K),
                Xtr[sel], Ytr[sel], Xte[:NTEST], seed=61, ard=False)

    # ---- shapes the kernels special-case (kuf_fold in the 2nd / 3rd k4 block; K = 1; odd K) ---------------------
    for name, D, K, S, M, model in (("shape_d5_k1", 5, 1, 6, 40, "SMGP"), ("shape_d6_k7", 6, 7, 5, 33, "SMGP"),
                                    ("shape_d10_k2", 10, 2, 6, 36, "SMGPModified")):
        rng = np.random.default_rng(70 + D)
        N = 96
        X = rng.standard_normal((N, D))
        comp = rng.integers(0, K, N)
        om, ph = rng.uniform(0.5, 1.5, (K, D)), rng.uniform(0, 6, K)
        Y = (np.sin((X * om[comp]).sum(1) + ph[comp]) + 1.5 * comp + 0.1 * rng.standard_normal(N))[:, None]
        Xpool = rng.standard_normal((2 * M, D))
        both_states(name, model, "gaussian", K, S, 4096,
                    default_layer(Xpool[:M], K, 1.0, 2.0 * np.ones(D)), default_layer(Xpool[M:], K, 0.5, 2.5 * np.ones(D)),
                    0.1 + 0.05 * np.arange(K), (0.3 + 0.05 * np.arange(K)) if model == "SMGPModified" else None,
                    X, Y, rng.standard_normal((NTEST, D)), seed=80 + D, ard=True, init=False)


if __name__ == "__main__":
    main()
