#!/usr/bin/env python
"""Replay the reference's demos (BASELINE.json configs #1-#3) as full training runs on libmgp.

    python examples/replay_demos.py --demo tf2|multiclass|john_doe|tf2_2d|tf2_modified|tf2_modified_multiclass|john_doe_multi_class
                                    [--iters N] [--squash S] [--out trajectory.json]

Each replay is the demo's model block with the imports swapped for this package at the demo's own
hyper-parameters:
    tf2         demos/demo_tf2.py:24-58                          SMGP + GaussianModified, N 1500, batch 500, 2000 Adam steps
    multiclass  demos/demo_tf2_2d_modified_multiclass.py:25-56   SMGPModified + MultiClass(RobustMax), N 500, 2000 steps
    john_doe    demos/demo_john_doe.py:29-60                     SMGP, 445 train rows (one batch), 10000 steps
and the reference's four other demos (round 2; data and k-means centroids in tests/golden/datasets/demo_datasets.npz):
    tf2_2d                   demos/demo_tf2_2d.py:22-60                   SMGP, 2-D inputs, K 3, N 500, 2000 steps
    tf2_modified             demos/demo_tf2_modified.py:22-60             SMGPModified + Gaussian experts, N 1500, 4000 steps
    tf2_modified_multiclass  demos/demo_tf2_modified_multiclass.py:22-64  SMGPModified + MultiClass(RobustMax), 1-D, K 2, 2000 steps
    john_doe_multi_class     demos/demo_john_doe_multi_class.py:23-67     SMGPModified + MultiClass(RobustMax), 445 rows, K 2, 2000 steps
with lr 0.005, 25 MC samples, 25 inducing points from k-means (seeds 0 / 1), `DeviceMinibatches` standing in for
tf.data's shuffle(N).batch(B).repeat() and `run_adam` logging the ELBO of a fresh minibatch every 5 iterations, as
utils/training_utils.py:4-28 does.  The data sets are the outputs of the reference's OWN loaders
(utils/dataset_utils.py, run unmodified by tests/golden/make_golden.py) kept in the golden fixtures — the GPU box has no
/root/reference.  The published anchors are the ELBO curves of final_figs/ (BASELINE.md §1); they are the one piece of
evidence about the reference that does not come from this repo's restatement of GPflow.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

DEMOS = {
    # name: (fixture holding the loader's full training set, fixture holding the scipy k-means centroids the demo starts
    #        from, K, kernels (variance, lengthscale) pred / assign, iterations)
    "tf2": ("demo_tf2_full.init", "demo_tf2_full.init", 3, (0.5, 0.5), (0.1, 1.0), 2000),
    "multiclass": ("demo_tf2_2d_modified_multiclass_fullbatch.pert", "demo_tf2_2d_modified_multiclass.init", 2, (0.1, 1.0),
                   (0.1, 1.0), 2000),
    "john_doe": ("demo_john_doe_fullbatch.pert", "demo_john_doe.init", 4, (0.1, 1.0), (0.1, 1.0), 10000),
    # the four demos outside BASELINE.json's configs: data and centroids from tests/golden/datasets/demo_datasets.npz
    # (the reference's own loaders + scipy k-means, tests/golden/make_demo_datasets.py)
    "tf2_2d": (None, None, 3, (0.1, 1.0), (0.1, 1.0), 2000),
    "tf2_modified": (None, None, 3, (0.5, 0.5), (0.1, 1.0), 4000),
    "tf2_modified_multiclass": (None, None, 2, (0.1, 1.0), (0.1, 1.0), 2000),
    "john_doe_multi_class": (None, None, 2, (0.1, 1.0), (0.1, 1.0), 2000),
}
# model class / expert likelihood per demo (default: SMGP with GaussianModified experts)
KIND = {"multiclass": ("SMGPModified", "multiclass"), "tf2_modified": ("SMGPModified", "gaussian"),
        "tf2_modified_multiclass": ("SMGPModified", "multiclass"), "john_doe_multi_class": ("SMGPModified", "multiclass")}
DATASETS = os.path.join(GOLDEN, "datasets", "demo_datasets.npz")
# Read off final_figs/*.png: the ELBO run_adam logged at a few iterations, with the half-width of the
# band the test accepts.  One log is ONE minibatch of 500 under ONE draw of 25 relaxed one-hot samples per point: the
# figures' own point-to-point scatter is ~ +-0.1 early on (demo_tf2.png: -2.85, -2.75, -2.47, -2.40 at iterations 5-20),
# and two independent replays of demo_tf2 here (CPU oracle + torch Adam; kernels + fused Adam) gave -2.65 / -2.66 at
# iteration 5 and -2.72 / -2.62 at iteration 10, so the first-log band is +-0.25, not the +-0.1 a smooth curve would allow.
ANCHORS = {
    "tf2": {"first": (-2.85, 0.25), "final_at_least": -0.25, "figure": "final_figs/demo_tf2.png",
            "curve": {15: (-2.47, 0.2), 50: (-2.0, 0.2), 100: (-1.72, 0.2), 500: (-1.33, 0.2), 1000: (-0.72, 0.25),
                      1500: (-0.3, 0.25)}},
    "multiclass": {"first": (-4.3, 0.1), "final_at_least": 0.8, "figure": "final_figs/demo_tf2_2d_modified_multiclass_2.png",
                   "curve": {100: (-1.9, 0.2), 400: (-1.2, 0.2), 1000: (0.0, 0.3), 1450: (0.85, 0.3)}},
    # round 2: the four other demos.  final_figs/demo_tf2_2d_2.png: -228 at the first log, -85 / -22 / -6 at iterations
    # 250 / 500 / 1000, ~ -3 at 2000;  demo_tf2_modified.png: -5.35 -> -3 by iteration 150, plateau -2.8 (300-700), -1.2 at
    # 1500, ~ -1.0 from 2000 on;  demo_tf2_modified_multiclass.png: -4.5 -> -1.9 (100) -> -1.1 (250) -> -0.5 (500) -> +0.65
    # (1000) -> +1.3 (1500) -> +1.45 (2000), spikes to -8 / -13;  demo_JohnDoe_*_multi_class_2.png: -4.3 -> plateau -2.5
    # (100-500) -> -1.3 (1000) -> +0.3 (1500) -> +1.45 (2000).
    "tf2_2d": {"first": (-228.0, 12.0), "final_at_least": -6.0, "figure": "final_figs/demo_tf2_2d_2.png",
               "curve": {250: (-85.0, 15.0), 500: (-22.0, 8.0), 1000: (-6.0, 3.0)}},
    "tf2_modified": {"first": (-5.35, 0.5), "final_at_least": -1.6, "figure": "final_figs/demo_tf2_modified.png",
                     "curve": {150: (-3.0, 0.3), 500: (-2.8, 0.3), 1000: (-2.3, 0.4)}},
    "tf2_modified_multiclass": {"first": (-4.5, 0.3), "final_at_least": 1.0,
                                "figure": "final_figs/demo_tf2_modified_multiclass.png",
                                "curve": {100: (-1.9, 0.3), 250: (-1.1, 0.3), 500: (-0.5, 0.3), 1000: (0.65, 0.35),
                                          1500: (1.3, 0.4)}},
    "john_doe_multi_class": {"first": (-4.3, 0.3), "final_at_least": 0.9,
                             "figure": "final_figs/demo_JohnDoe_RightArmSeam_stumpsX_stumpsY_multi_class_2.png",
                             # (the published run's train / test split is unknown — dataset_utils.py:76 fixes no
                             #  random_state — and this replay leaves the plateau ~150 iterations earlier: only the
                             #  plateau's level is anchored)
                             "curve": {250: (-2.5, 0.35)}},
    "john_doe": {"first": (-6.0, 0.6), "final_at_least": 1.5,
                 "figure": "final_figs/demo_JohnDoe_RightArmSeam_stumpsX_stumpsY_2.png",
                 "curve": {1000: (-1.5, 0.4), 2000: (-1.5, 0.4)}},     # the plateau before the break-out (figure: ~4000)
}


def load_training_set(demo):
    if DEMOS[demo][0] is None:
        d = np.load(DATASETS)
        return np.asarray(d[f"{demo}.X"], dtype=np.float64), np.asarray(d[f"{demo}.Y"], dtype=np.float64).reshape(-1, 1)
    d = np.load(os.path.join(GOLDEN, DEMOS[demo][0] + ".npz"))
    return np.asarray(d["X"], dtype=np.float64), np.asarray(d["Y"], dtype=np.float64).reshape(-1, 1)


def build(demo, Xtrain, seed=0, inducing="reference"):
    """inducing = "reference": the centroids scipy.cluster.vq.kmeans(Xtrain, 25, seed=0 / 1) gave the reference's demo
    (recorded in the golden fixtures: the replay then starts from the demo's exact initial state); "device": this
    package's k-means (mgp_kmeans_iterate), i.e. the whole pipeline on the GPU."""
    import modulatedgps_b200 as mg
    _, zfix, K, (pv, pl), (av, al), _ = DEMOS[demo]
    num_ind, num_samples, num_data = 25, 25, Xtrain.shape[0]
    pred_kernel = mg.SquaredExponential(variance=pv, lengthscales=pl)
    assign_kernel = mg.SquaredExponential(variance=av, lengthscales=al)
    if inducing == "device":
        Z, Z_assign = mg.kmeans(Xtrain, num_ind, seed=0)[0], mg.kmeans(Xtrain, num_ind, seed=1)[0]
    elif zfix is None:
        d = np.load(DATASETS)
        Z, Z_assign = np.asarray(d[f"{demo}.Z"]), np.asarray(d[f"{demo}.Z_assign"])
    else:
        d = np.load(os.path.join(GOLDEN, zfix + ".npz"))
        Z, Z_assign = np.asarray(d["pred.Z"]), np.asarray(d["assign.Z"])
    model_kind, expert = KIND.get(demo, ("SMGP", "gaussian"))
    if model_kind == "SMGPModified":
        lik = (mg.MultiClass(num_classes=K, invlink=mg.RobustMax(num_classes=K)) if expert == "multiclass"
               else mg.GaussianModified(variance=0.5, D=K))
        assign_lik = mg.GaussianModified(variance=0.5, D=K)
        pred_layer = mg.SVGPModified(kernel=pred_kernel, likelihood=lik, inducing_variable=Z, num_latent_gps=K, whiten=True)
        assign_layer = mg.SVGPModified(kernel=assign_kernel, likelihood=assign_lik, inducing_variable=Z_assign,
                                       num_latent_gps=K, whiten=True)
        model = mg.SMGPModified(likelihood=lik, assign_likelihood=assign_lik, pred_layer=pred_layer,
                                assign_layer=assign_layer, K=K, num_samples=num_samples, num_data=num_data)
    else:
        lik = mg.GaussianModified(variance=0.5, D=K)
        pred_layer = mg.SVGPModified(kernel=pred_kernel, likelihood=lik, inducing_variable=Z, num_latent_gps=K, whiten=True)
        assign_layer = mg.SVGPModified(kernel=assign_kernel, likelihood=lik, inducing_variable=Z_assign, num_latent_gps=K,
                                       whiten=True)
        model = mg.SMGP(likelihood=lik, pred_layer=pred_layer, assign_layer=assign_layer, K=K, num_samples=num_samples,
                        num_data=num_data)
    model.seed = seed
    return model


def replay(demo, iters=None, squash=None, seed=0, quiet=True, inducing="reference"):
    """Returns {"iters": [...], "elbos": [...], ...}: what run_adam logged."""
    import modulatedgps_b200 as mg
    from modulatedgps_b200 import _lib
    Xtrain, Ytrain = load_training_set(demo)
    model = build(demo, Xtrain, seed, inducing)
    ctx = _lib.get_context()
    ctx.set_robustmax_squash(_lib.ROBUSTMAX_CDF_SQUASH if squash is None else squash)
    try:
        train_iter = mg.DeviceMinibatches(Xtrain, Ytrain, 500, seed=seed)
        num_iter = DEMOS[demo][5] if iters is None else int(iters)
        sink = io.StringIO()
        with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
            its, elbos = mg.run_adam(model, num_iter, train_iter, 0.005, compile=False)
        ctx.check_status()
    finally:
        ctx.set_robustmax_squash(_lib.ROBUSTMAX_CDF_SQUASH)
    Xte = Xtrain[:: max(1, Xtrain.shape[0] // 200)]
    assign = np.asarray(model.predict_assign(Xte))
    return {"demo": demo, "iters": list(map(int, its)), "elbos": list(map(float, elbos)), "num_iter": num_iter,
            "robustmax_squash": _lib.ROBUSTMAX_CDF_SQUASH if squash is None else squash, "anchors": ANCHORS[demo],
            "inducing": inducing, "seed": seed,
            "assign_argmax_counts": np.bincount(np.argmax(assign, 1), minlength=DEMOS[demo][2]).tolist()}


def summarise(rec):
    e = np.asarray(rec["elbos"])
    tail = e[-100:] if e.size >= 100 else e
    return {"first_log": float(e[0]), "median_last_100_logs": float(np.median(tail)), "max": float(e.max()), "min": float(e.min())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--demo", required=True, choices=sorted(DEMOS))
    ap.add_argument("--iters", type=int, default=None)
    ap.add_argument("--squash", type=float, default=None, help="RobustMax CDF squash (default: MGP_ROBUSTMAX_CDF_SQUASH)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--inducing", default="reference", choices=["reference", "device"])
    ap.add_argument("--out", default=None)
    ap.add_argument("--verbose", action="store_true", help="print run_adam's iteration log, as the reference does")
    args = ap.parse_args()
    rec = replay(args.demo, args.iters, args.squash, args.seed, quiet=not args.verbose, inducing=args.inducing)
    rec["summary"] = summarise(rec)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rec, f)
    print(json.dumps({k: rec[k] for k in ("demo", "num_iter", "robustmax_squash", "anchors", "summary", "assign_argmax_counts")}))


if __name__ == "__main__":
    main()
