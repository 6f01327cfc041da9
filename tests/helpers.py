"""Shared helpers for the parity tests: golden-file loading and the norm-wise relative error used as
the parity bar (||a-b||_inf / max(||b||_inf, tiny) <= 1e-9 per tensor, argmax exact)."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9   # north_star: "within 1e-9 relative in float64"


def golden_names():
    """The model-case fixtures (hp_*.npz are high-precision spot-check fixtures with their own layout)."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith("hp_"))


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: d[k] for k in d.files}
    case = {"model": str(g["meta.model"]), "lik": str(g["meta.lik"]), "K": int(g["meta.K"]), "S": int(g["meta.S"]),
            "num_data": float(g["meta.num_data"])}
    for lname in ("pred", "assign"):
        case[lname] = {k: g[f"{lname}.{k}"] for k in ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")}
    case["lik_var"] = g.get("lik_var")
    case["assign_lik_var"] = g.get("assign_lik_var")
    return case, g


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), np.finfo(np.float64).tiny))


def grad_keys(g):
    return sorted(k[len("out.grad."):] for k in g if k.startswith("out.grad."))
