"""Model construction from a plain parameter dictionary, and the synthetic workloads of BASELINE.json configs #4 / #5
(SURVEY.md §8d), shared by bench.py, tools/, __graft_entry__.smoke() and the tests.

A *case* is a dict with the keys the golden fixtures use:
    model "SMGP" | "SMGPModified", lik "gaussian" | "multiclass", K, S, num_data,
    pred / assign: {variance, lengthscales ([] or [D]), Z [M, D], q_mu [M, K], q_sqrt [K, M, M]}   (constrained values)
    lik_var [K] | None, assign_lik_var [K] | None
"""
from __future__ import annotations

import math

import numpy as np

CONFIG4 = {"N": 1 << 20, "D": 2, "M": 256, "K": 4, "S": 16}     # BASELINE.json configs[3]
CONFIG5 = {"N": 1 << 24, "D": 8, "M": 1024, "K": 8, "S": 32}    # BASELINE.json configs[4]


def model_from_case(case):
    """Build the SMGP / SMGPModified the case describes through the public classes, exactly as a demo's model block does
    (demos/demo_tf2.py:36-49, demo_tf2_2d_modified_multiclass.py:36-53)."""
    from . import (GaussianModified, MultiClass, RobustMax, SMGP, SMGPModified, SquaredExponential, SVGPModified)
    K = int(case["K"])

    def layer(p, lik):
        ls = np.asarray(p["lengthscales"], dtype=np.float64)
        kern = SquaredExponential(variance=float(p["variance"]), lengthscales=float(ls) if ls.ndim == 0 else ls)
        return SVGPModified(kernel=kern, likelihood=lik, inducing_variable=np.asarray(p["Z"]), num_latent_gps=K,
                            whiten=True, q_mu=np.asarray(p["q_mu"]), q_sqrt=np.tril(np.asarray(p["q_sqrt"])))

    def gauss(v):
        g = GaussianModified(variance=1.0, D=K)
        g.variance.assign(np.asarray(v, dtype=np.float64).reshape(1, K))
        return g

    if case["model"] == "SMGP":
        lik = gauss(case["lik_var"])
        return SMGP(likelihood=lik, pred_layer=layer(case["pred"], lik), assign_layer=layer(case["assign"], lik), K=K,
                    num_samples=int(case["S"]), num_data=case["num_data"])
    lik = MultiClass(K, invlink=RobustMax(K)) if case["lik"] == "multiclass" else gauss(case["lik_var"])
    alik = gauss(case["assign_lik_var"])
    return SMGPModified(likelihood=lik, assign_likelihood=alik, pred_layer=layer(case["pred"], lik),
                        assign_layer=layer(case["assign"], alik), K=K, num_samples=int(case["S"]),
                        num_data=case["num_data"])


def _variational_state(rng, M, K):
    """q_mu = 0.3 N(0,1); q_sqrt_k = I + 0.05 tril(N(0,1)) with the diagonal kept > 0."""
    q = np.stack([np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))) for _ in range(K)])
    idx = np.arange(M)
    q[:, idx, idx] = np.abs(q[:, idx, idx]) + 0.05
    return 0.3 * rng.standard_normal((M, K)), q


def _targets(rng, X, K, om, ph):
    comp = rng.integers(0, K, X.shape[0])
    return (np.sin((X * om[comp]).sum(1) + ph[comp]) + 1.5 * comp + 0.1 * rng.standard_normal(X.shape[0]))[:, None]


def config4_workload(n_points, seed=0, m=CONFIG4["M"], k=CONFIG4["K"], d=CONFIG4["D"], s=CONFIG4["S"], num_data=None):
    """Config #4 synthetic inputs and parameters: X ~ U[0, sqrt(M))^2, y = sin(w_c . x + phi_c) + 1.5 c +
    0.1 eps, Z = sqrt(M) x sqrt(M) jittered grid, pred kernel (1.0, [1, 1]), assign kernel (0.5, [1.5, 1.5])."""
    rng = np.random.default_rng(seed)
    side = int(round(math.sqrt(m)))
    X = rng.uniform(0.0, float(side), (n_points, d))
    r1 = np.random.default_rng(1)
    om, ph = r1.uniform(0.5, 1.5, (k, d)), r1.uniform(0, 2 * np.pi, k)
    Y = _targets(rng, X, k, om, ph)
    r2 = np.random.default_rng(2)
    g = np.linspace(0.5, side - 0.5, side)
    grid = np.stack(np.meshgrid(g, g, indexing="ij"), -1).reshape(-1, d)

    def layer(var, ls):
        Z = grid + r2.uniform(-0.2, 0.2, grid.shape)
        q = np.stack([np.eye(m) + 0.05 * np.tril(r2.standard_normal((m, m))) for _ in range(k)])
        idx = np.arange(m)
        q[:, idx, idx] = np.abs(q[:, idx, idx]) + 0.05
        return {"variance": np.float64(var), "lengthscales": np.asarray(ls, dtype=np.float64), "Z": Z,
                "q_mu": 0.3 * r2.standard_normal((m, k)), "q_sqrt": q}

    case = {"model": "SMGP", "lik": "gaussian", "K": k, "S": s, "num_data": float(num_data or CONFIG4["N"]),
            "pred": layer(1.0, [1.0] * d), "assign": layer(0.5, [1.5] * d), "lik_var": 0.1 + 0.05 * np.arange(k),
            "assign_lik_var": None}
    return case, X, Y


def config5_parameters(m=CONFIG5["M"], k=CONFIG5["K"], d=CONFIG5["D"], s=CONFIG5["S"], num_data=None, pool=1 << 16):
    """Config #5 parameters: Z = M distinct rows drawn (rng 2) from the first `pool` rows of the
    X ~ N(0, I_8) stream of rng 0 — every rank regenerates the same pool, so Z does not depend on the sharding —
    lengthscales [2.5] * 8 (pred) / [3.0] * 8 (assign), other parameters as config #4."""
    Xpool = np.random.default_rng(0).standard_normal((pool, d))
    r2 = np.random.default_rng(2)

    def layer(var, ls):
        Z = Xpool[r2.choice(pool, m, replace=False)]
        q_mu, q = _variational_state(r2, m, k)
        return {"variance": np.float64(var), "lengthscales": ls * np.ones(d), "Z": Z, "q_mu": q_mu, "q_sqrt": q}

    return {"model": "SMGP", "lik": "gaussian", "K": k, "S": s, "num_data": float(num_data or CONFIG5["N"]),
            "pred": layer(1.0, 2.5), "assign": layer(0.5, 3.0), "lik_var": 0.1 + 0.05 * np.arange(k),
            "assign_lik_var": None}


def config5_points(lo, hi, k=CONFIG5["K"], d=CONFIG5["D"], block=1 << 16):
    """Rows [lo, hi) of the config-#5 data set, generated block by block from per-block seeds so that a rank can make
    its own shard of the 2^24 points without materialising the rest (block b uses default_rng([0, b]))."""
    r1 = np.random.default_rng(1)
    om, ph = r1.uniform(0.5, 1.5, (k, d)), r1.uniform(0, 2 * np.pi, k)
    Xs, Ys = [], []
    b0, b1 = lo // block, (hi + block - 1) // block
    for b in range(b0, b1):
        rng = np.random.default_rng([0, b])
        Xb = rng.standard_normal((block, d))
        Yb = _targets(rng, Xb, k, om, ph)
        a, e = max(lo - b * block, 0), min(hi - b * block, block)
        Xs.append(Xb[a:e])
        Ys.append(Yb[a:e])
    return np.concatenate(Xs), np.concatenate(Ys)


def synthetic_case(N, D, M, K, S, seed, model="SMGP", ls_assign=1.1):
    """Config-#4-style (D = 2, square M: jittered grid) or config-#5-style (otherwise: Z sampled like X) workload at an
    N small enough for the CPU checker with explicit noise.  SURVEY.md asks for cond(Kuu) <~ 1e4 so that 1e-9 is above the conditioning noise
    floor; its assign lengthscale 1.5 on a unit grid gives cond 4e6, so the strict cases use 1.1 (cond 1.3e4)."""
    rng = np.random.default_rng(seed)
    side = int(round(math.sqrt(M)))
    if D == 2 and side * side == M:
        gx = np.linspace(0.5, side - 0.5, side)
        grid = np.stack(np.meshgrid(gx, gx, indexing="ij"), -1).reshape(-1, 2)
        Zp, Za = grid + rng.uniform(-0.2, 0.2, grid.shape), grid + rng.uniform(-0.2, 0.2, grid.shape)
        X = rng.uniform(0, side, (N, D))
        lsp, lsa = np.array([1.0, 1.0]), np.array([ls_assign, ls_assign])
    else:
        X = rng.standard_normal((N, D))
        pool = rng.standard_normal((2 * M, D))
        Zp, Za = pool[:M], pool[M:]
        lsp, lsa = 2.5 * np.ones(D), 3.0 * np.ones(D)
    comp = rng.integers(0, K, N)
    Y = (np.sin(X.sum(1) + comp) + 1.5 * comp + 0.1 * rng.standard_normal(N))[:, None]

    def layer(Z, var, ls):
        q = np.stack([np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))) for _ in range(K)])
        idx = np.arange(M)
        q[:, idx, idx] = np.abs(q[:, idx, idx]) + 0.05
        return {"variance": np.float64(var), "lengthscales": ls, "Z": Z, "q_mu": 0.3 * rng.standard_normal((M, K)), "q_sqrt": q}

    case = {"model": model, "lik": "gaussian", "K": K, "S": S, "num_data": float(N), "pred": layer(Zp, 1.0, lsp),
            "assign": layer(Za, 0.5, lsa), "lik_var": 0.1 + 0.05 * np.arange(K),
            "assign_lik_var": (0.4 + 0.1 * np.arange(K)) if model != "SMGP" else None}
    z = rng.standard_normal((S, N, K))
    u = rng.uniform(np.finfo(np.float64).tiny, 1.0, (S, N, K))
    return case, X, Y, z, u
