"""Config #5's 16K-point parity subsample ("parity on a 16K-point subsample"):
the first 16384 points of the config-#5 data set at the config's OWN parameters and shapes (D = 8, M = 1024, K = 8,
S = 32), explicit noise, CUDA path against the CPU oracle (oracle/svgp_mixture.py, autograd) on ELBO and every gradient.
Too heavy for the routine suite (the oracle holds K x M x N doubles with its autograd graph: a few GB, ~1-2 minutes on
the box's host cores); run once per round on the GPU box:

    python tools/parity_cfg5_16k.py > profiles/rNN_parity_cfg5_16k.json

Tolerance per tensor: max(1e-9, 100 eps cond(Kuu)), reported beside the error."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(n=16384):
    from modulatedgps_b200 import _lib, workloads as W
    from oracle import svgp_mixture as O
    case = W.config5_parameters(num_data=n)
    X, Y = W.config5_points(0, n)
    K, S = case["K"], case["S"]
    rng = np.random.default_rng(3)
    z = rng.standard_normal((S, n, K))
    u = rng.uniform(np.finfo(np.float64).tiny, 1.0, (S, n, K))
    model = W.model_from_case(case)
    t0 = time.perf_counter()
    elbo, grads = model.elbo_and_grads(X, Y, noise=(z, u))
    torch.cuda.synchronize()
    _lib.get_context().check_status()
    t_gpu = time.perf_counter() - t0
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    ref, rg = O.elbo_and_grads(case["model"], case["lik"], O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                               O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"])
    t_cpu = time.perf_counter() - t0
    eps = np.finfo(np.float64).eps
    cond = {name: float(np.linalg.cond(O.kuu(O.layer_from_numpy(case[name])).numpy())) for name in ("pred", "assign")}
    tol = {name: max(1e-9, 100 * eps * c) for name, c in cond.items()}
    out = {"points": n, "D": X.shape[1], "M": int(case["pred"]["Z"].shape[0]), "K": K, "S": S, "cond_kuu": cond,
           "elbo": float(elbo), "elbo_oracle": float(ref), "elbo_rel_err": abs(float(elbo) - float(ref)) / abs(float(ref)),
           "grad_rel_err": {}, "tolerance": {}, "first_gpu_call_s": t_gpu, "oracle_s": t_cpu}
    ok = out["elbo_rel_err"] <= 1e-9
    for k, r in rg.items():
        mine = grads[k].cpu().numpy().reshape(r.shape)
        e = float(np.max(np.abs(mine - r)) / max(np.max(np.abs(r)), np.finfo(np.float64).tiny))
        t = tol[k.split(".")[0]] if "." in k else max(tol.values())
        out["grad_rel_err"][k], out["tolerance"][k] = e, t
        ok = ok and e <= t
    out["worst_grad_rel_err"] = max(out["grad_rel_err"].values())
    out["ok"] = bool(ok)
    print(json.dumps(out))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(int(sys.argv[1]) if len(sys.argv) > 1 else 16384))
