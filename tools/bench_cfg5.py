"""BASELINE config #5 shapes (D = 8, M = 1024, K = 8, S = 32) on ONE B200 at a reduced N: a stress run of the
large-M code path (tile width 16, single-buffer ring, chunked workspace), not the headline bench.
    python tools/bench_cfg5.py [N]     (default N = 262144 = 1/64 of the 16M-point config, 1/8 of one rank's shard)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers_gpu import build_model  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
D, M, K, S = 8, 1024, 8, 32
rng = np.random.default_rng(0)
X = rng.standard_normal((N, D))
comp = rng.integers(0, K, N)
Y = (np.sin(X.sum(1) + comp) + 1.5 * comp + 0.1 * rng.standard_normal(N))[:, None]
r2 = np.random.default_rng(2)
Zp, Za = X[r2.choice(N, M, replace=False)], X[r2.choice(N, M, replace=False)]


def layer(Z, var, ls):
    q = np.stack([np.eye(M) + 0.05 * np.tril(r2.standard_normal((M, M))) for _ in range(K)])
    idx = np.arange(M)
    q[:, idx, idx] = np.abs(q[:, idx, idx]) + 0.05
    return {"variance": np.float64(var), "lengthscales": ls * np.ones(D), "Z": Z, "q_mu": 0.3 * r2.standard_normal((M, K)), "q_sqrt": q}


case = {"model": "SMGP", "lik": "gaussian", "K": K, "S": S, "num_data": float(N), "pred": layer(Zp, 1.0, 2.5),
        "assign": layer(Za, 0.5, 3.0), "lik_var": 0.1 + 0.05 * np.arange(K), "assign_lik_var": None}
dev = torch.device("cuda", 0)
model = build_model(case)
model.seed = 3
Xd, Yd = torch.as_tensor(X).to(dev), torch.as_tensor(Y).to(dev)


def step():
    for v in model.trainable_variables:
        v.grad = None
    loss = model._training_loss((Xd, Yd), n_global=N, point_offset=0)
    loss.backward()
    return loss


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 3
e0.record()
for _ in range(steps):
    loss = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
flops_pt = 6 * (K + 1) * M * M + 12 * M * (D + 2 * K + 1)
print(json.dumps({"workload": f"config #5 shapes at N={N}: D={D}, M={M}, K={K}, S={S}, 1 B200", "ms_per_step": ms,
                  "points_per_s": N / (ms * 1e-3), "alg_tflops": flops_pt * N / (ms * 1e-3) / 1e12,
                  "frac_of_fp64_peak_37": flops_pt * N / (ms * 1e-3) / 37e12, "loss": float(loss.detach())}))
