"""Stage-level diagnostics on the GPU: Kuu / Cholesky / L^-1 against torch, predict_f at M=256."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C  # noqa: E402

import modulatedgps_b200 as mg  # noqa: E402
from modulatedgps_b200 import _lib  # noqa: E402
from modulatedgps_b200.models import _LayerView  # noqa: E402
from oracle import svgp_mixture as O  # noqa: E402


def stage_check(M, D, K, seed=0, ls=1.0):
    rng = np.random.default_rng(seed)
    side = int(round(M ** (1.0 / D))) if D <= 2 else 0
    if D == 2 and side * side == M:
        gx = np.linspace(0.5, side - 0.5, side)
        Z = np.stack(np.meshgrid(gx, gx, indexing="ij"), -1).reshape(-1, 2) + rng.uniform(-0.2, 0.2, (M, 2))
    else:
        Z = rng.standard_normal((M, D)) * 2.0
    q_mu = 0.3 * rng.standard_normal((M, K))
    q_sqrt = np.stack([np.eye(M) + 0.05 * np.tril(rng.standard_normal((M, M))) for _ in range(K)])
    kern = mg.SquaredExponential(variance=0.8, lengthscales=ls * np.ones(D))
    layer = mg.SVGPModified(kernel=kern, likelihood=None, inducing_variable=Z, num_latent_gps=K, q_mu=q_mu, q_sqrt=q_sqrt)
    view = _LayerView(layer)
    ctx = _lib.get_context()
    Kuu = torch.empty(M, M, dtype=torch.float64, device="cuda")
    L = torch.empty_like(Kuu)
    Linv = torch.empty_like(Kuu)
    ctx.check(ctx.lib.mgp_debug_kuu_chol(ctx.handle, C.byref(view.struct), _lib.ptr(Kuu), _lib.ptr(L), _lib.ptr(Linv)))
    ctx.check_status()
    ol = {"variance": torch.tensor(0.8, dtype=torch.float64), "lengthscales": torch.as_tensor(ls * np.ones(D)),
          "Z": torch.as_tensor(Z), "q_mu": torch.as_tensor(q_mu), "q_sqrt": torch.as_tensor(q_sqrt)}
    Kref = O.kuu(ol)
    Lref = torch.linalg.cholesky(Kref)
    Liref = torch.linalg.solve_triangular(Lref, torch.eye(M, dtype=torch.float64), upper=False)
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print(f"M={M} D={D} K={K} cond={np.linalg.cond(Kref.numpy()):.2e}  Kuu {rel(Kuu.cpu(), Kref):.2e}  L {rel(L.cpu(), Lref):.2e}  "
          f"Linv {rel(Linv.cpu(), Liref):.2e}  |L Linv - I| {float((L.cpu() @ Linv.cpu() - torch.eye(M, dtype=torch.float64)).abs().max()):.2e}")
    X = rng.uniform(Z.min(), Z.max(), (1000, D))
    fm, fv = layer.predict_f(X)
    rm, rv = O.conditional(torch.as_tensor(X), ol)
    print(f"    predict_f N=1000: mean {rel(fm.cpu(), rm):.2e} var {rel(fv.cpu(), rv):.2e}")


if __name__ == "__main__":
    for (M, D, K) in ((25, 1, 3), (64, 2, 4), (256, 2, 4), (100, 8, 8), (1024, 8, 8)):
        try:
            stage_check(M, D, K, ls=1.0 if D <= 2 else 3.0)
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
