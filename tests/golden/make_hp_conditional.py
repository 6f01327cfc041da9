"""Independent high-precision spot check of one SVGP conditional: mpmath at 60 significant
digits, at the BENCH configuration's conditioning (config #4's assign layer: M = 256 jittered-grid inducing points,
lengthscale 1.5 -> cond(Kuu) ~ 3e6, where float64 oracle and float64 kernel may legitimately disagree at 1e-9).

    python tests/golden/make_hp_conditional.py        (authoring container; ~2 minutes of pure-Python mpmath)

Writes tests/golden/hp_conditional.npz: the layer's parameters, 8 test points and fmean / fvar [8, K] rounded to float64
from the 60-digit values.  Nothing here shares code with oracle/ or the kernels: the SE kernel is evaluated from
(x - z)^2 directly (no |x|^2 + |z|^2 - 2 x.z cancellation), the Cholesky and the substitutions are plain loops.
Formulas: gpflow SquaredExponential.K, Kuu + 1e-6 I, base_conditional(white=True, full_cov=False),
call sites MixtureGPs/models.py:133-143.
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from modulatedgps_b200.workloads import config4_workload
    mp.mp.dps = 60
    case, X, _ = config4_workload(4096, seed=0)
    layer = case["assign"]
    Xs = X[:8]
    Z, q_mu, q_sqrt = layer["Z"], layer["q_mu"], np.tril(layer["q_sqrt"])
    M, D = Z.shape
    K = q_mu.shape[1]
    var = mp.mpf(float(layer["variance"]))
    ls = [mp.mpf(float(v)) for v in np.asarray(layer["lengthscales"]).reshape(-1)]
    if len(ls) == 1:
        ls = ls * D
    Zm = [[mp.mpf(float(Z[i, d])) for d in range(D)] for i in range(M)]

    def k(a, b):
        r2 = sum(((a[d] - b[d]) / ls[d]) ** 2 for d in range(D))
        return var * mp.exp(-r2 / 2)

    Kuu = [[k(Zm[i], Zm[j]) for j in range(i + 1)] for i in range(M)]
    for i in range(M):
        Kuu[i][i] += mp.mpf("1e-6")
    L = [[mp.mpf(0)] * (i + 1) for i in range(M)]          # lower Cholesky, plain loops
    for j in range(M):
        s = Kuu[j][j] - sum(L[j][p] ** 2 for p in range(j))
        L[j][j] = mp.sqrt(s)
        for i in range(j + 1, M):
            L[i][j] = (Kuu[i][j] - sum(L[i][p] * L[j][p] for p in range(j))) / L[j][j]
        if j % 32 == 0:
            print("cholesky column", j, flush=True)
    Lq = [[[mp.mpf(float(q_sqrt[c, i, j])) for j in range(M)] for i in range(M)] for c in range(K)]
    qm = [[mp.mpf(float(q_mu[i, c])) for c in range(K)] for i in range(M)]
    fmean = np.zeros((len(Xs), K))
    fvar = np.zeros((len(Xs), K))
    for n, x in enumerate(Xs):
        xm = [mp.mpf(float(v)) for v in x]
        kn = [k(Zm[i], xm) for i in range(M)]
        a = [mp.mpf(0)] * M                                  # a = L^-1 k_n
        for i in range(M):
            a[i] = (kn[i] - sum(L[i][p] * a[p] for p in range(i))) / L[i][i]
        asq = sum(v * v for v in a)
        for c in range(K):
            fmean[n, c] = float(sum(a[i] * qm[i][c] for i in range(M)))
            b = [sum(Lq[c][i][j] * a[i] for i in range(j, M)) for j in range(M)]     # b = Lq_c^T a (Lq lower)
            fvar[n, c] = float(var - asq + sum(v * v for v in b))
        print("point", n, flush=True)
    np.savez(os.path.join(HERE, "hp_conditional.npz"), X=Xs, fmean=fmean, fvar=fvar, digits=60,
             **{f"layer.{k_}": np.asarray(v, dtype=np.float64) for k_, v in layer.items()})
    print("wrote hp_conditional.npz")


if __name__ == "__main__":
    main()
