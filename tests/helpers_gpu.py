"""Build modulatedgps_b200 models from golden/oracle case dicts (GPU tests only)."""
import numpy as np

import modulatedgps_b200 as mg


def build_model(case):
    K = int(case["K"])

    def layer(p, lik):
        ls = np.asarray(p["lengthscales"], dtype=np.float64)
        kern = mg.SquaredExponential(variance=float(p["variance"]), lengthscales=float(ls) if ls.ndim == 0 else ls)
        return mg.SVGPModified(kernel=kern, likelihood=lik, inducing_variable=np.asarray(p["Z"]), num_latent_gps=K,
                               whiten=True, q_mu=np.asarray(p["q_mu"]), q_sqrt=np.tril(np.asarray(p["q_sqrt"])))

    def gauss(v):
        g = mg.GaussianModified(variance=1.0, D=K)
        g.variance.assign(np.asarray(v, dtype=np.float64).reshape(1, K))
        return g

    if case["model"] == "SMGP":
        lik = gauss(case["lik_var"])
        return mg.SMGP(likelihood=lik, pred_layer=layer(case["pred"], lik), assign_layer=layer(case["assign"], lik), K=K,
                       num_samples=int(case["S"]), num_data=case["num_data"])
    lik = mg.MultiClass(K, invlink=mg.RobustMax(K)) if case["lik"] == "multiclass" else gauss(case["lik_var"])
    alik = gauss(case["assign_lik_var"])
    return mg.SMGPModified(likelihood=lik, assign_likelihood=alik, pred_layer=layer(case["pred"], lik),
                           assign_layer=layer(case["assign"], alik), K=K, num_samples=int(case["S"]),
                           num_data=case["num_data"])


def unconstrained_grad_dict(model):
    """{golden key: gradient on the unconstrained variable} after loss.backward()."""
    out = {}
    for lname, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
        out[f"{lname}.variance"] = layer.kernel.variance.unconstrained_variable.grad
        out[f"{lname}.lengthscales"] = layer.kernel.lengthscales.unconstrained_variable.grad
        out[f"{lname}.Z"] = layer.inducing_variable.Z.unconstrained_variable.grad
        out[f"{lname}.q_mu"] = layer.q_mu.unconstrained_variable.grad
        out[f"{lname}.q_sqrt"] = layer.q_sqrt.unconstrained_variable.grad
    if hasattr(model.likelihood.likelihood, "variance"):
        out["lik_var"] = model.likelihood.likelihood.variance.unconstrained_variable.grad
    if hasattr(model, "assign_likelihood"):
        out["assign_lik_var"] = model.assign_likelihood.likelihood.variance.unconstrained_variable.grad
    return out
