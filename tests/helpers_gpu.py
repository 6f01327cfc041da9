"""Build modulatedgps_b200 models from golden/oracle case dicts (GPU tests only)."""


from modulatedgps_b200.workloads import model_from_case as build_model  # noqa: F401,E402


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
