"""GPU tests added in round 2: the gaps the round-1 review listed (VERDICT.md "what's weak" 2-4, 9).

  * the throughput-mode noise generator: Philox4x32-10 known answers (Random123 kat_vectors) and the distribution of the
    z / u draws the fused Monte-Carlo pass consumes (moments, tails, KS, independence of the Box-Muller pair);
  * the likelihood method surface (BroadcastingLikelihood.variational_expectations / predict_mean_and_var,
    GaussianModified._variational_expectations / _predict_mean_and_var / _scalar_log_prob / _predict_log_density,
    MultiClass._variational_expectations / _predict_mean_and_var) against the oracle's formulas;
  * the stale-precompute guard across mgp_elbo_local / mgp_elbo_finish: two models interleaved on one context, and
    parameters changed in place between the two calls;
  * a scalar likelihood variance with K > 1 (the reference's constructor default), Adam's non-finite guard;
  * the RobustMax squash override (the A/B of DESIGN.md §3) and the 60-digit conditional spot check.
"""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch

from tests.helpers import RTOL, load_golden, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hg():
    from tests import helpers_gpu
    return helpers_gpu


def _case(*a, **kw):
    from modulatedgps_b200.workloads import synthetic_case
    return synthetic_case(*a, **kw)


# ---------------------------------------------------------------------------------------------------------------
# Philox
# ---------------------------------------------------------------------------------------------------------------
def test_philox4x32_10_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output."""
    from modulatedgps_b200 import _lib
    kat = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    ctx = _lib.get_context()
    inp = torch.tensor(np.array([c + k for c, k, _ in kat], dtype=np.uint32).view(np.int32), device="cuda")
    out = torch.zeros(len(kat), 4, dtype=torch.int32, device="cuda")
    ctx.check(ctx.lib.mgp_debug_philox(ctx.handle, C.c_void_p(inp.data_ptr()), len(kat), C.c_void_p(out.data_ptr())))
    got = out.cpu().numpy().view(np.uint32)
    assert np.array_equal(got, np.array([o for _, _, o in kat], dtype=np.uint32)), [[hex(v) for v in r] for r in got]


def _draws(seed, N, S, K, stream=0, offset=0):
    from modulatedgps_b200 import _lib
    ctx = _lib.get_context()
    z = torch.empty(S, N, K, dtype=torch.float64, device="cuda")
    u = torch.empty(S, N, K, dtype=torch.float64, device="cuda")
    nz = _lib.MgpNoise(None, None, seed, offset)
    ctx.check(ctx.lib.mgp_debug_noise(ctx.handle, C.byref(nz), N, S, K, stream, _lib.ptr(z), _lib.ptr(u)))
    return z, u


@pytest.mark.parametrize("K", [1, 4, 7])
def test_throughput_mode_noise_has_the_right_distribution(K):
    """>= 1e7 draws per array of what mc_pass consumes in throughput mode (the benchmarked path): z ~ N(0, 1),
    u ~ U(0, 1), all components / samples / points independent.  Bounds are ~5 sigma of the estimator at this sample
    size, so a correct generator fails with probability < 1e-5 and a biased one (wrong Box-Muller scale, 32-bit
    truncation of u, components sharing a counter) fails by orders of magnitude."""
    from scipy import stats
    N, S = 1 << 18, 16 if K < 7 else 8
    z, u = _draws(12345, N, S, K)
    n = z.numel()
    assert n >= 1e7 / 2.5 and bool(torch.isfinite(z).all())
    se = 1.0 / math.sqrt(n)
    zz = z.reshape(-1)
    assert abs(float(zz.mean())) < 5 * se
    assert abs(float(zz.var()) - 1.0) < 5 * math.sqrt(2.0) * se
    assert abs(float((zz ** 3).mean())) < 5 * math.sqrt(15.0) * se                      # skewness
    assert abs(float((zz ** 4).mean()) - 3.0) < 5 * math.sqrt(96.0) * se               # kurtosis
    for t in (2.0, 3.0, 4.0):                                                          # two-sided tail mass
        p = 2 * stats.norm.sf(t)
        assert abs(float((zz.abs() > t).double().mean()) - p) < 5 * math.sqrt(p / n) + 1e-12
    uu = u.reshape(-1)
    assert float(uu.min()) > 0.0 and float(uu.max()) < 1.0                             # log(-log u) stays finite
    assert abs(float(uu.mean()) - 0.5) < 5 * se / math.sqrt(12.0)
    assert abs(float(uu.var()) - 1.0 / 12.0) < 5 * se / math.sqrt(180.0)
    assert float((uu < 1e-6).double().mean()) < 1e-6 + 5 * math.sqrt(1e-6 / n)         # 53-bit mantissa: no mass at 0
    sub = slice(0, 2_000_000)
    assert stats.kstest(zz[sub].cpu().numpy(), "norm").pvalue > 1e-4
    assert stats.kstest(uu[sub].cpu().numpy(), "uniform").pvalue > 1e-4
    # independence: components (k, k + 1 share one Box-Muller radius: cosine / sine branches), z vs u, samples, points
    pairs = []
    if K > 1:
        pairs += [(z[..., 0], z[..., 1]), (z[..., 0] ** 2, z[..., 1] ** 2), (u[..., 0], u[..., 1]), (z[..., K - 1], u[..., 0])]
    pairs += [(z[..., 0], u[..., 0]), (z[0], z[1]), (z[:, :-1], z[:, 1:]), (u[:, :-1], u[:, 1:])]
    for a, b in pairs:
        a, b = a.reshape(-1), b.reshape(-1)
        r = float(((a - a.mean()) * (b - b.mean())).mean() / (a.std() * b.std()))
        assert abs(r) < 5 / math.sqrt(a.numel()), r
    # the stream index and the seed select different, uncorrelated streams; the point offset shifts the same stream
    z1, _ = _draws(12345, 4096, 2, K, stream=1)
    z2, _ = _draws(12346, 4096, 2, K)
    z3, _ = _draws(12345, 4096 - 100, 2, K, offset=100)
    assert not torch.equal(z1, z[:2, :4096]) and not torch.equal(z2, z[:2, :4096])
    assert torch.equal(z3, z[:2, 100:4096])


def test_w_sample_in_philox_mode_is_the_documented_function_of_the_draws(hg):
    """mgp_w_sample without explicit noise == the oracle's relaxed one-hot evaluated on mgp_debug_noise's draws for the
    same (seed, offset): the throughput path uses exactly the stream the distribution test examines."""
    from modulatedgps_b200 import _lib
    from oracle import svgp_mixture as O
    N, K, S = 300, 3, 5
    case, X, Y, _, _ = _case(N, 2, 36, K, S, seed=4)
    model = hg.build_model(case)
    model.seed, model._step = 9, 0
    W = model.W_dist(X).sample(1)[0].reshape(S, N, K)
    z, u = _draws((9 << 20) + 1, N, S, K)
    mu_a, var_a = O.conditional(O.as_t(X), O.layer_from_numpy(case["assign"]))
    Wref = O.relaxed_onehot_weights(mu_a, var_a, z.cpu(), u.cpu())
    assert np.abs(W.cpu().numpy() - Wref.numpy()).max() <= 1e-9


# ---------------------------------------------------------------------------------------------------------------
# likelihood method surface
# ---------------------------------------------------------------------------------------------------------------
def test_gaussian_modified_methods_match_the_reference_formulas():
    import modulatedgps_b200 as mg
    rng = np.random.default_rng(0)
    S, N, K = 5, 37, 4
    Fmu, Fvar = rng.standard_normal((S, N, K)), rng.uniform(0.05, 2.0, (S, N, K))
    Y = rng.standard_normal((N, 1))
    var = rng.uniform(0.2, 1.5, K)
    lik = mg.GaussianModified(variance=1.0, D=K)
    lik.variance.assign(var.reshape(1, K))
    ve_ref = -0.5 * np.log(2 * np.pi) - 0.5 * np.log(var) - 0.5 * ((Y[None] - Fmu) ** 2 + Fvar) / var   # likelihoods.py:39-41
    b = mg.BroadcastingLikelihood(lik)
    ve = b.variational_expectations([], Fmu, Fvar, Y)
    assert tuple(ve.shape) == (S, N, K) and relerr(np.asarray(ve), ve_ref) <= 1e-13
    # the pass-through call the reference makes: Y expanded to [1, N, 1] (broadcasting_lik.py:22-24)
    ve2 = lik._variational_expectations([], Fmu, Fvar, Y[None])
    assert torch.equal(ve2, ve)
    mean, v = b.predict_mean_and_var([], Fmu, Fvar)
    assert np.array_equal(np.asarray(mean), Fmu) and relerr(np.asarray(v), Fvar + var) <= 1e-15   # likelihoods.py:31-32
    lp = lik._scalar_log_prob([], Fmu, Y[None])
    assert relerr(np.asarray(lp), -0.5 * np.log(2 * np.pi) - 0.5 * np.log(var) - 0.5 * (Y[None] - Fmu) ** 2 / var) <= 1e-13
    pld = lik._predict_log_density([], Fmu, Fvar, Y[None])
    s2 = Fvar + var
    assert tuple(pld.shape) == (S, N)
    assert relerr(np.asarray(pld), (-0.5 * np.log(2 * np.pi) - 0.5 * np.log(s2) - 0.5 * (Y[None] - Fmu) ** 2 / s2).sum(-1)) <= 1e-13
    assert np.array_equal(np.asarray(lik._conditional_mean([], Fmu)), Fmu)
    held = lik.variance.value().detach().cpu().numpy().reshape(-1)           # what the softplus round trip of `assign` left
    assert relerr(held, var) <= 1e-13
    assert np.array_equal(np.asarray(lik._conditional_variance([], Fmu)), np.broadcast_to(held, Fmu.shape))
    # the reference's default constructor: ONE variance shared by all components
    lik1 = mg.GaussianModified(variance=0.7)
    ve1 = lik1._variational_expectations([], Fmu, Fvar, Y[None])
    assert relerr(np.asarray(ve1), -0.5 * np.log(2 * np.pi) - 0.5 * np.log(0.7) - 0.5 * ((Y[None] - Fmu) ** 2 + Fvar) / 0.7) <= 1e-13
    with pytest.raises(ValueError):
        b.variational_expectations([], Fmu, Fvar, Y[:-1])


@pytest.mark.parametrize("K", [2, 3, 5])
def test_multiclass_methods_match_the_oracle(K):
    import modulatedgps_b200 as mg
    from oracle import svgp_mixture as O
    rng = np.random.default_rng(K)
    S, N = 3, 41
    Fmu, Fvar = rng.standard_normal((S, N, K)), rng.uniform(1e-12, 1.5, (S, N, K))
    Fvar[0, :3] = 0.0                                           # below safe_sqrt's 1e-10 clip
    Y = rng.integers(0, K, (N, 1)).astype(np.float64)
    lik = mg.MultiClass(K, invlink=mg.RobustMax(K))
    b = mg.BroadcastingLikelihood(lik)
    ve = b.variational_expectations([], Fmu, Fvar, Y)
    assert tuple(ve.shape) == (S, N, 1)                         # broadcasting_lik.py:26-37: reshaped to [S, N, -1]
    Yt = np.tile(Y[None], (S, 1, 1)).reshape(S * N, 1)
    ref = O.multiclass_ve(O.as_t(Yt), O.as_t(Fmu.reshape(S * N, K)), O.as_t(Fvar.reshape(S * N, K))).numpy()
    assert relerr(np.asarray(ve).reshape(-1), ref) <= 1e-12
    mean, var = b.predict_mean_and_var([], Fmu, Fvar)
    rm, rv = O.multiclass_predict_mean_and_var(O.as_t(Fmu.reshape(S * N, K)), O.as_t(Fvar.reshape(S * N, K)))
    assert relerr(np.asarray(mean).reshape(S * N, K), rm.numpy()) <= 1e-12
    assert relerr(np.asarray(var).reshape(S * N, K), rv.numpy()) <= 1e-12


def test_robustmax_squash_override_matches_the_oracle_at_both_values(hg):
    """The A/B knob of DESIGN.md §3: with mgp_set_robustmax_squash(1e-6) the kernels reproduce the oracle evaluated with
    1e-6, with the default they reproduce the 1e-4 goldens; the two ELBOs differ where the test can see it."""
    from modulatedgps_b200 import _lib
    from oracle import svgp_mixture as O
    case, g = load_golden("demo_tf2_2d_modified_multiclass.pert")
    model = hg.build_model(case)
    ctx = _lib.get_context()
    e_default, _ = model.elbo_and_grads(g["X"], g["Y"], noise=(g["z"], g["u"]))
    assert abs(float(e_default) - float(g["out.elbo"])) <= RTOL * abs(float(g["out.elbo"]))
    saved = O.ROBUSTMAX_CDF_SQUASH
    try:
        O.ROBUSTMAX_CDF_SQUASH = 1e-6
        ctx.set_robustmax_squash(1e-6)
        e6, g6 = model.elbo_and_grads(g["X"], g["Y"], noise=(g["z"], g["u"]))
        ref, rg = O.elbo_and_grads(case["model"], case["lik"], O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                                   None, O.as_t(case["assign_lik_var"]), g["X"], g["Y"], g["z"], g["u"], case["num_data"])
    finally:
        O.ROBUSTMAX_CDF_SQUASH = saved
        ctx.set_robustmax_squash(_lib.ROBUSTMAX_CDF_SQUASH)
    assert abs(float(e6) - ref) <= RTOL * abs(ref)
    assert relerr(g6["pred.q_mu"].cpu().numpy(), rg["pred.q_mu"]) <= RTOL
    assert abs(float(e6) - float(e_default)) > 1e-7 * abs(ref)
    with pytest.raises(_lib.MgpError):
        ctx.set_robustmax_squash(0.5)


# ---------------------------------------------------------------------------------------------------------------
# precompute guard across the two-phase C-ABI
# ---------------------------------------------------------------------------------------------------------------
def _two_phase(ctx, model, X, Y, noise, between=None):
    from modulatedgps_b200 import _lib
    from modulatedgps_b200.models import _LayerView
    pv, av = _LayerView(model.pred_layer), _LayerView(model.assign_layer)
    K, S, N = pv.K, int(model.num_samples), X.shape[0]
    cfg = _lib.MgpElboCfg(model._model_kind, model.likelihood.kind, S, 0, float(model.temperature), float(model.num_data), N)
    likv = model.likelihood.component_variances(K)
    z, u = (torch.as_tensor(a, device="cuda").contiguous() for a in noise)
    nz = _lib.MgpNoise(z.data_ptr(), u.data_ptr(), 0, 0)
    rb = torch.empty(int(ctx.lib.mgp_reduce_buffer_len(C.byref(pv.struct), C.byref(av.struct))), dtype=torch.float64, device="cuda")
    ctx.check(ctx.lib.mgp_elbo_local(ctx.handle, C.byref(cfg), C.byref(pv.struct), C.byref(av.struct), _lib.ptr(likv), None,
                                     _lib.ptr(X), _lib.ptr(Y), N, C.byref(nz), _lib.ptr(rb)))

    def finish():
        pg, pgs = pv.grad_buffers()
        ag, ags = av.grad_buffers()
        elbo = torch.empty(1, dtype=torch.float64, device="cuda")
        glik = torch.zeros(K, dtype=torch.float64, device="cuda")
        ctx.check(ctx.lib.mgp_elbo_finish(ctx.handle, C.byref(cfg), C.byref(pv.struct), C.byref(av.struct), _lib.ptr(likv), None,
                                          _lib.ptr(rb), _lib.ptr(elbo), C.byref(pgs), C.byref(ags), _lib.ptr(glik), None))
        grads = {f"pred.{k}": v for k, v in pg.items()}
        grads.update({f"assign.{k}": v for k, v in ag.items()})
        grads["lik_var"] = glik
        return elbo, grads

    return finish, (pv, av, z, u, likv, rb)


def test_two_models_interleaved_on_one_context(hg):
    """local(A), local(B), finish(A), finish(B) on ONE context: finish must not run A's Cholesky backward on B's
    factorisation (round 1 trusted a bare `pre_valid` flag).  Different M on purpose."""
    from modulatedgps_b200 import _lib
    ctx = _lib.get_context()
    ca, Xa, Ya, za, ua = _case(500, 2, 64, 4, 6, seed=31)
    cb, Xb, Yb, zb, ub = _case(400, 2, 36, 4, 6, seed=32)
    ma, mb = hg.build_model(ca), hg.build_model(cb)
    ea, ga = ma.elbo_and_grads(Xa, Ya, noise=(za, ua))
    eb, gb = mb.elbo_and_grads(Xb, Yb, noise=(zb, ub))
    ga, gb = {k: v.clone() for k, v in ga.items()}, {k: v.clone() for k, v in gb.items()}
    dev = lambda a: torch.as_tensor(a, device="cuda").contiguous()
    fin_a, keep_a = _two_phase(ctx, ma, dev(Xa), dev(Ya).reshape(-1), (za, ua))
    fin_b, keep_b = _two_phase(ctx, mb, dev(Xb), dev(Yb).reshape(-1), (zb, ub))
    e1, g1 = fin_a()
    e2, g2 = fin_b()
    e3, g3 = fin_a()                                           # and again, after B's finish replaced the factorisation
    ctx.check_status()
    for (e, g), (er, gr) in (((e1, g1), (ea, ga)), ((e2, g2), (eb, gb)), ((e3, g3), (ea, ga))):
        assert abs(float(e) - float(er)) <= 1e-12 * abs(float(er))
        for k in ("pred.Z", "pred.q_sqrt", "pred.lengthscales", "assign.Z", "assign.q_mu", "assign.variance", "lik_var"):
            assert relerr(g[k].cpu().numpy(), gr[k].cpu().numpy().reshape(g[k].shape)) <= 1e-11, k


def test_parameters_changed_between_local_and_finish_are_reported(hg):
    import modulatedgps_b200 as mg
    from modulatedgps_b200 import _lib
    ctx = _lib.get_context()
    case, X, Y, z, u = _case(300, 2, 36, 3, 4, seed=33)
    model = hg.build_model(case)
    dev = lambda a: torch.as_tensor(a, device="cuda").contiguous()
    finish, keep = _two_phase(ctx, model, dev(X), dev(Y).reshape(-1), (z, u))
    e_ok, _ = finish()
    ctx.check_status()
    assert math.isfinite(float(e_ok))
    finish, keep = _two_phase(ctx, model, dev(X), dev(Y).reshape(-1), (z, u))
    with torch.no_grad():
        keep[0].q_mu.mul_(1.0 + 1e-12)                         # an optimiser step in place, between the two phases
    e_bad, _ = finish()
    assert math.isnan(float(e_bad))
    with pytest.raises(mg.MgpError) as err:
        ctx.check_status()
    assert err.value.code == _lib.MGP_ERR_STALE_PRECOMPUTE
    ctx.check_status()                                         # the flag is cleared once reported


# ---------------------------------------------------------------------------------------------------------------
# training-side fixes
# ---------------------------------------------------------------------------------------------------------------
def test_scalar_likelihood_variance_with_several_components(hg):
    """GaussianModified(variance=v) with D=None (the reference's default) broadcasts ONE variance over the K components:
    its gradient is the SUM of the per-component gradients, on the autograd path and in FusedAdam."""
    import copy
    import modulatedgps_b200 as mg
    from oracle import svgp_mixture as O
    case, X, Y, z, u = _case(200, 2, 25, 3, 4, seed=41)
    case["lik_var"] = 0.3 * np.ones(3)

    def build():
        lik = mg.GaussianModified(variance=0.3)
        assert tuple(lik.variance.shape) == ()
        mk = lambda p: mg.SVGPModified(kernel=mg.SquaredExponential(float(p["variance"]), p["lengthscales"]), likelihood=lik,
                                       inducing_variable=p["Z"], num_latent_gps=3, q_mu=p["q_mu"], q_sqrt=np.tril(p["q_sqrt"]))
        return lik, mg.SMGP(lik, mk(case["pred"]), mk(case["assign"]), K=3, num_samples=4, num_data=case["num_data"])

    lik, model = build()
    ref, rg = O.elbo_and_grads("SMGP", "gaussian", O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                               O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"])
    loss = model._training_loss((X, Y), noise=(z, u))
    loss.backward()
    th = lik.variance.unconstrained_variable
    expect = -rg["lik_var"].sum() * float(torch.sigmoid(th.detach()))
    assert tuple(th.grad.shape) == tuple(th.shape)
    assert abs(float(th.grad) - expect) <= 1e-9 * abs(expect)
    # one fused Adam step moves theta by -lr * sign(g) (first step of Adam), with g the SUMMED gradient
    lik2, model2 = build()
    th2 = lik2.variance.unconstrained_variable
    before = float(th2.detach())
    mg.FusedAdam(model2, 0.01).minimize((X, Y), noise=(z, u))
    step = float(th2.detach()) - before
    assert abs(step + 0.01 * np.sign(expect)) <= 1e-6, (step, expect)


def test_adam_skips_the_update_when_the_elbo_is_not_finite(hg):
    """A failed Cholesky gives NaN ELBO and gradients; the fused update must leave theta, m, v untouched (no host sync
    involved) and run_adam / check_status must raise where the reference's tf.linalg.cholesky would."""
    import modulatedgps_b200 as mg
    from modulatedgps_b200 import _lib
    case, X, Y, z, u = _case(64, 2, 36, 3, 4, seed=4)
    case["pred"]["variance"] = np.float64(1e12)                 # jitter drowns: Kuu + 1e-6 I is numerically singular
    case["pred"]["Z"][1] = case["pred"]["Z"][0]
    model = hg.build_model(case)
    opt = mg.FusedAdam(model, 0.01)
    before = [v.detach().clone() for v in model.trainable_variables]
    loss = opt.minimize((X, Y), noise=(z, u))
    assert not math.isfinite(float(loss))
    for a, b in zip(before, model.trainable_variables):
        assert torch.equal(a, b.detach())
    for st in opt.state.values():
        assert float(st["m"].abs().sum()) == 0.0 and float(st["v"].abs().sum()) == 0.0
    with pytest.raises(mg.NotPositiveDefiniteError):
        _lib.get_context().check_status()


# ---------------------------------------------------------------------------------------------------------------
# 60-digit spot check at the bench's conditioning
# ---------------------------------------------------------------------------------------------------------------
def test_conditional_against_60_digit_arithmetic(hg):
    """tests/golden/hp_conditional.npz (mpmath, 60 digits) for config #4's ASSIGN layer as benchmarked (lengthscale 1.5,
    cond(Kuu) ~ 3e6): who is right when the float64 oracle and the float64 kernel disagree at 1e-9?  Both must sit within
    a small multiple of eps * cond of the 60-digit value, and the kernel (explicit L^-1, DMMA accumulation order) must not
    be more than 10x further from it than the oracle's triangular solves."""
    import modulatedgps_b200 as mg
    from oracle import svgp_mixture as O
    d = np.load(os.path.join(ROOT, "tests", "golden", "hp_conditional.npz"))
    p = {k: d["layer." + k] for k in ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")}
    K = p["q_mu"].shape[1]
    lik = mg.GaussianModified(variance=1.0, D=K)
    layer = mg.SVGPModified(kernel=mg.SquaredExponential(float(p["variance"]), p["lengthscales"]), likelihood=lik,
                            inducing_variable=p["Z"], num_latent_gps=K, q_mu=p["q_mu"], q_sqrt=np.tril(p["q_sqrt"]))
    fm, fv = layer.predict_f(d["X"])
    om, ov = O.conditional(O.as_t(d["X"]), O.layer_from_numpy(p))
    cond = float(np.linalg.cond(O.kuu(O.layer_from_numpy(p)).numpy()))
    floor = np.finfo(np.float64).eps * cond
    errs = {"kernel.fmean": relerr(np.asarray(fm), d["fmean"]), "kernel.fvar": relerr(np.asarray(fv), d["fvar"]),
            "oracle.fmean": relerr(om.numpy(), d["fmean"]), "oracle.fvar": relerr(ov.numpy(), d["fvar"])}
    print("60-digit spot check, cond(Kuu) = %.3g, eps*cond = %.3g: %s" % (cond, floor, errs))
    for k, e in errs.items():
        assert e <= floor, (k, e, floor)
    assert errs["kernel.fmean"] <= 10 * max(errs["oracle.fmean"], 1e-15)
    assert errs["kernel.fvar"] <= 10 * max(errs["oracle.fvar"], 1e-15)


# ---------------------------------------------------------------------------------------------------------------
# training-trajectory parity: kernels + fused Adam vs oracle + autograd + TF's Adam written out
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,steps", [("demo_tf2_full.init", 30), ("demo_tf2_2d_modified_multiclass.init", 20)])
def test_training_trajectory_matches_the_oracle(name, steps, hg):
    """The demos' training loop (utils/training_utils.py:6-10) step for step on EXPLICIT noise and a fixed minibatch order:
    libmgp's forward+backward + mgp_adam_step against the CPU oracle differentiated by torch autograd through GPflow's
    bijectors (softplus, fill-triangular) and updated by TF 2.10's Adam rule.  Losses and parameters must agree at every
    step — an end-to-end check of everything between `run_adam` and the kernels that a one-shot gradient comparison cannot
    give (optimizer state, bijector chain rule, parameter write-back, noise indexing across steps)."""
    import modulatedgps_b200 as mg
    from oracle import svgp_mixture as O
    case, g = load_golden(name)
    X, Y = g["X"], g["Y"]
    N, K, S = X.shape[0], case["K"], 5
    case = dict(case, S=S)
    B = min(500, N)
    rng = np.random.default_rng(7)
    tiny = np.finfo(np.float64).tiny
    batches = []
    for _ in range(steps):
        idx = rng.permutation(N)[:B]
        batches.append((X[idx], Y[idx], rng.standard_normal((S, B, K)), rng.uniform(tiny, 1.0, (S, B, K))))
    # ---- device
    model = hg.build_model(case)
    opt = mg.FusedAdam(model, 0.005)
    dev_losses = [float(opt.minimize((xb, yb), noise=(z, u))) for xb, yb, z, u in batches]
    # ---- oracle: unconstrained leaves, GPflow's bijectors, TF Adam
    leaves = {}

    def mk(key, v, kind):
        v = torch.as_tensor(np.asarray(v), dtype=torch.float64)
        if kind == "sp":
            u = O.softplus_inverse(v)
        elif kind == "tri":
            M = v.shape[-1]
            idx = O.fill_triangular_index(M)
            ii, jj = np.tril_indices(M)
            u = torch.zeros(v.shape[0], M * (M + 1) // 2, dtype=torch.float64)
            u[:, idx[ii, jj]] = v[:, ii, jj]
        else:
            u = v.clone()
        leaves[key] = (u.detach().clone().requires_grad_(True), kind)

    for l in ("pred", "assign"):
        mk(l + ".variance", case[l]["variance"], "sp")
        mk(l + ".lengthscales", case[l]["lengthscales"], "sp")
        mk(l + ".Z", case[l]["Z"], "id")
        mk(l + ".q_mu", case[l]["q_mu"], "id")
        mk(l + ".q_sqrt", case[l]["q_sqrt"], "tri")
    for key in ("lik_var", "assign_lik_var"):
        if case[key] is not None:
            mk(key, case[key], "sp")

    def cons(key):
        u, kind = leaves[key]
        if kind == "sp":
            return torch.nn.functional.softplus(u)
        if kind == "tri":
            n = u.shape[-1]
            M = int(round((math.sqrt(8 * n + 1) - 1) / 2))
            idx = O.fill_triangular_index(M)
            ii, jj = np.tril_indices(M)
            out = torch.zeros(u.shape[0], M, M, dtype=torch.float64)
            out[:, ii, jj] = u[:, idx[ii, jj]]
            return out
        return u

    params = [u for u, _ in leaves.values()]
    m1 = [torch.zeros_like(p) for p in params]
    v2 = [torch.zeros_like(p) for p in params]
    lr, b1, b2, eps = 0.005, 0.9, 0.999, 1e-7
    ref_losses = []
    for step, (xb, yb, z, u) in enumerate(batches, 1):
        for p in params:
            p.grad = None
        lay = lambda l: {k: cons(f"{l}.{k}") for k in ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")}
        loss = -O.elbo(case["model"], case["lik"], lay("pred"), lay("assign"), cons("lik_var") if "lik_var" in leaves else None,
                       cons("assign_lik_var") if "assign_lik_var" in leaves else None, xb, yb, z, u, case["num_data"])
        loss.backward()
        ref_losses.append(float(loss))
        lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        with torch.no_grad():
            for p, mm, vv in zip(params, m1, v2):
                if p.grad is None:
                    continue
                mm += (p.grad - mm) * (1 - b1)
                vv += (p.grad * p.grad - vv) * (1 - b2)
                p -= lr_t * mm / (vv.sqrt() + eps)
    assert np.max(np.abs(np.array(dev_losses) - np.array(ref_losses)) / np.abs(ref_losses)) <= 1e-7, (dev_losses[-3:], ref_losses[-3:])
    assert dev_losses[-1] < dev_losses[0]
    dev_params = {"pred.Z": model.pred_layer.inducing_variable.Z, "assign.q_mu": model.assign_layer.q_mu,
                  "pred.q_sqrt": model.pred_layer.q_sqrt, "assign.variance": model.assign_layer.kernel.variance,
                  "pred.lengthscales": model.pred_layer.kernel.lengthscales}
    for key, p in dev_params.items():
        ref = cons(key).detach().numpy()
        assert relerr(np.asarray(p.value().detach().cpu()).reshape(ref.shape), ref) <= 1e-6, key


# ---------------------------------------------------------------------------------------------------------------
# the fused forward kernel (kept selectable: measured slower than cond_fwd_a + cond_fwd_b)
# ---------------------------------------------------------------------------------------------------------------
def test_fused_forward_kernel_matches_the_goldens(hg, monkeypatch):
    """MGP_FUSED_FWD=1 routes 32-point-tile layers through cond_fwd_fused (generator warps + L^-1 product + K passes in
    one persistent kernel); results must be the two-kernel path's."""
    from modulatedgps_b200 import _lib
    from oracle import svgp_mixture as O
    ctx = _lib.get_context()
    monkeypatch.setenv("MGP_FUSED_FWD", "1")
    for name in ("demo_tf2.pert", "synth4_small.pert", "shape_d5_k1.pert", "demo_john_doe_fullbatch.pert"):
        case, g = load_golden(name)
        model = hg.build_model(case)
        elbo, grads = model.elbo_and_grads(g["X"], g["Y"], noise=(g["z"], g["u"]))
        assert abs(float(elbo) - float(g["out.elbo"])) <= RTOL * abs(float(g["out.elbo"])), name
        for k in ("pred.Z", "pred.q_sqrt", "assign.q_mu", "assign.lengthscales"):
            r = g["out.grad." + k]
            assert relerr(grads[k].cpu().numpy().reshape(r.shape), r) <= RTOL, (name, k)
        fm, fv = model.pred_layer.predict_f(g["Xtest"])
        assert relerr(np.asarray(fm), g["out.predict_f.pred.mean"]) <= RTOL and relerr(np.asarray(fv), g["out.predict_f.pred.var"]) <= RTOL
    case, X, Y, z, u = _case(3000, 2, 256, 4, 16, seed=3000)
    model = hg.build_model(case)
    e1, g1 = model.elbo_and_grads(X, Y, noise=(z, u))
    g1 = {k: v.clone() for k, v in g1.items()}
    monkeypatch.delenv("MGP_FUSED_FWD")
    e0, g0 = model.elbo_and_grads(X, Y, noise=(z, u))
    ctx.check_status()
    assert abs(float(e1) - float(e0)) <= 1e-12 * abs(float(e0))
    for k in g0:
        assert relerr(g1[k].cpu().numpy(), g0[k].cpu().numpy()) <= 1e-10, k


# ---------------------------------------------------------------------------------------------------------------
# the alternative forms of cond_fwd_a / cond_bwd_b (selectable for A/B timing) compute the same numbers
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("env", [{"MGP_NO_KUF_STASH": "1"}, {"MGP_FWD_A_PIPE": "1", "MGP_BWD_B_RING": "1"},
                                 {"MGP_FWD_A_PIPE": "1", "MGP_NO_KUF_STASH": "1"}, {"MGP_FWD_B_16W": "1"}])
def test_alternative_kernel_forms_agree(hg, monkeypatch, env):
    """Default: barrier-phased cond_fwd_a that keeps its Kuf tiles, two-CTA cond_bwd_b that reads them back.
    MGP_FWD_A_PIPE / MGP_BWD_B_RING select the one-CTA software-pipelined / ring forms (measured slower),
    MGP_NO_KUF_STASH makes cond_bwd_b generate Kuf again.  The Kuf values are bit-identical either way; only the order
    of the E-sum accumulation differs."""
    from modulatedgps_b200 import _lib
    ctx = _lib.get_context()
    cases = [_case(3000, 2, 256, 4, 16, seed=3000), _case(700, 5, 96, 3, 8, seed=7), _case(1500, 3, 400, 2, 4, seed=11),
             _case(900, 2, 324, 2, 4, seed=13)]
    for case, X, Y, z, u in cases:
        model = hg.build_model(case)
        e0, g0 = model.elbo_and_grads(X, Y, noise=(z, u))
        g0 = {k: v.clone() for k, v in g0.items()}
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e1, g1 = model.elbo_and_grads(X, Y, noise=(z, u))
        g1 = {k: v.clone() for k, v in g1.items()}
        for k in env:
            monkeypatch.delenv(k)
        ctx.check_status()
        assert abs(float(e1) - float(e0)) <= 1e-12 * abs(float(e0))
        for k in g0:
            assert relerr(g1[k].cpu().numpy(), g0[k].cpu().numpy()) <= 1e-10, k


def test_host_batch_stream_delivers_every_batch_in_order():
    """HostBatchStream: host batches come out as device tensors, in order, with equal values; the copy of the next batch
    is enqueued before the current one is handed out (two buffer pairs, reused two batches later)."""
    import modulatedgps_b200 as mg
    rng = np.random.default_rng(0)
    batches = [(rng.standard_normal((50 + 7 * i, 3)), rng.standard_normal((50 + 7 * i, 1))) for i in range(5)]
    pinned = [(torch.as_tensor(x).pin_memory(), torch.as_tensor(y).pin_memory()) for x, y in batches[:2]] + batches[2:]
    got = []
    for Xd, Yd in mg.HostBatchStream(iter(pinned)):
        assert Xd.is_cuda and Yd.is_cuda and Xd.dtype == torch.float64
        got.append((Xd.cpu().numpy().copy(), Yd.cpu().numpy().copy()))
    assert len(got) == len(batches)
    for (x, y), (gx, gy) in zip(batches, got):
        assert np.array_equal(x, gx) and np.array_equal(y, gy)


def test_config5_parity_on_the_16k_point_subsample():
    """config #5 (D = 8, M = 1024, K = 8, S = 32) at its own parameters, the first 16384 points of its
    data set, explicit noise: ELBO and every gradient against the CPU oracle.  cond(Kuu) is 1.8e7 / 1.3e8 here (Z drawn
    from X ~ N(0, I_8) at lengthscales 2.5 / 3: not the near-diagonal Kuu the survey expected), so the bound is the
    conditioning noise floor max(1e-9, 100 eps cond); measured: ELBO 4e-14, worst gradient 9.0e-10
    (profiles/r02_parity_cfg5_16k.json, tools/parity_cfg5_16k.py)."""
    from modulatedgps_b200 import _lib, workloads as W
    from oracle import svgp_mixture as O
    n = 16384
    case = W.config5_parameters(num_data=n)
    X, Y = W.config5_points(0, n)
    K, S = case["K"], case["S"]
    rng = np.random.default_rng(3)
    z = rng.standard_normal((S, n, K))
    u = rng.uniform(np.finfo(np.float64).tiny, 1.0, (S, n, K))
    model = W.model_from_case(case)
    elbo, grads = model.elbo_and_grads(X, Y, noise=(z, u))
    _lib.get_context().check_status()
    ref, rg = O.elbo_and_grads(case["model"], case["lik"], O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                               O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"])
    eps = np.finfo(np.float64).eps
    tol = {name: max(RTOL, 100 * eps * float(np.linalg.cond(O.kuu(O.layer_from_numpy(case[name])).numpy())))
           for name in ("pred", "assign")}
    assert abs(float(elbo) - ref) <= RTOL * abs(ref)
    for k, r in rg.items():
        mine = grads[k].cpu().numpy().reshape(r.shape)
        t = tol[k.split(".")[0]] if "." in k else max(tol.values())
        assert relerr(mine, r) <= t, (k, relerr(mine, r), t)


def test_config4_parity_on_a_64k_point_subsample():
    """config #4 AS BENCHMARKED (D = 2, M = 256, K = 4, S = 16; jittered-grid Z, assign lengthscale 1.5:
    cond(Kuu) = 3e6), 65536 points with explicit noise, ELBO and every gradient against the CPU oracle; the bound is
    max(1e-9, 100 eps cond) per layer."""
    from modulatedgps_b200 import _lib, workloads as W
    from oracle import svgp_mixture as O
    n = 1 << 16
    case, X, Y = W.config4_workload(n, seed=0, num_data=n)
    K, S = case["K"], case["S"]
    rng = np.random.default_rng(3)
    z = rng.standard_normal((S, n, K))
    u = rng.uniform(np.finfo(np.float64).tiny, 1.0, (S, n, K))
    model = W.model_from_case(case)
    elbo, grads = model.elbo_and_grads(X, Y, noise=(z, u))
    _lib.get_context().check_status()
    ref, rg = O.elbo_and_grads(case["model"], case["lik"], O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                               O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"])
    eps = np.finfo(np.float64).eps
    tol = {name: max(RTOL, 100 * eps * float(np.linalg.cond(O.kuu(O.layer_from_numpy(case[name])).numpy())))
           for name in ("pred", "assign")}
    assert abs(float(elbo) - ref) <= RTOL * abs(ref)
    worst = {}
    for k, r in rg.items():
        mine = grads[k].cpu().numpy().reshape(r.shape)
        t = tol[k.split(".")[0]] if "." in k else max(tol.values())
        worst[k] = relerr(mine, r)
        assert worst[k] <= t, (k, worst[k], t)
    print("config #4, 64K points: worst gradient rel. err.", max(worst.values()), worst)


def test_survey_invariants_hold_on_the_cuda_path(hg):
    """SURVEY.md §4's invariants, on the kernels (tests/test_oracle_golden.py checks them on the oracle):
    predict_y variance = predict_f variance + sigma^2_k (likelihoods.py:32); predict_assign rows sum to 1 (models.py:88);
    the ELBO's data term is a per-point MEAN — duplicating the minibatch (and its noise) leaves it unchanged — while the KL
    term scales with 1 / num_data only (models.py:76,79); the data term does not depend on S when all S samples are equal."""
    case, g = load_golden("demo_john_doe.pert")
    model = hg.build_model(case)
    Xt = g["Xtest"]
    fm, fv = model.pred_layer.predict_f(Xt)
    my, vy = model.predict_y(Xt, S=1)
    lik_var = np.asarray(case["lik_var"]).reshape(1, -1)
    assert relerr(np.asarray(vy[0]) - np.asarray(fv), np.broadcast_to(lik_var, np.asarray(fv).shape)) <= 1e-12
    assert relerr(np.asarray(my[0]), np.asarray(fm)) <= 1e-14
    assert np.allclose(np.asarray(model.predict_assign(Xt)).sum(1), 1.0, atol=1e-14)
    X, Y, z, u = g["X"], g["Y"], g["z"], g["u"]
    kl = float(model.pred_layer.prior_kl()) + float(model.assign_layer.prior_kl())
    e1, _ = model.elbo_and_grads(X, Y, noise=(z, u))
    X2, Y2 = np.concatenate([X, X]), np.concatenate([Y, Y])
    z2, u2 = np.concatenate([z, z], 1), np.concatenate([u, u], 1)
    e2, _ = model.elbo_and_grads(X2, Y2, noise=(z2, u2))
    assert abs(float(e1) - float(e2)) <= 1e-12 * abs(float(e1))          # mean over points; KL / num_data unchanged
    data_term = float(e1) + kl / case["num_data"]
    case10 = dict(case, num_data=10.0 * case["num_data"])
    e10, _ = hg.build_model(case10).elbo_and_grads(X, Y, noise=(z, u))
    assert abs(float(e10) - (data_term - kl / case10["num_data"])) <= 1e-12 * abs(float(e10))
    zs, us = np.repeat(z[:1], 5, 0), np.repeat(u[:1], 5, 0)               # five identical samples == one sample
    case5, case1s = dict(case, S=5), dict(case, S=1)
    e5, _ = hg.build_model(case5).elbo_and_grads(X, Y, noise=(zs, us))
    e1s, _ = hg.build_model(case1s).elbo_and_grads(X, Y, noise=(z[:1], u[:1]))
    assert abs(float(e5) - float(e1s)) <= 1e-12 * abs(float(e1s))
