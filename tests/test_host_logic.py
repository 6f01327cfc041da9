"""CPU: host-side logic of the package — the C-ABI library loads and exports every symbol include/mgp.h
declares (no compute without a GPU), the product path fails loudly without CUDA, the bijector / fill-triangular
maps match the TFP conventions, and the bench workload generator is deterministic."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from tests.helpers import relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from modulatedgps_b200 import build
    return build.build()


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "mgp.h")).read()
    declared = set(re.findall(r"\b(mgp_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"mgp_layer", "mgp_layer_grad", "mgp_noise", "mgp_elbo_cfg", "mgp_ctx"}
    assert len(declared) >= 18, declared
    lib = ctypes.CDLL(built_lib)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/mgp.h but not exported by libmgp.so"
    from modulatedgps_b200 import _lib
    assert set(_lib.exported_symbols()) == declared, set(_lib.exported_symbols()) ^ declared
    _lib.load_library()


def test_ctypes_struct_layout_matches_header():
    from modulatedgps_b200 import _lib
    assert ctypes.sizeof(_lib.MgpLayer) == 4 * 4 + 5 * 8
    assert ctypes.sizeof(_lib.MgpLayerGrad) == 5 * 8
    assert ctypes.sizeof(_lib.MgpNoise) == 4 * 8
    assert ctypes.sizeof(_lib.MgpElboCfg) == 4 * 4 + 2 * 8 + 8
    assert _lib.MgpElboCfg.temperature.offset == 16 and _lib.MgpElboCfg.n_global.offset == 32


def test_header_is_plain_c_and_ctypes_mirrors_it_field_by_field(tmp_path):
    """include/mgp.h compiles as C (no C++ / torch types at the boundary) and every field of every ctypes mirror sits at
    the offset the C compiler gives it."""
    import shutil
    import subprocess
    from modulatedgps_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    mirrors = {"mgp_layer": _lib.MgpLayer, "mgp_layer_grad": _lib.MgpLayerGrad, "mgp_noise": _lib.MgpNoise,
               "mgp_elbo_cfg": _lib.MgpElboCfg, "mgp_adam_slot": _lib.MgpAdamSlot}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "mgp.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(built_lib):
    import modulatedgps_b200 as mg
    from modulatedgps_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.get_context()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mg.SquaredExponential(1.0, 1.0)          # parameters live on the device
    h = ctypes.c_void_p()
    lib = _lib.load_library()
    assert lib.mgp_ctx_create(0, None, ctypes.byref(h)) != 0 and not h.value


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "modulatedgps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_fill_triangular_matches_tfp_convention():
    from modulatedgps_b200.parameter import fill_triangular_index
    idx = fill_triangular_index(3)
    x = np.array([1, 2, 3, 4, 5, 6])
    mat = np.where(idx >= 0, x[np.maximum(idx, 0)], 0)
    assert mat.tolist() == [[4, 0, 0], [6, 5, 0], [3, 2, 1]]        # TFP docstring example
    from oracle.svgp_mixture import fill_triangular_index as oracle_idx
    for m in (1, 2, 5, 25):
        assert np.array_equal(fill_triangular_index(m), oracle_idx(m))


def test_bijectors_roundtrip_and_chain_rule_on_cpu_tensors():
    from modulatedgps_b200.parameter import FillTriangular, Softplus
    sp = Softplus()
    y = torch.tensor([1e-3, 0.1, 0.5, 3.0, 40.0], dtype=torch.float64)
    x = sp.inverse(y)
    assert torch.allclose(sp.forward(x), y, rtol=1e-14, atol=0)
    xr = x.clone().requires_grad_(True)
    sp.forward(xr).sum().backward()
    assert torch.allclose(sp.grad_to_unconstrained(torch.ones_like(x), x), xr.grad, rtol=1e-14)
    ft = FillTriangular()
    v = torch.arange(1.0, 16.0, dtype=torch.float64).reshape(1, 15).repeat(2, 1)
    Lm = ft.forward(v)
    assert Lm.shape == (2, 5, 5) and torch.equal(torch.triu(Lm, 1), torch.zeros_like(Lm))
    assert torch.equal(ft.inverse(Lm), v)
    g = torch.randn(2, 5, 5, dtype=torch.float64)
    vr = v.clone().requires_grad_(True)
    (ft.forward(vr) * g).sum().backward()
    assert torch.equal(ft.grad_to_unconstrained(g, v), vr.grad)


def test_constrained_shapes_need_no_kernel():
    """Parameter.shape (used once per parameter and step by the gradient plumbing) comes from the bijector's
    forward_shape, not from evaluating the bijector."""
    from modulatedgps_b200.parameter import FillTriangular, Identity, Softplus
    assert Softplus().forward_shape((1, 4)) == (1, 4) and Identity().forward_shape((7, 3)) == (7, 3)
    ft = FillTriangular()
    for m in (1, 2, 5, 25, 256):
        n = m * (m + 1) // 2
        assert ft.forward_shape((4, n)) == (4, m, m)
        assert tuple(ft.forward(torch.zeros(4, n, dtype=torch.float64)).shape) == (4, m, m)


def test_exp_tab_is_accurate_to_an_ulp(tmp_path):
    """The table exponential of the Kuf kernels (csrc/stream_kernels.cu::exp_tab, csrc/exp_tab.h), restated in C with
    the same constants and operation order, stays within 1.1 ulp of expl() over [-700, 700] (2 M samples here;
    tools/exp_tab_check.c runs 20 M by default)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "exp_tab_check")
    subprocess.run(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "exp_tab_check.c"), "-lm"], check=True)
    r = subprocess.run([exe, "2000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # the header is what tools/gen_exp_tab.py generates
    gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_exp_tab.py")], capture_output=True, text=True, check=True)
    assert gen.stdout == open(os.path.join(ROOT, "modulatedgps_b200", "csrc", "exp_tab.h")).read()


def test_gauss_hermite_header_matches_numpy():
    src = open(os.path.join(ROOT, "modulatedgps_b200", "csrc", "gh20.h")).read()
    nums = [float(t) for t in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", src.split("GH20_X[20]")[1])]
    x, w = np.polynomial.hermite.hermgauss(20)
    assert np.array_equal(np.array(nums[:20]), x) and np.array_equal(np.array(nums[20:40]), w)
    # the statically initialised __constant__ copies the kernels read (no first-use upload, no per-device flag)
    assert np.array_equal(np.array(nums[40:60]), x) and np.array_equal(np.array(nums[60:80]), w / np.sqrt(np.pi))
    assert "cudaMemcpyToSymbol(" not in open(os.path.join(ROOT, "modulatedgps_b200", "csrc", "mc_pass.cu")).read()


def test_bench_workload_is_deterministic_and_well_conditioned():
    import bench
    cfg = bench.Cfg(4, points=2048)
    c1, X1, Y1 = bench.make_workload(cfg, 0, 2048)
    c2, X2, Y2 = bench.make_workload(cfg, 0, 2048)
    assert np.array_equal(X1, X2) and np.array_equal(Y1, Y2) and np.array_equal(c1["pred"]["Z"], c2["pred"]["Z"])
    assert X1.shape == (2048, 2) and c1["pred"]["Z"].shape == (256, 2) and c1["pred"]["q_sqrt"].shape == (4, 256, 256)
    assert bench.Cfg(4).flops_per_point() == 1_999_872
    assert bench.Cfg(5).flops_per_point() == 56_930_304
    assert bench.Cfg(4).N == 1 << 20 and bench.Cfg(5).N == 1 << 24
    alg, executed = bench.Cfg(4).kernel_flops_per_point()
    assert alg["cond_bwd_a"] == 2 * 4 * 256 * 256 and abs(executed["syrk"] / alg["syrk"] - 528 / 514) < 1e-12
    assert abs(executed["cond_fwd_a"] / alg["cond_fwd_a"] - 17 / 16) < 1e-12
    from oracle import svgp_mixture as O
    for lname in ("pred", "assign"):
        cond = np.linalg.cond(O.kuu(O.layer_from_numpy(c1[lname])).numpy())
        assert cond < 1e7, (lname, cond)


def test_config5_shards_are_independent_of_the_sharding():
    """Rank r of R generates rows [r N / R, (r + 1) N / R) of the SAME data set whatever R is (bench.py --config 5)."""
    from modulatedgps_b200 import workloads as W
    Xa, Ya = W.config5_points(0, 3 * (1 << 16) // 2)
    Xb, Yb = W.config5_points((1 << 16) - 5, (1 << 16) + 7)
    assert np.array_equal(Xa[(1 << 16) - 5:(1 << 16) + 7], Xb) and np.array_equal(Ya[(1 << 16) - 5:(1 << 16) + 7], Yb)
    assert Xa.shape == (3 * (1 << 16) // 2, 8) and Ya.shape == (3 * (1 << 16) // 2, 1)
    case = W.config5_parameters(m=64, k=8)
    assert case["pred"]["Z"].shape == (64, 8) and np.unique(case["pred"]["Z"], axis=0).shape[0] == 64


def test_robustmax_squash_is_one_named_constant():
    """include/mgp.h, the ctypes layer, the oracle and the shim carry the same RobustMax CDF squash (gpflow's literal 1e-4),
    and the kernels take it from the context, not from a literal of their own."""
    from modulatedgps_b200 import _lib
    from oracle import svgp_mixture as O
    hdr = open(os.path.join(ROOT, "include", "mgp.h")).read()
    val = float(re.search(r"#define MGP_ROBUSTMAX_CDF_SQUASH\s+(\S+)", hdr).group(1))
    assert val == _lib.ROBUSTMAX_CDF_SQUASH == O.ROBUSTMAX_CDF_SQUASH == 1e-4
    shim = open(os.path.join(ROOT, "oracle", "shim", "gpflow", "likelihoods.py")).read()
    assert float(re.search(r"^ROBUSTMAX_CDF_SQUASH = (\S+)", shim, re.M).group(1)) == val
    src = open(os.path.join(ROOT, "modulatedgps_b200", "csrc", "mc_pass.cu")).read()
    assert "2e-6" not in src and "1e-6;" not in src.split("robustmax_prob")[1].split("sample_weights")[0]


def test_second_build_does_not_run_nvcc():
    """build() must hit its cache when nothing changed (round 1 hashed its own stamp file and recompiled every time)."""
    from modulatedgps_b200 import build as B
    B.build()                                  # whatever state the tree is in: brings the cache up to date
    before = B.nvcc_invocations
    lib = B.build()
    assert B.nvcc_invocations == before and os.path.exists(lib)
    assert not os.path.exists(os.path.join(ROOT, "modulatedgps_b200", "csrc", ".build_stamp"))


def test_literal_oracle_equals_the_deduplicated_one():
    """The S-tiled formulation the reference's graph executes (bench.py's `literal` CPU timing) is the same function."""
    from modulatedgps_b200.workloads import synthetic_case
    from oracle import svgp_mixture as O
    for model in ("SMGP", "SMGPModified"):
        case, X, Y, z, u = synthetic_case(50, 2, 16, 3, 4, seed=2, model=model)
        args = (case["model"], "gaussian", O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"]),
                O.as_t(case["lik_var"]), None if case["assign_lik_var"] is None else O.as_t(case["assign_lik_var"]),
                X, Y, z, u, case["num_data"])
        e0, g0 = O.elbo_and_grads(*args)
        e1, g1 = O.elbo_and_grads(*args, literal=True)
        assert abs(e0 - e1) <= 1e-13 * abs(e0)
        for k in g0:
            assert relerr(g1[k], g0[k]) <= 1e-11, k


def test_oracle_conditional_against_60_digit_arithmetic():
    """tests/golden/hp_conditional.npz (mpmath, 60 digits; tests/golden/make_hp_conditional.py) at the bench's own
    conditioning, cond(Kuu) ~ 3e6: the float64 oracle's forward error must be a small multiple of eps * cond."""
    from oracle import svgp_mixture as O
    d = np.load(os.path.join(ROOT, "tests", "golden", "hp_conditional.npz"))
    layer = {k: d["layer." + k] for k in ("variance", "lengthscales", "Z", "q_mu", "q_sqrt")}
    fm, fv = O.conditional(O.as_t(d["X"]), O.layer_from_numpy(layer))
    cond = float(np.linalg.cond(O.kuu(O.layer_from_numpy(layer)).numpy()))
    assert 1e6 < cond < 1e7
    floor = np.finfo(np.float64).eps * cond
    assert relerr(fm.numpy(), d["fmean"]) <= floor and relerr(fv.numpy(), d["fvar"]) <= floor


def test_reference_arm_prints_the_contract_line_and_other_ranks_stay_silent():
    """`bench.py --impl reference` (the CPU arm the driver launches beside the GPU arm): ONE JSON line on rank 0 with the
    same metric / unit / config as the native arm, `impl`, `cpu_baseline` and a zero-copy `e2e`; ranks > 0 exit 0 without
    output.  Run on a reduced point count so that it takes seconds."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"]
    env = dict(os.environ, RANK="0")
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "elbo_fwd_bwd_points_per_s" and d["unit"] == "points/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["config"]["N"] == 1 << 20 and d["config"]["M"] == 256 and d["config"]["K"] == 4 and d["config"]["S"] == 16
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0 and abs(d["value"] - 8192 / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    silent = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RANK="1"), timeout=600)
    assert silent.returncode == 0 and silent.stdout.strip() == ""
