"""CPU, world_size 2, gloo: the data-parallel protocol of the N > 1 path.

The CUDA kernels cannot run here, so each rank evaluates its shard with the ORACLE in the same "shard sums over
the GLOBAL batch size" form that mgp_elbo_local produces; the host-side pieces under test are the product's own:
shard_bounds, global_count_and_offset, all_reduce_sum_ (modulatedgps_b200/parallel.py) and the rule that the KL
term is added once after the reduction."""
import math
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from modulatedgps_b200 import parallel
        from oracle import svgp_mixture as O
        from tests.helpers import load_golden
        case, g = load_golden("synth4_small.pert")
        X, Y, z, u = g["X"][:n_total], g["Y"][:n_total], g["z"][:, :n_total], g["u"][:, :n_total]
        lo, hi = parallel.shard_bounds(n_total, world, rank)
        n_glob, off = parallel.global_count_and_offset(hi - lo)
        assert (n_glob, off) == (n_total, lo)
        pred, assign = O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"])
        # shard sums: data term over the global N, no KL (num_data = inf)
        val, grads = O.elbo_and_grads("SMGP", "gaussian", pred, assign, O.as_t(case["lik_var"]), None, X[lo:hi], Y[lo:hi],
                                      z[:, lo:hi], u[:, lo:hi], math.inf, n_total=n_glob)
        keys = sorted(grads)
        flat = torch.cat([torch.tensor([val], dtype=torch.float64)] + [torch.as_tensor(grads[k]).reshape(-1) for k in keys])
        parallel.all_reduce_sum_(flat)               # the single collective
        # replicated finish: KL and its gradients once
        kl_p, kl_a = (O.gauss_kl_white(pred), O.gauss_kl_white(assign))
        elbo = float(flat[0]) - float(kl_p + kl_a) / case["num_data"]
        if rank == 0:
            np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
            np.save(os.path.join(out_dir, "elbo.npy"), np.array([elbo]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [192, 101])
def test_two_rank_shard_sums_match_single_process(tmp_path, n_total):
    from oracle import svgp_mixture as O
    from tests.helpers import load_golden, relerr
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True)
    case, g = load_golden("synth4_small.pert")
    X, Y, z, u = g["X"][:n_total], g["Y"][:n_total], g["z"][:, :n_total], g["u"][:, :n_total]
    pred, assign = O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"])
    full, gfull = O.elbo_and_grads("SMGP", "gaussian", pred, assign, O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"])
    assert abs(float(np.load(tmp_path / "elbo.npy")[0]) - full) <= 1e-12 * abs(full)
    # data-term gradients: the all-reduced shard sums equal the single-process gradients minus the KL part
    _, gdata = O.elbo_and_grads("SMGP", "gaussian", pred, assign, O.as_t(case["lik_var"]), None, X, Y, z, u, math.inf)
    flat = np.load(tmp_path / "flat.npy")
    ref = np.concatenate([np.asarray(gdata[k]).reshape(-1) for k in sorted(gdata)])
    assert relerr(flat[1:], ref) <= 1e-11


def test_shard_bounds_cover_the_range():
    from modulatedgps_b200.parallel import shard_bounds
    for n, w in ((1 << 20, 8), (101, 2), (5, 8), (0, 4)):
        edges = [shard_bounds(n, w, r) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in edges]
        assert max(sizes) - min(sizes) <= 1
