"""Multi-GPU check of the two data-parallel collectives (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank holds a contiguous shard of a config-#4-style workload.  The ELBO and every gradient from
  (a) torch.distributed all_reduce between mgp_elbo_local and mgp_elbo_finish, and
  (b) the communicator attached to the libmgp context (mgp_ctx_set_comm: ncclAllReduce issued by the C library)
must agree with (c) the single-GPU evaluation of all points on rank 0 to 1e-11, and with each other — bit for bit on 2
ranks (a two-term sum has one order), to 1e-11 on more: the torch form reduces ONE buffer, libmgp two (per layer), and
NCCL picks its algorithm and chunking per call, so the summation order over > 2 ranks differs in the last bit (8 GPUs:
not bit-identical — 1.4e-12 apart — and both 5e-12 from the single-GPU evaluation; profiles/r02_dist_check_8gpu.json).
Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from modulatedgps_b200 import _lib
    from modulatedgps_b200.parallel import shard_bounds
    from modulatedgps_b200.workloads import config4_workload, model_from_case
    N = 40000 + 7                                   # ragged on purpose
    case, X, Y = config4_workload(N, seed=0, m=64, num_data=N)
    lo, hi = shard_bounds(N, world, rank)
    Xs, Ys = torch.as_tensor(X[lo:hi], device=dev), torch.as_tensor(Y[lo:hi], device=dev)

    def run(collective):
        model = model_from_case(case)
        model.seed, model._step = 5, 0
        if collective is not None:
            model.enable_data_parallel(collective=collective)
            e, g = model.elbo_and_grads(Xs, Ys, n_global=N, point_offset=lo)
        else:
            e, g = model.elbo_and_grads(torch.as_tensor(X, device=dev), torch.as_tensor(Y, device=dev))
        torch.cuda.synchronize()
        _lib.get_context(dev).check_status()
        out = {k: v.detach().cpu().numpy().copy() for k, v in g.items()}
        out["elbo"] = np.array(float(e))
        if collective == "nccl":
            _lib.get_context(dev).set_comm(None)
        return out

    a, b = run("torch"), run("nccl")
    c = run(None) if rank == 0 else None
    same = all(np.array_equal(a[k], b[k]) for k in a)
    ab = max(float(np.max(np.abs(a[k] - b[k])) / max(np.max(np.abs(b[k])), 1e-300)) for k in a)
    worst, worst_torch = 0.0, 0.0
    if rank == 0:
        for k in a:
            den = max(np.max(np.abs(c[k])), 1e-300)
            worst = max(worst, float(np.max(np.abs(b[k].reshape(c[k].shape) - c[k])) / den))
            worst_torch = max(worst_torch, float(np.max(np.abs(a[k].reshape(c[k].shape) - c[k])) / den))
    close = same if world == 2 else ab <= 1e-11
    flag = torch.tensor([1 if same else 0, 1 if close else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    abt = torch.tensor([ab], dtype=torch.float64, device=dev)
    dist.all_reduce(abt, op=dist.ReduceOp.MAX)
    ok = bool(int(flag[1])) and worst <= 1e-11 and worst_torch <= 1e-11
    if rank == 0:
        print(json.dumps({"world": world, "torch_equals_nccl_bitwise_on_every_rank": bool(int(flag[0])),
                          "torch_vs_nccl_worst_rel_diff": float(abt), "worst_rel_err_vs_single_gpu": worst,
                          "worst_rel_err_vs_single_gpu_torch_collective": worst_torch, "elbo": float(b["elbo"]), "ok": ok}))
    dist.destroy_process_group()
    return 0 if (bool(int(flag[1])) and (rank != 0 or ok)) else 1


if __name__ == "__main__":
    sys.exit(main())
