"""Multi-GPU check of the two data-parallel collectives (run under torchrun on >= 2 GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py

Every rank holds a contiguous shard of a config-#4-style workload.  The ELBO and every gradient from
  (a) torch.distributed all_reduce between mgp_elbo_local and mgp_elbo_finish, and
  (b) the communicator attached to the libmgp context (mgp_ctx_set_comm: ncclAllReduce issued by the C library)
must agree with each other bit for bit and with (c) the single-GPU evaluation of all points on rank 0 to 1e-11.
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
    worst = 0.0
    if rank == 0:
        for k in a:
            den = max(np.max(np.abs(c[k])), 1e-300)
            worst = max(worst, float(np.max(np.abs(b[k].reshape(c[k].shape) - c[k])) / den))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "torch_equals_nccl_bitwise_on_every_rank": bool(int(flag)),
                          "worst_rel_err_vs_single_gpu": worst, "elbo": float(b["elbo"]), "ok": bool(int(flag)) and worst <= 1e-11}))
    dist.destroy_process_group()
    return 0 if (bool(int(flag)) and (rank != 0 or worst <= 1e-11)) else 1


if __name__ == "__main__":
    sys.exit(main())
