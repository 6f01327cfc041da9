"""Timeline of one ELBO step (torch.profiler / CUPTI): kernel list in launch order with the idle gaps between them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload  # noqa: E402
from modulatedgps_b200.workloads import model_from_case as build_model  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    dev = torch.device("cuda", 0)
    case, X, Y = make_workload(n, seed=0)
    case["num_data"] = float(n)
    Xd, Yd = torch.as_tensor(X).to(dev), torch.as_tensor(Y).to(dev)
    model = build_model(case)
    model.seed = 3

    def step():
        for v in model.trainable_variables:
            v.grad = None
        loss = model._training_loss((Xd, Yd), n_global=n, point_offset=0)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(nsteps):
            step()
        torch.cuda.synchronize()
    cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and (e.time_range.end - e.time_range.start) > 300]
    cpu.sort(key=lambda e: e.time_range.start)
    for e in cpu[:60]:
        print(f"   cpu {e.time_range.start / 1e3:12.3f} ms  dur {(e.time_range.end - e.time_range.start) / 1e3:8.3f} ms  {e.name[:70]}")
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    prev_end = t0
    total_gap = 0.0
    for e in evs:
        gap = e.time_range.start - prev_end
        if gap > 0:
            total_gap += gap
        dur = e.time_range.end - e.time_range.start
        if dur > 2000 or gap > 100 or os.environ.get('GAP_ALL'):
            print(f"{(e.time_range.start - t0) / 1e3:9.3f} ms  dur {dur / 1e3:8.3f} ms  gap {gap / 1e3:7.3f} ms  {e.name[:70]}")
        prev_end = max(prev_end, e.time_range.end)
    print(f"span {(prev_end - t0) / 1e3:.3f} ms, idle gaps {total_gap / 1e3:.3f} ms, kernels {len(evs)}")


if __name__ == "__main__":
    main()
