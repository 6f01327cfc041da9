"""Per-warp phase breakdown of cond_fwd_b / cond_bwd_a / cond_bwd_b from a TEMPORARY build with -DMGP_PHASE_TIMERS
(clock64 sums per warp, modulatedgps_b200/csrc/stream_kernels.cu):
    NVCC_EXTRA=-DMGP_PHASE_TIMERS python -m modulatedgps_b200.build --force ; gpurun -- python tools/phase_timers.py
Prints, per kernel, the share of each phase in the consumer warps' time (mean over CTAs and warps of the LAST launch)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from modulatedgps_b200 import _lib  # noqa: E402
from modulatedgps_b200.workloads import config4_workload, model_from_case  # noqa: E402

NAMES = {0: ("cond_fwd_b", ["wait full", "multiply", "store B_k + colsq", "reduce8 + partial store", "fmean + arrive", "loop head"]),
         1: ("cond_bwd_a", ["wait full", "sc load / zero", "multiply", "ck*sc FMA", "epilogue (last stage)", "hand-off + loop head"]),
         2: ("cond_bwd_b", ["wait full", "multiply", "Kuf load + DMUL", "E-sum DMMA + RED", "tfree wait + next load (warps 0,1)", "loop head"])}


def main():
    case, X, Y = config4_workload(1 << 20, seed=0, num_data=1 << 20)
    model = model_from_case(case)
    Xd, Yd = torch.as_tensor(X).cuda(), torch.as_tensor(Y).cuda()
    os.environ["MGP_SERIAL_LAYERS"] = "1"          # read at context creation: one layer's kernels at a time
    ctx = _lib.get_context(Xd.device)
    for _ in range(3):
        loss = model._training_loss((Xd, Yd), n_global=1 << 20, point_offset=0)
        loss.backward()
    torch.cuda.synchronize()
    out = np.zeros((4, 160, 17, 8), dtype=np.int64)
    rc = ctx.lib.mgp_debug_phase_dump(out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    for kid, (name, phases) in NAMES.items():
        a = out[kid, :148, :8, :].astype(np.float64)      # consumer warps 0..7
        tot = a.sum(-1)
        print(f"## {name}: mean clocks per consumer warp {tot.mean():.0f} (min {tot.min():.0f}, max {tot.max():.0f})")
        for i, ph in enumerate(phases):
            print(f"   {ph:38s} {100 * a[..., i].sum() / tot.sum():5.1f} %   per warp mean {a[..., i].mean():10.0f} clk   "
                  f"warp 0 {a[:, 0, i].mean():9.0f}  warp 7 {a[:, 7, i].mean():9.0f}")


if __name__ == "__main__":
    main()
