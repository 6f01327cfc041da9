"""Per-step host wall times of the end-to-end loop (host buffers in, loss out), to find where e2e time goes."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload
from modulatedgps_b200.workloads import model_from_case as build_model

n = 1 << 20
dev = torch.device("cuda", 0)
case, X, Y = make_workload(n, seed=0)
case["num_data"] = float(n)
Xh = torch.as_tensor(X).contiguous().pin_memory()
Yh = torch.as_tensor(Y).contiguous().pin_memory()
model = build_model(case)
model.seed = 3

def step(xd, yd):
    for v in model.trainable_variables:
        v.grad = None
    loss = model._training_loss((xd, yd), n_global=n, point_offset=0)
    loss.backward()
    return loss

from modulatedgps_b200 import _lib
ctx = _lib.get_context(dev)
_orig = ctx.lib.mgp_elbo_fwd_bwd
c_ms = [0.0]
class _Wrap:
    def __call__(self, *a):
        t = time.perf_counter(); r = _orig(*a); c_ms[0] = 1e3 * (time.perf_counter() - t); return r
ctx.lib.mgp_elbo_fwd_bwd = _Wrap()
for _ in range(3):
    step(Xh.to(dev), Yh.to(dev)).item()
rows = []
for i in range(20):
    t0 = time.perf_counter()
    xd = Xh.to(dev, non_blocking=True); yd = Yh.to(dev, non_blocking=True)
    t1 = time.perf_counter()
    loss = step(xd, yd)
    t2 = time.perf_counter()
    _ = loss.item()
    t3 = time.perf_counter()
    rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), c_ms[0], 1e3 * (t3 - t2), 1e3 * (t3 - t0)))
for r in rows:
    print("h2d-issue %.2f  launch %.2f (C call %.2f)  wait %.2f  total %.2f" % r)
