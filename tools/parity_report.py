"""Print a parity table (CUDA path vs the golden vectors of the reference run) for every golden case.
Diagnostic companion of tests/test_gpu_parity.py: never asserts, catches per-case failures, so one GPU call
shows everything.    python tools/parity_report.py [name-substring]"""
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests.helpers import golden_names, grad_keys, load_golden, relerr  # noqa: E402
from modulatedgps_b200.workloads import model_from_case as build_model
from tests.helpers_gpu import unconstrained_grad_dict  # noqa: E402


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    worst = 0.0
    for name in golden_names():
        if pat not in name:
            continue
        case, g = load_golden(name)
        print(f"== {name}  model={case['model']} lik={case['lik']} N={g['X'].shape[0]} M={case['pred']['Z'].shape[0]} "
              f"K={case['K']} S={case['S']} cond=({float(g['cond.pred']):.1e},{float(g['cond.assign']):.1e})")
        try:
            model = build_model(case)
            for lname, layer in (("pred", model.pred_layer), ("assign", model.assign_layer)):
                fm, fv = layer.predict_f(g["Xtest"])
                e1, e2 = relerr(np.asarray(fm), g[f"out.predict_f.{lname}.mean"]), relerr(np.asarray(fv), g[f"out.predict_f.{lname}.var"])
                worst = max(worst, e1, e2)
                print(f"   predict_f.{lname}: mean {e1:.2e}  var {e2:.2e}")
            elbo, grads = model.elbo_and_grads(g["X"], g["Y"], noise=(g["z"], g["u"]))
            ee = abs(float(elbo) - float(g["out.elbo"])) / abs(float(g["out.elbo"]))
            worst = max(worst, ee)
            print(f"   elbo {float(elbo):+.15e} ref {float(g['out.elbo']):+.15e} rel {ee:.2e}")
            for k in grad_keys(g):
                ref = g["out.grad." + k]
                mine = grads[k].cpu().numpy().reshape(ref.shape)
                if np.max(np.abs(ref)) < 1e-12:
                    print(f"   grad {k:22s} ref~0  max|mine| {np.max(np.abs(mine)):.2e}")
                else:
                    e = relerr(mine, ref)
                    worst = max(worst, e)
                    print(f"   grad {k:22s} rel {e:.2e}   |ref|max {np.max(np.abs(ref)):.3e}")
            loss = model._training_loss((g["X"], g["Y"]), noise=(g["z"], g["u"]))
            loss.backward()
            gu = unconstrained_grad_dict(model)
            for k in grad_keys(g):
                ref = -g["out.gradu." + k]          # golden holds d ELBO; the loss is -ELBO
                mine = gu[k].cpu().numpy().reshape(ref.shape)
                if np.max(np.abs(ref)) >= 1e-12:
                    e = relerr(mine, ref)
                    worst = max(worst, e)
                    if e > 1e-9:
                        print(f"   gradu {k:21s} rel {e:.2e}")
            my, vy = model.predict_y(g["Xtest"], S=2)
            pa, am = model.predict_assign_with_argmax(g["Xtest"])
            e1, e2 = relerr(np.asarray(my[0]), g["out.predict_y.mean"]), relerr(np.asarray(vy[0]), g["out.predict_y.var"])
            e3 = relerr(np.asarray(pa), g["out.predict_assign.probs"])
            same = bool(np.array_equal(np.asarray(am), g["out.predict_assign.argmax"]))
            worst = max(worst, e1, e2, e3)
            print(f"   predict_y mean {e1:.2e} var {e2:.2e}; predict_assign {e3:.2e} argmax_exact={same}")
            if "out.predict_samples.y" in g:
                S2 = g["sample.z_assign"].shape[0]
                sy, sf = model.predict_samples(g["Xtest"], S=S2, noise=(g["sample.z_assign"], g["sample.u"], g["sample.z_pred"]))
                e1, e2 = relerr(np.asarray(sy), g["out.predict_samples.y"]), relerr(np.asarray(sf), g["out.predict_samples.f"])
                worst = max(worst, e1, e2)
                print(f"   predict_samples y {e1:.2e} f {e2:.2e}")
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
            worst = float("inf")
    print(f"WORST relative error over all printed tensors: {worst:.3e}")


if __name__ == "__main__":
    main()
