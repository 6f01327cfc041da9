"""BASELINE.json configs #1-#3 replayed as FULL training runs (examples/replay_demos.py) against the ELBO curves the
reference publishes in final_figs/  — the one piece of evidence about the reference that does not pass
through this repo's own restatement of GPflow.  Anchors (read off the figures):

    demo_tf2                          -2.85 +- 0.1 at the first log (iteration 5);  >= -0.25 at iteration 2000
    demo_tf2_2d_modified_multiclass   -4.3 at the first log;  >= +0.8 at iteration 2000 (figure: ~ +1.05)
    demo_john_doe                     ~ -6 at the start;  >= +1.5 (median of the last 100 logs) at iteration 10000

The multiclass demo is replayed with the RobustMax CDF squash at gpflow's 1e-4 (shipped) and at 1e-6 (round 1's value):
both trajectories are written next to each other (gpurun_out/r02_replay_*.json -> profiles/) for DESIGN.md §3.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(demo, squash=None):
    import sys
    sys.path.insert(0, ROOT)
    from examples import replay_demos as R
    rec = R.replay(demo, squash=squash)
    rec["summary"] = R.summarise(rec)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        tag = demo if squash is None else f"{demo}_squash{squash:g}"
        with open(os.path.join(out, f"r02_replay_{tag}.json"), "w") as f:
            json.dump(rec, f)
    return rec, R.ANCHORS[demo]


def _check(rec, anchors):
    e = np.asarray(rec["elbos"])
    assert rec["iters"][0] == 5 and rec["iters"][-1] == rec["num_iter"] and np.isfinite(e).all()
    first, tol = anchors["first"]
    assert abs(e[0] - first) <= tol, (e[0], first)
    assert np.median(e[-100:]) >= anchors["final_at_least"], (np.median(e[-100:]), anchors)


def test_demo_tf2_reaches_the_published_elbo():
    rec, anchors = _run("tf2")
    _check(rec, anchors)
    assert min(rec["assign_argmax_counts"]) > 0          # all three components end up used (final_figs/demo_tf2.png)


def test_demo_multiclass_reaches_the_published_elbo_with_either_squash():
    a, anchors = _run("multiclass")
    b, _ = _run("multiclass", squash=1e-6)
    _check(a, anchors)
    _check(b, anchors)
    # same noise stream, same data order: the two runs differ only through the squash, and only far below what the
    # published figure resolves (the decision between 1e-4 and 1e-6 cannot come from the figure: DESIGN.md §3)
    ea, eb = np.asarray(a["elbos"]), np.asarray(b["elbos"])
    assert np.max(np.abs(ea - eb)[:40]) < 0.05
    # the curve's shape: the plateau near -1.2 around iterations 300-450, zero crossing before iteration 1300
    at = dict(zip(a["iters"], a["elbos"]))
    assert -1.6 <= at[400] <= -0.8 and max(a["elbos"][: 1300 // 5]) > 0.0


def test_demo_john_doe_reaches_the_published_elbo():
    rec, anchors = _run("john_doe")
    _check(rec, anchors)
