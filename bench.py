#!/usr/bin/env python
"""Headline benchmark: ELBO forward+backward points/s of the SVGP mixture at BASELINE.json config #4
(N = 2^20 points in total, D = 2, M = 256 inducing points per layer, K = 4 components, S = 16 MC samples),
float64, synthetic data, on 1/2/4/8 B200 of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] ...                # the reference arm: CPU restatement on host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One "step" = everything needed to hand an optimiser -ELBO and all its gradients: Kuu build +
Cholesky of both layers, the streamed conditional forward, the fused MC pass, the full backward, the single
all-reduce (N > 1) and the replicated Cholesky/kernel backward.  Rank 0 prints ONE JSON line.

  value   whole-job points/s with X, Y resident in HBM            (device-timed, max over ranks)
  e2e     the same metric through the public Python API with HOST (pinned) X, Y: H2D copy of the step's inputs
          and a D2H read of the loss inside the timed region
  roofline  the dominant kernel against the measured FP64 DMMA peak (profiles/r01_fp64_peak_microbench.txt);
            MEASURED_PEAKS.json holds HBM and bf16 peaks only, so the FP64 denominator is this repo's own
            measurement on the same pool, stated in the object
  cpu_baseline  oracle/svgp_mixture.py (the CPU float64 restatement of the TF2/GPflow path; TF itself is not
            installable in this image) timed on the box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP64_PEAK_TFLOPS = 37.0      # measured DMMA.8x8x4 issue-rate peak on this pool's B200 (profiles/r01_fp64_peak_microbench.txt)
DGEMM_TFLOPS = 35.4          # cuBLAS DGEMM 8192^3 on the same box (profiles/r01_dgemm_peak.json)


class Cfg:
    """The benchmarked configuration: BASELINE.json configs[3] (#4, the one `metric` is quoted on; default) or
    configs[4] (#5, the Kuf-streaming / FP64 DMMA stress shape) — sizes from modulatedgps_b200.workloads."""

    def __init__(self, which, points=None):
        from modulatedgps_b200.workloads import CONFIG4, CONFIG5
        c = {4: CONFIG4, 5: CONFIG5}[which]
        self.which = which
        self.N, self.D, self.M, self.K, self.S = (points or c["N"]), c["D"], c["M"], c["K"], c["S"]
        self.N_full = c["N"]

    def flops_per_point(self):
        """algorithmic flops per point, two layers, forward + backward."""
        return 6 * (self.K + 1) * self.M * self.M + 12 * self.M * (self.D + 2 * self.K + 1)

    def executed_algorithm_flops_per_point(self):
        """The same count without the M^2 / point / layer of the L-bar reduction, which the S_k reformulation
         removes from the algorithm altogether: (4 + 6K) M^2 + the O(M) terms."""
        return (4 + 6 * self.K) * self.M * self.M + 12 * self.M * (self.D + 2 * self.K + 1)

    def kernel_flops_per_point(self):
        """(algorithmic, executed) flops per point, both layers, of each streaming kernel (SURVEY.md §8d / Appendix B,
        per layer: cond_fwd_a M^2, cond_fwd_b K M^2, syrk K M^2, cond_bwd_a K M^2, cond_bwd_b M^2).  Executed =
        algorithmic x the padding of the triangular blocking: 16-row blocks ((nb+1)/nb), with the two all-zero fragments
        of every diagonal 16 x 16 block skipped ((2nb+1)/2nb) in cond_fwd_b / cond_bwd_a / cond_bwd_b; SYRK computes
        64 x 64 tile pairs (36 of 64 fragments on diagonal pairs) for M(M+1)/2 / 64 algorithmic 8 x 8 fragments."""
        M, K = self.M, self.K
        Mp = (M + 31) // 32 * 32
        nb = Mp // 16
        nb64 = (Mp + 63) // 64
        syrk_pad = (64 * nb64 * (nb64 - 1) // 2 + 36 * nb64) / (M * (M + 1) / 2 / 64)
        unit = {"cond_fwd_a": 1, "cond_fwd_b": K, "syrk": K, "cond_bwd_a": K, "cond_bwd_b": 1}
        unit["cond_fwd"] = 1 + K          # cond_fwd_a + cond_fwd_b in one kernel (32-point tiles)
        alg = {k: 2 * u * M * M for k, u in unit.items()}
        pad = {"cond_fwd_a": (nb + 1) / nb * (Mp / M) ** 2, "cond_fwd_b": (2 * nb + 1) / (2 * nb) * (Mp / M) ** 2,
               "syrk": syrk_pad, "cond_bwd_a": (2 * nb + 1) / (2 * nb) * (Mp / M) ** 2,
               "cond_bwd_b": (2 * nb + 1) / (2 * nb) * (Mp / M) ** 2,
               "cond_fwd": (2 * nb + 1) / (2 * nb) * (Mp / M) ** 2}
        return alg, {k: alg[k] * pad[k] for k in alg}


def make_workload(cfg, lo, hi):
    """(case, X[lo:hi], Y[lo:hi]) of the configuration's synthetic data set."""
    from modulatedgps_b200 import workloads as W
    if cfg.which == 4:
        case, X, Y = W.config4_workload(cfg.N, seed=0, num_data=cfg.N)
        return case, X[lo:hi], Y[lo:hi]
    case = W.config5_parameters(num_data=cfg.N)
    X, Y = W.config5_points(lo, hi)
    return case, X, Y


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement on the host cores (reported baseline, and `--impl reference`)
# ----------------------------------------------------------------------------------------------------
def cpu_points_per_s(cfg, chunk=4096, reps=3, warmup=1, literal=False):
    """ELBO fwd+bwd of oracle/svgp_mixture.py on `chunk` of the configuration's points with every host core.
    literal=True: the S-tiled formulation the reference's TF graph executes (S-fold the conditional work)."""
    import torch
    from oracle import svgp_mixture as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case, X, Y = make_workload(cfg, 0, chunk)
    rng = np.random.default_rng(3)
    z = rng.standard_normal((cfg.S, chunk, cfg.K))
    u = rng.uniform(np.finfo(np.float64).tiny, 1.0, (cfg.S, chunk, cfg.K))
    pred, assign = O.layer_from_numpy(case["pred"]), O.layer_from_numpy(case["assign"])
    times = []
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        O.elbo_and_grads("SMGP", "gaussian", pred, assign, O.as_t(case["lik_var"]), None, X, Y, z, u, case["num_data"],
                         n_total=cfg.N, literal=literal)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return chunk / statistics.median(times), cores, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = Cfg(args.config, args.points)
    chunk = 8192 if cfg.which == 4 else 2048
    warm = max(0, args.warmup)
    value, cores, times = cpu_points_per_s(cfg, chunk=chunk, reps=max(1, args.steps), warmup=warm)
    sample = (f"{chunk} of the {cfg.N} config-#{cfg.which} points per step, ELBO fwd+bwd (torch autograd) of "
              f"oracle/svgp_mixture.py (de-duplicated over S), explicit noise, {cores} torch threads")
    line = {"impl": "reference", "metric": "elbo_fwd_bwd_points_per_s", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
            "ms_per_step": 1e3 * statistics.median(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, args.gpus),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the TF2/GPflow path; TensorFlow/GPflow are not installable in this image"},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def workload_config(cfg, n_gpus):
    return {"workload": f"BASELINE config #{cfg.which}: synthetic N={cfg.N} (total), D={cfg.D}, M={cfg.M}, K={cfg.K}, "
                        f"S={cfg.S}, SMGP + GaussianModified, ELBO fwd+bwd step, Philox noise on device",
            "N": cfg.N, "D": cfg.D, "M": cfg.M, "K": cfg.K, "S": cfg.S, "points_per_gpu": cfg.N // n_gpus,
            "sharding": f"dp{n_gpus}: contiguous row shards, parameters replicated, the flat reduce buffer all-reduced once "
                        "per step (per layer, behind the C-ABI)",
            "l2": "no explicit flush: each step streams the materialised A (2 layers x M x N/gpus x 8 B = "
                  f"{2 * cfg.M * (cfg.N // n_gpus) * 8 / 1e9:.1f} GB) through HBM, far beyond the 126 MB L2"}


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, reasons, power = [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        if sm:
            busy = [c for c, p in zip(sm, power) if p > 0.5 * max(power)] or sm
            out["sm_mhz"] = statistics.median(busy)
            out["power_w_max"] = max(power)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[4, 5],
                    help="BASELINE.json configs[3] (#4: N=2^20, D=2, M=256, K=4, S=16; the headline metric, default) or "
                         "configs[4] (#5: N=2^24, D=8, M=1024, K=8, S=32; meant for --gpus 8)")
    ap.add_argument("--points", type=int, default=None, help="total points (default: the configuration's own N)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="nccl", choices=["nccl", "torch"],
                    help="N > 1: 'nccl' = ncclAllReduce issued by libmgp on a communicator attached to its context "
                         "(mgp_ctx_set_comm); 'torch' = torch.distributed.all_reduce between the two C calls")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    # NCCL prints its version banner on STDOUT (C stdio) when the communicator is created; rank 0's stdout must stay
    # ONE JSON line, so file descriptor 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: modulatedgps_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    cfg = Cfg(args.config, args.points)
    n_local = cfg.N // world
    n_total = cfg.N = n_local * world

    import modulatedgps_b200 as mg  # noqa: F401
    from modulatedgps_b200 import _lib
    from modulatedgps_b200.workloads import model_from_case as build_model

    case, X, Y = make_workload(cfg, rank * n_local, (rank + 1) * n_local)
    Xh = torch.as_tensor(X).contiguous().pin_memory()
    Yh = torch.as_tensor(Y).contiguous().pin_memory()
    del X, Y
    Xd, Yd = Xh.to(dev), Yh.to(dev)
    model = build_model(case)
    model.seed = 3
    if world > 1:
        model.enable_data_parallel(collective=args.collective)
    ctx = _lib.get_context(dev)
    kw = dict(n_global=n_total, point_offset=rank * n_local)

    def step(xd, yd):
        for v in model.trainable_variables:
            v.grad = None
        loss = model._training_loss((xd, yd), **kw)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing -------------------------------------------------------------------
    # the clock sampler (an nvidia-smi -lms subprocess) is started BEFORE the warm-up so that its start-up (NVML
    # initialisation can stall the GPU for milliseconds) stays out of the timed region; samples under load are kept
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step(Xd, Yd)
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(Xd, Yd)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    launches = ctx.launch_count - launches0
    final_loss = float(loss.detach())
    # ---- the same K steps again with libmgp's per-stage CUDA-event timers on (roofline of the dominant kernel);
    #      kept out of the pass above so that the event records cannot perturb `value`
    ctx.timing_enable(True)
    ctx.timing_read(reset=True)
    for _ in range(args.steps):
        step(Xd, Yd)
    barrier()
    stages = ctx.timing_read(reset=True)
    ctx.timing_enable(False)

    # ---- end-to-end timing: host buffers in, loss out ------------------------------------------------
    # untimed: first-use allocations of the per-step device copies.  Same statement pattern as the timed loop below:
    # the previous step's xd / yd are still alive when the next pair is allocated, so the caching allocator needs
    # two pairs, and the second pair's cudaMalloc (60-80 ms with the 23 GB workspace resident, measured) must land here
    # The host batches go through the package's input stage (HostBatchStream): every step's X, Y are copied from pinned
    # host memory inside the timed region, on a copy stream, one step ahead of the step that consumes them.
    import itertools
    feed = mg.HostBatchStream(itertools.repeat((Xh, Yh)), dev)
    for _ in range(3):
        xd, yd = next(feed)
        loss = step(xd, yd)
        _ = loss.item()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e2.record()
    trace = []
    # Every step's loss is read back inside the timed region, ONE STEP LATE (as a training loop that logs its loss does):
    # the host enqueues step i + 1 while step i runs, so the device does not idle through Python between steps.
    # (a plain .item() one step late would not do: its D2H copy is enqueued BEHIND the step just launched and returns
    #  only when that step ends; the loss goes to pinned host memory right behind its own step, guarded by an event)
    loss_host = torch.zeros(2, dtype=torch.float64).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    losses = []
    for i in range(args.steps):
        ta = time.perf_counter()
        xd, yd = next(feed)
        tb = time.perf_counter()
        loss = step(xd, yd)
        loss_host[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_ev[i & 1].record()
        tc = time.perf_counter()
        if i > 0:
            loss_ev[(i - 1) & 1].synchronize()
            losses.append(float(loss_host[(i - 1) & 1]))
        trace.append((1e3 * (tb - ta), 1e3 * (tc - tb), 1e3 * (time.perf_counter() - tc)))
    loss_ev[(args.steps - 1) & 1].synchronize()
    losses.append(float(loss_host[(args.steps - 1) & 1]))
    e3.record()
    if os.environ.get("BENCH_TRACE") and rank == 0:
        for r in trace:
            print("e2e step: h2d-issue %.2f ms, launch %.2f ms, wait %.2f ms" % r, file=sys.stderr)
    barrier()
    t_host = time.perf_counter() - t_host0
    ms2 = torch.tensor([max(e2.elapsed_time(e3), 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_ms = float(ms2) / args.steps
    clocks = sampler.stop() if rank == 0 else None   # after the e2e pass: tearing the NVML client down stays untimed

    if rank == 0:
        value = n_total / (ms_per_step * 1e-3)
        per_stage = {k: (v[0] / args.steps) for k, v in stages.items()}
        # dominant kernel and its algorithmic flops per step on this rank (per layer:
        # cond_fwd_a M^2, cond_fwd_b K M^2, syrk K M^2, cond_bwd_a K M^2, cond_bwd_b M^2 (+ the M^2 of the L-bar
        # reduction, which this design folds into the S_k algebra); x2 layers)
        # (the survey's F_pt also counts an M^2 L-bar reduction per layer that the S_k formulation removes altogether:
        #  it is part of step_roofline's flops_per_point, not of any kernel's algorithmic work)
        alg, executed = cfg.kernel_flops_per_point()
        dom = max(alg, key=lambda k: per_stage.get(k, 0.0))
        dom_ms = per_stage[dom]
        achieved = alg[dom] * n_local / (dom_ms * 1e-3) / 1e12
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel from the committed `ncu --set full`
        # capture of this command at 1 GPU (profiles/r01_kernel_traffic.json); scaled by the shard size
        traffic = None
        for tname in ("r02_kernel_traffic.json", "r01_kernel_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
                if cfg.which == 4 and dom in tj["per_launch_dram_bytes"]:
                    traffic = tj["per_launch_dram_bytes"][dom] * n_local / tj["points_per_launch"]
                    break
            except (OSError, KeyError, ValueError):
                continue
        roofline = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                    "frac": achieved / FP64_PEAK_TFLOPS, "traffic": traffic,
                    "launches_per_step": 2, "alg_flops_per_launch": alg[dom] * n_local / 2,
                    "all_kernels_tflops": {k: alg[k] * n_local / (per_stage[k] * 1e-3) / 1e12 for k in alg if per_stage.get(k)},
                    "executed_tflops": executed[dom] * n_local / (dom_ms * 1e-3) / 1e12,
                    "peak_source": "FP64 DMMA issue-rate peak measured on this pool's B200 by tools/fp64_peak.cu "
                                   "(profiles/r01_fp64_peak_microbench.txt); MEASURED_PEAKS.json has no FP64 entry; "
                                   f"cuBLAS DGEMM on the same box: {DGEMM_TFLOPS} TFLOP/s",
                    "kernel_ms_per_step": dom_ms, "kernel_share_of_step": dom_ms / ms_per_step}
        step_roof_ms = cfg.flops_per_point() * n_local / (FP64_PEAK_TFLOPS * 1e12) * 1e3
        exec_roof_ms = cfg.executed_algorithm_flops_per_point() * n_local / (FP64_PEAK_TFLOPS * 1e12) * 1e3
        line = {"metric": "elbo_fwd_bwd_points_per_s", "value": value, "unit": "points/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, world), "clocks": clocks,
                "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "points/s", "ms_per_step": e2e_ms,
                        "host_wall_ms_per_step": 1e3 * t_host / args.steps,
                        "h2d_bytes_per_step": int(Xh.numel() * 8 + Yh.numel() * 8) * world, "d2h_bytes_per_step": 8 * world,
                        "api": "for Xd, Yd in modulatedgps_b200.HostBatchStream(host_batches): "
                               "SMGP._training_loss((Xd, Yd)) + loss.backward() + loss.item()",
                        "h2d": "every step's inputs are copied from pinned host memory inside the timed region, on a copy "
                               "stream, while the previous step computes (two device buffer pairs)",
                        "d2h": "every step's loss (8 bytes) is copied to pinned host memory behind its own step and read by the "
                               "host inside the timed region, one step late; the last one before the closing event",
                        "losses_read": len(losses)},
                "gpu_launches": int(launches), "roofline": roofline,
                "step_roofline": {"flops_per_point": cfg.flops_per_point(), "roof_ms_per_step": step_roof_ms,
                                  "frac_of_fp64_peak": step_roof_ms / ms_per_step, "peak_tflops": FP64_PEAK_TFLOPS,
                                  "e2e_frac_of_fp64_peak": step_roof_ms / e2e_ms,
                                  "executed_algorithm": {
                                      "flops_per_point": cfg.executed_algorithm_flops_per_point(),
                                      "roof_ms_per_step": exec_roof_ms, "frac_of_fp64_peak": exec_roof_ms / ms_per_step,
                                      "note": "SURVEY's F_pt counts an M^2/point/layer L-bar reduction that the S_k "
                                              "reformulation removes; this is the count of the algorithm that runs"}},
                "stages_ms_per_step": per_stage, "loss": final_loss}
        if not args.no_cpu_baseline and world == 1:
            chunk = 4096 if cfg.which == 4 else 1024
            v, cores, times = cpu_points_per_s(cfg, chunk=chunk, reps=3, warmup=1)
            line["cpu_baseline"] = {"value": v, "unit": "points/s", "cores": cores, "kind": "port",
                                    "sample": f"{chunk} of the {cfg.N} config-#{cfg.which} points, 3 timed repetitions after 1 "
                                              "warm-up, oracle/svgp_mixture.py ELBO fwd+bwd (torch CPU autograd, explicit "
                                              "noise), conditionals de-duplicated over S",
                                    "note": "CPU restatement of the TF2/GPflow path; TF/GPflow are not installable here"}
            if cfg.which == 4:
                # what the reference's TF graph literally executes: the conditionals on the S tiled copies of X, at the
                # demos' batch size
                vl, _, tl = cpu_points_per_s(cfg, chunk=500, reps=3, warmup=1, literal=True)
                line["cpu_baseline"]["literal"] = {
                    "value": vl, "unit": "points/s", "cores": cores, "kind": "port",
                    "sample": "500 of the config-#4 points (the demos' batch size), 3 timed repetitions after 1 warm-up, the "
                              "same oracle with the S = 16 tiled copies of X evaluated as the reference's graph does "
                              "(LTA [S, K, M, N] materialised)"}
        else:
            line["cpu_baseline"] = None
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)                        # NCCL also warns on stdout while the communicator is torn down
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
