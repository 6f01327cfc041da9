#!/bin/bash
# last verification of a round: full GPU suite, smoke, the bench line (kept as the round's 1-GPU JSON)
python -m pytest tests -m gpu -q 2>&1 | tail -n 3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench_1gpu.json')); print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['step_roofline']['frac_of_fp64_peak'], 'dominant', d['roofline']['kernel'], d['roofline']['frac'], {k:round(v,3) for k,v in d['stages_ms_per_step'].items()})"
