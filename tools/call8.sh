#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -n 5 > gpurun_out/r02i_tests.txt; cat gpurun_out/r02i_tests.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 600 gpurun_out/r02_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['step_roofline']['frac_of_fp64_peak'], d['roofline']['all_kernels_tflops'])"
