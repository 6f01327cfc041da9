#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02h_tests.txt; cat gpurun_out/r02h_tests.txt
bash tools/quick_stages.sh bwda
python bench.py --config 5 --points 262144 --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/cfg5.err > gpurun_out/r02_cfg5_1gpu_b.json; python -c "
import json; d=json.load(open('gpurun_out/r02_cfg5_1gpu_b.json')); print('cfg5', d['ms_per_step'], {k:round(v,2) for k,v in d['stages_ms_per_step'].items()})"
