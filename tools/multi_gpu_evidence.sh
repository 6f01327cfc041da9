#!/bin/bash
# multi-GPU evidence: bash tools/multi_gpu_evidence.sh N [cfg5]   (under gpurun --gpus N)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print('N=$N cfg4 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'Mpts/s', d['value']/1e6, {k:round(v,3) for k,v in d['stages_ms_per_step'].items()})"
if [ "$2" = "cfg5" ]; then
timeout 600 $TR bench.py --gpus $N --config 5 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg5_${N}gpu.json 2> gpurun_out/r02_bench_cfg5_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg5_${N}gpu.json')); print('N=$N cfg5 ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'Mpts/s', d['value']/1e6, 'frac', d['step_roofline']['frac_of_fp64_peak'], {k:round(v,1) for k,v in d['stages_ms_per_step'].items()})"
fi
