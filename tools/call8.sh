#!/bin/bash
# 8-GPU evidence: config #4 (headline) and config #5 at its full size
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
tail -c 1500 gpurun_out/r02_bench_8gpu.json
timeout 600 $TR bench.py --gpus 8 --config 5 --steps 2 --warmup 3 > gpurun_out/r02_bench_cfg5_8gpu.json 2> gpurun_out/r02_bench_cfg5_8gpu.err
tail -c 2500 gpurun_out/r02_bench_cfg5_8gpu.json; tail -n 5 gpurun_out/r02_bench_cfg5_8gpu.err
