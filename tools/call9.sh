#!/bin/bash
# round-2 evidence at 1 GPU: bench JSON, ncu launch list, --set full of the six streaming kernels, smoke under ncu
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_1gpu.json')); print('ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'frac', d['step_roofline']['frac_of_fp64_peak'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_ref.err; tail -c 300 gpurun_out/r02_bench_reference_arm.json
sed -i 's/-s 270 -c 200/-s 330 -c 260/' tools/profile.sh
bash tools/profile.sh r02 > gpurun_out/r02_profile.log 2>&1; tail -n 4 gpurun_out/r02_profile.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_ncu.log 2>&1; echo "smoke under ncu rc=$?"; tail -n 2 gpurun_out/r02_smoke_plain.log
