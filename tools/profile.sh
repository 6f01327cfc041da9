#!/bin/bash
# ncu evidence for the round (run under gpurun, 1 GPU).  $1 = round tag (e.g. r01)
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:cond_bwd_a|cond_fwd_b|syrk_kernel|cond_fwd_a|cond_bwd_b|mc_pass' -s 33 -c 11 -o /tmp/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
# the .ncu-rep with sources is larger than the 64 MiB that travels back: export the pages here
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page details --csv > gpurun_out/${TAG}_prof_details.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv > /tmp/${TAG}_prof_source.csv 2>/dev/null
gzip -c /tmp/${TAG}_prof_source.csv > gpurun_out/${TAG}_prof_source.csv.gz
ls -la /tmp/${TAG}_prof.ncu-rep gpurun_out/ | tail -n 12; tail -n 2 gpurun_out/${TAG}_plain.log | cut -c1-400
