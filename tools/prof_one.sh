#!/bin/bash
# ncu --set full of the kernels matching $1 (regex), $2 launches after skipping $3; exports raw + source pages.  $4 = tag
RE=${1:-syrk_kernel}; CNT=${2:-2}; SKIP=${3:-8}; TAG=${4:-prof}
CMD=${CMD:-"python bench.py --steps 1 --warmup 3 --no-cpu-baseline"}
ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -o /tmp/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page source --csv > /tmp/${TAG}_source.csv 2>/dev/null
gzip -c /tmp/${TAG}_source.csv > gpurun_out/${TAG}_source.csv.gz
ls -la gpurun_out/${TAG}*; tail -n 3 gpurun_out/${TAG}_ncu.log | cut -c1-300
