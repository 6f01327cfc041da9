#!/bin/bash
# durations (ncu, serialised) of the kernels matching regex $1 over one bench step.  $2 = number of launches to keep
RE=${1:-chol}; CNT=${2:-24}
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RE" -c $CNT --csv --log-file gpurun_out/klist.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
python - <<PY
import csv
rows=[l for l in open('gpurun_out/klist.csv') if not l.startswith('==')]
agg={}
for x in csv.DictReader(rows):
    agg.setdefault(x['Kernel Name'][:48]+' '+x['Grid Size'],[]).append(float(x['Metric Value'])/1e3)
for k,v in agg.items(): print('%-70s n=%d  min %.1f  med %.1f us' % (k, len(v), min(v), sorted(v)[len(v)//2]))
PY
