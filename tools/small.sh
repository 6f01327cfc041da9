#!/bin/bash
# per-GPU behaviour of an 8-GPU run, on one GPU: the config-#4 step on 2^20 / 8 points (no all-reduce)
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --points 131072 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('131072 pts: ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), 'sum', round(sum(d['stages_ms_per_step'].values()),3))
print({k: round(v,3) for k,v in d['stages_ms_per_step'].items()})"
