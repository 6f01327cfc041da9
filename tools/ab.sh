#!/bin/bash
# A/B of tuning hooks: each line "VAR=val VAR=val" is one configuration; prints stage times
run() {
  env $1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
s=d['stages_ms_per_step']
print('$1', 'step', round(d['ms_per_step'],2), {k: round(v,2) for k,v in s.items() if k in ('cond_fwd_a','cond_fwd_b','syrk','cond_bwd_a','cond_bwd_b','mc_pass')})"
}
for cfg in "$@"; do run "$cfg"; done
