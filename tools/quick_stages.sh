#!/bin/bash
# per-stage timings of the config-#4 step under the environment given on the command line: tools/quick_stages.sh TAG [VAR=VAL ...]
TAG=$1; shift
env "$@" timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$TAG', round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stages_ms_per_step'].items() if v})
"
