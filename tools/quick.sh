#!/bin/bash
# quick GPU check used during kernel work: parity tests, then a short bench with stage timers.  $1 = tag
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("ms/step", round(d["ms_per_step"], 2), "Mpts/s", round(d["value"] / 1e6, 3), "e2e ms", round(d["e2e"]["ms_per_step"], 2))
print({k: round(v, 2) for k, v in d["stages_ms_per_step"].items()})
PY
