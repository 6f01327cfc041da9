#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/r02j_tests.txt; cat gpurun_out/r02j_tests.txt
for p in 0 15 30 50 70; do bash tools/quick_stages.sh skew$p MGP_SKEW_PCT=$p; done
