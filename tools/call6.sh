#!/bin/bash
./tools/loop_bisect 2>&1 | grep "MODE 5\|MODE 3" | tee gpurun_out/r02_loop_deep.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02g_tests.txt; cat gpurun_out/r02g_tests.txt
bash tools/quick_stages.sh zpre
