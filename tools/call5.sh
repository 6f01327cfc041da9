#!/bin/bash
./tools/loop_bisect 2>&1 | grep "MODE 4\|MODE 3" > gpurun_out/r02_loop_ring.txt; cat gpurun_out/r02_loop_ring.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02f_tests.txt; cat gpurun_out/r02f_tests.txt
bash tools/quick_stages.sh pipe
bash tools/quick_stages.sh phased MGP_FWD_A_PHASED=1
