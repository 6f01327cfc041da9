#!/bin/bash
./tools/loop_bisect > gpurun_out/r02_loop_bisect.txt 2>&1; cat gpurun_out/r02_loop_bisect.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02e_tests.txt; cat gpurun_out/r02e_tests.txt
bash tools/quick_stages.sh pipe_ring
bash tools/quick_stages.sh phased_2cta MGP_FWD_A_PHASED=1 MGP_BWD_B_2CTA=1
bash tools/prof_one.sh 'cond_fwd_a_pipe|cond_bwd_b_ring' 4 12 r02_pr > gpurun_out/r02_pr_prof.log 2>&1
