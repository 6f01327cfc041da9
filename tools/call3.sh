#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02d_tests.txt; cat gpurun_out/r02d_tests.txt
bash tools/quick_stages.sh ring
bash tools/quick_stages.sh twocta_pf MGP_BWD_B_2CTA=1
bash tools/quick_stages.sh twocta_nopf MGP_BWD_B_2CTA=1 MGP_BWD_B_NO_PF=1
