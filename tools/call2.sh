#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 > gpurun_out/r02c_tests.txt; cat gpurun_out/r02c_tests.txt
bash tools/quick_stages.sh stash
bash tools/quick_stages.sh nostash MGP_NO_KUF_STASH=1
