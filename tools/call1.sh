./tools/mix_bench > gpurun_out/r02_mix_bench.txt 2>&1
for s in 1 2 3 4 5; do python examples/replay_demos.py --demo john_doe --seed $s --out gpurun_out/jd_seed$s.json > /dev/null 2>&1; done
python -m pytest tests -m gpu -q 2>&1 | tail -n 12 > gpurun_out/r02b_tests.txt
bash tools/prof_one.sh 'cond_fwd_a|cond_bwd_b' 4 12 r02_ab > gpurun_out/r02_ab_prof.log 2>&1
cat gpurun_out/r02_mix_bench.txt; cat gpurun_out/r02b_tests.txt
