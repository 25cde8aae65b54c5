#!/bin/bash
# round 2, run B (1 GPU): wide-input kernel -- GPU tests, bench with the strong block (config 3), ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
LDSR_TIMING=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench1.json 2> gpurun_out/r2b_bench1.err; echo "rc=$?" >> gpurun_out/r2b_bench1.err
python tools/profile_em.py synthetic 300 2 12 100 > gpurun_out/r2b_profile_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_wide.csv python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2b_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_wide_kernel -c 1 -o gpurun_out/em_r02_wide -f python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2b_ncu_full.log 2>&1
ls -la gpurun_out
