#!/bin/bash
# round 2, run C (1 GPU): wide-input kernel v2 + scan kernel -- GPU tests, config 3 / config 1 timing, ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
python tools/profile_em.py synthetic 1000 2 48 100 > gpurun_out/r2c_profile_full.log 2>&1
python tools/profile_em.py synthetic 300 2 12 100 > gpurun_out/r2c_profile_plain.log 2>&1
python bench.py --steps 3 --warmup 3 --no-strong > gpurun_out/r2c_bench_nostrong.json 2> gpurun_out/r2c_bench_nostrong.err
ncu --set full --clock-control none --import-source on -k regex:em_wide_kernel -c 1 -o gpurun_out/em_r02_wide2 -f python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2c_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_scan_kernel -c 1 -o gpurun_out/em_r02_scan -f python tools/profile_em.py np_restarts 300 1 1 100 > gpurun_out/r2c_ncu_scan.log 2>&1
cat gpurun_out/r2c_profile_full.log gpurun_out/r2c_profile_plain.log
