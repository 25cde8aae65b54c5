#!/bin/bash
# round 2, run F (1 GPU): wide kernel v3 (register-blocked row phases): tests, timing, ncu; rep timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
python tools/profile_em.py synthetic 1000 2 48 100 > gpurun_out/r2f_profile.log 2>&1
python tools/profile_em.py synthetic 300 2 12 100 >> gpurun_out/r2f_profile.log 2>&1
cat gpurun_out/r2f_profile.log
python tools/profile_rep.py > gpurun_out/r2f_rep.log 2>&1; cat gpurun_out/r2f_rep.log
ncu --set full --clock-control none --import-source on -k regex:em_wide_kernel -c 1 -o gpurun_out/em_r02_wide3 -f python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2f_ncu_full.log 2>&1
