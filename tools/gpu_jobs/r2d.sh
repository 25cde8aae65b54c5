#!/bin/bash
# round 2, run D (1 GPU): GPU tests; phase clocks of the wide kernel (development build); scan kernel timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2d_clocks.log 2>&1
cat gpurun_out/r2d_clocks.log
python tools/profile_em.py np_restarts 1000 3 1 100 > gpurun_out/r2d_scan_100.log 2>&1
python tools/profile_em.py np_restarts 1000 3 1 1 >> gpurun_out/r2d_scan_100.log 2>&1
python tools/profile_em.py np_restarts 1000 3 1 500 >> gpurun_out/r2d_scan_100.log 2>&1
cat gpurun_out/r2d_scan_100.log
