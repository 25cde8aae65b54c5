#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py -m gpu -q -x > gpurun_out/r3e_pytest.log 2>&1; tail -3 gpurun_out/r3e_pytest.log
echo "=== groups (default)" > gpurun_out/r3e_scan.log
python tools/profile_scan.py 100000 1 8 32 >> gpurun_out/r3e_scan.log 2>&1
echo "=== LDSR_SCAN_GROUPS=1 (one CTA per fit)" >> gpurun_out/r3e_scan.log
LDSR_SCAN_GROUPS=1 python tools/profile_scan.py 100000 1 8 >> gpurun_out/r3e_scan.log 2>&1
echo "=== T = 20000, 1000000" >> gpurun_out/r3e_scan.log
python tools/profile_scan.py 20000 1 >> gpurun_out/r3e_scan.log 2>&1
python tools/profile_scan.py 1000000 1 >> gpurun_out/r3e_scan.log 2>&1
cat gpurun_out/r3e_scan.log
