#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py -m gpu -q -x > gpurun_out/r3g_pytest.log 2>&1; tail -2 gpurun_out/r3g_pytest.log
for tb in 0 32 64 128; do
echo "=== LDSR_SCAN_TB=$tb" >> gpurun_out/r3g_scan.log
LDSR_SCAN_TB=$tb python tools/profile_scan.py 100000 1 8 >> gpurun_out/r3g_scan.log 2>&1
done
cat gpurun_out/r3g_scan.log
