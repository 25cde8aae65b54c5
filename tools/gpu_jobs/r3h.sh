#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_scan.py -m gpu -q -x > gpurun_out/r3h_pytest.log 2>&1; tail -3 gpurun_out/r3h_pytest.log
python tools/profile_scan.py 100000 1 8 32 > gpurun_out/r3h_scan.log 2>&1
python tools/profile_scan.py 20000 1 >> gpurun_out/r3h_scan.log 2>&1
python tools/profile_scan.py 1000000 1 >> gpurun_out/r3h_scan.log 2>&1
cat gpurun_out/r3h_scan.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3h_launches.csv python tools/scan_one.py > gpurun_out/r3h_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r3h_launches.csv | tail -14 | awk -F'","' '{print $5, $(NF-1), $NF}'
