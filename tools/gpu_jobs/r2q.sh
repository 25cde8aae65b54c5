#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
python tools/profile_em.py np_restarts 1000 3 1 100 > gpurun_out/r2q_scan.log 2>&1
python tools/profile_em.py synthetic 1000 3 1 10 >> gpurun_out/r2q_scan.log 2>&1
cat gpurun_out/r2q_scan.log
