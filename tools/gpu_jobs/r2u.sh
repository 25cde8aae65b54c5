#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; tail -2 gpurun_out/r2u_pytest.log
python tools/profile_em.py np_restarts 1000 4 1 100 > gpurun_out/r2u_scan.log 2>&1
python tools/profile_em.py np_restarts 1000 4 1 1 >> gpurun_out/r2u_scan.log 2>&1
cat gpurun_out/r2u_scan.log | cut -c1-200
