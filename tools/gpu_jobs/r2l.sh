#!/bin/bash
# round 2, run L (1 GPU): block-level scan stage of the d-dimensional smoother; new tests; LDS_rep single-kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; tail -4 gpurun_out/r2l_pytest.log
python tools/profile_scan.py > gpurun_out/r2l_scan.log 2>&1; cat gpurun_out/r2l_scan.log
python tools/profile_rep.py > gpurun_out/r2l_rep.log 2>&1; cat gpurun_out/r2l_rep.log
