#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "repeated_runs or shared_by_iterations or ranked_assignment" > gpurun_out/r3q_pytest.log 2>&1; tail -3 gpurun_out/r3q_pytest.log
