#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final3_pytest.log 2>&1; tail -2 gpurun_out/final3_pytest.log
python tools/profile_crossover.py > gpurun_out/final3_crossover.log 2>&1; cat gpurun_out/final3_crossover.log
