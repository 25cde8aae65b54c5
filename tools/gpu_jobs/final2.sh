#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests -m gpu -q > gpurun_out/final2_pytest.log 2>&1; tail -2 gpurun_out/final2_pytest.log
