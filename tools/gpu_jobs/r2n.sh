#!/bin/bash
mkdir -p gpurun_out
python tools/profile_em.py synthetic 300 2 12 100 > gpurun_out/r2n.log 2>&1
python tools/profile_em.py synthetic 1000 2 48 1000 >> gpurun_out/r2n.log 2>&1
cat gpurun_out/r2n.log
python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
