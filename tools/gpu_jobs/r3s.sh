#!/bin/bash
mkdir -p gpurun_out
python tools/profile_crossover.py wide > gpurun_out/r3s_crossover_wide.log 2>&1; cat gpurun_out/r3s_crossover_wide.log
