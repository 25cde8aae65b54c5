#!/bin/bash
mkdir -p gpurun_out
python tools/profile_crossover.py > gpurun_out/r3r_crossover.log 2>&1; cat gpurun_out/r3r_crossover.log
