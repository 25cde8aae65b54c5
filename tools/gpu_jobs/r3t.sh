#!/bin/bash
mkdir -p gpurun_out
echo "=== product" > gpurun_out/r3t.log
python tools/profile_crossover.py >> gpurun_out/r3t.log 2>&1
echo "=== 168 registers (3 CTAs per SM)" >> gpurun_out/r3t.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_r168.so python tools/profile_crossover.py >> gpurun_out/r3t.log 2>&1
grep "===\| 200 \| 400 \| 600 \| 900 " gpurun_out/r3t.log
