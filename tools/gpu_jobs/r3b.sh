#!/bin/bash
mkdir -p gpurun_out
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py np_restarts 1000 2 1 100 > gpurun_out/r3b_clocks.log 2>&1
echo "=== single fit" >> gpurun_out/r3b_clocks.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py np_restarts 1000 2 1 1 >> gpurun_out/r3b_clocks.log 2>&1
tail -70 gpurun_out/r3b_clocks.log
