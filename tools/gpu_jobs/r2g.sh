#!/bin/bash
# round 2, run G: wide kernel, blocking factor 1 / 2 / 4 in the row phases: phase clocks and plain timing
mkdir -p gpurun_out
for f in 1 2 4; do
  echo "=== FPL=$f (plain build)" >> gpurun_out/r2g_fpl.log
  LDSR_SO=$PWD/ldsr_b200/variants/lib_fpl$f.so python tools/profile_em.py synthetic 300 2 12 100 >> gpurun_out/r2g_fpl.log 2>&1
  echo "=== FPL=$f (phase clocks)" >> gpurun_out/r2g_fpl.log
  LDSR_SO=$PWD/ldsr_b200/variants/lib_clk$f.so python tools/profile_em.py synthetic 300 1 12 100 >> gpurun_out/r2g_fpl.log 2>&1
done
cat gpurun_out/r2g_fpl.log
