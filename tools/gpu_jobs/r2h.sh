#!/bin/bash
# round 2, run H: wide kernel v4 (Dv / Sxv over the list of observed-unit steps, 2 fits per lane): piece balancing sweep
mkdir -p gpurun_out
for cm in 720 900 1100 1400; do
  echo "=== cost_m=$cm" >> gpurun_out/r2h.log
  LDSR_WIDE_COST_M=$cm python tools/profile_em.py synthetic 300 2 12 100 >> gpurun_out/r2h.log 2>&1
done
echo "=== phase clocks, cost_m=720" >> gpurun_out/r2h.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py synthetic 300 1 12 100 >> gpurun_out/r2h.log 2>&1
echo "=== phase clocks, cost_m=1100" >> gpurun_out/r2h.log
LDSR_WIDE_COST_M=1100 LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py synthetic 300 1 12 100 >> gpurun_out/r2h.log 2>&1
echo "=== old time-split kernel (variant 3)" >> gpurun_out/r2h.log
LDSR_VARIANT=3 python tools/profile_em.py synthetic 300 2 12 100 >> gpurun_out/r2h.log 2>&1
cat gpurun_out/r2h.log
python -m pytest tests -m gpu -q -k "wide" > gpurun_out/r2h_pytest.log 2>&1; tail -3 gpurun_out/r2h_pytest.log
