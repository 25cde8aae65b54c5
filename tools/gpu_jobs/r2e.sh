#!/bin/bash
# round 2, run E (1 GPU): phase clocks of the wide kernel; scan kernel L=2 vs L=4; LDS_rep timing; bench line
mkdir -p gpurun_out
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2e_clocks.log 2>&1
cat gpurun_out/r2e_clocks.log
for L in 2 4; do
  echo "== LDSR_SCAN_L=$L" >> gpurun_out/r2e_scan.log
  LDSR_SCAN_L=$L python tools/profile_em.py np_restarts 1000 3 1 100 >> gpurun_out/r2e_scan.log 2>&1
  LDSR_SCAN_L=$L python tools/profile_em.py np_restarts 1000 3 1 300 >> gpurun_out/r2e_scan.log 2>&1
done
cat gpurun_out/r2e_scan.log
python tools/profile_rep.py > gpurun_out/r2e_rep.log 2>&1; cat gpurun_out/r2e_rep.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench1.json 2> gpurun_out/r2e_bench1.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2e_bench1.json
