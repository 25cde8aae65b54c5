#!/bin/bash
# round 2, run J (1 GPU): final ncu captures -- launch list of the bench, wide-input kernel, scan kernel, rep kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-strong --no-configs > gpurun_out/r2j_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_wide_kernel -c 1 -o gpurun_out/em_r02_wide4 -f python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r2j_ncu_wide.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_scan_kernel -c 1 -o gpurun_out/em_r02_scan2 -f python tools/profile_em.py np_restarts 300 1 1 100 > gpurun_out/r2j_ncu_scan.log 2>&1
ncu --set full --clock-control none -k regex:rep_kernel -c 1 -o gpurun_out/rep_r02 -f python tools/profile_rep.py 20000 > gpurun_out/r2j_ncu_rep.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:em_split_kernel -s 1 -c 1 -o gpurun_out/em_r02_split -f python tools/profile_em.py np_cv 300 1 > gpurun_out/r2j_ncu_split.log 2>&1
ls -la gpurun_out
