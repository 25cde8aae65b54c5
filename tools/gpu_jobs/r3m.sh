#!/bin/bash
# final ncu --set full captures of the wide-input kernel and the scan kernel (after their programs exited 0 without ncu)
mkdir -p gpurun_out
python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r3m_wide_plain.log 2>&1 && tail -1 gpurun_out/r3m_wide_plain.log
ncu --set full --clock-control none --import-source on -k regex:em_wide_kernel -c 1 -o gpurun_out/em_r02_wide_final -f python tools/profile_em.py synthetic 300 1 12 100 > gpurun_out/r3m_ncu_wide.log 2>&1; echo "ncu wide rc=$?"
python tools/profile_em.py np_restarts 300 1 1 100 > gpurun_out/r3m_scan_plain.log 2>&1 && tail -1 gpurun_out/r3m_scan_plain.log
ncu --set full --clock-control none --import-source on -k regex:em_scan_kernel -c 1 -o gpurun_out/em_r02_scan_final -f python tools/profile_em.py np_restarts 300 1 1 100 > gpurun_out/r3m_ncu_scan.log 2>&1; echo "ncu scan rc=$?"
ls -la gpurun_out/*_final.ncu-rep
