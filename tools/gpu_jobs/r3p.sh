#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:scan_smth_down_kernel --launch-skip 1 -c 1 -o gpurun_out/scan_smth_down -f python tools/scan_one.py > gpurun_out/r3p_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/scan_smth_down.ncu-rep
