#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2m_scan_launches.csv python tools/scan_one.py > gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
