#!/bin/bash
# round 2, run A (2 GPUs): GPU tests incl. the multi-device path, bench at N=1 and N=2, sanitizers
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
LDSR_TIMING=1 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench1.json 2> gpurun_out/r2a_bench1.err; echo "rc=$?" >> gpurun_out/r2a_bench1.err
LDSR_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench2.json 2> gpurun_out/r2a_bench2.err; echo "rc=$?" >> gpurun_out/r2a_bench2.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2a_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_memcheck.log
LDSR_MAX_GRID=2 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2a_memcheck_taskloop.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_memcheck_taskloop.log
SANITIZE_NITER=6 LDSR_MAX_GRID=2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py > gpurun_out/r2a_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_racecheck.log
tail -3 gpurun_out/r2a_pytest.log gpurun_out/r2a_memcheck.log gpurun_out/r2a_racecheck.log
