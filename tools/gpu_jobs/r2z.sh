#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shared_by_iterations or full_size or task_loop or neighbours or chunking or poll or cv_np413 or wide_multi" > gpurun_out/r2z_pytest.log 2>&1; tail -3 gpurun_out/r2z_pytest.log
LDSR_TIMING=1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-strong --no-configs > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"
grep "chunk\|ldsr_em_batch" gpurun_out/r2z_bench.err | tail -14
python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle_on_sample'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_split_kernel --launch-skip 10 --launch-count 6 -o gpurun_out/r2z_split python bench.py --steps 1 --warmup 1 --no-strong --no-configs --no-cpu > gpurun_out/r2z_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2z_split.ncu-rep
