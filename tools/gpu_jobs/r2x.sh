#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shared_by_iterations or full_size or task_loop or neighbours or chunking or poll or cv_np413" > gpurun_out/r2x_pytest.log 2>&1; tail -3 gpurun_out/r2x_pytest.log
for mode in rank norank; do
  if [ $mode = norank ]; then export LDSR_NO_RANK=1; else unset LDSR_NO_RANK; fi
  LDSR_TIMING=1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-strong --no-configs > gpurun_out/r2x_bench_$mode.json 2> gpurun_out/r2x_bench_$mode.err; echo "bench $mode rc=$?"
  grep "chunk" gpurun_out/r2x_bench_$mode.err | tail -12
  python -c "
import json
d=json.loads(open('gpurun_out/r2x_bench_$mode.json').read().strip().splitlines()[-1])
print('$mode value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle_on_sample'])
"
done
