#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r3d_pytest.log 2>&1; tail -2 gpurun_out/r3d_pytest.log
for i in 1 2; do
python bench.py --gpus 1 --steps 5 --warmup 3 --no-strong --no-cpu > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3d_bench.json').read().strip().splitlines()[-1])
c=d['configs']['config1']; print('c1', c['restarts_100']['ms'], c['single_fit']['ms'], 'headline', d['ms_per_step'])
"
done
