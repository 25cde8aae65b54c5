#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r3u_pytest.log 2>&1; tail -2 gpurun_out/r3u_pytest.log
python tools/profile_crossover.py > gpurun_out/r3u_crossover.log 2>&1; cat gpurun_out/r3u_crossover.log
python bench.py --gpus 1 --steps 5 --warmup 3 --no-strong --no-cpu > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3u_bench.json').read().strip().splitlines()[-1])
c=d['configs']['config1']; print('c1', c['restarts_100']['ms'], c['single_fit']['ms'], 'headline', d['ms_per_step'])
"
