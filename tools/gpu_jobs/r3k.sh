#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r3k_pytest.log 2>&1; tail -2 gpurun_out/r3k_pytest.log
LDSR_TIMING=1 python bench.py --gpus 1 --steps 10 --warmup 3 --no-strong --no-cpu > gpurun_out/r3k_bench.json 2> gpurun_out/r3k_bench.err
grep "ldsr_em_batch: build" gpurun_out/r3k_bench.err | tail -4
python -c "
import json
d=json.loads(open('gpurun_out/r3k_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
c=d['configs']; print('c1', c['config1']['restarts_100']['ms'], c['config1']['single_fit']['ms'], 'c5', c['config5']['scan_ms'])
"
