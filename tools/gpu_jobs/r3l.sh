#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide" > gpurun_out/r3l_pytest.log 2>&1; tail -2 gpurun_out/r3l_pytest.log
python tools/profile_em.py synthetic 300 3 12 100 > gpurun_out/r3l_wide.log 2>&1; tail -3 gpurun_out/r3l_wide.log
python bench.py --gpus 1 --steps 3 --warmup 2 --no-configs --no-cpu > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3l_bench.json').read().strip().splitlines()[-1])
s=d['strong']; print('strong', s['value'], s['ms_per_step'], s['roofline']['frac'], s['e2e_sharded']['value'], (s.get('cpu_baseline') or {}).get('gpu_matches_oracle_on_sample'))
"
