#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r3i_pytest.log 2>&1; tail -3 gpurun_out/r3i_pytest.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3i_ref.json 2> gpurun_out/r3i_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r3i_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle_on_sample'], d['cpu_baseline']['sample'])
s=d['strong']; print('strong', s['value'], s['ms_per_step'], s['roofline']['frac'], s['e2e_sharded']['value'], s['cpu_baseline']['gpu_matches_oracle_on_sample'])
c=d['configs']; print('c1', c['config1']['restarts_100']['ms'], c['config1']['single_fit']['ms'], c['config1']['restarts_100']['gpu_matches_oracle'], 'c4', c['config4']['device_ms'], c['config4']['wall_ms'], 'c5', c['config5']['scan_ms'])
r=json.loads(open('gpurun_out/r3i_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['sample'], r['config']==d['config'])
"
