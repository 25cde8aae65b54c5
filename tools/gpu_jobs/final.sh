#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'], d['cpu_baseline']['gpu_matches_oracle_on_sample'])
s=d['strong']; print('strong', s['value'], s['ms_per_step'], s['roofline']['frac'], s['e2e_sharded']['value'], s['cpu_baseline']['gpu_matches_oracle_on_sample'])
c=d['configs']; print('c1', c['config1']['restarts_100']['ms'], c['config1']['single_fit']['ms'], c['config1']['restarts_100']['gpu_matches_oracle'], 'c4', c['config4']['device_ms'], c['config4']['wall_ms'], 'c5', c['config5']['scan_ms'])
"
