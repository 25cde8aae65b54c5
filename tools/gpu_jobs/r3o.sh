#!/bin/bash
mkdir -p gpurun_out
t0=$(date +%s)
python bench.py > gpurun_out/r3o_default.json 2> gpurun_out/r3o_default.err; echo "default bench rc=$?"
t1=$(date +%s); echo "wall $((t1-t0)) s"
python -c "
import json
d=json.loads(open('gpurun_out/r3o_default.json').read().strip().splitlines()[-1])
print(d['steps'], d['warmup'], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
"
