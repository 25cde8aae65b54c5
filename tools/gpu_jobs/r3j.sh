#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 50 75 100 125 150 200 250; do
python bench.py --gpus 1 --steps 8 --warmup 3 --no-strong --no-configs --no-cpu --chunk $c > gpurun_out/r3j_bench_$c.json 2> gpurun_out/r3j_bench_$c.err
python -c "
import json
d=json.loads(open('gpurun_out/r3j_bench_$c.json').read().strip().splitlines()[-1])
print('chunk $c ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
"
done
