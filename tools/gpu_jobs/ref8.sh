#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/ref8.json 2> gpurun_out/ref8.err; echo "ref rc=$?"
wc -l gpurun_out/ref8.json; python -c "
import json
lines=[l for l in open('gpurun_out/ref8.json').read().strip().splitlines() if l.startswith('{')]
print(len(lines)); d=json.loads(lines[-1]); print(d['impl'], d['value'], d['n_gpus'], d['cpu_baseline'])
"
