#!/bin/bash
mkdir -p gpurun_out
LDSR_NO_SHARE=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not shared_by_iterations and not ranked_assignment" > gpurun_out/r3n_noshare.log 2>&1; tail -1 gpurun_out/r3n_noshare.log
LDSR_NO_RANK=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r3n_norank.log 2>&1; tail -1 gpurun_out/r3n_norank.log
LDSR_SEQ_TRAJ=1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r3n_seqtraj.log 2>&1; tail -1 gpurun_out/r3n_seqtraj.log
/usr/bin/time -v python bench.py > gpurun_out/r3n_default.json 2> gpurun_out/r3n_default.err; echo "default bench rc=$?"; grep "Elapsed (wall" gpurun_out/r3n_default.err
python -c "
import json
d=json.loads(open('gpurun_out/r3n_default.json').read().strip().splitlines()[-1])
print(d['steps'], d['warmup'], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])
"
