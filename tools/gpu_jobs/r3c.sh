#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r3c_pytest.log 2>&1; tail -3 gpurun_out/r3c_pytest.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_clk.so python tools/profile_em.py np_restarts 1000 2 1 100 > gpurun_out/r3c_clocks.log 2>&1
grep "thread  0\|niter" gpurun_out/r3c_clocks.log | tail -17
python bench.py --gpus 1 --steps 5 --warmup 3 --no-strong --no-cpu > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r3c_bench.json').read().strip().splitlines()[-1])
c=d['configs']; print('c1', c['config1'])
"
