#!/bin/bash
# round 2: the driver's sequence on one GPU -- build check, GPU tests, smoke, reference arm, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2s_smoke.log 2>&1; cat gpurun_out/r2s_smoke.log
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r2s_ref.json 2> gpurun_out/r2s_ref.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2s_bench.json
