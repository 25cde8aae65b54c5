#!/bin/bash
# round 2, run I: config 3 at full size, wide-input kernel vs time-split kernel; full GPU tests; bench line
mkdir -p gpurun_out
echo "=== wide-input kernel (auto)" > gpurun_out/r2i.log
python tools/profile_em.py synthetic 1000 2 48 1000 >> gpurun_out/r2i.log 2>&1
echo "=== time-split kernel (variant 3)" >> gpurun_out/r2i.log
LDSR_VARIANT=3 python tools/profile_em.py synthetic 1000 2 48 1000 >> gpurun_out/r2i.log 2>&1
cat gpurun_out/r2i.log
python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; tail -3 gpurun_out/r2i_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench1.json 2> gpurun_out/r2i_bench1.err; echo "bench rc=$?"
