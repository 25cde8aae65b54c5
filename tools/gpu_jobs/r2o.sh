#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2o.log
for v in nw10g1 nw12g1; do
  echo "=== $v" >> gpurun_out/r2o.log
  LDSR_SO=$PWD/ldsr_b200/variants/lib_$v.so python tools/profile_em.py synthetic 300 2 12 100 >> gpurun_out/r2o.log 2>&1
  LDSR_SO=$PWD/ldsr_b200/variants/lib_$v.so python tools/profile_em.py synthetic 1000 1 48 1000 >> gpurun_out/r2o.log 2>&1
done
cat gpurun_out/r2o.log
