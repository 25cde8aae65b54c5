#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2r.log
for i in 1 2; do
echo "== new" >> gpurun_out/r2r.log
python tools/profile_em.py np_restarts 1000 3 1 100 >> gpurun_out/r2r.log 2>&1
echo "== old (HEAD~2)" >> gpurun_out/r2r.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_oldscan.so python tools/profile_em.py np_restarts 1000 3 1 100 >> gpurun_out/r2r.log 2>&1
done
grep -v "^LDS.*: [4-9]" gpurun_out/r2r.log
