#!/bin/bash
# The GPU test suite against a -DLDSR_DEBUG_CHECKS build of the library (plan-consistency asserts inside the
# kernels: a failed check prints its source line and traps) -- the stand-in for compute-sanitizer memcheck,
# which is closed on this pool.  Build first:
#   LDSR_BUILD_TAG=dbg LDSR_NVCC_EXTRA="-DLDSR_DEBUG_CHECKS" python -m ldsr_b200.build
mkdir -p gpurun_out
LDSR_SO=$PWD/ldsr_b200/variants/lib_dbg.so python -m pytest tests -m gpu -q > gpurun_out/debug_checks_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/debug_checks_pytest.log
LDSR_SO=$PWD/ldsr_b200/variants/lib_dbg.so LDSR_MAX_GRID=2 python tools/sanitize.py >> gpurun_out/debug_checks_pytest.log 2>&1
echo "task-loop rc=$?" >> gpurun_out/debug_checks_pytest.log
grep -c "LDSR_CHECK failed" gpurun_out/debug_checks_pytest.log
tail -8 gpurun_out/debug_checks_pytest.log
