#!/usr/bin/env python3
"""Small driver for ncu: runs the device-resident EM path a few times on one workload.
    python tools/profile_em.py [workload] [niter] [repeats] [n_folds] [n_restarts]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldsr_b200 import _lib, workloads as W  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "np_cv"
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nfold = int(sys.argv[4]) if len(sys.argv) > 4 else 100
nrest = int(sys.argv[5]) if len(sys.argv) > 5 else 100
if name == "np_cv":
    w = W.np_cv(nfold, nrest)
elif name == "synthetic":
    w = W.synthetic_stations(n_stations=nfold, n_folds=10, n_restarts=nrest)
else:
    w = W.np_restarts(nrest)
plan = _lib.Plan(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"])
for i in range(reps):
    t = time.perf_counter()
    st = plan.em(niter, 1e-5)
    dt = time.perf_counter() - t
    print("%s niter=%d: %.2f ms, %s" % (w["name"], niter, dt * 1e3, st))
