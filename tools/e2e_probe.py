import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from ldsr_b200 import _lib, workloads as W
w = W.np_cv(100, 100)
ctx = _lib.Ctx(devices=[0])
for i in range(4):
    t0 = time.perf_counter()
    pb = _lib.PackedBatch(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"])
    t1 = time.perf_counter()
    out = _lib.EmOutputs(pb, 1000, False, True)
    t2 = time.perf_counter()
    r = _lib.em_batch(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"], 1000, 1e-5, ctx=ctx)
    t3 = time.perf_counter()
    print("python: pack %.3f ms, outputs %.3f ms, em_batch total %.3f ms" % ((t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3))
