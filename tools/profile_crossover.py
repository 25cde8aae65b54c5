#!/usr/bin/env python3
"""Where the small-batch scan kernel (one CTA per fit) stops paying against the time-split kernel (32 fits per
CTA): cvLDS-shaped jobs on NP-413 with 20 restarts per fold, 1000 iterations, device-resident (plan.em).
    python tools/profile_crossover.py [wide]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldsr_b200 import _lib, workloads as W  # noqa: E402

wide = len(sys.argv) > 1 and sys.argv[1] == "wide"  # one synthetic station, ten inputs, T = 400 (config 3's shape)
for folds in ((5, 10, 15, 20, 30, 40) if wide else (10, 20, 30, 45, 55, 60, 65, 70, 90, 120)):
    w = W.synthetic_stations(n_stations=1, n_folds=folds, n_restarts=20) if wide else W.np_cv(folds, 20)
    plan = _lib.Plan(w["series"], w["group_series"], w["held"], w["fit_group"], w["theta0"])
    out = []
    for variant in ((5, 4) if wide else (5, 3)):
        best = None
        try:
            for _ in range(4):
                t = time.perf_counter()
                st = plan.em(1000, 1e-5, variant=variant)
                dt = (time.perf_counter() - t) * 1e3
                best = dt if best is None else min(best, dt)
            out.append("%s %.2f ms" % (st["kernel"], best))
        except Exception as e:  # the scan kernel refuses more than SCAN_MAX_FITS
            out.append("variant %d: %s" % (variant, str(e)[:60]))
    print("%5d fits: %s" % (folds * 20, " | ".join(out)), flush=True)
