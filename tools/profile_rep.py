#!/usr/bin/env python3
"""BASELINE.json config 4: LDS_rep -- 100 000 stochastic replicates of the fitted NP model
(theta = data/theta.rda, u = v = t(NPpc), T = 813) -> 81.3 M (x, y, Q) triples = 1.95 GB, and
propagate() on 100 000 parameter sets.  Wall clock through the host-buffer C ABI (H2D + kernels +
D2H of all three arrays) and the implied HBM write rate of the kernels alone is in profiles/.
    python tools/profile_rep.py [n_reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldsr_b200 import _lib  # noqa: E402
from tests import data  # noqa: E402

n_reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
d = data.load("np.json")
th = data.theta_of(d["theta"])
y, u, mu, inst = data.np_case(1, 1200)
T = y.size
for want in (("simX", "simY", "simQ"), ("simQ",)):
    for rep in range(2):
        t0 = time.perf_counter()
        r = _lib.rep_batch(th, u, u, T, n_reps, seed=20261018, mu=mu, want=want)
        dt = time.perf_counter() - t0
    nb = sum(a.nbytes for k, a in r.items() if k != "device_ms")
    print("LDS_rep %d replicates x T=%d, outputs %s: %.1f ms wall (%.2f GB to the host, %.1f M triples/s); "
          "kernels %.3f ms on the device = %.0f GB/s of output"
          % (n_reps, T, "+".join(want), dt * 1e3, nb / 1e9, n_reps * T / dt / 1e6, r["device_ms"],
             nb / (r["device_ms"] * 1e-3) / 1e9))
# the same replicates as `set.seed(1); LDS_rep(...)` in R: R's stream generated on the device (r_rng.cuh)
for rep in range(2):
    t0 = time.perf_counter()
    r = _lib.rep_batch(th, u, u, T, n_reps, mu=mu, want=("simQ",), r_seed=1)
    dt = time.perf_counter() - t0
print("LDS_rep %d replicates, simQ, R's own stream (set.seed(1)) generated on the device: %.1f ms wall "
      "(%.1f M normals); replicate kernels %.3f ms" % (n_reps, dt * 1e3, n_reps * (1 + 2 * T) / 1e6, r["device_ms"]))
# propagate on n_reps thetas (perturbed copies of the NP theta)
rng = np.random.default_rng(1)
n_th = min(n_reps, 100_000)
ths = np.tile(th, (n_th, 1)) * (1 + 0.01 * rng.standard_normal((n_th, th.size)))
ser = [dict(y=y, u=u, v=u)]
for rep in range(2):
    t0 = time.perf_counter()
    r = _lib.propagate_batch(ser, [0], [np.array([], dtype=np.int32)], np.zeros(n_th, dtype=np.int32), ths)
    dt = time.perf_counter() - t0
print("propagate on %d parameter sets x T=%d: %.1f ms wall (%.2f M (theta,step)/s)" % (n_th, T, dt * 1e3, n_th * T / dt / 1e6))
