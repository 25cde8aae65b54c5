#!/usr/bin/env python3
"""BASELINE.json config 5: state dimension 4, 20 proxies, T = 100 000, 10 % missing, 1..32 parameter
sets: associative-scan smoother vs the sequential recursion on the GPU (device time, CUDA events),
max abs difference of X, V, lik.
    python tools/profile_scan.py [T] [n_fits ...]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldsr_b200 import _lib  # noqa: E402


def model(rng, d, p, q, T):
    M = rng.standard_normal((d, d))
    A = 0.9 * M / np.max(np.abs(np.linalg.eigvals(M)))
    B = 0.3 * rng.standard_normal((d, p)) / np.sqrt(p)
    Cc = rng.standard_normal(d)
    D = 0.3 * rng.standard_normal(q) / np.sqrt(q)
    L = 0.4 * rng.standard_normal((d, d))
    Q = L @ L.T + 0.1 * np.eye(d)
    return np.concatenate([A.ravel(), B.ravel(), Cc, D, Q.ravel(), [0.3], np.zeros(d), np.eye(d).ravel()])


T = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
fits = [int(a) for a in sys.argv[2:]] or [1, 8, 32]
d, p, q = 4, 20, 20
rng = np.random.default_rng(20261018)
u = rng.standard_normal((p, T))
v = rng.standard_normal((q, T))
y = rng.standard_normal(T)
y[rng.uniform(size=T) < 0.1] = np.nan
for nf in fits:
    th = np.stack([model(rng, d, p, q, T) for _ in range(nf)])
    want = ("X", "V") if nf <= 8 else ()
    _lib.smoother_d(d, y, u, v, th, method=1, want=())  # warm-up
    seq = _lib.smoother_d(d, y, u, v, th, method=0, want=want)
    best = {}
    for chunk in (0, 4, 8, 16, 32, 64, 128):
        r = _lib.smoother_d(d, y, u, v, th, method=1, chunk=chunk, want=want)
        best[chunk] = r["kernel_ms"]
    sc = _lib.smoother_d(d, y, u, v, th, method=1, want=want)
    dl = float(np.max(np.abs(sc["lik"] - seq["lik"]) / np.abs(seq["lik"])))
    dx = float(np.max(np.abs(sc["X"] - seq["X"]))) if want else float("nan")
    dvv = float(np.max(np.abs(sc["V"] - seq["V"]))) if want else float("nan")
    print("d=%d p=q=%d T=%d fits=%d: sequential %.2f ms, scan %.3f ms (x%.0f); chunk sweep %s; max|dX| %.2e max|dV| %.2e rel dlik %.2e"
          % (d, p, T, nf, seq["kernel_ms"], sc["kernel_ms"], seq["kernel_ms"] / sc["kernel_ms"],
             {k: round(v, 3) for k, v in best.items()}, dx, dvv, dl))
