#!/usr/bin/env python3
"""Small runs of every EM kernel path, meant for compute-sanitizer:

    compute-sanitizer --tool memcheck  --error-exitcode 1 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 1 python tools/sanitize.py

Cases: (1) the cvLDS job in small (NP-413, folds x restarts, several groups per CTA), time-split and
lane kernels; (2) wide inputs (padded width 10 and 24); (3) the task loop (LDSR_MAX_GRID=2 in the
environment makes CTAs take several tasks and re-use their shared memory); (4) the single-step kernels,
replicates and the scan smoother.  Every result is compared with the oracle, so a hazard that changes a
number fails here even if the tool misses it."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ldsr_b200 import _lib  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests import data  # noqa: E402
from tests.test_gpu_parity import rand_theta0  # noqa: E402

NITER = int(os.environ.get("SANITIZE_NITER", "12"))


def check(ser, gs, held, fg, th0, niter, **kw):
    g = _lib.em_batch(ser, gs, held, fg, th0, niter, 1e-5, **kw)
    o = O.em_batch(ser, gs, held, fg, th0, niter, 1e-5)
    assert np.array_equal(g["iters"], o["iters"]) and np.array_equal(g["best"], o["best"])
    assert np.allclose(g["lik"], o["lik"], rtol=1e-9, atol=0)
    assert np.allclose(g["theta"], o["theta"], rtol=1e-6, atol=1e-12, equal_nan=True)


def main():
    y, u, mu, inst = data.np_case(401, 1600)
    rng = np.random.default_rng(413)
    n_folds, n_rest = 12, 8  # 96 fits = 3 CTAs of the time-split kernel, 4 groups per CTA
    held = [np.sort(rng.choice(inst, 11, replace=False)) for _ in range(n_folds)]
    fg = np.repeat(np.arange(n_folds), n_rest)
    th0 = rand_theta0(rng, 3, 3, n_folds * n_rest)
    ser = [dict(y=y, u=u, v=u)]
    for variant in (3, 2, 1):
        check(ser, np.zeros(n_folds, dtype=int), held, fg, th0, NITER, variant=variant, chunk_iters=5)
        print("np413 variant %d ok" % variant, flush=True)
    for p, q in ((10, 10), (18, 20)):
        T = 150
        uu = rng.standard_normal((p, T))
        vv = rng.standard_normal((q, T)) if p != q else uu
        yy = 0.3 * rng.standard_normal(T)
        yy[:60] = np.nan
        t0 = rand_theta0(rng, p, q, 40)
        hh = [np.array([70, 71, 100]), np.array([], dtype=int)]
        for variant in (3, 2):
            check([dict(y=yy, u=uu, v=vv)], [0, 0], hh, np.repeat([0, 1], 20), t0, NITER, variant=variant, chunk_iters=5)
        print("wide %dx%d ok" % (p, q), flush=True)
    # single-step kernels, replicates, scan smoother
    th = rand_theta0(rng, 3, 3, 5)
    s = _lib.smoother_batch(ser, [0], [held[0]], np.zeros(5, dtype=int), th)
    so = O.smoother_batch(ser, [0], [held[0]], np.zeros(5, dtype=int), th) if hasattr(O, "smoother_batch") else None
    if so is not None:
        assert np.allclose(s["lik"], so["lik"], rtol=1e-9)
    _lib.propagate_batch(ser, [0], [held[0]], np.zeros(5, dtype=int), th)
    r = _lib.rep_batch(th[0], u, u, y.size, 300, seed=1, mu=mu)
    assert all(np.isfinite(v).all() for v in r.values())  # incl. the device_ms scalar
    d, T = 3, 700
    A = 0.5 * np.eye(d)
    thd = np.concatenate([A.ravel(), 0.1 * rng.standard_normal(d * 2), rng.standard_normal(d), [0.1, 0.2],
                          np.eye(d).ravel(), [0.3], np.zeros(d), np.eye(d).ravel()])
    yd = rng.standard_normal(T)
    yd[rng.uniform(size=T) < 0.2] = np.nan
    ud = rng.standard_normal((2, T))
    a = _lib.smoother_d(d, yd, ud, ud, thd, method=0)
    b = _lib.smoother_d(d, yd, ud, ud, thd, method=1, chunk=16)
    assert np.allclose(a["lik"], b["lik"], rtol=1e-9) and np.allclose(a["X"], b["X"], atol=1e-9)
    print("single-step, replicates, scan ok", flush=True)


if __name__ == "__main__":
    main()
