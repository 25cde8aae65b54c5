"""Synthetic/bundled workloads of BASELINE.json's configs, shaped for ldsr_em_batch.

All inputs are deterministic (numpy default_rng(20261018 + offset)); initial values follow
make_init's distribution and draw order (R/LDS_reconstruction.R:14-30).  The NP data come from
tests/golden/np.json (the reference's data/NPannual.rda + data/NPpc.rda decoded by
tools/rda_to_golden.py) -- /root/reference is never read at run time.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED = 20261018


def _np_data():
    with open(os.path.join(ROOT, "tests", "golden", "np.json")) as f:
        return json.load(f)


def make_init_array(rng, p, q, n):
    """n flat thetas [A, B(p), C, D(q), Q=1, R=1, mu1=0, V1=1] in make_init's draw order (per restart:
    runif A, runif(p,-1,1) B, runif C, runif(q,-1,1) D).  One vectorised draw: numpy's Generator takes
    one 64-bit word per double whether asked for scalars or an array, and uniform(a, b) is
    a + (b - a) * random(), so this is bit-identical to drawing restart by restart."""
    r = rng.random((n, p + q + 2))
    out = np.empty((n, p + q + 6))
    out[:, 0] = r[:, 0]
    out[:, 1:1 + p] = -1.0 + 2.0 * r[:, 1:1 + p]
    out[:, 1 + p] = r[:, 1 + p]
    out[:, 2 + p:2 + p + q] = -1.0 + 2.0 * r[:, 2 + p:2 + p + q]
    out[:, 2 + p + q:] = (1.0, 1.0, 0.0, 1.0)
    return out


def np_series(first_row=401, start_year=1600):
    """y,u,v of LDS_reconstruction(NPannual, u=v=t(NPpc[first_row:813]), start.year) --
    T=413 for the defaults (config 1/2).  Returns (series dict, mu, inst)."""
    d = _np_data()
    pc = np.array([d["NPpc"][k] for k in ("PC1", "PC9", "PC13")])
    u = np.ascontiguousarray(pc[:, first_row - 1:])
    T = u.shape[1]
    years = np.arange(start_year, start_year + T)
    obs = np.log(np.array(d["NPannual"]["Qa"]))
    mu = float(obs.mean())
    inst = np.nonzero(np.isin(years, np.array(d["NPannual"]["year"])))[0]
    y = np.full(T, np.nan)
    y[inst] = obs - mu
    return dict(y=y, u=u, v=u, p=3, q=3), mu, inst


def np_restarts(n_restarts=100, seed=SEED):
    """Config 1: LDS_reconstruction(NPannual, NPpc, NPpc, start.year=1600, num.restarts=100):
    1 group x n_restarts fits."""
    s, mu, inst = np_series()
    rng = np.random.default_rng(seed)
    th0 = make_init_array(rng, 3, 3, n_restarts)
    return dict(name="LDS_reconstruction NPannual/NPpc T=413 %d restarts" % n_restarts, series=[s],
                group_series=np.zeros(1, dtype=np.int32), held=[np.zeros(0, dtype=np.int32)],
                fit_group=np.zeros(n_restarts, dtype=np.int32), theta0=th0, mu=mu, inst=inst)


def np_cv(n_folds=100, n_restarts=100, seed=SEED):
    """Config 2: cvLDS on NPannual/NPpc, n_folds make_Z(contiguous=FALSE, frac=0.25) hold-out folds
    (11 of the 46 observations each) x n_restarts restarts."""
    s, mu, inst = np_series()
    rng = np.random.default_rng(seed)
    k = int(np.floor(inst.size * 0.25))
    held = [np.sort(rng.choice(inst, k, replace=False)).astype(np.int32) for _ in range(n_folds)]
    th0 = make_init_array(rng, 3, 3, n_folds * n_restarts)
    return dict(name="cvLDS NPannual/NPpc T=413 %d folds x %d restarts" % (n_folds, n_restarts), series=[s],
                group_series=np.zeros(n_folds, dtype=np.int32), held=held,
                fit_group=np.repeat(np.arange(n_folds, dtype=np.int32), n_restarts), theta0=th0, mu=mu, inst=inst)


def synthetic_stations(n_stations=48, T=400, p=10, n_folds=100, n_restarts=100, seed=SEED):
    """Config 3: stations with u == v ~ N(0, diag(4/i)), instrumental period = last
    60..85 steps, y simulated from a random stable LDS and centred; n_folds of 25 % held out."""
    series, group_series, held, fit_group, th0 = [], [], [], [], []
    for st in range(n_stations):
        rng = np.random.default_rng(seed + 1 + st)
        lam = 4.0 / np.arange(1, p + 1)
        u = rng.standard_normal((p, T)) * np.sqrt(lam)[:, None]
        A, Cc = rng.uniform(0.3, 0.9), rng.uniform(0.01, 0.1)
        B = rng.uniform(-1, 1, p) / np.sqrt(p) * 0.1
        D = rng.uniform(-1, 1, p) / np.sqrt(p) * 0.1
        x = 0.0
        yfull = np.empty(T)
        for t in range(T):
            yfull[t] = Cc * x + D @ u[:, t] + rng.standard_normal() * 0.1
            x = A * x + B @ u[:, t] + rng.standard_normal()
        n_inst = int(rng.integers(60, 86))
        inst = np.arange(T - n_inst, T)
        y = np.full(T, np.nan)
        y[inst] = yfull[inst] - yfull[inst].mean()
        series.append(dict(y=y, u=u, v=u, p=p, q=p))
        k = int(np.floor(n_inst * 0.25))
        for _ in range(n_folds):
            g = len(group_series)
            group_series.append(st)
            held.append(np.sort(rng.choice(inst, k, replace=False)).astype(np.int32))
            th0.append(make_init_array(rng, p, p, n_restarts))
            fit_group.append(np.full(n_restarts, g, dtype=np.int32))
    return dict(name="synthetic %d stations T=%d p=q=%d %d folds x %d restarts" % (n_stations, T, p, n_folds,
                                                                                   n_restarts),
                series=series, group_series=np.asarray(group_series, dtype=np.int32), held=held,
                fit_group=np.concatenate(fit_group), theta0=np.concatenate(th0))


def subset(w, n_groups):
    """The first n_groups groups of a workload (bounded CPU sample of the same workload)."""
    n_groups = min(n_groups, len(w["group_series"]))
    keep = w["fit_group"] < n_groups
    out = dict(w)
    out.update(group_series=w["group_series"][:n_groups], held=w["held"][:n_groups],
               fit_group=w["fit_group"][keep], theta0=w["theta0"][keep])
    return out


def take_groups(w, groups):
    """The workload restricted to `groups` (sorted group ids): one rank's shard of a job.  Fits keep
    their order; `fits` in the result are their indices in the full job."""
    groups = np.asarray(groups, dtype=np.int64)
    keep = np.nonzero(np.isin(w["fit_group"], groups))[0]
    out = dict(w)
    out.update(group_series=w["group_series"][groups], held=[w["held"][g] for g in groups],
               fit_group=np.searchsorted(groups, w["fit_group"][keep]).astype(np.int32),
               theta0=w["theta0"][keep], fits=keep, groups=groups)
    return out


def flops_per_iter(T, p, q, n_obs):
    """F_iter = T(6p+23) + n_obs(4q+21) + S(p+1) + S(q+1), S(n) = 2/3 n^3 + 2 n^2."""
    S = lambda n: 2.0 / 3.0 * n ** 3 + 2.0 * n ** 2
    return T * (6 * p + 23) + n_obs * (4 * q + 21) + S(p + 1) + S(q + 1)


def algorithmic_flops(w, iters):
    """Sum over fits of iters*F_iter + T(2q+1) (final Ys), SURVEY.md 8d."""
    tot = 0.0
    gs = w["group_series"]
    for g in range(len(gs)):
        s = w["series"][gs[g]]
        T = s["y"].size
        n_obs = int(np.isfinite(s["y"]).sum()) - len(np.unique(w["held"][g]))
        f = flops_per_iter(T, s["p"], s["q"], n_obs)
        it = iters[w["fit_group"] == g] if len(gs) < 2000 else None
        if it is None:
            raise ValueError("use algorithmic_flops_fast for many groups")
        tot += float(it.sum()) * f + it.size * T * (2 * s["q"] + 1)
    return tot
