"""Python mirror of the reference's R API for the EM path -- same function names, argument
meaning, defaults and error behaviour -- on top of the CUDA library (ldsr_b200/_lib.py).

R is not installed in the build image, so this module plays the role the R wrappers in
ldsr_b200/r/ play for an R user; the parity tests are written against it
so that they read like tests/testthat/test-LDS-EM.R.

Conventions carried over from R:
  * u, v are [p, T] / [q, T] arrays (R matrices), or None (R: NULL -> `matrix(0)` sentinel,
    R/LDS_reconstruction.R:131-136); an ensemble is a list of such arrays.
  * theta is a dict with keys A, B, C, D, Q, R, mu1, V1 (R: list of 1x1 / 1xp matrices).
  * y is a length-T vector with NaN for missing (R: NA).
  * Random numbers: the reference draws initial values from R's global RNG (make_init) inside
    each worker.  Here the caller passes a numpy Generator; draws are made in the reference's
    order (fold-major, then restart, then A, B[1..p], C, D[1..q]) -- see LDS_EM_restart/cvLDS.
Every numeric result comes from the GPU; there is no CPU path in this module.
"""
import numpy as np

from . import _lib

_Z05 = 1.6448536269514722  # qnorm(0.95)


# ---------------------------------------------------------------------------------------------
# theta helpers
# ---------------------------------------------------------------------------------------------
def theta_to_vec(theta):
    """R/LDS_GA.R:6-16 order: [A, B, C, D, Q, R, mu1, V1]."""
    return np.concatenate([np.ravel(theta[k]) for k in ("A", "B", "C", "D", "Q", "R", "mu1", "V1")]).astype(float)


def vec_to_theta(vec, p, q):
    vec = np.asarray(vec, dtype=float)
    return dict(A=float(vec[0]), B=vec[1:1 + p].copy(), C=float(vec[1 + p]), D=vec[2 + p:2 + p + q].copy(),
                Q=float(vec[2 + p + q]), R=float(vec[3 + p + q]), mu1=float(vec[4 + p + q]),
                V1=float(vec[5 + p + q]))


def _dims(u, v):
    """nrow(u), nrow(v) with the matrix(0) sentinel counting as one row."""
    p = 1 if u is None else np.atleast_2d(u).shape[0]
    q = 1 if v is None else np.atleast_2d(v).shape[0]
    return p, q


def make_init(p, q, num_restarts, rng):
    """R/LDS_reconstruction.R:14-30: A~U(0,1), B~U(-1,1)^p, C~U(0,1), D~U(-1,1)^q, Q=R=V1=1, mu1=0,
    drawn per restart in that order."""
    out = []
    for _ in range(num_restarts):
        A = rng.uniform()
        B = rng.uniform(-1, 1, p)
        Cc = rng.uniform()
        D = rng.uniform(-1, 1, q)
        out.append(dict(A=A, B=B, C=Cc, D=D, Q=1.0, R=1.0, mu1=0.0, V1=1.0))
    return out


def make_Z(obs, nRuns=30, frac=0.1, contiguous=True, rng=None):
    """R/utils.R:83-101.  Returns 1-based index vectors like R (indices into obs)."""
    rng = rng or np.random.default_rng()
    obs = np.asarray(obs, dtype=float)
    obsInd = np.nonzero(~np.isnan(obs))[0] + 1
    if frac == 1:
        return [np.array([i]) for i in obsInd]
    n = obsInd.size
    k = int(np.floor(n * frac))
    if contiguous:
        maxInd = n - k
        if maxInd < nRuns:  # the reference's fallback (utils.R:92-95); negative k is its bug, we refuse
            maxInd = nRuns
            k = n - nRuns
            if k < 0 or maxInd + k > n:
                raise ValueError("make_Z: cannot make %d contiguous folds out of %d points "
                                 "(the reference misbehaves here, R/utils.R:92-95)" % (nRuns, n))
        starts = np.sort(rng.choice(np.arange(1, maxInd + 1), nRuns, replace=False))
        return [obsInd[x - 1:x + k] for x in starts]  # obsInd[x:(x+k)]: k+1 points
    return [np.sort(rng.choice(obsInd, k, replace=False)) for _ in range(nRuns)]


def _series(y, u, v):
    p, q = _dims(u, v)
    return dict(y=np.asarray(y, dtype=float).ravel(), u=None if u is None else np.atleast_2d(u),
                v=None if v is None else np.atleast_2d(v), p=p, q=q)


def _fit_dict(r, row, T):
    sl = slice(row, row + T)
    return dict(X=r["X"][sl].copy(), Y=r["Y"][sl].copy(), V=r["V"][sl].copy(), J=r["J"][sl].copy())


# ---------------------------------------------------------------------------------------------
# The reference's four .Call functions (R/RcppExports.R:15-59)
# ---------------------------------------------------------------------------------------------
def Kalman_smoother(y, u, v, theta, stdlik=True):
    s = _series(y, u, v)
    r = _lib.smoother_batch([s], [0], None, [0], theta_to_vec(theta)[None, :], stdlik)
    return dict(X=r["X"][0], Y=r["Y"][0], V=r["V"][0], J=r["J"][0], lik=float(r["lik"][0]))


def Mstep(y, u, v, fit):
    s = _series(y, u, v)
    r = _lib.mstep_batch([s], [0], None, [0], [fit["X"]], [fit["V"]], [fit["J"]], s["p"] + s["q"] + 6)
    if r["status"][0] == _lib.FIT_SINGULAR:
        raise _lib.LdsrError(_lib.ERR_ARG, "Mstep: inv(): matrix is singular")
    return vec_to_theta(r["theta"][0], s["p"], s["q"])


def LDS_EM(y, u, v, theta0, niter=1000, tol=1e-5):
    s = _series(y, u, v)
    r = _lib.em_batch([s], [0], None, [0], theta_to_vec(theta0)[None, :], niter, tol, want_liks=True)
    if r["status"][0] == _lib.FIT_SINGULAR:
        raise _lib.LdsrError(_lib.ERR_ARG, "LDS_EM: inv(): matrix is singular")
    n = int(r["iters"][0])
    fit = _fit_dict(r, 0, s["y"].size)
    fit["lik"] = float(r["lik"][0])
    return dict(theta=vec_to_theta(r["theta"][0], s["p"], s["q"]), fit=fit, liks=r["liks"][0, :n].copy(),
                lik=float(r["lik"][0]))


def propagate(theta, u, v, y, stdlik=True):
    s = _series(y, u, v)
    r = _lib.propagate_batch([s], [0], None, [0], theta_to_vec(theta)[None, :], stdlik)
    return dict(X=r["X"][0], Y=r["Y"][0], V=r["V"][0], lik=float(r["lik"][0]))


# ---------------------------------------------------------------------------------------------
# Restart / cross-validation fan-out: ONE batched GPU call instead of foreach %dopar%
# ---------------------------------------------------------------------------------------------
def LDS_EM_restart(y, u, v, init, niter=1000, tol=1e-5, return_init=True, n_devices=1):
    """R/LDS_reconstruction.R:42-62: all restarts as one batch, selection on the device."""
    s = _series(y, u, v)
    th0 = np.stack([theta_to_vec(t) for t in init])
    r = _lib.em_batch([s], [0], None, np.zeros(len(init), dtype=np.int32), th0, niter, tol,
                      n_devices=n_devices)
    if np.any(r["status"] == _lib.FIT_SINGULAR):
        raise _lib.LdsrError(_lib.ERR_ARG, "LDS_EM_restart: inv(): matrix is singular")
    b = int(r["best"][0])
    if b < 0:
        raise _lib.LdsrError(_lib.ERR_ARG, "LDS_EM_restart: no restart produced a finite likelihood")
    fit = _fit_dict(r, 0, s["y"].size)
    fit["lik"] = float(r["lik"][b])
    ans = dict(theta=vec_to_theta(r["theta"][b], s["p"], s["q"]), fit=fit, lik=float(r["lik"][b]),
               iters=int(r["iters"][b]), all_lik=r["lik"], all_iters=r["iters"], best=b)
    if return_init:
        ans["init"] = init[b]
    return ans


def _prep_uv(u, v):
    """Argument checks of LDS_reconstruction / cvLDS (R/LDS_reconstruction.R:128-157, :315-340).
    Returns (single, list of u, list of v, N)."""
    single = not isinstance(u, (list, tuple)) and not isinstance(v, (list, tuple))
    if single:
        if u is None and v is None:
            raise ValueError("u and v cannot both be NULL")
        if u is not None and v is not None and np.atleast_2d(u).shape[1] != np.atleast_2d(v).shape[1]:
            raise ValueError("u and v must have the same number of time steps.")
        N = np.atleast_2d(u if u is not None else v).shape[1]
        return True, [u], [v], N
    if u is None:
        u = [None] * len(v)
    if v is None:
        v = [None] * len(u)
    for a, b in zip(u, v):
        if a is not None and b is not None and np.atleast_2d(a).shape[1] != np.atleast_2d(b).shape[1]:
            raise ValueError("u and v must have the same number of time steps.")
    first = u[0] if u[0] is not None else v[0]
    return False, list(u), list(v), np.atleast_2d(first).shape[1]


def _boxcox_lambda(y):
    """MASS::boxcox(y ~ 1, plotit = FALSE): profile log-likelihood on lambda = seq(-2, 2, 0.1)."""
    y = np.asarray(y, dtype=float)
    n = y.size
    logy = np.log(y)
    ydot = np.exp(logy.mean())
    lam = np.round(np.arange(-2, 2.0001, 0.1), 10)
    ll = []
    for l in lam:
        yt = (np.log(y) * ydot) if abs(l) < 1e-12 else ((y / ydot) ** l - 1) / l * ydot
        rss = np.sum((yt - yt.mean()) ** 2)
        ll.append(-n / 2 * np.log(rss / n))
    return float(lam[int(np.argmax(ll))])


def _transform(Qa, transform):
    y = np.asarray(Qa, dtype=float)
    lam = None
    if transform == "log":
        y = np.log(y)
    elif transform == "boxcox":
        lam = _boxcox_lambda(y)
        if abs(lam) < 0.01:
            lam = 0.0
            y = np.log(y)
        else:
            y = (y ** lam - 1) / lam
    elif transform != "none":
        raise ValueError('Accepted transformations are "log", "boxcox" and "none" only.')
    return y, lam


def inv_boxcox(x, lam):
    return np.exp(x) if lam == 0 else (x * lam + 1) ** (1 / lam)


def _make_y(years_Qa, obs, start_year, N):
    """R/LDS_reconstruction.R:179-183: centre and NA-pad to the study horizon."""
    end_year = start_year + N - 1
    if end_year < years_Qa[-1]:
        raise ValueError("The last year of u is earlier than the last year of the instrumental period.")
    mu = float(np.nanmean(obs))
    y = np.concatenate([np.full(int(years_Qa[0] - start_year), np.nan), obs - mu,
                        np.full(int(end_year - years_Qa[-1]), np.nan)])
    years = np.arange(start_year, end_year + 1)
    return y, mu, years


def _construct_rec(fit, theta, mu, transform, lam, years):
    """R/LDS_reconstruction.R:190-212 through ldsr_construct_rec_batch (one member)."""
    out, _ = _lib.construct_rec(fit["X"], fit["V"], fit["Y"], float(np.ravel(theta["C"])[0]),
                                float(np.ravel(theta["R"])[0]), mu, transform, 0.0 if lam is None else float(lam))
    rec = dict(year=years)
    rec.update({k: out[0, j] for j, k in enumerate(_lib.REC_COLUMNS)})
    return rec


def LDS_reconstruction(Qa, u, v, start_year, method="EM", transform="log", init=None, num_restarts=50,
                       return_init=False, niter=1000, tol=1e-5, return_raw=False, rng=None, n_devices=1):
    """R/LDS_reconstruction.R:122-258.  Qa: dict(year=..., Qa=...).  Single model or ensemble
    (u, v lists); all members x restarts go to the GPU as one batch."""
    if method != "EM":
        raise ValueError("only method = 'EM' is implemented (GA/BFGS are experimental in the reference)")
    single, us, vs, N = _prep_uv(u, v)
    rng = rng or np.random.default_rng()
    if init is None:
        init = [make_init(*_dims(a, b), num_restarts, rng) for a, b in zip(us, vs)]
    elif single:
        init = [init]
    obs, lam = _transform(Qa["Qa"], transform)
    y, mu, years = _make_y(np.asarray(Qa["year"]), obs, start_year, N)

    series = [_series(y, a, b) for a, b in zip(us, vs)]
    stride = max(s["p"] + s["q"] + 6 for s in series)
    th0, fg = [], []
    for m, ini in enumerate(init):
        for t in ini:
            vec = theta_to_vec(t)
            th0.append(np.pad(vec, (0, stride - vec.size)))
            fg.append(m)
    r = _lib.em_batch(series, np.arange(len(series)), None, fg, np.stack(th0), niter, tol, n_devices=n_devices)
    if np.any(r["status"] == _lib.FIT_SINGULAR):
        raise _lib.LdsrError(_lib.ERR_ARG, "LDS_reconstruction: inv(): matrix is singular")
    members = []
    offs = np.cumsum([0] + [len(i) for i in init])
    for m, s in enumerate(series):
        b = int(r["best"][m])
        if b < 0:
            raise _lib.LdsrError(_lib.ERR_ARG, "no restart produced a finite likelihood")
        theta = vec_to_theta(r["theta"][b], s["p"], s["q"])
        fit = _fit_dict(r, int(r["traj_ptr"][m]), N)
        ans = dict(rec=_construct_rec(fit, theta, mu, transform, lam, years), theta=theta, lik=float(r["lik"][b]))
        if return_raw:
            raw = propagate(theta, s["u"], s["v"], y)
            ans["rec2"] = _construct_rec(raw, theta, mu, transform, lam, years)
        if return_init:
            ans["init"] = init[m][b - offs[m]]
        if transform == "boxcox":
            ans["lambda"] = lam
        members.append(ans)
    if single:
        return members[0]
    out = dict(rec=dict(year=years, X=np.mean([m["rec"]["X"] for m in members], axis=0),
                        Q=np.mean([m["rec"]["Q"] for m in members], axis=0)), ensemble=members)
    if return_raw:
        out["rec_raw"] = dict(year=years, X=np.mean([m["rec2"]["X"] for m in members], axis=0),
                              Q=np.mean([m["rec2"]["Q"] for m in members], axis=0))
    return out


def one_lds_cv(z, instPeriod, mu, y, u, v, method="EM", num_restarts=20, niter=1000, tol=1e-6, use_raw=False,
               rng=None):
    """R/LDS_reconstruction.R:270-285.  z: 1-based indices into the instrumental period;
    instPeriod: 1-based indices of the instrumental period in the whole record."""
    rng = rng or np.random.default_rng()
    y = np.array(y, dtype=float).ravel()
    y[np.asarray(instPeriod)[np.asarray(z) - 1] - 1] = np.nan
    init = make_init(*_dims(u, v), num_restarts, rng)
    res = LDS_EM_restart(y, u, v, init, niter, tol, return_init=False)
    idx = np.asarray(instPeriod) - 1
    if use_raw:
        return propagate(res["theta"], u, v, y)["Y"][idx] + mu
    return res["fit"]["Y"][idx] + mu


def _nse(sim, obs):
    return 1 - np.sum((sim - obs) ** 2) / np.sum((obs - obs.mean()) ** 2)


def calculate_metrics(sim, obs, z):
    """R/utils.R:56-70 with the metric definitions of src/utils.cpp:13-97.  z is 1-based."""
    sim, obs = np.asarray(sim, dtype=float), np.asarray(obs, dtype=float)
    zi = np.asarray(z) - 1
    keep = np.ones(obs.size, dtype=bool)
    keep[zi] = False
    tr_obs, tr_sim = obs[keep], sim[keep]
    ok = ~np.isnan(tr_obs)
    tr_obs, tr_sim = tr_obs[ok], tr_sim[ok]
    vs, vo = sim[zi], obs[zi]
    rmse = np.sqrt(np.mean((vs - vo) ** 2))
    r = np.corrcoef(vs, vo)[0, 1] if vs.size > 1 else np.nan
    alpha, beta = np.std(vs, ddof=1) / np.std(vo, ddof=1) if vs.size > 1 else np.nan, vs.mean() / vo.mean()
    return dict(R2=_nse(tr_sim, tr_obs),
                RE=1 - np.sum((vs - vo) ** 2) / np.sum((vo - tr_obs.mean()) ** 2),
                CE=_nse(vs, vo), nRMSE=rmse / np.nanmean(obs),
                KGE=1 - np.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2))


def tbrm(x, C=9.0):
    """Tukey's biweight robust mean, one step from the median, as dplR::tbrm (the reference's default
    aggregator of metrics.dist, R/LDS_reconstruction.R:397-398).  dplR is not part of the reference tree
    and is not installed here: this follows its published definition (Mosteller & Tukey 1977; weights
    (1 - u^2)^2 with u = (x - median)/(C * MAD + 1e-6), |u| < 1) -- unpinned.  The reference's stored NPcv
    result was aggregated with the plain mean (its `metrics` equal colMeans(metrics.dist))."""
    x = np.asarray(x, dtype=float)
    x = x[~np.isnan(x)]
    if x.size == 0:
        return float("nan")
    m = np.median(x)
    u = (x - m) / (C * np.median(np.abs(x - m)) + 1e-6)
    w = np.where(np.abs(u) < 1, (1 - u * u) ** 2, 0.0)
    return float(np.sum(w * x) / np.sum(w))


def cvLDS(Qa, u, v, start_year, method="EM", transform="log", num_restarts=50, Z=None, metric_space="original",
          use_raw=False, niter=1000, tol=1e-5, rng=None, n_devices=1, use_robust_mean=True):
    """R/LDS_reconstruction.R:308-409.  All folds x members x restarts run as ONE batch; initial
    values are drawn fold-major exactly as `foreach(z = Z) ... one_lds_cv` would under
    registerDoSEQ (:373-381, :275)."""
    if method != "EM":
        raise ValueError("only method = 'EM' is implemented")
    if use_raw:
        raise ValueError("use_raw is experimental in the reference ('don't use'); not implemented")
    single, us, vs, N = _prep_uv(u, v)
    rng = rng or np.random.default_rng()
    obs, lam = _transform(Qa["Qa"], transform)
    qyears = np.asarray(Qa["year"])
    if Z is None:
        Z = make_Z(Qa["Qa"], rng=rng)
    if not isinstance(Z, (list, tuple)):
        raise ValueError("Please provide the cross-validation folds (Z) in a list.")
    y, mu, years = _make_y(qyears, obs, start_year, N)
    inst = np.nonzero(np.isin(years, qyears))[0]  # 0-based instPeriod

    series = [_series(y, a, b) for a, b in zip(us, vs)]
    stride = max(s["p"] + s["q"] + 6 for s in series)
    group_series, held, fit_group, th0 = [], [], [], []
    for z in Z:  # fold-major, member-minor: the reference's nested foreach order
        for m, s in enumerate(series):
            g = len(group_series)
            group_series.append(m)
            held.append(inst[np.asarray(z) - 1])
            for t in make_init(s["p"], s["q"], num_restarts, rng):
                vec = theta_to_vec(t)
                th0.append(np.pad(vec, (0, stride - vec.size)))
                fit_group.append(g)
    r = _lib.em_batch(series, group_series, held, fit_group, np.stack(th0), niter, tol, n_devices=n_devices)
    if np.any(r["status"] == _lib.FIT_SINGULAR):
        raise _lib.LdsrError(_lib.ERR_ARG, "cvLDS: inv(): matrix is singular")
    nm = len(series)
    Ycv = []
    for k in range(len(Z)):
        cols = []
        for m in range(nm):
            g = k * nm + m
            if r["best"][g] < 0:
                raise _lib.LdsrError(_lib.ERR_ARG, "fold %d: no restart produced a finite likelihood" % k)
            row = int(r["traj_ptr"][g])
            cols.append(r["Y"][row + inst] + mu)  # fit$Y[instPeriod] + mu   (:283)
        Ycv.append(np.mean(cols, axis=0))        # .final = rowMeans          (:379)
    if metric_space == "original":
        if transform == "log":
            Ycv = [np.exp(a) for a in Ycv]
        elif transform == "boxcox":
            Ycv = [inv_boxcox(a, lam) for a in Ycv]
        target = np.asarray(Qa["Qa"], dtype=float)
    else:
        target = obs
    # mapply(calculate_metrics, ...) (:395) for all folds in one device call
    keys = _lib.METRIC_NAMES
    dist = _lib.cv_metrics(np.stack(Ycv), target, Z)
    metrics_dist = {k: dist[:, j].copy() for j, k in enumerate(keys)}
    mean_func = tbrm if use_robust_mean else np.mean  # R/LDS_reconstruction.R:397
    return dict(metrics_dist=metrics_dist, metrics={k: float(mean_func(metrics_dist[k])) for k in keys},
                target=dict(year=qyears, y=target), Ycv=np.stack(Ycv, axis=1), Z=Z,
                best=r["best"], lik=r["lik"], iters=r["iters"])


# ---------------------------------------------------------------------------------------------
# Stochastic replicates (R/stochastics.R)
# ---------------------------------------------------------------------------------------------
def LDS_rep(theta, u=None, v=None, years=None, num_reps=100, mu=0.0, exp_trans=True, seed=0, z=None, r_seed=None):
    """R/stochastics.R:58-63.  Returns long-format columns like the reference's data.table
    (year, simX, simY, simQ, rep).  Noise: r_seed = s gives the replicates of `set.seed(s); LDS_rep(...)`
    in R (R's own stream, generated on the device); z (optional): [num_reps, 1+2n] standard normals in
    the reference's draw order; otherwise the counter-based device generator keyed by `seed`."""
    n = len(years)
    if u is None:  # is.null(u): both input terms are dropped (stochastics.R:28-32)
        v = None
    p, q = (0 if u is None else np.atleast_2d(u).shape[0]), (0 if v is None else np.atleast_2d(v).shape[0])
    th = np.concatenate([[theta["A"]], np.ravel(theta["B"])[:p], [theta["C"]], np.ravel(theta["D"])[:q],
                         [theta["Q"], theta["R"], theta["mu1"], theta["V1"]]]).astype(float)
    r = _lib.rep_batch(th, u, v, n, num_reps, z=z, seed=seed, mu=mu, exp_trans=exp_trans, p=p, q=q, r_seed=r_seed)
    return dict(year=np.tile(np.asarray(years), num_reps), simX=r["simX"].ravel(), simY=r["simY"].ravel(),
                simQ=r["simQ"].ravel(), rep=np.repeat(np.arange(1, num_reps + 1), n))


def one_LDS_rep(rep_num, theta, u=None, v=None, years=None, mu=0.0, exp_trans=True, seed=0, z=None):
    """R/stochastics.R:18-47."""
    out = LDS_rep(theta, u, v, years, 1, mu, exp_trans, seed=seed + rep_num, z=None if z is None else z[None, :])
    out["rep"][:] = rep_num
    return out


def Kalman_smoother_d(y, u, v, theta, stdlik=True, method=1):
    """Kalman / RTS smoother for a state of dimension d = nrow(theta$A) in 2..4 (d = 1 works too).
    No counterpart in the reference, whose state is scalar (src/EM.cpp:20); same conventions as
    Kalman_smoother (src/EM.cpp:22-131).  theta: dict(A [d,d], B [d,p], C [d], D [q], Q [d,d], R, mu1 [d],
    V1 [d,d]).  method 1 = associative scan over time, 0 = sequential.  Returns dict(X [d,T], Y [T],
    V [T,d,d], lik) -- mirrors ldsr_b200.R::Kalman_smoother_d."""
    A = np.atleast_2d(np.asarray(theta["A"], dtype=float))
    d = A.shape[0]
    blocks = [A.ravel()]
    if u is not None:
        blocks.append(np.asarray(theta["B"], dtype=float).reshape(d, -1).ravel())
    blocks.append(np.asarray(theta["C"], dtype=float).ravel())
    if v is not None:
        blocks.append(np.asarray(theta["D"], dtype=float).ravel())
    blocks += [np.asarray(theta["Q"], dtype=float).reshape(d, d).ravel(), [float(np.ravel(theta["R"])[0])],
               np.asarray(theta["mu1"], dtype=float).ravel(), np.asarray(theta["V1"], dtype=float).reshape(d, d).ravel()]
    r = _lib.smoother_d(d, np.ravel(y), u, v, np.concatenate(blocks), stdlik=stdlik, method=method)
    return dict(X=r["X"][0].T, Y=r["Y"][0], V=r["V"][0], lik=float(r["lik"][0]))
