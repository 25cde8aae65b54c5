"""Pins the CPU oracle (oracle/ldsr_oracle.c) to every known-answer value the reference holds
for the EM hot path: tests/testthat/test-LDS-EM.R:21-41 and the NPlds fixture (R/sysdata.rda)."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import data


def rel(a, b):
    return abs(a - b) / abs(b)


def test_first_two_iterations_p1():
    # test-LDS-EM.R:21-35 ; testthat tolerance 1e-6 is relative
    y, u, th0, kat = data.p1_case()
    tol = kat["tolerance"]
    s1 = O.kalman_smoother(y, u, u, th0)
    th1 = O.mstep(y, u, u, s1)
    s2 = O.kalman_smoother(y, u, u, th1)
    th2 = O.mstep(y, u, u, s2)
    assert rel(s1["lik"], kat["smooth1_lik"]) < tol
    assert rel(s1["X"][0], kat["smooth1_X_1_85"][0]) < tol
    assert rel(s1["X"][84], kat["smooth1_X_1_85"][1]) < tol
    for th, g in ((th1, kat["theta1"]), (th2, kat["theta2"])):
        d = O.theta_split(th, 7, 7)
        assert rel(d["A"], g["A"]) < tol
        assert abs(d["C"] - g["C"]) < 1e-6  # printed to 6 decimals; C is ~1e-2
        assert rel(d["Q"], g["Q"]) < tol
    assert rel(s2["lik"], kat["smooth2_lik"]) < 5e-6  # golden printed with 6 significant digits


def test_convergence_p1():
    # test-LDS-EM.R:37-41
    y, u, th0, kat = data.p1_case()
    fit = O.em(y, u, u, th0, kat["em"]["niter"], kat["em"]["tol"])
    assert len(fit["liks"]) == kat["em"]["n_liks"] == 68
    assert abs(fit["lik"] - kat["em"]["lik"]) < 1e-6
    assert fit["lik"] == fit["liks"][-1] == fit["fit"]["lik"]
    assert np.all(np.diff(fit["liks"]) > -1e-12)  # EM is monotone


def test_nplds_fixture_estep():
    # NPlds = LDS_reconstruction(NPannual, t(NPpc), t(NPpc), start.year = 1200): its stored theta,
    # run through ONE E-step, must reproduce the stored lik / X / Q (T=813, 767 missing steps).
    g = data.load("nplds.json")
    y, u, mu, inst = data.np_case(1, 1200)
    assert y.size == 813 and inst[0] == 760 and inst.size == 46
    th = data.theta_of(g["theta"])
    s = O.kalman_smoother(y, u, u, th)
    assert rel(s["lik"], g["lik"]) < 1e-9
    assert np.max(np.abs(s["X"] - np.array(g["rec"]["X"]))) < 1e-12
    Q = np.exp(s["Y"] + mu)
    assert np.max(np.abs(Q / np.array(g["rec"]["Q"]) - 1)) < 1e-12
    # CI of X is 1.96*sqrt(V)  (R/LDS_reconstruction.R:197) -> pins the smoothed variance too
    Xl = s["X"] - 1.96 * np.sqrt(s["V"])
    assert np.max(np.abs(Xl - np.array(g["rec"]["Xl"]))) < 1e-11


def test_sentinels_and_unequal_dims():
    # test-LDS-EM.R:58-77: u=NULL, v=NULL, nrow(u) != nrow(v) must run; B/D come back zero
    y, u, mu, inst = data.np_case(601, 1800)
    rng = np.random.default_rng(1)
    th = np.concatenate([[0.5], rng.uniform(-1, 1, 1), [0.5], rng.uniform(-1, 1, 3), [1, 1, 0, 1]])
    f = O.em(y, None, u, th, 50, 1e-5, p=1, q=3)
    assert np.all(O.theta_split(f["theta"], 1, 3)["B"] == 0) and np.isfinite(f["lik"])
    th = np.concatenate([[0.5], rng.uniform(-1, 1, 3), [0.5], rng.uniform(-1, 1, 1), [1, 1, 0, 1]])
    f = O.em(y, u, None, th, 50, 1e-5, p=3, q=1)
    assert np.all(O.theta_split(f["theta"], 3, 1)["D"] == 0) and np.isfinite(f["lik"])
    th = np.concatenate([[0.5], rng.uniform(-1, 1, 2), [0.5], rng.uniform(-1, 1, 3), [1, 1, 0, 1]])
    f = O.em(y, u[:2], u, th, 50, 1e-5)
    assert np.isfinite(f["lik"]) and len(f["liks"]) >= 3


def test_selection_rule():
    # R/LDS_reconstruction.R:50-58
    assert O.select([1.0, 3.0, 2.0], [1.0, -1.0, 1.0]) == 2      # best among C>0, not global best
    assert O.select([1.0, 3.0, 2.0], [-1.0, -1.0, -1.0]) == 1    # no C>0: which.max
    assert O.select([np.nan, 1.0, 1.0], [1.0, 1.0, 1.0]) == 1    # NaN skipped, first of a tie
    assert O.select([np.nan, np.nan], [1.0, 1.0]) == -1
    assert O.select([2.0, np.nan, 1.0], [0.0, np.nan, 0.0]) == 0  # C == 0 is not > 0


def test_em_batch_matches_single_fits():
    y, u, mu, inst = data.np_case(601, 1800)
    rng = np.random.default_rng(7)
    n = 6
    th0 = np.stack([np.concatenate([[rng.uniform()], rng.uniform(-1, 1, 3), [rng.uniform()],
                                    rng.uniform(-1, 1, 3), [1, 1, 0, 1]]) for _ in range(n)])
    held = [inst[[0, 5, 9]], inst[[20, 21, 22, 23]]]
    r = O.em_batch([dict(y=y, u=u, v=u)], [0, 0], held, [0, 0, 0, 1, 1, 1], th0, 60, 1e-5, n_threads=2)
    for f in range(n):
        yy = y.copy()
        yy[held[f // 3]] = np.nan
        s = O.em(yy, u, u, th0[f], 60, 1e-5)
        assert s["lik"] == r["lik"][f] and len(s["liks"]) == r["iters"][f]
        assert np.array_equal(s["theta"], r["theta"][f])
    for g in range(2):
        sl = slice(3 * g, 3 * g + 3)
        assert r["best"][g] == 3 * g + O.select(r["lik"][sl], r["theta"][sl, 4])


def test_propagate_and_rep_definitions():
    # no golden in the reference ("parity unpinned"): check against the defining recursions
    d = data.load("np.json")
    th = data.theta_of(d["theta"])
    y, u, mu, inst = data.np_case(1, 1200)
    pr = O.propagate(th, u, u, y)
    t = O.theta_split(th, 3, 3)
    X = np.zeros(813)
    V = np.zeros(813)
    X[0], V[0] = t["mu1"], t["V1"]
    for k in range(1, 813):
        X[k] = t["A"] * X[k - 1] + t["B"] @ u[:, k - 1]
        V[k] = t["A"] ** 2 * V[k - 1] + t["Q"]
    assert np.allclose(pr["X"], X, rtol=1e-13) and np.allclose(pr["V"], V, rtol=1e-13)
    assert np.allclose(pr["Y"], t["C"] * X + t["D"] @ u, rtol=1e-13, atol=1e-15)
    z = np.random.default_rng(3).standard_normal(1 + 2 * 813)
    r = O.rep(th, u, u, 813, z, mu=mu)
    x = z[0] * np.sqrt(t["V1"])
    for k in range(813):
        assert abs(r["simX"][k] - x) <= 1e-12 * max(1, abs(x))
        x = t["A"] * x + t["B"] @ u[:, k] + z[1 + k] * np.sqrt(t["Q"])
    assert np.allclose(r["simQ"], np.exp(r["simY"] + mu))


def test_general_d_reduces_to_1d():
    """oracle/ldsr_oracle_d.c (general state dimension; no reference counterpart, src/EM.cpp:20 is
    scalar-state) at d = 1 must reproduce the pinned 1-D oracle: on the NPlds fixture (T = 813, 767
    missing steps) and on the fully observed P1 series."""
    g = data.load("nplds.json")
    y, u, mu, inst = data.np_case(1, 1200)
    th = data.theta_of(g["theta"])
    cases = [(y, u, u, th, 3, 3)]
    y1, u1, th1, _ = data.p1_case()
    cases.append((y1, u1, u1, th1, 7, 7))
    for yy, uu, vv, t, p, q in cases:
        o1 = O.kalman_smoother(yy, uu, vv, t)
        thd = O.theta_d_flat([[t[0]]], [t[1:1 + p]], [t[1 + p]], t[2 + p:2 + p + q], [[t[2 + p + q]]], t[3 + p + q],
                             [t[4 + p + q]], [[t[5 + p + q]]])
        od = O.smoother_d(1, yy, uu, vv, thd)
        assert abs(od["lik"] - o1["lik"]) < 1e-13 * max(1.0, abs(o1["lik"]))
        assert np.max(np.abs(od["X"][:, 0] - o1["X"])) < 1e-12
        assert np.max(np.abs(od["V"][:, 0, 0] - o1["V"])) < 1e-12
        assert np.max(np.abs(od["Y"] - o1["Y"])) < 1e-12
    assert abs(od["lik"] - (-11.678657)) < 1e-6  # tests/testthat/test-LDS-EM.R:26 through the d = 1 path


def test_cv_metrics_match_the_stored_cvLDS_result():
    """The reference's own stored cvLDS output (R/sysdata.rda::NPcv) holds, for 30 folds of 12 points,
    the cross-validated flows Ycv, the folds Z and the five skill metrics of every fold: a golden
    vector for calculate_metrics (R/utils.R:56-70) + src/utils.cpp:13-97."""
    g = data.load("npcv.json")
    n = len(g["target"]["y"])
    Y = np.array(g["Ycv"]["Y"]).reshape(-1, n)
    assert Y.shape[0] == len(g["Z"]) == 30
    m = O.cv_metrics(Y, np.array(g["target"]["y"]), g["Z"])
    for j, name in enumerate(O.METRIC_NAMES):
        want = np.array(g["metrics_dist"][name])
        assert np.allclose(m[:, j], want, rtol=1e-10, atol=1e-12), (name, np.max(np.abs(m[:, j] - want)))
    # exp_trans path: log-space input gives the same numbers
    m2 = O.cv_metrics(np.log(Y), np.array(g["target"]["y"]), g["Z"], exp_trans=True)
    assert np.allclose(m2, m, rtol=1e-12)


def test_construct_rec_reproduces_the_stored_reconstruction():
    """All six columns of the reference's stored LDS_reconstruction result (R/sysdata.rda::NPlds$rec:
    X, Xl, Xu, Q, Ql, Qu over T = 813) from the E-step with the stored theta: pins construct_rec
    (R/LDS_reconstruction.R:190-212) incl. exp_ci's qlnorm quantiles (R/utils.R:122-125)."""
    g = data.load("nplds.json")
    y, u, mu, inst = data.np_case(1, 1200)
    th = data.theta_of(g["theta"])
    s = O.kalman_smoother(y, u, u, th)
    out, mean = O.construct_rec(s["X"], s["V"], s["Y"], th[4], th[9], mu, "log")
    for j, name in enumerate(O.REC_COLUMNS):
        want = np.array(g["rec"][name])
        assert np.allclose(out[0, j], want, rtol=1e-11, atol=1e-12), (name, np.max(np.abs(out[0, j] - want)))
    assert np.allclose(mean[0], out[0, 0]) and np.allclose(mean[1], out[0, 3])
