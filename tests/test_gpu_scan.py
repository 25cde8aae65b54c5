"""GPU tests of the general-state-dimension smoother (BASELINE.json config 5; beyond the reference,
which is scalar-state only: src/EM.cpp:20).  The CUDA sequential recursion and the associative scan
are compared with oracle/ldsr_oracle_d.c, which at d = 1 reproduces the pinned 1-D oracle
(tests/test_oracle_golden.py::test_general_d_reduces_to_1d)."""
import numpy as np
import pytest

from ldsr_b200 import _lib
from oracle import oracle as O
from tests import data

pytestmark = pytest.mark.gpu


def random_model(rng, d, p, q, T, miss=0.3, lead=0):
    """A stable random LDS with scalar output, simulated; returns y, u, v, theta_flat."""
    M = rng.standard_normal((d, d))
    A = 0.9 * M / max(1e-9, np.max(np.abs(np.linalg.eigvals(M))))
    B = 0.3 * rng.standard_normal((d, p))
    Cc = rng.standard_normal(d)
    D = 0.3 * rng.standard_normal(q)
    Lq = 0.4 * rng.standard_normal((d, d))
    Q = Lq @ Lq.T + 0.1 * np.eye(d)
    R = 0.3
    mu1 = 0.2 * rng.standard_normal(d)
    Lv = 0.5 * rng.standard_normal((d, d))
    V1 = Lv @ Lv.T + 0.5 * np.eye(d)
    u = rng.standard_normal((p, T))
    v = rng.standard_normal((q, T))
    x = mu1 + np.linalg.cholesky(V1) @ rng.standard_normal(d)
    y = np.empty(T)
    cq = np.linalg.cholesky(Q)
    for t in range(T):
        if t > 0:
            x = A @ x + B @ u[:, t - 1] + cq @ rng.standard_normal(d)
        y[t] = Cc @ x + D @ v[:, t] + np.sqrt(R) * rng.standard_normal()
    y[rng.uniform(size=T) < miss] = np.nan
    y[:lead] = np.nan
    return y, u, v, O.theta_d_flat(A, B, Cc, D, Q, R, mu1, V1)


@pytest.mark.parametrize("d", [1, 2, 3, 4])
# (2600, 4), (4133, 5), (6000, 0): 512 chunks and more -- the scan stages run as 32 groups per fit in three launches
# (group totals, one-warp scan of the totals, apply), with ragged last groups; 6000 with the automatic chunk length
@pytest.mark.parametrize("T,chunk", [(2, 0), (37, 8), (200, 0), (513, 16), (1000, 7), (2600, 4), (4133, 5), (6000, 0)])
def test_scan_and_sequential_match_oracle(d, T, chunk):
    rng = np.random.default_rng(1000 * d + T)
    p, q = 5, 4
    y, u, v, th = random_model(rng, d, p, q, T, miss=0.3, lead=min(T // 4, 20))
    if not np.isfinite(y).any():
        y[-1] = 0.1
    o = O.smoother_d(d, y, u, v, th)
    for method in (0, 1):
        g = _lib.smoother_d(d, y, u, v, th, method=method, chunk=chunk)
        assert abs(g["lik"][0] - o["lik"]) < 1e-10 * max(1.0, abs(o["lik"])), (method, g["lik"][0], o["lik"])
        assert np.max(np.abs(g["X"][0] - o["X"])) < 1e-9
        assert np.max(np.abs(g["V"][0] - o["V"])) < 1e-9
        assert np.max(np.abs(g["Y"][0] - o["Y"])) < 1e-9


def test_d1_matches_the_reference_golden_through_the_scan():
    # tests/testthat/test-LDS-EM.R:26-27 (lik and X[1], X[85] of the first E-step) via d = 1, both methods
    y, u, th0, kat = data.p1_case()
    p = q = 7
    thd = O.theta_d_flat([[th0[0]]], [th0[1:1 + p]], [th0[1 + p]], th0[2 + p:2 + p + q], [[th0[2 + p + q]]],
                         th0[3 + p + q], [th0[4 + p + q]], [[th0[5 + p + q]]])
    for method in (0, 1):
        g = _lib.smoother_d(1, y, u, u, thd, method=method, chunk=8)
        assert abs(g["lik"][0] - kat["smooth1_lik"]) < 1e-6
        assert abs(g["X"][0, 0, 0] - kat["smooth1_X_1_85"][0]) < 1e-6
        assert abs(g["X"][0, 84, 0] - kat["smooth1_X_1_85"][1]) < 1e-6


def test_no_inputs_and_several_fits():
    rng = np.random.default_rng(7)
    d, T = 3, 300
    ths, ys = [], None
    for k in range(5):
        y, u, v, th = random_model(rng, d, 2, 2, T)
        ys = y if ys is None else ys
        ths.append(th)
    ths = np.stack(ths)
    # same y/u/v for all fits (the API is one series, many parameter sets)
    g = _lib.smoother_d(d, ys, u, v, ths, method=1, chunk=16)
    for k in range(5):
        o = O.smoother_d(d, ys, u, v, ths[k])
        assert abs(g["lik"][k] - o["lik"]) < 1e-10 * max(1.0, abs(o["lik"]))
        assert np.max(np.abs(g["X"][k] - o["X"])) < 1e-9 and np.max(np.abs(g["V"][k] - o["V"])) < 1e-9
    # u = v = None: theta has no B, D blocks
    A = 0.5 * np.eye(d)
    th = O.theta_d_flat(A, np.zeros((d, 0)), np.ones(d), np.zeros(0), np.eye(d), 0.5, np.zeros(d), np.eye(d))
    g = _lib.smoother_d(d, ys, None, None, th, method=1)
    o = O.smoother_d(d, ys, None, None, th)
    assert abs(g["lik"][0] - o["lik"]) < 1e-10 * max(1.0, abs(o["lik"])) and np.max(np.abs(g["X"][0] - o["X"])) < 1e-9


def test_config5_long_series_scan_vs_sequential():
    # BASELINE.json config 5 in small: d = 4, 20 proxies, 10 % missing; full length on the device only
    rng = np.random.default_rng(5)
    d, p, q, T = 4, 20, 20, 100_000
    y, u, v, th = random_model(rng, d, p, q, T, miss=0.1)
    s = _lib.smoother_d(d, y, u, v, th, method=0)
    g = _lib.smoother_d(d, y, u, v, th, method=1)
    assert abs(g["lik"][0] - s["lik"][0]) < 1e-10 * abs(s["lik"][0])
    assert np.max(np.abs(g["X"] - s["X"])) < 1e-9 and np.max(np.abs(g["V"] - s["V"])) < 1e-9
    # and the first 5 000 steps against the CPU oracle run on the same prefix are the filter only up to
    # the smoothing horizon: compare the filtered-dominated start of a shorter problem instead
    T2 = 5000
    o = O.smoother_d(d, y[:T2], u[:, :T2], v[:, :T2], th)
    g2 = _lib.smoother_d(d, y[:T2], u[:, :T2], v[:, :T2], th, method=1)
    assert np.max(np.abs(g2["X"][0] - o["X"])) < 1e-9 and abs(g2["lik"][0] - o["lik"]) < 1e-10 * abs(o["lik"])


def test_argument_errors():
    y = np.zeros(10)
    th = np.zeros(2 * 25 + 5 + 1 + 5 + 25)
    with pytest.raises(_lib.LdsrError) as e:
        _lib.smoother_d(5, y, None, None, th)
    assert e.value.code == _lib.ERR_UNSUPPORTED
    with pytest.raises(_lib.LdsrError):
        _lib.smoother_d(2, y, None, None, np.zeros(3))  # theta too short


def test_api_mirror_kalman_smoother_d():
    import ldsr_b200 as L
    rng = np.random.default_rng(3)
    d, p, q, T = 2, 3, 2, 120
    y, u, v, th = random_model(rng, d, p, q, T)
    o = O.smoother_d(d, y, u, v, th)
    k = 0
    blocks = {}
    for name, shape in (("A", (d, d)), ("B", (d, p)), ("C", (d,)), ("D", (q,)), ("Q", (d, d)), ("R", (1,)),
                        ("mu1", (d,)), ("V1", (d, d))):
        n = int(np.prod(shape))
        blocks[name] = th[k:k + n].reshape(shape)
        k += n
    g = L.Kalman_smoother_d(y, u, v, blocks)
    assert g["X"].shape == (d, T) and np.max(np.abs(g["X"].T - o["X"])) < 1e-9
    assert abs(g["lik"] - o["lik"]) < 1e-10 * max(1.0, abs(o["lik"]))
