"""The reference's random numbers: R's default generators after set.seed (Mersenne-Twister +
Inversion).  R is the reference's host language and is not under /root/reference, so the pins are
values any R session prints (the set.seed(42) / set.seed(1) / set.seed(123) lines below are in
countless R tutorials and answers; they fix the seeding, the twister, the tempering, the uniform
scaling and the inversion in one go) and, for qnorm (AS 241), scipy's ndtri over the whole range.
CPU tests check the Python restatement (oracle/r_rng.py) and the library's host generator against
them; GPU tests check the device stream against the host generator."""
import numpy as np
import pytest

from ldsr_b200 import _lib
import ldsr_b200 as L
from oracle import r_rng as RO

KNOWN = [  # (seed, kind, values as R prints them with options(digits = 7..9))
    (42, "runif", [0.9148060, 0.9370754, 0.2861395]),
    (123, "runif", [0.2875775, 0.7883051]),
    (1, "rnorm", [-0.6264538, 0.1836433, -0.8356286]),
    (123, "rnorm", [-0.56047565, -0.23017749, 1.55870831]),
    (42, "rnorm", [1.37095845, -0.56469817]),
]


def _close(a, b):
    return np.allclose(a, b, rtol=0, atol=6e-8)


def test_python_restatement_reproduces_known_r_output():
    for seed, kind, vals in KNOWN:
        g = RO.RMT(seed)
        got = [g.unif() if kind == "runif" else g.norm() for _ in vals]
        assert _close(got, vals), (seed, kind, got)


def test_library_host_generator_reproduces_known_r_output_and_the_restatement():
    for seed, kind, vals in KNOWN:
        g = _lib.RRandom(seed)
        got = g.runif(len(vals)) if kind == "runif" else g.rnorm(len(vals))
        assert _close(got, vals), (seed, kind, got)
    # long mixed stream, bit for bit against the restatement (several regenerations of the 624 words)
    g, o = _lib.RRandom(2024), RO.RMT(2024)
    a = np.concatenate([g.runif(700), g.rnorm(400), g.runif(3, -1, 1), g.rnorm(900)])
    b = np.array([o.unif() for _ in range(700)] + [o.norm() for _ in range(400)]
                 + [-1 + 2 * o.unif() for _ in range(3)] + [o.norm() for _ in range(900)])
    assert np.array_equal(a, b)


def test_qnorm_as241_against_scipy_over_the_whole_range():
    from scipy.special import ndtri
    ps = np.concatenate([np.linspace(1e-9, 1 - 1e-9, 20001), 10.0 ** -np.arange(1, 300, 7.0),
                         np.exp(np.linspace(np.log(1e-40), np.log(1e-9), 2001))])
    mine = np.array([RO.qnorm(p) for p in ps])
    ref = ndtri(ps)
    big = np.abs(ref) > 1e-6
    assert np.max(np.abs(mine[big] / ref[big] - 1)) < 5e-15
    assert np.max(np.abs(mine[~big] - ref[~big])) < 1e-15


def test_make_init_with_r_generator_draws_in_the_reference_order():
    # set.seed(7); make_init(3, 2, 2): per restart runif(1), runif(p,-1,1), runif(1), runif(q,-1,1)
    # (R/LDS_reconstruction.R:14-30)
    init = L.make_init(3, 2, 2, L.RRandom(7))
    o = RO.RMT(7)
    for th in init:
        assert th["A"] == o.unif()
        assert np.array_equal(th["B"], [-1 + 2 * o.unif() for _ in range(3)])
        assert th["C"] == o.unif()
        assert np.array_equal(th["D"], [-1 + 2 * o.unif() for _ in range(2)])


def test_sample_reproduces_known_r_output_and_make_Z_uses_it():
    # R >= 3.6.0 (sample.kind = "Rejection")
    known = [(42, 10, 10, [1, 5, 10, 8, 2, 4, 6, 9, 7, 3]), (123, 10, 10, [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]),
             (1, 10, 10, [9, 4, 7, 1, 2, 5, 3, 10, 6, 8]), (123, 100, 5, [31, 79, 51, 14, 67])]
    for seed, n, k, vals in known:
        assert RO.sample_int(RO.RMT(seed), n, k) == vals
        assert list(_lib.RRandom(seed).sample_int(n, k)) == vals
    # make_Z (R/utils.R:83-101): contiguous = sort(sample(1:maxInd, nRuns)) then obsInd[x:(x+k)];
    # scattered = replicate(nRuns, sort(sample(obsInd, k)))
    obs = np.r_[np.full(5, np.nan), np.arange(46.0)]
    Z = L.make_Z(obs, nRuns=4, frac=0.25, contiguous=True, rng=L.RRandom(9))
    o = RO.RMT(9)
    starts = sorted(RO.sample_int(o, 46 - 11, 4))
    assert [list(z) for z in Z] == [list(range(5 + x, 5 + x + 12)) for x in starts]
    Z = L.make_Z(obs, nRuns=3, frac=0.25, contiguous=False, rng=L.RRandom(9))
    o = RO.RMT(9)
    for z in Z:
        assert list(z) == sorted(5 + i for i in RO.sample_int(o, 46, 11))


@pytest.mark.gpu
def test_device_stream_equals_the_host_generator():
    n = 5 * 312 + 17  # several regenerations and a ragged tail
    d = _lib.r_rnorm_device(1, n)
    assert _close(d[:3], [-0.6264538, 0.1836433, -0.8356286])
    h = _lib.RRandom(1).rnorm(n)
    central = np.abs(h) < 1.4395  # |p - 0.5| <= 0.425: pure polynomial arithmetic, identical rounding
    assert np.array_equal(d[central], h[central])
    assert np.allclose(d, h, rtol=2e-15, atol=0)  # tails go through log(): CUDA's vs libm's, last bits
    big = _lib.r_rnorm_device(99, 2_000_000)
    hb = _lib.RRandom(99).rnorm(2_000_000)
    assert np.allclose(big, hb, rtol=2e-15, atol=0)
    cb = np.abs(hb) < 1.4395
    assert np.array_equal(big[cb], hb[cb])
    assert abs(big.mean()) < 3e-3 and abs(big.std() - 1) < 3e-3


@pytest.mark.gpu
def test_lds_rep_with_r_seed_equals_replaying_r_draws():
    # set.seed(11); LDS_rep(...): one stream across the replicates, per replicate x1, n state draws,
    # n observation draws (R/stochastics.R:23-26, 60-61)
    rng = np.random.default_rng(0)
    n, reps = 60, 7
    u = rng.normal(size=(2, n))
    theta = dict(A=0.8, B=np.array([0.3, -0.2]), C=0.5, D=np.array([0.1, 0.05]), Q=1.3, R=0.2, mu1=0.0, V1=1.5)
    years = np.arange(1900, 1900 + n)
    a = L.LDS_rep(theta, u, u, years, num_reps=reps, mu=0.4, r_seed=11)
    z = _lib.RRandom(11).rnorm(reps * (1 + 2 * n)).reshape(reps, 1 + 2 * n)
    b = L.LDS_rep(theta, u, u, years, num_reps=reps, mu=0.4, z=z)
    for k in ("simX", "simY", "simQ"):
        assert np.allclose(a[k], b[k], rtol=1e-13, atol=0)
