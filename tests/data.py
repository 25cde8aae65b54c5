"""Shared inputs for the tests: the reference's bundled data decoded to tests/golden/*.json
(by tools/rda_to_golden.py) and the y/u/v constructions of the reference's R code."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def p1_case():
    """tests/testthat/test-LDS-EM.R:3-16: y 1x85 (no NA), u=v=t(P1pc[322:406]) 7x85, theta0."""
    p1, kat = load("p1.json"), load("kat.json")
    obs = np.log(np.array(p1["P1annual"]["Qa"]))
    y = obs - obs.mean()
    pc = np.array([p1["P1pc"][k] for k in p1["P1pc"]])
    lo, hi = kat["P1pc_rows_1based"]
    u = np.ascontiguousarray(pc[:, lo - 1:hi])
    t0 = kat["theta0"]
    th0 = np.concatenate([[t0["A"]], t0["B"], [t0["C"]], t0["D"], [t0["Q"], t0["R"], t0["mu1"], t0["V1"]]])
    return y, u, th0, kat


def np_case(first_row=1, start_year=1200):
    """y/u/v as LDS_reconstruction builds them (R/LDS_reconstruction.R:164-183) for
    u = v = t(NPpc[first_row:813]) and the given start.year.  Returns y[T] (NaN outside the
    instrumental period), u[3,T], mu, inst (0-based indices of the instrumental period)."""
    d = load("np.json")
    pc = np.array([d["NPpc"][k] for k in ("PC1", "PC9", "PC13")])
    u = np.ascontiguousarray(pc[:, first_row - 1:])
    T = u.shape[1]
    years = np.arange(start_year, start_year + T)
    qa_year = np.array(d["NPannual"]["year"])
    obs = np.log(np.array(d["NPannual"]["Qa"]))
    mu = obs.mean()
    y = np.full(T, np.nan)
    inst = np.nonzero(np.isin(years, qa_year))[0]
    y[inst] = obs - mu
    return y, u, mu, inst


def theta_of(d):
    return np.concatenate([d["A"], d["B"], d["C"], d["D"], d["Q"], d["R"], d["mu1"], d["V1"]]).astype(float)
