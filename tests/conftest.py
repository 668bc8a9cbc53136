"""Shared fixtures: KAT models built with statespacer_b200.sysmat, random models,
and the handles to the three CPU checkers (NumPy restatement, C restatement, compiled reference)."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from statespacer_b200 import sysmat  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_series(name):
    return np.loadtxt(os.path.join(GOLDEN, name), comments="#")


def load_fed(first="1985-01", last="2000-12"):
    rows = []
    with open(os.path.join(GOLDEN, "fedyieldcurve.csv")) as f:
        next(f)
        for line in f:
            parts = line.strip().split(";")
            if first <= parts[0] <= last:
                rows.append([float(x) for x in parts[1:]])
    return np.array(rows)


class Case:
    """One LogLikC/KalmanC call: y (N x p), mask, SysMat."""

    def __init__(self, name, y, sm, expect=None):
        self.name = name
        self.y = np.asarray(y, dtype=np.float64).reshape(len(y), -1)
        self.isna = np.isnan(self.y)
        self.sm = sm
        self.expect = expect

    def args(self):
        return (self.y, self.isna) + tuple(self.sm.args())


def kat_cases():
    nile = load_series("nile.txt")
    air = np.log(load_series("airpassengers.txt"))
    fed = load_fed()
    var_nile = float(np.var(nile, ddof=1))
    cases = [
        # docs/articles/intro.html:168,170  (optim prints -loglik/N)
        Case("KAT-1a", nile, sysmat.nile_local_level(var_nile, var_nile), expect=-6.623273),
        Case("KAT-1b", nile, sysmat.nile_local_level(15100.252, 1468.724), expect=-6.334646),
        # docs/articles/boxjenkins.html:241-242
        Case("KAT-4a", air, sysmat.airline([0.5 * math.log(float(np.var(np.diff(air), ddof=1))), 0.0, 0.0]),
             expect=1.034434),
    ]
    # docs/articles/selfspec.html:306-307, params vignettes/selfspec.Rmd:162-166
    dns_par = DNS_PARAM
    sm, _ = sysmat.dns_yield_curve(dns_par)
    cases.append(Case("KAT-3", fed, sm, expect=7.005830))
    return cases


# vignettes/selfspec.Rmd:162-166 (initial == optimum, docs/articles/selfspec.html:306-307)
DNS_PARAM = np.array([
    -1.27000018, -2.82637721, -1.35280609, -1.74569078, -0.89761151, -0.77767132, 1.01894143, 0.42965982,
    13.69072496, 3.46050373, -10.36767668, -0.07334641, 6.68658053, 0.76975206, 0.02852844, 0.50448668,
    2.99984132, 8.16851107, -2.28360681, -0.45333494])


def random_model(rng, m, p, r=None, d=0, tv=(), N=30, dense_q=True):
    """Random well-conditioned model: stable-ish T, d diffuse states, optional time-varying matrices."""
    r = r or m
    T = rng.normal(size=(m, m)) * (0.6 / math.sqrt(m))
    T[np.arange(m), np.arange(m)] += 0.5
    Z = rng.normal(size=(p, m))
    R = rng.normal(size=(m, r)) * 0.5
    A = rng.normal(size=(r, r))
    Q = A @ A.T / r + 0.1 * np.eye(r) if dense_q else np.diag(rng.uniform(0.1, 1.0, r))
    a = rng.normal(size=m)
    P_inf = np.zeros((m, m))
    P_inf[np.arange(d), np.arange(d)] = 1.0
    C = rng.normal(size=(m, m))
    P_star = C @ C.T / m
    P_star[:d, :] = 0
    P_star[:, :d] = 0

    def tvify(X, name, scale=0.05):
        if name not in tv:
            return X[:, :, None]
        return np.stack([X + scale * rng.normal(size=X.shape) if name != "Q"
                         else X * (1.0 + 0.3 * math.sin(i)) for i in range(N)], axis=2)

    return sysmat.SysMat(Z=tvify(Z, "Z"), T=tvify(T, "T"), R=tvify(R, "R"), Q=tvify(Q, "Q"), a=a, P_inf=P_inf,
                         P_star=P_star)


def simulate_y(rng, sm, N, missing=0.0):
    """Draw y from the model itself (finite start for diffuse states)."""
    m, p, r = sm.m, sm.p, sm.r
    a = sm.a + rng.normal(size=m)
    y = np.zeros((N, p))
    for t in range(N):
        Z = sm.Z[:, :, t if sm.Z.shape[2] > 1 else 0]
        T = sm.T[:, :, t if sm.T.shape[2] > 1 else 0]
        R = sm.R[:, :, t if sm.R.shape[2] > 1 else 0]
        Q = sm.Q[:, :, t if sm.Q.shape[2] > 1 else 0]
        y[t] = Z @ a
        L = np.linalg.cholesky(Q + 1e-12 * np.eye(r))
        a = T @ a + R @ (L @ rng.normal(size=r))
    if missing > 0:
        y[rng.uniform(size=y.shape) < missing] = np.nan
    return y


def config2_standin(seed=2, N=192, missing=0.05):
    """Synthetic stand-in for BASELINE.json config 2 (the Seatbelts data are not available offline; SURVEY.md
    8d): bivariate local level + slope + trigonometric seasonal(12) + three shared regressors, m = 34 incl. the two
    residual states, 32 exactly-diffuse states, time-varying Z, 5 % missing observations."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, 3))

    def cov(scale):
        A = rng.normal(size=(2, 2))
        return scale * (A @ A.T / 2 + 0.5 * np.eye(2))

    sm = sysmat.bsm_multivariate_regressors([X, X], cov(1.0), cov(0.1), cov(0.01), cov(0.05))
    a = np.zeros(sm.m)
    a[2:4] = rng.normal(size=2) * 3
    a[-6:] = rng.normal(size=6)
    y = np.zeros((N, 2))
    Lq = np.linalg.cholesky(sm.Q[:, :, 0] + 1e-12 * np.eye(sm.m))
    for t in range(N):
        y[t] = sm.Z[:, :, t] @ a
        a = sm.T[:, :, 0] @ a + Lq @ rng.normal(size=sm.m)
    y[rng.uniform(size=y.shape) < missing] = np.nan
    return Case("config2", y, sm)


@pytest.fixture(scope="session")
def kats():
    return kat_cases()
