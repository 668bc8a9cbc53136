"""Pins the CPU checkers (oracle/) to the reference's own golden vectors:
numbers printed in the rendered vignettes under /root/reference/docs/articles/."""
import math

import numpy as np
import pytest

from conftest import Case, load_series
from oracle import cport, kalman_np, ref
from statespacer_b200 import sysmat



def airline_fitted():
    """Printed MA coefficients (s = 12 first, then s = 1: sarima_list order) back to the free parameters:
    coefficient = -A / sqrt(1 + A^2)  (R/AnsleyKohn.R:16, R/Components-BoxJenkins.R:647-700)."""
    def raw(theta):
        return -theta / math.sqrt(1 - theta * theta)
    return sysmat.airline([0.5 * math.log(0.001347882), raw(-0.55694248), raw(-0.40188859)])


CHECKERS = [("numpy", kalman_np), ("c", cport)] + ([("reference", ref)] if ref.available() else [])


@pytest.mark.parametrize("name,mod", CHECKERS)
def test_loglik_golden(kats, name, mod):
    for c in kats:
        got = mod.LogLikC(*c.args())
        # the vignettes print 6-7 significant digits
        assert abs(got - c.expect) < 6e-7 * max(1.0, abs(c.expect)), (c.name, name, got, c.expect)


@pytest.mark.parametrize("name,mod", [x for x in CHECKERS if x[0] != "c"])
def test_kalman_airline_fitted(name, mod):
    """KAT-4b: docs/articles/boxjenkins.html:242,269-271 -- fitted airline model, KalmanC sum(loglik)."""
    air = np.log(load_series("airpassengers.txt"))
    sm = airline_fitted()
    c = Case("KAT-4b", air, sm)
    out = mod.KalmanC(*c.args(), False)
    ll = np.nansum(out["loglik"])
    assert out["initialisation_steps"] == 13
    assert abs(ll - 232.750284785) < 2e-5          # parameters are printed with 8 digits
    assert abs(mod.LogLikC(*c.args()) * len(air) - ll) < 1e-9
    k = 3
    aic = (-2 * ll + 2 * (k + 13)) / len(air)       # R/StateSpaceEval.R:298-301
    assert abs(aic - (-3.010420622)) < 1e-6
    # optim printed -loglik/N = -1.616321 at the optimum
    assert abs(ll / len(air) - 1.616321) < 1e-6


def test_cholesky_golden():
    """KAT-U: docs/reference/Cholesky.html:136-159."""
    cov = sysmat.cholesky_cov([2.0, 4.0, 1.0], 2)
    assert np.allclose(cov, [[54.59815, 54.59815], [54.59815, 3035.55614]], rtol=1e-7)


def test_dns_phi_golden():
    """stationary VAR(1) coefficient of the DNS model."""
    from conftest import DNS_PARAM
    _, Phi = sysmat.dns_yield_curve(DNS_PARAM)
    assert np.all(np.abs(np.linalg.eigvals(Phi)) < 1)


def test_collapse_transform_kat3():
    """The collapse transform (R/statespacer.R:815-900) is likelihood preserving: KAT-3 with collapse = TRUE (p = 3,
    m = 9, vignettes/selfspec.Rmd:87-90) gives the printed 7.005830 like the uncollapsed model (p = 8, m = 14)."""
    from conftest import DNS_PARAM, load_fed
    from oracle import collapse_np
    fed = load_fed()
    sm, _ = sysmat.dns_yield_curve(DNS_PARAM)
    yk, Hs, la = collapse_np.collapse(fed, sm.Q[:8, :8, 0], sm.Z[:, 8:11, 0])
    smc = sysmat.dns_collapsed(DNS_PARAM, Hs[:, :, 0])
    assert (smc.p, smc.m) == (3, 9)
    ll = kalman_np.LogLikC(yk, np.isnan(yk), *smc.args()) + la / len(fed)
    assert abs(ll - 7.005830) < 6e-7 * 7
    assert abs(ll - kalman_np.LogLikC(fed, np.isnan(fed), *sm.args())) < 1e-11 * 7


def test_forecast_recursion_local_level_closed_form():
    """R/predict.R:205-246 on a local level model: the forecast is flat and its variance grows by Q per step."""
    from oracle import forecast_np
    H, Qv, a, P = 2.0, 0.5, 3.0, 1.25
    y_fc, a_fc, P_fc, F_fc = forecast_np.forecast(6, [a], [[P]], np.ones((1, 1, 1)), np.ones((1, 1, 1)),
                                                  np.ones((1, 1, 1)), np.full((1, 1, 1), Qv), np.full((1, 1, 1), H))
    assert np.allclose(y_fc[:, 0], a) and np.allclose(a_fc[:, 0], a)
    assert np.allclose(P_fc[0, 0], P + Qv * np.arange(6)) and np.allclose(F_fc[0, 0], P + Qv * np.arange(6) + H)


def test_numderiv_richardson_restatement():
    """oracle/numderiv_np.py (numDeriv::hessian, Richardson, defaults) on a function with a known Hessian."""
    from oracle import numderiv_np
    A = np.array([[2.0, 0.3, -0.1], [0.3, 1.5, 0.2], [-0.1, 0.2, 0.7]])
    f = lambda x: 0.5 * x @ A @ x + 0.25 * np.sum(x ** 4) + np.sin(x[0] * x[1])
    x = np.array([0.4, -0.8, 1.3])
    Han = A + np.diag(3 * x ** 2)
    Han[0, 0] += -np.sin(x[0] * x[1]) * x[1] ** 2
    Han[1, 1] += -np.sin(x[0] * x[1]) * x[0] ** 2
    Han[0, 1] += np.cos(x[0] * x[1]) - np.sin(x[0] * x[1]) * x[0] * x[1]
    Han[1, 0] = Han[0, 1]
    assert np.max(np.abs(numderiv_np.hessian(f, x) - Han)) < 1e-7
