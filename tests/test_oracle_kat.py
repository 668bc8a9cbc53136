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
