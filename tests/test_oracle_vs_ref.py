"""Cross-checks of the three CPU checkers on the parts the reference's artefacts leave unpinned
(SURVEY.md 8c): the NumPy and C restatements against the reference's own sources compiled unmodified
into oracle/_ref/libssref.so, plus the reference-internal identity FastSmootherC(y) == KalmanC smoothed
state."""
import numpy as np
import pytest

from conftest import Case, config2_standin, random_model, simulate_y
from oracle import cport, kalman_np, ref

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libssref.so not built")

SHAPES = [
    dict(m=2, p=1, d=1, tv=()),
    dict(m=5, p=2, d=2, tv=("Z",)),
    dict(m=8, p=3, d=3, tv=("T", "Q")),
    dict(m=14, p=1, d=13, tv=(), bsm=True),   # config-5 structure (random dense T with 13 diffuse states is ill-conditioned)
    dict(m=6, p=2, d=0, tv=("R",)),
]


def _case(seed, m, p, d, tv, N=40, missing=0.1, bsm=False):
    rng = np.random.default_rng(seed)
    if bsm:
        from statespacer_b200 import sysmat
        sm = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]) + 0.25 * rng.normal(size=4))
    else:
        sm = random_model(rng, m, p, d=d, tv=tv, N=N)
    return Case(f"rand{seed}", simulate_y(rng, sm, N, missing), sm)


def _close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    scale = max(1.0, np.nanmax(np.abs(b))) if b.size else 1.0
    assert np.nanmax(np.abs(a - b), initial=0.0) <= tol * scale, np.nanmax(np.abs(a - b))


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
def test_loglik_all_checkers(shape):
    for seed in range(3):
        c = _case(seed, **shape)
        r = ref.LogLikC(*c.args())
        assert abs(kalman_np.LogLikC(*c.args()) - r) <= 1e-10 * max(1, abs(r))
        assert abs(cport.LogLikC(*c.args()) - r) <= 1e-10 * max(1, abs(r))


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
def test_kalman_numpy_vs_ref(shape):
    c = _case(11, **shape)
    a = kalman_np.KalmanC(*c.args(), False)
    b = ref.KalmanC(*c.args(), False)
    assert a["initialisation_steps"] == b["initialisation_steps"]
    for k in ("loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil", "V", "P_inf_pred", "P_star_pred",
              "P_inf_fil", "P_star_fil", "yfit", "v", "eta", "Fmat", "eta_var", "a_fc", "P_inf_fc", "P_star_fc",
              "P_fc", "r_vec", "Nmat"):
        _close(a[k], b[k], 1e-8 if k in ("P_pred", "P_fil", "Fmat", "P_fc") else 1e-9)


@needs_ref
@pytest.mark.parametrize("shape", SHAPES)
def test_fast_smoother_checkers(shape):
    c = _case(5, missing=0.0, **shape)
    steps = ref.KalmanC(*c.args(), False)["initialisation_steps"]
    rng = np.random.default_rng(0)
    nsim = 3
    y3 = np.stack([c.y + 0.1 * rng.normal(size=c.y.shape) * (s > 0) for s in range(nsim)], axis=2)
    b = ref.FastSmootherC(y3, *c.sm.args(), steps, True)
    for mod in (kalman_np, cport):
        a = mod.FastSmootherC(y3, *c.sm.args(), steps, True)
        for k in ("a_smooth", "a_t", "eta"):
            _close(a[k], b[k])
    # reference-internal identity: draw 0 is the data itself
    ks = ref.KalmanC(*c.args(), False)
    _close(b["a_smooth"][:, :, 0], ks["a_smooth"], 1e-8)


@needs_ref
@pytest.mark.parametrize("repeat_Q,draw_initial,eta_only", [(1, False, False), (3, True, False), (1, False, True)])
def test_simulate_checkers(repeat_Q, draw_initial, eta_only):
    rng = np.random.default_rng(3)
    N, nsim, r = 12, 5, 2
    r_tot = r * repeat_Q
    m, p = 4, 2
    sm = random_model(rng, m, p, r=r_tot, N=N, tv=("T",))
    A = rng.normal(size=(r, r))
    Q = A @ A.T + 0.1 * np.eye(r)
    n = (m * nsim if draw_initial else 0) + N * r_tot * nsim
    draws = rng.normal(size=n)
    if eta_only:
        Z = T = R = np.zeros((1, 1, 1))
        a = np.zeros(1)
        P = np.zeros((1, 1))
    else:
        Z, T, R, a, P = sm.Z, sm.T, sm.R, sm.a, sm.P_star + 0.01 * np.eye(m)
    b = ref.SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P, draw_initial, eta_only, True, draws)
    a1 = kalman_np.SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P, draw_initial, eta_only, True, draws)
    a2 = cport.SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P, draw_initial, eta_only, True, draws)
    keys = ("eta",) if eta_only else ("y", "a", "a_t", "eta")
    for k in keys:
        _close(a1[k], b[k], 1e-10)
        _close(a2[k], b[k], 1e-10)


@needs_ref
def test_extract_checkers():
    rng = np.random.default_rng(9)
    m, nsim, N, p = 5, 4, 7, 2
    a = rng.normal(size=(m, nsim, N))
    for Z in (rng.normal(size=(p, m, 1)), rng.normal(size=(p, m, N))):
        b = ref.ExtractComponentC(a, Z)
        _close(kalman_np.ExtractComponentC(a, Z), b, 1e-12)
        _close(cport.ExtractComponentC(a, Z), b, 1e-12)


@needs_ref
def test_config2_shape_all_checkers():
    """BASELINE.json config 2 in its synthetic stand-in (p = 2, N = 192, m = 34, 32 diffuse states, time-varying Z,
    missing observations): the three CPU checkers agree on the loglikelihood, the diffuse period and the smoothed
    state."""
    c = config2_standin()
    assert c.sm.m == 34 and c.sm.Z.shape == (2, 34, 192)
    r = ref.LogLikC(*c.args())
    assert np.isfinite(r)
    assert abs(cport.LogLikC(*c.args()) - r) <= 1e-9 * max(1, abs(r))
    assert abs(kalman_np.LogLikC(*c.args()) - r) <= 1e-9 * max(1, abs(r))
    a, b = kalman_np.KalmanC(*c.args(), False), ref.KalmanC(*c.args(), False)
    assert a["initialisation_steps"] == b["initialisation_steps"] >= 32
    for k in ("loglik", "a_smooth", "V", "eta", "a_fc"):
        _close(a[k], b[k], 1e-7)
