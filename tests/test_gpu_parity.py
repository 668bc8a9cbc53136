"""Parity of the CUDA path (through the C ABI) with the CPU checkers.  Needs a B200: -m gpu.

Tolerances: 1e-9 relative on loglikelihoods and on smoothed states / variances (north star);
simulation smoother draws bit-exact against the pinned-order C oracle given the same variates."""
import math
import os

import numpy as np
import pytest

ROOT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from conftest import Case, DNS_PARAM, config2_standin, load_fed, load_series, random_model, simulate_y
from oracle import cport, kalman_np, ref
from statespacer_b200 import api, native, sysmat

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def rel(a, b):
    """Largest error of a against b, the whole array on one scale (loglikelihoods, per-draw arrays)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    if a.size == 0:
        return 0.0
    scale = max(1.0, float(np.nanmax(np.abs(b), initial=0.0)))
    return float(np.nanmax(np.abs(a - b), initial=0.0)) / scale


def entry_scale(b, ref_scale=None):
    """Scale of every entry of a KalmanC output for an ELEMENT-WISE comparison.

    m x m x N cubes (P_pred, V, Nmat, ...): per time slice, the entries within 1e-4 of the slice maximum are compared
    against that maximum, all the others against the largest of THEIR class -- a slice of P_pred / P_fil / Fmat in the
    diffuse phase mixes kappa = 1e7 scaled entries with O(1) ones (src/KalmanC.cpp:63,147-150), and one scale per slice
    would let the O(1) entries be wrong by 1e-2.  N x k arrays (a_smooth, r_vec, eta, ...): every column (a state, a
    disturbance) on its own scale over time.
    Floors: an entry below 1e-4 of its slice's maximum, or of `ref_scale` (the magnitude of the numbers the array is
    computed from; default: the array's own maximum), is the residue of a cancellation -- a filtered covariance of an
    ARIMA model is exactly zero in exact arithmetic -- and is held to 1e-4 of that scale instead of to itself:
    1e-9 there means 1e-13 absolute, a few hundred roundings."""
    B = np.abs(np.nan_to_num(np.asarray(b, dtype=float), nan=0.0))
    g = 1e-4 * max(float(ref_scale) if ref_scale is not None else float(B.max(initial=0.0)), 1e-300)
    if B.ndim == 3:
        sl = B.reshape(-1, B.shape[2])
        top = sl.max(axis=0, keepdims=True)
        big = sl >= 1e-4 * top
        rest = np.where(big, 0.0, sl).max(axis=0, keepdims=True)
        sc = np.where(big, top, np.maximum(rest, 1e-4 * top))
        sc = np.maximum(sc, g).reshape(B.shape)
    elif B.ndim == 2:
        sc = np.maximum(B.max(axis=0, keepdims=True), g) * np.ones_like(B)
    else:
        sc = np.full(B.shape, max(B.max(initial=0.0), g))
    return sc


def assert_entries(name, got, want, other=None, tol=RTOL, c=8.0, ref_scale=None, allow=None):
    """|got - want| <= tol element-wise (entry_scale).  `other`: the same quantity from the second CPU checker; where
    the two checkers themselves disagree by more than tol -- the recursion is ill-conditioned there, e.g. 32 exactly
    diffuse states, or a smoothed variance nine digits below the predicted one -- the bound is c times their
    disagreement (parity contract).  `allow`: absolute allowance per entry (broadcast)."""
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    assert np.array_equal(np.isnan(got), np.isnan(want)), (name, "NaN pattern differs")
    if got.size == 0:
        return
    sc = entry_scale(want, ref_scale)
    err = np.nan_to_num(np.abs(got - want) / sc)
    bound = np.full(err.shape, tol)
    if other is not None:
        bound = bound + c * np.nan_to_num(np.abs(np.asarray(other, dtype=float) - want) / sc)
    if allow is not None:
        bound = bound + allow / sc
    worst = np.unravel_index(np.argmax(err - bound), err.shape)
    assert np.all(err <= bound), (name, worst, float(err[worst]), float(bound[worst]), float(want[worst]))


def kalman_compare(name, got, want, alt, keys, p, tol=RTOL):
    """All arrays of one KalmanC result.  Covariances are floored by the magnitude of P_star / P_inf (what they are
    computed from).  P_pred, P_fil and Fmat of a diffuse time point are P_star + 1e7 P_inf (src/KalmanC.cpp:147-150,
    162,378-382): the rounding residue of an entry of P_inf that is zero in exact arithmetic (1e-17) arrives there as
    1e-10, in the reference as in any implementation, so those slices carry an absolute allowance of
    kappa x 16 eps x max|P_inf| -- P_star and P_inf themselves are held to 1e-9 element by element."""
    ps = float(np.nanmax(np.abs(want["P_star_pred"]), initial=0.0))
    pi = float(np.nanmax(np.abs(want["P_inf_pred"]), initial=0.0))
    N = np.asarray(want["P_pred"]).shape[2]
    t_init = min(N, -(-int(want["initialisation_steps"]) // p) + 1)
    kap = np.zeros((1, 1, N))
    kap[:, :, :t_init] = 1e7 * 16 * np.finfo(float).eps * pi
    refs = {k: ps for k in ("P_pred", "P_fil", "V", "P_star_pred", "P_star_fil", "P_star_fc", "P_fc")}
    refs.update({k: pi for k in ("P_inf_pred", "P_inf_fil", "P_inf_fc")})
    for k in keys:
        allow = kap if k in ("P_pred", "P_fil", "Fmat") else None
        rs = refs.get(k)
        if k == "Fmat" and t_init < N:
            rs = float(np.nanmax(np.abs(np.asarray(want[k])[:, :, t_init:]), initial=0.0))
        assert_entries((name, k), got[k], want[k], alt[k] if alt is not None else None, tol=tol, ref_scale=rs, allow=allow)


def loglik_want(*args):
    """The oracle's loglikelihood and the bound a GPU value is held to: 1e-9 relative (north star), widened by 8 x the
    disagreement of the two CPU checkers -- the pinned-order C restatement and the reference's own C++ -- where the
    recursion is conditioned so badly that they disagree by more themselves (parity contract)."""
    want = cport.LogLikC(*args)
    alt = ref.LogLikC(*args) if ref.available() else kalman_np.LogLikC(*args)
    return want, RTOL * max(1.0, abs(want)) + 8.0 * abs(alt - want)


# ---- LogLikC ------------------------------------------------------------------------------------------
def test_loglik_golden(kats):
    for c in kats:
        got = api.LogLikC(*c.args())
        assert abs(got - c.expect) < 6e-7 * max(1, abs(c.expect)), (c.name, got)
        assert abs(got - cport.LogLikC(*c.args())) <= RTOL * max(1, abs(got)), c.name


SHAPES = [
    dict(m=1, p=1, d=0, tv=()),
    dict(m=2, p=1, d=1, tv=()),
    dict(m=5, p=2, d=2, tv=("Z",)),
    dict(m=8, p=3, d=3, tv=("T", "Q")),
    dict(m=6, p=2, d=0, tv=("R",)),
    dict(m=20, p=4, d=2, tv=("Z", "T", "R", "Q")),
    dict(m=40, p=2, d=1, tv=()),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_loglik_random_models(shape):
    for seed in range(3):
        rng = np.random.default_rng(100 + seed)
        N = 35
        sm = random_model(rng, shape["m"], shape["p"], d=shape["d"], tv=shape["tv"], N=N)
        c = Case("r", simulate_y(rng, sm, N, missing=0.15), sm)
        want = cport.LogLikC(*c.args())
        got = api.LogLikC(*c.args())
        assert abs(got - want) <= RTOL * max(1, abs(want)), (shape, seed, got, want)
        # mask given separately from the values (R passes y_isna next to y)
        y0 = np.where(c.isna, 0.0, c.y)
        assert api.LogLikC(y0, c.isna, *sm.args()) == got


def test_loglik_edge_cases():
    sm = sysmat.nile_local_level(2.0, 1.0)
    y = np.full((6, 1), np.nan)
    assert math.isnan(api.LogLikC(y, np.isnan(y), *sm.args()))           # all missing -> NA (src/LogLikC.cpp:211)
    y1 = np.array([[3.0]])                                                # N = 1: only a diffuse step -> NA
    assert math.isnan(api.LogLikC(y1, np.isnan(y1), *sm.args())) and math.isnan(
        cport.LogLikC(y1, np.isnan(y1), *sm.args()))
    y3 = np.array([[3.0], [np.nan], [2.5]])
    assert abs(api.LogLikC(y3, np.isnan(y3), *sm.args()) - cport.LogLikC(y3, np.isnan(y3), *sm.args())) < 1e-12
    sm0 = sysmat.nile_local_level(0.0, 0.0)                               # F = 0 everywhere after init -> NA
    y2 = np.ones((5, 1))
    assert math.isnan(api.LogLikC(y2, np.isnan(y2), *sm0.args())) == math.isnan(
        cport.LogLikC(y2, np.isnan(y2), *sm0.args()))
    with pytest.raises(ValueError):
        api.LogLikC(np.ones((5, 2)), np.zeros((5, 2), bool), *sm.args())
    with pytest.raises(native.SsgpuError):
        big = random_model(np.random.default_rng(0), 65, 1)
        api.LogLikC(np.ones((3, 1)), np.zeros((3, 1), bool), *big.args())


# ---- KalmanC ------------------------------------------------------------------------------------------
KALMAN_KEYS = ("loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil", "V", "P_inf_pred", "P_star_pred",
               "P_inf_fil", "P_star_fil", "yfit", "v", "eta", "Fmat", "eta_var", "a_fc", "P_inf_fc", "P_star_fc",
               "P_fc", "r_vec", "Nmat")


def _kalman_check(c, tol=RTOL, diagnostics=False):
    """Every array of KalmanC against the compiled reference, element-wise at 1e-9; the NumPy restatement is the second
    checker that widens the bound where the two CPU implementations disagree themselves."""
    alt = kalman_np.KalmanC(*c.args(), diagnostics)
    want = ref.KalmanC(*c.args(), diagnostics) if ref.available() else alt
    got = api.KalmanC(*c.args(), diagnostics)
    assert got["initialisation_steps"] == want["initialisation_steps"]
    kalman_compare(c.name, got, want, alt, KALMAN_KEYS, c.y.shape[1], tol=tol)
    return got, want, alt


def test_kalman_golden_models(kats):
    for c in kats:
        _kalman_check(c)
    air = np.log(load_series("airpassengers.txt"))
    def raw(t):
        return -t / math.sqrt(1 - t * t)
    sm = sysmat.airline([0.5 * math.log(0.001347882), raw(-0.55694248), raw(-0.40188859)])
    got = api.KalmanC(air[:, None], np.isnan(air[:, None]), *sm.args(), False)
    assert got["initialisation_steps"] == 13
    assert abs(np.nansum(got["loglik"]) - 232.750284785) < 2e-5      # docs/articles/boxjenkins.html:269-271


@pytest.mark.parametrize("shape", SHAPES[:6])
def test_kalman_random_models(shape):
    rng = np.random.default_rng(7)
    N = 30
    sm = random_model(rng, shape["m"], shape["p"], d=shape["d"], tv=shape["tv"], N=N)
    _kalman_check(Case("r", simulate_y(rng, sm, N, missing=0.1), sm))


@pytest.mark.parametrize("shape", [dict(m=2, p=1, d=1, tv=()), dict(m=5, p=2, d=2, tv=("Z",)),
                                   dict(m=8, p=3, d=0, tv=("T", "Q")), dict(m=14, p=8, d=0, tv=())])
def test_kalman_diagnostics(shape):
    """diagnostics = TRUE: v_norm, e, D, Tstat_* (src/KalmanC.cpp:164-187,528-544)."""
    rng = np.random.default_rng(21)
    N = 30
    sm = random_model(rng, shape["m"], shape["p"], d=shape["d"], tv=shape["tv"], N=N)
    c = Case("diag", simulate_y(rng, sm, N, missing=0.1), sm)
    got, want, alt = _kalman_check(c, diagnostics=True)
    for k in ("v_norm", "e", "D", "Tstat_observation"):
        assert_entries(k, got[k], want[k], alt[k])
    assert_entries("Tstat_state", got["Tstat_state"][1:], want["Tstat_state"][1:], alt["Tstat_state"][1:])


def test_kalman_diagnostics_nile():
    nile = load_series("nile.txt")[:, None]
    sm = sysmat.nile_local_level(15100.252, 1468.724)
    got = api.KalmanC(nile, np.isnan(nile), *sm.args(), True)
    want = kalman_np.KalmanC(nile, np.isnan(nile), *sm.args(), True)
    for k in ("v_norm", "e", "D", "Tstat_observation"):
        assert_entries(k, got[k], want[k])
    assert math.isnan(got["v_norm"][0, 0]) and not math.isnan(got["v_norm"][1, 0])   # diffuse first step


def test_kalman_missing_airline():
    """SURVEY.md appendix C probe: airline with 6 NaNs -> initialisation_steps = 19."""
    air = np.log(load_series("airpassengers.txt")).copy()
    air[[5, 6, 50, 51, 52, 100]] = np.nan
    sm = sysmat.airline([0.5 * math.log(0.00135), 0.3, 0.4])
    c = Case("air-na", air, sm)
    got = api.KalmanC(*c.args(), False)
    assert got["initialisation_steps"] == 19
    assert abs(np.nansum(got["loglik"]) - api.LogLikC(*c.args()) * len(air)) < 1e-9 * 300
    _kalman_check(c)


def test_config2_shape_seatbelts_standin():
    """BASELINE.json config 2 (synthetic stand-in): bivariate BSM + regressors, p = 2, N = 192,
    m = 34, 32 exactly-diffuse states, time-varying Z, 5 % missing -- LogLikC, the whole KalmanC output and a small
    batch (generic kernel: m > 23, time-varying Z) against the oracle / the compiled reference."""
    c = config2_standin()
    want, tol = loglik_want(*c.args())
    got = api.LogLikC(*c.args())
    assert abs(got - want) <= tol, (got, want, tol)
    _kalman_check(c)
    k = api.KalmanC(*c.args(), False)
    assert abs(np.nansum(k["loglik"]) - got * len(c.y)) <= tol * len(c.y)
    # a batch: the same series scaled (scale c: loglik shifts by -(n_obs - d_obs) log(c) / N, see the properties test)
    ys = np.stack([c.y, 2.0 * c.y, c.y], axis=2)
    Qb = np.stack([c.sm.Q[:, :, 0], 4.0 * c.sm.Q[:, :, 0], c.sm.Q[:, :, 0]])
    Pb = np.stack([c.sm.P_star, 4.0 * c.sm.P_star, c.sm.P_star])
    out = api.loglik_batch(ys, c.sm, Q=Qb, P_star=Pb)
    assert api.last_kernel().startswith("loglik_wsm_kernel<TT=5")       # m = 34: warp + shared-memory kernel
    gen = api.loglik_batch(ys, c.sm, Q=Qb, P_star=Pb, kernel="generic")
    assert np.max(np.abs(out - gen)) <= 2 * tol
    assert out[0] == out[2] and abs(out[0] - want) <= tol
    n_obs = int(np.sum(~np.isnan(c.y)))
    assert abs(out[1] - (out[0] - (n_obs - 32) * math.log(2.0) / len(c.y))) <= 2 * tol


# ---- batched loglikelihood ---------------------------------------------------------------------------
def _bsm_batch(rng, B, N, V, missing):
    thetas = 0.5 * np.log([1.0, 0.1, 0.01, 0.05]) + 0.25 * rng.normal(size=(V, B, 4))
    var = np.exp(2 * thetas)
    Qd = np.concatenate([var[..., :3], np.repeat(var[..., 3:4], 11, axis=-1)], axis=-1)   # (V, B, 14)
    Pd = np.zeros((V, B, 14))
    Pd[..., 0] = var[..., 0]
    y = np.cumsum(rng.normal(size=(N, B)), axis=0) + 3 * np.sin(np.arange(N) * 2 * np.pi / 12)[:, None]
    y[rng.uniform(size=y.shape) < missing] = np.nan
    return thetas, Qd, Pd, y


KERNEL_NAMES = {"mma": "loglik_mma_kernel<TT=2,tiles=0x9,", "mma_dense": "loglik_mma_kernel<TT=2,tiles=0xf,",
                "cols": "loglik_cols", "generic": "fwd_generic", "blk": "loglik_thr_kernel<NB=7,",
                "auto": "loglik_blk_kernel<L=8,NB=7,",     # 210 evaluations: below the size the dispatch hands to the thread kernel
                "blk_lanes": "loglik_blk_kernel<L=8,NB=7,"}


@pytest.mark.parametrize("kernel", ["mma", "mma_dense", "cols", "generic", "blk", "blk_lanes", "auto"])
def test_batch_config5_shape(kernel):
    rng = np.random.default_rng(5)
    B, N, V = 70, 60, 3
    thetas, Qd, Pd, y = _bsm_batch(rng, B, N, V, 0.05)
    got = api.loglik_batch(y, sysmat.bsm12(thetas[0, 0]), V=V, Q_diag=Qd, P_star_diag=Pd, kernel=kernel)
    assert api.last_kernel().startswith(KERNEL_NAMES[kernel])
    assert got.shape == (V, B)
    for v in range(V):
        for b in range(0, B, 7):
            sm = sysmat.bsm12(thetas[v, b])
            want = cport.LogLikC(y[:, b:b + 1], np.isnan(y[:, b:b + 1]), *sm.args())
            assert abs(got[v, b] - want) <= RTOL * max(1, abs(want)), (v, b, got[v, b], want)


@pytest.mark.parametrize("m", list(range(1, 24)))
def test_batch_every_state_dim(m):
    """Register kernels at every supported state dimension (tensor kernel m <= 23, column kernel m <= 15),
    per-series dense Q / P_star / a, p in {1, 2}, diffuse states, missing values."""
    rng = np.random.default_rng(40 + m)
    B, N, p = 9, 25, 1 if m % 3 else 2
    d = min(m - 1, 2) if m > 1 else 0
    sm = random_model(rng, m, p, d=d, N=N)
    y = np.stack([simulate_y(rng, sm, N, missing=0.1) for _ in range(B)], axis=2)       # (N, p, B)
    # per-series dense Q and P_star, a, P_inf
    Qb = np.stack([sm.Q[:, :, 0] * (1 + 0.1 * b) for b in range(B)])
    Pb = np.stack([sm.P_star * (1 + 0.05 * b) for b in range(B)])
    ab = np.stack([sm.a + 0.01 * b for b in range(B)])
    got = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab, kernel="mma")
    gen = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab, kernel="generic")
    cols = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab, kernel="cols") if m <= 15 else got
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=Qb[b][:, :, None], a=ab[b], P_inf=sm.P_inf, P_star=Pb[b])
        yb = y[:, :, b]
        want, tol = loglik_want(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= tol, (m, b, got[b], want)
        assert abs(gen[b] - want) <= tol, (m, b, gen[b], want)
        assert abs(cols[b] - want) <= tol, (m, b, cols[b], want)


def _cycle(rho, lam, var):
    """Stationary stochastic cycle (R/Components-Unobserved.R: cycle, damping rho < 1): a 2 x 2 block of T, proper
    initial covariance var / (1 - rho^2) I."""
    c, s = rho * math.cos(lam), rho * math.sin(lam)
    return dict(Z=np.array([[1.0, 0.0]]), T=np.array([[c, s], [-s, c]]), R=np.eye(2), Q=var * np.eye(2), a=np.zeros(2),
                P_inf=np.zeros((2, 2)), P_star=var / (1 - rho * rho) * np.eye(2))


def _structural_models():
    S = sysmat
    return {
        "level (m=2, 1 block)": S._combine(1, [[0.7]], [S.comp_local_level(0.3)]),
        "level+slope (m=3, 2 blocks)": S._combine(1, [[0.7]], [S.comp_level_slope(0.3, 0.02)]),
        "level+seasonal4 (m=5, 3 blocks)": S._combine(1, [[0.5]], [S.comp_local_level(0.2), S.comp_bsm_trig(4, 0.05)]),
        "slope+seasonal7 (m=9, 5 blocks)": S._combine(1, [[0.5]], [S.comp_level_slope(0.2, 0.01), S.comp_bsm_trig(7, 0.05)]),
        "slope+seasonal12 (m=14, 7 blocks)": S.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05])),
        "slope+seasonal12+cycle (m=16, 8 blocks)": S._combine(
            1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(12, 0.05), _cycle(0.9, 0.4, 0.3)]),
        "slope+seasonal12+seasonal7 (m=20, 10 blocks)": S._combine(
            1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(12, 0.05), S.comp_bsm_trig(7, 0.03)]),
        "slope+seasonal24 (m=26, 13 blocks)": S._combine(
            1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(24, 0.05)]),
        "slope+seasonal24+seasonal7 (m=32, 16 blocks)": S._combine(
            1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(24, 0.05), S.comp_bsm_trig(7, 0.03)]),
    }


@pytest.mark.parametrize("name", list(_structural_models()))
def test_batch_blk_kernel(name):
    """Block-row kernel (T block diagonal with blocks of at most 2 x 2 after re-ordering the states) at every lane
    count, per-evaluation diagonal Q / P_star and a, exactly diffuse states, missing values (groups of one warp
    leave the diffuse phase at different time points), against the oracle and the generic kernel."""
    sm = _structural_models()[name]
    rng = np.random.default_rng(len(name))
    B, m = 37, sm.m
    N = 45 if m <= 16 else 3 * m                       # long enough for the d = m - 1 diffuse states to be identified
    y = np.stack([simulate_y(rng, sm, N, missing=0.12)[:, 0] for _ in range(B)], axis=1)      # (N, B)
    y[:3, 5] = np.nan                                                   # a series that starts with missing values
    scale = rng.uniform(0.5, 2.0, size=(B, 1))
    Qd = np.diag(sm.Q[:, :, 0])[None, :] * scale
    Pd = np.diag(sm.P_star)[None, :] * scale
    ab = rng.normal(size=(B, m)) * 0.1
    got = api.loglik_batch(y, sm, Q_diag=Qd, P_star_diag=Pd, a=ab, kernel="blk")
    # 2 to 8 blocks: the lanes run the diffuse phase and hand over to the thread-per-evaluation kernel
    first = "loglik_thr_kernel<" if 3 <= m <= 16 else "loglik_blk_kernel<"
    assert api.last_kernel().startswith(first), api.last_kernel()
    gen = api.loglik_batch(y, sm, Q_diag=Qd, P_star_diag=Pd, a=ab, kernel="generic")
    auto = api.loglik_batch(y, sm, Q_diag=Qd, P_star_diag=Pd, a=ab)            # a small batch: the lanes
    assert api.last_kernel().startswith("loglik_blk_kernel<") and np.allclose(auto, got, rtol=1e-12, atol=0, equal_nan=True)
    lanes = api.loglik_batch(y, sm, Q_diag=Qd, P_star_diag=Pd, a=ab, kernel="blk_lanes")
    assert api.last_kernel().startswith("loglik_blk_kernel<"), api.last_kernel()
    assert np.all(np.isfinite(got)) and np.all(np.isfinite(lanes))
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=np.diag(Qd[b])[:, :, None], a=ab[b], P_inf=sm.P_inf,
                            P_star=np.diag(Pd[b]))
        yb = y[:, b:b + 1]
        want = cport.LogLikC(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (name, b, got[b], want)
        assert abs(lanes[b] - want) <= RTOL * max(1, abs(want)), (name, b, lanes[b], want)
        assert abs(gen[b] - want) <= RTOL * max(1, abs(want)), (name, b, gen[b], want)


@pytest.mark.parametrize("name", ["level_slope", "level_seasonal4", "bsm12", "bsm12_cycle"])
def test_thread_kernel_residual_shortcut(name):
    """The thread-per-evaluation kernel drops the residual state of the reference's layout (loglik_thr.cu, EPS): with a
    diffuse phase, from a start WITHOUT one (P_inf = 0; P_star[eps, eps] different from H is honoured at the first time
    point), and not at all when a[eps] != 0 or P_star couples the residual -- every variant against the oracle."""
    rng = np.random.default_rng(len(name) + 3)
    comps = {"level_slope": [sysmat.comp_level_slope(0.3, 0.02)],
             "level_seasonal4": [sysmat.comp_local_level(0.2), sysmat.comp_bsm_trig(4, 0.05)],
             "bsm12": [sysmat.comp_level_slope(0.1, 0.01), sysmat.comp_bsm_trig(12, 0.05)],
             "bsm12_cycle": [sysmat.comp_level_slope(0.1, 0.01), sysmat.comp_bsm_trig(12, 0.05),
                             _cycle(0.9, 0.4, 0.3)]}[name]            # 8 blocks, the residual paired with a seasonal state
    sm = sysmat._combine(1, [[0.6]], comps)
    m, N, B = sm.m, 60, 48
    y = np.stack([simulate_y(rng, sm, N, missing=0.1)[:, 0] for _ in range(B)], axis=1)
    y[:4, 3] = np.nan

    def check(model, dropped, **kw):
        got = api.loglik_batch(y, model, kernel="blk", **kw)
        k = api.last_kernel()
        assert k.startswith("loglik_thr_kernel<") and ("residual state dropped" in k) == dropped, k
        for b in range(0, B, 5):
            mb = model
            if "P_star_diag" in kw:
                mb = sysmat.SysMat(Z=model.Z, T=model.T, R=model.R, Q=np.diag(kw["Q_diag"][b])[:, :, None], a=model.a,
                                   P_inf=model.P_inf, P_star=np.diag(kw["P_star_diag"][b]))
            yb = y[:, b:b + 1]
            want = cport.LogLikC(yb, np.isnan(yb), *mb.args())
            assert math.isfinite(want) and abs(got[b] - want) <= RTOL * max(1, abs(want)), (name, dropped, b, got[b], want)
        return got

    base = check(sm, True)
    lanes = api.loglik_batch(y, sm, kernel="blk_lanes")
    assert np.allclose(base, lanes, rtol=1e-11, atol=0)
    # no diffuse phase: the kernel starts from the caller's P_star, whose residual entry need not be H
    P0 = np.diag(rng.uniform(0.5, 2.0, m))
    P0[0, 0] = 1.7
    flat = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=sm.a, P_inf=np.zeros((m, m)), P_star=P0)
    check(flat, True)
    scale = rng.uniform(0.5, 2.0, size=(B, 1))
    check(flat, True, Q_diag=np.diag(sm.Q[:, :, 0])[None, :] * scale, P_star_diag=np.diag(P0)[None, :] * scale)
    # what the shortcut assumes does not hold: a[eps] != 0, P_star couples the residual with another state
    a1 = sm.a.copy()
    a1[0] = 0.3
    check(sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=a1, P_inf=sm.P_inf, P_star=sm.P_star), False)
    P1 = P0.copy()
    P1[0, 1] = P1[1, 0] = 0.2
    check(sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=sm.a, P_inf=np.zeros((m, m)), P_star=P1), False)


@pytest.mark.parametrize("seasonal", [12, 4, 24])
def test_batch_blk_kernel_time_varying_z(seasonal):
    """Univariate structural model with two explanatory variables (R/Components-Regressors.R:60-78: constant, exactly
    diffuse coefficient states whose row of Z is the regressor value at t): time-varying Z on the block-row kernel,
    against the oracle and the generic kernel; the automatic dispatch picks it."""
    rng = np.random.default_rng(60 + seasonal)
    N, B = (70 if seasonal <= 12 else 110), 23
    base = sysmat._combine(1, [[0.8]], [sysmat.comp_level_slope(0.1, 0.01), sysmat.comp_bsm_trig(seasonal, 0.05)])
    m0 = base.m
    X = rng.normal(size=(N, 2))
    m = m0 + 2
    Z = np.zeros((1, m, N))
    Z[0, :m0, :] = base.Z[0, :, 0][:, None]
    Z[0, m0:, :] = X.T
    T = sysmat.block_diag(base.T[:, :, 0], np.eye(2))
    R = np.concatenate([np.eye(m0), np.zeros((2, m0))], axis=0)
    P_inf = sysmat.block_diag(base.P_inf, np.eye(2))
    P_star = sysmat.block_diag(base.P_star, np.zeros((2, 2)))
    sm = sysmat.SysMat(Z=Z, T=T[:, :, None], R=R[:, :, None], Q=base.Q, a=np.zeros(m), P_inf=P_inf, P_star=P_star)
    beta = np.array([1.5, -0.7])
    y = np.stack([simulate_y(rng, base, N, missing=0.08)[:, 0] + X @ beta for _ in range(B)], axis=1)
    scale = rng.uniform(0.5, 2.0, size=(B, 1))
    Qd = np.diag(base.Q[:, :, 0])[None, :] * scale
    got = api.loglik_batch(y, sm, Q_diag=Qd, kernel="blk")
    assert "time-varying Z" in api.last_kernel(), api.last_kernel()
    # up to 8 blocks the regular phase runs one thread per evaluation (the row of Z loaded per step), beyond on the lanes
    first = "loglik_thr_kernel<" if m <= 16 else "loglik_blk_kernel<"
    assert api.last_kernel().startswith(first), api.last_kernel()
    lanes = api.loglik_batch(y, sm, Q_diag=Qd, kernel="blk_lanes")
    assert api.last_kernel().startswith("loglik_blk_kernel<") and "time-varying Z" in api.last_kernel()
    gen = api.loglik_batch(y, sm, Q_diag=Qd, kernel="generic")
    auto = api.loglik_batch(y, sm, Q_diag=Qd)                                   # a small batch: the lanes
    assert api.last_kernel().startswith("loglik_blk_kernel<") and np.allclose(auto, got, rtol=1e-12, atol=0)
    assert np.all(np.isfinite(got))
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=np.diag(Qd[b])[:, :, None], a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star)
        yb = y[:, b:b + 1]
        want, tol = loglik_want(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= tol, (b, got[b], want)
        assert abs(lanes[b] - want) <= tol, (b, lanes[b], want)
        assert abs(gen[b] - want) <= tol, (b, gen[b], want)
        if b < 3:                                     # the single-series entry point takes the same kernel
            one = api.LogLikC(yb, np.isnan(yb), *smb.args())
            assert api.last_kernel().startswith("loglik_blk_kernel<") and abs(one - got[b]) <= 1e-12 * abs(one)


def test_batch_blk_kernel_bivariate_dense_blocks():
    """p = 2: two local levels with correlated disturbances.  T = diag(0, 0, 1, 1); R Q R' = blkdiag(H, Q_level) has
    dense 2 x 2 blocks, which pairs the residuals and the levels; shared dense Q, per-series dense P_star and a."""
    rng = np.random.default_rng(77)
    H = np.array([[1.0, 0.3], [0.3, 0.5]])
    Ql = np.array([[0.2, -0.05], [-0.05, 0.1]])
    sm = sysmat.SysMat(Z=np.concatenate([np.eye(2), np.eye(2)], axis=1)[:, :, None],
                       T=np.diag([0.0, 0.0, 1.0, 1.0])[:, :, None], R=np.eye(4)[:, :, None],
                       Q=sysmat.block_diag(H, Ql)[:, :, None], a=np.zeros(4),
                       P_inf=np.diag([0.0, 0.0, 1.0, 1.0]), P_star=sysmat.block_diag(H, np.zeros((2, 2))))
    B, N = 21, 40
    y = np.stack([simulate_y(rng, sm, N, missing=0.15) for _ in range(B)], axis=2)            # (N, 2, B)
    Pb = np.stack([sm.P_star * (1 + 0.05 * b) for b in range(B)])
    ab = rng.normal(size=(B, 4)) * 0.1
    got = api.loglik_batch(y, sm, P_star=Pb, a=ab, kernel="blk")
    assert api.last_kernel().startswith("loglik_blk_kernel<L=2,NB=2,"), api.last_kernel()
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=ab[b], P_inf=sm.P_inf, P_star=Pb[b])
        yb = y[:, :, b]
        want = cport.LogLikC(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (b, got[b], want)


@pytest.mark.parametrize("N", [1, 2, 14, 40])
def test_batch_blk_kernel_degenerate_series(N):
    """Series of one warp that behave differently: all observations missing (NA, src/LogLikC.cpp:211-213), missing
    until late (the group leaves the diffuse phase long after its neighbours, or never), zero variances (F = 0 after the
    initialisation), very short series (N = 1: only a diffuse step)."""
    rng = np.random.default_rng(50 + N)
    B = 13
    thetas, Qd, Pd, y = _bsm_batch(rng, B, N, 1, 0.1)
    Qd, Pd = Qd[0].copy(), Pd[0].copy()
    y[:, 1] = np.nan                                  # never observed
    y[: max(N - 3, 0), 2] = np.nan                    # observed only at the end: still diffuse when the data stop
    y[: N // 2, 3] = np.nan
    Qd[4], Pd[4] = 0.0, 0.0                           # no noise at all
    got = api.loglik_batch(y, sysmat.bsm12(thetas[0, 0]), Q_diag=Qd, P_star_diag=Pd, kernel="blk")
    assert api.last_kernel().startswith("loglik_thr_kernel<NB=7,"), api.last_kernel()
    lanes = api.loglik_batch(y, sysmat.bsm12(thetas[0, 0]), Q_diag=Qd, P_star_diag=Pd, kernel="blk_lanes")
    assert np.allclose(got, lanes, rtol=RTOL, atol=RTOL, equal_nan=True)
    gen = api.loglik_batch(y, sysmat.bsm12(thetas[0, 0]), Q_diag=Qd, P_star_diag=Pd, kernel="generic")
    for b in range(B):
        sm = sysmat.bsm12(thetas[0, b])
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=np.diag(Qd[b])[:, :, None], a=sm.a, P_inf=sm.P_inf,
                            P_star=np.diag(Pd[b]))
        yb = y[:, b:b + 1]
        want = cport.LogLikC(yb, np.isnan(yb), *smb.args())
        assert math.isnan(got[b]) == math.isnan(want), (N, b, got[b], want)
        assert math.isnan(gen[b]) == math.isnan(want), (N, b, gen[b], want)
        if not math.isnan(want):
            assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (N, b, got[b], want)


def test_batch_blk_kernel_not_eligible():
    """A transition matrix that couples a state with two others (SARIMA companion block, dense T) is refused by
    the explicit selector and passed on by the automatic dispatch."""
    rng = np.random.default_rng(3)
    sm = random_model(rng, 6, 1, d=1, N=20, dense_q=False)
    y = np.stack([simulate_y(rng, sm, 20)[:, 0] for _ in range(4)], axis=1)
    with pytest.raises(RuntimeError, match="block kernel not eligible"):
        api.loglik_batch(y, sm, kernel="blk")
    auto = api.loglik_batch(y, sm)
    assert not api.last_kernel().startswith("loglik_blk")
    assert np.all(np.isfinite(auto))


BSM_Q_INDEX = [0, 1, 2] + [3] * 11       # config 5: Q = diag(s2_eps, s2_level, s2_slope, s2_seas x 11)
BSM_P_INDEX = [0] + [-1] * 13            # P_star = diag(s2_eps, 0, ...)


def test_loglik_grad_batch_from_theta():
    """theta -> variances on the device + central-difference gradient from one launch
    against the oracle evaluated at the same 1 + 2k parameter vectors with matrices built on the host."""
    import torch
    rng = np.random.default_rng(31)
    B, N, k, h = 40, 80, 4, 1e-3
    thetas, _, _, y = _bsm_batch(rng, B, N, 1, 0.03)
    theta = thetas[0]                                                   # (B, 4)
    sm0 = sysmat.bsm12(theta[0])
    ll, gr = api.loglik_grad_batch(y, sm0, theta, BSM_Q_INDEX, BSM_P_INDEX, h=h)
    assert api.last_kernel().startswith("loglik_blk_kernel<L=8,NB=7,")          # 360 evaluations: the lanes
    for b in range(0, B, 5):
        yb = y[:, b:b + 1]
        f = lambda th: cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(th).args())
        want = f(theta[b])
        assert abs(ll[b] - want) <= RTOL * max(1, abs(want))
        for i in range(k):
            e = np.zeros(k)
            e[i] = h
            g = (f(theta[b] + e) - f(theta[b] - e)) / (2 * h)
            assert abs(gr[b, i] - g) <= 1e-9 * max(1, abs(want)) / h * 2 + 1e-9 * abs(g), (b, i, gr[b, i], g)
    # device-resident variants, bit-identical to the host-pointer entry point
    dev = torch.device("cuda")
    model = api.BatchModel(sm0, N, B, 1, device=dev)
    y_d, th_d = torch.from_numpy(y).to(dev), torch.from_numpy(theta).to(dev)
    ll_d, gr_d = api.loglik_grad_batch_dev(y_d, model, th_d, BSM_Q_INDEX, BSM_P_INDEX, h=h)
    assert np.array_equal(ll_d.cpu().numpy(), ll) and np.array_equal(gr_d.cpu().numpy(), gr)
    pts = torch.stack([th_d, th_d + 0.1, th_d - 0.2])
    out = api.loglik_theta_batch_dev(y_d, model, pts, BSM_Q_INDEX, BSM_P_INDEX).cpu().numpy()
    assert np.array_equal(out[0], ll)
    yb = y[:, 7:8]
    want = cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(theta[7] - 0.2).args())
    assert abs(out[2, 7] - want) <= RTOL * max(1, abs(want))
    # argument errors
    with pytest.raises(api.native.SsgpuError):
        api.loglik_grad_batch(y, sm0, theta, [9] * 14, BSM_P_INDEX)


def test_fit_batch_bfgs():
    """The batched caller of the path (statespacer_b200/fit.py): every series' loglikelihood ends no lower than at
    its start values and no lower than at the data-generating parameters (up to optimiser tolerance), gradients
    vanish, and for a few series the optimum agrees with scipy's BFGS on the oracle objective."""
    import torch
    from scipy.optimize import minimize
    from statespacer_b200 import fit
    rng = np.random.default_rng(77)
    B, N, k = 48, 120, 4
    truth = 0.5 * np.log([1.0, 0.1, 0.01, 0.05]) + 0.2 * rng.normal(size=(B, k))
    y = np.zeros((N, B))
    for b in range(B):
        y[:, b] = simulate_y(rng, sysmat.bsm12(truth[b]), N)[:, 0]
    dev = torch.device("cuda")
    sm0 = sysmat.bsm12(truth[0])
    model = api.BatchModel(sm0, N, B, 1, device=dev)
    y_d = torch.from_numpy(y).to(dev)
    start = torch.from_numpy(np.tile(0.5 * np.log([0.5, 0.5, 0.05, 0.05]), (B, 1))).to(dev)
    ll0, _ = api.loglik_grad_batch_dev(y_d, model, start, BSM_Q_INDEX, BSM_P_INDEX)
    ll_true, _ = api.loglik_grad_batch_dev(y_d, model, torch.from_numpy(truth).to(dev), BSM_Q_INDEX, BSM_P_INDEX)
    res = fit.fit_batch(y_d, model, start, BSM_Q_INDEX, BSM_P_INDEX, max_iter=80)
    ll = res.loglik.cpu().numpy()
    assert np.all(np.isfinite(ll)) and np.all(ll >= ll0.cpu().numpy() - 1e-12)
    assert np.mean(ll >= ll_true.cpu().numpy() - 1e-6) > 0.9
    # a variance that wants to be zero runs off to theta -> -inf along a flat ridge: judge convergence by the value
    for b in (0, 7, 19):
        yb = y[:, b:b + 1]
        obj = lambda th: -cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(th).args())
        ref_opt = minimize(obj, start[b].cpu().numpy(), method="BFGS", options={"gtol": 1e-6, "eps": 1e-3})
        assert ll[b] >= -ref_opt.fun - 5e-5 * max(1.0, abs(ref_opt.fun)), (b, ll[b], -ref_opt.fun)
    assert res.evaluations == (1 + 2 * k) * (res.iterations + 1) + 7 * res.iterations
    # every series carries an outcome code; `converged` means the gradient test was met, nothing else
    st = res.status.cpu().numpy()
    assert set(st.tolist()) <= {fit.CONVERGED, fit.ITERATION_LIMIT, fit.LINE_SEARCH_FAILED}
    assert np.array_equal(res.converged.cpu().numpy(), st == fit.CONVERGED)
    gmax = res.grad.abs().amax(1).cpu().numpy()
    assert np.all(gmax[st == fit.CONVERGED] < 1e-5)
    # multistart: three start vectors per series advance in lock-step in the same launches; the winner per series is at
    # least as good as the single start from above (which is one of the three), and every start's value is reported
    starts = torch.stack([start, start + 0.7, torch.from_numpy(truth).to(dev)])
    ms = fit.fit_batch(y_d, model, starts, BSM_Q_INDEX, BSM_P_INDEX, max_iter=80)
    assert ms.loglik_starts.shape == (3, B) and ms.start.shape == (B,)
    best = ms.loglik.cpu().numpy()
    assert np.allclose(best, ms.loglik_starts.max(0).values.cpu().numpy(), rtol=0, atol=0)
    assert np.all(best >= ll - 2e-5 * np.maximum(1.0, np.abs(ll)))
    chk, _ = api.loglik_grad_batch_dev(y_d, model, ms.theta, BSM_Q_INDEX, BSM_P_INDEX)    # theta and value belong together
    assert np.allclose(chk.cpu().numpy(), best, rtol=1e-12, atol=1e-12)
    bad = start.clone()
    bad[3] = float("nan")                                        # a start the objective cannot be evaluated at
    rb = fit.fit_batch(y_d, model, bad, BSM_Q_INDEX, BSM_P_INDEX, max_iter=5)
    assert int(rb.status[3]) == fit.NON_FINITE and not bool(rb.converged[3])


def test_hessian_batch():
    """One-launch central-difference Hessian (the sweep behind statespacer()'s standard errors) against the same
    differences of the oracle objective."""
    import torch
    from statespacer_b200 import fit
    rng = np.random.default_rng(5)
    B, N, k, h = 6, 90, 4, 1e-2
    thetas, _, _, y = _bsm_batch(rng, B, N, 1, 0.0)
    theta = thetas[0]
    dev = torch.device("cuda")
    model = api.BatchModel(sysmat.bsm12(theta[0]), N, B, 1, device=dev)
    Hm = fit.hessian_batch(torch.from_numpy(y).to(dev), model, torch.from_numpy(theta).to(dev), BSM_Q_INDEX,
                           BSM_P_INDEX, h=h).cpu().numpy()
    assert np.allclose(Hm, Hm.transpose(0, 2, 1))
    b = 3
    yb = y[:, b:b + 1]
    f = lambda th: cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(th).args())
    E = np.eye(k) * h
    for i in range(k):
        want = (f(theta[b] + E[i]) - 2 * f(theta[b]) + f(theta[b] - E[i])) / h ** 2
        assert abs(Hm[b, i, i] - want) <= 1e-6 * max(1.0, abs(want)), (i, Hm[b, i, i], want)
    want = (f(theta[b] + E[0] + E[2]) - f(theta[b] + E[0] - E[2]) - f(theta[b] - E[0] + E[2])
            + f(theta[b] - E[0] - E[2])) / (4 * h * h)
    assert abs(Hm[b, 0, 2] - want) <= 1e-6 * max(1.0, abs(want))


def _block_diagonal_model(rng, m, p, d, N, scatter=True):
    """Random model whose T is block diagonal with blocks of size 1-3 scattered over the state vector, so
    that the tensor kernel has to re-order the states to reach a tile-diagonal T."""
    sm = random_model(rng, m, p, d=d, N=N)
    order = rng.permutation(m) if scatter else np.arange(m)
    T = np.zeros((m, m))
    i = 0
    while i < m:
        k = min(m - i, int(rng.integers(1, 4)))
        idx = order[i:i + k]
        blk = rng.normal(size=(k, k)) * 0.4 + 0.6 * np.eye(k)
        blk *= 0.95 / max(0.95, np.max(np.abs(np.linalg.eigvals(blk))))      # keep the recursion well conditioned
        T[np.ix_(idx, idx)] = blk
        i += k
    return sysmat.SysMat(Z=sm.Z, T=T[:, :, None], R=sm.R, Q=sm.Q, a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star)


def test_hessian_richardson_batch():
    """fit.hessian_richardson_batch (the numDeriv::hessian sweep of R/statespacer.R:1057-1062, 81 parameter vectors per
    series in one launch) against the restated numDeriv algorithm driven by the oracle objective."""
    import torch
    from oracle import numderiv_np
    from statespacer_b200 import fit
    rng = np.random.default_rng(17)
    B, N = 6, 120
    thetas, _, _, y = _bsm_batch(rng, B, N, 1, 0.02)
    theta = thetas[0]
    dev = torch.device("cuda", 0)
    model = api.BatchModel(sysmat.bsm12(theta[0]), N, B, 1, device=dev)
    Hm = fit.hessian_richardson_batch(torch.tensor(y, device=dev), model, torch.tensor(theta, device=dev), BSM_Q_INDEX,
                                      BSM_P_INDEX).cpu().numpy()
    for b in (0, 3, 5):
        yb = y[:, b:b + 1]
        f = lambda th: cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(th).args())
        want = numderiv_np.hessian(f, theta[b])
        assert np.max(np.abs(Hm[b] - want)) <= 1e-6 * max(1.0, np.max(np.abs(want))), (b, Hm[b], want)
        assert np.allclose(Hm[b], Hm[b].T)


@pytest.mark.parametrize("m,p", [(9, 1), (10, 2), (14, 1), (15, 1), (16, 2), (17, 1), (20, 1), (23, 3)])
def test_batch_tile_sparse_state_order(m, p):
    """Tensor kernel with the states re-ordered internally (block-diagonal T -> all-zero tiles skipped) against
    the same kernel in the caller's order with every tile multiplied, and against the oracle."""
    rng = np.random.default_rng(700 + m)
    B, N = 10, 40
    sm = _block_diagonal_model(rng, m, p, d=min(3, m - 1), N=N)
    perm, mask, n_dmma = api.plan_state_order(sm.T[:, :, 0])
    tt = (m + 1 + 7) // 8
    use = 0x9 if tt == 2 else 0x111                      # the tile-diagonal instantiation
    assert mask & ~use == 0, hex(mask)
    y = np.stack([simulate_y(rng, sm, N, missing=0.1) for _ in range(B)], axis=2)
    Qb = np.stack([sm.Q[:, :, 0] * (1 + 0.1 * b) for b in range(B)])
    got = api.loglik_batch(y, sm, Q=Qb, kernel="mma")
    assert f"tiles={hex(use)}," in api.last_kernel(), api.last_kernel()
    dense = api.loglik_batch(y, sm, Q=Qb, kernel="mma_dense")
    assert f"tiles={hex((1 << tt * tt) - 1)}," in api.last_kernel()
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=Qb[b][:, :, None], a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star)
        yb = y[:, :, b]
        want, tol = loglik_want(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= tol, (m, b, got[b], want)
        assert abs(dense[b] - want) <= tol, (m, b, dense[b], want)


@pytest.mark.parametrize("m,p,tvz,block", [(9, 1, True, False), (24, 1, False, False), (31, 2, True, True),
                                             (34, 2, True, True), (40, 1, False, True), (47, 3, True, False),
                                             (56, 1, False, True), (63, 2, True, True)])
def test_batch_wsm_kernel(m, p, tvz, block):
    """Warp-per-evaluation kernel with the covariance in shared memory (24 <= m <= 63, or time-varying Z at any m):
    dense and block-diagonal T (re-ordered to tile-sparse), per-series Q / P_star / a, diffuse states, missing
    values, against the oracle and the generic kernel."""
    rng = np.random.default_rng(900 + m)
    B, N = 6, 24
    d = 3
    sm = _block_diagonal_model(rng, m, p, d=d, N=N) if block else random_model(rng, m, p, d=d, N=N)
    if tvz:
        Zt = np.stack([sm.Z[:, :, 0] + 0.05 * rng.normal(size=(p, m)) for _ in range(N)], axis=2)
        sm = sysmat.SysMat(Z=Zt, T=sm.T, R=sm.R, Q=sm.Q, a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star)
    y = np.stack([simulate_y(rng, sm, N, missing=0.1) for _ in range(B)], axis=2)
    Qb = np.stack([sm.Q[:, :, 0] * (1 + 0.1 * b) for b in range(B)])
    Pb = np.stack([sm.P_star * (1 + 0.05 * b) for b in range(B)])
    ab = np.stack([sm.a + 0.01 * b for b in range(B)])
    got = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab, kernel="wsm")
    assert api.last_kernel().startswith(f"loglik_wsm_kernel<TT={(m + 8) // 8},"), api.last_kernel()
    gen = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab, kernel="generic")
    if m > 23 or tvz:
        auto = api.loglik_batch(y, sm, Q=Qb, P_star=Pb, a=ab)
        assert api.last_kernel().startswith("loglik_wsm") and np.array_equal(auto, got)
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=Qb[b][:, :, None], a=ab[b], P_inf=sm.P_inf, P_star=Pb[b])
        yb = y[:, :, b]
        want, tol = loglik_want(yb, np.isnan(yb), *smb.args())
        assert abs(got[b] - want) <= tol, (m, b, got[b], want)
        assert abs(gen[b] - want) <= tol, (m, b, gen[b], want)


def test_batch_multivariate_dns():
    """Config 3 shape (p = 8, m = 14) replicated with noise; all matrices shared."""
    fed = load_fed()
    sm, _ = sysmat.dns_yield_curve(DNS_PARAM)
    rng = np.random.default_rng(1)
    B = 12
    y = np.stack([fed + (0.01 * rng.normal(size=fed.shape) if b else 0.0) for b in range(B)], axis=2)
    y[rng.uniform(size=y.shape) < 0.03] = np.nan
    y[:, :, 0] = fed
    auto = api.loglik_batch(y, sm, kernel="auto")
    assert api.last_kernel().startswith("loglik_cols")                     # p = 8: the column kernel is preferred
    got = api.loglik_batch(y, sm, kernel="mma")
    assert api.last_kernel().startswith("loglik_mma")
    assert abs(got[0] - 7.005830) < 6e-7 * 7                               # docs/articles/selfspec.html:306-307
    cols = api.loglik_batch(y, sm, kernel="cols")
    assert np.array_equal(auto, cols)
    for b in range(B):
        yb = y[:, :, b]
        want = cport.LogLikC(yb, np.isnan(yb), *sm.args())
        assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (b, got[b], want)
        assert abs(cols[b] - want) <= RTOL * max(1, abs(want)), (b, cols[b], want)


def _residual_layout_model(rng, p, mc, d, r_c=None, h_zero=(), n_det=0):
    """A model in the reference's own state layout (R/GetSysMat.R:1388-1426): [p residuals ; mc component states],
    dense T_c, Z_c, R_c, Q_c, d diffuse component states, diagonal H (entries in h_zero set to 0)."""
    r_c = r_c or mc
    T = rng.normal(size=(mc, mc)) * (0.5 / math.sqrt(mc))
    T[np.arange(mc), np.arange(mc)] += 0.6
    T[:d, :] = 0.0
    T[:, :d] = 0.0
    T[np.arange(d), np.arange(d)] = 1.0                                   # diffuse states: random walks
    if d >= 2:
        T[0, 1] = 1.0                                                     # ... the first one with a slope
    Z = rng.normal(size=(p, mc)) * (rng.uniform(size=(p, mc)) < 0.8)
    Z[:, :1] = 1.0
    A = rng.normal(size=(r_c, r_c))
    Q = A @ A.T / r_c + 0.05 * np.eye(r_c)
    R = np.eye(mc)[:, :r_c] if r_c <= mc else rng.normal(size=(mc, r_c)) * 0.5
    C = rng.normal(size=(mc, mc))
    P_star = C @ C.T / mc
    P_star[:d, :] = 0.0
    P_star[:, :d] = 0.0
    P_inf = np.zeros((mc, mc))
    P_inf[np.arange(d), np.arange(d)] = 1.0
    H = np.diag(rng.uniform(0.05, 0.5, p))
    for j in h_zero:
        H[j, j] = 0.0
    a = rng.normal(size=mc)
    if n_det:
        # deterministic states behind the stochastic ones: known start, no disturbance, driven by each other only, feeding
        # the stochastic states and the observations (the means of a VAR, fixed effects)
        Tdd = np.eye(n_det) * 0.97 + 0.02 * rng.normal(size=(n_det, n_det))
        Tsd = rng.normal(size=(mc, n_det)) * 0.2
        T = np.block([[T, Tsd], [np.zeros((n_det, mc)), Tdd]])
        Z = np.concatenate([Z, rng.normal(size=(p, n_det))], axis=1)
        R = np.concatenate([R, np.zeros((n_det, R.shape[1]))], axis=0)
        P_star = sysmat.block_diag(P_star, np.zeros((n_det, n_det)))
        P_inf = sysmat.block_diag(P_inf, np.zeros((n_det, n_det)))
        a = np.concatenate([a, rng.normal(size=n_det)])
    comp = dict(Z=Z, T=T, R=R, Q=Q, a=a, P_inf=P_inf, P_star=P_star)
    return sysmat._combine(p, H, [comp])


@pytest.mark.parametrize("p,mc,d,n_det", [(1, 1, 0, 0), (1, 3, 1, 0), (2, 2, 2, 0), (3, 4, 0, 0), (3, 5, 2, 0), (8, 6, 0, 0),
                                          (4, 6, 3, 0), (2, 7, 1, 0), (5, 8, 2, 0), (16, 3, 1, 0),
                                          (3, 4, 0, 2), (2, 3, 1, 3), (4, 5, 2, 6), (8, 3, 0, 3)])
def test_batch_core_kernel(p, mc, d, n_det):
    """Thread-per-series kernel on the component states (residual states eliminated, loglik_core.cu) against the oracle
    run on the FULL state: dense T_c / Q_c, exact diffuse phase with d diffuse states, missing observations, a series
    that is missing throughout a stretch spanning the diffuse phase, one observation without noise (H_jj = 0); n_det
    deterministic component states (known start, no disturbance) that the launcher takes out of the recursion."""
    rng = np.random.default_rng(1000 + 37 * p + 5 * mc + d + 101 * n_det)
    N, B = 45, 40
    sm = _residual_layout_model(rng, p, mc, d, h_zero=(p - 1,) if p >= 3 else (), n_det=n_det)
    y = np.stack([simulate_y(rng, sm, N, missing=0.08) for _ in range(B)], axis=2)
    y[:6, :, 1] = np.nan                                                   # diffuse phase starts late
    y[:, 0, 2] = np.nan                                                    # first observed variable never seen
    y[3:, :, 3] = np.nan                                                   # too few observations to leave the diffuse phase
    got = api.loglik_batch(y, sm, kernel="core")
    # the n_det deterministic states leave the recursion (tabulated offsets): the kernel runs on mc states
    assert api.last_kernel().startswith(f"loglik_core_kernel<MC={mc}"), api.last_kernel()
    assert ("deterministic states tabulated" in api.last_kernel()) == (n_det > 0)
    other = api.loglik_batch(y, sm, kernel="generic")
    for b in range(B):
        yb = y[:, :, b]
        want, tol = loglik_want(yb, np.isnan(yb), *sm.args())
        if math.isnan(want):
            assert math.isnan(got[b]), (b, got[b])
            continue
        assert abs(got[b] - want) <= tol, (p, mc, d, b, got[b], want)
        assert abs(other[b] - want) <= tol


def test_batch_core_kernel_dns_and_dispatch():
    """BASELINE.json config 3 (dynamic Nelson-Siegel on the FedYieldCurve window, p = 8, m = 14, no diffuse states): a
    6 x 6 problem per thread.  KAT-3 through the kernel, every series against the oracle, automatic choice from 2048
    series on, refusal of what does not have the residual structure."""
    fed = load_fed()
    sm, _ = sysmat.dns_yield_curve(DNS_PARAM)
    rng = np.random.default_rng(5)
    B = 2048
    y = np.stack([fed + (0.01 * rng.normal(size=fed.shape) if b else 0.0) for b in range(B)], axis=2)
    y[rng.uniform(size=y.shape) < 0.03] = np.nan
    y[:, :, 0] = fed
    got = api.loglik_batch(y, sm)
    # 14 states = 8 residuals + 3 factors + their 3 constant means: a 3 x 3 problem per thread
    assert api.last_kernel().startswith("loglik_core_kernel<MC=3>") and "deterministic states tabulated" in api.last_kernel()
    assert abs(got[0] - 7.005830) < 6e-7 * 7                               # docs/articles/selfspec.html:306-307
    assert np.array_equal(got, api.loglik_batch(y, sm, kernel="core"))
    cols = api.loglik_batch(y[:, :, :64], sm)
    assert api.last_kernel().startswith("loglik_cols")                     # small batches stay on the lane kernels
    assert np.allclose(cols, got[:64], rtol=1e-11, atol=0)
    for b in list(range(24)) + [B - 1]:
        yb = y[:, :, b]
        want = cport.LogLikC(yb, np.isnan(yb), *sm.args())
        assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (b, got[b], want)
    # the local level (config 1 batched): AUTO takes it to the p = 1 instantiation from 2048 series on
    nile = sysmat.nile_local_level(15100.252, 1468.724)
    yl = np.cumsum(rng.normal(size=(100, 2048)), axis=0) * 40.0 + 900.0
    yl[rng.uniform(size=yl.shape) < 0.05] = np.nan
    yl[:, 7] = np.nan
    gl = api.loglik_batch(yl, nile)
    assert api.last_kernel().startswith("loglik_core_kernel<MC=1,diffuse,p1>"), api.last_kernel()
    lanes = api.loglik_batch(yl, nile, kernel="blk")
    assert api.last_kernel().startswith("loglik_blk_kernel<L=1,NB=1,")
    assert np.isnan(gl[7]) and np.allclose(gl, lanes, rtol=1e-12, atol=0, equal_nan=True)
    for b in range(0, 2048, 199):
        yb = yl[:, b:b + 1]
        want = cport.LogLikC(yb, np.isnan(yb), *nile.args())
        assert (math.isnan(want) and math.isnan(gl[b])) or abs(gl[b] - want) <= RTOL * max(1, abs(want)), (b, gl[b], want)
    # correlated observation noise (H not diagonal) couples the residual states: refused, AUTO moves on
    bad = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q.copy(), a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star.copy())
    bad.Q[0, 1, 0] = bad.Q[1, 0, 0] = 1e-4
    bad.P_star[0, 1] = bad.P_star[1, 0] = 1e-4
    with pytest.raises(native.SsgpuError):
        api.loglik_batch(y[:, :, :8], bad, kernel="core")
    auto = api.loglik_batch(y, bad)
    assert not api.last_kernel().startswith("loglik_core")
    yb = y[:, :, 5]
    want, tol = loglik_want(yb, np.isnan(yb), *bad.args())
    assert abs(auto[5] - want) <= tol


def test_batch_generic_time_varying_and_per_series():
    rng = np.random.default_rng(3)
    N, B, m, p = 20, 6, 7, 2
    sm = random_model(rng, m, p, d=1, tv=("Z", "T"), N=N)
    y = np.stack([simulate_y(rng, sm, N, missing=0.1) for _ in range(B)], axis=2)
    got = api.loglik_batch(y, sm)
    assert api.last_kernel().startswith("fwd_generic")
    for b in range(B):
        want, tol = loglik_want(y[:, :, b], np.isnan(y[:, :, b]), *sm.args())
        assert abs(got[b] - want) <= tol


def test_batch_properties_large():
    """Size-independent properties at a size the oracle would not finish: (i) replicated series give
    identical values, (ii) scaling y by c and all variances by c^2 shifts loglik by -(n_obs - d)*log(c)/N
    (the d = 13 exact-diffuse steps contribute -log(F_inf)/2, which does not scale: src/LogLikC.cpp:153)."""
    rng = np.random.default_rng(9)
    B, N, V = 20000, 100, 1
    thetas, Qd, Pd, y = _bsm_batch(rng, 64, N, V, 0.02)
    reps = B // 64 + 1
    yB = np.tile(y, (1, reps))[:, :B].copy()
    QB = np.tile(Qd[0], (reps, 1))[:B].copy()
    PB = np.tile(Pd[0], (reps, 1))[:B].copy()
    sm = sysmat.bsm12(thetas[0, 0])
    got = api.loglik_batch(yB, sm, Q_diag=QB, P_star_diag=PB)
    assert np.array_equal(got[:64], got[64:128]) and np.array_equal(got[:B % 64], got[B - B % 64:])
    c = 3.0
    got_c = api.loglik_batch(c * yB, sm, Q_diag=c * c * QB, P_star_diag=c * c * PB)
    n_obs = np.sum(~np.isnan(yB), axis=0)
    assert np.max(np.abs(got_c - (got - (n_obs - 13) * math.log(c) / N))) < 1e-9 * np.max(np.abs(got))
    want = cport.LogLikC(y[:, 5:6], np.isnan(y[:, 5:6]), *sysmat.bsm12(thetas[0, 5]).args())
    assert abs(got[5] - want) <= RTOL * abs(want)


def test_full_size_block_row_vs_tensor_kernel():
    """BASELINE.json config 5 at its full per-GPU size (125,000 series x 9 parameter vectors x T = 1000, the bench's
    own synthetic shard): the thread-per-evaluation kernel (default), the lane kernel and the tensor kernel -- three independent implementations of the
    recursion -- agree to 1e-9 on every loglikelihood and on the central-difference gradients, and a sample of series
    agrees with the oracle."""
    import torch
    import bench
    dev = torch.device("cuda", 0)
    B = bench.B_PER_GPU
    sm, y, theta = bench.make_shard(torch, dev, B, 0)
    y[torch.rand(y.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) < 0.01] = float("nan")
    model = api.BatchModel(sm, bench.T_STEPS, B, 1, device=dev)
    out = {}
    for kernel in ("blk", "blk_lanes", "mma"):
        out[kernel] = api.loglik_grad_batch_dev(y, model, theta, bench.Q_INDEX, bench.P_INDEX, h=bench.H_FD, kernel=kernel)
    ll_b, gr_b = out["blk"]
    ll_l, gr_l = out["blk_lanes"]
    ll_m, gr_m = out["mma"]
    assert bool(torch.isfinite(ll_b).all()) and bool(torch.isfinite(gr_b).all())
    assert float(((ll_b - ll_m).abs() / ll_m.abs().clamp(min=1.0)).max()) < RTOL
    assert float(((ll_l - ll_m).abs() / ll_m.abs().clamp(min=1.0)).max()) < RTOL
    assert float((gr_l - gr_m).abs().max()) < 1e-6
    # gradients are differences of loglik / N at h = 1e-3: 1e-9 on the values is 1e-6 on the quotient
    assert float((gr_b - gr_m).abs().max()) < 1e-6
    yh, th = y[:, :4].cpu().numpy(), theta[:4].cpu().numpy()
    for b in range(4):
        yb = yh[:, b:b + 1]
        want = cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(th[b]).args())
        assert abs(float(ll_b[b]) - want) <= RTOL * max(1, abs(want))


# ---- simulation smoother -----------------------------------------------------------------------------
def _sim_case(rng, m, p, r, N, tv=()):
    sm = random_model(rng, m, p, r=r, d=0, tv=tv, N=N)
    return sm


@pytest.mark.parametrize("repeat_Q,draw_initial,eta_only,tv", [(1, False, False, ()), (3, True, False, ("T",)),
                                                               (1, False, True, ()), (2, True, False, ("Z", "R"))])
def test_simulate_bit_exact(repeat_Q, draw_initial, eta_only, tv):
    rng = np.random.default_rng(3)
    N, nsim, r = 15, 37, 2
    r_tot = r * repeat_Q
    m, p = 5, 2
    sm = random_model(rng, m, p, r=r_tot, N=N, tv=tv)
    A = rng.normal(size=(r, r))
    Q = A @ A.T + 0.1 * np.eye(r)
    n = (m * nsim if draw_initial else 0) + N * r_tot * nsim
    draws = rng.normal(size=n)
    P = sm.P_star + 0.01 * np.eye(m)
    Q_root = cport.sym_root(Q)[:, :, None]
    P_root = cport.sym_root(P)
    if eta_only:
        Z = T = R = np.zeros((1, 1, 1))
        a = np.zeros(1)
        n = N * r_tot * nsim
        draws = draws[:n]
        args = (nsim, repeat_Q, N, a, Z, T, R, Q, np.zeros((1, 1)), False, True, True, draws)
    else:
        args = (nsim, repeat_Q, N, sm.a, sm.Z, sm.T, sm.R, Q, P, draw_initial, False, True, draws)
    want = cport.SimulateC(*args, Q_root=Q_root, P_root=P_root)
    got = api.SimulateC(*args, Q_root=Q_root, P_root=P_root)
    keys = ("eta",) if eta_only else ("y", "a", "a_t", "eta")
    for k in keys:
        assert np.array_equal(got[k], want[k]), k                         # bit-exact
    if ref.available():
        r0 = ref.SimulateC(*args)
        got2 = api.SimulateC(*args)                                       # library's own Jacobi roots
        for k in keys:
            assert rel(got2[k], r0[k]) < 1e-10, k


def test_simulate_wrong_draw_count():
    rng = np.random.default_rng(0)
    sm = random_model(rng, 3, 1, r=2, N=5)
    with pytest.raises(native.SsgpuError) as e:
        api.SimulateC(4, 1, 5, sm.a, sm.Z, sm.T, sm.R, np.eye(2), sm.P_star, False, False, False, np.zeros(7))
    assert e.value.status == 3


@pytest.mark.parametrize("shape", [dict(m=2, p=1, d=1, tv=()), dict(m=5, p=2, d=2, tv=("Z",)),
                                   dict(m=8, p=3, d=3, tv=("T", "Q")), dict(m=28, p=1, d=0, tv=()),
                                   dict(m=6, p=2, d=0, tv=("R",))])
def test_fast_smoother(shape):
    rng = np.random.default_rng(5)
    N, nsim = 30, 41
    sm = random_model(rng, shape["m"], shape["p"], d=shape["d"], tv=shape["tv"], N=N)
    c = Case("fs", simulate_y(rng, sm, N, missing=0.0), sm)
    steps = kalman_np.KalmanC(*c.args(), False)["initialisation_steps"]
    y3 = np.stack([c.y + 0.1 * rng.normal(size=c.y.shape) * (s > 0) for s in range(nsim)], axis=2)
    want = cport.FastSmootherC(y3, *sm.args(), steps, True)
    got = api.FastSmootherC(y3, *sm.args(), steps, True)
    for k in ("a_smooth", "a_t", "eta"):
        assert np.array_equal(got[k], want[k]), (k, rel(got[k], want[k]))  # bit-exact vs the pinned-order oracle
    if ref.available():
        r0 = ref.FastSmootherC(y3, *sm.args(), steps, True)
        for k in ("a_smooth", "a_t", "eta"):
            assert_entries(k, got[k], r0[k], ref_scale=float(np.max(np.abs(r0["a_smooth"]))) if k != "eta" else None)
    # reference-internal identity: draw 0 is the data -> KalmanC's smoothed state
    ks = api.KalmanC(*c.args(), False)
    assert_entries("a_smooth of draw 0", got["a_smooth"][:, :, 0], ks["a_smooth"])


def test_fast_smoother_airline_sim_smoother_pipeline():
    """Config 4 in small: SimulateC -> FastSmootherC -> mean correction -> ExtractComponentC
    (R/SimSmoother.R:53-769) on the airline model, bit-exact against the C oracle."""
    air = np.log(load_series("airpassengers.txt"))[:, None]
    sm = sysmat.airline([0.5 * math.log(0.00135), 0.4, 0.55])
    N, nsim = len(air), 64
    steps = api.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"]
    rng = np.random.default_rng(4)
    comp_m = sm.m - 1
    Qc = sm.Q[1:, 1:, :]
    draws = rng.normal(size=N * 1 * nsim)
    Zc, Tc, Rc = sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :]
    Q_root = cport.sym_root(Qc[:, :, 0])[:, :, None]
    sims = {}
    for name, mod in (("gpu", api), ("cpu", cport)):
        s = mod.SimulateC(nsim, 1, N, np.zeros(comp_m), Zc, Tc, Rc, Qc, np.zeros((comp_m, comp_m)), False, False,
                          True, draws, Q_root=Q_root)
        fs = mod.FastSmootherC(s["y"], *sm.args(), steps, True)
        comp = mod.ExtractComponentC(fs["a_t"], sm.Z)
        sims[name] = (s, fs, comp)
    for k in ("y", "a", "a_t", "eta"):
        assert np.array_equal(sims["gpu"][0][k], sims["cpu"][0][k]), k
    for k in ("a_smooth", "a_t", "eta"):
        assert np.array_equal(sims["gpu"][1][k], sims["cpu"][1][k]), k
    assert np.array_equal(sims["gpu"][2], sims["cpu"][2])


@pytest.mark.parametrize("case", ["bsm12", "random_tv"])
def test_kalman_batch(case):
    """Batched KalmanC (filter + smoother of B series in one call, only the requested arrays returned) against the
    single-series entry point and the compiled reference."""
    rng = np.random.default_rng(8)
    if case == "bsm12":
        B, N = 19, 40
        thetas, Qd, Pd, y = _bsm_batch(rng, B, N, 1, 0.08)
        sm = sysmat.bsm12(thetas[0, 0])
        per = dict(Q_diag=Qd[0], P_star_diag=Pd[0])
        sms = [sysmat.bsm12(thetas[0, b]) for b in range(B)]
        y3 = y[:, None, :]
    else:
        B, N = 7, 30
        sm = random_model(rng, 6, 2, d=2, tv=("Z", "Q"), N=N)
        y3 = np.stack([simulate_y(rng, sm, N, missing=0.1) for _ in range(B)], axis=2)
        ab = rng.normal(size=(B, 6))
        per = dict(a=ab)
        sms = [sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=ab[b], P_inf=sm.P_inf, P_star=sm.P_star) for b in range(B)]
    want = ("loglik", "a_smooth", "V", "a_fil", "eta", "eta_var", "r_vec", "Nmat", "P_fc", "a_fc", "v", "Fmat")
    got = api.kalman_batch(y3, sm, want=want, **per)
    for b in range(B):
        yb = y3[:, :, b]
        one = api.KalmanC(yb, np.isnan(yb), *sms[b].args(), False)
        assert got["initialisation_steps"][b] == one["initialisation_steps"]
        for k in want:
            assert np.array_equal(got[k][b], np.asarray(one[k]).reshape(got[k][b].shape), equal_nan=True), (case, b, k)
        if ref.available() and b % 3 == 0:
            r0 = ref.KalmanC(yb, np.isnan(yb), *sms[b].args(), False)
            alt = kalman_np.KalmanC(yb, np.isnan(yb), *sms[b].args(), False)
            kalman_compare((case, b), one, r0, alt, [k for k in want if k in KALMAN_KEYS], yb.shape[1])
    lean = api.kalman_batch(y3, sm, want=("a_smooth",), **per)              # intermediates kept as device scratch
    assert np.array_equal(lean["a_smooth"], got["a_smooth"])
    with pytest.raises(RuntimeError, match="diagnostics"):
        api.kalman_batch(y3, sm, want=("a_smooth", "v_norm"), **per)


@pytest.mark.parametrize("tv", [(), ("H",), ("Z",), ("H", "Z")])
def test_collapse_transform(tv):
    """ssgpu_collapse (R/statespacer.R:815-900) against the NumPy restatement for time-invariant and time-varying H / Z,
    missing observations and all-zero collapsed values; KAT-3 through the collapsed model on the GPU."""
    from oracle import collapse_np
    rng = np.random.default_rng(21)
    N, p, m2 = 50, 7, 3
    mk = lambda base, name, f: np.stack([f(base, i) for i in range(N)], axis=2) if name in tv else base[:, :, None]
    A = rng.normal(size=(p, p))
    H = mk(A @ A.T / p + np.eye(p), "H", lambda B, i: B * (1.0 + 0.2 * math.sin(i)))
    Z = mk(rng.normal(size=(p, m2)), "Z", lambda B, i: B + 0.05 * math.cos(i))
    y = rng.normal(size=(N, p)) * 2
    y[rng.uniform(size=y.shape) < 0.1] = np.nan
    y[7] = np.nan                                                       # a time point without observations: y* = 0 -> NA
    ys, Hs, la = api.collapse(y, H, Z)
    ys0, Hs0, la0 = collapse_np.collapse(y, H, Z)
    assert rel(ys, ys0) < RTOL and np.all(np.isnan(ys[7]))
    assert rel(Hs, Hs0) < RTOL
    assert abs(la - la0) <= RTOL * max(1, abs(la0))
    if not tv:
        fed = load_fed()
        sm, _ = sysmat.dns_yield_curve(DNS_PARAM)
        yk, Hk, lk = api.collapse(fed, sm.Q[:8, :8, 0], sm.Z[:, 8:11, 0])
        smc = sysmat.dns_collapsed(DNS_PARAM, Hk[:, :, 0])
        ll = api.LogLikC(yk, np.isnan(yk), *smc.args()) + lk / len(fed)
        assert abs(ll - 7.005830) < 6e-7 * 7                            # docs/articles/selfspec.html:306-307
    with pytest.raises(RuntimeError, match="m2 < p"):
        api.collapse(y[:, :3], H[:3, :3], Z[:3])


@pytest.mark.parametrize("tv", [False, True])
def test_forecast_recursion(tv):
    """ssgpu_forecast (R/predict.R:205-246) started from the GPU KalmanC's a_fc / P_fc, against the NumPy restatement."""
    from oracle import forecast_np
    rng = np.random.default_rng(33)
    h, N = 12, 40
    sm = random_model(rng, 7, 2, d=2, N=N, tv=())
    y = simulate_y(rng, sm, N, missing=0.05)
    ks = api.KalmanC(y, np.isnan(y), *sm.args(), False)
    rep = (lambda X: np.stack([X[:, :, 0] * (1 + 0.05 * i) for i in range(h)], axis=2)) if tv else (lambda X: X)
    Z, T, R, Q = rep(sm.Z), rep(sm.T), rep(sm.R), rep(sm.Q)
    H = rep((np.eye(2) * 0.3)[:, :, None])
    got = api.forecast(h, ks["a_fc"], ks["P_fc"], Z, T, R, Q, H)
    want = forecast_np.forecast(h, ks["a_fc"], ks["P_fc"], Z, T, R, Q, H)
    for g, w in zip(got, want):
        assert rel(g, w) < RTOL


def _sim_smoother_case(kind, rng):
    """Inputs of one SimSmoother() call: the components as R/SimSmoother.R would pass them to SimulateC, the full model
    with the residual states in front and the smoothed estimates of the observed data."""
    S = sysmat
    if kind == "airline":
        air = np.log(load_series("airpassengers.txt"))[:, None]
        sm = S.airline([0.5 * math.log(0.00135), 0.4, 0.55], H=0.0002)
        c = S.comp_sarima_univariate((12, 1), (0, 0), (1, 1), (1, 1), [0.5 * math.log(0.00135), 0.4, 0.55])
        comps = [dict(a=c["a"], Z=c["Z"], T=c["T"], R=c["R"], Q=c["Q"], P_star=c["P_star"], repeat_Q=1, draw_initial=True)]
        y, p = air, 1
        extract = [c["Z"]]
    elif kind == "bivariate":  # two local levels with correlated disturbances and residuals, p = 2, time-varying H
        H = np.array([[1.0, 0.3], [0.3, 0.5]])
        Ql = np.array([[0.2, -0.05], [-0.05, 0.1]])
        N0 = 50
        Ht = np.stack([H * (1.0 + 0.3 * math.sin(i)) for i in range(N0)], axis=2)
        Hq = Ht[:, :, list(range(1, N0)) + [0]]                           # R/statespacer.R:805: H shifted by one in Q
        sm = S.SysMat(Z=np.concatenate([np.eye(2), np.eye(2)], axis=1)[:, :, None],
                      T=np.diag([0.0, 0.0, 1.0, 1.0])[:, :, None], R=np.eye(4)[:, :, None],
                      Q=np.stack([S.block_diag(Hq[:, :, i], Ql) for i in range(N0)], axis=2), a=np.zeros(4),
                      P_inf=np.diag([0.0, 0.0, 1.0, 1.0]), P_star=S.block_diag(Ht[:, :, 0], np.zeros((2, 2))))
        comps = [dict(a=np.zeros(2), Z=np.eye(2), T=np.eye(2), R=np.eye(2), Q=Ql, P_star=np.zeros((1, 1)), repeat_Q=1,
                      draw_initial=False)]
        y, p = simulate_y(rng, sm, N0, missing=0.1), 2
        extract = [np.eye(2)]
        ks = api.KalmanC(y, np.isnan(y), *sm.args(), False)
        return dict(N=N0, p=2, comps=comps, H=Ht, full=tuple(sm.args()), steps=int(ks["initialisation_steps"]),
                    a_hat=ks["a_smooth"][:, 2:], eps_hat=ks["a_smooth"][:, :2], eta_hat=ks["eta"][:, 2:], extract=extract)
    else:  # level + slope, trigonometric seasonal (repeat_Q = s - 1, R/SimSmoother.R:188) and a cycle, p = 1
        cl, cs, cc = S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(4, 0.05), _cycle(0.9, 0.4, 0.3)
        sm = S._combine(1, [[0.5]], [cl, cs, cc])
        comps = [dict(a=cl["a"], Z=cl["Z"], T=cl["T"], R=cl["R"], Q=cl["Q"], P_star=np.zeros((1, 1)), repeat_Q=1,
                      draw_initial=False),
                 dict(a=cs["a"], Z=cs["Z"], T=cs["T"], R=cs["R"], Q=np.array([[0.05]]), P_star=np.zeros((1, 1)), repeat_Q=3,
                      draw_initial=False),
                 dict(a=cc["a"], Z=cc["Z"], T=cc["T"], R=cc["R"], Q=np.array([[0.3]]), P_star=cc["P_star"], repeat_Q=2,
                      draw_initial=True)]
        y, p = simulate_y(rng, sm, 60, missing=0.1), 1
        pad = lambda Zc, lo, k: np.concatenate([np.zeros((1, lo)), Zc, np.zeros((1, sm.m - 1 - lo - k))], axis=1)
        extract = [pad(cl["Z"], 0, 2), pad(cs["Z"], 2, 3), pad(cc["Z"], 5, 2)]
    ks = api.KalmanC(y, np.isnan(y), *sm.args(), False)
    a_hat, eps_hat, eta_hat = ks["a_smooth"][:, p:], ks["a_smooth"][:, :p], ks["eta"][:, p:]
    H = sm.Q[:p, :p, :]
    return dict(N=len(y), p=p, comps=comps, H=H, full=tuple(sm.args()), steps=int(ks["initialisation_steps"]),
                a_hat=a_hat, eps_hat=eps_hat, eta_hat=eta_hat, extract=extract)


@pytest.mark.parametrize("kind", ["airline", "structural", "bivariate"])
def test_sim_smoother_one_call(kind):
    """ssgpu_SimSmoother (R/SimSmoother.R:34-769 as one call, per-draw arrays resident on the device) against the
    same function strung together from the C oracle's routines: bit-exact with shared symmetric roots, 1e-9 against
    the compiled reference with its own SVD roots."""
    from oracle import sim_smoother
    rng = np.random.default_rng(12)
    c = _sim_smoother_case(kind, rng)
    nsim, N = 37, c["N"]
    n_draws = sum((np.asarray(k["a"]).size * nsim if k["draw_initial"] else 0)
                  + N * np.asarray(k["Q"]).shape[0] * k["repeat_Q"] * nsim for k in c["comps"]) + N * c["p"] * nsim
    draws = rng.normal(size=n_draws)
    roots = [dict(Q_root=np.stack([cport.sym_root(np.atleast_2d(k["Q"]))], axis=2),
                  P_root=cport.sym_root(np.atleast_2d(k["P_star"])) if k["draw_initial"] else None) for k in c["comps"]]
    roots.append({})
    args = (nsim, N, c["p"], c["comps"], c["H"], c["full"], c["steps"], c["a_hat"], c["eps_hat"], c["eta_hat"], draws)
    got = api.SimSmoother(*args, extract=c["extract"], roots=roots)
    want = sim_smoother.SimSmoother(cport, *args, extract=c["extract"], roots=roots)
    # the root of the residual variance is the library's own (Jacobi) against the oracle's (SVD): identical bits for a
    # scalar H, 1e-9 for the 2 x 2 time-varying H of the bivariate case
    same = (lambda a, b: np.array_equal(a, b)) if c["p"] == 1 else (lambda a, b: rel(a, b) < 1e-9)
    for k in ("epsilon", "a", "eta"):
        assert same(got[k], want[k]), k
    assert len(got["components"]) == len(c["extract"])
    for g, w in zip(got["components"], want["components"]):
        assert same(g, w)
    if ref.available():
        r0 = sim_smoother.SimSmoother(ref, *args, extract=c["extract"])
        got2 = api.SimSmoother(*args, extract=c["extract"])                # the library's own Jacobi roots
        for k in ("epsilon", "a", "eta"):
            assert rel(got2[k], r0[k]) < RTOL, k
        for g, w in zip(got2["components"], r0["components"]):
            assert rel(g, w) < RTOL
    with pytest.raises(RuntimeError, match="n_draws"):
        api.SimSmoother(*args[:-1], draws[:-1], extract=c["extract"])


def test_simulate_model_predict_simulations():
    """ssgpu_SimulateModel (predict()'s simulations over the forecast period, R/predict.R:142-856) against the same
    calls strung together from the C oracle: bit-exact; the component extraction equals each component's own y."""
    from oracle import sim_smoother
    rng = np.random.default_rng(23)
    c = _sim_smoother_case("structural", rng)
    for k in c["comps"]:                                # predict() starts every component from a_fc / P_fc
        mc = np.asarray(k["a"]).size
        A = rng.normal(size=(mc, mc))
        k["a"], k["P_star"], k["draw_initial"] = rng.normal(size=mc), A @ A.T / mc, True
    nsim, h = 33, 24
    n_draws = sum(np.asarray(k["a"]).size * nsim + h * np.asarray(k["Q"]).shape[0] * k["repeat_Q"] * nsim
                  for k in c["comps"]) + h * c["p"] * nsim
    draws = rng.normal(size=n_draws)
    roots = [dict(Q_root=np.stack([cport.sym_root(np.atleast_2d(k["Q"]))], axis=2),
                  P_root=cport.sym_root(np.atleast_2d(k["P_star"]))) for k in c["comps"]] + [{}]
    got = api.SimulateModel(nsim, h, c["p"], c["comps"], c["H"], draws, extract=c["extract"], roots=roots)
    want = sim_smoother.SimulateModel(cport, nsim, h, c["p"], c["comps"], c["H"], draws, extract=c["extract"], roots=roots)
    for k in ("y", "epsilon", "a", "eta"):
        assert np.array_equal(got[k], want[k]), k
    for g, w, own in zip(got["components"], want["components"], want["own_y"]):
        assert np.array_equal(g, w) and np.array_equal(g, own)


@pytest.mark.parametrize("nsim", [100, 40000, 80000])
def test_sim_smoother_device_resident_pipeline(nsim):
    """The *_dev entry points (draws stay on the GPU in the draw-contiguous layout) give the same bits as the
    host-pointer entry points; nsim picks the 32-, 64- and 128-thread variants of the per-draw kernels.  A small
    prefix of the draws is also checked against the C oracle."""
    import torch
    rng = np.random.default_rng(11)
    N, m, p, r = 12, 6, 2, 3
    sm = random_model(rng, m, p, r=r, d=2, N=N)
    sm.T[:, :, 0][np.abs(sm.T[:, :, 0]) < 0.25] = 0.0            # some structural zeros for the sparse loops
    y_obs = simulate_y(rng, sm, N)
    steps = api.KalmanC(y_obs, np.isnan(y_obs), *sm.args(), False)["initialisation_steps"]
    draws = rng.normal(size=m * nsim + N * r * nsim)
    Q_root = cport.sym_root(sm.Q[:, :, 0])[:, :, None]
    P_root = cport.sym_root(sm.P_star)
    host = api.SimulateC(nsim, 1, N, sm.a, sm.Z, sm.T, sm.R, sm.Q, sm.P_star, True, False, False, draws,
                         Q_root=Q_root, P_root=P_root)
    fs_h = api.FastSmootherC(host["y"], *sm.args(), steps, False)
    dev = api.SimulateC_dev(nsim, 1, N, sm.a, sm.Z, sm.T, sm.R, sm.Q, sm.P_star, True, False,
                            torch.from_numpy(draws).cuda(), Q_root=Q_root, P_root=P_root)
    fs_d = api.FastSmootherC_dev(dev["y"], *sm.args(), steps)
    comp_d = api.ExtractComponentC_dev(fs_d["a_smooth"], sm.Z)
    to_np = lambda t: t.cpu().numpy().transpose(0, 1, 2)          # (N, K, nsim), same index order as R's array
    for k in ("y", "a", "eta"):
        assert np.array_equal(to_np(dev[k]), host[k]), k
    for k in ("a_smooth", "eta"):
        assert np.array_equal(to_np(fs_d[k]), fs_h[k]), k
    comp_h = np.einsum("pk,tks->tps", sm.Z[:, :, 0], fs_h["a_smooth"])
    assert np.allclose(to_np(comp_d), comp_h, rtol=1e-12, atol=1e-12)
    flat = api.draw_layout_dev(fs_d["a_smooth"], True).cpu().numpy()
    assert np.array_equal(flat, fs_h["a_smooth"].ravel(order="F"))
    k = 16                                                        # oracle on the first draws
    sub = np.concatenate([draws[:m * nsim].reshape(nsim, m)[:k].ravel(),
                          draws[m * nsim:].reshape(N, nsim, r)[:, :k].ravel()])
    ref = cport.SimulateC(k, 1, N, sm.a, sm.Z, sm.T, sm.R, sm.Q, sm.P_star, True, False, False, sub, Q_root=Q_root,
                          P_root=P_root)
    assert np.array_equal(ref["y"], host["y"][:, :, :k]) and np.array_equal(ref["a"], host["a"][:, :, :k])
    fs_r = cport.FastSmootherC(ref["y"], *sm.args(), steps, False)
    assert np.array_equal(fs_r["a_smooth"], fs_h["a_smooth"][:, :, :k])
    assert np.array_equal(fs_r["eta"], fs_h["eta"][:, :, :k])


def test_sim_smoother_large_state_shared_memory_kernels():
    """m = 36 > 32: the per-draw kernels that keep the state vectors in shared memory (the register variants stop
    at m = 32), bit-exact against the C oracle."""
    rng = np.random.default_rng(21)
    N, m, p, r, nsim = 9, 36, 2, 4, 70
    sm = random_model(rng, m, p, r=r, d=2, N=N)
    sm.T[:, :, 0][np.abs(sm.T[:, :, 0]) < 0.12] = 0.0
    y_obs = simulate_y(rng, sm, N)
    steps = api.KalmanC(y_obs, np.isnan(y_obs), *sm.args(), False)["initialisation_steps"]
    draws = rng.normal(size=m * nsim + N * r * nsim)
    Q_root = cport.sym_root(sm.Q[:, :, 0])[:, :, None]
    P_root = cport.sym_root(sm.P_star)
    out = {}
    for name, mod in (("gpu", api), ("cpu", cport)):
        s_ = mod.SimulateC(nsim, 1, N, sm.a, sm.Z, sm.T, sm.R, sm.Q, sm.P_star, True, False, True, draws,
                           Q_root=Q_root, P_root=P_root)
        out[name] = (s_, mod.FastSmootherC(s_["y"], *sm.args(), steps, True))
    for k in ("y", "a", "a_t", "eta"):
        assert np.array_equal(out["gpu"][0][k], out["cpu"][0][k]), k
    for k in ("a_smooth", "a_t", "eta"):
        assert np.array_equal(out["gpu"][1][k], out["cpu"][1][k]), k


def test_fast_smoother_lanes_variant_bit_exact():
    """The opt-in four-lanes-per-draw smoother kernel (SSGPU_SIM_LANES=1, read once per process: run in a child) gives
    the same bits as the pinned-order C oracle on the airline pipeline."""
    import subprocess
    import sys
    code = ("import sys, math, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from conftest import load_series\nfrom oracle import cport\nfrom statespacer_b200 import api, sysmat\n"
            "air = np.log(load_series('airpassengers.txt'))[:, None]\n"
            "sm = sysmat.airline([0.5 * math.log(0.00135), 0.4, 0.55])\n"
            "steps = api.KalmanC(air, np.isnan(air), *sm.args(), False)['initialisation_steps']\n"
            "y = np.random.default_rng(3).normal(size=(len(air), 1, 70)).cumsum(axis=0) * 0.05 + air[:, :, None]\n"
            "g = api.FastSmootherC(y, *sm.args(), steps, True)\nw = cport.FastSmootherC(y, *sm.args(), steps, True)\n"
            "assert all(np.array_equal(g[k], w[k]) for k in ('a_smooth', 'a_t', 'eta'))\nprint('lanes ok')\n"
            % (ROOT_DIR, ROOT_DIR + '/tests'))
    env = dict(os.environ, SSGPU_SIM_LANES="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "lanes ok" in r.stdout, r.stderr[-2000:]


def test_extract_component():
    rng = np.random.default_rng(9)
    m, nsim, N, p = 5, 70, 9, 2
    a = rng.normal(size=(m, nsim, N))
    for Z in (rng.normal(size=(p, m, 1)), rng.normal(size=(p, m, N))):
        assert np.array_equal(api.ExtractComponentC(a, Z), cport.ExtractComponentC(a, Z))
    with pytest.raises(ValueError):
        api.ExtractComponentC(a, rng.normal(size=(p, m + 1, 1)))


def test_fp64_peak_probe():
    tf, ms = api.fp64_peak(0, 200, 2)
    assert 5 < tf < 80 and ms > 0


# ---- BASELINE.json config 4 on its own model: SARIMA(1,1,1)(0,1,1)_12, bench.sim_model() --------------------
def _config4():
    """The model the sim workload of bench.py runs (sysmat.airline(..., ar = (0, 1)): one non-seasonal AR coefficient;
    R/Components-BoxJenkins.R:364-482), its component (state without the residual) and the smoothed quantities
    SimSmoother() takes from the fitted object."""
    import bench
    air, sm = bench.sim_model()
    N, mc = len(air), sm.m - 1
    ks = api.KalmanC(air, np.isnan(air), *sm.args(), False)
    comp = dict(Z=sm.Z[:, 1:, :], T=sm.T[1:, 1:, :], R=sm.R[1:, 1:, :], Q=sm.Q[1:, 1:, :], H=sm.Q[:1, :1, :])
    return air, sm, N, mc, ks, comp


def test_config4_exact_model_three_routines_4096_draws():
    """SimulateC (component and residual) -> FastSmootherC -> ExtractComponentC through the reference-shaped entry
    points on config 4's own model at 4,096 draws: every array bit-identical to the pinned-order C oracle, 1e-9 against
    the reference's C++ (compiled unmodified; run in chunks of draws on the host threads -- a draw does not depend on
    its neighbours)."""
    from concurrent.futures import ThreadPoolExecutor
    air, sm, N, mc, ks, c = _config4()
    steps0 = int(ks["initialisation_steps"])
    assert steps0 == 13
    nsim = 4096
    rng = np.random.default_rng(4)
    d1, d2 = rng.normal(size=N * nsim), rng.normal(size=N * nsim)
    one = np.zeros((1, 1, 1))
    Q_root = cport.sym_root(c["Q"][:, :, 0])[:, :, None]
    H_root = cport.sym_root(c["H"][:, :, 0])[:, :, None]

    def run(mod, n, e1, e2, roots=True):
        kw1 = dict(Q_root=Q_root) if roots else {}
        kw2 = dict(Q_root=H_root) if roots else {}
        s1 = mod.SimulateC(n, 1, N, np.zeros(mc), c["Z"], c["T"], c["R"], c["Q"], np.zeros((mc, mc)), False, False, True, e1, **kw1)
        s2 = mod.SimulateC(n, 1, N, np.zeros(1), one, one, one, c["H"], np.zeros((1, 1)), False, True, False, e2, **kw2)
        fs = mod.FastSmootherC(s1["y"] + s2["eta"], *sm.args(), steps0, True)
        return s1, s2, fs, mod.ExtractComponentC(fs["a_t"], sm.Z)

    g1, g2, gfs, gcomp = run(api, nsim, d1, d2)
    w1, w2, wfs, wcomp = run(cport, nsim, d1, d2)
    for k in ("y", "a", "a_t", "eta"):
        assert np.array_equal(g1[k], w1[k]), k
    assert np.array_equal(g2["eta"], w2["eta"])
    for k in ("a_smooth", "a_t", "eta"):
        assert np.array_equal(gfs[k], wfs[k]), k
    assert np.array_equal(gcomp, wcomp)
    assert np.all(np.isfinite(gcomp)) and gcomp.shape == (N, 1, nsim)
    if ref.available():
        chunk = 256
        e1, e2 = d1.reshape(N, nsim), d2.reshape(N, nsim)           # r_tot = 1: draw s of time t at [t, s]

        def piece(i):
            sl = slice(i * chunk, (i + 1) * chunk)
            return run(ref, chunk, np.ascontiguousarray(e1[:, sl]).ravel(), np.ascontiguousarray(e2[:, sl]).ravel(), roots=False)

        with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
            pieces = list(ex.map(piece, range(nsim // chunk)))
        for i, (r1, r2, rfs, rcomp) in enumerate(pieces):
            sl = slice(i * chunk, (i + 1) * chunk)
            for k in ("y", "a", "eta"):
                assert rel(g1[k][:, :, sl], r1[k]) < RTOL, (i, k)
            for k in ("a_smooth", "eta"):
                assert rel(gfs[k][:, :, sl], rfs[k]) < RTOL, (i, k)
            assert rel(gcomp[:, :, sl], rcomp) < RTOL, i


def test_config4_exact_model_sim_smoother_one_call_4096_draws():
    """ssgpu_SimSmoother on config 4's own model at 4,096 draws: bit-identical to R/SimSmoother.R:34-769 strung together
    from the C oracle's routines (shared symmetric roots), 1e-9 against the same composition of the reference's C++ on a
    prefix of the draws."""
    from oracle import sim_smoother
    import bench  # noqa: F401
    air, sm, N, mc, ks, _ = _config4()
    cs = sysmat.comp_sarima_univariate((12, 1), (0, 1), (1, 1), (1, 1), [0.5 * math.log(0.00135), 0.3, -0.4, -0.55])
    comps = [dict(a=cs["a"], Z=cs["Z"], T=cs["T"], R=cs["R"], Q=cs["Q"], P_star=cs["P_star"], repeat_Q=1, draw_initial=True)]
    nsim = 4096
    draws = np.random.default_rng(44).normal(size=mc * nsim + N * nsim + N * nsim)
    roots = [dict(Q_root=np.stack([cport.sym_root(np.atleast_2d(cs["Q"]))], axis=2), P_root=cport.sym_root(cs["P_star"])), {}]
    args = (nsim, N, 1, comps, sm.Q[:1, :1, :], tuple(sm.args()), int(ks["initialisation_steps"]), ks["a_smooth"][:, 1:],
            ks["a_smooth"][:, :1], ks["eta"][:, 1:], draws)
    got = api.SimSmoother(*args, extract=[cs["Z"]], roots=roots)
    want = sim_smoother.SimSmoother(cport, *args, extract=[cs["Z"]], roots=roots)
    for k in ("epsilon", "a", "eta"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["components"][0], want["components"][0])
    if ref.available():
        k0 = 128                                                        # the first draws through the reference's own code
        sub = np.concatenate([draws[:mc * nsim].reshape(nsim, mc)[:k0].ravel(),
                              draws[mc * nsim:mc * nsim + N * nsim].reshape(N, nsim)[:, :k0].ravel(),
                              draws[mc * nsim + N * nsim:].reshape(N, nsim)[:, :k0].ravel()])
        r0 = sim_smoother.SimSmoother(ref, k0, *args[1:-1], sub, extract=[cs["Z"]])
        got2 = api.SimSmoother(*args, extract=[cs["Z"]])                 # the library's own Jacobi roots
        for k in ("epsilon", "a", "eta"):
            assert rel(got2[k][:, :, :k0], r0[k]) < RTOL, k
        assert rel(got2["components"][0][:, :, :k0], r0["components"][0]) < RTOL


def test_config4_shards_concatenate_to_one_call():
    """BASELINE.json config 4 is 100,000 draws sharded over 8 GPUs: 8 shards of 12,500 draws, each run as its own call
    on its slice of the variates, concatenate to the BITS of one 100,000-draw call (device-resident path, the one
    bench.py --workload sim times): a draw's result does not depend on which shard, CTA or lane computed it."""
    import torch
    air, sm, N, mc, ks, c = _config4()
    steps0 = int(ks["initialisation_steps"])
    S, n_sh = 100_000, 8
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    d1 = torch.randn(N * S, generator=g, device="cuda", dtype=torch.float64)
    d2 = torch.randn(N * S, generator=g, device="cuda", dtype=torch.float64)
    one = np.zeros((1, 1, 1))

    def run(n, e1, e2):
        s1 = api.SimulateC_dev(n, 1, N, np.zeros(mc), c["Z"], c["T"], c["R"], c["Q"], np.zeros((mc, mc)), False, False, e1)
        s2 = api.SimulateC_dev(n, 1, N, np.zeros(1), one, one, one, c["H"], np.zeros((1, 1)), False, True, e2, want=("eta",))
        fs = api.FastSmootherC_dev(s1["y"] + s2["eta"], *sm.args(), steps0)
        comp = api.ExtractComponentC_dev(fs["a_smooth"], sm.Z)
        return s1["a"], fs["a_smooth"], fs["eta"], comp

    whole = run(S, d1, d2)
    assert all(bool(torch.isfinite(x).all()) for x in whole)
    e1, e2 = d1.view(N, S), d2.view(N, S)                               # r_tot = 1: draw s of time t at [t, s]
    sh = S // n_sh
    for i in range(n_sh):
        sl = slice(i * sh, (i + 1) * sh)
        part = run(sh, e1[:, sl].contiguous().view(-1), e2[:, sl].contiguous().view(-1))
        for name, w, q in zip(("a", "a_smooth", "eta", "component"), whole, part):
            assert torch.equal(w[:, :, sl], q), (i, name)
        del part
    # and a prefix against the C oracle (host entry points give the same bits as the device path: tested above)
    k0 = 32
    o1 = cport.SimulateC(k0, 1, N, np.zeros(mc), c["Z"], c["T"], c["R"], c["Q"], np.zeros((mc, mc)), False, False, True,
                         e1[:, :k0].contiguous().view(-1).cpu().numpy())
    assert np.array_equal(o1["a"], whole[0][:, :, :k0].cpu().numpy())


def test_r_glue_row_layout_multivariate():
    """r/R/ssgpu_glue.R loglik_and_gradient() hands LogLikBatchC one row per series, matrix(t(y), nrow = 1): R stores the
    N x p matrix y column-major (t + N j), so t(y) flattened has observation (t, j) at t p + j -- the time-major layout
    the batched kernels read.  Emulated here for p = 2 (a plain matrix(y, nrow = 1) puts (t, j) at t + N j and gives a
    different, wrong, value)."""
    rng = np.random.default_rng(17)
    N, p, m = 30, 2, 5
    sm = random_model(rng, m, p, d=1, N=N)
    y = simulate_y(rng, sm, N, missing=0.1)                       # (N, p)
    r_storage = np.asfortranarray(y).ravel(order="F")            # what R holds for y
    row_t = np.asfortranarray(y.T).ravel(order="F")              # matrix(t(y), nrow = 1)
    assert np.array_equal(row_t, y.reshape(-1), equal_nan=True)  # index t * p + j
    want, tol = loglik_want(y, np.isnan(y), *sm.args())
    got = api.loglik_batch(row_t.reshape(N, p, 1), sm)[0]
    assert abs(got - want) <= tol
    wrong = api.loglik_batch(r_storage.reshape(N, p, 1), sm)[0]   # matrix(y, nrow = 1): observations scrambled
    assert not abs(wrong - want) <= 1e-6 * abs(want)


def test_single_process_device_sharding():
    """ssgpu_use_devices: the host-pointer objective + gradient call shards its series over device slots with one worker
    thread, stream and staging ring per slot and writes each range straight into the caller's arrays.  On a one-GPU box
    the slots share the device (SSGPU_ALLOW_SHARED_DEVICE): same code path, and the results must be the bits of the
    unsharded call -- pinned and pageable host memory, a batch that does not divide evenly."""
    import ctypes as C
    import torch
    rng = np.random.default_rng(91)
    B, N, k = 1000, 60, 4
    thetas, _, _, y = _bsm_batch(rng, B, N, 1, 0.03)
    theta = np.ascontiguousarray(thetas[0])
    sm0 = sysmat.bsm12(theta[0])
    hmodel = api.BatchModel(sm0, N, B, 1)
    qi, pi = np.asarray(BSM_Q_INDEX, dtype=np.int32), np.asarray(BSM_P_INDEX, dtype=np.int32)

    def call(yh, th):
        ll, gr = np.empty(B), np.empty((B, k))
        native.check(native.lib().ssgpu_loglik_grad_batch(C.c_void_p(yh), B, C.byref(hmodel.c), C.c_void_p(th), k,
                                                          qi.ctypes.data, pi.ctypes.data, C.c_double(1e-3), 0,
                                                          ll.ctypes.data, gr.ctypes.data))
        return ll, gr

    y_pin = torch.from_numpy(y).pin_memory()
    ll1, gr1 = call(y.ctypes.data, theta.ctypes.data)
    os.environ["SSGPU_ALLOW_SHARED_DEVICE"] = "1"
    try:
        assert api.use_devices(3) == 3
        ll3, gr3 = call(y.ctypes.data, theta.ctypes.data)                   # pageable: rows through the slots' rings
        ll3p, gr3p = call(y_pin.data_ptr(), theta.ctypes.data)              # pinned: strided DMA per slot
    finally:
        api.use_devices(1)
        del os.environ["SSGPU_ALLOW_SHARED_DEVICE"]
    assert np.array_equal(ll1, ll3) and np.array_equal(gr1, gr3)
    assert np.array_equal(ll1, ll3p) and np.array_equal(gr1, gr3p)
    want = cport.LogLikC(y[:, 700:701], np.isnan(y[:, 700:701]), *sysmat.bsm12(theta[700]).args())
    assert abs(ll3[700] - want) <= RTOL * max(1, abs(want))


def test_time_varying_z_bulk_copy_ring():
    """SSGPU_BLK_ZTMA=1: the rows of a time-varying Z reach the lane kernel through a per-warp ring of shared-memory stages
    filled by cp.async.bulk (TMA) with mbarrier completion instead of per-lane loads.  Opt-in (measured slower for 256-byte
    rows); the environment switch is read once per process, hence the subprocess.  Same bits as the default
    path, for a series length that is not a multiple of the 8-row stage and one shorter than a stage."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from conftest import simulate_y
from statespacer_b200 import api, sysmat
rng = np.random.default_rng(3)
for N in (61, 5):
    base = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05])); m0 = base.m; m = m0 + 2
    X = rng.normal(size=(N, 2))
    Z = np.zeros((1, m, N)); Z[0, :m0, :] = base.Z[0, :, 0][:, None]; Z[0, m0:, :] = X.T
    sm = sysmat.SysMat(Z=Z, T=sysmat.block_diag(base.T[:, :, 0], np.eye(2))[:, :, None],
                       R=np.concatenate([np.eye(m0), np.zeros((2, m0))], axis=0)[:, :, None], Q=base.Q, a=np.zeros(m),
                       P_inf=sysmat.block_diag(base.P_inf, np.eye(2)), P_star=sysmat.block_diag(base.P_star, np.zeros((2, 2))))
    y = np.stack([simulate_y(rng, base, N, missing=0.08)[:, 0] + X @ np.array([1.5, -0.7]) for _ in range(45)], axis=1)
    got = api.loglik_batch(y, sm, kernel="blk_lanes")
    assert "time-varying Z" in api.last_kernel() and api.last_kernel().startswith("loglik_blk_kernel<")
    np.save(sys.argv[1] + "_%%d.npy" %% N, got)
''' % (ROOT_DIR, ROOT_DIR)
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        outs = {}
        for tag, env in (("plain", {}), ("ring", {"SSGPU_BLK_ZTMA": "1"})):
            r = subprocess.run([sys.executable, "-c", code, os.path.join(d, tag)], env={**os.environ, **env},
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs[tag] = [np.load(os.path.join(d, f"{tag}_{N}.npy")) for N in (61, 5)]
        for a, b in zip(outs["plain"], outs["ring"]):
            assert np.array_equal(a, b, equal_nan=True) and np.all(np.isfinite(a[~np.isnan(a)]))
        assert np.all(np.isfinite(outs["plain"][0]))


def test_thread_kernel_shortcuts_against_the_generic_instantiation():
    """The thread kernel's two structural shortcuts against its generic instantiation in subprocesses (the switches are
    read once per process): SSGPU_THR_LS=0 -- the level + slope block by sums is bit for bit the generic product
    (1 * x is exact) --, SSGPU_THR_ZS=0 -- so is leaving out the multiply-adds with z = 0 -- and SSGPU_THR_EPS=0 --
    dropping the residual state changes only the order of a few sums (1e-12); SSGPU_THR_SHARED_DIFFUSE=0 -- every
    evaluation's diffuse phase on the lanes instead of the shared diffuse steps in the thread kernel (1e-11)."""
    import subprocess
    import sys
    import tempfile
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
from conftest import simulate_y
from statespacer_b200 import api, sysmat
rng = np.random.default_rng(8)
sm = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]))
y = np.stack([simulate_y(rng, sm, 80, missing=0.05)[:, 0] for _ in range(96)], axis=1)
got = api.loglik_batch(y, sm, kernel="blk")
open(sys.argv[1] + ".txt", "w").write(api.last_kernel())
np.save(sys.argv[1] + ".npy", got)
''' % (ROOT_DIR, ROOT_DIR)
    with tempfile.TemporaryDirectory() as d:
        outs, names = {}, {}
        for tag, env in (("both", {}), ("no_ls", {"SSGPU_THR_LS": "0"}), ("no_zs", {"SSGPU_THR_ZS": "0"}),
                         ("no_shared", {"SSGPU_THR_SHARED_DIFFUSE": "0"}), ("no_self", {"SSGPU_THR_SELF_SETUP": "0"}),
                         ("no_eps", {"SSGPU_THR_EPS": "0", "SSGPU_THR_LS": "0"})):
            r = subprocess.run([sys.executable, "-c", code, os.path.join(d, tag)], env={**os.environ, **env},
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs[tag] = np.load(os.path.join(d, tag + ".npy"))
            names[tag] = open(os.path.join(d, tag + ".txt")).read()
        assert "residual state dropped, level + slope block by sums" in names["both"] and "zero z skipped" in names["both"], names
        assert "zero z skipped" not in names["no_zs"] and "level + slope" in names["no_zs"], names
        assert np.array_equal(outs["both"], outs["no_zs"])                 # fma(x, 0, acc) == acc
        assert "residual state dropped" in names["no_ls"] and "level + slope" not in names["no_ls"], names
        assert "dropped" not in names["no_eps"] and "level + slope" not in names["no_eps"], names
        assert np.all(np.isfinite(outs["both"]))
        # about half of the 96 series have no missing value among their first 13 observations: their diffuse steps run in
        # the thread kernel with P_inf's share as constants, the others' on the lanes
        assert "13 shared diffuse steps" in names["both"] and "shared diffuse" not in names["no_shared"], names
        assert np.allclose(outs["both"], outs["no_shared"], rtol=1e-11, atol=0)
        # ... and without a hand-over record: the thread kernel sets their state up from the model itself (R = I: same bits
        # as the lane kernel's record)
        assert np.array_equal(outs["both"], outs["no_self"])
        assert np.array_equal(outs["both"], outs["no_ls"])
        assert np.allclose(outs["both"], outs["no_eps"], rtol=1e-12, atol=0)


def test_loglik_from_theta_with_ldl_covariance_blocks():
    """SURVEY.md 8f row 1 for multivariate models: the variance matrices are L D L' built ON THE DEVICE from theta
    (R/Cholesky.R:148-167) -- a bivariate local level with correlated observation errors and correlated level
    disturbances (p = 2, m = 4; Q = blkdiag(H, Q_level), P_star = blkdiag(H, 0): three blocks, k = 6) and a trivariate
    one (3 x 3 blocks, k = 12) -- against the oracle on matrices built by sysmat.cholesky_cov; the central-difference
    gradient against differences of the oracle; the <= 1e-7 -> NA rule (:159-161)."""
    import torch
    dev = torch.device("cuda")
    rng = np.random.default_rng(23)
    for p in (2, 3):
        npar = p + p * (p - 1) // 2
        k, N, B, m = 2 * npar, 40, 9, 2 * p
        Z = np.concatenate([np.eye(p), np.eye(p)], axis=1)
        T = sysmat.block_diag(np.zeros((p, p)), np.eye(p))

        def build(th):
            H, Ql = sysmat.cholesky_cov(th[:npar], p), sysmat.cholesky_cov(th[npar:], p)
            if H is None or Ql is None:
                return None
            return sysmat.SysMat(Z=Z[:, :, None], T=T[:, :, None], R=np.eye(m)[:, :, None], Q=sysmat.block_diag(H, Ql)[:, :, None],
                                 a=np.zeros(m), P_inf=sysmat.block_diag(np.zeros((p, p)), np.eye(p)),
                                 P_star=sysmat.block_diag(H, np.zeros((p, p))))

        theta = np.concatenate([0.5 * np.log(rng.uniform(0.3, 2.0, size=(B, p))), 0.4 * rng.normal(size=(B, npar - p)),
                                0.5 * np.log(rng.uniform(0.05, 0.5, size=(B, p))), 0.4 * rng.normal(size=(B, npar - p))], axis=1)
        theta[4, 0] = 0.5 * math.log(5e-8)                       # a variance below 1e-7: NA
        tpl = build(theta[0])
        y = np.stack([simulate_y(rng, tpl, N, missing=0.1) for _ in range(B)], axis=2)       # (N, p, B)
        model = api.BatchModel(tpl, N, B, 1, device=dev)
        blocks = [("Q", 0, p, 0), ("Q", p, p, npar), ("P_star", 0, p, 0)]
        y_d, th_d = torch.from_numpy(y).to(dev), torch.from_numpy(theta).to(dev)
        ll, gr = api.loglik_cov_batch_dev(y_d, model, th_d, blocks, h=1e-3)
        ll, gr = ll.cpu().numpy(), gr.cpu().numpy()
        assert math.isnan(ll[4]) and np.all(np.isnan(gr[4]))
        for b in range(B):
            if b == 4:
                assert build(theta[b]) is None
                continue
            yb = y[:, :, b]
            f = lambda th: cport.LogLikC(yb, np.isnan(yb), *build(th).args())
            want = f(theta[b])
            assert abs(ll[b] - want) <= RTOL * max(1, abs(want)), (p, b, ll[b], want)
            for i in range(0, k, 3):
                e = np.zeros(k)
                e[i] = 1e-3
                g = (f(theta[b] + e) - f(theta[b] - e)) / 2e-3
                assert abs(gr[b, i] - g) <= 1e-6 * max(1, abs(g)), (p, b, i, gr[b, i], g)
        # V arbitrary vectors per series, no gradient
        pts = np.stack([theta, theta + 0.05, theta - 0.07])
        out = api.loglik_cov_batch_dev(y_d, model, torch.from_numpy(pts).to(dev), blocks, grad=False).cpu().numpy()
        assert out.shape == (3, B) and np.array_equal(out[0], ll, equal_nan=True)
        want = cport.LogLikC(y[:, :, 2], np.isnan(y[:, :, 2]), *build(pts[1, 2]).args())
        assert abs(out[1, 2] - want) <= RTOL * max(1, abs(want))


def test_sim_smoother_sharded_over_device_slots():
    """ssgpu_use_devices + ssgpu_SimSmoother: the draws are cut into contiguous ranges (one worker thread per device slot;
    on a one-GPU box the slots share the device), each range's variates gathered from the reference's consumption order and
    its results written into its block of R's arrays: the bits of the unsharded call, for a model whose components draw
    initial states and repeat Q (structural) and for config 4's SARIMA."""
    from oracle import sim_smoother  # noqa: F401
    rng = np.random.default_rng(12)
    for kind in ("structural", "airline"):
        c = _sim_smoother_case(kind, rng)
        nsim, N = 333, c["N"]
        n_draws = sum((np.asarray(k["a"]).size * nsim if k["draw_initial"] else 0)
                      + N * np.asarray(k["Q"]).shape[0] * k["repeat_Q"] * nsim for k in c["comps"]) + N * c["p"] * nsim
        draws = rng.normal(size=n_draws)
        args = (nsim, N, c["p"], c["comps"], c["H"], c["full"], c["steps"], c["a_hat"], c["eps_hat"], c["eta_hat"], draws)
        one = api.SimSmoother(*args, extract=c["extract"])
        os.environ["SSGPU_ALLOW_SHARED_DEVICE"] = "1"
        try:
            assert api.use_devices(3) == 3
            three = api.SimSmoother(*args, extract=c["extract"])
        finally:
            api.use_devices(1)
            del os.environ["SSGPU_ALLOW_SHARED_DEVICE"]
        for k in ("epsilon", "a", "eta"):
            assert np.array_equal(one[k], three[k]), (kind, k)
        for g, w in zip(three["components"], one["components"]):
            assert np.array_equal(g, w), kind


def test_thread_kernel_correlated_block_and_scratch_ranges():
    """The thread-per-evaluation kernel with (i) correlated disturbances INSIDE a 2 x 2 block (level and slope
    disturbances correlated: R Q R' has off-diagonal entries in the block, the QOFF instantiation) and (ii) a hand-over
    scratch so small that the batch is cut into many ranges of series (SSGPU_THR_SCRATCH_MB, read once per process: hence
    the subprocess) -- same values as one range, and as the oracle."""
    import subprocess
    import sys
    import tempfile
    rng = np.random.default_rng(71)
    B, N = 41, 50
    sm = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]))
    Q = sm.Q[:, :, 0].copy()
    Q[1, 2] = Q[2, 1] = 0.6 * math.sqrt(Q[1, 1] * Q[2, 2])
    smq = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=Q[:, :, None], a=sm.a, P_inf=sm.P_inf, P_star=sm.P_star)
    y = np.stack([simulate_y(rng, smq, N, missing=0.1)[:, 0] for _ in range(B)], axis=1)
    ab = rng.normal(size=(B, 14)) * 0.1
    got = api.loglik_batch(y, smq, a=ab, kernel="blk")
    assert api.last_kernel().startswith("loglik_thr_kernel<NB=7,"), api.last_kernel()
    lanes = api.loglik_batch(y, smq, a=ab, kernel="blk_lanes")
    for b in range(B):
        smb = sysmat.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=Q[:, :, None], a=ab[b], P_inf=sm.P_inf, P_star=sm.P_star)
        want = cport.LogLikC(y[:, b:b + 1], np.isnan(y[:, b:b + 1]), *smb.args())
        assert abs(got[b] - want) <= RTOL * max(1, abs(want)), (b, got[b], want)
        assert abs(lanes[b] - want) <= RTOL * max(1, abs(want)), (b, lanes[b], want)
    # ranges of series: 1 MB of scratch holds ~100 evaluations of 9 parameter vectors
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from statespacer_b200 import api, sysmat
rng = np.random.default_rng(5)
B, N, V = 700, 30, 9
thetas = 0.5 * np.log([1.0, 0.1, 0.01, 0.05]) + 0.25 * rng.normal(size=(V, B, 4))
var = np.exp(2 * thetas)
Qd = np.concatenate([var[..., :3], np.repeat(var[..., 3:4], 11, axis=-1)], axis=-1)
Pd = np.zeros((V, B, 14)); Pd[..., 0] = var[..., 0]
y = np.cumsum(rng.normal(size=(N, B)), axis=0)
y[rng.uniform(size=y.shape) < 0.05] = np.nan
got = api.loglik_batch(y, sysmat.bsm12(thetas[0, 0]), V=V, Q_diag=Qd, P_star_diag=Pd, kernel="blk")
assert api.last_kernel().startswith("loglik_thr_kernel<NB=7,")
from oracle import cport
for v, b in ((0, 0), (3, 350), (8, 698), (0, 699)):
    yb = y[:, b:b + 1]
    want = cport.LogLikC(yb, np.isnan(yb), *sysmat.bsm12(thetas[v, b]).args())
    assert (np.isnan(want) and np.isnan(got[v, b])) or abs(got[v, b] - want) <= 1e-9 * max(1, abs(want)), (v, b, got[v, b], want)
np.save(sys.argv[1], got)
''' % ROOT_DIR
    with tempfile.TemporaryDirectory() as d:
        outs = []
        for tag, env in (("one", {}), ("many", {"SSGPU_THR_SCRATCH_MB": "1"})):
            r = subprocess.run([sys.executable, "-c", code, os.path.join(d, tag + ".npy")], env={**os.environ, **env},
                               capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            outs.append(np.load(os.path.join(d, tag + ".npy")))
        # a range boundary changes which evaluations share a warp of the lane kernel, i.e. the time point at which an
        # evaluation is handed over (the warp's last group to leave the diffuse phase decides): same arithmetic, a few steps
        # computed by the other kernel -- equal to rounding, not to the bit
        assert outs[0].shape == (9, 700) and np.allclose(outs[0], outs[1], rtol=1e-12, atol=0, equal_nan=True)
        assert np.isnan(outs[0][:, 699]).all() and np.mean(np.isfinite(outs[0])) > 0.9    # series 699: NA in the reference too


@pytest.mark.parametrize("B,N,V", [(64, 100, 1), (1000, 61, 1), (96, 5, 1), (130, 37, 3), (37, 40, 1)])
def test_local_level_observations_through_bulk_copy_ring(B, N, V):
    """One lane per series (local level, config 1's model batched): the observations of a warp's series reach it through a
    ring of shared-memory stages filled by cp.async.bulk with mbarrier completion (TMA), 32 time points ahead -- on when
    the number of series is even (16-byte rows), plain loads otherwise (B = 37).  Series lengths that are not a multiple
    of the 8-row stage or shorter than one stage, several parameter vectors per series (narrower rows), missing values;
    against the oracle and the generic kernel."""
    rng = np.random.default_rng(B + N)
    sm = sysmat.nile_local_level(1.3, 0.4)
    y = np.cumsum(rng.normal(size=(N, B)), axis=0) + rng.normal(size=(N, B))
    y[rng.uniform(size=y.shape) < 0.1] = np.nan
    y[:, 3] = np.nan                                      # a series that is never observed: NA
    sc = rng.uniform(0.5, 2.0, size=(V, B, 1))
    Qd = np.array([1.3, 0.4])[None, None, :] * sc
    Pd = np.array([1.3, 0.0])[None, None, :] * sc
    got = api.loglik_batch(y, sm, V=V, Q_diag=Qd if V > 1 else Qd[0], P_star_diag=Pd if V > 1 else Pd[0], kernel="blk")
    assert api.last_kernel().startswith("loglik_blk_kernel<L=1,NB=1,"), api.last_kernel()
    gen = api.loglik_batch(y, sm, V=V, Q_diag=Qd if V > 1 else Qd[0], P_star_diag=Pd if V > 1 else Pd[0], kernel="generic")
    got, gen = got.reshape(V, B), gen.reshape(V, B)
    assert np.all(np.isnan(got[:, 3])) and np.all(np.isfinite(np.delete(got, 3, axis=1)))
    assert np.allclose(got, gen, rtol=1e-12, atol=1e-12, equal_nan=True)
    for v in range(V):
        for b in range(0, B, max(1, B // 9)):
            smb = sysmat.nile_local_level(1.3 * sc[v, b, 0], 0.4 * sc[v, b, 0])
            want = cport.LogLikC(y[:, b:b + 1], np.isnan(y[:, b:b + 1]), *smb.args())
            assert (math.isnan(want) and math.isnan(got[v, b])) or abs(got[v, b] - want) <= RTOL * max(1, abs(want)), (v, b)
