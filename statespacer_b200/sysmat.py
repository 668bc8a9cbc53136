"""Host-side system-matrix builders in the layout of statespacer's GetSysMat / CombineSysMat.

These are tiny NumPy helpers (no time loop, no data) that produce exactly the dense matrices the
reference's R code hands to its native routines, so that tests and bench.py can restate the
BASELINE.json configs without R.  Layout (R/GetSysMat.R:1378-1426, R/statespacer.R:776-807):

    state = [ eps (p residual states) ; component states ... ]
    Z = [I_p | Z_comp]        T = blkdiag(0_p, T_comp)      R = blkdiag(I_p, R_comp)
    Q = blkdiag(H, Q_comp)    a = [0 ; a_comp]              P_inf = blkdiag(0, P_inf_comp)
    P_star = blkdiag(H, P_star_comp)

All 3-D matrices are ``(rows, cols, slices)`` like the R arrays (slices = 1 when time-invariant,
R/statespacer.R:903-914).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class SysMat:
    """The `*_kal` matrices of R/GetSysMat.R:1456-1485."""
    Z: np.ndarray       # (p, m, 1|N)
    T: np.ndarray       # (m, m, 1|N)
    R: np.ndarray       # (m, r, 1|N)
    Q: np.ndarray       # (r, r, 1|N)
    a: np.ndarray       # (m,)
    P_inf: np.ndarray   # (m, m)
    P_star: np.ndarray  # (m, m)

    @property
    def p(self):
        return self.Z.shape[0]

    @property
    def m(self):
        return self.Z.shape[1]

    @property
    def r(self):
        return self.R.shape[1]

    def args(self):
        """Positional tail shared by LogLikC/KalmanC: (a, P_inf, P_star, Z, T, R, Q)."""
        return (self.a, self.P_inf, self.P_star, self.Z, self.T, self.R, self.Q)


def block_diag(*mats):
    """R/BlockMatrix.R for 2-D inputs."""
    mats = [np.atleast_2d(np.asarray(x, dtype=np.float64)) for x in mats]
    rows = sum(x.shape[0] for x in mats)
    cols = sum(x.shape[1] for x in mats)
    out = np.zeros((rows, cols))
    i = j = 0
    for x in mats:
        out[i:i + x.shape[0], j:j + x.shape[1]] = x
        i += x.shape[0]
        j += x.shape[1]
    return out


def cholesky_cov(param, n):
    """LDL' parameterisation of an n x n covariance with a full lower-triangular format
    (R/Cholesky.R:148-167): diag(D) = exp(2*param[:n]); strict lower triangle of L (column-major
    order) = param[n:].  Returns None where the reference returns NA (any diagonal <= 1e-7)."""
    param = np.asarray(param, dtype=np.float64)
    n_low = n * (n - 1) // 2
    assert param.shape[0] == n + n_low
    d = np.exp(2 * param[:n])
    if np.any(d <= 1e-7):
        return None
    L = np.eye(n)
    idx = 0
    for j in range(n):                  # R fills logical-indexed positions in column-major order
        for i in range(j + 1, n):
            L[i, j] = param[n + idx]
            idx += 1
    return (L * d[None, :]) @ L.T       # tcrossprod(L %*% D, L)


def pac_univariate(A):
    """R/AnsleyKohn.R:16."""
    A = np.asarray(A, dtype=np.float64)
    return A / np.sqrt(1 + A ** 2)


def transform_pac_univariate(P):
    """R/AnsleyKohn.R:52-63 (Durbin-Levinson on partial autocorrelations)."""
    new = np.array(P, dtype=np.float64).copy()
    n = new.shape[0]
    for i in range(1, n):
        old = new.copy()
        for j in range(1, i + 1):
            new[j - 1] = old[j - 1] - new[i] * old[i - j]
    return new


def coeff_arma_univariate(A, ar, ma):
    """R/AnsleyKohn.R:172-184: stationary AR / invertible MA coefficients from free parameters."""
    P = pac_univariate(np.asarray(A, dtype=np.float64).reshape(-1))
    out = {}
    if ar > 0:
        out["ar"] = transform_pac_univariate(P[:ar])
    if ma > 0:
        out["ma"] = -transform_pac_univariate(P[ar:ar + ma])
    return out


def coeff_var1(A, variance):
    """Multivariate AR(1) coefficient matrix, R/AnsleyKohn.R:21,69,187-206 with ar=1, ma=0."""
    A = np.asarray(A, dtype=np.float64)
    k = A.shape[0]
    U = np.linalg.cholesky(np.eye(k) + A @ A.T).T           # chol() returns the upper factor
    P = np.linalg.solve(U, np.eye(k)).T @ A                  # crossprod(solve(U), A)
    sigma = np.eye(k) - P @ P.T
    L = np.linalg.cholesky(variance)
    L_ar = np.linalg.cholesky(sigma)
    return L @ np.linalg.inv(L_ar) @ P @ L_ar @ np.linalg.inv(L)


def _combine(p, H, comps):
    """comps: list of dicts with Z (p, k), T, R, Q, a, P_inf, P_star."""
    H = np.atleast_2d(np.asarray(H, dtype=np.float64))
    Z = np.concatenate([np.eye(p)] + [c["Z"] for c in comps], axis=1)
    T = block_diag(np.zeros((p, p)), *[c["T"] for c in comps])
    R = block_diag(np.eye(p), *[c["R"] for c in comps])
    Q = block_diag(H, *[c["Q"] for c in comps])
    a = np.concatenate([np.zeros(p)] + [np.asarray(c["a"], dtype=np.float64).reshape(-1) for c in comps])
    P_inf = block_diag(np.zeros((p, p)), *[c["P_inf"] for c in comps])
    P_star = block_diag(H, *[c["P_star"] for c in comps])
    return SysMat(Z[:, :, None], T[:, :, None], R[:, :, None], Q[:, :, None], a, P_inf, P_star)


# ---------------------------------------------------------------------------------------------
# components (univariate p = 1 forms; R/Components-Unobserved.R:33-60,138-184,476-536)
# ---------------------------------------------------------------------------------------------
def comp_local_level(var_level):
    return dict(Z=np.ones((1, 1)), T=np.ones((1, 1)), R=np.ones((1, 1)), Q=np.array([[var_level]]),
                a=np.zeros(1), P_inf=np.ones((1, 1)), P_star=np.zeros((1, 1)))


def comp_level_slope(var_level, var_slope):
    return dict(Z=np.array([[1.0, 0.0]]), T=np.array([[1.0, 1.0], [0.0, 1.0]]), R=np.eye(2),
                Q=np.diag([var_level, var_slope]), a=np.zeros(2), P_inf=np.eye(2), P_star=np.zeros((2, 2)))


def comp_bsm_trig(s, var_seas):
    even = s % 2 == 0
    ln = s // 2 - 1 if even else s // 2
    lam = 2 * math.pi * np.arange(1, ln + 1) / s
    C, S = np.diag(np.cos(lam)), np.diag(np.sin(lam))
    T = np.block([[C, S], [-S, C]])
    Z = np.concatenate([np.ones((1, ln)), np.zeros((1, ln))], axis=1)
    if even:
        T = block_diag(T, -np.ones((1, 1)))
        Z = np.concatenate([Z, np.ones((1, 1))], axis=1)
    k = s - 1
    return dict(Z=Z, T=T, R=np.eye(k), Q=var_seas * np.eye(k), a=np.zeros(k), P_inf=np.eye(k),
                P_star=np.zeros((k, k)))


def comp_sarima_univariate(s, ar, i, ma, param):
    """Univariate SARIMA component (R/Components-BoxJenkins.R:364-482 fixed part, :486-756 update
    part, n_sarima = 1).  param = [0.5*log(sigma2), AR/MA free params per seasonality ...]."""
    s, ar, i, ma = (np.asarray(x, dtype=int) for x in (s, ar, i, ma))
    param = np.asarray(param, dtype=np.float64)
    r = int(max(np.sum(s * ar), np.sum(s * ma) + 1))
    s_aux = np.repeat(s, i)
    n_ns = int(np.sum(s_aux))
    sigma2 = math.exp(2 * param[0])
    rest = param[1:]
    AR = np.zeros(int(np.sum(s * ar)))
    MA = np.zeros(int(np.sum(s * ma)))
    ar_poly: list[int] = []
    ma_poly: list[int] = []
    for k in range(len(s)):
        na, nm = int(ar[k]), int(ma[k])
        if na + nm == 0:
            continue
        coeff = coeff_arma_univariate(rest[:na + nm], na, nm)
        rest = rest[na + nm:]
        if na > 0:
            old = AR.copy()
            add = [int(s[k]) * q for q in range(1, na + 1)]
            for q, lag in enumerate(add):
                AR[lag - 1] -= coeff["ar"][q]
            for q in range(1, na + 1):
                for lag0 in ar_poly:
                    lag = int(s[k]) * q + lag0
                    AR[lag - 1] -= coeff["ar"][q - 1] * old[lag0 - 1]
                    add.append(lag)
            ar_poly = list(dict.fromkeys(ar_poly + add))
        if nm > 0:
            old = MA.copy()
            add = [int(s[k]) * q for q in range(1, nm + 1)]
            for q, lag in enumerate(add):
                MA[lag - 1] += coeff["ma"][q]
            for q in range(1, nm + 1):
                for lag0 in ma_poly:
                    lag = int(s[k]) * q + lag0
                    MA[lag - 1] += coeff["ma"][q - 1] * old[lag0 - 1]
                    add.append(lag)
            ma_poly = list(dict.fromkeys(ma_poly + add))
    # stationary block
    T1 = np.zeros((r, r))
    for q in range(r - 1):
        T1[q, q + 1] = 1.0
    if AR.shape[0] > 0:
        T1[:AR.shape[0], 0] = -AR
    R1 = np.zeros((r, 1))
    R1[0, 0] = 1.0
    R1[1:1 + MA.shape[0], 0] = MA
    if np.sum(ar) + np.sum(ma) == 0:
        P_st = np.array([[sigma2]])
        assert r == 1
        P_st = R1 @ np.array([[sigma2]]) @ R1.T
    else:
        K = np.kron(T1, T1)
        vec = (R1 * sigma2) @ R1.T
        P_st = np.linalg.solve(np.eye(r * r) - K, vec.reshape(-1, order="F")).reshape((r, r), order="F")
    # non-stationary block
    Z = np.zeros((1, n_ns + r))
    T = np.zeros((n_ns + r, n_ns + r))
    T[n_ns:, n_ns:] = T1
    off = 0
    # T_list column blocks: for each s_aux[q]: [0_{x-1}, 1]; then [1]; then zeros(r-1)
    tl = []
    for x in s_aux:
        v = np.zeros(int(x))
        v[-1] = 1.0
        tl.append(v)
    row = 0
    for q, x in enumerate(s_aux):
        x = int(x)
        Z[0, off + x - 1] = 1.0
        full = np.concatenate(tl + [np.ones(1), np.zeros(r - 1)])
        T[row, :] = full
        for w in range(x - 1):
            T[row + 1 + w, off + w] = 1.0
        tl[q] = np.zeros(x)
        row += x
        off += x
    Z[0, n_ns] = 1.0
    R = np.concatenate([np.zeros((n_ns, 1)), R1], axis=0)
    P_inf = block_diag(np.eye(n_ns), np.zeros((r, r))) if n_ns > 0 else np.zeros((r, r))
    P_star = block_diag(np.zeros((n_ns, n_ns)), P_st) if n_ns > 0 else P_st
    return dict(Z=Z, T=T, R=R, Q=np.array([[sigma2]]), a=np.zeros(n_ns + r), P_inf=P_inf, P_star=P_star,
                ar=-AR, ma=MA)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs
# ---------------------------------------------------------------------------------------------
def nile_local_level(H, Q_level):
    """Config 1 (R/statespacer.R:199-202): m = r = 2."""
    return _combine(1, [[H]], [comp_local_level(Q_level)])


def airline(param, s=(12, 1), ar=(0, 0), i=(1, 1), ma=(1, 1), H=0.0):
    """KAT-4 / config 4 style SARIMA with H_format = matrix(0) (vignettes/boxjenkins.Rmd:97-104)."""
    return _combine(1, [[H]], [comp_sarima_univariate(s, ar, i, ma, param)])


def bsm12(theta, s=12):
    """Config 5: local level + slope + trigonometric seasonal, m = r = 14 for s=12.
    theta = 0.5*log of (var_eps, var_level, var_slope, var_seasonal) (R/Cholesky.R:158)."""
    v = np.exp(2 * np.asarray(theta, dtype=np.float64))
    return _combine(1, [[v[0]]], [comp_level_slope(v[1], v[2]), comp_bsm_trig(s, v[3])])


DNS_MATURITY = np.array([3, 6, 12, 24, 36, 60, 84, 120], dtype=np.float64)


def dns_yield_curve(param):
    """Config 3: dynamic Nelson-Siegel self-specified model (vignettes/selfspec.Rmd:94-138),
    collapse=FALSE form: p = 8, m = r = 14."""
    param = np.asarray(param, dtype=np.float64)
    lam = math.exp(2 * param[0])
    sigma2 = math.exp(2 * param[1])
    H = sigma2 * np.eye(8)
    lm = lam * DNS_MATURITY
    z = np.exp(-lm)
    Zf = np.ones((8, 3))
    Zf[:, 1] = (1 - z) / lm
    Zf[:, 2] = Zf[:, 1] - z
    Qf = cholesky_cov(param[2:8], 3)
    Phi = coeff_var1(param[8:17].reshape((3, 3), order="F"), Qf)
    K = np.kron(Phi, Phi)
    Pst = np.linalg.solve(np.eye(9) - K, Qf.reshape(-1, order="F")).reshape((3, 3), order="F")
    comp = dict(
        Z=np.concatenate([Zf, np.zeros((8, 3))], axis=1),
        T=np.block([[Phi, np.eye(3) - Phi], [np.zeros((3, 3)), np.eye(3)]]),
        R=np.eye(6), Q=block_diag(Qf, np.zeros((3, 3))),
        a=np.concatenate([param[17:20], param[17:20]]),
        P_inf=np.zeros((6, 6)), P_star=block_diag(Pst, np.zeros((3, 3))))
    return _combine(8, H, [comp]), Phi


def dns_collapsed(param, H_star):
    """Config 3 with collapse = TRUE (R/statespacer.R:380-405,815-832): the 8 yields are collapsed onto the m2 = 3
    factors, state = [3 residual states of the collapsed observations ; 6 component states], p = 3, m = r = 9.
    `H_star` (3 x 3) is the variance of the collapsed observation error, A* H A*' from the collapse transform."""
    sm, _ = dns_yield_curve(param)
    c = dict(Z=None, T=sm.T[8:, 8:, 0], R=sm.R[8:, 8:, 0], Q=sm.Q[8:, 8:, 0], a=sm.a[8:], P_inf=sm.P_inf[8:, 8:],
             P_star=sm.P_star[8:, 8:])
    Hs = np.asarray(H_star, dtype=np.float64).reshape(3, 3)
    Z = np.concatenate([np.eye(3), np.eye(3), np.zeros((3, 3))], axis=1)               # Z_kal_tf, :399-402
    T = block_diag(np.zeros((3, 3)), c["T"])
    R = block_diag(np.eye(3), c["R"])
    Q = block_diag(Hs, c["Q"])
    a = np.concatenate([np.zeros(3), c["a"]])
    P_inf = block_diag(np.zeros((3, 3)), c["P_inf"])
    P_star = block_diag(Hs, c["P_star"])
    return SysMat(Z[:, :, None], T[:, :, None], R[:, :, None], Q[:, :, None], a, P_inf, P_star)


def bsm_multivariate_regressors(X_list, H, Q_level, Q_slope, Q_seas, s=12):
    """Config 2 shape (Seatbelts-style): p dependent variables with local level + slope
    (R/Components-Unobserved.R:138-184), trigonometric seasonal (:476-536) and explanatory variables with
    time-varying Z (R/Components-Regressors.R:40-90).  X_list[i] is the N x k_i regressor matrix of series i.
    State = [eps (p) | level (p) | slope (p) | seasonal ((s-1) p, frequency-major) | coefficients (sum k_i)];
    H, Q_level, Q_slope, Q_seas are p x p covariances (the seasonal one repeated for its s-1 blocks,
    R/Components-Unobserved.R:560-575)."""
    p = len(X_list)
    N = X_list[0].shape[0]
    I = np.eye(p)
    even = s % 2 == 0
    ln = s // 2 - 1 if even else s // 2
    lam = 2 * math.pi * np.arange(1, ln + 1) / s
    T1 = block_diag(*[math.cos(x) * I for x in lam])
    T2 = block_diag(*[math.sin(x) * I for x in lam])
    Ts = np.block([[T1, T2], [-T2, T1]])
    Zs = np.concatenate([np.tile(I, (1, ln)), np.zeros((p, ln * p))], axis=1)
    if even:
        Ts = block_diag(Ts, -I)
        Zs = np.concatenate([Zs, I], axis=1)
    ks = (s - 1) * p
    kx = sum(x.shape[1] for x in X_list)
    Tl = np.block([[I, I], [np.zeros((p, p)), I]])
    Zl = np.concatenate([I, np.zeros((p, p))], axis=1)
    T = block_diag(np.zeros((p, p)), Tl, Ts, np.eye(kx))
    m = T.shape[0]
    Z = np.zeros((p, m, N))
    Z[:, :p, :] = I[:, :, None]
    Z[:, p:3 * p, :] = Zl[:, :, None]
    Z[:, 3 * p:3 * p + ks, :] = Zs[:, :, None]
    off = 3 * p + ks
    for i, X in enumerate(X_list):
        Z[i, off:off + X.shape[1], :] = X.T
        off += X.shape[1]
    Q = block_diag(np.asarray(H, dtype=np.float64), Q_level, Q_slope, *([np.asarray(Q_seas)] * (s - 1)),
                   np.zeros((kx, kx)))
    P_inf = block_diag(np.zeros((p, p)), np.eye(m - p))
    P_star = block_diag(np.asarray(H, dtype=np.float64), np.zeros((m - p, m - p)))
    return SysMat(Z, T[:, :, None], np.eye(m)[:, :, None], Q[:, :, None], np.zeros(m), P_inf, P_star)
