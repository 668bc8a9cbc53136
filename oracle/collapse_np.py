"""TEST INFRASTRUCTURE -- NumPy restatement of the collapse transform of R/statespacer.R:815-900 (the observation
vector y_t (p) is replaced by y*_t = A*_t y_t (m2 < p) with A* = (Z'H^-1 Z)^-1 Z'H^-1, plus the loglikelihood
contribution of the part that is projected away).  The product never imports this file.

Pinned by KAT-3 (docs/articles/selfspec.html:306-307): the dynamic Nelson-Siegel model gives the same -loglik/N =
7.005830 with collapse = TRUE (p = 3, m = 9) as without (p = 8, m = 14); tests/test_oracle_kat.py."""
import numpy as np


def collapse(y, H, Z):
    """y: N x p with NaN = missing; H: p x p x {1|N}; Z: p x m2 x {1|N} (the columns of Z_kal that are collapsed onto,
    R/statespacer.R:765-773).  Returns y_kal (N x m2, NaN where R writes NA), H_star (m2 x m2 x {1|N}), loglik_add."""
    y = np.asarray(y, dtype=np.float64)
    N, p = y.shape
    H = np.asarray(H, dtype=np.float64).reshape(p, p, -1)
    Z = np.asarray(Z, dtype=np.float64)
    Z = Z.reshape(p, Z.shape[1], -1)
    m2 = Z.shape[1]
    nH, nZ = H.shape[2], Z.shape[2]
    ns = max(nH, nZ)
    y_temp = np.where(np.isnan(y), 0.0, y)                              # :407-409
    y_kal = np.zeros((N, m2))
    H_star = np.zeros((m2, m2, ns))
    loglik_add = 0.0
    for i in range(N):                                                   # :815-897, all four branches
        H_t = H[:, :, i if nH > 1 else 0]
        Z_t = Z[:, :, i if nZ > 1 else 0]
        Hinv = np.linalg.inv(H_t)
        ZtHinv = Z_t.T @ Hinv
        A_star = np.linalg.solve(ZtHinv @ Z_t, ZtHinv)
        y_kal[i] = A_star @ y_temp[i]
        Hs = A_star @ H_t @ A_star.T
        if ns > 1 or i == 0:
            H_star[:, :, i if ns > 1 else 0] = Hs
        e = y_temp[i] - Z_t @ y_kal[i]
        loglik_add += 0.5 * np.log(np.linalg.det(Hs) / np.linalg.det(H_t)) - 0.5 * e @ Hinv @ e
    y_out = np.where(y_kal == 0.0, np.nan, y_kal)                        # :830, :851 ...
    loglik_add -= 0.5 * (np.sum(~np.isnan(y)) - np.sum(~np.isnan(y_out))) * np.log(2 * np.pi)
    return y_out, H_star, loglik_add
