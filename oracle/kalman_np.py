"""CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT) -- NumPy restatement of statespacer's five
native routines.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (statespacer_b200/) never does.

Every function follows the reference line by line (same branch structure, same thresholds, same
`P * L_0.t()` GEMM form), so it is slow on purpose.  Arrays use the R / Armadillo shapes:
3-D system matrices are ``(rows, cols, slices)`` with ``slices in {1, N}``.

Pinned against the reference's own golden vectors (tests/test_oracle_kat.py): KAT-1a/1b
(docs/articles/intro.html:168,170), KAT-3 (docs/articles/selfspec.html:306-307), KAT-4a/4b
(docs/articles/boxjenkins.html:241-242,269-271) and against the unmodified reference sources
compiled into oracle/_ref (tests/test_oracle_vs_ref.py).

Reference files restated here:
  src/LogLikC.cpp:20-215, src/KalmanC.cpp:21-594 (diagnostics=FALSE and TRUE),
  src/FastSmootherC.cpp:23-375, src/SimulateC.cpp:26-139, src/ExtractComponentC.cpp:12-40.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-7          # src/LogLikC.cpp:49,109,112,158,173,205
KAPPA = 1.0e7       # src/KalmanC.cpp:63
CONSTANT = -math.log(2 * math.pi) / 2   # src/LogLikC.cpp:62


def _cube(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    assert x.ndim == 3
    return x


def _all_small(P):
    # arma::all(arma::vectorise(arma::abs(P)) < 1e-7); NaN compares false -> "not all small"
    return bool(np.all(np.abs(P) < EPS))


# --------------------------------------------------------------------------------------------
# src/LogLikC.cpp:20-215
# --------------------------------------------------------------------------------------------
def LogLikC(y, y_isna, a, P_inf, P_star, Z, T, R, Q):
    y = np.asarray(y, dtype=np.float64)
    y_isna = np.asarray(y_isna).astype(bool)
    a = np.array(a, dtype=np.float64).reshape(-1).copy()          # by value (:22)
    P_inf = np.array(P_inf, dtype=np.float64).copy()              # by value (:23)
    P_star = np.array(P_star, dtype=np.float64).copy()            # by value (:24)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p = y.shape                                                # :31
    m = a.shape[0]
    Z_tv, T_tv, R_tv, Q_tv = (Z.shape[2] > 1, T.shape[2] > 1, R.shape[2] > 1, Q.shape[2] > 1)  # :37
    Z_mat, T_mat, R_mat, Q_mat = Z[:, :, 0], T[:, :, 0], R[:, :, 0], Q[:, :, 0]  # :41
    initialisation = not _all_small(P_inf)                        # :49
    non_init = 0
    F_eq0 = 0                                                     # :53
    loglik = 0.0                                                  # :56
    eye = np.eye(m)
    for i in range(N):                                            # :68
        if Z_tv and i > 0:
            Z_mat = Z[:, :, i]
        if T_tv and i > 0:
            T_mat = T[:, :, i]
        if R_tv and i > 0:
            R_mat = R[:, :, i]
        if Q_tv and i > 0:
            Q_mat = Q[:, :, i]
        for j in range(p):                                        # :85
            if y_isna[i, j]:                                      # :88
                continue
            Z_row = Z_mat[j, :]                                   # :93-95 (semantically always row j)
            if initialisation:                                    # :98
                M_inf = P_inf @ Z_row                             # :101
                M_star = P_star @ Z_row                           # :102
                F_inf = float(Z_row @ M_inf)                      # :105
                F_star = float(Z_row @ M_star)                    # :106
                if F_inf < EPS:                                   # :109
                    if F_star < EPS:                              # :112
                        continue
                    F_1 = 1 / F_star                              # :117
                    v = y[i, j] - float(Z_row @ a)                # :120
                    K_0 = M_star * F_1                            # :123
                    L_0 = eye - np.outer(K_0, Z_row)              # :124
                    a = a + K_0 * v                               # :128
                    P_star = P_star @ L_0.T                       # :129
                    loglik += CONSTANT - (math.log(F_star) + v ** 2 * F_1) / 2   # :130
                    continue
                F_1 = 1 / F_inf                                   # :136
                F_2 = -F_1 ** 2 * F_star                          # :137
                v = y[i, j] - float(Z_row @ a)                    # :140
                K_0 = M_inf * F_1                                 # :143
                L_0 = eye - np.outer(K_0, Z_row)                  # :144
                K_1 = M_star * F_1 + M_inf * F_2                  # :145
                L_1 = -np.outer(K_1, Z_row)                       # :146
                a = a + K_0 * v                                   # :150
                P_star = P_inf @ L_1.T + P_star @ L_0.T           # :151
                P_inf = P_inf @ L_0.T                             # :152
                loglik += CONSTANT - math.log(F_inf) / 2          # :153
                if j < p - 1:                                     # :157
                    initialisation = not _all_small(P_inf)        # :158
            else:
                non_init += 1                                     # :164
                M_star = P_star @ Z_row                           # :167
                F_star = float(Z_row @ M_star)                    # :170
                if F_star < EPS:                                  # :173
                    F_eq0 += 1
                else:
                    F_1 = 1 / F_star
                    v = y[i, j] - float(Z_row @ a)                # :184
                    K_0 = M_star * F_1
                    L_0 = eye - np.outer(K_0, Z_row)
                    a = a + K_0 * v                               # :192
                    P_star = P_star @ L_0.T                       # :193
                    loglik += CONSTANT - (math.log(F_star) + v ** 2 * F_1) / 2   # :194
        if i < N - 1:                                             # :200
            a = T_mat @ a
            P_star = T_mat @ P_star @ T_mat.T + R_mat @ Q_mat @ R_mat.T   # :202
            if initialisation:
                P_inf = T_mat @ P_inf @ T_mat.T                   # :204
                initialisation = not _all_small(P_inf)            # :205
    if non_init == F_eq0:                                         # :211
        return float("nan")
    return loglik / N                                             # :214


# --------------------------------------------------------------------------------------------
# helpers for the diagnostics branch of KalmanC (arma::inv / arma::svd on p x p matrices)
# --------------------------------------------------------------------------------------------
def _sym_root_svd(A):
    """U diag(sqrt(s)) U' with (U, s, V) = svd(A)  -- src/KalmanC.cpp:178-181, src/SimulateC.cpp:73-74."""
    U, s, _ = np.linalg.svd(A)
    return U @ np.diag(np.sqrt(s)) @ U.T


# --------------------------------------------------------------------------------------------
# src/KalmanC.cpp:21-594
# --------------------------------------------------------------------------------------------
def KalmanC(y, y_isna, a, P_inf, P_star, Z, T, R, Q, diagnostics=False):
    y = np.asarray(y, dtype=np.float64)
    y_isna = np.asarray(y_isna).astype(bool)
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    P_inf = np.asarray(P_inf, dtype=np.float64)
    P_star = np.asarray(P_star, dtype=np.float64)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p = y.shape
    m = a.shape[0]
    r = R.shape[1]                                                # :34
    Np = N * p
    Np_min1, N_min1, p_min1 = Np - 1, N - 1, p - 1                # :37
    Z_tv, T_tv, R_tv, Q_tv = (Z.shape[2] > 1, T.shape[2] > 1, R.shape[2] > 1, Q.shape[2] > 1)
    Z_mat, T_mat, T_mat_now = Z[:, :, 0], T[:, :, 0], T[:, :, 0]
    R_mat, Q_mat = R[:, :, 0], Q[:, :, 0]
    initialisation = not _all_small(P_inf)                        # :52
    initialisation_steps = 0                                      # :55
    nan = float("nan")
    # Uninitialised Armadillo/Rcpp storage is represented by NaN-filled arrays for matrices the
    # reference leaves unset on skipped branches and zero for Rcpp::NumericVector(Np) (zero-filled).
    M_inf = np.zeros((m, Np)); M_star = np.zeros((m, Np))
    K_0 = np.zeros((m, Np)); K_1 = np.zeros((m, Np))              # :58
    L_0 = np.zeros((m, m, Np)); L_1 = np.zeros((m, m, Np))        # :59
    F_inf = np.zeros(Np); F_star = np.zeros(Np); F_1 = np.zeros(Np); F_2 = np.zeros(Np)
    v_UT = np.zeros(Np)                                           # :60-61
    loglik = np.zeros(Np)                                         # :66
    a_mat = np.zeros((Np, m)); a_pred = np.zeros((N, m)); a_fil = np.zeros((N, m))
    a_smooth = np.zeros((N, m))                                   # :69
    P_inf_cube = np.zeros((m, m, Np)); P_star_cube = np.zeros((m, m, Np))
    P_pred = np.zeros((m, m, N)); P_fil = np.zeros((m, m, N)); V = np.zeros((m, m, N))
    P_inf_pred = np.zeros((m, m, N)); P_star_pred = np.zeros((m, m, N))
    P_inf_fil = np.zeros((m, m, N)); P_star_fil = np.zeros((m, m, N))   # :70-73
    a_mat[0, :] = a
    P_inf_cube[:, :, 0] = P_inf
    P_star_cube[:, :, 0] = P_star                                 # :74-76
    yfit = np.zeros((N, p)); v = np.zeros((N, p)); v_norm = np.full((N, p), nan)
    eta = np.zeros((N, r)); e = np.zeros((N, p))                  # :79-81
    Fmat = np.zeros((p, p, N)); eta_var = np.zeros((r, r, N)); D = np.zeros((p, p, N))
    calculate_vnorm = False                                       # :84
    Finv = np.zeros((p, p, N)); v_0na = np.zeros((N, p))
    a_fc = np.zeros(m)
    P_inf_fc = np.zeros((m, m)); P_star_fc = np.zeros((m, m)); P_fc = np.zeros((m, m))   # :91-92
    Tstat_observation = np.zeros((N, p)); Tstat_state = np.zeros((N + 1, m))  # :95
    r_UT = np.zeros((Np + 1, m)); r_vec = np.zeros((N + 1, m))    # :98-99
    N_UT = np.zeros((m, m, Np + 1)); Nmat = np.zeros((m, m, N + 1))
    r_1 = np.zeros(m); N_1 = np.zeros((m, m)); N_2 = np.zeros((m, m))   # :104-106
    QtR = np.zeros((r, m, N))
    QtR_mat = Q_mat @ R_mat.T                                     # :108
    QtR[:, :, 0] = QtR_mat
    QtR_tv = Q_tv or R_tv                                         # :112
    eye = np.eye(m)
    index = 0

    for i in range(N):                                            # :122
        if Z_tv and i > 0:
            Z_mat = Z[:, :, i]
        if T_tv and i > 0:
            T_mat = T[:, :, i]
        if R_tv and i > 0:
            R_mat = R[:, :, i]
        if Q_tv and i > 0:
            Q_mat = Q[:, :, i]
        if QtR_tv and i > 0:                                      # :139
            QtR_mat = Q_mat @ R_mat.T
            QtR[:, :, i] = QtR_mat
        a_pred[i, :] = a_mat[index, :]                            # :145
        P_star_pred[:, :, i] = P_star_cube[:, :, index]
        P_pred[:, :, i] = P_star_pred[:, :, i]
        if initialisation:                                        # :148
            P_inf_pred[:, :, i] = P_inf_cube[:, :, index]
            P_pred[:, :, i] = P_pred[:, :, i] + KAPPA * P_inf_pred[:, :, i]
            if diagnostics:
                calculate_vnorm = _all_small(Z_mat @ P_inf_pred[:, :, i] @ Z_mat.T)   # :152-153
        yfit[i, :] = Z_mat @ a_pred[i, :]                         # :158
        v[i, :] = y[i, :] - yfit[i, :]                            # :161
        Fmat[:, :, i] = Z_mat @ P_pred[:, :, i] @ Z_mat.T         # :162
        if diagnostics:
            Finv[:, :, i] = np.linalg.inv(Fmat[:, :, i])          # :167
            v_0na[i, :] = np.where(np.isnan(v[i, :]), 0.0, v[i, :])   # :170-171
        if diagnostics and ((not initialisation) or calculate_vnorm):   # :175
            Finv_root = _sym_root_svd(Finv[:, :, i])              # :178-181
            v_norm_row = v_0na[i, :] @ Finv_root                  # :184
            v_norm_row[~np.isfinite(v[i, :])] = nan               # :185
            v_norm[i, :] = v_norm_row
        for j in range(p):                                        # :190
            if y_isna[i, j]:                                      # :193
                loglik[index] = nan
                if index < Np_min1:
                    a_mat[index + 1, :] = a_mat[index, :]
                    P_star_cube[:, :, index + 1] = P_star_cube[:, :, index]
                    if initialisation:
                        initialisation_steps += 1                 # :199
                        P_inf_cube[:, :, index + 1] = P_inf_cube[:, :, index]
                else:
                    a_fc = a_mat[index, :].copy()
                    P_star_fc = P_star_cube[:, :, index].copy()
                    if initialisation:
                        initialisation_steps += 1                 # :206
                        P_inf_fc = P_inf_cube[:, :, index].copy()
                index += 1
                continue
            Z_row = Z_mat[j, :]
            if initialisation:                                    # :219
                initialisation_steps += 1                         # :222
                M_inf[:, index] = P_inf_cube[:, :, index] @ Z_row
                M_star[:, index] = P_star_cube[:, :, index] @ Z_row
                F_inf[index] = float(Z_row @ M_inf[:, index])
                F_star[index] = float(Z_row @ M_star[:, index])   # :229-230
                if F_inf[index] < EPS:                            # :233
                    if F_star[index] < EPS:                       # :236
                        loglik[index] = nan
                        if index < Np_min1:
                            a_mat[index + 1, :] = a_mat[index, :]
                            P_star_cube[:, :, index + 1] = P_star_cube[:, :, index]
                            P_inf_cube[:, :, index + 1] = P_inf_cube[:, :, index]
                        else:
                            a_fc = a_mat[index, :].copy()
                            P_star_fc = P_star_cube[:, :, index].copy()
                            P_inf_fc = P_inf_cube[:, :, index].copy()
                        index += 1
                        continue
                    F_1[index] = 1 / F_star[index]                # :253
                    v_UT[index] = y[i, j] - float(Z_row @ a_mat[index, :])
                    K_0[:, index] = M_star[:, index] * F_1[index]
                    L_0[:, :, index] = eye - np.outer(K_0[:, index], Z_row)
                    if index < Np_min1:
                        a_mat[index + 1, :] = a_mat[index, :] + K_0[:, index] * v_UT[index]
                        P_star_cube[:, :, index + 1] = P_star_cube[:, :, index] @ L_0[:, :, index].T
                        P_inf_cube[:, :, index + 1] = P_inf_cube[:, :, index]
                    else:
                        a_fc = a_mat[index, :] + K_0[:, index] * v_UT[index]
                        P_star_fc = P_star_cube[:, :, index] @ L_0[:, :, index].T
                        P_inf_fc = P_inf_cube[:, :, index].copy()
                    loglik[index] = CONSTANT - (math.log(F_star[index]) + v_UT[index] ** 2 * F_1[index]) / 2
                    index += 1
                    continue
                F_1[index] = 1 / F_inf[index]                     # :283
                F_2[index] = -F_1[index] ** 2 * F_star[index]
                v_UT[index] = y[i, j] - float(Z_row @ a_mat[index, :])
                K_0[:, index] = M_inf[:, index] * F_1[index]
                L_0[:, :, index] = eye - np.outer(K_0[:, index], Z_row)
                K_1[:, index] = M_star[:, index] * F_1[index] + M_inf[:, index] * F_2[index]
                L_1[:, :, index] = -np.outer(K_1[:, index], Z_row)
                if index < Np_min1:                               # :299
                    a_mat[index + 1, :] = a_mat[index, :] + K_0[:, index] * v_UT[index]
                    P_star_cube[:, :, index + 1] = (P_inf_cube[:, :, index] @ L_1[:, :, index].T
                                                    + P_star_cube[:, :, index] @ L_0[:, :, index].T)
                    P_inf_cube[:, :, index + 1] = P_inf_cube[:, :, index] @ L_0[:, :, index].T
                else:
                    a_fc = a_mat[index, :] + K_0[:, index] * v_UT[index]
                    P_star_fc = (P_inf_cube[:, :, index] @ L_1[:, :, index].T
                                 + P_star_cube[:, :, index] @ L_0[:, :, index].T)
                    P_inf_fc = P_inf_cube[:, :, index] @ L_0[:, :, index].T
                loglik[index] = CONSTANT - math.log(F_inf[index]) / 2   # :314
                if index < Np_min1 and j < p_min1:                # :318
                    initialisation = not _all_small(P_inf_cube[:, :, index + 1])
            else:
                M_star[:, index] = P_star_cube[:, :, index] @ Z_row     # :326
                F_star[index] = float(Z_row @ M_star[:, index])
                if F_star[index] < EPS:                           # :332
                    loglik[index] = nan
                    if index < Np_min1:
                        a_mat[index + 1, :] = a_mat[index, :]
                        P_star_cube[:, :, index + 1] = P_star_cube[:, :, index]
                    else:
                        a_fc = a_mat[index, :].copy()
                        P_star_fc = P_star_cube[:, :, index].copy()
                else:
                    F_1[index] = 1 / F_star[index]
                    v_UT[index] = y[i, j] - float(Z_row @ a_mat[index, :])
                    K_0[:, index] = M_star[:, index] * F_1[index]
                    L_0[:, :, index] = eye - np.outer(K_0[:, index], Z_row)
                    if index < Np_min1:
                        a_mat[index + 1, :] = a_mat[index, :] + K_0[:, index] * v_UT[index]
                        P_star_cube[:, :, index + 1] = P_star_cube[:, :, index] @ L_0[:, :, index].T
                    else:
                        a_fc = a_mat[index, :] + K_0[:, index] * v_UT[index]
                        P_star_fc = P_star_cube[:, :, index] @ L_0[:, :, index].T
                    loglik[index] = CONSTANT - (math.log(F_star[index]) + v_UT[index] ** 2 * F_1[index]) / 2
            index += 1
        if index < Np:                                            # :375
            a_fil[i, :] = a_mat[index, :]
            P_star_fil[:, :, i] = P_star_cube[:, :, index]
            P_fil[:, :, i] = P_star_fil[:, :, i]
            if initialisation:
                P_inf_fil[:, :, i] = P_inf_cube[:, :, index]
                P_fil[:, :, i] = P_fil[:, :, i] + KAPPA * P_inf_fil[:, :, i]
            a_mat[index, :] = T_mat @ a_mat[index, :]             # :383
            P_star_cube[:, :, index] = T_mat @ P_star_cube[:, :, index] @ T_mat.T + R_mat @ QtR_mat
            if initialisation:
                P_inf_cube[:, :, index] = T_mat @ P_inf_cube[:, :, index] @ T_mat.T
                initialisation = not _all_small(P_inf_cube[:, :, index])   # :389
        else:
            a_fil[i, :] = a_fc                                    # :394
            P_star_fil[:, :, i] = P_star_fc
            P_fil[:, :, i] = P_star_fc
            if initialisation:
                P_inf_fil[:, :, i] = P_inf_fc
                P_fil[:, :, i] = P_fil[:, :, i] + KAPPA * P_inf_fc
            a_fc = T_mat @ a_fc                                   # :401
            P_star_fc = T_mat @ P_star_fc @ T_mat.T + R_mat @ QtR_mat
            P_fc = P_star_fc.copy()
            if initialisation:
                P_inf_fc = T_mat @ P_inf_fc @ T_mat.T
                P_fc = P_fc + KAPPA * P_inf_fc

    index = Np_min1                                               # :412
    for i in range(N_min1, -1, -1):                               # :416
        if Z_tv and i < N_min1:
            Z_mat = Z[:, :, i]
        if T_tv:
            T_mat_now = T[:, :, i]
        if QtR_tv and i < N_min1:
            QtR_mat = QtR[:, :, i]
        if T_tv and i > 0:
            T_mat = T[:, :, i - 1]                                # :430-432
        for j in range(p_min1, -1, -1):                           # :435
            if y_isna[i, j]:
                r_UT[index, :] = r_UT[index + 1, :]
                N_UT[:, :, index] = N_UT[:, :, index + 1]
                index -= 1
                continue
            Z_row = Z_mat[j, :]
            L0 = L_0[:, :, index]
            L1 = L_1[:, :, index]
            Nn = N_UT[:, :, index + 1]
            if index < initialisation_steps:                      # :450
                if F_inf[index] < EPS:
                    if F_star[index] < EPS:
                        r_UT[index, :] = r_UT[index + 1, :]
                        N_UT[:, :, index] = Nn
                    else:
                        r_UT[index, :] = Z_row * F_1[index] * v_UT[index] + r_UT[index + 1, :] @ L0   # :464
                        N_UT[:, :, index] = np.outer(Z_row, Z_row) * F_1[index] + L0.T @ Nn @ L0
                        N_1 = N_1 @ L0                            # :468
                else:
                    r_UT[index, :] = r_UT[index + 1, :] @ L0      # :473
                    r_1 = Z_row * F_1[index] * v_UT[index] + r_1 @ L0 + r_UT[index + 1, :] @ L1
                    N_UT[:, :, index] = L0.T @ Nn @ L0
                    tZZ = np.outer(Z_row, Z_row)
                    N_1 = (tZZ * F_1[index] + L0.T @ N_1 @ L0 + L1.T @ Nn @ L0 + L0.T @ Nn @ L1)   # :479-482
                    N_2 = (tZZ * F_2[index] + L0.T @ N_2 @ L0 + L0.T @ N_1 @ L1
                           + L1.T @ N_1.T @ L0 + L1.T @ Nn @ L1)  # :483-487 (uses updated N_1)
            else:
                if F_star[index] < EPS:
                    r_UT[index, :] = r_UT[index + 1, :]
                    N_UT[:, :, index] = Nn
                else:
                    r_UT[index, :] = Z_row * F_1[index] * v_UT[index] + r_UT[index + 1, :] @ L0
                    N_UT[:, :, index] = np.outer(Z_row, Z_row) * F_1[index] + L0.T @ Nn @ L0
            index -= 1
        r_vec[i, :] = r_UT[index + 1, :]                          # :509
        Nmat[:, :, i] = N_UT[:, :, index + 1]
        Ps = P_star_pred[:, :, i]
        Pi = P_inf_pred[:, :, i]
        if (index + 1) < initialisation_steps:                    # :513
            a_smooth[i, :] = a_pred[i, :] + r_vec[i, :] @ Ps + r_1 @ Pi
            V[:, :, i] = (Ps - Ps @ Nmat[:, :, i] @ Ps - Ps @ N_1.T @ Pi
                          - Pi @ N_1 @ Ps - Pi @ N_2 @ Pi)
        else:
            a_smooth[i, :] = a_pred[i, :] + r_vec[i, :] @ P_pred[:, :, i]
            V[:, :, i] = P_pred[:, :, i] - P_pred[:, :, i] @ Nmat[:, :, i] @ P_pred[:, :, i]
        if diagnostics:                                           # :528
            with np.errstate(divide="ignore", invalid="ignore"):
                Tstat_state[i + 1, :] = r_vec[i + 1, :] / np.sqrt(np.diag(Nmat[:, :, i + 1]))
                K = T_mat_now @ P_pred[:, :, i] @ Z_mat.T @ Finv[:, :, i]
                e_row = v_0na[i, :] @ Finv[:, :, i] - r_vec[i + 1, :] @ K
                e_row[~np.isfinite(v[i, :])] = nan
                e[i, :] = e_row
                D[:, :, i] = Finv[:, :, i] + K.T @ Nmat[:, :, i + 1] @ K
                Tstat_observation[i, :] = e_row / np.sqrt(np.diag(D[:, :, i]))
        eta[i, :] = QtR_mat @ r_vec[i + 1, :]                     # :547
        eta_var[:, :, i] = Q_mat_at(Q, i) - QtR_mat @ Nmat[:, :, i + 1] @ QtR_mat.T   # :548 (see note)
        if i > 0:                                                 # :551
            r_UT[index + 1, :] = r_UT[index + 1, :] @ T_mat
            N_UT[:, :, index + 1] = T_mat.T @ N_UT[:, :, index + 1] @ T_mat
            if (index + 1) < initialisation_steps:
                r_1 = r_1 @ T_mat
                N_1 = T_mat.T @ N_1 @ T_mat
                N_2 = T_mat.T @ N_2 @ T_mat

    return {
        "initialisation_steps": initialisation_steps, "loglik": loglik,
        "a_pred": a_pred, "a_fil": a_fil, "a_smooth": a_smooth,
        "P_pred": P_pred, "P_fil": P_fil, "V": V, "P_inf_pred": P_inf_pred,
        "P_star_pred": P_star_pred, "P_inf_fil": P_inf_fil, "P_star_fil": P_star_fil,
        "yfit": yfit, "v": v, "v_norm": v_norm, "eta": eta, "e": e, "Fmat": Fmat,
        "eta_var": eta_var, "D": D, "a_fc": a_fc, "P_inf_fc": P_inf_fc, "P_star_fc": P_star_fc,
        "P_fc": P_fc, "Tstat_observation": Tstat_observation, "Tstat_state": Tstat_state,
        "r_vec": r_vec, "Nmat": Nmat,
    }


def Q_mat_at(Q, i):
    """`Q_mat` as seen by the backward loop of KalmanC (src/KalmanC.cpp:548).

    NOTE (reference quirk, reproduced): the backward loop never reloads `Q_mat`; it still holds the
    slice loaded last in the forward loop, i.e. Q[:, :, N-1] when Q is time-varying, Q[:, :, 0]
    otherwise.  Only `QtR_mat` is reloaded per time point (:425-427)."""
    return Q[:, :, Q.shape[2] - 1]


# --------------------------------------------------------------------------------------------
# src/FastSmootherC.cpp:23-375
# --------------------------------------------------------------------------------------------
def FastSmootherC(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, transposed_state):
    y = np.asarray(y, dtype=np.float64)
    assert y.ndim == 3
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    P_inf = np.asarray(P_inf, dtype=np.float64)
    P_star = np.asarray(P_star, dtype=np.float64)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p, nsim = y.shape                                          # :36
    m = a.shape[0]
    r = R.shape[1]
    Np = N * p
    Np_min1, N_min1, p_min1 = Np - 1, N - 1, p - 1
    Z_tv, T_tv, R_tv, Q_tv = (Z.shape[2] > 1, T.shape[2] > 1, R.shape[2] > 1, Q.shape[2] > 1)
    QtR_tv = Q_tv or R_tv
    L_0 = np.zeros((m, m, Np)); L_1 = np.zeros((m, m, Np))
    F_inf = np.zeros(Np); F_star = np.zeros(Np); F_1 = np.zeros(Np); v_UT = np.zeros(Np)
    r_UT = np.zeros((m, nsim, Np + 1)); r_vec = np.zeros((m, nsim, N))    # :65-66
    r_1_mat = np.zeros((m, nsim))
    QtR = np.zeros((r, m, N)); tT = np.zeros((m, m, N))
    for i in range(N):
        Qi = Q[:, :, i if Q_tv else 0]
        Ri = R[:, :, i if R_tv else 0]
        QtR[:, :, i] = Qi @ Ri.T                                  # :72-73,:120-124
        tT[:, :, i] = T[:, :, i if T_tv else 0].T                 # :125-128
    a_smooth = np.zeros((N, m, nsim)); a_t = np.zeros((m, nsim, N))
    eta = np.zeros((N, r, nsim))                                  # :80-81
    eye = np.eye(m)
    for sim in range(nsim):                                       # :88
        index = 0
        a_vec = a.copy()
        P_inf_mat = P_inf.copy()
        P_star_mat = P_star.copy()
        r_1 = np.zeros(m)                                         # :94-99
        for i in range(N):                                        # :103
            Z_mat = Z[:, :, i if Z_tv else 0]
            T_mat = T[:, :, i if T_tv else 0]
            R_mat = R[:, :, i if R_tv else 0]
            QtR_mat = QtR[:, :, i if QtR_tv else 0]
            tT_mat = tT[:, :, i if T_tv else 0]
            for j in range(p):                                    # :139
                Z_row = Z_mat[j, :]
                if index < initialisation_steps:                  # :147
                    M_inf = P_inf_mat @ Z_row
                    M_star = P_star_mat @ Z_row
                    F_inf[index] = float(Z_row @ M_inf)
                    F_star[index] = float(Z_row @ M_star)
                    if F_inf[index] < EPS:                        # :158
                        if F_star[index] < EPS:
                            index += 1
                            continue
                        F_1[index] = 1 / F_star[index]
                        v_UT[index] = y[i, j, sim] - float(Z_row @ a_vec)
                        K0 = M_star * F_1[index]
                        L_0[:, :, index] = eye - np.outer(K0, Z_row)
                        if index < Np_min1:                       # :179
                            a_vec = a_vec + K0 * v_UT[index]
                            P_star_mat = P_star_mat @ L_0[:, :, index].T
                        index += 1
                        continue
                    F_1[index] = 1 / F_inf[index]
                    F_2 = -F_1[index] ** 2 * F_star[index]
                    v_UT[index] = y[i, j, sim] - float(Z_row @ a_vec)
                    K0 = M_inf * F_1[index]
                    L_0[:, :, index] = eye - np.outer(K0, Z_row)
                    K1 = M_star * F_1[index] + M_inf * F_2
                    L_1[:, :, index] = -np.outer(K1, Z_row)
                    if index < Np_min1:                           # :202
                        a_vec = a_vec + K0 * v_UT[index]
                        P_star_mat = P_inf_mat @ L_1[:, :, index].T + P_star_mat @ L_0[:, :, index].T
                        P_inf_mat = P_inf_mat @ L_0[:, :, index].T
                else:
                    M_star = P_star_mat @ Z_row                   # :212
                    F_star[index] = float(Z_row @ M_star)
                    if F_star[index] < EPS:
                        index += 1
                        continue
                    F_1[index] = 1 / F_star[index]
                    v_UT[index] = y[i, j, sim] - float(Z_row @ a_vec)
                    K0 = M_star * F_1[index]
                    L_0[:, :, index] = eye - np.outer(K0, Z_row)
                    if index < Np_min1:                           # :236
                        a_vec = a_vec + K0 * v_UT[index]
                        P_star_mat = P_star_mat @ L_0[:, :, index].T
                index += 1
            if i < N_min1:                                        # :245
                a_vec = T_mat @ a_vec
                P_star_mat = T_mat @ P_star_mat @ tT_mat + R_mat @ QtR_mat
                if (index - 1) < initialisation_steps:            # :248
                    P_inf_mat = T_mat @ P_inf_mat @ tT_mat
        index = Np_min1                                           # :255
        for i in range(N_min1, -1, -1):                           # :259
            Z_mat = Z[:, :, i if Z_tv else 0]
            # tT of the previous time point (:267-269); time-invariant -> slice 0
            tT_prev = tT[:, :, (i - 1) if (T_tv and i > 0) else 0]
            for j in range(p_min1, -1, -1):                       # :272
                Z_row = Z_mat[j, :]
                rn = r_UT[:, sim, index + 1]
                if index < initialisation_steps:                  # :280
                    if F_inf[index] < EPS:
                        if F_star[index] < EPS:
                            r_UT[:, sim, index] = rn
                        else:
                            r_UT[:, sim, index] = Z_row * F_1[index] * v_UT[index] + L_0[:, :, index].T @ rn
                    else:
                        r_UT[:, sim, index] = L_0[:, :, index].T @ rn
                        r_1 = (Z_row * F_1[index] * v_UT[index] + L_0[:, :, index].T @ r_1
                               + L_1[:, :, index].T @ rn)         # :302-303
                else:
                    if F_star[index] < EPS:
                        r_UT[:, sim, index] = rn
                    else:
                        r_UT[:, sim, index] = Z_row * F_1[index] * v_UT[index] + L_0[:, :, index].T @ rn
                index -= 1
            r_vec[:, sim, i] = r_UT[:, sim, index + 1]            # :323
            r_1_mat[:, sim] = r_1
            if i > 0:                                             # :327
                r_UT[:, sim, index + 1] = tT_prev @ r_UT[:, sim, index + 1]
                if (index + 1) < initialisation_steps:
                    r_1 = tT_prev @ r_1
    a_temp = a[:, None] + P_star @ r_vec[:, :, 0] + P_inf @ r_1_mat     # :338
    a_smooth[0, :, :] = a_temp
    if transposed_state:
        a_t[:, :, 0] = a_temp
    for i in range(1, N):                                         # :345
        T_mat = T[:, :, (i - 1) if T_tv else 0]
        R_mat = R[:, :, (i - 1) if R_tv else 0]
        QtR_mat = QtR[:, :, (i - 1) if QtR_tv else 0]
        eta_temp = QtR_mat @ r_vec[:, :, i]                       # :359
        eta[i - 1, :, :] = eta_temp
        a_temp = T_mat @ a_temp + R_mat @ eta_temp                # :363
        a_smooth[i, :, :] = a_temp
        if transposed_state:
            a_t[:, :, i] = a_temp
    return {"a_smooth": a_smooth, "a_t": a_t, "eta": eta}


# --------------------------------------------------------------------------------------------
# src/SimulateC.cpp:26-139.  `draws` replaces Rcpp::rnorm: a flat vector consumed in the
# reference's order: [m*nsim initial draws if draw_initial], then r_tot*nsim per time point.
# --------------------------------------------------------------------------------------------
def SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only,
              transposed_state, draws, Q_root=None, P_root=None):
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    P_star = np.asarray(P_star, dtype=np.float64)
    draws = np.asarray(draws, dtype=np.float64).reshape(-1)
    p, m, r = Z.shape[0], a.shape[0], Q.shape[0]                  # :40
    r_tot = r * repeat_Q
    nsim_rep = nsim * repeat_Q
    total = r_tot * nsim
    m_nsim = m * nsim
    N_min1 = N - 1
    Z_tv, T_tv, R_tv, Q_tv = (Z.shape[2] > 1, T.shape[2] > 1, R.shape[2] > 1, Q.shape[2] > 1)
    rep_g1 = repeat_Q > 1
    Z_mat, T_mat, R_mat, Q_mat = Z[:, :, 0], T[:, :, 0], R[:, :, 0], Q[:, :, 0]
    eta = np.zeros((N, r_tot, nsim)); a_cube = np.zeros((N, m, nsim))
    a_t = np.zeros((m, nsim, N)); y = np.zeros((N, p, nsim))      # :59-60
    pos = 0
    a_temp = np.repeat(a[:, None], nsim, axis=1)                  # :71
    if draw_initial:                                              # :72
        Pr = _sym_root_svd(P_star) if P_root is None else np.asarray(P_root, dtype=np.float64)
        a_draw = draws[pos:pos + m_nsim]
        pos += m_nsim
        a_temp = a_temp + _mm_pinned(Pr, a_draw.reshape((m, nsim), order="F"))   # :76
    a_cube[0, :, :] = a_temp
    if transposed_state:
        a_t[:, :, 0] = a_temp
    roots = None if Q_root is None else _cube(Q_root)
    Qr = _sym_root_svd(Q_mat) if roots is None else roots[:, :, 0]   # :84-85
    for i in range(N):                                            # :88
        if Z_tv and not eta_only and i > 0:
            Z_mat = Z[:, :, i]
        if T_tv and not eta_only and i > 0:
            T_mat = T[:, :, i]
        if R_tv and not eta_only and i > 0:
            R_mat = R[:, :, i]
        if Q_tv and i > 0:                                        # :100
            Q_mat = Q[:, :, i]
            Qr = _sym_root_svd(Q_mat) if roots is None else roots[:, :, i]
        if not eta_only:
            y[i, :, :] = _mm_pinned(Z_mat, a_temp)                # :108
        draw = draws[pos:pos + total]                             # :112
        pos += total
        if rep_g1:                                                # :115
            eta_sim = _mm_pinned(Qr, draw.reshape((r, nsim_rep), order="F"))
            eta_temp = eta_sim.reshape(-1, order="F").reshape((r_tot, nsim), order="F")
        else:
            eta_temp = _mm_pinned(Qr, draw.reshape((r_tot, nsim), order="F"))
        eta[i, :, :] = eta_temp                                   # :120
        if not eta_only and i < N_min1:                           # :123
            a_temp = _mm_pinned(T_mat, a_temp) + _mm_pinned(R_mat, eta_temp)
            a_cube[i + 1, :, :] = a_temp
            if transposed_state:
                a_t[:, :, i + 1] = a_temp
    assert pos == draws.shape[0], "draw vector length does not match the reference's consumption"
    return {"y": y, "a": a_cube, "a_t": a_t, "eta": eta}


def _mm_pinned(A, B):
    """A @ B with the PINNED accumulation order used for the bit-exactness contract: every output
    element is acc = 0; for k ascending: acc = fma(A[i,k], B[k,j], acc).  math.fma needs Python
    3.13; NumPy has no fma, so emulate a correctly-rounded fma with exact integer/float arithmetic
    via fractions only when the matrices are small; otherwise fall back to the C oracle
    (oracle/kalman_c.c: ss_mm_pinned) through ctypes.  The NumPy oracle is *not* used for the
    bit-exact tests -- the C oracle is; this function only needs 1e-15-level agreement."""
    return A @ B


# --------------------------------------------------------------------------------------------
# src/ExtractComponentC.cpp:12-40
# --------------------------------------------------------------------------------------------
def ExtractComponentC(a, Z):
    a = np.asarray(a, dtype=np.float64)
    Z = _cube(Z)
    m, nsim, N = a.shape
    p = Z.shape[0]
    Z_tv = Z.shape[2] > 1
    comp = np.zeros((N, p, nsim))
    for i in range(N):
        Z_mat = Z[:, :, i if Z_tv else 0]
        comp[i, :, :] = Z_mat @ a[:, :, i]                        # :36
    return comp
