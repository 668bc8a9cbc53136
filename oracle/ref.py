"""ctypes binding of oracle/_ref/libssref.so -- the reference's own five C++ sources compiled
unmodified against oracle/refshim (TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT).

Same call signatures as oracle/kalman_np.py so tests can swap one for the other.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libssref.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ssref_last_error.restype = C.c_char_p
    return _lib


def _f(x):
    return np.asfortranarray(np.asarray(x, dtype=np.float64))


def _p(x):
    return x.ctypes.data_as(_dp) if x is not None else None


def _cube(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    return np.asfortranarray(x)


def _check(rc):
    if rc != 0:
        raise RuntimeError("libssref: " + lib().ssref_last_error().decode())


def num_threads():
    return lib().ssref_num_threads()


def LogLikC(y, y_isna, a, P_inf, P_star, Z, T, R, Q):
    y = _f(y)
    isna = np.asfortranarray(np.asarray(y_isna).astype(np.int32))
    a = _f(np.asarray(a).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p = y.shape
    m, r = a.shape[0], R.shape[1]
    out = C.c_double()
    _check(lib().ssref_loglik(_p(y), isna.ctypes.data_as(_ip), N, p, m, r, _p(a), _p(P_inf), _p(P_star),
                              _p(Z), Z.shape[2], _p(T), T.shape[2], _p(R), R.shape[2], _p(Q), Q.shape[2],
                              C.byref(out)))
    return out.value


def loglik_batch(y_bn, a, P_inf, P_star_b, Z, T, R, Q_b, threads=0):
    """y_bn: (B, N) p=1 series-major; P_star_b (B, m, m) and Q_b (B, r, r) per series (C-order rows =
    series; each matrix symmetric so layout order is immaterial).  Returns (B,) loglik/N."""
    y = np.ascontiguousarray(y_bn, dtype=np.float64)
    B, N = y.shape
    isna = np.ascontiguousarray(np.isnan(y).astype(np.int32))
    a = _f(np.asarray(a).reshape(-1))
    m = a.shape[0]
    Z2, T2, R2 = (np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(x.shape[0], x.shape[1]))
                  for x in (Z, T, R))
    r = R2.shape[1]
    Ps = np.ascontiguousarray(P_star_b, dtype=np.float64)
    Qb = np.ascontiguousarray(Q_b, dtype=np.float64)
    out = np.zeros(B)
    fn = lib().ssref_loglik_batch
    fn.argtypes = [_dp, _ip, C.c_long, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                   C.c_int, C.c_int, _dp]
    _check(fn(_p(y), isna.ctypes.data_as(_ip), B, N, 1, m, r, _p(a), _p(_f(P_inf)), _p(Ps), _p(Z2), _p(T2), _p(R2),
              _p(Qb), 1, threads, _p(out)))
    return out


class _KalmanOut(C.Structure):
    _names = ["loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil", "V", "P_inf_pred", "P_star_pred",
              "P_inf_fil", "P_star_fil", "yfit", "v", "v_norm", "eta", "e", "Fmat", "eta_var", "D", "a_fc",
              "P_inf_fc", "P_star_fc", "P_fc", "Tstat_observation", "Tstat_state", "r_vec", "Nmat"]
    _fields_ = [("initialisation_steps", _ip)] + [(n, _dp) for n in _names]


def kalman_shapes(N, p, m, r):
    return {
        "loglik": (N * p,), "a_pred": (N, m), "a_fil": (N, m), "a_smooth": (N, m),
        "P_pred": (m, m, N), "P_fil": (m, m, N), "V": (m, m, N), "P_inf_pred": (m, m, N),
        "P_star_pred": (m, m, N), "P_inf_fil": (m, m, N), "P_star_fil": (m, m, N),
        "yfit": (N, p), "v": (N, p), "v_norm": (N, p), "eta": (N, r), "e": (N, p), "Fmat": (p, p, N),
        "eta_var": (r, r, N), "D": (p, p, N), "a_fc": (m,), "P_inf_fc": (m, m), "P_star_fc": (m, m),
        "P_fc": (m, m), "Tstat_observation": (N, p), "Tstat_state": (N + 1, m), "r_vec": (N + 1, m),
        "Nmat": (m, m, N + 1),
    }


def KalmanC(y, y_isna, a, P_inf, P_star, Z, T, R, Q, diagnostics=False):
    y = _f(y)
    isna = np.asfortranarray(np.asarray(y_isna).astype(np.int32))
    a = _f(np.asarray(a).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p = y.shape
    m, r = a.shape[0], R.shape[1]
    out = {k: np.zeros(s, order="F") for k, s in kalman_shapes(N, p, m, r).items()}
    steps = C.c_int()
    st = _KalmanOut()
    st.initialisation_steps = C.pointer(steps)
    for k in _KalmanOut._names:
        setattr(st, k, _p(out[k]))
    _check(lib().ssref_kalman(_p(y), isna.ctypes.data_as(_ip), N, p, m, r, _p(a), _p(P_inf), _p(P_star),
                              _p(Z), Z.shape[2], _p(T), T.shape[2], _p(R), R.shape[2], _p(Q), Q.shape[2],
                              int(bool(diagnostics)), C.byref(st)))
    out["initialisation_steps"] = steps.value
    return out


def FastSmootherC(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, transposed_state):
    y = np.asfortranarray(np.asarray(y, dtype=np.float64))
    a = _f(np.asarray(a).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p, nsim = y.shape
    m, r = a.shape[0], R.shape[1]
    a_smooth = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r, nsim), order="F")
    _check(lib().ssref_fast_smoother(_p(y), N, p, nsim, m, r, _p(a), _p(P_inf), _p(P_star), _p(Z), Z.shape[2],
                                     _p(T), T.shape[2], _p(R), R.shape[2], _p(Q), Q.shape[2],
                                     int(initialisation_steps), int(bool(transposed_state)), _p(a_smooth), _p(a_t),
                                     _p(eta)))
    return {"a_smooth": a_smooth, "a_t": a_t, "eta": eta}


def SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only, transposed_state, draws):
    a = _f(np.asarray(a).reshape(-1))
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    P_star = _f(np.atleast_2d(P_star))
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1)
    p, m = Z.shape[0], a.shape[0]
    r = Q.shape[0]
    r_tot = r * repeat_Q
    y = np.zeros((N, p, nsim), order="F")
    a_out = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r_tot, nsim), order="F")
    fn = lib().ssref_simulate
    fn.argtypes = [C.c_int] * 8 + [_dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp,
                                   C.c_int, C.c_int, C.c_int, _dp, C.c_long, _dp, _dp, _dp, _dp]
    _check(fn(nsim, repeat_Q, N, p, m, R.shape[1], r, P_star.shape[0], _p(a), _p(Z), Z.shape[2], _p(T), T.shape[2],
              _p(R), R.shape[2], _p(Q), Q.shape[2], _p(P_star), int(bool(draw_initial)), int(bool(eta_only)),
              int(bool(transposed_state)), _p(draws), draws.shape[0], _p(y), _p(a_out), _p(a_t), _p(eta)))
    return {"y": y, "a": a_out, "a_t": a_t, "eta": eta}


def ExtractComponentC(a, Z):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    Z = _cube(Z)
    m, nsim, N = a.shape
    p = Z.shape[0]
    out = np.zeros((N, p, nsim), order="F")
    _check(lib().ssref_extract(_p(a), m, nsim, N, _p(Z), p, Z.shape[2], _p(out)))
    return out
