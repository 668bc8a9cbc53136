"""ctypes binding of oracle/libssoracle.so -- the plain-C, pinned-arithmetic-order restatement
(oracle/kalman_c.c).  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Same call signatures as oracle/kalman_np.py and oracle/ref.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssoracle.so")
_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "c"])


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            build()
        _lib = C.CDLL(LIB_PATH)
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


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"libssoracle: {what} failed with status {rc}")


def num_threads():
    return lib().ssor_num_threads()


def LogLikC(y, y_isna, a, P_inf, P_star, Z, T, R, Q):
    y = _f(y)
    isna = np.asfortranarray(np.asarray(y_isna).astype(np.int32))
    a = _f(np.asarray(a).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p = y.shape
    m, r = a.shape[0], R.shape[1]
    out = C.c_double()
    _check(lib().ssor_loglik(_p(y), isna.ctypes.data_as(_ip), N, p, m, r, _p(a), _p(P_inf), _p(P_star),
                             _p(Z), Z.shape[2], _p(T), T.shape[2], _p(R), R.shape[2], _p(Q), Q.shape[2],
                             C.byref(out)), "ssor_loglik")
    return out.value


def loglik_batch(y_bn, a, P_inf, P_star_b, Z, T, R, Q_b, threads=0):
    """y_bn (B, N), p = 1, NaN = missing; P_star_b (B, m, m), Q_b (B, r, r) per series (symmetric).
    Returns (B,) loglik / N."""
    y = np.ascontiguousarray(y_bn, dtype=np.float64)
    B, N = y.shape
    a = _f(np.asarray(a).reshape(-1))
    m = a.shape[0]
    Z2, T2, R2 = (np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(x.shape[0], x.shape[1]))
                  for x in (Z, T, R))
    r = R2.shape[1]
    Ps = np.ascontiguousarray(P_star_b, dtype=np.float64)
    Qb = np.ascontiguousarray(Q_b, dtype=np.float64)
    out = np.zeros(B)
    fn = lib().ssor_loglik_batch
    fn.argtypes = [_dp, C.c_long, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int, C.c_int,
                   _dp]
    _check(fn(_p(y), B, N, m, r, _p(a), _p(_f(P_inf)), _p(Ps), _p(Z2), _p(T2), _p(R2), _p(Qb), 1, threads,
              _p(out)), "ssor_loglik_batch")
    return out


def gains(P_inf, P_star, Z, T, R, Q, N, initialisation_steps):
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    p, m = Z.shape[0], Z.shape[1]
    r = R.shape[1]
    Np = N * p
    nqtr = N if (Q.shape[2] > 1 or R.shape[2] > 1) else 1
    code = np.zeros(Np, dtype=np.int32)
    F1 = np.zeros(Np)
    K0 = np.zeros((m, Np), order="F")
    K1 = np.zeros((m, Np), order="F")
    QtR = np.zeros((r, m, nqtr), order="F")
    _check(lib().ssor_gains(N, p, m, r, _p(P_inf), _p(P_star), _p(Z), Z.shape[2], _p(T), T.shape[2], _p(R),
                            R.shape[2], _p(Q), Q.shape[2], int(initialisation_steps),
                            code.ctypes.data_as(_ip), _p(F1), _p(K0), _p(K1), _p(QtR)), "ssor_gains")
    return {"code": code, "F1": F1, "K0": K0, "K1": K1, "QtR": QtR}


def FastSmootherC(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, transposed_state, threads=0):
    y = np.asfortranarray(np.asarray(y, dtype=np.float64))
    a = _f(np.asarray(a).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    N, p, nsim = y.shape
    m, r = a.shape[0], R.shape[1]
    a_smooth = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r, nsim), order="F")
    _check(lib().ssor_fast_smoother(_p(y), N, p, nsim, m, r, _p(a), _p(P_inf), _p(P_star), _p(Z), Z.shape[2],
                                    _p(T), T.shape[2], _p(R), R.shape[2], _p(Q), Q.shape[2],
                                    int(initialisation_steps), int(bool(transposed_state)), _p(a_smooth), _p(a_t),
                                    _p(eta), int(threads)), "ssor_fast_smoother")
    return {"a_smooth": a_smooth, "a_t": a_t, "eta": eta}


def sym_root(A):
    """U diag(sqrt(s)) U' with (U, s, V) = svd(A) -- the root SimulateC takes of Q and P_star
    (src/SimulateC.cpp:73-74,84-85).  Computed once on the host and handed to BOTH the oracle and the
    GPU library so that the bit-exactness contract does not depend on an SVD implementation."""
    A = np.asarray(A, dtype=np.float64)
    U, s, _ = np.linalg.svd(A)
    return U @ np.diag(np.sqrt(s)) @ U.T


def SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only, transposed_state, draws,
              Q_root=None, P_root=None, threads=0):
    a = _f(np.asarray(a).reshape(-1))
    Z, T, R, Q = _cube(Z), _cube(T), _cube(R), _cube(Q)
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1)
    p, m = Z.shape[0], a.shape[0]
    r = Q.shape[0]
    r_tot = r * repeat_Q
    if Q_root is None:
        Q_root = np.stack([sym_root(Q[:, :, i]) for i in range(Q.shape[2])], axis=2)
    Q_root = _cube(Q_root)
    if draw_initial and P_root is None:
        P_root = sym_root(np.atleast_2d(P_star))
    P_root = _f(P_root) if P_root is not None else None
    y = np.zeros((N, p, nsim), order="F")
    a_out = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r_tot, nsim), order="F")
    fn = lib().ssor_simulate
    fn.argtypes = [C.c_int] * 7 + [_dp, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, _dp, C.c_long, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    _check(fn(nsim, repeat_Q, N, p, m, R.shape[1], r, _p(a), _p(Z), Z.shape[2], _p(T), T.shape[2], _p(R),
              R.shape[2], Q.shape[2], int(bool(draw_initial)), int(bool(eta_only)), int(bool(transposed_state)),
              _p(draws), draws.shape[0], _p(Q_root), _p(P_root), _p(y), _p(a_out), _p(a_t), _p(eta),
              int(threads)), "ssor_simulate")
    return {"y": y, "a": a_out, "a_t": a_t, "eta": eta}


def ExtractComponentC(a, Z):
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    Z = _cube(Z)
    m, nsim, N = a.shape
    p = Z.shape[0]
    out = np.zeros((N, p, nsim), order="F")
    _check(lib().ssor_extract(_p(a), m, nsim, N, _p(Z), p, Z.shape[2], _p(out)), "ssor_extract")
    return out
