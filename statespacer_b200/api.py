"""Host-side mirror of the reference's native interface (R/RcppExports.R:11-99) on top of libssgpu.

The five functions keep the reference's names, argument order and meaning (arrays in R's shapes:
3-D system matrices are (rows, cols, 1|N)); results come back as NumPy arrays / dicts keyed like the
lists the reference returns (src/KalmanC.cpp:563-593, src/FastSmootherC.cpp:370-374,
src/SimulateC.cpp:133-138).  `loglik_batch` / `loglik_batch_dev` are the new batched entry points that
replace the sequential LogLikC calls of the optim loop (R/statespacer.R:917-927,977-982).

Errors: inconsistent shapes raise (the reference throws through BEGIN_RCPP/END_RCPP); numerical
degeneracy is not an error (NaN where the reference returns NA).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import native
from .native import (KERNEL_AUTO, KERNEL_BLK, KERNEL_BLK_LANES, KERNEL_COLS, KERNEL_CORE, KERNEL_GENERIC, KERNEL_MMA, KERNEL_MMA_DENSE,  # noqa: F401
                     KERNEL_WSM)

_KERNELS = {"auto": KERNEL_AUTO, "generic": KERNEL_GENERIC, "cols": KERNEL_COLS, "mma": KERNEL_MMA,
            "mma_dense": KERNEL_MMA_DENSE, "wsm": KERNEL_WSM, "blk": KERNEL_BLK, "blk_lanes": KERNEL_BLK_LANES, "core": KERNEL_CORE}


def _f(x):
    return np.asfortranarray(np.asarray(x, dtype=np.float64))


def _cube(x, name):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2:
        x = x[:, :, None]
    if x.ndim != 3:
        raise ValueError(f"{name} must be a matrix or a 3-D array")
    return np.asfortranarray(x)


def _ptr(x):
    return None if x is None else x.ctypes.data


def _dims(a, P_inf, P_star, Z, T, R, Q, N):
    m = a.shape[0]
    p, r = Z.shape[0], R.shape[1]
    ok = (P_inf.shape == (m, m) and P_star.shape == (m, m) and Z.shape[:2] == (p, m) and T.shape[:2] == (m, m)
          and R.shape[0] == m and Q.shape[:2] == (r, r))
    if not ok:
        raise ValueError("inconsistent system matrix dimensions")
    for X, nm in ((Z, "Z"), (T, "T"), (R, "R"), (Q, "Q")):
        if X.shape[2] not in (1, N):
            raise ValueError(f"{nm} must have 1 or N slices")
    return m, p, r


def _prep(a, P_inf, P_star, Z, T, R, Q):
    a = _f(np.asarray(a, dtype=np.float64).reshape(-1))
    return a, _f(P_inf), _f(P_star), _cube(Z, "Z"), _cube(T, "T"), _cube(R, "R"), _cube(Q, "Q")


def use_devices(n=0):
    """Shard the host-pointer batch entry points over n devices of this process (0 = all visible); returns the count."""
    native.check(native.lib().ssgpu_use_devices(int(n)))
    out = C.c_int(0)
    native.check(native.lib().ssgpu_devices_in_use(C.byref(out)))
    return out.value


def last_kernel() -> str:
    return native.lib().ssgpu_last_kernel().decode()


def launch_count() -> int:
    return int(native.lib().ssgpu_launch_count())


def set_device(device: int) -> None:
    native.check(native.lib().ssgpu_set_device(int(device)))


def fp64_peak(mode=0, iters=2000, repeats=5):
    """Measured FP64 peak of the current device in TFLOP/s (mode 0 DFMA, 1 DMMA, 2 both)."""
    tf, ms = C.c_double(), C.c_double()
    native.check(native.lib().ssgpu_fp64_peak(mode, iters, repeats, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def plan_state_order(T):
    """Internal state order of the tensor kernel for a shared transition matrix T (host-only helper):
    (perm, tile_mask, dmma_per_step), see include/ssgpu.h."""
    T = _f(np.asarray(T, dtype=np.float64).reshape(np.shape(T)[0], np.shape(T)[1]))
    m = T.shape[0]
    perm = np.zeros(m, dtype=np.int32)
    mask, n = C.c_uint(), C.c_int()
    native.check(native.lib().ssgpu_plan_state_order(T.ctypes.data, m, perm.ctypes.data, C.byref(mask), C.byref(n)))
    return perm, mask.value, n.value


# ---- the five reference routines ------------------------------------------------------------------
def LogLikC(y, y_isna, a, P_inf, P_star, Z, T, R, Q):
    """src/LogLikC.cpp:20-215: loglik / N, NaN where the reference returns NA_REAL."""
    y = _f(y)
    if y.ndim != 2:
        raise ValueError("y must be an N x p matrix")
    isna = None if y_isna is None else np.asfortranarray(np.asarray(y_isna).astype(np.int32))
    a, P_inf, P_star, Z, T, R, Q = _prep(a, P_inf, P_star, Z, T, R, Q)
    N = y.shape[0]
    m, p, r = _dims(a, P_inf, P_star, Z, T, R, Q, N)
    if y.shape[1] != p or (isna is not None and isna.shape != y.shape):
        raise ValueError("y / y_isna do not match Z")
    out = C.c_double()
    native.check(native.lib().ssgpu_LogLikC(_ptr(y), _ptr(isna), N, p, m, r, _ptr(a), _ptr(P_inf), _ptr(P_star),
                                            _ptr(Z), Z.shape[2], _ptr(T), T.shape[2], _ptr(R), R.shape[2],
                                            _ptr(Q), Q.shape[2], C.byref(out)))
    return out.value


def kalman_shapes(N, p, m, r, diagnostics=False):
    extra = {"v_norm": (N, p), "e": (N, p), "D": (p, p, N), "Tstat_observation": (N, p),
             "Tstat_state": (N + 1, m)} if diagnostics else {}
    return {**extra,
        "loglik": (N * p,), "a_pred": (N, m), "a_fil": (N, m), "a_smooth": (N, m),
        "P_pred": (m, m, N), "P_fil": (m, m, N), "V": (m, m, N), "P_inf_pred": (m, m, N),
        "P_star_pred": (m, m, N), "P_inf_fil": (m, m, N), "P_star_fil": (m, m, N),
        "yfit": (N, p), "v": (N, p), "eta": (N, r), "Fmat": (p, p, N), "eta_var": (r, r, N), "a_fc": (m,),
        "P_inf_fc": (m, m), "P_star_fc": (m, m), "P_fc": (m, m), "r_vec": (N + 1, m), "Nmat": (m, m, N + 1),
    }


def KalmanC(y, y_isna, a, P_inf, P_star, Z, T, R, Q, diagnostics=False):
    """src/KalmanC.cpp:21-594: filter + smoother (+ diagnostics) for one series."""
    y = _f(y)
    isna = None if y_isna is None else np.asfortranarray(np.asarray(y_isna).astype(np.int32))
    a, P_inf, P_star, Z, T, R, Q = _prep(a, P_inf, P_star, Z, T, R, Q)
    N = y.shape[0]
    m, p, r = _dims(a, P_inf, P_star, Z, T, R, Q, N)
    if y.shape[1] != p:
        raise ValueError("y does not match Z")
    out = {k: np.zeros(s, order="F") for k, s in kalman_shapes(N, p, m, r, diagnostics).items()}
    steps = C.c_int()
    st = native.KalmanOut()
    st.initialisation_steps = C.pointer(steps)
    for k, arr in out.items():
        setattr(st, k, arr.ctypes.data_as(native.dp))
    native.check(native.lib().ssgpu_KalmanC(_ptr(y), _ptr(isna), N, p, m, r, _ptr(a), _ptr(P_inf), _ptr(P_star),
                                            _ptr(Z), Z.shape[2], _ptr(T), T.shape[2], _ptr(R), R.shape[2],
                                            _ptr(Q), Q.shape[2], int(bool(diagnostics)), C.byref(st)))
    out["initialisation_steps"] = steps.value
    return out


def FastSmootherC(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, transposed_state):
    """src/FastSmootherC.cpp:23-375: y is N x p x nsim."""
    y = np.asfortranarray(np.asarray(y, dtype=np.float64))
    if y.ndim != 3:
        raise ValueError("y must be N x p x nsim")
    a, P_inf, P_star, Z, T, R, Q = _prep(a, P_inf, P_star, Z, T, R, Q)
    N, p_y, nsim = y.shape
    m, p, r = _dims(a, P_inf, P_star, Z, T, R, Q, N)
    if p_y != p:
        raise ValueError("y does not match Z")
    a_smooth = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r, nsim), order="F")
    native.check(native.lib().ssgpu_FastSmootherC(
        _ptr(y), N, p, nsim, m, r, _ptr(a), _ptr(P_inf), _ptr(P_star), _ptr(Z), Z.shape[2], _ptr(T), T.shape[2],
        _ptr(R), R.shape[2], _ptr(Q), Q.shape[2], int(initialisation_steps), int(bool(transposed_state)),
        _ptr(a_smooth), _ptr(a_t) if transposed_state else None, _ptr(eta)))
    return {"a_smooth": a_smooth, "a_t": a_t, "eta": eta}


def sym_root(A):
    A = _f(np.atleast_2d(A))
    out = np.zeros_like(A, order="F")
    native.check(native.lib().ssgpu_sym_root(_ptr(A), A.shape[0], _ptr(out)))
    return out


def SimulateC(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only, transposed_state, draws,
              Q_root=None, P_root=None):
    """src/SimulateC.cpp:26-139 with the normal variates injected: `draws` are the values Rcpp::rnorm
    would return, in the reference's call order ([m*nsim initial draws], then r_tot*nsim per time point)."""
    a = _f(np.asarray(a, dtype=np.float64).reshape(-1))
    Z, T, R, Q = _cube(Z, "Z"), _cube(T, "T"), _cube(R, "R"), _cube(Q, "Q")
    P_star = _f(np.atleast_2d(P_star))
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1)
    p, m, r = Z.shape[0], a.shape[0], Q.shape[0]
    r_tot = r * repeat_Q
    Q_root = None if Q_root is None else _cube(Q_root, "Q_root")
    P_root = None if P_root is None else _f(P_root)
    y = np.zeros((N, p, nsim), order="F")
    a_out = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    eta = np.zeros((N, r_tot, nsim), order="F")
    native.check(native.lib().ssgpu_SimulateC(
        nsim, repeat_Q, N, p, m, R.shape[1], r, P_star.shape[0], _ptr(a), _ptr(Z), Z.shape[2], _ptr(T), T.shape[2],
        _ptr(R), R.shape[2], _ptr(Q), Q.shape[2], _ptr(P_star), int(bool(draw_initial)), int(bool(eta_only)),
        int(bool(transposed_state)), _ptr(draws), draws.shape[0], _ptr(Q_root), _ptr(P_root), _ptr(y), _ptr(a_out),
        _ptr(a_t) if transposed_state else None, _ptr(eta)))
    return {"y": y, "a": a_out, "a_t": a_t, "eta": eta}


def ExtractComponentC(a, Z):
    """src/ExtractComponentC.cpp:12-40: a is m x nsim x N."""
    a = np.asfortranarray(np.asarray(a, dtype=np.float64))
    Z = _cube(Z, "Z")
    if a.ndim != 3 or Z.shape[1] != a.shape[0] or Z.shape[2] not in (1, a.shape[2]):
        raise ValueError("a must be m x nsim x N and match Z")
    m, nsim, N = a.shape
    p = Z.shape[0]
    out = np.zeros((N, p, nsim), order="F")
    native.check(native.lib().ssgpu_ExtractComponentC(_ptr(a), m, nsim, N, _ptr(Z), p, Z.shape[2], _ptr(out)))
    return out


def plan_thread_kernel(sm, N):
    """Host only: what the thread-per-evaluation kernel does with the univariate model `sm` (shared matrices) over N time
    points -- dict(n_blocks, residual_dropped, level_slope_by_sums, zero_z_skipped, shared_diffuse_steps, flop_per_step)."""
    f = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))
    T, R, Q, Z = f(sm.T[:, :, 0]), f(sm.R[:, :, 0]), f(sm.Q[:, :, 0]), f(sm.Z[0, :, 0])
    a, Pi, Ps = f(sm.a), f(sm.P_inf), f(sm.P_star)
    out = [C.c_int(0) for _ in range(6)]
    native.check(native.lib().ssgpu_plan_thread_kernel(_ptr(T), _ptr(R), _ptr(Q), _ptr(Z), _ptr(a), _ptr(Pi), _ptr(Ps), T.shape[0],
                                                       R.shape[1], int(N), *[C.addressof(o) for o in out]))
    keys = ("n_blocks", "residual_dropped", "level_slope_by_sums", "zero_z_skipped", "shared_diffuse_steps", "flop_per_step")
    return dict(zip(keys, (o.value for o in out)))


def plan_core(sm):
    """Host only: what the core kernel does with the model `sm` (shared, time-invariant matrices) --
    dict(eligible, n_stochastic, n_deterministic, diffuse)."""
    f = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))
    T, R, Q, Z = f(sm.T[:, :, 0]), f(sm.R[:, :, 0]), f(sm.Q[:, :, 0]), f(sm.Z[:, :, 0])
    a, Pi, Ps = f(sm.a), f(sm.P_inf), f(sm.P_star)
    out = [C.c_int(0) for _ in range(4)]
    native.check(native.lib().ssgpu_plan_core(_ptr(T), _ptr(R), _ptr(Q), _ptr(Z), _ptr(a), _ptr(Pi), _ptr(Ps), T.shape[0],
                                              R.shape[1], Z.shape[0], *[C.addressof(o) for o in out]))
    return dict(zip(("eligible", "n_stochastic", "n_deterministic", "diffuse"), (o.value for o in out)))


def core_flops_per_step(m, p):
    """FP64 operations per series and time point of the core kernel's regular phase (0 = outside its range)."""
    return int(native.lib().ssgpu_core_flops_per_step(int(m), int(p)))


def plan_blocks(T, R, Q=None, p=1):
    """Block decomposition of the block-row kernel for shared T (m x m), R (m x r) and the pattern of Q (None =
    diagonal) -- host-only helper.  Returns None when the model does not decompose, else a dict with n_blocks, lanes,
    flop_per_step and perm (internal position -> state index, -1 = padding)."""
    T = _f(np.asarray(T, dtype=np.float64).reshape(np.shape(T)[0], np.shape(T)[1]))
    R = _f(np.asarray(R, dtype=np.float64).reshape(np.shape(R)[0], np.shape(R)[1]))
    Qm = None if Q is None else _f(np.asarray(Q, dtype=np.float64).reshape(np.shape(Q)[0], np.shape(Q)[1]))
    perm = np.full(32, -1, dtype=np.int32)
    nb, lanes, flops = C.c_int(), C.c_int(), C.c_int()
    native.check(native.lib().ssgpu_plan_blocks(_ptr(T), _ptr(R), _ptr(Qm), T.shape[0], R.shape[1], int(p), perm.ctypes.data,
                                                C.addressof(nb), C.addressof(lanes), C.addressof(flops)))
    if nb.value == 0:
        return None
    return {"n_blocks": nb.value, "lanes": lanes.value, "flop_per_step": flops.value, "perm": perm[:2 * nb.value]}


def collapse(y, H, Z):
    """Collapse transform of R/statespacer.R:815-900 (include/ssgpu.h section 6e).  y N x p (NaN = missing),
    H p x p x {1|N}, Z p x m2 x {1|N}.  Returns y_star (N x m2, NaN where R writes NA), H_star (m2 x m2 x {1|N}),
    loglik_add."""
    y = _f(y)
    H, Z = _cube(H, "H"), _cube(Z, "Z")
    N, p = y.shape
    m2 = Z.shape[1]
    ns = N if (H.shape[2] > 1 or Z.shape[2] > 1) else 1
    y_star = np.zeros((N, m2), order="F")
    H_star = np.zeros((m2, m2, ns), order="F")
    la = C.c_double()
    native.check(native.lib().ssgpu_collapse(_ptr(y), N, p, m2, _ptr(H), H.shape[2], _ptr(Z), Z.shape[2], _ptr(y_star),
                                             _ptr(H_star), C.addressof(la)))
    return y_star, H_star, la.value


def forecast(h, a1, P1, Z, T, R, Q, H):
    """Forecast recursion of R/predict.R:205-246 from KalmanC's a_fc / P_fc.  Returns y_fc (h x p), a_fc (h x m),
    P_fc (m x m x h), Fmat_fc (p x p x h)."""
    a1 = _f(np.asarray(a1, dtype=np.float64).reshape(-1))
    P1 = _f(P1)
    Z, T, R, Q, H = _cube(Z, "Z"), _cube(T, "T"), _cube(R, "R"), _cube(Q, "Q"), _cube(H, "H")
    p, m, r = Z.shape[0], a1.shape[0], Q.shape[0]
    y_fc, a_fc = np.zeros((h, p), order="F"), np.zeros((h, m), order="F")
    P_fc, F_fc = np.zeros((m, m, h), order="F"), np.zeros((p, p, h), order="F")
    native.check(native.lib().ssgpu_forecast(h, p, m, r, _ptr(a1), _ptr(P1), _ptr(Z), Z.shape[2], _ptr(T), T.shape[2],
                                             _ptr(R), R.shape[2], _ptr(Q), Q.shape[2], _ptr(H), H.shape[2], _ptr(y_fc),
                                             _ptr(a_fc), _ptr(P_fc), _ptr(F_fc)))
    return y_fc, a_fc, P_fc, F_fc


def _sim_components(comps, roots, p, hold):
    """ssgpu_sim_component array of the SimulateC calls of R/SimSmoother.R:65-553 / R/predict.R:262-833."""
    roots = roots or [{} for _ in range(len(comps) + 1)]
    if roots[len(comps)]:
        raise ValueError("roots of the residual variance cannot be injected")
    arr = (native.SimComponent * len(comps))()
    m = r = 0
    for i, (c, rt) in enumerate(zip(comps, roots)):
        a = _f(np.asarray(c["a"], dtype=np.float64).reshape(-1))
        Z, T, R, Q = _cube(c["Z"], "Z"), _cube(c["T"], "T"), _cube(c["R"], "R"), _cube(c["Q"], "Q")
        P_star = _f(np.atleast_2d(c["P_star"]))
        if c["draw_initial"] and P_star.shape[0] != a.shape[0]:
            raise ValueError("draw_initial needs an m x m P_star")
        if Z.shape[0] != p or R.shape[1] != Q.shape[0] * c["repeat_Q"]:
            raise ValueError("component matrices do not match")
        Q_root = None if rt.get("Q_root") is None else _cube(rt["Q_root"], "Q_root")
        P_root = None if rt.get("P_root") is None else _f(rt["P_root"])
        arr[i] = native.SimComponent(a.shape[0], Q.shape[0], int(c["repeat_Q"]), int(bool(c["draw_initial"])), hold(a),
                                     hold(Z), hold(T), hold(R), hold(Q), hold(P_star), Z.shape[2], T.shape[2],
                                     R.shape[2], Q.shape[2], hold(Q_root) if Q_root is not None else None,
                                     hold(P_root) if P_root is not None else None)
        m += a.shape[0]
        r += Q.shape[0] * int(c["repeat_Q"])
    return arr, m, r


def SimulateModel(nsim, N, p, comps, H, draws, extract=(), roots=None):
    """The unconditional half of SimSmoother() as one call -- predict()'s simulations over the forecast period
    (R/predict.R:142-856; include/ssgpu.h section 6d).  Returns y, epsilon (N, p, nsim), a (N, m, nsim), eta (N, r, nsim)
    and the extracted components."""
    keep = []

    def hold(x):
        keep.append(x)
        return _ptr(x)

    arr, m, r = _sim_components(comps, roots, p, hold)
    H = _cube(H, "H")
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1)
    y = np.zeros((N, p, nsim), order="F")
    eps = np.zeros((N, p, nsim), order="F")
    a_out = np.zeros((N, m, nsim), order="F")
    eta = np.zeros((N, r, nsim), order="F")
    ne = len(extract)
    Zs = [_cube(Z, "Z_padded") for Z in extract]
    outs = [np.zeros((N, p, nsim), order="F") for _ in extract]
    zp = (C.c_void_p * max(ne, 1))(*[_ptr(Z) for Z in Zs])
    nz = (C.c_int * max(ne, 1))(*[Z.shape[2] for Z in Zs])
    op = (C.c_void_p * max(ne, 1))(*[_ptr(o) for o in outs])
    native.check(native.lib().ssgpu_SimulateModel(int(nsim), int(N), int(p), len(comps), arr, _ptr(H), H.shape[2],
                                                  _ptr(draws), draws.shape[0], _ptr(y), _ptr(eps), _ptr(a_out), _ptr(eta),
                                                  ne, zp, nz, op))
    return {"y": y, "epsilon": eps, "a": a_out, "eta": eta, "components": outs}


def SimSmoother(nsim, N, p, comps, H, full, initialisation_steps, a_hat, eps_hat, eta_hat, draws, extract=(), roots=None):
    """R/SimSmoother.R:34-769 as one call (include/ssgpu.h section 6d): the unconditional draws of every component,
    the simulated data, their smoothed states, the mean corrections and the extracted components without the per-draw
    arrays leaving the device.  Arguments as oracle/sim_smoother.py (comps: dicts a, Z, T, R, Q, P_star, repeat_Q,
    draw_initial; full: (a, P_inf, P_star, Z, T, R, Q) with the residual states in front; draws in the reference's
    consumption order).  Returns epsilon (N, p, nsim), a (N, m, nsim), eta (N, r, nsim), components [(N, p, nsim)]."""
    keep = []

    def hold(x):
        keep.append(x)
        return _ptr(x)

    arr, m, r = _sim_components(comps, roots, p, hold)
    fa, fPi, fPs, fZ, fT, fR, fQ = full
    fa = _f(np.asarray(fa, dtype=np.float64).reshape(-1))
    fPi, fPs = _f(fPi), _f(fPs)
    fZ, fT, fR, fQ = _cube(fZ, "Z"), _cube(fT, "T"), _cube(fR, "R"), _cube(fQ, "Q")
    fm = native.SimModel(fa.shape[0], fQ.shape[0], hold(fa), hold(fPi), hold(fPs), hold(fZ), hold(fT), hold(fR), hold(fQ),
                         fZ.shape[2], fT.shape[2], fR.shape[2], fQ.shape[2])
    H = _cube(H, "H")
    a_hat, eps_hat, eta_hat = _f(a_hat), _f(eps_hat), _f(eta_hat)
    if a_hat.shape != (N, m) or eps_hat.shape != (N, p) or eta_hat.shape != (N, r):
        raise ValueError("smoothed estimates must be N x m, N x p, N x r")
    draws = np.ascontiguousarray(draws, dtype=np.float64).reshape(-1)
    eps = np.zeros((N, p, nsim), order="F")
    a_out = np.zeros((N, m, nsim), order="F")
    eta = np.zeros((N, r, nsim), order="F")
    ne = len(extract)
    Zs = [_cube(Z, "Z_padded") for Z in extract]
    outs = [np.zeros((N, p, nsim), order="F") for _ in extract]
    zp = (C.c_void_p * max(ne, 1))(*[_ptr(Z) for Z in Zs])
    nz = (C.c_int * max(ne, 1))(*[Z.shape[2] for Z in Zs])
    op = (C.c_void_p * max(ne, 1))(*[_ptr(o) for o in outs])
    native.check(native.lib().ssgpu_SimSmoother(
        int(nsim), int(N), int(p), len(comps), arr, _ptr(H), H.shape[2], C.byref(fm), int(initialisation_steps),
        _ptr(a_hat), _ptr(eps_hat), _ptr(eta_hat), _ptr(draws), draws.shape[0], _ptr(eps), _ptr(a_out), _ptr(eta),
        ne, zp, nz, op))
    return {"epsilon": eps, "a": a_out, "eta": eta, "components": outs}


# ---- simulation smoother on device-resident draws (torch CUDA float64 tensors, draw-contiguous layout) -----
def _stream_handle(stream):
    import torch
    st = stream if stream is not None else torch.cuda.current_stream()
    # torch's default stream has handle 0, which the C ABI reads as "the library's own stream": name the legacy
    # default stream explicitly (cudaStreamLegacy == 0x1) so that the work is ordered with torch's
    return C.c_void_p(st.cuda_stream or 1)


def SimulateC_dev(nsim, repeat_Q, N, a, Z, T, R, Q, P_star, draw_initial, eta_only, draws, Q_root=None, P_root=None,
                  stream=None, want=("y", "a", "eta")):
    """SimulateC on the device: `draws` is a CUDA tensor in the reference's consumption order; returns CUDA
    tensors y (N, p, nsim), a (N, m, nsim), eta (N, r_tot, nsim) -- C-contiguous, i.e. draws fastest."""
    import torch
    a = _f(np.asarray(a, dtype=np.float64).reshape(-1))
    Z, T, R, Q = _cube(Z, "Z"), _cube(T, "T"), _cube(R, "R"), _cube(Q, "Q")
    P_star = _f(np.atleast_2d(P_star))
    p, m, r = Z.shape[0], a.shape[0], Q.shape[0]
    r_tot = r * repeat_Q
    Q_root = None if Q_root is None else _cube(Q_root, "Q_root")
    P_root = None if P_root is None else _f(P_root)
    dev = draws.device
    mk = lambda k: torch.empty(N, k, nsim, device=dev, dtype=torch.float64)
    y = mk(p) if ("y" in want and not eta_only) else None
    ao = mk(m) if ("a" in want and not eta_only) else None
    eta = mk(r_tot) if "eta" in want else None
    dp_ = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    native.check(native.lib().ssgpu_SimulateC_dev(
        nsim, repeat_Q, N, p, m, R.shape[1], r, P_star.shape[0], _ptr(a), _ptr(Z), Z.shape[2], _ptr(T), T.shape[2],
        _ptr(R), R.shape[2], _ptr(Q), Q.shape[2], _ptr(P_star), int(bool(draw_initial)), int(bool(eta_only)),
        dp_(draws), draws.numel(), _ptr(Q_root), _ptr(P_root), dp_(y), dp_(ao), dp_(eta), _stream_handle(stream)))
    return {"y": y, "a": ao, "eta": eta}


def FastSmootherC_dev(y, a, P_inf, P_star, Z, T, R, Q, initialisation_steps, stream=None):
    """FastSmootherC on the device: y is a CUDA tensor (N, p, nsim), draws fastest; returns a_smooth (N, m, nsim)
    and eta (N, r, nsim) in the same layout."""
    import torch
    a = _f(np.asarray(a, dtype=np.float64).reshape(-1))
    P_inf, P_star = _f(P_inf), _f(P_star)
    Z, T, R, Q = _cube(Z, "Z"), _cube(T, "T"), _cube(R, "R"), _cube(Q, "Q")
    N, p, nsim = y.shape
    m, r = a.shape[0], Q.shape[0]
    if not y.is_contiguous():
        raise ValueError("y must be contiguous (N, p, nsim)")
    a_s = torch.empty(N, m, nsim, device=y.device, dtype=torch.float64)
    eta = torch.empty(N, r, nsim, device=y.device, dtype=torch.float64)
    native.check(native.lib().ssgpu_FastSmootherC_dev(
        C.c_void_p(y.data_ptr()), N, p, nsim, m, r, _ptr(a), _ptr(P_inf), _ptr(P_star), _ptr(Z), Z.shape[2], _ptr(T),
        T.shape[2], _ptr(R), R.shape[2], _ptr(Q), Q.shape[2], int(initialisation_steps), C.c_void_p(a_s.data_ptr()),
        C.c_void_p(eta.data_ptr()), _stream_handle(stream)))
    return {"a_smooth": a_s, "eta": eta}


def ExtractComponentC_dev(a, Z, stream=None):
    """ExtractComponentC on the device: a is a CUDA tensor (N, m, nsim), draws fastest -> (N, p, nsim)."""
    import torch
    Z = _cube(Z, "Z")
    N, m, nsim = a.shape
    p = Z.shape[0]
    out = torch.empty(N, p, nsim, device=a.device, dtype=torch.float64)
    native.check(native.lib().ssgpu_ExtractComponentC_dev(C.c_void_p(a.data_ptr()), m, nsim, N, _ptr(Z), p, Z.shape[2],
                                                          C.c_void_p(out.data_ptr()), _stream_handle(stream)))
    return out


def last_sim_kernel_ms():
    """(simulate, fast smoother per-draw, extract) kernel times of the last launches, ms."""
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    native.check(native.lib().ssgpu_last_sim_kernel_ms(C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def draw_layout_dev(x, to_r, stream=None):
    """(N, K, nsim) draws-fastest CUDA tensor <-> R's column-major N x K x nsim array (returned as a flat tensor)."""
    import torch
    if to_r:
        N, K, nsim = x.shape
        out = torch.empty(N * K * nsim, device=x.device, dtype=torch.float64)
    else:
        raise ValueError("use to_r=True; the reverse direction needs the shape: call native.lib() directly")
    native.check(native.lib().ssgpu_draw_layout_dev(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), N, K, nsim,
                                                    1, _stream_handle(stream)))
    return out


# ---- batched loglikelihood ----------------------------------------------------------------------------
class BatchModel:
    """An ssgpu_model plus the arrays that keep its pointers alive.

    Shared matrices come from `sm` (a sysmat.SysMat); per-evaluation overrides are arrays whose leading
    dimensions are (V, B) or (B,): `P_star_diag` / `Q_diag` (..., m) / (..., r) hold diagonals only,
    `P_star` / `Q` / `P_inf` (..., k, k) dense symmetric blocks, `a` (..., m).  Arrays are NumPy (host
    entry point) or torch CUDA tensors (device entry point) -- never mixed."""

    def __init__(self, sm, N, B, V=1, a=None, P_inf=None, P_star=None, Q=None, P_star_diag=None, Q_diag=None,
                 device=None):
        self.keep = []
        self.device = device
        Z, T, R, Qs = (_cube(x, n) for x, n in ((sm.Z, "Z"), (sm.T, "T"), (sm.R, "R"), (sm.Q, "Q")))
        a0 = _f(np.asarray(sm.a, dtype=np.float64).reshape(-1))
        m, p, r = _dims(a0, _f(sm.P_inf), _f(sm.P_star), Z, T, R, Qs, N)
        self.N, self.p, self.m, self.r, self.B, self.V = N, p, m, r, B, V
        mo = native.Model()
        mo.N, mo.p, mo.m, mo.r = N, p, m, r
        mo.Z = self._shared(Z, p, m)
        mo.T = self._shared(T, m, m)
        mo.R = self._shared(R, m, r)
        mo.a = self._per_eval(a, m, 1, False, vec=True) if a is not None else self._shared(a0.reshape(m, 1, 1), m, 1)
        mo.P_inf = (self._per_eval(P_inf, m, m, False) if P_inf is not None
                    else self._shared(_f(sm.P_inf)[:, :, None], m, m))
        if P_star_diag is not None:
            mo.P_star = self._per_eval(P_star_diag, m, m, True)
        elif P_star is not None:
            mo.P_star = self._per_eval(P_star, m, m, False)
        else:
            mo.P_star = self._shared(_f(sm.P_star)[:, :, None], m, m)
        if Q_diag is not None:
            mo.Q = self._per_eval(Q_diag, r, r, True)
        elif Q is not None:
            mo.Q = self._per_eval(Q, r, r, False)
        else:
            mo.Q = self._shared(Qs, r, r)
        self.c = mo

    def _shared(self, cube, rows, cols):
        flat = np.asfortranarray(cube).reshape(-1, order="F")
        if self.device is None:
            self.keep.append(flat)
            ptr = flat.ctypes.data
        else:
            import torch
            t = torch.from_numpy(flat.copy()).to(self.device)
            self.keep.append(t)
            ptr = t.data_ptr()
        return native.Matrix(ptr, rows, cols, 0, cube.shape[2], 0, 0)

    def _per_eval(self, x, rows, cols, diag, vec=False):
        block = rows if diag else rows * cols
        is_torch = not isinstance(x, np.ndarray) and hasattr(x, "data_ptr")
        shape = tuple(x.shape)
        tail = 1 if (diag or vec) else 2
        lead = shape[:len(shape) - tail]
        if int(np.prod(shape[len(lead):])) != block:
            raise ValueError("per-evaluation block has the wrong size")
        if lead == (self.B,):
            sv = 0
        elif lead == (self.V, self.B):
            sv = self.B * block
        else:
            raise ValueError(f"per-evaluation arrays must lead with (B,) or (V, B); got {shape}")
        if is_torch:
            if self.device is None:
                raise ValueError("torch tensors need device=...")
            t = x.contiguous()
            if str(t.dtype) != "torch.float64" or not t.is_cuda:
                raise ValueError("device arrays must be float64 CUDA tensors")
            self.keep.append(t)
            ptr = t.data_ptr()
        else:
            arr = np.ascontiguousarray(x, dtype=np.float64)
            if self.device is None:
                self.keep.append(arr)
                ptr = arr.ctypes.data
            else:
                import torch
                t = torch.from_numpy(arr).to(self.device)
                self.keep.append(t)
                ptr = t.data_ptr()
        # dense per-evaluation blocks are symmetric (covariances), so C order == column-major
        return native.Matrix(ptr, rows, cols, 1 if diag else 0, 1, block, sv)


def _y_dims(y, p):
    shape = tuple(y.shape)
    if len(shape) == 2:
        Np, B = shape
    elif len(shape) == 3:
        Np, B = shape[0] * shape[1], shape[2]
        if shape[1] != p:
            raise ValueError("y must be (N, p, B)")
    else:
        raise ValueError("y must be (N*p, B) or (N, p, B), time-major and batch-contiguous")
    if Np % p:
        raise ValueError("y rows must be a multiple of p")
    return Np // p, B


def loglik_batch(y, sm, V=1, kernel="auto", **per_eval):
    """Batched LogLikC on host arrays.  y: (N, B) or (N, p, B) C-contiguous (time-major, batch-contiguous),
    NaN = missing.  Returns (B,) for V == 1 else (V, B): loglik / N per evaluation."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, B = _y_dims(y, np.asarray(sm.Z).shape[0])
    model = BatchModel(sm, N, B, V, **per_eval)
    out = np.zeros((V, B))
    native.check(native.lib().ssgpu_loglik_batch(y.ctypes.data, B, V, C.byref(model.c), _KERNELS[kernel],
                                                 out.ctypes.data))
    return out[0] if V == 1 else out


def loglik_batch_dev(y, model: BatchModel, out, kernel="auto", stream=None):
    """Batched LogLikC on device-resident data (torch CUDA float64 tensors); asynchronous on `stream`
    (a torch.cuda.Stream, default: the current one)."""
    import torch
    if not (y.is_cuda and y.dtype == torch.float64 and y.is_contiguous() and out.is_cuda
            and out.dtype == torch.float64 and out.is_contiguous()):
        raise ValueError("y and out must be contiguous float64 CUDA tensors")
    if out.numel() != model.B * model.V:
        raise ValueError("out must hold B * V values")
    s = stream if stream is not None else torch.cuda.current_stream(y.device)
    # torch's default stream has handle 0, which the C ABI reads as "the library's own stream": name the
    # legacy default stream explicitly (cudaStreamLegacy == 0x1) so that the launch is ordered after torch's work
    handle = s.cuda_stream or 1
    native.check(native.lib().ssgpu_loglik_batch_dev(y.data_ptr(), model.B, model.V, C.byref(model.c),
                                                     _KERNELS[kernel], out.data_ptr(), handle))
    return out


def kalman_batch(y, sm, want=("loglik", "a_smooth", "V"), **per_series):
    """Batched KalmanC on host arrays (include/ssgpu.h section 2b).  y: (N, B) or (N, p, B), NaN = missing; per-series
    overrides as for loglik_batch with leading dimension (B,).  Returns a dict: for every name in `want` an array
    (B, ...) whose trailing dimensions are the reference's (column-major per series), plus initialisation_steps (B,)."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, B = _y_dims(y, np.asarray(sm.Z).shape[0])
    model = BatchModel(sm, N, B, 1, **per_series)
    p, m, r = model.p, model.m, model.r
    out = native.KalmanOut()
    res, flat = {}, {}
    for name in want:
        shape = kalman_shapes(N, p, m, r, True)[name]
        flat[name] = np.zeros((B, int(np.prod(shape))))
        setattr(out, name, flat[name].ctypes.data_as(native.dp))
        res[name] = shape
    steps = np.zeros(B, dtype=np.int32)
    out.initialisation_steps = steps.ctypes.data_as(native.ip)
    native.check(native.lib().ssgpu_kalman_batch(y.ctypes.data, B, C.byref(model.c), C.byref(out)))
    result = {k: flat[k].reshape((B,) + res[k][::-1]).transpose((0,) + tuple(range(len(res[k]), 0, -1))) for k in want}
    result["initialisation_steps"] = steps
    return result


def kalman_batch_dev(y, model: BatchModel, want=("loglik", "a_smooth", "V"), stream=None):
    """Batched KalmanC on device-resident data: y a contiguous CUDA tensor (N*p, B); returns CUDA tensors (B, size)
    -- block b is series b in the reference's column-major layout -- plus initialisation_steps (B,) int32."""
    import torch
    N, p, m, r, B = model.N, model.p, model.m, model.r, model.B
    out = native.KalmanOut()
    res = {}
    for name in want:
        n = int(np.prod(kalman_shapes(N, p, m, r, True)[name]))
        res[name] = torch.empty(B, n, device=y.device, dtype=torch.float64)
        setattr(out, name, C.cast(C.c_void_p(res[name].data_ptr()), native.dp))
    steps = torch.zeros(B, device=y.device, dtype=torch.int32)
    out.initialisation_steps = C.cast(C.c_void_p(steps.data_ptr()), native.ip)
    native.check(native.lib().ssgpu_kalman_batch_dev(y.data_ptr(), B, C.byref(model.c), C.byref(out), _stream_handle(stream)))
    res["initialisation_steps"] = steps
    return res


# ---- theta-driven objective / gradient (include/ssgpu.h section 6c) ---------------------------------------
def _index_vec(x, n, name):
    v = np.ascontiguousarray(x, dtype=np.int32).reshape(-1)
    if v.shape[0] != n:
        raise ValueError(f"{name} must have {n} entries")
    return v


def loglik_grad_batch(y, sm, theta, q_index, p_star_index, h=1e-3, kernel="auto"):
    """Loglikelihood and central-difference gradient for B series from one launch of 1 + 2k evaluations per
    series; variances exp(2 theta[q_index[j]]) are placed on the diagonals of Q / P_star on the device
    (-1 keeps the entry of sm.Q / sm.P_star).  y (N, B) or (N, p, B); theta (B, k).  Returns (loglik (B,), grad (B, k))."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    N, B = _y_dims(y, np.asarray(sm.Z).shape[0])
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(B, -1)
    k = theta.shape[1]
    model = BatchModel(sm, N, B, 1)
    qi, pi = _index_vec(q_index, model.r, "q_index"), _index_vec(p_star_index, model.m, "p_star_index")
    ll, gr = np.zeros(B), np.zeros((B, k))
    native.check(native.lib().ssgpu_loglik_grad_batch(y.ctypes.data, B, C.byref(model.c), theta.ctypes.data, k,
                                                      qi.ctypes.data, pi.ctypes.data, float(h), _KERNELS[kernel],
                                                      ll.ctypes.data, gr.ctypes.data))
    return ll, gr


def loglik_grad_batch_dev(y, model: BatchModel, theta, q_index, p_star_index, h=1e-3, kernel="auto", stream=None):
    """The same on device-resident data: y, theta (B, k) are torch CUDA float64 tensors, `model` a BatchModel
    built with device=... (shared matrices only).  Returns CUDA tensors (loglik (B,), grad (B, k))."""
    import torch
    B, k = theta.shape
    qi, pi = _index_vec(q_index, model.r, "q_index"), _index_vec(p_star_index, model.m, "p_star_index")
    ll = torch.empty(B, device=y.device, dtype=torch.float64)
    gr = torch.empty(B, k, device=y.device, dtype=torch.float64)
    native.check(native.lib().ssgpu_loglik_grad_batch_dev(
        C.c_void_p(y.data_ptr()), B, C.byref(model.c), C.c_void_p(theta.data_ptr()), k, qi.ctypes.data, pi.ctypes.data,
        float(h), _KERNELS[kernel], C.c_void_p(ll.data_ptr()), C.c_void_p(gr.data_ptr()), _stream_handle(stream)))
    return ll, gr


def loglik_theta_batch_dev(y, model: BatchModel, theta, q_index, p_star_index, kernel="auto", stream=None):
    """V arbitrary parameter vectors per series: theta (V, B, k) CUDA tensor -> (V, B) loglik / N."""
    import torch
    V, B, k = theta.shape
    qi, pi = _index_vec(q_index, model.r, "q_index"), _index_vec(p_star_index, model.m, "p_star_index")
    out = torch.empty(V, B, device=y.device, dtype=torch.float64)
    native.check(native.lib().ssgpu_loglik_theta_batch_dev(
        C.c_void_p(y.data_ptr()), B, V, C.byref(model.c), C.c_void_p(theta.contiguous().data_ptr()), k,
        qi.ctypes.data, pi.ctypes.data, _KERNELS[kernel], C.c_void_p(out.data_ptr()), _stream_handle(stream)))
    return out


def loglik_cov_batch_dev(y, model: BatchModel, theta, blocks, h=1e-3, grad=True, kernel="auto", stream=None):
    """Objective (and central-difference gradient) from parameter vectors with full covariance blocks: every block
    (target, offset, n, theta_offset) -- target "Q" or "P_star" -- is L D L' built on the device from
    theta[theta_offset : theta_offset + n + n (n - 1) / 2] (R/Cholesky.R:148-167; include/ssgpu.h section 6d).
    grad=True: theta (B, k) -> (loglik (B,), gradient (B, k)); grad=False: theta (V, B, k) -> (V, B).
    NaN where the reference's Cholesky() returns NA (a variance <= 1e-7)."""
    import torch
    arr = (native.CovBlock * len(blocks))()
    for i, (target, offset, n, t0) in enumerate(blocks):
        arr[i] = native.CovBlock({"Q": 0, "P_star": 1}[target], int(offset), int(n), int(t0))
    theta = theta.contiguous()
    if grad:
        B, k = theta.shape
        ll = torch.empty(B, device=y.device, dtype=torch.float64)
        gr = torch.empty(B, k, device=y.device, dtype=torch.float64)
        native.check(native.lib().ssgpu_loglik_cov_batch_dev(
            C.c_void_p(y.data_ptr()), B, 0, C.byref(model.c), C.c_void_p(theta.data_ptr()), k, arr, len(blocks),
            C.c_double(h), 1, _KERNELS[kernel], C.c_void_p(ll.data_ptr()), C.c_void_p(gr.data_ptr()), _stream_handle(stream)))
        return ll, gr
    V, B, k = theta.shape
    out = torch.empty(V, B, device=y.device, dtype=torch.float64)
    native.check(native.lib().ssgpu_loglik_cov_batch_dev(
        C.c_void_p(y.data_ptr()), B, V, C.byref(model.c), C.c_void_p(theta.data_ptr()), k, arr, len(blocks),
        C.c_double(0.0), 0, _KERNELS[kernel], C.c_void_p(out.data_ptr()), None, _stream_handle(stream)))
    return out
