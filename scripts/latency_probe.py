"""Development probe: single-call latency of the five legacy entry points (what statespacer() / StateSpaceEval()
see per call) next to the reference's own C++ on one host core."""
import sys, os, time, math
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import kat_cases, config2_standin
from statespacer_b200 import api
from oracle import ref

def timeit(f, n):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e6

cases = kat_cases() + [config2_standin()]
print(f"{'case':10s} {'N':>4s} {'p':>2s} {'m':>3s} | LogLikC gpu / ref (us) | KalmanC gpu / ref (us)")
for c in cases:
    a = c.args()
    g1 = timeit(lambda: api.LogLikC(*a), 50)
    r1 = timeit(lambda: ref.LogLikC(*a), 20) if ref.available() else float("nan")
    g2 = timeit(lambda: api.KalmanC(*a, False), 20)
    r2 = timeit(lambda: ref.KalmanC(*a, False), 5) if ref.available() else float("nan")
    print(f"{c.name:10s} {c.y.shape[0]:4d} {c.y.shape[1]:2d} {c.sm.m:3d} | {g1:9.0f} / {r1:9.0f}        | {g2:9.0f} / {r2:9.0f}")

# the library calls alone (arguments converted once, ctypes call in the loop): what a compiled caller such as the .Call
# shim pays per call, without numpy conversions on either side
import ctypes as C
from statespacer_b200 import native
from oracle import ref as _ref
print("\nraw C-ABI calls (us per call, arguments prepared once):")
for c in cases[:4]:
    y, isna, a, P_inf, P_star, Z, T, R, Q = c.args()
    y = np.asfortranarray(y, dtype=np.float64); isna = np.asfortranarray(np.asarray(isna).astype(np.int32))
    a = np.asfortranarray(np.asarray(a, dtype=np.float64).reshape(-1))
    P_inf, P_star = np.asfortranarray(P_inf), np.asfortranarray(P_star)
    cube = lambda x: np.asfortranarray(x if np.ndim(x) == 3 else np.asarray(x)[:, :, None])
    Z, T, R, Q = cube(Z), cube(T), cube(R), cube(Q)
    N, p = y.shape; m, r = a.shape[0], R.shape[1]
    out = C.c_double()
    vp = lambda x: C.c_void_p(x.ctypes.data)
    args = (vp(y), vp(isna), N, p, m, r, vp(a), vp(P_inf), vp(P_star), vp(Z), Z.shape[2], vp(T), T.shape[2], vp(R), R.shape[2],
            vp(Q), Q.shape[2], C.byref(out))
    L = native.lib()
    g = timeit(lambda: L.ssgpu_LogLikC(*args), 200)
    gv = out.value
    rl = _ref.lib()
    rargs = (y.ctypes.data_as(C.POINTER(C.c_double)), isna.ctypes.data_as(C.POINTER(C.c_int)), N, p, m, r) + tuple(
        x.ctypes.data_as(C.POINTER(C.c_double)) for x in (a, P_inf, P_star)) + (
        Z.ctypes.data_as(C.POINTER(C.c_double)), Z.shape[2], T.ctypes.data_as(C.POINTER(C.c_double)), T.shape[2],
        R.ctypes.data_as(C.POINTER(C.c_double)), R.shape[2], Q.ctypes.data_as(C.POINTER(C.c_double)), Q.shape[2], C.byref(out))
    rr = timeit(lambda: rl.ssref_loglik(*rargs), 100)
    print(f"{c.name:10s} N={N:4d} m={m:3d}  ssgpu_LogLikC {g:8.1f}   reference LogLikC {rr:8.1f}   (values {gv:.9f} / {out.value:.9f})")
