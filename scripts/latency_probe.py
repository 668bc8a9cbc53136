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
