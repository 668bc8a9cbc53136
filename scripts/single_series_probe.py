"""Development probe: latency of single-series calls on structural models (what statespacer()'s optim loop pays per objective)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from statespacer_b200 import api, sysmat
from oracle import ref
rng = np.random.default_rng(0)
th = 0.5 * np.log([1.0, 0.1, 0.01, 0.05])
cases = [("local level N=100", sysmat.nile_local_level(1.3, 0.4), 100), ("local level N=1000", sysmat.nile_local_level(1.3, 0.4), 1000),
         ("bsm12 N=144", sysmat.bsm12(th), 144), ("bsm12 N=1000", sysmat.bsm12(th), 1000)]
for name, sm, N in cases:
    y = np.cumsum(rng.normal(size=(N, 1)), axis=0)
    na = np.isnan(y)
    f = lambda: api.LogLikC(y, na, *sm.args())
    g = lambda: ref.LogLikC(y, na, *sm.args())
    f(); g()
    t0 = time.perf_counter()
    for _ in range(100): f()
    tg = (time.perf_counter() - t0) / 100 * 1e6
    t0 = time.perf_counter()
    for _ in range(20): g()
    tr = (time.perf_counter() - t0) / 20 * 1e6
    # the gradient the optimiser needs: 1 + 2k evaluations in one call
    k = 4 if sm.m == 14 else 2
    print(f"{name:20s} LogLikC {tg:8.1f} us (reference {tr:8.1f} us)   kernel {api.last_kernel()[:44]}")
