"""Config-4-shaped probe of the simulation smoother through the host-pointer API (development aid):
airline SARIMA on log AirPassengers, N = 144, m = 28, nsim draws."""
import math
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_series  # noqa: E402
from statespacer_b200 import api, sysmat  # noqa: E402

nsim = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
air = np.log(load_series("airpassengers.txt"))[:, None]
sm = sysmat.airline([0.5 * math.log(0.00135), 0.4, 0.55])
N = len(air)
steps = api.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"]
rng = np.random.default_rng(4)
cm = sm.m - 1
draws = rng.normal(size=N * nsim)
for rep in range(2):
    t0 = time.perf_counter()
    s = api.SimulateC(nsim, 1, N, np.zeros(cm), sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :], sm.Q[1:, 1:, :],
                      np.zeros((cm, cm)), False, False, True, draws)
    t1 = time.perf_counter()
    fs = api.FastSmootherC(s["y"], *sm.args(), steps, True)
    t2 = time.perf_counter()
    comp = api.ExtractComponentC(fs["a_t"], sm.Z)
    t3 = time.perf_counter()
    print(f"nsim={nsim} simulate {t1 - t0:.3f}s fast_smoother {t2 - t1:.3f}s extract {t3 - t2:.3f}s (host API, incl. copies)")
