"""Element-wise errors of ssgpu_KalmanC against the compiled reference next to the disagreement of the two CPU checkers
(tests/test_gpu_parity.py entry_scale), per output array and test case: where does 1e-9 hold, where is it conditioning."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import Case, config2_standin, kat_cases, load_series, random_model, simulate_y  # noqa: E402
from oracle import kalman_np, ref  # noqa: E402
from statespacer_b200 import api, sysmat  # noqa: E402
import test_gpu_parity as T  # noqa: E402


def report(c, diag=False):
    alt = kalman_np.KalmanC(*c.args(), diag)
    want = ref.KalmanC(*c.args(), diag)
    got = api.KalmanC(*c.args(), diag)
    keys = T.KALMAN_KEYS + (("v_norm", "e", "D", "Tstat_observation") if diag else ())
    for k in keys:
        try:
            T.kalman_compare(c.name, got, want, alt, (k,), c.y.shape[1])
        except AssertionError as e:
            print("FAIL", str(e).splitlines()[0][:300])
        try:
            T.kalman_compare(c.name, got, want, None, (k,), c.y.shape[1])
        except AssertionError as e:
            print("  needs the second checker:", str(e).splitlines()[0][:200])


for c in kat_cases():
    report(c)
for i, shape in enumerate(T.SHAPES[:6]):
    rng = np.random.default_rng(7)
    sm = random_model(rng, shape["m"], shape["p"], d=shape["d"], tv=shape["tv"], N=30)
    report(Case(f"shape{i}", simulate_y(rng, sm, 30, missing=0.1), sm))
air = np.log(load_series("airpassengers.txt")).copy()
air[[5, 6, 50, 51, 52, 100]] = np.nan
report(Case("air-na", air, sysmat.airline([0.5 * math.log(0.00135), 0.3, 0.4])))
report(config2_standin())
