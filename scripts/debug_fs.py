import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import random_model, simulate_y
from statespacer_b200 import api
from oracle import cport
for (m, p, d, nsim) in [(2, 1, 1, 41), (2, 1, 1, 64), (2, 1, 0, 41), (3, 1, 1, 41), (6, 2, 2, 41)]:
    rng = np.random.default_rng(300 + m)
    N = 30
    sm = random_model(rng, m, p, d=d, N=N)
    y0 = simulate_y(rng, sm, N)
    steps = api.KalmanC(y0, np.isnan(y0), *sm.args(), False)["initialisation_steps"]
    y = np.stack([simulate_y(rng, sm, N) for _ in range(nsim)], axis=2)
    g = api.FastSmootherC(y, *sm.args(), steps, True)
    c = cport.FastSmootherC(y, *sm.args(), steps, True)
    bad = [s for s in range(nsim) if not np.array_equal(g["a_smooth"][:, :, s], c["a_smooth"][:, :, s])]
    tb = [t for t in range(N) if not np.array_equal(g["a_smooth"][t], c["a_smooth"][t])]
    print(m, p, d, nsim, "steps", steps, "bad draws", len(bad), bad[:5], "bad t", tb[:6], "eta ok", np.array_equal(g["eta"], c["eta"]))
