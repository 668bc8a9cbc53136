"""Development probe: where the host time of one device-resident simulation-smoother step goes."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from statespacer_b200 import api
air, sm = bench.sim_model()
N, S = len(air), 12500
steps0 = int(api.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"])
mc = sm.m - 1
Zc, Tc, Rc, Qc = sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :], sm.Q[1:, 1:, :]
Qe = sm.Q[:1, :1, :]; one = np.zeros((1, 1, 1))
dev = torch.device("cuda")
d1 = torch.randn(N * S, device=dev, dtype=torch.float64); d2 = torch.randn(N * S, device=dev, dtype=torch.float64)
def t(f, name, acc):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); acc[name] = acc.get(name, 0) + time.perf_counter() - t0; return r
for rep in range(4):
    acc = {}
    s1 = t(lambda: api.SimulateC_dev(S, 1, N, np.zeros(mc), Zc, Tc, Rc, Qc, np.zeros((mc, mc)), False, False, d1), "simulate", acc)
    s2 = t(lambda: api.SimulateC_dev(S, 1, N, np.zeros(1), one, one, one, Qe, np.zeros((1, 1)), False, True, d2, want=("eta",)), "simulate_eps", acc)
    y = t(lambda: s1["y"] + s2["eta"], "add", acc)
    fs = t(lambda: api.FastSmootherC_dev(y, *sm.args(), steps0), "smoother", acc)
    comp = t(lambda: api.ExtractComponentC_dev(fs["a_smooth"], sm.Z), "extract", acc)
    print({k: round(v * 1e3, 3) for k, v in acc.items()}, "kernel ms", [round(x, 3) for x in api.last_sim_kernel_ms()])
