"""Development probe: wall time of every call of one device-resident simulation-smoother step (bench.py --workload sim)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from statespacer_b200 import api
dev = torch.device("cuda", 0)
air, sm = bench.sim_model()
N, S = len(air), int(sys.argv[1]) if len(sys.argv) > 1 else 12500
steps0 = int(api.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"])
mc = sm.m - 1
Zc, Tc, Rc, Qc = sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :], sm.Q[1:, 1:, :]
Qe, one = sm.Q[:1, :1, :], np.zeros((1, 1, 1))
d1 = torch.randn(N * S, device=dev, dtype=torch.float64)
d2 = torch.randn(N * S, device=dev, dtype=torch.float64)
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    acc.setdefault(name, []).append(time.perf_counter() - t0)
for it in range(12):
    t0 = time.perf_counter(); s1 = api.SimulateC_dev(S, 1, N, np.zeros(mc), Zc, Tc, Rc, Qc, np.zeros((mc, mc)), False, False, d1); tick("simulate", t0)
    t0 = time.perf_counter(); s2 = api.SimulateC_dev(S, 1, N, np.zeros(1), one, one, one, Qe, np.zeros((1, 1)), False, True, d2, want=("eta",)); tick("residual", t0)
    t0 = time.perf_counter(); y = s1["y"] + s2["eta"]; tick("add", t0)
    t0 = time.perf_counter(); fs = api.FastSmootherC_dev(y, *sm.args(), steps0); tick("fast_smoother", t0)
    t0 = time.perf_counter(); comp = api.ExtractComponentC_dev(fs["a_smooth"], sm.Z); tick("extract", t0)
    t0 = time.perf_counter(); del s1, s2, y, fs, comp; tick("free", t0)
for k, v in acc.items():
    print(f"{k:14s} median {1e3 * sorted(v[2:])[len(v[2:]) // 2]:7.3f} ms   max {1e3 * max(v[2:]):7.3f} ms")
print("kernel ms (simulate, fs, extract):", api.last_sim_kernel_ms())
