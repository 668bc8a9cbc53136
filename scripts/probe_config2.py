"""Development probe: batched loglik throughput at the config-2 shape (p = 2, N = 192, m = 34, time-varying Z,
5 % missing) and config-3 shape (DNS p = 8, m = 14) -- which kernel serves them and how fast."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import random_model, simulate_y, load_fed, DNS_PARAM
from statespacer_b200 import api, sysmat

def run(name, sm, y, kernel="auto", reps=3):
    dev = torch.device("cuda")
    N, p, B = y.shape
    model = api.BatchModel(sm, N, B, 1, device=dev)
    yd = torch.from_numpy(np.ascontiguousarray(y)).to(dev)
    out = torch.empty(1, B, device=dev, dtype=torch.float64)
    api.loglik_batch_dev(yd, model, out, kernel=kernel)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        api.loglik_batch_dev(yd, model, out, kernel=kernel)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    m = sm.m
    falg = 4 * m ** 3 + (4 * p + 3) * m ** 2 + 7 * p * m
    print(f"{name:28s} kernel={api.last_kernel():45s} B={B} {B * N / dt:.3e} series*timesteps/s  ({B * N * falg / dt / 1e12:.2f} TFLOP/s algorithmic)")

rng = np.random.default_rng(2)
N, p, m = 192, 2, 34
sm = random_model(rng, m, p, d=4, tv=("Z",), N=N)
y1 = simulate_y(rng, sm, N, missing=0.05)
B = 4096
y = np.repeat(y1[:, :, None], B, axis=2) + 0.01 * rng.normal(size=(N, p, B))
run("config2 shape (tv Z, m=34)", sm, y)
sm2 = random_model(rng, m, p, d=4, N=N)
run("m=34 time-invariant", sm2, y)
fed = load_fed()
smd, _ = sysmat.dns_yield_curve(DNS_PARAM)
B = 20000
yd = np.repeat(fed[:, :, None], B, axis=2) + 0.01 * rng.normal(size=fed.shape + (B,))
run("config3 DNS p=8 m=14 (mma)", smd, yd)
run("config3 DNS (cols)", smd, yd, kernel="cols")
run("config3 DNS (generic)", smd, yd[:, :, :4096], kernel="generic")
run("config2 shape (wsm forced)", sm, y, kernel="wsm")
run("config2 shape (generic)", sm, y, kernel="generic")
from conftest import config2_standin
c2 = config2_standin()
B = 32768
y2 = np.repeat(c2.y[:, :, None], B, axis=2) * (1 + 0.01 * rng.normal(size=(1, 1, B)))
run("config2 stand-in BSM+regressors", c2.sm, y2)
run("config2 stand-in (generic)", c2.sm, y2[:, :, :2048], kernel="generic")
