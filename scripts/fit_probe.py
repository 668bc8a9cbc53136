"""Development probe: maximum-likelihood fit of every series of one config-5 shard (statespacer_b200/fit.py)."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from statespacer_b200 import api, fit
B = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
dev = torch.device("cuda")
sm, y, theta_true = bench.make_shard(torch, dev, B, 0)
model = api.BatchModel(sm, bench.T_STEPS, B, 1, device=dev)
start = torch.tensor(bench.MU, device=dev).expand(B, 4).contiguous() + 0.3
torch.cuda.synchronize(); t0 = time.perf_counter()
res = fit.fit_batch(y, model, start, bench.Q_INDEX, bench.P_INDEX, max_iter=60)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
ll_true, _ = api.loglik_grad_batch_dev(y, model, theta_true, bench.Q_INDEX, bench.P_INDEX)
err = (res.theta - theta_true).abs().median(0).values.cpu().numpy()
print(f"B={B} T={bench.T_STEPS}: {res.iterations} iterations, {res.evaluations} filter runs per series, {dt:.1f} s "
      f"= {B * bench.T_STEPS * res.evaluations / dt:.3e} series*timesteps/s; converged {float(res.converged.double().mean()):.3f}; "
      f"loglik >= loglik(truth) for {float((res.loglik >= ll_true - 1e-9).double().mean()):.3f}; median |theta - truth| {err}")
