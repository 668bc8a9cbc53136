"""Development probe: univariate structural model with two explanatory variables (time-varying Z, m = 16) --
block-row kernel against the warp / shared-memory kernel that served the shape before."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from statespacer_b200 import api, sysmat
rng = np.random.default_rng(1)
N, B = 500, 60000
base = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]))
m0 = base.m
X = rng.normal(size=(N, 2))
m = m0 + 2
Z = np.zeros((1, m, N)); Z[0, :m0, :] = base.Z[0, :, 0][:, None]; Z[0, m0:, :] = X.T
sm = sysmat.SysMat(Z=Z, T=sysmat.block_diag(base.T[:, :, 0], np.eye(2))[:, :, None],
                   R=np.concatenate([np.eye(m0), np.zeros((2, m0))], axis=0)[:, :, None], Q=base.Q, a=np.zeros(m),
                   P_inf=sysmat.block_diag(base.P_inf, np.eye(2)), P_star=sysmat.block_diag(base.P_star, np.zeros((2, 2))))
dev = torch.device("cuda", 0)
y = torch.randn(N, B, device=dev, dtype=torch.float64).cumsum(0) * 0.3
model = api.BatchModel(sm, N, B, 1, device=dev)
out = torch.empty(B, device=dev, dtype=torch.float64)
res = {}
for kernel in ("blk", "wsm", "generic"):
    api.loglik_batch_dev(y, model, out, kernel=kernel); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): api.loglik_batch_dev(y, model, out, kernel=kernel)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    res[kernel] = out.clone()
    print(f"{kernel:8s} {api.last_kernel():60s} {B * N / dt:.3e} series*timesteps/s")
print("max rel diff blk vs wsm", float(((res["blk"] - res["wsm"]).abs() / res["wsm"].abs().clamp(min=1)).max()))
