"""Small cases through every kernel touched in round 2, for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`
(thread-per-evaluation kernel at 3 and 7 blocks incl. the hand-over, lane kernel with the observation ring and with the
time-varying-Z ring, LDL' covariance blocks, sharded objective)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
from conftest import simulate_y  # noqa: E402
from statespacer_b200 import api, sysmat  # noqa: E402

rng = np.random.default_rng(0)
B, N = 70, 37
th = 0.5 * np.log([1.0, 0.1, 0.01, 0.05]) + 0.2 * rng.normal(size=(3, B, 4))
var = np.exp(2 * th)
Qd = np.concatenate([var[..., :3], np.repeat(var[..., 3:4], 11, axis=-1)], axis=-1)
Pd = np.zeros((3, B, 14)); Pd[..., 0] = var[..., 0]
y = np.cumsum(rng.normal(size=(N, B)), axis=0)
y[rng.uniform(size=y.shape) < 0.1] = np.nan
a = api.loglik_batch(y, sysmat.bsm12(th[0, 0]), V=3, Q_diag=Qd, P_star_diag=Pd, kernel="blk")
b = api.loglik_batch(y, sysmat.bsm12(th[0, 0]), V=3, Q_diag=Qd, P_star_diag=Pd, kernel="blk_lanes")
print("bsm12", api.last_kernel()[:40], float(np.nanmax(np.abs(a - b))))
sm5 = sysmat._combine(1, [[0.5]], [sysmat.comp_local_level(0.2), sysmat.comp_bsm_trig(4, 0.05)])
y5 = np.stack([simulate_y(rng, sm5, N, missing=0.1)[:, 0] for _ in range(B)], axis=1)
print("m=5", api.loglik_batch(y5, sm5, kernel="blk")[:2], api.last_kernel()[:30])
ll = sysmat.nile_local_level(1.3, 0.4)
yl = np.cumsum(rng.normal(size=(N, 64)), axis=0)
print("level", api.loglik_batch(yl, ll, kernel="blk")[:2], api.last_kernel()[:30])
os.environ["SSGPU_ALLOW_SHARED_DEVICE"] = "1"
api.use_devices(2)
ll2, gr2 = api.loglik_grad_batch(y[:, :64].copy(), sysmat.bsm12(th[0, 0]), th[0, :64].copy(), [0, 1, 2] + [3] * 11, [0] + [-1] * 13)
api.use_devices(1)
print("sharded", ll2[:2])
dev = torch.device("cuda")
p, npar = 2, 3
Z = np.concatenate([np.eye(p), np.eye(p)], axis=1); T = sysmat.block_diag(np.zeros((p, p)), np.eye(p))
H = sysmat.cholesky_cov(np.array([0.1, -0.2, 0.3]), 2)
tpl = sysmat.SysMat(Z=Z[:, :, None], T=T[:, :, None], R=np.eye(4)[:, :, None], Q=sysmat.block_diag(H, H)[:, :, None], a=np.zeros(4),
                    P_inf=sysmat.block_diag(np.zeros((2, 2)), np.eye(2)), P_star=sysmat.block_diag(H, np.zeros((2, 2))))
yb = np.stack([simulate_y(rng, tpl, N) for _ in range(9)], axis=2)
model = api.BatchModel(tpl, N, 9, 1, device=dev)
thc = torch.from_numpy(0.2 * rng.normal(size=(9, 6))).to(dev)
print("cov", api.loglik_cov_batch_dev(torch.from_numpy(yb).to(dev), model, thc, [("Q", 0, 2, 0), ("Q", 2, 2, 3), ("P_star", 0, 2, 0)])[0][:2])
torch.cuda.synchronize()
print("done")
