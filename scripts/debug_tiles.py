"""Development probe: tensor kernel (tile-sparse / dense order) vs the oracle on block-diagonal models."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import simulate_y
from test_gpu_parity import _block_diagonal_model
from statespacer_b200 import api, sysmat
from oracle import cport
for m, p, d, miss in [(17, 1, 3, 0.1), (17, 1, 0, 0.0), (17, 1, 3, 0.0), (17, 1, 0, 0.1), (20, 1, 3, 0.1), (14, 1, 3, 0.1), (17, 2, 3, 0.1)]:
    rng = np.random.default_rng(700 + m)
    B, N = 4, 40
    sm = _block_diagonal_model(rng, m, p, d=d, N=N)
    y = np.stack([simulate_y(rng, sm, N, missing=miss) for _ in range(B)], axis=2)
    got = api.loglik_batch(y, sm, kernel="mma"); k1 = api.last_kernel()
    dense = api.loglik_batch(y, sm, kernel="mma_dense")
    gen = api.loglik_batch(y, sm, kernel="generic")
    want = np.array([cport.LogLikC(y[:, :, b], np.isnan(y[:, :, b]), *sm.args()) for b in range(B)])
    print(m, p, d, miss, k1, "sparse", np.max(np.abs(got - want)), "dense", np.max(np.abs(dense - want)), "generic", np.max(np.abs(gen - want)))
