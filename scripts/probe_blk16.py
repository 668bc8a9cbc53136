"""Development probe: univariate structural models with 17 <= m <= 32 (two seasonal patterns, hourly seasonal) --
block-row kernel with 16 lanes per evaluation against the warp / shared-memory kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from statespacer_b200 import api, sysmat as S
dev = torch.device("cuda", 0)
N, B = 400, 40000
models = {"slope+seasonal12+seasonal7 (m=20)": S._combine(1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(12, 0.05), S.comp_bsm_trig(7, 0.03)]),
          "slope+seasonal24 (m=26)": S._combine(1, [[1.0]], [S.comp_level_slope(0.1, 0.01), S.comp_bsm_trig(24, 0.05)])}
for name, sm in models.items():
    y = torch.randn(N, B, device=dev, dtype=torch.float64).cumsum(0) * 0.3
    model = api.BatchModel(sm, N, B, 1, device=dev)
    out = torch.empty(B, device=dev, dtype=torch.float64)
    res = {}
    for kernel in ("blk", "wsm" if sm.m > 23 else "mma"):
        api.loglik_batch_dev(y, model, out, kernel=kernel); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): api.loglik_batch_dev(y, model, out, kernel=kernel)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        res[kernel] = out.clone()
        print(f"{name:36s} {api.last_kernel():58s} {B * N / dt:.3e} series*timesteps/s")
    a, b = list(res.values())
    print("   max rel diff", float(((a - b).abs() / b.abs().clamp(min=1)).max()))
