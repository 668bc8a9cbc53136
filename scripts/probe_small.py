"""Development probe: small structural models (m <= 9) -- block-row kernel (1 to 8 lanes per evaluation) against the
tensor and column kernels."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from statespacer_b200 import api, sysmat as S
dev = torch.device("cuda", 0)
N, B = 500, int(os.environ.get("PROBE_B", 200000))
models = {"level (m=2)": S._combine(1, [[0.7]], [S.comp_local_level(0.3)]),
          "level+slope (m=3)": S._combine(1, [[0.7]], [S.comp_level_slope(0.3, 0.02)]),
          "level+seasonal4 (m=5)": S._combine(1, [[0.5]], [S.comp_local_level(0.2), S.comp_bsm_trig(4, 0.05)]),
          "slope+seasonal7 (m=9)": S._combine(1, [[0.5]], [S.comp_level_slope(0.2, 0.01), S.comp_bsm_trig(7, 0.05)])}
for name, sm in models.items():
    y = torch.randn(N, B, device=dev, dtype=torch.float64).cumsum(0) * 0.3
    model = api.BatchModel(sm, N, B, 1, device=dev)
    out = torch.empty(B, device=dev, dtype=torch.float64)
    ref = None
    for kernel in ("blk", "core", "mma", "cols"):
        api.loglik_batch_dev(y, model, out, kernel=kernel); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3): api.loglik_batch_dev(y, model, out, kernel=kernel)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        d = 0.0 if ref is None else float(((out - ref).abs() / ref.abs().clamp(min=1)).max())
        ref = out.clone() if ref is None else ref
        print(f"{name:24s} {api.last_kernel()[:52]:52s} {B * N / dt:.3e} series*timesteps/s  (max rel diff {d:.1e})")
