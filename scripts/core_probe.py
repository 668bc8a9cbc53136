"""Config 3 (dynamic Nelson-Siegel, p = 8, m = 14) on the core kernel: device-resident rate for the library SSGPU_LIB
names.  usage: [SSGPU_LIB=...] python scripts/core_probe.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import DNS_PARAM, load_fed  # noqa: E402
from statespacer_b200 import api, sysmat  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
dev = torch.device("cuda:0")
api.set_device(0)
sm, _ = sysmat.dns_yield_curve(DNS_PARAM)
y1 = load_fed()
N, p = y1.shape
g = torch.Generator(device=dev)
g.manual_seed(11)
yd = (torch.from_numpy(np.ascontiguousarray(y1)).to(dev)[:, :, None] *
      (1.0 + 0.01 * torch.randn(1, 1, B, generator=g, device=dev, dtype=torch.float64))).contiguous()
model = api.BatchModel(sm, N, B, 1, device=dev)
out = torch.empty(1, B, device=dev, dtype=torch.float64)
stream = torch.cuda.current_stream(dev)
for _ in range(3):
    api.loglik_batch_dev(yd, model, out, stream=stream)
torch.cuda.synchronize(dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10):
    api.loglik_batch_dev(yd, model, out, stream=stream)
e1.record(stream)
torch.cuda.synchronize(dev)
ms = e0.elapsed_time(e1) / 10
print(os.environ.get("SSGPU_LIB", "default"), api.last_kernel(), f"{ms:.3f} ms", f"{B * N / ms * 1e3:.4g} series*timesteps/s",
      float(out[0, 0]))
