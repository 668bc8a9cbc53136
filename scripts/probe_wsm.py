"""Development probe: one launch of loglik_wsm_kernel at the config-2 stand-in (for ncu)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from conftest import config2_standin
from statespacer_b200 import api
c2 = config2_standin()
rng = np.random.default_rng(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
y2 = np.repeat(c2.y[:, :, None], B, axis=2) * (1 + 0.01 * rng.normal(size=(1, 1, B)))
dev = torch.device("cuda")
model = api.BatchModel(c2.sm, 192, B, 1, device=dev)
yd = torch.from_numpy(np.ascontiguousarray(y2)).to(dev)
out = torch.empty(1, B, device=dev, dtype=torch.float64)
for _ in range(2):
    api.loglik_batch_dev(yd, model, out, kernel="wsm")
torch.cuda.synchronize()
print(api.last_kernel(), float(out[0, 0]))
