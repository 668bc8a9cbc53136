"""End-to-end probe of ssgpu_loglik_grad_batch (host pointers, pinned y): wall time per call and its parts."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import bench
from statespacer_b200 import api, native
dev = torch.device("cuda", 0)
B = 125000
sm, y, theta = bench.make_shard(torch, dev, B, 0)
y_h = torch.empty(y.shape, dtype=torch.float64, pin_memory=True); y_h.copy_(y)
th_h = theta.cpu().pin_memory()
ll_h = torch.empty(B, dtype=torch.float64, pin_memory=True); gr_h = torch.empty(B, 4, dtype=torch.float64, pin_memory=True)
hm = api.BatchModel(sm, bench.T_STEPS, B, 1)
qi, pi = np.asarray(bench.Q_INDEX, dtype=np.int32), np.asarray(bench.P_INDEX, dtype=np.int32)
def call():
    native.check(native.lib().ssgpu_loglik_grad_batch(y_h.data_ptr(), B, C.byref(hm.c), th_h.data_ptr(), 4, qi.ctypes.data, pi.ctypes.data, 1e-3, 0, ll_h.data_ptr(), gr_h.data_ptr()))
for i in range(2): call()
torch.cuda.synchronize()
for i in range(3):
    t0 = time.perf_counter(); call(); print("call %.1f ms" % (1e3 * (time.perf_counter() - t0)))
yd = torch.empty_like(y)
torch.cuda.synchronize(); t0 = time.perf_counter(); yd.copy_(y_h, non_blocking=True); torch.cuda.synchronize(); print("plain H2D of y %.1f ms" % (1e3 * (time.perf_counter() - t0)))
# the same call on pageable host arrays (what R passes): staged through the library's ring of pinned buffers
y_p = y.cpu().numpy().copy(); th_p = theta.cpu().numpy().copy(); ll_p = np.empty(B); gr_p = np.empty((B, 4))
def call_p():
    native.check(native.lib().ssgpu_loglik_grad_batch(y_p.ctypes.data, B, C.byref(hm.c), th_p.ctypes.data, 4, qi.ctypes.data, pi.ctypes.data, 1e-3, 0, ll_p.ctypes.data, gr_p.ctypes.data))
for i in range(2): call_p()
for i in range(3):
    t0 = time.perf_counter(); call_p(); print("pageable call %.1f ms" % (1e3 * (time.perf_counter() - t0)))
print("pageable == pinned:", np.array_equal(ll_p, ll_h.numpy()))
