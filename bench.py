#!/usr/bin/env python
"""Headline benchmark: batched Kalman loglikelihood + finite-difference gradient (BASELINE.json config 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload loglik|sim]

Workload per GPU (weak scaling): 125,000 series (1M / 8) of length T = 1000 from the
local level + slope + trigonometric seasonal (s = 12) model, m = r = 14 incl. the residual state, 13
exactly-diffuse states; per-series parameters theta_b in R^4; one step = the loglikelihood at theta_b and
at the 2k = 8 central-difference points (h = 1e-3) for every series = 9 filter runs per series =
1.125e9 series*timesteps per GPU, plus (N > 1) the NCCL gather of per-series loglik + gradient.

metric   series*timesteps / s, FP64, whole job, inputs resident in HBM (CUDA events on the launch stream); the
         step is the public call api.loglik_grad_batch_dev (ssgpu_loglik_grad_batch_dev): theta in, loglik +
         gradient out, the 9 variance sets per series built on the device
e2e      the same through the host-pointer C ABI (ssgpu_loglik_grad_batch): pinned host y / theta are copied in
         and loglik + gradient copied out inside the timed region
roofline dominant kernel (loglik_blk_kernel for this model) against the FP64 peak measured in this run by
         ssgpu_fp64_peak (MEASURED_PEAKS.json carries no FP64 figure): achieved / frac count the FP64 operations the
         structure-exploiting kernel issues (the kernel reports flop_per_step in its name); the dense algorithm's
         count of SURVEY.md 8d is reported beside it (dense_algorithm_*)
cpu_baseline  the reference's own LogLikC.cpp (oracle/_ref, compiled unmodified) on all host cores over a
         bounded sample of the same workload (value_1core: one thread, port_value: the allocation-free C
         restatement on all cores); --impl reference runs only the all-core leg and prints its own line.

--workload sim runs BASELINE.json config 4 instead (simulation smoother draws, see the section below).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MU = 0.5 * np.log(np.array([1.0, 0.1, 0.01, 0.05]))
T_STEPS = 1000
M = 14
K_PAR = 4
V = 1 + 2 * K_PAR
H_FD = 1e-3
B_PER_GPU = 125_000
# config 5: which parameter each diagonal entry of Q / P_star is the variance exp(2 theta) of (-1: fixed)
Q_INDEX = [0, 1, 2] + [3] * 11
P_INDEX = [0] + [-1] * 13
METRIC = "kalman_loglik_series_timesteps_per_s"
UNIT = "series*timesteps/s"


def f_alg(m, p):
    """Algorithmic flops per series*timestep, dense generic algorithm."""
    return 4 * m ** 3 + (4 * p + 3) * m ** 2 + 7 * p * m


# Scalar FP64 instructions per warp and time step in the p = 1 hot loop of loglik_mma_kernel (cuobjdump -sass and
# ncu smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on, profiles/README.md): 23 DFMA, 11 DADD, 4 DMUL
SCALAR_FLOPS_PER_STEP = (23 * 2 + 11 + 4) * 32

def executed_flops(kernel_name):
    """Flops the tensor kernel issues per series*timestep: DMMAs (8x8x4: 512 flops each, padding included) + scalar
    FP64.  None for the other kernels."""
    import re
    m = re.search(r"dmma_per_step=(\d+)", kernel_name or "")
    if m:
        return int(m.group(1)) * 512 + SCALAR_FLOPS_PER_STEP
    m = re.search(r"flop_per_step=(\d+)", kernel_name or "")     # block kernels: counted by the library itself
    return int(m.group(1)) if m else None


def executed_fp64_instructions(kernel_name):
    """FP64 instructions (DFMA + DMUL + DADD, one issue slot each) per series*timestep of the thread-per-evaluation kernel:
    the library reports the count of its instantiation in the kernel name (loglik_thr.cu thr_flops_per_step -- the counts
    of the SASS of its loop, profiles/sass_thr_r2_hotloop.txt); None for the other kernels."""
    m = re.search(r"fp64_instr_per_step=(\d+)", kernel_name or "")
    return int(m.group(1)) if m else None


def variants(theta):
    """(V, B, 4): theta, theta + h e_i, theta - h e_i  (stats::optim central differences, ndeps = 1e-3)."""
    out = [theta]
    for i in range(K_PAR):
        for sgn in (1.0, -1.0):
            t = theta.copy()
            t[:, i] += sgn * H_FD
            out.append(t)
    return np.stack(out)


def diag_params(theta_v):
    """Q and P_star diagonals (V, B, 14) of the config-5 model from theta (R/Cholesky.R:158: var = exp(2 theta))."""
    var = np.exp(2 * theta_v)
    Qd = np.concatenate([var[..., :3], np.repeat(var[..., 3:4], 11, axis=-1)], axis=-1)
    Pd = np.zeros_like(Qd)
    Pd[..., 0] = var[..., 0]
    return np.ascontiguousarray(Qd), np.ascontiguousarray(Pd)


def config(n_gpus, B, extra=None):
    c = {"workload": "config5: local level + slope + trig seasonal(12), m=r=14, d=13 diffuse, T=1000, "
                     "loglik + central-difference gradient (9 parameter vectors per series)",
         "series_per_gpu": B, "series_total": B * n_gpus, "timesteps": T_STEPS, "param_vectors": V,
         "l2": "inputs larger than L2 (y = %.2f GB per GPU), no flush" % (B * T_STEPS * 8 / 1e9),
         "parallelism": f"series sharded over {n_gpus} GPU(s), no collective in the recursion; one all-gather of "
                        "per-series [loglik, gradient] per step when n_gpus > 1"}
    if extra:
        c.update(extra)
    return c


# ---- the reference / CPU arm ----------------------------------------------------------------------------
def cpu_sample(n_series, seed=0):
    """A bounded sample of the workload in the CPU checkers' layout."""
    from statespacer_b200 import sysmat
    rng = np.random.default_rng(20240501 + seed)
    theta = MU + 0.25 * rng.normal(size=(n_series, 4))
    sm = sysmat.bsm12(MU)
    Tm, Zm = sm.T[:, :, 0], sm.Z[0, :, 0]
    var = np.exp(2 * theta)
    sd = np.sqrt(np.concatenate([var[:, :3], np.repeat(var[:, 3:4], 11, axis=1)], axis=1))
    a = np.zeros((n_series, M))
    a[:, 1] = rng.normal(size=n_series)
    a[:, 0] = rng.normal(size=n_series) * sd[:, 0]
    y = np.zeros((n_series, T_STEPS))
    for t in range(T_STEPS):
        y[:, t] = a @ Zm
        a = a @ Tm.T + rng.normal(size=(n_series, M)) * sd
    tv = variants(theta)                                  # (V, n, 4)
    Qd, Pd = diag_params(tv)
    yy = np.tile(y, (V, 1))                               # evaluation e = v * n + b
    Qfull = np.zeros((V * n_series, M, M))
    Pfull = np.zeros((V * n_series, M, M))
    idx = np.arange(M)
    Qfull[:, idx, idx] = Qd.reshape(-1, M)
    Pfull[:, idx, idx] = Pd.reshape(-1, M)
    return sm, yy, Pfull, Qfull


try:
    ORIG_AFFINITY = os.sched_getaffinity(0)     # before any NUMA binding of the GPU legs
except AttributeError:
    ORIG_AFFINITY = None


def host_cores():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the CPU arm is told the
    real count explicitly (omp_set_num_threads inside the checkers).  The GPU legs bind the process to the NUMA node of
    its GPU (bind_to_gpu_numa_node); the CPU legs call this first and get every core back."""
    if ORIG_AFFINITY is not None:
        try:
            os.sched_setaffinity(0, ORIG_AFFINITY)
        except OSError:
            pass
        return len(ORIG_AFFINITY)
    return os.cpu_count() or 1


def time_cpu(mod, sm, yy, Pfull, Qfull, threads=0):
    t0 = time.perf_counter()
    out = mod.loglik_batch(yy, sm.a, sm.P_inf, Pfull, sm.Z[:, :, 0], sm.T[:, :, 0], sm.R[:, :, 0], Qfull, threads)
    dt = time.perf_counter() - t0
    return yy.shape[0] * T_STEPS / dt, dt, out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport, ref
    mod, kind = (ref, "reference") if ref.available() else (cport, "port")
    cores = host_cores()
    # BASELINE.md 3.4 asks for >= 8192 series per sample; the whole `--steps K --warmup W` run has to end within a few
    # minutes, so the sample is sized from a probe: 8192 series when K + W steps of it fit ~150 s, fewer otherwise
    smp, yyp, Pp, Qp = cpu_sample(64 * max(1, cores // 4))
    rate, _, _ = time_cpu(mod, smp, yyp, Pp, Qp, cores)
    n_series = args.cpu_series or int(min(8192, max(32 * cores, rate * 150.0 / ((args.steps + min(args.warmup, 1)) * V * T_STEPS))))
    sm, yy, Pfull, Qfull = cpu_sample(n_series)
    for _ in range(min(args.warmup, 1)):
        time_cpu(mod, sm, yy, Pfull, Qfull, cores)
    times = []
    for _ in range(args.steps):
        _, dt, _ = time_cpu(mod, sm, yy, Pfull, Qfull, cores)
        times.append(dt)
    total = sum(times)
    value = args.steps * yy.shape[0] * T_STEPS / total
    sample = f"{n_series} series x {V} parameter vectors x {T_STEPS} timesteps per step"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config(args.gpus, B_PER_GPU, {"cpu_sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                             "what": "src/LogLikC.cpp compiled unmodified (oracle/_ref), OpenMP over series"
                             if kind == "reference" else "oracle/kalman_c.c"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- the GPU arm ----------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: an NVML polling thread in this process
    (a resident `nvidia-smi -lms` stalled the driver for tens of milliseconds per query, visible in a 15 ms timed
    region), `nvidia-smi` only as the fallback when pynvml is missing."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=100):
        self.index = index
        self.interval_ms = interval_ms
        self.proc = None
        self.thread = None
        self.rows = []
        self.stop_flag = False

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _poll(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        smax = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((float(sm), float(smax), pw, [k for k, b in bits.items() if rs & b]))
            except Exception:
                pass
            time.sleep(self.interval_ms / 1000.0)

    def start(self):
        if os.environ.get("SSGPU_BENCH_NO_SAMPLER"):      # diagnostic runs only: no clocks line
            return
        try:
            import threading

            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", str(self.interval_ms), "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            sm = [r[0] for r in self.rows]
            reasons = sorted({k for r in self.rows for k in r[3]})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(r[1] for r in self.rows) if sm else None,
                    "power_w_max": max(r[2] for r in self.rows) if sm else None, "samples": len(sm), "reasons": reasons,
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi"}


def make_shard(torch, dev, B, seed):
    """Synthetic shard generated on the device: y (T, B) time-major, theta (B, 4)."""
    from statespacer_b200 import sysmat
    g = torch.Generator(device=dev)
    g.manual_seed(20240501 + seed)
    theta = torch.tensor(MU, device=dev) + 0.25 * torch.randn(B, 4, generator=g, device=dev, dtype=torch.float64)
    sm = sysmat.bsm12(MU)
    Tm = torch.tensor(sm.T[:, :, 0], device=dev)
    Zm = torch.tensor(sm.Z[0, :, 0], device=dev)
    var = torch.exp(2 * theta)
    sd = torch.sqrt(torch.cat([var[:, :3], var[:, 3:4].expand(B, 11)], dim=1))
    g.manual_seed(77 + seed)
    a = torch.zeros(B, M, device=dev, dtype=torch.float64)
    a[:, 1] = torch.randn(B, generator=g, device=dev, dtype=torch.float64)
    a[:, 0] = torch.randn(B, generator=g, device=dev, dtype=torch.float64) * sd[:, 0]
    y = torch.empty(T_STEPS, B, device=dev, dtype=torch.float64)
    for t in range(T_STEPS):
        y[t] = a @ Zm
        a = a @ Tm.T + torch.randn(B, M, generator=g, device=dev, dtype=torch.float64) * sd
    return sm, y, theta


def gpu_setup():
    """One process per GPU (RANK / LOCAL_RANK / WORLD_SIZE from torchrun); the NCCL group is created once per process."""
    import torch
    import torch.distributed as dist
    from statespacer_b200 import api
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (statespacer_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    bind_to_gpu_numa_node(torch, local)
    api.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run on (and therefore first-touch pinned staging memory from) the CPUs of the NUMA node the
    GPU's PCIe root hangs off -- with 8 ranks each uploading 1 GB per step, copies that cross the socket interconnect are
    what the end-to-end figure loses.  Best effort: silently skipped where sysfs does not tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:          # nvml pads the domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        cpus = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        allowed = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            allowed.update(range(int(lo), int(hi or lo) + 1))
        allowed &= set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
    except Exception:
        pass


def committed_traffic(kernel_prefix, grid):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel from the `ncu --set full` summaries
    committed under profiles/ (scripts/ncu_summary.py), matched by kernel name and grid size; None when no capture of
    this kernel at this size is committed.  DRAM bytes cannot be counted in-run without a profiler."""
    import csv
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_*_summary.csv"))):
        try:
            rows = list(csv.reader(open(path)))
        except OSError:
            continue
        cur = None
        for r in rows + [["kernel", "", ""]]:
            if r and r[0] == "kernel":
                if cur and kernel_prefix in cur.get("name", "") and cur.get("grid") == grid and "rd" in cur and "wr" in cur:
                    best = {"bytes": cur["rd"] + cur["wr"], "source": os.path.relpath(path, ROOT)}
                cur = {"name": r[2] if len(r) > 2 else ""}
            elif cur is not None and len(r) >= 3:
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[1])
                if r[0] == "dram__bytes_read.sum" and scale:
                    cur["rd"] = float(r[2]) * scale
                elif r[0] == "dram__bytes_write.sum" and scale:
                    cur["wr"] = float(r[2]) * scale
                elif r[0] == "launch__grid_size":
                    cur["grid"] = int(float(r[2]))
    return best


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from statespacer_b200 import api

    world, rank, local, dev = gpu_setup()
    B = args.series
    if args.scaling == "strong":           # a fixed job (--series-total, default 1M series = config 5) cut into world shards
        B = args.series_total // world
    sm, y, theta = make_shard(torch, dev, B, rank)
    if args.missing > 0:
        g = torch.Generator(device=dev)
        g.manual_seed(5 + rank)
        y[torch.rand(y.shape, generator=g, device=dev) < args.missing] = float("nan")
    # the public call of the path: theta (B, 4) in, loglik (B,) + central-difference gradient (B, 4) out; the
    # 9 variance sets per series are built on the device (ssgpu_loglik_grad_batch_dev)
    model = api.BatchModel(sm, T_STEPS, B, 1, device=dev)
    from statespacer_b200 import shard
    stream = torch.cuda.current_stream(dev)
    res = {}

    def gather():
        # per-series [loglik, gradient] formed locally, then the one collective of the path (NCCL all-gather)
        local = torch.cat([res["ll"][:, None], res["gr"]], dim=1)
        return shard.gather_series(local, B * world)

    def kernel_call():
        res["ll"], res["gr"] = api.loglik_grad_batch_dev(y, model, theta, Q_INDEX, P_INDEX, h=H_FD,
                                                         kernel=args.kernel, stream=stream)

    def step():
        kernel_call()
        if world > 1:
            gather()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync()
    # FP64 roofline denominator, measured on this device in this run
    peak_tf, _ = api.fp64_peak(0, 4000, 5)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = api.launch_count()
    kern_ms = []
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(stream)
        kernel_call()
        k1.record(stream)
        kern_ms.append((k0, k1))
        if world > 1:
            gather()
    e1.record(stream)
    sync()
    launches = api.launch_count() - l0
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    kernel_ms = statistics.mean(a.elapsed_time(b) for a, b in kern_ms)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    units_per_step = B * V * T_STEPS * world
    value = units_per_step * args.steps / (elapsed_ms * 1e-3)
    finite = bool(torch.isfinite(res["ll"]).all().item() and torch.isfinite(res["gr"]).all().item())
    kernel_name = api.last_kernel()

    # ---- end to end through the host-pointer C ABI (pinned host buffers) ----
    e2e_steps = max(1, min(args.steps, 3))
    y_h = torch.empty(y.shape, dtype=torch.float64, pin_memory=True)
    y_h.copy_(y)
    th_h = theta.cpu().pin_memory()
    ll_h = torch.empty(B, dtype=torch.float64, pin_memory=True)
    gr_h = torch.empty(B, K_PAR, dtype=torch.float64, pin_memory=True)
    import ctypes as C
    from statespacer_b200 import native
    hmodel = api.BatchModel(sm, T_STEPS, B, 1)
    qi, pi = np.asarray(Q_INDEX, dtype=np.int32), np.asarray(P_INDEX, dtype=np.int32)

    def e2e_step():
        native.check(native.lib().ssgpu_loglik_grad_batch(y_h.data_ptr(), B, C.byref(hmodel.c), th_h.data_ptr(), K_PAR,
                                                          qi.ctypes.data, pi.ctypes.data, H_FD,
                                                          api._KERNELS[args.kernel], ll_h.data_ptr(), gr_h.data_ptr()))

    e2e_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = units_per_step * e2e_steps / e2e_s
    e2e_match = bool(np.array_equal(ll_h.numpy(), res["ll"].cpu().numpy())
                     and np.array_equal(gr_h.numpy(), res["gr"].cpu().numpy()))
    h2d = y_h.numel() * 8 + th_h.numel() * 8 + 8 * (M * M * 3 + M * 3)
    d2h = (ll_h.numel() + gr_h.numel()) * 8
    # the same call on PAGEABLE host arrays (what R hands over): the library stages them through its pinned ring
    y_p, th_p = y_h.numpy().copy(), th_h.numpy().copy()
    ll_p, gr_p = np.empty(B), np.empty((B, K_PAR))

    def e2e_pageable_step():
        native.check(native.lib().ssgpu_loglik_grad_batch(y_p.ctypes.data, B, C.byref(hmodel.c), th_p.ctypes.data, K_PAR,
                                                          qi.ctypes.data, pi.ctypes.data, H_FD,
                                                          api._KERNELS[args.kernel], ll_p.ctypes.data, gr_p.ctypes.data))

    e2e_pageable_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pageable_step()
    sync()
    e2e_p_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_p_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_p_s = float(t.item())
    e2e_pageable_value = units_per_step * e2e_steps / e2e_p_s
    e2e_pageable_match = bool(np.array_equal(ll_p, ll_h.numpy()) and np.array_equal(gr_p, gr_h.numpy()))

    # ONE process driving several GPUs through the C ABI (ssgpu_use_devices): what an R session can do.  Weak scaling:
    # args.single_process devices x B series, host arrays pinned, results written straight into the caller's arrays.
    sp = None
    if world == 1 and args.single_process > 1:
        nd = api.use_devices(args.single_process)
        Bt = B * nd
        y_all = torch.empty(T_STEPS, Bt, dtype=torch.float64, pin_memory=True)
        th_all = torch.empty(Bt, K_PAR, dtype=torch.float64, pin_memory=True)
        for d in range(nd):                    # the same shard on every device: results must repeat bit for bit
            y_all[:, d * B:(d + 1) * B].copy_(y_h)
            th_all[d * B:(d + 1) * B].copy_(th_h)
        ll_all = torch.empty(Bt, dtype=torch.float64, pin_memory=True)
        gr_all = torch.empty(Bt, K_PAR, dtype=torch.float64, pin_memory=True)
        big = api.BatchModel(sm, T_STEPS, Bt, 1)

        def sp_step():
            native.check(native.lib().ssgpu_loglik_grad_batch(y_all.data_ptr(), Bt, C.byref(big.c), th_all.data_ptr(), K_PAR,
                                                              qi.ctypes.data, pi.ctypes.data, H_FD,
                                                              api._KERNELS[args.kernel], ll_all.data_ptr(), gr_all.data_ptr()))

        sp_step()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sp_step()
        sp_s = time.perf_counter() - t0
        same = all(bool(torch.equal(ll_all[d * B:(d + 1) * B], ll_h)) and bool(torch.equal(gr_all[d * B:(d + 1) * B], gr_h))
                   for d in range(nd))
        sp = {"n_devices": nd, "value": Bt * V * T_STEPS * e2e_steps / sp_s, "unit": UNIT, "steps": e2e_steps,
              "series_total": Bt, "every_device_bit_identical_to_single_device": same,
              "h2d_bytes_per_step": y_all.numel() * 8 + th_all.numel() * 8, "d2h_bytes_per_step": (Bt + Bt * K_PAR) * 8,
              "note": "ssgpu_use_devices + ssgpu_loglik_grad_batch from one process: one worker thread, stream and staging "
                      "ring per device, no collective"}
        api.use_devices(1)
        del y_all, th_all

    if rank == 0:
        flops = f_alg(M, 1) * B * V * T_STEPS
        achieved = flops / (kernel_ms * 1e-3) / 1e12
        f_exec = executed_flops(kernel_name)
        executed = f_exec * B * V * T_STEPS / (kernel_ms * 1e-3) / 1e12 if f_exec else None
        hbm_peak = None
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        thr = kernel_name.startswith("loglik_thr_kernel")
        # every lane of the thread kernel works on its own evaluation; the lane kernel at 7 blocks idles 1 lane in 8
        useful = 1.0 if thr else 7.0 / 8.0 if kernel_name.startswith("loglik_blk_kernel<L=8,NB=7") else 1.0
        n_inst = executed_fp64_instructions(kernel_name)
        # share of the FP64 pipe's issue slots the kernel fills: one FP64 instruction of a warp occupies a
        # sub-partition's pipe for 2 cycles, whether it is a DFMA (2 flop) or a DMUL / DADD (1 flop)
        pipe_frac = None
        if n_inst:
            props = torch.cuda.get_device_properties(dev)
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            pipe_frac = n_inst * 2.0 * B * V * T_STEPS / 32.0 / (props.multi_processor_count * 4 * sm_hz * kernel_ms * 1e-3)
        # state record of blk_handoff.cuh: 135 doubles + 5 ints per evaluation at 7 blocks (written by the lane kernel, read here)
        handoff = 1.0 * (135 * 8 + 5 * 4) * B * V if thr else 0.0
        grid = (B * V + 127) // 128 if thr else None
        traffic = committed_traffic("loglik_thr_kernel" if thr else kernel_name.split("<")[0], grid) if grid else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config(world, B, {"kernel": kernel_name, "missing_fraction": args.missing,
                                            "results_finite": finite}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "bit_identical_to_resident_path": e2e_match, "host_memory": "pinned",
                        "pageable": {"value": e2e_pageable_value, "unit": UNIT, "steps": e2e_steps,
                                     "bit_identical_to_pinned": e2e_pageable_match,
                                     "note": "the same call on pageable host arrays (what R passes): staged through "
                                             "the library's ring of pinned buffers"},
                        "single_process_multi_gpu": sp},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "kernel": kernel_name,
                             "achieved": executed if executed else achieved, "peak": peak_tf, "unit": "TFLOP/s",
                             "frac": (executed if executed else achieved) / peak_tf,
                             "traffic": traffic["bytes"] if traffic else None,
                             "traffic_source": traffic["source"] if traffic else
                             "no ncu capture of this kernel at this size under profiles/",
                             "traffic_algorithmic": 8.0 * B * T_STEPS + 8.0 * B * V + (handoff if thr else 8.0 * B * V * 2 * M),
                             "traffic_algorithmic_note": "of the dominant kernel's launch: y once + results + "
                             + ("the state record the lane kernel leaves behind after the diffuse phase, read once (%d B per "
                                "evaluation; the lane kernel's own launch writes it and reads the variance sets: %.2f GB)"
                                % (handoff / (B * V), (handoff + 8.0 * B * V * 2 * M + 8.0 * B * 13) / 1e9) if thr
                                else "the variance sets"),
                             "useful_frac": (executed if executed else achieved) / peak_tf * useful,
                             "fp64_pipe_frac": pipe_frac,
                             "kernel_ms": kernel_ms,
                             "flops_executed_per_series_timestep": f_exec,
                             "dense_algorithm_flops_per_series_timestep": f_alg(M, 1),
                             "dense_algorithm_equivalent_tflops": achieved,
                             "dense_algorithm_equivalent_frac": achieved / peak_tf,
                             "note": "achieved/frac count the FP64 operations the kernel ISSUES (fma = 2, mul = add = 1; every "
                                     "lane): the kernels exploit the block-diagonal T and the symmetry of P (SURVEY.md "
                                     "8d: a structure-exploiting kernel reports its own executed-flop count).  "
                                     "fp64_pipe_frac is the share of FP64 issue slots filled (30 % of the kernel's FP64 "
                                     "instructions are DMUL / DADD: one flop, same slot); useful_frac discounts lanes "
                                     "that only pad.  dense_algorithm_*: the same run priced at the dense "
                                     "algorithm's 4m^3 + 7m^2 + 7m flops per series*timestep, for comparison with a "
                                     "dense kernel only -- it exceeds the peak when the structure is exploited",
                             "peak_source": "ssgpu_fp64_peak DFMA chains, measured in this run (MEASURED_PEAKS.json "
                                            "has no FP64 figure)",
                             "hbm_algorithmic_gbs": 8.0 * B * T_STEPS / (kernel_ms * 1e-3) / 1e9,
                             "hbm_peak_gbs": hbm_peak}}
        if world == 1 and not args.no_cpu:
            from oracle import cport, ref
            mod, kind = (ref, "reference") if ref.available() else (cport, "port")
            cores = host_cores()
            n_series = max(8192, 256 * cores) if not args.cpu_series else args.cpu_series   # >= 8192 series
            smc, yy, Pfull, Qfull = cpu_sample(n_series)
            v_ref, dt, o_ref = time_cpu(mod, smc, yy, Pfull, Qfull, cores)
            v_port, _, o_port = time_cpu(cport, smc, yy, Pfull, Qfull, cores)
            n1 = 64 * V                                      # the 1-core rate next to the all-core one
            v_one, _, _ = time_cpu(mod, smc, yy[:n1], Pfull[:n1], Qfull[:n1], 1)
            line["cpu_baseline"] = {"value": v_ref, "unit": UNIT, "cores": cores, "kind": kind, "value_1core": v_one,
                                    "sample": f"{n_series} series x {V} parameter vectors x {T_STEPS} timesteps "
                                              f"({dt:.1f} s)",
                                    "port_value": v_port,
                                    "port_vs_reference_max_abs_diff": float(np.max(np.abs(o_ref - o_port)))}
        return line
    return None


# ---- config 4: simulation smoother draws (--workload sim) ---------------------------------------------
SIM_METRIC = "sim_smoother_draw_timesteps_per_s"
SIM_UNIT = "draws*timesteps/s"
NSIM_PER_GPU = 12_500          # 100,000 draws over 8 GPUs (BASELINE.json config 4)


def sim_model():
    """SARIMA(1,1,1)(0,1,1)_12 on log AirPassengers: N = 144, m = 28 incl. the residual state."""
    import math
    from statespacer_b200 import sysmat
    air = np.log(np.loadtxt(os.path.join(ROOT, "tests", "golden", "airpassengers.txt"), comments="#"))[:, None]
    sm = sysmat.airline([0.5 * math.log(0.00135), 0.3, -0.4, -0.55], s=(12, 1), ar=(0, 1), i=(1, 1), ma=(1, 1))
    return air, sm


def sim_bytes(N, p, m, r, mc, rc):
    """Mandatory HBM bytes per draw*timestep of the three kernels, draw-contiguous layout, no a_t copy."""
    simulate = 8 * (rc + p + mc + rc) + 8 * (p + p)        # component + residual (eta_only) call
    smoother = 8 * (p + m + r)
    extract = 8 * (m + p)
    return simulate, smoother, extract


def sim_config(n_gpus, nsim, extra=None):
    c = {"workload": "config4: SimSmoother draws, SARIMA(1,1,1)(0,1,1)12 on log AirPassengers, N=144, m=28, d=13 "
                     "diffuse: SimulateC (component + residual) -> FastSmootherC -> ExtractComponentC per draw",
         "nsim_per_gpu": nsim, "nsim_total": nsim * n_gpus, "timesteps": 144,
         "l2": "per-draw arrays larger than L2 (%.2f GB of outputs per GPU), no flush" % (nsim * 144 * 8 * 90 / 1e9),
         "parallelism": f"draws sharded over {n_gpus} GPU(s), no collective"}
    if extra:
        c.update(extra)
    return c


def run_sim_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cport, ref
    mod, kind = (ref, "reference") if ref.available() else (cport, "port")
    cores = host_cores()
    air, sm = sim_model()
    N = len(air)
    from oracle import kalman_np
    steps0 = int(kalman_np.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"])
    per_thread = args.cpu_draws
    mc = sm.m - 1
    Zc, Tc, Rc, Qc = sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :], sm.Q[1:, 1:, :]
    Qe = sm.Q[:1, :1, :]
    one = np.zeros((1, 1, 1))

    def chunk(seed):
        rng = np.random.default_rng(seed)
        d1, d2 = rng.normal(size=N * per_thread), rng.normal(size=N * per_thread)
        s1 = mod.SimulateC(per_thread, 1, N, np.zeros(mc), Zc, Tc, Rc, Qc, np.zeros((mc, mc)), False, False, True, d1)
        s2 = mod.SimulateC(per_thread, 1, N, np.zeros(1), one, one, one, Qe, np.zeros((1, 1)), False, True, False, d2)
        y = s1["y"] + s2["eta"]
        fs = mod.FastSmootherC(y, *sm.args(), steps0, True)
        return mod.ExtractComponentC(fs["a_t"], sm.Z)

    def step():
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            list(ex.map(chunk, range(cores)))
        return time.perf_counter() - t0

    for _ in range(min(args.warmup, 1)):
        step()
    times = [step() for _ in range(args.steps)]
    value = cores * per_thread * N * args.steps / sum(times)
    sample = f"{cores} threads x {per_thread} draws x {N} timesteps per step (one chunk per thread)"
    print(json.dumps({"impl": "reference", "metric": SIM_METRIC, "value": value, "unit": SIM_UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                      "data": "synthetic", "config": sim_config(args.gpus, args.nsim, {"cpu_sample": sample}),
                      "cpu_baseline": {"value": value, "unit": SIM_UNIT, "cores": cores, "kind": kind, "sample": sample},
                      "e2e": {"value": value, "unit": SIM_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


def run_sim_gpu(args):
    import torch
    import torch.distributed as dist
    from statespacer_b200 import api

    world, rank, local, dev = gpu_setup()
    air, sm = sim_model()
    N, S = len(air), args.nsim
    steps0 = int(api.KalmanC(air, np.isnan(air), *sm.args(), False)["initialisation_steps"])
    mc = sm.m - 1
    Zc, Tc, Rc, Qc = sm.Z[:, 1:, :], sm.T[1:, 1:, :], sm.R[1:, 1:, :], sm.Q[1:, 1:, :]
    Qe = sm.Q[:1, :1, :]
    one = np.zeros((1, 1, 1))
    g = torch.Generator(device=dev)
    g.manual_seed(4 + rank)
    d1 = torch.randn(N * S, generator=g, device=dev, dtype=torch.float64)
    d2 = torch.randn(N * S, generator=g, device=dev, dtype=torch.float64)
    stream = torch.cuda.current_stream(dev)
    kt = {"simulate": [], "smoother": [], "extract": []}

    def step(record=False):
        s1 = api.SimulateC_dev(S, 1, N, np.zeros(mc), Zc, Tc, Rc, Qc, np.zeros((mc, mc)), False, False, d1)
        if record:
            kt["simulate"].append(api.last_sim_kernel_ms()[0])
        s2 = api.SimulateC_dev(S, 1, N, np.zeros(1), one, one, one, Qe, np.zeros((1, 1)), False, True, d2,
                               want=("eta",))
        y = s1["y"] + s2["eta"]
        fs = api.FastSmootherC_dev(y, *sm.args(), steps0)
        comp = api.ExtractComponentC_dev(fs["a_smooth"], sm.Z)
        if record:
            ms = api.last_sim_kernel_ms()
            kt["smoother"].append(ms[1])
            kt["extract"].append(ms[2])
        return s1, fs, comp

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    out = None
    for _ in range(args.warmup + 1):
        # the same ownership pattern as the timed loop (the previous step's results are alive while the next step runs):
        # otherwise torch's caching allocator meets a new peak inside the timed region and stalls it with cudaMalloc calls
        out = step()
    sync()
    # a step is a few ms of kernels between short host-synchronous calls: clocks are polled through NVML in-process
    sampler = ClockSampler(local, interval_ms=20)
    sampler.start()
    time.sleep(0.1)
    l0 = api.launch_count()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        out = step(record=True)
    e1.record(stream)
    sync()
    launches = api.launch_count() - l0
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    units = S * N * world
    value = units * args.steps / (elapsed_ms * 1e-3)
    finite = bool(torch.isfinite(out[2]).all().item() and torch.isfinite(out[1]["a_smooth"]).all().item())
    del out

    # ---- end to end through the host-pointer C ABI (R's array layouts in host memory) ----
    Se = min(S, args.e2e_nsim)
    d1h, d2h = d1[:N * Se].cpu().numpy(), d2[:N * Se].cpu().numpy()

    def e2e_step():
        s1 = api.SimulateC(Se, 1, N, np.zeros(mc), Zc, Tc, Rc, Qc, np.zeros((mc, mc)), False, False, True, d1h)
        s2 = api.SimulateC(Se, 1, N, np.zeros(1), one, one, one, Qe, np.zeros((1, 1)), False, True, False, d2h)
        fs = api.FastSmootherC(s1["y"] + s2["eta"], *sm.args(), steps0, True)
        return api.ExtractComponentC(fs["a_t"], sm.Z)

    def timed(fn):
        fn()
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    legacy_value = Se * N * world / timed(e2e_step)
    m, r, p = sm.m, sm.r, 1
    legacy_h2d = 8 * Se * N * (1 + 1 + p + m)                # variates x2, y, a_t
    legacy_d2h = 8 * Se * N * ((p + 2 * mc + 1) + 1 + (2 * m + r) + p)

    # the call a user of the package makes: SimSmoother() (R/SimSmoother.R) as ONE call of the C ABI -- variates in,
    # epsilon / a / eta / the extracted component out; the per-draw intermediates never leave the device
    from statespacer_b200 import sysmat
    ks = api.KalmanC(air, np.isnan(air), *sm.args(), False)
    import math
    cs = sysmat.comp_sarima_univariate((12, 1), (0, 1), (1, 1), (1, 1), [0.5 * math.log(0.00135), 0.3, -0.4, -0.55])
    comps = [dict(a=cs["a"], Z=cs["Z"], T=cs["T"], R=cs["R"], Q=cs["Q"], P_star=cs["P_star"], repeat_Q=1,
                  draw_initial=True)]
    n_dr = mc * Se + N * Se + N * Se
    gh = np.random.default_rng(4 + rank)
    draws_h = gh.normal(size=n_dr)

    def one_call():
        return api.SimSmoother(Se, N, 1, comps, sm.Q[:1, :1, :], tuple(sm.args()), steps0, ks["a_smooth"][:, 1:],
                               ks["a_smooth"][:, :1], ks["eta"][:, 1:], draws_h, extract=[cs["Z"]])

    e2e_value = Se * N * world / timed(one_call)
    h2d = 8 * n_dr
    d2h = 8 * Se * N * (p + mc + (r - 1) + p)

    if rank == 0:
        b_sim, b_fs, b_ext = sim_bytes(N, p, m, r, mc, 1)
        hbm_peak, peak_src = 6545.3, "fallback"
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        fs_ms = statistics.mean(kt["smoother"])
        achieved = b_fs * S * N / (fs_ms * 1e-3) / 1e9
        fs_traffic = committed_traffic("fs_draw_reg_kernel", (S + 31) // 32)
        line = {"metric": SIM_METRIC, "value": value, "unit": SIM_UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": sim_config(world, S, {"results_finite": finite}),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": SIM_UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": 1, "nsim": Se,
                        "note": "ssgpu_SimSmoother: R's SimSmoother() as one call of the host-pointer C ABI (variates in; "
                                "epsilon, a, eta and the extracted component out, R array layouts, pageable host memory)",
                        "legacy_three_call": {"value": legacy_value, "h2d_bytes_per_step": legacy_h2d,
                                              "d2h_bytes_per_step": legacy_d2h,
                                              "note": "SimulateC x2 -> FastSmootherC -> ExtractComponentC through the "
                                                      "five legacy entry points, every array crossing PCIe"}},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "fs_draw_reg_kernel", "achieved": achieved, "peak": hbm_peak,
                             "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": fs_traffic["bytes"] if fs_traffic else None,
                             "traffic_source": fs_traffic["source"] if fs_traffic else
                             "no ncu capture of this kernel at this size under profiles/",
                             "peak_source": peak_src,
                             "bytes_per_draw_timestep": b_fs, "kernel_ms": fs_ms,
                             "other_kernels": {
                                 "simulate_kernel": {"kernel_ms": statistics.mean(kt["simulate"]),
                                                     "bytes_per_draw_timestep": 8 * (1 + p + mc + 1),
                                                     "achieved": 8 * (1 + p + mc + 1) * S * N
                                                     / (statistics.mean(kt["simulate"]) * 1e-3) / 1e9},
                                 "extract_kernel": {"kernel_ms": statistics.mean(kt["extract"]),
                                                    "bytes_per_draw_timestep": b_ext,
                                                    "achieved": b_ext * S * N
                                                    / (statistics.mean(kt["extract"]) * 1e-3) / 1e9}},
                             "pipeline_bytes_per_draw_timestep": b_sim + b_fs + b_ext,
                             "pipeline_achieved": (b_sim + b_fs + b_ext) * S * N * args.steps
                             / (elapsed_ms * 1e-3) / 1e9}}
        if world == 1 and not args.no_cpu:
            import contextlib
            import io
            buf = io.StringIO()
            a2 = argparse.Namespace(**vars(args))
            a2.steps, a2.warmup = 1, 0
            with contextlib.redirect_stdout(buf):
                run_sim_reference(a2)
            line["cpu_baseline"] = json.loads(buf.getvalue())["cpu_baseline"]
        return line
    return None


# ---- batched KalmanC (--workload kalman): filter + smoother of many series, HBM-bound by its outputs -------------
KAL_METRIC, KAL_UNIT = "kalman_filter_smoother_series_timesteps_per_s", "series*timesteps/s"
KAL_SETS = {"lean": ("loglik", "a_smooth", "V"),
            "full": ("loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil", "V", "P_inf_pred", "P_star_pred",
                     "P_inf_fil", "P_star_fil", "yfit", "v", "eta", "Fmat", "eta_var", "r_vec", "Nmat")}


def run_kalman_gpu(args):
    """StateSpaceEval() over a batch (SURVEY.md 8d, row `KalmanC` batched): config-5 model, B series per GPU, the
    arrays of --kalman-out written to HBM.  Roofline: HBM, algorithmic bytes = requested outputs + y."""
    import torch
    import torch.distributed as dist
    from statespacer_b200 import api
    world, rank, local, dev = gpu_setup()
    B = args.kalman_series
    sm, y, theta = make_shard(torch, dev, B, rank)
    var = torch.exp(2 * theta)
    Qd = torch.cat([var[:, :3], var[:, 3:4].expand(B, 11)], dim=1).contiguous()
    Pd = torch.zeros(B, M, device=dev, dtype=torch.float64)
    Pd[:, 0] = var[:, 0]
    model = api.BatchModel(sm, T_STEPS, B, 1, Q_diag=Qd, P_star_diag=Pd, device=dev)
    want = KAL_SETS[args.kalman_out]
    stream = torch.cuda.current_stream(dev)
    res = {}

    def step():
        res["o"] = api.kalman_batch_dev(y, model, want=want, stream=stream)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.1)
    l0 = api.launch_count()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync()
    launches = api.launch_count() - l0
    clocks = sampler.stop()
    elapsed_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    value = B * T_STEPS * world * args.steps / (elapsed_ms * 1e-3)
    shapes = api.kalman_shapes(T_STEPS, 1, M, M, False)
    out_bytes = 8 * sum(int(np.prod(shapes[k])) for k in want)
    finite = bool(torch.isfinite(res["o"]["a_smooth"]).all().item())
    # end to end: host arrays in, the requested arrays out
    Be = min(B, args.kalman_e2e_series)
    y_h = y[:, :Be].contiguous().cpu().numpy()
    Qh, Ph = Qd[:Be].cpu().numpy(), Pd[:Be].cpu().numpy()
    api.kalman_batch(y_h, sm, want=want, Q_diag=Qh, P_star_diag=Ph)
    t0 = time.perf_counter()
    api.kalman_batch(y_h, sm, want=want, Q_diag=Qh, P_star_diag=Ph)
    e2e_s = time.perf_counter() - t0
    if rank == 0:
        hbm_peak, peak_src = 6545.3, "fallback"
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        alg = (out_bytes + 8 * T_STEPS) * B
        achieved = alg * args.steps / (elapsed_ms * 1e-3) / 1e9
        line = {"metric": KAL_METRIC, "value": value, "unit": KAL_UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "KalmanC batched (filter + smoother): config-5 model, m=r=14, d=13 diffuse, T=1000",
                           "series_per_gpu": B, "outputs": list(want), "bytes_out_per_series_timestep": out_bytes / T_STEPS,
                           "l2": "outputs larger than L2 (%.1f GB per GPU), no flush" % (out_bytes * B / 1e9),
                           "results_finite": finite},
                "clocks": clocks,
                "e2e": {"value": Be * T_STEPS * world / e2e_s, "unit": KAL_UNIT, "h2d_bytes_per_step": 8 * Be * (T_STEPS + 2 * M),
                        "d2h_bytes_per_step": out_bytes * Be, "steps": 1, "series": Be,
                        "note": "ssgpu_kalman_batch, host pointers, pageable memory"},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": "fwd_generic_kernel<KALMAN> + bwd_generic_kernel",
                             "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                             "traffic": None, "peak_source": peak_src,
                             "note": "algorithmic bytes = the requested output arrays + y; both kernels of a step together"}}
        if world == 1 and not args.no_cpu:
            from concurrent.futures import ThreadPoolExecutor
            from oracle import cport, ref
            mod, kind = (ref, "reference") if ref.available() else (cport, "port")
            cores = host_cores()
            from statespacer_b200 import sysmat
            ys = y[:, :cores * 2].cpu().numpy()
            ths = theta[:cores * 2].cpu().numpy()

            def one(b):
                smb = sysmat.bsm12(ths[b])
                yb = ys[:, b:b + 1]
                return mod.KalmanC(yb, np.isnan(yb), *smb.args(), False)["a_smooth"][0, 0]

            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(one, range(cores * 2)))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": cores * 2 * T_STEPS / dt, "unit": KAL_UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{cores * 2} series x {T_STEPS} timesteps, one KalmanC call per series, "
                                              f"{cores} threads ({dt:.1f} s)"}
        return line
    return None


# ---- configs 2, 3 and 4 as batched loglikelihoods (--workload models) ------------------------------------------------
def model_cases():
    """BASELINE.json configs 2 (synthetic stand-in: bivariate BSM + regressors), 3 (dynamic Nelson-Siegel on
    FedYieldCurve) and 4's model (SARIMA(1,1,1)(0,1,1)12 on log AirPassengers) as concrete inputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import DNS_PARAM, config2_standin, load_fed
    from statespacer_b200 import sysmat
    c2 = config2_standin()
    air, sm4 = sim_model()
    sm3, _ = sysmat.dns_yield_curve(DNS_PARAM)
    nile = np.loadtxt(os.path.join(ROOT, "tests", "golden", "nile.txt"), comments="#")[:, None]
    sm1 = sysmat.nile_local_level(15100.252, 1468.724)                     # KAT-1a parameters
    return [("config1", "Nile local level, p=1, N=100, m=2 (batched by replication, SURVEY.md 8d)", sm1, nile, 1_000_000),
            ("config2", "bivariate BSM + 3 shared regressors (stand-in), p=2, N=192, m=34, time-varying Z, 5% missing", c2.sm, c2.y, 16384),
            ("config3", "dynamic Nelson-Siegel, FedYieldCurve 1985-2000, p=8, N=192, m=14", sm3, load_fed(), 262144),
            ("config4_loglik", "SARIMA(1,1,1)(0,1,1)12 on log AirPassengers, p=1, N=144, m=28 (companion-form T)", sm4, air, 16384)]


def run_models_gpu(args):
    """Batched LogLikC of configs 2, 3 and 4's model: the data of the config replicated B times with 1 % noise on the
    values, all matrices shared.  Roofline: FP64, algorithmic flops of the dense recursion (SURVEY.md 8d F_alg) -- these T
    are not block diagonal with 2 x 2 blocks, so the dense algorithm is what the kernels run."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from statespacer_b200 import api
    world, rank, local, dev = gpu_setup()
    peak_tf, _ = api.fp64_peak(0, 4000, 5)
    stream = torch.cuda.current_stream(dev)
    lines = {}
    for name, what, sm, y1, B in model_cases():
        N, p = y1.shape
        g = torch.Generator(device=dev)
        g.manual_seed(11 + rank)
        yd = torch.from_numpy(np.ascontiguousarray(y1)).to(dev)[:, :, None] * (
            1.0 + 0.01 * torch.randn(1, 1, B, generator=g, device=dev, dtype=torch.float64))
        yd = yd.contiguous()
        model = api.BatchModel(sm, N, B, 1, device=dev)
        out = torch.empty(1, B, device=dev, dtype=torch.float64)
        for _ in range(max(1, args.warmup)):
            api.loglik_batch_dev(yd, model, out, stream=stream)
        torch.cuda.synchronize(dev)
        l0 = api.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            api.loglik_batch_dev(yd, model, out, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / args.steps
        launches = api.launch_count() - l0
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        value = B * N * world / (ms * 1e-3)
        falg = f_alg(sm.m, p)
        # end to end: host arrays in, loglik out (pageable memory)
        Be = min(B, 4096)
        yh = yd[:, :, :Be].contiguous().cpu().numpy()
        api.loglik_batch(yh, sm)
        t0 = time.perf_counter()
        got = api.loglik_batch(yh, sm)
        e2e_s = time.perf_counter() - t0
        hbm_peak = 6545.3
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": f"{name}: {what}", "series_per_gpu": B, "timesteps": N,
                                                 "kernel": api.last_kernel(),
                                                 "results_finite": bool(np.all(np.isfinite(got)))},
                "e2e": {"value": Be * N * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": yh.nbytes,
                        "d2h_bytes_per_step": 8 * Be, "steps": 1, "series": Be, "note": "ssgpu_loglik_batch, pageable host arrays"},
                "gpu_launches": launches,
                "roofline": {"bound": "fp64", "kernel": api.last_kernel(), "achieved": value / world * falg / 1e12,
                             "peak": peak_tf, "unit": "TFLOP/s", "frac": value / world * falg / 1e12 / peak_tf, "traffic": None,
                             "flops_per_series_timestep": falg,
                             "note": "dense-algorithm flops 4m^3 + (4p+3)m^2 + 7pm per series*timestep (SURVEY.md 8d); the "
                                     "tensor / warp kernels skip the all-zero 8 x 8 tiles of T, so the executed count is lower",
                             "peak_source": "ssgpu_fp64_peak DFMA chains, measured in this run",
                             # a two-state model sits at the ridge: the HBM side of the roofline as well
                             "hbm_achieved_gbs": value / world * 8.0 * p / 1e9, "hbm_peak_gbs": hbm_peak,
                             "hbm_frac": value / world * 8.0 * p / 1e9 / hbm_peak}}
        if api.last_kernel().startswith("loglik_core"):
            # structure-exploiting kernel: its own executed count is the hardware fraction, the dense
            # figure stays beside it as an algorithm-equivalent rate
            mcore = int(re.search(r"MC=(\d+)", api.last_kernel()).group(1))   # component states left in the recursion
            fex = api.core_flops_per_step(p + mcore, p)
            rl = line["roofline"]
            rl["dense_algorithm_equivalent_frac"] = rl["frac"]
            rl["achieved"] = value / world * fex / 1e12
            rl["frac"] = rl["achieved"] / peak_tf
            rl["flops_per_series_timestep"] = fex
            rl["dense_flops_per_series_timestep"] = falg
            rl["note"] = ("executed FP64 operations (fma = 2) per series*timestep of loglik_core_kernel: the p residual states "
                          "are eliminated, deterministic component states tabulated, the remaining component states dense; dense_algorithm_equivalent_frac divides the "
                          "dense-algorithm count 4m^3 + (4p+3)m^2 + 7pm by the same time and is not a hardware fraction")
        if world == 1 and not args.no_cpu and rank == 0:
            from oracle import cport, ref
            mod, kind = (ref, "reference") if ref.available() else (cport, "port")
            cores = host_cores()
            n = cores * max(2, int(2e5 / (N * p * sm.m ** 2)))        # about a second of host work per model
            ys = yd[:, :, :n].cpu().numpy()
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(lambda b: mod.LogLikC(ys[:, :, b], np.isnan(ys[:, :, b]), *sm.args()), range(n)))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n * N / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{n} series x {N} timesteps, one LogLikC call per series on {cores} threads ({dt:.2f} s)"}
        lines[name] = line
        del yd, out, model
    return lines if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--series", type=int, default=B_PER_GPU, help="series per GPU")
    ap.add_argument("--missing", type=float, default=0.0, help="fraction of NaN observations (variant B: 0.02)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "blk", "blk_lanes", "core", "mma", "mma_dense", "cols", "wsm", "generic"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="all", choices=["all", "loglik", "sim", "kalman", "models"],
                    help="all (default): the config-5 loglik line (headline fields) with the sim and kalman lines attached "
                         "under `workloads`; loglik: config 5 alone; sim: config 4 simulation smoother draws; "
                         "kalman: batched filter + smoother (KalmanC) of config-5 series")
    ap.add_argument("--kalman-series", type=int, default=2048, help="series per GPU (--workload kalman)")
    ap.add_argument("--kalman-e2e-series", type=int, default=256, help="series of the end-to-end leg (--workload kalman)")
    ap.add_argument("--kalman-out", default="lean", choices=list(KAL_SETS), help="output arrays (--workload kalman)")
    ap.add_argument("--nsim", type=int, default=NSIM_PER_GPU, help="draws per GPU (--workload sim)")
    ap.add_argument("--e2e-nsim", type=int, default=NSIM_PER_GPU, help="draws of the end-to-end leg (--workload sim)")
    ap.add_argument("--cpu-draws", type=int, default=128, help="draws per host thread of the CPU arm (--workload sim)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): --series per GPU; strong: --series-total series cut into one shard per GPU")
    ap.add_argument("--series-total", type=int, default=1_000_000, help="series of the whole job (--scaling strong)")
    ap.add_argument("--single-process", type=int, default=0,
                    help="also time the end-to-end call from ONE process on this many GPUs (ssgpu_use_devices; --gpus 1 only)")
    ap.add_argument("--cpu-series", type=int, default=0, help="series of the CPU sample (0 = sized automatically)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_sim_reference(args) if args.workload == "sim" else run_reference(args)
        return
    import torch.distributed as dist
    if args.workload == "sim":
        line = run_sim_gpu(args)
    elif args.workload == "kalman":
        line = run_kalman_gpu(args)
    elif args.workload == "models":
        line = run_models_gpu(args)
    elif args.workload == "loglik":
        line = run_gpu(args)
    else:
        # the headline line (config 5) carries the other two workloads of the path as sub-results, each with its own
        # roofline, cpu_baseline and e2e; their step counts are bounded so that the default run stays within minutes
        line = run_gpu(args)
        sub = argparse.Namespace(**vars(args))
        sub.steps, sub.warmup = min(args.steps, 5), min(args.warmup, 3)
        sim = run_sim_gpu(sub)
        kal = run_kalman_gpu(sub)
        mods = run_models_gpu(sub)
        if line is not None:
            line["workloads"] = {"sim": sim, "kalman": kal, **(mods or {})}
    if line is not None:
        print(json.dumps(line))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
