"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/ssgpu.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import math

import numpy as np
import pytest

from conftest import ROOT
from statespacer_b200 import native

HEADER = os.path.join(ROOT, "include", "ssgpu.h")

pytestmark = pytest.mark.skipif(not os.path.exists(native.LIB_PATH), reason="libssgpu.so not built")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ssgpu_[A-Za-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = native.lib()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ssgpu.h but not exported"
    assert sorted(native.SYMBOLS) == names


def test_version_and_error_string():
    L = native.lib()
    assert L.ssgpu_version() == 100
    assert isinstance(L.ssgpu_last_error(), bytes)


def test_argument_errors_do_not_need_a_device():
    L = native.lib()
    out = C.c_double()
    y = np.zeros(4)
    rc = L.ssgpu_LogLikC(y.ctypes.data, None, 4, 1, 0, 1, None, None, None, None, 1, None, 1, None, 1, None, 1,
                         C.byref(out))
    assert rc == 1 and L.ssgpu_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from statespacer_b200 import api, sysmat
    sm = sysmat.nile_local_level(1.0, 1.0)
    y = np.arange(10.0)[:, None]
    with pytest.raises(native.SsgpuError) as e:
        api.LogLikC(y, np.isnan(y), *sm.args())
    assert e.value.status == 2


# ---- host logic of the tensor kernel: state order that makes T tile-sparse (no device needed) ----------------
def _tile_mask(T, perm):
    m = T.shape[0]
    tt = (m + 1 + 7) // 8
    Tp = T[np.ix_(perm, perm)]
    mask = 0
    for i in range(m):
        for k in range(m):
            if Tp[i, k] != 0:
                mask |= 1 << ((i // 8) * tt + k // 8)
    return mask


def test_state_order_config5_is_tile_diagonal():
    """Config 5 (R/Components-Unobserved.R:150-159,504-519): the trigonometric pairs (j, j+5) straddle the two
    8-state tiles in the caller's order; re-ordered, T is tile diagonal and a time update needs 14 DMMAs, not 28."""
    from statespacer_b200 import api, sysmat
    T = sysmat.bsm12(np.zeros(4)).T[:, :, 0]
    ident = np.arange(14)
    assert _tile_mask(T, ident) == 0xF
    perm, mask, n = api.plan_state_order(T)
    assert sorted(perm) == list(range(14))
    assert mask == 0x9 and _tile_mask(T, perm) == 0x9 and n == 14
    # coupled states stay in one tile
    for i in range(14):
        for k in range(14):
            if T[perm[i], perm[k]] != 0:
                assert i // 8 == k // 8


def test_state_order_dense_and_oversized_components():
    from statespacer_b200 import api, sysmat
    rng = np.random.default_rng(0)
    T = rng.normal(size=(12, 12))
    perm, mask, n = api.plan_state_order(T)
    assert list(perm) == list(range(12)) and mask == 0xF and n == 28
    # airline model: one 27-state companion component cannot be packed, m = 28 is outside the kernel anyway
    T = np.zeros((20, 20))
    T[np.arange(19), np.arange(1, 20)] = 1.0          # a 20-state shift chain: one component > 8
    perm, mask, n = api.plan_state_order(T)
    assert list(perm) == list(range(20)) and mask == 0x1FF
    # three tiles, components of size 3 + 3 + 3 + 2 + 2 + 2 + 1 + 1 = 17 pack into 8 / 8 / 1
    T = np.zeros((17, 17))
    blocks = [[0, 9, 16], [1, 2, 3], [4, 10, 11], [5, 6], [7, 8], [12, 13], [14], [15]]
    for blk in blocks:
        for i in blk:
            for k in blk:
                T[i, k] = rng.normal()
    perm, mask, n = api.plan_state_order(T)
    assert sorted(perm) == list(range(17)) and mask == 0x111 and _tile_mask(T, perm) == 0x111
    # diagonal T (regression coefficients, local levels): already tile diagonal, caller's order kept
    perm, mask, n = api.plan_state_order(np.eye(10))
    assert list(perm) == list(range(10)) and mask == 0x9


def test_state_order_rejects_bad_arguments():
    L = native.lib()
    assert L.ssgpu_plan_state_order(None, 5, None, None, None) == 1
    T = np.eye(30)
    assert L.ssgpu_plan_state_order(T.ctypes.data, 30, None, None, None) == 1


def test_r_call_shim_type_checks():
    """r/src/ssgpu_rcall.c (the .Call shim a maintainer drops into the package) against declaration-only stand-ins
    of R's headers: catches signature drift between the shim and include/ssgpu.h in an image without R."""
    import subprocess
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-std=c11", "-I", os.path.join(ROOT, "tests", "rstub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "r", "src", "ssgpu_rcall.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "r", "src", "ssgpu_rcall.c")).read()
    for name, arity in (("ExtractComponentC", 2), ("FastSmootherC", 10), ("KalmanC", 10), ("LogLikC", 9),
                        ("SimulateC", 12)):                       # src/RcppExports.cpp:108-115
        assert f'{{"_statespacer_{name}", (DL_FUNC)&_statespacer_{name}, {arity}}}' in src


def test_plan_blocks_host_helper():
    """ssgpu_plan_blocks (host only): the block decomposition of the block-row kernel.  Config 5 is 7 blocks of coupled
    pairs (slope pair, five seasonal rotations) plus the paired-up singles; a SARIMA companion matrix, a dense Q that
    couples three states, or 33 uncoupled states do not decompose."""
    from statespacer_b200 import api, sysmat
    sm = sysmat.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]))
    T, R, Q = sm.T[:, :, 0], sm.R[:, :, 0], sm.Q[:, :, 0]
    pl = api.plan_blocks(T, R)
    assert pl["n_blocks"] == 7 and pl["lanes"] == 8 and pl["flop_per_step"] == 1470   # 609 DFMA + 251 DMUL + 1 DADD
    perm = pl["perm"].reshape(7, 2)
    assert sorted(perm.ravel().tolist()) == list(range(14))
    for a, b in perm:                                   # nothing of T outside the blocks
        others = [k for k in range(14) if k not in (a, b)]
        assert not T[np.ix_([a, b], others)].any() and not T[np.ix_(others, [a, b])].any()
    assert [1, 2] in perm.tolist()                      # level and slope share a block
    assert api.plan_blocks(T, R, Q)["n_blocks"] == 7    # diagonal Q given explicitly
    air = sysmat.airline([0.5 * math.log(0.00135), 0.4, 0.55])
    assert api.plan_blocks(air.T[:, :, 0], air.R[:, :, 0]) is None
    Qd = Q.copy()
    Qd[1, 2] = Qd[2, 1] = Qd[1, 3] = Qd[3, 1] = 0.01    # level disturbance correlated with two others
    assert api.plan_blocks(T, R, Qd) is None
    assert api.plan_blocks(np.eye(17), np.eye(17))["n_blocks"] == 9 and api.plan_blocks(np.eye(17), np.eye(17))["lanes"] == 16
    assert api.plan_blocks(np.eye(33), np.eye(33)) is None
    assert api.plan_blocks(np.eye(3), np.eye(3))["n_blocks"] == 2 and api.plan_blocks(np.eye(3), np.eye(3))["lanes"] == 2


def test_plan_thread_kernel_host_helper():
    """ssgpu_plan_thread_kernel (host only): the structural shortcuts of the thread-per-evaluation kernel and its shared
    diffuse steps, decided on host copies of the matrices.  The number of shared diffuse steps must be the oracle's
    `initialisation_steps` of a series without missing values (src/KalmanC.cpp:199-222: P_inf vanishes after exactly
    that many observations)."""
    from oracle import kalman_np
    from statespacer_b200 import api, sysmat
    S = sysmat
    sm = S.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05]))
    pl = api.plan_thread_kernel(sm, 1000)
    assert pl == dict(n_blocks=7, residual_dropped=1, level_slope_by_sums=1, zero_z_skipped=1, shared_diffuse_steps=13,
                      flop_per_step=992)                                  # 392 DFMA + 189 DMUL + 19 DADD
    assert api.plan_thread_kernel(sm, 10)["shared_diffuse_steps"] == 0     # shorter than the diffuse phase: the lanes
    rng = np.random.default_rng(2)
    models = [sm, S._combine(1, [[0.7]], [S.comp_level_slope(0.3, 0.02)]),
              S._combine(1, [[0.5]], [S.comp_local_level(0.2), S.comp_bsm_trig(4, 0.05)]),
              S._combine(1, [[0.5]], [S.comp_level_slope(0.2, 0.01), S.comp_bsm_trig(7, 0.05)])]
    for mdl in models:
        N = 40
        y = rng.normal(size=(N, 1)).cumsum(axis=0)
        want = int(kalman_np.KalmanC(y, np.isnan(y), *mdl.args(), False)["initialisation_steps"])
        got = api.plan_thread_kernel(mdl, N)
        assert got["shared_diffuse_steps"] == want and got["residual_dropped"] == 1 and got["zero_z_skipped"] == 1, (got, want)
    # one block only (local level): not the thread kernel's; a[residual] != 0: the residual stays, and with it the zero-z
    # shortcut (instantiated together); a SARIMA companion matrix does not decompose into blocks at all
    assert api.plan_thread_kernel(S.nile_local_level(2.0, 1.0), 100)["flop_per_step"] == 0
    a1 = sm.a.copy()
    a1[0] = 0.3
    pl = api.plan_thread_kernel(S.SysMat(Z=sm.Z, T=sm.T, R=sm.R, Q=sm.Q, a=a1, P_inf=sm.P_inf, P_star=sm.P_star), 1000)
    assert (pl["residual_dropped"], pl["level_slope_by_sums"], pl["zero_z_skipped"], pl["shared_diffuse_steps"]) == (0, 1, 0, 13)
    assert api.plan_thread_kernel(S.airline([0.5 * math.log(0.00135), 0.4, 0.55]), 144)["n_blocks"] == 0


def test_plan_thread_kernel_random_structural_models():
    """Property check of the host planner over random univariate structural models (level or level + slope, up to two
    trigonometric seasonals, an optional stationary cycle; 2 to 8 blocks): the shared diffuse steps equal the oracle's
    `initialisation_steps`, the residual is dropped, the level + slope shortcut applies exactly when the model has a
    slope, and the zero-loading shortcut as long as no two loaded singles share a block."""
    from oracle import kalman_np
    from statespacer_b200 import api, sysmat as S
    rng = np.random.default_rng(23)
    seen = 0
    for trial in range(40):
        slope = bool(rng.integers(0, 2))
        comps = [S.comp_level_slope(0.3, 0.02) if slope else S.comp_local_level(0.3)]
        singles = 0 if slope else 1                      # uncoupled loaded states besides the residual: the level ...
        for s_ in rng.choice([3, 4, 6, 7, 12], size=rng.integers(0, 3), replace=False):
            comps.append(S.comp_bsm_trig(int(s_), 0.05))
            singles += int(s_) % 2 == 0                  # ... and the Nyquist term of an even seasonal
        if rng.integers(0, 2):
            c, sn = 0.9 * math.cos(0.4), 0.9 * math.sin(0.4)
            comps.append(dict(Z=np.array([[1.0, 0.0]]), T=np.array([[c, sn], [-sn, c]]), R=np.eye(2), Q=0.3 * np.eye(2),
                              a=np.zeros(2), P_inf=np.zeros((2, 2)), P_star=0.3 / (1 - 0.81) * np.eye(2)))
        sm = S._combine(1, [[0.6]], comps)
        if sm.m > 16 or sm.m < 3:
            continue
        N = 3 * sm.m
        pl = api.plan_thread_kernel(sm, N)
        if not 2 <= pl["n_blocks"] <= 8:
            continue
        y = rng.normal(size=(N, 1)).cumsum(axis=0)
        want = int(kalman_np.KalmanC(y, np.isnan(y), *sm.args(), False)["initialisation_steps"])
        assert pl["shared_diffuse_steps"] == (want if want <= 16 else 0), (trial, pl, want)
        # the planner pairs the uncoupled states two by two, the residual first: from three loaded singles on, two of them
        # share a block and the zero-loading shortcut does not apply
        assert pl["residual_dropped"] == 1 and pl["level_slope_by_sums"] == int(slope), (trial, pl)
        assert pl["zero_z_skipped"] == int(singles <= 2), (trial, singles, pl)
        seen += 1
    assert seen >= 20


def test_plan_core_host_helper():
    """ssgpu_plan_core (host only): the residual structure the core kernel rests on and its split of the component states.
    Config 3 (dynamic Nelson-Siegel: 8 residuals + 3 factors + their 3 constant means) is a 3 x 3 problem without
    diffuse phase; the local level one diffuse state; 13 component states with a variance, correlated observation noise
    or a loading matrix that does not start with the identity are refused."""
    from conftest import DNS_PARAM
    from statespacer_b200 import api, sysmat as S
    dns, _ = S.dns_yield_curve(DNS_PARAM)
    assert api.plan_core(dns) == dict(eligible=1, n_stochastic=3, n_deterministic=3, diffuse=0)
    assert api.plan_core(S.nile_local_level(2.0, 1.0)) == dict(eligible=1, n_stochastic=1, n_deterministic=0, diffuse=1)
    assert api.plan_core(S.bsm12(0.5 * np.log([1.0, 0.1, 0.01, 0.05])))["eligible"] == 0
    bad = S.SysMat(Z=dns.Z, T=dns.T, R=dns.R, Q=dns.Q.copy(), a=dns.a, P_inf=dns.P_inf, P_star=dns.P_star.copy())
    bad.Q[0, 1, 0] = bad.Q[1, 0, 0] = bad.P_star[0, 1] = bad.P_star[1, 0] = 1e-4
    assert api.plan_core(bad)["eligible"] == 0
    Z2 = dns.Z.copy()
    Z2[0, 1, 0] = 0.5
    assert api.plan_core(S.SysMat(Z=Z2, T=dns.T, R=dns.R, Q=dns.Q, a=dns.a, P_inf=dns.P_inf, P_star=dns.P_star))["eligible"] == 0
    # a mean that is driven by a factor is not deterministic any more, and drags the means it drives along
    T2 = dns.T.copy()
    T2[8 + 3, 8 + 0, 0] = 0.1
    pl = api.plan_core(S.SysMat(Z=dns.Z, T=T2, R=dns.R, Q=dns.Q, a=dns.a, P_inf=dns.P_inf, P_star=dns.P_star))
    assert pl == dict(eligible=1, n_stochastic=4, n_deterministic=2, diffuse=0), pl


def test_core_flops_helper():
    """ssgpu_core_flops_per_step (host only): executed FP64 operations of the core kernel's regular phase -- config 3 with
    its 6 component states (2,082) and with the 3 deterministic means tabulated (648); 0 outside the kernel's range."""
    from statespacer_b200 import api
    assert api.core_flops_per_step(14, 8) == 2082 and api.core_flops_per_step(11, 8) == 648
    assert api.core_flops_per_step(2, 1) > 0 and api.core_flops_per_step(30, 1) == 0 and api.core_flops_per_step(20, 17) == 0


def test_plan_blocks_random_models():
    """Property check of the host planner over random models: a T made of 1x1 and 2x2 blocks, states shuffled, R = I
    and a diagonal Q must decompose into closed blocks covering every state once; one extra coupling that chains three
    states must be refused."""
    from statespacer_b200 import api
    rng = np.random.default_rng(11)
    for trial in range(60):
        sizes = rng.integers(1, 3, size=rng.integers(1, 12)).tolist()
        m = int(sum(sizes))
        T = np.zeros((m, m))
        o = 0
        for sz in sizes:
            blk = rng.normal(size=(sz, sz))
            if sz == 2 and rng.random() < 0.3:
                blk[1, 0] = 0.0                             # level/slope-like: coupled one way only
            T[o:o + sz, o:o + sz] = blk
            o += sz
        shuffle = rng.permutation(m)
        Ts = T[np.ix_(shuffle, shuffle)]
        pl = api.plan_blocks(Ts, np.eye(m))
        assert pl is not None, (trial, sizes)
        perm = pl["perm"].reshape(-1, 2)
        real = [k for k in perm.ravel().tolist() if k >= 0]
        assert sorted(real) == list(range(m)), (trial, perm)
        assert pl["n_blocks"] <= 16 and pl["lanes"] in (1, 2, 4, 8, 16) and pl["lanes"] >= min(pl["n_blocks"], 16) / 2
        for a, b in perm:
            inside = [k for k in (a, b) if k >= 0]
            others = [k for k in range(m) if k not in inside]
            assert not Ts[np.ix_(inside, others)].any() and not Ts[np.ix_(others, inside)].any(), (trial, perm)
        pairs = [i for i, sz in enumerate(sizes) if sz == 2]
        if pairs and m >= 3:                                # chain a third state onto a coupled pair
            o = int(sum(sizes[:pairs[0]]))
            third = (o + 2) % m
            Tc = T.copy()
            Tc[o, third] = 0.5
            assert api.plan_blocks(Tc[np.ix_(shuffle, shuffle)], np.eye(m)) is None, (trial, sizes)


def test_div_small_margin():
    """mm_dmma.cuh div_small(q, d) = trunc(__fdividef(q + 0.5f, d)) replaces q / d in the generic kernels for
    0 <= q < 4096, 1 <= d <= 64.  __fdividef is within 2 ulp of the quotient; the truncation is exact as long as
    (q + 0.5) / d stays further from every integer than that error -- checked here over the whole range with a
    64-ulp allowance."""
    q = np.arange(4096, dtype=np.float64)[:, None]
    d = np.arange(1, 65, dtype=np.float64)[None, :]
    x = (q + 0.5) / d
    err = 64 * np.spacing(x.astype(np.float32)).astype(np.float64)
    assert np.all(np.floor(x - err) == np.floor(q / d)) and np.all(np.floor(x + err) == np.floor(q / d))
    assert np.all(np.trunc(((q + 0.5).astype(np.float32) / d.astype(np.float32))) == np.floor(q / d))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own LogLikC on the host cores) runs without a GPU and prints one
    JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-series", "64"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "kalman_loglik_series_timesteps_per_s"
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
