"""TEST INFRASTRUCTURE -- CPU restatement of R/SimSmoother.R:34-769 on top of the per-routine checkers.

The product never imports this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may.

`mod` is one of the checkers with the reference's five routines (oracle.cport: pinned-order C restatement,
oracle.ref: the reference's own C++ compiled unmodified, oracle.kalman_np); the function below strings their calls
together exactly as the R function does, with the normal variates injected in the order the reference consumes them
(src/SimulateC.cpp:75,112 per call; the calls in the order of R/SimSmoother.R).  Parity: pinned by the routines it
calls (tests/test_oracle_kat.py, tests/test_oracle_vs_ref.py); the R function itself has no golden output in the
reference."""
import numpy as np


def SimSmoother(mod, nsim, N, p, comps, H, full, initialisation_steps, a_hat, eps_hat, eta_hat, draws, extract=(),
                roots=None):
    """comps: list of dicts with a, Z, T, R, Q, P_star, repeat_Q, draw_initial (one SimulateC call each,
    R/SimSmoother.R:65-553); H: p x p x {1|N} residual variance (:556-573); full: (a, P_inf, P_star, Z, T, R, Q) of the
    model with the p residual states in front (:579-615); *_hat: object$smoothed$a / $epsilon / $eta;
    extract: Z_padded matrices (:645-766).  roots: optional per-call dicts with Q_root / P_root (shared with the
    implementation under test so that bit-exactness does not hinge on an SVD).
    Returns epsilon (N, p, nsim), a (N, m, nsim), eta (N, r, nsim) and the list of extracted components."""
    draws = np.asarray(draws, dtype=np.float64).reshape(-1)
    m = sum(np.asarray(c["a"]).size for c in comps)
    r = sum(np.asarray(c["Q"]).shape[0] * c["repeat_Q"] for c in comps)
    y = np.zeros((N, p, nsim), order="F")                                           # :53-61
    eta = np.zeros((N, r, nsim), order="F")
    a = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    off = ai = ei = 0
    roots = roots or [{} for _ in range(len(comps) + 1)]
    for c, rt in zip(comps, roots):
        mc = np.asarray(c["a"]).size
        r_tot = np.asarray(c["Q"]).shape[0] * c["repeat_Q"]
        cnt = (mc * nsim if c["draw_initial"] else 0) + N * r_tot * nsim
        sim = mod.SimulateC(nsim, c["repeat_Q"], N, c["a"], c["Z"], c["T"], c["R"], c["Q"], c["P_star"],
                            c["draw_initial"], False, True, draws[off:off + cnt], **rt)
        off += cnt
        y = y + sim["y"]                                                             # :104
        eta[:, ei:ei + r_tot, :] = sim["eta"]
        a[:, ai:ai + mc, :] = sim["a"]
        a_t[ai:ai + mc, :, :] = sim["a_t"]
        ai += mc
        ei += r_tot
    one = np.zeros((1, 1, 1))
    cnt = N * p * nsim
    sim = mod.SimulateC(nsim, 1, N, np.zeros(1), one, one, one, H, np.zeros((1, 1)), False, True, False,
                        draws[off:off + cnt], **roots[len(comps)])                   # :559-573
    off += cnt
    assert off == draws.size
    epsilon = sim["eta"]
    y = y + epsilon                                                                  # :575-576
    fs = mod.FastSmootherC(y, *full, initialisation_steps, True)                    # :617-628
    a_smooth = fs["a_smooth"][:, p:, :]
    eps_smooth = fs["a_smooth"][:, :p, :]
    eta_smooth = fs["eta"][:, p:, :]
    a_t_smooth = fs["a_t"][p:, :, :]
    a = (a - a_smooth) + np.asarray(a_hat)[:, :, None]                               # :634-642
    epsilon = (epsilon - eps_smooth) + np.asarray(eps_hat)[:, :, None]
    eta = (eta - eta_smooth) + np.asarray(eta_hat)[:, :, None]
    a_t = (a_t - a_t_smooth) + np.transpose(np.asarray(a_hat), (1, 0))[:, None, :]
    comps_out = [mod.ExtractComponentC(a_t, Z) for Z in extract]                     # :645-766
    return {"epsilon": epsilon, "a": a, "eta": eta, "components": comps_out}


def SimulateModel(mod, nsim, N, p, comps, H, draws, extract=(), roots=None):
    """The simulations of predict() over the forecast period (R/predict.R:142-856): one SimulateC call per component
    (started from object$predicted$a_fc / P_fc with draw_initial = TRUE, transposed_state = FALSE), stacked into
    y_sim / a_sim / eta_sim exactly like SimSmoother(), the residual draws last (:836-862).  extract: padded Z matrices;
    Z_c a_sim per draw equals the component's own sim$y (sim_list$level, ... at :300-301)."""
    draws = np.asarray(draws, dtype=np.float64).reshape(-1)
    m = sum(np.asarray(c["a"]).size for c in comps)
    r = sum(np.asarray(c["Q"]).shape[0] * c["repeat_Q"] for c in comps)
    y = np.zeros((N, p, nsim), order="F")                                           # :145-147
    eta = np.zeros((N, r, nsim), order="F")
    a = np.zeros((N, m, nsim), order="F")
    a_t = np.zeros((m, nsim, N), order="F")
    off = ai = ei = 0
    roots = roots or [{} for _ in range(len(comps) + 1)]
    own = []
    for c, rt in zip(comps, roots):
        mc = np.asarray(c["a"]).size
        r_tot = np.asarray(c["Q"]).shape[0] * c["repeat_Q"]
        cnt = (mc * nsim if c["draw_initial"] else 0) + N * r_tot * nsim
        sim = mod.SimulateC(nsim, c["repeat_Q"], N, c["a"], c["Z"], c["T"], c["R"], c["Q"], c["P_star"],
                            c["draw_initial"], False, True, draws[off:off + cnt], **rt)
        off += cnt
        own.append(sim["y"])                                                         # sim_list$<component>
        y = y + sim["y"]
        eta[:, ei:ei + r_tot, :] = sim["eta"]
        a[:, ai:ai + mc, :] = sim["a"]
        a_t[ai:ai + mc, :, :] = sim["a_t"]
        ai += mc
        ei += r_tot
    one = np.zeros((1, 1, 1))
    cnt = N * p * nsim
    sim = mod.SimulateC(nsim, 1, N, np.zeros(1), one, one, one, H, np.zeros((1, 1)), False, True, False,
                        draws[off:off + cnt], **roots[len(comps)])                   # :846-860
    assert off + cnt == draws.size
    epsilon = sim["eta"]
    y = y + epsilon
    return {"y": y, "epsilon": epsilon, "a": a, "eta": eta, "own_y": own,
            "components": [mod.ExtractComponentC(a_t, Z) for Z in extract]}
