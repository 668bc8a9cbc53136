"""Batched maximum-likelihood fits: the caller of the batched loglikelihood path.

statespacer() hands the objective LogLikC to stats::optim / optimx with method "BFGS" and numerical gradients
(R/statespacer.R:917-982): per series, per iteration, 1 + 2k + (line search) sequential objective calls.  Here all
B series advance in lock-step: one launch of ssgpu_loglik_grad_batch_dev gives every series its value and
central-difference gradient, one launch of ssgpu_loglik_theta_batch_dev evaluates every series' line-search
candidates, and the BFGS bookkeeping (k x k per series) is a handful of batched torch operations.  Series are
independent, so a multi-GPU fit is a shard per rank (statespacer_b200.shard) with nothing to exchange.

Only host-side orchestration lives here; the arithmetic of the objective is the CUDA path of include/ssgpu.h.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import api


# per-series outcome of fit_batch (stats::optim reports a convergence code as well; R/statespacer.R:984-997 prints it)
CONVERGED, ITERATION_LIMIT, LINE_SEARCH_FAILED, NON_FINITE = 0, 1, 2, 3


@dataclass
class FitResult:
    theta: "torch.Tensor"       # (B, k) parameters at the optimum (multistart: of the best start)
    loglik: "torch.Tensor"      # (B,)  loglik / N there (what LogLikC returns)
    grad: "torch.Tensor"        # (B, k) central-difference gradient there
    iterations: int
    converged: "torch.Tensor"   # (B,) bool: status == CONVERGED (max |grad| < gtol)
    evaluations: int            # filter runs per series (for the series*timesteps bookkeeping)
    status: "torch.Tensor" = None   # (B,) int: CONVERGED, ITERATION_LIMIT, LINE_SEARCH_FAILED (no candidate improved,
                                    # also after the inverse Hessian was reset to the identity), NON_FINITE (objective
                                    # not finite at the start values)
    start: "torch.Tensor" = None    # (B,) multistart: index of the start that won; None for a single start
    loglik_starts: "torch.Tensor" = None  # (S, B) multistart: loglik / N reached from every start


def fit_batch(y, model, theta0, q_index, p_star_index, max_iter=60, h=1e-3, gtol=1e-5, kernel="auto",
              steps=(1.0, 0.5, 0.25, 0.125, 0.0625, 1.0 / 64, 1.0 / 256), max_step=2.0, c1=1e-4):
    """Maximise loglik / N over theta for every series of the batch (BFGS, backtracking line search).

    y (N, B) or (N, p, B) CUDA float64 tensor, `model` an api.BatchModel built with device=... (shared matrices),
    theta0 (B, k) CUDA tensor of start values, q_index / p_star_index as for api.loglik_grad_batch_dev.
    theta0 of shape (S, B, k): MULTISTART -- S start vectors per series, all S x B optimisations advance in lock-step
    and share the launches (the S (1 + 2k) gradient points and the S line searches of a series are parameter vectors of
    that series: y is read once for all of them), the best start per series is returned."""
    import torch

    multi = theta0.dim() == 3
    S = theta0.shape[0] if multi else 1
    B, k = theta0.shape[-2:]
    dev = theta0.device
    theta = theta0.reshape(S * B, k).clone()                     # row s * B + b
    n_rows = S * B
    eye = torch.eye(k, device=dev, dtype=torch.float64).expand(n_rows, k, k)
    H = eye.clone()                                               # inverse-Hessian approximation of -loglik
    alphas = torch.tensor(steps, device=dev, dtype=torch.float64)
    L = alphas.numel()
    offs = torch.zeros(1 + 2 * k, k, device=dev, dtype=torch.float64)
    for i in range(k):
        offs[1 + 2 * i, i], offs[2 + 2 * i, i] = h, -h

    def theta_batch(pts):                                         # pts (V, S * B, k) -> (V, S * B): one launch
        V = pts.shape[0]
        if not multi:
            return api.loglik_theta_batch_dev(y, model, pts.contiguous(), q_index, p_star_index, kernel=kernel)
        arranged = pts.reshape(V, S, B, k).reshape(V * S, B, k).contiguous()
        return api.loglik_theta_batch_dev(y, model, arranged, q_index, p_star_index, kernel=kernel).reshape(V, S * B)

    def value_grad(th):
        if not multi:                                             # variance sets and differences formed on the device
            return api.loglik_grad_batch_dev(y, model, th, q_index, p_star_index, h=h, kernel=kernel)
        fv = theta_batch(th[None] + offs[:, None, :])
        return fv[0], ((fv[1::2] - fv[2::2]) / (2.0 * h)).transpose(0, 1).contiguous()

    ll, gr = value_grad(theta)
    f, g = -ll, -gr
    n_eval = 1 + 2 * k
    status = torch.full((n_rows,), ITERATION_LIMIT, device=dev, dtype=torch.int64)
    status[~torch.isfinite(f)] = NON_FINITE
    done = ~torch.isfinite(f)
    reset = torch.zeros(n_rows, device=dev, dtype=torch.bool)     # H already reset to the identity after a failed search
    it = 0
    for it in range(1, max_iter + 1):
        d = -torch.bmm(H, g.unsqueeze(2)).squeeze(2)
        slope = (g * d).sum(1)
        bad = ~(slope < 0)                                        # not a descent direction: steepest descent
        if bad.any():
            H = torch.where(bad[:, None, None], eye, H)
            d = torch.where(bad[:, None], -g, d)
            slope = (g * d).sum(1)
        scale = torch.clamp(max_step / d.abs().amax(1).clamp_min(1e-300), max=1.0)   # variances are exp(2 theta)
        d = d * scale[:, None]
        slope = slope * scale
        cand = theta[None] + alphas[:, None, None] * d[None]      # (L, rows, k)
        fc = -theta_batch(cand)
        n_eval += L
        fc = torch.where(torch.isfinite(fc), fc, torch.full_like(fc, float("inf")))
        armijo = fc <= f[None] + c1 * alphas[:, None] * slope[None]
        first = torch.where(armijo.any(0), armijo.float().argmax(0), fc.argmin(0))   # largest step passing, else best
        f_new = fc.gather(0, first[None]).squeeze(0)
        improved = (f_new < f) & ~done
        a_sel = alphas[first]
        theta_new = torch.where(improved[:, None], theta + a_sel[:, None] * d, theta)
        ll_new, gr_new = value_grad(theta_new)
        n_eval += 1 + 2 * k
        g_new = -gr_new
        s = theta_new - theta
        yv = g_new - g
        sy = (s * yv).sum(1)
        upd = improved & (sy > 1e-12)
        rho = torch.where(upd, 1.0 / sy.clamp_min(1e-300), torch.zeros_like(sy))
        I_rsy = eye - rho[:, None, None] * s.unsqueeze(2) * yv.unsqueeze(1)
        H_new = torch.bmm(torch.bmm(I_rsy, H), I_rsy.transpose(1, 2)) + rho[:, None, None] * s.unsqueeze(2) * s.unsqueeze(1)
        H = torch.where(upd[:, None, None], H_new, H)
        theta, f, g = theta_new, torch.where(improved, -ll_new, f), torch.where(improved[:, None], g_new, g)
        small = g.abs().amax(1) < gtol
        failed = ~improved & ~done & ~small
        # a search that found nothing along the BFGS direction gets one more along the gradient (H := I) before the
        # series is given up; a second failure in a row is final
        give_up = failed & reset
        retry = failed & ~reset
        H = torch.where(retry[:, None, None], eye, H)
        reset = (reset | retry) & ~improved
        status = torch.where(small & ~done, torch.full_like(status, CONVERGED), status)
        status = torch.where(give_up, torch.full_like(status, LINE_SEARCH_FAILED), status)
        done = done | small | give_up
        if bool(done.all()):
            break
    if not multi:
        return FitResult(theta=theta, loglik=-f, grad=-g, iterations=it, converged=status == CONVERGED,
                         evaluations=n_eval, status=status)
    fs = torch.where(torch.isfinite(f), f, torch.full_like(f, float("inf"))).reshape(S, B)
    best = fs.argmin(0)
    pick = lambda x: x.reshape((S, B) + x.shape[1:]).gather(
        0, best.reshape((1, B) + (1,) * (x.dim() - 1)).expand((1, B) + x.shape[1:])).squeeze(0)
    st = pick(status)
    return FitResult(theta=pick(theta), loglik=-pick(f), grad=-pick(g), iterations=it, converged=st == CONVERGED,
                     evaluations=S * n_eval, status=st, start=best, loglik_starts=-fs)


def hessian_batch(y, model, theta, q_index, p_star_index, h=1e-2, kernel="auto"):
    """Central-difference Hessian of loglik / N at theta for every series from ONE launch of 1 + 2k + 2k(k-1)
    parameter vectors per series (the points of the sweep statespacer() hands to numDeriv::hessian for the standard
    errors, R/statespacer.R:1057-1062; numDeriv refines them by Richardson extrapolation, this is the plain second
    difference).  Returns (B, k, k)."""
    import torch

    B, k = theta.shape
    dev = theta.device
    offsets = [torch.zeros(k, device=dev, dtype=torch.float64)]
    index = {}
    for i in range(k):
        for sgn in (1.0, -1.0):
            e = torch.zeros(k, device=dev, dtype=torch.float64)
            e[i] = sgn * h
            index[(i, sgn)] = len(offsets)
            offsets.append(e)
    for i in range(k):
        for j in range(i + 1, k):
            for si in (1.0, -1.0):
                for sj in (1.0, -1.0):
                    e = torch.zeros(k, device=dev, dtype=torch.float64)
                    e[i], e[j] = si * h, sj * h
                    index[(i, j, si, sj)] = len(offsets)
                    offsets.append(e)
    pts = theta[None] + torch.stack(offsets)[:, None, :]          # (V, B, k)
    f = api.loglik_theta_batch_dev(y, model, pts.contiguous(), q_index, p_star_index, kernel=kernel)
    Hm = torch.empty(B, k, k, device=dev, dtype=torch.float64)
    for i in range(k):
        Hm[:, i, i] = (f[index[(i, 1.0)]] - 2.0 * f[0] + f[index[(i, -1.0)]]) / (h * h)
        for j in range(i + 1, k):
            v = (f[index[(i, j, 1.0, 1.0)]] - f[index[(i, j, 1.0, -1.0)]] - f[index[(i, j, -1.0, 1.0)]]
                 + f[index[(i, j, -1.0, -1.0)]]) / (4.0 * h * h)
            Hm[:, i, j] = v
            Hm[:, j, i] = v
    return Hm


def hessian_richardson_batch(y, model, theta, q_index, p_star_index, kernel="auto"):
    """numDeriv::hessian(method = "Richardson") of loglik / N at theta for every series -- the sweep behind
    statespacer()'s standard errors (R/statespacer.R:1057-1062; numDeriv's genD with its defaults eps = 1e-4, d = 0.1,
    r = 4, v = 2) -- with all 1 + 2 k r + k (k - 1) r parameter vectors of a series (81 for k = 4) evaluated in ONE
    batched launch; the Richardson tables are a few torch operations.  Returns (B, k, k)."""
    import torch

    B, k = theta.shape
    dev = theta.device
    eps, d, r, v = 1e-4, 0.1, 4, 2.0
    zero_tol = (torch.finfo(torch.float64).eps / 7e-7) ** 0.5
    h0 = (d * theta).abs() + eps * (theta.abs() < zero_tol).to(torch.float64)        # (B, k)
    pts, where = [theta], {}

    def add(key, offs):
        where[key] = len(pts)
        pts.append(theta + offs)

    eye = torch.eye(k, device=dev, dtype=torch.float64)
    for i in range(k):
        for it in range(r):
            step = eye[i][None] * h0 / v ** it
            add((i, i, it, 1), step)
            add((i, i, it, -1), -step)
    for i in range(k):
        for j in range(i):
            for it in range(r):
                step = (eye[i] + eye[j])[None] * h0 / v ** it
                add((i, j, it, 1), step)
                add((i, j, it, -1), -step)
    f = api.loglik_theta_batch_dev(y, model, torch.stack(pts).contiguous(), q_index, p_star_index, kernel=kernel)
    f0 = f[0]

    def richardson(tab):                                 # tab: list of r tensors (B,)
        tab = list(tab)
        for m_ in range(1, r):
            for kk in range(r - m_):
                tab[kk] = (tab[kk + 1] * 4 ** m_ - tab[kk]) / (4 ** m_ - 1)
        return tab[0]

    Hm = torch.zeros(B, k, k, device=dev, dtype=torch.float64)
    for i in range(k):
        hi = [h0[:, i] / v ** it for it in range(r)]
        Hm[:, i, i] = richardson([(f[where[(i, i, it, 1)]] - 2 * f0 + f[where[(i, i, it, -1)]]) / hi[it] ** 2
                                  for it in range(r)])
    for i in range(k):
        for j in range(i):
            tab = []
            for it in range(r):
                hi, hj = h0[:, i] / v ** it, h0[:, j] / v ** it
                tab.append((f[where[(i, j, it, 1)]] - 2 * f0 + f[where[(i, j, it, -1)]] - Hm[:, i, i] * hi ** 2
                            - Hm[:, j, j] * hj ** 2) / (2 * hi * hj))
            Hm[:, i, j] = Hm[:, j, i] = richardson(tab)
    return Hm
