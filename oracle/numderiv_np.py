"""TEST INFRASTRUCTURE -- restatement of numDeriv::hessian(method = "Richardson") as statespacer calls it for the
standard errors (R/statespacer.R:1057-1062: `numDeriv::hessian(func = result$loglik_fun, x = model_par)`, all
method.args at their defaults).  numDeriv (>= 2016.8-1.1, DESCRIPTION:28) is a third-party dependency that is not in
/root/reference; this follows its published algorithm (genD.default, Richardson extrapolation: eps = 1e-4, d = 0.1,
zero.tol = sqrt(.Machine$double.eps / 7e-7), r = 4, v = 2).  Parity unpinned by the reference's artefacts (no Hessian
is printed in the vignettes): checked against the analytic Hessian of a quadratic-plus-quartic function.
The product never imports this file."""
import numpy as np

EPS, D, R_IT, V = 1e-4, 0.1, 4, 2.0
ZERO_TOL = np.sqrt(np.finfo(float).eps / 7e-7)


def hessian(func, x):
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    f0 = func(x)
    h0 = np.abs(D * x) + EPS * (np.abs(x) < ZERO_TOL)
    Hdiag = np.zeros(n)
    for i in range(n):                                  # F''(i, i) with successively halved steps
        h = h0.copy()
        Ha = np.zeros(R_IT)
        e = np.zeros(n)
        e[i] = 1.0
        for k in range(R_IT):
            f1, f2 = func(x + e * h), func(x - e * h)
            Ha[k] = (f1 - 2 * f0 + f2) / h[i] ** 2
            h = h / V
        for m in range(1, R_IT):                        # Richardson extrapolation
            for k in range(R_IT - m):
                Ha[k] = (Ha[k + 1] * 4 ** m - Ha[k]) / (4 ** m - 1)
        Hdiag[i] = Ha[0]
    H = np.diag(Hdiag)
    for i in range(n):                                  # lower half
        for j in range(i):
            h = h0.copy()
            Da = np.zeros(R_IT)
            e = np.zeros(n)
            e[i] = e[j] = 1.0
            for k in range(R_IT):
                f1, f2 = func(x + e * h), func(x - e * h)
                Da[k] = (f1 - 2 * f0 + f2 - Hdiag[i] * h[i] ** 2 - Hdiag[j] * h[j] ** 2) / (2 * h[i] * h[j])
                h = h / V
            for m in range(1, R_IT):
                for k in range(R_IT - m):
                    Da[k] = (Da[k + 1] * 4 ** m - Da[k]) / (4 ** m - 1)
            H[i, j] = H[j, i] = Da[0]
    return H
