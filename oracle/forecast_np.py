"""TEST INFRASTRUCTURE -- NumPy restatement of the forecast recursion of R/predict.R:205-246.  The product never
imports this file.  Parity: unpinned by the reference's own artefacts (no forecast value is printed in the vignettes);
checked against the closed form of the local level model in tests/test_oracle_kat.py."""
import numpy as np


def _sl(X, i):
    return X[:, :, i if X.shape[2] > 1 else 0]


def forecast(h, a1, P1, Z, T, R, Q, H):
    """a1 (m), P1 (m x m): KalmanC's a_fc / P_fc; Z p x m x {1|h}, T, R, Q, H likewise.  Returns y_fc (h x p),
    a_fc (h x m), P_fc (m x m x h), Fmat_fc (p x p x h)."""
    cube = lambda X: np.asarray(X, dtype=np.float64).reshape(np.shape(X)[0], np.shape(X)[1], -1)
    Z, T, R, Q, H = cube(Z), cube(T), cube(R), cube(Q), cube(H)
    p, m = Z.shape[:2]
    a = np.asarray(a1, dtype=np.float64).reshape(-1).copy()
    P = np.asarray(P1, dtype=np.float64).copy()
    y_fc, a_fc = np.zeros((h, p)), np.zeros((h, m))
    P_fc, F_fc = np.zeros((m, m, h)), np.zeros((p, p, h))
    for i in range(h):
        Zi, Ti, Ri, Qi, Hi = _sl(Z, i), _sl(T, i), _sl(R, i), _sl(Q, i), _sl(H, i)
        a_fc[i], P_fc[:, :, i] = a, P
        y_fc[i] = Zi @ a                                                   # :237
        F_fc[:, :, i] = (Zi @ P) @ Zi.T + Hi                               # :238
        if i < h - 1:                                                      # :241-245
            a = Ti @ a
            P = (Ti @ P) @ Ti.T + (Ri @ Qi) @ Ri.T
    return y_fc, a_fc, P_fc, F_fc
