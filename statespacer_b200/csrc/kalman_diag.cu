// KalmanC(diagnostics = TRUE): normalised prediction errors, smoothing errors and t-statistics
// (src/KalmanC.cpp:151-154,164-187,528-544).  Every quantity of time point i depends only on arrays the
// filter / smoother passes have already written (Fmat, v, P_pred, P_inf_pred, r_vec, Nmat), so the time points
// are independent: one thread per time point.
//   Finv   = inv(Fmat_i)                                   Gauss-Jordan with partial pivoting (arma::inv)
//   v_norm = v0 Finv^(1/2)   if !initialisation || all(|Z P_inf_pred Z'| < 1e-7)   (root by cyclic Jacobi; the
//            reference's U diag(sqrt(s)) U' from arma::svd is the same matrix for a symmetric positive Finv)
//   K      = T_i P_pred_i Z_i' Finv,   e = v0 Finv - r_{i+1} K,   D = Finv + K' N_{i+1} K
//   Tstat_state[i+1] = r_{i+1} / sqrt(diag N_{i+1}),   Tstat_observation[i] = e / sqrt(diag D)
// with v0 = v with missing entries replaced by 0, and NaN (NA) written back where v is not finite.
#include "common.cuh"

namespace ssgpu {

constexpr int kDiagMaxP = 16;

struct DiagArgs {
  DModel md;
  ssgpu_kalman_out k;   // device pointers
  const int* code;      // [Np] bit 8 of entry t*p: filter was diffuse when time point t started
  double* Kscr;         // [N][m*p] scratch
  double* NKscr;        // [N][m*p] scratch
};

__device__ void jacobi_root(double* a, double* v, double* out, int n) {
  // a: symmetric n x n (destroyed), v: scratch, out = V diag(sqrt|lambda|) V'
  for (int i = 0; i < n * n; i++) v[i] = 0.0;
  for (int i = 0; i < n; i++) v[i + i * n] = 1.0;
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) a[i + j * n] = a[j + i * n] = 0.5 * (a[i + j * n] + a[j + i * n]);
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        const double x = a[i + j * n] * a[i + j * n];
        if (i == j) diag += x; else off += x;
      }
    if (off == 0.0 || off <= 1e-32 * diag) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = a[p + q * n];
        if (apq == 0.0) continue;
        const double theta = (a[q + q * n] - a[p + p * n]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = a[k + p * n], akq = a[k + q * n];
          a[k + p * n] = c * akp - s * akq;
          a[k + q * n] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = a[p + k * n], aqk = a[q + k * n];
          a[p + k * n] = c * apk - s * aqk;
          a[q + k * n] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = v[k + p * n], vkq = v[k + q * n];
          v[k + p * n] = c * vkp - s * vkq;
          v[k + q * n] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double acc = 0.0;
      for (int k = 0; k < n; k++) acc += v[i + k * n] * sqrt(fabs(a[k + k * n])) * v[j + k * n];
      out[i + j * n] = acc;
    }
}

// inv(A) by Gauss-Jordan with partial pivoting; A (n x n, column-major) is destroyed
__device__ void gj_inverse(double* A, double* inv, int n) {
  for (int i = 0; i < n * n; i++) inv[i] = 0.0;
  for (int i = 0; i < n; i++) inv[i + i * n] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (fabs(A[r + c * n]) > fabs(A[piv + c * n])) piv = r;
    if (piv != c)
      for (int k = 0; k < n; k++) {
        double t = A[c + k * n]; A[c + k * n] = A[piv + k * n]; A[piv + k * n] = t;
        t = inv[c + k * n]; inv[c + k * n] = inv[piv + k * n]; inv[piv + k * n] = t;
      }
    const double d = 1.0 / A[c + c * n];
    for (int k = 0; k < n; k++) {
      A[c + k * n] *= d;
      inv[c + k * n] *= d;
    }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      const double f = A[r + c * n];
      if (f == 0.0) continue;
      for (int k = 0; k < n; k++) {
        A[r + k * n] -= f * A[c + k * n];
        inv[r + k * n] -= f * inv[c + k * n];
      }
    }
  }
}

__global__ void kalman_diag_kernel(const DiagArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const DModel& md = A.md;
  const int N = md.N, p = md.p, m = md.m;
  if (i >= N) return;
  const ssgpu_kalman_out& K = A.k;
  const int mm = m * m, pp = p * p;
  double Fm[kDiagMaxP * kDiagMaxP], Finv[kDiagMaxP * kDiagMaxP], W1[kDiagMaxP * kDiagMaxP], W2[kDiagMaxP * kDiagMaxP];
  double v0[kDiagMaxP], tmp[kDiagMaxP];
  bool vbad[kDiagMaxP];
  const double* Zi = md.Z.block(0, 0, md.Z.tv() ? i : 0);
  const double* Ti = md.T.block(0, 0, md.T.tv() ? i : 0);
  const double* Pp = K.P_pred + (long long)i * mm;
  const double* Nn = K.Nmat + (long long)(i + 1) * mm;
  for (int q = 0; q < pp; q++) Fm[q] = K.Fmat[(long long)i * pp + q];
  gj_inverse(Fm, Finv, p);
  for (int j = 0; j < p; j++) {
    const double vv = K.v[i + (long long)N * j];
    vbad[j] = !isfinite(vv);
    v0[j] = isnan(vv) ? 0.0 : vv;  // :170-171 replaces NA only
  }
  // v_norm (:151-154,175-187)
  bool do_vnorm = true;
  if (A.code[(long long)i * p] & 256) {  // diffuse at the start of this time point: test Z P_inf_pred Z'
    const double* Pi = K.P_inf_pred + (long long)i * mm;
    for (int a = 0; a < p && do_vnorm; a++)
      for (int b = 0; b < p && do_vnorm; b++) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) {
          double zp = 0.0;
          for (int l = 0; l < m; l++) zp = fma(Zi[a + l * p], Pi[l + k * m], zp);
          acc = fma(zp, Zi[b + k * p], acc);
        }
        if (!(fabs(acc) < kEps)) do_vnorm = false;
      }
  }
  if (K.v_norm) {
    if (do_vnorm) {
      for (int q = 0; q < pp; q++) W1[q] = Finv[q];
      jacobi_root(W1, W2, Fm, p);  // Fm <- Finv^(1/2)
      for (int j = 0; j < p; j++) {
        double acc = 0.0;
        for (int k = 0; k < p; k++) acc = fma(v0[k], Fm[k + j * p], acc);
        K.v_norm[i + (long long)N * j] = vbad[j] ? nan("") : acc;
      }
    } else {
      for (int j = 0; j < p; j++) K.v_norm[i + (long long)N * j] = nan("");
    }
  }
  // Tstat_state (:531-532)
  if (K.Tstat_state) {
    for (int k = 0; k < m; k++)
      K.Tstat_state[(i + 1) + (long long)(N + 1) * k] = K.r_vec[(i + 1) + (long long)(N + 1) * k] / sqrt(Nn[k + k * m]);
    if (i == 0)
      for (int k = 0; k < m; k++) K.Tstat_state[(long long)(N + 1) * k] = 0.0;  // row 0 is never written (:95)
  }
  // K = T P_pred Z' Finv (m x p), NK = N_{i+1} K
  double* Ks = A.Kscr + (long long)i * m * p;
  double* NKs = A.NKscr + (long long)i * m * p;
  for (int b = 0; b < p; b++) {
    // zb = Z' Finv[:,b]; pb = P zb; kb = T pb   (NKs column b used as scratch for zb, pb)
    for (int k = 0; k < m; k++) {
      double acc = 0.0;
      for (int a = 0; a < p; a++) acc = fma(Zi[a + k * p], Finv[a + b * p], acc);
      NKs[k + b * m] = acc;
    }
    for (int k = 0; k < m; k++) {
      double acc = 0.0;
      for (int l = 0; l < m; l++) acc = fma(Pp[k + l * m], NKs[l + b * m], acc);
      Ks[k + b * m] = acc;
    }
    for (int k = 0; k < m; k++) {
      double acc = 0.0;
      for (int l = 0; l < m; l++) acc = fma(mat_get(Ti, md.T.diag, m, k, l), Ks[l + b * m], acc);
      NKs[k + b * m] = acc;
    }
    for (int k = 0; k < m; k++) Ks[k + b * m] = NKs[k + b * m];
  }
  for (int b = 0; b < p; b++)
    for (int k = 0; k < m; k++) {
      double acc = 0.0;
      for (int l = 0; l < m; l++) acc = fma(Nn[k + l * m], Ks[l + b * m], acc);
      NKs[k + b * m] = acc;
    }
  // e = v0 Finv - r_{i+1} K ; D = Finv + K' N K ; Tstat_observation = e / sqrt(diag D)   (:535-543)
  for (int b = 0; b < p; b++) {
    double e1 = 0.0, e2 = 0.0;
    for (int a = 0; a < p; a++) e1 = fma(v0[a], Finv[a + b * p], e1);
    for (int k = 0; k < m; k++) e2 = fma(K.r_vec[(i + 1) + (long long)(N + 1) * k], Ks[k + b * m], e2);
    tmp[b] = vbad[b] ? nan("") : e1 - e2;
    if (K.e) K.e[i + (long long)N * b] = tmp[b];
  }
  for (int b = 0; b < p; b++)
    for (int a = 0; a < p; a++) {
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(Ks[k + a * m], NKs[k + b * m], acc);
      const double d = Finv[a + b * p] + acc;
      if (K.D) K.D[(long long)i * pp + a + b * p] = d;
      if (a == b && K.Tstat_observation) K.Tstat_observation[i + (long long)N * b] = tmp[b] / sqrt(d);
    }
}

int launch_kalman_diag(const DModel& md, const ssgpu_kalman_out* kout, const int* code, double* scratch,
                       cudaStream_t st) {
  SS_ARG(md.p <= kDiagMaxP, "KalmanC diagnostics: p > 16 is not supported");
  DiagArgs A{};
  A.md = md;
  A.k = *kout;
  A.code = code;
  A.Kscr = scratch;
  A.NKscr = scratch + (size_t)md.N * md.m * md.p;
  kalman_diag_kernel<<<(md.N + 63) / 64, 64, 0, st>>>(A);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu
