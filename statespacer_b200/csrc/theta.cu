// Device-side construction of the per-evaluation variances from parameter vectors, and the finite-difference
// gradient of the batched loglikelihood.
//
// statespacer's default parameterisation of a univariate variance is  sigma^2 = exp(2 theta)  (R/Cholesky.R:148-167,
// placed on the diagonals of Q and P_star by R/statespacer.R:444-509,776-807).  For such models -- local level,
// slope, trigonometric seasonal, cycle-free BSMs -- the map theta -> (Q, P_star) is two index vectors, so the host
// ships k doubles per evaluation instead of r + m, and R never rebuilds system matrices inside the optim loop
// (R/statespacer.R:917-927 does that once per objective call).
#include "common.cuh"

namespace ssgpu {

// One thread per evaluation (v, b): Qd[(v*B + b)*r + j] = q_index[j] >= 0 ? exp(2 th[q_index[j]]) : q_fixed[j];
// same for P_star.  fd != 0: theta is [b*k + i] and variant v is theta (v = 0), theta + h e_i (v = 1 + 2i),
// theta - h e_i (v = 2 + 2i) -- stats::optim's central differences; fd == 0: theta is [(v*B + b)*k + i].
__global__ void expand_theta_kernel(const double* theta, long long B, int V, int k, int r, int m, const int* q_index,
                                    const int* p_index, const double* q_fixed, const double* p_fixed, double h, int fd,
                                    double* Qd, double* Pd) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * V) return;
  const long long b = e % B;
  const int v = (int)(e / B);
  const double* th = fd ? theta + b * k : theta + e * k;
  const int shifted = fd && v > 0 ? (v - 1) / 2 : -1;
  const double shift = (fd && v > 0) ? ((v & 1) ? h : -h) : 0.0;
  for (int j = 0; j < r; j++) {
    const int i = q_index[j];
    Qd[e * r + j] = i >= 0 ? exp(2.0 * (th[i] + (i == shifted ? shift : 0.0))) : q_fixed[j];
  }
  for (int j = 0; j < m; j++) {
    const int i = p_index[j];
    Pd[e * m + j] = i >= 0 ? exp(2.0 * (th[i] + (i == shifted ? shift : 0.0))) : p_fixed[j];
  }
}

// loglik[b] = out[b], grad[b*k + i] = (out[(1+2i)*B + b] - out[(2+2i)*B + b]) / (2h)
__global__ void fd_gradient_kernel(const double* out, long long B, int k, double h, double* loglik, double* grad) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  loglik[b] = out[b];
  for (int i = 0; i < k; i++) grad[b * k + i] = (out[(long long)(1 + 2 * i) * B + b] - out[(long long)(2 + 2 * i) * B + b]) / (2.0 * h);
}

// ---- covariance blocks from the LDL' parameterisation (R/Cholesky.R:148-167) ---------------------------------------
// A block of n x n of Q or P_star is  L D L'  with  D = diag(exp(2 theta[t0 .. t0+n-1]))  and the strict lower triangle
// of the unit lower-triangular L filled column by column from theta[t0+n ...] (R fills `chol_L[format_n0_lt] <- param`
// in column-major order, :151-153).  A diagonal entry of D <= 1e-7 makes the reference return NA (:159-161): the
// evaluation is flagged invalid and the block written as the identity so that the filter runs on finite numbers; the
// caller overwrites the loglikelihood with NaN afterwards.
// One thread per evaluation: Qe / Pe (dense, column-major, r x r / m x m) = template with the blocks overwritten.
__global__ void expand_cov_kernel(const double* theta, long long B, int V, int k, int r, int m, const ssgpu_cov_block* blocks,
                                  int n_blocks, const double* q_tpl, const double* p_tpl, double h, int fd, double* Qe,
                                  double* Pe, int* valid) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * V) return;
  const long long b = e % B;
  const int v = (int)(e / B);
  const double* th = fd ? theta + b * k : theta + e * k;
  const int shifted = fd && v > 0 ? (v - 1) / 2 : -1;
  const double shift = (fd && v > 0) ? ((v & 1) ? h : -h) : 0.0;
  auto par = [&](int i) { return th[i] + (i == shifted ? shift : 0.0); };
  double* Q = Qe + (size_t)e * r * r;
  for (int i = 0; i < r * r; i++) Q[i] = q_tpl[i];
  double* P = Pe ? Pe + (size_t)e * m * m : nullptr;
  if (P)
    for (int i = 0; i < m * m; i++) P[i] = p_tpl[i];
  int ok = 1;
  for (int bi = 0; bi < n_blocks; bi++) {
    const ssgpu_cov_block bl = blocks[bi];
    const int n = bl.n, ld = bl.target == 0 ? r : m;
    double* M = (bl.target == 0 ? Q : P) + bl.offset + (size_t)bl.offset * ld;
    double d[8], Lm[8][8];
    bool good = true;
    for (int i = 0; i < n; i++) {
      d[i] = exp(2.0 * par(bl.theta_offset + i));
      good = good && !(d[i] <= 1e-7);  // R/Cholesky.R:159
    }
    int idx = bl.theta_offset + n;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) Lm[i][j] = i == j ? 1.0 : 0.0;
    for (int j = 0; j < n; j++)
      for (int i = j + 1; i < n; i++) Lm[i][j] = par(idx++);
    ok = ok && good;
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n; i++) {
        double acc = 0.0;  // (L D L')[i][j] = sum_l L[i][l] d[l] L[j][l], l ascending (tcrossprod(L %*% D, L), :164)
        for (int l = 0; l <= (i < j ? i : j); l++) acc = fma(Lm[i][l] * d[l], Lm[j][l], acc);
        M[i + (size_t)j * ld] = good ? acc : (i == j ? 1.0 : 0.0);
      }
  }
  valid[e] = ok;
}

// out[e] = NaN where the parameter vector of evaluation e is not admissible
__global__ void mask_invalid_kernel(double* out, const int* valid, long long E) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E && !valid[e]) out[e] = nan("");
}

int launch_expand_cov(const double* theta, long long B, int V, int k, int r, int m, const ssgpu_cov_block* blocks,
                      int n_blocks, const double* q_tpl, const double* p_tpl, double h, int fd, double* Qe, double* Pe,
                      int* valid, cudaStream_t st) {
  const long long E = B * V;
  expand_cov_kernel<<<(unsigned)((E + 63) / 64), 64, 0, st>>>(theta, B, V, k, r, m, blocks, n_blocks, q_tpl, p_tpl, h, fd,
                                                             Qe, Pe, valid);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_mask_invalid(double* out, const int* valid, long long E, cudaStream_t st) {
  mask_invalid_kernel<<<(unsigned)((E + 255) / 256), 256, 0, st>>>(out, valid, E);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_expand_theta(const double* theta, long long B, int V, int k, int r, int m, const int* q_index,
                        const int* p_index, const double* q_fixed, const double* p_fixed, double h, int fd, double* Qd,
                        double* Pd, cudaStream_t st) {
  const long long E = B * V;
  expand_theta_kernel<<<(unsigned)((E + 127) / 128), 128, 0, st>>>(theta, B, V, k, r, m, q_index, p_index, q_fixed,
                                                                  p_fixed, h, fd, Qd, Pd);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_fd_gradient(const double* out, long long B, int k, double h, double* loglik, double* grad, cudaStream_t st) {
  fd_gradient_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(out, B, k, h, loglik, grad);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu
