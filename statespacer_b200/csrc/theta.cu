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
