// Out-of-sample forecast recursion of R/predict.R:205-246:
//     y_fc[i] = Z_i a_i,   F_i = Z_i P_i Z_i' + H_i,   a_{i+1} = T_i a_i,   P_{i+1} = T_i P_i T_i' + R_i Q_i R_i',
// started from KalmanC's a_fc / P_fc.  Sequential in i and small (one call per predict()): one CTA, the covariance in
// shared memory, every product element-parallel with k-ascending fma chains.
#include "common.cuh"

namespace ssgpu {

struct FcArgs {
  int h, p, m, r, nZ, nT, nR, nQ, nH;
  const double *a1, *P1, *Z, *T, *R, *Q, *H;
  double *y_fc, *a_fc, *P_fc, *F_fc;
};

__global__ void __launch_bounds__(256) forecast_kernel(const FcArgs A) {
  extern __shared__ double sh[];
  const int m = A.m, p = A.p, r = A.r, h = A.h, tid = threadIdx.x, NT = blockDim.x;
  const int mm = m * m;
  double* P = sh;            // m x m
  double* W = P + mm;        // max(m x m, p x m, m x r)
  double* av = W + (mm > p * m ? (mm > m * r ? mm : m * r) : (p * m > m * r ? p * m : m * r));
  double* a2 = av + m;
  for (int q = tid; q < mm; q += NT) P[q] = A.P1[q];
  for (int q = tid; q < m; q += NT) av[q] = A.a1[q];
  __syncthreads();
  for (int i = 0; i < h; i++) {
    const double* Z = A.Z + (A.nZ > 1 ? (size_t)i * p * m : 0);
    const double* T = A.T + (A.nT > 1 ? (size_t)i * mm : 0);
    const double* R = A.R + (A.nR > 1 ? (size_t)i * m * r : 0);
    const double* Q = A.Q + (A.nQ > 1 ? (size_t)i * r * r : 0);
    const double* H = A.H + (A.nH > 1 ? (size_t)i * p * p : 0);
    for (int q = tid; q < m; q += NT) A.a_fc[i + (size_t)h * q] = av[q];
    for (int q = tid; q < mm; q += NT) A.P_fc[(size_t)i * mm + q] = P[q];
    for (int j = tid; j < p; j += NT) {  // y_fc[i, ] = Z a
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(Z[j + (size_t)k * p], av[k], acc);
      A.y_fc[i + (size_t)h * j] = acc;
    }
    for (int q = tid; q < p * m; q += NT) {  // W = Z P (p x m)
      const int j = q % p, c = q / p;
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(Z[j + (size_t)k * p], P[k + c * m], acc);
      W[q] = acc;
    }
    __syncthreads();
    for (int q = tid; q < p * p; q += NT) {  // F = (Z P) Z' + H
      const int j = q % p, c = q / p;
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(W[j + k * p], Z[c + (size_t)k * p], acc);
      A.F_fc[(size_t)i * p * p + q] = acc + H[q];
    }
    __syncthreads();
    if (i == h - 1) break;
    for (int q = tid; q < m; q += NT) {  // a = T a
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(T[q + (size_t)k * m], av[k], acc);
      a2[q] = acc;
    }
    for (int q = tid; q < mm; q += NT) {  // W = T P
      const int j = q % m, c = q / m;
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(T[j + (size_t)k * m], P[k + c * m], acc);
      W[q] = acc;
    }
    __syncthreads();
    for (int q = tid; q < mm; q += NT) {  // P = (T P) T'
      const int j = q % m, c = q / m;
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(W[j + k * m], T[c + (size_t)k * m], acc);
      P[q] = acc;
    }
    for (int q = tid; q < m; q += NT) av[q] = a2[q];
    __syncthreads();
    for (int q = tid; q < m * r; q += NT) {  // W = R Q (m x r)
      const int j = q % m, c = q / m;
      double acc = 0.0;
      for (int k = 0; k < r; k++) acc = fma(R[j + (size_t)k * m], Q[k + (size_t)c * r], acc);
      W[q] = acc;
    }
    __syncthreads();
    for (int q = tid; q < mm; q += NT) {  // P += (R Q) R'
      const int j = q % m, c = q / m;
      double acc = 0.0;
      for (int k = 0; k < r; k++) acc = fma(W[j + k * m], R[c + (size_t)k * m], acc);
      P[q] = P[q] + acc;
    }
    __syncthreads();
  }
}

}  // namespace ssgpu

using namespace ssgpu;

extern "C" int ssgpu_forecast(int h, int p, int m, int r, const double* a1, const double* P1, const double* Z, int nZ,
                              const double* T, int nT, const double* R, int nR, const double* Q, int nQ, const double* H,
                              int nH, double* y_fc, double* a_fc, double* P_fc, double* Fmat_fc) {
  SS_ARG(a1 && P1 && Z && T && R && Q && H && y_fc && a_fc && P_fc && Fmat_fc, "forecast: null pointer argument");
  SS_ARG(h >= 1 && p >= 1 && m >= 1 && r >= 1 && m <= SSGPU_MAX_M && p <= SSGPU_MAX_M && r <= SSGPU_MAX_M,
         "forecast: dimensions out of range");
  SS_ARG((nZ == 1 || nZ == h) && (nT == 1 || nT == h) && (nR == 1 || nR == h) && (nQ == 1 || nQ == h) &&
             (nH == 1 || nH == h), "forecast: system matrices must have 1 or forecast_period slices");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  const size_t mm = (size_t)m * m;
  const size_t in[7] = {(size_t)m, mm, (size_t)p * m * nZ, mm * nT, (size_t)m * r * nR, (size_t)r * r * nQ, (size_t)p * p * nH};
  const double* src[7] = {a1, P1, Z, T, R, Q, H};
  const size_t out[4] = {(size_t)h * p, (size_t)h * m, mm * h, (size_t)p * p * h};
  double* dst[4] = {y_fc, a_fc, P_fc, Fmat_fc};
  size_t total = 0;
  for (size_t n : in) total += pad(n);
  for (size_t n : out) total += pad(n);
  DevBuf blk;
  SS_TRY(blk.alloc(total * sizeof(double), st));
  double* b = blk.as<double>();
  const double* din[7];
  for (int i = 0; i < 7; i++) {
    SS_CUDA(cudaMemcpyAsync(b, src[i], in[i] * sizeof(double), cudaMemcpyHostToDevice, st));
    din[i] = b;
    b += pad(in[i]);
  }
  double* dout[4];
  for (int i = 0; i < 4; i++) {
    dout[i] = b;
    b += pad(out[i]);
  }
  FcArgs A{h, p, m, r, nZ, nT, nR, nQ, nH, din[0], din[1], din[2], din[3], din[4], din[5], din[6],
           dout[0], dout[1], dout[2], dout[3]};
  size_t w = mm;
  if ((size_t)p * m > w) w = (size_t)p * m;
  if ((size_t)m * r > w) w = (size_t)m * r;
  const size_t smem = (mm + w + 2 * (size_t)m) * sizeof(double);
  SS_CUDA(cudaFuncSetAttribute(forecast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  forecast_kernel<<<1, 256, smem, st>>>(A);
  count_launch();
  SS_CUDA(cudaGetLastError());
  for (int i = 0; i < 4; i++)
    SS_CUDA(cudaMemcpyAsync(dst[i], dout[i], out[i] * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}
