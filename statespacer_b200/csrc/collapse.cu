// Collapse transform of R/statespacer.R:815-900: y_t (p observations) -> y*_t = A*_t y_t (m2 < p) with
// A* = (Z' H^-1 Z)^-1 Z' H^-1, H* = A* H A*', and the loglikelihood contribution of the part projected away
//     loglik_add = sum_t [ log(det H*_t / det H_t) / 2 - e_t' H_t^-1 e_t / 2 ] - (#obs(y) - #obs(y*)) log(2 pi) / 2,
// e_t = y_t - Z_t y*_t (missing y entered as 0, :407-409; y* == 0 becomes NA, :830).  R does this in interpreted code
// inside the optimiser's objective for every parameter vector of a `collapse = TRUE` model.
//
//   collapse_prepare_kernel   one warp per (H_t, Z_t) slice: H^-1, A*, H*, log(det H* / det H) by Gauss-Jordan with
//                             partial pivoting, lane j owning column j of the augmented matrix (p <= 16)
//   collapse_apply_kernel     one thread per time point: y*_t, e_t' H^-1 e_t, the observation counts
//   collapse_reduce_kernel    one warp: the sums, in a fixed order
#include "common.cuh"

namespace ssgpu {

constexpr int kColMaxP = 16;

// In-place inverse of the n x n matrix in a[.][.] (row stride kColMaxP) by Gauss-Jordan on [A | I]; lane j < 2n owns
// column j.  Returns det(A) (0 when singular: the inverse is then garbage, as solve() would stop in R).
__device__ double warp_inverse(double (*aug)[2 * kColMaxP], int n, int lane) {
  double det = 1.0;
  for (int k = 0; k < n; k++) {
    // pivot: largest |a[i][k]|, i >= k (every lane computes the same answer)
    int piv = k;
    double best = fabs(aug[k][k]);
    for (int i = k + 1; i < n; i++) {
      const double v = fabs(aug[i][k]);
      if (v > best) {
        best = v;
        piv = i;
      }
    }
    __syncwarp();
    if (piv != k) {
      if (lane < 2 * n) {
        const double t = aug[k][lane];
        aug[k][lane] = aug[piv][lane];
        aug[piv][lane] = t;
      }
      det = -det;
    }
    __syncwarp();
    const double d = aug[k][k];
    det *= d;
    if (d == 0.0) return 0.0;
    __syncwarp();
    if (lane < 2 * n) aug[k][lane] = aug[k][lane] / d;
    // column k before any lane touches it (lane k rewrites it below)
    double fcol[kColMaxP];
#pragma unroll
    for (int i = 0; i < kColMaxP; i++) fcol[i] = i < n ? aug[i][k] : 0.0;
    __syncwarp();
    if (lane < 2 * n) {
      const double pk = aug[k][lane];
#pragma unroll
      for (int i = 0; i < kColMaxP; i++)
        if (i < n && i != k) aug[i][lane] = fma(-fcol[i], pk, aug[i][lane]);
    }
    __syncwarp();
  }
  return det;
}

struct CollapseArgs {
  const double* y;  // N x p column-major, NaN = missing
  const double* H;  // p x p x nH
  const double* Z;  // p x m2 x nZ
  int N, p, m2, nH, nZ, ns;
  double* A_star;   // [ns][m2 x p]
  double* Hinv;     // [ns][p x p]
  double* H_star;   // [ns][m2 x m2]  (output)
  double* ld;       // [ns] log(det H* / det H)
  double* y_star;   // N x m2 column-major (output)
  double* term;     // [N]
  int* cnt;         // [2 N]
  double* out;      // loglik_add
};

__global__ void __launch_bounds__(32) collapse_prepare_kernel(const CollapseArgs A) {
  __shared__ double aug[kColMaxP][2 * kColMaxP];
  __shared__ double Hi[kColMaxP][kColMaxP], ZtH[kColMaxP][kColMaxP], As[kColMaxP][kColMaxP], tmp[kColMaxP][kColMaxP];
  const int s = blockIdx.x, lane = threadIdx.x, p = A.p, m2 = A.m2;
  const double* H = A.H + (A.nH > 1 ? (size_t)s * p * p : 0);
  const double* Z = A.Z + (A.nZ > 1 ? (size_t)s * p * m2 : 0);
  // H^-1, det H
  for (int i = 0; i < p; i++)
    if (lane < 2 * p) aug[i][lane] = lane < p ? H[i + (size_t)lane * p] : (lane - p == i ? 1.0 : 0.0);
  __syncwarp();
  const double detH = warp_inverse(aug, p, lane);
  for (int i = 0; i < p; i++)
    if (lane < p) Hi[i][lane] = aug[i][p + lane];
  __syncwarp();
  // ZtHinv = Z' H^-1 (m2 x p): lane = column
  for (int i = 0; i < m2; i++)
    if (lane < p) {
      double acc = 0.0;
      for (int k = 0; k < p; k++) acc = fma(Z[k + (size_t)i * p], Hi[k][lane], acc);
      ZtH[i][lane] = acc;
    }
  __syncwarp();
  // G = ZtHinv Z (m2 x m2) -> inverse
  for (int i = 0; i < m2; i++)
    if (lane < 2 * m2) {
      double acc = lane - m2 == i ? 1.0 : 0.0;
      if (lane < m2) {
        acc = 0.0;
        for (int k = 0; k < p; k++) acc = fma(ZtH[i][k], Z[k + (size_t)lane * p], acc);
      }
      aug[i][lane] = acc;
    }
  __syncwarp();
  warp_inverse(aug, m2, lane);
  // A* = G^-1 ZtHinv (m2 x p)
  for (int i = 0; i < m2; i++)
    if (lane < p) {
      double acc = 0.0;
      for (int k = 0; k < m2; k++) acc = fma(aug[i][m2 + k], ZtH[k][lane], acc);
      As[i][lane] = acc;
      A.A_star[(size_t)s * m2 * p + i + (size_t)lane * m2] = acc;
    }
  for (int i = 0; i < p; i++)
    if (lane < p) A.Hinv[(size_t)s * p * p + i + (size_t)lane * p] = Hi[i][lane];
  __syncwarp();
  // H* = A* H A*': tmp = A* H (m2 x p), then tmp A*'
  for (int i = 0; i < m2; i++)
    if (lane < p) {
      double acc = 0.0;
      for (int k = 0; k < p; k++) acc = fma(As[i][k], H[k + (size_t)lane * p], acc);
      tmp[i][lane] = acc;
    }
  __syncwarp();
  for (int i = 0; i < m2; i++)
    if (lane < 2 * m2) {
      double acc = lane - m2 == i ? 1.0 : 0.0;
      if (lane < m2) {
        acc = 0.0;
        for (int k = 0; k < p; k++) acc = fma(tmp[i][k], As[lane][k], acc);
        A.H_star[(size_t)s * m2 * m2 + i + (size_t)lane * m2] = acc;
      }
      aug[i][lane] = acc;
    }
  __syncwarp();
  const double detHs = warp_inverse(aug, m2, lane);
  if (lane == 0) A.ld[s] = log(detHs / detH);
}

__global__ void collapse_apply_kernel(const CollapseArgs A) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= A.N) return;
  const int p = A.p, m2 = A.m2, N = A.N;
  const int s = A.ns > 1 ? t : 0;
  const double* As = A.A_star + (size_t)s * m2 * p;
  const double* Hi = A.Hinv + (size_t)s * p * p;
  const double* Z = A.Z + (A.nZ > 1 ? (size_t)t * p * m2 : 0);
  double yt[kColMaxP], ys[kColMaxP], e[kColMaxP];
  int n_obs = 0, n_star = 0;
  for (int j = 0; j < p; j++) {
    const double v = A.y[t + (size_t)N * j];
    n_obs += !isnan(v);
    yt[j] = isnan(v) ? 0.0 : v;  // :407-409
  }
  for (int i = 0; i < m2; i++) {
    double acc = 0.0;
    for (int j = 0; j < p; j++) acc = fma(As[i + (size_t)j * m2], yt[j], acc);
    ys[i] = acc;
    n_star += acc != 0.0;
    A.y_star[t + (size_t)N * i] = acc == 0.0 ? nan("") : acc;  // y_kal[y_kal == 0] <- NA
  }
  for (int j = 0; j < p; j++) {
    double acc = 0.0;
    for (int i = 0; i < m2; i++) acc = fma(Z[j + (size_t)i * p], ys[i], acc);
    e[j] = yt[j] - acc;
  }
  double q = 0.0;
  for (int j = 0; j < p; j++) {
    double acc = 0.0;
    for (int k = 0; k < p; k++) acc = fma(Hi[j + (size_t)k * p], e[k], acc);
    q = fma(e[j], acc, q);
  }
  A.term[t] = 0.5 * A.ld[s] - 0.5 * q;
  A.cnt[2 * t] = n_obs;
  A.cnt[2 * t + 1] = n_star;
}

__global__ void __launch_bounds__(32) collapse_reduce_kernel(const CollapseArgs A) {
  const int lane = threadIdx.x;
  double acc = 0.0;
  long long dn = 0;
  for (int t = lane; t < A.N; t += 32) {
    acc += A.term[t];
    dn += A.cnt[2 * t] - A.cnt[2 * t + 1];
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    dn += __shfl_xor_sync(0xffffffffu, dn, o);
  }
  if (lane == 0) *A.out = acc - 0.5 * (double)dn * 1.8378770664093454835606594728112;  // log(2 pi)
}

int launch_collapse(const CollapseArgs& A, cudaStream_t st) {
  collapse_prepare_kernel<<<A.ns, 32, 0, st>>>(A);
  collapse_apply_kernel<<<(A.N + 127) / 128, 128, 0, st>>>(A);
  collapse_reduce_kernel<<<1, 32, 0, st>>>(A);
  count_launch(3);
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu

using namespace ssgpu;

extern "C" int ssgpu_collapse(const double* y, int N, int p, int m2, const double* H, int nH, const double* Z, int nZ,
                              double* y_star, double* H_star, double* loglik_add) {
  SS_ARG(y && H && Z && y_star && H_star && loglik_add, "collapse: null pointer argument");
  SS_ARG(N >= 1 && m2 >= 1 && p > m2, "collapse: needs 1 <= m2 < p (R/statespacer.R:386-395)");
  SS_ARG(p <= kColMaxP, "collapse: p > 16 is not supported");
  SS_ARG((nH == 1 || nH == N) && (nZ == 1 || nZ == N), "collapse: H and Z must have 1 or N slices");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  const int ns = (nH > 1 || nZ > 1) ? N : 1;
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  const size_t n_y = (size_t)N * p, n_H = (size_t)p * p * nH, n_Z = (size_t)p * m2 * nZ, n_ys = (size_t)N * m2;
  const size_t n_A = (size_t)ns * m2 * p, n_Hi = (size_t)ns * p * p, n_Hs = (size_t)ns * m2 * m2;
  const size_t total = pad(n_y) + pad(n_H) + pad(n_Z) + pad(n_ys) + pad(n_A) + pad(n_Hi) + pad(n_Hs) + pad(ns) + pad(N) +
                       pad(N) + 32;
  DevBuf blk;
  SS_TRY(blk.alloc(total * sizeof(double), st));
  double* b = blk.as<double>();
  CollapseArgs A{};
  A.N = N; A.p = p; A.m2 = m2; A.nH = nH; A.nZ = nZ; A.ns = ns;
  double* yd = b; b += pad(n_y);
  double* Hd = b; b += pad(n_H);
  double* Zd = b; b += pad(n_Z);
  A.y_star = b; b += pad(n_ys);
  A.A_star = b; b += pad(n_A);
  A.Hinv = b; b += pad(n_Hi);
  A.H_star = b; b += pad(n_Hs);
  A.ld = b; b += pad(ns);
  A.term = b; b += pad(N);
  A.cnt = reinterpret_cast<int*>(b); b += pad(N);
  A.out = b;
  SS_CUDA(cudaMemcpyAsync(yd, y, n_y * sizeof(double), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(Hd, H, n_H * sizeof(double), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(Zd, Z, n_Z * sizeof(double), cudaMemcpyHostToDevice, st));
  A.y = yd; A.H = Hd; A.Z = Zd;
  SS_TRY(launch_collapse(A, st));
  SS_CUDA(cudaMemcpyAsync(y_star, A.y_star, n_ys * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(H_star, A.H_star, n_Hs * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(loglik_add, A.out, sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}
