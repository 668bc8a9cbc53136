// Simulation smoother kernels: one thread per draw, draws contiguous in memory.
//
//   simulate_kernel        src/SimulateC.cpp:71-131      unconditional draws from injected normal variates
//   fs_draw_kernel         src/FastSmootherC.cpp:88-368  innovations, backward r recursion, forward reconstruction
//   extract_kernel         src/ExtractComponentC.cpp:31-37
//   permute_* kernels      R's array layouts  <->  the draw-contiguous layout used on the device
//
// Device ("native") layout of every per-draw array: X[(t*K + k)*S + s], S = nsim, so that a warp's
// 32 draws read and write 256 contiguous bytes.  R's layouts (N x K x nsim and m x nsim x N, column
// major) are produced by tiled shared-memory transposes.
//
// The FastSmootherC gain sequence (F, K0, K1 per univariate step) does not depend on y; it is computed
// ONCE by fwd_generic_kernel<MODE_GAINS> and shared by all draws, where the reference recomputes the
// whole covariance recursion for every draw (src/FastSmootherC.cpp:88,150-251).
// Arithmetic follows the pinned order of oracle/kalman_c.c exactly (fma chains, k ascending): the
// draws are bit-identical to the C oracle's given the same variates.
#include "common.cuh"

namespace ssgpu {

constexpr int kSimThreads = 128;

// per-thread vector v[k] kept in shared memory as sh[k*blockDim.x + tid]: conflict-free
struct ShVec {
  double* p;
  __device__ __forceinline__ double& operator[](int k) const { return p[k * kSimThreads]; }
};

__device__ __forceinline__ const double* slice_of(const DMat& M, int t) { return M.block(0, 0, M.tv() ? t : 0); }

// ---- SimulateC (SimArgs in common.cuh) ----------------------------------------------------------
__global__ void __launch_bounds__(kSimThreads) simulate_kernel(const SimArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  const long long s = (long long)blockIdx.x * kSimThreads + tid;
  if (s >= A.S) return;
  const int m = A.m, p = A.p, r = A.r, r_tot = A.r * A.repeat_Q, rR = A.rR, N = A.N;
  const long long S = A.S;
  ShVec av{sh + tid}, a2{sh + (size_t)m * kSimThreads + tid}, et{sh + (size_t)2 * m * kSimThreads + tid};
  if (!A.eta_only) {
    for (int i = 0; i < m; i++) {
      if (A.draw_initial) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(A.P_root[i + k * m], A.d0[k + (long long)m * s], acc);
        av[i] = A.a0[i] + acc;
      } else {
        av[i] = A.a0[i];
      }
    }
  }
  const long long total = (long long)r_tot * S;
  for (int t = 0; t < N; t++) {
    const double* Qr = A.Q_root + (A.nQ > 1 ? (long long)t * r * r : 0);
    const double* dd = A.dt + total * t + (long long)r_tot * s;
    for (int q = 0; q < A.repeat_Q; q++)
      for (int rr = 0; rr < r; rr++) {
        double acc = 0.0;
        for (int l = 0; l < r; l++) acc = fma(Qr[rr + l * r], dd[l + r * q], acc);
        et[rr + r * q] = acc;
        if (A.eta) A.eta[((long long)t * r_tot + rr + r * q) * S + s] = acc;
      }
    if (A.eta_only) continue;
    const double* Zt = slice_of(A.Z, t);
    const double* Tt = slice_of(A.T, t);
    const double* Rt = slice_of(A.R, t);
    if (A.a)
      for (int k = 0; k < m; k++) A.a[((long long)t * m + k) * S + s] = av[k];
    for (int j = 0; j < p; j++) {
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(Zt[j + k * p], av[k], acc);
      if (A.y) A.y[((long long)t * p + j) * S + s] = acc;
    }
    if (t < N - 1) {
      for (int i = 0; i < m; i++) {
        double acc = 0.0, acc1 = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Tt[i + k * m], av[k], acc);
        for (int k = 0; k < rR; k++) acc1 = fma(Rt[i + k * m], et[k], acc1);
        a2[i] = acc + acc1;
      }
      for (int i = 0; i < m; i++) av[i] = a2[i];
    }
  }
}

// ---- FastSmootherC per-draw passes ----------------------------------------------------------------
__global__ void __launch_bounds__(kSimThreads) fs_draw_kernel(const FsArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  const long long s = (long long)blockIdx.x * kSimThreads + tid;
  if (s >= A.S) return;
  const int m = A.m, p = A.p, r = A.r, N = A.N;
  const long long S = A.S;
  const size_t vs = (size_t)m * kSimThreads;
  ShVec av{sh + tid}, a2{sh + vs + tid}, rr{sh + 2 * vs + tid}, r1{sh + 3 * vs + tid}, r2{sh + 4 * vs + tid},
      et{sh + 5 * vs + tid};

  // forward: innovations v = y - z'a with the shared gains (src/FastSmootherC.cpp:139-251)
  for (int k = 0; k < m; k++) av[k] = A.a0[k];
  for (int t = 0; t < N; t++) {
    const double* Zt = slice_of(A.Z, t);
    const double* Tt = slice_of(A.T, t);
    for (int j = 0; j < p; j++) {
      const long long index = (long long)t * p + j;
      if ((A.g.code[index] & 255) == 0) continue;
      double za = 0.0;
      for (int k = 0; k < m; k++) za = fma(Zt[j + k * p], av[k], za);
      const double vv = A.y[index * S + s] - za;
      A.v[index * S + s] = vv;
      const double* k0 = A.g.K0 + index * m;
      for (int k = 0; k < m; k++) av[k] = fma(k0[k], vv, av[k]);
    }
    if (t < N - 1) {
      for (int i = 0; i < m; i++) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Tt[i + k * m], av[k], acc);
        a2[i] = acc;
      }
      for (int i = 0; i < m; i++) av[i] = a2[i];
    }
  }

  // backward: r_t (and r^(1) while diffuse), src/FastSmootherC.cpp:259-334
  for (int k = 0; k < m; k++) {
    rr[k] = 0.0;
    r1[k] = 0.0;
  }
  for (int t = N - 1; t >= 0; t--) {
    const double* Zt = slice_of(A.Z, t);
    for (int j = p - 1; j >= 0; j--) {
      const long long index = (long long)t * p + j;
      const int c = A.g.code[index] & 255;
      if (c == 0) continue;
      const double* k0 = A.g.K0 + index * m;
      const double* k1 = A.g.K1 + index * m;
      const double c1 = A.g.F1[index] * A.v[index * S + s];
      double d0 = 0.0;
      for (int k = 0; k < m; k++) d0 = fma(k0[k], rr[k], d0);
      if (c == 1) {
        const double sc = c1 - d0;
        for (int k = 0; k < m; k++) rr[k] = fma(Zt[j + k * p], sc, rr[k]);
      } else {
        double d1 = 0.0, d2 = 0.0;
        for (int k = 0; k < m; k++) {
          d1 = fma(k0[k], r1[k], d1);
          d2 = fma(k1[k], rr[k], d2);
        }
        const double s1 = (c1 - d1) - d2, s0 = -d0;
        for (int k = 0; k < m; k++) {
          const double z = Zt[j + k * p];
          r1[k] = fma(z, s1, r1[k]);
          rr[k] = fma(z, s0, rr[k]);
        }
      }
    }
    for (int k = 0; k < m; k++) A.a_smooth[((long long)t * m + k) * S + s] = rr[k];
    if (t > 0) {
      const double* Tp = slice_of(A.T, t - 1);
      const bool both = ((long long)t * p) < A.init_steps;
      for (int jj = 0; jj < m; jj++) {
        double acc = 0.0, acc1 = 0.0;
        for (int k = 0; k < m; k++) {
          acc = fma(Tp[k + jj * m], rr[k], acc);
          if (both) acc1 = fma(Tp[k + jj * m], r1[k], acc1);
        }
        r2[jj] = acc;
        a2[jj] = acc1;
      }
      for (int k = 0; k < m; k++) {
        rr[k] = r2[k];
        if (both) r1[k] = a2[k];
      }
    }
  }

  // forward reconstruction (src/FastSmootherC.cpp:338-367); rr still holds r_0
  for (int i = 0; i < m; i++) {
    double acc = 0.0, acc1 = 0.0;
    for (int k = 0; k < m; k++) {
      acc = fma(A.P_star[i + k * m], rr[k], acc);
      acc1 = fma(A.P_inf[i + k * m], r1[k], acc1);
    }
    av[i] = (A.a0[i] + acc) + acc1;
  }
  for (int t = 0; t < N; t++) {
    if (t > 0) {
      const double* Tp = slice_of(A.T, t - 1);
      const double* Rp = slice_of(A.R, t - 1);
      const double* Qp = A.g.QtR + (A.qtr_tv ? (long long)(t - 1) * r * m : 0);
      for (int k = 0; k < m; k++) rr[k] = A.a_smooth[((long long)t * m + k) * S + s];  // r_t
      for (int kk = 0; kk < r; kk++) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Qp[kk + k * r], rr[k], acc);
        et[kk] = acc;
        A.eta[((long long)(t - 1) * r + kk) * S + s] = acc;
      }
      for (int i = 0; i < m; i++) {
        double acc = 0.0, acc1 = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Tp[i + k * m], av[k], acc);
        for (int k = 0; k < r; k++) acc1 = fma(Rp[i + k * m], et[k], acc1);
        a2[i] = acc + acc1;
      }
      for (int i = 0; i < m; i++) av[i] = a2[i];
    }
    for (int k = 0; k < m; k++) A.a_smooth[((long long)t * m + k) * S + s] = av[k];
  }
  for (int kk = 0; kk < r; kk++) A.eta[((long long)(N - 1) * r + kk) * S + s] = 0.0;  // :81
}

// ---- ExtractComponentC ----------------------------------------------------------------------------
// a: R layout m x nsim x N (a_native == 0) or native [(t*m+k)*S+s]; out native [(t*p+j)*S+s]
__global__ void __launch_bounds__(kSimThreads) extract_kernel(const double* a, int a_native, DMat Z, int m, int p, int S,
                                                              int N, double* out) {
  const long long s = (long long)blockIdx.x * kSimThreads + threadIdx.x;
  const int t = blockIdx.y;
  if (s >= S) return;
  const double* Zt = slice_of(Z, t);
  for (int j = 0; j < p; j++) {
    double acc = 0.0;
    for (int k = 0; k < m; k++) {
      const double ak = a_native ? a[((long long)t * m + k) * S + s] : a[k + (long long)m * (s + (long long)S * t)];
      acc = fma(Zt[j + k * p], ak, acc);
    }
    out[((long long)t * p + j) * S + s] = acc;
  }
}

// ---- layout changes -------------------------------------------------------------------------------
// native X[(t*K+k)*S+s]  <->  R array [t + N*(k + K*s)]; one 32x32 (t, s) tile per block, blockIdx.z = k
template <bool TO_R>
__global__ void permute_tks_kernel(const double* src, double* dst, int N, int K, long long S) {
  __shared__ double tile[32][33];
  const int k = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int t0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (TO_R) {
    for (int dy = ty; dy < 32; dy += 8) {
      const int t = t0 + dy;
      const long long s = s0 + tx;
      if (t < N && s < S) tile[dy][tx] = src[((long long)t * K + k) * S + s];
    }
    __syncthreads();
    for (int dy = ty; dy < 32; dy += 8) {
      const long long s = s0 + dy;
      const int t = t0 + tx;
      if (t < N && s < S) dst[t + (long long)N * (k + (long long)K * s)] = tile[tx][dy];
    }
  } else {
    for (int dy = ty; dy < 32; dy += 8) {
      const long long s = s0 + dy;
      const int t = t0 + tx;
      if (t < N && s < S) tile[tx][dy] = src[t + (long long)N * (k + (long long)K * s)];
    }
    __syncthreads();
    for (int dy = ty; dy < 32; dy += 8) {
      const int t = t0 + dy;
      const long long s = s0 + tx;
      if (t < N && s < S) dst[((long long)t * K + k) * S + s] = tile[dy][tx];
    }
  }
}

// native X[(t*m+k)*S+s] -> a_t [k + m*(s + S*t)]; one 32x32 (k, s) tile per block, blockIdx.z = t
__global__ void permute_at_kernel(const double* src, double* dst, int m, long long S) {
  __shared__ double tile[32][33];
  const int t = blockIdx.z;
  const long long s0 = (long long)blockIdx.x * 32;
  const int k0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int dy = ty; dy < 32; dy += 8) {
    const int k = k0 + dy;
    const long long s = s0 + tx;
    if (k < m && s < S) tile[dy][tx] = src[((long long)t * m + k) * S + s];
  }
  __syncthreads();
  for (int dy = ty; dy < 32; dy += 8) {
    const long long s = s0 + dy;
    const int k = k0 + tx;
    if (k < m && s < S) dst[k + (long long)m * (s + S * t)] = tile[tx][dy];
  }
}

// ---- launchers --------------------------------------------------------------------------------------
static unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

int launch_simulate(const SimArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * kSimThreads * (2 * (size_t)A.m + (size_t)A.r * A.repeat_Q);
  SS_ARG(smem <= 227 * 1024, "SimulateC: state too large for shared memory");
  SS_CUDA(cudaFuncSetAttribute(simulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  simulate_kernel<<<nblk(A.S, kSimThreads), kSimThreads, smem, st>>>(A);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_fs_draws(const FsArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * kSimThreads * (5 * (size_t)A.m + (size_t)A.r);
  SS_ARG(smem <= 227 * 1024, "FastSmootherC: state too large for shared memory (m > 44)");
  SS_CUDA(cudaFuncSetAttribute(fs_draw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fs_draw_kernel<<<nblk(A.S, kSimThreads), kSimThreads, smem, st>>>(A);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_extract(const double* a, int a_native, const DMat& Z, int m, int p, int S, int N, double* out,
                   cudaStream_t st) {
  SS_ARG(N <= 65535, "ExtractComponentC: N too large");
  extract_kernel<<<dim3(nblk(S, kSimThreads), N), kSimThreads, 0, st>>>(a, a_native, Z, m, p, S, N, out);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_permute_tks(const double* src, double* dst, int N, int K, long long S, bool to_r, cudaStream_t st) {
  SS_ARG(K <= 65535 && nblk(N, 32) <= 65535, "permute: dimension too large");
  dim3 grid(nblk(S, 32), nblk(N, 32), K), blk(32, 8);
  if (to_r)
    permute_tks_kernel<true><<<grid, blk, 0, st>>>(src, dst, N, K, S);
  else
    permute_tks_kernel<false><<<grid, blk, 0, st>>>(src, dst, N, K, S);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_permute_at(const double* src, double* dst, int N, int m, long long S, cudaStream_t st) {
  SS_ARG(N <= 65535, "permute: N too large");
  permute_at_kernel<<<dim3(nblk(S, 32), nblk(m, 32), N), dim3(32, 8), 0, st>>>(src, dst, m, S);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu
