// Backward pass of KalmanC: state / disturbance smoother with the exact diffuse recursions
// (r0, r1, N0, N1, N2), restating src/KalmanC.cpp:416-560 for one series in one CTA.
//
// The reference stores L0 = I - K0 z and L1 = -K1 z as dense m x m slices and multiplies with
// them; here only K0, K1 (from the forward pass scratch) are kept and every product with L0 / L1 is
// applied as a rank-1 correction:
//   X L0 = X - (X K0) z'        L0' X = X - z (K0' X)        X L1 = -(X K1) z'     L1' X = -z (K1' X)
// which is the same arithmetic in O(m^2) instead of O(m^3).  The reference's quirks are reproduced
// on purpose: N_1 = N_1 L0 only (no r_1 / N_2 update) in the
// F_inf < eps <= F_star branch (:464-468); N_2 uses the already updated N_1 (:483-487); eta_var uses
// the Q slice left over from the forward loop (:548); the regime test is index < initialisation_steps.
#include <cstdlib>

#include "common.cuh"
#include "mm_dmma.cuh"

namespace ssgpu {

struct BwdArgs {
  DModel md;
  ssgpu_kalman_out k;  // device pointers
  KalmanScratch s;
};

__device__ __forceinline__ double dotv(const double* a, const double* b, int m) {
  double acc = 0.0;
  for (int k = 0; k < m; k++) acc = fma(a[k], b[k], acc);
  return acc;
}

// X <- L0' X L0  (in place), L0 = I - k0 z'.  u, w: m-vectors of scratch.
__device__ void sandwich_L0(double* X, const double* k0, const double* z, double* u, double* w, int m, int tid, int NT) {
  for (int i = tid; i < m; i += NT) {
    double acc = 0.0;
    for (int k = 0; k < m; k++) acc = fma(X[i + k * m], k0[k], acc);
    u[i] = acc;
  }
  __syncthreads();
  for (int q = tid; q < m * m; q += NT) {
    const int c = div_small(q, m);
    X[q] -= u[q - c * m] * z[c];
  }
  __syncthreads();
  for (int j = tid; j < m; j += NT) {
    double acc = 0.0;
    for (int i = 0; i < m; i++) acc = fma(k0[i], X[i + j * m], acc);
    w[j] = acc;
  }
  __syncthreads();
  for (int q = tid; q < m * m; q += NT) {
    const int c = div_small(q, m);
    X[q] -= z[q - c * m] * w[c];
  }
  __syncthreads();
}

// X <- T' X T (in place) using tmp; DMMA tiles (mm_dmma.cuh), same bits as the scalar k-ascending fma loops
__device__ void sandwich_T(double* X, const double* Ts, double* tmp, int m, int tid, int NT) {
  mm_dmma<false, false>(tmp, X, Ts, nullptr, m, tid, NT);  // tmp = X T
  __syncthreads();
  mm_dmma<true, false>(X, Ts, tmp, nullptr, m, tid, NT);  // X = T' tmp
  __syncthreads();
}

template <bool SHIFT>
struct BwdView;
template <>
struct BwdView<false> {
  const BwdArgs& A;
  __device__ BwdView(const BwdArgs& a, long long) : A(a) {}
};
template <>
struct BwdView<true> {
  BwdArgs A;
  __device__ BwdView(const BwdArgs& a, long long b) : A(a) {
    kalman_shift(A.k, A.s, b, A.md.N, A.md.p, A.md.m, A.md.r, A.md.Q.tv() || A.md.R.tv());
  }
};

template <bool SHIFT>
__global__ void __launch_bounds__(512) bwd_generic_kernel(const BwdArgs A_) {
  extern __shared__ double sm[];
  const long long b = blockIdx.x;  // batched KalmanC: one CTA per series, the b-th block of every array
  const BwdView<SHIFT> view(A_, b);
  const BwdArgs& A = view.A;
  const DModel& md = A.md;
  const int m = md.m, p = md.p, r = md.r, N = md.N;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int mm = m * m;
  double* Nn = sm;
  double* N1 = Nn + mm;
  double* N2 = N1 + mm;
  double* Ts = N2 + mm;
  double* tA = Ts + mm;
  double* tB = tA + mm;
  double* Zs = tB + mm;       // p*m
  double* vec = Zs + p * m;   // 12 m-vectors
  double* rv = vec;
  double* r1 = rv + m;
  double* k0 = r1 + m;
  double* k1 = k0 + m;
  double* u = k1 + m;
  double* w = u + m;
  double* h = w + m;
  double* s = h + m;
  double* g = s + m;
  double* qv = g + m;
  double* q2 = qv + m;
  double* zr = q2 + m;
  const ssgpu_kalman_out& K = A.k;
  const int init_steps = *K.initialisation_steps;
  const bool qtr_tv = md.Q.tv() || md.R.tv();

  for (int q = tid; q < mm; q += NT) {
    Nn[q] = 0.0;
    N1[q] = 0.0;
    N2[q] = 0.0;
    Ts[q] = mat_get(md.T.block(b, 0, 0), md.T.diag, m, q % m, q / m);
  }
  for (int q = tid; q < m; q += NT) {
    rv[q] = 0.0;
    r1[q] = 0.0;
  }
  for (int q = tid; q < p * m; q += NT) Zs[q] = md.Z.block(b, 0, md.Z.tv() ? N - 1 : 0)[q];
  // r_vec row N and Nmat slice N are zero (src/KalmanC.cpp:98-101); eta row N-1 = QtR * 0
  for (int q = tid; q < m; q += NT)
    if (K.r_vec) K.r_vec[N + (long long)(N + 1) * q] = 0.0;
  for (int q = tid; q < mm; q += NT)
    if (K.Nmat) K.Nmat[(long long)N * mm + q] = 0.0;
  const double* Qlast = md.Q.block(b, 0, md.Q.tv() ? N - 1 : 0);
  for (int q = tid; q < r; q += NT)
    if (K.eta) K.eta[(N - 1) + (long long)N * q] = 0.0;
  for (int q = tid; q < r * r; q += NT)
    if (K.eta_var) K.eta_var[(long long)(N - 1) * r * r + q] = mat_get(Qlast, md.Q.diag, r, q % r, q / r);
  __syncthreads();

  for (int t = N - 1; t >= 0; t--) {
    if (md.Z.tv() && t < N - 1) {
      for (int q = tid; q < p * m; q += NT) Zs[q] = md.Z.block(b, 0, t)[q];
      __syncthreads();
    }
    for (int j = p - 1; j >= 0; j--) {
      const int index = t * p + j;
      const int code = A.s.code[index] & 255;
      if (code == 0) continue;  // missing or no information: r, N carried over (:438-441,458-460,494-496)
      const bool diffuse = index < init_steps;
      const double F1 = A.s.F1[index], F2 = A.s.F2[index], vv = A.s.v[index];
      for (int q = tid; q < m; q += NT) {
        zr[q] = Zs[j + q * p];
        k0[q] = A.s.K0[(long long)index * m + q];
        k1[q] = A.s.K1[(long long)index * m + q];
      }
      __syncthreads();
      if (code == 1) {
        // r = z F1 v + r L0 ; N = z z' F1 + L0' N L0   (:464-467, :500-503)
        const double rk = dotv(rv, k0, m);
        __syncthreads();
        for (int q = tid; q < m; q += NT) rv[q] = zr[q] * F1 * vv + (rv[q] - rk * zr[q]);
        sandwich_L0(Nn, k0, zr, u, w, m, tid, NT);
        for (int q = tid; q < mm; q += NT) {
          const int c = div_small(q, m);
          Nn[q] = zr[q - c * m] * zr[c] * F1 + Nn[q];
        }
        if (diffuse) {  // N_1 = N_1 L0   (:468)
          for (int i = tid; i < m; i += NT) {
            double acc = 0.0;
            for (int k = 0; k < m; k++) acc = fma(N1[i + k * m], k0[k], acc);
            u[i] = acc;
          }
          __syncthreads();
          for (int q = tid; q < mm; q += NT) N1[q] -= u[q % m] * zr[q / m];
        }
        __syncthreads();
      } else {
        // diffuse step with F_inf >= eps (:472-487)
        const double rk0 = dotv(rv, k0, m), rk1 = dotv(rv, k1, m), r1k0 = dotv(r1, k0, m);
        // vectors from the OLD N: h' = K1' N, s = N K1
        for (int i = tid; i < m; i += NT) {
          double a1 = 0.0, a2 = 0.0;
          for (int k = 0; k < m; k++) {
            a1 = fma(k1[k], Nn[k + i * m], a1);
            a2 = fma(Nn[i + k * m], k1[k], a2);
          }
          h[i] = a1;
          s[i] = a2;
        }
        __syncthreads();
        const double hk0 = dotv(h, k0, m), sk0 = dotv(s, k0, m), c = dotv(k1, s, m);
        __syncthreads();
        for (int q = tid; q < m; q += NT) {
          const double rold = rv[q];
          rv[q] = rold - rk0 * zr[q];                                              // r L0
          r1[q] = zr[q] * F1 * vv + (r1[q] - r1k0 * zr[q]) + (-rk1 * zr[q]);       // z F1 v + r1 L0 + r L1
          g[q] = h[q] - hk0 * zr[q];                                               // (K1' N) L0
          qv[q] = s[q] - zr[q] * sk0;                                              // L0' (N K1)
        }
        __syncthreads();
        sandwich_L0(Nn, k0, zr, u, w, m, tid, NT);   // N = L0' N L0
        sandwich_L0(N1, k0, zr, u, w, m, tid, NT);   // L0' N_1 L0
        for (int q = tid; q < mm; q += NT) {
          const int i = q % m, jj = q / m;
          // + z z' F1 + L1' N L0 (= -z g') + L0' N L1 (= -q z')
          N1[q] = zr[i] * zr[jj] * F1 + N1[q] + (-zr[i] * g[jj]) + (-qv[i] * zr[jj]);
        }
        __syncthreads();
        // s2 = N_1(new) K1 ; q2 = L0' s2
        for (int i = tid; i < m; i += NT) {
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(N1[i + k * m], k1[k], acc);
          u[i] = acc;
        }
        __syncthreads();
        const double s2k0 = dotv(u, k0, m);
        __syncthreads();
        for (int q = tid; q < m; q += NT) q2[q] = u[q] - zr[q] * s2k0;
        __syncthreads();
        sandwich_L0(N2, k0, zr, u, w, m, tid, NT);   // L0' N_2 L0
        for (int q = tid; q < mm; q += NT) {
          const int i = q % m, jj = q / m;
          // z z' F2 + L0' N_2 L0 + L0' N_1 L1 (= -q2 z') + L1' N_1' L0 (= -z q2') + L1' N L1 (= c z z')
          N2[q] = zr[i] * zr[jj] * F2 + N2[q] + (-q2[i] * zr[jj]) + (-zr[i] * q2[jj]) + c * zr[i] * zr[jj];
        }
        __syncthreads();
      }
    }

    // save r and N of this time point (:509-510)
    for (int q = tid; q < m; q += NT)
      if (K.r_vec) K.r_vec[t + (long long)(N + 1) * q] = rv[q];
    for (int q = tid; q < mm; q += NT)
      if (K.Nmat) K.Nmat[(long long)t * mm + q] = Nn[q];

    // smoothed state and variance (:513-526)
    const bool diffuse_t = t * p < init_steps;
    const double* Pst = K.P_star_pred + (long long)t * mm;
    const double* Pin = K.P_inf_pred + (long long)t * mm;
    for (int jj = tid; jj < m; jj += NT) {
      double acc = 0.0;
      for (int i = 0; i < m; i++) acc = fma(rv[i], Pst[i + jj * m], acc);
      if (diffuse_t) {
        double acc2 = 0.0;
        for (int i = 0; i < m; i++) acc2 = fma(r1[i], Pin[i + jj * m], acc2);
        acc += acc2;
      }
      if (K.a_smooth) K.a_smooth[t + (long long)N * jj] = K.a_pred[t + (long long)N * jj] + acc;
    }
    if (K.V && !diffuse_t) {
      // V = P* - P* (N P*): P*_pred of this time point staged once (the scalar loops below read it m-fold from
      // global memory), two DMMA products -- the same k-ascending fma chains, bit for bit
      for (int q = tid; q < mm; q += NT) tB[q] = Pst[q];
      __syncthreads();
      mm_dmma<false, false>(tA, Nn, tB, nullptr, m, tid, NT);
      __syncthreads();
      mm_dmma<false, false, true>(K.V + (long long)t * mm, tB, tA, tB, m, tid, NT);
      __syncthreads();
    } else if (K.V) {
      for (int q = tid; q < mm; q += NT) {
        const int jj = div_small(q, m), kk = q - jj * m;
        double a1 = 0.0, a2 = 0.0;
        for (int l = 0; l < m; l++) {
          a1 = fma(Nn[kk + l * m], Pst[l + jj * m], a1);
          if (diffuse_t) {
            a1 = fma(N1[l + kk * m], Pin[l + jj * m], a1);     // N_1' P_inf
            a2 = fma(N1[kk + l * m], Pst[l + jj * m], a2);     // N_1 P_star
            a2 = fma(N2[kk + l * m], Pin[l + jj * m], a2);     // N_2 P_inf
          }
        }
        tA[q] = a1;
        tB[q] = a2;
      }
      __syncthreads();
      for (int q = tid; q < mm; q += NT) {
        const int jj = div_small(q, m), i = q - jj * m;
        double acc = 0.0;
        for (int kk = 0; kk < m; kk++) {
          acc = fma(Pst[i + kk * m], tA[kk + jj * m], acc);
          if (diffuse_t) acc = fma(Pin[i + kk * m], tB[kk + jj * m], acc);
        }
        K.V[(long long)t * mm + q] = Pst[q] - acc;
      }
      __syncthreads();
    }

    // smoothed disturbance of the PREVIOUS time point: eta[t-1] = QtR_{t-1} r_t,
    // eta_var[t-1] = Q - QtR_{t-1} N_t QtR_{t-1}'   (:547-548 evaluated at i = t-1)
    if (t > 0) {
      const double* QtR = A.s.QtR + (long long)(qtr_tv ? t - 1 : 0) * r * m;  // r x m
      for (int q = tid; q < r; q += NT) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(QtR[q + k * r], rv[k], acc);
        if (K.eta) K.eta[(t - 1) + (long long)N * q] = acc;
      }
      if (K.eta_var) {
        for (int q = tid; q < r * m; q += NT) {  // tA (r x m) = QtR N
          const int i = q % r, jj = q / r;
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(QtR[i + k * r], Nn[k + jj * m], acc);
          tA[q] = acc;
        }
        __syncthreads();
        for (int q = tid; q < r * r; q += NT) {
          const int i = q % r, jj = q / r;
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(tA[i + k * r], QtR[jj + k * r], acc);
          K.eta_var[(long long)(t - 1) * r * r + q] = mat_get(Qlast, md.Q.diag, r, i, jj) - acc;
        }
        __syncthreads();
      }
      // r and N for the previous time point (:551-558), T of time t-1
      if (md.T.tv()) {
        for (int q = tid; q < mm; q += NT)
          Ts[q] = mat_get(md.T.block(b, 0, t - 1), md.T.diag, m, q % m, q / m);
        __syncthreads();
      }
      for (int jj = tid; jj < m; jj += NT) {
        double acc = 0.0, acc1 = 0.0;
        for (int k = 0; k < m; k++) {
          acc = fma(rv[k], Ts[k + jj * m], acc);
          acc1 = fma(r1[k], Ts[k + jj * m], acc1);
        }
        u[jj] = acc;
        w[jj] = acc1;
      }
      __syncthreads();
      for (int q = tid; q < m; q += NT) {
        rv[q] = u[q];
        if (diffuse_t) r1[q] = w[q];
      }
      __syncthreads();
      sandwich_T(Nn, Ts, tA, m, tid, NT);
      if (diffuse_t) {
        sandwich_T(N1, Ts, tA, m, tid, NT);
        sandwich_T(N2, Ts, tA, m, tid, NT);
      }
    }
  }
}

static size_t bwd_smem_bytes(const DModel& md) {
  const size_t m = md.m, p = md.p, r = md.r;
  size_t mm = m * m;
  if (r * m > mm) mm = r * m;
  return sizeof(double) * (6 * mm + p * m + 12 * m);
}

int launch_bwd_generic(const DModel& md, const ssgpu_kalman_out* kout, const KalmanScratch* scratch,
                       cudaStream_t st, long long B) {
  BwdArgs A{};
  A.md = md;
  A.k = *kout;
  A.s = *scratch;
  SS_ARG(md.r <= md.m, "KalmanC: r > m is not supported by the smoother kernel");
  const size_t smem = bwd_smem_bytes(md);
  SS_ARG(smem <= 227 * 1024, "KalmanC smoother: model does not fit in shared memory");
  int nt = ((md.m * md.m + 31) / 32) * 32;
  if (nt > 512) nt = 512;
  if (nt < 32) nt = 32;
  if (B > 1) {  // batched: more, smaller CTAs per SM (cheaper barriers, the batch supplies the parallelism)
    nt = ((md.m * md.m / 4 + 31) / 32) * 32;
    if (nt > 256) nt = 256;
    if (nt < 64) nt = 64;
    static const int forced = [] {
      const char* e = getenv("SSGPU_GENERIC_NT");
      return e ? atoi(e) : 0;
    }();
    if (forced > 0) nt = forced;
  }
  if (B > 1) {
    SS_CUDA(cudaFuncSetAttribute(bwd_generic_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_generic_kernel<true><<<(unsigned)B, nt, smem, st>>>(A);
  } else {
    SS_CUDA(cudaFuncSetAttribute(bwd_generic_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bwd_generic_kernel<false><<<1, nt, smem, st>>>(A);
  }
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu
