// Generic forward filter: univariate-treatment Kalman filter with exact diffuse initialisation,
// one CTA per evaluation, any (m <= 64, p, r), time-varying or per-evaluation system matrices.
//
// Restates the recursion of src/LogLikC.cpp:68-208 (MODE_LOGLIK), of the forward loop of
// src/KalmanC.cpp:122-409 (MODE_KALMAN, which additionally stores every intermediate the backward
// pass and the R caller need) and the y-independent part of src/FastSmootherC.cpp:103-252
// (MODE_GAINS: regime chosen by index < initialisation_steps, gains written once for all draws).
// The arithmetic order is the pinned one of oracle/kalman_c.c (explicit fma chains, k ascending).  The covariance lives in shared memory; work inside a CTA is
// element-parallel (thread e owns elements e, e+NT, ... of every m x m object), all scalar decisions
// (F_inf / F_star thresholds, P_inf ~ 0 test) are block-uniform.
//
// This is the "any shape" kernel: the throughput kernel for the batched headline workload is
// loglik_cols.cu.  Rank-1 forms are used for the measurement update:
//   P* L0' = P* - (P* z) K0'            (src/LogLikC.cpp:129,193)
//   P_inf L1' + P* L0' = P* - M* K0' - M_inf K1'   (:151),  P_inf L0' = P_inf - M_inf K0' (:152)
// which equal the reference's GEMM forms exactly in real arithmetic and to O(eps) in FP64.
#include <cstdlib>

#include "common.cuh"
#include "mm_dmma.cuh"

namespace ssgpu {

struct FwdArgs {
  DModel md;
  const double* y;  // [(t*p + j) * B + b]
  long long B;
  int V;
  double* out;  // [E] loglik / N (LOGLIK mode)
  ssgpu_kalman_out k;  // device pointers (KALMAN mode)
  KalmanScratch s;
  int init_steps;  // MODE_GAINS: src/FastSmootherC.cpp:147
};

constexpr int MODE_LOGLIK = 0, MODE_KALMAN = 1, MODE_GAINS = 2;

__device__ __forceinline__ void load_mat(double* dst, const double* blk, int diag, int rows, int cols, int tid,
                                         int NT) {
  const int n = rows * cols;
  for (int e = tid; e < n; e += NT) {
    if (diag) {
      const int i = e % rows, j = e / rows;
      dst[e] = i == j ? blk[i] : 0.0;
    } else {
      dst[e] = blk[e];
    }
  }
}

// C (m x m) = A (m x m) * B (m x m), all column-major in shared memory: DMMA tiles (mm_dmma.cuh), same bits as
// the scalar loop  acc = fma(A[i + k m], B[k + j m], acc), k ascending
__device__ __forceinline__ void mm_nn(double* C, const double* A, const double* Bm, int m, int tid, int NT) {
  mm_dmma<false, false>(C, A, Bm, nullptr, m, tid, NT);
}
// C = add + A * B'   (add may be nullptr)
__device__ __forceinline__ void mm_nt(double* C, const double* A, const double* Bm, const double* add, int m, int tid,
                                      int NT) {
  mm_dmma<false, true>(C, A, Bm, add, m, tid, NT);
}

template <int MODE, bool SHIFT = false>
__global__ void __launch_bounds__(512) fwd_generic_kernel(const FwdArgs A_) {
  constexpr bool KALMAN = MODE == MODE_KALMAN, GAINS = MODE == MODE_GAINS;
  extern __shared__ double sm[];
  const long long e = blockIdx.x;
  const long long b = e % A_.B, v = e / A_.B;
  // batched KalmanC (SHIFT): this CTA's series owns the b-th block of every output and scratch array.  The shifted
  // argument block lives in shared memory: as a thread-private copy ptxas rematerialised every pointer (base +
  // b * block) at each use, 10 % of the instructions of an issue-bound kernel.  Single-series calls read the
  // parameter bank itself.
  __shared__ __align__(16) unsigned char shifted_args[SHIFT ? sizeof(FwdArgs) : 16];
  if constexpr (SHIFT) {
    if (threadIdx.x == 0) {
      FwdArgs* S = reinterpret_cast<FwdArgs*>(shifted_args);
      *S = A_;
      kalman_shift(S->k, S->s, b, A_.md.N, A_.md.p, A_.md.m, A_.md.r, A_.md.Q.tv() || A_.md.R.tv());
    }
    __syncthreads();
  }
  const FwdArgs& A = SHIFT ? *reinterpret_cast<const FwdArgs*>(shifted_args) : A_;
  const DModel& md = A.md;
  const int m = md.m, p = md.p, r = md.r, N = md.N;
  const int tid = threadIdx.x, NT = blockDim.x;
  const int mm = m * m;
  int wsz = mm;
  if (m * r > wsz) wsz = m * r;
  if (p * m > wsz) wsz = p * m;
  double* Ps = sm;
  double* Pi = Ps + mm;
  double* W = Pi + mm;
  double* Ts = W + wsz;
  double* RQR = Ts + mm;
  double* Zs = RQR + mm;
  double* av = Zs + p * m;
  double* av2 = av + m;
  double* Ms = av2 + m;
  double* Mi = Ms + m;
  double* k0v = Mi + m;
  double* k1v = k0v + m;
  double* scal = k1v + m;  // F_star, F_inf, z'a of the current observation

  if (!GAINS) load_mat(av, md.a.block(b, v, 0), 0, m, 1, tid, NT);
  load_mat(Pi, md.P_inf.block(b, v, 0), md.P_inf.diag, m, m, tid, NT);
  load_mat(Ps, md.P_star.block(b, v, 0), md.P_star.diag, m, m, tid, NT);
  load_mat(Zs, md.Z.block(b, v, 0), 0, p, m, tid, NT);
  load_mat(Ts, md.T.block(b, v, 0), md.T.diag, m, m, tid, NT);
  __syncthreads();

  // initialisation = !all(|P_inf| < 1e-7)   (src/LogLikC.cpp:49)
  auto pinf_small = [&]() {
    int ok = 1;
    for (int q = tid; q < mm; q += NT) ok &= (fabs(Pi[q]) < kEps) ? 1 : 0;
    return __syncthreads_and(ok);
  };
  bool init = GAINS ? true : !pinf_small();

  // Q_t R_t' (r x m, into W) then R_t (Q_t R_t')  -> RQR   (src/KalmanC.cpp:108,385)
  auto build_rqr = [&](int t) {
    const double* Qb = md.Q.block(b, v, md.Q.tv() ? t : 0);
    const double* Rb = md.R.block(b, v, md.R.tv() ? t : 0);
    for (int q = tid; q < r * m; q += NT) {
      const int k = q % r, j = q / r;
      double acc = 0.0;
      if (md.Q.diag) {
        acc = Qb[k] * mat_get(Rb, md.R.diag, m, j, k);
      } else {
        for (int l = 0; l < r; l++) acc = fma(Qb[k + l * r], mat_get(Rb, md.R.diag, m, j, l), acc);
      }
      W[q] = acc;
      if (KALMAN || GAINS) A.s.QtR[(long long)((md.Q.tv() || md.R.tv()) ? t : 0) * r * m + q] = acc;
    }
    __syncthreads();
    for (int q = tid; q < mm; q += NT) {
      const int i = q % m, j = q / m;
      double acc = 0.0;
      if (md.R.diag) {
        acc = i < r ? Rb[i] * W[i + j * r] : 0.0;
      } else {
        for (int k = 0; k < r; k++) acc = fma(Rb[i + (long long)k * m], W[k + j * r], acc);
      }
      RQR[q] = acc;
    }
    __syncthreads();
  };
  build_rqr(0);

  double loglik = 0.0;
  int non_init = 0, F_eq0 = 0, init_steps = 0;
  const bool rq_tv = md.Q.tv() || md.R.tv();

  for (int t = 0; t < N; t++) {
    if (t > 0) {
      if (md.Z.tv()) load_mat(Zs, md.Z.block(b, v, t), 0, p, m, tid, NT);
      if (md.T.tv()) load_mat(Ts, md.T.block(b, v, t), md.T.diag, m, m, tid, NT);
      if (md.Z.tv() || md.T.tv()) __syncthreads();
      if (rq_tv) build_rqr(t);
    }

    if (KALMAN) {
      // predicted quantities (src/KalmanC.cpp:145-162)
      const ssgpu_kalman_out& K = A.k;
      for (int q = tid; q < m; q += NT)
        if (K.a_pred) K.a_pred[t + (long long)N * q] = av[q];
      for (int q = tid; q < mm; q += NT) {
        const double ps = Ps[q], pi = init ? Pi[q] : 0.0;
        const long long o = (long long)t * mm + q;
        if (K.P_star_pred) K.P_star_pred[o] = ps;
        if (K.P_inf_pred) K.P_inf_pred[o] = pi;
        if (K.P_pred) K.P_pred[o] = init ? ps + kKappa * pi : ps;
      }
      // yfit = Z a_pred ; v = y - yfit
      for (int j = tid; j < p; j += NT) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Zs[j + k * p], av[k], acc);
        if (K.yfit) K.yfit[t + (long long)N * j] = acc;
        if (K.v) K.v[t + (long long)N * j] = A.y[((long long)t * p + j) * A.B + b] - acc;
      }
      // Fmat = Z P_pred Z'   via ZP (p x m) in W
      if (K.Fmat) {
        for (int q = tid; q < p * m; q += NT) {
          const int i = q % p, l = q / p;
          double acc = 0.0;
          for (int k = 0; k < m; k++) {
            const double pp = init ? Ps[k + l * m] + kKappa * Pi[k + l * m] : Ps[k + l * m];
            acc = fma(Zs[i + k * p], pp, acc);
          }
          W[q] = acc;
        }
        __syncthreads();
        for (int q = tid; q < p * p; q += NT) {
          const int i = q % p, j = q / p;
          double acc = 0.0;
          for (int l = 0; l < m; l++) acc = fma(W[i + l * p], Zs[j + l * p], acc);
          K.Fmat[(long long)t * p * p + q] = acc;
        }
        __syncthreads();
      }
    }

    for (int j = 0; j < p; j++) {
      const int index = t * p + j;
      const double yv = GAINS ? 0.0 : A.y[((long long)t * p + j) * A.B + b];
      if (GAINS) init = index < A.init_steps;
      if (!GAINS && isnan(yv)) {  // src/LogLikC.cpp:88-90, src/KalmanC.cpp:193-211
        if (KALMAN) {
          if (init) init_steps++;
          if (tid == 0) {
            if (A.k.loglik) A.k.loglik[index] = nan("");
            A.s.code[index] = init ? 256 : 0;
          }
        }
        continue;
      }
      // M = P z (M* and, while diffuse, M_inf): DMMA chains, one 8-row tile per warp (mm_dmma.cuh)
      if (!init)
        for (int i = tid; i < m; i += NT) Mi[i] = 0.0;
      mv_dmma2(Ms, Ps, Mi, init ? Pi : nullptr, Zs + j, p, m, tid, NT);
      __syncthreads();
      // F = z'M and z'a: one warp runs the three chains (k ascending) and publishes them -- every warp doing so
      // for itself is 4 m broadcast loads per warp and makes the shared-memory pipe the bottleneck of the kernel
      if (tid < 32) {
        double fs = 0.0, fi = 0.0, zz = 0.0;
        for (int k = 0; k < m; k++) {
          const double z = Zs[j + k * p];
          fs = fma(z, Ms[k], fs);
          fi = fma(z, Mi[k], fi);
          if (!GAINS) zz = fma(z, av[k], zz);
        }
        if (tid == 0) {
          scal[0] = fs;
          scal[1] = fi;
          scal[2] = zz;
        }
      }
      __syncthreads();
      const double F_star = scal[0], F_inf = scal[1], za = scal[2];
      int code = 0;
      double F1 = 0.0, F2 = 0.0, vv = 0.0, ll = nan("");
      if (init) {
        if (KALMAN) init_steps++;
        if (F_inf < kEps) {
          if (!(F_star < kEps)) code = 1;  // NaN takes the update branch, like the reference's else
        } else {
          code = 2;
        }
      } else {
        non_init++;
        if (F_star < kEps)
          F_eq0++;
        else
          code = 1;
      }
      // the likelihood term is thread 0's alone (it owns loglik and the stores): the logarithm is ~40 instructions
      // that every other warp of an issue-bound kernel can skip
      if (code == 1) {
        F1 = 1.0 / F_star;
        vv = yv - za;
        if (tid == 0) {
          ll = kLogConst - (log(F_star) + vv * vv * F1) / 2;
          loglik += ll;
        }
      } else if (code == 2) {
        F1 = 1.0 / F_inf;
        F2 = -(F1 * F1) * F_star;
        vv = yv - za;
        if (tid == 0) {
          ll = kLogConst - log(F_inf) / 2;
          loglik += ll;
        }
      }
      if ((KALMAN || GAINS) && tid == 0) {
        if (KALMAN && A.k.loglik) A.k.loglik[index] = ll;
        A.s.code[index] = code | (init ? 256 : 0);
        A.s.F1[index] = F1;
        if (KALMAN) {
          A.s.F2[index] = F2;
          A.s.v[index] = vv;
        }
      }
      // K0 = M F1 (code 1: M*, code 2: M_inf); K1 = M* F1 + M_inf F2 (code 2 only)
      for (int i = tid; i < m; i += NT) {
        const double k0 = code == 0 ? 0.0 : (code == 1 ? Ms[i] : Mi[i]) * F1;
        const double k1 = code == 2 ? Ms[i] * F1 + Mi[i] * F2 : 0.0;
        k0v[i] = k0;
        k1v[i] = k1;
        if (KALMAN || GAINS) {
          A.s.K0[(long long)index * m + i] = k0;
          A.s.K1[(long long)index * m + i] = k1;
        }
      }
      __syncthreads();
      if (code == 0) continue;  // block-uniform
      for (int q = tid; q < mm; q += NT) {
        const int c = div_small(q, m), i = q - c * m;
        if (code == 1) {
          Ps[q] = Ps[q] - Ms[i] * k0v[c];
        } else {
          Ps[q] = (Ps[q] - Ms[i] * k0v[c]) - Mi[i] * k1v[c];
          Pi[q] = Pi[q] - Mi[i] * k0v[c];
        }
      }
      if (!GAINS)
        for (int i = tid; i < m; i += NT) av[i] = fma(k0v[i], vv, av[i]);
      __syncthreads();
      if (!GAINS && code == 2 && j < p - 1) init = !pinf_small();  // src/LogLikC.cpp:157-159
    }

    if (KALMAN) {
      // filtered quantities (src/KalmanC.cpp:375-382,394-400)
      const ssgpu_kalman_out& K = A.k;
      for (int q = tid; q < m; q += NT)
        if (K.a_fil) K.a_fil[t + (long long)N * q] = av[q];
      for (int q = tid; q < mm; q += NT) {
        const double ps = Ps[q], pi = init ? Pi[q] : 0.0;
        const long long o = (long long)t * mm + q;
        if (K.P_star_fil) K.P_star_fil[o] = ps;
        if (K.P_inf_fil) K.P_inf_fil[o] = pi;
        if (K.P_fil) K.P_fil[o] = init ? ps + kKappa * pi : ps;
      }
    }

    // transition to the next time point (src/LogLikC.cpp:200-207; KalmanC also forecasts after t = N-1)
    if (t < N - 1 || KALMAN) {
      if (!GAINS)
        for (int i = tid; i < m; i += NT) {
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(Ts[i + k * m], av[k], acc);
          av2[i] = acc;
        }
      mm_nn(W, Ts, Ps, m, tid, NT);
      __syncthreads();
      mm_nt(Ps, W, Ts, RQR, m, tid, NT);
      if (!GAINS)
        for (int i = tid; i < m; i += NT) av[i] = av2[i];
      __syncthreads();
      if (GAINS) init = (t * p + p - 1) < A.init_steps;  // src/FastSmootherC.cpp:248
      if (init) {
        mm_nn(W, Ts, Pi, m, tid, NT);
        __syncthreads();
        mm_nt(Pi, W, Ts, nullptr, m, tid, NT);
        __syncthreads();
        if (!GAINS && t < N - 1) init = !pinf_small();
      }
      if (KALMAN && t == N - 1) {
        const ssgpu_kalman_out& K = A.k;
        for (int q = tid; q < m; q += NT)
          if (K.a_fc) K.a_fc[q] = av[q];
        for (int q = tid; q < mm; q += NT) {
          const double ps = Ps[q], pi = init ? Pi[q] : 0.0;
          if (K.P_star_fc) K.P_star_fc[q] = ps;
          if (K.P_inf_fc) K.P_inf_fc[q] = pi;
          if (K.P_fc) K.P_fc[q] = init ? ps + kKappa * pi : ps;
        }
      }
    }
  }

  if (tid == 0) {
    if (KALMAN) {
      if (A.k.initialisation_steps) *A.k.initialisation_steps = init_steps;
    } else if (!GAINS) {
      A.out[e] = (non_init == F_eq0) ? nan("") : loglik / N;  // src/LogLikC.cpp:211-214
    }
  }
}

static size_t fwd_smem_bytes(const DModel& md) {
  const int m = md.m, p = md.p, r = md.r;
  size_t wsz = (size_t)m * m;
  if ((size_t)m * r > wsz) wsz = (size_t)m * r;
  if ((size_t)p * m > wsz) wsz = (size_t)p * m;
  return sizeof(double) * (4 * (size_t)m * m + wsz + (size_t)p * m + 6 * (size_t)m + 8);
}

// Threads per CTA.  One series: as many as there are matrix elements (latency).  A batch: a quarter of that, capped
// at 256 -- barriers and the per-warp bookkeeping grow with the warp count while the SM is filled by other CTAs
// anyway (measured at m = 34: 64 / 128 / 256 / 512 threads -> 2.3 / 3.1 / 3.4 / 2.5 e7 series*timesteps/s; at
// m = 14: 5.7 / 5.1 / 3.2 / 3.2 e7).
static int pick_threads(int m, bool batched = false) {
  static const int forced = [] {
    const char* e = getenv("SSGPU_GENERIC_NT");
    return e ? atoi(e) : 0;
  }();
  int nt = batched ? ((m * m / 4 + 31) / 32) * 32 : ((m * m + 31) / 32) * 32;
  const int lo = batched ? 64 : 32, hi = batched ? 256 : 512;
  if (nt < lo) nt = lo;
  if (nt > hi) nt = hi;
  if (forced > 0) nt = forced;
  return nt;
}

int launch_fwd_generic(const DModel& md, const double* y, long long B, int V, double* out_loglik,
                       const ssgpu_kalman_out* kout, const KalmanScratch* scratch, cudaStream_t st) {
  FwdArgs A{};
  A.md = md;
  A.y = y;
  A.B = B;
  A.V = V;
  A.out = out_loglik;
  const size_t smem = fwd_smem_bytes(md);
  SS_ARG(smem <= 227 * 1024, "generic filter: model does not fit in shared memory (m, p or r too large)");
  const long long E = B * V;
  SS_ARG(E > 0 && E < (1LL << 31), "generic filter: batch size out of range");
  const int nt = pick_threads(md.m, E > 1);
  if (kout) {
    SS_ARG(V == 1 && scratch, "KalmanC runs one parameter vector per series");
    A.k = *kout;
    A.s = *scratch;
    if (B > 1) {
      SS_CUDA(cudaFuncSetAttribute(fwd_generic_kernel<MODE_KALMAN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
      fwd_generic_kernel<MODE_KALMAN, true><<<(unsigned)B, nt, smem, st>>>(A);
    } else {
      SS_CUDA(cudaFuncSetAttribute(fwd_generic_kernel<MODE_KALMAN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
      fwd_generic_kernel<MODE_KALMAN, false><<<1, nt, smem, st>>>(A);
    }
    set_last_kernel("fwd_generic_kernel<KALMAN>");
  } else {
    SS_CUDA(cudaFuncSetAttribute(fwd_generic_kernel<MODE_LOGLIK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    fwd_generic_kernel<MODE_LOGLIK><<<(unsigned)E, nt, smem, st>>>(A);
    set_last_kernel("fwd_generic_kernel<LOGLIK>");
  }
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

// FastSmootherC gain sequence: code / F1 / K0 / K1 / QtR of `scratch` (F2, v unused)
int launch_gains_generic(const DModel& md, int init_steps, const KalmanScratch* scratch, cudaStream_t st) {
  FwdArgs A{};
  A.md = md;
  A.B = 1;
  A.V = 1;
  A.s = *scratch;
  A.init_steps = init_steps;
  const size_t smem = fwd_smem_bytes(md);
  SS_ARG(smem <= 227 * 1024, "generic filter: model does not fit in shared memory (m, p or r too large)");
  SS_CUDA(cudaFuncSetAttribute(fwd_generic_kernel<MODE_GAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fwd_generic_kernel<MODE_GAINS><<<1, pick_threads(md.m), smem, st>>>(A);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

}  // namespace ssgpu
