// Batched loglikelihood with the p residual states of the reference's state layout eliminated: ONE THREAD PER SERIES
// on the dense covariance of the remaining "core" states, held once (upper triangle) in registers.
//
// Every model the reference builds has the state [eps (p residuals) ; components] with Z = [I_p | Z_c],
// T = blkdiag(0_p, T_c), R Q R' = blkdiag(H, W_c), a = [0 ; a_c], P_inf = blkdiag(0, Pinf_c), P_star = blkdiag(H, Pstar_c)
// (R/GetSysMat.R:1388-1426, R/statespacer.R:776-807).  For a diagonal H the recursion of
// src/LogLikC.cpp:68-208 on that state reduces EXACTLY to the textbook form on the components alone:
//
//   before observation j of a time point is processed, row / column eps_j of P_star is H_jj e_j, of P_inf zero, and
//   a[eps_j] = 0 (T wipes the residual rows at every time update, and an update with z_i = [e_i ; zc_i], i != j,
//   changes row eps_j by multiples of M[eps_j] = P[eps_j, eps_i] + P[eps_j, c] zc_i = 0).  Hence
//   M_c = C zc_j,  F = H_jj + zc_j' C zc_j,  v = y_j - zc_j' a_c,  and the update of (a_c, C) needs nothing else; what the
//   update writes into the residual rows of the observations already processed is never read again.
//
// So a dynamic Nelson-Siegel model (BASELINE.json config 3: p = 8, m = 14) is a 6 x 6 problem with 8 scalar updates per
// time point: 21 + 6 doubles of state per series, all in registers, no exchange between lanes, no shared memory.
// The constants (T_c, Z_c, H, W_c) are the same for every thread at every instruction and come from the kernel
// parameters (constant bank / uniform registers).  The exact diffuse phase (src/LogLikC.cpp:98-159,203-206) runs in the
// same thread on a second triangle for P_inf (template DIFF; models without diffuse states get the leaner kernel).
//
// Updates in the symmetric rank-1 / rank-2 form (equal to the reference's P L_0' products to
// 1e-15):  P -= M M'/F;  diffuse: P_star -= F1 (Mi Ms' + Ms Mi') + F2 Mi Mi',  P_inf -= F1 Mi Mi'.
#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace ssgpu {

constexpr int kCoreMaxMC = 8;   // core states (m - p)
constexpr int kCoreMaxP = 16;   // observations per time point
constexpr int kCoreThreads = 128;
// observations are fetched this many univariate steps ahead of their use (a ring in registers).  A thread of a small
// model finishes a step in a few hundred cycles while a load from HBM takes longer than that: with one load in flight per
// thread the local level runs at 1.7 TB/s of observations whatever the arithmetic costs.  Large cores hide the load
// behind the ~100 FP64 instructions of an observation and have no registers to spare.
#ifndef SSGPU_CORE_AHEAD_SMALL
#define SSGPU_CORE_AHEAD_SMALL 4
#endif
#ifndef SSGPU_CORE_AHEAD_MID
#define SSGPU_CORE_AHEAD_MID 1
#endif
__host__ __device__ constexpr int core_ahead(int mc) { return mc <= 2 ? SSGPU_CORE_AHEAD_SMALL : mc <= 4 ? SSGPU_CORE_AHEAD_MID : 1; }
constexpr int kCoreMaxNS = kCoreMaxMC * (kCoreMaxMC + 1) / 2;

struct CoreArgs {
  const double* y;  // [(t * p + j) * B + b]
  long long B, E;
  int N, p;
  double* out;                      // [e], e = v * B + b
  double T[kCoreMaxMC * kCoreMaxMC];  // T_c row-major, row stride MC
  double W[kCoreMaxNS];             // R Q R' of the core, upper triangle
  double Z[kCoreMaxP][kCoreMaxMC];  // Z_c
  double H[kCoreMaxP];
  double a0[kCoreMaxMC], P0[kCoreMaxNS], Pi0[kCoreMaxNS];
  // deterministic component states taken out of the recursion (see the launcher): their share of z'a per observation,
  // [t * p + j], and of T a per time update, [t * MC + i]; null = none.  Device memory, the same for every series.
  const double* coff;
  const double* boff;
};

// position of (k, l) in the row-major upper triangle
__host__ __device__ constexpr int core_sym(int MC, int k, int l) {
  return k <= l ? k * MC - k * (k - 1) / 2 + (l - k) : l * MC - l * (l - 1) / 2 + (k - l);
}

// see loglik_thr.cu: constants read through an offset the optimiser cannot see stay in the constant bank instead of
// being hoisted into vector registers for the whole time loop
__device__ __forceinline__ int core_opaque0() {
  int o;
  asm volatile("mov.u32 %0, 0;" : "=r"(o));
  return o;
}

__device__ __forceinline__ double core_rsqrt(double x) {  // x finite, normal, >= 1e-7
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  return r;
}

// X <- T X T' (+ W): column j of X T' first (MC dot products), then the entries (i <= j, j) of T (X T')
template <int MC, bool ADDW>
__device__ __forceinline__ void core_time_update(double (&X)[MC * (MC + 1) / 2], const double* __restrict__ Tc,
                                                 const double* __restrict__ Wq) {
  constexpr int NS = MC * (MC + 1) / 2;
  double Y[NS];
#pragma unroll
  for (int j = 0; j < MC; j++) {
    double col[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
      double acc = X[core_sym(MC, k, 0)] * Tc[j * MC];
#pragma unroll
      for (int l = 1; l < MC; l++) acc = fma(X[core_sym(MC, k, l)], Tc[j * MC + l], acc);
      col[k] = acc;
    }
#pragma unroll
    for (int i = 0; i <= j; i++) {
      double acc = ADDW ? fma(Tc[i * MC], col[0], Wq[core_sym(MC, i, j)]) : Tc[i * MC] * col[0];
#pragma unroll
      for (int k = 1; k < MC; k++) acc = fma(Tc[i * MC + k], col[k], acc);
      Y[core_sym(MC, i, j)] = acc;
    }
  }
#pragma unroll
  for (int i = 0; i < NS; i++) X[i] = Y[i];
}

// CTAs per SM the register budget is held to (measured on config 3, 6 core states: 124 registers at 4 CTAs 1.22e10
// series*timesteps/s, left to ptxas 156 - 192 registers and 0.84 - 1.11e10)
#ifndef SSGPU_CORE_MINB_SMALL
#define SSGPU_CORE_MINB_SMALL 8
#endif
constexpr int core_min_blocks(int mc, bool diff) {
  return diff ? (mc <= 4 ? 4 : 2) : (mc <= 3 ? SSGPU_CORE_MINB_SMALL : mc <= 6 ? 4 : 2);
}

// OFF: deterministic component states have been taken out (A.coff / A.boff, see the launcher)
// P1: one observation per time point (p = 1), known at compile time: no loop over the observations, the one row of Z_c
// and H are loaded once
template <int MC, bool DIFF, bool OFF, bool P1>
__global__ void __launch_bounds__(kCoreThreads, core_min_blocks(MC, DIFF)) loglik_core_kernel(const CoreArgs A) {
  constexpr int NS = MC * (MC + 1) / 2;
  const long long w = (long long)blockIdx.x * kCoreThreads + threadIdx.x;
  if (w >= A.E) return;
  const long long b = w % A.B;
  const int N = A.N, p = P1 ? 1 : A.p;
  double C[NS], a[MC];
  double Ci[DIFF ? NS : 1];
#pragma unroll
  for (int i = 0; i < NS; i++) {
    C[i] = A.P0[i];
    if constexpr (DIFF) Ci[i] = A.Pi0[i];
  }
#pragma unroll
  for (int k = 0; k < MC; k++) a[k] = A.a0[k];
  double prod = 1.0, sumq = 0.0;
  int esum = 0, n_terms = 0, non_init = 0, F_eq0 = 0;

  // log F is summed as the logarithm of a running product: mantissa in prod, exponent in esum
  auto acc_log = [&](double F) {
    prod *= F;
    const int hi = __double2hiint(prod);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    esum += ex;
    prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
  };

  constexpr int kCoreAhead = core_ahead(MC);
  const long long Np = (long long)N * p;
  long long q = 0;
  const double* yp = A.y + b;
  double yq[kCoreAhead];
#pragma unroll
  for (int i = 0; i < kCoreAhead; i++) yq[i] = i < Np ? yp[(long long)i * A.B] : 0.0;
  yp += (long long)kCoreAhead * A.B;
  auto take_y = [&]() {
    const double yv = yq[0];
#pragma unroll
    for (int i = 0; i + 1 < kCoreAhead; i++) yq[i] = yq[i + 1];
    if (q + kCoreAhead < Np) yq[kCoreAhead - 1] = *yp;
    yp += A.B;
    q++;
    return yv;
  };

  // regular update with observation j (src/LogLikC.cpp:162-195); missing or F < 1e-7: a no-op through 1/sqrt(F) := 0
  int t = 0;  // time point (the lambdas below index the offset tables with it)
  // the row of Z_c and H_jj of the coming observation are read one observation ahead (an indexed read of the
  // constant bank right in front of the dot products that need it costs its latency 8 times per time point)
  double zc[MC], hc;
  auto load_z = [&](int j) {
    const double* const zz = &A.Z[0][0] + j * kCoreMaxMC + core_opaque0();
#pragma unroll
    for (int k = 0; k < MC; k++) zc[k] = zz[k];
    hc = (&A.H[0] + core_opaque0())[j];
  };
  load_z(0);
  auto regular_obs = [&](int j, double yv) {
    double z[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) z[k] = zc[k];
    const double hj = hc;
    if constexpr (!P1) load_z(j + 1 < p ? j + 1 : 0);
    double M[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
      double acc = C[core_sym(MC, k, 0)] * z[0];
#pragma unroll
      for (int l = 1; l < MC; l++) acc = fma(C[core_sym(MC, k, l)], z[l], acc);
      M[k] = acc;
    }
    double F = hj, za = 0.0;
#pragma unroll
    for (int k = 0; k < MC; k++) {
      F = fma(z[k], M[k], F);
      za = fma(z[k], a[k], za);
    }
    const bool obs = !isnan(yv);  // :88-90
    const bool upd = obs && !(F < kEps);  // :173
    double r = core_rsqrt(upd ? F : 1.0);
    r = upd ? r : 0.0;
    double off = 0.0;
    if constexpr (OFF) off = A.coff[(long long)t * p + j];
    const double cv = r * (upd ? yv - za - off : 0.0);
    acc_log(upd ? F : 1.0);
    n_terms += upd;
    non_init += obs;
    F_eq0 += obs && !upd;
    sumq = fma(cv, cv, sumq);
    double s[MC];
#pragma unroll
    for (int k = 0; k < MC; k++) {
      s[k] = M[k] * r;
      a[k] = fma(s[k], cv, a[k]);
    }
#pragma unroll
    for (int k = 0; k < MC; k++) {
#pragma unroll
      for (int l = k; l < MC; l++) C[core_sym(MC, k, l)] = fma(-s[k], s[l], C[core_sym(MC, k, l)]);
    }
  };

  auto time_update = [&](bool with_inf) {
    const double* const Tc = &A.T[0] + core_opaque0();
    double ta[MC];
#pragma unroll
    for (int i = 0; i < MC; i++) {
      double acc = Tc[i * MC] * a[0];
#pragma unroll
      for (int k = 1; k < MC; k++) acc = fma(Tc[i * MC + k], a[k], acc);
      ta[i] = acc;
    }
#pragma unroll
    for (int i = 0; i < MC; i++) {
      a[i] = ta[i];
      if constexpr (OFF) a[i] += A.boff[(long long)t * MC + i];
    }
    core_time_update<MC, true>(C, Tc, &A.W[0] + core_opaque0());
    if constexpr (DIFF) {
      if (with_inf) core_time_update<MC, false>(Ci, Tc, nullptr);
    }
  };

  if constexpr (DIFF) {
    auto inf_small = [&]() {  // all(|P_inf| < 1e-7), NaN compares false (:49,158,205)
      bool ok = true;
#pragma unroll
      for (int i = 0; i < NS; i++) ok = ok && (fabs(Ci[i]) < kEps);
      return ok;
    };
    bool init = !inf_small();
    for (; t < N && init; t++) {
#pragma unroll 1
      for (int j = 0; j < p; j++) {
        const double yv = take_y();
        if (!init) {
          if constexpr (!P1) load_z(j);  // the diffuse branch below reads its row directly and does not keep the look-ahead current
          regular_obs(j, yv);
          continue;
        }
        if (isnan(yv)) continue;  // :88-90
        const double* const z = &A.Z[0][0] + j * kCoreMaxMC;
        double Ms[MC], Mi[MC];
#pragma unroll
        for (int k = 0; k < MC; k++) {
          double as = 0.0, ai = 0.0;
#pragma unroll
          for (int l = 0; l < MC; l++) {
            as = fma(C[core_sym(MC, k, l)], z[l], as);
            ai = fma(Ci[core_sym(MC, k, l)], z[l], ai);
          }
          Ms[k] = as;
          Mi[k] = ai;
        }
        double Fs = A.H[j], Fi = 0.0, za = 0.0;
#pragma unroll
        for (int k = 0; k < MC; k++) {
          Fs = fma(z[k], Ms[k], Fs);
          Fi = fma(z[k], Mi[k], Fi);
          za = fma(z[k], a[k], za);
        }
        if (Fi < kEps) {                // :109
          if (Fs < kEps) continue;      // :112
          const double F1 = 1.0 / Fs;   // :117-130
          double off = 0.0;
        if constexpr (OFF) off = A.coff[(long long)t * p + j];
        const double v = yv - za - off, g = F1 * v;
#pragma unroll
          for (int k = 0; k < MC; k++) a[k] = fma(Ms[k], g, a[k]);
#pragma unroll
          for (int k = 0; k < MC; k++) {
            const double sk = Ms[k] * F1;
#pragma unroll
            for (int l = k; l < MC; l++) C[core_sym(MC, k, l)] = fma(-sk, Ms[l], C[core_sym(MC, k, l)]);
          }
          acc_log(Fs);
          n_terms++;
          sumq = fma(v * v, F1, sumq);
          continue;
        }
        const double F1 = 1.0 / Fi, F2 = -(F1 * F1) * Fs;  // :136-153
        double off = 0.0;
        if constexpr (OFF) off = A.coff[(long long)t * p + j];
        const double v = yv - za - off, g = F1 * v;
#pragma unroll
        for (int k = 0; k < MC; k++) a[k] = fma(Mi[k], g, a[k]);
#pragma unroll
        for (int k = 0; k < MC; k++) {
          const double ik = Mi[k] * F1, sk = Ms[k] * F1, jk = Mi[k] * F2;
#pragma unroll
          for (int l = k; l < MC; l++) {
            const int e = core_sym(MC, k, l);
            C[e] = fma(-jk, Mi[l], fma(-sk, Mi[l], fma(-ik, Ms[l], C[e])));
            Ci[e] = fma(-ik, Mi[l], Ci[e]);
          }
        }
        acc_log(Fi);
        n_terms++;
        if (j < p - 1) init = !inf_small();  // :157-159
      }
      if (t < N - 1) {  // :200-206
        time_update(init);
        if (init) init = !inf_small();
      }
    }
  }
  for (; t < N; t++) {
#pragma unroll 1
    for (int j = 0; j < p; j++) regular_obs(j, take_y());
    if (t < N - 1) time_update(false);
  }
  double res;
  if (non_init == F_eq0) {
    res = nan("");  // src/LogLikC.cpp:211-213
  } else {
    const double sumlog = (double)esum * 0.69314718055994530942 + log(prod);
    res = ((double)n_terms * kLogConst - 0.5 * (sumlog + sumq)) / (double)N;
  }
  A.out[w] = res;
}

// FP64 operations (fma = 2) per series and time point of the regular phase
int core_flops_per_step(int mc, int p) {
  const int ns = mc * (mc + 1) / 2;
  // per observation: M (mc mul + mc (mc - 1) fma), F and z'a (2 mc fma), 1/sqrt(F) (4 mul + 6 fma), cv (1 mul + 1 add),
  // product (1 mul), sumq (1 fma), s and a (mc mul + mc fma), rank-1 (ns fma)
  const int fma_o = mc * (mc - 1) + 2 * mc + 6 + 1 + mc + ns, mul_o = mc + 4 + 1 + 1 + mc, add_o = 1;
  // time update: T a (mc mul + mc (mc - 1) fma), X T' (mc^2 mul + mc^2 (mc - 1) fma), T (X T') + W (ns mc fma)
  const int fma_t = mc * (mc - 1) + mc * mc * (mc - 1) + ns * mc, mul_t = mc + mc * mc;
  return p * (2 * fma_o + mul_o + add_o) + 2 * fma_t + mul_t;
}

bool core_eligible(const DModel& md, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  const int mc = md.m - md.p;
  if (mc < 1 || mc > 2 * kCoreMaxMC) return no("m - p outside 1 .. 16 (at most 8 of them with a variance)");
  if (md.p < 1 || md.p > kCoreMaxP) return no("p outside 1 .. 16");
  if (!md.Z.shared() || !md.T.shared() || !md.R.shared() || !md.Q.shared() || !md.a.shared() || !md.P_inf.shared() ||
      !md.P_star.shared())
    return no("system matrices must be shared by the batch");
  if (md.Z.tv() || md.T.tv() || md.R.tv() || md.Q.tv()) return no("time-varying system matrices");
  return true;
}

template <int MC>
static int launch_core_mc(const CoreArgs& A, bool diffuse, cudaStream_t st) {
  const long long blocks = (A.E + kCoreThreads - 1) / kCoreThreads;
  SS_ARG(blocks < (1LL << 31), "core kernel: batch too large for one launch");
  const unsigned g = (unsigned)blocks;
  const bool off = A.coff != nullptr, p1 = A.p == 1 && !off;
  if (diffuse && p1)
    loglik_core_kernel<MC, true, false, true><<<g, kCoreThreads, 0, st>>>(A);
  else if (p1)
    loglik_core_kernel<MC, false, false, true><<<g, kCoreThreads, 0, st>>>(A);
  else if (diffuse && off)
    loglik_core_kernel<MC, true, true, false><<<g, kCoreThreads, 0, st>>>(A);
  else if (diffuse)
    loglik_core_kernel<MC, true, false, false><<<g, kCoreThreads, 0, st>>>(A);
  else if (off)
    loglik_core_kernel<MC, false, true, false><<<g, kCoreThreads, 0, st>>>(A);
  else
    loglik_core_kernel<MC, false, false, false><<<g, kCoreThreads, 0, st>>>(A);
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

// Host analysis of a model for the core kernel, on host copies of its matrices (`md` supplies the storage flags): R Q R',
// the residual structure the elimination rests on (bad = why it does not hold, else null), and the split of the
// component states into those that carry a variance and the deterministic ones.
struct CoreAnalysis {
  std::vector<double> W;        // R Q R', m x m
  std::vector<int> sidx, didx;  // component indices (0 = first state behind the residuals): stochastic / deterministic
  const char* bad = nullptr;
};

static void core_analyse(const DModel& md, const std::vector<double>& Th, const std::vector<double>& Rh,
                         const std::vector<double>& Zh, const std::vector<double>& Qh, const std::vector<double>& ah,
                         const std::vector<double>& Pih, const std::vector<double>& Psh, CoreAnalysis& out) {
  const int m = md.m, r = md.r, p = md.p, mc = m - p;
  auto Tm = [&](int i, int k) { return mat_get(Th.data(), md.T.diag, m, i, k); };
  auto Rm = [&](int i, int k) { return mat_get(Rh.data(), md.R.diag, m, i, k); };
  auto Zm = [&](int i, int k) { return mat_get(Zh.data(), md.Z.diag, p, i, k); };
  auto Qm = [&](int i, int k) { return mat_get(Qh.data(), md.Q.diag, r, i, k); };
  auto Pim = [&](int i, int k) { return mat_get(Pih.data(), md.P_inf.diag, m, i, k); };
  auto Psm = [&](int i, int k) { return mat_get(Psh.data(), md.P_star.diag, m, i, k); };
  // R Q R' (m x m)
  std::vector<double> RQ((size_t)m * r, 0.0);
  std::vector<double>& W = out.W;
  W.assign((size_t)m * m, 0.0);
  for (int i = 0; i < m; i++)
    for (int k = 0; k < r; k++) {
      const double rik = Rm(i, k);
      if (rik == 0.0) continue;
      for (int l = 0; l < r; l++) RQ[i + (size_t)l * m] += rik * Qm(k, l);
    }
  for (int i = 0; i < m; i++)
    for (int l = 0; l < r; l++) {
      const double x = RQ[i + (size_t)l * m];
      if (x == 0.0) continue;
      for (int k = 0; k < m; k++) W[i + (size_t)k * m] += x * Rm(k, l);
    }
  const char* bad = nullptr;
  for (int j = 0; j < p && !bad; j++) {
    for (int k = 0; k < m; k++) {
      if (Tm(j, k) != 0.0 || Tm(k, j) != 0.0) bad = "T couples the residual states";
      if (k != j && (W[j + (size_t)k * m] != 0.0 || W[k + (size_t)j * m] != 0.0)) bad = "R Q R' couples a residual state (H not diagonal)";
      if (Pim(j, k) != 0.0 || Pim(k, j) != 0.0) bad = "P_inf is not zero in the residual rows";
      if (k != j && (Psm(j, k) != 0.0 || Psm(k, j) != 0.0)) bad = "P_star couples a residual state";
      if (k < p && Zm(j, k) != (k == j ? 1.0 : 0.0)) bad = "Z does not start with the identity";
    }
    if (Psm(j, j) != W[j + (size_t)j * m]) bad = "P_star and R Q R' differ in the residual block";
    if (ah[j] != 0.0) bad = "a is not zero in the residual rows";
  }
  out.bad = bad;
  if (bad) return;
  // Deterministic component states: no initial variance (P_star and P_inf rows zero), no disturbance (R Q R' row zero)
  // and driven by deterministic states only (T[d, s] = 0 for every other state s).  Their rows of P stay zero for ever
  // and their values are the same known numbers for every series -- the constants of a VAR with a mean (the dynamic
  // Nelson-Siegel model of config 3 carries three), fixed effects.  They leave the recursion: what they add to z'a per
  // observation and to T a per time update is tabulated here once (N x p and N x mc_s doubles) and the kernel runs on
  // the remaining states.
  std::vector<int> det((size_t)mc, 1);
  for (int i = 0; i < mc; i++)
    for (int k = 0; k < mc && det[i]; k++)
      if (Psm(p + i, p + k) != 0.0 || Psm(p + k, p + i) != 0.0 || Pim(p + i, p + k) != 0.0 || Pim(p + k, p + i) != 0.0 ||
          W[(p + i) + (size_t)(p + k) * m] != 0.0 || W[(p + k) + (size_t)(p + i) * m] != 0.0)
        det[i] = 0;
  for (bool changed = true; changed;) {
    changed = false;
    for (int i = 0; i < mc; i++) {
      if (!det[i]) continue;
      for (int k = 0; k < mc; k++)
        if (!det[k] && Tm(p + i, p + k) != 0.0) {
          det[i] = 0;
          changed = true;
          break;
        }
    }
  }
  std::vector<int>& sidx = out.sidx;
  std::vector<int>& didx = out.didx;
  sidx.clear();
  didx.clear();
  for (int i = 0; i < mc; i++) (det[i] ? didx : sidx).push_back(i);
  if (sidx.empty()) {  // nothing stochastic but the residuals: keep one state in the recursion
    sidx.push_back(didx.back());
    didx.pop_back();
  }
}

// applicable != null: a model without the residual structure is not an error; *applicable = false, nothing runs
int launch_loglik_core(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                       const HostShared* hs, bool* applicable) {
  std::string why;
  if (applicable) *applicable = true;
  SS_ARG(core_eligible(md, &why), "core kernel not eligible: " + why);
  const int m = md.m, p = md.p;
  auto count = [](const DMat& M) { return (size_t)(M.diag ? M.rows : M.rows * M.cols); };
  std::vector<double> Th(count(md.T)), Rh(count(md.R)), Zh(count(md.Z)), Qh(count(md.Q)), ah(count(md.a)),
      Pih(count(md.P_inf)), Psh(count(md.P_star));
  bool need_sync = false;
  auto fetch = [&](std::vector<double>& dst, const double* host, const double* dev) -> int {
    if (host) {
      std::copy(host, host + dst.size(), dst.begin());
    } else {
      SS_CUDA(cudaMemcpyAsync(dst.data(), dev, dst.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
      need_sync = true;
    }
    return SSGPU_OK;
  };
  SS_TRY(fetch(Th, hs ? hs->T : nullptr, md.T.p));
  SS_TRY(fetch(Rh, hs ? hs->R : nullptr, md.R.p));
  SS_TRY(fetch(Zh, hs ? hs->Z : nullptr, md.Z.p));
  SS_TRY(fetch(Qh, hs ? hs->Q : nullptr, md.Q.p));
  SS_TRY(fetch(ah, nullptr, md.a.p));
  SS_TRY(fetch(Pih, nullptr, md.P_inf.p));
  SS_TRY(fetch(Psh, nullptr, md.P_star.p));
  if (need_sync) SS_CUDA(cudaStreamSynchronize(st));
  auto Tm = [&](int i, int k) { return mat_get(Th.data(), md.T.diag, m, i, k); };
  auto Zm = [&](int i, int k) { return mat_get(Zh.data(), md.Z.diag, p, i, k); };
  auto Pim = [&](int i, int k) { return mat_get(Pih.data(), md.P_inf.diag, m, i, k); };
  auto Psm = [&](int i, int k) { return mat_get(Psh.data(), md.P_star.diag, m, i, k); };
  CoreAnalysis an;
  core_analyse(md, Th, Rh, Zh, Qh, ah, Pih, Psh, an);
  const char* const bad = an.bad;
  if (bad) {
    if (applicable) {
      *applicable = false;
      return SSGPU_OK;
    }
    SS_ARG(false, std::string("core kernel not eligible: ") + bad);
  }
  const std::vector<double>& W = an.W;
  const std::vector<int>&sidx = an.sidx, &didx = an.didx;
  const int ms = (int)sidx.size(), mdn = (int)didx.size();
  if (ms > kCoreMaxMC) {
    if (applicable) {
      *applicable = false;
      return SSGPU_OK;
    }
    SS_ARG(false, "core kernel not eligible: more than 8 component states with a variance");
  }
  CoreArgs A{};
  A.y = y;
  A.B = B;
  A.E = B * V;
  A.N = md.N;
  A.p = p;
  A.out = out;
  DevBuf offs;  // released stream-ordered behind the kernel
  if (mdn > 0) {
    const int N = md.N;
    std::vector<double> tab((size_t)N * p + (size_t)N * ms), ad((size_t)mdn), nx((size_t)mdn);
    for (int k = 0; k < mdn; k++) ad[k] = ah[p + didx[k]];
    for (int t = 0; t < N; t++) {
      for (int j = 0; j < p; j++) {
        double acc = 0.0;
        for (int k = 0; k < mdn; k++) acc += Zm(j, p + didx[k]) * ad[k];
        tab[(size_t)t * p + j] = acc;
      }
      for (int i = 0; i < ms; i++) {
        double acc = 0.0;
        for (int k = 0; k < mdn; k++) acc += Tm(p + sidx[i], p + didx[k]) * ad[k];
        tab[(size_t)N * p + (size_t)t * ms + i] = acc;
      }
      for (int i = 0; i < mdn; i++) {
        double acc = 0.0;
        for (int k = 0; k < mdn; k++) acc += Tm(p + didx[i], p + didx[k]) * ad[k];
        nx[i] = acc;
      }
      ad = nx;
    }
    SS_TRY(offs.alloc(tab.size() * sizeof(double), st));
    SS_CUDA(cudaMemcpyAsync(offs.get(), tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaStreamSynchronize(st));  // the host vector dies with this frame
    A.coff = offs.as<double>();
    A.boff = offs.as<double>() + (size_t)N * p;
  }
  bool diffuse = false;
  for (int i = 0; i < ms; i++) {
    const int ci = p + sidx[i];
    A.a0[i] = ah[ci];
    for (int k = 0; k < ms; k++) {
      const int ck = p + sidx[k];
      A.T[i * ms + k] = Tm(ci, ck);
      if (k >= i) {
        const int e = core_sym(ms, i, k);
        A.W[e] = W[ci + (size_t)ck * m];
        A.P0[e] = Psm(ci, ck);
        A.Pi0[e] = Pim(ci, ck);
        // all(|P_inf| < 1e-7) with NaN comparing false (src/LogLikC.cpp:49)
        if (!(std::fabs(A.Pi0[e]) < kEps)) diffuse = true;
      }
    }
  }
  for (int j = 0; j < p; j++) {
    A.H[j] = W[j + (size_t)j * m];
    for (int k = 0; k < ms; k++) A.Z[j][k] = Zm(j, p + sidx[k]);
  }
  int rc = SSGPU_E_ARG;
  switch (ms) {
#define SS_CASE(MM)                            \
  case MM:                                     \
    rc = launch_core_mc<MM>(A, diffuse, st);   \
    break;
    SS_CASE(1) SS_CASE(2) SS_CASE(3) SS_CASE(4) SS_CASE(5) SS_CASE(6) SS_CASE(7) SS_CASE(8)
#undef SS_CASE
  }
  SS_TRY(rc);
  count_launch(1);
  static thread_local char name[96];
  snprintf(name, sizeof name, "loglik_core_kernel<MC=%d%s%s>%s", ms, diffuse ? ",diffuse" : "", (p == 1 && mdn == 0) ? ",p1" : "",
           mdn > 0 ? " deterministic states tabulated" : "");
  set_last_kernel(name);
  return SSGPU_OK;
}

}  // namespace ssgpu

extern "C" int ssgpu_core_flops_per_step(int m, int p) {
  const int mc = m - p;
  if (mc < 1 || mc > ssgpu::kCoreMaxMC || p < 1 || p > ssgpu::kCoreMaxP) return 0;
  return ssgpu::core_flops_per_step(mc, p);
}

// Host-only helper (no device needed): what the core kernel does with a model whose matrices are shared by the batch.
extern "C" int ssgpu_plan_core(const double* T, const double* R, const double* Q, const double* Z, const double* a,
                               const double* P_inf, const double* P_star, int m, int r, int p, int* eligible,
                               int* n_stochastic, int* n_deterministic, int* diffuse) {
  using namespace ssgpu;
  SS_ARG(T && R && Q && Z && a && P_inf && P_star && m >= 1 && r >= 1 && p >= 1, "ssgpu_plan_core: bad arguments");
  for (int* o : {eligible, n_stochastic, n_deterministic, diffuse})
    if (o) *o = 0;
  DModel md{};
  md.N = 1;
  md.p = p;
  md.m = m;
  md.r = r;  // every matrix dense, shared and time-invariant: strides 0, diag 0
  std::string why;
  if (!core_eligible(md, &why)) return SSGPU_OK;
  const std::vector<double> Th(T, T + (size_t)m * m), Rh(R, R + (size_t)m * r), Zh(Z, Z + (size_t)p * m),
      Qh(Q, Q + (size_t)r * r), ah(a, a + (size_t)m), Pih(P_inf, P_inf + (size_t)m * m), Psh(P_star, P_star + (size_t)m * m);
  CoreAnalysis an;
  core_analyse(md, Th, Rh, Zh, Qh, ah, Pih, Psh, an);
  if (an.bad || (int)an.sidx.size() > kCoreMaxMC) return SSGPU_OK;
  if (eligible) *eligible = 1;
  if (n_stochastic) *n_stochastic = (int)an.sidx.size();
  if (n_deterministic) *n_deterministic = (int)an.didx.size();
  bool dif = false;
  for (int i : an.sidx)
    for (int k : an.sidx)
      if (!(std::fabs(P_inf[(p + i) + (size_t)(p + k) * m]) < kEps)) dif = true;
  if (diffuse) *diffuse = dif ? 1 : 0;
  return SSGPU_OK;
}
