// Regular phase of the batched loglikelihood for block-diagonal T (blocks of at most 2 x 2), ONE THREAD PER
// EVALUATION, the symmetric covariance held once (blocks on and above the diagonal) in registers and, for what does
// not fit, in shared memory.
//
// Same recursion and model class as loglik_blk.cu (src/LogLikC.cpp:162-207; p = 1, Z / T / R shared by the batch and
// time-invariant, 2 to 8 blocks: 3 <= m <= 16 -- BASELINE.json config 5 is 7 blocks; up to 4 blocks everything lives in
// registers at 4 CTAs per SM).  The lane kernel of
// loglik_blk.cu runs the exact diffuse phase (about d of the N time points, src/LogLikC.cpp:98-159) and leaves the
// state behind (blk_handoff.cuh); this kernel runs the remaining time points.
//
// Why one thread per evaluation.  Spread over 8 lanes, an evaluation pays for the exchange of M = P z between the
// lanes (36 shuffles per step), for 8 copies of the scalar part of a step (1/F, the loglikelihood terms), for one
// idle lane in 8 at 7 blocks, and every lane needs its own copies of T_J and z in vector registers: 1,168
// FP64 instructions per series and time step.  Here T and z are the same for all 32 lanes at every instruction, so ptxas
// keeps them in UNIFORM registers (DFMA takes a uniform-register operand) and the vector registers hold P alone;
// nothing is exchanged, nothing is padded, nothing is computed twice: 861 FP64 instructions per series and time step
// in the generic instantiation -- 600 with the structure of a reference model taken out of the arithmetic (EPS, LS, ZS
// below; config 5), 83 % of all instructions of the loop.
//
// One sweep per time step (p = 1): with M = P z known from the previous sweep,
//     tm = T M,  F = z'M,  v = y - z'a,  s = tm / sqrt(F),  a <- T a + s (v / sqrt(F)),
// then block by block  P_IJ <- T_I P_IJ T_J' (+ R Q R' on the diagonal) - s_I s_J'   [= T (P - M M'/F) T' + R Q R']
// and at once the block's share of the NEXT step's M:  M_I += P_IJ z_J,  M_J += P_IJ' z_I.  Every block is read once
// and written once per step; the blocks are independent, so the FP64 pipe never waits for the reciprocal square root.
// A missing observation or F < 1e-7 (src/LogLikC.cpp:88-90,173-176) is a no-op through 1/sqrt(F) := 0.
//
// Storage (7 blocks): 7 diagonal blocks x 3 + 21 blocks above the diagonal x 4 = 105 doubles.  255 registers hold
// the diagonal blocks, the first NOFF - S blocks above the diagonal, s and the M being accumulated; the last S blocks,
// a and the diagonal of R Q R' live in shared memory as [slot][thread] (conflict-free, one read and one write per
// step).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "blk_handoff.cuh"
#include "common.cuh"

namespace ssgpu {

#ifndef SSGPU_THR_THREADS
#define SSGPU_THR_THREADS 128
#endif
constexpr int kThrThreads = SSGPU_THR_THREADS;  // 2 x 128 or 4 x 64 threads per SM: 255 registers either way
constexpr int kThrMaxNB = 8;
#ifndef SSGPU_THR_BIG_THREADS
#define SSGPU_THR_BIG_THREADS 256  // resident threads per SM the kernels of 5 to 8 blocks are sized for (255 registers)
#endif

struct ThrArgs {
  const double* y;  // [t * ldy + b]
  long long B, E, ldy;
  int N;
  const double* hd;
  const int* hi;
  double* out;  // [v * ldy + b]
  double cst[kThrMaxNB][6];  // T_I row-major (4), z_I (2): internal order of the block plan
  const double* zt;          // time-varying Z (explanatory variables): [t][32] in internal order, device memory; else null
  ThrDiffuse ud;             // the shared diffuse steps of evaluations handed over at t = 0 (marked -1); du = 0: none
};

__device__ __forceinline__ double thr_rsqrt(double x) {
  // MUFU.RSQ64H seed and three coupled Newton steps (x is finite, normal and >= 1e-7 here); relative error of the
  // seed <= 2^-20, after the steps below the rounding of the last fma
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double h = 0.5 * x;
  double e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
#ifndef SSGPU_THR_NEWTON2
  e = fma(-h * r, r, 0.5);
  r = fma(r, e, r);
#endif
  return r;
}

// An offset of zero the optimiser cannot see through.  T and z are read from the kernel parameters through it inside
// the time loop: ptxas then serves them from uniform registers / the constant bank at the instruction that uses
// them.  Read at fixed parameter offsets they are hoisted out of the loop into 84 VECTOR registers and the
// covariance spills.
__device__ __forceinline__ int thr_opaque0() {
  int o;
  asm volatile("mov.u32 %0, 0;" : "=r"(o));
  return o;
}

// S: blocks above the diagonal kept in shared memory (the last S in row-major order)
// ZTV: Z changes with t (R/Components-Regressors.R:60-78): the row of a time point is loaded by every thread for ITS time
// point (the threads of a warp differ by a few steps at most: a handful of L1-resident lines) into registers that replace
// the uniform-register copies of z; F = z'M and z'a of a step are formed right behind the sweep that accumulated M, while
// that row is still at hand, and carried into the next iteration.
//
// EPS: position 1 of the LAST block is the residual state of the reference's layout (z = 1, T row and column zero, no
// coupling through R Q R': R/GetSysMat.R:1388-1426).  At the start of every time point its row of P is H e' and its
// entry of a is 0, and its row of T P T' vanishes: the kernel keeps nothing for it -- the last block is the scalar of
// its partner, the blocks (I, last) are 2 x 1, F gets H added.  At 7 blocks (config 5) that is 738 instead of 861 FP64
// instructions per series and time step.  Only P[eps, eps] of the FIRST time point is taken from the hand-over (a model
// without diffuse phase starts from the caller's P_star); the launcher checks what the kernel assumes of P_star and a.
//
// LS: block 0 is a local level with slope, T_0 = [[1, 1], [0, 1]] (R/Components-Unobserved.R:150-159).  Its products are
// sums: T_0 X = [x0 + x1 ; x1] row-wise -- one DADD where the generic block spends two multiplications and two
// multiply-adds, bit for bit the same result (1 * x is exact).  48 FP64 instructions fewer per step at 7 blocks.
//
// ZS: the second state of every block has z = 0 (level + slope, trigonometric seasonal pairs, cycles: Z = (1, 0) per pair,
// R/Components-Unobserved.R:143,491-500): the multiply-adds with that zero are left out -- fma(x, 0, acc) == acc, so
// the same bits again.  90 FP64 instructions fewer per step at 7 blocks.
template <int NB, int S, bool QOFF, bool ZTV = false, bool EPS = false, bool LS = false, bool ZS = false>
__global__ void __launch_bounds__(kThrThreads, (NB <= 4 ? 512 : SSGPU_THR_BIG_THREADS) / kThrThreads) loglik_thr_kernel(const ThrArgs A) {
  static_assert(!(ZS && ZTV), "sparse-z shortcut: time-invariant Z");
  auto z1fma = [](double x, double z, double acc) { return ZS ? acc : fma(x, z, acc); };
  static_assert(!(EPS && ZTV), "residual shortcut: time-invariant Z");
  static_assert(!(LS && ZTV), "level + slope shortcut: time-invariant Z");
  constexpr int L = NB - 1;
  constexpr BlkHandoff HL(NB, QOFF);
  constexpr int NOFF = HL.noff();
  constexpr int RO = NOFF - S;  // blocks above the diagonal in registers
  constexpr int QW = HL.qw;
  static_assert(S >= 0 && S <= NOFF, "shared-memory blocks");
  const long long w = (long long)blockIdx.x * kThrThreads + threadIdx.x;
  if (w >= A.E) return;
  const long long E = A.E;
  int t = A.hi[w];
  const int N = A.N;
  if (t >= N) return;  // finished by the lane kernel
  // handed over as it stood at t = 0: no missing value among its first du observations, so its diffuse steps are the
  // shared ones (below)
  const bool pro = t < 0;
  if (pro) t = 0;
  const long long V = E / A.B, b = w / V, v = w % V;
  extern __shared__ double thr_sm[];  // [slot][thread]: a (2 NB), R Q R' (QW NB), S blocks (4 each)
  double* const smt = thr_sm + threadIdx.x;
  auto SA = [&](int i) -> double& { return smt[i * kThrThreads]; };
  auto SQ = [&](int i) -> double& { return smt[(2 * NB + i) * kThrThreads]; };
  auto SP = [&](int i) -> double& { return smt[((2 + QW) * NB + i) * kThrThreads]; };
  const double* const hd = A.hd + w;
  double D[NB][3], O[RO > 0 ? RO : 1][4], M[NB][2];
#pragma unroll
  for (int i = 0; i < 2 * NB; i++) SA(i) = hd[(long long)(HL.a() + i) * E];
#pragma unroll
  for (int i = 0; i < QW * NB; i++) SQ(i) = hd[(long long)(HL.q() + i) * E];
#pragma unroll
  for (int I = 0; I < NB; I++) {
#pragma unroll
    for (int e = 0; e < 3; e++) D[I][e] = hd[(long long)(HL.d() + 3 * I + e) * E];
  }
#pragma unroll
  for (int i = 0; i < NOFF; i++) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const double x = hd[(long long)(HL.o() + 4 * i + e) * E];
      if (i < RO)
        O[i][e] = x;
      else
        SP((i - RO) * 4 + e) = x;
    }
  }
  double prod = hd[(long long)HL.s() * E], sumq = hd[(long long)(HL.s() + 1) * E];
  int esum = A.hi[E + w], n_upd = 0, n_obs = 0;

  double zr[ZTV ? NB : 1][2];  // ZTV: the row of Z the M being accumulated belongs to
  auto load_z = [&](int tt) {
    if constexpr (ZTV) {
      const double2* zp = reinterpret_cast<const double2*>(A.zt + (size_t)tt * 32);
#pragma unroll
      for (int I = 0; I < NB; I++) {
        const double2 z2 = zp[I];
        zr[I][0] = z2.x;
        zr[I][1] = z2.y;
      }
    }
  };
  load_z(t);
  // M = P z of the first time point (afterwards every sweep leaves the next step's M behind)
  {
    int oi = 0;
#pragma unroll
    for (int I = 0; I < NB; I++) M[I][0] = M[I][1] = 0.0;
#pragma unroll
    for (int I = 0; I < NB; I++) {
      const double zi0 = ZTV ? zr[ZTV ? I : 0][0] : A.cst[I][4], zi1 = ZTV ? zr[ZTV ? I : 0][1] : A.cst[I][5];
      if (EPS && I == L) {
        M[I][0] = fma(D[I][0], zi0, M[I][0]);
        continue;
      }
      M[I][0] = z1fma(D[I][1], zi1, fma(D[I][0], zi0, M[I][0]));
      M[I][1] = z1fma(D[I][2], zi1, fma(D[I][1], zi0, M[I][1]));
#pragma unroll
      for (int J = I + 1; J < NB; J++, oi++) {
        const double zj0 = ZTV ? zr[ZTV ? J : 0][0] : A.cst[J][4], zj1 = ZTV ? zr[ZTV ? J : 0][1] : A.cst[J][5];
        double p0, p1, p2, p3;
        if (oi < RO) {
          p0 = O[oi < RO ? oi : 0][0], p1 = O[oi < RO ? oi : 0][1], p2 = O[oi < RO ? oi : 0][2], p3 = O[oi < RO ? oi : 0][3];
        } else {
          const int o = (oi - RO) * 4;
          p0 = SP(o), p1 = SP(o + 1), p2 = SP(o + 2), p3 = SP(o + 3);
        }
        if (EPS && J == L) {
          M[I][0] = fma(p0, zj0, M[I][0]);
          M[I][1] = fma(p2, zj0, M[I][1]);
          M[J][0] = z1fma(p2, zi1, fma(p0, zi0, M[J][0]));
          continue;
        }
        M[I][0] = z1fma(p1, zj1, fma(p0, zj0, M[I][0]));
        M[I][1] = z1fma(p3, zj1, fma(p2, zj0, M[I][1]));
        M[J][0] = z1fma(p2, zi1, fma(p0, zi0, M[J][0]));
        M[J][1] = z1fma(p3, zi1, fma(p1, zi0, M[J][1]));
      }
    }
  }

  // loglik terms of one observation (src/LogLikC.cpp:173-195): r = 1/sqrt(F) or 0, cv = v / sqrt(F)
  double r, cv;
  auto observe = [&](double yv, double F, double za) {
    const bool obs = !isnan(yv);  // :88-90
    const bool upd = obs && !(F < kEps);
    r = thr_rsqrt(upd ? F : 1.0);
    r = upd ? r : 0.0;
    cv = r * (upd ? yv - za : 0.0);
    prod *= upd ? F : 1.0;
    const int hi = __double2hiint(prod);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    esum += ex;
    prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
    n_upd += upd;
    n_obs += obs;
    sumq = fma(cv, cv, sumq);
  };

  const double* yp = A.y + b + (long long)t * A.ldy;
  double ynext = *yp;
  // EPS: P[eps, eps] z_eps, the residual's share of F -- the hand-over's value at the first time point, H afterwards
  double heps = EPS ? D[L][2] : 0.0;
  const double hq = EPS ? SQ(QW * L + QW - 1) : 0.0;
  double Fc = 0.0, zac = 0.0;  // ZTV: F = z'M and z'a of the coming step
  if constexpr (ZTV) {
#pragma unroll
    for (int I = 0; I < NB; I++) {
      Fc = fma(zr[ZTV ? I : 0][1], M[I][1], fma(zr[ZTV ? I : 0][0], M[I][0], Fc));
      zac = fma(zr[ZTV ? I : 0][1], SA(2 * I + 1), fma(zr[ZTV ? I : 0][0], SA(2 * I), zac));
    }
  }
  // ---- shared diffuse steps (src/LogLikC.cpp:136-153,200-205).  P_inf, F_inf and K0 = M_inf / F_inf depend on Z, T and
  // P_inf only: for an evaluation without a missing value in the diffuse window they are constants of the model, computed
  // on the host (loglik_blk.cu) -- u_t = T M_inf and F1_t = 1 / F_inf per step.  What is left per evaluation is P_star:
  //     F* = z'M*,  F2 = -F1^2 F*,  v = y - z'a,  w = T M*,  g = F1 w + (F2 / 2) u,
  //     P* <- T P* T' + R Q R' - u g' - g u'     [= T (P* - F1 (M_inf M*' + M* M_inf') - F2 M_inf M_inf') T' + R Q R'],
  //     a  <- T a + u F1 v,                       loglik term: -log(F_inf) / 2, a constant of the model.
  // The same sweep as the regular phase with a rank-2 correction; the residual / level + slope / zero-z shortcuts hold
  // here as well (u and g vanish in the residual's row like s does).
  int n_dif = 0;
  if constexpr (!ZTV) {
    if (pro) {
      const int du = A.ud.du;  // < N
      for (; t < du; t++) {
        const double yv = ynext;
        yp += A.ldy;
        ynext = *yp;
#include "loglik_thr_vectors.inc"
        const double* const ut = &A.ud.u[0][0] + t * 16 + thr_opaque0();
        const double F1 = (&A.ud.F1[0] + thr_opaque0())[t], hF2 = -0.5 * (F1 * F1) * F, gv = F1 * (yv - za);
        double g[NB][2], uu[NB][2];
#pragma unroll
        for (int I = 0; I < NB; I++) {
          uu[I][0] = ut[2 * I];
          g[I][0] = fma(F1, M[I][0], hF2 * uu[I][0]);
          SA(2 * I) = fma(uu[I][0], gv, ta[I][0]);
          M[I][0] = 0.0;
          if (EPS && I == L) continue;
          uu[I][1] = ut[2 * I + 1];
          g[I][1] = fma(F1, M[I][1], hF2 * uu[I][1]);
          SA(2 * I + 1) = fma(uu[I][1], gv, ta[I][1]);
          M[I][1] = 0.0;
        }
#define THR_RK(I, a, J, b, x) fma(-g[I][a], uu[J][b], fma(-uu[I][a], g[J][b], (x)))
#include "loglik_thr_sweep.inc"
#undef THR_RK
      }
      n_dif = du;
      esum += A.ud.esumF;
      prod *= A.ud.prodF;
    }
  }
  for (; t < N - 1; t++) {
    const double yv = ynext;
    yp += A.ldy;
    ynext = *yp;
#include "loglik_thr_vectors.inc"
    observe(yv, F, za);
    load_z(t + 1);  // ZTV: the row the sweep below accumulates the next M with
    zac = 0.0;
    double s[NB][2];
#pragma unroll
    for (int I = 0; I < NB; I++) {
      if (EPS && I == L) {
        s[I][0] = M[I][0] * r;
        SA(2 * I) = fma(s[I][0], cv, ta[I][0]);
        M[I][0] = 0.0;
        continue;
      }
      s[I][0] = M[I][0] * r;
      s[I][1] = M[I][1] * r;
      const double n0 = fma(s[I][0], cv, ta[I][0]), n1 = fma(s[I][1], cv, ta[I][1]);
      SA(2 * I) = n0;
      SA(2 * I + 1) = n1;
      if constexpr (ZTV) zac = fma(zr[ZTV ? I : 0][1], n1, fma(zr[ZTV ? I : 0][0], n0, zac));
      M[I][0] = 0.0;
      M[I][1] = 0.0;
    }
#define THR_RK(I, a, J, b, x) fma(-s[I][a], s[J][b], (x))
#include "loglik_thr_sweep.inc"
#undef THR_RK
    if constexpr (ZTV) {
      Fc = 0.0;
#pragma unroll
      for (int I = 0; I < NB; I++) Fc = fma(zr[ZTV ? I : 0][1], M[I][1], fma(zr[ZTV ? I : 0][0], M[I][0], Fc));
    }
  }
  {  // the last observation (src/LogLikC.cpp:200 -- P and a are not needed afterwards)
    double F = EPS ? heps : Fc, za = zac;
    if constexpr (!ZTV) {
#pragma unroll
      for (int I = 0; I < NB; I++) {
        if (EPS && I == L) {
          F = fma(A.cst[I][4], M[I][0], F);
          za = fma(A.cst[I][4], SA(2 * I), za);
          continue;
        }
        F = z1fma(M[I][1], A.cst[I][5], fma(A.cst[I][4], M[I][0], F));
        za = z1fma(SA(2 * I + 1), A.cst[I][5], fma(A.cst[I][4], SA(2 * I), za));
      }
    }
    observe(ynext, F, za);
  }
  const int n_terms = A.hi[2 * E + w] + n_upd + n_dif, non_init = A.hi[3 * E + w] + n_obs, F_eq0 = A.hi[4 * E + w] + (n_obs - n_upd);
  double res;
  if (non_init == F_eq0) {
    res = nan("");  // src/LogLikC.cpp:211-213
  } else {
    const double sumlog = (double)esum * 0.69314718055994530942 + log(prod);
    res = ((double)n_terms * kLogConst - 0.5 * (sumlog + sumq)) / (double)N;
  }
  A.out[v * A.ldy + b] = res;
}

// ---- records of the marked evaluations, straight from the model ------------------------------------------------------
// An evaluation the lane kernel has only marked (-1: no missing value in its diffuse window, loglik_blk.cu) has no
// hand-over record yet.  When its state at t = 0 is a matter of look-ups -- a shared by the batch, P_star and Q stored as
// diagonals, R with at most one non-zero per row and column, so that R Q R' = diag(R[i, k]^2 Q[k]) (the launcher has
// checked all of that) -- this kernel writes the record, one thread per evaluation, every store coalesced; the lanes,
// 8 per evaluation with their strided stores and their general set-up code, took 1.5 ms for it at config 5, this 0.3 ms.
// (Building the state inside loglik_thr_kernel instead saved this launch but cost its loop 2 %: ptxas scheduled the
// same 722 instructions differently.)
struct ThrSetupArgs {
  long long B, E;
  double* hd;
  int* hi;
  ThrSelf self;
};

template <int NB>
__global__ void __launch_bounds__(128) loglik_thr_setup_kernel(const ThrSetupArgs A) {
  constexpr BlkHandoff HL(NB, false);
  const long long w = (long long)blockIdx.x * 128 + threadIdx.x;
  if (w >= A.E) return;
  if (A.hi[w] != -1) return;
  const long long E = A.E, V = E / A.B, b = w / V, v = w % V;
  const double* const Qb = A.self.Q.block(b, v, 0);
  const double* const Pb = A.self.P_star.block(b, v, 0);
  double* const hd = A.hd + w;
#pragma unroll
  for (int i = 0; i < 2 * NB; i++) {
    hd[(long long)(HL.a() + i) * E] = A.self.a0[i];
    hd[(long long)(HL.q() + i) * E] = A.self.qsrc[i] >= 0 ? Qb[A.self.qsrc[i]] * A.self.qscale[i] : 0.0;
  }
#pragma unroll
  for (int I = 0; I < NB; I++) {
    hd[(long long)(HL.d() + 3 * I) * E] = A.self.psrc[2 * I] >= 0 ? Pb[A.self.psrc[2 * I]] : 0.0;
    hd[(long long)(HL.d() + 3 * I + 1) * E] = 0.0;
    hd[(long long)(HL.d() + 3 * I + 2) * E] = A.self.psrc[2 * I + 1] >= 0 ? Pb[A.self.psrc[2 * I + 1]] : 0.0;
  }
#pragma unroll 4
  for (int i = 0; i < 4 * HL.noff(); i++) hd[(long long)(HL.o() + i) * E] = 0.0;
  hd[(long long)HL.s() * E] = 1.0;
  hd[(long long)(HL.s() + 1) * E] = 0.0;
  A.hi[E + w] = 0;
  A.hi[2 * E + w] = 0;
  A.hi[3 * E + w] = 0;
  A.hi[4 * E + w] = 0;
}

int launch_thr_setup(int nb, const ThrSelf& self, long long B, long long E, double* hd, int* hi, cudaStream_t st) {
  ThrSetupArgs A{B, E, hd, hi, self};
  const long long blocks = (E + 127) / 128;
  SS_ARG(blocks < (1LL << 31), "thread kernel: batch too large for one launch");
  switch (nb) {
#define SS_SETUP_CASE(NBV) \
  case NBV:                \
    loglik_thr_setup_kernel<NBV><<<(unsigned)blocks, 128, 0, st>>>(A); \
    break;
    SS_SETUP_CASE(2) SS_SETUP_CASE(3) SS_SETUP_CASE(4) SS_SETUP_CASE(5) SS_SETUP_CASE(6) SS_SETUP_CASE(7) SS_SETUP_CASE(8)
#undef SS_SETUP_CASE
    default:
      SS_ARG(false, "thread kernel: 2 to 8 blocks");
  }
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

// blocks above the diagonal that go to shared memory: what 255 registers do not hold next to s, the M being
// accumulated and the temporaries of a block (sized with ptxas -v: no spill at these values)
#ifndef SSGPU_THR_S7
#define SSGPU_THR_S7 10
#endif
#ifndef SSGPU_THR_S8
#define SSGPU_THR_S8 17
#endif
constexpr int thr_smem_blocks(int nb) { return nb == 8 ? SSGPU_THR_S8 : nb == 7 ? SSGPU_THR_S7 : nb == 6 ? 3 : 0; }
// time-varying Z: 2 NB more vector registers hold the row of Z, so a few more blocks move to shared memory (0, 1 or 3
// extra blocks measure the same at 8 blocks: 8.2e9 series*timesteps/s; 5 extra 6.5e9)
constexpr int thr_smem_blocks_ztv(int nb, bool ztv) { return thr_smem_blocks(nb) + ((ztv && nb >= 5) ? (nb >= 7 ? 3 : 2) : 0); }

// residual shortcut: the nb - 1 blocks next to the last one are 2 x 1 and the last diagonal block is a scalar -- 14 doubles
// fewer in registers at 7 blocks, so fewer blocks in shared memory (sized with ptxas -v like the others)
#ifndef SSGPU_THR_S7E
#define SSGPU_THR_S7E 7
#endif
#ifndef SSGPU_THR_S8E
#define SSGPU_THR_S8E 16
#endif
constexpr int thr_smem_blocks_eps(int nb, bool ztv, bool eps) {
  return !eps ? thr_smem_blocks_ztv(nb, ztv) : nb == 8 ? SSGPU_THR_S8E : nb == 7 ? SSGPU_THR_S7E : nb == 6 ? 1 : 0;
}

int thr_smem_blocks_host(int nb, bool ztv, bool eps) { return thr_smem_blocks_eps(nb, ztv, eps); }

size_t thr_handoff_bytes(int nb, bool qoff, long long E) {
  const BlkHandoff hl(nb, qoff);
  return (size_t)E * (hl.doubles() * sizeof(double) + kBlkHandoffInts * sizeof(int));
}

// FP64 operations (fma = 2) per series and time step: the counts of the SASS of the loop
// instr: the number of FP64 instructions (fma = mul = add = 1 issue slot) instead of the operations (fma = 2)
int thr_flops_per_step(int nb, bool qoff, bool eps, bool ls, bool zs, bool instr) {
  const int noff = nb * (nb - 1) / 2;
  // vectors: F, z'a 4 fma; tm, T a 2 mul + 2 fma each; s 2 mul; a 2 fma    (per block row)
  const int fma_v = 4 + 4 + 2, mul_v = 4 + 2;
  // 1/sqrt(F): 3 x (1 mul + 2 fma) + 1 mul; cv 1 mul + 1 add; prod 1 mul; sumq 1 fma
  const int fma_s = 7, mul_s = 6, add_s = 1;
  // diagonal block: 4 mul + 4 fma, 3 outputs (5 fma + 1 mul, all fma with correlated R Q R'), rank-1 3 fma, M 4 fma
  const int fma_d = 4 + (qoff ? 6 : 5) + 3 + 4, mul_d = 4 + (qoff ? 0 : 1);
  // block above the diagonal: 4 mul + 4 fma, 4 mul + 4 fma, rank-1 4 fma, M 8 fma
  const int fma_o = 4 + 4 + 4 + 8, mul_o = 8;
  if (eps) {
    // the last block is a scalar (vectors 3 fma + 3 mul, diagonal 3 fma + 1 mul), the nb - 1 blocks next to it are
    // 2 x 1 (2 mul + 2 fma, 2 mul, rank-1 2 fma, M 4 fma)
    int fma = (nb - 1) * fma_v + 3 + fma_s + (nb - 1) * fma_d + 3 + (noff - (nb - 1)) * fma_o + (nb - 1) * 8;
    int mul = (nb - 1) * mul_v + 3 + mul_s + (nb - 1) * mul_d + 1 + (noff - (nb - 1)) * mul_o + (nb - 1) * 4;
    int add = add_s;
    if (ls) {  // block 0 = level + slope: vectors 2 add for 4 mul + 4 fma, diagonal 5 (4 with correlated R Q R': 6) add for
               // 5 mul + 9 fma (4 mul + 10 fma), full neighbours 2 add for 4 mul + 4 fma, the 2 x 1 neighbour 1 add for 2 + 2
      fma -= 4 + (qoff ? 10 : 9) + (nb - 2) * 4 + 2;
      mul -= 4 + (qoff ? 4 : 5) + (nb - 2) * 4 + 2;
      add += 2 + (qoff ? 6 : 5) + (nb - 2) * 2 + 1;
    }
    // z = 0 for the second state of every block: 2 (vectors) + 2 (diagonal) fma per full block row, 4 per full block above the
    // diagonal, 1 per 2 x 1 block
    if (zs) fma -= (nb - 1) * 4 + (noff - (nb - 1)) * 4 + (nb - 1);
    return (instr ? fma : 2 * fma) + mul + add;
  }
  int fma = nb * fma_v + fma_s + nb * fma_d + noff * fma_o, mul = nb * mul_v + mul_s + nb * mul_d + noff * mul_o, add = add_s;
  if (zs) fma -= nb * 4 + noff * 4;
  if (ls) {
    fma -= 4 + (qoff ? 10 : 9) + (nb - 1) * 4;
    mul -= 4 + (qoff ? 4 : 5) + (nb - 1) * 4;
    add += 2 + (qoff ? 6 : 5) + (nb - 1) * 2;
  }
  return (instr ? fma : 2 * fma) + mul + add;
}

template <int NB, bool QOFF, bool ZTV = false, bool EPS = false, bool LS = false, bool ZS = false>
static int launch_thr_one(const ThrArgs& A, cudaStream_t st) {
  // time-varying Z: 2 NB more vector registers hold the row of Z, so two more blocks move to shared memory (NB >= 5)
  constexpr int S = thr_smem_blocks_eps(NB, ZTV, EPS);
  constexpr BlkHandoff HL(NB, QOFF);
  const size_t smem = (size_t)((2 + HL.qw) * NB + 4 * S) * kThrThreads * sizeof(double);
  // a function attribute belongs to the device it was set on (one slot per device: api.cu)
  SS_CUDA(cudaFuncSetAttribute(loglik_thr_kernel<NB, S, QOFF, ZTV, EPS, LS, ZS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long blocks = (A.E + kThrThreads - 1) / kThrThreads;
  SS_ARG(blocks < (1LL << 31), "thread kernel: batch too large for one launch");
  loglik_thr_kernel<NB, S, QOFF, ZTV, EPS, LS, ZS><<<(unsigned)blocks, kThrThreads, smem, st>>>(A);
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

#ifndef SSGPU_THR_MIN_NB
#define SSGPU_THR_MIN_NB 2
#endif
bool thr_supported(int nb, int p, bool ztv) { (void)ztv; return p == 1 && nb >= SSGPU_THR_MIN_NB && nb <= kThrMaxNB; }

// y, out: of the first series of the launch; hd / hi: filled by the lane kernel for the same E evaluations
// eps: position 1 of the last block of the plan is the residual state (see the kernel), time-invariant Z only
int launch_loglik_thr(const BlkPlan& pl, const double* z_internal, const double* zt, const double* y, long long B,
                      long long E, long long ldy, int N, const double* hd, const int* hi, double* out, cudaStream_t st,
                      bool eps, bool ls, bool zs, const ThrDiffuse* ud) {
  ThrArgs A{};
  if (ud) A.ud = *ud;
  A.zt = zt;
  A.y = y;
  A.B = B;
  A.E = E;
  A.ldy = ldy;
  A.N = N;
  A.hd = hd;
  A.hi = hi;
  A.out = out;
  for (int I = 0; I < pl.nb; I++) {
    for (int k = 0; k < 4; k++) A.cst[I][k] = pl.Tb[I][k];
    A.cst[I][4] = z_internal[2 * I];
    A.cst[I][5] = z_internal[2 * I + 1];
  }
  const bool qoff = pl.q_offdiag != 0;
  SS_ARG(!((eps || ls || zs) && zt), "thread kernel: the structural shortcuts need a time-invariant Z");
  zs = zs && eps;  // instantiated together with the residual shortcut only (the reference's layout always has the residual)
#define SS_THR_CASE(NBV)                                                                                                  \
  case NBV:                                                                                                               \
    if (eps && ls && zs)                                                                                                  \
      return qoff ? launch_thr_one<NBV, true, false, true, true, true>(A, st)                                             \
                  : launch_thr_one<NBV, false, false, true, true, true>(A, st);                                           \
    if (eps && zs)                                                                                                        \
      return qoff ? launch_thr_one<NBV, true, false, true, false, true>(A, st)                                            \
                  : launch_thr_one<NBV, false, false, true, false, true>(A, st);                                          \
    if (eps && ls)                                                                                                        \
      return qoff ? launch_thr_one<NBV, true, false, true, true>(A, st) : launch_thr_one<NBV, false, false, true, true>(A, st); \
    if (ls) return qoff ? launch_thr_one<NBV, true, false, false, true>(A, st) : launch_thr_one<NBV, false, false, false, true>(A, st); \
    if (eps) return qoff ? launch_thr_one<NBV, true, false, true>(A, st) : launch_thr_one<NBV, false, false, true>(A, st); \
    return zt ? (qoff ? launch_thr_one<NBV, true, true>(A, st) : launch_thr_one<NBV, false, true>(A, st))                 \
              : (qoff ? launch_thr_one<NBV, true>(A, st) : launch_thr_one<NBV, false>(A, st));
  switch (pl.nb) {
    SS_THR_CASE(2) SS_THR_CASE(3) SS_THR_CASE(4) SS_THR_CASE(5) SS_THR_CASE(6) SS_THR_CASE(7) SS_THR_CASE(8)
  }
#undef SS_THR_CASE
  SS_ARG(false, "thread kernel: 2 to 8 blocks");
}

}  // namespace ssgpu
