// Batched loglikelihood for models whose transition matrix is block diagonal with blocks of at most 2 x 2:
// structure-exploiting DFMA kernel, L lanes per evaluation, the SYMMETRIC covariance distributed over the lanes in
// circulant block ownership.
//
// Restates the recursion of src/LogLikC.cpp:68-214 (univariate treatment, exact diffuse initialisation,
// NaN = missing) for models whose Z, T, R are shared by the batch, T and R time-invariant, Z time-invariant or -- with
// p = 1 -- one row per time point (explanatory variables) (per-evaluation a, P_inf, P_star, Q allowed) and whose T,
// after re-ordering the states, is  blkdiag(T_0, ..., T_{NB-1})  with 1 x 1 or 2 x 2 blocks, NB <= 8 (m <= 16;
// NB <= 16, m <= 32 when p = 1), and whose R Q R' has no entry outside those blocks.  That is every structural model
// statespacer builds from residuals, level, slope, trigonometric seasonals, cycles and regression coefficients
// (R/GetSysMat.R:1388-1426, R/Components-Unobserved.R:143-159,491-519): T couples a state with at most one partner.
// BASELINE.json config 5 (m = 14) is 7 such blocks.
//
// Why this shape.  With P cut into 2 x 2 blocks P_IJ the time update  P <- T P T' + R Q R'  (src/LogLikC.cpp:202)
// is  P_IJ <- T_I P_IJ T_J' (+ (RQR')_II)  block by block: 16 multiply-adds per block, 4 m^2 per step instead of the
// 4 m^3 of the dense product, and it needs NO data from any other block.  P is symmetric, so only the blocks on and
// above the diagonal are kept: lane I of a group of L lanes owns  P_{I,(I+k) mod NB},  k = 0 .. H = NB / 2  -- the
// diagonal block and the next H blocks of block row I, wrapping around.  Every unordered pair {I, J} has exactly one
// owner (NB even: the pairs at distance NB / 2 would have two; the lanes I >= NB / 2 keep theirs at zero).  That is
// H + 1 blocks (4 of the 7 x 7 at config 5: 32 registers instead of 56) and 16 (H + 1) multiply-adds of time update
// per lane, all from the lane's own registers; T_I, T_{I+k} and z_{I+k} are per-lane constants in registers.
// Only the measurement update couples the lanes:  M = P z  needs, for row I, the own blocks times z_{I+k} (local) and
// the transposed blocks of the H lanes behind times THEIR z (H shuffled 2-vectors);  F = z'M = sum_I z_I'(2 m_I - u_I)
// (m_I the local part of M_I, u_I = P_II z_I) is formed from local data, so its xor-butterfly over the group starts
// before the exchange of M;  P_IJ -= (M_I / F) M_J'  needs M_{I+k} from the H lanes ahead (H shuffled 2-vectors).
// 32 / L evaluations share a warp; the exact diffuse phase (about d of the N steps) runs in a select-based loop that
// all groups of the warp execute until the last of them has left it, with P_inf parked in shared memory so that the
// registers are sized for the regular phase.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "blk_handoff.cuh"
#include "common.cuh"

namespace ssgpu {

constexpr int kBlkThreads = 128;   // CTA of the register-constant instantiations (NB <= 8)
constexpr int kBlkThreadsCS = 64;  // NB > 8
constexpr int kBlkMaxNB = 16;  // blocks: 1..8 with T_J and z in registers, 9..16 (p = 1) with both in shared memory
constexpr int kBlkMaxP = 8;

// resident CTAs per SM the register budget is sized for: 7 and 8 blocks (5 own blocks at most) 3 x 128 threads at
// 168 registers, fewer blocks 4 x 128 at 128 registers, more than 8 blocks 4 x 64 at 255
#ifndef SSGPU_BLK_MINB
#define SSGPU_BLK_MINB 3
#endif
constexpr int blk_minb(int nb) { return nb > 8 ? 4 : nb > 4 ? SSGPU_BLK_MINB : 4; }

struct BlkArgs {
  DModel md;
  const double* y;  // [(t*p + j) * ldy + b]
  long long B;      // series of this launch
  long long E;      // B * V
  long long ldy;    // series of the whole batch (a launch may cover a range of them)
  double* out;      // [v * ldy + b]
  double* hd;       // hand-over to the thread kernel (blk_handoff.cuh), null = run to the end
  int* hi;
  signed char perm[2 * kBlkMaxNB];     // internal position 2I+e holds state perm[2I+e] of the caller, -1 = padding
  double Tb[kBlkMaxNB][4];             // T_J[row][col] at [2*row + col], internal order
  double z[kBlkMaxP][2 * kBlkMaxNB];   // row j of Z, internal order
  const double* zt;                    // time-varying Z (p = 1): [t][2 * kBlkMaxNB] in internal order, device memory
  int self_setup;                      // ... and its record is written by loglik_thr_setup_kernel, not here
  int du;                              // hand-over only: an evaluation without a missing value among its first du observations
                                       // is handed over as it stands at t = 0, marked -1 (the thread kernel runs its diffuse steps)
};

__device__ __forceinline__ double blk_rcp(double x) {
  // MUFU.RCP64H seed and two Newton steps (x is finite, normal and >= 1e-7 here)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// ---- bulk-asynchronous (TMA) staging of per-step parameter rows: cp.async.bulk + mbarrier ------------------------------
// Time-varying Z (explanatory variables, R/Components-Regressors.R:60-78) is one 256-byte row per time point, the same
// for every evaluation.  Each warp keeps a ring of kBlkZStages rows in shared memory; lane 0 arms the stage's mbarrier
// with the byte count and issues ONE bulk copy for the row kBlkZStages steps ahead (the copy engine, not the LSU, moves
// it and signals the barrier), the lanes wait on the barrier's phase parity and read their z_{I+k} from shared memory.
// One copy per row costs more than it saves (arm + issue + wait every step: measured 4.3e9 against 7.5e9
// series*timesteps/s with plain loads on the m = 16 regression model), so a stage is a block of kBlkZRows rows: one copy,
// one wait and one refill per kBlkZRows steps.
constexpr int kBlkZStages = 2;
constexpr int kBlkZRows = 8;
constexpr int kBlkZRowBytes = 2 * kBlkMaxNB * (int)sizeof(double);
// Ring of observations (YTMA): per warp kBlkYStages stages of kBlkYRows time points; a row holds the y of the warp's
// series (at most 32, contiguous in the time-major array) widened to 16-byte alignment: 32 + 2 doubles.  32 time points of
// look-ahead: a DRAM access is ~800 cycles, a step of the local level model ~100.
constexpr int kBlkYStages = 4;
constexpr int kBlkYRows = 8;
constexpr int kBlkYRowDoubles = 34;
__device__ __forceinline__ unsigned blk_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void blk_bar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(blk_smem_u32(bar)));
}
__device__ __forceinline__ void blk_bar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void blk_bulk_rows(double* dst, const double* src, int bytes, unsigned long long* bar) {
  const unsigned b = blk_smem_u32(bar), d = blk_smem_u32(dst);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
               "r"(bytes), "r"(b)
               : "memory");
}
__device__ __forceinline__ void blk_bar_wait(unsigned long long* bar, unsigned parity) {
  const unsigned b = blk_smem_u32(bar);
  unsigned ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(b), "r"(parity)
                 : "memory");
  } while (!ok);
}

__device__ __forceinline__ double2 blk_shfl(double2 x, int src) {
  return make_double2(__shfl_sync(0xffffffffu, x.x, src), __shfl_sync(0xffffffffu, x.y, src));
}

// QOFF: the diagonal blocks of R Q R' have off-diagonal entries (correlated disturbances of a multivariate model)
// ZTV: Z changes with t (explanatory variables, R/Components-Regressors.R:60-78); p = 1 only
// HO: leave after the diffuse phase and hand the state over (A.hd, A.hi) to the thread kernel
// ZTMA: the rows of a time-varying Z arrive through the cp.async.bulk ring instead of per-lane loads (SSGPU_BLK_ZTMA=1)
// YTMA: the observations of the warp's series arrive through a cp.async.bulk ring (one-lane-per-series models: the step is
// ~60 instructions and the kernel otherwise waits for DRAM)
template <int L, int NB, bool P1, bool QOFF, bool ZTV = false, bool HO = false, bool ZTMA = false, bool YTMA = false>
__global__ void __launch_bounds__(NB > 8 ? kBlkThreadsCS : kBlkThreads, blk_minb(NB)) loglik_blk_kernel(const BlkArgs A) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / L;       // evaluations per warp
  constexpr int H = NB / 2;       // own blocks: (I, I + k mod NB), k = 0 .. H
  constexpr int K = H + 1;
  constexpr bool EVEN = NB > 1 && NB % 2 == 0;  // the pairs at distance H have two candidate owners: lanes I < H keep them
  // NB > 8 (17 <= m <= 32, p = 1): T_J comes from shared memory (per-lane 8-byte loads) instead of 4 H registers pairs
  constexpr bool CS = NB > 8;
  constexpr bool ZREG = P1 && !ZTV && !CS;  // z_{I+k} in registers for the whole kernel
  constexpr int BT = CS ? kBlkThreadsCS : kBlkThreads;
  static_assert(!CS || P1, "more than 8 blocks: p = 1 only");
  static_assert(NB <= L, "one lane per block row");
  const DModel& md = A.md;
  const int N = md.N, p = P1 ? 1 : md.p, M = md.m, r = md.r;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int I = lane % L, g = lane / L;
  __shared__ double tb_s[CS ? NB * 4 : 1];
  __shared__ double2 zs_s[ZREG || ZTV ? 1 : kBlkMaxP * NB];  // rows of Z, internal order, as block pairs
  if constexpr (CS) {
    for (int e = threadIdx.x; e < NB * 4; e += BT) tb_s[e] = (&A.Tb[0][0])[e];
  }
  if constexpr (!(ZREG || ZTV)) {
    for (int e = threadIdx.x; e < p * NB; e += BT) {
      const int j = e / NB, J = e % NB;
      zs_s[e] = make_double2(A.z[j][2 * J], A.z[j][2 * J + 1]);
    }
  }
  if constexpr (CS || !(ZREG || ZTV)) __syncthreads();  // before any warp leaves
  // evaluations are numbered series-major (w = b*V + v) so that the V parameter vectors of a series run side by
  // side and share its y through L1/L2; results keep the public order e = v*B + b
  const long long w0 = ((long long)blockIdx.x * (BT / 32) + warp) * G;
  if (w0 >= A.E) return;  // whole warp
  // every lane stays alive (shuffles): groups beyond the last evaluation shadow it and do not store
  bool live = w0 + g < A.E;
  const long long w_id = live ? w0 + g : A.E - 1;
  const long long V = A.E / A.B, b = w_id / V, v = w_id % V;
  const long long e_id = v * A.ldy + b;
  const bool own = I < NB;  // L > NB: the surplus lanes carry zero blocks
  // Hand-over instantiation: an evaluation that has no missing value among its first du observations does not need the
  // lanes -- its diffuse steps are the shared ones of the thread kernel (loglik_thr.cu).  It is marked -1; when the
  // record of such evaluations is written from the model by loglik_thr_setup_kernel (self_setup) nothing else is left to do here,
  // and a warp made of such evaluations only leaves right away.
  bool handed = false;
  if constexpr (HO) {
    if (A.du > 0) {
      bool clean = true;
      for (int tt = I; tt < A.du; tt += L) clean = clean && !isnan(A.y[(long long)tt * A.ldy + b]);
      const unsigned bal = __ballot_sync(FULL, clean);
      const unsigned gm = (L == 32 ? FULL : ((1u << L) - 1u)) << (g * L);
      handed = (bal & gm) == gm;
      if (A.self_setup) {
        if (handed && live && I == 0) A.hi[w_id] = -1;
        if (__all_sync(FULL, handed)) return;
      }
    }
  }
  // block column of the k-th own block; the lane it is shared with; the lane that shares ITS k-th block with this one
  int Jk[K], srcF[K], srcB[K];
  bool keep[K];
#pragma unroll
  for (int k = 0; k < K; k++) {
    const int jf = I + k < NB ? I + k : I + k - NB, jb = I - k >= 0 ? I - k : I - k + NB;
    Jk[k] = own ? jf : 0;
    srcF[k] = own ? lane - I + jf : lane;
    srcB[k] = own ? lane - I + jb : lane;
    keep[k] = own && !(EVEN && k == H && I >= H);
    // opaque to the optimiser: otherwise ptxas recomputes the wrapped lane numbers inside the hot loop
    asm volatile("" : "+r"(srcF[k]), "+r"(srcB[k]));
  }

  // P_inf: parked in shared memory [element][thread] so that the registers are sized for the regular phase -- except in
  // the hand-over instantiation, which ends with the diffuse phase and keeps it in registers (its launch was LSU-bound
  // on this traffic: 72 % of the wavefront rate, FP64 pipe 47 %)
  __shared__ double pinf_s[HO ? 1 : K * 4][HO ? 1 : BT];
  double pinf_r[HO ? K * 4 : 1];
  const int tx = threadIdx.x;
#define PINF(e) (*(HO ? &pinf_r[HO ? (e) : 0] : &pinf_s[HO ? 0 : (e)][HO ? 0 : tx]))
  // ring of blocks of rows of the time-varying Z, one ring per warp (see blk_bulk_rows above)
  __shared__ __align__(128) double zring_s[ZTMA ? BT / 32 : 1][ZTMA ? kBlkZStages : 1][ZTMA ? kBlkZRows : 1][2 * kBlkMaxNB];
  __shared__ __align__(8) unsigned long long zbar_s[ZTMA ? BT / 32 : 1][ZTMA ? kBlkZStages : 1];
  int zblk = -1;  // block of rows the lanes have waited for
  auto z_issue = [&](int blk) {  // lane 0: rows blk * kBlkZRows ... into stage blk % kBlkZStages
    const int row0 = blk * kBlkZRows, nr = min(kBlkZRows, N - row0);
    if (nr > 0)
      blk_bulk_rows(&zring_s[warp][blk % kBlkZStages][0][0], A.zt + (size_t)row0 * (2 * kBlkMaxNB), nr * kBlkZRowBytes,
                    &zbar_s[warp][blk % kBlkZStages]);
  };
  if constexpr (ZTMA) {
    if (lane == 0) {
#pragma unroll
      for (int sg = 0; sg < kBlkZStages; sg++) blk_bar_init(&zbar_s[warp][sg]);
      blk_bar_fence_init();
#pragma unroll
      for (int sg = 0; sg < kBlkZStages; sg++) z_issue(sg);
    }
    __syncwarp();
  }
  // block b + kBlkZStages takes the place of block b once every lane has used the last row of b (call at the end of step t)
  auto z_refill = [&](int t) {
    if constexpr (ZTMA) {
      if ((t & (kBlkZRows - 1)) == kBlkZRows - 1) {
        __syncwarp();
        if (lane == 0) z_issue(t / kBlkZRows + kBlkZStages);
      }
    }
  };
  // ring of observations, one per warp (YTMA; p = 1).  The series of the warp's evaluations are b_lo .. b_hi, contiguous
  // in every row of y; the copy starts at the even index below b_lo and has an even length (16-byte rule of
  // cp.async.bulk; the launcher has checked that ldy is even and y 16-byte aligned, so no row is overrun).
  __shared__ __align__(128) double yring_s[YTMA ? BT / 32 : 1][YTMA ? kBlkYStages : 1][YTMA ? kBlkYRows : 1][YTMA ? kBlkYRowDoubles : 1];
  __shared__ __align__(8) unsigned long long ybar_s[YTMA ? BT / 32 : 1][YTMA ? kBlkYStages : 1];
  int yblk = -1, ycol = 0;
  long long ybase = 0;
  int ybytes = 0;
  auto y_issue = [&](int blk) {  // lane 0: rows blk * kBlkYRows ... of the warp's column block into stage blk % kBlkYStages
    const int row0 = blk * kBlkYRows, nr = min(kBlkYRows, N - row0);
    if (nr <= 0) return;
    unsigned long long* bar = &ybar_s[warp][blk % kBlkYStages];
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(blk_smem_u32(bar)), "r"(nr * ybytes) : "memory");
    for (int rr = 0; rr < nr; rr++) {
      const unsigned d = blk_smem_u32(&yring_s[warp][blk % kBlkYStages][rr][0]);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d),
                   "l"(A.y + ybase + (long long)(row0 + rr) * A.ldy), "r"(ybytes), "r"(blk_smem_u32(bar))
                   : "memory");
    }
  };
  if constexpr (YTMA) {
    const long long b_lo = __shfl_sync(FULL, b, 0), b_hi = __shfl_sync(FULL, b, 31);  // evaluations are series-major: b ascends
    ybase = b_lo & ~1LL;
    ybytes = (int)(((b_hi - ybase + 1) + 1) & ~1LL) * (int)sizeof(double);
    ycol = (int)(b - ybase);
    if (lane == 0) {
#pragma unroll
      for (int sg = 0; sg < kBlkYStages; sg++) blk_bar_init(&ybar_s[warp][sg]);
      blk_bar_fence_init();
#pragma unroll
      for (int sg = 0; sg < kBlkYStages; sg++) y_issue(sg);
    }
    __syncwarp();
  }
  // y of time point t (every t exactly once, ascending): from the ring, or from device memory
  auto y_get = [&](int t) {
    if constexpr (YTMA) {
      const int blk = t / kBlkYRows, sg = blk % kBlkYStages;
      if (blk != yblk) {
        blk_bar_wait(&ybar_s[warp][sg], (unsigned)(blk / kBlkYStages) & 1u);
        yblk = blk;
      }
      return yring_s[warp][sg][t & (kBlkYRows - 1)][ycol];
    } else {
      return A.y[(long long)t * A.ldy + b];
    }
  };
  auto y_refill = [&](int t) {  // at the end of step t: the block after the last row of a stage has been used
    if constexpr (YTMA) {
      if ((t & (kBlkYRows - 1)) == kBlkYRows - 1) {
        __syncwarp();
        if (lane == 0) y_issue(t / kBlkYRows + kBlkYStages);
      }
    }
  };

  // ---- own blocks of P_star, P_inf, own T blocks, own part of a, own diagonal block of R Q R' ---------------------
  const int i0 = own ? (int)A.perm[2 * I] : -1, i1 = own ? (int)A.perm[2 * I + 1] : -1;
  auto elem = [&](const double* blk, int diag, int i, int k) { return (i >= 0 && k >= 0) ? mat_get(blk, diag, M, i, k) : 0.0; };
  double P[K][4];
  {
    const double* Ps = md.P_star.block(b, v, 0);
    const double* Pi = md.P_inf.block(b, v, 0);
#pragma unroll
    for (int k = 0; k < K; k++) {
      const int k0 = keep[k] ? (int)A.perm[2 * Jk[k]] : -1, k1 = keep[k] ? (int)A.perm[2 * Jk[k] + 1] : -1;
      P[k][0] = elem(Ps, md.P_star.diag, i0, k0);
      P[k][1] = elem(Ps, md.P_star.diag, i0, k1);
      P[k][2] = elem(Ps, md.P_star.diag, i1, k0);
      P[k][3] = elem(Ps, md.P_star.diag, i1, k1);
      PINF(k * 4 + 0) = elem(Pi, md.P_inf.diag, i0, k0);
      PINF(k * 4 + 1) = elem(Pi, md.P_inf.diag, i0, k1);
      PINF(k * 4 + 2) = elem(Pi, md.P_inf.diag, i1, k0);
      PINF(k * 4 + 3) = elem(Pi, md.P_inf.diag, i1, k1);
    }
  }
  double a0, a1;
  {
    const double* ab = md.a.block(b, v, 0);
    a0 = i0 >= 0 ? ab[i0] : 0.0;
    a1 = i1 >= 0 ? ab[i1] : 0.0;
  }
  const int Iz = own ? I : 0;
  const double t00 = own ? A.Tb[Iz][0] : 0.0, t01 = own ? A.Tb[Iz][1] : 0.0, t10 = own ? A.Tb[Iz][2] : 0.0,
               t11 = own ? A.Tb[Iz][3] : 0.0;
  // T_{I+k}, k = 1 .. H: per-lane constants in vector registers for the whole kernel (NB <= 8)
  double TJ[CS ? 1 : K][4];
  if constexpr (!CS) {
#pragma unroll
    for (int k = 1; k < K; k++) {
#pragma unroll
      for (int e = 0; e < 4; e++) TJ[k][e] = A.Tb[Jk[k]][e];
    }
  }
  // (R Q R')[i][c] for the own diagonal block (src/LogLikC.cpp:202, hoisted: Q and R are time-invariant here; the
  // host has checked that R Q R' has no entry outside the diagonal blocks)
  double q0, q1, q2, q3;
  {
    const double* Qb = md.Q.block(b, v, 0);
    const double* Rb = md.R.block(0, 0, 0);
    auto rqr = [&](int i, int c) {
      double acc = 0.0;
      if (i < 0 || c < 0) return acc;
      for (int k = 0; k < r; k++) {
        const double rik = mat_get(Rb, md.R.diag, M, i, k);
        if (rik == 0.0) continue;
        double qr;
        if (md.Q.diag) {
          qr = Qb[k] * mat_get(Rb, md.R.diag, M, c, k);
        } else {
          qr = 0.0;
          for (int l = 0; l < r; l++) qr = fma(Qb[k + (long long)l * r], mat_get(Rb, md.R.diag, M, c, l), qr);
        }
        acc = fma(rik, qr, acc);
      }
      return acc;
    };
    q0 = rqr(i0, i0);
    q1 = QOFF ? rqr(i0, i1) : 0.0;
    q2 = QOFF ? rqr(i1, i0) : 0.0;
    q3 = rqr(i1, i1);
  }

  auto group_sum = [&](double x) {
#pragma unroll
    for (int o = 1; o < L; o <<= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
  };
  auto group_all = [&](bool ok) {
    const unsigned bal = __ballot_sync(FULL, ok);
    const unsigned gm = (L == 32 ? FULL : ((1u << L) - 1u)) << (g * L);
    return (bal & gm) == gm;
  };

  // loglik = n_terms*const - (sum log F + sum v^2/F)/2, log F accumulated as mantissa * 2^esum
  double prod = 1.0, sumq = 0.0;
  int esum = 0, n_terms = 0, non_init = 0, F_eq0 = 0;
  auto acc_log = [&](double F) {
    prod *= F;
    const int hi = __double2hiint(prod);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    esum += ex;
    prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
  };

  // z_{I+k} of observation j at time t: registers (p = 1, time-invariant), the row of the step in device memory
  // (time-varying, L1-resident: the same few hundred bytes for every evaluation) or shared memory (p > 1, NB > 8)
  double2 ZK[K];
  if constexpr (ZREG) {
#pragma unroll
    for (int k = 0; k < K; k++) ZK[k] = own ? make_double2(A.z[0][2 * Jk[k]], A.z[0][2 * Jk[k] + 1]) : make_double2(0.0, 0.0);
  }
  auto load_z = [&](int t, int j) {
    if constexpr (ZTV && !ZTMA) {  // default: the row of the step from device memory (L1-resident, per-lane 16-byte loads)
      const double2* zr = reinterpret_cast<const double2*>(A.zt + (size_t)t * (2 * kBlkMaxNB));
#pragma unroll
      for (int k = 0; k < K; k++) ZK[k] = own ? zr[Jk[k]] : make_double2(0.0, 0.0);
    } else if constexpr (ZTV) {
      const int blk = t / kBlkZRows, sg = blk % kBlkZStages;
      if (blk != zblk) {  // first row of a block: its copy has landed when the stage's barrier has flipped
        blk_bar_wait(&zbar_s[warp][sg], (unsigned)(blk / kBlkZStages) & 1u);
        zblk = blk;
      }
      const double2* zr = reinterpret_cast<const double2*>(zring_s[warp][sg][t & (kBlkZRows - 1)]);
#pragma unroll
      for (int k = 0; k < K; k++) ZK[k] = own ? zr[Jk[k]] : make_double2(0.0, 0.0);
    } else if constexpr (!ZREG) {
#pragma unroll
      for (int k = 0; k < K; k++) ZK[k] = own ? zs_s[j * NB + Jk[k]] : make_double2(0.0, 0.0);
    }
  };

  // local part of M = X z for a symmetric X held like P (X[k][e] = element e of the k-th own block):
  //   ml = sum_k X_{I,I+k} z_{I+k}  (stays in row I),   u0 = X_II z_I,   c[k] = X_{I,I+k}' z_I  (belongs to row I+k)
  auto products = [&](const double (&X)[K][4], double2& ml, double2& u0, double2 (&c)[K]) {
    const double2 zi = ZK[0];
    u0.x = fma(X[0][1], zi.y, X[0][0] * zi.x);
    u0.y = fma(X[0][3], zi.y, X[0][2] * zi.x);
    ml = u0;
#pragma unroll
    for (int k = 1; k < K; k++) {
      const double2 zk = ZK[k];
      ml.x = fma(X[k][1], zk.y, fma(X[k][0], zk.x, ml.x));
      ml.y = fma(X[k][3], zk.y, fma(X[k][2], zk.x, ml.y));
      c[k].x = fma(X[k][2], zi.y, X[k][0] * zi.x);
      c[k].y = fma(X[k][3], zi.y, X[k][1] * zi.x);
    }
  };
  // M_I = ml + the parts the H lanes behind hold for row I
  auto complete = [&](double2 ml, const double2 (&c)[K]) {
#pragma unroll
    for (int k = 1; k < K; k++) {
      const double2 ci = blk_shfl(c[k], srcB[k]);
      ml.x += ci.x;
      ml.y += ci.y;
    }
    return ml;
  };
  // x_{I+k}, k = 0 .. H, of a 2-vector per block row (zero where the k-th block is not kept)
  auto ahead = [&](double2 x, double2 (&xk)[K]) {
    xk[0] = x;
#pragma unroll
    for (int k = 1; k < K; k++) {
      const double2 f = blk_shfl(x, srcF[k]);
      xk[k] = (EVEN && k == H && !keep[k]) ? make_double2(0.0, 0.0) : f;
    }
  };

  // P_{I,I+k} <- T_I P_{I,I+k} T_{I+k}' (+ the own diagonal block of R Q R' when k == 0); X[r][c] = sum_l Y[r][l] T_J[c][l]
  auto time_block = [&](int k, double& p0, double& p1, double& p2, double& p3, bool with_q) {
    const double y00 = fma(t01, p2, t00 * p0), y01 = fma(t01, p3, t00 * p1);
    const double y10 = fma(t11, p2, t10 * p0), y11 = fma(t11, p3, t10 * p1);
    if (k == 0) {
      if (with_q) {
        p0 = fma(y01, t01, fma(y00, t00, q0));
        p1 = QOFF ? fma(y01, t11, fma(y00, t10, q1)) : fma(y01, t11, y00 * t10);
        p2 = QOFF ? fma(y11, t01, fma(y10, t00, q2)) : fma(y11, t01, y10 * t00);
        p3 = fma(y11, t11, fma(y10, t10, q3));
      } else {
        p0 = fma(y01, t01, y00 * t00);
        p1 = fma(y01, t11, y00 * t10);
        p2 = fma(y11, t01, y10 * t00);
        p3 = fma(y11, t11, y10 * t10);
      }
    } else {
      double tj0, tj1, tj2, tj3;
      if constexpr (CS) {
        const double* tp = tb_s + Jk[k] * 4;
        tj0 = tp[0], tj1 = tp[1], tj2 = tp[2], tj3 = tp[3];
      } else {
        tj0 = TJ[CS ? 0 : k][0], tj1 = TJ[CS ? 0 : k][1], tj2 = TJ[CS ? 0 : k][2], tj3 = TJ[CS ? 0 : k][3];
      }
      p0 = fma(y01, tj1, y00 * tj0);
      p1 = fma(y01, tj3, y00 * tj2);
      p2 = fma(y11, tj1, y10 * tj0);
      p3 = fma(y11, tj3, y10 * tj2);
    }
  };

  // ---- hand-over (HO): the state at the start of time point `tmark` goes to the record of blk_handoff.cuh
  auto hand_over = [&](int tmark) {
    constexpr BlkHandoff HL(NB, QOFF);
    if (live && own) {
      double* const hd = A.hd + w_id;
      const long long E = A.E;
      hd[(HL.a() + 2 * I) * E] = a0;
      hd[(HL.a() + 2 * I + 1) * E] = a1;
      hd[(HL.q() + HL.qw * I) * E] = q0;
      if constexpr (QOFF) hd[(HL.q() + HL.qw * I + 1) * E] = q1;
      hd[(HL.q() + HL.qw * I + HL.qw - 1) * E] = q3;
      hd[(HL.d() + 3 * I) * E] = P[0][0];
      hd[(HL.d() + 3 * I + 1) * E] = P[0][1];
      hd[(HL.d() + 3 * I + 2) * E] = P[0][3];
#pragma unroll
      for (int k = 1; k < K; k++) {
        if (!keep[k]) continue;
        const int J = Jk[k];
        const bool up = J > I;  // held as (I, J): as is; as (I, J) with J < I: the transpose of block (J, I)
        const int oi = up ? HL.off_index(I, J) : HL.off_index(J, I);
        double* const ho = hd + (long long)(HL.o() + 4 * oi) * E;
        ho[0] = P[k][0];
        ho[E] = up ? P[k][1] : P[k][2];
        ho[2 * E] = up ? P[k][2] : P[k][1];
        ho[3 * E] = P[k][3];
      }
      if (I == 0) {
        hd[HL.s() * E] = prod;
        hd[(HL.s() + 1) * E] = sumq;
        int* const hi = A.hi + w_id;
        hi[0] = tmark;
        hi[E] = esum;
        hi[2 * E] = n_terms;
        hi[3 * E] = non_init;
        hi[4 * E] = F_eq0;
      }
    }
  };

  const long long Np = (long long)N * p;
  int t = 0;
  // a marked evaluation (see the top of the kernel): handed over as it stands unless the thread kernel sets it up itself;
  // its group then idles through the loop below while another group of its warp needs it
  if constexpr (HO) {
    if (handed) {
      if (!A.self_setup) hand_over(-1);
      live = false;
    }
  }
  {
    // ---- exact diffuse phase (src/LogLikC.cpp:98-159), select-based: the groups of a warp may leave it at
    // different time points (missing observations), so every group runs this loop until the last one is done ----
    bool small = true;
#pragma unroll
    for (int e = 0; e < K * 4; e++) small = small && (fabs(PINF(e)) < kEps);
    bool init = !group_all(small) && !handed;  // :49
    while (t < N && __any_sync(FULL, init)) {
      for (int j = 0; j < p; j++) {
        const double yv = YTMA ? y_get(t) : A.y[((long long)t * p + j) * A.ldy + b];
        const bool obs = !isnan(yv);  // :88-90
        load_z(t, j);
        const double2 zi = ZK[0];
        double2 ml, u0, cs[K], ci[K];
        products(P, ml, u0, cs);
        const double2 ms = complete(ml, cs);
        double Pi[K][4];
#pragma unroll
        for (int e = 0; e < K * 4; e++) Pi[e / 4][e % 4] = PINF(e);
        products(Pi, ml, u0, ci);
        const double2 mi = complete(ml, ci);
        const double F_star = group_sum(fma(zi.y, ms.y, zi.x * ms.x));
        const double F_inf = group_sum(fma(zi.y, mi.y, zi.x * mi.x));
        const double za = group_sum(fma(zi.y, a1, zi.x * a0));
        // the three outcomes of :109-153 and the regular step of :162-195 (a group whose P_inf has vanished)
        const bool dif = init && obs && !(F_inf < kEps);
        const bool star = obs && !dif && !(F_star < kEps);
        non_init += obs && !init;
        F_eq0 += obs && !init && !star;
        const double F1i = dif ? 1.0 / F_inf : 0.0, F2 = dif ? -(F1i * F1i) * F_star : 0.0;
        const double F1s = star ? 1.0 / F_star : 0.0;
        const double vv = (dif || star) ? yv - za : 0.0;
        if (dif) acc_log(F_inf);  // no v^2 term (:153)
        if (star) acc_log(F_star);
        n_terms += dif || star;
        sumq = fma(vv * vv, F1s, sumq);
        // K0 = M_inf F1 (diffuse) or M_star F1 (star); K1 = M_star F1 + M_inf F2 (diffuse only)
        // P*_IJ -= M*_I K0_J' + M_inf,I K1_J';  P_inf,IJ -= M_inf,I K0_J' (diffuse only)
        double2 msk[K], mik[K];
        ahead(ms, msk);
        ahead(mi, mik);
#pragma unroll
        for (int k = 0; k < K; k++) {
          const double2 sJ = msk[k], iJ = mik[k];
          const double k00 = fma(F1s, sJ.x, F1i * iJ.x), k01 = fma(F1s, sJ.y, F1i * iJ.y);    // K0_J
          const double k10 = fma(F2, iJ.x, F1i * sJ.x), k11 = fma(F2, iJ.y, F1i * sJ.y);      // K1_J
          P[k][0] = fma(-mi.x, k10, fma(-ms.x, k00, P[k][0]));
          P[k][1] = fma(-mi.x, k11, fma(-ms.x, k01, P[k][1]));
          P[k][2] = fma(-mi.y, k10, fma(-ms.y, k00, P[k][2]));
          P[k][3] = fma(-mi.y, k11, fma(-ms.y, k01, P[k][3]));
          const double c0 = F1i * iJ.x, c1 = F1i * iJ.y;
          PINF(k * 4 + 0) = fma(-mi.x, c0, PINF(k * 4 + 0));
          PINF(k * 4 + 1) = fma(-mi.x, c1, PINF(k * 4 + 1));
          PINF(k * 4 + 2) = fma(-mi.y, c0, PINF(k * 4 + 2));
          PINF(k * 4 + 3) = fma(-mi.y, c1, PINF(k * 4 + 3));
        }
        a0 = fma(fma(F1s, ms.x, F1i * mi.x), vv, a0);
        a1 = fma(fma(F1s, ms.y, F1i * mi.y), vv, a1);
        if (j < p - 1) {  // :157-159 (only after a diffuse update)
          small = true;
#pragma unroll
          for (int e = 0; e < K * 4; e++) small = small && (fabs(PINF(e)) < kEps);
          const bool gone = group_all(small);
          if (dif && gone) init = false;
        }
      }
      if (t < N - 1) {  // :200-207
        small = true;
#pragma unroll
        for (int k = 0; k < K; k++) {
          time_block(k, P[k][0], P[k][1], P[k][2], P[k][3], true);
          double i0v = PINF(k * 4 + 0), i1v = PINF(k * 4 + 1), i2v = PINF(k * 4 + 2), i3v = PINF(k * 4 + 3);
          time_block(k, i0v, i1v, i2v, i3v, false);
          PINF(k * 4 + 0) = i0v;
          PINF(k * 4 + 1) = i1v;
          PINF(k * 4 + 2) = i2v;
          PINF(k * 4 + 3) = i3v;
          small = small && (fabs(i0v) < kEps) && (fabs(i1v) < kEps) && (fabs(i2v) < kEps) && (fabs(i3v) < kEps);
        }
        const double na0 = fma(t01, a1, t00 * a0), na1 = fma(t11, a1, t10 * a0);
        a0 = na0;
        a1 = na1;
        const bool gone = group_all(small);
        if (init && gone) init = false;
      }
      z_refill(t);
      y_refill(t);
      t++;
    }
  }

  if constexpr (HO) {
    // every group of the warp is through the diffuse phase (or at t == N).  An evaluation that is already complete
    // (t == N) writes its result below as usual and the thread kernel skips it.
    hand_over(t);
    if (t < N) return;
  }

  // ---- regular phase: the hot loop (src/LogLikC.cpp:162-207), branch-free ------------------------------------
  // A missing observation or F < 1e-7 (:88-90, :173-176) makes the update a no-op through 1/F := v := 0.
  // loglik terms of one observation; returns 1/F and v of the update (both 0 for a no-op)
  double F1, vv;
  auto observe = [&](double yv, double F_star, double za) {
    const bool obs = !isnan(yv);
    const bool upd = obs && !(F_star < kEps);
    F1 = upd ? blk_rcp(F_star) : 0.0;
    vv = upd ? yv - za : 0.0;
    acc_log(upd ? F_star : 1.0);
    n_terms += upd;
    non_init += obs;
    F_eq0 += obs && !upd;
    sumq = fma(vv * vv, F1, sumq);
  };
  // F = z'Pz = sum_I z_I'(2 ml_I - u0_I) and z'a over the group: needs no data of another lane
  auto innovation_sums = [&](double2 ml, double2 u0, double& F_star, double& za) {
    const double2 zi = ZK[0];
    const double w0 = fma(2.0, ml.x, -u0.x), w1 = fma(2.0, ml.y, -u0.y);
    F_star = group_sum(fma(zi.y, w1, zi.x * w0));
    za = group_sum(fma(zi.y, a1, zi.x * a0));
  };
  if constexpr (HO) {
    // nothing left: the thread kernel runs the regular phase (an evaluation that reaches this point has t == N)
  } else if constexpr (P1) {
    // p = 1: the time update does not wait for 1/F.  With M = P z, k = M / F:
    //     T (P - k M') T' + RQR'  =  (T P T' + RQR')  -  (T M)(T M)' / F,        T (a + k v) = T a + (T M) v / F,
    // so the multiply-adds of T P T' are issued while F travels through the butterfly and the reciprocal;
    // (T M)_{I+k} is what the lanes ahead hand over and the rank-1 correction comes last.
    if (t < N) {
      // observations are fetched two steps ahead of their use
      double ynext = YTMA ? 0.0 : A.y[(long long)t * A.ldy + b];
      double ynext2 = (!YTMA && t + 1 < N) ? A.y[(long long)(t + 1) * A.ldy + b] : 0.0;
      double2 ml, u0, c[K];
      double F_star, za;
      for (; t < N - 1; t++) {  // one straight-line block per step
        double yv;
        if constexpr (YTMA) {
          yv = y_get(t);
        } else {
          yv = ynext;
          ynext = ynext2;
          if (t + 2 < N) ynext2 = A.y[(long long)(t + 2) * A.ldy + b];
        }
        load_z(t, 0);
        products(P, ml, u0, c);
        innovation_sums(ml, u0, F_star, za);
        const double2 m = complete(ml, c);
        const double2 tm = make_double2(fma(t01, m.y, t00 * m.x), fma(t11, m.y, t10 * m.x));  // (T M)_I
        double2 tmk[K];
        ahead(tm, tmk);
#pragma unroll
        for (int k = 0; k < K; k++) time_block(k, P[k][0], P[k][1], P[k][2], P[k][3], true);
        const double ta0 = fma(t01, a1, t00 * a0), ta1 = fma(t11, a1, t10 * a0);
        observe(yv, F_star, za);
        const double g0 = tm.x * F1, g1 = tm.y * F1;
#pragma unroll
        for (int k = 0; k < K; k++) {
          P[k][0] = fma(-g0, tmk[k].x, P[k][0]);
          P[k][1] = fma(-g0, tmk[k].y, P[k][1]);
          P[k][2] = fma(-g1, tmk[k].x, P[k][2]);
          P[k][3] = fma(-g1, tmk[k].y, P[k][3]);
        }
        a0 = fma(g0, vv, ta0);
        a1 = fma(g1, vv, ta1);
        z_refill(t);
        y_refill(t);
      }
      // the last observation (:200 -- P and a are not needed afterwards)
      load_z(N - 1, 0);
      products(P, ml, u0, c);
      innovation_sums(ml, u0, F_star, za);
      observe(YTMA ? y_get(N - 1) : ynext, F_star, za);
    }
  } else if (t < N) {
    long long idx = (long long)t * p;
    double ynext = A.y[idx * A.ldy + b];  // observations are fetched one univariate step ahead of their use
    for (; t < N; t++) {
#pragma unroll 1
      for (int j = 0; j < p; j++, idx++) {
        const double yv = ynext;
        if (idx + 1 < Np) ynext = A.y[(idx + 1) * A.ldy + b];
        load_z(t, j);
        double2 ml, u0, c[K], mk[K];
        double F_star, za;
        products(P, ml, u0, c);
        innovation_sums(ml, u0, F_star, za);
        const double2 m = complete(ml, c);
        ahead(m, mk);
        observe(yv, F_star, za);
        const double k0 = m.x * F1, k1 = m.y * F1;
#pragma unroll
        for (int k = 0; k < K; k++) {
          P[k][0] = fma(-k0, mk[k].x, P[k][0]);
          P[k][1] = fma(-k0, mk[k].y, P[k][1]);
          P[k][2] = fma(-k1, mk[k].x, P[k][2]);
          P[k][3] = fma(-k1, mk[k].y, P[k][3]);
        }
        a0 = fma(k0, vv, a0);
        a1 = fma(k1, vv, a1);
      }
      if (t < N - 1) {  // :200-204 (after the last observation P and a are not needed any more)
#pragma unroll
        for (int k = 0; k < K; k++) time_block(k, P[k][0], P[k][1], P[k][2], P[k][3], true);
        const double na0 = fma(t01, a1, t00 * a0), na1 = fma(t11, a1, t10 * a0);
        a0 = na0;
        a1 = na1;
      }
    }
  }

  if (I == 0 && live) {
    double res;
    if (non_init == F_eq0) {
      res = nan("");  // src/LogLikC.cpp:211-213
    } else {
      const double sumlog = (double)esum * 0.69314718055994530942 + log(prod);
      res = ((double)n_terms * kLogConst - 0.5 * (sumlog + sumq)) / (double)N;
    }
    A.out[e_id] = res;
  }
}

#undef PINF

// ---- host side -----------------------------------------------------------------------------------
// States i, k share a block when T or R Q R' couples them.  `Rh` m x r, `qpat` r x r (non-zero pattern of Q as 0/1,
// null = diagonal).  Pairs first (ascending), then the uncoupled states two by two.
bool plan_blocks(const double* Th, int t_diag, const double* Rh, int r_diag, const unsigned char* qpat, int m, int r,
                 BlkPlan* pl, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  std::vector<int> partner(m, -1);
  bool q_off = false;
  auto couple = [&](int i, int k) {
    if (i == k) return true;
    if (partner[i] == k && partner[k] == i) return true;
    if (partner[i] >= 0 || partner[k] >= 0) return false;
    partner[i] = k;
    partner[k] = i;
    return true;
  };
  for (int i = 0; i < m; i++)
    for (int k = 0; k < m; k++)
      if (i != k && mat_get(Th, t_diag, m, i, k) != 0.0 && !couple(i, k)) return no("T couples a state with two others");
  // R Q R' [i][c] != 0 needs k, l with R[i][k], Q[k][l], R[c][l] all non-zero
  for (int i = 0; i < m; i++)
    for (int c = 0; c < m; c++) {
      if (i == c) continue;
      bool nz = false;
      for (int k = 0; k < r && !nz; k++) {
        if (mat_get(Rh, r_diag, m, i, k) == 0.0) continue;
        for (int l = 0; l < r && !nz; l++) {
          const bool q = qpat ? qpat[k + (size_t)l * r] != 0 : k == l;
          nz = q && mat_get(Rh, r_diag, m, c, l) != 0.0;
        }
      }
      if (nz && !couple(i, c)) return no("R Q R' couples a state with two others");
      q_off = q_off || nz;
    }
  int nb = 0;
  std::memset(pl, 0, sizeof *pl);
  for (int i = 0; i < 2 * kBlkMaxNB; i++) pl->perm[i] = -1;
  auto put = [&](int i, int k) {
    if (nb >= kBlkMaxNB) return false;
    pl->perm[2 * nb] = (signed char)i;
    pl->perm[2 * nb + 1] = (signed char)k;
    nb++;
    return true;
  };
  for (int i = 0; i < m; i++)
    if (partner[i] > i && !put(i, partner[i])) return no("more than 8 blocks");
  int pend = -1;
  for (int i = 0; i < m; i++) {
    if (partner[i] >= 0) continue;
    if (pend < 0) {
      pend = i;
    } else {
      if (!put(pend, i)) return no("more than 8 blocks");
      pend = -1;
    }
  }
  if (pend >= 0 && !put(pend, -1)) return no("more than 8 blocks");
  pl->nb = nb;
  pl->q_offdiag = q_off ? 1 : 0;
  for (int J = 0; J < nb; J++)
    for (int rr = 0; rr < 2; rr++)
      for (int cc = 0; cc < 2; cc++) {
        const int i = pl->perm[2 * J + rr], k = pl->perm[2 * J + cc];
        pl->Tb[J][2 * rr + cc] = (i >= 0 && k >= 0) ? mat_get(Th, t_diag, m, i, k) : 0.0;
      }
  return true;
}

bool blk_eligible(const DModel& md, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  if (md.m > 2 * kBlkMaxNB) return no("m > 32");
  if (md.p > kBlkMaxP) return no("p > 8");
  if (!md.Z.shared() || !md.T.shared() || !md.R.shared()) return no("Z, T, R must be shared by the batch");
  if (md.T.tv() || md.R.tv() || md.Q.tv()) return no("time-varying T, R or Q");
  if (md.Z.tv() && md.p != 1) return no("time-varying Z needs p = 1");
  if (!md.Q.diag && !md.Q.shared()) return no("per-evaluation Q must be stored as a diagonal");
  return true;
}

template <int L, int NB, bool P1, bool QOFF, bool ZTV = false, bool HO = false, bool ZTMA = false, bool YTMA = false>
static void launch_blk_one(const BlkArgs& A, unsigned blocks, cudaStream_t st) {
  if constexpr (L == 1 && P1 && !ZTV && !HO && !YTMA) {
    // one lane per series (local level and the like): observations through the bulk-copy ring unless SSGPU_BLK_YTMA=0
    // (rows must be 16-byte aligned at even series offsets: even number of series, aligned y)
    static const bool ytma = [] {
      const char* e = getenv("SSGPU_BLK_YTMA");
      return !(e && e[0] == '0');
    }();
    if (ytma && A.ldy % 2 == 0 && (reinterpret_cast<unsigned long long>(A.y) & 15ull) == 0)
      return launch_blk_one<L, NB, P1, QOFF, ZTV, HO, ZTMA, true>(A, blocks, st);
  }
  if constexpr (ZTV && !ZTMA && NB <= 8) {
    // SSGPU_BLK_ZTMA=1: time-varying Z through the bulk-copy ring (measured slower than the per-lane loads: DESIGN.md)
    static const bool ztma = [] {
      const char* e = getenv("SSGPU_BLK_ZTMA");
      return e && e[0] == '1';
    }();
    if (ztma) return launch_blk_one<L, NB, P1, QOFF, ZTV, HO, true>(A, blocks, st);
  }
  // SSGPU_BLK_PAD_SMEM: unused dynamic shared memory per CTA -- an occupancy knob for experiments (e.g. 120000 leaves
  // one CTA per SM: the step time of a warp that has a sub-partition to itself)
  static const int pad = [] {
    const char* e = getenv("SSGPU_BLK_PAD_SMEM");
    return e ? atoi(e) : 0;
  }();
  if (pad > 0)
    cudaFuncSetAttribute(loglik_blk_kernel<L, NB, P1, QOFF, ZTV, HO, ZTMA, YTMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
  loglik_blk_kernel<L, NB, P1, QOFF, ZTV, HO, ZTMA, YTMA><<<blocks, NB > 8 ? kBlkThreadsCS : kBlkThreads, pad, st>>>(A);
}

template <int L, int NB>
static void launch_blk_nb(const BlkArgs& A, bool q_offdiag, unsigned blocks, cudaStream_t st) {
  if constexpr (NB > 8) {  // p = 1, diagonal R Q R' (checked by the caller)
    if (A.zt)
      launch_blk_one<L, NB, true, false, true>(A, blocks, st);
    else
      launch_blk_one<L, NB, true, false>(A, blocks, st);
    return;
  } else if (A.hd) {  // diffuse prologue of the thread kernel (2 to 8 blocks, p = 1)
    if constexpr (NB >= 2) {
      if (A.zt) {
        if (q_offdiag)
          launch_blk_one<L, NB, true, true, true, true>(A, blocks, st);
        else
          launch_blk_one<L, NB, true, false, true, true>(A, blocks, st);
      } else if (q_offdiag) {
        launch_blk_one<L, NB, true, true, false, true>(A, blocks, st);
      } else {
        launch_blk_one<L, NB, true, false, false, true>(A, blocks, st);
      }
    }
  } else if (A.md.p == 1 && A.zt) {
    if (q_offdiag)
      launch_blk_one<L, NB, true, true, true>(A, blocks, st);
    else
      launch_blk_one<L, NB, true, false, true>(A, blocks, st);
  } else if (A.md.p == 1) {
    if (q_offdiag)
      launch_blk_one<L, NB, true, true>(A, blocks, st);
    else
      launch_blk_one<L, NB, true, false>(A, blocks, st);
  } else {
    if (q_offdiag)
      launch_blk_one<L, NB, false, true>(A, blocks, st);
    else
      launch_blk_one<L, NB, false, false>(A, blocks, st);
  }
}

int blk_lanes(int nb) { return nb > 8 ? 16 : nb > 4 ? 8 : nb > 2 ? 4 : nb > 1 ? 2 : 1; }

// FP64 operations (fma = 2) the kernel issues per series and time step in the regular phase, all L lanes counted
// (H = nb / 2 blocks beside the diagonal one per lane; the counts are those of the SASS of the hot loop)
int blk_flops_per_step(int nb, int p) {
  const int L = blk_lanes(nb), H = nb / 2, K = H + 1;
  int lg = 0;
  while ((1 << lg) < L) lg++;
  // per observation: local products (u0, ml, c), F partial 2 ml - u0, z'a, completion of M, 1/F, loglik terms, rank-1
  const int fma_obs = (2 + 4 * H + 2 * H) + 2 + 1 + 1 + 4 + 1 + 4 * K + 2, mul_obs = 2 + 2 * H + 1 + 1 + 1 + 1 + 2,
            add_obs = 2 * lg + 2 * H + 1;
  // per time step: T_I P_IJ T_J' for the K own blocks (R Q R' added on the diagonal one), T a
  const int fma_t = 4 + 6 + 8 * H + 2, mul_t = 4 + 2 + 8 * H + 2;
  const int tm_fma = p == 1 ? 2 : 0, tm_mul = p == 1 ? 2 : 0;  // p = 1: (T M)_I
  return L * (p * (2 * fma_obs + mul_obs + add_obs) + 2 * (fma_t + tm_fma) + mul_t + tm_mul);
}
// the same without the lanes and blocks that only pad (L > nb: surplus lanes; nb even: the unowned half of the
// blocks at distance nb / 2)
int blk_useful_flops_per_step(int nb, int p) {
  const int L = blk_lanes(nb);
  return (int)((long long)blk_flops_per_step(nb, p) * nb / L);
}

// Host side of the thread kernel's structural shortcuts and of the shared diffuse steps (loglik_thr.cu): everything is
// decided on host copies of the shared matrices.  `md` supplies the storage flags only (shared / diagonal).  The plan's
// blocks are re-ordered in place; ud and self are filled.
void plan_thread_shortcuts(const DModel& md, int nZ, const std::vector<double>& Th, const std::vector<double>& Rh,
                           const std::vector<double>& Zh, const std::vector<double>& ah, const std::vector<double>& Psh,
                           const std::vector<double>& Pih, const std::vector<unsigned char>& qpat, BlkPlan& pl,
                           ThrDiffuse& ud, ThrSelf& self) {
  const int m = md.m, r = md.r, p = md.p;
  // The residual state of the reference's layout (z = 1, nothing in its row and column of T, uncoupled by R Q R'; p = 1,
  // time-invariant Z): moved to position 1 of the last block, where the thread kernel drops it (loglik_thr.cu, EPS).  The
  // kernel assumes P_star[eps, .] = 0 off the diagonal and a[eps] = 0 for a model that starts without a diffuse phase:
  // P_star stored as a diagonal or checked here, a checked here.
  static const bool eps_on = [] {
    const char* e = getenv("SSGPU_THR_EPS");
    return !(e && e[0] == '0');
  }();
  pl.eps = 0;
  if (eps_on && p == 1 && nZ == 1 && pl.nb >= 2 && pl.nb <= kBlkMaxNB / 2 && md.a.shared() && (md.P_star.diag || md.P_star.shared())) {
    int e = -1;
    for (int i = 0; i < m && e < 0; i++) {
      if (Zh[i] != 1.0) continue;
      bool lone = true;
      for (int k = 0; k < m && lone; k++)
        lone = mat_get(Th.data(), md.T.diag, m, i, k) == 0.0 && mat_get(Th.data(), md.T.diag, m, k, i) == 0.0;
      if (lone) e = i;
    }
    int blk = -1, pos = 0;
    for (int i = 0; i < 2 * pl.nb && e >= 0; i++)
      if (pl.perm[i] == e) blk = i / 2, pos = i % 2;
    // uncoupled: its block is one of the paired-up singles (T and R Q R' off the diagonal of the block are zero)
    bool ok = blk >= 0 && pl.Tb[blk][1] == 0.0 && pl.Tb[blk][2] == 0.0;
    if (ok) {
      const int other = pl.perm[2 * blk + 1 - pos];
      if (other >= 0)
        for (int k = 0; k < r && ok; k++) {  // R Q R'[e][other] == 0: no disturbance shared through a non-zero of Q
          if (mat_get(Rh.data(), md.R.diag, m, e, k) == 0.0) continue;
          for (int l = 0; l < r && ok; l++) {
            const bool q = qpat.empty() ? k == l : qpat[k + (size_t)l * r] != 0;
            ok = !(q && mat_get(Rh.data(), md.R.diag, m, other, l) != 0.0);
          }
        }
    }
    if (ok) {
      ok = !ah.empty() && ah[e] == 0.0 && (md.P_star.diag || !Psh.empty());
      for (int k = 0; k < m && ok && !Psh.empty(); k++)
        ok = k == e || (Psh[e + (size_t)k * m] == 0.0 && Psh[k + (size_t)e * m] == 0.0);
    }
    if (ok) {
      const int last = pl.nb - 1;
      for (int c = 0; c < 2; c++) std::swap(pl.perm[2 * blk + c], pl.perm[2 * last + c]);
      for (int c = 0; c < 4; c++) std::swap(pl.Tb[blk][c], pl.Tb[last][c]);
      if (pos == 0) {
        std::swap(pl.perm[2 * last], pl.perm[2 * last + 1]);
        std::swap(pl.Tb[last][0], pl.Tb[last][3]);
        std::swap(pl.Tb[last][1], pl.Tb[last][2]);
      }
      pl.eps = 1;
    }
  }
  // A local level with slope, T_J = [[1, 1], [0, 1]] in either order of its two states: moved to block 0, where the
  // thread kernel replaces its products by sums (loglik_thr.cu, LS)
  static const bool ls_on = [] {
    const char* e = getenv("SSGPU_THR_LS");
    return !(e && e[0] == '0');
  }();
  pl.ls = 0;
  if (ls_on && p == 1 && nZ == 1 && pl.nb >= 2 && pl.nb <= kBlkMaxNB / 2) {
    for (int J = 0; J < pl.nb - (pl.eps ? 1 : 0) && !pl.ls; J++) {
      const double* t = pl.Tb[J];
      const bool upper = t[0] == 1.0 && t[1] == 1.0 && t[2] == 0.0 && t[3] == 1.0;
      const bool lower = t[0] == 1.0 && t[1] == 0.0 && t[2] == 1.0 && t[3] == 1.0;
      if (!upper && !lower) continue;
      if (lower) {
        std::swap(pl.perm[2 * J], pl.perm[2 * J + 1]);
        std::swap(pl.Tb[J][1], pl.Tb[J][2]);
      }
      for (int c = 0; c < 2; c++) std::swap(pl.perm[2 * J + c], pl.perm[c]);
      for (int c = 0; c < 4; c++) std::swap(pl.Tb[J][c], pl.Tb[0][c]);
      pl.ls = 1;
    }
  }
  // z = 0 for the second state of every block (after turning blocks whose FIRST state has the zero around): the thread
  // kernel leaves those multiply-adds out (loglik_thr.cu, ZS)
  static const bool zs_on = [] {
    const char* e = getenv("SSGPU_THR_ZS");
    return !(e && e[0] == '0');
  }();
  pl.zs = 0;
  if (zs_on && p == 1 && nZ == 1 && pl.eps) {
    auto zof = [&](int i) { return pl.perm[i] >= 0 ? Zh[pl.perm[i]] : 0.0; };
    bool all = true;
    for (int J = 0; J < pl.nb - 1; J++) {  // the last block is (partner, residual): the kernel never reads z of position 1
      if (zof(2 * J + 1) != 0.0 && zof(2 * J) == 0.0 && !(pl.ls && J == 0)) {
        std::swap(pl.perm[2 * J], pl.perm[2 * J + 1]);
        std::swap(pl.Tb[J][0], pl.Tb[J][3]);
        std::swap(pl.Tb[J][1], pl.Tb[J][2]);
      }
      all = all && zof(2 * J + 1) == 0.0;
    }
    pl.zs = all ? 1 : 0;
  }
  // Shared diffuse steps (loglik_thr.cu): P_inf, F_inf and the diffuse gains depend on neither the data nor the
  // variances, so for an evaluation without a missing value in the diffuse window they are constants of the model.  The
  // recursion of P_inf alone (src/LogLikC.cpp:101,105,136,143-144,152,204-205) runs here on the host: step t gives
  // u_t = T M_inf and 1 / F_inf; it ends when P_inf has vanished (du steps), and is given up when some F_inf < 1e-7.
  static const bool uni_on = [] {
    const char* e = getenv("SSGPU_THR_SHARED_DIFFUSE");
    return !(e && e[0] == '0');
  }();
  ud = ThrDiffuse{};
  if (uni_on && p == 1 && nZ == 1 && !Pih.empty() && pl.nb >= 2 && pl.nb <= kBlkMaxNB / 2) {
    std::vector<double> Pi((size_t)m * m), Mi((size_t)m), X((size_t)m * m);
    for (int i = 0; i < m; i++)
      for (int k = 0; k < m; k++) Pi[i + (size_t)k * m] = mat_get(Pih.data(), md.P_inf.diag, m, i, k);
    auto small = [&]() {
      for (double x : Pi)
        if (!(std::fabs(x) < kEps)) return false;
      return true;
    };
    bool init = !small(), okd = true;
    int du = 0;
    double prodF = 1.0;
    int esumF = 0;
    while (init && okd && du < kThrMaxDu && du < md.N - 1) {
      double Fi = 0.0;
      for (int i = 0; i < m; i++) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc += Pi[i + (size_t)k * m] * Zh[k];
        Mi[i] = acc;
      }
      for (int i = 0; i < m; i++) Fi += Zh[i] * Mi[i];
      if (Fi < kEps) {
        okd = false;
        break;
      }
      const double F1 = 1.0 / Fi;
      ud.F1[du] = F1;
      int ex;
      prodF = std::frexp(prodF * Fi, &ex) * 2.0;  // mantissa in [1, 2)
      esumF += ex - 1;
      for (int i = 0; i < 2 * pl.nb; i++) {  // u = T M_inf, internal order
        double acc = 0.0;
        if (pl.perm[i] >= 0)
          for (int k = 0; k < m; k++) acc += mat_get(Th.data(), md.T.diag, m, pl.perm[i], k) * Mi[k];
        ud.u[du][i] = acc;
      }
      for (int i = 0; i < m; i++)
        for (int k = 0; k < m; k++) Pi[i + (size_t)k * m] -= Mi[i] * Mi[k] * F1;
      for (int i = 0; i < m; i++)  // X = T P_inf
        for (int k = 0; k < m; k++) {
          double acc = 0.0;
          for (int l = 0; l < m; l++) acc += mat_get(Th.data(), md.T.diag, m, i, l) * Pi[l + (size_t)k * m];
          X[i + (size_t)k * m] = acc;
        }
      for (int i = 0; i < m; i++)  // P_inf = X T'
        for (int k = 0; k < m; k++) {
          double acc = 0.0;
          for (int l = 0; l < m; l++) acc += X[i + (size_t)l * m] * mat_get(Th.data(), md.T.diag, m, k, l);
          Pi[i + (size_t)k * m] = acc;
        }
      du++;
      init = !small();
    }
    if (okd && !init && du > 0) {
      ud.du = du;
      ud.prodF = prodF;
      ud.esumF = esumF;
    }
  }
  // Marked evaluations: a small kernel writes their hand-over record at t = 0 from the model (loglik_thr.cu) when
  // that is a matter of look-ups -- a shared, P_star and Q stored as diagonals, R with at most one non-zero per row and
  // column (so R Q R' = diag(R[i, k]^2 Q[k]): no correlated disturbances)
  static const bool self_on = [] {
    const char* e = getenv("SSGPU_THR_SELF_SETUP");
    return !(e && e[0] == '0');
  }();
  self = ThrSelf{};
  if (self_on && ud.du > 0 && !pl.q_offdiag && md.Q.diag && md.P_star.diag && !ah.empty()) {
    bool ok = true;
    std::vector<int> col_used((size_t)r, 0);
    for (int i = 0; i < 2 * pl.nb && ok; i++) {
      const int st_i = pl.perm[i];
      self.qsrc[i] = -1;
      self.psrc[i] = (signed char)st_i;
      self.qscale[i] = 0.0;
      self.a0[i] = st_i >= 0 ? ah[st_i] : 0.0;
      if (st_i < 0) continue;
      for (int k = 0; k < r && ok; k++) {
        const double rik = mat_get(Rh.data(), md.R.diag, m, st_i, k);
        if (rik == 0.0) continue;
        ok = self.qsrc[i] < 0 && !col_used[k];  // a second non-zero in the row, or the disturbance drives another state too
        col_used[k] = 1;
        self.qsrc[i] = (signed char)k;
        self.qscale[i] = rik * rik;
      }
    }
    if (ok) {
      self.on = 1;
      self.Q = md.Q;
      self.P_star = md.P_star;
    }
  }
}

int launch_loglik_blk(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                      const HostShared* hs, bool* applicable, bool lanes_only, long long call_evals) {
  std::string why;
  if (applicable) *applicable = true;
  SS_ARG(blk_eligible(md, &why), "block kernel not eligible: " + why);
  const int m = md.m, r = md.r, p = md.p;
  // the shared matrices are a few hundred bytes: fetch what the caller has not passed along (host copies)
  const int nZ = md.Z.tv() ? md.N : 1;
  std::vector<double> Th((size_t)(md.T.diag ? m : m * m)), Rh((size_t)(md.R.diag ? m : m * r)), Zh((size_t)p * m * nZ), Qh;
  std::vector<unsigned char> qpat;
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
  SS_TRY(fetch(Zh, (hs && nZ == 1) ? hs->Z : nullptr, md.Z.p));
  if (!md.Q.diag) {
    Qh.resize((size_t)r * r);
    SS_TRY(fetch(Qh, hs ? hs->Q : nullptr, md.Q.p));
  }
  // what the thread kernel's residual shortcut looks at (p = 1, time-invariant Z): a, and a dense P_star, when shared
  std::vector<double> ah, Psh, Pih;
  if (p == 1 && nZ == 1) {
    if (md.a.shared()) {
      ah.resize((size_t)m);
      SS_TRY(fetch(ah, hs ? hs->a : nullptr, md.a.p));
    }
    if (md.P_star.shared() && !md.P_star.diag) {
      Psh.resize((size_t)m * m);
      SS_TRY(fetch(Psh, hs ? hs->P_star : nullptr, md.P_star.p));
    }
    if (md.P_inf.shared()) {
      Pih.resize((size_t)(md.P_inf.diag ? m : m * m));
      SS_TRY(fetch(Pih, hs ? hs->P_inf : nullptr, md.P_inf.p));
    }
  }
  if (need_sync) SS_CUDA(cudaStreamSynchronize(st));
  if (!md.Q.diag) {
    qpat.resize((size_t)r * r);
    for (size_t i = 0; i < qpat.size(); i++) qpat[i] = Qh[i] != 0.0;
  }
  BlkPlan pl;
  const bool planned = plan_blocks(Th.data(), md.T.diag, Rh.data(), md.R.diag, qpat.empty() ? nullptr : qpat.data(), m, r,
                                   &pl, &why);
  if (!planned && applicable) {  // automatic dispatch: the caller moves on to the next kernel
    *applicable = false;
    return SSGPU_OK;
  }
  SS_ARG(planned, "block kernel not eligible: " + why);
  ThrDiffuse ud{};
  ThrSelf self{};
  plan_thread_shortcuts(md, nZ, Th, Rh, Zh, ah, Psh, Pih, qpat, pl, ud, self);
  BlkArgs A{};
  A.md = md;
  A.y = y;
  A.B = B;
  A.E = B * V;
  A.out = out;
  for (int i = 0; i < 2 * kBlkMaxNB; i++) A.perm[i] = pl.perm[i];
  std::memcpy(A.Tb, pl.Tb, sizeof A.Tb);
  for (int j = 0; j < p; j++)
    for (int i = 0; i < 2 * kBlkMaxNB; i++) A.z[j][i] = pl.perm[i] >= 0 ? Zh[j + (size_t)pl.perm[i] * p] : 0.0;
  DevBuf ztd;  // released stream-ordered behind the kernel
  if (nZ > 1) {
    std::vector<double> zt((size_t)nZ * 2 * kBlkMaxNB, 0.0);
    for (int t = 0; t < nZ; t++)
      for (int i = 0; i < 2 * kBlkMaxNB; i++)
        if (pl.perm[i] >= 0) zt[(size_t)t * 2 * kBlkMaxNB + i] = Zh[(size_t)t * p * m + (size_t)pl.perm[i] * p];
    SS_TRY(ztd.alloc(zt.size() * sizeof(double), st));
    SS_CUDA(cudaMemcpyAsync(ztd.get(), zt.data(), zt.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaStreamSynchronize(st));  // the host vector dies with this frame
    A.zt = ztd.as<double>();
  }
  if (pl.nb > 8 && (p != 1 || pl.q_offdiag)) {  // more than 8 blocks: univariate models with a diagonal R Q R' only
    why = "more than 8 blocks needs p = 1 and a diagonal R Q R'";
    if (applicable) {
      *applicable = false;
      return SSGPU_OK;
    }
    SS_ARG(false, "block kernel not eligible: " + why);
  }
  A.ldy = B;
  const int L = blk_lanes(pl.nb), G = 32 / L, wpb = (pl.nb > 8 ? kBlkThreadsCS : kBlkThreads) / 32;
  auto launch_lanes = [&](const BlkArgs& Ax) -> int {
    const long long blocks = (Ax.E + (long long)wpb * G - 1) / ((long long)wpb * G);
    SS_ARG(blocks < (1LL << 31), "block kernel: batch too large for one launch");
    const unsigned nblk = (unsigned)blocks;
    switch (pl.nb) {
      case 1: launch_blk_nb<1, 1>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 2: launch_blk_nb<2, 2>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 3: launch_blk_nb<4, 3>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 4: launch_blk_nb<4, 4>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 5: launch_blk_nb<8, 5>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 6: launch_blk_nb<8, 6>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 7: launch_blk_nb<8, 7>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 8: launch_blk_nb<8, 8>(Ax, pl.q_offdiag != 0, nblk, st); break;
      case 9: launch_blk_nb<16, 9>(Ax, false, nblk, st); break;
      case 10: launch_blk_nb<16, 10>(Ax, false, nblk, st); break;
      case 11: launch_blk_nb<16, 11>(Ax, false, nblk, st); break;
      case 12: launch_blk_nb<16, 12>(Ax, false, nblk, st); break;
      case 13: launch_blk_nb<16, 13>(Ax, false, nblk, st); break;
      case 14: launch_blk_nb<16, 14>(Ax, false, nblk, st); break;
      case 15: launch_blk_nb<16, 15>(Ax, false, nblk, st); break;
      default: launch_blk_nb<16, 16>(Ax, false, nblk, st); break;
    }
    SS_CUDA(cudaGetLastError());
    count_launch();
    return SSGPU_OK;
  };
  static thread_local char name[320];
  // SSGPU_BLK_THREAD=0: the lane kernel alone (experiments, A/B runs)
  static const bool thr_on = [] {
    const char* e = getenv("SSGPU_BLK_THREAD");
    return !(e && e[0] == '0');
  }();
  // A few evaluations (a single series in statespacer()'s optim loop: 1 + 2k of them) finish sooner on the lanes: one launch
  // instead of two, and a step is 243 instructions of a warp instead of 1,134 of a thread.  The automatic dispatch therefore
  // takes the thread kernel from kThrMinEvals evaluations on (about one wave of the lane kernel); SSGPU_KERNEL_BLK always.
  constexpr long long kThrMinEvals = 8192;
  const bool explicit_blk = applicable == nullptr;
  // (decided on the size of the whole call, so that ranges of one call -- upload chunks, device slots -- agree bit for bit)
  if (thr_on && !lanes_only && thr_supported(pl.nb, p, nZ > 1) && (explicit_blk || std::max(A.E, call_evals) >= kThrMinEvals)) {
    // Two launches per range of series: the lane kernel runs the exact diffuse phase and leaves the state behind, the
    // thread kernel (loglik_thr.cu) runs the regular phase.  The state record is ~1.1 KB per evaluation at 7 blocks:
    // ranges of series are sized for SSGPU_THR_SCRATCH_MB (default 2048) of scratch.
    static const long long cap = [] {
      const char* e = getenv("SSGPU_THR_SCRATCH_MB");
      return (long long)(e ? std::max(1, atoi(e)) : 2048) << 20;
    }();
    const size_t per_eval = thr_handoff_bytes(pl.nb, pl.q_offdiag != 0, 1);
    const long long Bc_max = std::min<long long>(B, std::max<long long>(1, cap / (long long)(per_eval * V)));
    const BlkHandoff hl(pl.nb, pl.q_offdiag != 0);
    DevBuf scratch;  // released stream-ordered behind the last kernel
    SS_TRY(scratch.alloc((size_t)Bc_max * V * per_eval, st));
    for (long long b0 = 0; b0 < B; b0 += Bc_max) {
      const long long Bc = std::min(Bc_max, B - b0), Ec = Bc * V;
      BlkArgs Ac = A;
      for (DMat* dm : {&Ac.md.a, &Ac.md.P_inf, &Ac.md.P_star, &Ac.md.Q}) dm->p += b0 * dm->sb;
      Ac.y = y + b0;
      Ac.out = out + b0;
      Ac.B = Bc;
      Ac.E = Ec;
      Ac.hd = scratch.as<double>();
      Ac.hi = reinterpret_cast<int*>(Ac.hd + (size_t)hl.doubles() * Ec);
      Ac.du = ud.du;
      Ac.self_setup = self.on;
      ThrSelf selfc = self;  // per-evaluation variances of the range: same strides, shifted base
      selfc.Q.p += b0 * selfc.Q.sb;
      selfc.P_star.p += b0 * selfc.P_star.sb;
      SS_TRY(launch_lanes(Ac));
      if (selfc.on) {
        SS_TRY(launch_thr_setup(pl.nb, selfc, Bc, Ec, Ac.hd, Ac.hi, st));
        count_launch();
      }
      SS_TRY(launch_loglik_thr(pl, &A.z[0][0], A.zt, Ac.y, Bc, Ec, B, md.N, Ac.hd, Ac.hi, Ac.out, st, pl.eps != 0, pl.ls != 0, pl.zs != 0, &ud));
      count_launch();
    }
    snprintf(name, sizeof name,
             "loglik_thr_kernel<NB=%d,S=%d,flop_per_step=%d,fp64_instr_per_step=%d>%s%s after loglik_blk_kernel<L=%d,NB=%d> (diffuse phase)",
             pl.nb, thr_smem_blocks_host(pl.nb, nZ > 1, pl.eps != 0),
             thr_flops_per_step(pl.nb, pl.q_offdiag != 0, pl.eps != 0, pl.ls != 0, pl.zs != 0),
             thr_flops_per_step(pl.nb, pl.q_offdiag != 0, pl.eps != 0, pl.ls != 0, pl.zs != 0, true), nZ > 1 ? " time-varying Z" : "",
             pl.eps ? (pl.ls ? " residual state dropped, level + slope block by sums" : " residual state dropped")
                    : (pl.ls ? " level + slope block by sums" : ""),
             L, pl.nb);
    if (pl.zs) {
      const size_t len = strlen(name);
      snprintf(name + len, sizeof name - len, ", zero z skipped");
    }
    if (ud.du > 0) {
      const size_t len = strlen(name);
      snprintf(name + len, sizeof name - len, ", %d shared diffuse steps", ud.du);
    }
    set_last_kernel(name);
    return SSGPU_OK;
  }
  A.B = B;
  SS_TRY(launch_lanes(A));
  snprintf(name, sizeof name, "loglik_blk_kernel<L=%d,NB=%d,flop_per_step=%d>%s", L, pl.nb, blk_flops_per_step(pl.nb, p),
           nZ > 1 ? " time-varying Z" : "");
  set_last_kernel(name);
  return SSGPU_OK;
}

}  // namespace ssgpu

// Host-only helper (no device needed): the block decomposition the block-row kernel would use.
extern "C" int ssgpu_plan_blocks(const double* T, const double* R, const double* Q, int m, int r, int p, int* perm,
                                 int* n_blocks, int* lanes, int* flop_per_step) {
  using namespace ssgpu;
  SS_ARG(T && R && m >= 1 && r >= 1 && p >= 1, "ssgpu_plan_blocks: bad arguments");
  BlkPlan pl;
  std::string why;
  std::vector<unsigned char> qpat;
  if (Q) {
    qpat.resize((size_t)r * r);
    for (size_t i = 0; i < qpat.size(); i++) qpat[i] = Q[i] != 0.0;
  }
  const bool ok = m <= 2 * kBlkMaxNB && p <= kBlkMaxP &&
                  plan_blocks(T, 0, R, 0, Q ? qpat.data() : nullptr, m, r, &pl, &why);
  if (n_blocks) *n_blocks = ok ? pl.nb : 0;
  if (!ok) return SSGPU_OK;  // not decomposable: the automatic dispatch uses another kernel
  if (perm)
    for (int i = 0; i < 2 * kBlkMaxNB; i++) perm[i] = pl.perm[i];
  if (lanes) *lanes = blk_lanes(pl.nb);
  // what the automatic dispatch runs: 2 to 8 blocks with p = 1 go to the thread-per-evaluation kernel after the diffuse
  // phase (time-invariant Z assumed here: the planner does not see Z)
  if (flop_per_step)
    *flop_per_step = thr_supported(pl.nb, p, false) ? thr_flops_per_step(pl.nb, pl.q_offdiag != 0) : blk_flops_per_step(pl.nb, p);
  return SSGPU_OK;
}

// Host-only helper (no device needed): what the thread-per-evaluation kernel would do with a univariate model whose
// matrices are shared by the batch -- the structural shortcuts that apply and the number of shared diffuse steps.
extern "C" int ssgpu_plan_thread_kernel(const double* T, const double* R, const double* Q, const double* Z, const double* a,
                                        const double* P_inf, const double* P_star, int m, int r, int N, int* n_blocks,
                                        int* residual_dropped, int* level_slope_by_sums, int* zero_z_skipped,
                                        int* shared_diffuse_steps, int* flop_per_step) {
  using namespace ssgpu;
  SS_ARG(T && R && Z && a && P_inf && P_star && m >= 1 && r >= 1 && N >= 1, "ssgpu_plan_thread_kernel: bad arguments");
  for (int* o : {n_blocks, residual_dropped, level_slope_by_sums, zero_z_skipped, shared_diffuse_steps, flop_per_step})
    if (o) *o = 0;
  BlkPlan pl;
  std::string why;
  std::vector<unsigned char> qpat;
  if (Q) {
    qpat.resize((size_t)r * r);
    for (size_t i = 0; i < qpat.size(); i++) qpat[i] = Q[i] != 0.0;
  }
  if (m > 2 * kBlkMaxNB || !plan_blocks(T, 0, R, 0, Q ? qpat.data() : nullptr, m, r, &pl, &why)) return SSGPU_OK;
  if (n_blocks) *n_blocks = pl.nb;
  if (!thr_supported(pl.nb, 1, false)) return SSGPU_OK;
  DModel md{};
  md.N = N;
  md.p = 1;
  md.m = m;
  md.r = r;  // every matrix dense and shared: strides 0, diag 0
  const std::vector<double> Th(T, T + (size_t)m * m), Rh(R, R + (size_t)m * r), Zh(Z, Z + (size_t)m), ah(a, a + (size_t)m),
      Psh(P_star, P_star + (size_t)m * m), Pih(P_inf, P_inf + (size_t)m * m);
  ThrDiffuse ud{};
  ThrSelf self{};
  plan_thread_shortcuts(md, 1, Th, Rh, Zh, ah, Psh, Pih, qpat, pl, ud, self);
  if (residual_dropped) *residual_dropped = pl.eps;
  if (level_slope_by_sums) *level_slope_by_sums = pl.ls;
  if (zero_z_skipped) *zero_z_skipped = pl.zs;
  if (shared_diffuse_steps) *shared_diffuse_steps = ud.du;
  if (flop_per_step) *flop_per_step = thr_flops_per_step(pl.nb, pl.q_offdiag != 0, pl.eps != 0, pl.ls != 0, pl.zs != 0);
  return SSGPU_OK;
}
