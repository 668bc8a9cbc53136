// Batched loglikelihood for models whose transition matrix is block diagonal with blocks of at most 2 x 2:
// structure-exploiting DFMA kernel, L lanes per evaluation, the covariance distributed by block rows.
//
// Restates the recursion of src/LogLikC.cpp:68-214 (univariate treatment, exact diffuse initialisation,
// NaN = missing) for models whose Z, T, R are shared by the batch, T and R time-invariant, Z time-invariant or -- with
// p = 1 -- one row per time point (explanatory variables) (per-evaluation a, P_inf, P_star, Q allowed) and whose T, after re-ordering the states, is  blkdiag(T_0, ..., T_{NB-1})  with
// 1 x 1 or 2 x 2 blocks, NB <= 8 (m <= 16; NB <= 16, m <= 32 when p = 1), and whose R Q R' has no entry outside those blocks.  That is every
// structural model statespacer builds from residuals, level, slope, trigonometric seasonals, cycles and
// regression coefficients (R/GetSysMat.R:1388-1426, R/Components-Unobserved.R:143-159,491-519): T couples a state
// with at most one partner.  BASELINE.json config 5 (m = 14) is 7 such blocks.
//
// Why this shape.  With P cut into 2 x 2 blocks P_IJ the time update  P <- T P T' + R Q R'  (src/LogLikC.cpp:202)
// is  P_IJ <- T_I P_IJ T_J' (+ (RQR')_II)  block by block: 16 multiply-adds per block, 4 m^2 per step instead of the
// 4 m^3 of the dense product (the FP64 tensor kernel of loglik_mma.cu multiplies whole 8 x 8 tiles: 7,168 flop per
// step at m = 14 where this kernel issues about 1,800), and it needs NO data from any other block.  Lane I of a
// group of L lanes therefore owns block row I (NB blocks, 8 NB registers) and runs the whole time update from its
// own registers; T_J and z are the same for every evaluation and come from the kernel parameters (constant bank).
// Only the measurement update couples the block rows:  M = P z  is computed row-wise in place, all-gathered through
// a few hundred bytes of shared memory (one 16-byte store and NB broadcast 16-byte loads per lane), F = z'M and z'a
// are xor-butterfly sums over the group (bit-identical in every lane), and  P_IJ -= (M_I / F) M_J'  is local again.
// 32 / L evaluations share a warp; the exact diffuse phase (about d of the N steps) runs in a select-based
// loop that all groups of the warp execute until the last of them has left it, with P_inf parked in shared memory
// so that the registers are sized for the regular phase.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace ssgpu {

#ifndef SSGPU_BLK_MINB
#define SSGPU_BLK_MINB 2  // resident CTAs per SM the register budget is sized for (2: 255 registers, 3: 168, 4: 128)
#endif
constexpr int kBlkThreads = 128;   // CTA of the register-constant instantiations (NB <= 8)
constexpr int kBlkThreadsCS = 64;  // NB > 8: P_inf of a CTA (16 NB doubles per thread) has to fit the static 48 KB
constexpr int kBlkMaxNB = 16;  // blocks: 1..8 with T_J and z in registers, 9..16 (p = 1) with both in shared memory
constexpr int kBlkMaxP = 8;

struct BlkArgs {
  DModel md;
  const double* y;  // [(t*p + j) * B + b]
  long long B;
  long long E;  // B * V
  double* out;
  signed char perm[2 * kBlkMaxNB];     // internal position 2I+e holds state perm[2I+e] of the caller, -1 = padding
  double Tb[kBlkMaxNB][4];             // T_J[row][col] at [2*row + col], internal order
  double z[kBlkMaxP][2 * kBlkMaxNB];   // row j of Z, internal order
  const double* zt;                    // time-varying Z (p = 1): [t][2 * kBlkMaxNB] in internal order, device memory
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

// QOFF: the diagonal blocks of R Q R' have off-diagonal entries (correlated disturbances of a multivariate model)
// ZTV: Z changes with t (explanatory variables, R/Components-Regressors.R:60-78); p = 1 only
template <int L, int NB, bool P1, bool QOFF, bool ZTV = false>
__global__ void __launch_bounds__(NB > 8 ? kBlkThreadsCS : kBlkThreads, NB > 8 ? 4 : SSGPU_BLK_MINB)
    loglik_blk_kernel(const BlkArgs A) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / L;  // evaluations per warp
  // NB > 8 (17 <= m <= 32, p = 1): the block row alone is 8 NB <= 128 registers, so T_J and z live in shared memory
  // (8-byte broadcast loads) and the R Q R' addends are selected per step instead of hoisted
  constexpr bool CS = NB > 8;
  constexpr int BT = CS ? kBlkThreadsCS : kBlkThreads;
  static_assert(!CS || P1, "more than 8 blocks: p = 1 only");
  const DModel& md = A.md;
  const int N = md.N, p = P1 ? 1 : md.p, M = md.m, r = md.r;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int I = lane % L, g = lane / L;
  __shared__ double tb_s[CS ? NB * 4 : 1], zs_s[CS ? NB * 2 : 1];
  if constexpr (CS) {  // before any warp leaves: filled by the whole CTA
    for (int e = threadIdx.x; e < NB * 4; e += BT) tb_s[e] = (&A.Tb[0][0])[e];
    for (int e = threadIdx.x; e < NB * 2; e += BT) zs_s[e] = A.z[0][e];
    __syncthreads();
  }
  // evaluations are numbered series-major (w = b*V + v) so that the V parameter vectors of a series run side by
  // side and share its y through L1/L2; results keep the public order e = v*B + b
  const long long w0 = ((long long)blockIdx.x * (BT / 32) + warp) * G;
  if (w0 >= A.E) return;  // whole warp
  // every lane stays alive (shuffles, __syncwarp): groups beyond the last evaluation shadow it and do not store
  const bool live = w0 + g < A.E;
  const long long w_id = live ? w0 + g : A.E - 1;
  const long long V = A.E / A.B, b = w_id / V, v = w_id % V;
  const long long e_id = v * A.B + b;
  const bool own = I < NB;  // L > NB: the surplus lanes carry zero blocks

  __shared__ double2 xch_s[BT / 32][2][NB][G];
  __shared__ double pinf_s[NB * 4][BT];
  double2(*xch)[NB][G] = xch_s[warp];
  double(*pinf)[BT] = pinf_s;
  const int tx = threadIdx.x;

  // ---- own block row of P_star, P_inf, own T block, own part of a and z, own diagonal block of R Q R' ------------
  const int i0 = own ? (int)A.perm[2 * I] : -1, i1 = own ? (int)A.perm[2 * I + 1] : -1;
  auto elem = [&](const double* blk, int diag, int i, int k) { return (i >= 0 && k >= 0) ? mat_get(blk, diag, M, i, k) : 0.0; };
  double P[NB][4];
  {
    const double* Ps = md.P_star.block(b, v, 0);
    const double* Pi = md.P_inf.block(b, v, 0);
#pragma unroll
    for (int J = 0; J < NB; J++) {
      const int k0 = A.perm[2 * J], k1 = A.perm[2 * J + 1];
      P[J][0] = elem(Ps, md.P_star.diag, i0, k0);
      P[J][1] = elem(Ps, md.P_star.diag, i0, k1);
      P[J][2] = elem(Ps, md.P_star.diag, i1, k0);
      P[J][3] = elem(Ps, md.P_star.diag, i1, k1);
      pinf[J * 4 + 0][tx] = elem(Pi, md.P_inf.diag, i0, k0);
      pinf[J * 4 + 1][tx] = elem(Pi, md.P_inf.diag, i0, k1);
      pinf[J * 4 + 2][tx] = elem(Pi, md.P_inf.diag, i1, k0);
      pinf[J * 4 + 3][tx] = elem(Pi, md.P_inf.diag, i1, k1);
    }
  }
  double a0, a1;
  {
    const double* ab = md.a.block(b, v, 0);
    a0 = i0 >= 0 ? ab[i0] : 0.0;
    a1 = i1 >= 0 ? ab[i1] : 0.0;
  }
  const double t00 = A.Tb[I][0], t01 = A.Tb[I][1], t10 = A.Tb[I][2], t11 = A.Tb[I][3];  // zero beyond NB
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
    q1 = rqr(i0, i1);
    q2 = rqr(i1, i0);
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

  // T_J (and z when p = 1) are the same for every evaluation.  They are kept in VECTOR registers for the whole
  // kernel (4 NB + 2 NB doubles; the kernel is sized for 2 CTAs per SM, 255 registers), loaded once through an index
  // that ptxas cannot prove warp-uniform (`lz` is always 0 but derives from the per-lane evaluation number).
  // What was measured on the way (config 5, one B200):
  //   constants in shared memory, 16-byte broadcast loads: an LDS.128 costs ~3 LSU wavefronts even when every lane
  //     reads the same address; LSU pipe 77 % busy, FP64 pipe 58 %, 167 ms per sweep;
  //   constants through the uniform datapath (LDCU -> uniform-register operand): left loop-invariant, ptxas hoists
  //     all 42 of a step, runs out of uniform registers and spills them (300 extra instructions per step); made
  //     loop-variant, it funnels every constant through ONE uniform register pair, LDCU / 2 DFMA / LDCU ..., each
  //     load's latency exposed (IDC 56 % busy, `short_scoreboard` the top stall), FP64 pipe 60 %, 162 ms.
  const int lz = (int)(w_id >> 62);
  double TJ[CS ? 1 : NB][4], ZJ[CS ? 1 : NB][2];
  if constexpr (!CS) {
    const double* tbp = &A.Tb[0][0] + lz;
    const double* zp = &A.z[0][0] + lz;
#pragma unroll
    for (int J = 0; J < NB; J++) {
#pragma unroll
      for (int k = 0; k < 4; k++) TJ[J][k] = tbp[J * 4 + k];
      ZJ[J][0] = P1 ? zp[2 * J] : 0.0;
      ZJ[J][1] = P1 ? zp[2 * J + 1] : 0.0;
    }
  }
  // P_IJ <- T_I P_IJ T_J' (+ the own diagonal block of R Q R' when J == I); X[r][c] = sum_k Y[r][k] T_J[c][k].
  // The selected addends (dg ? q : 0) are loop-invariant and hoisted by the compiler: 2 NB doubles when R Q R' is
  // diagonal (QOFF = false).
  int zv = 0;  // NB > 8: always 0, but refreshed from a per-lane counter every step so that the selected R Q R' addends
               // are not hoisted into 4 NB registers
  auto time_block = [&](int J, double& p0, double& p1, double& p2, double& p3, bool with_q) {
    const double y00 = fma(t01, p2, t00 * p0), y01 = fma(t01, p3, t00 * p1);
    const double y10 = fma(t11, p2, t10 * p0), y11 = fma(t11, p3, t10 * p1);
    const bool dg = with_q && J == (CS ? I + zv : I);
    const double tj0 = CS ? tb_s[J * 4 + 0] : TJ[CS ? 0 : J][0], tj1 = CS ? tb_s[J * 4 + 1] : TJ[CS ? 0 : J][1];
    const double tj2 = CS ? tb_s[J * 4 + 2] : TJ[CS ? 0 : J][2], tj3 = CS ? tb_s[J * 4 + 3] : TJ[CS ? 0 : J][3];
    p0 = fma(y01, tj1, fma(y00, tj0, dg ? q0 : 0.0));
    p1 = fma(y01, tj3, fma(y00, tj2, (QOFF && dg) ? q1 : 0.0));
    p2 = fma(y11, tj1, fma(y10, tj0, (QOFF && dg) ? q2 : 0.0));
    p3 = fma(y11, tj3, fma(y10, tj2, dg ? q3 : 0.0));
  };

  const long long Np = (long long)N * p;
  int buf = 0;
  int t = 0;
  {
    // ---- exact diffuse phase (src/LogLikC.cpp:98-159), select-based: the groups of a warp may leave it at
    // different time points (missing observations), so every group runs this loop until the last one is done ----
    bool small = true;
#pragma unroll
    for (int e = 0; e < NB * 4; e++) small = small && (fabs(pinf[e][tx]) < kEps);
    bool init = !group_all(small);  // :49
    while (t < N && __any_sync(FULL, init)) {
      for (int j = 0; j < p; j++) {
        const double yv = A.y[((long long)t * p + j) * A.B + b];
        const bool obs = !isnan(yv);  // :88-90
        const double* zrow = ZTV ? A.zt + (size_t)t * (2 * kBlkMaxNB) : &A.z[j][0];
        const double zi0 = zrow[2 * I], zi1 = zrow[2 * I + 1];
        double ms0 = 0.0, ms1 = 0.0, mi0 = 0.0, mi1 = 0.0;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          const double z0 = zrow[2 * J], z1 = zrow[2 * J + 1];
          ms0 = fma(P[J][1], z1, fma(P[J][0], z0, ms0));
          ms1 = fma(P[J][3], z1, fma(P[J][2], z0, ms1));
          mi0 = fma(pinf[J * 4 + 1][tx], z1, fma(pinf[J * 4 + 0][tx], z0, mi0));
          mi1 = fma(pinf[J * 4 + 3][tx], z1, fma(pinf[J * 4 + 2][tx], z0, mi1));
        }
        const double F_star = group_sum(fma(zi1, ms1, zi0 * ms0));
        const double F_inf = group_sum(fma(zi1, mi1, zi0 * mi0));
        const double za = group_sum(fma(zi1, a1, zi0 * a0));
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
        if (own) xch[buf][I][g] = make_double2(ms0, ms1);
        if (own) xch[buf ^ 1][I][g] = make_double2(mi0, mi1);
        __syncwarp();
#pragma unroll
        for (int J = 0; J < NB; J++) {
          const double2 sJ = xch[buf][J][g], iJ = xch[buf ^ 1][J][g];
          const double k00 = fma(F1s, sJ.x, F1i * iJ.x), k01 = fma(F1s, sJ.y, F1i * iJ.y);    // K0_J
          const double k10 = fma(F2, iJ.x, F1i * sJ.x), k11 = fma(F2, iJ.y, F1i * sJ.y);      // K1_J
          P[J][0] = fma(-mi0, k10, fma(-ms0, k00, P[J][0]));
          P[J][1] = fma(-mi0, k11, fma(-ms0, k01, P[J][1]));
          P[J][2] = fma(-mi1, k10, fma(-ms1, k00, P[J][2]));
          P[J][3] = fma(-mi1, k11, fma(-ms1, k01, P[J][3]));
          const double c0 = F1i * iJ.x, c1 = F1i * iJ.y;
          pinf[J * 4 + 0][tx] = fma(-mi0, c0, pinf[J * 4 + 0][tx]);
          pinf[J * 4 + 1][tx] = fma(-mi0, c1, pinf[J * 4 + 1][tx]);
          pinf[J * 4 + 2][tx] = fma(-mi1, c0, pinf[J * 4 + 2][tx]);
          pinf[J * 4 + 3][tx] = fma(-mi1, c1, pinf[J * 4 + 3][tx]);
        }
        __syncwarp();  // both buffers are rewritten by the next observation
        a0 = fma(fma(F1s, ms0, F1i * mi0), vv, a0);
        a1 = fma(fma(F1s, ms1, F1i * mi1), vv, a1);
        if (j < p - 1) {  // :157-159 (only after a diffuse update)
          small = true;
#pragma unroll
          for (int e = 0; e < NB * 4; e++) small = small && (fabs(pinf[e][tx]) < kEps);
          const bool gone = group_all(small);
          if (dif && gone) init = false;
        }
      }
      if (t < N - 1) {  // :200-207
        small = true;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          time_block(J, P[J][0], P[J][1], P[J][2], P[J][3], true);
          double i0v = pinf[J * 4 + 0][tx], i1v = pinf[J * 4 + 1][tx], i2v = pinf[J * 4 + 2][tx], i3v = pinf[J * 4 + 3][tx];
          time_block(J, i0v, i1v, i2v, i3v, false);
          pinf[J * 4 + 0][tx] = i0v;
          pinf[J * 4 + 1][tx] = i1v;
          pinf[J * 4 + 2][tx] = i2v;
          pinf[J * 4 + 3][tx] = i3v;
          small = small && (fabs(i0v) < kEps) && (fabs(i1v) < kEps) && (fabs(i2v) < kEps) && (fabs(i3v) < kEps);
        }
        const double na0 = fma(t01, a1, t00 * a0), na1 = fma(t11, a1, t10 * a0);
        a0 = na0;
        a1 = na1;
        const bool gone = group_all(small);
        if (init && gone) init = false;
      }
      t++;
    }
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
  if constexpr (P1) {
    // p = 1: the time update does not wait for 1/F.  With M = P z, k = M / F:
    //     T (P - k M') T' + RQR'  =  (T P T' + RQR')  -  (T M)(T M)' / F,        T (a + k v) = T a + (T M) v / F,
    // so the 16 NB multiply-adds of T P T' are issued while F = z'M travels through the butterfly and the
    // reciprocal; lane I publishes T_I M_I instead of M_I and the rank-1 correction comes last.  (At 2 warps per
    // sub-partition the FP64 pipe otherwise idles through that chain: 148 ms per sweep of config 5 before, see
    // profiles/README.md.)
    if (t < N) {
      double zi0 = A.z[0][2 * I], zi1 = A.z[0][2 * I + 1];
      // time-varying Z: the row of step t in internal order, the same for every evaluation (16-byte broadcast loads)
      const double2* zcur = nullptr;  // NB > 8 with time-varying Z: the row of the step, read where it is used
      auto load_z = [&](int tt) {
        if constexpr (ZTV) {
          const double2* zr = reinterpret_cast<const double2*>(A.zt + (size_t)tt * (2 * kBlkMaxNB));
#pragma unroll
          if constexpr (!CS) {
#pragma unroll
            for (int J = 0; J < NB; J++) {
              const double2 zz = zr[J];
              ZJ[J][0] = zz.x;
              ZJ[J][1] = zz.y;
            }
          }
          zcur = zr;
          const double2 zo = zr[I];
          zi0 = zo.x;
          zi1 = zo.y;
        }
      };
      // observations are fetched two steps ahead of their use (one step, 600 cycles at 2 warps per sub-partition,
      // does not always cover a DRAM miss: 2.4 % of the stall samples sat on this load)
      double ynext = A.y[(long long)t * A.B + b];
      double ynext2 = t + 1 < N ? A.y[(long long)(t + 1) * A.B + b] : 0.0;
      // M_I = sum_J P_IJ z_J (two chains per row), F = z'M and z'a over the group
      double m0, m1, F_star, za;
      auto innovation = [&]() {
        double m0a = 0.0, m0b = 0.0, m1a = 0.0, m1b = 0.0;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          double z0, z1;
          if constexpr (!CS) {
            z0 = ZJ[J][0];
            z1 = ZJ[J][1];
          } else if constexpr (ZTV) {
            const double2 zz = zcur[J];
            z0 = zz.x;
            z1 = zz.y;
          } else {
            z0 = zs_s[2 * J];
            z1 = zs_s[2 * J + 1];
          }
          m0a = fma(P[J][0], z0, m0a);
          m0b = fma(P[J][1], z1, m0b);
          m1a = fma(P[J][2], z0, m1a);
          m1b = fma(P[J][3], z1, m1b);
        }
        m0 = m0a + m0b;
        m1 = m1a + m1b;
      };
      for (; t < N - 1; t++) {  // one straight-line block per step: no branch between the butterfly and the products
        const double yv = ynext;
        ynext = ynext2;
        if (t + 2 < N) ynext2 = A.y[(long long)(t + 2) * A.B + b];
        if constexpr (CS) zv = n_terms >> 30;
        load_z(t);
        innovation();
        const double tm0 = fma(t01, m1, t00 * m0), tm1 = fma(t11, m1, t10 * m0);  // (T M)_I
        if constexpr (L > 1) {  // one lane per evaluation owns every block: nothing to exchange
          if (own) xch[buf][I][g] = make_double2(tm0, tm1);
        }
        F_star = group_sum(fma(zi1, m1, zi0 * m0));
        za = group_sum(fma(zi1, a1, zi0 * a0));
        if constexpr (L > 1) __syncwarp();
#pragma unroll
        for (int J = 0; J < NB; J++) time_block(J, P[J][0], P[J][1], P[J][2], P[J][3], true);
        const double ta0 = fma(t01, a1, t00 * a0), ta1 = fma(t11, a1, t10 * a0);
        observe(yv, F_star, za);
        const double g0 = tm0 * F1, g1 = tm1 * F1;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          const double2 mJ = L > 1 ? xch[buf][J][g] : make_double2(tm0, tm1);  // (T M)_J
          P[J][0] = fma(-g0, mJ.x, P[J][0]);
          P[J][1] = fma(-g0, mJ.y, P[J][1]);
          P[J][2] = fma(-g1, mJ.x, P[J][2]);
          P[J][3] = fma(-g1, mJ.y, P[J][3]);
        }
        buf ^= 1;
        a0 = fma(g0, vv, ta0);
        a1 = fma(g1, vv, ta1);
      }
      // the last observation (:200 -- P and a are not needed afterwards)
      load_z(N - 1);
      innovation();
      F_star = group_sum(fma(zi1, m1, zi0 * m0));
      za = group_sum(fma(zi1, a1, zi0 * a0));
      observe(ynext, F_star, za);
    }
  } else if (t < N) {
    long long idx = (long long)t * p;
    double ynext = A.y[idx * A.B + b];  // observations are fetched one univariate step ahead of their use
    for (; t < N; t++) {
      const int zo = t >> 30;  // z from the parameter bank; always 0 (N < 2^30), but loop-variant: no hoisting
#pragma unroll 1
      for (int j = 0; j < p; j++, idx++) {
        const double yv = ynext;
        if (idx + 1 < Np) ynext = A.y[(idx + 1) * A.B + b];
        const double* zp = &A.z[0][0] + zo + j * (2 * kBlkMaxNB);
        const double zi0 = A.z[j][2 * I], zi1 = A.z[j][2 * I + 1];  // own part: a per-lane address, kept apart from zo
        double m0a = 0.0, m0b = 0.0, m1a = 0.0, m1b = 0.0;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          const double z0 = zp[2 * J], z1 = zp[2 * J + 1];
          m0a = fma(P[J][0], z0, m0a);
          m0b = fma(P[J][1], z1, m0b);
          m1a = fma(P[J][2], z0, m1a);
          m1b = fma(P[J][3], z1, m1b);
        }
        const double m0 = m0a + m0b, m1 = m1a + m1b;
        if (own) xch[buf][I][g] = make_double2(m0, m1);
        const double F_star = group_sum(fma(zi1, m1, zi0 * m0));
        const double za = group_sum(fma(zi1, a1, zi0 * a0));
        __syncwarp();
        observe(yv, F_star, za);
        const double k0 = m0 * F1, k1 = m1 * F1;
#pragma unroll
        for (int J = 0; J < NB; J++) {
          const double2 mJ = xch[buf][J][g];
          P[J][0] = fma(-k0, mJ.x, P[J][0]);
          P[J][1] = fma(-k0, mJ.y, P[J][1]);
          P[J][2] = fma(-k1, mJ.x, P[J][2]);
          P[J][3] = fma(-k1, mJ.y, P[J][3]);
        }
        buf ^= 1;
        a0 = fma(k0, vv, a0);
        a1 = fma(k1, vv, a1);
      }
      if (t < N - 1) {  // :200-204 (after the last observation P and a are not needed any more)
#pragma unroll
        for (int J = 0; J < NB; J++) time_block(J, P[J][0], P[J][1], P[J][2], P[J][3], true);
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

template <int L, int NB, bool P1, bool QOFF, bool ZTV = false>
static void launch_blk_one(const BlkArgs& A, unsigned blocks, cudaStream_t st) {
  // SSGPU_BLK_PAD_SMEM: unused dynamic shared memory per CTA -- an occupancy knob for experiments (e.g. 120000 leaves
  // one CTA per SM: the step time of a warp that has a sub-partition to itself)
  static const int pad = [] {
    const char* e = getenv("SSGPU_BLK_PAD_SMEM");
    return e ? atoi(e) : 0;
  }();
  if (pad > 0)
    cudaFuncSetAttribute(loglik_blk_kernel<L, NB, P1, QOFF, ZTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
  loglik_blk_kernel<L, NB, P1, QOFF, ZTV><<<blocks, NB > 8 ? kBlkThreadsCS : kBlkThreads, pad, st>>>(A);
}

template <int L, int NB>
static void launch_blk_nb(const BlkArgs& A, bool q_offdiag, unsigned blocks, cudaStream_t st) {
  if constexpr (NB > 8) {  // p = 1, diagonal R Q R' (checked by the caller)
    if (A.zt)
      launch_blk_one<L, NB, true, false, true>(A, blocks, st);
    else
      launch_blk_one<L, NB, true, false>(A, blocks, st);
    return;
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
int blk_flops_per_step(int nb, int p) {
  const int L = blk_lanes(nb);
  int lg = 0;
  while ((1 << lg) < L) lg++;
  const int fma_obs = 4 * nb + 2 + 4 + 4 * nb + 2 + 1, mul_obs = 2 + 2 + 3, add_obs = 2 + 2 * lg;
  const int fma_t = 12 * nb + 2, mul_t = 4 * nb + 2;
  const int tm = p == 1 ? 2 * 2 + 2 : 0;  // p = 1: (T M)_I
  return L * (p * (2 * fma_obs + mul_obs + add_obs) + 2 * fma_t + mul_t + tm);
}

int launch_loglik_blk(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                      const HostShared* hs, bool* applicable) {
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
  const int L = blk_lanes(pl.nb), G = 32 / L, wpb = (pl.nb > 8 ? kBlkThreadsCS : kBlkThreads) / 32;
  const long long blocks = (A.E + (long long)wpb * G - 1) / ((long long)wpb * G);
  SS_ARG(blocks < (1LL << 31), "block kernel: batch too large for one launch");
  const unsigned nblk = (unsigned)blocks;
  switch (pl.nb) {
    case 1: launch_blk_nb<1, 1>(A, pl.q_offdiag != 0, nblk, st); break;
    case 2: launch_blk_nb<2, 2>(A, pl.q_offdiag != 0, nblk, st); break;
    case 3: launch_blk_nb<4, 3>(A, pl.q_offdiag != 0, nblk, st); break;
    case 4: launch_blk_nb<4, 4>(A, pl.q_offdiag != 0, nblk, st); break;
    case 5: launch_blk_nb<8, 5>(A, pl.q_offdiag != 0, nblk, st); break;
    case 6: launch_blk_nb<8, 6>(A, pl.q_offdiag != 0, nblk, st); break;
    case 7: launch_blk_nb<8, 7>(A, pl.q_offdiag != 0, nblk, st); break;
    case 8: launch_blk_nb<8, 8>(A, pl.q_offdiag != 0, nblk, st); break;
    case 9: launch_blk_nb<16, 9>(A, false, nblk, st); break;
    case 10: launch_blk_nb<16, 10>(A, false, nblk, st); break;
    case 11: launch_blk_nb<16, 11>(A, false, nblk, st); break;
    case 12: launch_blk_nb<16, 12>(A, false, nblk, st); break;
    case 13: launch_blk_nb<16, 13>(A, false, nblk, st); break;
    case 14: launch_blk_nb<16, 14>(A, false, nblk, st); break;
    case 15: launch_blk_nb<16, 15>(A, false, nblk, st); break;
    default: launch_blk_nb<16, 16>(A, false, nblk, st); break;
  }
  SS_CUDA(cudaGetLastError());
  count_launch();
  static thread_local char name[96];
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
  if (flop_per_step) *flop_per_step = blk_flops_per_step(pl.nb, p);
  return SSGPU_OK;
}
