// Batched loglikelihood on the FP64 tensor path: one evaluation (series b, parameter vector v) per
// warp, every m x m object held in registers as mma.sync.m8n8k4.f64 accumulator fragments.
//
// Restates the recursion of src/LogLikC.cpp:68-214 (univariate treatment, exact diffuse
// initialisation, NaN = missing) for models whose Z, T, R are shared by the batch and time-invariant
// (per-evaluation a, P_inf, P_star, Q allowed), m <= 23.
//
// Why this shape.  The time update  P <- T P T' + R Q R'  (src/LogLikC.cpp:202) is 4 m^3 of the
// 4 m^3 + 7 m^2 flops of a step: a dense FP64 contraction.  sm_100 has no tcgen05 kind for f64; its FP64
// tensor path is DMMA (mma.sync.m8n8k4.f64), measured at the same 37 TFLOP/s as the DFMA pipe
// (ssgpu_fp64_peak mode 1) but needing 1/32 of the instructions and no operand traffic at all:
//   * a D x D matrix X (D = 8 ceil((m+1)/8)) lives in "C layout": lane (g = lane/4, q = lane%4) holds
//     X[8rt+g][8ct+2q+e], e = 0,1 -- the accumulator fragment of tile (rt, ct);
//   * with the k index of the contraction permuted as k = 8ct + 2q + e  <->  (k-step (ct,e), slot q),
//     those same registers ARE the A fragment of X and the B fragment of X': any product  X W'  of two
//     C-layout matrices is 2 TT^3 DMMAs with no shuffle, no shared memory, no constant load;
//   * the state vector rides along as row m of  G = [P ; a'] :
//         Y  = T G'              = [T P , T a]            (a-row of G becomes column m of Y, unused)
//         G' = [Y ; a'] T' + RQR' = [T P T' + R Q R' ; (T a)']
//     (column m of Y meets the zero padding column m of T, so it drops out);
//   * the measurement update is a rank-1 correction in place: row sums by two quad shuffles, F = z'Pz and
//     the gather of M[c] by a few more; v, K, a += K v come out of the same formulas through row m.
// One evaluation per warp makes every branch of the recursion (diffuse / regular / skipped step) warp
// uniform: there is no divergence and no predication anywhere.
// Measurement update in the rank-1 forms P - M K0' (src/LogLikC.cpp:129,193), P* - M* K0' - M_inf K1',
// P_inf - M_inf K0' (:151-152); sum(log F) carried as (mantissa product, exponent sum).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tile_plan.cuh"

namespace ssgpu {

constexpr int kMmaThreads = 128;  // 4 warps = 4 evaluations per CTA
constexpr int kMmaMaxM = 23;

struct MmaArgs {
  DModel md;
  const double* y;  // [(t*p + j) * B + b]
  long long B;
  long long E;  // B * V
  double* out;
  signed char perm[24];  // internal state order: position i holds state perm[i] (plan_tiles)
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// D x D matrix in C layout: c[rt][ct][e] = X[8rt+g][8ct+2q+e]
template <int TT>
struct Frag {
  double c[TT][TT][2];
};

// TM: which 8 x 8 tiles of the (padded, internally ordered) T hold a non-zero, bit rt*TT + ct.  The host orders
// the states so that T becomes as tile-sparse as its coupling graph allows (plan_tiles below: statespacer's
// T is block diagonal by construction, R/GetSysMat.R:1388-1426) and the DMMAs of all-zero tiles are not
// compiled in -- a template parameter, not a branch, so the DMMA stream stays one straight-line block.
template <int TT, unsigned TM>
__device__ __forceinline__ constexpr bool tile_on(int rt, int ct) {
  return (TM >> (rt * TT + ct)) & 1u;
}

// Y = init + T W'   (all C layout): tile (rt, nt) += T(rt, ct) W(nt, ct)'
template <int TT, unsigned TM>
__device__ __forceinline__ void mul_t_wt(Frag<TT>& Y, const Frag<TT>& T, const Frag<TT>& W) {
  // k-steps outermost: the DMMAs that accumulate into one tile are as far apart as the tile count allows
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int ct = 0; ct < TT; ct++)
#pragma unroll
      for (int rt = 0; rt < TT; rt++)
#pragma unroll
        for (int nt = 0; nt < TT; nt++)
          if (tile_on<TT, TM>(rt, ct)) dmma(Y.c[rt][nt][0], Y.c[rt][nt][1], T.c[rt][ct][e], W.c[nt][ct][e]);
}

// Z = init + Y T', only the tiles on or above the diagonal (rt <= nt): for a symmetric result the rest is
// mirrored by the caller -- TT(TT-1) TT fewer DMMAs, paid with a few shuffles on otherwise idle pipes.
template <int TT, unsigned TM>
__device__ __forceinline__ void mul_y_tt_upper(Frag<TT>& Z, const Frag<TT>& Y, const Frag<TT>& T) {
#pragma unroll
  for (int e = 0; e < 2; e++)
#pragma unroll
    for (int ct = 0; ct < TT; ct++)
#pragma unroll
      for (int rt = 0; rt < TT; rt++)
#pragma unroll
        for (int nt = rt; nt < TT; nt++)
          if (tile_on<TT, TM>(nt, ct)) dmma(Z.c[rt][nt][0], Z.c[rt][nt][1], Y.c[rt][ct][e], T.c[nt][ct][e]);
}

// 1 / x for a finite, normal x well inside the exponent range (F >= 1e-7 here): MUFU.RCP64H seed (about 20
// bits) and two Newton steps; no special-case branch, the result is within an ulp or two of the rounded one.
__device__ __forceinline__ double rcp_newton(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

template <int TT, unsigned TM, bool P1, int MINB>
__global__ void __launch_bounds__(kMmaThreads, MINB) loglik_mma_kernel(const MmaArgs A) {
  constexpr unsigned FULL = 0xffffffffu;
  // The state vector rides along as the LAST row of the D x D fragment matrix (row MA = D - 1, a compile-time
  // position: tile row TT-1, g = 7; as a column: tile TT-1, q = 3, e = 1).  Rows / columns M .. MA-1 are zero
  // padding and stay exactly zero through every operation below, so the hot loop carries no "i < M" masks.
  constexpr int MA = 8 * TT - 1, LA = TT - 1;
  const DModel& md = A.md;
  const int N = md.N, p = P1 ? 1 : md.p, M = md.m, r = md.r;
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
  const bool lane_a = g == 7;  // this lane holds row MA in the fragments of tile row LA
  // warps are numbered series-major (w = b*V + v) so that the V parameter vectors of a series run side by
  // side and share its y through L1/L2; results keep the public order e = v*B + b
  const long long w_id = (long long)blockIdx.x * (kMmaThreads / 32) + (threadIdx.x >> 5);
  if (w_id >= A.E) return;  // whole warp
  const long long V = A.E / A.B, b = w_id / V, v = w_id % V;
  const long long e_id = v * A.B + b;

  auto row_of = [&](int rt) { return 8 * rt + g; };
  auto col_of = [&](int ct, int e) { return 8 * ct + 2 * q + e; };
  auto pm = [&](int i) { return (int)A.perm[i]; };  // internal position -> state index of the caller

  // ---- shared matrices into fragments: T (zero padded), z --------------------------------------
  Frag<TT> T, G;
  // R Q R' (accumulator init of the second product) waits in shared memory: [element][lane], conflict-free
  __shared__ double rqr_s[kMmaThreads / 32][TT * TT * 2][32];
  double(*rqr)[32] = rqr_s[threadIdx.x >> 5];
  {
    const double* Tb = md.T.block(0, 0, 0);
    const double* ab = md.a.block(b, v, 0);
    const double* Ps = md.P_star.block(b, v, 0);
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = row_of(rt), k = col_of(ct, e);
          const bool in = i < M && k < M;
          T.c[rt][ct][e] = in ? mat_get(Tb, md.T.diag, M, pm(i), pm(k)) : 0.0;
          G.c[rt][ct][e] = in ? mat_get(Ps, md.P_star.diag, M, pm(i), pm(k)) : ((i == MA && k < M) ? ab[pm(k)] : 0.0);
        }
    // R Q R' (src/LogLikC.cpp:202, hoisted: Q and R are time-invariant here)
    const double* Qb = md.Q.block(b, v, 0);
    const double* Rb = md.R.block(0, 0, 0);
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = row_of(rt), cidx = col_of(ct, e);
          double acc = 0.0;
          if (i < M && cidx < M) {
            const int pi = pm(i), pc = pm(cidx);
            for (int k = 0; k < r; k++) {
              const double rik = mat_get(Rb, md.R.diag, M, pi, k);
              if (rik == 0.0) continue;
              double qr;
              if (md.Q.diag) {
                qr = Qb[k] * mat_get(Rb, md.R.diag, M, pc, k);
              } else {
                qr = 0.0;
                for (int l = 0; l < r; l++) qr = fma(Qb[k + (long long)l * r], mat_get(Rb, md.R.diag, M, pc, l), qr);
              }
              acc = fma(rik, qr, acc);
            }
          }
          rqr[(rt * TT + ct) * 2 + e][lane] = acc;
        }
  }
  // row j of Z as column values zc[ct][e] = z[col] and row values zr[rt] = z[row] (rows >= M: 0)
  double zc[TT][2], zr[TT];
  auto load_z = [&](int j) {
    const double* Zb = md.Z.block(0, 0, 0);
#pragma unroll
    for (int ct = 0; ct < TT; ct++) {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int k = col_of(ct, e);
        zc[ct][e] = k < M ? Zb[j + (long long)pm(k) * p] : 0.0;
      }
      const int i = row_of(ct);
      zr[ct] = i < M ? Zb[j + (long long)pm(i) * p] : 0.0;
    }
  };
  if (P1) load_z(0);

  // sum over the quad (columns) / over the 8 quads (rows)
  auto quad_sum = [&](double x) {
    x += __shfl_xor_sync(FULL, x, 1);
    x += __shfl_xor_sync(FULL, x, 2);
    return x;
  };
  auto rows_sum = [&](double x) {
    x += __shfl_xor_sync(FULL, x, 4);
    x += __shfl_xor_sync(FULL, x, 8);
    x += __shfl_xor_sync(FULL, x, 16);
    return x;
  };
  // m[rt] = (X z)[row]: known to the 4 lanes of the quad that owns the row (row MA: z'a)
  auto row_dots = [&](const Frag<TT>& X, double (&m)[TT]) {
#pragma unroll
    for (int rt = 0; rt < TT; rt++) {
      double acc = 0.0;
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) acc = fma(X.c[rt][ct][e], zc[ct][e], acc);
      m[rt] = quad_sum(acc);
    }
  };
  // mc[ct][e] = m[col]: row `col` belongs to quad (2q+e), slot ct.  Column MA receives z'a instead of a zero;
  // it only ever meets the zero column MA of T and is overwritten by the next time update.
  auto to_cols = [&](const double (&m)[TT], double (&mc)[TT][2]) {
#pragma unroll
    for (int ct = 0; ct < TT; ct++)
#pragma unroll
      for (int e = 0; e < 2; e++) mc[ct][e] = __shfl_sync(FULL, m[ct], (2 * q + e) * 4);
  };
  // z'a = entry MA of m
  auto z_dot_a = [&](const double (&m)[TT]) { return __shfl_sync(FULL, m[LA], 7 * 4); };
  auto abs_small = [&](const Frag<TT>& X) {
    bool ok = true;
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) ok = ok && (fabs(X.c[rt][ct][e]) < kEps);
    return __all_sync(FULL, ok) != 0;
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

  // y: 32 flat indices (t*p + j) at a time, one per lane, one chunk ahead
  const long long Np = (long long)N * p;
  auto load_y = [&](long long qq) { return qq < Np ? A.y[qq * A.B + b] : 0.0; };
  double ybuf = 0.0, ynext = load_y(lane);
  long long qi = 0;
  auto next_y = [&]() {
    const int ql = (int)(qi & 31);
    if (ql == 0) {
      ybuf = ynext;
      ynext = load_y(qi + 32 + lane);
    }
    qi++;
    return __shfl_sync(FULL, ybuf, ql);
  };

  // rank-1 measurement update G <- G - rc mc' with rc = M F1 on the state rows and -v F1 on row MA (a += K0 v)
  auto rank1 = [&](const double (&ms)[TT], double F1, double vv) {
    double mc[TT][2];
    to_cols(ms, mc);
#pragma unroll
    for (int rt = 0; rt < TT; rt++) {
      double rc = ms[rt] * F1;
      if (rt == LA) rc = lane_a ? -vv * F1 : rc;
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) G.c[rt][ct][e] = fma(-rc, mc[ct][e], G.c[rt][ct][e]);
    }
  };

  // P* L0' = P* - M K0' applied to G = [P* ; a'] (src/LogLikC.cpp:181-194 / 117-130)
  auto update_star = [&](const double (&ms)[TT], double F_star, double yv) {
    const double F1 = __drcp_rn(F_star);
    const double vv = yv - z_dot_a(ms);
    rank1(ms, F1, vv);
    acc_log(F_star);
    n_terms++;
    sumq = fma(vv * vv, F1, sumq);
  };

  // Second product of the time update, X = init + [Y ; a'] T', exploiting the symmetry of T P T' + R Q R':
  // tiles below the diagonal are not multiplied but mirrored from above (8x8 tile transpose in accumulator
  // layout: element (2q+e, g) sits in lane (2q+e)*4 + g/2, half g%2), and the part of the state row that
  // falls into them, (T a)[n], is fetched from column MA of Y = [T P , T a] where the first product left it.
  auto second_product = [&](Frag<TT>& X, const Frag<TT>& Y, bool keep_a) {
    mul_y_tt_upper<TT, TM>(X, Y, T);
    if constexpr (TT > 1) {
#pragma unroll
      for (int rt = 1; rt < TT; rt++)
#pragma unroll
        for (int nt = 0; nt < rt; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int src = (2 * q + e) * 4 + (g >> 1);
            const double lo = __shfl_sync(FULL, X.c[nt][rt][0], src);
            const double hi = __shfl_sync(FULL, X.c[nt][rt][1], src);
            double val = (g & 1) ? hi : lo;  // X[8nt + 2q+e][8rt + g]
            if (rt == LA) {
              // (T a)[8nt + 2q + e] = Y[8nt+2q+e][MA]
              const double ta = __shfl_sync(FULL, Y.c[nt][LA][1], (2 * q + e) * 4 + 3);
              if (keep_a && lane_a) val = ta;
            }
            X.c[rt][nt][e] = val;
          }
    }
  };

  // [Y ; a'] T' with Y = T G'  (+ R Q R')
  auto time_update = [&](Frag<TT>& X, bool with_rqr, bool keep_a) {
    Frag<TT> Y;
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++) Y.c[rt][ct][0] = Y.c[rt][ct][1] = 0.0;
    mul_t_wt<TT, TM>(Y, T, X);  // T X' = [T P , T a]
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          if (rt == LA && keep_a && lane_a) Y.c[rt][ct][e] = X.c[rt][ct][e];  // row MA of Y <- a'
          X.c[rt][ct][e] = with_rqr ? rqr[(rt * TT + ct) * 2 + e][lane] : 0.0;
        }
    second_product(X, Y, keep_a);  // [T P ; a'] T' = [T P T' ; (T a)']
  };

  int t = 0;
  {
    // ---- exact diffuse phase (src/LogLikC.cpp:98-159) ------------------------------------------
    Frag<TT> Ginf;
    const double* Pi = md.P_inf.block(b, v, 0);
#pragma unroll
    for (int rt = 0; rt < TT; rt++)
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int i = row_of(rt), k = col_of(ct, e);
          Ginf.c[rt][ct][e] = (i < M && k < M) ? mat_get(Pi, md.P_inf.diag, M, pm(i), pm(k)) : 0.0;
        }
    bool init = !abs_small(Ginf);  // :49
    for (; t < N && init; t++) {
      for (int j = 0; j < p; j++) {
        const double yv = next_y();
        if (isnan(yv)) continue;  // :88-90
        if (!P1) load_z(j);
        double ms[TT];
        row_dots(G, ms);
        double fs = 0.0;
#pragma unroll
        for (int rt = 0; rt < TT; rt++) fs = fma(zr[rt], ms[rt], fs);
        const double F_star = rows_sum(fs);  // lanes of equal q: every row once
        if (!init) {  // P_inf vanished earlier in this time point (p > 1): regular step
          non_init++;
          if (F_star < kEps)
            F_eq0++;
          else
            update_star(ms, F_star, yv);
          continue;
        }
        double mi[TT];
        row_dots(Ginf, mi);
        double fi = 0.0;
#pragma unroll
        for (int rt = 0; rt < TT; rt++) fi = fma(zr[rt], mi[rt], fi);
        const double F_inf = rows_sum(fi);
        if (F_inf < kEps) {
          if (F_star < kEps) continue;  // :112-113
          update_star(ms, F_star, yv);  // :117-131
          continue;
        }
        // :136-153  K0 = M_inf F1, K1 = M* F1 + M_inf F2
        const double F1 = __drcp_rn(F_inf), F2 = -(F1 * F1) * F_star;
        const double vv = yv - z_dot_a(ms);
        double msc[TT][2], mic[TT][2];
        to_cols(ms, msc);
        to_cols(mi, mic);
#pragma unroll
        for (int rt = 0; rt < TT; rt++) {
          // row MA: M* -> -v, M_inf -> 0  (a += K0 v)
          const double rs = (rt == LA && lane_a) ? -vv : ms[rt];
          const double ri = (rt == LA && lane_a) ? 0.0 : mi[rt];
#pragma unroll
          for (int ct = 0; ct < TT; ct++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const double k0 = mic[ct][e] * F1, k1 = msc[ct][e] * F1 + mic[ct][e] * F2;
              // P* - M* K0' - M_inf K1'
              G.c[rt][ct][e] = fma(-ri, k1, fma(-rs, k0, G.c[rt][ct][e]));
              Ginf.c[rt][ct][e] = fma(-ri, k0, Ginf.c[rt][ct][e]);
            }
        }
        acc_log(F_inf);  // no v^2 term (:153)
        n_terms++;
        if (j < p - 1) init = !abs_small(Ginf);  // :157-159
      }
      if (t < N - 1) {  // :200-207
        time_update(G, true, true);
        if (init) {
          time_update(Ginf, false, false);
          init = !abs_small(Ginf);
        }
      }
    }
  }

  // ---- regular phase: the hot loop ---------------------------------------------------------------
  if constexpr (P1) {
    // Software-pipelined, branch-free form for p = 1.  With P+ = T P T' + R Q R' (the second product below),
    //     M+ = P+ z = [T P] (T'z) + (R Q R') z,        z'a+ = a'(T'z):
    // the next observation's M, F = z'M and 1/F only need the FIRST product Y = [T P ; a'], so their shuffle /
    // reciprocal chain is issued next to the DMMAs of the second product instead of after them.  A missing
    // observation or F < 1e-7 (src/LogLikC.cpp:88-90,173-176) turns the update into a no-op through
    // F1 = v = 0 rather than through a branch, so that the whole step is one straight-line block.
    double wc[TT][2], c0[TT];
    {
      const double* Tb = md.T.block(0, 0, 0);
      const double* Zb = md.Z.block(0, 0, 0);
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) {  // w = T'z
          const int k = col_of(ct, e);
          double acc = 0.0;
          if (k < M)
            for (int i = 0; i < M; i++) acc = fma(mat_get(Tb, md.T.diag, M, i, pm(k)), Zb[(long long)i], acc);
          wc[ct][e] = acc;
        }
#pragma unroll
      for (int rt = 0; rt < TT; rt++) {  // c0 = (R Q R') z
        double acc = 0.0;
#pragma unroll
        for (int ct = 0; ct < TT; ct++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc = fma(rqr[(rt * TT + ct) * 2 + e][lane], zc[ct][e], acc);
        c0[rt] = quad_sum(acc);
      }
    }
    // the per-lane constants of the hot loop wait in shared memory like R Q R' (registers are what limits the
    // number of resident warps): [w (2 TT) | c0 (TT) | z rows (TT)][lane]
    __shared__ double cst_s[kMmaThreads / 32][4 * TT][32];
    double(*cst)[32] = cst_s[threadIdx.x >> 5];
#pragma unroll
    for (int ct = 0; ct < TT; ct++) {
      cst[2 * ct][lane] = wc[ct][0];
      cst[2 * ct + 1][lane] = wc[ct][1];
      cst[2 * TT + ct][lane] = 0.25 * c0[ct];  // a quarter per lane of the quad: the quad sum adds it back
      cst[3 * TT + ct][lane] = zr[ct];
    }
    __syncwarp();
    double ms[TT], F_star = 1.0, F1 = 1.0, zat = 0.0;
    // F = z'M, 1/F, z'a from ms
    auto prepare = [&]() {
      double fs = 0.0;
#pragma unroll
      for (int rt = 0; rt < TT; rt++) fs = fma(cst[3 * TT + rt][lane], ms[rt], fs);
      F_star = rows_sum(fs);
      F1 = rcp_newton(F_star < kEps ? 1.0 : F_star);
      zat = z_dot_a(ms);
    };
    // loglik terms of one observation (:88-90, :173-195); returns the update's 1/F and v, both 0 for a no-op
    double F1e, vv;
    auto observe = [&](double yv) {
      const bool obs = !isnan(yv);
      const bool upd = obs && !(F_star < kEps);
      F1e = upd ? F1 : 0.0;
      vv = upd ? yv - zat : 0.0;
      acc_log(upd ? F_star : 1.0);
      n_terms += upd;
      non_init += obs;
      F_eq0 += obs && !upd;
      sumq = fma(vv * vv, F1e, sumq);
    };
    if (t < N) {
      row_dots(G, ms);
      prepare();
    }
    // y in chunks of 32 time points, one chunk ahead
    int t0 = t;
    double ycur, ynxt = load_y(t0 + lane);
    while (t0 < N) {
      ycur = ynxt;
      ynxt = load_y((long long)t0 + 32 + lane);
      const int nl = min(32, N - t0);
      for (int l = 0; l < nl; l++) {
        observe(__shfl_sync(FULL, ycur, l));
        if (t0 + l < N - 1) {  // :200-207 (after the last observation P and a are not needed any more)
          rank1(ms, F1e, vv);
          Frag<TT> Y;
#pragma unroll
          for (int rt = 0; rt < TT; rt++)
#pragma unroll
            for (int ct = 0; ct < TT; ct++) Y.c[rt][ct][0] = Y.c[rt][ct][1] = 0.0;
          mul_t_wt<TT, TM>(Y, T, G);  // T G' = [T P , T a]
#pragma unroll
          for (int ct = 0; ct < TT; ct++)
#pragma unroll
            for (int e = 0; e < 2; e++)
              if (lane_a) Y.c[LA][ct][e] = G.c[LA][ct][e];  // row MA of Y <- a'
          // next observation's M (state rows) and z'a (row MA) from Y
#pragma unroll
          for (int rt = 0; rt < TT; rt++) {
            double acc = cst[2 * TT + rt][lane];
#pragma unroll
            for (int ct = 0; ct < TT; ct++) {
              acc = fma(Y.c[rt][ct][0], cst[2 * ct][lane], acc);
              acc = fma(Y.c[rt][ct][1], cst[2 * ct + 1][lane], acc);
            }
            ms[rt] = quad_sum(acc);
          }
#pragma unroll
          for (int rt = 0; rt < TT; rt++)
#pragma unroll
            for (int ct = 0; ct < TT; ct++)
#pragma unroll
              for (int e = 0; e < 2; e++) G.c[rt][ct][e] = rqr[(rt * TT + ct) * 2 + e][lane];
          second_product(G, Y, true);  // [T P ; a'] T' + R Q R'
          prepare();
        }
      }
      t0 += nl;
    }
  } else {
    for (; t < N; t++) {
      for (int j = 0; j < p; j++) {
        const double yv = next_y();
        if (isnan(yv)) continue;
        load_z(j);
        non_init++;
        double ms[TT];
        row_dots(G, ms);
        double fs = 0.0;
#pragma unroll
        for (int rt = 0; rt < TT; rt++) fs = fma(zr[rt], ms[rt], fs);
        const double F_star = rows_sum(fs);
        if (F_star < kEps)
          F_eq0++;  // :173-176
        else
          update_star(ms, F_star, yv);
      }
      if (t < N - 1) time_update(G, true, true);
    }
  }

  if (lane == 0) {
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
bool mma_eligible(const DModel& md, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  if (md.m > kMmaMaxM) return no("m > 23");
  if (!md.Z.shared() || !md.T.shared() || !md.R.shared()) return no("Z, T, R must be shared by the batch");
  if (md.Z.tv() || md.T.tv() || md.R.tv() || md.Q.tv()) return no("time-varying system matrices");
  return true;
}

template <int TT, unsigned TM, int MINB>
static void launch_mma_tt(const MmaArgs& A, unsigned blocks, cudaStream_t st) {
  if (A.md.p == 1)
    loglik_mma_kernel<TT, TM, true, MINB><<<blocks, kMmaThreads, 0, st>>>(A);
  else
    loglik_mma_kernel<TT, TM, false, MINB><<<blocks, kMmaThreads, 0, st>>>(A);
}

// DMMAs per time update for tile mask `mask` (first product: all column tiles; second: upper tiles only)
int mma_dmma_per_step(int tt, unsigned mask) {
  int n = 0;
  for (int rt = 0; rt < tt; rt++)
    for (int ct = 0; ct < tt; ct++)
      if ((mask >> (rt * tt + ct)) & 1u) n += 2 * tt + 2 * (rt + 1);  // T(rt,ct): tt tiles of Y; T(nt=rt,ct): rt+1 tiles
  return n;
}

int launch_loglik_mma(const DModel& md, const double* y, long long B, int V, double* out, bool tile_sparse,
                      cudaStream_t st, const double* T_host) {
  std::string why;
  SS_ARG(mma_eligible(md, &why), "tensor kernel not eligible: " + why);
  MmaArgs A{};
  A.md = md;
  A.y = y;
  A.B = B;
  A.E = B * V;
  A.out = out;
  const int wpb = kMmaThreads / 32;
  const long long blocks = (A.E + wpb - 1) / wpb;
  SS_ARG(blocks < (1LL << 31), "tensor kernel: batch too large for one launch");
  const int tt = (md.m + 1 + 7) / 8;
  // T is shared and small: fetch it to plan the state order (a few microseconds next to a batched launch)
  std::vector<double> Th((size_t)(md.T.diag ? md.m : md.m * md.m));
  if (tile_sparse && tt > 1) {
    if (T_host) {
      std::copy(T_host, T_host + Th.size(), Th.begin());
    } else {
      SS_CUDA(cudaMemcpyAsync(Th.data(), md.T.p, Th.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaStreamSynchronize(st));
    }
  }
  TilePlan pl = plan_tiles(Th, md.T.diag, md.m, tt, tile_sparse);
  const unsigned dense = (1u << (tt * tt)) - 1u;
  constexpr unsigned kDiag2 = 0x9u, kDiag3 = 0x111u;
  const unsigned plmask = (unsigned)pl.mask;
  unsigned use = dense;
  if (tt == 2 && (plmask & ~kDiag2) == 0) use = kDiag2;
  if (tt == 3 && (plmask & ~kDiag3) == 0) use = kDiag3;
  if (use == dense)  // nothing to gain: keep the caller's order
    for (int i = 0; i < 24; i++) pl.perm[i] = (signed char)(i < md.m ? i : 0);
  for (int i = 0; i < 24; i++) A.perm[i] = pl.perm[i];
  // resident CTAs per SM the register budget is sized for (tuning knob)
  static const int minb = [] {
    const char* e = getenv("SSGPU_MMA_MINB");
    return e ? atoi(e) : 5;  // 96 registers, 5 warps per sub-partition: measured best on B200 (profiles/README.md)
  }();
  if (tt == 1)
    launch_mma_tt<1, 0x1u, 6>(A, (unsigned)blocks, st);
  else if (tt == 2) {
    if (use == kDiag2) {
      if (minb <= 4)
        launch_mma_tt<2, kDiag2, 4>(A, (unsigned)blocks, st);
      else if (minb == 5)
        launch_mma_tt<2, kDiag2, 5>(A, (unsigned)blocks, st);
      else
        launch_mma_tt<2, kDiag2, 6>(A, (unsigned)blocks, st);
    } else {
      launch_mma_tt<2, 0xFu, 4>(A, (unsigned)blocks, st);
    }
  } else {
    if (use == kDiag3)
      launch_mma_tt<3, kDiag3, 2>(A, (unsigned)blocks, st);
    else
      launch_mma_tt<3, 0x1FFu, 2>(A, (unsigned)blocks, st);
  }
  SS_CUDA(cudaGetLastError());
  count_launch();
  static thread_local char name[96];
  snprintf(name, sizeof name, "loglik_mma_kernel<TT=%d,tiles=0x%x,dmma_per_step=%d>", tt, use,
           mma_dmma_per_step(tt, use));
  set_last_kernel(name);
  return SSGPU_OK;
}

}  // namespace ssgpu

extern "C" int ssgpu_plan_state_order(const double* T, int m, int* perm, unsigned* tile_mask, int* dmma_per_step) {
  using namespace ssgpu;
  SS_ARG(T != nullptr && m >= 1 && m <= kMmaMaxM, "ssgpu_plan_state_order: need T and 1 <= m <= 23");
  const int tt = (m + 1 + 7) / 8;
  std::vector<double> Th(T, T + (size_t)m * m);
  const TilePlan pl = plan_tiles(Th, 0, m, tt, true);
  if (perm)
    for (int i = 0; i < m; i++) perm[i] = pl.perm[i];
  if (tile_mask) *tile_mask = (unsigned)pl.mask;
  if (dmma_per_step) *dmma_per_step = mma_dmma_per_step(tt, (unsigned)pl.mask);
  return SSGPU_OK;
}
