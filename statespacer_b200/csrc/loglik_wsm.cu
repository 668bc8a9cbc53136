// Batched loglikelihood for larger state dimensions (m <= 63): one evaluation per warp, the covariance in SHARED
// memory, the time update on the FP64 tensor path.
//
// Same recursion as loglik_mma.cu (src/LogLikC.cpp:68-214: univariate treatment, exact diffuse initialisation,
// NaN = missing) and the same data layout -- every D x D matrix (D = 8 TT >= m + 1) as mma.sync.m8n8k4.f64
// accumulator fragments, element X[8rt+g][8ct+2q+e] owned by lane (g, q) -- but the fragments of P live in shared
// memory as [tile][e][lane] (each lane reads and writes only its own slots: conflict-free, no transposition ever)
// because TT^2 tiles no longer fit the register file:
//   * per warp two fragment matrices (current P and the one being built), 2 TT^2 512 B; T once per CTA;
//   * the time update goes row tile by row tile:  Y(rt, .) = T(rt, .) G' is accumulated in registers and, being in
//     accumulator layout, IS the A operand of the second product  X(rt, .) = Y(rt, .) T' + R Q R'(rt, .);  only T's
//     and G's fragments are fetched from shared memory (one 8-byte load per lane and DMMA operand);
//   * the state vector rides along as row D-1 of G (loglik_mma.cu);
//   * R Q R' per evaluation and, during the d diffuse steps, P_inf wait in a global scratch area sized by the
//     number of resident warps (the kernel is persistent: warps loop over evaluations);
//   * the host orders the states so that a block-diagonal T becomes tile diagonal (tile_plan.cuh) and the kernel
//     skips all-zero tiles of T through a warp-uniform mask test.
// Serves models whose Z, T, R are shared by the batch with T, R, Q time-invariant; Z may be time-varying
// (explanatory variables, R/Components-Regressors.R:60-78): BASELINE.json config 2 (m = 34, p = 2).
#include <cstdlib>

#include "common.cuh"
#include "mm_dmma.cuh"
#include "tile_plan.cuh"

namespace ssgpu {

struct WsmArgs {
  DModel md;
  const double* y;  // [(t*p + j) * B + b]
  long long B, E;
  double* out;
  double* scratch;  // per resident warp: rqr | P_inf, each TT*TT*2*32 doubles
  unsigned long long mask;
  int wpb;  // warps per CTA
  signed char perm[64];
};

// DIAG: T is tile diagonal in the internal state order -- the common case after re-ordering, compiled without any
// mask test (ptxas turns the warp-uniform test of the general path into PREDICATED DMMAs in the fully unrolled second
// product, and a predicated-off DMMA still occupies the FP64 datapath).
template <int TT, bool DIAG>
__global__ void __launch_bounds__(256) loglik_wsm_kernel(const WsmArgs A) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NF = TT * TT * 2;  // fragment slots per matrix
  constexpr int MA = 8 * TT - 1, LA = TT - 1;
  extern __shared__ double sh[];
  const DModel& md = A.md;
  const int N = md.N, p = md.p, M = md.m, r = md.r;
  const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
  const bool lane_a = g == 7;
  double* Tf = sh;                                         // [NF][32]
  double* Ga = sh + (size_t)NF * 32 * (1 + 2 * warp);      // [NF][32]
  double* Gb = Ga + (size_t)NF * 32;
  const long long w_res = (long long)blockIdx.x * A.wpb + warp, n_res = (long long)gridDim.x * A.wpb;
  double* rqr = A.scratch + (size_t)w_res * 2 * NF * 32;
  double* Ia = rqr + (size_t)NF * 32;

  auto fi = [&](int rt, int ct, int e) { return ((rt * TT + ct) * 2 + e) * 32 + lane; };
  auto row_of = [&](int rt) { return 8 * rt + g; };
  auto col_of = [&](int ct, int e) { return 8 * ct + 2 * q + e; };
  auto pm = [&](int i) { return (int)A.perm[i]; };
  auto t_on = [&](int rt, int ct) { return (A.mask >> (rt * TT + ct)) & 1ull; };

  // T (shared by the batch) once per CTA
  {
    const double* Tb = md.T.block(0, 0, 0);
    for (int tile = warp; tile < TT * TT; tile += A.wpb) {
      const int rt = tile / TT, ct = tile % TT;
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int i = row_of(rt), k = col_of(ct, e);
        Tf[fi(rt, ct, e)] = (i < M && k < M) ? mat_get(Tb, md.T.diag, M, pm(i), pm(k)) : 0.0;
      }
    }
  }
  __syncthreads();

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

  const long long V = A.E / A.B;
  for (long long w_id = w_res; w_id < A.E; w_id += n_res) {
    const long long b = w_id / V, v = w_id % V, e_id = v * A.B + b;
    double* G = Ga;
    double* Gn = Gb;
    double* Gi = Ia;
    // ---- P_star, a, P_inf, R Q R' of this evaluation ----
    {
      const double* ab = md.a.block(b, v, 0);
      const double* Ps = md.P_star.block(b, v, 0);
      const double* Pi = md.P_inf.block(b, v, 0);
      const double* Qb = md.Q.block(b, v, 0);
      const double* Rb = md.R.block(0, 0, 0);
      for (int rt = 0; rt < TT; rt++)
        for (int ct = 0; ct < TT; ct++)
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int i = row_of(rt), k = col_of(ct, e);
            const bool in = i < M && k < M;
            G[fi(rt, ct, e)] = in ? mat_get(Ps, md.P_star.diag, M, pm(i), pm(k)) : ((i == MA && k < M) ? ab[pm(k)] : 0.0);
            Gi[fi(rt, ct, e)] = in ? mat_get(Pi, md.P_inf.diag, M, pm(i), pm(k)) : 0.0;
            double acc = 0.0;
            if (in) {
              const int pi_ = pm(i), pc = pm(k);
              for (int kk = 0; kk < r; kk++) {
                const double rik = mat_get(Rb, md.R.diag, M, pi_, kk);
                if (rik == 0.0) continue;
                double qr;
                if (md.Q.diag) {
                  qr = Qb[kk] * mat_get(Rb, md.R.diag, M, pc, kk);
                } else {
                  qr = 0.0;
                  for (int l = 0; l < r; l++) qr = fma(Qb[kk + (long long)l * r], mat_get(Rb, md.R.diag, M, pc, l), qr);
                }
                acc = fma(rik, qr, acc);
              }
            }
            rqr[fi(rt, ct, e)] = acc;
          }
    }
    __syncwarp();

    // z of the current observation: column values zc[ct][e] and row values zr[rt]; when Z changes from one
    // observation to the next (p > 1 or time-varying Z) the next row is fetched one observation ahead (zcn, zrn)
    double zc[TT][2], zr[TT], zcn[TT][2], zrn[TT];
    int pmc[TT][2], pmr[TT];  // state index behind this lane's columns / rows (-1: padding)
#pragma unroll
    for (int ct = 0; ct < TT; ct++) {
#pragma unroll
      for (int e = 0; e < 2; e++) pmc[ct][e] = col_of(ct, e) < M ? pm(col_of(ct, e)) : -1;
      pmr[ct] = row_of(ct) < M ? pm(row_of(ct)) : -1;
    }
    auto fetch_z = [&](long long idx) {  // flat observation index t*p + j -> zcn, zrn
      const int t = (int)(idx / p), j = (int)(idx % p);
      const double* Zb = md.Z.block(0, 0, md.Z.tv() ? (t < N ? t : N - 1) : 0);
#pragma unroll
      for (int ct = 0; ct < TT; ct++) {
#pragma unroll
        for (int e = 0; e < 2; e++) zcn[ct][e] = pmc[ct][e] >= 0 ? Zb[j + (long long)pmc[ct][e] * p] : 0.0;
        zrn[ct] = pmr[ct] >= 0 ? Zb[j + (long long)pmr[ct] * p] : 0.0;
      }
    };
    auto advance_z = [&]() {
#pragma unroll
      for (int ct = 0; ct < TT; ct++) {
        zc[ct][0] = zcn[ct][0];
        zc[ct][1] = zcn[ct][1];
        zr[ct] = zrn[ct];
      }
    };
    const bool z_fixed = p == 1 && !md.Z.tv();
    fetch_z(0);
    advance_z();

    // m[rt] = (X z)[row] (row MA of G: z'a)
    auto row_dots = [&](const double* X, double (&m)[TT]) {
#pragma unroll
      for (int rt = 0; rt < TT; rt++) {
        double acc = 0.0;
#pragma unroll
        for (int ct = 0; ct < TT; ct++)
#pragma unroll
          for (int e = 0; e < 2; e++) acc = fma(X[fi(rt, ct, e)], zc[ct][e], acc);
        m[rt] = quad_sum(acc);
      }
    };
    auto to_cols = [&](const double (&m)[TT], double (&mc)[TT][2]) {
#pragma unroll
      for (int ct = 0; ct < TT; ct++)
#pragma unroll
        for (int e = 0; e < 2; e++) mc[ct][e] = __shfl_sync(FULL, m[ct], (2 * q + e) * 4);
    };
    auto dot_rows = [&](const double (&m)[TT]) {
      double fs = 0.0;
#pragma unroll
      for (int rt = 0; rt < TT; rt++) fs = fma(zr[rt], m[rt], fs);
      return rows_sum(fs);
    };
    auto abs_small = [&](const double* X) {
      bool ok = true;
      for (int s = 0; s < NF; s++) ok = ok && (fabs(X[s * 32 + lane]) < kEps);
      return __all_sync(FULL, ok) != 0;
    };

    double prod = 1.0, sumq = 0.0;
    int esum = 0, n_terms = 0, non_init = 0, F_eq0 = 0;
    auto acc_log = [&](double F) {
      prod *= F;
      const int hi = __double2hiint(prod);
      const int ex = ((hi >> 20) & 0x7ff) - 1023;
      esum += ex;
      prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
      n_terms++;
    };

    // y: 32 flat indices (t*p + j) at a time, one chunk ahead
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

    // P* <- P* - M K0', a += K0 v   (src/LogLikC.cpp:181-194 / 117-130), in place on G
    auto update_star = [&](const double (&ms)[TT], double F_star, double yv) {
      const double F1 = 1.0 / F_star;
      const double vv = yv - __shfl_sync(FULL, ms[LA], 7 * 4);
      double mc[TT][2];
      to_cols(ms, mc);
#pragma unroll
      for (int rt = 0; rt < TT; rt++) {
        double rc = ms[rt] * F1;
        if (rt == LA) rc = lane_a ? -vv * F1 : rc;
#pragma unroll
        for (int ct = 0; ct < TT; ct++)
#pragma unroll
          for (int e = 0; e < 2; e++) G[fi(rt, ct, e)] = fma(-rc, mc[ct][e], G[fi(rt, ct, e)]);
      }
      acc_log(F_star);
      sumq = fma(vv * vv, F1, sumq);
    };

    // Xn <- [T X ; a'] T' (+ R Q R'), row tile by row tile (X, Xn: fragment matrices of this warp)
    auto time_update = [&](const double* X, double* Xn, bool with_rqr, bool keep_a) {
      for (int rt = 0; rt < TT; rt++) {
        // the accumulator init of the second product comes from global scratch: fetch it first so that its
        // latency hides behind the DMMAs of the first product
        double xv[TT][2];
#pragma unroll
        for (int nt = 0; nt < TT; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) xv[nt][e] = with_rqr ? rqr[fi(rt, nt, e)] : 0.0;
        double yv_[TT][2];
#pragma unroll
        for (int nt = 0; nt < TT; nt++) yv_[nt][0] = yv_[nt][1] = 0.0;
        // Y(rt, nt) = sum_ct T(rt, ct) X(nt, ct)'
        if constexpr (DIAG) {
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const double a = Tf[fi(rt, rt, e)];
#pragma unroll
            for (int nt = 0; nt < TT; nt++) dmma_884(yv_[nt][0], yv_[nt][1], a, X[fi(nt, rt, e)]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 2; e++)
#pragma unroll
            for (int ct = 0; ct < TT; ct++)
              if (t_on(rt, ct)) {
                const double a = Tf[fi(rt, ct, e)];
#pragma unroll
                for (int nt = 0; nt < TT; nt++) dmma_884(yv_[nt][0], yv_[nt][1], a, X[fi(nt, ct, e)]);
              }
        }
        if (keep_a && rt == LA) {
#pragma unroll
          for (int nt = 0; nt < TT; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++)
              if (lane_a) yv_[nt][e] = X[fi(LA, nt, e)];  // row MA of Y <- a'
        }
        // X'(rt, nt) = init + sum_ct Y(rt, ct) T(nt, ct)'
        if constexpr (DIAG) {
#pragma unroll
          for (int e = 0; e < 2; e++)
#pragma unroll
            for (int nt = 0; nt < TT; nt++) dmma_884(xv[nt][0], xv[nt][1], yv_[nt][e], Tf[fi(nt, nt, e)]);
        } else {
#pragma unroll
          for (int e = 0; e < 2; e++)
#pragma unroll
            for (int ct = 0; ct < TT; ct++)
#pragma unroll
              for (int nt = 0; nt < TT; nt++)
                if (t_on(nt, ct)) dmma_884(xv[nt][0], xv[nt][1], yv_[ct][e], Tf[fi(nt, ct, e)]);
        }
#pragma unroll
        for (int nt = 0; nt < TT; nt++)
#pragma unroll
          for (int e = 0; e < 2; e++) Xn[fi(rt, nt, e)] = xv[nt][e];
      }
      __syncwarp();
    };

    bool init = !abs_small(Gi);  // src/LogLikC.cpp:49
    for (int t = 0; t < N; t++) {
      for (int j = 0; j < p; j++) {
        const double yv = next_y();
        if (!z_fixed) {
          const long long idx = (long long)t * p + j;
          if (idx > 0) advance_z();           // the row fetched during the previous observation
          if (idx + 1 < Np) fetch_z(idx + 1);  // in flight while this observation is processed
        }
        if (isnan(yv)) continue;  // :88-90
        double ms[TT];
        row_dots(G, ms);
        const double F_star = dot_rows(ms);
        if (!init) {  // regular step (:164-195)
          non_init++;
          if (F_star < kEps)
            F_eq0++;
          else
            update_star(ms, F_star, yv);
          continue;
        }
        double mi[TT];
        row_dots(Gi, mi);
        const double F_inf = dot_rows(mi);
        if (F_inf < kEps) {
          if (F_star < kEps) continue;  // :112-113
          update_star(ms, F_star, yv);  // :117-131
          continue;
        }
        // :136-153  K0 = M_inf F1, K1 = M* F1 + M_inf F2
        const double F1 = 1.0 / F_inf, F2 = -(F1 * F1) * F_star;
        const double vv = yv - __shfl_sync(FULL, ms[LA], 7 * 4);
        double msc[TT][2], mic[TT][2];
        to_cols(ms, msc);
        to_cols(mi, mic);
#pragma unroll
        for (int rt = 0; rt < TT; rt++) {
          const double rs = (rt == LA && lane_a) ? -vv : ms[rt];
          const double ri = (rt == LA && lane_a) ? 0.0 : mi[rt];
#pragma unroll
          for (int ct = 0; ct < TT; ct++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const double k0 = mic[ct][e] * F1, k1 = msc[ct][e] * F1 + mic[ct][e] * F2;
              G[fi(rt, ct, e)] = fma(-ri, k1, fma(-rs, k0, G[fi(rt, ct, e)]));
              Gi[fi(rt, ct, e)] = fma(-ri, k0, Gi[fi(rt, ct, e)]);
            }
        }
        acc_log(F_inf);  // no v^2 term (:153)
        if (j < p - 1) init = !abs_small(Gi);  // :157-159
      }
      if (t < N - 1) {  // :200-207
        time_update(G, Gn, true, true);
        double* tmp = G;
        G = Gn;
        Gn = tmp;
        if (init) {
          // P_inf lives in global scratch; its product is staged through the shared buffer the update of P* has
          // just vacated (Gn), so that the DMMA operands come from shared memory, not from L2
          for (int s_ = 0; s_ < NF; s_++) Gn[s_ * 32 + lane] = Gi[s_ * 32 + lane];
          time_update(Gn, Gi, false, false);
          init = !abs_small(Gi);
        }
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
    __syncwarp();
  }
}

// ---- host side -----------------------------------------------------------------------------------
bool wsm_eligible(const DModel& md, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  if (md.m > 63) return no("m > 63");
  if (!md.Z.shared() || !md.T.shared() || !md.R.shared()) return no("Z, T, R must be shared by the batch");
  if (md.T.tv() || md.R.tv() || md.Q.tv()) return no("time-varying T, R or Q");
  return true;
}

template <int TT>
static int launch_wsm_tt(WsmArgs& A, cudaStream_t st, DevBuf& scratch) {
  unsigned long long diag = 0;
  for (int t = 0; t < TT; t++) diag |= 1ull << (t * TT + t);
  const bool is_diag = (A.mask & ~diag) == 0;
  constexpr size_t frag = (size_t)TT * TT * 2 * 32 * sizeof(double);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t budget = 227 * 1024 - 1024;
  // warps per CTA and CTAs per SM: as many resident warps as shared memory allows, at most 8 per CTA
  int wpb = 8;
  while (wpb > 1 && frag * (1 + 2 * wpb) > budget) wpb--;
  SS_ARG(frag * (1 + 2 * wpb) <= budget, "loglik_wsm: state dimension too large for shared memory");
  const size_t smem = frag * (1 + 2 * wpb);
  int ctas_per_sm = (int)(budget / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm * wpb > 32) ctas_per_sm = 32 / wpb;
  long long blocks = (A.E + wpb - 1) / wpb;
  if (blocks > (long long)sms * ctas_per_sm) blocks = (long long)sms * ctas_per_sm;
  A.wpb = wpb;
  SS_TRY(scratch.alloc((size_t)blocks * wpb * 2 * frag, st));
  A.scratch = scratch.as<double>();
  if (is_diag) {
    SS_CUDA(cudaFuncSetAttribute(loglik_wsm_kernel<TT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loglik_wsm_kernel<TT, true><<<(unsigned)blocks, wpb * 32, smem, st>>>(A);
  } else {
    SS_CUDA(cudaFuncSetAttribute(loglik_wsm_kernel<TT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    loglik_wsm_kernel<TT, false><<<(unsigned)blocks, wpb * 32, smem, st>>>(A);
  }
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_loglik_wsm(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                      const double* T_host) {
  std::string why;
  SS_ARG(wsm_eligible(md, &why), "warp/shared-memory kernel not eligible: " + why);
  WsmArgs A{};
  A.md = md;
  A.y = y;
  A.B = B;
  A.E = B * V;
  A.out = out;
  const int tt = (md.m + 1 + 7) / 8;
  std::vector<double> Th((size_t)(md.T.diag ? md.m : md.m * md.m));
  if (T_host) {
    std::copy(T_host, T_host + Th.size(), Th.begin());
  } else {
    SS_CUDA(cudaMemcpyAsync(Th.data(), md.T.p, Th.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
  }
  static const bool reorder = [] {
    const char* e = getenv("SSGPU_WSM_REORDER");
    return e ? atoi(e) != 0 : true;
  }();
  const TilePlan pl = plan_tiles(Th, md.T.diag, md.m, tt, reorder);
  A.mask = pl.mask;
  for (int i = 0; i < 64; i++) A.perm[i] = pl.perm[i];
  DevBuf scratch;
  int rc;
  switch (tt) {
    case 1: rc = launch_wsm_tt<1>(A, st, scratch); break;
    case 2: rc = launch_wsm_tt<2>(A, st, scratch); break;
    case 3: rc = launch_wsm_tt<3>(A, st, scratch); break;
    case 4: rc = launch_wsm_tt<4>(A, st, scratch); break;
    case 5: rc = launch_wsm_tt<5>(A, st, scratch); break;
    case 6: rc = launch_wsm_tt<6>(A, st, scratch); break;
    case 7: rc = launch_wsm_tt<7>(A, st, scratch); break;
    default: rc = launch_wsm_tt<8>(A, st, scratch); break;
  }
  SS_TRY(rc);
  count_launch();
  int n_on = 0;
  for (int i = 0; i < tt * tt; i++) n_on += (int)((A.mask >> i) & 1ull);
  static thread_local char name[96];
  snprintf(name, sizeof name, "loglik_wsm_kernel<TT=%d,tiles_on=%d/%d>", tt, n_on, tt * tt);
  set_last_kernel(name);
  return SSGPU_OK;
}

}  // namespace ssgpu
