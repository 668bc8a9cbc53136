// Batched loglikelihood, "column-per-lane" register kernel -- the throughput path behind
// ssgpu_loglik_batch[_dev] for models whose Z, T, R are shared by the whole batch and
// time-invariant (per-evaluation a, P_inf, P_star, Q are allowed), m <= 15.
//
// Restates the recursion of src/LogLikC.cpp:68-214 (univariate treatment, exact diffuse
// initialisation, NaN = missing).  Mapping:
//   * one evaluation (series b, parameter vector v) per group of G = pow2 >= m+1 lanes;
//   * lane c < m keeps column c of P_star (and of P_inf while the filter is diffuse) in registers,
//     lane m keeps the state vector a as an extra column: with the columns side by side
//         M = [P a]' z           gives  P z  and  z'a   in one sweep,
//         [P a] -= M g'          is the rank-1 covariance update and  a += K v  at once
//                                (g_c = M_c / F for c < m, g_m = -v / F),
//         [P a] <- T [P a]       is the first half of T P T' and the whole of  a <- T a;
//   * T, Z, R live in __constant__ memory: every DFMA of the two T-products takes its matrix operand
//     from a uniform register filled by an LDCU issued next to it (no vector register, no LSU traffic);
//   * T P T' = T (T P)' needs the transpose of T P: lanes exchange it through a padded,
//     conflict-free shared-memory tile; P z is all-gathered through shared memory as well;
//   * sum(log F) is carried as (mantissa product, exponent sum): one log() per evaluation.
// The measurement update is applied as the rank-1 forms P - M K0' (src/LogLikC.cpp:129,193) and
// P* - M* K0' - M_inf K1', P_inf - M_inf K0' (:151-152): equal to the reference's GEMM forms in
// real arithmetic, O(eps) apart in FP64.
#include <type_traits>

#include <mutex>

#include "common.cuh"

namespace ssgpu {

constexpr int kColsMaxM = 15;
constexpr int kColsThreads = 128;
#ifndef SS_COLS_MINB
#define SS_COLS_MINB 3
#endif
constexpr int kColsMinBlocks = SS_COLS_MINB;
// __constant__ layout (doubles): T column-major, even pitch [k*ldt(M) + i] | Z [j*M + k] | R [k*M + i]
constexpr int kOffT = 0, kOffZ = 1024, kOffR = 3072, kConstDoubles = 4096;
constexpr int kColsMaxP = 2048 / 16, kColsMaxR = 32;

__constant__ __align__(16) double cMat[kConstDoubles];
__host__ __device__ constexpr int ldt(int m) { return m + (m & 1); }

struct ColsArgs {
  DModel md;
  const double* y;  // [(t*p + j) * B + b]
  long long B;
  long long E;  // B * V
  double* out;
};

// repack the shared matrices for the constant bank (one tiny launch per batch call)
__global__ void cols_pack_kernel(DModel md, double* dst) {
  const int m = md.m, p = md.p, r = md.r;
  const double* Tb = md.T.block(0, 0, 0);
  const double* Zb = md.Z.block(0, 0, 0);
  const double* Rb = md.R.block(0, 0, 0);
  for (int q = threadIdx.x; q < kConstDoubles; q += blockDim.x) dst[q] = 0.0;
  __syncthreads();
  for (int q = threadIdx.x; q < m * m; q += blockDim.x) {
    const int i = q % m, k = q / m;
    dst[kOffT + k * ldt(m) + i] = mat_get(Tb, md.T.diag, m, i, k);
  }
  for (int q = threadIdx.x; q < p * m; q += blockDim.x) {
    const int j = q / m, k = q % m;
    dst[kOffZ + q] = Zb[j + (long long)k * p];
  }
  for (int q = threadIdx.x; q < r * m; q += blockDim.x) {
    const int k = q / m, i = q % m;
    dst[kOffR + q] = md.R.diag ? (i == k ? Rb[i] : 0.0) : Rb[i + (long long)k * m];
  }
}

// y += T x with x read from shared memory and T from the constant bank.  The k loop is a REAL loop
// (unrolled by 2 only): ptxas then software-pipelines one LDCU per DFMA straight into the DFMA's
// uniform-register operand.  Fully unrolled, its scheduler hoists all m^2 constant loads to the top of
// the block, runs out of uniform registers and bounces T through vector registers (LDC + R2UR), which
// is what capped the first version of this kernel at 24 % of the FP64 peak (profiles/).
// Each y[i] is summed over k in ascending order.
template <int M>
__device__ __forceinline__ void mul_T(const double* xs, double (&y)[M]) {
#pragma unroll 2
  for (int k = 0; k < M; k++) {
    const double xk = xs[k];
#pragma unroll
    for (int i = 0; i < M; i++) y[i] = fma(cMat[kOffT + k * ldt(M) + i], xk, y[i]);
  }
}

template <int M, int G, bool P1, int MINB>
__global__ void __launch_bounds__(kColsThreads, MINB) loglik_cols_kernel(const ColsArgs A) {
  constexpr int LD = G + 1;  // padded pitch: conflict-free by rows and by columns
  // transpose tile (M rows) + one private row per lane (G rows) + 2 double-buffered gather vectors
  constexpr int GROUP_SMEM = (M + G) * LD + 4 * G;
  constexpr int GPB = kColsThreads / G;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double smem[];

  const DModel& md = A.md;
  const int N = md.N, p = P1 ? 1 : md.p, r = md.r;
  const int tid = threadIdx.x;
  const int gi = tid / G, c = tid % G;
  // The whole kernel is warp-convergent (no early exit, no group-dependent branch: every per-group
  // decision of the recursion is a select) so that ptxas keeps T on the uniform datapath (LDCU -> UR
  // operand of the DFMA).  Tail groups recompute the last evaluation and skip the store.
  const long long e_raw = (long long)blockIdx.x * GPB + gi;
  const bool live = e_raw < A.E;
  const long long w_id = live ? e_raw : A.E - 1;
  // groups are numbered series-major (w = b*V + v): the V parameter vectors of a series run side by side
  // and share its y through L1/L2; results keep the public order e = v*B + b
  const long long V = A.E / A.B, b = w_id / V, v = w_id % V;
  const long long e = v * A.B + b;
  const unsigned gshift = (tid % 32) / G * G, gbits = (G == 32) ? FULL : ((1u << G) - 1u);
  double* S = smem + gi * GROUP_SMEM;
  double* X = S + M * LD + c * LD;  // this lane's private row
  double* Mv = S + (M + G) * LD;    // [2 parities][2 (star, inf)][G]
  const bool is_col = c < M, is_a = c == M;
  const int cc = is_col ? c : M - 1;  // row this lane reads back from the transpose tile

  // ---- the evaluation's own a / P_star column and its column of R Q R' ------------------------
  double col[M], rqr[M];
  {
    const double* ab = md.a.block(b, v, 0);
    const double* Ps = md.P_star.block(b, v, 0);
#pragma unroll
    for (int k = 0; k < M; k++) {
      col[k] = is_col ? mat_get(Ps, md.P_star.diag, M, k, c) : (is_a ? ab[k] : 0.0);
      rqr[k] = 0.0;
    }
    // rqr[i] = sum_k R[i,k] (Q R')[k,c]   (src/LogLikC.cpp:202, hoisted: Q and R are time-invariant here)
    const double* Qb = md.Q.block(b, v, 0);
    for (int k = 0; k < r; k++) {
      double qr;
      if (md.Q.diag) {
        qr = Qb[k] * cMat[kOffR + k * M + cc];
      } else {
        qr = 0.0;
        for (int l = 0; l < r; l++) qr = fma(Qb[k + (long long)l * r], cMat[kOffR + l * M + cc], qr);
      }
      qr = is_col ? qr : 0.0;
#pragma unroll
      for (int i = 0; i < M; i++) rqr[i] = fma(cMat[kOffR + k * M + i], qr, rqr[i]);
    }
  }

  // loglik = n_terms*const - (sum log F + sum v^2/F)/2, log F accumulated as mantissa * 2^esum
  double prod = 1.0, sumq = 0.0;
  int esum = 0, n_terms = 0, non_init = 0, F_eq0 = 0, par = 0;
  auto acc_log = [&](double F, bool counted) {  // F == 1 (skipped step) is a no-op
    prod *= F;
    const int hi = __double2hiint(prod);
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    esum += ex;
    prod = __hiloint2double(hi - (ex << 20), __double2loint(prod));
    n_terms += counted ? 1 : 0;
  };

  // y is read G flat indices (q = t*p + j) at a time, one per lane, one chunk ahead
  const long long Np = (long long)N * p;
  auto load_y = [&](long long qq) { return qq < Np ? A.y[qq * A.B + b] : 0.0; };
  double ybuf = 0.0, ynext = load_y(c);
  long long q = 0;
  auto next_y = [&]() {
    const int ql = (int)(q & (G - 1));
    if (ql == 0) {
      ybuf = ynext;
      ynext = load_y(q + G + c);
    }
    q++;
    return __shfl_sync(FULL, ybuf, ql, G);
  };

  // M = [P a]' z into the gather buffer; returns this lane's entry
  auto sweep = [&](const double (&x)[M], const double* z, double* dst) {
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) acc = fma(x[k], z[k], acc);
    dst[c] = acc;
    return acc;
  };

  // x <- T x T' (+ R Q R'), columns exchanged through the shared tile.  For the [P a] block lane m
  // keeps T a (the first product) instead of the second one.  Both products run through ONE instance
  // of the inner loop (rep = 0, 1): a second textual copy in the same loop nest is not given the uniform
  // datapath by ptxas and falls back to LDC into vector registers.
  auto sandwich = [&](double (&x)[M], auto with_rqr, auto keep_first) {
    double y[M];
#pragma unroll
    for (int k = 0; k < M; k++) X[k] = x[k];
#pragma unroll 1
    for (int rep = 0; rep < 2; rep++) {
      const double* xs = rep ? S + cc * LD : X;  // rep 1: row c of T X
#pragma unroll
      for (int i = 0; i < M; i++) y[i] = (decltype(with_rqr)::value && rep) ? rqr[i] : 0.0;
      mul_T<M>(xs, y);
      if (rep == 0) {  // lanes < m: column c of T X;  lane m: T a
#pragma unroll
        for (int i = 0; i < M; i++) S[i * LD + c] = y[i];
        __syncwarp();
      }
    }
    // lanes < m: column c of T X T' [+ R Q R'];  lane m reads its own first product (T a) back
#pragma unroll
    for (int k = 0; k < M; k++) {
      const double first = decltype(keep_first)::value ? S[k * LD + c] : 0.0;
      x[k] = is_col ? y[k] : first;
    }
    __syncwarp();
  };
  using yes = std::true_type;
  using no = std::false_type;

  int t = 0;
#ifndef SS_EXP_NO_DIFFUSE
  {
    // ---- exact diffuse phase (src/LogLikC.cpp:98-159): P_inf columns live only in this block ----
    double pin[M];
    const double* Pi = md.P_inf.block(b, v, 0);
#pragma unroll
    for (int k = 0; k < M; k++) pin[k] = is_col ? mat_get(Pi, md.P_inf.diag, M, k, c) : 0.0;
    auto pinf_small = [&]() {  // all(|P_inf| < 1e-7) over the group, evaluated by the whole warp
      bool ok = true;
#pragma unroll
      for (int k = 0; k < M; k++) ok = ok && (fabs(pin[k]) < kEps);
      const unsigned bal = __ballot_sync(FULL, ok);
      return ((bal >> gshift) & gbits) == gbits;
    };
    bool init = !pinf_small();  // :49

    // Every group of the warp stays in this general loop until the last one has left its diffuse
    // phase, so that t is warp-uniform when the regular loop starts.  The three outcomes of a step
    // (:112 nothing, :117-131 update with F_star, :136-153 update with F_inf) and the regular update of
    // a group that is already past its diffuse phase are one code path with selected coefficients.
    for (; t < N && __any_sync(FULL, init); t++) {
      for (int j = 0; j < p; j++) {
        const double yv = next_y();
        const bool obs = !isnan(yv);  // :88-90
        const double* z = cMat + kOffZ + (P1 ? 0 : j * M);
        double* mv = Mv + par * 2 * G;
        par ^= 1;
        const double ms = sweep(col, z, mv);
        const double mi = sweep(pin, z, mv + G);
        __syncwarp();
        double F_star = 0.0, F_inf = 0.0;
#pragma unroll
        for (int k = 0; k < M; k++) {
          F_star = fma(z[k], mv[k], F_star);
          F_inf = fma(z[k], mv[G + k], F_inf);
        }
        F_inf = init ? F_inf : 0.0;
        const double za = mv[M];
        const bool s_tiny = F_star < kEps;
        const bool c2 = obs && init && !(F_inf < kEps);
        const bool c1 = obs && !c2 && !s_tiny;
        non_init += (obs && !init) ? 1 : 0;
        F_eq0 += (obs && !init && s_tiny) ? 1 : 0;
        const double Fs = c2 ? F_inf : (c1 ? F_star : 1.0);
        const double F1 = __drcp_rn(Fs);
        const double F2 = c2 ? -(F1 * F1) * F_star : 0.0;
        const double vv = (c1 || c2) ? yv - za : 0.0;
        const double gv = -(F1 * vv);
        // coefficient of M_star / M_inf in this lane's column update, and of M_inf in its P_inf update
        const double k0 = (c2 ? mi : ms) * F1;  // K0[c]
        const double cs = is_col ? ((c1 || c2) ? k0 : 0.0) : ((is_a && c1) ? gv : 0.0);
        const double ci = is_col ? (c2 ? ms * F1 + mi * F2 : 0.0) : ((is_a && c2) ? gv : 0.0);
        const double cp = (is_col && c2) ? k0 : 0.0;
#pragma unroll
        for (int k = 0; k < M; k++) {
          const double Msk = mv[k], Mik = c2 ? mv[G + k] : 0.0;
          // lanes < m: P*[k,c] - M*[k] K0[c] - M_inf[k] K1[c];  lane m: a[k] + K0[k] v
          col[k] = fma(-Mik, ci, fma(-Msk, cs, col[k]));
          pin[k] = fma(-Mik, cp, pin[k]);
        }
        acc_log(Fs, c1 || c2);
        sumq = fma(c1 ? vv * vv : 0.0, F1, sumq);  // no v^2 term in the F_inf branch (:153)
        if (j < p - 1) {                            // :157-159
          const bool small = pinf_small();
          init = c2 ? !small : init;
        }
      }
      if (t < N - 1) {  // :200-207
        sandwich(col, yes{}, yes{});
        sandwich(pin, no{}, no{});
        const bool small = pinf_small();
        init = init && !small;
#pragma unroll
        for (int k = 0; k < M; k++) pin[k] = init ? pin[k] : 0.0;
      }
    }
  }
#endif

  // ---- regular phase: the hot loop, branch-free (missing values and F ~ 0 are handled by selects:
  // a skipped update is g = 0, F = 1); the last time point (no time update, :200) is peeled ---------
  auto hot_obs = [&](int j) {
    const double yv = next_y();
    const bool obs = !isnan(yv);  // :88-90
    const double* z = cMat + kOffZ + (P1 ? 0 : j * M);
    double* mv = Mv + par * 2 * G;
    par ^= 1;
    const double ms = sweep(col, z, mv);
    __syncwarp();
    double Ms[M];
    double F_star = 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) {
      Ms[k] = mv[k];
      F_star = fma(z[k], Ms[k], F_star);
    }
    const double za = mv[M];
    const bool tiny = F_star < kEps;  // :173-176
    const bool upd = obs && !tiny;
    non_init += obs ? 1 : 0;
    F_eq0 += (obs && tiny) ? 1 : 0;
    const double Fs = upd ? F_star : 1.0;
    const double F1 = __drcp_rn(Fs), vv = upd ? yv - za : 0.0;
    const double g = upd ? (is_col ? ms * F1 : (is_a ? -(F1 * vv) : 0.0)) : 0.0;
#pragma unroll
    for (int k = 0; k < M; k++) col[k] = fma(-Ms[k], g, col[k]);
    acc_log(Fs, upd);
    sumq = fma(vv * vv, F1, sumq);
  };
  const int t_hot = __reduce_max_sync(FULL, t);  // == t in every lane; tells ptxas it is uniform
  for (int tt = t_hot; tt < N - 1; tt++) {
    for (int j = 0; j < p; j++) hot_obs(j);
    sandwich(col, yes{}, yes{});
  }
  if (t_hot < N)
    for (int j = 0; j < p; j++) hot_obs(j);

  if (c == 0) {
    double res;
    if (non_init == F_eq0) {
      res = nan("");  // src/LogLikC.cpp:211-213
    } else {
      const double sumlog = (double)esum * 0.69314718055994530942 + log(prod);
      res = ((double)n_terms * kLogConst - 0.5 * (sumlog + sumq)) / (double)N;
    }
    if (live) A.out[e] = res;
  }
}

// ---- host side -----------------------------------------------------------------------------------
static int group_lanes(int m) {
  int g = 2;
  while (g < m + 1) g *= 2;
  return g;
}

bool cols_eligible(const DModel& md, std::string* why) {
  auto no = [&](const char* s) {
    if (why) *why = s;
    return false;
  };
  if (md.m > kColsMaxM) return no("m > 15");
  if (md.r > kColsMaxR) return no("r > 32");
  if (md.p > kColsMaxP) return no("p too large for the constant bank");
  if (!md.Z.shared() || !md.T.shared() || !md.R.shared()) return no("Z, T, R must be shared by the batch");
  if (md.Z.tv() || md.T.tv() || md.R.tv() || md.Q.tv()) return no("time-varying system matrices");
  return true;
}

template <int M, int G, bool P1>
static int launch_cols_inst(const ColsArgs& A, cudaStream_t st) {
  constexpr int GPB = kColsThreads / G;
  constexpr size_t smem = sizeof(double) * GPB * ((M + G) * (G + 1) + 4 * G);
  const long long blocks = (A.E + GPB - 1) / GPB;
  SS_ARG(blocks < (1LL << 31), "column kernel: batch too large for one launch");
  loglik_cols_kernel<M, G, P1, kColsMinBlocks><<<(unsigned)blocks, kColsThreads, smem, st>>>(A);
  return SSGPU_OK;
}

template <int M>
static int launch_cols_m(const ColsArgs& A, cudaStream_t st) {
  constexpr int G = M + 1 <= 2 ? 2 : M + 1 <= 4 ? 4 : M + 1 <= 8 ? 8 : 16;
  if (A.md.p == 1) return launch_cols_inst<M, G, true>(A, st);
  return launch_cols_inst<M, G, false>(A, st);
}

// T, Z, R travel through ONE __constant__ block per device.  Two launches on different streams of one device would
// race on it, so they are chained: a launch waits (on its own stream) for the event its predecessor recorded behind
// its kernel before it overwrites the block.  Host side: one mutex per device slot around the enqueue.
static double* g_packs[16] = {nullptr};
static cudaEvent_t g_cols_done[16] = {nullptr};
static std::mutex g_cols_mu[16];

int launch_loglik_cols(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st) {
  std::string why;
  SS_ARG(cols_eligible(md, &why), "column kernel not eligible: " + why);
  const int slot = dev_slot();
  std::lock_guard<std::mutex> lk(g_cols_mu[slot]);
  double*& g_pack = g_packs[slot];
  if (!g_pack) SS_CUDA(cudaMalloc(&g_pack, sizeof(double) * kConstDoubles));
  if (!g_cols_done[slot])
    SS_CUDA(cudaEventCreateWithFlags(&g_cols_done[slot], cudaEventDisableTiming));
  else
    SS_CUDA(cudaStreamWaitEvent(st, g_cols_done[slot], 0));  // the previous launch has finished reading the block
  cols_pack_kernel<<<1, 256, 0, st>>>(md, g_pack);
  SS_CUDA(cudaGetLastError());
  SS_CUDA(cudaMemcpyToSymbolAsync(cMat, g_pack, sizeof(double) * kConstDoubles, 0, cudaMemcpyDeviceToDevice, st));
  ColsArgs A{};
  A.md = md;
  A.y = y;
  A.B = B;
  A.E = B * V;
  A.out = out;
  (void)group_lanes;
  int rc = SSGPU_E_ARG;
  switch (md.m) {
#define SS_CASE(MM) \
  case MM:          \
    rc = launch_cols_m<MM>(A, st); \
    break;
#ifdef SS_COLS_ONLY_M  // development builds: one instantiation
    SS_CASE(SS_COLS_ONLY_M)
#else
    SS_CASE(1) SS_CASE(2) SS_CASE(3) SS_CASE(4) SS_CASE(5) SS_CASE(6) SS_CASE(7) SS_CASE(8)
    SS_CASE(9) SS_CASE(10) SS_CASE(11) SS_CASE(12) SS_CASE(13) SS_CASE(14) SS_CASE(15)
#endif
#undef SS_CASE
    default:
      set_error("column kernel: unsupported m");
      return SSGPU_E_ARG;
  }
  SS_TRY(rc);
  SS_CUDA(cudaGetLastError());
  SS_CUDA(cudaEventRecord(g_cols_done[slot], st));  // the next launch on this device overwrites cMat behind it
  count_launch(2);
  set_last_kernel("loglik_cols_kernel");
  return SSGPU_OK;
}

}  // namespace ssgpu
