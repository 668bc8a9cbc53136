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
#include <cstdlib>

#include "common.cuh"

namespace ssgpu {

constexpr int kSimThreads = 128;  // extract kernel; the per-draw kernels pick 32 / 64 / 128 (launchers below)

// per-thread vector v[k] kept in shared memory as sh[k*BT + tid]: conflict-free
template <int BT>
struct ShVec {
  double* p;
  __device__ __forceinline__ double& operator[](int k) const { return p[k * BT]; }
};

__device__ __forceinline__ const double* slice_of(const DMat& M, int t) { return M.block(0, 0, M.tv() ? t : 0); }

// sum over the pattern entries k of row i: Mt[i + k*ld] * x[k], k ascending, starting from +0
template <class Vec>
__device__ __forceinline__ double sp_row_dot(const SpPat& P, int i, const double* Mt, int ld, const Vec& x) {
  double acc = 0.0;
  const int e1 = P.ptr[i + 1];
  for (int e = P.ptr[i]; e < e1; e++) {
    const int k = P.idx[e];
    acc = fma(Mt[i + k * ld], x[k], acc);
  }
  return acc;
}

// sum over the W entries of ELL row i: val[i*W + w] * x[idx[i*W + w]], k ascending, starting from +0
template <class Vec>
__device__ __forceinline__ double ell_row_dot(const Ell& E, const double* vt, int i, const Vec& x) {
  double acc = 0.0;
#pragma unroll 4
  for (int w = 0; w < E.W; w++) acc = fma(vt[w * E.rows + i], x[E.idx[w * E.rows + i]], acc);
  return acc;
}

// ---- SimulateC (SimArgs in common.cuh) ----------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(BT) simulate_kernel(const SimArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  const long long s = (long long)blockIdx.x * BT + tid;
  if (s >= A.S) return;
  const int m = A.m, p = A.p, r = A.r, r_tot = A.r * A.repeat_Q, N = A.N;
  const int a_ld = A.a_ld ? A.a_ld : m, eta_ld = A.eta_ld ? A.eta_ld : r_tot;
  const long long S = A.S;
  ShVec<BT> av{sh + tid}, a2{sh + (size_t)m * BT + tid}, et{sh + (size_t)2 * m * BT + tid};
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
        if (A.eta) A.eta[((long long)t * eta_ld + A.eta_off + rr + r * q) * S + s] = acc;
        if (A.eta_only && A.y) A.y[((long long)t * r_tot + rr + r * q) * S + s] += acc;
      }
    if (A.eta_only) continue;
    const double* Zt = slice_of(A.Z, t);
    const double* Tt = A.Te.vals(t);
    const double* Rt = A.Re.vals(t);
    if (A.a)
      for (int k = 0; k < m; k++) A.a[((long long)t * a_ld + A.a_off + k) * S + s] = av[k];
    for (int j = 0; j < p; j++) {
      const double acc = sp_row_dot(A.Zp, j, Zt, p, av);
      if (A.y) {
        double* yp = A.y + ((long long)t * p + j) * S + s;
        *yp = A.y_acc ? *yp + acc : acc;
      }
    }
    if (t < N - 1) {
      for (int i = 0; i < m; i++) a2[i] = ell_row_dot(A.Te, Tt, i, av) + ell_row_dot(A.Re, Rt, i, et);
      for (int i = 0; i < m; i++) av[i] = a2[i];
    }
  }
}

// ---- FastSmootherC per-draw passes ----------------------------------------------------------------
template <int BT>
__global__ void __launch_bounds__(BT) fs_draw_kernel(const FsArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  const long long s = (long long)blockIdx.x * BT + tid;
  if (s >= A.S) return;
  const int m = A.m, p = A.p, r = A.r, N = A.N;
  const long long S = A.S;
  const size_t vs = (size_t)m * BT;
  // r2 (backward pass only) shares its storage with av (forward passes only)
  ShVec<BT> av{sh + tid}, a2{sh + vs + tid}, rr{sh + 2 * vs + tid}, r1{sh + 3 * vs + tid}, r2{sh + tid},
      et{sh + 4 * vs + tid};

  // forward: innovations v = y - z'a with the shared gains (src/FastSmootherC.cpp:139-251)
  for (int k = 0; k < m; k++) av[k] = A.a0[k];
  for (int t = 0; t < N; t++) {
    const double* Zt = slice_of(A.Z, t);
    const double* Tt = A.Te.vals(t);
    for (int j = 0; j < p; j++) {
      const long long index = (long long)t * p + j;
      if ((A.g.code[index] & 255) == 0) continue;
      const double za = sp_row_dot(A.Zp, j, Zt, p, av);
      const double vv = A.y[index * S + s] - za;
      A.v[index * S + s] = vv;
      const double* k0 = A.g.K0 + index * m;
      for (int k = 0; k < m; k++) av[k] = fma(k0[k], vv, av[k]);
    }
    if (t < N - 1) {
      for (int i = 0; i < m; i++) a2[i] = ell_row_dot(A.Te, Tt, i, av);
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
      const int z0 = A.Zp.ptr[j], z1 = A.Zp.ptr[j + 1];
      if (c == 1) {
        const double sc = c1 - d0;
        for (int e = z0; e < z1; e++) {
          const int k = A.Zp.idx[e];
          rr[k] = fma(Zt[j + k * p], sc, rr[k]);
        }
      } else {
        double d1 = 0.0, d2 = 0.0;
        for (int k = 0; k < m; k++) {
          d1 = fma(k0[k], r1[k], d1);
          d2 = fma(k1[k], rr[k], d2);
        }
        const double s1 = (c1 - d1) - d2, s0 = -d0;
        for (int e = z0; e < z1; e++) {
          const int k = A.Zp.idx[e];
          const double z = Zt[j + k * p];
          r1[k] = fma(z, s1, r1[k]);
          rr[k] = fma(z, s0, rr[k]);
        }
      }
    }
    if (t > 0) {
      // eta_{t-1} = Q_{t-1} R_{t-1}' r_t (src/FastSmootherC.cpp:345-352) is formed here, while r_t is at hand:
      // the forward reconstruction then needs r doubles per step instead of a stored N x m array of r_t
      const double* Qp = A.g.QtR + (A.qtr_tv ? (long long)(t - 1) * r * m : 0);
      for (int kk = 0; kk < r; kk++) A.eta[((long long)(t - 1) * r + kk) * S + s] = sp_row_dot(A.Qp, kk, Qp, r, rr);
      const double* Tp = A.Tce.vals(t - 1);
      const bool both = ((long long)t * p) < A.init_steps;
      for (int jj = 0; jj < m; jj++) {  // T' r: column jj of T
        r2[jj] = ell_row_dot(A.Tce, Tp, jj, rr);
        a2[jj] = both ? ell_row_dot(A.Tce, Tp, jj, r1) : 0.0;
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
      const double* Tp = A.Te.vals(t - 1);
      const double* Rp = A.Re.vals(t - 1);
      for (int kk = 0; kk < r; kk++) et[kk] = A.eta[((long long)(t - 1) * r + kk) * S + s];
      for (int i = 0; i < m; i++) a2[i] = ell_row_dot(A.Te, Tp, i, av) + ell_row_dot(A.Re, Rp, i, et);
      for (int i = 0; i < m; i++) av[i] = a2[i];
    }
    for (int k = 0; k < m; k++) A.a_smooth[((long long)t * m + k) * S + s] = av[k];
  }
  for (int kk = 0; kk < r; kk++) A.eta[((long long)(N - 1) * r + kk) * S + s] = 0.0;  // :81
}


// ---- register-resident variants (m <= 32, r_tot <= 32) -------------------------------------------------
// The state vectors live in registers (MP = 8 / 16 / 32 padded entries, every loop over the state fully unrolled);
// shared memory only holds the gather source of the sparse products: x is parked once per product in xs[k][tid]
// and row i accumulates  sum_w val[w][i] * xs[idx[w][i]]  with w (dynamic, = k ascending) OUTSIDE the unrolled row
// loop, so the MP accumulators advance side by side and every load of a sweep is independent.  The operations and
// their order per result are those of the shared-memory kernels above and of oracle/kalman_c.c: same bits.
constexpr int kRegBT = 32;

template <int MP>
__device__ __forceinline__ void ell_apply(const Ell& E, const double* vt, const ShVec<kRegBT>& xs, int m, double (&out)[MP]) {
#pragma unroll
  for (int i = 0; i < MP; i++) out[i] = 0.0;
  for (int w = 0; w < E.W; w++) {
    const int* id = E.idx + w * E.rows;
    const double* vl = vt + w * E.rows;
#pragma unroll
    for (int i = 0; i < MP; i++)
      if (i < m) out[i] = fma(vl[i], xs[id[i]], out[i]);
  }
}

// ELL table staged in shared memory once per CTA, rows padded to MP with (offset 0, value 0.0): byte offsets of the
// gather source instead of indices, and the values too when the matrix is time-invariant.  The inner loop is then
// two broadcast shared-memory loads at immediate offsets, the gather and the fma -- no bounds predicate, no 64-bit
// address arithmetic, no global load.  (Time-varying matrices keep their values in global memory.)
struct EllS {
  const double* vs;  // shared memory [W][MP], valid when st == 0
  const double* vg;  // global per-slice blocks [W][rows], used when st != 0
  const int* off;    // shared memory [W][MP]: idx * kRegBT * 8
  const unsigned* rm;  // shared memory [W]: bit i set when entry w of row i is a stored one (not padding); st == 0 only
  int W, rows;
  long long st;
};

template <int MP>
__device__ __forceinline__ EllS stage_ell(const Ell& E, double*& dcur, int*& icur, int tid) {
  EllS s;
  s.W = E.W;
  s.rows = E.rows;
  s.st = E.st;
  s.vg = E.val;
  for (int e = tid; e < E.W * MP; e += kRegBT) {
    const int w = e / MP, i = e % MP;
    icur[e] = i < E.rows ? E.idx[w * E.rows + i] * kRegBT * 8 : 0;
    if (E.st == 0) dcur[e] = i < E.rows ? E.val[w * E.rows + i] : 0.0;
  }
  s.off = icur;
  icur += E.W * MP;
  // row masks of the sweeps: a padded entry is (index 0, value 0.0) and fma(0.0, x[0], acc) == acc, so leaving it out
  // changes no bit.  (A stored entry that happens to be an exact 0.0 is left out as well: same argument.)
  unsigned* rm = reinterpret_cast<unsigned*>(icur);
  if (E.st == 0 && tid < E.W) {
    unsigned mk = 0;
    for (int i = 0; i < E.rows && i < 32; i++)
      if (E.val[tid * E.rows + i] != 0.0) mk |= 1u << i;
    rm[tid] = mk;
  }
  s.rm = rm;
  icur += (E.W + 1) & ~1;  // keeps the next table 8-byte aligned
  s.vs = dcur;
  if (E.st == 0) dcur += E.W * MP;
  return s;
}

template <int MP>
__device__ __forceinline__ void ell_apply_s(const EllS& E, int t, const double* x0, int m, double (&out)[MP]) {
#pragma unroll
  for (int i = 0; i < MP; i++) out[i] = 0.0;
  const char* xb = reinterpret_cast<const char*>(x0);
  if (E.st == 0) {
    for (int w = 0; w < E.W; w++) {
      const int* of = E.off + w * MP;
      const double* vl = E.vs + w * MP;
      // groups of four rows without a stored entry in this sweep are branched over (the same decision for every draw):
      // a SARIMA companion matrix has 28 + 2 + 1 entries in its three sweeps of 32
      const unsigned mk = E.rm[w];
#pragma unroll
      for (int i0 = 0; i0 < MP; i0 += 4) {
        if (((mk >> i0) & 0xFu) == 0u) continue;
#pragma unroll
        for (int i = i0; i < i0 + 4 && i < MP; i++)
          out[i] = fma(vl[i], *reinterpret_cast<const double*>(xb + of[i]), out[i]);
      }
    }
  } else {
    const double* vt = E.vg + (long long)t * E.st;
    for (int w = 0; w < E.W; w++) {
      const int* of = E.off + w * MP;
      const double* vl = vt + w * E.rows;
#pragma unroll
      for (int i = 0; i < MP; i++)
        if (i < m) out[i] = fma(vl[i], *reinterpret_cast<const double*>(xb + of[i]), out[i]);
    }
  }
}

template <int MP>
__device__ __forceinline__ void park(const double (&x)[MP], const ShVec<kRegBT>& xs) {
#pragma unroll
  for (int k = 0; k < MP; k++) xs[k] = x[k];
}

template <int MP>
__global__ void __launch_bounds__(kRegBT) simulate_reg_kernel(const SimArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  // every lane stays alive (warp votes below); lanes beyond the last draw shadow draw S-1 and do not store
  const long long s_raw = (long long)blockIdx.x * kRegBT + tid;
  const bool live = s_raw < A.S;
  const long long s = live ? s_raw : A.S - 1;
  const int m = A.m, p = A.p, r = A.r, r_tot = A.r * A.repeat_Q, N = A.N;
  const int a_ld = A.a_ld ? A.a_ld : m, eta_ld = A.eta_ld ? A.eta_ld : r_tot;
  const long long S = A.S;
  ShVec<kRegBT> xs{sh + tid}, es{sh + (size_t)MP * kRegBT + tid};
  EllS Te{}, Re{};
  if (!A.eta_only) {
    double* dcur = sh + (size_t)2 * MP * kRegBT;
    int* icur = reinterpret_cast<int*>(dcur + (A.Te.st ? 0 : A.Te.W * MP) + (A.Re.st ? 0 : A.Re.W * MP));
    Te = stage_ell<MP>(A.Te, dcur, icur, tid);
    Re = stage_ell<MP>(A.Re, dcur, icur, tid);
  }
  __syncwarp();
  double av[MP];
#pragma unroll
  for (int i = 0; i < MP; i++) av[i] = 0.0;
  if (!A.eta_only) {
    if (A.draw_initial) {
      for (int k = 0; k < m; k++) {
        const double dk = A.d0[k + (long long)m * s];
#pragma unroll
        for (int i = 0; i < MP; i++)
          if (i < m) av[i] = fma(A.P_root[i + k * m], dk, av[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < MP; i++)
      if (i < m) av[i] = A.a0[i] + av[i];
  }
  const long long total = (long long)r_tot * S;
  for (int t = 0; t < N; t++) {
    const double* Qr = A.Q_root + (A.nQ > 1 ? (long long)t * r * r : 0);
    const double* dd = A.dt + total * t + (long long)r_tot * s;
    for (int q = 0; q < A.repeat_Q; q++)
      for (int rr = 0; rr < r; rr++) {
        double acc = 0.0;
        for (int l = 0; l < r; l++) acc = fma(Qr[rr + l * r], dd[l + r * q], acc);
        es[rr + r * q] = acc;
        if (A.eta && live) A.eta[((long long)t * eta_ld + A.eta_off + rr + r * q) * S + s] = acc;
        if (A.eta_only && A.y && live) A.y[((long long)t * r_tot + rr + r * q) * S + s] += acc;
      }
    if (A.eta_only) continue;
    const double* Zt = slice_of(A.Z, t);
    if (A.a && live) {
#pragma unroll
      for (int k = 0; k < MP; k++)
        if (k < m) A.a[((long long)t * a_ld + A.a_off + k) * S + s] = av[k];
    }
    for (int j = 0; j < p; j++) {
      double acc = 0.0;
      // non-zero entries of the row of Z (the same for every draw): groups of four zero entries are branched over
      const unsigned zm = __ballot_sync(0xffffffffu, tid < m && Zt[j + (tid < m ? tid : 0) * p] != 0.0);
#pragma unroll
      for (int k0 = 0; k0 < MP; k0 += 4) {
        if (((zm >> k0) & 0xFu) == 0u) continue;
#pragma unroll
        for (int k = k0; k < k0 + 4 && k < MP; k++)
          if (k < m) {
            const double z = Zt[j + k * p];
            if (z != 0.0) acc = fma(z, av[k], acc);
          }
      }
      if (A.y && live) {
        double* yp = A.y + ((long long)t * p + j) * S + s;
        *yp = A.y_acc ? *yp + acc : acc;
      }
    }
    if (t < N - 1) {
      double o2[MP];
      park<MP>(av, xs);
      ell_apply_s<MP>(Te, t, &xs[0], m, av);
      ell_apply_s<MP>(Re, t, &es[0], m, o2);
#pragma unroll
      for (int i = 0; i < MP; i++) av[i] = av[i] + o2[i];
    }
  }
}

template <int MP>
__global__ void __launch_bounds__(kRegBT) fs_draw_reg_kernel(const FsArgs A) {
  extern __shared__ double sh[];
  const int tid = threadIdx.x;
  const int m = A.m, p = A.p, r = A.r, N = A.N;
  const long long S = A.S;
  // every lane stays alive (the rows below are fetched cooperatively); lanes beyond the last draw shadow draw S-1
  // and do not store
  const long long s_raw = (long long)blockIdx.x * kRegBT + tid;
  const bool live = s_raw < S;
  const long long s = live ? s_raw : S - 1;
  // xs: gather source; es: eta_t for R eta; r1: the diffuse-phase vector r^(1) (touched on d of the N steps only)
  // es holds the r disturbances of a step only (r rows, not MP): 20 instead of 28 KB per CTA at config 4, 11 instead of
  // 8 resident CTAs per SM -- what bounds the kernel at 100,000 draws
  ShVec<kRegBT> xs{sh + tid}, r1{sh + (size_t)MP * kRegBT + tid}, es{sh + (size_t)2 * MP * kRegBT + tid};
  double* dcur = sh + ((size_t)2 * MP + r) * kRegBT;
  // the gain row K0 and the row of Z of an observation are the same for all draws: lane k fetches element k one
  // observation ahead and parks it in shared memory (two buffers), every lane then reads the m values as broadcast
  // loads at immediate offsets -- one global load per lane and step instead of 2 m uniform ones in the chain
  double* kb = dcur;          // [2][MP]
  double* zb = dcur + 2 * MP; // [2][MP]
  dcur += 4 * MP;
  int* icur = reinterpret_cast<int*>(dcur + (A.Te.st ? 0 : A.Te.W * MP) + (A.Tce.st ? 0 : A.Tce.W * MP) +
                                     (A.Re.st ? 0 : A.Re.W * MP));
  const EllS Te = stage_ell<MP>(A.Te, dcur, icur, tid);
  const EllS Tce = stage_ell<MP>(A.Tce, dcur, icur, tid);
  const EllS Re = stage_ell<MP>(A.Re, dcur, icur, tid);
  const long long Np = (long long)N * p;
  auto fetch_k = [&](long long index) { return (tid < m && index >= 0 && index < Np) ? A.g.K0[index * m + tid] : 0.0; };
  auto fetch_z = [&](long long index) {
    if (!(tid < m && index >= 0 && index < Np)) return 0.0;
    // p = 1 (the usual case) needs no 64-bit division in the look-ahead fetch
    const int t = p == 1 ? (int)index : (int)(index / p), j = p == 1 ? 0 : (int)(index % p);
    return slice_of(A.Z, t)[j + tid * p];
  };
  double av[MP];

  // forward: innovations v = y - z'a with the shared gains (src/FastSmootherC.cpp:139-251)
#pragma unroll
  for (int k = 0; k < MP; k++) av[k] = k < m ? A.a0[k] : 0.0;
  double y_next = A.y[s];  // observations are fetched one univariate step ahead of their use
  int code_next = A.g.code[0];  // and so is the step's code: the branch on it heads every step's dependence chain
  if (tid < MP) {
    kb[tid] = fetch_k(0);
    zb[tid] = fetch_z(0);
  }
  __syncwarp();
  int cur = 0;
  for (int t = 0; t < N; t++) {
    for (int j = 0; j < p; j++) {
      const long long index = (long long)t * p + j;
      const double y_now = y_next;
      const int code_now = code_next;
      if (index + 1 < Np) {
        y_next = A.y[(index + 1) * S + s];
        code_next = A.g.code[index + 1];
      }
      const double k_n = fetch_k(index + 1), z_n = fetch_z(index + 1);  // in flight during this observation
      const double* kc = kb + cur * MP;
      const double* zc = zb + cur * MP;
      if ((code_now & 255) != 0) {
        // which entries of the row of Z are non-zero: the same for every draw, so whole groups of four states are
        // branched over (Z of a structural or SARIMA model has a handful of ones)
        const unsigned zm = __ballot_sync(0xffffffffu, tid < MP && zc[tid < MP ? tid : 0] != 0.0);
        double za = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < MP; k0 += 4) {
          if (((zm >> k0) & 0xFu) == 0u) continue;
#pragma unroll
          for (int k = k0; k < k0 + 4 && k < MP; k++) {
            const double z = zc[k];
            if (z != 0.0) za = fma(z, av[k], za);
          }
        }
        const double vv = y_now - za;
        if (live) A.v[index * S + s] = vv;
#pragma unroll
        for (int k = 0; k < MP; k++) av[k] = fma(kc[k], vv, av[k]);
      }
      if (tid < MP) {
        kb[(cur ^ 1) * MP + tid] = k_n;
        zb[(cur ^ 1) * MP + tid] = z_n;
      }
      __syncwarp();
      cur ^= 1;
    }
    if (t < N - 1) {
      park<MP>(av, xs);
      ell_apply_s<MP>(Te, t, &xs[0], m, av);
    }
  }

  // backward: r_t (and r^(1) while diffuse), src/FastSmootherC.cpp:259-334
  double rr[MP];
#pragma unroll
  for (int k = 0; k < MP; k++) rr[k] = 0.0;
  for (int k = 0; k < m; k++) r1[k] = 0.0;
  double v_next = A.v[(Np - 1) * S + s];  // innovations, one univariate step ahead (steps without one read junk)
  code_next = A.g.code[Np - 1];
  double f1_next = A.g.F1[Np - 1];
  __syncwarp();
  if (tid < MP) {
    kb[tid] = fetch_k(Np - 1);
    zb[tid] = fetch_z(Np - 1);
  }
  __syncwarp();
  cur = 0;
  for (int t = N - 1; t >= 0; t--) {
    for (int j = p - 1; j >= 0; j--) {
      const long long index = (long long)t * p + j;
      const double v_now = v_next, f1_now = f1_next;
      const int c = code_next & 255;
      if (index > 0) {
        v_next = A.v[(index - 1) * S + s];
        code_next = A.g.code[index - 1];
        f1_next = A.g.F1[index - 1];
      }
      const double k_n = fetch_k(index - 1), z_n = fetch_z(index - 1);
      const double* kc = kb + cur * MP;
      const double* zc = zb + cur * MP;
      if (c != 0) {
        const double c1 = f1_now * v_now;
        double d0 = 0.0;
#pragma unroll
        for (int k = 0; k < MP; k++) d0 = fma(kc[k], rr[k], d0);
        if (c == 1) {
          const double sc = c1 - d0;
          const unsigned zm = __ballot_sync(0xffffffffu, tid < MP && zc[tid < MP ? tid : 0] != 0.0);
#pragma unroll
          for (int k0 = 0; k0 < MP; k0 += 4) {
            if (((zm >> k0) & 0xFu) == 0u) continue;
#pragma unroll
            for (int k = k0; k < k0 + 4 && k < MP; k++) {
              const double z = zc[k];
              if (z != 0.0) rr[k] = fma(z, sc, rr[k]);
            }
          }
        } else {
          const double* k1 = A.g.K1 + index * m;
          double d1 = 0.0, d2 = 0.0;
#pragma unroll
          for (int k = 0; k < MP; k++)
            if (k < m) {
              d1 = fma(kc[k], r1[k], d1);
              d2 = fma(k1[k], rr[k], d2);
            }
          const double s1 = (c1 - d1) - d2, s0 = -d0;
#pragma unroll
          for (int k = 0; k < MP; k++)
            if (k < m) {
              const double z = zc[k];
              if (z != 0.0) {
                r1[k] = fma(z, s1, r1[k]);
                rr[k] = fma(z, s0, rr[k]);
              }
            }
        }
      }
      if (tid < MP) {
        kb[(cur ^ 1) * MP + tid] = k_n;
        zb[(cur ^ 1) * MP + tid] = z_n;
      }
      __syncwarp();
      cur ^= 1;
    }
    if (t > 0) {
      const double* Tp = A.Tce.vals(t - 1);
      const bool both = ((long long)t * p) < A.init_steps;
      park<MP>(rr, xs);
      // eta_{t-1} = Q_{t-1} R_{t-1}' r_t right away (src/FastSmootherC.cpp:345-352): no N x m array of r_t
      const double* Qp = A.g.QtR + (A.qtr_tv ? (long long)(t - 1) * r * m : 0);
      for (int kk = 0; kk < r; kk++) {
        const double ev = sp_row_dot(A.Qp, kk, Qp, r, xs);
        if (live) A.eta[((long long)(t - 1) * r + kk) * S + s] = ev;
      }
      ell_apply_s<MP>(Tce, t - 1, &xs[0], m, rr);  // T' r
      if (both) {  // T' r^(1): rare (the d diffuse steps), kept out of the registers; xs is free again here
        for (int i = 0; i < m; i++) xs[i] = ell_row_dot(A.Tce, Tp, i, r1);
        for (int i = 0; i < m; i++) r1[i] = xs[i];
      }
    }
  }

  // forward reconstruction (src/FastSmootherC.cpp:338-367); rr still holds r_0
  park<MP>(rr, xs);
#pragma unroll
  for (int i = 0; i < MP; i++) {
    av[i] = 0.0;
    if (i < m) {
      double acc = 0.0, acc1 = 0.0;
      for (int k = 0; k < m; k++) {
        acc = fma(A.P_star[i + k * m], xs[k], acc);
        acc1 = fma(A.P_inf[i + k * m], r1[k], acc1);
      }
      av[i] = (A.a0[i] + acc) + acc1;
    }
  }
  // eta_{t-1} was written by this thread in the backward pass; it is fetched one step ahead of its use (up to four
  // disturbances in registers -- the load otherwise sits at the head of every step's dependence chain)
  double en[4] = {0.0, 0.0, 0.0, 0.0};
  const bool pre = r <= 4;
  if (pre && N > 1) {
#pragma unroll
    for (int kk = 0; kk < 4; kk++)
      if (kk < r) en[kk] = A.eta[(long long)kk * S + s];
  }
  for (int t = 0; t < N; t++) {
    if (t > 0) {
      if (pre) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
          if (kk < r) es[kk] = en[kk];
        if (t + 1 < N) {
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            if (kk < r) en[kk] = A.eta[((long long)t * r + kk) * S + s];
        }
      } else {
        for (int kk = 0; kk < r; kk++) es[kk] = A.eta[((long long)(t - 1) * r + kk) * S + s];
      }
      double o2[MP];
      park<MP>(av, xs);
      ell_apply_s<MP>(Te, t - 1, &xs[0], m, av);
      ell_apply_s<MP>(Re, t - 1, &es[0], m, o2);
#pragma unroll
      for (int i = 0; i < MP; i++) av[i] = av[i] + o2[i];
    }
    if (live) {
#pragma unroll
      for (int k = 0; k < MP; k++)
        if (k < m) A.a_smooth[((long long)t * m + k) * S + s] = av[k];
    }
  }
  if (live)
    for (int kk = 0; kk < r; kk++) A.eta[((long long)(N - 1) * r + kk) * S + s] = 0.0;  // :81
}


// ---- FastSmootherC per draw, FOUR lanes per draw (m <= 32) -------------------------------------------------------
// One thread per draw leaves a 12,500-draw shard with 2.6 warps per SM and every warp bound by its own latency
// (about 600 dependent-ish instructions per time step and pass).  Here a draw is spread over LPD = 4 lanes by ROWS:
// lane `sub` owns the state rows [sub*RPL, (sub+1)*RPL) (RPL = MP / 4) of a, r and r^(1), so the sparse products
// T a, T'r, R eta, the gain updates a += K0 v, r += z (.) and the stores are a quarter per lane, and four times as
// many warps are resident.  The dot products z'a, K0'r, K0'r^(1), K1'r stay ONE fma chain in ascending k -- the pinned
// order of oracle/kalman_c.c -- that travels through the four lanes (lane q continues the chain of lane q-1; blocks of
// z that are all zero are skipped, a warp-uniform decision since z is shared by the draws): bit-identical results.
// MEASURED (config 4, one B200): 1.04 ms at 12,500 draws and 7.6 ms at 100,000 against 0.97 / 4.3 ms for one thread
// per draw.  Both kernels are bound by the latency of ONE warp (fs_draw_reg_kernel takes 0.89 ms for 64 draws and
// 0.96 ms for 12,500: 1,041 instructions per warp, time step and pass at 0.24 IPC); spreading a draw over four
// lanes leaves 539 instructions per warp-step for a quarter of the draws (the chain rounds, the block masks and the
// table loads are paid per warp, not per draw) and pushes the LSU pipe to 85 % of its wavefront rate.  Kept as an
// opt-in variant (SSGPU_SIM_LANES=1) because it passes the same bit-exact tests; not the default.
constexpr int kLPD = 4;

template <int MP>
__global__ void __launch_bounds__(kRegBT) fs_draw_lanes_kernel(const FsArgs A) {
  constexpr int RPL = MP / kLPD;     // rows per lane
  constexpr int DPC = kRegBT / kLPD; // draws per CTA
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double sh[];
  const int tid = threadIdx.x, sub = tid & (kLPD - 1), dl = tid / kLPD;
  const int m = A.m, p = A.p, r = A.r, N = A.N;
  const long long S = A.S;
  const long long s_raw = (long long)blockIdx.x * DPC + dl;
  const bool live = s_raw < S;
  const long long s = live ? s_raw : S - 1;  // every lane stays alive; surplus draws shadow the last one
  const int row0 = sub * RPL, gbase = tid & ~(kLPD - 1);
  // gather sources of the sparse products, one column per draw: v[k] at [k*DPC + dl]
  double* xs = sh + dl;
  double* r1 = sh + (size_t)MP * DPC + dl;
  double* es = sh + (size_t)2 * MP * DPC + dl;
  double* dcur = sh + (size_t)3 * MP * DPC;
  double* kb = dcur;           // [2][MP] gain row K0 of the current / next observation
  double* zb = dcur + 2 * MP;  // [2][MP] row of Z
  dcur += 4 * MP;
  int* icur = reinterpret_cast<int*>(dcur + (A.Te.st ? 0 : A.Te.W * MP) + (A.Tce.st ? 0 : A.Tce.W * MP) +
                                     (A.Re.st ? 0 : A.Re.W * MP));
  // ELL tables in shared memory (stage_ell computes byte offsets for a stride of kRegBT doubles: rescale to DPC)
  EllS Te = stage_ell<MP>(A.Te, dcur, icur, tid);
  EllS Tce = stage_ell<MP>(A.Tce, dcur, icur, tid);
  EllS Re = stage_ell<MP>(A.Re, dcur, icur, tid);
  __syncwarp();
  const long long Np = (long long)N * p;
  auto fetch_k = [&](long long index) { return (tid < m && index >= 0 && index < Np) ? A.g.K0[index * m + tid] : 0.0; };
  auto fetch_z = [&](long long index) {
    if (!(tid < m && index >= 0 && index < Np)) return 0.0;
    // p = 1 (the usual case) needs no 64-bit division in the look-ahead fetch
    const int t = p == 1 ? (int)index : (int)(index / p), j = p == 1 ? 0 : (int)(index % p);
    return slice_of(A.Z, t)[j + tid * p];
  };
  // rows row0 .. row0+RPL-1 of  E x  with x parked at xp (one column per draw)
  auto apply = [&](const EllS& E, int t, const double* xp, double (&out)[RPL]) {
#pragma unroll
    for (int j = 0; j < RPL; j++) out[j] = 0.0;
    const char* xb = reinterpret_cast<const char*>(xp);
    const double* vt = E.st ? E.vg + (long long)t * E.st : nullptr;
    for (int w = 0; w < E.W; w++) {
#pragma unroll
      for (int j = 0; j < RPL; j++) {
        const int i = row0 + j;
        const int off = E.off[w * MP + i] / (kRegBT / DPC);  // byte offset of x[idx] in a DPC-wide column layout
        const double val = E.st ? (i < m ? vt[w * E.rows + i] : 0.0) : E.vs[w * MP + i];
        out[j] = fma(val, *reinterpret_cast<const double*>(xb + off), out[j]);
      }
    }
  };
  auto park = [&](const double (&x)[RPL], double* dst) {
#pragma unroll
    for (int j = 0; j < RPL; j++) dst[(row0 + j) * DPC] = x[j];
  };
  // acc + sum_k c[k] x[k], k ascending over all MP rows, x distributed by rows: the chain visits the four lanes in
  // order.  skip_zero: entries with c[k] == 0 are left out (z'a; the oracle skips them too); nzmask: bit q set when
  // block q of c has a non-zero (all-zero blocks are skipped altogether, identical for every draw of the warp)
  auto chain = [&](const double* c, const double (&x)[RPL], bool skip_zero, unsigned nzmask) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < kLPD; q++) {
      if (!((nzmask >> q) & 1u)) continue;
      if (sub == q) {
#pragma unroll
        for (int j = 0; j < RPL; j++) {
          const double cv = c[row0 + j];
          if (!skip_zero || cv != 0.0) acc = fma(cv, x[j], acc);
        }
      }
      acc = __shfl_sync(FULL, acc, gbase + q);
    }
    return acc;
  };
  auto block_mask = [&](const double* c) {  // which quarter of c holds a non-zero (c is shared by all draws)
    bool nz = false;
#pragma unroll
    for (int j = 0; j < RPL; j++) nz = nz || (c[row0 + j] != 0.0);
    const unsigned bal = __ballot_sync(FULL, nz);
    return bal & ((1u << kLPD) - 1u);  // lanes 0..3 = sub 0..3 of draw 0: c is the same for every draw
  };

  double av[RPL];
  // forward: innovations v = y - z'a with the shared gains (src/FastSmootherC.cpp:139-251)
#pragma unroll
  for (int j = 0; j < RPL; j++) av[j] = (row0 + j) < m ? A.a0[row0 + j] : 0.0;
  double y_next = A.y[s];
  if (tid < MP) {
    kb[tid] = fetch_k(0);
    zb[tid] = fetch_z(0);
  }
  __syncwarp();
  int cur = 0;
  for (int t = 0; t < N; t++) {
    for (int j = 0; j < p; j++) {
      const long long index = (long long)t * p + j;
      const double y_now = y_next;
      if (index + 1 < Np) y_next = A.y[(index + 1) * S + s];
      const double k_n = fetch_k(index + 1), z_n = fetch_z(index + 1);
      const double* kc = kb + cur * MP;
      const double* zc = zb + cur * MP;
      if ((A.g.code[index] & 255) != 0) {
        const double za = chain(zc, av, true, block_mask(zc));
        const double vv = y_now - za;
        if (live && sub == 0) A.v[index * S + s] = vv;
#pragma unroll
        for (int q = 0; q < RPL; q++) av[q] = fma(kc[row0 + q], vv, av[q]);
      }
      __syncwarp();
      if (tid < MP) {
        kb[(cur ^ 1) * MP + tid] = k_n;
        zb[(cur ^ 1) * MP + tid] = z_n;
      }
      __syncwarp();
      cur ^= 1;
    }
    if (t < N - 1) {
      park(av, xs);
      __syncwarp();
      apply(Te, t, xs, av);
      __syncwarp();
    }
  }

  // backward: r_t (and r^(1) while diffuse), src/FastSmootherC.cpp:259-334
  double rr[RPL], rq[RPL];  // r and r^(1), own rows
#pragma unroll
  for (int j = 0; j < RPL; j++) rr[j] = rq[j] = 0.0;
  double v_next = A.v[(Np - 1) * S + s];
  __syncwarp();
  if (tid < MP) {
    kb[tid] = fetch_k(Np - 1);
    zb[tid] = fetch_z(Np - 1);
  }
  __syncwarp();
  cur = 0;
  for (int t = N - 1; t >= 0; t--) {
    for (int j = p - 1; j >= 0; j--) {
      const long long index = (long long)t * p + j;
      const double v_now = v_next;
      if (index > 0) v_next = A.v[(index - 1) * S + s];
      const double k_n = fetch_k(index - 1), z_n = fetch_z(index - 1);
      const double* kc = kb + cur * MP;
      const double* zc = zb + cur * MP;
      const int c = A.g.code[index] & 255;
      if (c != 0) {
        const double c1 = A.g.F1[index] * v_now;
        const double d0 = chain(kc, rr, false, (1u << kLPD) - 1u);
        if (c == 1) {
          const double sc = c1 - d0;
#pragma unroll
          for (int q = 0; q < RPL; q++) {
            const double z = zc[row0 + q];
            if (z != 0.0) rr[q] = fma(z, sc, rr[q]);
          }
        } else {
          const double* k1 = A.g.K1 + index * m;
          double k1v[RPL];
#pragma unroll
          for (int q = 0; q < RPL; q++) k1v[q] = (row0 + q) < m ? k1[row0 + q] : 0.0;
          // d1 = K0'r^(1) and d2 = K1'r, each its own ascending chain
          double d1 = 0.0, d2 = 0.0;
#pragma unroll
          for (int qq = 0; qq < kLPD; qq++) {
            if (sub == qq) {
#pragma unroll
              for (int q = 0; q < RPL; q++)
                if (row0 + q < m) {
                  d1 = fma(kc[row0 + q], rq[q], d1);
                  d2 = fma(k1v[q], rr[q], d2);
                }
            }
            d1 = __shfl_sync(FULL, d1, gbase + qq);
            d2 = __shfl_sync(FULL, d2, gbase + qq);
          }
          const double s1 = (c1 - d1) - d2, s0 = -d0;
#pragma unroll
          for (int q = 0; q < RPL; q++) {
            const double z = zc[row0 + q];
            if (z != 0.0) {
              rq[q] = fma(z, s1, rq[q]);
              rr[q] = fma(z, s0, rr[q]);
            }
          }
        }
      }
      __syncwarp();
      if (tid < MP) {
        kb[(cur ^ 1) * MP + tid] = k_n;
        zb[(cur ^ 1) * MP + tid] = z_n;
      }
      __syncwarp();
      cur ^= 1;
    }
    if (t > 0) {
      const bool both = ((long long)t * p) < A.init_steps;
      park(rr, xs);
      if (both) park(rq, r1);
      __syncwarp();
      // eta_{t-1} = Q_{t-1} R_{t-1}' r_t right away (src/FastSmootherC.cpp:345-352): row kk by lane kk % 4
      const double* Qp = A.g.QtR + (A.qtr_tv ? (long long)(t - 1) * r * m : 0);
      for (int kk = sub; kk < r; kk += kLPD) {
        double acc = 0.0;
        const int e1 = A.Qp.ptr[kk + 1];
        for (int e = A.Qp.ptr[kk]; e < e1; e++) {
          const int k = A.Qp.idx[e];
          acc = fma(Qp[kk + k * r], xs[k * DPC], acc);
        }
        if (live) A.eta[((long long)(t - 1) * r + kk) * S + s] = acc;
      }
      apply(Tce, t - 1, xs, rr);  // T' r
      if (both) apply(Tce, t - 1, r1, rq);
      __syncwarp();
    }
  }

  // forward reconstruction (src/FastSmootherC.cpp:338-367); rr holds r_0, rq r^(1)_0
  park(rr, xs);
  park(rq, r1);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < RPL; j++) {
    const int i = row0 + j;
    av[j] = 0.0;
    if (i < m) {
      double acc = 0.0, acc1 = 0.0;
      for (int k = 0; k < m; k++) {
        acc = fma(A.P_star[i + k * m], xs[k * DPC], acc);
        acc1 = fma(A.P_inf[i + k * m], r1[k * DPC], acc1);
      }
      av[j] = (A.a0[i] + acc) + acc1;
    }
  }
  __syncwarp();
  for (int t = 0; t < N; t++) {
    if (t > 0) {
      for (int kk = sub; kk < r; kk += kLPD) es[kk * DPC] = A.eta[((long long)(t - 1) * r + kk) * S + s];
      double o2[RPL];
      park(av, xs);
      __syncwarp();
      apply(Te, t - 1, xs, av);
      apply(Re, t - 1, es, o2);
#pragma unroll
      for (int j = 0; j < RPL; j++) av[j] = av[j] + o2[j];
      __syncwarp();
    }
    if (live) {
#pragma unroll
      for (int j = 0; j < RPL; j++)
        if (row0 + j < m) A.a_smooth[((long long)t * m + row0 + j) * S + s] = av[j];
    }
  }
  if (live)
    for (int kk = sub; kk < r; kk += kLPD) A.eta[((long long)(N - 1) * r + kk) * S + s] = 0.0;  // :81
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

// ---- mean correction of R/SimSmoother.R:634-642 -------------------------------------------------------
// x <- (x - smoothed(simulated data)) + smoothed(observed data), draw-contiguous layout; hat is R's N x K matrix
__global__ void __launch_bounds__(kSimThreads) sim_correct_kernel(double* x, const double* sm, const double* hat, int N,
                                                                   int K, int K_sm, int off_sm, long long S) {
  const long long s = (long long)blockIdx.x * kSimThreads + threadIdx.x;
  const int k = blockIdx.y, t = blockIdx.z;
  if (s >= S) return;
  const long long i = ((long long)t * K + k) * S + s;
  x[i] = (x[i] - sm[((long long)t * K_sm + off_sm + k) * S + s]) + hat[t + (long long)N * k];
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

// ---- non-zero patterns (host) -----------------------------------------------------------------------
size_t PatPack::finish(const std::vector<std::vector<int>>& lists) {
  const size_t off = w_.size();
  int run = 0;
  for (const auto& l : lists) {
    w_.push_back(run);
    run += (int)l.size();
  }
  w_.push_back(run);
  for (const auto& l : lists) w_.insert(w_.end(), l.begin(), l.end());
  return off;
}

size_t PatPack::rows_of(const double* M, int rows, int cols, int n_slices) {
  std::vector<std::vector<int>> lists(rows);
  const size_t sl = (size_t)rows * cols;
  for (int i = 0; i < rows; i++)
    for (int k = 0; k < cols; k++) {
      bool nz = false;
      for (int t = 0; t < n_slices && !nz; t++) nz = M[i + (size_t)k * rows + t * sl] != 0.0;
      if (nz) lists[i].push_back(k);
    }
  return finish(lists);
}

size_t PatPack::cols_of(const double* M, int rows, int cols, int n_slices) {
  std::vector<std::vector<int>> lists(cols);
  const size_t sl = (size_t)rows * cols;
  for (int j = 0; j < cols; j++)
    for (int k = 0; k < rows; k++) {
      bool nz = false;
      for (int t = 0; t < n_slices && !nz; t++) nz = M[k + (size_t)j * rows + t * sl] != 0.0;
      if (nz) lists[j].push_back(k);
    }
  return finish(lists);
}

// (Q_t R_t')[kk][k] = sum_l Q_t[kk][l] R_t[k][l] can only be non-zero where some l has both factors non-zero
size_t PatPack::qtr_of(const double* Q, int nQ, const double* R, int nR, int m, int r) {
  std::vector<std::vector<int>> lists(r);
  const int n = nQ > nR ? nQ : nR;
  for (int kk = 0; kk < r; kk++)
    for (int k = 0; k < m; k++) {
      bool nz = false;
      for (int t = 0; t < n && !nz; t++) {
        const double* Qt = Q + (size_t)(nQ > 1 ? t : 0) * r * r;
        const double* Rt = R + (size_t)(nR > 1 ? t : 0) * m * r;
        for (int l = 0; l < r && !nz; l++) nz = Qt[kk + (size_t)l * r] != 0.0 && Rt[k + (size_t)l * m] != 0.0;
      }
      if (nz) lists[kk].push_back(k);
    }
  return finish(lists);
}

EllPack::Ref EllPack::add(const double* M, int rows, int cols, int n_slices, bool by_cols) {
  const int nr = by_cols ? cols : rows, nc = by_cols ? rows : cols;
  const size_t sl = (size_t)rows * cols;
  auto at = [&](int i, int k, int t) { return by_cols ? M[k + (size_t)i * rows + t * sl] : M[i + (size_t)k * rows + t * sl]; };
  std::vector<std::vector<int>> lists(nr);
  int W = 1;
  for (int i = 0; i < nr; i++) {
    for (int k = 0; k < nc; k++) {
      bool nz = false;
      for (int t = 0; t < n_slices && !nz; t++) nz = at(i, k, t) != 0.0;
      if (nz) lists[i].push_back(k);
    }
    if ((int)lists[i].size() > W) W = (int)lists[i].size();
  }
  Ref r{i_.size(), v_.size(), W, nr, n_slices > 1 ? (long long)nr * W : 0};
  for (int w = 0; w < W; w++)
    for (int i = 0; i < nr; i++) i_.push_back(w < (int)lists[i].size() ? lists[i][w] : 0);
  for (int t = 0; t < n_slices; t++)
    for (int w = 0; w < W; w++)
      for (int i = 0; i < nr; i++) v_.push_back(w < (int)lists[i].size() ? at(i, lists[i][w], t) : 0.0);
  while (v_.size() % 32) v_.push_back(0.0);
  while (i_.size() % 32) i_.push_back(0);
  return r;
}

size_t PatPack::dense(int rows, int cols) {
  std::vector<std::vector<int>> lists(rows);
  for (int i = 0; i < rows; i++)
    for (int k = 0; k < cols; k++) lists[i].push_back(k);
  return finish(lists);
}

// ---- launchers --------------------------------------------------------------------------------------
static unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

// Threads per CTA of the per-draw kernels: small CTAs keep every SM busy when a GPU's share of the draws is small
// (nsim = 1e5 over 8 GPUs is 12,500 draws = 98 CTAs of 128 threads on 148 SMs, but 391 CTAs of 32)
// Also, the per-thread state vectors live in shared memory ((4m + r) doubles for the smoother): 32-thread CTAs pack
// an SM's 227 KB more tightly than 128-thread ones once m is large.
static int pick_bt(long long S, size_t doubles_per_thread) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  if (doubles_per_thread * 8 > 256) return 32;
  if (S >= (long long)sms * 128 * 4) return 128;
  if (S >= (long long)sms * 64 * 4) return 64;
  return 32;
}

// CUDA events around the last launch of each per-draw kernel (ssgpu_last_sim_kernel_ms: bench.py's live
// per-kernel time on the launching stream)
struct KernelTimer {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool used = false;
  void start(cudaStream_t st) {
    if (!e0) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
    }
    cudaEventRecord(e0, st);
  }
  void stop(cudaStream_t st) {
    cudaEventRecord(e1, st);
    used = true;
  }
  double ms() {
    if (!used) return -1.0;
    float f = 0.f;
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&f, e0, e1);
    return f;
  }
};
static KernelTimer g_t_sim, g_t_fs, g_t_ext;

void last_sim_kernel_ms(double* sim, double* fs, double* ext) {
  if (sim) *sim = g_t_sim.ms();
  if (fs) *fs = g_t_fs.ms();
  if (ext) *ext = g_t_ext.ms();
}

template <int BT>
static int launch_simulate_bt(const SimArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * BT * (2 * (size_t)A.m + (size_t)A.r * A.repeat_Q);
  SS_ARG(smem <= 227 * 1024, "SimulateC: state too large for shared memory");
  SS_CUDA(cudaFuncSetAttribute(simulate_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g_t_sim.start(st);
  simulate_kernel<BT><<<nblk(A.S, BT), BT, smem, st>>>(A);
  g_t_sim.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

template <int MP>
static int launch_simulate_reg(const SimArgs& A, cudaStream_t st) {
  auto tab = [](const Ell& E) {
    return (size_t)E.W * MP * ((E.st ? 0 : sizeof(double)) + sizeof(int)) + sizeof(unsigned) * ((E.W + 1) & ~1);
  };
  const size_t smem = sizeof(double) * kRegBT * 2 * MP + (A.eta_only ? 0 : tab(A.Te) + tab(A.Re)) + 16;
  SS_CUDA(cudaFuncSetAttribute(simulate_reg_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g_t_sim.start(st);
  simulate_reg_kernel<MP><<<nblk(A.S, kRegBT), kRegBT, smem, st>>>(A);
  g_t_sim.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

static bool reg_variant_enabled() {
  static const bool on = [] {
    const char* e = getenv("SSGPU_SIM_REG");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

int launch_simulate(const SimArgs& A, cudaStream_t st) {
  const int r_tot = A.r * A.repeat_Q;
  if (reg_variant_enabled() && A.m <= 32 && r_tot <= 32) {
    const int need = A.m > r_tot ? A.m : r_tot;
    return need <= 8 ? launch_simulate_reg<8>(A, st) : need <= 16 ? launch_simulate_reg<16>(A, st)
                                                                   : launch_simulate_reg<32>(A, st);
  }
  const int bt = pick_bt(A.S, 2 * (size_t)A.m + (size_t)A.r * A.repeat_Q);
  return bt == 128 ? launch_simulate_bt<128>(A, st) : bt == 64 ? launch_simulate_bt<64>(A, st) : launch_simulate_bt<32>(A, st);
}

template <int BT>
static int launch_fs_bt(const FsArgs& A, cudaStream_t st) {
  const size_t smem = sizeof(double) * BT * (4 * (size_t)A.m + (size_t)A.r);
  SS_ARG(smem <= 227 * 1024, "FastSmootherC: state too large for shared memory");
  SS_CUDA(cudaFuncSetAttribute(fs_draw_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g_t_fs.start(st);
  fs_draw_kernel<BT><<<nblk(A.S, BT), BT, smem, st>>>(A);
  g_t_fs.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

template <int MP>
static int launch_fs_reg(const FsArgs& A, cudaStream_t st) {
  auto tab = [](const Ell& E) {
    return (size_t)E.W * MP * ((E.st ? 0 : sizeof(double)) + sizeof(int)) + sizeof(unsigned) * ((E.W + 1) & ~1);
  };
  const size_t smem = sizeof(double) * (kRegBT * (2 * MP + (size_t)A.r) + 4 * MP) + tab(A.Te) + tab(A.Tce) + tab(A.Re) + 16;
  SS_ARG(smem <= 200 * 1024, "FastSmootherC: sparse tables too large for shared memory");
  SS_CUDA(cudaFuncSetAttribute(fs_draw_reg_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g_t_fs.start(st);
  fs_draw_reg_kernel<MP><<<nblk(A.S, kRegBT), kRegBT, smem, st>>>(A);
  g_t_fs.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

template <int MP>
static int launch_fs_lanes(const FsArgs& A, cudaStream_t st) {
  constexpr int DPC = kRegBT / kLPD;
  auto tab = [](const Ell& E) {
    return (size_t)E.W * MP * ((E.st ? 0 : sizeof(double)) + sizeof(int)) + sizeof(unsigned) * ((E.W + 1) & ~1);
  };
  const size_t smem = sizeof(double) * (DPC * 3 * MP + 4 * MP) + tab(A.Te) + tab(A.Tce) + tab(A.Re) + 16;
  SS_ARG(smem <= 200 * 1024, "FastSmootherC: sparse tables too large for shared memory");
  SS_CUDA(cudaFuncSetAttribute(fs_draw_lanes_kernel<MP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  g_t_fs.start(st);
  fs_draw_lanes_kernel<MP><<<nblk(A.S, DPC), kRegBT, smem, st>>>(A);
  g_t_fs.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

static bool lanes_variant_enabled() {
  static const bool on = [] {
    const char* e = getenv("SSGPU_SIM_LANES");
    return e ? atoi(e) != 0 : false;  // measured slower than one thread per draw (below): opt-in only
  }();
  return on;
}

int launch_fs_draws(const FsArgs& A, cudaStream_t st) {
  if (reg_variant_enabled() && lanes_variant_enabled() && A.m <= 32 && A.r <= 32) {
    const int need = A.m > A.r ? A.m : A.r;
    return need <= 8 ? launch_fs_lanes<8>(A, st) : need <= 16 ? launch_fs_lanes<16>(A, st) : launch_fs_lanes<32>(A, st);
  }
  if (reg_variant_enabled() && A.m <= 32 && A.r <= 32) {
    const int need = A.m > A.r ? A.m : A.r;
    return need <= 8 ? launch_fs_reg<8>(A, st) : need <= 16 ? launch_fs_reg<16>(A, st) : launch_fs_reg<32>(A, st);
  }
  const int bt = pick_bt(A.S, 4 * (size_t)A.m + (size_t)A.r);
  return bt == 128 ? launch_fs_bt<128>(A, st) : bt == 64 ? launch_fs_bt<64>(A, st) : launch_fs_bt<32>(A, st);
}

int launch_extract(const double* a, int a_native, const DMat& Z, int m, int p, int S, int N, double* out,
                   cudaStream_t st) {
  SS_ARG(N <= 65535, "ExtractComponentC: N too large");
  g_t_ext.start(st);
  extract_kernel<<<dim3(nblk(S, kSimThreads), N), kSimThreads, 0, st>>>(a, a_native, Z, m, p, S, N, out);
  g_t_ext.stop(st);
  count_launch();
  SS_CUDA(cudaGetLastError());
  return SSGPU_OK;
}

int launch_sim_correct(double* x, const double* sm, const double* hat, int N, int K, int K_sm, int off_sm, long long S,
                       cudaStream_t st) {
  SS_ARG(K <= 65535 && N <= 65535, "SimSmoother: dimension too large");
  sim_correct_kernel<<<dim3(nblk(S, kSimThreads), K, N), kSimThreads, 0, st>>>(x, sm, hat, N, K, K_sm, off_sm, S);
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
