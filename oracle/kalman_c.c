/* CPU oracle in plain C (TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT).
 *
 * Allocation-free restatement of statespacer's Kalman hot path with a PINNED arithmetic order, so
 * that the CUDA kernels can be checked bit for bit where the contract asks for it (simulation
 * smoother draws) and to 1e-9 elsewhere.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; statespacer_b200/ never does.
 *
 * Pinned order (compiled with -ffp-contract=off; the CUDA side with -fmad=false):
 *   every inner product / matrix product element is  acc = 0; for k ascending: acc = fma(x_k, y_k, acc);
 *   every other operation is the single IEEE add / mul / div written in the source.
 *
 * It follows the reference's recursions but applies the measurement update in the rank-1 form
 *   P L0' = P - (P z) K0'      P_inf L1' + P* L0' = P* - M* K0' - M_inf K1'      (src/LogLikC.cpp:129,151-152,193)
 * and computes the FastSmootherC gain sequence once instead of once per draw (it does not depend on
 * y, src/FastSmootherC.cpp:150-251).  Agreement with the unmodified reference sources (oracle/_ref)
 * and with the line-by-line NumPy restatement (oracle/kalman_np.py) is checked in
 * tests/test_oracle_c.py to 1e-10; the reference's golden vectors pin it in tests/test_oracle_kat.py.
 *
 * Reference files: src/LogLikC.cpp:68-214, src/FastSmootherC.cpp:88-368, src/SimulateC.cpp:71-131,
 * src/ExtractComponentC.cpp:16-37.  All arrays column-major with the R shapes.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EPS 1e-7
#define LOGC (-0.91893853320467274178032973640562)

static inline const double* slice(const double* X, int n_slices, long sz, int t) {
  return n_slices > 1 ? X + sz * t : X;
}

/* W (r x m) = Q R' ; RQR (m x m) = R W */
static void build_rqr(const double* Q, const double* R, int m, int r, double* W, double* RQR) {
  for (int j = 0; j < m; j++)
    for (int k = 0; k < r; k++) {
      double acc = 0.0;
      for (int l = 0; l < r; l++) acc = fma(Q[k + l * r], R[j + l * m], acc);
      W[k + j * r] = acc;
    }
  if (RQR)
    for (int j = 0; j < m; j++)
      for (int i = 0; i < m; i++) {
        double acc = 0.0;
        for (int k = 0; k < r; k++) acc = fma(R[i + k * m], W[k + j * r], acc);
        RQR[i + j * m] = acc;
      }
}

/* P <- T P T' (+ add), W scratch m x m */
static void sandwich(double* P, const double* T, const double* add, int m, double* W) {
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) {
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(T[i + k * m], P[k + j * m], acc);
      W[i + j * m] = acc;
    }
  for (int j = 0; j < m; j++)
    for (int i = 0; i < m; i++) {
      double acc = 0.0;
      for (int k = 0; k < m; k++) acc = fma(W[i + k * m], T[j + k * m], acc);
      P[i + j * m] = add ? acc + add[i + j * m] : acc;
    }
}

static int all_small(const double* P, int n) {
  for (int i = 0; i < n; i++)
    if (!(fabs(P[i]) < EPS)) return 0;
  return 1;
}

/* Forward pass.  y != NULL: loglikelihood mode (regime by the P_inf test, NaN = missing), result in
 * *loglik.  y == NULL: gain mode of FastSmootherC (regime by index < init_steps, no missing values):
 * code[Np] (0 none, 1 F_star update, 2 F_inf update), F1[Np], K0/K1 [Np*m], QtR [nQtR*r*m].
 * work: 4*m*m + r*m + 6*m doubles. */
static void forward(const double* y, int N, int p, int m, int r, const double* a0, const double* P_inf0,
                    const double* P_star0, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                    const double* Q, int nQ, int init_steps, double* loglik, int* code, double* F1o, double* K0o,
                    double* K1o, double* QtRo, double* work) {
  const int mm = m * m;
  double* Ps = work;
  double* Pi = Ps + mm;
  double* W = Pi + mm;
  double* RQR = W + mm;
  double* QtR = RQR + mm; /* r x m */
  double* a = QtR + r * m;
  double* a2 = a + m;
  double* Ms = a2 + m;
  double* Mi = Ms + m;
  double* k0 = Mi + m;
  double* k1 = k0 + m;
  memcpy(Ps, P_star0, sizeof(double) * mm);
  memcpy(Pi, P_inf0, sizeof(double) * mm);
  if (a0) memcpy(a, a0, sizeof(double) * m);
  const int gains = (y == NULL);
  int init = gains ? 1 : !all_small(Pi, mm);
  const int rq_tv = nQ > 1 || nR > 1;
  double ll = 0.0;
  int non_init = 0, F_eq0 = 0;
  build_rqr(Q, R, m, r, QtR, RQR);
  if (gains) memcpy(QtRo, QtR, sizeof(double) * r * m);
  for (int t = 0; t < N; t++) {
    const double* Zt = slice(Z, nZ, (long)p * m, t);
    const double* Tt = slice(T, nT, (long)mm, t);
    if (t > 0 && rq_tv) {
      build_rqr(slice(Q, nQ, (long)r * r, t), slice(R, nR, (long)m * r, t), m, r, QtR, RQR);
      if (gains) memcpy(QtRo + (long)t * r * m, QtR, sizeof(double) * r * m);
    }
    for (int j = 0; j < p; j++) {
      const int index = t * p + j;
      double yv = 0.0;
      if (gains) {
        init = index < init_steps;
        code[index] = 0;
        F1o[index] = 0.0;
      } else {
        yv = y[t + (long)N * j];
        if (isnan(yv)) continue;
      }
      for (int i = 0; i < m; i++) {
        double as = 0.0, ai = 0.0;
        for (int k = 0; k < m; k++) {
          const double z = Zt[j + k * p];
          as = fma(Ps[i + k * m], z, as);
          if (init) ai = fma(Pi[i + k * m], z, ai);
        }
        Ms[i] = as;
        Mi[i] = ai;
      }
      double F_star = 0.0, F_inf = 0.0, za = 0.0;
      for (int k = 0; k < m; k++) {
        const double z = Zt[j + k * p];
        F_star = fma(z, Ms[k], F_star);
        F_inf = fma(z, Mi[k], F_inf);
        if (!gains) za = fma(z, a[k], za);
      }
      int c = 0;
      if (init) {
        if (F_inf < EPS) {
          if (!(F_star < EPS)) c = 1;
        } else
          c = 2;
      } else {
        non_init++;
        if (F_star < EPS)
          F_eq0++;
        else
          c = 1;
      }
      if (c == 0) {
        if (gains)
          for (int i = 0; i < m; i++) K0o[(long)index * m + i] = K1o[(long)index * m + i] = 0.0;
        continue;
      }
      double F1, F2 = 0.0, vv = yv - za;
      if (c == 1) {
        F1 = 1.0 / F_star;
        for (int i = 0; i < m; i++) {
          k0[i] = Ms[i] * F1;
          k1[i] = 0.0;
        }
        ll += LOGC - (log(F_star) + vv * vv * F1) / 2;
      } else {
        F1 = 1.0 / F_inf;
        F2 = -(F1 * F1) * F_star;
        for (int i = 0; i < m; i++) {
          k0[i] = Mi[i] * F1;
          k1[i] = Ms[i] * F1 + Mi[i] * F2;
        }
        ll += LOGC - log(F_inf) / 2;
      }
      if (gains) {
        code[index] = c;
        F1o[index] = F1;
        for (int i = 0; i < m; i++) {
          K0o[(long)index * m + i] = k0[i];
          K1o[(long)index * m + i] = k1[i];
        }
      }
      for (int cc = 0; cc < m; cc++)
        for (int i = 0; i < m; i++) {
          const int q = i + cc * m;
          if (c == 1) {
            Ps[q] = Ps[q] - Ms[i] * k0[cc];
          } else {
            Ps[q] = (Ps[q] - Ms[i] * k0[cc]) - Mi[i] * k1[cc];
            Pi[q] = Pi[q] - Mi[i] * k0[cc];
          }
        }
      if (!gains) {
        for (int i = 0; i < m; i++) a[i] = fma(k0[i], vv, a[i]);
        if (c == 2 && j < p - 1) init = !all_small(Pi, mm);
      }
    }
    if (t < N - 1) {
      if (!gains) {
        for (int i = 0; i < m; i++) {
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(Tt[i + k * m], a[k], acc);
          a2[i] = acc;
        }
        memcpy(a, a2, sizeof(double) * m);
      }
      sandwich(Ps, Tt, RQR, m, W);
      if (gains ? ((t * p + p - 1) < init_steps) : init) {
        sandwich(Pi, Tt, NULL, m, W);
        if (!gains) init = !all_small(Pi, mm);
      }
    }
  }
  if (loglik) *loglik = (non_init == F_eq0) ? NAN : ll / N;
}

static size_t forward_work(int m, int r) { return (size_t)4 * m * m + (size_t)r * m + 6 * (size_t)m; }

int ssor_loglik(const double* y, const int* y_isna, int N, int p, int m, int r, const double* a, const double* P_inf,
                const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                const double* Q, int nQ, double* out) {
  double* work = (double*)malloc(sizeof(double) * (forward_work(m, r) + (size_t)N * p));
  if (!work) return 1;
  double* yy = work + forward_work(m, r);
  for (long i = 0; i < (long)N * p; i++) yy[i] = (y_isna && y_isna[i]) ? NAN : y[i];
  forward(yy, N, p, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, -1, out, NULL, NULL, NULL, NULL, NULL, work);
  free(work);
  return 0;
}

int ssor_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Batch of independent evaluations (p = 1, time-invariant shared a, P_inf, Z, T, R; per-series diagonal-or-dense
 * P_star and Q): y [B][N] series-major with NaN = missing.  The hand-tuned CPU port timed beside the GPU. */
int ssor_loglik_batch(const double* y, long B, int N, int m, int r, const double* a, const double* P_inf,
                      const double* P_star, const double* Z, const double* T, const double* R, const double* Q,
                      int per_series, int threads, double* out) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  int fail = 0;
#pragma omp parallel
  {
    double* work = (double*)malloc(sizeof(double) * forward_work(m, r));
    if (!work) fail = 1;
#pragma omp for schedule(dynamic, 16)
    for (long b = 0; b < B; b++) {
      if (!work) continue;
      const double* Ps = per_series ? P_star + b * (long)m * m : P_star;
      const double* Qb = per_series ? Q + b * (long)r * r : Q;
      forward(y + b * (long)N, N, 1, m, r, a, P_inf, Ps, Z, 1, T, 1, R, 1, Qb, 1, -1, out + b, NULL, NULL, NULL, NULL,
              NULL, work);
    }
    free(work);
  }
  return fail;
}

/* FastSmootherC gain sequence (y-independent part of src/FastSmootherC.cpp:103-252).
 * code[Np], F1[Np], K0/K1 [Np*m], QtR [nQtR*r*m] with nQtR = (nQ>1||nR>1) ? N : 1. */
int ssor_gains(int N, int p, int m, int r, const double* P_inf, const double* P_star, const double* Z, int nZ,
               const double* T, int nT, const double* R, int nR, const double* Q, int nQ, int init_steps, int* code,
               double* F1, double* K0, double* K1, double* QtR) {
  double* work = (double*)malloc(sizeof(double) * forward_work(m, r));
  if (!work) return 1;
  forward(NULL, N, p, m, r, NULL, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, init_steps, NULL, code, F1, K0, K1, QtR,
          work);
  free(work);
  return 0;
}

/* FastSmootherC (src/FastSmootherC.cpp:23-375): y N x p x nsim -> a_smooth N x m x nsim, a_t m x nsim x N
 * (if transposed_state), eta N x r x nsim (row N-1 = 0). */
int ssor_fast_smoother(const double* y, int N, int p, int nsim, int m, int r, const double* a, const double* P_inf,
                       const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                       const double* Q, int nQ, int init_steps, int transposed_state, double* a_smooth, double* a_t,
                       double* eta, int threads) {
  const long Np = (long)N * p;
  const int qtr_tv = nQ > 1 || nR > 1;
  int* code = (int*)malloc(sizeof(int) * Np);
  double* F1 = (double*)malloc(sizeof(double) * Np);
  double* K0 = (double*)malloc(sizeof(double) * Np * m);
  double* K1 = (double*)malloc(sizeof(double) * Np * m);
  double* QtR = (double*)malloc(sizeof(double) * (qtr_tv ? N : 1) * r * m);
  if (!code || !F1 || !K0 || !K1 || !QtR) return 1;
  ssor_gains(N, p, m, r, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, init_steps, code, F1, K0, K1, QtR);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    double* v = (double*)malloc(sizeof(double) * Np);
    double* rv = (double*)malloc(sizeof(double) * (long)N * m);
    double* av = (double*)malloc(sizeof(double) * 6 * m + sizeof(double) * r);
    double* a2 = av + m;
    double* rr = a2 + m;
    double* r1 = rr + m;
    double* r2 = r1 + m;
    double* r12 = r2 + m;
    double* et = r12 + m;
#pragma omp for schedule(static)
    for (int s = 0; s < nsim; s++) {
      const double* ys = y + (long)s * Np;
      memcpy(av, a, sizeof(double) * m);
      /* forward: innovations (:139-244) */
      for (int t = 0; t < N; t++) {
        const double* Zt = slice(Z, nZ, (long)p * m, t);
        const double* Tt = slice(T, nT, (long)m * m, t);
        for (int j = 0; j < p; j++) {
          const long index = (long)t * p + j;
          v[index] = 0.0;
          if (code[index] == 0) continue;
          double za = 0.0;
          for (int k = 0; k < m; k++) za = fma(Zt[j + k * p], av[k], za);
          const double vv = ys[t + (long)N * j] - za;
          v[index] = vv;
          const double* k0 = K0 + index * m;
          for (int k = 0; k < m; k++) av[k] = fma(k0[k], vv, av[k]);
        }
        if (t < N - 1) {
          for (int i = 0; i < m; i++) {
            double acc = 0.0;
            for (int k = 0; k < m; k++) acc = fma(Tt[i + k * m], av[k], acc);
            a2[i] = acc;
          }
          memcpy(av, a2, sizeof(double) * m);
        }
      }
      /* backward: r recursion (:259-334) */
      for (int k = 0; k < m; k++) rr[k] = r1[k] = 0.0;
      for (int t = N - 1; t >= 0; t--) {
        const double* Zt = slice(Z, nZ, (long)p * m, t);
        for (int j = p - 1; j >= 0; j--) {
          const long index = (long)t * p + j;
          const int c = code[index];
          if (c == 0) continue;
          const double* k0 = K0 + index * m;
          const double* k1 = K1 + index * m;
          const double c1 = F1[index] * v[index];
          double d0 = 0.0;
          for (int k = 0; k < m; k++) d0 = fma(k0[k], rr[k], d0);
          if (c == 1) {
            const double sc = c1 - d0;
            for (int k = 0; k < m; k++) rr[k] = fma(Zt[j + k * p], sc, rr[k]);
          } else {
            double d1 = 0.0, d2 = 0.0;
            for (int k = 0; k < m; k++) {
              d1 = fma(k0[k], r1[k], d1);
              d2 = fma(k1[k], rr[k], d2);
            }
            const double s1 = (c1 - d1) - d2, s0 = -d0;
            for (int k = 0; k < m; k++) {
              const double z = Zt[j + k * p];
              r1[k] = fma(z, s1, r1[k]);
              rr[k] = fma(z, s0, rr[k]);
            }
          }
        }
        memcpy(rv + (long)t * m, rr, sizeof(double) * m);
        if (t > 0) {
          const double* Tp = slice(T, nT, (long)m * m, t - 1);
          const int both = ((long)t * p) < init_steps;
          for (int jj = 0; jj < m; jj++) {
            double acc = 0.0, acc1 = 0.0;
            for (int k = 0; k < m; k++) {
              acc = fma(Tp[k + jj * m], rr[k], acc);
              if (both) acc1 = fma(Tp[k + jj * m], r1[k], acc1);
            }
            r2[jj] = acc;
            r12[jj] = acc1;
          }
          memcpy(rr, r2, sizeof(double) * m);
          if (both) memcpy(r1, r12, sizeof(double) * m);
        }
      }
      /* forward reconstruction (:338-367) */
      for (int i = 0; i < m; i++) {
        double acc = 0.0, acc1 = 0.0;
        for (int k = 0; k < m; k++) {
          acc = fma(P_star[i + k * m], rv[k], acc);
          acc1 = fma(P_inf[i + k * m], r1[k], acc1);
        }
        av[i] = (a[i] + acc) + acc1;
      }
      for (int t = 0; t < N; t++) {
        if (t > 0) {
          const double* Tp = slice(T, nT, (long)m * m, t - 1);
          const double* Rp = slice(R, nR, (long)m * r, t - 1);
          const double* Qp = QtR + (qtr_tv ? (long)(t - 1) * r * m : 0);
          for (int kk = 0; kk < r; kk++) {
            double acc = 0.0;
            for (int k = 0; k < m; k++) acc = fma(Qp[kk + k * r], rv[(long)t * m + k], acc);
            et[kk] = acc;
            eta[(t - 1) + (long)N * (kk + (long)r * s)] = acc;
          }
          for (int i = 0; i < m; i++) {
            double acc = 0.0, acc1 = 0.0;
            for (int k = 0; k < m; k++) acc = fma(Tp[i + k * m], av[k], acc);
            for (int k = 0; k < r; k++) acc1 = fma(Rp[i + k * m], et[k], acc1);
            a2[i] = acc + acc1;
          }
          memcpy(av, a2, sizeof(double) * m);
        }
        for (int k = 0; k < m; k++) {
          a_smooth[t + (long)N * (k + (long)m * s)] = av[k];
          if (transposed_state && a_t) a_t[k + (long)m * (s + (long)nsim * t)] = av[k];
        }
      }
      for (int kk = 0; kk < r; kk++) eta[(N - 1) + (long)N * (kk + (long)r * s)] = 0.0;
    }
    free(v);
    free(rv);
    free(av);
  }
  free(code);
  free(F1);
  free(K0);
  free(K1);
  free(QtR);
  return 0;
}

/* SimulateC (src/SimulateC.cpp:26-139) with injected normal variates and precomputed symmetric roots
 * Q_root (r x r x nQ) and P_root (mP x mP, only read if draw_initial).  With eta_only only eta is written. */
int ssor_simulate(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, const double* a, const double* Z, int nZ,
                  const double* T, int nT, const double* R, int nR, int nQ, int draw_initial, int eta_only,
                  int transposed_state, const double* draws, long n_draws, const double* Q_root, const double* P_root,
                  double* y, double* a_out, double* a_t, double* eta, int threads) {
  const int r_tot = r * repeat_Q;
  const long total = (long)r_tot * nsim;
  const long need = (draw_initial ? (long)m * nsim : 0) + total * N;
  if (need != n_draws) return 3;
  if (!eta_only && rR != r_tot) return 1;
  const double* d0 = draws;
  const double* dt = draws + (draw_initial ? (long)m * nsim : 0);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel
  {
    double* av = (double*)malloc(sizeof(double) * (2 * m + r_tot + 1));
    double* a2 = av + m;
    double* et = a2 + m;
#pragma omp for schedule(static)
    for (int s = 0; s < nsim; s++) {
      if (!eta_only) {
        for (int i = 0; i < m; i++) {
          if (draw_initial) {
            double acc = 0.0;
            for (int k = 0; k < m; k++) acc = fma(P_root[i + k * m], d0[k + (long)m * s], acc);
            av[i] = a[i] + acc;
          } else
            av[i] = a[i];
        }
      }
      for (int t = 0; t < N; t++) {
        const double* Qr = slice(Q_root, nQ, (long)r * r, t);
        const double* dd = dt + total * t + (long)r_tot * s;
        for (int q = 0; q < repeat_Q; q++)
          for (int rr = 0; rr < r; rr++) {
            double acc = 0.0;
            for (int l = 0; l < r; l++) acc = fma(Qr[rr + l * r], dd[l + r * q], acc);
            et[rr + r * q] = acc;
            if (eta) eta[t + (long)N * (rr + r * q + (long)r_tot * s)] = acc;
          }
        if (eta_only) continue;
        const double* Zt = slice(Z, nZ, (long)p * m, t);
        const double* Tt = slice(T, nT, (long)m * m, t);
        const double* Rt = slice(R, nR, (long)m * rR, t);
        for (int k = 0; k < m; k++) {
          if (a_out) a_out[t + (long)N * (k + (long)m * s)] = av[k];
          if (transposed_state && a_t) a_t[k + (long)m * (s + (long)nsim * t)] = av[k];
        }
        for (int j = 0; j < p; j++) {
          double acc = 0.0;
          for (int k = 0; k < m; k++) acc = fma(Zt[j + k * p], av[k], acc);
          if (y) y[t + (long)N * (j + (long)p * s)] = acc;
        }
        if (t < N - 1) {
          for (int i = 0; i < m; i++) {
            double acc = 0.0, acc1 = 0.0;
            for (int k = 0; k < m; k++) acc = fma(Tt[i + k * m], av[k], acc);
            for (int k = 0; k < rR; k++) acc1 = fma(Rt[i + k * m], et[k], acc1);
            a2[i] = acc + acc1;
          }
          memcpy(av, a2, sizeof(double) * m);
        }
      }
    }
    free(av);
  }
  return 0;
}

/* ExtractComponentC (src/ExtractComponentC.cpp:12-40): a m x nsim x N, Z p x m x nZ -> N x p x nsim */
int ssor_extract(const double* a, int m, int nsim, int N, const double* Z, int p, int nZ, double* out) {
  for (int s = 0; s < nsim; s++)
    for (int t = 0; t < N; t++) {
      const double* Zt = slice(Z, nZ, (long)p * m, t);
      const double* at = a + (long)m * (s + (long)nsim * t);
      for (int j = 0; j < p; j++) {
        double acc = 0.0;
        for (int k = 0; k < m; k++) acc = fma(Zt[j + k * p], at[k], acc);
        out[t + (long)N * (j + (long)p * s)] = acc;
      }
    }
  return 0;
}
