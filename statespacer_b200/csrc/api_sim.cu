// C ABI of libssgpu, simulation smoother part: ssgpu_FastSmootherC, ssgpu_SimulateC,
// ssgpu_ExtractComponentC (include/ssgpu.h sections 3-5) -- host staging around the kernels of sim.cu.
#include <cstring>

#include <functional>

#include "common.cuh"

namespace ssgpu {

static int upload(DevBuf& d, const double* src, size_t n, cudaStream_t st) {
  SS_TRY(d.alloc(n * sizeof(double), st));
  if (n) SS_TRY(copy_h2d(d.get(), src, n * sizeof(double), st));
  return SSGPU_OK;
}

static DMat dmat3(const double* dev, int rows, int cols, int n_slices) {
  DMat d{};
  d.p = dev;
  d.rows = rows;
  d.cols = cols;
  d.st = n_slices > 1 ? (long long)rows * cols : 0;
  return d;
}

// U diag(sqrt(|lambda|)) U' of a symmetric matrix by cyclic Jacobi rotations: the matrix that
// src/SimulateC.cpp:73-74,84-85 obtains from arma::svd (U diag(sqrt(s)) U', s = |lambda|).
static void sym_root_host(const double* A, int n, double* out) {
  const size_t N = (size_t)n;
  std::vector<double> a(A, A + N * N), v(N * N, 0.0);
  for (size_t i = 0; i < N; i++) v[i + i * N] = 1.0;
  for (size_t i = 0; i < N; i++)
    for (size_t j = i + 1; j < N; j++) a[i + j * N] = a[j + i * N] = 0.5 * (a[i + j * N] + a[j + i * N]);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, diag = 0.0;
    for (size_t i = 0; i < N; i++)
      for (size_t j = 0; j < N; j++) (i == j ? diag : off) += a[i + j * N] * a[i + j * N];
    if (off == 0.0 || off <= 1e-32 * diag) break;
    for (size_t p = 0; p + 1 < N; p++)
      for (size_t q = p + 1; q < N; q++) {
        const double apq = a[p + q * N];
        if (apq == 0.0) continue;
        const double theta = (a[q + q * N] - a[p + p * N]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
        for (size_t k = 0; k < N; k++) {
          const double akp = a[k + p * N], akq = a[k + q * N];
          a[k + p * N] = c * akp - sn * akq;
          a[k + q * N] = sn * akp + c * akq;
        }
        for (size_t k = 0; k < N; k++) {
          const double apk = a[p + k * N], aqk = a[q + k * N];
          a[p + k * N] = c * apk - sn * aqk;
          a[q + k * N] = sn * apk + c * aqk;
        }
        for (size_t k = 0; k < N; k++) {
          const double vkp = v[k + p * N], vkq = v[k + q * N];
          v[k + p * N] = c * vkp - sn * vkq;
          v[k + q * N] = sn * vkp + c * vkq;
        }
      }
  }
  for (size_t i = 0; i < N; i++)
    for (size_t j = 0; j < N; j++) {
      double acc = 0.0;
      for (size_t k = 0; k < N; k++) acc += v[i + k * N] * std::sqrt(std::fabs(a[k + k * N])) * v[j + k * N];
      out[i + j * N] = acc;
    }
}

// pads every block to 32 doubles so device sub-buffers stay 256-byte aligned
struct HostPack {
  std::vector<double> buf;
  size_t add(const double* src, size_t n) {
    const size_t off = buf.size();
    buf.resize(off + (n + 31) / 32 * 32, 0.0);
    if (n) std::memcpy(buf.data() + off, src, n * sizeof(double));
    return off;
  }
};

}  // namespace ssgpu

using namespace ssgpu;

extern "C" int ssgpu_sym_root(const double* A, int n, double* out) {
  SS_ARG(A && out && n >= 1, "sym_root: bad arguments");
  sym_root_host(A, n, out);
  return SSGPU_OK;
}

namespace ssgpu {

static int check_fs(const void* y, int N, int p, int nsim, int m, int r, const double* a, const double* P_inf,
                    const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                    const double* Q, int nQ) {
  SS_ARG(y && a && P_inf && P_star && Z && T && R && Q, "null pointer argument");
  SS_ARG(N >= 1 && p >= 1 && m >= 1 && r >= 1 && nsim >= 1, "N, p, m, r, nsim must be positive");
  SS_ARG(m <= SSGPU_MAX_M, "state dimension m exceeds SSGPU_MAX_M");
  SS_ARG((nZ == 1 || nZ == N) && (nT == 1 || nT == N) && (nR == 1 || nR == N) && (nQ == 1 || nQ == N),
         "system matrices must have 1 or N slices");
  return SSGPU_OK;
}

// FastSmootherC on device-resident draws in the draw-contiguous layout: yN [(t*p+j)*S+s] -> aN [(t*m+k)*S+s],
// eN [(t*r+kk)*S+s].  System matrices are host pointers.  Everything is enqueued on st; the temporaries are
// stream-ordered allocations released behind the kernels.
static int fast_smoother_core(const double* yN, int N, int p, int nsim, int m, int r, const double* a,
                              const double* P_inf, const double* P_star, const double* Z, int nZ, const double* T, int nT,
                              const double* R, int nR, const double* Q, int nQ, int initialisation_steps, double* aN,
                              double* eN, cudaStream_t st) {
  HostPack pk;
  const size_t oa = pk.add(a, m), opi = pk.add(P_inf, (size_t)m * m), ops = pk.add(P_star, (size_t)m * m);
  const size_t oz = pk.add(Z, (size_t)p * m * nZ), ot = pk.add(T, (size_t)m * m * nT);
  const size_t orr = pk.add(R, (size_t)m * r * nR), oq = pk.add(Q, (size_t)r * r * nQ);
  DevBuf mats;
  SS_TRY(upload(mats, pk.buf.data(), pk.buf.size(), st));
  const double* d = mats.as<double>();
  DModel md{};
  md.N = N; md.p = p; md.m = m; md.r = r;
  md.a = dmat3(d + oa, m, 1, 1);
  md.P_inf = dmat3(d + opi, m, m, 1);
  md.P_star = dmat3(d + ops, m, m, 1);
  md.Z = dmat3(d + oz, p, m, nZ);
  md.T = dmat3(d + ot, m, m, nT);
  md.R = dmat3(d + orr, m, r, nR);
  md.Q = dmat3(d + oq, r, r, nQ);

  const size_t Np = (size_t)N * p, ns = (size_t)nsim;
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  // gain sequence shared by all draws
  const bool qtr_tv = nQ > 1 || nR > 1;
  DevBuf gb;
  SS_TRY(gb.alloc(sizeof(double) * (pad(Np) * 2 + pad(Np * m) * 2 + pad((size_t)(qtr_tv ? N : 1) * r * m)), st));
  KalmanScratch g{};
  double* base = gb.as<double>();
  g.F1 = base; base += pad(Np);
  g.code = reinterpret_cast<int*>(base); base += pad(Np);
  g.K0 = base; base += pad(Np * m);
  g.K1 = base; base += pad(Np * m);
  g.QtR = base;
  SS_TRY(launch_gains_generic(md, initialisation_steps, &g, st));
  DevBuf vN, pats;
  SS_TRY(vN.alloc(sizeof(double) * Np * ns, st));
  FsArgs A{};
  A.S = nsim; A.N = N; A.p = p; A.m = m; A.r = r;
  A.init_steps = initialisation_steps;
  A.qtr_tv = qtr_tv;
  A.a0 = md.a.p; A.P_inf = md.P_inf.p; A.P_star = md.P_star.p;
  A.Z = md.Z; A.T = md.T; A.R = md.R;
  PatPack pp;
  const size_t pz = pp.rows_of(Z, p, m, nZ), pq = pp.qtr_of(Q, nQ, R, nR, m, r);
  SS_TRY(pats.alloc(pp.words().size() * sizeof(int), st));
  SS_CUDA(cudaMemcpyAsync(pats.get(), pp.words().data(), pp.words().size() * sizeof(int), cudaMemcpyHostToDevice, st));
  const int* pb = pats.as<int>();
  A.Zp = pp.at(pb, pz, p);
  A.Qp = pp.at(pb, pq, r);
  EllPack ep;
  const EllPack::Ref et = ep.add(T, m, m, nT, false), etc = ep.add(T, m, m, nT, true), er = ep.add(R, m, r, nR, false);
  DevBuf elli, ellv;
  SS_TRY(elli.alloc(ep.ints().size() * sizeof(int), st));
  SS_TRY(ellv.alloc(ep.vals().size() * sizeof(double), st));
  SS_CUDA(cudaMemcpyAsync(elli.get(), ep.ints().data(), ep.ints().size() * sizeof(int), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(ellv.get(), ep.vals().data(), ep.vals().size() * sizeof(double), cudaMemcpyHostToDevice, st));
  A.Te = ep.at(et, elli.as<int>(), ellv.as<double>());
  A.Tce = ep.at(etc, elli.as<int>(), ellv.as<double>());
  A.Re = ep.at(er, elli.as<int>(), ellv.as<double>());
  A.g = g;
  A.y = yN; A.v = vN.as<double>(); A.a_smooth = aN; A.eta = eN;
  SS_TRY(launch_fs_draws(A, st));
  // the host staging vectors die with this frame: make sure the (pageable) copies have been consumed
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

}  // namespace ssgpu

extern "C" int ssgpu_FastSmootherC_dev(const double* y, int N, int p, int nsim, int m, int r, const double* a,
                                       const double* P_inf, const double* P_star, const double* Z, int nZ,
                                       const double* T, int nT, const double* R, int nR, const double* Q, int nQ,
                                       int initialisation_steps, double* a_smooth, double* eta, void* stream) {
  SS_TRY(check_fs(y, N, p, nsim, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ));
  SS_ARG(a_smooth && eta, "null output pointer");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  return fast_smoother_core(y, N, p, nsim, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, initialisation_steps,
                            a_smooth, eta, st);
}

extern "C" int ssgpu_FastSmootherC(const double* y, int N, int p, int nsim, int m, int r, const double* a,
                                   const double* P_inf, const double* P_star, const double* Z, int nZ, const double* T,
                                   int nT, const double* R, int nR, const double* Q, int nQ, int initialisation_steps,
                                   int transposed_state, double* a_smooth, double* a_t, double* eta) {
  SS_TRY(check_fs(y, N, p, nsim, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ));
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  const size_t Np = (size_t)N * p, ns = (size_t)nsim;
  // per-draw arrays in the draw-contiguous layout
  DevBuf yR, yN, aN, eN, outR;
  SS_TRY(upload(yR, y, Np * ns, st));
  SS_TRY(yN.alloc(sizeof(double) * Np * ns, st));
  SS_TRY(launch_permute_tks(yR.as<double>(), yN.as<double>(), N, p, nsim, false, st));
  SS_TRY(aN.alloc(sizeof(double) * (size_t)N * m * ns, st));
  SS_TRY(eN.alloc(sizeof(double) * (size_t)N * r * ns, st));
  SS_TRY(fast_smoother_core(yN.as<double>(), N, p, nsim, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ,
                            initialisation_steps, aN.as<double>(), eN.as<double>(), st));
  SS_TRY(outR.alloc(sizeof(double) * (size_t)N * (m > r ? m : r) * ns, st));
  if (a_smooth) {
    SS_TRY(launch_permute_tks(aN.as<double>(), outR.as<double>(), N, m, nsim, true, st));
    SS_TRY(copy_d2h(a_smooth, outR.get(), sizeof(double) * (size_t)N * m * ns, st));
  }
  if (transposed_state && a_t) {
    SS_TRY(launch_permute_at(aN.as<double>(), outR.as<double>(), N, m, nsim, st));
    SS_TRY(copy_d2h(a_t, outR.get(), sizeof(double) * (size_t)N * m * ns, st));
  }
  if (eta) {
    SS_TRY(launch_permute_tks(eN.as<double>(), outR.as<double>(), N, r, nsim, true, st));
    SS_TRY(copy_d2h(eta, outR.get(), sizeof(double) * (size_t)N * r * ns, st));
  }
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

namespace ssgpu {

static int check_sim(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP, const double* a,
                     const double* Z, int nZ, const double* T, int nT, const double* R, int nR, const double* Q, int nQ,
                     const double* P_star, int draw_initial, int eta_only, const double* draws, long long n_draws,
                     const double* P_root) {
  SS_ARG(nsim >= 1 && repeat_Q >= 1 && N >= 1 && r >= 1 && m >= 1, "nsim, repeat_Q, N, r, m must be positive");
  SS_ARG(Q && draws, "null pointer argument");
  SS_ARG(nQ == 1 || nQ == N, "Q must have 1 or N slices");
  const int r_tot = r * repeat_Q;
  const size_t ns = (size_t)nsim, total = (size_t)r_tot * ns;
  // the reference draws the initial state whenever draw_initial is set, eta_only or not (src/SimulateC.cpp:72-76)
  const size_t n_init = draw_initial ? (size_t)m * ns : 0;
  if (n_draws != (long long)(n_init + total * N)) {
    set_error("SimulateC: n_draws does not match the reference's consumption ([m*nsim] + N*r_tot*nsim)");
    return SSGPU_E_DRAWS;
  }
  if (!eta_only) {
    SS_ARG(a && Z && T && R, "null pointer argument");
    SS_ARG(p >= 1 && m <= SSGPU_MAX_M, "p, m out of range");
    SS_ARG(rR == r_tot, "R must have r * repeat_Q columns");
    SS_ARG((nZ == 1 || nZ == N) && (nT == 1 || nT == N) && (nR == 1 || nR == N),
           "system matrices must have 1 or N slices");
    SS_ARG(!draw_initial || P_root || (P_star && mP == m), "draw_initial needs an m x m P_star");
  }
  return SSGPU_OK;
}

// placement of one component's outputs inside the stacked arrays of the whole model (SimArgs, common.cuh)
struct SimPlace {
  int a_ld, a_off, eta_ld, eta_off, y_acc;
};

// SimulateC on device-resident variates (reference order) writing the draw-contiguous layout; yN / aN / eN may be
// null.  System matrices are host pointers.
static int simulate_core(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, const double* a, const double* Z,
                         int nZ, const double* T, int nT, const double* R, int nR, const double* Q, int nQ,
                         const double* P_star, int draw_initial, int eta_only, const double* draws_dev,
                         const double* Q_root, const double* P_root, double* yN, double* aN, double* eN,
                         cudaStream_t st, const SimPlace* place = nullptr) {
  const size_t ns = (size_t)nsim;
  const size_t n_init = draw_initial ? (size_t)m * ns : 0;
  // symmetric roots (host, once) unless supplied
  std::vector<double> qroot((size_t)r * r * nQ), proot;
  for (int i = 0; i < nQ; i++) {
    double* dst = qroot.data() + (size_t)i * r * r;
    if (Q_root)
      std::memcpy(dst, Q_root + (size_t)i * r * r, sizeof(double) * r * r);
    else
      sym_root_host(Q + (size_t)i * r * r, r, dst);
  }
  HostPack pk;
  const size_t oq = pk.add(qroot.data(), qroot.size());
  size_t oa = 0, oz = 0, ot = 0, orr = 0, op = 0;
  if (!eta_only) {
    oa = pk.add(a, m);
    oz = pk.add(Z, (size_t)p * m * nZ);
    ot = pk.add(T, (size_t)m * m * nT);
    orr = pk.add(R, (size_t)m * rR * nR);
    if (draw_initial) {
      proot.resize((size_t)m * m);
      if (P_root)
        std::memcpy(proot.data(), P_root, sizeof(double) * m * m);
      else
        sym_root_host(P_star, m, proot.data());
      op = pk.add(proot.data(), proot.size());
    }
  }
  DevBuf mats, pats, elli, ellv;
  SS_TRY(upload(mats, pk.buf.data(), pk.buf.size(), st));
  PatPack pp;
  EllPack ep;
  const double* d = mats.as<double>();
  SimArgs A{};
  A.S = nsim; A.N = N; A.p = eta_only ? 1 : p; A.m = eta_only ? 1 : m; A.rR = rR; A.r = r; A.repeat_Q = repeat_Q;
  A.nQ = nQ; A.draw_initial = draw_initial; A.eta_only = eta_only;
  A.Q_root = d + oq;
  A.d0 = draws_dev;
  A.dt = draws_dev + n_init;
  if (!eta_only) {
    A.a0 = d + oa;
    A.Z = dmat3(d + oz, p, m, nZ);
    A.T = dmat3(d + ot, m, m, nT);
    A.R = dmat3(d + orr, m, rR, nR);
    A.P_root = d + op;
    const size_t pz = pp.rows_of(Z, p, m, nZ);
    SS_TRY(pats.alloc(pp.words().size() * sizeof(int), st));
    SS_CUDA(cudaMemcpyAsync(pats.get(), pp.words().data(), pp.words().size() * sizeof(int), cudaMemcpyHostToDevice, st));
    A.Zp = pp.at(pats.as<int>(), pz, p);
    const EllPack::Ref et = ep.add(T, m, m, nT, false), er = ep.add(R, m, rR, nR, false);
    SS_TRY(elli.alloc(ep.ints().size() * sizeof(int), st));
    SS_TRY(ellv.alloc(ep.vals().size() * sizeof(double), st));
    SS_CUDA(cudaMemcpyAsync(elli.get(), ep.ints().data(), ep.ints().size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemcpyAsync(ellv.get(), ep.vals().data(), ep.vals().size() * sizeof(double), cudaMemcpyHostToDevice, st));
    A.Te = ep.at(et, elli.as<int>(), ellv.as<double>());
    A.Re = ep.at(er, elli.as<int>(), ellv.as<double>());
    A.y = yN;
    A.a = aN;
  } else {
    A.y = yN;  // residuals: y += eta (R/SimSmoother.R:575-576), null otherwise
  }
  A.eta = eN;
  if (place) {
    A.a_ld = place->a_ld; A.a_off = place->a_off; A.eta_ld = place->eta_ld; A.eta_off = place->eta_off;
    A.y_acc = place->y_acc;
  }
  SS_TRY(launch_simulate(A, st));
  SS_CUDA(cudaStreamSynchronize(st));  // host staging vectors die with this frame
  return SSGPU_OK;
}

}  // namespace ssgpu

extern "C" int ssgpu_SimulateC_dev(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP, const double* a,
                                   const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                                   const double* Q, int nQ, const double* P_star, int draw_initial, int eta_only,
                                   const double* draws, long long n_draws, const double* Q_root, const double* P_root,
                                   double* y, double* a_out, double* eta, void* stream) {
  SS_TRY(check_sim(nsim, repeat_Q, N, p, m, rR, r, mP, a, Z, nZ, T, nT, R, nR, Q, nQ, P_star, draw_initial, eta_only,
                   draws, n_draws, P_root));
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  return simulate_core(nsim, repeat_Q, N, p, m, rR, r, a, Z, nZ, T, nT, R, nR, Q, nQ, P_star, draw_initial, eta_only,
                       draws, Q_root, P_root, y, a_out, eta, st);
}

extern "C" int ssgpu_SimulateC(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP, const double* a,
                               const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                               const double* Q, int nQ, const double* P_star, int draw_initial, int eta_only,
                               int transposed_state, const double* draws, long long n_draws, const double* Q_root,
                               const double* P_root, double* y, double* a_out, double* a_t, double* eta) {
  SS_TRY(check_sim(nsim, repeat_Q, N, p, m, rR, r, mP, a, Z, nZ, T, nT, R, nR, Q, nQ, P_star, draw_initial, eta_only,
                   draws, n_draws, P_root));
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  const int r_tot = r * repeat_Q;
  const size_t ns = (size_t)nsim;
  DevBuf dr, yN, aN, eN, outR;
  SS_TRY(upload(dr, draws, (size_t)n_draws, st));
  size_t biggest = (size_t)N * r_tot * ns;
  if (!eta_only) {
    SS_TRY(yN.alloc(sizeof(double) * (size_t)N * p * ns, st));
    SS_TRY(aN.alloc(sizeof(double) * (size_t)N * m * ns, st));
    if ((size_t)N * m * ns > biggest) biggest = (size_t)N * m * ns;
    if ((size_t)N * p * ns > biggest) biggest = (size_t)N * p * ns;
  }
  SS_TRY(eN.alloc(sizeof(double) * (size_t)N * r_tot * ns, st));
  SS_TRY(simulate_core(nsim, repeat_Q, N, p, m, rR, r, a, Z, nZ, T, nT, R, nR, Q, nQ, P_star, draw_initial, eta_only,
                       dr.as<double>(), Q_root, P_root, eta_only ? nullptr : yN.as<double>(),
                       eta_only ? nullptr : aN.as<double>(), eN.as<double>(), st));
  SS_TRY(outR.alloc(sizeof(double) * biggest, st));
  if (eta) {
    SS_TRY(launch_permute_tks(eN.as<double>(), outR.as<double>(), N, r_tot, nsim, true, st));
    SS_TRY(copy_d2h(eta, outR.get(), sizeof(double) * (size_t)N * r_tot * ns, st));
  }
  if (!eta_only) {
    if (y) {
      SS_TRY(launch_permute_tks(yN.as<double>(), outR.as<double>(), N, p, nsim, true, st));
      SS_TRY(copy_d2h(y, outR.get(), sizeof(double) * (size_t)N * p * ns, st));
    }
    if (a_out) {
      SS_TRY(launch_permute_tks(aN.as<double>(), outR.as<double>(), N, m, nsim, true, st));
      SS_TRY(copy_d2h(a_out, outR.get(), sizeof(double) * (size_t)N * m * ns, st));
    }
    if (transposed_state && a_t) {
      SS_TRY(launch_permute_at(aN.as<double>(), outR.as<double>(), N, m, nsim, st));
      SS_TRY(copy_d2h(a_t, outR.get(), sizeof(double) * (size_t)N * m * ns, st));
    }
  }
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

// ---- SimSmoother as one call (R/SimSmoother.R:34-769) ----------------------------------------------------------
// Everything between the variates and the returned arrays stays on the device: the component draws land directly in
// the stacked arrays of the whole model, y is accumulated in place, the smoother runs on it, the mean correction and
// the component extraction read the same buffers; only epsilon / a / eta / the extracted components cross PCIe.
// full == nullptr: the unconditional half only (predict()'s simulations, R/predict.R:142-856): no smoother, no mean
// correction; y_out (N x p x nsim) receives the simulated data
static int sim_smoother_impl(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps, const double* H,
                             int nH, const ssgpu_sim_model* full, int initialisation_steps, const double* a_hat,
                             const double* eps_hat, const double* eta_hat, const double* draws, long long n_draws,
                             double* y_out, double* epsilon, double* a_out, double* eta, int n_extract,
                             const double* const* Z_extract, const int* nZ_extract, double* const* extracted) {
  SS_ARG(nsim >= 1 && N >= 1 && p >= 1 && n_comp >= 1 && comps && H && draws, "SimSmoother: bad arguments");
  SS_ARG(nH == 1 || nH == N, "SimSmoother: H must have 1 or N slices");
  SS_ARG(!full || (a_hat && eps_hat && eta_hat), "SimSmoother: smoothed estimates of the observed data are required");
  SS_ARG(n_extract == 0 || (Z_extract && nZ_extract && extracted), "SimSmoother: null extraction arguments");
  int m = 0, r = 0;
  long long need = 0;
  const size_t ns = (size_t)nsim;
  for (int c = 0; c < n_comp; c++) {
    const ssgpu_sim_component& C = comps[c];
    SS_ARG(C.m >= 1 && C.r >= 1 && C.repeat_Q >= 1, "SimSmoother: component dimensions must be positive");
    m += C.m;
    r += C.r * C.repeat_Q;
    need += (long long)(C.draw_initial ? (size_t)C.m * ns : 0) + (long long)N * C.r * C.repeat_Q * (long long)ns;
  }
  need += (long long)N * p * (long long)ns;  // residuals
  if (n_draws != need) {
    set_error("SimSmoother: n_draws does not match the reference's consumption (components in order, then N*p*nsim)");
    return SSGPU_E_DRAWS;
  }
  if (full) {
    SS_ARG(full->m == m + p && full->r == r + p, "SimSmoother: the full model must be the components plus p residual states");
    SS_TRY(check_fs(draws, N, p, nsim, full->m, full->r, full->a, full->P_inf, full->P_star, full->Z, full->nZ, full->T,
                    full->nT, full->R, full->nR, full->Q, full->nQ));
  }
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  DevBuf dr, yN, aN, eN, epsN, afN, efN, hats, outR, zD, cN;
  SS_TRY(upload(dr, draws, (size_t)n_draws, st));
  SS_TRY(yN.alloc(sizeof(double) * (size_t)N * p * ns, st));
  SS_TRY(aN.alloc(sizeof(double) * (size_t)N * m * ns, st));
  SS_TRY(eN.alloc(sizeof(double) * (size_t)N * r * ns, st));
  SS_TRY(epsN.alloc(sizeof(double) * (size_t)N * p * ns, st));
  // unconditional draws, component by component (R/SimSmoother.R:65-553), then the residuals (:556-576)
  size_t doff = 0;
  int a_off = 0, e_off = 0;
  for (int c = 0; c < n_comp; c++) {
    const ssgpu_sim_component& C = comps[c];
    const int r_tot = C.r * C.repeat_Q;
    const size_t cnt = (C.draw_initial ? (size_t)C.m * ns : 0) + (size_t)N * r_tot * ns;
    SS_TRY(check_sim(nsim, C.repeat_Q, N, p, C.m, r_tot, C.r, C.m, C.a, C.Z, C.nZ, C.T, C.nT, C.R, C.nR, C.Q, C.nQ,
                     C.P_star, C.draw_initial, 0, draws + doff, (long long)cnt, C.P_root));
    const SimPlace pl{m, a_off, r, e_off, c > 0};
    SS_TRY(simulate_core(nsim, C.repeat_Q, N, p, C.m, r_tot, C.r, C.a, C.Z, C.nZ, C.T, C.nT, C.R, C.nR, C.Q, C.nQ,
                         C.P_star, C.draw_initial, 0, dr.as<double>() + doff, C.Q_root, C.P_root, yN.as<double>(),
                         aN.as<double>(), eN.as<double>(), st, &pl));
    doff += cnt;
    a_off += C.m;
    e_off += r_tot;
  }
  SS_TRY(simulate_core(nsim, 1, N, 1, 1, p, p, nullptr, nullptr, 1, nullptr, 1, nullptr, 1, H, nH, nullptr, 0, 1,
                       dr.as<double>() + doff, nullptr, nullptr, yN.as<double>(), nullptr, epsN.as<double>(), st));
  HostPack hp;
  if (full) {
    // smoothed state and disturbances of the simulated data (:617-632)
    SS_TRY(afN.alloc(sizeof(double) * (size_t)N * full->m * ns, st));
    SS_TRY(efN.alloc(sizeof(double) * (size_t)N * full->r * ns, st));
    SS_TRY(fast_smoother_core(yN.as<double>(), N, p, nsim, full->m, full->r, full->a, full->P_inf, full->P_star, full->Z,
                              full->nZ, full->T, full->nT, full->R, full->nR, full->Q, full->nQ, initialisation_steps,
                              afN.as<double>(), efN.as<double>(), st));
    // mean corrections (:634-642)
    const size_t oa = hp.add(a_hat, (size_t)N * m), oe = hp.add(eps_hat, (size_t)N * p), oh = hp.add(eta_hat, (size_t)N * r);
    SS_TRY(upload(hats, hp.buf.data(), hp.buf.size(), st));
    SS_TRY(launch_sim_correct(aN.as<double>(), afN.as<double>(), hats.as<double>() + oa, N, m, full->m, p, nsim, st));
    SS_TRY(launch_sim_correct(epsN.as<double>(), afN.as<double>(), hats.as<double>() + oe, N, p, full->m, 0, nsim, st));
    SS_TRY(launch_sim_correct(eN.as<double>(), efN.as<double>(), hats.as<double>() + oh, N, r, full->r, p, nsim, st));
  }
  // results in R's layouts
  const int kmax = m > r ? (m > p ? m : p) : (r > p ? r : p);
  SS_TRY(outR.alloc(sizeof(double) * (size_t)N * kmax * ns, st));
  struct Out {
    double* host;
    const double* dev;
    int K;
  } outs[4] = {{epsilon, epsN.as<double>(), p}, {a_out, aN.as<double>(), m}, {eta, eN.as<double>(), r},
               {y_out, yN.as<double>(), p}};
  for (const Out& o : outs) {
    if (!o.host) continue;
    SS_TRY(launch_permute_tks(o.dev, outR.as<double>(), N, o.K, nsim, true, st));
    SS_TRY(copy_d2h(o.host, outR.get(), sizeof(double) * (size_t)N * o.K * ns, st));
  }
  // components (:645-766): Z_padded a_t per draw, from the corrected states
  if (n_extract > 0) SS_TRY(cN.alloc(sizeof(double) * (size_t)N * p * ns, st));
  for (int e = 0; e < n_extract; e++) {
    SS_ARG(Z_extract[e] && extracted[e] && (nZ_extract[e] == 1 || nZ_extract[e] == N), "SimSmoother: bad extraction matrix");
    SS_TRY(upload(zD, Z_extract[e], (size_t)p * m * nZ_extract[e], st));
    SS_TRY(launch_extract(aN.as<double>(), 1, dmat3(zD.as<double>(), p, m, nZ_extract[e]), m, p, nsim, N, cN.as<double>(), st));
    SS_TRY(launch_permute_tks(cN.as<double>(), outR.as<double>(), N, p, nsim, true, st));
    SS_TRY(copy_d2h(extracted[e], outR.get(), sizeof(double) * (size_t)N * p * ns, st));
  }
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

// Draws [lo, hi) of a call as their own call: the variates of the reference's consumption order (per component
// [m x nsim initial draws] then per time point [r_tot x nsim], draws slowest inside each block; the residuals last) are
// gathered for the range, the outputs are R arrays with the draw index slowest, so the range is a contiguous tail block.
static int sim_smoother_range(long long lo, long long hi, int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps,
                              const double* H, int nH, const ssgpu_sim_model* full, int initialisation_steps,
                              const double* a_hat, const double* eps_hat, const double* eta_hat, const double* draws,
                              double* y_out, double* epsilon, double* a_out, double* eta, int n_extract,
                              const double* const* Z_extract, const int* nZ_extract, double* const* extracted) {
  const size_t ns = (size_t)nsim, n = (size_t)(hi - lo);
  std::vector<double> sub;
  size_t off = 0;
  int m = 0, r = 0;
  auto take = [&](size_t width, size_t blocks) {  // `blocks` blocks of width x nsim doubles each
    for (size_t t = 0; t < blocks; t++) {
      const double* src = draws + off + t * width * ns + width * (size_t)lo;
      sub.insert(sub.end(), src, src + width * n);
    }
    off += blocks * width * ns;
  };
  for (int c = 0; c < n_comp; c++) {
    const ssgpu_sim_component& C = comps[c];
    if (C.draw_initial) take((size_t)C.m, 1);
    take((size_t)C.r * C.repeat_Q, (size_t)N);
    m += C.m;
    r += C.r * C.repeat_Q;
  }
  take((size_t)p, (size_t)N);
  std::vector<double*> ext(n_extract > 0 ? n_extract : 0);
  for (int i = 0; i < n_extract; i++) ext[i] = extracted[i] + (size_t)N * p * (size_t)lo;
  auto at = [&](double* x, size_t K) { return x ? x + (size_t)N * K * (size_t)lo : nullptr; };
  return sim_smoother_impl((int)n, N, p, n_comp, comps, H, nH, full, initialisation_steps, a_hat, eps_hat, eta_hat,
                           sub.data(), (long long)sub.size(), at(y_out, p), at(epsilon, p), at(a_out, m), at(eta, r), n_extract,
                           Z_extract, nZ_extract, n_extract > 0 ? ext.data() : nullptr);
}

// Several devices in use (ssgpu_use_devices): the draws are cut into contiguous ranges, one worker thread per device.
static int sim_smoother_sharded(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps, const double* H,
                                int nH, const ssgpu_sim_model* full, int initialisation_steps, const double* a_hat,
                                const double* eps_hat, const double* eta_hat, const double* draws, long long n_draws,
                                double* y_out, double* epsilon, double* a_out, double* eta, int n_extract,
                                const double* const* Z_extract, const int* nZ_extract, double* const* extracted) {
  if (devices_in_use() > 1 && nsim >= 64 * devices_in_use() && n_comp >= 1 && comps && draws) {
    long long need = (long long)N * p * nsim;
    for (int c = 0; c < n_comp; c++)
      need += (long long)(comps[c].draw_initial ? (long long)comps[c].m * nsim : 0) +
              (long long)N * comps[c].r * comps[c].repeat_Q * (long long)nsim;
    if (need == n_draws)  // a wrong count is reported by the unsharded path below
      return for_each_device(nsim, 32, [&](int, long long lo, long long hi) {
        return sim_smoother_range(lo, hi, nsim, N, p, n_comp, comps, H, nH, full, initialisation_steps, a_hat, eps_hat, eta_hat,
                                  draws, y_out, epsilon, a_out, eta, n_extract, Z_extract, nZ_extract, extracted);
      });
  }
  return sim_smoother_impl(nsim, N, p, n_comp, comps, H, nH, full, initialisation_steps, a_hat, eps_hat, eta_hat, draws,
                           n_draws, y_out, epsilon, a_out, eta, n_extract, Z_extract, nZ_extract, extracted);
}

extern "C" int ssgpu_SimSmoother(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps, const double* H,
                                 int nH, const ssgpu_sim_model* full, int initialisation_steps,
                                 const double* a_hat, const double* eps_hat, const double* eta_hat, const double* draws,
                                 long long n_draws, double* epsilon, double* a_out, double* eta, int n_extract,
                                 const double* const* Z_extract, const int* nZ_extract, double* const* extracted) {
  SS_ARG(full != nullptr, "SimSmoother: the full model is required");
  return sim_smoother_sharded(nsim, N, p, n_comp, comps, H, nH, full, initialisation_steps, a_hat, eps_hat, eta_hat, draws,
                              n_draws, nullptr, epsilon, a_out, eta, n_extract, Z_extract, nZ_extract, extracted);
}

extern "C" int ssgpu_SimulateModel(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps, const double* H,
                                   int nH, const double* draws, long long n_draws, double* y, double* epsilon,
                                   double* a_out, double* eta, int n_extract, const double* const* Z_extract,
                                   const int* nZ_extract, double* const* extracted) {
  return sim_smoother_sharded(nsim, N, p, n_comp, comps, H, nH, nullptr, 0, nullptr, nullptr, nullptr, draws, n_draws, y,
                              epsilon, a_out, eta, n_extract, Z_extract, nZ_extract, extracted);
}

extern "C" int ssgpu_ExtractComponentC_dev(const double* a, int m, int nsim, int N, const double* Z, int p, int nZ,
                                           double* out, void* stream) {
  SS_ARG(a && Z && out, "null pointer argument");
  SS_ARG(m >= 1 && nsim >= 1 && N >= 1 && p >= 1 && (nZ == 1 || nZ == N), "bad dimensions");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  DevBuf zD;
  SS_TRY(upload(zD, Z, (size_t)p * m * nZ, st));
  SS_TRY(launch_extract(a, 1, dmat3(zD.as<double>(), p, m, nZ), m, p, nsim, N, out, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

extern "C" int ssgpu_last_sim_kernel_ms(double* simulate_ms, double* fast_smoother_ms, double* extract_ms) {
  last_sim_kernel_ms(simulate_ms, fast_smoother_ms, extract_ms);
  return SSGPU_OK;
}

extern "C" int ssgpu_draw_layout_dev(const double* src, double* dst, int N, int K, long long nsim, int to_r, void* stream) {
  SS_ARG(src && dst && N >= 1 && K >= 1 && nsim >= 1, "bad arguments");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  return launch_permute_tks(src, dst, N, K, nsim, to_r != 0, st);
}

extern "C" int ssgpu_ExtractComponentC(const double* a, int m, int nsim, int N, const double* Z, int p, int nZ,
                                       double* out) {
  SS_ARG(a && Z && out, "null pointer argument");
  SS_ARG(m >= 1 && nsim >= 1 && N >= 1 && p >= 1 && (nZ == 1 || nZ == N), "bad dimensions");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  DevBuf aD, zD, oN, oR;
  const size_t ns = (size_t)nsim;
  SS_TRY(upload(aD, a, (size_t)m * ns * N, st));
  SS_TRY(upload(zD, Z, (size_t)p * m * nZ, st));
  SS_TRY(oN.alloc(sizeof(double) * (size_t)N * p * ns, st));
  SS_TRY(oR.alloc(sizeof(double) * (size_t)N * p * ns, st));
  SS_TRY(launch_extract(aD.as<double>(), 0, dmat3(zD.as<double>(), p, m, nZ), m, p, nsim, N, oN.as<double>(), st));
  SS_TRY(launch_permute_tks(oN.as<double>(), oR.as<double>(), N, p, nsim, true, st));
  SS_TRY(copy_d2h(out, oR.get(), sizeof(double) * (size_t)N * p * ns, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}
