// C entry points around the UNMODIFIED reference sources (TEST / BASELINE INFRASTRUCTURE).
//
// Linked together with /root/reference/src/{LogLikC,KalmanC,FastSmootherC,SimulateC,
// ExtractComponentC}.cpp (compiled in place against oracle/refshim/RcppArmadillo.h, see
// oracle/Makefile) into oracle/_ref/libssref.so.  Nothing from the reference is copied here: this
// file only declares the five functions with the signatures they have in the reference
// (src/RcppExports.cpp:14-105) and converts flat column-major buffers to the stand-in types.
//
// All arrays are column-major with the R shapes; 3-D system matrices carry a slice count
// (1 = time-invariant, N = time-varying).
#include <RcppArmadillo.h>

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif

// reference declarations (definitions live in /root/reference/src/*.cpp)
double LogLikC(const Rcpp::NumericMatrix& y, const Rcpp::LogicalMatrix& y_isna, arma::colvec a, arma::mat P_inf,
               arma::mat P_star, const arma::cube& Z, const arma::cube& T, const arma::cube& R, const arma::cube& Q);
Rcpp::List KalmanC(const arma::mat& y, const Rcpp::LogicalMatrix& y_isna, const arma::colvec& a,
                   const arma::mat& P_inf, const arma::mat& P_star, const arma::cube& Z, const arma::cube& T,
                   const arma::cube& R, const arma::cube& Q, const bool& diagnostics);
Rcpp::List FastSmootherC(const arma::cube& y, const arma::colvec& a, const arma::mat& P_inf, const arma::mat& P_star,
                         const arma::cube& Z, const arma::cube& T, const arma::cube& R, const arma::cube& Q,
                         const int& initialisation_steps, const bool& transposed_state);
Rcpp::List SimulateC(const int& nsim, const int& repeat_Q, const int& N, const arma::colvec& a, const arma::cube& Z,
                     const arma::cube& T, const arma::cube& R, const arma::cube& Q, const arma::mat& P_star,
                     const bool& draw_initial, const bool& eta_only, const bool& transposed_state);
arma::cube ExtractComponentC(const arma::cube& a, const arma::cube& Z);

namespace {
thread_local std::string g_err;

void put_mat(const Rcpp::Value& v, double* dst) {
  if (!dst) return;
  if (v.kind == Rcpp::Value::MAT)
    std::memcpy(dst, v.m.memptr(), v.m.n_elem * sizeof(double));
  else if (v.kind == Rcpp::Value::CUBE)
    v.c.copy_out(dst);
  else if (v.kind == Rcpp::Value::VEC)
    std::memcpy(dst, v.vec.data(), v.vec.size() * sizeof(double));
  else
    throw std::runtime_error("put_mat: unexpected value kind");
}
}  // namespace

extern "C" {

const char* ssref_last_error() { return g_err.c_str(); }

int ssref_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// LogLikC (src/LogLikC.cpp:20).  Returns 0 on success.
int ssref_loglik(const double* y, const int* y_isna, int N, int p, int m, int r, const double* a, const double* P_inf,
                 const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                 const double* Q, int nQ, double* out) {
  try {
    *out = LogLikC(Rcpp::NumericMatrix(y, N, p), Rcpp::LogicalMatrix(y_isna, N, p), arma::colvec(a, m),
                   arma::mat(P_inf, m, m), arma::mat(P_star, m, m), arma::cube(Z, p, m, nZ), arma::cube(T, m, m, nT),
                   arma::cube(R, m, r, nR), arma::cube(Q, r, r, nQ));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Batch of independent LogLikC calls (one series per task), OpenMP over series: the CPU baseline.
// y: [B][N*p] (series-major, each N x p column-major); per-series P_star/Q (stride m*m / r*r) when
// per_series != 0, shared otherwise.  Time-invariant matrices only (config-5 shape).
int ssref_loglik_batch(const double* y, const int* y_isna, long B, int N, int p, int m, int r, const double* a,
                       const double* P_inf, const double* P_star, const double* Z, const double* T, const double* R,
                       const double* Q, int per_series, int threads, double* out) {
  int fail = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (long b = 0; b < B; b++) {
    const double* Ps = per_series ? P_star + b * (long)m * m : P_star;
    const double* Qb = per_series ? Q + b * (long)r * r : Q;
    if (ssref_loglik(y + b * (long)N * p, y_isna + b * (long)N * p, N, p, m, r, a, P_inf, Ps, Z, 1, T, 1, R, 1, Qb, 1,
                     out + b))
      fail = 1;
  }
  return fail;
}

// KalmanC (src/KalmanC.cpp:21).  Any output pointer may be NULL.
struct ssref_kalman_out {
  int* initialisation_steps;
  double *loglik, *a_pred, *a_fil, *a_smooth, *P_pred, *P_fil, *V, *P_inf_pred, *P_star_pred, *P_inf_fil, *P_star_fil,
      *yfit, *v, *v_norm, *eta, *e, *Fmat, *eta_var, *D, *a_fc, *P_inf_fc, *P_star_fc, *P_fc, *Tstat_observation,
      *Tstat_state, *r_vec, *Nmat;
};

int ssref_kalman(const double* y, const int* y_isna, int N, int p, int m, int r, const double* a, const double* P_inf,
                 const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                 const double* Q, int nQ, int diagnostics, ssref_kalman_out* o) {
  try {
    Rcpp::List res = KalmanC(arma::mat(y, N, p), Rcpp::LogicalMatrix(y_isna, N, p), arma::colvec(a, m),
                             arma::mat(P_inf, m, m), arma::mat(P_star, m, m), arma::cube(Z, p, m, nZ),
                             arma::cube(T, m, m, nT), arma::cube(R, m, r, nR), arma::cube(Q, r, r, nQ),
                             diagnostics != 0);
    const Rcpp::List& nested = *res.at("nested").l;
    if (o->initialisation_steps) *o->initialisation_steps = nested.at("initialisation_steps").i;
    put_mat(nested.at("loglik"), o->loglik);
    put_mat(nested.at("a_pred"), o->a_pred);
    put_mat(nested.at("a_fil"), o->a_fil);
    put_mat(nested.at("a_smooth"), o->a_smooth);
    put_mat(nested.at("P_pred"), o->P_pred);
    put_mat(nested.at("P_fil"), o->P_fil);
    put_mat(nested.at("V"), o->V);
    put_mat(nested.at("P_inf_pred"), o->P_inf_pred);
    put_mat(res.at("P_star_pred"), o->P_star_pred);
    put_mat(res.at("P_inf_fil"), o->P_inf_fil);
    put_mat(res.at("P_star_fil"), o->P_star_fil);
    put_mat(res.at("yfit"), o->yfit);
    put_mat(res.at("v"), o->v);
    put_mat(res.at("v_norm"), o->v_norm);
    put_mat(res.at("eta"), o->eta);
    put_mat(res.at("e"), o->e);
    put_mat(res.at("Fmat"), o->Fmat);
    put_mat(res.at("eta_var"), o->eta_var);
    put_mat(res.at("D"), o->D);
    put_mat(res.at("a_fc"), o->a_fc);
    put_mat(res.at("P_inf_fc"), o->P_inf_fc);
    put_mat(res.at("P_star_fc"), o->P_star_fc);
    put_mat(res.at("P_fc"), o->P_fc);
    put_mat(res.at("Tstat_observation"), o->Tstat_observation);
    put_mat(res.at("Tstat_state"), o->Tstat_state);
    put_mat(res.at("r_vec"), o->r_vec);
    put_mat(res.at("Nmat"), o->Nmat);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// FastSmootherC (src/FastSmootherC.cpp:23).  y: N x p x nsim.
int ssref_fast_smoother(const double* y, int N, int p, int nsim, int m, int r, const double* a, const double* P_inf,
                        const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                        const double* Q, int nQ, int initialisation_steps, int transposed_state, double* a_smooth,
                        double* a_t, double* eta) {
  try {
    Rcpp::List res =
        FastSmootherC(arma::cube(y, N, p, nsim), arma::colvec(a, m), arma::mat(P_inf, m, m), arma::mat(P_star, m, m),
                      arma::cube(Z, p, m, nZ), arma::cube(T, m, m, nT), arma::cube(R, m, r, nR), arma::cube(Q, r, r, nQ),
                      initialisation_steps, transposed_state != 0);
    put_mat(res.at("a_smooth"), a_smooth);
    if (transposed_state) put_mat(res.at("a_t"), a_t);
    put_mat(res.at("eta"), eta);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// SimulateC (src/SimulateC.cpp:26) with the normal variates injected (`draws`, consumed in the
// reference's Rcpp::rnorm call order).  Z: p x m, T: m x m, R: m x rR, Q: r x r, P_star: mP x mP
// (the eta_only caller passes 1x1x1 dummies exactly like R/SimSmoother.R:559-573).
int ssref_simulate(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP, const double* a,
                   const double* Z, int nZ, const double* T, int nT, const double* R, int nR, const double* Q, int nQ,
                   const double* P_star, int draw_initial, int eta_only, int transposed_state, const double* draws,
                   long n_draws, double* y, double* a_out, double* a_t, double* eta) {
  try {
    Rcpp::DrawStream& s = Rcpp::draw_stream();
    s.p = draws;
    s.n = n_draws;
    s.pos = 0;
    Rcpp::List res = SimulateC(nsim, repeat_Q, N, arma::colvec(a, m), arma::cube(Z, p, m, nZ), arma::cube(T, m, m, nT),
                               arma::cube(R, m, rR, nR), arma::cube(Q, r, r, nQ), arma::mat(P_star, mP, mP),
                               draw_initial != 0, eta_only != 0, transposed_state != 0);
    if (s.pos != s.n) throw std::runtime_error("ssref_simulate: injected draws not fully consumed");
    if (!eta_only) {
      put_mat(res.at("y"), y);
      put_mat(res.at("a"), a_out);
      if (transposed_state) put_mat(res.at("a_t"), a_t);
    }
    put_mat(res.at("eta"), eta);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ExtractComponentC (src/ExtractComponentC.cpp:12).  a: m x nsim x N.
int ssref_extract(const double* a, int m, int nsim, int N, const double* Z, int p, int nZ, double* out) {
  try {
    arma::cube c = ExtractComponentC(arma::cube(a, m, nsim, N), arma::cube(Z, p, m, nZ));
    c.copy_out(out);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // extern "C"
