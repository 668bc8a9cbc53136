/* .Call shim: the five registered native routines of statespacer (src/RcppExports.cpp:108-115, same
 * names, same arity) plus `_statespacer_LogLikBatchC`, `_statespacer_LogLikGradBatchC` and `_statespacer_SimSmootherC`, implemented on libssgpu (include/ssgpu.h) with
 * R's plain C API -- no Rcpp, no Armadillo.
 *
 * Build inside the package (src/Makevars):
 *     PKG_CPPFLAGS = -I$(SSGPU_HOME)/include
 *     PKG_LIBS     = -L$(SSGPU_HOME)/statespacer_b200 -lssgpu -Wl,-rpath,$(SSGPU_HOME)/statespacer_b200
 * and delete LogLikC.cpp, KalmanC.cpp, FastSmootherC.cpp, SimulateC.cpp, ExtractComponentC.cpp and
 * RcppExports.cpp from src/.  R is not installed in the image this repository is developed in: there this file
 * is only type-checked against declaration-only stand-ins of R's headers (tests/rstub, tests/test_abi.py);
 * everything below the boundary is exercised through the same C ABI from Python (tests/test_gpu_parity.py).
 *
 * Contract kept from the reference: inputs are read-only; results are freshly allocated and
 * PROTECTed; 3-D arguments carry a length-3 dim; the RNG state is read/written around SimulateC exactly like
 * Rcpp::RNGScope does, and the normal variates are drawn on the host with norm_rand() in the reference's
 * order (src/SimulateC.cpp:75,112) so that R's random stream is unchanged; every call synchronises before
 * returning; a non-zero status becomes an R error after all device work of the call has been released.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <string.h>

#include "ssgpu.h"

static void chk(int rc) {
  if (rc != SSGPU_OK) Rf_error("libssgpu: %s", ssgpu_last_error());
}

/* dims of a matrix / 3-D array argument: rows, cols, slices (1 if 2-D) */
static void dims3(SEXP x, int* r, int* c, int* s) {
  SEXP d = Rf_getAttrib(x, R_DimSymbol);
  if (d == R_NilValue) {
    *r = Rf_length(x); *c = 1; *s = 1;
  } else {
    *r = INTEGER(d)[0];
    *c = Rf_length(d) > 1 ? INTEGER(d)[1] : 1;
    *s = Rf_length(d) > 2 ? INTEGER(d)[2] : 1;
  }
}

/* The reference's Rcpp signatures (NumericMatrix, arma::cube, LogicalMatrix) coerce what they are handed; REAL() /
 * LOGICAL() / INTEGER() on a vector of another type is undefined behaviour.  The R stubs (r/R/ssgpu_glue.R) coerce
 * with storage.mode<-; a caller that reaches .Call directly with an integer matrix gets an R error here instead. */
static void need(SEXP x, SEXPTYPE type, const char* what) {
  if (x != R_NilValue && TYPEOF(x) != type)
    Rf_error("libssgpu shim: argument `%s` must be of type %s (the stubs of ssgpu_glue.R coerce it)", what,
             type == REALSXP ? "double" : type == LGLSXP ? "logical" : "integer");
}
static void need_model(SEXP a, SEXP P_inf, SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q) {
  need(a, REALSXP, "a"); need(P_inf, REALSXP, "P_inf"); need(P_star, REALSXP, "P_star"); need(Z, REALSXP, "Z");
  need(T, REALSXP, "T"); need(R, REALSXP, "R"); need(Q, REALSXP, "Q");
  int m, mc, ms, im, imc, ims, zr, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ;
  dims3(P_star, &m, &mc, &ms); dims3(P_inf, &im, &imc, &ims); dims3(Z, &zr, &zc, &nZ); dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR); dims3(Q, &qr, &qc, &nQ);
  /* Armadillo throws on these in the reference (src/LogLikC.cpp:101,202); here they would be out-of-bounds reads */
  if (mc != m || im != m || imc != m || Rf_length(a) != m || zc != m || tr != m || tc != m || rr != m || qr != r || qc != r)
    Rf_error("libssgpu shim: non-conformable system matrices (m = %d: a %d, P_inf %d x %d, Z . x %d, T %d x %d, R %d x %d, Q %d x %d)",
             m, Rf_length(a), im, imc, zc, tr, tc, rr, r, qr, qc);
}

static SEXP alloc3(int a, int b, int c) {
  SEXP x = PROTECT(Rf_alloc3DArray(REALSXP, a, b, c));
  return x;
}

SEXP _statespacer_LogLikC(SEXP y, SEXP y_isna, SEXP a, SEXP P_inf, SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q) {
  need(y, REALSXP, "y"); need(y_isna, LGLSXP, "y_isna"); need_model(a, P_inf, P_star, Z, T, R, Q);
  int N, p, one, m, mc, ms, zr, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ;
  dims3(y, &N, &p, &one);
  dims3(P_star, &m, &mc, &ms);
  dims3(Z, &zr, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  double out = NA_REAL;
  chk(ssgpu_LogLikC(REAL(y), LOGICAL(y_isna), N, p, m, r, REAL(a), REAL(P_inf), REAL(P_star), REAL(Z), nZ,
                    REAL(T), nT, REAL(R), nR, REAL(Q), nQ, &out));
  return Rf_ScalarReal(ISNAN(out) ? NA_REAL : out); /* src/LogLikC.cpp:211-213 */
}

SEXP _statespacer_KalmanC(SEXP y, SEXP y_isna, SEXP a, SEXP P_inf, SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q,
                          SEXP diagnostics) {
  need(y, REALSXP, "y"); need(y_isna, LGLSXP, "y_isna"); need_model(a, P_inf, P_star, Z, T, R, Q);
  int N, p, one, m, mc, ms, zr, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ, np = 0;
  dims3(y, &N, &p, &one);
  dims3(P_star, &m, &mc, &ms);
  dims3(Z, &zr, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  ssgpu_kalman_out o;
  memset(&o, 0, sizeof o);
  int steps = 0;
  o.initialisation_steps = &steps;
  /* the list of src/KalmanC.cpp:563-593 */
  SEXP loglik = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)N * p)); np++;
  SEXP a_pred = PROTECT(Rf_allocMatrix(REALSXP, N, m)); np++;
  SEXP a_fil = PROTECT(Rf_allocMatrix(REALSXP, N, m)); np++;
  SEXP a_smooth = PROTECT(Rf_allocMatrix(REALSXP, N, m)); np++;
  SEXP P_pred = alloc3(m, m, N); np++;
  SEXP P_fil = alloc3(m, m, N); np++;
  SEXP V = alloc3(m, m, N); np++;
  SEXP P_inf_pred = alloc3(m, m, N); np++;
  SEXP P_star_pred = alloc3(m, m, N); np++;
  SEXP P_inf_fil = alloc3(m, m, N); np++;
  SEXP P_star_fil = alloc3(m, m, N); np++;
  SEXP yfit = PROTECT(Rf_allocMatrix(REALSXP, N, p)); np++;
  SEXP v = PROTECT(Rf_allocMatrix(REALSXP, N, p)); np++;
  SEXP eta = PROTECT(Rf_allocMatrix(REALSXP, N, r)); np++;
  SEXP Fmat = alloc3(p, p, N); np++;
  SEXP eta_var = alloc3(r, r, N); np++;
  SEXP a_fc = PROTECT(Rf_allocVector(REALSXP, m)); np++;
  SEXP P_inf_fc = PROTECT(Rf_allocMatrix(REALSXP, m, m)); np++;
  SEXP P_star_fc = PROTECT(Rf_allocMatrix(REALSXP, m, m)); np++;
  SEXP P_fc = PROTECT(Rf_allocMatrix(REALSXP, m, m)); np++;
  SEXP r_vec = PROTECT(Rf_allocMatrix(REALSXP, N + 1, m)); np++;
  SEXP Nmat = alloc3(m, m, N + 1); np++;
  /* diagnostics arrays exist in the list either way (src/KalmanC.cpp:563-593); v_norm defaults to NA (:81) */
  SEXP v_norm = PROTECT(Rf_allocMatrix(REALSXP, N, p)); np++;
  SEXP e = PROTECT(Rf_allocMatrix(REALSXP, N, p)); np++;
  SEXP D = alloc3(p, p, N); np++;
  SEXP Tstat_observation = PROTECT(Rf_allocMatrix(REALSXP, N, p)); np++;
  SEXP Tstat_state = PROTECT(Rf_allocMatrix(REALSXP, N + 1, m)); np++;
  for (R_xlen_t i = 0; i < (R_xlen_t)N * p; i++) { REAL(v_norm)[i] = NA_REAL; REAL(e)[i] = 0; REAL(Tstat_observation)[i] = 0; }
  memset(REAL(D), 0, sizeof(double) * (size_t)p * p * N);
  memset(REAL(Tstat_state), 0, sizeof(double) * (size_t)(N + 1) * m);
  if (Rf_asLogical(diagnostics)) {
    o.v_norm = REAL(v_norm); o.e = REAL(e); o.D = REAL(D);
    o.Tstat_observation = REAL(Tstat_observation); o.Tstat_state = REAL(Tstat_state);
  }
  o.loglik = REAL(loglik); o.a_pred = REAL(a_pred); o.a_fil = REAL(a_fil); o.a_smooth = REAL(a_smooth);
  o.P_pred = REAL(P_pred); o.P_fil = REAL(P_fil); o.V = REAL(V); o.P_inf_pred = REAL(P_inf_pred);
  o.P_star_pred = REAL(P_star_pred); o.P_inf_fil = REAL(P_inf_fil); o.P_star_fil = REAL(P_star_fil);
  o.yfit = REAL(yfit); o.v = REAL(v); o.eta = REAL(eta); o.Fmat = REAL(Fmat); o.eta_var = REAL(eta_var);
  o.a_fc = REAL(a_fc); o.P_inf_fc = REAL(P_inf_fc); o.P_star_fc = REAL(P_star_fc); o.P_fc = REAL(P_fc);
  o.r_vec = REAL(r_vec); o.Nmat = REAL(Nmat);
  int rc = ssgpu_KalmanC(REAL(y), LOGICAL(y_isna), N, p, m, r, REAL(a), REAL(P_inf), REAL(P_star), REAL(Z), nZ,
                         REAL(T), nT, REAL(R), nR, REAL(Q), nQ, Rf_asLogical(diagnostics), &o);
  if (rc != SSGPU_OK) { UNPROTECT(np); chk(rc); }
  for (R_xlen_t i = 0; i < (R_xlen_t)N * p; i++)
    if (ISNAN(REAL(loglik)[i])) REAL(loglik)[i] = NA_REAL; /* src/KalmanC.cpp:194 */
  const char* nested_names[] = {"initialisation_steps", "loglik", "a_pred", "a_fil", "a_smooth", "P_pred", "P_fil",
                                "V", "P_inf_pred", ""};
  SEXP nested = PROTECT(Rf_mkNamed(VECSXP, nested_names)); np++;
  SET_VECTOR_ELT(nested, 0, Rf_ScalarInteger(steps));
  SET_VECTOR_ELT(nested, 1, loglik); SET_VECTOR_ELT(nested, 2, a_pred); SET_VECTOR_ELT(nested, 3, a_fil);
  SET_VECTOR_ELT(nested, 4, a_smooth); SET_VECTOR_ELT(nested, 5, P_pred); SET_VECTOR_ELT(nested, 6, P_fil);
  SET_VECTOR_ELT(nested, 7, V); SET_VECTOR_ELT(nested, 8, P_inf_pred);
  if (Rf_asLogical(diagnostics))
    for (R_xlen_t i = 0; i < (R_xlen_t)N * p; i++) {
      if (ISNAN(REAL(v_norm)[i])) REAL(v_norm)[i] = NA_REAL;
      if (ISNAN(REAL(e)[i])) REAL(e)[i] = NA_REAL;
    }
  const char* names[] = {"nested", "P_star_pred", "P_inf_fil", "P_star_fil", "yfit", "v", "v_norm", "eta", "e", "Fmat",
                         "eta_var", "D", "a_fc", "P_inf_fc", "P_star_fc", "P_fc", "Tstat_observation", "Tstat_state",
                         "r_vec", "Nmat", ""};
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names)); np++;
  SEXP vals[] = {nested, P_star_pred, P_inf_fil, P_star_fil, yfit, v, v_norm, eta, e, Fmat, eta_var, D, a_fc, P_inf_fc,
                 P_star_fc, P_fc, Tstat_observation, Tstat_state, r_vec, Nmat};
  for (int i = 0; i < 20; i++) SET_VECTOR_ELT(res, i, vals[i]);
  UNPROTECT(np);
  return res;
}

SEXP _statespacer_FastSmootherC(SEXP y, SEXP a, SEXP P_inf, SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q,
                                SEXP initialisation_steps, SEXP transposed_state) {
  need(y, REALSXP, "y"); need_model(a, P_inf, P_star, Z, T, R, Q);
  int N, p, nsim, m, mc, ms, zr, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ;
  dims3(y, &N, &p, &nsim);
  dims3(P_star, &m, &mc, &ms);
  dims3(Z, &zr, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  const int tst = Rf_asLogical(transposed_state);
  SEXP a_smooth = alloc3(N, m, nsim), a_t = alloc3(m, nsim, N), eta = alloc3(N, r, nsim);
  if (!tst) memset(REAL(a_t), 0, sizeof(double) * (size_t)m * nsim * N);
  int rc = ssgpu_FastSmootherC(REAL(y), N, p, nsim, m, r, REAL(a), REAL(P_inf), REAL(P_star), REAL(Z), nZ, REAL(T),
                               nT, REAL(R), nR, REAL(Q), nQ, Rf_asInteger(initialisation_steps), tst, REAL(a_smooth),
                               tst ? REAL(a_t) : NULL, REAL(eta));
  if (rc != SSGPU_OK) { UNPROTECT(3); chk(rc); }
  const char* names[] = {"a_smooth", "a_t", "eta", ""}; /* src/FastSmootherC.cpp:370-374 */
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, a_smooth); SET_VECTOR_ELT(res, 1, a_t); SET_VECTOR_ELT(res, 2, eta);
  UNPROTECT(4);
  return res;
}

SEXP _statespacer_SimulateC(SEXP nsim_, SEXP repeat_Q_, SEXP N_, SEXP a, SEXP Z, SEXP T, SEXP R, SEXP Q, SEXP P_star,
                            SEXP draw_initial_, SEXP eta_only_, SEXP transposed_state_) {
  need(a, REALSXP, "a"); need(Z, REALSXP, "Z"); need(T, REALSXP, "T"); need(R, REALSXP, "R"); need(Q, REALSXP, "Q");
  need(P_star, REALSXP, "P_star");
  const int nsim = Rf_asInteger(nsim_), repeat_Q = Rf_asInteger(repeat_Q_), N = Rf_asInteger(N_);
  const int draw_initial = Rf_asLogical(draw_initial_), eta_only = Rf_asLogical(eta_only_),
            tst = Rf_asLogical(transposed_state_);
  int p, zc, nZ, tr, tc, nT, rr, rR, nR, r, qc, nQ, mP, mPc, mPs;
  const int m = Rf_length(a);
  dims3(Z, &p, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &rR, &nR);
  dims3(Q, &r, &qc, &nQ);
  dims3(P_star, &mP, &mPc, &mPs);
  const int r_tot = r * repeat_Q;
  /* the variates Rcpp::rnorm would return, in the reference's call order (src/SimulateC.cpp:75,112) */
  const R_xlen_t n_draws = (draw_initial ? (R_xlen_t)m * nsim : 0) + (R_xlen_t)N * r_tot * nsim;
  double* draws = (double*)R_alloc(n_draws, sizeof(double));
  GetRNGstate();
  for (R_xlen_t i = 0; i < n_draws; i++) draws[i] = norm_rand();
  PutRNGstate();
  SEXP y = alloc3(N, p, nsim), a_out = alloc3(N, m, nsim), a_t = alloc3(m, nsim, N), eta = alloc3(N, r_tot, nsim);
  if (!tst || eta_only) memset(REAL(a_t), 0, sizeof(double) * (size_t)m * nsim * N);
  if (eta_only) {
    memset(REAL(y), 0, sizeof(double) * (size_t)N * p * nsim);
    memset(REAL(a_out), 0, sizeof(double) * (size_t)N * m * nsim);
  }
  int rc = ssgpu_SimulateC(nsim, repeat_Q, N, p, m, rR, r, mP, REAL(a), REAL(Z), nZ, REAL(T), nT, REAL(R), nR, REAL(Q),
                           nQ, REAL(P_star), draw_initial, eta_only, tst, draws, (long long)n_draws, NULL, NULL, REAL(y),
                           REAL(a_out), tst ? REAL(a_t) : NULL, REAL(eta));
  if (rc != SSGPU_OK) { UNPROTECT(4); chk(rc); }
  const char* names[] = {"y", "a", "a_t", "eta", ""}; /* src/SimulateC.cpp:133-138 */
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, y); SET_VECTOR_ELT(res, 1, a_out); SET_VECTOR_ELT(res, 2, a_t); SET_VECTOR_ELT(res, 3, eta);
  UNPROTECT(5);
  return res;
}

SEXP _statespacer_ExtractComponentC(SEXP a, SEXP Z) {
  need(a, REALSXP, "a"); need(Z, REALSXP, "Z");
  int m, nsim, N, p, zc, nZ;
  dims3(a, &m, &nsim, &N);
  dims3(Z, &p, &zc, &nZ);
  SEXP out = alloc3(N, p, nsim);
  int rc = ssgpu_ExtractComponentC(REAL(a), m, nsim, N, REAL(Z), p, nZ, REAL(out));
  if (rc != SSGPU_OK) { UNPROTECT(1); chk(rc); }
  UNPROTECT(1);
  return out;
}

/* New: E = B x V evaluations in one launch.  y: (N*p) x B matrix (time-major rows, one column per series:
 * R's column-major storage of t(y_b) stacked is NOT batch-contiguous, so the R glue passes t(Y) with Y the
 * B x (N*p) matrix -- see r/R/ssgpu_glue.R).  Q_diag, P_star_diag: r x (B*V) and m x (B*V) matrices of
 * per-evaluation diagonals (evaluation e = v*B + b), or NULL to use the shared Q / P_star. */
SEXP _statespacer_LogLikBatchC(SEXP yt, SEXP V_, SEXP a, SEXP P_inf, SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q,
                               SEXP Q_diag, SEXP P_star_diag) {
  need(yt, REALSXP, "Y"); need_model(a, P_inf, P_star, Z, T, R, Q); need(Q_diag, REALSXP, "Q_diag");
  need(P_star_diag, REALSXP, "P_star_diag");
  int B, Np, one, m, mc, ms, p, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ;
  const int V = Rf_asInteger(V_);
  dims3(yt, &B, &Np, &one); /* B x (N*p): element [b + B*q] = y_b[q]  == y[q*B + b] */
  dims3(P_star, &m, &mc, &ms);
  dims3(Z, &p, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  ssgpu_model mo;
  memset(&mo, 0, sizeof mo);
  mo.N = Np / p; mo.p = p; mo.m = m; mo.r = r;
  mo.a = (ssgpu_matrix){REAL(a), m, 1, 0, 1, 0, 0};
  mo.P_inf = (ssgpu_matrix){REAL(P_inf), m, m, 0, 1, 0, 0};
  mo.P_star = P_star_diag == R_NilValue ? (ssgpu_matrix){REAL(P_star), m, m, 0, 1, 0, 0}
                                        : (ssgpu_matrix){REAL(P_star_diag), m, m, 1, 1, m, (long long)m * B};
  mo.Z = (ssgpu_matrix){REAL(Z), p, m, 0, nZ, 0, 0};
  mo.T = (ssgpu_matrix){REAL(T), m, m, 0, nT, 0, 0};
  mo.R = (ssgpu_matrix){REAL(R), m, r, 0, nR, 0, 0};
  mo.Q = Q_diag == R_NilValue ? (ssgpu_matrix){REAL(Q), r, r, 0, nQ, 0, 0}
                              : (ssgpu_matrix){REAL(Q_diag), r, r, 1, 1, r, (long long)r * B};
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, B, V));
  int rc = ssgpu_loglik_batch(REAL(yt), B, V, &mo, SSGPU_KERNEL_AUTO, REAL(out));
  if (rc != SSGPU_OK) { UNPROTECT(1); chk(rc); }
  for (R_xlen_t i = 0; i < (R_xlen_t)B * V; i++)
    if (ISNAN(REAL(out)[i])) REAL(out)[i] = NA_REAL;
  UNPROTECT(1);
  return out;
}

/* New: objective and central-difference gradient for B series from parameter vectors (ssgpu.h section 6c).
 * yt: B x (N*p) as for LogLikBatchC; theta: k x B matrix (column b = theta of series b, i.e. [b*k + i]);
 * q_index / p_star_index: integer vectors (0-based parameter index per diagonal entry of Q / P_star, -1 = keep
 * the entry of the shared matrix).  Returns list(value = numeric(B), gradient = k x B matrix). */
SEXP _statespacer_LogLikGradBatchC(SEXP yt, SEXP theta, SEXP q_index, SEXP p_star_index, SEXP h_, SEXP a, SEXP P_inf,
                                   SEXP P_star, SEXP Z, SEXP T, SEXP R, SEXP Q) {
  need(yt, REALSXP, "Y"); need(theta, REALSXP, "theta"); need(q_index, INTSXP, "q_index");
  need(p_star_index, INTSXP, "p_star_index"); need_model(a, P_inf, P_star, Z, T, R, Q);
  int B, Np, one, m, mc, ms, p, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ, k, Bt, one2;
  dims3(yt, &B, &Np, &one);
  dims3(theta, &k, &Bt, &one2);
  dims3(P_star, &m, &mc, &ms);
  dims3(Z, &p, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  if (Bt != B || LENGTH(q_index) != r || LENGTH(p_star_index) != m) Rf_error("LogLikGradBatchC: inconsistent sizes");
  ssgpu_model mo;
  memset(&mo, 0, sizeof mo);
  mo.N = Np / p; mo.p = p; mo.m = m; mo.r = r;
  mo.a = (ssgpu_matrix){REAL(a), m, 1, 0, 1, 0, 0};
  mo.P_inf = (ssgpu_matrix){REAL(P_inf), m, m, 0, 1, 0, 0};
  mo.P_star = (ssgpu_matrix){REAL(P_star), m, m, 0, 1, 0, 0};
  mo.Z = (ssgpu_matrix){REAL(Z), p, m, 0, nZ, 0, 0};
  mo.T = (ssgpu_matrix){REAL(T), m, m, 0, nT, 0, 0};
  mo.R = (ssgpu_matrix){REAL(R), m, r, 0, nR, 0, 0};
  mo.Q = (ssgpu_matrix){REAL(Q), r, r, 0, nQ, 0, 0};
  SEXP value = PROTECT(Rf_allocVector(REALSXP, B));
  SEXP grad = PROTECT(Rf_allocMatrix(REALSXP, k, B));
  int rc = ssgpu_loglik_grad_batch(REAL(yt), B, &mo, REAL(theta), k, INTEGER(q_index), INTEGER(p_star_index),
                                   Rf_asReal(h_), SSGPU_KERNEL_AUTO, REAL(value), REAL(grad));
  if (rc != SSGPU_OK) { UNPROTECT(2); chk(rc); }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 2));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, 2));
  SET_STRING_ELT(names, 0, Rf_mkChar("value"));
  SET_STRING_ELT(names, 1, Rf_mkChar("gradient"));
  SET_VECTOR_ELT(out, 0, value);
  SET_VECTOR_ELT(out, 1, grad);
  Rf_setAttrib(out, R_NamesSymbol, names);
  UNPROTECT(4);
  return out;
}

/* element `name` of a named R list (R_NilValue if absent) */
static SEXP list_get(SEXP lst, const char* name) {
  SEXP names = Rf_getAttrib(lst, R_NamesSymbol);
  for (int i = 0; i < Rf_length(lst); i++)
    if (names != R_NilValue && strcmp(CHAR(STRING_ELT(names, i)), name) == 0) return VECTOR_ELT(lst, i);
  return R_NilValue;
}

/* New: SimSmoother() as ONE call (ssgpu.h section 6d; replaces the SimulateC / FastSmootherC / ExtractComponentC round
 * trips of R/SimSmoother.R:65-766).  comps: list of lists(a, Z, T, R, Q, P_star, repeat_Q, draw_initial) in the order
 * R/SimSmoother.R calls SimulateC; H: p x p x {1|N}; full: list(a, P_inf, P_star, Z, T, R, Q) of :579-615;
 * a_hat / eps_hat / eta_hat: object$smoothed$a / $epsilon / $eta; Z_extract: list of Z_padded arrays (:645-766).
 * The normal variates are drawn here with norm_rand() in exactly the order the reference's calls would consume them, so
 * R's random stream -- and therefore every draw -- is the one the CPU package produces for the same seed. */
SEXP _statespacer_SimSmootherC(SEXP nsim_, SEXP N_, SEXP p_, SEXP comps, SEXP H, SEXP full, SEXP steps_, SEXP a_hat,
                               SEXP eps_hat, SEXP eta_hat, SEXP Z_extract) {
  const int nsim = Rf_asInteger(nsim_), N = Rf_asInteger(N_), p = Rf_asInteger(p_), nc = Rf_length(comps),
            ne = Rf_length(Z_extract);
  ssgpu_sim_component* cs = (ssgpu_sim_component*)R_alloc(nc > 0 ? nc : 1, sizeof(ssgpu_sim_component));
  int m = 0, r = 0, d0, d1, d2;
  R_xlen_t n_draws = 0;
  for (int c = 0; c < nc; c++) {
    SEXP L = VECTOR_ELT(comps, c);
    SEXP a = list_get(L, "a"), Z = list_get(L, "Z"), T = list_get(L, "T"), R = list_get(L, "R"), Q = list_get(L, "Q"),
         P = list_get(L, "P_star");
    memset(&cs[c], 0, sizeof cs[c]);
    cs[c].m = Rf_length(a);
    cs[c].repeat_Q = Rf_asInteger(list_get(L, "repeat_Q"));
    cs[c].draw_initial = Rf_asLogical(list_get(L, "draw_initial"));
    cs[c].a = REAL(a); cs[c].Z = REAL(Z); cs[c].T = REAL(T); cs[c].R = REAL(R); cs[c].Q = REAL(Q); cs[c].P_star = REAL(P);
    dims3(Z, &d0, &d1, &cs[c].nZ);
    dims3(T, &d0, &d1, &cs[c].nT);
    dims3(R, &d0, &d1, &cs[c].nR);
    dims3(Q, &cs[c].r, &d1, &cs[c].nQ);
    m += cs[c].m;
    r += cs[c].r * cs[c].repeat_Q;
    n_draws += (cs[c].draw_initial ? (R_xlen_t)cs[c].m * nsim : 0) + (R_xlen_t)N * cs[c].r * cs[c].repeat_Q * nsim;
  }
  n_draws += (R_xlen_t)N * p * nsim;
  const int conditional = full != R_NilValue; /* NULL: predict()'s unconditional simulations (ssgpu_SimulateModel) */
  ssgpu_sim_model fm;
  memset(&fm, 0, sizeof fm);
  if (conditional) {
  fm.a = REAL(list_get(full, "a")); fm.P_inf = REAL(list_get(full, "P_inf")); fm.P_star = REAL(list_get(full, "P_star"));
  fm.Z = REAL(list_get(full, "Z")); fm.T = REAL(list_get(full, "T")); fm.R = REAL(list_get(full, "R"));
  fm.Q = REAL(list_get(full, "Q"));
  dims3(list_get(full, "Z"), &d0, &fm.m, &fm.nZ);
  dims3(list_get(full, "T"), &d0, &d1, &fm.nT);
  dims3(list_get(full, "R"), &d0, &fm.r, &fm.nR);
  dims3(list_get(full, "Q"), &d0, &d1, &fm.nQ);
  }
  dims3(H, &d0, &d1, &d2);
  const int nH = d2;
  double* draws = (double*)R_alloc(n_draws, sizeof(double));
  GetRNGstate();
  for (R_xlen_t i = 0; i < n_draws; i++) draws[i] = norm_rand();
  PutRNGstate();
  SEXP eps = alloc3(N, p, nsim), a_out = alloc3(N, m, nsim), eta = alloc3(N, r, nsim);
  SEXP ysim = alloc3(N, p, conditional ? 0 : nsim);
  SEXP ext = PROTECT(Rf_allocVector(VECSXP, ne));
  const double** zp = (const double**)R_alloc(ne > 0 ? ne : 1, sizeof(double*));
  double** op = (double**)R_alloc(ne > 0 ? ne : 1, sizeof(double*));
  int* nz = (int*)R_alloc(ne > 0 ? ne : 1, sizeof(int));
  for (int e = 0; e < ne; e++) {
    SEXP Z = VECTOR_ELT(Z_extract, e);
    SEXP o = Rf_alloc3DArray(REALSXP, N, p, nsim);
    SET_VECTOR_ELT(ext, e, o); /* protected through ext */
    zp[e] = REAL(Z);
    op[e] = REAL(o);
    dims3(Z, &d0, &d1, &nz[e]);
  }
  int rc = conditional
               ? ssgpu_SimSmoother(nsim, N, p, nc, cs, REAL(H), nH, &fm, Rf_asInteger(steps_), REAL(a_hat), REAL(eps_hat),
                                   REAL(eta_hat), draws, (long long)n_draws, REAL(eps), REAL(a_out), REAL(eta), ne, zp, nz, op)
               : ssgpu_SimulateModel(nsim, N, p, nc, cs, REAL(H), nH, draws, (long long)n_draws, REAL(ysim), REAL(eps),
                                     REAL(a_out), REAL(eta), ne, zp, nz, op);
  if (rc != SSGPU_OK) { UNPROTECT(5); chk(rc); }
  const char* names[] = {"epsilon", "a", "eta", "components", "y", ""};
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, eps); SET_VECTOR_ELT(res, 1, a_out); SET_VECTOR_ELT(res, 2, eta); SET_VECTOR_ELT(res, 3, ext);
  SET_VECTOR_ELT(res, 4, ysim);
  UNPROTECT(6);
  return res;
}

/* New: the collapse transform of R/statespacer.R:815-900 (y N x p with NA, H p x p [x N], Z p x m2 [x N]) ->
 * list(y = N x m2 with NA, H_star = m2 x m2 x {1|N}, loglik_add). */
SEXP _statespacer_CollapseC(SEXP y, SEXP H, SEXP Z) {
  int N, p, one, hr, hc, nH, zr, m2, nZ;
  dims3(y, &N, &p, &one);
  dims3(H, &hr, &hc, &nH);
  dims3(Z, &zr, &m2, &nZ);
  const int ns = (nH > 1 || nZ > 1) ? N : 1;
  double* yy = (double*)R_alloc((size_t)N * p, sizeof(double));
  for (R_xlen_t i = 0; i < (R_xlen_t)N * p; i++) yy[i] = ISNAN(REAL(y)[i]) ? R_NaN : REAL(y)[i];
  SEXP ys = PROTECT(Rf_allocMatrix(REALSXP, N, m2));
  SEXP Hs = alloc3(m2, m2, ns);
  double la = 0.0;
  int rc = ssgpu_collapse(yy, N, p, m2, REAL(H), nH, REAL(Z), nZ, REAL(ys), REAL(Hs), &la);
  if (rc != SSGPU_OK) { UNPROTECT(2); chk(rc); }
  for (R_xlen_t i = 0; i < (R_xlen_t)N * m2; i++)
    if (ISNAN(REAL(ys)[i])) REAL(ys)[i] = NA_REAL;
  const char* names[] = {"y", "H_star", "loglik_add", ""};
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, ys); SET_VECTOR_ELT(res, 1, Hs); SET_VECTOR_ELT(res, 2, Rf_ScalarReal(la));
  UNPROTECT(3);
  return res;
}

/* New: the forecast recursion of R/predict.R:205-246 -> list(y_fc, a_fc, P_fc, Fmat_fc). */
SEXP _statespacer_ForecastC(SEXP h_, SEXP a1, SEXP P1, SEXP Z, SEXP T, SEXP R, SEXP Q, SEXP H) {
  const int h = Rf_asInteger(h_), m = Rf_length(a1);
  int p, zc, nZ, tr, tc, nT, rr, r, nR, qr, qc, nQ, hr, hc, nH;
  dims3(Z, &p, &zc, &nZ);
  dims3(T, &tr, &tc, &nT);
  dims3(R, &rr, &r, &nR);
  dims3(Q, &qr, &qc, &nQ);
  dims3(H, &hr, &hc, &nH);
  SEXP y_fc = PROTECT(Rf_allocMatrix(REALSXP, h, p));
  SEXP a_fc = PROTECT(Rf_allocMatrix(REALSXP, h, m));
  SEXP P_fc = alloc3(m, m, h), F_fc = alloc3(p, p, h);
  int rc = ssgpu_forecast(h, p, m, r, REAL(a1), REAL(P1), REAL(Z), nZ, REAL(T), nT, REAL(R), nR, REAL(Q), nQ, REAL(H), nH,
                          REAL(y_fc), REAL(a_fc), REAL(P_fc), REAL(F_fc));
  if (rc != SSGPU_OK) { UNPROTECT(4); chk(rc); }
  const char* names[] = {"y_fc", "a_fc", "P_fc", "Fmat_fc", ""};
  SEXP res = PROTECT(Rf_mkNamed(VECSXP, names));
  SET_VECTOR_ELT(res, 0, y_fc); SET_VECTOR_ELT(res, 1, a_fc); SET_VECTOR_ELT(res, 2, P_fc); SET_VECTOR_ELT(res, 3, F_fc);
  UNPROTECT(5);
  return res;
}

/* New: shard the batched objective over n GPUs of this R process (n <= 0: all visible); returns the count in use. */
SEXP _statespacer_UseDevices(SEXP n_) {
  int n = 0;
  chk(ssgpu_use_devices(Rf_asInteger(n_)));
  chk(ssgpu_devices_in_use(&n));
  return Rf_ScalarInteger(n);
}

static const R_CallMethodDef CallEntries[] = {
    {"_statespacer_ExtractComponentC", (DL_FUNC)&_statespacer_ExtractComponentC, 2},
    {"_statespacer_FastSmootherC", (DL_FUNC)&_statespacer_FastSmootherC, 10},
    {"_statespacer_KalmanC", (DL_FUNC)&_statespacer_KalmanC, 10},
    {"_statespacer_LogLikC", (DL_FUNC)&_statespacer_LogLikC, 9},
    {"_statespacer_SimulateC", (DL_FUNC)&_statespacer_SimulateC, 12},
    {"_statespacer_LogLikBatchC", (DL_FUNC)&_statespacer_LogLikBatchC, 11},
    {"_statespacer_LogLikGradBatchC", (DL_FUNC)&_statespacer_LogLikGradBatchC, 12},
    {"_statespacer_SimSmootherC", (DL_FUNC)&_statespacer_SimSmootherC, 11},
    {"_statespacer_CollapseC", (DL_FUNC)&_statespacer_CollapseC, 3},
    {"_statespacer_ForecastC", (DL_FUNC)&_statespacer_ForecastC, 8},
    {"_statespacer_UseDevices", (DL_FUNC)&_statespacer_UseDevices, 1},
    {NULL, NULL, 0}};

void R_init_statespacer(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
