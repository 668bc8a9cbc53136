/* ssgpu.h -- C ABI of libssgpu.so: the B200 (sm_100a) implementation of statespacer's Kalman hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain pointers and sizes (no R, Rcpp,
 * Armadillo or torch types) and returns an int status (0 = ok, see SSGPU_E_*); the message of the
 * last failure on the calling thread is available from ssgpu_last_error().  Numerical degeneracy is
 * NOT an error: like the reference, LogLikC yields NaN (R's NA_real_) and KalmanC writes NaN into
 * loglik[index] (src/LogLikC.cpp:210-213, src/KalmanC.cpp:194,239,335).
 *
 * Layouts follow R / Armadillo exactly for the five legacy entry points (column-major; 3-D system
 * matrices are rows x cols x n_slices with n_slices in {1, N}, "time-varying" iff n_slices > 1,
 * src/LogLikC.cpp:37-38).  The batched entry points define their own batch-contiguous layout.
 *
 * Reference interface replaced (all in /root/reference):
 *   ssgpu_LogLikC            <- _statespacer_LogLikC            src/RcppExports.cpp:68-84,  src/LogLikC.cpp:20-28
 *   ssgpu_KalmanC            <- _statespacer_KalmanC            src/RcppExports.cpp:48-65,  src/KalmanC.cpp:21-30
 *   ssgpu_FastSmootherC      <- _statespacer_FastSmootherC      src/RcppExports.cpp:28-45,  src/FastSmootherC.cpp:23-32
 *   ssgpu_SimulateC          <- _statespacer_SimulateC          src/RcppExports.cpp:87-106, src/SimulateC.cpp:26-37
 *   ssgpu_ExtractComponentC  <- _statespacer_ExtractComponentC  src/RcppExports.cpp:16-25,  src/ExtractComponentC.cpp:12-13
 *   ssgpu_loglik_batch[_dev] <- (new) the sequential LogLikC calls of the optim loop, R/statespacer.R:917-927,977-982
 *
 * Threading: entry points may be called from any host thread but serialise on one internal CUDA
 * stream per device; every host-pointer entry point synchronises before returning (R's .Call
 * contract).  There is no CPU fallback: without a CUDA device every compute entry
 * point returns SSGPU_E_CUDA.
 */
#ifndef SSGPU_H
#define SSGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define SSGPU_OK 0
#define SSGPU_E_ARG 1     /* inconsistent sizes / null pointers / unsupported shape (m > SSGPU_MAX_M ...) */
#define SSGPU_E_CUDA 2    /* CUDA runtime failure (no device, out of memory, launch failure) */
#define SSGPU_E_DRAWS 3   /* injected normal variates: wrong count */

#define SSGPU_MAX_M 64    /* largest supported state dimension (incl. the p residual states) */

/* Library / device management -------------------------------------------------------------- */
int ssgpu_version(void);                     /* 100*major + minor */
const char* ssgpu_last_error(void);          /* thread-local, never NULL */
int ssgpu_set_device(int device);            /* device used by subsequent calls of this process */
int ssgpu_device_count(int* count);
/* Multi-GPU from ONE process (R is one process: R/statespacer.R:977-982 calls the objective from the main R thread).
 * n devices starting at the one of ssgpu_set_device (n <= 0: all visible).  The host-pointer batch entry point
 * ssgpu_loglik_grad_batch then shards its series over them in contiguous ranges -- one worker thread, stream and pinned
 * staging ring per device; every device writes its series' results straight into the caller's arrays (the recursion has
 * no exchange step: nothing is gathered, no collective is needed).  Results are bit-identical to the single-device call.
 * ssgpu_SimSmoother / ssgpu_SimulateModel shard their draws the same way (the variates of a range are gathered on the
 * host, the outputs of a range are a contiguous block of R's arrays).  The `_dev` entry points stay on the device of
 * ssgpu_set_device. */
int ssgpu_use_devices(int n);
int ssgpu_devices_in_use(int* n);
int ssgpu_release(void);                     /* free cached device workspaces and pinned staging */
/* number of kernels this library launched since process start (bench.py's gpu_launches) */
long long ssgpu_launch_count(void);

/* ---- 1. LogLikC ---------------------------------------------------------------------------
 * y N x p (NaN where missing), y_isna N x p (int, non-zero = missing; may be NULL -> derived from
 * isnan(y)), a[m], P_inf[m*m], P_star[m*m], Z p x m x nZ, T m x m x nT, R m x r x nR, Q r x r x nQ.
 * *loglik = loglik / N, or NaN if F_star < 1e-7 at every post-initialisation step. */
int ssgpu_LogLikC(const double* y, const int* y_isna, int N, int p, int m, int r,
                  const double* a, const double* P_inf, const double* P_star,
                  const double* Z, int nZ, const double* T, int nT,
                  const double* R, int nR, const double* Q, int nQ, double* loglik);

/* ---- 2. KalmanC ---------------------------------------------------------------------------
 * Outputs mirror the list returned at src/KalmanC.cpp:563-593; any pointer may be NULL (skipped).
 * Shapes (column-major): loglik[N*p] (index = t*p + j); a_pred,a_fil,a_smooth N x m;
 * P_pred,P_fil,V,P_inf_pred,P_star_pred,P_inf_fil,P_star_fil m x m x N; yfit,v,v_norm,e N x p;
 * eta N x r; Fmat,D p x p x N; eta_var r x r x N; a_fc[m]; P_inf_fc,P_star_fc,P_fc m x m;
 * Tstat_observation N x p; Tstat_state (N+1) x m (row 0 is not written, as in the reference);
 * r_vec (N+1) x m; Nmat m x m x (N+1). */
typedef struct ssgpu_kalman_out {
  int* initialisation_steps;
  double *loglik, *a_pred, *a_fil, *a_smooth, *P_pred, *P_fil, *V, *P_inf_pred, *P_star_pred,
      *P_inf_fil, *P_star_fil, *yfit, *v, *v_norm, *eta, *e, *Fmat, *eta_var, *D, *a_fc,
      *P_inf_fc, *P_star_fc, *P_fc, *Tstat_observation, *Tstat_state, *r_vec, *Nmat;
} ssgpu_kalman_out;

int ssgpu_KalmanC(const double* y, const int* y_isna, int N, int p, int m, int r,
                  const double* a, const double* P_inf, const double* P_star,
                  const double* Z, int nZ, const double* T, int nT,
                  const double* R, int nR, const double* Q, int nQ,
                  int diagnostics, ssgpu_kalman_out* out);

/* ---- 3. FastSmootherC ---------------------------------------------------------------------
 * y N x p x nsim (no missing values, src/FastSmootherC.cpp:23-32).  a_smooth N x m x nsim,
 * a_t m x nsim x N (written only if transposed_state; may be NULL otherwise), eta N x r x nsim
 * (row N-1 = 0). */
int ssgpu_FastSmootherC(const double* y, int N, int p, int nsim, int m, int r,
                        const double* a, const double* P_inf, const double* P_star,
                        const double* Z, int nZ, const double* T, int nT,
                        const double* R, int nR, const double* Q, int nQ,
                        int initialisation_steps, int transposed_state,
                        double* a_smooth, double* a_t, double* eta);

/* ---- 4. SimulateC -------------------------------------------------------------------------
 * Z p x m x nZ, T m x m x nT, R m x rR x nR, Q r x r x nQ (block-repeated repeat_Q times,
 * r_tot = r*repeat_Q), P_star mP x mP (read only if draw_initial).  With eta_only the caller passes
 * 1 x 1 x 1 dummies for Z, T, R exactly like R/SimSmoother.R:559-573 and only `eta` is written.
 * `draws` are the standard normal variates the reference would obtain from Rcpp::rnorm, in its call
 * order (src/SimulateC.cpp:75,112): [m*nsim initial draws if draw_initial] then, for every time
 * point, r_tot*nsim draws (r_tot fastest); n_draws must equal that count exactly.
 * Outputs: y N x p x nsim, a N x m x nsim, a_t m x nsim x N (only if transposed_state),
 * eta N x r_tot x nsim.  Any output pointer may be NULL.
 * Q_root / P_root: optional precomputed symmetric roots U sqrt(S) U' (r x r x nQ, mP x mP); when
 * NULL they are computed on the host by a Jacobi eigen-decomposition (src/SimulateC.cpp:73-74,84-85). */
int ssgpu_SimulateC(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP,
                    const double* a, const double* Z, int nZ, const double* T, int nT,
                    const double* R, int nR, const double* Q, int nQ, const double* P_star,
                    int draw_initial, int eta_only, int transposed_state,
                    const double* draws, long long n_draws,
                    const double* Q_root, const double* P_root,
                    double* y, double* a_out, double* a_t, double* eta);

/* U diag(sqrt(|lambda|)) U' of a symmetric n x n matrix (host, cyclic Jacobi): the root SimulateC takes of
 * Q and P_star when Q_root / P_root are NULL. */
int ssgpu_sym_root(const double* A, int n, double* out);

/* ---- 5. ExtractComponentC -----------------------------------------------------------------
 * a m x nsim x N, Z p x m x nZ  ->  out N x p x nsim. */
int ssgpu_ExtractComponentC(const double* a, int m, int nsim, int N,
                            const double* Z, int p, int nZ, double* out);

/* ---- 6. Batched loglikelihood (new) -------------------------------------------------------
 * E = B * V evaluations: B series x V parameter vectors ("variants": finite-difference points,
 * multistart candidates); evaluation e = v * B + b.
 * y is time-major and batch-contiguous: y[(t*p + j) * B + b], NaN = missing.
 * Every system matrix is described by an ssgpu_matrix: dense column-major (rows x cols) or, with
 * diag != 0, only its diagonal (rows doubles per slice); n_slices in {1, N}; the element offset of
 * the block used by evaluation (b, v) is b*stride_series + v*stride_variant (0/0 = shared by the
 * whole batch), slices of one block are contiguous.
 * out[e] = loglik/N or NaN, exactly as ssgpu_LogLikC would return for that evaluation. */
typedef struct ssgpu_matrix {
  const double* data;
  int rows, cols;
  int diag;
  int n_slices;
  long long stride_series;
  long long stride_variant;
} ssgpu_matrix;

typedef struct ssgpu_model {
  int N, p, m, r;
  ssgpu_matrix a;       /* m x 1 */
  ssgpu_matrix P_inf;   /* m x m */
  ssgpu_matrix P_star;  /* m x m */
  ssgpu_matrix Z;       /* p x m */
  ssgpu_matrix T;       /* m x m */
  ssgpu_matrix R;       /* m x r */
  ssgpu_matrix Q;       /* r x r */
} ssgpu_model;

/* kernel selection for the batched path */
#define SSGPU_KERNEL_AUTO 0     /* block-row kernel when the model qualifies (the core kernel, below, first when p >= 2
                                   and the batch is large), else the core kernel, else the tensor kernel (column
                                   kernel first when p >= 4 and it qualifies), else column kernel, else the
                                   warp/shared-memory kernel, else generic */
#define SSGPU_KERNEL_GENERIC 1  /* one CTA per evaluation, any shape */
#define SSGPU_KERNEL_COLS 2     /* column-per-lane DFMA kernel (m <= 15); SSGPU_E_ARG if not eligible */
#define SSGPU_KERNEL_MMA 3      /* warp-per-evaluation FP64 tensor (DMMA) kernel (m <= 23); SSGPU_E_ARG if not eligible.
                                   States are re-ordered internally so that the all-zero 8 x 8 tiles of a block-
                                   diagonal T are skipped (results equal up to summation order) */
#define SSGPU_KERNEL_MMA_DENSE 4 /* same kernel, caller's state order, every tile multiplied */
#define SSGPU_KERNEL_WSM 5      /* warp per evaluation, covariance in shared memory, FP64 tensor time update (m <= 63,
                                   shared Z / T / R, Z may be time-varying); SSGPU_E_ARG if not eligible */

#define SSGPU_KERNEL_BLK 6      /* block-row DFMA kernel: T block diagonal with blocks of at most 2 x 2 after re-ordering
                                   the states (residuals, level, slope, trigonometric seasonals, cycles, regression
                                   coefficients), R Q R' inside the blocks; at most 8 blocks (m <= 16) with p <= 8, or
                                   up to 16 blocks (m <= 32) with p = 1 and a diagonal R Q R'; Z may be time-varying
                                   when p = 1 (explanatory variables); SSGPU_E_ARG if not eligible.  AUTO tries it
                                   first.  With 2 to 8 blocks, p = 1 and a time-invariant Z (BASELINE.json config 5)
                                   the lanes run the exact diffuse phase only and a second kernel, one thread per
                                   evaluation, runs the regular phase */
#define SSGPU_KERNEL_BLK_LANES 7 /* the same model class on the lane kernel alone (second implementation the tests
                                   cross-check; A/B runs) */

#define SSGPU_KERNEL_CORE 8     /* one thread per series on the component states alone: the p residual states every
                                   reference model starts with (Z = [I_p | Z_c], T = blkdiag(0_p, T_c), R Q R' =
                                   blkdiag(H, W_c), H diagonal; R/GetSysMat.R:1388-1426) are eliminated exactly, the
                                   dense covariance of the remaining states lives in registers (p <= 16, m - p <= 16
                                   of which at most 8 carry a variance: component states with a known start, no
                                   disturbance and no stochastic driver are tabulated on the host; everything shared
                                   by the batch and time-invariant; exact diffuse phase included).  BASELINE.json
                                   config 3 (dynamic Nelson-Siegel, p = 8, m = 14) is a 3 x 3 problem here.  SSGPU_E_ARG if not eligible.  AUTO takes it for batches of 2048 evaluations
                                   or more: before the block-row kernel when p >= 2 or m = 2 (local level), behind it
                                   otherwise */

/* host pointers everywhere (y, the matrices' data, out); copies in, runs, copies out, synchronises */
int ssgpu_loglik_batch(const double* y, long long B, int V, const ssgpu_model* model, int kernel,
                       double* out);
/* device pointers everywhere; asynchronous on `stream` (a cudaStream_t, NULL = the library's
 * stream); `out` may be read after the stream is synchronised. */
int ssgpu_loglik_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, int kernel,
                           double* out, void* stream);
/* name of the kernel the last ssgpu_loglik_batch[_dev] call on this thread launched */
const char* ssgpu_last_kernel(void);

/* ---- 2b. Batched KalmanC (new) ---------------------------------------------------------------
 * Filter and smoother of B independent series in one call (StateSpaceEval() over a fitted batch, R/StateSpaceEval.R:156-167
 * once per series in the reference).  y is time-major and batch-contiguous as in section 6, y[(t*p + j)*B + b]; `model`
 * as in section 6 with V = 1 (per-series a / P_inf / P_star / Q / ... through stride_series; shared matrices with
 * stride 0).  Every non-NULL member of `out` receives B consecutive blocks, block b holding series b in exactly the
 * layout of ssgpu_KalmanC (initialisation_steps: B ints).  The diagnostics arrays (v_norm, e, D, Tstat_*) must be NULL.
 * Only what is asked for crosses to the caller; P_star_pred, P_inf_pred and a_pred are kept as device scratch for
 * the backward pass when they are not (2 m^2 N + N m doubles per series: size the batch accordingly).
 * ssgpu_kalman_batch: host pointers everywhere; ssgpu_kalman_batch_dev: device pointers, work enqueued on `stream`,
 * returns after completion. */
int ssgpu_kalman_batch(const double* y, long long B, const ssgpu_model* model, const ssgpu_kalman_out* out);
int ssgpu_kalman_batch_dev(const double* y, long long B, const ssgpu_model* model, const ssgpu_kalman_out* out,
                           void* stream);

/* ---- 6c. Batched loglikelihood and gradient from parameter vectors (new) ---------------------
 * The objective of the optim loop (R/statespacer.R:917-927) rebuilds the system matrices from theta in R for
 * every call, and stats::optim takes 2k further calls per gradient (R/statespacer.R:977-982).  For models whose
 * parameters enter as variances exp(2 theta_i) on the diagonals of Q and P_star (the default univariate
 * parameterisation, R/Cholesky.R:148-167 -- local level, slope, trigonometric seasonal) the matrices are built on
 * the device from theta itself: q_index[r] / p_star_index[m] give, per diagonal entry, the parameter it is the
 * variance of, or -1 to keep the entry of the model's (shared, diagonal) Q / P_star.  a, P_inf, Z, T, R are the
 * shared matrices of `model`.
 *   ssgpu_loglik_grad_batch      host pointers; theta [b*k + i]; loglik[b] and the central-difference gradient
 *                                grad[b*k + i] = (l(theta + h e_i) - l(theta - h e_i)) / 2h from ONE launch of
 *                                1 + 2k evaluations per series (h = 1e-3 is stats::optim's ndeps default); with y in
 *                                pinned host memory and B >= 16384 the upload is pipelined with the kernels over
 *                                8 chunks of series, pageable y goes through the library's pinned staging ring
 *   ssgpu_loglik_grad_batch_dev  the same on device pointers (y, theta, loglik, grad), asynchronous on `stream`
 *   ssgpu_loglik_theta_batch_dev V arbitrary parameter vectors per series, theta [(v*B + b)*k + i] -> out[v*B + b]
 *                                (multistart candidates, the points of a numDeriv Hessian sweep) */
int ssgpu_loglik_grad_batch(const double* y, long long B, const ssgpu_model* model, const double* theta, int k,
                            const int* q_index, const int* p_star_index, double h, int kernel,
                            double* loglik, double* grad);
int ssgpu_loglik_grad_batch_dev(const double* y, long long B, const ssgpu_model* model, const double* theta, int k,
                                const int* q_index, const int* p_star_index, double h, int kernel,
                                double* loglik, double* grad, void* stream);
int ssgpu_loglik_theta_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, const double* theta,
                                 int k, const int* q_index, const int* p_star_index, int kernel, double* out,
                                 void* stream);

/* ---- 6d. ... with full covariance blocks: the LDL' parameterisation on the device (new) ------------------------------
 * Multivariate models parameterise every variance matrix as L D L' (R/Cholesky.R:148-167): D = diag(exp(2 theta_i)), L
 * unit lower triangular with its strict lower triangle filled column by column from the following parameters.  A block
 * descriptor places such an n x n matrix (n <= 8) on the diagonal of Q (target 0) or P_star (target 1) at row = column
 * `offset`, built from theta[theta_offset ...] (n + n(n-1)/2 parameters); n = 1 is the plain exp(2 theta) of section 6c.
 * Entries of the model's (shared, dense) Q / P_star outside the blocks are kept.  A parameter vector with a diagonal of
 * D <= 1e-7 is not admissible: the reference returns NA (R/Cholesky.R:159-161), this entry point NaN.
 *   fd != 0: theta [b*k + i]; out[b] = loglik / N, grad[b*k + i] central differences with step h (V is ignored)
 *   fd == 0: theta [(v*B + b)*k + i] -> out[v*B + b]; grad is not touched
 * Device pointers (y, theta, out, grad), asynchronous on `stream`; a, P_inf, Z, T, R, Q, P_star of `model` shared. */
typedef struct {
  int target;       /* 0 = Q, 1 = P_star */
  int offset;       /* first row = first column of the block */
  int n;            /* 1 .. 8 */
  int theta_offset; /* first of its n + n(n-1)/2 parameters */
} ssgpu_cov_block;
int ssgpu_loglik_cov_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, const double* theta, int k,
                               const ssgpu_cov_block* blocks, int n_blocks, double h, int fd, int kernel, double* out,
                               double* grad, void* stream);

/* ---- 6b. Simulation smoother on device-resident draws (new) ---------------------------------
 * The three per-draw routines again, for callers that keep the draws on the GPU between the steps of
 * R/SimSmoother.R:91-769 (simulate every component -> add -> smooth -> extract) instead of round-tripping
 * N x m x nsim arrays through host memory.  Per-draw arrays are DEVICE pointers in the draw-contiguous layout the
 * kernels use, X[(t*K + k)*nsim + s] (K = p, m, r or r_tot); `draws` is a device pointer in the reference's
 * consumption order (as for ssgpu_SimulateC); system matrices stay HOST pointers (a few KB, copied before the call
 * returns).  Work is enqueued on `stream` (a cudaStream_t, NULL = the library's stream); the calls return after
 * the kernels have completed.  Arithmetic is identical to the host-pointer entry points (bit-exact).
 * ssgpu_draw_layout_dev converts between the draw-contiguous layout and R's N x K x nsim array on the device
 * (to_r != 0: draw-contiguous -> R). */
int ssgpu_SimulateC_dev(int nsim, int repeat_Q, int N, int p, int m, int rR, int r, int mP,
                        const double* a, const double* Z, int nZ, const double* T, int nT,
                        const double* R, int nR, const double* Q, int nQ, const double* P_star,
                        int draw_initial, int eta_only, const double* draws, long long n_draws,
                        const double* Q_root, const double* P_root,
                        double* y, double* a_out, double* eta, void* stream);
int ssgpu_FastSmootherC_dev(const double* y, int N, int p, int nsim, int m, int r,
                            const double* a, const double* P_inf, const double* P_star,
                            const double* Z, int nZ, const double* T, int nT,
                            const double* R, int nR, const double* Q, int nQ,
                            int initialisation_steps, double* a_smooth, double* eta, void* stream);
int ssgpu_ExtractComponentC_dev(const double* a, int m, int nsim, int N,
                                const double* Z, int p, int nZ, double* out, void* stream);
int ssgpu_draw_layout_dev(const double* src, double* dst, int N, int K, long long nsim, int to_r, void* stream);
/* device time (ms, CUDA events on the launching stream) of the last simulate / per-draw smoother / extract kernel
 * this process launched; -1 if none yet.  Waits for that kernel. */
int ssgpu_last_sim_kernel_ms(double* simulate_ms, double* fast_smoother_ms, double* extract_ms);

/* ---- 6d. SimSmoother as one call (new) ---------------------------------------------------------
 * R/SimSmoother.R:34-769 draws every component of the model unconditionally with SimulateC, adds the draws up to
 * simulated data, smooths those with FastSmootherC, corrects the draws by the difference to the smoothed estimates of
 * the observed data and extracts the components -- five .Call round trips per component through N x m x nsim host
 * arrays.  Here the whole function is one call: only the normal variates go in and only what SimSmoother() returns
 * comes out; the per-draw arrays never leave the device.
 *   comps[c]        one SimulateC call of R/SimSmoother.R:65-553 (arguments as for ssgpu_SimulateC; R has
 *                   r*repeat_Q columns); their states / disturbances are stacked in this order
 *   H               p x p x nH residual variance (the Q of the eta_only call, :556-573)
 *   full            the model FastSmootherC runs on (:579-632): residual states first, m = p + sum m_c, r = p + sum r_c
 *   a_hat N x m, eps_hat N x p, eta_hat N x r   object$smoothed$a / $epsilon / $eta (column-major)
 *   draws           standard normal variates in the order the reference consumes them: per component
 *                   [m_c*nsim if draw_initial] then N*r_tot_c*nsim, and N*p*nsim for the residuals last
 *   epsilon N x p x nsim, a N x m x nsim, eta N x r x nsim   results (any may be NULL)
 *   Z_extract[e] p x m x nZ_extract[e]  ->  extracted[e] N x p x nsim   (the Z_padded matrices of :645-766)
 * Arithmetic is that of the three routines (bit-identical to calling them in R's order). */
typedef struct ssgpu_sim_component {
  int m, r, repeat_Q, draw_initial;
  const double *a, *Z, *T, *R, *Q, *P_star;   /* Z p x m x nZ, T m x m x nT, R m x (r*repeat_Q) x nR, Q r x r x nQ, P_star m x m */
  int nZ, nT, nR, nQ;
  const double *Q_root, *P_root;              /* optional precomputed symmetric roots (NULL: computed on the host) */
} ssgpu_sim_component;

typedef struct ssgpu_sim_model {
  int m, r;
  const double *a, *P_inf, *P_star, *Z, *T, *R, *Q;
  int nZ, nT, nR, nQ;
} ssgpu_sim_model;

int ssgpu_SimSmoother(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps,
                      const double* H, int nH, const ssgpu_sim_model* full, int initialisation_steps,
                      const double* a_hat, const double* eps_hat, const double* eta_hat,
                      const double* draws, long long n_draws,
                      double* epsilon, double* a_out, double* eta,
                      int n_extract, const double* const* Z_extract, const int* nZ_extract, double* const* extracted);

/* The unconditional half of the above as its own call: predict()'s simulations over the forecast period
 * (R/predict.R:142-856: one SimulateC call per component started from object$predicted$a_fc / P_fc with
 * draw_initial = TRUE, the residual draws last; y_sim, a_sim, eta_sim stacked exactly like SimSmoother()).  N is the
 * forecast period.  y N x p x nsim (sum of the components' y and the residuals), epsilon, a, eta as above (no mean
 * correction); extracted[e] = Z_extract[e] a per draw -- with a component's padded Z this is that component's own
 * simulated y (sim_list$level, ...). */
int ssgpu_SimulateModel(int nsim, int N, int p, int n_comp, const ssgpu_sim_component* comps,
                        const double* H, int nH, const double* draws, long long n_draws,
                        double* y, double* epsilon, double* a_out, double* eta,
                        int n_extract, const double* const* Z_extract, const int* nZ_extract, double* const* extracted);

/* ---- 6e. The steps on either side of the path (new) ------------------------------------
 * ssgpu_collapse: the collapse transform of R/statespacer.R:815-900, done by R inside the optimiser's objective for
 * every parameter vector of a `collapse = TRUE` model.  y N x p (column-major, NaN = missing, entered as 0 like
 * :407-409), H p x p x nH, Z p x m2 x nZ (the columns of Z_kal that are collapsed onto), n* in {1, N}, m2 < p <= 16.
 * y_star N x m2 (NaN where R writes NA: exact zeros), H_star m2 x m2 x max(nH, nZ), *loglik_add as at :823-832.
 * ssgpu_forecast: the forecast recursion of R/predict.R:205-246 from KalmanC's a_fc / P_fc: y_fc h x p, a_fc h x m,
 * P_fc m x m x h, Fmat_fc p x p x h; Z p x m, T m x m, R m x r, Q r x r, H p x p with 1 or h slices each. */
int ssgpu_collapse(const double* y, int N, int p, int m2, const double* H, int nH, const double* Z, int nZ,
                   double* y_star, double* H_star, double* loglik_add);
int ssgpu_forecast(int h, int p, int m, int r, const double* a1, const double* P1, const double* Z, int nZ,
                   const double* T, int nT, const double* R, int nR, const double* Q, int nQ, const double* H, int nH,
                   double* y_fc, double* a_fc, double* P_fc, double* Fmat_fc);

/* Host-only helper (no device needed): the internal state order the tensor kernel would use for a shared
 * m x m transition matrix T (column-major, m <= 23).  perm[i] = caller's index of the state at internal
 * position i; *tile_mask has bit rt*TT + ct set when the 8 x 8 tile (rt, ct) of the re-ordered T holds a
 * non-zero (TT = ceil((m+1)/8)); *dmma_per_step = DMMA instructions one time update then issues.  Any output
 * pointer may be NULL. */
int ssgpu_plan_state_order(const double* T, int m, int* perm, unsigned* tile_mask, int* dmma_per_step);
/* The same for the block-row kernel (SSGPU_KERNEL_BLK): T m x m, R m x r, Q r x r or NULL (= diagonal), all
 * column-major.  *n_blocks = number of 2 x 2 blocks, or 0 when the model does not decompose (a state coupled with two
 * others through T or R Q R', more than 16 blocks, p > 8); perm[32]: internal position 2I+e -> state index of the
 * caller, -1 = padding; *lanes = lanes per evaluation of the lane kernel (the diffuse phase; with 2 to 8 blocks and p = 1 the regular phase runs one thread per evaluation); *flop_per_step = FP64 operations the kernel of the regular phase issues per series
 * and time step (fma = 2, all lanes). */
int ssgpu_plan_blocks(const double* T, const double* R, const double* Q, int m, int r, int p, int* perm,
                      int* n_blocks, int* lanes, int* flop_per_step);
/* Host-only helper (no device needed): what the thread-per-evaluation kernel behind SSGPU_KERNEL_BLK does with a
 * univariate model (p = 1; Z a row of m, everything column-major and shared by the batch, Q NULL = diagonal):
 * *n_blocks as ssgpu_plan_blocks (0: the model does not decompose; the other outputs are 0 then, and when the block
 * count is outside 2 .. 8); *residual_dropped, *level_slope_by_sums, *zero_z_skipped: the structural shortcuts that
 * apply; *shared_diffuse_steps: the length of the diffuse phase the kernel runs itself for series
 * without a missing value in it (0: P_inf does not vanish within 16 steps, or some F_inf < 1e-7); *flop_per_step: FP64
 * operations per series and time step of the resulting instantiation.  Any output pointer may be NULL. */
int ssgpu_plan_thread_kernel(const double* T, const double* R, const double* Q, const double* Z, const double* a,
                             const double* P_inf, const double* P_star, int m, int r, int N, int* n_blocks,
                             int* residual_dropped, int* level_slope_by_sums, int* zero_z_skipped,
                             int* shared_diffuse_steps, int* flop_per_step);
/* Host-only helper (no device needed): what the core kernel (SSGPU_KERNEL_CORE) does with a model whose matrices (all
 * dense, column-major, time-invariant) are shared by the batch.  *eligible: the residual structure holds (Z = [I_p | .],
 * T and P_inf zero in the residual rows and columns, H diagonal and equal in P_star and R Q R', a zero there) and at most
 * 8 component states carry a variance; *n_stochastic: component states left in the recursion; *n_deterministic: component
 * states with a known start, no disturbance and no stochastic driver, tabulated on the host; *diffuse: the instantiation
 * with the exact diffuse phase runs.  Any output pointer may be NULL. */
int ssgpu_plan_core(const double* T, const double* R, const double* Q, const double* Z, const double* a,
                    const double* P_inf, const double* P_star, int m, int r, int p, int* eligible, int* n_stochastic,
                    int* n_deterministic, int* diffuse);
/* FP64 operations (fma = 2) the core kernel (SSGPU_KERNEL_CORE) issues per series and time point of the regular
 * phase for a model with m states of which p are residuals; 0 when m - p or p is outside the kernel's range. */
int ssgpu_core_flops_per_step(int m, int p);

/* ---- 7. Measurement helpers ---------------------------------------------------------------
 * FP64 FMA-pipe peak of the current device, measured with a register-resident DFMA chain kernel
 * (`iters` x 256 dependent-free DFMAs per thread, grid = SMs x 8 CTAs x 256 threads): *tflops is
 * 2 * DFMAs / elapsed.  mode 0 = DFMA, 1 = DMMA (mma.sync.m8n8k4.f64), 2 = both interleaved. */
int ssgpu_fp64_peak(int mode, int iters, int repeats, double* tflops, double* ms);

#ifdef __cplusplus
}
#endif
#endif /* SSGPU_H */
