// Shared declarations for libssgpu (sm_100a).  Internal -- the public boundary is include/ssgpu.h.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

#include "../../include/ssgpu.h"

namespace ssgpu {

constexpr double kEps = 1e-7;      // src/LogLikC.cpp:49,109,112,158,173,205
constexpr double kKappa = 1.0e7;   // src/KalmanC.cpp:63
// -log(2*pi)/2, src/LogLikC.cpp:62
constexpr double kLogConst = -0.91893853320467274178032973640562;

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(int n = 1);
void set_last_kernel(const char* name);

#define SS_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) return ::ssgpu::cuda_fail(e__, #call, __FILE__, __LINE__); \
  } while (0)

#define SS_ARG(cond, msg)      \
  do {                         \
    if (!(cond)) {             \
      ::ssgpu::set_error(msg); \
      return SSGPU_E_ARG;      \
    }                          \
  } while (0)

#define SS_TRY(call)            \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != SSGPU_OK) return rc__; \
  } while (0)

// ---- device-side matrix descriptor ---------------------------------------------------------
// One system matrix of a batch: the block of evaluation (b, v) starts at p + b*sb + v*sv; slice t
// of a time-varying matrix is st elements further (st = 0 when time-invariant).
struct DMat {
  const double* p;
  long long sb, sv, st;
  int rows, cols, diag;
  __host__ __device__ const double* block(long long b, long long v, int t) const {
    return p + b * sb + v * sv + (long long)t * st;
  }
  __host__ __device__ bool shared() const { return sb == 0 && sv == 0; }
  __host__ __device__ bool tv() const { return st != 0; }
};

__host__ __device__ inline double mat_get(const double* blk, int diag, int rows, int i, int j) {
  if (diag) return i == j ? blk[i] : 0.0;
  return blk[i + (long long)j * rows];
}

struct DModel {
  int N, p, m, r;
  DMat a, P_inf, P_star, Z, T, R, Q;
};

// Scratch written by the forward pass of KalmanC for its backward pass (one series), per
// univariate step index = t*p + j.
struct KalmanScratch {
  int* code;    // [Np] 0 = no update (missing / F ~ 0), 1 = update with F_star, 2 = diffuse update
                //      with F_inf (src/KalmanC.cpp:280-315); bit 8 set = step lies in the diffuse phase
  double* F1;   // [Np]
  double* F2;   // [Np]
  double* v;    // [Np]  v_UT
  double* K0;   // [Np*m]
  double* K1;   // [Np*m]
  double* QtR;  // [N*r*m] Q_t R_t' per time point (src/KalmanC.cpp:107-109,139-142), r x m col-major
};

// Batched KalmanC: series b owns the b-th block of every output / scratch array (each block in the reference's
// layout for one series).  Null members stay null.
__host__ __device__ inline void kalman_shift(ssgpu_kalman_out& k, KalmanScratch& s, long long b, int N, int p, int m,
                                             int r, bool qtr_tv) {
  const long long Np = (long long)N * p, mm = (long long)m * m, Nm = (long long)N * m;
  auto sh = [&](double*& x, long long n) {
    if (x) x += b * n;
  };
  if (k.initialisation_steps) k.initialisation_steps += b;
  sh(k.loglik, Np); sh(k.a_pred, Nm); sh(k.a_fil, Nm); sh(k.a_smooth, Nm);
  sh(k.P_pred, mm * N); sh(k.P_fil, mm * N); sh(k.V, mm * N); sh(k.P_inf_pred, mm * N); sh(k.P_star_pred, mm * N);
  sh(k.P_inf_fil, mm * N); sh(k.P_star_fil, mm * N);
  sh(k.yfit, Np); sh(k.v, Np); sh(k.v_norm, Np); sh(k.e, Np);
  sh(k.eta, (long long)N * r); sh(k.Fmat, (long long)p * p * N); sh(k.eta_var, (long long)r * r * N);
  sh(k.D, (long long)p * p * N);
  sh(k.a_fc, m); sh(k.P_inf_fc, mm); sh(k.P_star_fc, mm); sh(k.P_fc, mm);
  sh(k.Tstat_observation, Np); sh(k.Tstat_state, (long long)(N + 1) * m);
  sh(k.r_vec, (long long)(N + 1) * m); sh(k.Nmat, mm * (N + 1));
  if (s.code) s.code += b * Np;
  sh(s.F1, Np); sh(s.F2, Np); sh(s.v, Np); sh(s.K0, Np * m); sh(s.K1, Np * m);
  sh(s.QtR, (long long)(qtr_tv ? N : 1) * r * m);
}

// ---- device memory: stream-ordered allocations from the default pool (cached by the driver) ---
cudaStream_t lib_stream();       // compute stream of the calling thread's device slot
cudaStream_t lib_copy_stream();  // second stream of the slot (uploads that overlap kernels)
int ensure_device();
int dev_slot();                  // device slot of the calling thread (0 unless a worker of for_each_device)
int devices_in_use();
// fn(slot, lo, hi) on one worker thread per device slot in use, for the contiguous ranges of 0 .. n (multiples of align)
int for_each_device(long long n, long long align, const std::function<int(int, long long, long long)>& fn);

class DevBuf {
 public:
  DevBuf() {}
  ~DevBuf() { free(); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  int alloc(size_t bytes, cudaStream_t st);
  void free();
  void borrow(void* p) {  // memory owned by someone else (the per-device call arena): never freed here
    free();
    p_ = p;
    borrowed_ = true;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p_); }
  void* get() const { return p_; }

 private:
  void* p_ = nullptr;
  cudaStream_t st_ = nullptr;
  bool borrowed_ = false;
};

// ---- large host <-> device copies through pinned staging when the host side is pageable (xfer.cu) ----------
int copy_d2h(void* dst, const void* src_dev, size_t bytes, cudaStream_t st);   // returns when dst holds the data
int copy_h2d(void* dst_dev, const void* src, size_t bytes, cudaStream_t st);   // pinned src: valid until st is synchronised
// the same for `rows` rows of `width` bytes each, `spitch` bytes apart in the caller's memory, dense on the device
int copy_h2d_2d(void* dst_dev, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t st);
void xfer_release();
bool host_is_pinned(const void* p);

// ---- launchers (implemented in the .cu files) ----------------------------------------------
// Generic forward filter, one CTA per evaluation (fwd_generic.cu).  kout == nullptr -> loglik only
// (out_loglik[e]); otherwise B = V = 1 and every non-null member of kout (DEVICE pointers) is filled.
int launch_fwd_generic(const DModel& md, const double* y, long long B, int V, double* out_loglik,
                       const ssgpu_kalman_out* kout, const KalmanScratch* scratch, cudaStream_t st);
int launch_gains_generic(const DModel& md, int init_steps, const KalmanScratch* scratch, cudaStream_t st);
// KalmanC diagnostics (kalman_diag.cu); scratch: 2*N*m*p doubles; needs P_pred, P_inf_pred, Fmat, v, r_vec, Nmat
int launch_kalman_diag(const DModel& md, const ssgpu_kalman_out* kout, const int* code, double* scratch,
                       cudaStream_t st);
int launch_bwd_generic(const DModel& md, const ssgpu_kalman_out* kout, const KalmanScratch* scratch,
                       cudaStream_t st, long long B = 1);

// ---- simulation smoother (sim.cu) --------------------------------------------------------------
// Non-zero pattern of a (possibly time-varying) matrix, the union over its slices: the entries of row i (or of
// column i for a by-columns pattern) are idx[ptr[i]] .. idx[ptr[i+1]-1], ascending.  Values are read from the
// dense matrix itself.  Skipping an exact zero leaves an fma chain bit-identical (fma(0, x, acc) == acc for
// finite x), so the per-draw kernels stay bit-exact against the dense-loop oracle while a SARIMA companion
// matrix costs its ~30 non-zeros instead of m^2 = 784 multiply-adds.
struct SpPat {
  const int* ptr;
  const int* idx;
};

// host-side builder: all patterns of one call are packed into one int buffer and uploaded together
class PatPack {
 public:
  // by rows of a rows x cols x n_slices column-major array; returns the offset of the pattern
  size_t rows_of(const double* M, int rows, int cols, int n_slices);
  size_t cols_of(const double* M, int rows, int cols, int n_slices);
  // structural pattern of Q_t R_t' (r x m) by rows
  size_t qtr_of(const double* Q, int nQ, const double* R, int nR, int m, int r);
  size_t dense(int rows, int cols);
  const std::vector<int>& words() const { return w_; }
  SpPat at(const int* dev_base, size_t off, int n_rows) const { return SpPat{dev_base + off, dev_base + off + n_rows + 1}; }

 private:
  size_t finish(const std::vector<std::vector<int>>& lists);
  std::vector<int> w_;
};

// Fixed-width (ELL) form of a sparse matrix for the per-draw kernels: row i (or column i of the matrix for a
// by-columns build) holds W (index, value) pairs, k ascending, padded with (0, 0.0) -- fma(0, x[0], acc) == acc, so
// padding changes no bit.  W is the largest non-zero count of a row over all slices' union pattern; a uniform W
// lets the inner loop be unrolled with independent loads.  Time-varying matrices carry one value block per slice.
struct Ell {
  const int* idx;     // [W][rows]: entry w of row i at w*rows + i
  const double* val;  // [n_slices][W][rows]
  int W, rows;
  long long st;       // value stride between slices (0 = time-invariant)
  __host__ __device__ const double* vals(int t) const { return val + (st ? (long long)t * st : 0); }
};

class EllPack {
 public:
  struct Ref {
    size_t io, vo;
    int W, rows;
    long long st;
  };
  Ref add(const double* M, int rows, int cols, int n_slices, bool by_cols);
  const std::vector<int>& ints() const { return i_; }
  const std::vector<double>& vals() const { return v_; }
  Ell at(const Ref& r, const int* ib, const double* vb) const { return Ell{ib + r.io, vb + r.vo, r.W, r.rows, r.st}; }

 private:
  std::vector<int> i_;
  std::vector<double> v_;
};

struct SimArgs {
  int S, N, p, m, rR, r, repeat_Q, nQ;
  int draw_initial, eta_only;
  const double* a0;
  DMat Z, T, R;
  SpPat Zp;              // by rows
  Ell Te, Re;            // by rows
  const double* Q_root;  // r x r x nQ
  const double* P_root;  // m x m
  const double* d0;      // [k + m*s]
  const double* dt;      // [t][kk + r_tot*s]
  double *y, *a, *eta;   // native layouts, may be null
  // where the outputs land when several components write into the stacked arrays of the whole model
  // (R/SimSmoother.R:104-110: a[, a_indices, ] <- sim$a, eta[, eta_indices, ] <- sim$eta, y <- y + sim$y):
  // a[(t*a_ld + a_off + k)*S + s], eta[(t*eta_ld + eta_off + kk)*S + s]; 0 / 0 = the component's own m / r_tot.
  // y_acc: add to y instead of storing (with eta_only: y += eta, R/SimSmoother.R:575-576)
  int a_ld, a_off, eta_ld, eta_off, y_acc;
};

struct FsArgs {
  int S, N, p, m, r, init_steps, qtr_tv;
  const double* a0;
  const double* P_inf;
  const double* P_star;
  DMat Z, T, R;
  SpPat Zp, Qp;      // Z, QtR by rows
  Ell Te, Tce, Re;   // T, R by rows; T by columns (Tce)
  KalmanScratch g;   // code, F1, K0, K1, QtR
  const double* y;   // native [(t*p+j)*S + s]
  double* v;         // scratch, native [(t*p+j)*S + s]
  double* a_smooth;  // native [(t*m+k)*S + s]
  double* eta;       // native [(t*r+kk)*S + s]
};

void last_sim_kernel_ms(double* sim, double* fs, double* ext);
// x[(t*K + k)*S + s] = (x[..] - sm[(t*K_sm + off_sm + k)*S + s]) + hat[t + N*k]   (R/SimSmoother.R:634-642)
int launch_sim_correct(double* x, const double* sm, const double* hat, int N, int K, int K_sm, int off_sm, long long S,
                       cudaStream_t st);
int launch_simulate(const SimArgs& A, cudaStream_t st);
int launch_fs_draws(const FsArgs& A, cudaStream_t st);
int launch_extract(const double* a, int a_native, const DMat& Z, int m, int p, int S, int N, double* out,
                   cudaStream_t st);
// native X[(t*K+k)*S+s] <-> R array [t + N*(k + K*s)]
int launch_permute_tks(const double* src, double* dst, int N, int K, long long S, bool to_r, cudaStream_t st);
// native X[(t*m+k)*S+s] -> a_t [k + m*(s + S*t)]
int launch_permute_at(const double* src, double* dst, int N, int m, long long S, cudaStream_t st);

// Warp-per-evaluation FP64 tensor kernel (loglik_mma.cu)
bool mma_eligible(const DModel& md, std::string* why);
// T_host: optional host copy of the shared T (saves the device-to-host fetch the state-order planner needs)
int launch_loglik_mma(const DModel& md, const double* y, long long B, int V, double* out, bool tile_sparse,
                      cudaStream_t st, const double* T_host = nullptr);

// Warp-per-evaluation kernel with the covariance in shared memory, m <= 63 (loglik_wsm.cu)
bool wsm_eligible(const DModel& md, std::string* why);
int launch_loglik_wsm(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                      const double* T_host = nullptr);

// theta -> variances on the device, finite-difference gradient (theta.cu)
int launch_expand_theta(const double* theta, long long B, int V, int k, int r, int m, const int* q_index,
                        const int* p_index, const double* q_fixed, const double* p_fixed, double h, int fd, double* Qd,
                        double* Pd, cudaStream_t st);
int launch_expand_cov(const double* theta, long long B, int V, int k, int r, int m, const ssgpu_cov_block* blocks,
                      int n_blocks, const double* q_tpl, const double* p_tpl, double h, int fd, double* Qe, double* Pe,
                      int* valid, cudaStream_t st);
int launch_mask_invalid(double* out, const int* valid, long long E, cudaStream_t st);
int launch_fd_gradient(const double* out, long long B, int k, double h, double* loglik, double* grad, cudaStream_t st);

// Block-row DFMA kernel for transition matrices that are block diagonal with blocks of at most 2 x 2 (loglik_blk.cu)
struct HostShared {  // optional host copies of the shared, time-invariant matrices (null = fetch from the device)
  const double *T, *R, *Z, *Q;
  const double *a = nullptr, *P_star = nullptr, *P_inf = nullptr;  // in the storage of the device copy (diag or dense)
};
struct BlkPlan {
  int nb;                // number of 2 x 2 blocks (a lone state is padded with an empty partner)
  int q_offdiag;         // R Q R' has off-diagonal entries inside the blocks
  signed char perm[32];  // internal position 2I+e holds state perm[2I+e] of the caller, -1 = padding
  int eps;               // position 1 of the last block is the residual state (set by the launcher, which sees Z)
  int ls;                // block 0 is a local level with slope: T_0 = [[1, 1], [0, 1]] (set by the launcher)
  int zs;                // z = 0 for the second state of every block (set by the launcher)
  double Tb[16][4];      // T_J[row][col] at [2*row + col]
};
bool plan_blocks(const double* Th, int t_diag, const double* Rh, int r_diag, const unsigned char* qpat, int m, int r,
                 BlkPlan* pl, std::string* why);
bool blk_eligible(const DModel& md, std::string* why);
int blk_lanes(int nb);
int blk_flops_per_step(int nb, int p);
// applicable != null: a model whose T / R Q R' does not decompose is not an error; *applicable = false, nothing runs
int launch_loglik_blk(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                      const HostShared* hs = nullptr, bool* applicable = nullptr, bool lanes_only = false,
                      long long call_evals = 0);  // evaluations of the whole API call (a launch may cover a part): kernel choice

// Thread-per-evaluation kernel for the regular phase of the same model class (loglik_thr.cu); the lane kernel
// above runs the diffuse phase first and hands the state over (blk_handoff.cuh)
// The diffuse steps of evaluations without a missing value in the diffuse window share P_inf, F_inf and K0 (they depend
// on Z, T and P_inf only): per-step constants computed on the host (loglik_blk.cu, launcher), passed by value
constexpr int kThrMaxDu = 16;
struct ThrDiffuse {
  int du;                      // diffuse steps (0 = none)
  int esumF;                   // prod of F_inf over the steps = prodF * 2^esumF, prodF in [1, 2)
  double prodF;
  double F1[kThrMaxDu];        // 1 / F_inf of step t
  double u[kThrMaxDu][16];     // T M_inf of step t, internal order of the block plan
};
// The record of a marked evaluation written from the model itself: see loglik_thr.cu, loglik_thr_setup_kernel
struct ThrSelf {
  int on;
  DMat Q, P_star;             // stored as diagonals (per evaluation or shared)
  signed char qsrc[16];       // internal position i: R Q R'[i][i] = Q[qsrc[i]] * qscale[i]; -1 = no disturbance
  signed char psrc[16];       // P_star[i][i] = P_star diag [psrc[i]]; -1 = padding
  double qscale[16];
  double a0[16];              // a, internal order (shared by the batch)
};
bool thr_supported(int nb, int p, bool ztv);
size_t thr_handoff_bytes(int nb, bool qoff, long long E);
int thr_flops_per_step(int nb, bool qoff, bool eps = false, bool ls = false, bool zs = false, bool instr = false);
int thr_smem_blocks_host(int nb, bool ztv, bool eps = false);
int launch_loglik_thr(const BlkPlan& pl, const double* z_internal, const double* zt, const double* y, long long B,
                      long long E, long long ldy, int N, const double* hd, const int* hi, double* out, cudaStream_t st,
                      bool eps = false, bool ls = false, bool zs = false, const ThrDiffuse* ud = nullptr);
int launch_thr_setup(int nb, const ThrSelf& self, long long B, long long E, double* hd, int* hi, cudaStream_t st);

// Thread-per-series kernel on the component states alone, the p residual states of the reference's layout eliminated
// (loglik_core.cu): m - p <= 8, p <= 16, everything shared and time-invariant, diagonal H
bool core_eligible(const DModel& md, std::string* why);
int core_flops_per_step(int mc, int p);
int launch_loglik_core(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st,
                       const HostShared* hs = nullptr, bool* applicable = nullptr);

// Column-per-thread register kernel (loglik_cols.cu)
bool cols_eligible(const DModel& md, std::string* why);
int launch_loglik_cols(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st);

}  // namespace ssgpu
