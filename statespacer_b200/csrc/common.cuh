// Shared declarations for libssgpu (sm_100a).  Internal -- the public boundary is include/ssgpu.h.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
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

// ---- device memory: stream-ordered allocations from the default pool (cached by the driver) ---
cudaStream_t lib_stream();
int ensure_device();

class DevBuf {
 public:
  DevBuf() {}
  ~DevBuf() { free(); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  int alloc(size_t bytes, cudaStream_t st);
  void free();
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p_); }
  void* get() const { return p_; }

 private:
  void* p_ = nullptr;
  cudaStream_t st_ = nullptr;
};

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
                       cudaStream_t st);

// ---- simulation smoother (sim.cu) --------------------------------------------------------------
struct SimArgs {
  int S, N, p, m, rR, r, repeat_Q, nQ;
  int draw_initial, eta_only;
  const double* a0;
  DMat Z, T, R;
  const double* Q_root;  // r x r x nQ
  const double* P_root;  // m x m
  const double* d0;      // [k + m*s]
  const double* dt;      // [t][kk + r_tot*s]
  double *y, *a, *eta;   // native layouts, may be null
};

struct FsArgs {
  int S, N, p, m, r, init_steps, qtr_tv;
  const double* a0;
  const double* P_inf;
  const double* P_star;
  DMat Z, T, R;
  KalmanScratch g;   // code, F1, K0, K1, QtR
  const double* y;   // native [(t*p+j)*S + s]
  double* v;         // scratch, native [(t*p+j)*S + s]
  double* a_smooth;  // native [(t*m+k)*S + s]; holds r_t between the backward and the last pass
  double* eta;       // native [(t*r+kk)*S + s]
};

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
int launch_loglik_mma(const DModel& md, const double* y, long long B, int V, double* out, bool tile_sparse,
                      cudaStream_t st);

// Column-per-thread register kernel (loglik_cols.cu)
bool cols_eligible(const DModel& md, std::string* why);
int launch_loglik_cols(const DModel& md, const double* y, long long B, int V, double* out, cudaStream_t st);

}  // namespace ssgpu
