// C ABI of libssgpu (include/ssgpu.h): argument checks, host<->device staging, kernel dispatch.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <cctype>
#include <functional>
#include <mutex>
#include <thread>
#ifdef __linux__
#include <pthread.h>
#include <sched.h>
#endif

#include "common.cuh"

namespace ssgpu {

// ---- error / bookkeeping ---------------------------------------------------------------------
static thread_local std::string g_err;
static thread_local const char* g_last_kernel = "";
static std::atomic<long long> g_launches{0};
static std::mutex g_mu;
// One context per device the process uses.  Slot 0 is the device of ssgpu_set_device (the only one unless
// ssgpu_use_devices asks for more); the host-pointer batch entry points then shard their series / draws over the slots
// with one worker thread per device.  A thread works on the slot named by t_slot (0 for every caller thread).
constexpr int kMaxDev = 16;
struct DevCtx {
  int device = -1;
  cudaStream_t stream = nullptr, copy = nullptr;
  bool ready = false;
};
static DevCtx g_ctx[kMaxDev];
static int g_device = 0;    // device of slot 0
static int g_n_use = 1;     // slots the sharded entry points use
static bool g_ready = false;
static thread_local int t_slot = 0;

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }
void set_last_kernel(const char* name) { g_last_kernel = name; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  g_err = buf;
  cudaGetLastError();  // clear sticky-less errors
  return SSGPU_E_CUDA;
}

int dev_slot() { return t_slot; }

// Worker threads of a sharded call run on (and first-touch their pinned staging ring from) the CPUs of the NUMA node
// their device's PCIe root hangs off: staging copies that cross the socket interconnect are what an 8-GPU call from
// one process loses first.  Best effort (Linux sysfs); a thread that cannot be placed stays where it is.
static void bind_thread_to_device_numa(int device) {
#ifdef __linux__
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  for (char* c = bus; *c; c++) *c = (char)tolower(*c);
  char path[96];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bus);
  FILE* f = fopen(path, "r");
  if (!f) return;
  char line[1024] = {0};
  const bool ok = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!ok) return;
  cpu_set_t now, want;
  CPU_ZERO(&want);
  if (pthread_getaffinity_np(pthread_self(), sizeof now, &now) != 0) return;
  int n_set = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int lo = 0, hi = 0;
    const int got = sscanf(tok, "%d-%d", &lo, &hi);
    if (got == 1) hi = lo;
    if (got < 1) continue;
    for (int c = lo; c <= hi && c < CPU_SETSIZE; c++)
      if (CPU_ISSET(c, &now)) {
        CPU_SET(c, &want);
        n_set++;
      }
  }
  if (n_set > 0) pthread_setaffinity_np(pthread_self(), sizeof want, &want);
#else
  (void)device;
#endif
}

int ensure_device() {
  std::lock_guard<std::mutex> lk(g_mu);
  DevCtx& c = g_ctx[t_slot];
  if (c.ready) {
    SS_CUDA(cudaSetDevice(c.device));
    return SSGPU_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    g_err = "no CUDA device available (libssgpu has no CPU fallback)";
    return SSGPU_E_CUDA;
  }
  c.device = (g_device + t_slot) % n;
  SS_CUDA(cudaSetDevice(c.device));
  SS_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  SS_CUDA(cudaStreamCreateWithFlags(&c.copy, cudaStreamNonBlocking));
  cudaMemPool_t pool;
  SS_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
  unsigned long long thr = ~0ULL;  // keep freed blocks cached in the pool
  SS_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  c.ready = true;
  if (t_slot == 0) g_ready = true;
  return SSGPU_OK;
}

cudaStream_t lib_stream() { return g_ctx[t_slot].stream; }
cudaStream_t lib_copy_stream() { return g_ctx[t_slot].copy; }

// Runs fn(slot, lo, hi) for the contiguous ranges [lo, hi) of 0 .. n that the slots in use own (multiples of `align`),
// one worker thread per slot; returns the first failure with its message.
int for_each_device(long long n, long long align, const std::function<int(int, long long, long long)>& fn) {
  int use;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    use = g_n_use;
  }
  const long long per = ((n + use - 1) / use + align - 1) / align * align;
  std::vector<std::thread> th;
  std::vector<int> rc(use, SSGPU_OK);
  std::vector<std::string> msg(use);
  const char* kname[kMaxDev] = {nullptr};
  for (int s = 0; s < use; s++) {
    const long long lo = s * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([&, s, lo, hi]() {
      t_slot = s;
      rc[s] = ensure_device();
      if (rc[s] == SSGPU_OK) bind_thread_to_device_numa(g_ctx[s].device);
      if (rc[s] == SSGPU_OK) rc[s] = fn(s, lo, hi);
      if (rc[s] != SSGPU_OK) msg[s] = g_err;
      kname[s] = g_last_kernel;
    });
  }
  for (auto& t : th) t.join();
  if (kname[0]) g_last_kernel = kname[0];
  for (int s = 0; s < use; s++)
    if (rc[s] != SSGPU_OK) {
      g_err = "device slot " + std::to_string(s) + ": " + msg[s];
      return rc[s];
    }
  return SSGPU_OK;
}
int devices_in_use() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_n_use;
}

int DevBuf::alloc(size_t bytes, cudaStream_t st) {
  free();
  st_ = st;
  if (bytes == 0) bytes = 8;
  SS_CUDA(cudaMallocAsync(&p_, bytes, st));
  return SSGPU_OK;
}
void DevBuf::free() {
  if (p_ && !borrowed_) cudaFreeAsync(p_, st_);
  p_ = nullptr;
  borrowed_ = false;
}

// ---- call arena: what a single small call needs without touching the allocator -----------------------------------
// One per device slot: a pinned host block the arguments are packed into, a device block of the same size they are
// copied to with ONE asynchronous copy, and a pinned word the scalar result comes back into.  ssgpu_LogLikC on a
// 100-point series otherwise spends more time in cudaMallocAsync / cudaFreeAsync and in the driver's staging of two
// pageable copies than in its kernel.  A call takes the arena only if it is free (try_lock) and large enough.
constexpr size_t kArenaBytes = 1u << 20;
struct CallArena {
  std::mutex mu;
  void* host = nullptr;
  void* dev = nullptr;
  double* res_host = nullptr;
  size_t used = 0;  // bytes of the blocks taken by the upload of the current call (the rest is free for its results)
  bool ready = false;
  int init() {
    if (ready) return SSGPU_OK;
    SS_CUDA(cudaHostAlloc(&host, kArenaBytes, cudaHostAllocDefault));
    SS_CUDA(cudaMalloc(&dev, kArenaBytes + 256));
    SS_CUDA(cudaHostAlloc((void**)&res_host, 256, cudaHostAllocDefault));
    ready = true;
    return SSGPU_OK;
  }
  double* res_dev() const { return reinterpret_cast<double*>(static_cast<char*>(dev) + kArenaBytes); }
};
static CallArena g_arena[kMaxDev];
struct ArenaLease {
  CallArena* a = nullptr;
  ArenaLease() {
    CallArena& c = g_arena[t_slot];
    if (c.mu.try_lock()) {
      if (c.init() == SSGPU_OK) {
        a = &c;
        c.used = 0;
      } else {
        c.mu.unlock();
      }
    }
  }
  ~ArenaLease() {
    if (a) a->mu.unlock();
  }
};

// ---- host-side packing of many small arrays into one H2D copy ------------------------------
class Packer {
 public:
  // returns the element offset (in doubles) of the copied block
  size_t add(const double* src, size_t n) {
    const size_t off = buf_.size();
    const size_t padded = (n + 31) / 32 * 32;  // keep every block 256-byte aligned
    buf_.resize(off + padded, 0.0);
    if (n) std::memcpy(buf_.data() + off, src, n * sizeof(double));
    return off;
  }
  size_t reserve(size_t n) {
    const size_t off = buf_.size();
    buf_.resize(off + (n + 31) / 32 * 32, 0.0);
    return off;
  }
  double* host(size_t off) { return buf_.data() + off; }
  size_t size() const { return buf_.size(); }
  int upload(DevBuf& d, cudaStream_t st, CallArena* arena = nullptr) {
    const size_t bytes = buf_.size() * sizeof(double);
    if (arena && bytes <= kArenaBytes) {  // pinned staging, persistent device block: one truly asynchronous copy
      std::memcpy(arena->host, buf_.data(), bytes);
      d.borrow(arena->dev);
      SS_CUDA(cudaMemcpyAsync(arena->dev, arena->host, bytes, cudaMemcpyHostToDevice, st));
      arena->used = bytes;
      return SSGPU_OK;
    }
    SS_TRY(d.alloc(bytes, st));
    SS_CUDA(cudaMemcpyAsync(d.get(), buf_.data(), bytes, cudaMemcpyHostToDevice, st));
    return SSGPU_OK;
  }

 private:
  std::vector<double> buf_;
};

static DMat make_dmat(const double* dev, int rows, int cols, int n_slices) {
  DMat d{};
  d.p = dev;
  d.rows = rows;
  d.cols = cols;
  d.diag = 0;
  d.sb = d.sv = 0;
  d.st = n_slices > 1 ? (long long)rows * cols : 0;
  return d;
}

static int check_single(int N, int p, int m, int r, int nZ, int nT, int nR, int nQ) {
  SS_ARG(N >= 1 && p >= 1 && m >= 1 && r >= 1, "N, p, m, r must be positive");
  SS_ARG(m <= SSGPU_MAX_M, "state dimension m exceeds SSGPU_MAX_M");
  SS_ARG((nZ == 1 || nZ == N) && (nT == 1 || nT == N) && (nR == 1 || nR == N) && (nQ == 1 || nQ == N),
         "system matrices must have 1 or N slices");
  return SSGPU_OK;
}

// y (N x p column-major, optional mask) -> time-major [t*p + j] with NaN for missing
static void pack_y_single(const double* y, const int* isna, int N, int p, double* dst) {
  for (int t = 0; t < N; t++)
    for (int j = 0; j < p; j++) {
      const double v = y[t + (size_t)N * j];
      dst[(size_t)t * p + j] = (isna && isna[t + (size_t)N * j]) ? nan("") : v;
    }
}

struct SingleModel {
  Packer pk;
  DevBuf dev;
  DModel md;
  size_t oy = 0;
};

// uploads y + the 7 matrices of one model, fills md with device pointers
static int stage_single(SingleModel& S, const double* y, const int* y_isna, int N, int p, int m, int r, const double* a,
                        const double* P_inf, const double* P_star, const double* Z, int nZ, const double* T, int nT,
                        const double* R, int nR, const double* Q, int nQ, cudaStream_t st, CallArena* arena = nullptr) {
  Packer& pk = S.pk;
  S.oy = pk.reserve((size_t)N * p);
  pack_y_single(y, y_isna, N, p, pk.host(S.oy));
  const size_t oa = pk.add(a, m), opi = pk.add(P_inf, (size_t)m * m), ops = pk.add(P_star, (size_t)m * m);
  const size_t oz = pk.add(Z, (size_t)p * m * nZ), ot = pk.add(T, (size_t)m * m * nT);
  const size_t orr = pk.add(R, (size_t)m * r * nR), oq = pk.add(Q, (size_t)r * r * nQ);
  SS_TRY(pk.upload(S.dev, st, arena));
  const double* d = S.dev.as<double>();
  DModel& md = S.md;
  md.N = N;
  md.p = p;
  md.m = m;
  md.r = r;
  md.a = make_dmat(d + oa, m, 1, 1);
  md.P_inf = make_dmat(d + opi, m, m, 1);
  md.P_star = make_dmat(d + ops, m, m, 1);
  md.Z = make_dmat(d + oz, p, m, nZ);
  md.T = make_dmat(d + ot, m, m, nT);
  md.R = make_dmat(d + orr, m, r, nR);
  md.Q = make_dmat(d + oq, r, r, nQ);
  return SSGPU_OK;
}

// ---- ssgpu_matrix (public) -> DMat ------------------------------------------------------------
static int to_dmat(const ssgpu_matrix& M, int rows, int cols, int N, const char* name, DMat* out) {
  std::string nm(name);
  SS_ARG(M.data != nullptr, nm + ": null data pointer");
  SS_ARG(M.rows == rows && M.cols == cols, nm + ": unexpected dimensions");
  SS_ARG(M.n_slices == 1 || M.n_slices == N, nm + ": n_slices must be 1 or N");
  SS_ARG(!M.diag || rows == cols, nm + ": diag storage needs a square matrix");
  SS_ARG(M.stride_series >= 0 && M.stride_variant >= 0, nm + ": negative strides are not supported");
  out->p = M.data;
  out->rows = rows;
  out->cols = cols;
  out->diag = M.diag ? 1 : 0;
  out->sb = M.stride_series;
  out->sv = M.stride_variant;
  out->st = M.n_slices > 1 ? (M.diag ? rows : (long long)rows * cols) : 0;
  return SSGPU_OK;
}

static int to_dmodel(const ssgpu_model* mo, DModel* md) {
  SS_ARG(mo != nullptr, "null model");
  SS_ARG(mo->N >= 1 && mo->p >= 1 && mo->m >= 1 && mo->r >= 1, "N, p, m, r must be positive");
  SS_ARG(mo->m <= SSGPU_MAX_M, "state dimension m exceeds SSGPU_MAX_M");
  md->N = mo->N;
  md->p = mo->p;
  md->m = mo->m;
  md->r = mo->r;
  SS_TRY(to_dmat(mo->a, mo->m, 1, mo->N, "a", &md->a));
  SS_ARG(mo->a.n_slices == 1 && !mo->a.diag, "a: must be a plain vector");
  SS_TRY(to_dmat(mo->P_inf, mo->m, mo->m, mo->N, "P_inf", &md->P_inf));
  SS_TRY(to_dmat(mo->P_star, mo->m, mo->m, mo->N, "P_star", &md->P_star));
  SS_ARG(mo->P_inf.n_slices == 1 && mo->P_star.n_slices == 1, "P_inf / P_star cannot be time-varying");
  SS_TRY(to_dmat(mo->Z, mo->p, mo->m, mo->N, "Z", &md->Z));
  SS_ARG(!mo->Z.diag, "Z: diag storage is not supported");
  SS_TRY(to_dmat(mo->T, mo->m, mo->m, mo->N, "T", &md->T));
  SS_TRY(to_dmat(mo->R, mo->m, mo->r, mo->N, "R", &md->R));
  SS_TRY(to_dmat(mo->Q, mo->r, mo->r, mo->N, "Q", &md->Q));
  return SSGPU_OK;
}

static size_t extent(const ssgpu_matrix& M, long long B, int V) {
  const size_t slice = M.diag ? (size_t)M.rows : (size_t)M.rows * M.cols;
  return (size_t)((B - 1) * M.stride_series + (V - 1) * M.stride_variant) + slice * M.n_slices;
}

static bool auto_blk_enabled() {
  static const bool on = [] {
    const char* e = getenv("SSGPU_AUTO_BLK");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}

static bool auto_core_enabled() {
  static const bool on = [] {
    const char* e = getenv("SSGPU_AUTO_CORE");
    return e ? atoi(e) != 0 : true;
  }();
  return on;
}
constexpr long long kCoreMinEvals = 2048;

static int dispatch_batch(const DModel& md, const double* y_dev, long long B, int V, int kernel, double* out_dev,
                          cudaStream_t st, const HostShared* hs = nullptr, long long call_evals = 0) {
  SS_ARG(B >= 1 && V >= 1, "B and V must be positive");
  std::string why;
  const double* T_host = hs ? hs->T : nullptr;
  // block-diagonal T with blocks of at most 2 x 2 (structural models): the block-row kernel issues 4 m^2 instead of
  // 4 m^3 multiply-adds per time update
  if (kernel == SSGPU_KERNEL_BLK) return launch_loglik_blk(md, y_dev, B, V, out_dev, st, hs);
  if (kernel == SSGPU_KERNEL_BLK_LANES) return launch_loglik_blk(md, y_dev, B, V, out_dev, st, hs, nullptr, true);
  if (kernel == SSGPU_KERNEL_CORE) return launch_loglik_core(md, y_dev, B, V, out_dev, st, hs);
  // thread-per-series kernel on the component states (residual states eliminated): worth it once the batch fills the
  // machine with threads; many observations per time point go there before the block-row kernel (p = 1 has the
  // thread-per-evaluation kernel behind the block planner)
  const bool core_auto = kernel == SSGPU_KERNEL_AUTO && auto_core_enabled() && std::max(B * V, call_evals) >= kCoreMinEvals &&
                         core_eligible(md, &why);
  // ... and so does the local level (m = 2: one component state, p = 1 known at compile time: 2.7e11 series*timesteps/s
  // against 2.0e11 on the one-lane-per-series instantiation of the block kernel)
  if (core_auto && (md.p >= 2 || md.m == 2)) {
    bool ran = true;
    SS_TRY(launch_loglik_core(md, y_dev, B, V, out_dev, st, hs, &ran));
    if (ran) return SSGPU_OK;
  }
  if (kernel == SSGPU_KERNEL_AUTO && auto_blk_enabled() && blk_eligible(md, &why)) {
    bool ran = true;
    SS_TRY(launch_loglik_blk(md, y_dev, B, V, out_dev, st, hs, &ran, false, call_evals));
    if (ran) return SSGPU_OK;
  }
  if (core_auto && md.p < 2 && md.m != 2) {
    bool ran = true;
    SS_TRY(launch_loglik_core(md, y_dev, B, V, out_dev, st, hs, &ran));
    if (ran) return SSGPU_OK;
  }
  const bool mma_ok = mma_eligible(md, &why);
  if (kernel == SSGPU_KERNEL_MMA || kernel == SSGPU_KERNEL_MMA_DENSE)
    SS_ARG(mma_ok, "tensor kernel not eligible: " + why);
  const bool cols_ok = cols_eligible(md, &why);
  // many observations per time point: the measurement updates dominate and the column kernel's single-sweep update
  // beats the tensor kernel's shuffle-based one (config 3, p = 8, m = 14: 5.0e8 vs 3.6e8 series*timesteps/s)
  const bool prefer_cols = kernel == SSGPU_KERNEL_AUTO && cols_ok && md.p >= 4;
  if (!prefer_cols &&
      (kernel == SSGPU_KERNEL_MMA || kernel == SSGPU_KERNEL_MMA_DENSE || (kernel == SSGPU_KERNEL_AUTO && mma_ok)))
    return launch_loglik_mma(md, y_dev, B, V, out_dev, kernel != SSGPU_KERNEL_MMA_DENSE, st, T_host);
  if (kernel == SSGPU_KERNEL_COLS) SS_ARG(cols_ok, "column kernel not eligible: " + why);
  if (kernel == SSGPU_KERNEL_COLS || (kernel == SSGPU_KERNEL_AUTO && cols_ok))
    return launch_loglik_cols(md, y_dev, B, V, out_dev, st);
  const bool wsm_ok = wsm_eligible(md, &why);
  if (kernel == SSGPU_KERNEL_WSM) SS_ARG(wsm_ok, "warp/shared-memory kernel not eligible: " + why);
  if (kernel == SSGPU_KERNEL_WSM || (kernel == SSGPU_KERNEL_AUTO && wsm_ok))
    return launch_loglik_wsm(md, y_dev, B, V, out_dev, st, T_host);
  SS_ARG(kernel == SSGPU_KERNEL_AUTO || kernel == SSGPU_KERNEL_GENERIC, "unknown kernel selector");
  return launch_fwd_generic(md, y_dev, B, V, out_dev, nullptr, nullptr, st);
}

}  // namespace ssgpu

using namespace ssgpu;

extern "C" {

int ssgpu_version(void) { return 100; }
const char* ssgpu_last_error(void) { return g_err.c_str(); }
const char* ssgpu_last_kernel(void) { return g_last_kernel; }
long long ssgpu_launch_count(void) { return g_launches.load(); }

int ssgpu_set_device(int device) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    SS_ARG(!g_ready || device == g_device, "ssgpu_set_device must be called before the first compute call");
    g_device = device;
  }
  return ensure_device();
}

int ssgpu_device_count(int* count) {
  SS_ARG(count != nullptr, "null count");
  cudaError_t e = cudaGetDeviceCount(count);
  if (e != cudaSuccess) {
    *count = 0;
    cudaGetLastError();
  }
  return SSGPU_OK;
}

int ssgpu_release(void) {
  if (!g_ready) return SSGPU_OK;
  for (int s = 0; s < kMaxDev; s++) {
    DevCtx& c = g_ctx[s];
    if (!c.ready) continue;
    SS_CUDA(cudaSetDevice(c.device));
    SS_CUDA(cudaStreamSynchronize(c.stream));
    SS_CUDA(cudaStreamSynchronize(c.copy));
    cudaMemPool_t pool;
    SS_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
    SS_CUDA(cudaMemPoolTrimTo(pool, 0));
  }
  for (int s2 = 0; s2 < kMaxDev; s2++) {  // call arenas (pinned staging + device block per slot)
    CallArena& ar = g_arena[s2];
    std::lock_guard<std::mutex> lk(ar.mu);
    if (!ar.ready) continue;
    if (g_ctx[s2].ready) cudaSetDevice(g_ctx[s2].device);
    cudaFreeHost(ar.host);
    cudaFree(ar.dev);
    cudaFreeHost(ar.res_host);
    ar.host = ar.dev = nullptr;
    ar.res_host = nullptr;
    ar.ready = false;
  }
  SS_CUDA(cudaSetDevice(g_ctx[0].device));
  xfer_release();
  return SSGPU_OK;
}

int ssgpu_use_devices(int n) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    count = 0;
  }
  SS_ARG(count >= 1, "no CUDA device available (libssgpu has no CPU fallback)");
  if (n <= 0) n = count;
  // SSGPU_ALLOW_SHARED_DEVICE=1 (tests on a one-GPU box): slots beyond the visible devices wrap around and share them
  const char* share = getenv("SSGPU_ALLOW_SHARED_DEVICE");
  SS_ARG(n <= kMaxDev && (n <= count || (share && share[0] == '1')), "ssgpu_use_devices: more devices than are visible");
  {
    std::lock_guard<std::mutex> lk(g_mu);
    g_n_use = n;
  }
  return ensure_device();
}

int ssgpu_devices_in_use(int* n) {
  SS_ARG(n != nullptr, "null pointer");
  *n = devices_in_use();
  return SSGPU_OK;
}

int ssgpu_LogLikC(const double* y, const int* y_isna, int N, int p, int m, int r, const double* a, const double* P_inf,
                  const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                  const double* Q, int nQ, double* loglik) {
  SS_ARG(y && a && P_inf && P_star && Z && T && R && Q && loglik, "null pointer argument");
  SS_TRY(check_single(N, p, m, r, nZ, nT, nR, nQ));
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  ArenaLease lease;  // no allocator call, no pageable copy when the per-device call arena is free
  SingleModel S;
  SS_TRY(stage_single(S, y, y_isna, N, p, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, st, lease.a));
  DevBuf out;
  if (lease.a)
    out.borrow(lease.a->res_host);  // pinned host memory is device-accessible (unified addressing): the kernel writes the
                                    // scalar result straight into it and no copy back is enqueued
  else
    SS_TRY(out.alloc(sizeof(double), st));
  // one series is a batch of one: the warp-level kernels serve it when the model qualifies (a few hundred cycles
  // per observation instead of the CTA kernel's barriers), the generic kernel otherwise
  const HostShared hs{nT == 1 ? T : nullptr, nR == 1 ? R : nullptr, nZ == 1 ? Z : nullptr, nQ == 1 ? Q : nullptr};
  SS_TRY(dispatch_batch(S.md, S.dev.as<double>() + S.oy, 1, 1, SSGPU_KERNEL_AUTO, out.as<double>(), st, &hs));
  if (!lease.a) SS_CUDA(cudaMemcpyAsync(loglik, out.get(), sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  if (lease.a) *loglik = *lease.a->res_host;
  return SSGPU_OK;
}

int ssgpu_KalmanC(const double* y, const int* y_isna, int N, int p, int m, int r, const double* a, const double* P_inf,
                  const double* P_star, const double* Z, int nZ, const double* T, int nT, const double* R, int nR,
                  const double* Q, int nQ, int diagnostics, ssgpu_kalman_out* out) {
  SS_ARG(y && a && P_inf && P_star && Z && T && R && Q && out, "null pointer argument");
  SS_TRY(check_single(N, p, m, r, nZ, nT, nR, nQ));
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  ArenaLease lease;  // small calls: arguments and results through the per-device arena (one copy each way, no allocator)
  SingleModel S;
  SS_TRY(stage_single(S, y, y_isna, N, p, m, r, a, P_inf, P_star, Z, nZ, T, nT, R, nR, Q, nQ, st, lease.a));

  // one device block for every output + the backward-pass scratch
  const size_t Np = (size_t)N * p, mm = (size_t)m * m;
  struct Field {
    double** host;
    double** dev;
    size_t n;
  };
  ssgpu_kalman_out d{};
  std::vector<Field> f = {
      {&out->loglik, &d.loglik, Np},
      {&out->a_pred, &d.a_pred, (size_t)N * m},
      {&out->a_fil, &d.a_fil, (size_t)N * m},
      {&out->a_smooth, &d.a_smooth, (size_t)N * m},
      {&out->P_pred, &d.P_pred, mm * N},
      {&out->P_fil, &d.P_fil, mm * N},
      {&out->V, &d.V, mm * N},
      {&out->P_inf_pred, &d.P_inf_pred, mm * N},
      {&out->P_star_pred, &d.P_star_pred, mm * N},
      {&out->P_inf_fil, &d.P_inf_fil, mm * N},
      {&out->P_star_fil, &d.P_star_fil, mm * N},
      {&out->yfit, &d.yfit, Np},
      {&out->v, &d.v, Np},
      {&out->eta, &d.eta, (size_t)N * r},
      {&out->Fmat, &d.Fmat, (size_t)p * p * N},
      {&out->eta_var, &d.eta_var, (size_t)r * r * N},
      {&out->a_fc, &d.a_fc, (size_t)m},
      {&out->P_inf_fc, &d.P_inf_fc, mm},
      {&out->P_star_fc, &d.P_star_fc, mm},
      {&out->P_fc, &d.P_fc, mm},
      {&out->r_vec, &d.r_vec, (size_t)(N + 1) * m},
      {&out->Nmat, &d.Nmat, mm * (N + 1)},
      {&out->v_norm, &d.v_norm, diagnostics ? Np : 0},
      {&out->e, &d.e, diagnostics ? Np : 0},
      {&out->D, &d.D, diagnostics ? (size_t)p * p * N : 0},
      {&out->Tstat_observation, &d.Tstat_observation, diagnostics ? Np : 0},
      {&out->Tstat_state, &d.Tstat_state, diagnostics ? (size_t)(N + 1) * m : 0},
  };
  size_t total = 0;
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  for (auto& x : f) total += pad(x.n);
  const bool qtr_tv = nQ > 1 || nR > 1;
  const size_t n_diag = diagnostics ? pad(2 * (size_t)N * m * p) : 0;
  const size_t n_scr = pad(Np) * 3 + pad(Np * m) * 2 + pad((size_t)(qtr_tv ? N : 1) * r * m) + pad(Np) + 32 + n_diag;
  DevBuf blk;
  // results + scratch behind the uploaded arguments in the arena's device block when they fit: the results then come
  // back with ONE copy into the arena's pinned block and are handed out from there (27 copies into pageable memory, each
  // staged and synchronised by the driver, were half of a Nile call's 1.4 ms)
  const size_t up = lease.a ? (lease.a->used + 255) / 256 * 256 : 0;
  const bool in_arena = lease.a && lease.a->used > 0 && up + (total + n_scr) * sizeof(double) <= kArenaBytes;
  if (in_arena)
    blk.borrow(static_cast<char*>(lease.a->dev) + up);
  else
    SS_TRY(blk.alloc((total + n_scr) * sizeof(double), st));
  double* base = blk.as<double>();
  size_t off = 0;
  for (auto& x : f) {
    *x.dev = base + off;
    off += pad(x.n);
  }
  KalmanScratch s{};
  d.initialisation_steps = reinterpret_cast<int*>(base + off); off += 32;  // right behind the results: one copy takes both
  s.F1 = base + off; off += pad(Np);
  s.F2 = base + off; off += pad(Np);
  s.v = base + off; off += pad(Np);
  s.K0 = base + off; off += pad(Np * m);
  s.K1 = base + off; off += pad(Np * m);
  s.QtR = base + off; off += pad((size_t)(qtr_tv ? N : 1) * r * m);
  s.code = reinterpret_cast<int*>(base + off); off += pad(Np);
  double* diag_scr = base + off; off += n_diag;

  const double* yd = S.dev.as<double>() + S.oy;
  SS_TRY(launch_fwd_generic(S.md, yd, 1, 1, nullptr, &d, &s, st));
  SS_TRY(launch_bwd_generic(S.md, &d, &s, st));
  if (diagnostics) SS_TRY(launch_kalman_diag(S.md, &d, s.code, diag_scr, st));
  if (in_arena) {
    char* hb = static_cast<char*>(lease.a->host) + up;
    SS_CUDA(cudaMemcpyAsync(hb, base, (total + 32) * sizeof(double), cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    for (auto& x : f)
      if (*x.host && x.n) std::memcpy(*x.host, hb + ((*x.dev) - base) * sizeof(double), x.n * sizeof(double));
    if (out->initialisation_steps)
      std::memcpy(out->initialisation_steps, hb + (reinterpret_cast<double*>(d.initialisation_steps) - base) * sizeof(double), sizeof(int));
    return SSGPU_OK;
  }
  for (auto& x : f)
    if (*x.host && x.n)
      SS_CUDA(cudaMemcpyAsync(*x.host, *x.dev, x.n * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (out->initialisation_steps)
    SS_CUDA(cudaMemcpyAsync(out->initialisation_steps, d.initialisation_steps, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

// ---- batched KalmanC (include/ssgpu.h section 2b) ---------------------------------------------------------------
namespace {

struct KField {
  double* ssgpu_kalman_out::*ptr;
  size_t n;        // doubles per series
  bool needed;     // read by the backward pass: allocated as scratch when the caller does not ask for it
  bool diagnostic;
};

std::vector<KField> kalman_fields(int N, int p, int m, int r) {
  const size_t Np = (size_t)N * p, mm = (size_t)m * m, Nm = (size_t)N * m;
  typedef ssgpu_kalman_out K;
  return {
      {&K::loglik, Np, false, false},       {&K::a_pred, Nm, true, false},          {&K::a_fil, Nm, false, false},
      {&K::a_smooth, Nm, false, false},     {&K::P_pred, mm * N, false, false},     {&K::P_fil, mm * N, false, false},
      {&K::V, mm * N, false, false},        {&K::P_inf_pred, mm * N, true, false},  {&K::P_star_pred, mm * N, true, false},
      {&K::P_inf_fil, mm * N, false, false}, {&K::P_star_fil, mm * N, false, false}, {&K::yfit, Np, false, false},
      {&K::v, Np, false, false},            {&K::eta, (size_t)N * r, false, false}, {&K::Fmat, (size_t)p * p * N, false, false},
      {&K::eta_var, (size_t)r * r * N, false, false}, {&K::a_fc, (size_t)m, false, false}, {&K::P_inf_fc, mm, false, false},
      {&K::P_star_fc, mm, false, false},    {&K::P_fc, mm, false, false},           {&K::r_vec, (size_t)(N + 1) * m, false, false},
      {&K::Nmat, mm * (N + 1), false, false}, {&K::v_norm, Np, false, true},        {&K::e, Np, false, true},
      {&K::D, (size_t)p * p * N, false, true}, {&K::Tstat_observation, Np, false, true},
      {&K::Tstat_state, (size_t)(N + 1) * m, false, true}};
}

// `want`: device pointers of the arrays the caller asked for (null = not wanted).  Runs forward + backward for B series.
int kalman_batch_core(const DModel& md, const double* y_dev, long long B, const ssgpu_kalman_out& want, cudaStream_t st) {
  const int N = md.N, p = md.p, m = md.m, r = md.r;
  SS_ARG(B >= 1 && B < (1LL << 31), "kalman_batch: batch size out of range");
  SS_ARG(!md.a.sv && !md.Q.sv, "kalman_batch: one parameter vector per series");
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  const size_t Np = (size_t)N * p, nb = (size_t)B;
  const bool qtr_tv = md.Q.tv() || md.R.tv();
  ssgpu_kalman_out d = want;
  size_t extra = 0;
  const auto fields = kalman_fields(N, p, m, r);
  for (const KField& f : fields) {
    SS_ARG(!(f.diagnostic && want.*(f.ptr)), "kalman_batch: the diagnostics arrays are not available in the batched call");
    if (f.needed && !(want.*(f.ptr))) extra += pad(f.n * nb);
  }
  const size_t per_series = pad(Np) * 3 + pad(Np * m) * 2 + pad((size_t)(qtr_tv ? N : 1) * r * m) + pad(Np);
  DevBuf blk;
  SS_TRY(blk.alloc((extra + per_series * nb + pad(nb)) * sizeof(double), st));
  double* base = blk.as<double>();
  size_t off = 0;
  for (const KField& f : fields)
    if (f.needed && !(want.*(f.ptr))) {
      d.*(f.ptr) = base + off;
      off += pad(f.n * nb);
    }
  // per-series scratch blocks are contiguous per array (kalman_shift strides by the array's own block size)
  KalmanScratch s{};
  s.F1 = base + off; off += Np * nb;
  s.F2 = base + off; off += Np * nb;
  s.v = base + off; off += Np * nb;
  s.K0 = base + off; off += Np * m * nb;
  s.K1 = base + off; off += Np * m * nb;
  s.QtR = base + off; off += (size_t)(qtr_tv ? N : 1) * r * m * nb;
  s.code = reinterpret_cast<int*>(base + off); off += (Np * nb + 1) / 2 + 32;
  if (!d.initialisation_steps) d.initialisation_steps = reinterpret_cast<int*>(base + off);
  SS_TRY(launch_fwd_generic(md, y_dev, B, 1, nullptr, &d, &s, st));
  SS_TRY(launch_bwd_generic(md, &d, &s, st, B));
  SS_CUDA(cudaStreamSynchronize(st));  // the scratch block is released behind the kernels
  return SSGPU_OK;
}

}  // namespace

int ssgpu_kalman_batch_dev(const double* y, long long B, const ssgpu_model* model, const ssgpu_kalman_out* out,
                           void* stream) {
  SS_ARG(y && out, "null pointer argument");
  DModel md{};
  SS_TRY(to_dmodel(model, &md));
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  return kalman_batch_core(md, y, B, *out, st);
}

int ssgpu_kalman_batch(const double* y, long long B, const ssgpu_model* model, const ssgpu_kalman_out* out) {
  SS_ARG(y && out && model, "null pointer argument");
  DModel md{};
  SS_TRY(to_dmodel(model, &md));
  SS_ARG(B >= 1, "B must be positive");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  const int N = md.N, p = md.p, m = md.m, r = md.r;
  const ssgpu_matrix* src[7] = {&model->a, &model->P_inf, &model->P_star, &model->Z, &model->T, &model->R, &model->Q};
  DMat* dst[7] = {&md.a, &md.P_inf, &md.P_star, &md.Z, &md.T, &md.R, &md.Q};
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  const size_t n_y = (size_t)N * p * (size_t)B, nb = (size_t)B;
  size_t total = pad(n_y) + pad(nb), ext[7];
  for (int i = 0; i < 7; i++) {
    ext[i] = extent(*src[i], B, 1);
    total += pad(ext[i]);
  }
  const auto fields = kalman_fields(N, p, m, r);
  for (const KField& f : fields)
    if (out->*(f.ptr)) total += pad(f.n * nb);
  DevBuf blk;
  SS_TRY(blk.alloc(total * sizeof(double), st));
  double* base = blk.as<double>();
  size_t off = 0;
  SS_TRY(copy_h2d(base, y, n_y * sizeof(double), st));
  off += pad(n_y);
  for (int i = 0; i < 7; i++) {
    SS_CUDA(cudaMemcpyAsync(base + off, src[i]->data, ext[i] * sizeof(double), cudaMemcpyHostToDevice, st));
    dst[i]->p = base + off;
    off += pad(ext[i]);
  }
  ssgpu_kalman_out d{};
  for (const KField& f : fields)
    if (out->*(f.ptr)) {
      d.*(f.ptr) = base + off;
      off += pad(f.n * nb);
    }
  if (out->initialisation_steps) d.initialisation_steps = reinterpret_cast<int*>(base + off);
  SS_TRY(kalman_batch_core(md, base, B, d, st));
  for (const KField& f : fields)
    if (out->*(f.ptr))
      SS_TRY(copy_d2h(out->*(f.ptr), d.*(f.ptr), f.n * nb * sizeof(double), st));
  if (out->initialisation_steps)
    SS_CUDA(cudaMemcpyAsync(out->initialisation_steps, d.initialisation_steps, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

int ssgpu_loglik_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, int kernel, double* out,
                           void* stream) {
  SS_ARG(y && out, "null pointer argument");
  DModel md{};
  SS_TRY(to_dmodel(model, &md));
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  return dispatch_batch(md, y, B, V, kernel, out, st);
}

int ssgpu_loglik_batch(const double* y, long long B, int V, const ssgpu_model* model, int kernel, double* out) {
  SS_ARG(y && out, "null pointer argument");
  DModel md{};
  SS_TRY(to_dmodel(model, &md));
  SS_ARG(B >= 1 && V >= 1, "B and V must be positive");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  const ssgpu_matrix* src[7] = {&model->a, &model->P_inf, &model->P_star, &model->Z, &model->T, &model->R, &model->Q};
  DMat* dst[7] = {&md.a, &md.P_inf, &md.P_star, &md.Z, &md.T, &md.R, &md.Q};
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  size_t n_y = (size_t)model->N * model->p * (size_t)B, total = pad(n_y) + pad((size_t)B * V);
  size_t ext[7];
  for (int i = 0; i < 7; i++) {
    ext[i] = extent(*src[i], B, V);
    total += pad(ext[i]);
  }
  DevBuf blk;
  SS_TRY(blk.alloc(total * sizeof(double), st));
  double* base = blk.as<double>();
  size_t off = 0;
  SS_TRY(copy_h2d(base, y, n_y * sizeof(double), st));
  off += pad(n_y);
  for (int i = 0; i < 7; i++) {
    SS_CUDA(cudaMemcpyAsync(base + off, src[i]->data, ext[i] * sizeof(double), cudaMemcpyHostToDevice, st));
    dst[i]->p = base + off;
    off += pad(ext[i]);
  }
  double* out_dev = base + off;
  SS_TRY(dispatch_batch(md, base, B, V, kernel, out_dev, st));
  SS_CUDA(cudaMemcpyAsync(out, out_dev, (size_t)B * V * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

// ---- theta-driven batched loglikelihood / gradient (include/ssgpu.h section 6c) ---------------------------
namespace {

// The shared parts of `model` (host or device pointers, see `on_device`) -> device; Q and P_star become
// per-evaluation diagonals built from theta.  Everything allocated here is released stream-ordered.
struct ThetaWork {
  DevBuf shared, maps, Qd, Pd, out;
  DModel md;
};

int theta_setup(const ssgpu_model* model, bool on_device, long long B, int V, int k, const double* theta_dev,
                const int* q_index, const int* p_index, double h, int fd, ThetaWork& W, cudaStream_t st) {
  SS_TRY(to_dmodel(model, &W.md));
  DModel& md = W.md;
  const int m = md.m, r = md.r;
  SS_ARG(q_index && p_index, "null index vector");
  SS_ARG(k >= 1, "k must be positive");
  for (int j = 0; j < r; j++) SS_ARG(q_index[j] < k, "q_index entry out of range");
  for (int j = 0; j < m; j++) SS_ARG(p_index[j] < k, "p_star_index entry out of range");
  SS_ARG(md.Q.shared() && md.P_star.shared() && !md.Q.tv(), "Q / P_star of the model must be the shared templates");
  SS_ARG(md.a.shared() && md.P_inf.shared() && md.Z.shared() && md.T.shared() && md.R.shared(),
         "theta-driven entry points take shared a, P_inf, Z, T, R");
  // fixed diagonal entries of Q and P_star (host copy), index vectors
  std::vector<double> qf(r), pf(m);
  {
    const size_t nq = md.Q.diag ? r : (size_t)r * r, np = md.P_star.diag ? m : (size_t)m * m;
    std::vector<double> qh(nq), ph(np);
    if (on_device) {
      SS_CUDA(cudaMemcpyAsync(qh.data(), md.Q.p, nq * sizeof(double), cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaMemcpyAsync(ph.data(), md.P_star.p, np * sizeof(double), cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaStreamSynchronize(st));
    } else {
      std::memcpy(qh.data(), md.Q.p, nq * sizeof(double));
      std::memcpy(ph.data(), md.P_star.p, np * sizeof(double));
    }
    for (int j = 0; j < r; j++) {
      qf[j] = md.Q.diag ? qh[j] : qh[j + (size_t)j * r];
      if (!md.Q.diag)
        for (int l = 0; l < r; l++) SS_ARG(l == j || qh[j + (size_t)l * r] == 0.0, "Q template must be diagonal");
    }
    for (int j = 0; j < m; j++) {
      pf[j] = md.P_star.diag ? ph[j] : ph[j + (size_t)j * m];
      if (!md.P_star.diag)
        for (int l = 0; l < m; l++) SS_ARG(l == j || ph[j + (size_t)l * m] == 0.0, "P_star template must be diagonal");
    }
  }
  auto pad = [](size_t n) { return (n + 31) / 32 * 32; };
  if (!on_device) {  // a, P_inf, Z, T, R: host -> device in one block
    const ssgpu_matrix* src[5] = {&model->a, &model->P_inf, &model->Z, &model->T, &model->R};
    DMat* dst[5] = {&md.a, &md.P_inf, &md.Z, &md.T, &md.R};
    size_t total = 0, ext[5];
    for (int i = 0; i < 5; i++) {
      ext[i] = extent(*src[i], 1, 1);
      total += pad(ext[i]);
    }
    SS_TRY(W.shared.alloc(total * sizeof(double), st));
    size_t off = 0;
    for (int i = 0; i < 5; i++) {
      SS_CUDA(cudaMemcpyAsync(W.shared.as<double>() + off, src[i]->data, ext[i] * sizeof(double), cudaMemcpyHostToDevice, st));
      dst[i]->p = W.shared.as<double>() + off;
      off += pad(ext[i]);
    }
  }
  // index vectors + fixed entries
  const size_t n_maps = pad((size_t)r + m) /* ints, 4 B */ + pad((size_t)r) + pad((size_t)m);
  SS_TRY(W.maps.alloc(n_maps * sizeof(double), st));
  int* qi = W.maps.as<int>();
  int* pi = qi + r;
  double* qfd = W.maps.as<double>() + pad((size_t)r + m);
  double* pfd = qfd + pad((size_t)r);
  SS_CUDA(cudaMemcpyAsync(qi, q_index, r * sizeof(int), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(pi, p_index, m * sizeof(int), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(qfd, qf.data(), r * sizeof(double), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(pfd, pf.data(), m * sizeof(double), cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaStreamSynchronize(st));  // the host vectors above die with this frame
  const size_t E = (size_t)B * V;
  SS_TRY(W.Qd.alloc(E * r * sizeof(double), st));
  SS_TRY(W.Pd.alloc(E * m * sizeof(double), st));
  SS_TRY(launch_expand_theta(theta_dev, B, V, k, r, m, qi, pi, qfd, pfd, h, fd, W.Qd.as<double>(), W.Pd.as<double>(), st));
  md.Q = DMat{W.Qd.as<double>(), (long long)r, (long long)B * r, 0, r, r, 1};
  md.P_star = DMat{W.Pd.as<double>(), (long long)m, (long long)B * m, 0, m, m, 1};
  return SSGPU_OK;
}

}  // namespace

int ssgpu_loglik_theta_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, const double* theta,
                                 int k, const int* q_index, const int* p_star_index, int kernel, double* out,
                                 void* stream) {
  SS_ARG(y && theta && out && B >= 1 && V >= 1, "bad arguments");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  ThetaWork W;
  SS_TRY(theta_setup(model, true, B, V, k, theta, q_index, p_star_index, 0.0, 0, W, st));
  return dispatch_batch(W.md, y, B, V, kernel, out, st);
}

int ssgpu_loglik_cov_batch_dev(const double* y, long long B, int V, const ssgpu_model* model, const double* theta, int k,
                               const ssgpu_cov_block* blocks, int n_blocks, double h, int fd, int kernel, double* out,
                               double* grad, void* stream) {
  SS_ARG(y && theta && out && model && blocks && B >= 1 && k >= 1 && n_blocks >= 1, "bad arguments");
  SS_ARG(!fd || (grad && h > 0.0), "fd: grad and a positive step are needed");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  if (fd) V = 1 + 2 * k;
  SS_ARG(V >= 1, "V must be positive");
  DModel md;
  SS_TRY(to_dmodel(model, &md));
  const int m = md.m, r = md.r;
  SS_ARG(md.Q.shared() && md.P_star.shared() && !md.Q.tv() && !md.Q.diag && !md.P_star.diag,
         "Q / P_star of the model must be the shared dense templates");
  SS_ARG(md.a.shared() && md.P_inf.shared() && md.Z.shared() && md.T.shared() && md.R.shared(),
         "theta-driven entry points take shared a, P_inf, Z, T, R");
  bool p_blocks = false;
  for (int i = 0; i < n_blocks; i++) {
    const ssgpu_cov_block& bl = blocks[i];
    SS_ARG(bl.target == 0 || bl.target == 1, "covariance block: target must be 0 (Q) or 1 (P_star)");
    SS_ARG(bl.n >= 1 && bl.n <= 8 && bl.offset >= 0 && bl.offset + bl.n <= (bl.target == 0 ? r : m),
           "covariance block outside its matrix (1 <= n <= 8)");
    SS_ARG(bl.theta_offset >= 0 && bl.theta_offset + bl.n + bl.n * (bl.n - 1) / 2 <= k, "covariance block: parameters out of range");
    p_blocks = p_blocks || bl.target == 1;
  }
  const size_t E = (size_t)B * V;
  DevBuf bd, Qe, Pe, valid, work;
  SS_TRY(bd.alloc(sizeof(ssgpu_cov_block) * n_blocks, st));
  SS_CUDA(cudaMemcpyAsync(bd.get(), blocks, sizeof(ssgpu_cov_block) * n_blocks, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaStreamSynchronize(st));  // `blocks` is the caller's
  SS_TRY(Qe.alloc(E * r * r * sizeof(double), st));
  if (p_blocks) SS_TRY(Pe.alloc(E * m * m * sizeof(double), st));
  SS_TRY(valid.alloc(E * sizeof(int), st));
  SS_TRY(launch_expand_cov(theta, B, V, k, r, m, bd.as<ssgpu_cov_block>(), n_blocks, md.Q.p, md.P_star.p, h, fd,
                           Qe.as<double>(), p_blocks ? Pe.as<double>() : nullptr, valid.as<int>(), st));
  md.Q = DMat{Qe.as<double>(), (long long)r * r, (long long)B * r * r, 0, r, r, 0};
  if (p_blocks) md.P_star = DMat{Pe.as<double>(), (long long)m * m, (long long)B * m * m, 0, m, m, 0};
  double* res = out;
  if (fd) {
    SS_TRY(work.alloc(E * sizeof(double), st));
    res = work.as<double>();
  }
  SS_TRY(dispatch_batch(md, y, B, V, kernel, res, st));
  SS_TRY(launch_mask_invalid(res, valid.as<int>(), (long long)E, st));
  if (fd) SS_TRY(launch_fd_gradient(res, B, k, h, out, grad, st));
  return SSGPU_OK;
}

int ssgpu_loglik_grad_batch_dev(const double* y, long long B, const ssgpu_model* model, const double* theta, int k,
                                const int* q_index, const int* p_star_index, double h, int kernel, double* loglik,
                                double* grad, void* stream) {
  SS_ARG(y && theta && loglik && grad && B >= 1 && h > 0.0, "bad arguments");
  SS_TRY(ensure_device());
  cudaStream_t st = stream ? (cudaStream_t)stream : lib_stream();
  ThetaWork W;
  const int V = 1 + 2 * k;
  SS_TRY(theta_setup(model, true, B, V, k, theta, q_index, p_star_index, h, 1, W, st));
  SS_TRY(W.out.alloc((size_t)B * V * sizeof(double), st));
  SS_TRY(dispatch_batch(W.md, y, B, V, kernel, W.out.as<double>(), st));
  return launch_fd_gradient(W.out.as<double>(), B, k, h, loglik, grad, st);
}

}  // extern "C"

namespace ssgpu {
// ssgpu_loglik_grad_batch for the series [b0, b0 + B) of a batch of ldy series, on the device of the calling thread's
// slot: y points at series b0 of the time-major array (row pitch ldy), theta / loglik / grad at that series' entries.
static int grad_batch_slice(const double* y, long long ldy, long long B, const ssgpu_model* model, const double* theta,
                            int k, const int* q_index, const int* p_star_index, double h, int kernel, double* loglik,
                            double* grad) {
  const long long call_evals = ldy * (1 + 2 * k);  // of the whole call: every range takes the kernel the whole would
  cudaStream_t st = lib_stream();
  const size_t Np = (size_t)model->N * model->p;
  const size_t n_y = Np * (size_t)B;
  DevBuf yd, thd, res;
  SS_TRY(yd.alloc(n_y * sizeof(double), st));
  SS_TRY(thd.alloc((size_t)B * k * sizeof(double), st));
  SS_TRY(res.alloc((size_t)B * (1 + k) * sizeof(double), st));
  const int V = 1 + 2 * k;
  double* ll = res.as<double>();
  double* gr = ll + B;
  ThetaWork W;
  // Pinned y and a large batch: the upload is pipelined with the kernels over chunks of series (chunk c of y is a
  // column block of the time-major array: one strided 2-D copy on a second stream; the kernels of chunk c wait for
  // its event only).  1 GB of y at config 5 otherwise sits in front of the first launch for 18 ms of the call.
  const bool pinned = host_is_pinned(y);
  const int nch = B >= 16384 ? 10 : 1;
  if (getenv("SSGPU_TRACE"))
    fprintf(stderr, "ssgpu_loglik_grad_batch: slot=%d B=%lld of %lld pinned=%d chunks=%d\n", dev_slot(), B, ldy, (int)pinned, nch);
  if (nch == 1) {
    // pageable y: rows of the column block travel through the slot's ring of pinned buffers (xfer.cu)
    SS_TRY(copy_h2d_2d(yd.get(), y, (size_t)ldy * sizeof(double), (size_t)B * sizeof(double), Np, st));
    SS_CUDA(cudaMemcpyAsync(thd.get(), theta, (size_t)B * k * sizeof(double), cudaMemcpyHostToDevice, st));
    SS_TRY(theta_setup(model, false, B, V, k, thd.as<double>(), q_index, p_star_index, h, 1, W, st));
    SS_TRY(W.out.alloc((size_t)B * V * sizeof(double), st));
    SS_TRY(dispatch_batch(W.md, yd.as<double>(), B, V, kernel, W.out.as<double>(), st, nullptr, call_evals));
    SS_TRY(launch_fd_gradient(W.out.as<double>(), B, k, h, ll, gr, st));
  } else {
    // A large batch is cut into chunks of series and the upload of chunk c + 1 overlaps the kernels of chunk c (chunk c of
    // y is a column block of the time-major array).  Pinned y: one strided 2-D copy per chunk on the slot's copy stream,
    // the kernels of the chunk wait for its event only.  Pageable y (what R passes): the chunk's rows are packed into the
    // slot's ring of pinned buffers by helper threads and leave it as contiguous copies (copy_h2d_2d); that call returns
    // when the chunk is on the device, the kernels are enqueued asynchronously, and the host goes on staging the next chunk
    // while they run.  1 GB of y at config 5 otherwise sits in front of the first launch.
    cudaStream_t cs = lib_copy_stream();
    // theta and the small tables first: the one host-to-device copy engine serves both streams in issue order
    SS_CUDA(cudaMemcpyAsync(thd.get(), theta, (size_t)B * k * sizeof(double), cudaMemcpyHostToDevice, st));
    SS_TRY(theta_setup(model, false, B, V, k, thd.as<double>(), q_index, p_star_index, h, 1, W, st));
    SS_TRY(W.out.alloc((size_t)B * V * sizeof(double), st));
    // Pinned y: chunk sizes 1/16, 1/16, 1/8, 1/4, 1/2 of the batch -- the first kernels start after 1/16 of the upload,
    // every later upload (55 GB/s) is shorter than the kernels it hides behind, and most of the work runs in large launches
    // (few partial waves).  Pageable y: the staged upload runs at about the speed of the kernels (1 GB in ~38 ms against
    // 43 ms at config 5), so the call ends one chunk's kernels after the last chunk has landed: 1/16, 1/16 and seven
    // chunks of 1/8 (60.7 -> 56.0 ms per call; with a ring of 4 instead of 8 staging buffers 70.3).
    long long cb[11] = {0};  // chunk c covers the series [cb[c], cb[c + 1])
    {
      long long u = (B / 16 + 31) / 32 * 32, wave_u = 0;
      // Whole waves per chunk: the thread-per-evaluation kernel keeps 256 threads per SM resident, and a chunk that ends
      // with a sliver of a wave pays a whole wave's time for it (1.4 ms at config 5, where 1/16 of the batch is 1.86
      // waves).  The unit becomes the multiple of a wave of series closest to 1/16 of the batch.
      {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            sms > 0) {
          const long long wave = (long long)sms * 256 / V / 32 * 32;  // series per wave, a multiple of 32
          if (wave >= 32 && B / 16 >= wave) {
            u = std::max<long long>(1, (B / 16 + wave / 2) / wave) * wave;
            wave_u = wave;
          }
        }
      }
      // pinned: the very first chunk is half a unit when that is still whole waves -- the kernels start after 1/32 of the upload
      const long long first = (wave_u > 0 && (u / wave_u) % 2 == 0) ? u / 2 : u;
      const long long cut_pinned[6] = {first, u, 2 * u, 4 * u, 8 * u, B};
      const long long cut_paged[9] = {u, 2 * u, 4 * u, 6 * u, 8 * u, 10 * u, 12 * u, 14 * u, B};
      const int nc = pinned ? 6 : 9;
      for (int c = 0; c < nc; c++) cb[c + 1] = std::min<long long>(B, pinned ? cut_pinned[c] : cut_paged[c]);
      for (int c = nc; c < nch; c++) cb[c + 1] = B;
    }
    // events die with the frame on every return path; the copy stream is drained first so that no queued copy
    // still refers to them (or to the caller's y)
    struct Events {
      std::vector<cudaEvent_t> v;
      cudaStream_t drain;
      ~Events() {
        cudaStreamSynchronize(drain);
        for (cudaEvent_t e : v)
          if (e) cudaEventDestroy(e);
      }
    } evs{std::vector<cudaEvent_t>(nch + 1, nullptr), cs};
    cudaEvent_t* ev = evs.v.data();
    cudaEvent_t& ready = evs.v[nch];
    SS_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    SS_CUDA(cudaEventRecord(ready, st));  // yd exists on the compute stream from here on
    SS_CUDA(cudaStreamWaitEvent(cs, ready, 0));
    HostShared hs{model->T.n_slices == 1 ? model->T.data : nullptr, model->R.n_slices == 1 ? model->R.data : nullptr,
                  model->Z.n_slices == 1 ? model->Z.data : nullptr, nullptr};
    if (model->a.stride_series == 0 && model->a.stride_variant == 0) hs.a = model->a.data;  // shared: no device round trip
    auto upload = [&](int c) -> int {
      const long long b0 = cb[c], bn = cb[c + 1] - cb[c];
      SS_CUDA(cudaEventCreateWithFlags(&ev[c], cudaEventDisableTiming));
      if (pinned)
        SS_CUDA(cudaMemcpy2DAsync(yd.as<double>() + Np * b0, bn * sizeof(double), y + b0, (size_t)ldy * sizeof(double),
                                  bn * sizeof(double), Np, cudaMemcpyHostToDevice, cs));
      else
        SS_TRY(copy_h2d_2d(yd.as<double>() + Np * b0, y + b0, (size_t)ldy * sizeof(double), bn * sizeof(double), Np, cs));
      SS_CUDA(cudaEventRecord(ev[c], cs));
      return SSGPU_OK;
    };
    int used = 0;
    for (int c = 0; c < nch && cb[c] < B; c++) used = c + 1;
    if (pinned)  // all copies queued up front: the copy engine never waits for the host
      for (int c = 0; c < used; c++) SS_TRY(upload(c));
    for (int c = 0; c < used; c++) {
      const long long b0 = cb[c], bn = cb[c + 1] - cb[c];
      if (!pinned) SS_TRY(upload(c));
      DModel mdc = W.md;  // per-evaluation variances of the chunk: same strides, shifted base
      mdc.Q.p += b0 * mdc.Q.sb;
      mdc.P_star.p += b0 * mdc.P_star.sb;
      SS_CUDA(cudaStreamWaitEvent(st, ev[c], 0));
      double* outc = W.out.as<double>() + (size_t)V * b0;
      SS_TRY(dispatch_batch(mdc, yd.as<double>() + Np * b0, bn, V, kernel, outc, st, &hs, call_evals));
      SS_TRY(launch_fd_gradient(outc, bn, k, h, ll + b0, gr + b0 * k, st));
    }
  }
  SS_CUDA(cudaMemcpyAsync(loglik, ll, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(grad, gr, (size_t)B * k * sizeof(double), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}
}  // namespace ssgpu

extern "C" {

int ssgpu_loglik_grad_batch(const double* y, long long B, const ssgpu_model* model, const double* theta, int k,
                            const int* q_index, const int* p_star_index, double h, int kernel, double* loglik,
                            double* grad) {
  SS_ARG(y && theta && loglik && grad && B >= 1 && h > 0.0 && model, "bad arguments");
  SS_TRY(ensure_device());
  // Several devices in use (ssgpu_use_devices): contiguous ranges of series, one worker thread, stream and staging ring
  // per device, every device writes its series' results straight into the caller's arrays -- the recursion has no
  // exchange step, so there is nothing to gather.
  if (devices_in_use() > 1 && B >= 64 * devices_in_use())
    return for_each_device(B, 32, [&](int, long long lo, long long hi) {
      return grad_batch_slice(y + lo, B, hi - lo, model, theta + lo * k, k, q_index, p_star_index, h, kernel, loglik + lo,
                              grad + lo * k);
    });
  return grad_batch_slice(y, B, B, model, theta, k, q_index, p_star_index, h, kernel, loglik, grad);
}

}  // extern "C"
