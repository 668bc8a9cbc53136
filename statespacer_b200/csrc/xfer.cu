// Host <-> device copies of large PAGEABLE arrays through pinned staging buffers.
//
// R allocates its arrays with malloc: a cudaMemcpy from / to them is staged by the driver through one small pinned
// buffer at 4-6 GB/s, which is what bounds every host-pointer entry point that returns per-draw or per-series arrays
// (SimSmoother(): 432 MB of results per 12,500 draws).  Here the transfer is cut into chunks that travel through a ring
// of pinned buffers: the DMA of chunk i runs while helper threads memcpy chunks i-1 .. i-7 between the ring and the
// caller's memory.  Pinned (registered) caller memory and small transfers take the plain cudaMemcpyAsync path.
#include <cstring>
#include <future>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ssgpu {

namespace {
constexpr size_t kChunk = 8u << 20;  // 8 MB
#ifndef SSGPU_XFER_RING
#define SSGPU_XFER_RING 8
#endif
constexpr int kRing = SSGPU_XFER_RING;  // slots in flight: one DMA + kRing - 1 helper threads copying between the ring and the caller's memory
constexpr size_t kMinStaged = 4u << 20;

struct Ring {
  void* buf[kRing] = {};
  cudaEvent_t ev[kRing] = {};
  bool ready = false;
  std::mutex mu;  // one staged transfer at a time (the entry points serialise on the library stream anyway)
  int init() {
    if (ready) return SSGPU_OK;
    for (int i = 0; i < kRing; i++) {
      SS_CUDA(cudaHostAlloc(&buf[i], kChunk, cudaHostAllocDefault));
      SS_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    ready = true;
    return SSGPU_OK;
  }
  void release() {
    for (int i = 0; i < kRing; i++) {
      if (buf[i]) cudaFreeHost(buf[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
      buf[i] = nullptr;
      ev[i] = nullptr;
    }
    ready = false;
  }
};
// one ring per device slot: the worker threads of a sharded call stage their slices side by side
Ring g_rings[16];
#define g_ring g_rings[dev_slot()]

bool is_pageable(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

bool host_is_pinned(const void* p) { return !is_pageable(p); }

void xfer_release() {
  for (Ring& r : g_rings) {
    std::lock_guard<std::mutex> lk(r.mu);
    r.release();
  }
}

// dst (host) <- src (device); returns when dst holds the data
int copy_d2h(void* dst, const void* src_dev, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return SSGPU_OK;
  if (bytes < kMinStaged || !is_pageable(dst)) {
    SS_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    return SSGPU_OK;
  }
  std::lock_guard<std::mutex> lk(g_ring.mu);
  SS_TRY(g_ring.init());
  const size_t n = (bytes + kChunk - 1) / kChunk;
  std::vector<std::future<cudaError_t>> done(n);
  cudaError_t err = cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);  // helper threads start on device 0: bind them to the caller's device (one process per GPU)
  for (size_t i = 0; i < n; i++) {
    const int slot = (int)(i % kRing);
    if (i >= (size_t)kRing) {
      const cudaError_t e = done[i - kRing].get();  // the ring slot is free again
      if (err == cudaSuccess) err = e;
    }
    const size_t off = i * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
    cudaError_t e = cudaMemcpyAsync(g_ring.buf[slot], static_cast<const char*>(src_dev) + off, len, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(g_ring.ev[slot], st);
    if (e != cudaSuccess && err == cudaSuccess) err = e;
    void* from = g_ring.buf[slot];
    cudaEvent_t ev = g_ring.ev[slot];
    char* to = static_cast<char*>(dst) + off;
    done[i] = std::async(std::launch::async, [from, to, len, ev, dev]() {
      cudaSetDevice(dev);
      const cudaError_t w = cudaEventSynchronize(ev);
      if (w == cudaSuccess) std::memcpy(to, from, len);
      return w;
    });
  }
  for (size_t i = n > (size_t)kRing ? n - kRing : 0; i < n; i++) {
    const cudaError_t e = done[i].get();
    if (err == cudaSuccess) err = e;
  }
  SS_CUDA(err);
  return SSGPU_OK;
}

// dst (device) <- src (host), enqueued on st.  Pageable src: staged through the pinned ring, src may be re-used when the
// call returns.  Pinned / registered src (and small copies): a plain cudaMemcpyAsync -- src must stay valid and unchanged
// until st is synchronised (every entry point of the library synchronises before it returns to its caller).
int copy_h2d(void* dst_dev, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return SSGPU_OK;
  if (bytes < kMinStaged || !is_pageable(src)) {
    SS_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
    return SSGPU_OK;
  }
  std::lock_guard<std::mutex> lk(g_ring.mu);
  SS_TRY(g_ring.init());
  const size_t n = (bytes + kChunk - 1) / kChunk;
  // chunk i: memcpy into ring slot (helper thread), then DMA from the slot; the slot is free once its DMA has finished
  std::vector<std::future<void>> filled(n);
  auto fill = [&](size_t i) {
    const size_t off = i * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
    void* to = g_ring.buf[i % kRing];
    const char* from = static_cast<const char*>(src) + off;
    filled[i] = std::async(std::launch::async, [to, from, len]() { std::memcpy(to, from, len); });
  };
  for (size_t i = 0; i < n && i < (size_t)kRing; i++) fill(i);
  for (size_t i = 0; i < n; i++) {
    const int slot = (int)(i % kRing);
    const size_t off = i * kChunk, len = bytes - off < kChunk ? bytes - off : kChunk;
    filled[i].get();
    SS_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_dev) + off, g_ring.buf[slot], len, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaEventRecord(g_ring.ev[slot], st));
    if (i + kRing < n) {
      SS_CUDA(cudaEventSynchronize(g_ring.ev[slot]));  // slot drained: refill it with chunk i + kRing
      fill(i + kRing);
    }
  }
  SS_CUDA(cudaStreamSynchronize(st));  // the ring is shared: the next transfer must not overwrite slots in flight
  return SSGPU_OK;
}

// dst (device, dense rows of `width` bytes) <- `rows` rows of the caller's array, `spitch` bytes apart.  Pinned source:
// one strided DMA.  Pageable source: the rows are packed into the ring's pinned slots by helper threads (a slot holds
// as many whole rows as fit) and leave it as contiguous copies.
int copy_h2d_2d(void* dst_dev, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t st) {
  if (rows == 0 || width == 0) return SSGPU_OK;
  if (spitch == width) return copy_h2d(dst_dev, src, width * rows, st);
  if (!is_pageable(src) || width > kChunk) {
    SS_CUDA(cudaMemcpy2DAsync(dst_dev, width, src, spitch, width, rows, cudaMemcpyHostToDevice, st));
    return SSGPU_OK;
  }
  Ring& ring = g_ring;
  std::lock_guard<std::mutex> lk(ring.mu);
  SS_TRY(ring.init());
  const size_t rpc = kChunk / width;  // rows per chunk
  const size_t n = (rows + rpc - 1) / rpc;
  std::vector<std::future<void>> filled(n);
  auto fill = [&](size_t i) {
    const size_t r0 = i * rpc, nr = rows - r0 < rpc ? rows - r0 : rpc;
    char* to = static_cast<char*>(ring.buf[i % kRing]);
    const char* from = static_cast<const char*>(src) + r0 * spitch;
    filled[i] = std::async(std::launch::async, [to, from, nr, width, spitch]() {
      for (size_t r = 0; r < nr; r++) std::memcpy(to + r * width, from + r * spitch, width);
    });
  };
  for (size_t i = 0; i < n && i < (size_t)kRing; i++) fill(i);
  for (size_t i = 0; i < n; i++) {
    const int slot = (int)(i % kRing);
    const size_t r0 = i * rpc, nr = rows - r0 < rpc ? rows - r0 : rpc;
    filled[i].get();
    SS_CUDA(cudaMemcpyAsync(static_cast<char*>(dst_dev) + r0 * width, ring.buf[slot], nr * width, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaEventRecord(ring.ev[slot], st));
    if (i + kRing < n) {
      SS_CUDA(cudaEventSynchronize(ring.ev[slot]));
      fill(i + kRing);
    }
  }
  SS_CUDA(cudaStreamSynchronize(st));
  return SSGPU_OK;
}

}  // namespace ssgpu
