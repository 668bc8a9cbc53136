// ssgpu_fp64_peak: measures the FP64 roofline denominator on the device the library runs on.
// MEASURED_PEAKS.json carries HBM and bf16 figures only, so the FP64
// FMA-pipe peak is measured with register-resident, dependence-free DFMA chains (mode 0), the legacy
// FP64 tensor path mma.sync.m8n8k4.f64 (mode 1) or both interleaved (mode 2).
#include "common.cuh"

namespace ssgpu {

constexpr int kPeakThreads = 256;
constexpr int kPeakChains = 16;

template <int MODE>
__global__ void __launch_bounds__(kPeakThreads) fp64_peak_kernel(int iters, double seed, double* sink) {
  double acc[kPeakChains];
  double c0 = 0.0, c1 = 0.0;
  const double x = seed + threadIdx.x * 1e-9, yv = 1.0 - 1e-9 * (blockIdx.x & 7);
#pragma unroll
  for (int i = 0; i < kPeakChains; i++) acc[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0 || MODE == 2) {
#pragma unroll
        for (int i = 0; i < kPeakChains; i++) acc[i] = fma(acc[i], yv, x);
      }
      if (MODE == 1 || MODE == 2) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c0), "+d"(c1)
                       : "d"(x), "d"(yv));
        }
      }
    }
  }
  double s = c0 + c1;
#pragma unroll
  for (int i = 0; i < kPeakChains; i++) s += acc[i];
  if (s == 123.456) sink[0] = s;  // keeps the chains alive without a store on the timed path
}

}  // namespace ssgpu

using namespace ssgpu;

extern "C" int ssgpu_fp64_peak(int mode, int iters, int repeats, double* tflops, double* ms) {
  SS_ARG(mode >= 0 && mode <= 2 && iters > 0 && repeats > 0 && tflops, "fp64_peak: bad arguments");
  SS_TRY(ensure_device());
  cudaStream_t st = lib_stream();
  int dev = 0, sms = 0;
  SS_CUDA(cudaGetDevice(&dev));
  SS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = sms * 8;
  DevBuf sink;
  SS_TRY(sink.alloc(8, st));
  cudaEvent_t e0, e1;
  SS_CUDA(cudaEventCreate(&e0));
  SS_CUDA(cudaEventCreate(&e1));
  auto launch = [&]() {
    if (mode == 0)
      fp64_peak_kernel<0><<<grid, kPeakThreads, 0, st>>>(iters, 0.5, sink.as<double>());
    else if (mode == 1)
      fp64_peak_kernel<1><<<grid, kPeakThreads, 0, st>>>(iters, 0.5, sink.as<double>());
    else
      fp64_peak_kernel<2><<<grid, kPeakThreads, 0, st>>>(iters, 0.5, sink.as<double>());
    count_launch();
  };
  launch();  // warm-up
  SS_CUDA(cudaStreamSynchronize(st));
  float best = 1e30f;
  for (int rep = 0; rep < repeats; rep++) {
    SS_CUDA(cudaEventRecord(e0, st));
    launch();
    SS_CUDA(cudaEventRecord(e1, st));
    SS_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    SS_CUDA(cudaEventElapsedTime(&t, e0, e1));
    if (t < best) best = t;
  }
  SS_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double threads = (double)grid * kPeakThreads;
  double flops = 0.0;
  if (mode == 0 || mode == 2) flops += threads * iters * 16.0 * kPeakChains * 2.0;
  if (mode == 1 || mode == 2) flops += threads / 32.0 * iters * 16.0 * 4.0 * 512.0;  // 2*8*8*4 per warp mma
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms) *ms = best;
  return SSGPU_OK;
}
