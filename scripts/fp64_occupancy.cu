// FP64 pipe utilisation against resident warps per sub-partition: independent DFMA chains (16 accumulators per
// thread, dependent distance 16 instructions), W warps per SMSP.  nvcc -arch=sm_100a -O3 -o fp64_occ fp64_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, int iters, double x) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, 1.0);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 1024);
  const int iters = 20000;
  for (int w = 1; w <= 8; w *= 2) {  // warps per sub-partition: CTA of 128*w threads, one CTA per SM
    const int threads = 128 * w;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<sms, threads>>>(out, 100, 1.0000001);
    cudaEventRecord(e0);
    k<<<sms, threads>>>(out, iters, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * iters * (double)threads * sms;
    printf("warps/SMSP %d: %.2f TFLOP/s\n", w, flops / (ms * 1e-3) / 1e12);
  }
  // 3 warps per SMSP: 384 threads
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<sms, 384>>>(out, iters, 1.0000001); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("warps/SMSP 3: %.2f TFLOP/s\n", 2.0 * 16 * iters * 384.0 * sms / (ms * 1e-3) / 1e12);
  }
  return 0;
}
