// Development probe: dependent-issue latency of DFMA, DADD, DMMA.8x8x4, SHFL (64-bit) and MUFU.RCP64H on sm_100,
// and DMMA issue interval per SM sub-partition.  nvcc -arch=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double c0 = threadIdx.x, c1 = 1.0, d0 = 2.0, d1 = 3.0, e0 = 4, e1 = 5, f0 = 6, f1 = 7;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {  // dependent DFMA
#pragma unroll
      for (int j = 0; j < 16; j++) c0 = fma(c0, a, b);
    } else if (MODE == 1) {  // dependent DADD
#pragma unroll
      for (int j = 0; j < 16; j++) c0 = c0 + a;
    } else if (MODE == 2) {  // dependent DMMA
#pragma unroll
      for (int j = 0; j < 16; j++) dmma(c0, c1, a, b);
    } else if (MODE == 3) {  // 4 independent DMMA chains
#pragma unroll
      for (int j = 0; j < 4; j++) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); }
    } else if (MODE == 4) {  // dependent 64-bit shuffle
#pragma unroll
      for (int j = 0; j < 16; j++) c0 = __shfl_xor_sync(0xffffffffu, c0, 4);
    } else if (MODE == 5) {  // shuffle + DADD (one reduction round)
#pragma unroll
      for (int j = 0; j < 16; j++) c0 = c0 + __shfl_xor_sync(0xffffffffu, c0, 4);
    } else if (MODE == 6) {  // DMMA whose A operand depends on the previous result (product chain)
#pragma unroll
      for (int j = 0; j < 16; j++) { double x0 = 0, x1 = 0; dmma(x0, x1, c0, b); c0 = x0; c1 = x1; }
    } else if (MODE == 7) {  // DFMA dependent on DMMA result and back
#pragma unroll
      for (int j = 0; j < 8; j++) { dmma(c0, c1, a, b); c0 = fma(c0, a, b); }
    } else if (MODE == 8) {  // 2 independent DFMA chains
#pragma unroll
      for (int j = 0; j < 8; j++) { c0 = fma(c0, a, b); d0 = fma(d0, a, b); }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps, int blocks, int per_iter) {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * 1024); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<blocks, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  k<MODE><<<blocks, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/CTA %2d CTAs %4d : %.2f cycles per op (per warp)\n", name, warps, blocks, (double)h / iters / per_iter);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("dependent DFMA", 1, 1, 16);
  run<1>("dependent DADD", 1, 1, 16);
  run<8>("2 independent DFMA chains (per op)", 1, 1, 16);
  run<2>("dependent DMMA (accumulator chain)", 1, 1, 16);
  run<6>("dependent DMMA (through the A operand)", 1, 1, 16);
  run<3>("4 independent DMMA chains (per DMMA)", 1, 1, 16);
  run<7>("DMMA -> DFMA -> DMMA (per pair)", 1, 1, 8);
  run<4>("dependent SHFL.64", 1, 1, 16);
  run<5>("SHFL.64 + DADD round", 1, 1, 16);
  // throughput: 4 warps per CTA = 1 per sub-partition, then 16 warps/SM
  run<3>("DMMA throughput, 1 warp per SMSP", 4, 1, 16);
  run<3>("DMMA throughput, 4 warps per SMSP", 16, 1, 16);
  run<0>("DFMA dependent, 4 warps per SMSP", 16, 1, 16);
  run<8>("DFMA 2 chains, 4 warps per SMSP", 16, 1, 16);
  run<2>("dependent DMMA, 4 warps per SMSP", 16, 1, 16);
  run<7>("DMMA->DFMA pairs, 4 warps per SMSP", 16, 1, 8);
  return 0;
}
