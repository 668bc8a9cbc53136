// Development probe: cost of alternating DMMA.8x8x4 and scalar FP64 on one SM sub-partition's FP64 datapath.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int ND, int NF>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
  double c[4][2], f[8];
  for (int i = 0; i < 4; i++) c[i][0] = c[i][1] = threadIdx.x + i;
  for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 0.5 + i;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ND; j++) dmma(c[j & 3][0], c[j & 3][1], a, b);
#pragma unroll
    for (int j = 0; j < NF; j++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[j & 7]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
  for (int i = 0; i < 8; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ND, int NF>
void run(int warps) {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * 64); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<ND, NF><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  k<ND, NF><<<1, 32 * warps>>>(out, cyc, iters, 1.0000001, 1e-9);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / iters, ideal = (16.0 * ND + 2.0 * NF) * (warps / 4);
  printf("%2d DMMA + %2d DFMA per iteration, %d warps/SMSP: %.1f cycles, ideal %.0f, ratio %.3f\n", ND, NF, warps / 4, per, ideal, per / ideal);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<8, 0>(4); run<0, 32>(4); run<8, 8>(4); run<8, 32>(4); run<1, 1>(4); run<1, 4>(4); run<2, 8>(4); run<14, 40>(4);
  run<8, 0>(16); run<0, 32>(16); run<8, 8>(16); run<8, 32>(16); run<1, 1>(16); run<1, 4>(16); run<2, 8>(16); run<14, 40>(16);
  return 0;
}
