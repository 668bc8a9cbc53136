// Which IEEE operation sequence does DMMA.8x8x4 perform?  Compares mma.sync.m8n8k4.f64 with candidate fma
// chains on random data (development probe; result recorded in DESIGN.md).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, const double* C, double* D) {
  const int lane = threadIdx.x, g = lane >> 2, q = lane & 3;
  double c0 = C[g * 8 + 2 * q], c1 = C[g * 8 + 2 * q + 1];
  const double a = A[g * 4 + q];   // A[g][q] row-major 8x4
  const double b = B[q * 8 + g];   // B[q][g] 4x8
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  D[g * 8 + 2 * q] = c0;
  D[g * 8 + 2 * q + 1] = c1;
}
int main() {
  double hA[32], hB[32], hC[64], hD[64];
  double *A, *B, *C, *D;
  cudaMalloc(&A, 256); cudaMalloc(&B, 256); cudaMalloc(&C, 512); cudaMalloc(&D, 512);
  int cnt[6] = {0, 0, 0, 0, 0, 0}, total = 0;
  srand(1);
  for (int trial = 0; trial < 200; trial++) {
    for (int i = 0; i < 32; i++) { hA[i] = (rand() / (double)RAND_MAX - 0.5) * pow(10.0, rand() % 7 - 3); hB[i] = (rand() / (double)RAND_MAX - 0.5) * pow(10.0, rand() % 7 - 3); }
    for (int i = 0; i < 64; i++) hC[i] = (rand() / (double)RAND_MAX - 0.5) * pow(10.0, rand() % 7 - 3);
    cudaMemcpy(A, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, 256, cudaMemcpyHostToDevice); cudaMemcpy(C, hC, 512, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(A, B, C, D);
    cudaMemcpy(hD, D, 512, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) {
      double a[4], b[4]; for (int kk = 0; kk < 4; kk++) { a[kk] = hA[i * 4 + kk]; b[kk] = hB[kk * 8 + j]; }
      const double c = hC[i * 8 + j], d = hD[i * 8 + j];
      double f0 = fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], fma(a[0], b[0], c))));      // k ascending from c
      double f1 = fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], c))));      // k descending from c
      double f2 = c + fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])));           // dot then add
      double f3 = fma(a[1], b[1], fma(a[0], b[0], c)); f3 = fma(a[3], b[3], fma(a[2], b[2], f3));
      long double e = (long double)c; for (int kk = 0; kk < 4; kk++) e += (long double)a[kk] * b[kk];
      double f4 = (double)e;                                                                    // (almost) exact sum rounded once
      double f5 = (fma(a[0], b[0], c) + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]);
      double f[6] = {f0, f1, f2, f3, f4, f5};
      for (int t = 0; t < 6; t++) cnt[t] += (f[t] == d);
      total++;
    }
  }
  printf("total %d  asc-from-c %d  desc-from-c %d  dot-then-add %d  asc(dup) %d  exact-rounded-once %d  pairwise %d\n", total, cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], cnt[5]);
  return 0;
}
