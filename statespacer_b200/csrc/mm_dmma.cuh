// m x m matrix products of the generic (one CTA per evaluation) kernels on the FP64 tensor path.
//
// C = op(A) op(B) (+ add) with every matrix column-major (ld = m) in shared memory.  Output tiles (8 x 8) are
// dealt round-robin to the warps of the CTA; each tile is one chain of DMMA.8x8x4 over k = 0, 4, 8, ...  A DMMA is an
// ascending fma chain starting from its accumulator (scripts/dmma_semantics.cu: 12,800 of 12,800 results bit-equal),
// so the chain reproduces  acc = fma(a_k, b_k, acc), k ascending  -- the pinned order of oracle/kalman_c.c -- bit
// for bit; entries beyond m enter as fma(0, 0, acc) == acc.  Against the element-parallel loops these replace
// (two 8-byte shared-memory loads per fma: bandwidth-bound) a DMMA reads 2 operands per lane for 8 fma per lane.
#pragma once

namespace ssgpu {

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// q / d for 0 <= q < 4096, 1 <= d <= 64 without the ~20-instruction integer division (the generic kernels are
// issue-bound: profiles/README.md): (q + 0.5) / d stays 1 / 128 away from every integer, far more than the error of
// the approximate single-precision quotient
__device__ __forceinline__ int div_small(int q, int d) { return __float2int_rz(__fdividef((float)q + 0.5f, (float)d)); }

// TA / TB: use the transpose of A / B.  All threads of the CTA must call (NT a multiple of 32); no barrier inside.
// SUB: C = add - A B (add required) instead of C = A B + add.
template <bool TA, bool TB, bool SUB = false>
__device__ __forceinline__ void mm_dmma(double* C, const double* A, const double* Bm, const double* add, int m, int tid,
                                        int NT) {
  const int warp = tid >> 5, nw = NT >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int TT = (m + 7) >> 3;
  for (int tile = warp; tile < TT * TT; tile += nw) {
    const int nt = div_small(tile, TT), rt = tile - nt * TT;
    const int i = 8 * rt + g, n = 8 * nt + g;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < m; k0 += 4) {
      const int k = k0 + q;
      const bool kin = k < m;
      const double a = (kin && i < m) ? (TA ? A[k + i * m] : A[i + k * m]) : 0.0;
      const double b = (kin && n < m) ? (TB ? Bm[n + k * m] : Bm[k + n * m]) : 0.0;
      dmma_884(c0, c1, a, b);
    }
    const int j0 = 8 * nt + 2 * q;
    if (i < m) {
      if constexpr (SUB) {
        if (j0 < m) C[i + j0 * m] = add[i + j0 * m] - c0;
        if (j0 + 1 < m) C[i + (j0 + 1) * m] = add[i + (j0 + 1) * m] - c1;
      } else {
        if (j0 < m) C[i + j0 * m] = add ? c0 + add[i + j0 * m] : c0;
        if (j0 + 1 < m) C[i + (j0 + 1) * m] = add ? c1 + add[i + (j0 + 1) * m] : c1;
      }
    }
  }
}

// y1 = A1 x (and y2 = A2 x when A2 != nullptr): matrix-vector products as DMMA chains with x in column 0 of the B
// operand (x[k] = xs[k * xstride]) -- the same ascending fma chain per row as the scalar loop, but spread over
// the warps of the CTA instead of m threads.
__device__ __forceinline__ void mv_dmma2(double* y1, const double* A1, double* y2, const double* A2, const double* xs,
                                         int xstride, int m, int tid, int NT) {
  const int warp = tid >> 5, nw = NT >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int TT = (m + 7) >> 3, n_tiles = A2 ? 2 * TT : TT;
  for (int tile = warp; tile < n_tiles; tile += nw) {
    const double* A = tile < TT ? A1 : A2;
    double* y = tile < TT ? y1 : y2;
    const int rt = tile < TT ? tile : tile - TT;
    const int i = 8 * rt + g;
    double c0 = 0.0, c1 = 0.0;
    for (int k0 = 0; k0 < m; k0 += 4) {
      const int k = k0 + q;
      const bool kin = k < m;
      const double a = (kin && i < m) ? A[i + k * m] : 0.0;
      const double b = (kin && g == 0) ? xs[k * xstride] : 0.0;
      dmma_884(c0, c1, a, b);
    }
    if (q == 0 && i < m) y[i] = c0;
  }
}

}  // namespace ssgpu
