// State record the lane kernel (loglik_blk.cu) leaves behind when an evaluation has come out of the exact diffuse
// phase, and the thread kernel (loglik_thr.cu) continues from: the regular phase of src/LogLikC.cpp:162-207 needs
// a, P_star, the hoisted diagonal blocks of R Q R' and the loglikelihood accumulators, nothing of P_inf.
// Doubles are stored slot-major, hd[slot * E + w] (w = evaluation number of the launch, series-major), so that the
// 32 evaluations of a warp of the thread kernel read 256 contiguous bytes per slot; ints likewise in hi.
#pragma once

namespace ssgpu {

struct BlkHandoff {
  int nb, qw;  // blocks; doubles of R Q R' kept per diagonal block: 2 (q00, q11) or 3 (q00, q01, q11)
  __host__ __device__ constexpr BlkHandoff(int nb_, bool qoff) : nb(nb_), qw(qoff ? 3 : 2) {}
  __host__ __device__ constexpr int noff() const { return nb * (nb - 1) / 2; }
  __host__ __device__ constexpr int a() const { return 0; }                  // a: 2 per block row
  __host__ __device__ constexpr int q() const { return 2 * nb; }             // qw per block row
  __host__ __device__ constexpr int d() const { return q() + qw * nb; }      // diagonal blocks: p00, p01, p11
  __host__ __device__ constexpr int o() const { return d() + 3 * nb; }       // blocks above the diagonal, row-major, 4 each
  __host__ __device__ constexpr int s() const { return o() + 4 * noff(); }   // prod, sumq (Sum log F as mantissa, Sum v^2/F)
  __host__ __device__ constexpr int doubles() const { return s() + 2; }
  // position of block (I, J), I < J, among the blocks above the diagonal
  __host__ __device__ constexpr int off_index(int I, int J) const { return I * nb - I * (I + 1) / 2 + (J - I - 1); }
};
// hi: t0 (first time point of the regular phase; N = the lane kernel has finished the evaluation), esum, n_terms,
// non_init, F_eq0
constexpr int kBlkHandoffInts = 5;

}  // namespace ssgpu
