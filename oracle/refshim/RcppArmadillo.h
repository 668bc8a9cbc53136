// Stand-in for <RcppArmadillo.h> (TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT).
//
// R, Rcpp and Armadillo are not installed in this image, so the reference's five native sources
// (/root/reference/src/{LogLikC,KalmanC,FastSmootherC,SimulateC,ExtractComponentC}.cpp) cannot be
// built the way CRAN builds them.  This header implements -- from scratch, eagerly, with naive
// column-major loops -- exactly the subset of the Armadillo / Rcpp API those five files use, so
// that they compile UNMODIFIED from where they lie (see oracle/Makefile) into oracle/_ref/libssref.so.
// That library is (a) a second oracle next to oracle/kalman_np.py and oracle/kalman_c.c and
// (b) the "reference" CPU baseline timed by bench.py.
//
// Semantics notes:
//  * All expressions evaluate eagerly to `mat`; `colvec` / `rowvec` are thin subclasses.
//  * `.row()`, `.col()` on non-const objects return write-through views that are themselves a
//    `mat` holding a copy of the viewed data (reads see the value at view creation -- sufficient for
//    every use in the reference, which never aliases a view across a write to its parent).
//  * `Rcpp::rnorm(n)` pops n values from an injected, thread-local stream (ssref_set_draws), in
//    call order: this is what makes SimulateC reproducible without R's RNG.
//  * Small matrices (<= 256 elements) live in inline storage, no heap traffic.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define NA_REAL (std::numeric_limits<double>::quiet_NaN())

namespace arma {

namespace fill {
struct fill_zeros {};
struct fill_eye {};
static const fill_zeros zeros{};
static const fill_eye eye{};
}  // namespace fill

typedef unsigned long long uword;

// ---- storage with a small inline buffer ---------------------------------------------------
class Store {
 public:
  static const size_t kInline = 256;
  Store() : n_(0), p_(buf_) {}
  explicit Store(size_t n) : n_(0), p_(buf_) { resize(n); }
  Store(const Store& o) : n_(0), p_(buf_) {
    resize(o.n_);
    if (n_) std::memcpy(p_, o.p_, n_ * sizeof(double));
  }
  Store& operator=(const Store& o) {
    if (this != &o) {
      resize(o.n_);
      if (n_) std::memcpy(p_, o.p_, n_ * sizeof(double));
    }
    return *this;
  }
  ~Store() {
    if (p_ != buf_) delete[] p_;
  }
  void resize(size_t n) {
    if (n == n_) return;
    if (p_ != buf_) delete[] p_;
    p_ = n > kInline ? new double[n] : buf_;
    n_ = n;
  }
  double* data() { return p_; }
  const double* data() const { return p_; }
  size_t size() const { return n_; }

 private:
  size_t n_;
  double* p_;
  double buf_[kInline];
};

class mat;
class umat_idx {
 public:
  std::vector<uword> idx;
};

class ElemView {
 public:
  ElemView(mat* m, mat* parent, uword prow, uword pcol, int kind, const umat_idx& ix)
      : m_(m), parent_(parent), prow_(prow), pcol_(pcol), kind_(kind), ix_(ix) {}
  void fill(double v);

 private:
  mat* m_;
  mat* parent_;
  uword prow_, pcol_;
  int kind_;
  umat_idx ix_;
};

class mat {
 public:
  uword n_rows, n_cols, n_elem;
  mat() : n_rows(0), n_cols(0), n_elem(0) {}
  mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), s_(r * c) {}
  mat(uword r, uword c, fill::fill_zeros) : n_rows(r), n_cols(c), n_elem(r * c), s_(r * c) { zeros(); }
  mat(uword r, uword c, fill::fill_eye) : n_rows(r), n_cols(c), n_elem(r * c), s_(r * c) {
    zeros();
    for (uword i = 0; i < std::min(r, c); i++) (*this)(i, i) = 1.0;
  }
  mat(const double* src, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), s_(r * c) {
    if (n_elem) std::memcpy(s_.data(), src, n_elem * sizeof(double));
  }
  mat(double* src, uword r, uword c) : mat(static_cast<const double*>(src), r, c) {}
  virtual ~mat() {}

  void set_size(uword r, uword c) {
    n_rows = r;
    n_cols = c;
    n_elem = r * c;
    s_.resize(n_elem);
  }
  double* memptr() { return s_.data(); }
  const double* memptr() const { return s_.data(); }
  double* begin() { return s_.data(); }
  const double* begin() const { return s_.data(); }
  double& operator()(uword i, uword j) { return s_.data()[i + j * n_rows]; }
  double operator()(uword i, uword j) const { return s_.data()[i + j * n_rows]; }
  double& operator()(uword i) { return s_.data()[i]; }
  double operator()(uword i) const { return s_.data()[i]; }
  virtual void zeros() {
    if (n_elem) std::memset(s_.data(), 0, n_elem * sizeof(double));
  }
  virtual void fill(double v) { std::fill(s_.data(), s_.data() + n_elem, v); }
  virtual void replace(double from, double to) {
    double* d = s_.data();
    if (std::isnan(from)) {
      for (uword i = 0; i < n_elem; i++)
        if (std::isnan(d[i])) d[i] = to;
    } else {
      for (uword i = 0; i < n_elem; i++)
        if (d[i] == from) d[i] = to;
    }
  }
  mat t() const {
    mat o(n_cols, n_rows);
    for (uword j = 0; j < n_cols; j++)
      for (uword i = 0; i < n_rows; i++) o(j, i) = (*this)(i, j);
    return o;
  }
  mat& operator+=(const mat& b) {
    check_same(b, "+=");
    double* d = s_.data();
    const double* e = b.memptr();
    for (uword i = 0; i < n_elem; i++) d[i] += e[i];
    return *this;
  }
  void check_same(const mat& b, const char* what) const {
    if (n_rows != b.n_rows || n_cols != b.n_cols) throw std::logic_error(std::string("size mismatch in ") + what);
  }

  class RowView;
  class ColView;
  RowView row(uword i);
  mat row(uword i) const;
  ColView col(uword j);
  mat col(uword j) const;
  virtual ElemView elem(const umat_idx& ix) { return ElemView(this, nullptr, 0, 0, 0, ix); }

 protected:
  Store s_;
};

// write-through views: a copy of the data plus a pointer to the parent
class mat::RowView : public mat {
 public:
  RowView(mat* parent, uword i) : mat(1, parent->n_cols), parent_(parent), i_(i) {
    for (uword j = 0; j < n_cols; j++) (*this)(0, j) = (*parent)(i, j);
  }
  RowView& operator=(const mat& v) {
    if (v.n_elem != n_cols) throw std::logic_error("row view assignment: size mismatch");
    for (uword j = 0; j < n_cols; j++) {
      (*parent_)(i_, j) = v(j);
      (*this)(0, j) = v(j);
    }
    return *this;
  }
  RowView& operator=(const RowView& v) { return *this = static_cast<const mat&>(v); }
  void fill(double v) override {
    for (uword j = 0; j < n_cols; j++) (*parent_)(i_, j) = (*this)(0, j) = v;
  }
  void zeros() override { fill(0.0); }
  void replace(double from, double to) override {
    mat::replace(from, to);
    for (uword j = 0; j < n_cols; j++) (*parent_)(i_, j) = (*this)(0, j);
  }
  ElemView elem(const umat_idx& ix) override { return ElemView(this, parent_, i_, 0, 1, ix); }

 private:
  mat* parent_;
  uword i_;
};

class mat::ColView : public mat {
 public:
  ColView(mat* parent, uword j) : mat(parent->n_rows, 1), parent_(parent), j_(j) {
    for (uword i = 0; i < n_rows; i++) (*this)(i, 0) = (*parent)(i, j);
  }
  ColView& operator=(const mat& v) {
    if (v.n_elem != n_rows) throw std::logic_error("col view assignment: size mismatch");
    for (uword i = 0; i < n_rows; i++) {
      (*parent_)(i, j_) = v(i);
      (*this)(i, 0) = v(i);
    }
    return *this;
  }
  ColView& operator=(const ColView& v) { return *this = static_cast<const mat&>(v); }
  void fill(double v) override {
    for (uword i = 0; i < n_rows; i++) (*parent_)(i, j_) = (*this)(i, 0) = v;
  }
  void zeros() override { fill(0.0); }
  ElemView elem(const umat_idx& ix) override { return ElemView(this, parent_, 0, j_, 2, ix); }

 private:
  mat* parent_;
  uword j_;
};

inline mat::RowView mat::row(uword i) { return RowView(this, i); }
inline mat mat::row(uword i) const {
  mat o(1, n_cols);
  for (uword j = 0; j < n_cols; j++) o(0, j) = (*this)(i, j);
  return o;
}
inline mat::ColView mat::col(uword j) { return ColView(this, j); }
inline mat mat::col(uword j) const { return mat(memptr() + j * n_rows, n_rows, 1); }

inline void ElemView::fill(double v) {
  for (uword k : ix_.idx) {
    (*m_)(k) = v;
    if (kind_ == 1) (*parent_)(prow_, k) = v;
    if (kind_ == 2) (*parent_)(k, pcol_) = v;
  }
}

class colvec : public mat {
 public:
  colvec() : mat() {}
  explicit colvec(uword n) : mat(n, 1) {}
  colvec(uword n, fill::fill_zeros z) : mat(n, 1, z) {}
  colvec(const double* src, uword n) : mat(src, n, 1) {}
  colvec(const mat& m) : mat(m) {
    if (m.n_cols != 1 && m.n_elem != 0) throw std::logic_error("colvec from non-column matrix");
  }
};
class rowvec : public mat {
 public:
  rowvec() : mat() {}
  explicit rowvec(uword n) : mat(1, n) {}
  rowvec(uword n, fill::fill_zeros z) : mat(1, n, z) {}
  rowvec(const mat& m) : mat(m) {
    if (m.n_rows != 1 && m.n_elem != 0) throw std::logic_error("rowvec from non-row matrix");
  }
};
typedef colvec vec;

// ---- arithmetic ---------------------------------------------------------------------------
inline mat operator*(const mat& a, const mat& b) {
  if (a.n_cols != b.n_rows) throw std::logic_error("matrix multiplication: incompatible dimensions");
  mat c(a.n_rows, b.n_cols, fill::zeros);
  const uword M = a.n_rows, K = a.n_cols, Nn = b.n_cols;
  const double* A = a.memptr();
  const double* B = b.memptr();
  double* C = c.memptr();
  for (uword j = 0; j < Nn; j++)
    for (uword k = 0; k < K; k++) {
      const double bkj = B[k + j * K];
      const double* Ak = A + k * M;
      double* Cj = C + j * M;
      for (uword i = 0; i < M; i++) Cj[i] += Ak[i] * bkj;
    }
  return c;
}
inline mat operator*(const mat& a, double s) {
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = a(i) * s;
  return c;
}
inline mat operator*(double s, const mat& a) { return a * s; }
inline mat operator+(const mat& a, const mat& b) {
  a.check_same(b, "addition");
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = a(i) + b(i);
  return c;
}
inline mat operator-(const mat& a, const mat& b) {
  a.check_same(b, "subtraction");
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = a(i) - b(i);
  return c;
}
inline mat operator-(const mat& a) {
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = -a(i);
  return c;
}
inline mat operator/(const mat& a, const mat& b) {
  a.check_same(b, "element-wise division");
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = a(i) / b(i);
  return c;
}
inline mat operator<(const mat& a, double s) {
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = a(i) < s ? 1.0 : 0.0;
  return c;
}

inline mat abs(const mat& a) {
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = std::fabs(a(i));
  return c;
}
inline mat sqrt(const mat& a) {
  mat c(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) c(i) = std::sqrt(a(i));
  return c;
}
inline mat vectorise(const mat& a) { return mat(a.memptr(), a.n_elem, 1); }
inline bool all(const mat& a) {
  for (uword i = 0; i < a.n_elem; i++)
    if (a(i) == 0.0) return false;
  return true;
}
inline double as_scalar(const mat& a) {
  if (a.n_elem != 1) throw std::logic_error("as_scalar(): expression must evaluate to 1 element");
  return a(0);
}
inline mat trans(const mat& a) { return a.t(); }
inline mat repmat(const mat& a, uword r, uword c) {
  mat o(a.n_rows * r, a.n_cols * c);
  for (uword cj = 0; cj < c; cj++)
    for (uword ri = 0; ri < r; ri++)
      for (uword j = 0; j < a.n_cols; j++)
        for (uword i = 0; i < a.n_rows; i++) o(ri * a.n_rows + i, cj * a.n_cols + j) = a(i, j);
  return o;
}
inline mat diagmat(const mat& v) {
  if (v.n_rows != 1 && v.n_cols != 1) {  // matrix input: keep the diagonal
    mat o(v.n_rows, v.n_cols, fill::zeros);
    for (uword i = 0; i < std::min(v.n_rows, v.n_cols); i++) o(i, i) = v(i, i);
    return o;
  }
  mat o(v.n_elem, v.n_elem, fill::zeros);
  for (uword i = 0; i < v.n_elem; i++) o(i, i) = v(i);
  return o;
}
inline colvec diagvec(const mat& a) {
  uword n = std::min(a.n_rows, a.n_cols);
  colvec o(n);
  for (uword i = 0; i < n; i++) o(i) = a(i, i);
  return o;
}
inline umat_idx find_nonfinite(const mat& a) {
  umat_idx r;
  for (uword i = 0; i < a.n_elem; i++)
    if (!std::isfinite(a(i))) r.idx.push_back(i);
  return r;
}

// inverse by Gauss-Jordan elimination with partial pivoting (stands in for LAPACK dgetrf/dgetri)
inline mat inv(const mat& a) {
  if (a.n_rows != a.n_cols) throw std::logic_error("inv(): matrix must be square");
  const uword n = a.n_rows;
  mat w(a), o(n, n, fill::eye);
  for (uword c = 0; c < n; c++) {
    uword piv = c;
    for (uword i = c + 1; i < n; i++)
      if (std::fabs(w(i, c)) > std::fabs(w(piv, c))) piv = i;
    if (w(piv, c) == 0.0 || !std::isfinite(w(piv, c))) throw std::runtime_error("inv(): matrix is singular");
    if (piv != c)
      for (uword j = 0; j < n; j++) {
        std::swap(w(c, j), w(piv, j));
        std::swap(o(c, j), o(piv, j));
      }
    const double d = 1.0 / w(c, c);
    for (uword j = 0; j < n; j++) {
      w(c, j) *= d;
      o(c, j) *= d;
    }
    for (uword i = 0; i < n; i++) {
      if (i == c) continue;
      const double f = w(i, c);
      if (f == 0.0) continue;
      for (uword j = 0; j < n; j++) {
        w(i, j) -= f * w(c, j);
        o(i, j) -= f * o(c, j);
      }
    }
  }
  return o;
}

// one-sided Jacobi (Hestenes) SVD, singular values sorted descending (stands in for LAPACK dgesdd)
inline bool svd(mat& U, colvec& s, mat& V, const mat& X) {
  const uword m = X.n_rows, n = X.n_cols;
  if (m != n) throw std::logic_error("svd(): this stand-in supports square matrices only");
  mat A(X);
  V = mat(n, n, fill::eye);
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0;
    for (uword p = 0; p + 1 < n; p++)
      for (uword q = p + 1; q < n; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (uword i = 0; i < m; i++) {
          alpha += A(i, p) * A(i, p);
          beta += A(i, q) * A(i, q);
          gamma += A(i, p) * A(i, q);
        }
        if (gamma == 0.0) continue;
        off = std::max(off, std::fabs(gamma) / std::sqrt(alpha * beta + 1e-300));
        if (std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
        for (uword i = 0; i < m; i++) {
          const double ap = A(i, p), aq = A(i, q);
          A(i, p) = c * ap - sn * aq;
          A(i, q) = sn * ap + c * aq;
        }
        for (uword i = 0; i < n; i++) {
          const double vp = V(i, p), vq = V(i, q);
          V(i, p) = c * vp - sn * vq;
          V(i, q) = sn * vp + c * vq;
        }
      }
    if (off < 1e-15) break;
  }
  std::vector<double> sv(n);
  for (uword j = 0; j < n; j++) {
    double nn = 0;
    for (uword i = 0; i < m; i++) nn += A(i, j) * A(i, j);
    sv[j] = std::sqrt(nn);
  }
  std::vector<uword> order(n);
  for (uword j = 0; j < n; j++) order[j] = j;
  std::stable_sort(order.begin(), order.end(), [&](uword x, uword y) { return sv[x] > sv[y]; });
  U = mat(m, n, fill::zeros);
  mat Vs(n, n);
  s = colvec(n);
  for (uword k = 0; k < n; k++) {
    const uword j = order[k];
    s(k) = sv[j];
    for (uword i = 0; i < n; i++) Vs(i, k) = V(i, j);
    if (sv[j] > 0)
      for (uword i = 0; i < m; i++) U(i, k) = A(i, j) / sv[j];
  }
  // complete U for zero singular values with an orthonormal basis (Gram-Schmidt on unit vectors)
  for (uword k = 0; k < n; k++) {
    if (s(k) > 0) continue;
    for (uword e = 0; e < m; e++) {
      std::vector<double> c(m, 0.0);
      c[e] = 1.0;
      for (uword q = 0; q < n; q++) {
        if (q == k) continue;
        double dot = 0;
        for (uword i = 0; i < m; i++) dot += U(i, q) * c[i];
        for (uword i = 0; i < m; i++) c[i] -= dot * U(i, q);
      }
      double nn = 0;
      for (uword i = 0; i < m; i++) nn += c[i] * c[i];
      if (nn > 1e-8) {
        nn = std::sqrt(nn);
        for (uword i = 0; i < m; i++) U(i, k) = c[i] / nn;
        break;
      }
    }
  }
  V = Vs;
  return true;
}

// ---- cube ---------------------------------------------------------------------------------
class cube {
 public:
  uword n_rows, n_cols, n_slices;
  cube() : n_rows(0), n_cols(0), n_slices(0) {}
  cube(uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s), sl_(s, mat(r, c)) {}
  cube(uword r, uword c, uword s, fill::fill_zeros z) : n_rows(r), n_cols(c), n_slices(s), sl_(s, mat(r, c, z)) {}
  cube(const double* src, uword r, uword c, uword s) : n_rows(r), n_cols(c), n_slices(s) {
    sl_.reserve(s);
    for (uword k = 0; k < s; k++) sl_.emplace_back(src + k * r * c, r, c);
  }
  mat& slice(uword k) { return sl_[k]; }
  const mat& slice(uword k) const { return sl_[k]; }
  double& operator()(uword i, uword j, uword k) { return sl_[k](i, j); }
  double operator()(uword i, uword j, uword k) const { return sl_[k](i, j); }

  class RowView {
   public:
    RowView(cube* c, uword i) : c_(c), i_(i) {}
    // subview_cube (1 x n_cols x n_slices) = Mat (n_cols x n_slices)
    RowView& operator=(const mat& v) {
      if (v.n_rows != c_->n_cols || v.n_cols != c_->n_slices)
        throw std::logic_error("cube row assignment: size mismatch");
      for (uword k = 0; k < c_->n_slices; k++)
        for (uword j = 0; j < c_->n_cols; j++) c_->sl_[k](i_, j) = v(j, k);
      return *this;
    }

   private:
    cube* c_;
    uword i_;
  };
  RowView row(uword i) { return RowView(this, i); }
  void copy_out(double* dst) const {
    for (uword k = 0; k < n_slices; k++)
      std::memcpy(dst + k * n_rows * n_cols, sl_[k].memptr(), n_rows * n_cols * sizeof(double));
  }

 private:
  std::vector<mat> sl_;
};

}  // namespace arma

// ---- Rcpp ---------------------------------------------------------------------------------
namespace Rcpp {

class NumericVector {
 public:
  NumericVector() {}
  explicit NumericVector(long n) : v_(n, 0.0) {}
  NumericVector(const std::vector<double>& v) : v_(v) {}
  double& operator()(long i) { return v_[i]; }
  double operator()(long i) const { return v_[i]; }
  double& operator[](long i) { return v_[i]; }
  double* begin() { return v_.data(); }
  const double* begin() const { return v_.data(); }
  long size() const { return (long)v_.size(); }
  const std::vector<double>& std() const { return v_; }

 private:
  std::vector<double> v_;
};

class NumericMatrix {
 public:
  NumericMatrix(const double* p, int r, int c) : p_(p), r_(r), c_(c) {}
  int nrow() const { return r_; }
  int ncol() const { return c_; }
  double operator()(int i, int j) const { return p_[i + (long)j * r_]; }

 private:
  const double* p_;
  int r_, c_;
};

class LogicalMatrix {
 public:
  LogicalMatrix(const int* p, int r, int c) : p_(p), r_(r), c_(c) {}
  int nrow() const { return r_; }
  int ncol() const { return c_; }
  int operator()(int i, int j) const { return p_[i + (long)j * r_]; }

 private:
  const int* p_;
  int r_, c_;
};

// injected normal variates (thread-local so the OpenMP batch driver stays safe)
struct DrawStream {
  const double* p = nullptr;
  long n = 0, pos = 0;
};
inline DrawStream& draw_stream() {
  static thread_local DrawStream s;
  return s;
}
inline NumericVector rnorm(long n) {
  DrawStream& s = draw_stream();
  if (s.pos + n > s.n) throw std::runtime_error("Rcpp::rnorm stand-in: injected draw stream exhausted");
  NumericVector v(n);
  std::memcpy(v.begin(), s.p + s.pos, n * sizeof(double));
  s.pos += n;
  return v;
}

class List;
struct Value {
  enum Kind { NONE, INT, VEC, MAT, CUBE, LIST } kind = NONE;
  int i = 0;
  std::vector<double> vec;
  arma::mat m;
  arma::cube c;
  std::shared_ptr<List> l;
  Value() {}
  Value(int x) : kind(INT), i(x) {}
  Value(const NumericVector& x) : kind(VEC), vec(x.std()) {}
  Value(const arma::mat& x) : kind(MAT), m(x) {}
  Value(const arma::cube& x) : kind(CUBE), c(x) {}
  Value(const List& x);
};
struct NamedValue {
  std::string name;
  Value v;
};
struct Named {
  std::string name;
  explicit Named(const char* n) : name(n) {}
  template <class T>
  NamedValue operator=(const T& x) const {
    return NamedValue{name, Value(x)};
  }
};
class List {
 public:
  std::map<std::string, Value> items;
  Value& operator[](const char* k) { return items[k]; }
  template <class... A>
  static List create(const A&... a) {
    List l;
    (void)std::initializer_list<int>{(l.items[a.name] = a.v, 0)...};
    return l;
  }
  const Value& at(const char* k) const {
    auto it = items.find(k);
    if (it == items.end()) throw std::runtime_error(std::string("List: no element ") + k);
    return it->second;
  }
};
inline Value::Value(const List& x) : kind(LIST), l(std::make_shared<List>(x)) {}

}  // namespace Rcpp
