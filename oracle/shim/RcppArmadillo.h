// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Minimal stand-in for <RcppArmadillo.h> so that the reference's native covariance
// loops (/root/reference/src/cf.cpp, compiled where it lies, never copied) build
// without R / Rcpp / Armadillo.  Only the handful of arma:: constructs that file
// uses are provided: a column-major dense matrix, a row proxy convertible to/from
// a 1 x n_cols matrix, element-wise -, % (Hadamard), scalar *, square() and sum().
// Semantics follow Armadillo's documented behaviour for these operations; the
// arithmetic order inside sum() is a plain left-to-right accumulation, as in
// Armadillo's scalar fallback for short rows.
#ifndef GPLITE_B200_ORACLE_ARMA_SHIM_H
#define GPLITE_B200_ORACLE_ARMA_SHIM_H

#include <cstddef>
#include <vector>
#include <algorithm>

namespace arma {

typedef unsigned long long uword;

struct mat;

struct rowref {
  mat* parent;
  uword r;
  rowref(mat* p, uword r_) : parent(p), r(r_) {}
  operator mat() const;
  rowref& operator=(const mat& m);
};

struct mat {
  uword n_rows, n_cols;
  std::vector<double> mem;
  mat() : n_rows(0), n_cols(0) {}
  mat(uword r, uword c) : n_rows(r), n_cols(c), mem((size_t)r * c, 0.0) {}
  double& operator()(uword i, uword j) { return mem[(size_t)j * n_rows + i]; }
  const double& operator()(uword i, uword j) const { return mem[(size_t)j * n_rows + i]; }
  rowref row(uword i) { return rowref(this, i); }
};

inline rowref::operator mat() const {
  mat out(1, parent->n_cols);
  for (uword j = 0; j < parent->n_cols; ++j) out(0, j) = (*parent)(r, j);
  return out;
}

inline rowref& rowref::operator=(const mat& m) {
  for (uword j = 0; j < parent->n_cols; ++j) (*parent)(r, j) = m.mem[j];
  return *this;
}

inline mat operator-(const mat& a, const mat& b) {
  mat out(a.n_rows, a.n_cols);
  for (size_t k = 0; k < out.mem.size(); ++k) out.mem[k] = a.mem[k] - b.mem[k];
  return out;
}

inline mat operator%(const mat& a, const mat& b) {
  mat out(a.n_rows, a.n_cols);
  for (size_t k = 0; k < out.mem.size(); ++k) out.mem[k] = a.mem[k] * b.mem[k];
  return out;
}

inline mat operator*(const mat& a, double s) {
  mat out(a.n_rows, a.n_cols);
  for (size_t k = 0; k < out.mem.size(); ++k) out.mem[k] = a.mem[k] * s;
  return out;
}

inline mat square(const mat& a) {
  mat out(a.n_rows, a.n_cols);
  for (size_t k = 0; k < out.mem.size(); ++k) out.mem[k] = a.mem[k] * a.mem[k];
  return out;
}

inline double sum(const mat& a) {
  double s = 0.0;
  for (size_t k = 0; k < a.mem.size(); ++k) s += a.mem[k];
  return s;
}

}  // namespace arma

#endif
