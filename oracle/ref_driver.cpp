// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" forwarding layer over the reference's four native covariance loops
// (declared here, defined in /root/reference/src/cf.cpp:9,48,86,125 which is compiled
// verbatim next to this file by oracle/Makefile).  Buffers are column-major doubles,
// exactly what R hands to .Call (R/RcppExports.R:4-18).
#include <RcppArmadillo.h>
#include <cstring>

arma::mat cf_sexp_c(arma::mat x1, arma::mat x2, double lscale, double magn, bool diag_only);
arma::mat cf_matern32_c(arma::mat x1, arma::mat x2, double lscale, double magn, bool diag_only);
arma::mat cf_matern52_c(arma::mat x1, arma::mat x2, double lscale, double magn, bool diag_only);
arma::mat cf_nn_c(arma::mat x1, arma::mat x2, double sigma0, double sigma, double magn,
                  bool diag_only);

static arma::mat wrap_in(const double* x, long n, long d) {
  arma::mat m((arma::uword)n, (arma::uword)d);
  std::memcpy(m.mem.data(), x, sizeof(double) * (size_t)n * (size_t)d);
  return m;
}

extern "C" {

// kind: 0 sexp(p1=lscale,p2=magn)  1 matern32  2 matern52  3 nn(p1=sigma0,p2=sigma,p3=magn)
// out: n1*n2 doubles column-major, or min(n1,n2) doubles when diag_only.
int gplite_ref_cf(int kind, const double* x1, long n1, const double* x2, long n2, long d,
                  double p1, double p2, double p3, int diag_only, double* out) {
  arma::mat a = wrap_in(x1, n1, d), b = wrap_in(x2, n2, d), K;
  switch (kind) {
    case 0: K = cf_sexp_c(a, b, p1, p2, diag_only != 0); break;
    case 1: K = cf_matern32_c(a, b, p1, p2, diag_only != 0); break;
    case 2: K = cf_matern52_c(a, b, p1, p2, diag_only != 0); break;
    case 3: K = cf_nn_c(a, b, p1, p2, p3, diag_only != 0); break;
    default: return 1;
  }
  std::memcpy(out, K.mem.data(), sizeof(double) * K.mem.size());
  return 0;
}

}  // extern "C"
