// l1_irls.hpp - C++ host mirror of the reference's RAL interface over the C ABI (include/ira.h).
//
// Drop-in for ral/l1_irls.hpp:89-112: the same free functions in namespace irotavg with the same
// argument order, in/out semantics and error behaviour (std::cerr + std::exit(-1) on failure, like
// ral/l1_irls.cpp:149-176,723-726), so that ral/test.cpp:286-302 and src/ViewGraph.cpp:1400-1417
// compile against it unchanged.  Everything numerical happens in libira.so (sm_100a CUDA).
//
// Matrix types.  With Eigen available (`__has_include(<Eigen/Dense>)`, the reference's own
// dependency) the typedefs are the reference's: Mat = Eigen::MatrixXd, Vec = Eigen::VectorXd.
// Without Eigen (this image) a minimal column-major Mat/Vec with the members the callers use
// (rows(), cols(), data(), operator()(i,j), setOnes(), resize) stands in, so the adapter and the
// C++ tests still build.  The functions are templates over any type with data()/rows()/cols() and
// column-major storage.
//
// `SpMat A` of the reference signatures is accepted and ignored: it is a pure function of (n, f, I)
// (ral/l1_irls.cpp:755-780) and the device rebuilds its pattern.  make_A is provided for callers that
// still want the matrix (returns the (+1, -1) column indices per row).
#ifndef IROTAVG_B200_L1_IRLS_HPP_
#define IROTAVG_B200_L1_IRLS_HPP_

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <utility>
#include <vector>

#include "ira.h"

#if !defined(IROTAVG_HAVE_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>) && __has_include(<Eigen/Sparse>)
#define IROTAVG_HAVE_EIGEN 1
#endif
#endif
#ifdef IROTAVG_HAVE_EIGEN
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/Sparse>
#endif

// Define IROTAVG_B200_NO_REFERENCE_NAMES to get only namespace ira_b200 (templates over the caller's
// own matrix types), e.g. to use ira_b200::irls next to the reference's ral/l1_irls.hpp while its l1ra /
// init_mst are still in use (option A).
#ifndef IROTAVG_B200_NO_REFERENCE_NAMES
namespace irotavg {

#define EPS 2.2204e-16                                   /* ral/l1_irls.hpp:40 */
const double DBL_MAX_ = std::numeric_limits<double>::max();

typedef std::vector<std::pair<int, int> > I_t;          /* ral/l1_irls.hpp:51 */

enum Cost { L2, L1, L15, L05, Geman_McClure, Huber, Pseudo_Huber, Andrews, Bisquare, Cauchy, Fair,
            Logistic, Talwar, Welsch };                  /* ral/l1_irls.hpp:56-57 */

inline std::ostream& operator<<(std::ostream& os, const Cost cost) {   /* ral/l1_irls.hpp:59-80 */
  static const char* names[] = {"L2", "L1", "L1.5", "L0.5", "Geman-McClure", "Huber", "Pseudo-Huber", "Andrews",
                                "Bisquare", "Cauchy", "Fair", "Logistic", "Talwar", "Welsch"};
  if (cost >= L2 && cost <= Welsch) os << names[cost];
  return os;
}

#ifdef IROTAVG_HAVE_EIGEN
// the reference's own typedefs, ral/l1_irls.hpp:43-50 (Long is SuiteSparse_long = long there)
typedef long Long;
typedef Eigen::SparseMatrix<double, Eigen::ColMajor, Long> SpMat;
typedef Eigen::MatrixXd Mat;
typedef Eigen::VectorXd Vec;
typedef Eigen::Vector3d Vec3;
typedef Eigen::Vector4d Vec4;
typedef Eigen::Quaterniond Quat;
typedef Eigen::Triplet<double> T;
#else
typedef long Long;
// Minimal column-major stand-ins (only what the reference's callers touch).
struct Mat {
  std::vector<double> v; long r = 0, c = 0;
  Mat() {}
  Mat(long rows_, long cols_) : v((size_t)rows_ * cols_, 0.0), r(rows_), c(cols_) {}
  static Mat Zero(long rows_, long cols_) { return Mat(rows_, cols_); }
  long rows() const { return r; }
  long cols() const { return c; }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double& operator()(long i, long j) { return v[(size_t)j * r + i]; }
  double operator()(long i, long j) const { return v[(size_t)j * r + i]; }
};
struct Vec {
  std::vector<double> v;
  Vec() {}
  explicit Vec(long n) : v((size_t)n, 0.0) {}
  long size() const { return (long)v.size(); }
  long rows() const { return (long)v.size(); }
  double* data() { return v.data(); }
  const double* data() const { return v.data(); }
  double& operator()(long i) { return v[(size_t)i]; }
  double operator()(long i) const { return v[(size_t)i]; }
  void setOnes() { for (auto& x : v) x = 1.0; }
};
#endif

#ifndef IROTAVG_HAVE_EIGEN
// The sparse incidence matrix of the reference, as the two column indices of every row (-1 = none).
struct SpMat {
  std::vector<int32_t> col_plus, col_minus;    // A(k, col_plus[k]) = +1, A(k, col_minus[k]) = -1
  long nrows = 0, ncols = 0;
  long rows() const { return nrows; }
  long cols() const { return ncols; }
};
#endif

}  // namespace irotavg
#endif  // IROTAVG_B200_NO_REFERENCE_NAMES

namespace ira_b200 {
typedef std::vector<std::pair<int, int> > I_t;
namespace detail {
inline ira_handle& handle() {
  static ira_handle h = nullptr;
  if (!h) {
    ira_status s = ira_create(&h, nullptr);
    if (s != IRA_OK) {
      std::cerr << "irotavg-b200: " << ira_status_string(s) << std::endl;
      std::exit(-1);
    }
  }
  return h;
}
inline void check(ira_status s, const char* what) {
  if (s != IRA_OK && s != IRA_ERR_NONFINITE) {
    std::cerr << what << " failed: " << ira_status_string(s) << ": " << ira_last_error(handle()) << std::endl;
    std::exit(-1);                                        // the reference's error behaviour
  }
}
// The reference solves every linear step exactly (SuiteSparseQR / UMFPACK); this library iterates.  A solve
// that stopped at cg_max_iters without reaching cg_rtol is therefore reported, never silent.
inline ira_stats& stats() { static ira_stats st; return st; }
inline void warn_unconverged(const char* what, const ira_stats& st) {
  if (st.cg_hit_max > 0) {
    double worst = 0.0;
    for (int k = 0; k < st.irls_iters && k < IRA_STATS_MAX_ITERS; ++k) worst = st.cg_relres[k] > worst ? st.cg_relres[k] : worst;
    std::cerr << "irotavg-b200 WARNING: " << what << ": " << st.cg_hit_max << " linear solve(s) stopped at the PCG iteration cap "
              << "without reaching cg_rtol (worst relative residual " << worst << "); the step is inexact where the reference's "
              << "is exact.  Raise ira_options.cg_max_iters." << std::endl;
  }
}
inline std::vector<int32_t> flatten(const I_t& I) {
  std::vector<int32_t> out(2 * I.size());
  for (size_t k = 0; k < I.size(); ++k) { out[2 * k] = I[k].first; out[2 * k + 1] = I[k].second; }
  return out;
}
}  // namespace detail

// SpMat make_A(n, f, I)   - ral/l1_irls.hpp:92, ral/l1_irls.cpp:755-780
template <class SpMatT>
inline SpMatT make_A_as(const int n, const int f, const I_t& I) {
  SpMatT A;
  A.nrows = (long)I.size(); A.ncols = n - f;
  A.col_plus.resize(I.size()); A.col_minus.resize(I.size());
  const std::vector<int32_t> flat = detail::flatten(I);
  ira_status s = ira_make_A((int64_t)I.size(), n, f, flat.data(), A.col_plus.data(), A.col_minus.data());
  if (s != IRA_OK) { std::cerr << "make_A failed: " << ira_status_string(s) << std::endl; std::exit(-1); }
  return A;
}

// void irls(QQ, I, A, cost, sigma, Q, f, max_iters, change_th, weights, iteration, runtime)
//   - ral/l1_irls.hpp:103-106, ral/l1_irls.cpp:559-752.  `weights` must be sized to m by the caller
//   (ral/test.cpp:299), Q rows [f, n) are updated in place.
template <class MatT, class VecT, class SpMatT, class CostT>
inline void irls(const MatT& QQ, const I_t& I, const SpMatT& /*A*/, CostT cost, double sigma, MatT& Q, const int f,
                 const int max_iters, double change_th, VecT& weights, int& iteration, double& runtime) {
  const int64_t m = (int64_t)QQ.rows(), n = (int64_t)Q.rows();
  if ((int64_t)I.size() != m || (int64_t)weights.rows() != m) {
    std::cerr << "irls: I, QQ and weights disagree on the number of connections" << std::endl;
    std::exit(-1);
  }
  const std::vector<int32_t> flat = detail::flatten(I);
  int32_t it = 0;
  double rt = 0.0;
  ira_status s = ira_irls(detail::handle(), m, n, f, flat.data(), QQ.data(), m > 0 ? m : 1, Q.data(), n > 0 ? n : 1,
                          (int32_t)cost, sigma, max_iters, change_th, weights.data(), &it, &rt, &detail::stats());
  if (s == IRA_ERR_UNKNOWN_COST) { std::cerr << "Unknown cost!!" << std::endl; std::exit(-1); }   // :723-726
  detail::check(s, "irls");
  detail::warn_unconverged("irls", detail::stats());
  iteration = it;
  runtime = rt;
  if (it >= max_iters) std::cout << " Max Iteration" << std::endl;                                // :746-749
}

// void l1ra(QQ, I, A, Q, f, max_iters, change_th, iter, runtime)
//   - ral/l1_irls.hpp:98-101, ral/l1_irls.cpp:851-912.  Q rows [f, n) are updated in place.
template <class MatT, class SpMatT>
inline void l1ra(const MatT& QQ, const I_t& I, const SpMatT& /*A*/, MatT& Q, const int f, const int max_iters,
                 double change_th, int& iter, double& runtime) {
  const int64_t m = (int64_t)QQ.rows(), n = (int64_t)Q.rows();
  if ((int64_t)I.size() != m) {
    std::cerr << "l1ra: I and QQ disagree on the number of connections" << std::endl;
    std::exit(-1);
  }
  const std::vector<int32_t> flat = detail::flatten(I);
  int32_t it = 0;
  double rt = 0.0;
  ira_status s = ira_l1ra(detail::handle(), m, n, f, flat.data(), QQ.data(), m > 0 ? m : 1, Q.data(), n > 0 ? n : 1,
                          max_iters, change_th, &it, &rt, &detail::stats());
  detail::check(s, "l1ra");
  detail::warn_unconverged("l1ra", detail::stats());
  iter = it;
  runtime = rt;
}

// void init_mst(Q, QQ, I, f)   - ral/l1_irls.hpp:89-90, ral/l1_irls.cpp:915-979.  Rows [0, f) of Q are
// kept, the others are overwritten with the spanning-tree start (same tree as the reference's edge sweeps).
template <class MatT>
inline void init_mst(MatT& Q, const MatT& QQ, const I_t& I, const int f) {
  const int64_t m = (int64_t)QQ.rows(), n = (int64_t)Q.rows();
  const std::vector<int32_t> flat = detail::flatten(I);
  ira_status s = ira_init_mst(detail::handle(), m, n, f, flat.data(), QQ.data(), m > 0 ? m : 1, Q.data(),
                              n > 0 ? n : 1, nullptr);
  if (s == IRA_ERR_NOT_SPANNING) {                                                                 // :970-977
    std::cerr << ira_last_error(detail::handle()) << "\nConnected Nodes are given as output\n"
              << "Remove extra nodes and retry." << std::endl;
    std::exit(-1);
  }
  detail::check(s, "init_mst");
}

// void quat_normalised(Q, f)   - ral/l1_irls.hpp:112, ral/l1_irls.cpp:982-991
template <class MatT>
inline void quat_normalised(MatT& Q, const int f) {
  const int64_t n = (int64_t)Q.rows();
  ira_status s = ira_quat_normalised(Q.data(), n, n > 0 ? n : 1, f);
  if (s != IRA_OK) { std::cerr << "quat_normalised failed" << std::endl; std::exit(-1); }
}

}  // namespace ira_b200

#ifndef IROTAVG_B200_NO_REFERENCE_NAMES
namespace irotavg {
#ifdef IROTAVG_HAVE_EIGEN
// The reference's SpMat, filled exactly like ral/l1_irls.cpp:755-780 from the rule ira_make_A restates
// (row k: +1 at column j-f if j >= f; -1 at column i-f if additionally i >= f; a self-loop's second write wins).
inline SpMat make_A(const int n, const int f, const I_t& I) {
  const std::vector<int32_t> flat = ira_b200::detail::flatten(I);
  std::vector<int32_t> cp(I.size()), cm(I.size());
  ira_status s = ira_make_A((int64_t)I.size(), n, f, flat.data(), cp.data(), cm.data());
  if (s != IRA_OK) { std::cerr << "make_A failed: " << ira_status_string(s) << std::endl; std::exit(-1); }
  SpMat A((Long)I.size(), (Long)(n - f));
  for (size_t k = 0; k < I.size(); ++k) {
    if (cp[k] >= 0) A.coeffRef((Long)k, cp[k]) = 1;
    if (cm[k] >= 0) A.coeffRef((Long)k, cm[k]) = -1;
  }
  A.makeCompressed();
  return A;
}
#else
inline SpMat make_A(const int n, const int f, const I_t& I) { return ira_b200::make_A_as<SpMat>(n, f, I); }
#endif
using ira_b200::irls;
using ira_b200::l1ra;
using ira_b200::init_mst;
using ira_b200::quat_normalised;
}  // namespace irotavg
#endif
#endif  // IROTAVG_B200_L1_IRLS_HPP_
