// ral_ref_capi.cpp - a C entry-point layer over the REFERENCE's own functions (ral/l1_irls.hpp, compiled from
// /root/reference/ral/l1_irls.cpp together with this file by oracle/build_ref.py), so that the tests can call
// irotavg::irls / l1ra / init_mst / make_A / quat_normalised and the residual helpers through ctypes.
// Row-major (m x 4 / n x 4) buffers in and out; everything else is the reference's code.  TEST INFRASTRUCTURE ONLY.
#include "l1_irls.hpp"

namespace irotavg {   // defined in ral/l1_irls.cpp with external linkage, not declared in the header
Mat delta_rel(const I_t& I, const Mat& QQ, const Mat& Q);
void log_map(Mat& w);
void exp_map(Mat& W);
Vec4 quat_mult(const Vec4& q1, const Vec4& q2);
}  // namespace irotavg

using namespace irotavg;

static Mat from_rows(const double* p, long r, long c) {
  Mat M(r, c);
  for (long i = 0; i < r; ++i) for (long j = 0; j < c; ++j) M(i, j) = p[i * c + j];
  return M;
}
static void to_rows(const Mat& M, double* p) {
  const long r = M.rows(), c = M.cols();
  for (long i = 0; i < r; ++i) for (long j = 0; j < c; ++j) p[i * c + j] = M(i, j);
}
static I_t pairs(const int* I, long m) {
  I_t v; v.reserve((size_t)m);
  for (long k = 0; k < m; ++k) v.push_back(std::make_pair(I[2 * k], I[2 * k + 1]));
  return v;
}

extern "C" {

int ref_irls(long m, long n, int f, const int* I, const double* QQ, double* Q, int cost, double sigma, int max_iters,
             double change_th, double* weights, int* iters_out) {
  I_t Iv = pairs(I, m);
  Mat QQm = from_rows(QQ, m, 4), Qm = from_rows(Q, n, 4);
  SpMat A = make_A((int)n, f, Iv);
  Vec w(m);
  int iters = 0; double runtime = 0;
  irls(QQm, Iv, A, (Cost)cost, sigma, Qm, f, max_iters, change_th, w, iters, runtime);
  to_rows(Qm, Q);
  for (long k = 0; k < m; ++k) weights[k] = w(k);
  *iters_out = iters;
  return 0;
}

int ref_l1ra(long m, long n, int f, const int* I, const double* QQ, double* Q, int max_iters, double change_th,
             int* iters_out) {
  I_t Iv = pairs(I, m);
  Mat QQm = from_rows(QQ, m, 4), Qm = from_rows(Q, n, 4);
  SpMat A = make_A((int)n, f, Iv);
  int iters = 0; double runtime = 0;
  l1ra(QQm, Iv, A, Qm, f, max_iters, change_th, iters, runtime);
  to_rows(Qm, Q);
  *iters_out = iters;
  return 0;
}

int ref_init_mst(long m, long n, int f, const int* I, const double* QQ, double* Q) {
  I_t Iv = pairs(I, m);
  Mat QQm = from_rows(QQ, m, 4), Qm = from_rows(Q, n, 4);
  init_mst(Qm, QQm, Iv, f);
  to_rows(Qm, Q);
  return 0;
}

int ref_quat_normalised(long n, int f, double* Q) {
  Mat Qm = from_rows(Q, n, 4);
  quat_normalised(Qm, f);
  to_rows(Qm, Q);
  return 0;
}

// make_A as (row, col, value) triplets in column-major storage order; returns nnz (call with null outputs to size)
long ref_make_A(long m, long n, int f, const int* I, long* rows, long* cols, double* vals) {
  I_t Iv = pairs(I, m);
  SpMat A = make_A((int)n, f, Iv);
  const long nnz = A.nonZeros();
  if (rows && cols && vals) {
    const Long* op = A.outerIndexPtr(); const Long* ip = A.innerIndexPtr(); const double* vp = A.valuePtr();
    for (long j = 0; j < A.cols(); ++j)
      for (Long k = op[j]; k < op[j + 1]; ++k) { rows[k] = ip[k]; cols[k] = j; vals[k] = vp[k]; }
  }
  return nnz;
}

// w = log_map(delta_rel(I, QQ, Q)) (ral/l1_irls.cpp:592-593), m x 4 row-major out
int ref_residual(long m, long n, const int* I, const double* QQ, const double* Q, double* w_out) {
  I_t Iv = pairs(I, m);
  Mat QQm = from_rows(QQ, m, 4), Qm = from_rows(Q, n, 4);
  Mat w = delta_rel(Iv, QQm, Qm);
  log_map(w);
  to_rows(w, w_out);
  return 0;
}

int ref_exp_map(long n, double* W) {
  Mat Wm = from_rows(W, n, 4);
  exp_map(Wm);
  to_rows(Wm, W);
  return 0;
}

int ref_quat_mult(const double* a, const double* b, double* out) {
  Vec4 r = quat_mult(Vec4(a[0], a[1], a[2], a[3]), Vec4(b[0], b[1], b[2], b[3]));
  for (int k = 0; k < 4; ++k) out[k] = r(k);
  return 0;
}

}  // extern "C"
