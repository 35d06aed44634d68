// SuiteSparseQR.hpp - stand-in for SuiteSparseQR<double>(A, B, cc) = "X = A\B" (ral/l1_irls.cpp:550): the least-squares
// minimiser of ||A X - B||_F.  SPQR is a multifrontal sparse Householder QR; this stand-in runs the SAME factorisation
// densely (Householder reflections column by column, columns whose remaining norm is <= tol are "dead" exactly as in
// SPQR's rank detection with its default tolerance 20 (m+n) eps max_j ||A_j||_2; their X rows are 0 = SPQR's basic
// solution).  Dense storage: only for graphs of a few thousand edges.  TEST INFRASTRUCTURE ONLY.
#ifndef ORACLE_REF_SHIM_SUITESPARSEQR_HPP_
#define ORACLE_REF_SHIM_SUITESPARSEQR_HPP_
#include <cmath>
#include <cstdlib>
#include <vector>

#include "cholmod.h"

template <typename Entry>
cholmod_dense* SuiteSparseQR(cholmod_sparse* A, cholmod_dense* B, cholmod_common* cc) {
  (void)cc;
  const long m = (long)A->nrow, n = (long)A->ncol, nrhs = (long)B->ncol;
  const SuiteSparse_long* Ap = (const SuiteSparse_long*)A->p;
  const SuiteSparse_long* Ai = (const SuiteSparse_long*)A->i;
  const double* Ax = (const double*)A->x;
  std::vector<double> M((size_t)m * (size_t)n, 0.0);               // column-major dense copy
  double maxcol = 0.0;
  for (long j = 0; j < n; ++j) {
    double s = 0.0;
    for (SuiteSparse_long k = Ap[j]; k < Ap[j + 1]; ++k) { M[(size_t)j * m + Ai[k]] += Ax[k]; }
    for (long i = 0; i < m; ++i) s += M[(size_t)j * m + i] * M[(size_t)j * m + i];
    if (std::sqrt(s) > maxcol) maxcol = std::sqrt(s);
  }
  const double tol = 20.0 * (double)(m + n) * 2.220446049250313e-16 * maxcol;
  std::vector<double> C((size_t)m * (size_t)nrhs);
  const double* Bx = (const double*)B->x;
  for (long c = 0; c < nrhs; ++c)
    for (long i = 0; i < m; ++i) C[(size_t)c * m + i] = Bx[(size_t)c * B->d + i];
  std::vector<long> pivcol;                                         // live columns, in order
  std::vector<double> v((size_t)m);
  long rank = 0;
  for (long j = 0; j < n && rank < m; ++j) {
    double* col = &M[(size_t)j * m];
    double s = 0.0;
    for (long i = rank; i < m; ++i) s += col[i] * col[i];
    const double nrm = std::sqrt(s);
    if (nrm <= tol) continue;                                       // dead column: x_j = 0
    const double alpha = col[rank] > 0.0 ? -nrm : nrm;
    // v = x - alpha e1, H = I - 2 v v^T / (v^T v)
    for (long i = rank; i < m; ++i) v[(size_t)i] = col[i];
    v[(size_t)rank] -= alpha;
    double vtv = 0.0;
    for (long i = rank; i < m; ++i) vtv += v[(size_t)i] * v[(size_t)i];
    if (vtv > 0.0) {
#pragma omp parallel for schedule(static)
      for (long jj = j + 1; jj < n; ++jj) {
        double* cj = &M[(size_t)jj * m];
        double d = 0.0;
        for (long i = rank; i < m; ++i) d += v[(size_t)i] * cj[i];
        if (d != 0.0) { d = 2.0 * d / vtv; for (long i = rank; i < m; ++i) cj[i] -= d * v[(size_t)i]; }
      }
      for (long c = 0; c < nrhs; ++c) {
        double* cj = &C[(size_t)c * m];
        double d = 0.0;
        for (long i = rank; i < m; ++i) d += v[(size_t)i] * cj[i];
        d = 2.0 * d / vtv;
        for (long i = rank; i < m; ++i) cj[i] -= d * v[(size_t)i];
      }
    }
    col[rank] = alpha;
    for (long i = rank + 1; i < m; ++i) col[i] = 0.0;
    pivcol.push_back(j);
    ++rank;
  }
  cholmod_dense* X = (cholmod_dense*)std::malloc(sizeof(cholmod_dense));
  X->nrow = (size_t)n; X->ncol = (size_t)nrhs; X->nzmax = (size_t)(n * nrhs); X->d = (size_t)n;
  X->x = std::calloc((size_t)(n * nrhs) + 1, sizeof(double)); X->z = 0;
  X->xtype = CHOLMOD_REAL; X->dtype = CHOLMOD_DOUBLE;
  double* Xx = (double*)X->x;
  // back substitution on the rank x rank upper-triangular R (row r of R lives in rows r of the live columns)
  for (long c = 0; c < nrhs; ++c)
    for (long r = rank - 1; r >= 0; --r) {
      double s = C[(size_t)c * m + r];
      for (long q = r + 1; q < rank; ++q) s -= M[(size_t)pivcol[(size_t)q] * m + r] * Xx[(size_t)c * n + pivcol[(size_t)q]];
      Xx[(size_t)c * n + pivcol[(size_t)r]] = s / M[(size_t)pivcol[(size_t)r] * m + r];
    }
  return X;
}
#endif
