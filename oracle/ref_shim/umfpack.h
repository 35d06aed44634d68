/* umfpack.h - stand-in for the four UMFPACK calls of the reference's linsolve (ral/l1_irls.cpp:131-184): sparse LU
 * with partial pivoting of a square CSC matrix, here as a dense LU with partial pivoting (same factorisation, dense
 * storage: only for a few thousand unknowns).  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_UMFPACK_H_
#define ORACLE_REF_SHIM_UMFPACK_H_
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cholmod.h"

#define UMFPACK_INFO 90
#define UMFPACK_CONTROL 20
#define UMFPACK_A 0
#define UMFPACK_OK 0
#define UMFPACK_WARNING_singular_matrix 1
#define UMFPACK_ERROR_out_of_memory (-1)

typedef struct { long n; double* lu; long* piv; } ref_shim_umf_numeric;

static inline long umfpack_dl_symbolic(long n_row, long n_col, const long* Ap, const long* Ai, const double* Ax,
                                       void** Symbolic, const double* Control, double* Info) {
  (void)n_col; (void)Ap; (void)Ai; (void)Ax; (void)Control; (void)Info;
  long* s = (long*)malloc(sizeof(long));
  if (!s) return UMFPACK_ERROR_out_of_memory;
  *s = n_row;                                  /* the only thing _numeric needs from the symbolic object here */
  *Symbolic = s;
  return UMFPACK_OK;
}
static inline void umfpack_dl_free_symbolic(void** Symbolic) { if (Symbolic && *Symbolic) { free(*Symbolic); *Symbolic = 0; } }

static inline long umfpack_dl_numeric(const long* Ap, const long* Ai, const double* Ax, void* Symbolic, void** Numeric,
                                      const double* Control, double* Info) {
  (void)Control; (void)Info;
  const long n = *(const long*)Symbolic;
  ref_shim_umf_numeric* N = (ref_shim_umf_numeric*)malloc(sizeof(ref_shim_umf_numeric));
  if (!N) return UMFPACK_ERROR_out_of_memory;
  N->n = n;
  N->lu = (double*)calloc((size_t)n * (size_t)n + 1, sizeof(double));
  N->piv = (long*)malloc(sizeof(long) * (size_t)(n + 1));
  if (!N->lu || !N->piv) return UMFPACK_ERROR_out_of_memory;
  for (long j = 0; j < n; ++j)
    for (long k = Ap[j]; k < Ap[j + 1]; ++k) N->lu[(size_t)Ai[k] * n + j] += Ax[k];        /* row-major dense copy */
  long status = UMFPACK_OK;
  for (long k = 0; k < n; ++k) {
    long p = k; double best = fabs(N->lu[(size_t)k * n + k]);
    for (long i = k + 1; i < n; ++i) { const double a = fabs(N->lu[(size_t)i * n + k]); if (a > best) { best = a; p = i; } }
    N->piv[k] = p;
    if (best == 0.0) { status = UMFPACK_WARNING_singular_matrix; continue; }
    if (p != k)
      for (long j = 0; j < n; ++j) { const double t = N->lu[(size_t)k * n + j]; N->lu[(size_t)k * n + j] = N->lu[(size_t)p * n + j]; N->lu[(size_t)p * n + j] = t; }
    const double piv = N->lu[(size_t)k * n + k];
    const double* rk = N->lu + (size_t)k * n;
#pragma omp parallel for schedule(static)
    for (long i = k + 1; i < n; ++i) {
      double* ri = N->lu + (size_t)i * n;
      if (ri[k] != 0.0) {
        const double l = ri[k] / piv;
        ri[k] = l;
        for (long j = k + 1; j < n; ++j) ri[j] -= l * rk[j];
      }
    }
  }
  *Numeric = N;
  return status;
}
static inline void umfpack_dl_free_numeric(void** Numeric) {
  if (Numeric && *Numeric) { ref_shim_umf_numeric* N = (ref_shim_umf_numeric*)*Numeric; free(N->lu); free(N->piv); free(N); *Numeric = 0; }
}
static inline long umfpack_dl_solve(long sys, const long* Ap, const long* Ai, const double* Ax, double* X, const double* B,
                                    void* Numeric, const double* Control, double* Info) {
  (void)sys; (void)Ap; (void)Ai; (void)Ax; (void)Control; (void)Info;
  const ref_shim_umf_numeric* N = (const ref_shim_umf_numeric*)Numeric;
  const long n = N->n;
  memcpy(X, B, sizeof(double) * (size_t)n);
  for (long k = 0; k < n; ++k) {                /* P b: whole rows were swapped during the factorisation */
    const long p = N->piv[k];
    if (p != k) { const double t = X[k]; X[k] = X[p]; X[p] = t; }
  }
  for (long k = 0; k < n; ++k) {                /* L y = P b (unit lower triangle) */
    double s = X[k];
    for (long j = 0; j < k; ++j) s -= N->lu[(size_t)k * n + j] * X[j];
    X[k] = s;
  }
  for (long k = n - 1; k >= 0; --k) {
    double s = X[k];
    for (long j = k + 1; j < n; ++j) s -= N->lu[(size_t)k * n + j] * X[j];
    X[k] = s / N->lu[(size_t)k * n + k];
  }
  return UMFPACK_OK;
}
static inline void umfpack_dl_report_info(const double* Control, const double* Info) { (void)Control; (void)Info; }
static inline void umfpack_dl_report_status(const double* Control, long status) { (void)Control; (void)status; }
#endif
