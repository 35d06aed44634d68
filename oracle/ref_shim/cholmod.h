/* cholmod.h - stand-in for the CHOLMOD *types* the reference uses to hand matrices to SuiteSparseQR
 * (ral/l1_irls.cpp:50-96, 541-555).  No CHOLMOD algorithm is called by the reference.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_REF_SHIM_CHOLMOD_H_
#define ORACLE_REF_SHIM_CHOLMOD_H_
#include <stddef.h>
#include <stdlib.h>

typedef long SuiteSparse_long;

#define CHOLMOD_INT 0
#define CHOLMOD_INTLONG 1
#define CHOLMOD_LONG 2
#define CHOLMOD_PATTERN 0
#define CHOLMOD_REAL 1
#define CHOLMOD_DOUBLE 0

typedef struct cholmod_sparse_struct {
  size_t nrow, ncol, nzmax;
  void *p, *i, *nz, *x, *z;
  int stype, itype, xtype, dtype, sorted, packed;
} cholmod_sparse;

typedef struct cholmod_dense_struct {
  size_t nrow, ncol, nzmax, d;
  void *x, *z;
  int xtype, dtype;
} cholmod_dense;

typedef struct cholmod_common_struct { int print; int status; } cholmod_common;

static inline int cholmod_l_start(cholmod_common* c) { c->print = 0; c->status = 0; return 1; }
static inline int cholmod_l_finish(cholmod_common* c) { (void)c; return 1; }
static inline int cholmod_l_free_dense(cholmod_dense** X, cholmod_common* c) {
  (void)c;
  if (X && *X) { free((*X)->x); free(*X); *X = NULL; }
  return 1;
}
#endif
