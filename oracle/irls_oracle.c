/* Plain-C restatement of irotavg::irls() - the CPU oracle / CPU baseline ("port").
 *
 * TEST AND BENCH INFRASTRUCTURE ONLY: nothing under irotavg_b200/ or include/ links or calls this
 * file; tests/ use it as an independent checker of oracle/irls_oracle.py and bench.py times it as
 * the `cpu_baseline` / `--impl reference` arm.  Pinned through oracle/irls_oracle.py, which tests/test_ref_pin.py
 * holds to the reference's own ral/l1_irls.cpp compiled by oracle/build_ref.py (tests/test_c_oracle.py: C == numpy).
 *
 * Follows ral/l1_irls.cpp (paths relative to the reference tree):
 *   quat_mult :99-105   delta_rel :109-127   log_map :498-532   exp_map :471-492
 *   irls :559-752 (loop :590, weights :617-727, score :729, update :734-737)   make_A :755-780
 * The SuiteSparseQR least-squares solve (:550, third-party, unpinned system package) is restated
 * from its definition X = argmin ||D A X - D w||_F as Jacobi-preconditioned CG on the normal
 * equations A^T D^2 A X = A^T D^2 w, x0 = 0 - the only formulation that can run the 1M-edge random
 * graph at all (a sparse direct factor would need ~40 GB of fill).
 * OpenMP parallelises the edge, row and node loops ("all the host threads it can use"); the
 * reference itself is single-threaded.
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libirls_oracle.so oracle/irls_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORA_EPS 2.2204e-16
#define ORA_PI 3.141592653589793238462643383279502884

static void qmul(const double* a, const double* b, double* r) { /* :99-105, [x y z w] */
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  r[0] = aw * bx + ax * bw + ay * bz - az * by;
  r[1] = aw * by + ay * bw + az * bx - ax * bz;
  r[2] = aw * bz + az * bw + ax * by - ay * bx;
  r[3] = aw * bw - ax * bx - ay * by - az * bz;
}

static double robust_weight(int cost, double sigma, double e2, double old) { /* :617-727 */
  double tun, e, w;
  switch (cost) {
    case 0: return old;
    case 3: w = 1.0 / pow(e2, 3.0 / 8.0); return w > 1e4 ? 1e4 : w;
    case 1: w = 1.0 / sqrt(sqrt(e2)); return w > 1e4 ? 1e4 : w;
    case 2: w = 1.0 / sqrt(sqrt(sqrt(e2))); return w > 1e4 ? 1e4 : w;
    case 4: return 1.0 / (e2 + sigma * sigma);
    case 5: tun = 1.345 * sigma; e = sqrt(e2) / tun; return e >= 1.0 ? sqrt(1.0 / e) : old;
    case 6: return 1.0 / sqrt(sqrt(1.0 + e2 / (sigma * sigma)));
    case 7:
      tun = 1.339 * sigma; e = sqrt(e2) / tun; w = sqrt(sin(e) / e);
      if (e >= ORA_PI) w = 0.0; else if (e < 1e-4) w = 1.0;
      if (w < 1e-4) w = 1e-4;
      return w;
    case 8: tun = 4.685 * sigma; w = 1.0 - e2 / (tun * tun); return w < 1e-4 ? 1e-4 : w;
    case 9: tun = 2.385 * sigma; return 1.0 / sqrt(1.0 + e2 / (tun * tun));
    case 10: tun = 1.400 * sigma; return 1.0 / sqrt(1.0 + sqrt(e2) / tun);
    case 11: tun = 1.205 * sigma; e = sqrt(e2) / tun; return e < 1e-4 ? 1.0 : sqrt(tanh(e) / e);
    case 12: tun = 2.795 * sigma; return e2 < tun * tun ? 1.0001 : 0.0;
    case 13: tun = 2.985 * sigma; w = exp(-0.5 * e2 / (tun * tun)); return w < 1e-4 ? 1e-4 : w;
    default: return old;
  }
}

/* Row-major inputs: I (m x 2 int32), QQ (m x 4), Q (n x 4, updated in place).  Returns 0 on success,
 * -1 unknown cost, -2 allocation failure.  scores / cg_iters need max_iters entries (may be NULL). */
int ora_irls(int64_t m, int64_t n, int32_t f, const int32_t* I, const double* QQ, double* Q, int32_t cost,
             double sigma, int32_t max_iters, double change_th, double cg_rtol, int32_t cg_max_iters,
             int32_t threads, double* weights, int32_t* iters_out, double* scores, int32_t* cg_iters) {
  if (cost < 0 || cost > 13) return -1;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  /* CSR of A^T A over all nodes (make_A mask :755-780): row j gets (col i, +k) when j >= f; row i gets
   * (col j, -k) when additionally i >= f.  Counting sort keeps entries in edge order. */
  int64_t* rowptr = (int64_t*)calloc((size_t)n + 2, sizeof(int64_t));
  if (!rowptr) return -2;
  for (int64_t k = 0; k < m; ++k) {
    const int32_t i = I[2 * k], j = I[2 * k + 1];
    if (j >= f) { rowptr[j + 1]++; if (i >= f) rowptr[i + 1]++; }
  }
  for (int64_t r = 0; r < n; ++r) rowptr[r + 1] += rowptr[r];
  const int64_t nnz = rowptr[n];
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz + 1));
  int64_t* eid = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nnz + 1));
  double* ew2 = (double*)malloc(sizeof(double) * (size_t)(nnz + 1));
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n + 1));
  double* w = (double*)malloc(sizeof(double) * 3 * (size_t)(m + 1));
  double* vec = (double*)calloc((size_t)(n + 1) * 3 * 6 + (size_t)(n + 1), sizeof(double));
  if (!col || !eid || !ew2 || !cur || !w || !vec) return -2;
  double *X = vec, *R = X + 3 * n, *Z = R + 3 * n, *P = Z + 3 * n, *AP = P + 3 * n, *B = AP + 3 * n, *dinv = B + 3 * n;
  memcpy(cur, rowptr, sizeof(int64_t) * (size_t)n);
  for (int64_t k = 0; k < m; ++k) {
    const int32_t i = I[2 * k], j = I[2 * k + 1];
    if (j >= f) {
      col[cur[j]] = i; eid[cur[j]++] = k;
      if (i >= f) { col[cur[i]] = j; eid[cur[i]++] = ~k; }
    }
  }
  for (int64_t k = 0; k < m; ++k) weights[k] = 1.0;                                  /* :577 */
  double score = 1.7976931348623157e308;                                              /* :574 */
  int iters = 0;
  const int64_t nf = n - f;
  while (score > change_th && iters < max_iters) {                                    /* :590 */
    /* residuals (:592-593) */
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < m; ++k) {
      const double* qi = Q + 4 * (int64_t)I[2 * k];
      const double* qj = Q + 4 * (int64_t)I[2 * k + 1];
      double t[4], p[4];
      const double qjn[4] = {qj[0], qj[1], qj[2], -qj[3]};                             /* :114-115 */
      qmul(QQ + 4 * k, qi, t);
      qmul(qjn, t, p);
      const double s = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
      double theta = 2.0 * atan2(s, p[3]);
      if (theta < -ORA_PI) theta += 2.0 * ORA_PI; else if (theta >= ORA_PI) theta -= 2.0 * ORA_PI;
      if (s < ORA_EPS) { w[3 * k] = w[3 * k + 1] = w[3 * k + 2] = 0.0; }
      else { const double a = theta / s; w[3 * k] = p[0] * a; w[3 * k + 1] = p[1] * a; w[3 * k + 2] = p[2] * a; }
    }
    /* b = A^T D^2 w, diag, per-entry weights (:596-610) */
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
      double bx = 0, by = 0, bz = 0, d = 0;
      for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
        const int64_t id = eid[e];
        const int neg = id < 0;
        const int64_t k = neg ? ~id : id;
        const double w2 = weights[k] * weights[k];
        ew2[e] = w2; d += w2;
        const double sg = neg ? -w2 : w2;
        bx += sg * w[3 * k]; by += sg * w[3 * k + 1]; bz += sg * w[3 * k + 2];
      }
      B[3 * r] = bx; B[3 * r + 1] = by; B[3 * r + 2] = bz;
      dinv[r] = d > 0.0 ? 1.0 / d : 0.0;
    }
    /* Jacobi-PCG, 3 right-hand sides with independent alpha/beta (ls_solve :536-556 restated) */
    double rz[3] = {0, 0, 0}, bb[3] = {0, 0, 0};
    for (int64_t r = 0; r < n; ++r)
      for (int c = 0; c < 3; ++c) {
        const double b = B[3 * r + c];
        X[3 * r + c] = 0.0; R[3 * r + c] = b; Z[3 * r + c] = dinv[r] * b; P[3 * r + c] = Z[3 * r + c];
        rz[c] += b * Z[3 * r + c]; bb[c] += b * b;
      }
    int it = 0;
    double rr[3] = {bb[0], bb[1], bb[2]};
    while (it < cg_max_iters) {
      if (rr[0] <= cg_rtol * cg_rtol * bb[0] && rr[1] <= cg_rtol * cg_rtol * bb[1] && rr[2] <= cg_rtol * cg_rtol * bb[2]) break;
      double pap0 = 0, pap1 = 0, pap2 = 0;
#pragma omp parallel for schedule(static) reduction(+ : pap0, pap1, pap2)
      for (int64_t r = 0; r < n; ++r) {
        double ax = 0, ay = 0, az = 0;
        const double px = P[3 * r], py = P[3 * r + 1], pz = P[3 * r + 2];
        for (int64_t e = rowptr[r]; e < rowptr[r + 1]; ++e) {
          const double* pc = P + 3 * (int64_t)col[e];
          const double w2 = ew2[e];
          ax += w2 * (px - pc[0]); ay += w2 * (py - pc[1]); az += w2 * (pz - pc[2]);
        }
        AP[3 * r] = ax; AP[3 * r + 1] = ay; AP[3 * r + 2] = az;
        pap0 += px * ax; pap1 += py * ay; pap2 += pz * az;
      }
      const double pap[3] = {pap0, pap1, pap2};
      double al[3], rz0 = 0, rz1 = 0, rz2 = 0, rr0 = 0, rr1 = 0, rr2 = 0;
      for (int c = 0; c < 3; ++c) al[c] = pap[c] > 0.0 ? rz[c] / pap[c] : 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rz0, rz1, rz2, rr0, rr1, rr2)
      for (int64_t r = 0; r < n; ++r) {
        double t[3];
        for (int c = 0; c < 3; ++c) {
          X[3 * r + c] += al[c] * P[3 * r + c];
          R[3 * r + c] -= al[c] * AP[3 * r + c];
          t[c] = R[3 * r + c];
          Z[3 * r + c] = dinv[r] * t[c];
        }
        rz0 += t[0] * Z[3 * r]; rz1 += t[1] * Z[3 * r + 1]; rz2 += t[2] * Z[3 * r + 2];
        rr0 += t[0] * t[0]; rr1 += t[1] * t[1]; rr2 += t[2] * t[2];
      }
      const double rzn[3] = {rz0, rz1, rz2};
      double be[3];
      for (int c = 0; c < 3; ++c) { be[c] = rz[c] > 0.0 ? rzn[c] / rz[c] : 0.0; rz[c] = rzn[c]; }
      rr[0] = rr0; rr[1] = rr1; rr[2] = rr2;
#pragma omp parallel for schedule(static)
      for (int64_t r = 0; r < n; ++r)
        for (int c = 0; c < 3; ++c) P[3 * r + c] = Z[3 * r + c] + be[c] * P[3 * r + c];
      ++it;
    }
    /* E = A X - w, new weights (:614-727) */
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < m; ++k) {
      const int32_t i = I[2 * k], j = I[2 * k + 1];
      double ex = -w[3 * k], ey = -w[3 * k + 1], ez = -w[3 * k + 2];
      if (j >= f) {
        ex += X[3 * (int64_t)j]; ey += X[3 * (int64_t)j + 1]; ez += X[3 * (int64_t)j + 2];
        if (i >= f) { ex -= X[3 * (int64_t)i]; ey -= X[3 * (int64_t)i + 1]; ez -= X[3 * (int64_t)i + 2]; }
      }
      weights[k] = robust_weight(cost, sigma, ex * ex + ey * ey + ez * ez, weights[k]);
    }
    /* score, exp map, right-multiply (:729-737) */
    double ssum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : ssum)
    for (int64_t r = f; r < n; ++r) {
      const double vx = X[3 * r], vy = X[3 * r + 1], vz = X[3 * r + 2];
      const double th = sqrt(vx * vx + vy * vy + vz * vz);
      ssum += th;
      const double k2 = sin(0.5 * th) / th;
      double d[4] = {vx * k2, vy * k2, vz * k2, cos(0.5 * th)};
      for (int c = 0; c < 4; ++c) if (!isfinite(d[c])) d[c] = 0.0;                  /* :491 */
      double q[4];
      qmul(Q + 4 * r, d, q);
      memcpy(Q + 4 * r, q, sizeof q);
    }
    score = ssum / (double)nf;
    if (scores) scores[iters] = score;
    if (cg_iters) cg_iters[iters] = it;
    ++iters;
  }
  *iters_out = iters;
  free(rowptr); free(col); free(eid); free(ew2); free(cur); free(w); free(vec);
  return 0;
}

int ora_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
