// Window-sized problems: l1ra followed by irls in ONE launch of ONE thread block.
//
// ViewGraph::rotAvg's local calls (src/IRotAvg.cpp:377: rotAvg(10)) solve ~15 views / ~40 edges, ten
// thousand times in a row.  At that size the general pipeline is pure launch + synchronisation latency
// (~30 launches and several host round trips per l1ra iteration: 1.3 ms per call measured).  Here the whole
// call sequence of src/ViewGraph.cpp:1402-1417 - l1ra (ral/l1_irls.cpp:851-912) with its primal-dual L1
// decoder (:228-468), then irls (:559-752) - runs inside one block with every array in shared memory:
//   * 3 warps, warp c owns coordinate c: the three l1decode_pd calls of :890-892 are independent (own tau,
//     own step length, own back-tracking loop), so each warp runs the scalar algorithm for its coordinate
//     with warp-synchronous loops and shuffle reductions - no block barrier inside the decoder;
//   * the Newton systems H dx = w1p (UMFPACK LU in the reference, :131-184) and the IRLS normal equations
//     (SuiteSparseQR least squares, :536-556) are dense nf x nf SPD systems here: Cholesky in shared
//     memory, one matrix row per lane (nf <= 32) - an exact solve, like the reference's;
//   * matrices are assembled row-wise from a per-row entry list in fixed order, reductions are
//     butterfly shuffles: the call is bitwise repeatable;
//   * inputs are read from, and results written to, mapped pinned host memory: one launch and one stream
//     synchronisation per rotAvg call, no separate copies.
// Limits: n_total <= 64, n_free <= 32, m <= 256 (larger windows take the general pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ira_l1ra.cuh"

namespace ira {

constexpr int kSmN = 64, kSmNF = 32, kSmM = 256, kSmThreads = 96;
constexpr int kSmScores = 8;

struct SmallIn {            // header of the mapped input block; arrays follow (see small_in_bytes)
  int m, n, f, cost;
  int l1_max_iters, irls_max_iters, pad0, pad1;
  double l1_th, irls_th, sigma, pad2;
};
struct SmallOut {           // header of the mapped output block; Q (4n, column-major) and weights (m) follow
  int l1_iters, irls_iters, stuck, nonfinite;
  double l1_score_last, irls_score_last;
  double irls_score[kSmScores];
};
// input block:  SmallIn | int32 I[2 m] (padded to 8 B) | double QQ[4 m] column-major | double Q[4 n] column-major
__host__ __device__ inline size_t small_in_I(void) { return sizeof(SmallIn); }
__host__ __device__ inline size_t small_in_QQ(int m) { return sizeof(SmallIn) + (((size_t)2 * m * 4 + 7) & ~(size_t)7); }
__host__ __device__ inline size_t small_in_Q(int m) { return small_in_QQ(m) + (size_t)4 * m * 8; }
__host__ __device__ inline size_t small_in_bytes(int m, int n) { return small_in_Q(m) + (size_t)4 * n * 8; }
__host__ __device__ inline size_t small_out_bytes(int m, int n) { return sizeof(SmallOut) + (size_t)(4 * n + m) * 8; }

struct SmallWarp {          // one coordinate's decoder state
  double u[kSmM], ax[kSmM], l1[kSmM], l2[kSmM];
  double adx[kSmM], du[kSmM], dl1[kSmM], dl2[kSmM], sigx[kSmM], ev[kSmM];
  double H[kSmNF * kSmNF];  // column-major, lane = row
  double x[kSmNF], atv[kSmNF], atdv[kSmNF], rhs[kSmNF], dx[kSmNF];
};
struct SmallSmem {
  double4 QQ[kSmM];
  double4 Q[kSmN];
  double W[3][kSmM];        // residual components (rotation vectors)
  double X[3][kSmNF];       // increments per coordinate
  double wt[kSmM];          // square-root IRLS weights
  int2 I[kSmM];
  int hp[kSmM], hm[kSmM];   // make_AtA pattern: free index of j / of i, or -1 (:825-835)
  int cm[kSmM];             // make_A: -1 column only when BOTH endpoints are free (:770-776); +1 column = hp
  int rowptr[kSmNF + 1];
  int ent[2 * kSmM];        // per free row, in edge order: (k << 1) | side   (side 0: row is j, 1: row is i)
  int stuck;
  SmallWarp pw[3];
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// out[r] = (A^T v)[r] on make_A's pattern; returns |out|^2 (warp-uniform).  One row per lane.
__device__ __forceinline__ double small_At(const SmallSmem& s, const double* v, double* out, int nf, int lane) {
  double acc = 0.0;
  if (lane < nf) {
    for (int e = s.rowptr[lane]; e < s.rowptr[lane + 1]; ++e) {
      const int k = s.ent[e] >> 1;
      if (s.ent[e] & 1) { if (s.cm[k] >= 0) acc -= v[k]; }
      else acc += v[k];
    }
    out[lane] = acc;
  }
  __syncwarp();
  return warp_sum(lane < nf ? acc * acc : 0.0);
}

// H = P^T diag(d) P, column-major nf x nf (lane = row).  quirk_in_diag: make_AtA's pattern keeps the
// (free i, fixed j) edges on the diagonal (Newton matrix); make_A's drops them (IRLS normal equations).
__device__ __forceinline__ void small_assemble(const SmallSmem& s, const double* d, double* H, int nf, int lane,
                                               bool quirk_in_diag) {
  if (lane < nf) {
    for (int c = 0; c < nf; ++c) H[c * kSmNF + lane] = 0.0;
    double diag = 0.0;
    for (int e = s.rowptr[lane]; e < s.rowptr[lane + 1]; ++e) {
      const int k = s.ent[e] >> 1;
      const bool side_i = s.ent[e] & 1;
      const bool quirk = s.hm[k] >= 0 && s.cm[k] < 0;              // i free, j fixed
      if (quirk && !quirk_in_diag) continue;
      diag += d[k];
      const int other = side_i ? s.hp[k] : s.hm[k];
      if (other == lane) diag -= d[k];                             // self loop: (e_r - e_r) contributes nothing
      else if (other >= 0) H[other * kSmNF + lane] -= d[k];
    }
    H[lane * kSmNF + lane] = diag;
  }
  __syncwarp();
}

// In-place Cholesky H = L L^T (lower triangle, column-major, lane = row) and solve H x = b.
// A non-positive pivot (a free node without any weighted edge) pins that unknown to 0.
__device__ __forceinline__ void small_chol_solve(double* H, const double* b, double* x, int nf, int lane) {
  unsigned null_rows = 0u;
  for (int j = 0; j < nf; ++j) {
    const double piv = H[j * kSmNF + j];
    const bool ok = piv > 0.0;
    if (!ok) null_rows |= 1u << j;
    const double d = ok ? sqrt(piv) : 1.0;
    __syncwarp();
    double lij = 0.0;
    if (lane < nf && lane >= j) {
      lij = lane == j ? d : (ok ? H[j * kSmNF + lane] / d : 0.0);
      H[j * kSmNF + lane] = lij;
    }
    __syncwarp();
    for (int k = j + 1; k < nf; ++k) {
      const double lkj = H[j * kSmNF + k];
      if (lane < nf && lane >= k) H[k * kSmNF + lane] -= lij * lkj;
    }
    __syncwarp();
  }
  // forward: L y = b
  double y = lane < nf ? b[lane] : 0.0;
  for (int j = 0; j < nf; ++j) {
    const double yj = __shfl_sync(0xffffffffu, y, j) / H[j * kSmNF + j];
    if (lane == j) y = yj;
    else if (lane < nf && lane > j) y -= H[j * kSmNF + lane] * yj;
  }
  // backward: L^T x = y
  for (int j = nf - 1; j >= 0; --j) {
    const double xj = __shfl_sync(0xffffffffu, y, j) / H[j * kSmNF + j];
    if (lane == j) y = xj;
    else if (lane < j) y -= H[lane * kSmNF + j] * xj;
  }
  if (lane < nf) x[lane] = (null_rows >> lane) & 1u ? 0.0 : y;
  __syncwarp();
}

// l1decode_pd(x0 = 0, A, y, pdmaxiter, AtA) for one coordinate, one warp (ral/l1_irls.cpp:228-468).
// Result in pw.x.  Returns 1 when the back-tracking got stuck (:423-428).
__device__ int small_l1decode(SmallSmem& s, SmallWarp& pw, const double* y, int m, int nf, int pdmaxiter, int lane) {
  double ymax = -INFINITY;
  for (int k = lane; k < m; k += 32) ymax = fmax(ymax, fabs(y[k]));
  ymax = warp_max(ymax);
  double sdg = 0.0, rd2e = 0.0;
  for (int k = lane; k < m; k += 32) {                                   // :248-262
    const double u = 0.95 * fabs(y[k]) + 0.10 * ymax;
    const double f1 = -y[k] - u, f2 = y[k] - u;
    const double a1 = -1.0 / f1, a2 = -1.0 / f2;
    pw.u[k] = u; pw.ax[k] = 0.0; pw.l1[k] = a1; pw.l2[k] = a2; pw.ev[k] = a1 - a2;
    sdg += -(f1 * a1 + f2 * a2);
    const double rm = 1.0 - a1 - a2;
    rd2e += rm * rm;
  }
  if (lane < nf) pw.x[lane] = 0.0;
  __syncwarp();
  sdg = warp_sum(sdg);
  rd2e = warp_sum(rd2e);
  double tau = kPdMu * (double)(2 * m) / sdg;                            // :265
  double rd2n = small_At(s, pw.ev, pw.atv, nf, lane);
  double rc2 = 0.0;
  for (int k = lane; k < m; k += 32) {                                   // :267-281
    const double f1 = pw.ax[k] - y[k] - pw.u[k], f2 = -pw.ax[k] + y[k] - pw.u[k];
    const double r1 = -pw.l1[k] * f1 - 1.0 / tau, r2 = -pw.l2[k] * f2 - 1.0 / tau;
    rc2 += r1 * r1 + r2 * r2;
  }
  rc2 = warp_sum(rc2);
  double resnorm = sqrt(rd2n + rd2e + rc2);
  int pditer = 0;
  bool done = sdg < kPdTol || pditer >= pdmaxiter;                       // :284
  while (!done) {
    ++pditer;
    const double it = 1.0 / tau;
    for (int k = lane; k < m; k += 32) {                                 // :293-306
      const double f1 = pw.ax[k] - y[k] - pw.u[k], f2 = -pw.ax[k] + y[k] - pw.u[k];
      const double w2 = -1.0 - it * (1.0 / f1 + 1.0 / f2);
      const double sig1 = -pw.l1[k] / f1 - pw.l2[k] / f2, sig2 = pw.l1[k] / f1 - pw.l2[k] / f2;
      pw.sigx[k] = sig1 - sig2 * sig2 / sig1;
      pw.ev[k] = -it * (-1.0 / f1 + 1.0 / f2) - (sig2 / sig1) * w2;
    }
    __syncwarp();
    small_At(s, pw.ev, pw.rhs, nf, lane);                                // w1p
    small_assemble(s, pw.sigx, pw.H, nf, lane, true);                    // :308-317
    small_chol_solve(pw.H, pw.rhs, pw.dx, nf, lane);                     // :319
    double smin = INFINITY;
    for (int k = lane; k < m; k += 32) {                                 // :324-381
      double adx = 0.0;
      if (s.hp[k] >= 0) { adx = pw.dx[s.hp[k]]; if (s.cm[k] >= 0) adx -= pw.dx[s.cm[k]]; }
      const double f1 = pw.ax[k] - y[k] - pw.u[k], f2 = -pw.ax[k] + y[k] - pw.u[k];
      const double a1 = pw.l1[k], a2 = pw.l2[k];
      const double w2 = -1.0 - it * (1.0 / f1 + 1.0 / f2);
      const double sig1 = -a1 / f1 - a2 / f2, sig2 = a1 / f1 - a2 / f2;
      const double du = (w2 - sig2 * adx) / sig1;
      const double d1 = -(a1 / f1) * (adx - du) - a1 - it / f1;
      const double d2 = (a2 / f2) * (adx + du) - a2 - it / f2;
      if (d1 < 0.0) smin = fmin(smin, -a1 / d1);
      if (d2 < 0.0) smin = fmin(smin, -a2 / d2);
      const double p1 = adx - du, p2 = -adx - du;
      if (p1 > 0.0) smin = fmin(smin, -f1 / p1);
      if (p2 > 0.0) smin = fmin(smin, -f2 / p2);
      pw.adx[k] = adx; pw.du[k] = du; pw.dl1[k] = d1; pw.dl2[k] = d2; pw.ev[k] = d1 - d2;
    }
    __syncwarp();
    smin = warp_min(smin);
    double st = 0.99 * fmin(1.0, smin);
    small_At(s, pw.ev, pw.atdv, nf, lane);                               // :342
    int backiter = 0;
    double rd2n_t = 0.0, rd2e_t = 0.0;
    for (;;) {                                                           // :392-429
      double a = 0.0, b = 0.0;
      for (int k = lane; k < m; k += 32) {
        const double up = pw.u[k] + st * pw.du[k], axp = pw.ax[k] + st * pw.adx[k];
        const double l1p = pw.l1[k] + st * pw.dl1[k], l2p = pw.l2[k] + st * pw.dl2[k];
        const double f1 = axp - y[k] - up, f2 = -axp + y[k] - up;
        const double rm = 1.0 - l1p - l2p;
        const double r1 = -l1p * f1 - it, r2 = -l2p * f2 - it;
        a += rm * rm;
        b += r1 * r1 + r2 * r2;
      }
      double c = 0.0;
      if (lane < nf) { const double t = pw.atv[lane] + st * pw.atdv[lane]; c = t * t; }
      rd2e_t = warp_sum(a);
      const double rc2_t = warp_sum(b);
      rd2n_t = warp_sum(c);
      const bool suff = sqrt(rd2n_t + rd2e_t + rc2_t) <= (1.0 - kPdAlpha * st) * resnorm;
      ++backiter;
      if (backiter > 32) return 1;                                       // stuck: x stays the last iterate
      if (suff) break;
      st *= kPdBeta;
    }
    double sd = 0.0;
    for (int k = lane; k < m; k += 32) {                                 // :432-446
      pw.u[k] += st * pw.du[k]; pw.ax[k] += st * pw.adx[k];
      pw.l1[k] += st * pw.dl1[k]; pw.l2[k] += st * pw.dl2[k];
      const double f1 = pw.ax[k] - y[k] - pw.u[k], f2 = -pw.ax[k] + y[k] - pw.u[k];
      sd += -(f1 * pw.l1[k] + f2 * pw.l2[k]);
    }
    if (lane < nf) { pw.x[lane] += st * pw.dx[lane]; pw.atv[lane] += st * pw.atdv[lane]; }
    __syncwarp();
    sdg = warp_sum(sd);
    tau = kPdMu * (double)(2 * m) / sdg;                                 // :448
    rc2 = 0.0;
    for (int k = lane; k < m; k += 32) {                                 // :450-453
      const double f1 = pw.ax[k] - y[k] - pw.u[k], f2 = -pw.ax[k] + y[k] - pw.u[k];
      const double r1 = -pw.l1[k] * f1 - 1.0 / tau, r2 = -pw.l2[k] * f2 - 1.0 / tau;
      rc2 += r1 * r1 + r2 * r2;
    }
    rc2 = warp_sum(rc2);
    resnorm = sqrt(rd2n_t + rd2e_t + rc2);                               // rdual = rdp (:455-458)
    done = sdg < kPdTol || pditer >= pdmaxiter;                          // :460
  }
  return 0;
}

// score = mean |X_i| (:729, :894); Q_{i+f} <- Q_{i+f} (x) Exp(X_i) (:731-737, :896-902).  All threads.
__device__ __forceinline__ double small_update(SmallSmem& s, int n, int f, double* red /* [4] */) {
  const int nf = n - f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double part = 0.0;
  for (int i = tid; i < nf; i += kSmThreads) {
    double th;
    const double4 dq = exp_quat(s.X[0][i], s.X[1][i], s.X[2][i], &th);
    part += th;
    s.Q[i + f] = quat_mult(s.Q[i + f], dq);
  }
  part = warp_sum(part);
  if (lane == 0) red[warp] = part;
  __syncthreads();
  const double score = (red[0] + red[1] + red[2]) / (double)nf;
  __syncthreads();
  return score;
}

__device__ __forceinline__ void small_residual(SmallSmem& s, int m) {    // :592-593 / :885-887
  for (int k = threadIdx.x; k < m; k += kSmThreads) {
    const double4 w = edge_residual(s.Q[s.I[k].x], s.QQ[k], s.Q[s.I[k].y]);
    s.W[0][k] = w.x; s.W[1][k] = w.y; s.W[2][k] = w.z;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kSmThreads, 1)
k_small_l1ra_irls(const unsigned char* __restrict__ in, unsigned char* __restrict__ out) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  SmallSmem& s = *reinterpret_cast<SmallSmem*>(smem_raw);
  __shared__ double red[4];
  const SmallIn hd = *reinterpret_cast<const SmallIn*>(in);
  const int m = hd.m, n = hd.n, f = hd.f, nf = n - f;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int* Iin = reinterpret_cast<const int*>(in + small_in_I());
  const double* QQin = reinterpret_cast<const double*>(in + small_in_QQ(m));
  const double* Qin = reinterpret_cast<const double*>(in + small_in_Q(m));
  for (int k = tid; k < m; k += kSmThreads) {
    const int i = Iin[2 * k], j = Iin[2 * k + 1];
    s.I[k] = make_int2(i, j);
    s.QQ[k] = make_double4(QQin[k], QQin[m + k], QQin[2 * m + k], QQin[3 * m + k]);
    s.hp[k] = j >= f ? j - f : -1;
    s.hm[k] = i >= f ? i - f : -1;
    s.cm[k] = (j >= f && i >= f) ? i - f : -1;
    s.wt[k] = 1.0;
  }
  for (int i = tid; i < n; i += kSmThreads) s.Q[i] = make_double4(Qin[i], Qin[n + i], Qin[2 * n + i], Qin[3 * n + i]);
  if (tid == 0) s.stuck = 0;
  __syncthreads();
  if (warp == 0) {                                     // per-row entry lists, edge order
    int cnt = 0;
    if (lane < nf) for (int k = 0; k < m; ++k) cnt += (s.hp[k] == lane) + (s.hm[k] == lane);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane < nf) s.rowptr[lane + 1] = incl;
    if (lane == 0) s.rowptr[0] = 0;
    int pos = incl - cnt;
    if (lane < nf)
      for (int k = 0; k < m; ++k) {
        if (s.hp[k] == lane) s.ent[pos++] = k << 1;
        if (s.hm[k] == lane) s.ent[pos++] = (k << 1) | 1;
      }
  }
  __syncthreads();
  SmallWarp& pw = s.pw[warp];
  SmallOut* oh = reinterpret_cast<SmallOut*>(out);

  // ---- l1ra (:851-912) -----------------------------------------------------------------------------
  double score = 1.7976931348623157e308;
  int iter = 0;
  while (score >= hd.l1_th && iter < hd.l1_max_iters && nf > 0) {        // :877 (l1_step is 2 for ever)
    small_residual(s, m);
    if (m > 0) {
      const int st = small_l1decode(s, pw, s.W[warp], m, nf, 2, lane);   // :889-892
      if (st && lane == 0) atomicOr(&s.stuck, 1 << warp);
    } else if (lane < nf) pw.x[lane] = 0.0;
    if (lane < nf) s.X[warp][lane] = pw.x[lane];
    __syncthreads();
    score = small_update(s, n, f, red);
    ++iter;
    if (!isfinite(score)) break;
  }
  const int l1_iters = iter;
  const double l1_score = score;

  // ---- irls (:559-752) -------------------------------------------------------------------------------
  score = 1.7976931348623157e308;
  iter = 0;
  const bool l1_ok = isfinite(l1_score) || l1_iters == 0;
  while (l1_ok && score > hd.irls_th && iter < hd.irls_max_iters && nf > 0) {   // :590
    small_residual(s, m);
    for (int k = lane; k < m; k += 32) {                                 // D^2 and D^2 w (:596-610)
      const double w2 = s.wt[k] * s.wt[k];
      pw.sigx[k] = w2;
      pw.ev[k] = w2 * s.W[warp][k];
    }
    __syncwarp();
    small_At(s, pw.ev, pw.rhs, nf, lane);
    small_assemble(s, pw.sigx, pw.H, nf, lane, false);
    small_chol_solve(pw.H, pw.rhs, pw.x, nf, lane);                      // :612
    if (lane < nf) s.X[warp][lane] = pw.x[lane];
    __syncthreads();
    for (int k = tid; k < m; k += kSmThreads) {                          // E = A X - w, new weights (:614-727)
      double e[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        double ax = 0.0;
        if (s.hp[k] >= 0) { ax = s.X[c][s.hp[k]]; if (s.cm[k] >= 0) ax -= s.X[c][s.cm[k]]; }
        e[c] = ax - s.W[c][k];
      }
      s.wt[k] = robust_weight(hd.cost, hd.sigma, e[0] * e[0] + e[1] * e[1] + e[2] * e[2], s.wt[k]);
    }
    score = small_update(s, n, f, red);                                  // begins with its own barrier use
    if (iter < kSmScores && tid == 0) oh->irls_score[iter] = score;
    ++iter;
    if (!isfinite(score)) break;
  }

  // ---- results ---------------------------------------------------------------------------------------
  double* Qout = reinterpret_cast<double*>(out + sizeof(SmallOut));
  double* wout = Qout + 4 * n;
  for (int i = tid; i < n; i += kSmThreads) {
    const double4 q = s.Q[i];
    Qout[i] = q.x; Qout[n + i] = q.y; Qout[2 * n + i] = q.z; Qout[3 * n + i] = q.w;
  }
  for (int k = tid; k < m; k += kSmThreads) wout[k] = s.wt[k];
  if (tid == 0) {
    oh->l1_iters = l1_iters; oh->irls_iters = iter; oh->stuck = s.stuck;
    oh->nonfinite = (!isfinite(l1_score) && l1_iters > 0) || (!isfinite(score) && iter > 0);
    oh->l1_score_last = l1_score; oh->irls_score_last = score;
  }
}

}  // namespace ira
