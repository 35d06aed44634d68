// Two-level PCG for small and medium view graphs (n <= kCoarseMaxRows): Jacobi + an additive piecewise-constant
// coarse space over contiguous index blocks.  Replaces ls_solve (ral/l1_irls.cpp:536-556) and linsolve
// (:131-184, the Newton systems of l1decode_pd) where k_pcg_persistent_w3 / k_pcg_persistent_reg_mw needed thousands
// of iterations.
//
// Why.  A SLAM view graph is a chain: every view is tied to the previous few (src/IRotAvg.cpp:159: 4) plus rare loop
// closures.  Its grounded Laplacian has condition number ~ n^2, so (block-)Jacobi PCG needs O(n) iterations: 920 on
// a 3 000-view stream graph, 432 on the reference's bundled fixture, ~5 000 on the 9 500-view global rotAvg calls of
// config 5 (119 ms per call).  The slow modes are smooth ALONG THE CHAIN, i.e. in the
// node index (views are numbered in frame order), so a coarse space of indicator vectors of contiguous index blocks
// captures them whatever the local edge pattern ((k, k+1) or (k, k+4), (k, k+5) as in ral/data/ravg_input.txt):
//     M^-1 = D^-1 + P (P^T L P)^-1 P^T,        P[v][a] = 1 iff v is a free node of block a = v / B
// numpy study (tools/precond_study.py --coarse): 920 -> 115 / 70 iterations with 64 / 125 blocks on the stream graph,
// 432 -> 66 / 44 on the bundled fixture; 64 -> 45 / 34 on config 2, whose 4 645 loop closures already make it
// well conditioned; nothing on random graphs (config 3), which never take this path.
//
// How.  One cooperative kernel per solve, three right-hand sides with THREE weight sets (the Newton systems of the
// three coordinates differ; irls passes the same weights three times), vectors in L2-resident global memory:
//   set-up   A_c = P^T L P per coordinate, nc <= 64 blocks: one warp per coarse row walks its rows' SELL entries in
//            a fixed order (deterministic sums); every block then inverts the three nc x nc matrices in its own
//            shared memory (in-place Gauss-Jordan, 3 x 32 KB) - identical arithmetic in every block, no broadcast;
//   per PCG iteration (Chronopoulos-Gear, one fused reduction; one slice per warp, its rows' x r p s u in registers):
//     A  w = L u (gathers of u, software pipelined), partial dots -> grid reduction (barrier 1)
//     B  p, s, x, r updated in registers, r published            -> barrier 2
//     C  r_c = P^T r, one warp per block of the partition        -> barrier 3
//     D  y_c = A_c^-1 r_c from shared memory (every block, redundantly), u = D^-1 r + y_c[block(row)]   -> barrier 4
// Four barriers instead of two per iteration (2-3x the cost of a one-level iteration on these small grids) for 6-13x
// fewer iterations.
#pragma once
#include "ira_l1ra.cuh"

namespace ira {

constexpr int kCoarseMax = 64;          // coarse unknowns per coordinate (3 x 64 x 64 doubles = 96 KB of shared memory)
constexpr int kCoarseThreads = 384;     // >= kPcgNV warps: pcg_grid_reduce gives each of its 9 sums to one warp
constexpr int kCoarseMaxRows = 32768;   // larger graphs keep the one-level kernels (148 x 12 warps hold 56 832 rows)

struct PcgCoarseParams {
  PcgW3Params w;                        // matrix (3 weights per entry), vectors, partials, ctl
  const int* sell_pos;                  // row -> SELL position
  int f;                                // rows < f are fixed (never unknowns, never in P)
  int nc, bsz;                          // coarse size and rows per block of the partition: block(v) = v / bsz
  double* AC;                           // [3][nc][nc] assembled coarse matrices (global scratch)
  double4* RC;                          // [nc] coarse residual of the current iteration
};

// r_c = P^T r: warp a sums the rows of block a in a fixed order (lane-strided partial sums, butterfly).
__device__ __forceinline__ void coarse_restrict(const PcgCoarseParams& q, int gwarp, int nwarps, int lane) {
  const PcgW3Params& p = q.w;
  for (int a = gwarp; a < q.nc; a += nwarps) {
    const int v0 = max(a * q.bsz, q.f), v1 = min(p.n, (a + 1) * q.bsz);
    double sx = 0, sy = 0, sz = 0;
    for (int v = v0 + lane; v < v1; v += 32) {
      const double4 r = ld256(p.R + v), di = ld256(p.DINV + v);
      if (di.x != 0.0) sx += r.x;
      if (di.y != 0.0) sy += r.y;
      if (di.z != 0.0) sz += r.z;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) st256(q.RC + a, make_double4(sx, sy, sz, 0.0));
  }
}

// y_c = A_c^-1 r_c for the three coordinates, from the block's shared copy of the inverses (symmetric: column reads);
// r_c is staged through shared memory (yc doubles as the staging buffer: read fully before it is overwritten).
__device__ __forceinline__ void coarse_solve(const PcgCoarseParams& q, const double* ainv /* [3][nc][nc] */, double4* yc,
                                             double4* rc_s /* [nc] shared */) {
  const int nc = q.nc;
  for (int t = threadIdx.x; t < nc; t += blockDim.x) rc_s[t] = ld256(q.RC + t);
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * nc; t += blockDim.x) {
    const int c = t / nc, a = t % nc;
    const double* A = ainv + (size_t)c * nc * nc;
    double acc = 0.0;
    for (int j = 0; j < nc; ++j) acc += A[j * nc + a] * reinterpret_cast<const double*>(rc_s + j)[c];
    reinterpret_cast<double*>(yc + a)[c] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCoarseThreads, 1)
k_pcg_coarse_w3(const PcgCoarseParams q) {
  const PcgW3Params& p = q.w;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(32) unsigned char dyn_smem[];
  double* const ainv = reinterpret_cast<double*>(dyn_smem);                       // [3][nc][nc]
  double4* const yc = reinterpret_cast<double4*>(ainv + ((3 * q.nc * q.nc + 3) & ~3));   // [nc], 32-byte aligned
  double4* const rc_s = yc + q.nc;                                                // [nc]
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31;
  const int wpb = kCoarseThreads / 32;
  const int gwarp = blockIdx.x * wpb + (threadIdx.x >> 5);           // one slice per warp for the whole solve:
  const int nwarps = gridDim.x * wpb;                                 // the rows' x r p s u stay in registers
  const int nc = q.nc;
  int row = -1, width = 0;
  int64_t base = 0;
  if (gwarp < p.nslices) {
    row = p.sell_row[gwarp * kSellC + lane];
    width = p.slice_width[gwarp];
    base = (int64_t)p.slice_off[gwarp] + lane;
  }
  const bool in_p = row >= q.f;                                       // fixed rows are never coarse unknowns

  // ---- set-up 1: D^-1, x = 0, r = b; coarse matrices A_c = P^T L P, one warp per coarse row ---------------------
  double v[kPcgNV];
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, s0 = 0, s1 = 0, s2 = 0;
  double u0 = 0, u1 = 0, u2 = 0, d0 = 0, d1 = 0, d2 = 0;
  if (row >= 0) {
    const double4 b = ldg256(p.B + row), d = ldg256(p.diag3 + row);
    d0 = d.x > 0.0 ? 1.0 / d.x : 0.0; d1 = d.y > 0.0 ? 1.0 / d.y : 0.0; d2 = d.z > 0.0 ? 1.0 / d.z : 0.0;
    st256(p.DINV + row, make_double4(d0, d1, d2, 0.0));
    r0 = b.x; r1 = b.y; r2 = b.z;
    st256(p.R + row, b);
    v[0] = r0 * r0; v[1] = r1 * r1; v[2] = r2 * r2;
  }
  for (int a = gwarp; a < nc; a += nwarps) {
    // lane l owns the coarse columns l and l + 32; every entry of every row of block a is visited in a fixed order
    double acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    const int v0 = max(a * q.bsz, q.f), v1 = min(p.n, (a + 1) * q.bsz);
    for (int vv = v0; vv < v1; ++vv) {
      const int pos = q.sell_pos[vv];
      const int sl = pos / kSellC, ln = pos % kSellC;
      const int wd = p.slice_width[sl];
      const int64_t bs = (int64_t)p.slice_off[sl] + ln;
      for (int j0 = 0; j0 < wd; j0 += 32) {                         // 32 entries of the row at a time, one per lane
        const int j = j0 + lane;
        int col = vv;
        double4 w3 = make_double4(0, 0, 0, 0);
        if (j < wd) { col = __ldg(p.sell_col + bs + (int64_t)j * kSellC); w3 = ldg256(p.sell_w3 + bs + (int64_t)j * kSellC); }
        const int cnt = min(32, wd - j0);
        for (int e = 0; e < cnt; ++e) {                              // then added one after the other by the column's owner
          const int ce = __shfl_sync(0xffffffffu, col, e);
          const double wx = __shfl_sync(0xffffffffu, w3.x, e), wy = __shfl_sync(0xffffffffu, w3.y, e), wz = __shfl_sync(0xffffffffu, w3.z, e);
          if (ce == vv) continue;                                    // padding slot
          if ((a & 31) == lane) { acc[0][a >> 5] += wx; acc[1][a >> 5] += wy; acc[2][a >> 5] += wz; }   // L[v][v] part
          if (ce >= q.f) {                                           // -w couples to the column's block when it is a free node
            const int bc = ce / q.bsz;
            if ((bc & 31) == lane) { acc[0][bc >> 5] -= wx; acc[1][bc >> 5] -= wy; acc[2][bc >> 5] -= wz; }
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int bcol = lane + 32 * h;
        if (bcol < nc) q.AC[((size_t)c * nc + a) * nc + bcol] = acc[c][h];
      }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);                    // |b|^2; also publishes R, DINV, AC
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  // ---- set-up 2: every block inverts the three coarse matrices in its own shared memory (in-place Gauss-Jordan; SPD,
  //      no pivoting).  An empty block of the partition (all its nodes fixed) or a floating component leaves a zero
  //      pivot: that coarse unknown is switched off (row and column zeroed), the preconditioner stays SPD. --------------
  for (int t = threadIdx.x; t < 3 * nc * nc; t += blockDim.x) ainv[t] = __ldcg(q.AC + t);
  __shared__ double piv_inv[3];
  __shared__ double scale[3 * kCoarseMax];
  __shared__ int dead[3 * kCoarseMax];
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * nc; t += blockDim.x) {          // row scales of the assembled matrices
    double m = 0.0;
    for (int jj = 0; jj < nc; ++jj) m = fmax(m, fabs(ainv[(size_t)t * nc + jj]));
    scale[t] = m;
  }
  __syncthreads();
  const int nn = nc * nc;
  for (int k = 0; k < nc; ++k) {                                     // the three matrices step together
    if (threadIdx.x < 3) {
      const int c = threadIdx.x;
      const double pv = ainv[(size_t)c * nn + k * nc + k], sc = scale[c * nc + k];
      const int d = !(sc > 0.0) || !(pv > 1e-12 * sc);
      dead[c * nc + k] = d;
      piv_inv[c] = d ? 0.0 : 1.0 / pv;
    }
    __syncthreads();
    // column k of the other rows holds the multipliers; they are read here and rewritten after the barrier
    for (int t = threadIdx.x; t < 3 * nn; t += blockDim.x) {
      const int c = t / nn, ij = t % nn, ii = ij / nc, jj = ij % nc;
      const double pi = piv_inv[c];
      if (pi != 0.0 && ii != k && jj != k) {
        double* A = ainv + (size_t)c * nn;
        A[ij] -= A[ii * nc + k] * pi * A[k * nc + jj];
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 3 * nc; t += blockDim.x) {
      const int c = t / nc, ii = t % nc;
      const double pi = piv_inv[c];
      if (pi != 0.0) {
        double* A = ainv + (size_t)c * nn;
        if (ii != k) { A[ii * nc + k] = -A[ii * nc + k] * pi; A[k * nc + ii] = A[k * nc + ii] * pi; }
        else A[k * nc + k] = pi;
      }
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < 3 * nn; t += blockDim.x) {           // dead unknowns off
    const int c = t / nn, ij = t % nn;
    if (dead[c * nc + ij / nc] || dead[c * nc + ij % nc]) ainv[t] = 0.0;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * nn; t += blockDim.x) {           // symmetrise (Gauss-Jordan leaves rounding asymmetry)
    const int c = t / nn, ij = t % nn, ii = ij / nc, jj = ij % nc;
    if (ii < jj) {
      double* A = ainv + (size_t)c * nn;
      const double m = 0.5 * (A[ii * nc + jj] + A[jj * nc + ii]);
      A[ii * nc + jj] = m; A[jj * nc + ii] = m;
    }
  }
  __syncthreads();
  // ---- u0 = M^-1 b ---------------------------------------------------------------------------------------------
  coarse_restrict(q, gwarp, nwarps, lane);
  grid.sync();
  coarse_solve(q, ainv, yc, rc_s);
  if (row >= 0) {
    u0 = d0 * r0; u1 = d1 * r1; u2 = d2 * r2;
    if (in_p) {
      const double4 y = yc[row / q.bsz];
      if (d0 != 0.0) u0 += y.x;
      if (d1 != 0.0) u1 += y.y;
      if (d2 != 0.0) u2 += y.z;
    }
    st256(p.U + row, make_double4(u0, u1, u2, 0.0));
  }
  grid.sync();

  int it = 0;
  while (!sc_stop) {
    // ---- A: w = L u (software pipelined: the (col, w3) of batch k + 1 are requested before batch k's gathers are
    //      consumed), gamma = r.u, delta = u.w, |r|^2 --------------------------------------------------------------
    double w0 = 0, w1 = 0, w2 = 0;
    if (width > 0) {
      int c[4]; double4 w3[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int64_t o = base + (int64_t)t * kSellC;
        c[t] = __ldg(p.sell_col + o); w3[t] = ldg256(p.sell_w3 + o);
      }
      for (int j = 0; j < width; j += 4) {                              // widths are multiples of 4
        double4 uc[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) uc[t] = ld256(p.U + c[t]);
        int cn[4] = {0, 0, 0, 0}; double4 wn[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) wn[t] = make_double4(0, 0, 0, 0);
        if (j + 4 < width) {
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int64_t o = base + (int64_t)(j + 4 + t) * kSellC;
            cn[t] = __ldg(p.sell_col + o); wn[t] = ldg256(p.sell_w3 + o);
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          w0 += w3[t].x * (u0 - uc[t].x); w1 += w3[t].y * (u1 - uc[t].y); w2 += w3[t].z * (u2 - uc[t].z);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) { c[t] = cn[t]; w3[t] = wn[t]; }
      }
    }
    if (row >= 0) {
      v[0] = r0 * u0; v[1] = r1 * u1; v[2] = r2 * u2;
      v[3] = u0 * w0; v[4] = u1 * w1; v[5] = u2 * w2;
      v[6] = r0 * r0; v[7] = r1 * r1; v[8] = r2 * r2;
    } else {
#pragma unroll
      for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    // ---- B: p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s (registers); r published for the restriction ----
    if (row >= 0) {
      p0 = u0 + b0 * p0; p1 = u1 + b1 * p1; p2 = u2 + b2 * p2;
      s0 = w0 + b0 * s0; s1 = w1 + b1 * s1; s2 = w2 + b2 * s2;
      x0 += a0 * p0; x1 += a1 * p1; x2 += a2 * p2;
      r0 -= a0 * s0; r1 -= a1 * s1; r2 -= a2 * s2;
      st256(p.R + row, make_double4(r0, r1, r2, 0.0));
    }
    ++it;
    grid.sync();
    // ---- C: coarse residual; D: coarse solve and u = M^-1 r ---------------------------------------------------------
    coarse_restrict(q, gwarp, nwarps, lane);
    grid.sync();
    coarse_solve(q, ainv, yc, rc_s);
    if (row >= 0) {
      u0 = d0 * r0; u1 = d1 * r1; u2 = d2 * r2;
      if (in_p) {
        const double4 y = yc[row / q.bsz];
        if (d0 != 0.0) u0 += y.x;
        if (d1 != 0.0) u1 += y.y;
        if (d2 != 0.0) u2 += y.z;
      }
      st256(p.U + row, make_double4(u0, u1, u2, 0.0));
    }
    grid.sync();
  }
  if (row >= 0) st256(p.X + row, make_double4(x0, x1, x2, 0.0));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
  }
}

// irls on this path: the same weight for the three coordinates.
__global__ void __launch_bounds__(256)
k_w2_to_w3(const double* __restrict__ sell_w2, const double* __restrict__ diag, int64_t total, int n,
           double4* __restrict__ sell_w3, double4* __restrict__ diag3) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const double w = sell_w2[o];
    st256(sell_w3 + o, make_double4(w, w, w, 0.0));
  }
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const double d = diag[v];
    st256(diag3 + v, make_double4(d, d, d, 0.0));
  }
}

}  // namespace ira
