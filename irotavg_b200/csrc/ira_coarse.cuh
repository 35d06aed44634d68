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
//
// Tridiagonal variant (TRI, graphs of more than ~500 free nodes).  With 64 blocks the iteration count still grows like
// the block length (tools/precond_study.py --coarse-tri: 9 500-view stream graph, unit weights: 366 iterations with 64
// blocks, 108 with 250, 41 with 950).  For CONTIGUOUS index blocks P^T L P is tridiagonal up to the loop closures, and keeping only its tridiagonal
// part T (the far couplings stay on the diagonal: T is diagonally dominant, SPD) costs little: 138 / 96 / 78 / 65
// iterations with 250 / 500 / 950 / 1 900 blocks, the same again with the dense 64-block space added on top.  So
// nc grows to <= 1 024 blocks of >= 8 rows, and every thread block solves T y = r_c by cyclic reduction in its shared
// memory: the multipliers of the log2(nc) elimination levels are computed once per solve (5 doubles per coarse unknown
// and coordinate), an application is 2 log2(nc) + 1 short block-synchronous steps.  A dead pivot (floating component,
// empty block) switches that unknown off exactly as in the dense variant (T restricted to the others).  The SELL
// pattern of such a graph is built in index order, a block of the partition is a group of lanes of one warp, and the
// restriction happens in registers: three grid barriers per iteration instead of four.
#pragma once
#include "ira_l1ra.cuh"
#include "ira_plan.hpp"

namespace ira {

using plan::kCoarseMax;                 // coarse unknowns per coordinate, dense variant (ira_plan.hpp)
using plan::kCoarseMaxRows;
using plan::kTriMax;                    // tridiagonal variant
using plan::kTriMinBlock;
constexpr int kCoarseThreads = 384;     // >= kPcgNV warps: pcg_grid_reduce gives each of its 9 sums to one warp
static_assert(plan::kSellRows == kSellC, "ira_plan.hpp and the SELL kernels must agree on the slice height");

struct PcgCoarseParams {
  PcgW3Params w;                        // matrix (3 weights per entry), vectors, partials, ctl
  const int* sell_pos;                  // row -> SELL position
  int f;                                // rows < f are fixed (never unknowns, never in P)
  int nc, bsz;                          // coarse size and rows per block of the partition: block(v) = v / bsz
  double* AC;                           // [3][nc][nc] assembled coarse matrices (global scratch); TRI: [2][3][tri_n] diagonal, coupling to block a - 1
  double4* RC;                          // [nc] coarse residual of the current iteration
  int tri_n;                            // TRI: nc padded to a power of two (0: dense variant)
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


// ---- tridiagonal coarse operator -------------------------------------------------------------------------------------
// TRI runs on a SELL pattern built in INDEX order (ira_api.cu: build_sell with sell_index_order; view graphs have
// near-uniform degrees, so nothing is lost to padding): slice k holds rows 32k .. 32k + 31, a block of the partition
// (bsz = 8, 16 or 32 rows) is an aligned group of lanes of ONE warp, and r_c = P^T r comes straight from the registers:
// a fixed-order butterfly over the group, no barrier and no round trip of r through global memory.
__device__ __forceinline__ void coarse_restrict_tri(const PcgCoarseParams& q, int gwarp, int lane, double cx, double cy, double cz) {
  for (int o = q.bsz >> 1; o > 0; o >>= 1) {
    cx += __shfl_xor_sync(0xffffffffu, cx, o);
    cy += __shfl_xor_sync(0xffffffffu, cy, o);
    cz += __shfl_xor_sync(0xffffffffu, cz, o);
  }
  const int a = (gwarp * kSellC + lane) / q.bsz;
  if ((lane & (q.bsz - 1)) == 0 && a < q.nc) st256(q.RC + a, make_double4(cx, cy, cz, 0.0));
}

// Shared-memory layout of the TRI variant: six [3][N] arrays.
struct TriSmem {
  double *BI, *L1, *L2, *M1, *M2, *Y;      // 1 / pivot, coupling to j - s, coupling to j + s, forward multipliers, work vector
};
__device__ __forceinline__ TriSmem tri_smem(unsigned char* base, int N) {
  double* d = reinterpret_cast<double*>(base);
  TriSmem t;
  t.BI = d; t.L1 = d + 3 * N; t.L2 = d + 6 * N; t.M1 = d + 9 * N; t.M2 = d + 12 * N; t.Y = d + 15 * N;
  return t;
}

// Cyclic-reduction factorisation of the three tridiagonal matrices (diag TD, coupling to the previous unknown TL), in
// every block's shared memory, identical arithmetic everywhere.  Level l (stride s = 2^l): the active unknowns are
// i with (i + 1) % s == 0; those with (i + 1) % 2s == s are eliminated into their neighbours i +- s, which stay.
__device__ __forceinline__ void tri_factor(const PcgCoarseParams& q, const TriSmem& t) {
  const int N = q.tri_n;
  double* const D0 = t.Y;                                      // original diagonal: the scale of the dead-pivot test
  for (int k = threadIdx.x; k < 3 * N; k += blockDim.x) {
    const int c = k / N, a = k % N;
    const double d = a < q.nc ? __ldcg(q.AC + (size_t)c * N + a) : 1.0;
    t.BI[k] = d; D0[k] = d;
    t.L1[k] = a < q.nc ? __ldcg(q.AC + (size_t)(3 + c) * N + a) : 0.0;
    t.L2[k] = 0.0; t.M1[k] = 0.0; t.M2[k] = 0.0;
  }
  __syncthreads();
  for (int s = 1; s < N; s <<= 1) {
    const int cnt = N / (2 * s), off = N - N / s;
    for (int k = threadIdx.x; k < 3 * cnt; k += blockDim.x) {
      const int c = k / cnt, kk = k % cnt;
      const int i = 2 * s * (kk + 1) - 1, jl = i - s, jr = i + s;
      double* const B = t.BI + c * N; double* const lo = t.L1 + c * N;
      const double bl = B[jl], sl = D0[c * N + jl];
      const double il = (sl > 0.0 && bl > 1e-12 * sl) ? 1.0 / bl : 0.0;
      const double ci = lo[i];                                 // coupling i <-> jl
      const double k1 = ci * il;
      double k2 = 0.0, cr = 0.0;
      if (jr < N) {
        const double br = B[jr], sr = D0[c * N + jr];
        cr = lo[jr];                                           // coupling jr <-> i
        k2 = cr * ((sr > 0.0 && br > 1e-12 * sr) ? 1.0 / br : 0.0);
      }
      const double lo_jl = lo[jl];
      t.M1[c * N + off + kk] = k1; t.M2[c * N + off + kk] = k2;
      t.L2[c * N + jl] = ci;                                   // jl is solved at this level: its coupling to jl + s = i
      B[i] -= ci * k1 + cr * k2;
      lo[i] = -lo_jl * k1;                                     // coupling i <-> i - 2s
    }
    __syncthreads();
    // the eliminated unknowns keep 1 / pivot from here on (their pivots were read by both neighbours above; the next
    // level touches only unknowns that stay, so no barrier is needed after this pass)
    for (int k = threadIdx.x; k < 3 * cnt; k += blockDim.x) {
      const int c = k / cnt, j = s - 1 + 2 * s * (k % cnt);
      const double b = t.BI[c * N + j], sc = D0[c * N + j];
      t.BI[c * N + j] = (sc > 0.0 && b > 1e-12 * sc) ? 1.0 / b : 0.0;
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    const double b = t.BI[c * N + N - 1], sc = D0[c * N + N - 1];
    t.BI[c * N + N - 1] = (sc > 0.0 && b > 1e-12 * sc) ? 1.0 / b : 0.0;
  }
  __syncthreads();
}

// y_c = T^-1 r_c for the three coordinates: forward elimination with the stored multipliers, back substitution.
// N is a power of two: every index split is a shift.
__device__ __forceinline__ void tri_solve(const PcgCoarseParams& q, const TriSmem& t) {
  const int N = q.tri_n;
  const int lgN = 31 - __clz(N);
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    double4 r = make_double4(0, 0, 0, 0);
    if (a < q.nc) r = ld256(q.RC + a);
    t.Y[a] = r.x; t.Y[N + a] = r.y; t.Y[2 * N + a] = r.z;
  }
  __syncthreads();
  for (int l = 0; l < lgN; ++l) {
    const int s = 1 << l, lgc = lgN - l - 1, cnt = 1 << lgc, off = N - (N >> l);
    for (int k = threadIdx.x; k < 3 * cnt; k += blockDim.x) {
      const int c = k >> lgc, kk = k & (cnt - 1);
      const int i = ((kk + 1) << (l + 1)) - 1;
      double* const Y = t.Y + c * N;
      double y = Y[i] - t.M1[c * N + off + kk] * Y[i - s];
      if (i + s < N) y -= t.M2[c * N + off + kk] * Y[i + s];
      Y[i] = y;
    }
    __syncthreads();
  }
  if (threadIdx.x < 3) t.Y[threadIdx.x * N + N - 1] *= t.BI[threadIdx.x * N + N - 1];
  __syncthreads();
  for (int l = lgN - 1; l >= 0; --l) {
    const int s = 1 << l, lgc = lgN - l - 1, cnt = 1 << lgc;
    for (int k = threadIdx.x; k < 3 * cnt; k += blockDim.x) {
      const int c = k >> lgc, kk = k & (cnt - 1);
      const int j = s - 1 + (kk << (l + 1));
      double* const Y = t.Y + c * N;
      double y = Y[j] - t.L2[c * N + j] * Y[j + s];
      if (j >= s) y -= t.L1[c * N + j] * Y[j - s];
      Y[j] = y * t.BI[c * N + j];
    }
    __syncthreads();
  }
}

template <bool TRI>
__global__ void __launch_bounds__(kCoarseThreads, 1)
k_pcg_coarse_w3(const PcgCoarseParams q) {
  const PcgW3Params& p = q.w;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(32) unsigned char dyn_smem[];
  double* const ainv = reinterpret_cast<double*>(dyn_smem);                       // [3][nc][nc]          (dense variant)
  double4* const yc = reinterpret_cast<double4*>(ainv + ((3 * q.nc * q.nc + 3) & ~3));   // [nc], 32-byte aligned
  double4* const rc_s = yc + q.nc;                                                // [nc]
  const TriSmem tri = tri_smem(dyn_smem, q.tri_n);                                // six [3][tri_n] arrays (TRI variant)
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
  if (TRI) {
    // diagonal (everything that leaves block a, fixed nodes and far blocks included) and the coupling to block a - 1;
    // every lane walks its own row (index-ordered SELL), butterfly over the block's lanes: a fixed order.  T[a][a - 1] = T[a - 1][a] = the value row a computes.
    const int N = q.tri_n;
    if (gwarp < p.nslices) {
      double dg0 = 0, dg1 = 0, dg2 = 0, lo0 = 0, lo1 = 0, lo2 = 0;
      const int a = (gwarp * kSellC + lane) / q.bsz;
      if (in_p) {
        for (int j = 0; j < width; j += 4) {                          // widths are multiples of 4
          int col[4]; double4 w3[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            col[t] = __ldg(p.sell_col + base + (int64_t)(j + t) * kSellC);
            w3[t] = ldg256(p.sell_w3 + base + (int64_t)(j + t) * kSellC);
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (col[t] == row) continue;                             // padding slot
            const int bc = col[t] >= q.f ? col[t] / q.bsz : -1;      // a fixed node grounds the block
            if (bc != a) {
              dg0 += w3[t].x; dg1 += w3[t].y; dg2 += w3[t].z;
              if (bc >= 0 && bc == a - 1) { lo0 -= w3[t].x; lo1 -= w3[t].y; lo2 -= w3[t].z; }
            }
          }
        }
      }
      for (int o = q.bsz >> 1; o > 0; o >>= 1) {
        dg0 += __shfl_xor_sync(0xffffffffu, dg0, o); dg1 += __shfl_xor_sync(0xffffffffu, dg1, o); dg2 += __shfl_xor_sync(0xffffffffu, dg2, o);
        lo0 += __shfl_xor_sync(0xffffffffu, lo0, o); lo1 += __shfl_xor_sync(0xffffffffu, lo1, o); lo2 += __shfl_xor_sync(0xffffffffu, lo2, o);
      }
      if ((lane & (q.bsz - 1)) == 0 && a < nc) {
        q.AC[a] = dg0; q.AC[N + a] = dg1; q.AC[2 * N + a] = dg2;
        q.AC[3 * N + a] = lo0; q.AC[4 * N + a] = lo1; q.AC[5 * N + a] = lo2;
      }
    }
  } else
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
  __shared__ double piv_inv[3];
  __shared__ double scale[3 * kCoarseMax];
  __shared__ int dead[3 * kCoarseMax];
  if (TRI) {
    tri_factor(q, tri);
  } else {
  for (int t = threadIdx.x; t < 3 * nc * nc; t += blockDim.x) ainv[t] = __ldcg(q.AC + t);
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
  }
  // ---- u0 = M^-1 b ---------------------------------------------------------------------------------------------
  if (TRI) {
    if (gwarp < p.nslices)
      coarse_restrict_tri(q, gwarp, lane, in_p && d0 != 0.0 ? r0 : 0.0, in_p && d1 != 0.0 ? r1 : 0.0, in_p && d2 != 0.0 ? r2 : 0.0);
  } else {
    coarse_restrict(q, gwarp, nwarps, lane);
  }
  grid.sync();
  if (TRI) tri_solve(q, tri); else coarse_solve(q, ainv, yc, rc_s);
  if (row >= 0) {
    u0 = d0 * r0; u1 = d1 * r1; u2 = d2 * r2;
    if (in_p) {
      const int blk = row / q.bsz;
      const double4 y = TRI ? make_double4(tri.Y[blk], tri.Y[q.tri_n + blk], tri.Y[2 * q.tri_n + blk], 0.0) : yc[blk];
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
      if (!TRI) st256(p.R + row, make_double4(r0, r1, r2, 0.0));
    }
    ++it;
    // ---- C: coarse residual (TRI: from the registers, no barrier before it); D: coarse solve and u = M^-1 r ------------
    if (TRI) {
      if (gwarp < p.nslices)
        coarse_restrict_tri(q, gwarp, lane, in_p && d0 != 0.0 ? r0 : 0.0, in_p && d1 != 0.0 ? r1 : 0.0, in_p && d2 != 0.0 ? r2 : 0.0);
    } else {
      grid.sync();
      coarse_restrict(q, gwarp, nwarps, lane);
    }
    grid.sync();
    if (TRI) tri_solve(q, tri); else coarse_solve(q, ainv, yc, rc_s);
    if (row >= 0) {
      u0 = d0 * r0; u1 = d1 * r1; u2 = d2 * r2;
      if (in_p) {
        const int blk = row / q.bsz;
        const double4 y = TRI ? make_double4(tri.Y[blk], tri.Y[q.tri_n + blk], tri.Y[2 * q.tri_n + blk], 0.0) : yc[blk];
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
