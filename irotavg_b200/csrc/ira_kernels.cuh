// sm_100a kernels of the IRLS rotation-averaging hot path (multi-kernel pipeline).
//
// HBM layout (all FP64 / int32; "AoS4" = one 32-byte double4 record per edge or node so that a
// random gather of one endpoint is exactly one 32 B sector and one LDG.256):
//   I        int2  [m_pad]        edge endpoints (i, j) as the caller's std::pair<int,int> memory
//   QQ       double[4][m_pad]     relative rotations, column-major as the caller's Eigen matrix
//   weights  double[m_pad]        square-root IRLS weights (ral/l1_irls.cpp:577,617-727)
//   wres     double4[m_pad]       per edge (w_x, w_y, w_z, weights^2): residual + quadratic weight
//   Q        double4[n]           absolute rotations [x y z w]
//   CSR of A^T A over ALL n nodes (rows of fixed nodes are empty, their x stays 0):
//     rowptr int32[n+1], ent_col int32[nnz], ent_eid int32[nnz] (k for the +1 column, ~k for the
//     -1 column of row k of A; bit 30 of k flags an entry make_A drops but make_AtA keeps, see
//     k_csr_keys), ent_w2 double[nnz] (weights^2 of the entry's edge)
//   node vectors X, R, Z, P, AP, B: double4[n] (c0, c1, c2, pad) for the 3 right-hand sides.
//
// Reference lines restated are cited per kernel (paths relative to the reference tree).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ira_math.cuh"

namespace ira {

// -------------------------------------------------------------------------------------------
// PTX helpers: 256-bit global access (sm_100+), mbarrier + 1-D bulk async copy (TMA engine)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ double4 ldg256(const double4* p) {          // read-only path
  double4 v;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double4 ld256(const double4* p) {           // coherent
  double4 v;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st256(double4* p, const double4 v) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy executed by the TMA engine; completion is signalled on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// -------------------------------------------------------------------------------------------
// device-side control block of one irls() call
// -------------------------------------------------------------------------------------------
struct Ctl {
  double rz[3];        // r.z per column
  double alpha[3];
  double beta[3];
  double bnorm2[3];    // |b_c|^2
  double rnorm2[3];    // |r_c|^2
  double score;        // mean |X_i| over free nodes (ral/l1_irls.cpp:729)
  double rtol2;        // cg_rtol^2
  int cg_iters;
  int cg_max_iters;
  int done;            // 1: converged / capped; the remaining CG kernels of the chunk are no-ops
  unsigned int ticket; // last-block-done counter
  // persistent solve only: SM-clock cycles block 0 spent in the SpMV+reduce and update phases,
  // whole-kernel cycles and globaltimer nanoseconds (to convert cycles to time), SpMV phases run
  long long cyc_spmv, cyc_update, cyc_total, ns_total, pcg_spmv_phases;
  unsigned long long epoch;   // peer-memory solve (ira_peer.cuh): last cross-GPU barrier epoch used
};

constexpr int kRedMaxBlocks = 1184;      // 148 SMs x 8: upper bound on reduction grids
constexpr int kRedThreads = 256;

// Fixed-shape block reduction of NV doubles per thread (deterministic for a fixed launch shape).
template <int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sm /* [NV][32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) sm[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double t = lane < nwarps ? sm[k * 32 + lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      v[k] = t;
    }
  }
  __syncthreads();
}

// Grid-wide deterministic sum: each block publishes its NV partials, the block that takes the last
// ticket adds them in block order.  Returns true (block-uniformly) in that last block, with the
// totals in v[] valid for thread 0 of that block.
template <int NV>
__device__ __forceinline__ bool grid_reduce_last(double (&v)[NV], double* partials /* [grid][NV] */,
                                                 unsigned int* ticket, double* sm, int* sm_flag) {
  block_reduce<NV>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[blockIdx.x * NV + k] = v[k];
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    *sm_flag = (t == gridDim.x - 1);
  }
  __syncthreads();
  const bool last = *sm_flag != 0;
  if (last) {
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.0;
    for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] += __ldcg(&partials[b * NV + k]);
    }
    block_reduce<NV>(acc, sm);
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = acc[k];
    if (threadIdx.x == 0) *ticket = 0u;
  }
  return last;
}

// -------------------------------------------------------------------------------------------
// layout conversion at the boundary (column-major n x 4  <->  AoS4)
// -------------------------------------------------------------------------------------------
__global__ void k_colmajor_to_aos4(const double* __restrict__ src, int64_t ld, double4* __restrict__ dst,
                                   int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st256(dst + i, make_double4(src[i], src[ld + i], src[2 * ld + i], src[3 * ld + i]));
}
__global__ void k_aos4_to_colmajor(const double4* __restrict__ src, double* __restrict__ dst, int64_t ld,
                                   int64_t n, int ncols) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double4 v = ldg256(src + i);
    dst[i] = v.x; dst[ld + i] = v.y; dst[2 * ld + i] = v.z;
    if (ncols > 3) dst[3 * ld + i] = v.w;
  }
}
__global__ void k_fill_f64(double* __restrict__ p, double v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

// -------------------------------------------------------------------------------------------
// CSR pattern of A^T A (built once per upload; make_A's mask, ral/l1_irls.cpp:755-780)
// -------------------------------------------------------------------------------------------
// Two candidate entries per edge k = (i, j):  slot 2k   -> row j, column i   (+1 column of A's row k)
//                                             slot 2k+1 -> row i, column j   (-1 column)
// make_A (:755-780): row j exists iff j >= f; row i exists iff j >= f AND i >= f (the `continue` at :771).
// make_AtA (:811-848, the Newton matrix of l1ra) has no such `continue`: an edge (free i, fixed j) still
// loads H(i,i).  The pattern therefore keeps the row-i entry whenever i >= f and flags it (kEidQuirk)
// when j < f: irls() gives flagged entries weight 0, l1ra's Newton matrix uses them.
// Missing entries get the sentinel key n so a stable sort by key pushes them past the end.
constexpr int kEidQuirk = 1 << 30;
constexpr int kEidMask = kEidQuirk - 1;
__global__ void k_csr_keys(const int2* __restrict__ I, int64_t m, int n, int f, int* __restrict__ keys,
                           int* __restrict__ vals, int* __restrict__ bad) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
    const int2 e = I[k];
    if (e.x < 0 || e.x >= n || e.y < 0 || e.y >= n) { atomicMax(bad, 2); keys[2 * k] = n; keys[2 * k + 1] = n; }
    else if (e.x == e.y) { atomicMax(bad, 1); keys[2 * k] = n; keys[2 * k + 1] = n; }   // self-loop: rejected by the host
    else {
      keys[2 * k] = e.y >= f ? e.y : n;
      keys[2 * k + 1] = e.x >= f ? e.x : n;
    }
    vals[2 * k] = (int)(2 * k);
    vals[2 * k + 1] = (int)(2 * k + 1);
  }
}
// keys sorted (stable => entries of a row are in increasing edge order => deterministic sums).
__global__ void k_csr_finalize(const int* __restrict__ keys, const int* __restrict__ vals,
                               const int2* __restrict__ I, int64_t two_m, int n, int f, int* __restrict__ rowptr,
                               int* __restrict__ ent_col, int* __restrict__ ent_eid) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p <= two_m; p += (int64_t)gridDim.x * blockDim.x) {
    const int kprev = p > 0 ? keys[p - 1] : -1;
    const int kcur = p < two_m ? keys[p] : n + 1;
    if (kcur != kprev) {
      const int hi = kcur < n ? kcur : n;
      for (int r = kprev + 1; r <= hi; ++r) rowptr[r] = (int)p;
    }
    if (p < two_m && kcur < n) {
      const int src = vals[p];
      const int k = src >> 1;
      const int2 e = I[k];
      if (src & 1) { ent_col[p] = e.y; ent_eid[p] = ~(e.y < f ? (k | kEidQuirk) : k); }   // row i: neighbour j, A(k,i) = -1
      else         { ent_col[p] = e.x; ent_eid[p] = k; }    // row j: neighbour i, A(k,j) = +1
    }
  }
}

// -------------------------------------------------------------------------------------------
// residual kernel: w_k = Log( q~_j (x) QQ_k (x) Q_i )      (ral/l1_irls.cpp:592-593, 109-127, 498-532)
// -------------------------------------------------------------------------------------------
// Persistent CTAs walk 256-edge tiles.  One elected thread streams the tile's edge arrays
// (I, the 4 QQ columns, weights: 12 KB) into shared memory with 1-D bulk async copies (TMA
// engine, mbarrier completion), double-buffered, while the CTA's 256 threads each take one edge:
// two 32 B LDG.256 gathers of the endpoint quaternions, ~45 FP64 FMA + sqrt + atan2 + div, one
// 32 B STG.256 of (w, weights^2).  Algorithmic bytes per launch: 72 m + 56 n.
constexpr int kResTile = 256;
struct __align__(128) ResStage {
  int2 I[kResTile];
  double qq[4][kResTile];
  double wt[kResTile];
};
static_assert(sizeof(ResStage) == 12288, "stage size");

__global__ void __launch_bounds__(kResTile)
k_residual(const int2* __restrict__ I, const double* __restrict__ QQ, int64_t ldqq,
           const double* __restrict__ weights, const double4* __restrict__ Q, double4* __restrict__ wres,
           int64_t m, int ntiles, int store_theta) {
  __shared__ ResStage st[2];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int tile, int s) {
    const int64_t e0 = (int64_t)tile * kResTile;
    mbar_expect_tx(&bar[s], (uint32_t)sizeof(ResStage));
    bulk_g2s(st[s].I, I + e0, sizeof(int2) * kResTile, &bar[s]);
#pragma unroll
    for (int c = 0; c < 4; ++c) bulk_g2s(st[s].qq[c], QQ + c * ldqq + e0, sizeof(double) * kResTile, &bar[s]);
    bulk_g2s(st[s].wt, weights + e0, sizeof(double) * kResTile, &bar[s]);
  };

  int tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) issue(tile, 0);
  for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    const int next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) issue(next, s ^ 1);   // stage s^1 was released by the barrier below
    mbar_wait(&bar[s], (it >> 1) & 1);
    const int64_t e = (int64_t)tile * kResTile + tid;
    if (e < m) {
      const int2 ij = st[s].I[tid];
      const double4 qq = make_double4(st[s].qq[0][tid], st[s].qq[1][tid], st[s].qq[2][tid], st[s].qq[3][tid]);
      const double wt = st[s].wt[tid];
      const double4 qi = ldg256(Q + ij.x);
      const double4 qj = ldg256(Q + ij.y);
      double4 w = edge_residual(qi, qq, qj);
      if (!store_theta) w.w = wt * wt;
      st256(wres + e, w);
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------
// right-hand side, Jacobi diagonal and per-entry weights        (ral/l1_irls.cpp:596-610)
//   b = A^T D^2 w,  diag = diag(A^T D^2 A),  ent_w2[e] = weights[edge(e)]^2
// -------------------------------------------------------------------------------------------
// LPR lanes cooperate on one row: coalesced reads of the row's entries, one 32 B gather of
// (w_k, weights_k^2) per entry, butterfly reduction over the LPR lanes.
template <int LPR>
__global__ void __launch_bounds__(256)
k_rhs_diag(const int* __restrict__ rowptr, const int* __restrict__ ent_eid, const double4* __restrict__ wres,
           double* __restrict__ ent_w2, double4* __restrict__ B, double* __restrict__ diag, int n) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int base = warp * RPW; base < n; base += nwarps * RPW) {
    const int row = base + sub;
    double bx = 0, by = 0, bz = 0, d = 0;
    if (row < n) {
      const int e1 = rowptr[row + 1];
      for (int e = rowptr[row] + sl; e < e1; e += LPR) {
        const int eid = ent_eid[e];
        const bool neg = eid < 0;
        const int kk = neg ? ~eid : eid;
        double4 w = ldg256(wres + (kk & kEidMask));
        if (kk & kEidQuirk) w.w = 0.0;                       // make_A drops (free i, fixed j) edges
        ent_w2[e] = w.w;
        d += w.w;
        const double s = neg ? -w.w : w.w;
        bx += s * w.x; by += s * w.y; bz += s * w.z;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      bx += __shfl_xor_sync(0xffffffffu, bx, o);
      by += __shfl_xor_sync(0xffffffffu, by, o);
      bz += __shfl_xor_sync(0xffffffffu, bz, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if (row < n && sl == 0) {
      st256(B + row, make_double4(bx, by, bz, 0.0));
      diag[row] = d;
    }
  }
}

// CG start: x = 0, r = b, z = M^-1 r, p = z; rz = r.z, |b|^2.  (M = diag; rows with diag 0 - fixed
// nodes and nodes whose every edge has weight 0 - get M^-1 = 0 and stay at x = 0.)
__global__ void __launch_bounds__(kRedThreads)
k_cg_init(const double4* __restrict__ B, const double* __restrict__ diag, double* __restrict__ dinv,
          double4* __restrict__ X, double4* __restrict__ R, double4* __restrict__ Z, double4* __restrict__ P,
          int n, Ctl* ctl, double* partials, const int* __restrict__ mate, const double* __restrict__ pc1,
          const double* __restrict__ pc2, const int* __restrict__ mate2, const double* __restrict__ pc3) {
  __shared__ double sm[6 * 32];
  __shared__ int flag;
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 b = ldg256(B + i);
    const double d = diag[i];
    const double di = pc1 ? pc1[i] : (d > 0.0 ? 1.0 / d : 0.0);     // 2x2 block-Jacobi c1, or 1/d
    dinv[i] = di;
    double4 z = make_double4(di * b.x, di * b.y, di * b.z, 0.0);
    if (mate) {
      const int mt = mate[i];
      if (mt >= 0) {
        const double4 bm = ldg256(B + mt);
        const double c2 = pc2[i];
        z.x += c2 * bm.x; z.y += c2 * bm.y; z.z += c2 * bm.z;
        const int m2 = mate2[i];
        if (m2 >= 0) {
          const double4 b2 = ldg256(B + m2);
          const double c3 = pc3[i];
          z.x += c3 * b2.x; z.y += c3 * b2.y; z.z += c3 * b2.z;
        }
      }
    }
    st256(X + i, make_double4(0, 0, 0, 0));
    st256(R + i, b);
    st256(Z + i, z);
    st256(P + i, z);
    v[0] += b.x * z.x; v[1] += b.y * z.y; v[2] += b.z * z.z;
    v[3] += b.x * b.x; v[4] += b.y * b.y; v[5] += b.z * b.z;
  }
  if (grid_reduce_last<6>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    int zero = 1;
    for (int c = 0; c < 3; ++c) {
      ctl->rz[c] = v[c];
      ctl->bnorm2[c] = v[3 + c];
      ctl->rnorm2[c] = v[3 + c];
      ctl->alpha[c] = 0.0;
      ctl->beta[c] = 0.0;
      if (v[3 + c] > 0.0) zero = 0;
    }
    ctl->cg_iters = 0;
    ctl->done = zero || ctl->cg_max_iters <= 0;
  }
}

// -------------------------------------------------------------------------------------------
// SpMV: AP = (A^T D^2 A) P on 3 right-hand sides, + p.Ap and alpha       (ls_solve, :536-556)
//   (L p)_r = sum_{entries e of row r} w2_e (p_r - p_col(e))
// -------------------------------------------------------------------------------------------
// LPR lanes per row.  Per entry: 4 B column + 8 B weight streamed coalesced, one 32 B gather of
// the neighbour's p.  Algorithmic bytes per launch: 16 m + 48 n; as implemented
// 24 m (2 entries x 12 B) + 64 n streamed plus 64 m of gathered sectors served by L2.
// FUSE_DOT: also reduce p.Ap over the grid and let the last block write alpha (single-GPU path).
template <int LPR, bool FUSE_DOT>
__global__ void __launch_bounds__(256)
k_spmv(const int* __restrict__ rowptr, const int* __restrict__ ent_col, const double* __restrict__ ent_w2,
       const double4* __restrict__ P, double4* __restrict__ AP, int n, Ctl* ctl, double* partials) {
  if (ctl->done) return;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  double dot[3] = {0, 0, 0};
  for (int base = warp * RPW; base < n; base += nwarps * RPW) {
    const int row = base + sub;
    double ax = 0, ay = 0, az = 0;
    double4 pr = make_double4(0, 0, 0, 0);
    if (row < n) {
      const int e0 = rowptr[row], e1 = rowptr[row + 1];
      if (e1 > e0) pr = ldg256(P + row);
      for (int e = e0 + sl; e < e1; e += LPR) {
        const int c = ent_col[e];
        const double w2 = ent_w2[e];
        const double4 pc = ldg256(P + c);
        ax += w2 * (pr.x - pc.x);
        ay += w2 * (pr.y - pc.y);
        az += w2 * (pr.z - pc.z);
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
      az += __shfl_xor_sync(0xffffffffu, az, o);
    }
    if (row < n && sl == 0) {
      st256(AP + row, make_double4(ax, ay, az, 0.0));
      if (FUSE_DOT) { dot[0] += pr.x * ax; dot[1] += pr.y * ay; dot[2] += pr.z * az; }
    }
  }
  if (FUSE_DOT) {
    __shared__ double sm[3 * 32];
    __shared__ int flag;
    if (grid_reduce_last<3>(dot, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
      for (int c = 0; c < 3; ++c) ctl->alpha[c] = dot[c] > 0.0 ? ctl->rz[c] / dot[c] : 0.0;
    }
  }
}

// Sharded path: p.Ap after the all-reduce of AP.
__global__ void __launch_bounds__(kRedThreads)
k_cg_dot_pap(const double4* __restrict__ P, const double4* __restrict__ AP, int n, Ctl* ctl, double* partials) {
  if (ctl->done) return;
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  double v[3] = {0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 p = ldg256(P + i), a = ldg256(AP + i);
    v[0] += p.x * a.x; v[1] += p.y * a.y; v[2] += p.z * a.z;
  }
  if (grid_reduce_last<3>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) ctl->alpha[c] = v[c] > 0.0 ? ctl->rz[c] / v[c] : 0.0;
  }
}

// End of a PCG iteration (one thread): beta from the new r.z, convergence from |r|^2 (already in ctl).
__device__ __forceinline__ void cg_finish_iteration(Ctl* ctl, const double* rz_new) {
  int conv = 1;
  for (int c = 0; c < 3; ++c) {
    const double rz_old = ctl->rz[c];
    ctl->beta[c] = rz_old > 0.0 ? rz_new[c] / rz_old : 0.0;
    ctl->rz[c] = rz_new[c];
    if (!(ctl->rnorm2[c] <= ctl->rtol2 * ctl->bnorm2[c])) conv = 0;
  }
  const int it = ctl->cg_iters + 1;
  ctl->cg_iters = it;
  if (conv || it >= ctl->cg_max_iters) ctl->done = 1;
}

// x += alpha p; r -= alpha Ap; z = M^-1 r; then rz', |r|^2 -> beta, convergence flag.
__global__ void __launch_bounds__(kRedThreads)
k_cg_update(double4* __restrict__ X, double4* __restrict__ R, double4* __restrict__ Z,
            const double4* __restrict__ P, const double4* __restrict__ AP, const double* __restrict__ dinv,
            int n, Ctl* ctl, double* partials, int defer_z) {
  if (ctl->done) return;
  __shared__ double sm[6 * 32];
  __shared__ int flag;
  const double a0 = ctl->alpha[0], a1 = ctl->alpha[1], a2 = ctl->alpha[2];
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 p = ldg256(P + i), ap = ldg256(AP + i);
    double4 x = ld256(X + i), r = ld256(R + i);
    const double di = dinv[i];
    x.x += a0 * p.x; x.y += a1 * p.y; x.z += a2 * p.z;
    r.x -= a0 * ap.x; r.y -= a1 * ap.y; r.z -= a2 * ap.z;
    const double4 z = make_double4(di * r.x, di * r.y, di * r.z, 0.0);
    st256(X + i, x);
    st256(R + i, r);
    st256(Z + i, z);
    v[0] += r.x * z.x; v[1] += r.y * z.y; v[2] += r.z * z.z;
    v[3] += r.x * r.x; v[4] += r.y * r.y; v[5] += r.z * r.z;
  }
  if (grid_reduce_last<6>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) ctl->rnorm2[c] = v[3 + c];
    if (!defer_z) cg_finish_iteration(ctl, v);     // else k_cg_precond computes z and r.z
  }
}

// z = M^-1 r for the 2x2 block-Jacobi preconditioner (z_v = c1_v r_v + c2_v r_mate(v)); r must be
// complete, hence a kernel of its own after k_cg_update(defer_z = 1).  Then r.z -> beta, convergence.
__global__ void __launch_bounds__(kRedThreads)
k_cg_precond(const double4* __restrict__ R, double4* __restrict__ Z, const int* __restrict__ mate,
             const double* __restrict__ pc1, const double* __restrict__ pc2, int n, Ctl* ctl, double* partials,
             const int* __restrict__ mate2, const double* __restrict__ pc3) {
  if (ctl->done) return;
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  double v[3] = {0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 r = ldg256(R + i);
    const double c1 = pc1[i];
    double4 z = make_double4(c1 * r.x, c1 * r.y, c1 * r.z, 0.0);
    const int mt = mate[i];
    if (mt >= 0) {
      const double4 rm = ldg256(R + mt);
      const double c2 = pc2[i];
      z.x += c2 * rm.x; z.y += c2 * rm.y; z.z += c2 * rm.z;
      const int m2 = mate2[i];
      if (m2 >= 0) {
        const double4 r2 = ldg256(R + m2);
        const double c3 = pc3[i];
        z.x += c3 * r2.x; z.y += c3 * r2.y; z.z += c3 * r2.z;
      }
    }
    st256(Z + i, z);
    v[0] += r.x * z.x; v[1] += r.y * z.y; v[2] += r.z * z.z;
  }
  if (grid_reduce_last<3>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) cg_finish_iteration(ctl, v);
}

// p = z + beta p
__global__ void __launch_bounds__(256)
k_cg_p(double4* __restrict__ P, const double4* __restrict__ Z, int n, const Ctl* ctl) {
  if (ctl->done) return;
  const double b0 = ctl->beta[0], b1 = ctl->beta[1], b2 = ctl->beta[2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 z = ldg256(Z + i);
    double4 p = ld256(P + i);
    p.x = z.x + b0 * p.x; p.y = z.y + b1 * p.y; p.z = z.z + b2 * p.z;
    st256(P + i, p);
  }
}

// -------------------------------------------------------------------------------------------
// robust re-weighting: E_k = (A X)_k - w_k, weights_k = rho(|E_k|)        (ral/l1_irls.cpp:614-727)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_weights(const int2* __restrict__ I, const double4* __restrict__ wres, const double4* __restrict__ X,
          double* __restrict__ weights, int64_t m, int f, int cost, double sigma) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
    const int2 e = I[k];
    const double4 w = ldg256(wres + k);
    double ex = -w.x, ey = -w.y, ez = -w.z;
    if (e.y >= f) {                       // row k of A is empty when j is fixed (:770-771)
      const double4 xj = ldg256(X + e.y);
      ex += xj.x; ey += xj.y; ez += xj.z;
      if (e.x >= f) {
        const double4 xi = ldg256(X + e.x);
        ex -= xi.x; ey -= xi.y; ez -= xi.z;
      }
    }
    const double e2 = ex * ex + ey * ey + ez * ez;
    weights[k] = robust_weight(cost, sigma, e2, weights[k]);
  }
}

// -------------------------------------------------------------------------------------------
// node step: score = mean |X_i|, Q_i <- Q_i (x) Exp(X_i) for i >= f   (ral/l1_irls.cpp:729-737)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
k_update(double4* __restrict__ Q, const double4* __restrict__ X, int n, int f, Ctl* ctl, double* partials) {
  __shared__ double sm[32];
  __shared__ int flag;
  double v[1] = {0.0};
  for (int i = f + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 x = ldg256(X + i);
    double theta;
    const double4 dq = exp_quat(x.x, x.y, x.z, &theta);
    v[0] += theta;
    st256(Q + i, quat_mult(ld256(Q + i), dq));
  }
  if (grid_reduce_last<1>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0)
    ctl->score = v[0] / (double)(n - f);
}

// weights^2 into wres.w (probe path only)
__global__ void k_set_w2(double4* __restrict__ wres, const double* __restrict__ weights, int64_t m) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
    const double w = weights[k];
    reinterpret_cast<double*>(wres + k)[3] = w * w;
  }
}
// free-node n_free x 3 column-major  <->  AoS4 over all nodes (fixed rows zero)
__global__ void k_free_to_aos4(const double* __restrict__ src, int64_t nf, int f, double4* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double4 v = make_double4(0, 0, 0, 0);
    if (i >= f) { const int64_t r = i - f; v.x = src[r]; v.y = src[nf + r]; v.z = src[2 * nf + r]; }
    st256(dst + i, v);
  }
}

}  // namespace ira
