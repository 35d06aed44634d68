// Persistent (cooperative) PCG solve on a SELL-32-sigma copy of the CSR pattern.
//
// Why: the multi-kernel PCG of ira_kernels.cuh pays three launches and ~40 us per iteration at
// config 3, and its sub-warp-per-row SpMV is latency bound (ncu: long_scoreboard, 46 % L1TEX).
// Here one cooperative kernel runs the whole linear solve of one IRLS iteration
// (replaces ls_solve, ral/l1_irls.cpp:536-556):
//   * SELL-C-sigma, C = 32: rows are sorted by degree inside windows of 1024 rows, packed 32 to a
//     slice, entries stored slice-column-major, so a warp reads its slice's (col, w2) streams as
//     full 128 B / 256 B lines and every lane owns one row: no cross-lane reduction, no idle
//     lanes beyond the (small) padding, 4 independent 32 B gathers in flight per lane;
//   * Chronopoulos-Gear single-reduction CG: one fused grid reduction (gamma = r.u, delta = u.Au,
//     |r|^2) and two grid barriers per iteration; alpha/beta are recomputed redundantly by every
//     thread from block-ordered partial sums, so the iteration is deterministic;
//   * 3 right-hand sides share every SpMV, each with its own alpha/beta.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ira_kernels.cuh"

namespace ira {
namespace cg = cooperative_groups;

constexpr int kSellC = 32;          // rows per slice (one warp)
constexpr int kSellSigma = 1024;    // sorting window (rows)
constexpr int kSellPad = INT32_MIN; // ent_eid marker of a padding slot
constexpr int kPcgThreads = 768;   // 24 warps, <= 85 registers per thread

// ---- SELL construction ------------------------------------------------------------------------
// One block per window: bitonic sort of (degree desc, row asc); writes the row of every sorted
// position (-1 beyond n) and the width (max degree) of every slice.
__global__ void __launch_bounds__(kSellSigma)
k_sell_sort(const int* __restrict__ rowptr, int n, int* __restrict__ sell_row, int* __restrict__ slice_width,
            int* __restrict__ slice_cnt, int index_order) {
  __shared__ unsigned long long key[kSellSigma];
  const int t = threadIdx.x;
  const int row = blockIdx.x * kSellSigma + t;
  const unsigned int deg = row < n ? (unsigned int)(rowptr[row + 1] - rowptr[row]) : 0u;
  if (index_order) {
    // chain-like view graphs on the two-level TRI path (ira_coarse.cuh): position = row, a slice is as wide as its
    // widest row (degrees are near-uniform there)
    sell_row[row] = row < n ? row : -1;
    unsigned int w = deg;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if ((t & (kSellC - 1)) == 0) {
      const int wd = ((int)w + 3) & ~3;
      slice_width[row / kSellC] = wd;
      slice_cnt[row / kSellC] = wd * kSellC;
    }
    return;
  }
  // descending sort of this key = degree descending, then original order ascending; rows >= n last
  key[t] = row < n ? (((unsigned long long)deg + 1ull) << 32) | (unsigned long long)(kSellSigma - 1 - t) : 0ull;
  __syncthreads();
  for (int k = 2; k <= kSellSigma; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = t ^ j;
      if (ixj > t) {
        const unsigned long long a = key[t], b = key[ixj];
        const bool desc = (t & k) == 0;
        if (desc ? (a < b) : (a > b)) { key[t] = b; key[ixj] = a; }
      }
      __syncthreads();
    }
  }
  const unsigned long long kk = key[t];
  const int pos = blockIdx.x * kSellSigma + t;
  const int src = kk ? blockIdx.x * kSellSigma + (kSellSigma - 1 - (int)(kk & 0xffffffffull)) : -1;
  sell_row[pos] = src;
  if ((t & (kSellC - 1)) == 0) {
    const int w = kk ? (((int)(kk >> 32) - 1 + 3) & ~3) : 0;      // multiple of 4 (sell_row_apply)
    slice_width[pos / kSellC] = w;
    slice_cnt[pos / kSellC] = w * kSellC;
  }
}

// Copy the CSR entries of every sorted row into its slice column-major slots.
__global__ void __launch_bounds__(256)
k_sell_fill(const int* __restrict__ rowptr, const int* __restrict__ ent_col, const int* __restrict__ ent_eid,
            const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
            int npos, int* __restrict__ sell_col, int* __restrict__ sell_eid, double* __restrict__ sell_w2) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= npos) return;
  const int s = pos / kSellC, lane = pos % kSellC;
  const int row = sell_row[pos];
  const int width = slice_width[s];
  const int64_t base = (int64_t)slice_off[s] + lane;
  int e0 = 0, deg = 0;
  if (row >= 0) { e0 = rowptr[row]; deg = rowptr[row + 1] - e0; }
  for (int j = 0; j < width; ++j) {
    const int64_t o = base + (int64_t)j * kSellC;
    if (j < deg) { sell_col[o] = ent_col[e0 + j]; sell_eid[o] = ent_eid[e0 + j]; }
    else { sell_col[o] = row >= 0 ? row : 0; sell_eid[o] = kSellPad; }   // p_r - p_r = 0, w2 = 0
    sell_w2[o] = 0.0;
  }
}

// ---- rhs / diagonal / per-entry weights on the SELL pattern  (ral/l1_irls.cpp:596-610) ---------
__global__ void __launch_bounds__(256)
k_sell_rhs(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
           const int* __restrict__ sell_eid, const double4* __restrict__ wres, double* __restrict__ sell_w2,
           double4* __restrict__ B, double* __restrict__ diag, int nslices) {
  const int lane = threadIdx.x & 31;
  const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int s = gwarp; s < nslices; s += nwarps) {
    const int row = sell_row[s * kSellC + lane];
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    double bx = 0, by = 0, bz = 0, d = 0;
    for (int j = 0; j < width; j += 4) {                     // widths are multiples of 4: 4 independent gathers in flight
      int eid[4]; double4 w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) eid[q] = __ldg(sell_eid + base + (int64_t)(j + q) * kSellC);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kk = eid[q] < 0 ? ~eid[q] : eid[q];
        w[q] = eid[q] != kSellPad ? ldg256(wres + (kk & kEidMask)) : make_double4(0, 0, 0, 0);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        double w2 = 0.0;
        if (eid[q] != kSellPad) {
          const bool neg = eid[q] < 0;
          const int kk = neg ? ~eid[q] : eid[q];
          w2 = (kk & kEidQuirk) ? 0.0 : w[q].w;                // make_A drops (free i, fixed j) edges
          d += w2;
          const double sg = neg ? -w2 : w2;
          bx += sg * w[q].x; by += sg * w[q].y; bz += sg * w[q].z;
        }
        sell_w2[base + (int64_t)(j + q) * kSellC] = w2;
      }
    }
    if (row >= 0) {
      st256(B + row, make_double4(bx, by, bz, 0.0));
      diag[row] = d;
    }
  }
}

// Gather flavours (experiment knob, ira_options.spmv_variant): 0 = plain coherent ld.global (L1
// allocating), 1 = ld.global.L1::no_allocate (coherent at L2; the gathered vector has ~5 % L1 hit rate
// on a random graph, so allocation only churns the cache), 2 = ld.global.nc (read-only path; not legal
// inside the persistent kernel, where the vector is rewritten between barriers).
template <int V>
__device__ __forceinline__ double4 gather256(const double4* p) {
  double4 v;
  if (V == 1)
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  else if (V == 2)
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  else
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}

// One lane's row of (A^T D^2 A) v: sum_e w2_e (v_row - v_col(e)) over the slice-column-major slots
// base, base+32, ...  Slice widths are multiples of 4 (padding slots have col = row, w2 = 0).
// Software pipelined: the (col, w2) of batch k+1 are requested before batch k's gathers are
// consumed, so the dependent chain per batch is one gather latency, with 4 (or 8) independent 32 B
// gathers in flight per lane.
template <int V, int UNR>
__device__ __forceinline__ void sell_row_apply(const int* __restrict__ sell_col, const double* __restrict__ sell_w2,
                                               const double4* V_, int64_t base, int width, const double4 u,
                                               double& ax, double& ay, double& az) {
  ax = 0.0; ay = 0.0; az = 0.0;
  if (width <= 0) return;
  int c[4]; double w2[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t o = base + (int64_t)q * kSellC;
    c[q] = __ldg(sell_col + o);
    w2[q] = __ldg(sell_w2 + o);
  }
  int j = 0;
  if (UNR == 8) {
    for (; j + 8 <= width; j += 8) {
      int c2[4]; double w22[4]; double4 ua[4], ub[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t o = base + (int64_t)(j + 4 + q) * kSellC;
        c2[q] = __ldg(sell_col + o);
        w22[q] = __ldg(sell_w2 + o);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) ua[q] = gather256<V>(V_ + c[q]);
#pragma unroll
      for (int q = 0; q < 4; ++q) ub[q] = gather256<V>(V_ + c2[q]);
      int cn[4] = {0, 0, 0, 0}; double wn[4] = {0, 0, 0, 0};
      if (j + 8 < width) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t o = base + (int64_t)(j + 8 + q) * kSellC;
          cn[q] = __ldg(sell_col + o);
          wn[q] = __ldg(sell_w2 + o);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ax += w2[q] * (u.x - ua[q].x); ay += w2[q] * (u.y - ua[q].y); az += w2[q] * (u.z - ua[q].z);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ax += w22[q] * (u.x - ub[q].x); ay += w22[q] * (u.y - ub[q].y); az += w22[q] * (u.z - ub[q].z);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { c[q] = cn[q]; w2[q] = wn[q]; }
    }
  }
  for (; j < width; j += 4) {
    double4 uc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) uc[q] = gather256<V>(V_ + c[q]);
    int cn[4] = {0, 0, 0, 0}; double wn[4] = {0, 0, 0, 0};
    if (j + 4 < width) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t o = base + (int64_t)(j + 4 + q) * kSellC;
        cn[q] = __ldg(sell_col + o);
        wn[q] = __ldg(sell_w2 + o);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ax += w2[q] * (u.x - uc[q].x); ay += w2[q] * (u.y - uc[q].y); az += w2[q] * (u.z - uc[q].z);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = cn[q]; w2[q] = wn[q]; }
  }
}

// Stand-alone SELL SpMV (multi-kernel / sharded path and the roofline probe): AP = (A^T D^2 A) P,
// optionally fused with p.Ap and alpha exactly like k_spmv.  Slices are dealt to blocks round-robin
// (slice s -> block s % gridDim.x) so that every SM gets the same number of slices.
template <bool FUSE_DOT, int V, int UNR>
__global__ void __launch_bounds__(256)
k_spmv_sell(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
            const int* __restrict__ sell_col, const double* __restrict__ sell_w2, const double4* __restrict__ P,
            double4* __restrict__ AP, int nslices, Ctl* ctl, double* partials) {
  if (ctl->done) return;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double dot[3] = {0, 0, 0};
  for (int s = blockIdx.x + gridDim.x * warp; s < nslices; s += gridDim.x * wpb) {
    const int row = sell_row[s * kSellC + lane];
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    const double4 u = row >= 0 ? gather256<V>(P + row) : make_double4(0, 0, 0, 0);
    double ax, ay, az;
    sell_row_apply<V, UNR>(sell_col, sell_w2, P, base, width, u, ax, ay, az);
    if (row >= 0) {
      st256(AP + row, make_double4(ax, ay, az, 0.0));
      if (FUSE_DOT) { dot[0] += u.x * ax; dot[1] += u.y * ay; dot[2] += u.z * az; }
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

// ---- pairwise block-Jacobi ("mate") preconditioner ----------------------------------------------
// Robust IRLS weights (L1 above all: weights^2 = 1/|E| up to 1e8) tie a few nodes together with edges
// 10^3..10^6 times stiffer than the rest; Jacobi then leaves eigenvalues ~ soft/stiff and PCG needs
// 10^3..10^4 iterations.  Pair every node with its strongest neighbour when the pick is mutual and
// the normalised strength w2_vu / sqrt(d_v d_u) >= theta, and invert the pair's 2x2 diagonal block
// exactly (block-Jacobi with blocks of size 1 and 2).
// (numpy study: 5 600 -> ~65 PCG iterations at n = 10k after 20 L1 iterations.)
__global__ void __launch_bounds__(256)
k_pair_best(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
            const int* __restrict__ sell_col, const double* __restrict__ sell_w2, const double* __restrict__ diag,
            int nslices, double theta, unsigned long long* __restrict__ pair_key, double* __restrict__ pair_w2) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int s = blockIdx.x + gridDim.x * warp; s < nslices; s += gridDim.x * wpb) {
    const int row = sell_row[s * kSellC + lane];
    if (row < 0) continue;
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    const double dr = diag[row];
    unsigned long long best = 0ull;
    double bw = 0.0;
    if (dr > 0.0) {
      for (int j = 0; j < width; j += 4) {                   // 4 independent diag gathers in flight
        double w2[4], dc[4]; int c[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t o = base + (int64_t)(j + q) * kSellC;
          w2[q] = sell_w2[o];
          c[q] = sell_col[o];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) dc[q] = diag[c[q]];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (w2[q] > 0.0 && c[q] != row && dc[q] > 0.0) {      // fixed nodes (diag 0) never pair
            const float st = (float)(w2[q] / sqrt(dr * dc[q]));
            if (st >= (float)theta) {
              const unsigned long long key = ((unsigned long long)__float_as_uint(st) << 32) | (unsigned long long)(0xffffffffu - (unsigned)c[q]);
              if (key > best) { best = key; bw = w2[q]; }
            }
          }
        }
      }
    }
    pair_key[row] = best;
    pair_w2[row] = bw;
  }
}

// Sharded path: after the all-reduce(MAX) of pair_key, a rank keeps its candidate weight only if its
// local pick is the global one (then all-reduce(MAX) of pair_w2 delivers it to every rank).
__global__ void k_pair_select(const unsigned long long* __restrict__ key_local, const unsigned long long* __restrict__ key_global,
                              double* __restrict__ pair_w2, int n) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x)
    if (key_local[v] != key_global[v]) pair_w2[v] = 0.0;
}

// Mutual picks become 2x2 diagonal blocks [[d_v, -w], [-w, d_u]] of the block-Jacobi preconditioner;
// their exact inverse is applied as  z_v = c1_v r_v + c2_v r_mate(v)  with
//   det = d_v d_u - w^2 = w (a + b) + a b   (a = d_v - w, b = d_u - w: no cancellation for stiff w),
//   c1_v = d_u / det,  c2_v = w / det.      Unpaired nodes keep Jacobi: c1 = 1/d (0 if d = 0), c2 = 0.
__global__ void __launch_bounds__(256)
k_pair_mate(const unsigned long long* __restrict__ pair_key, const double* __restrict__ pair_w2,
            const double* __restrict__ diag, int n, int* __restrict__ mate, double* __restrict__ pc1,
            double* __restrict__ pc2, int* __restrict__ npairs, int* __restrict__ mate2, double* __restrict__ pc3) {
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    int m = -1;
    const double dv = diag[v];
    double c1 = dv > 0.0 ? 1.0 / dv : 0.0, c2 = 0.0;
    const unsigned long long kv = pair_key[v];
    if (kv) {
      const int u = (int)(0xffffffffu - (unsigned)(kv & 0xffffffffull));
      const unsigned long long ku = pair_key[u];
      if (ku && (int)(0xffffffffu - (unsigned)(ku & 0xffffffffull)) == v && (ku >> 32) == (kv >> 32)) {
        const double du = diag[u];
        const double w = pair_w2[v < u ? v : u];            // one value for both members: symmetric M
        const double a = dv - w, b = du - w;
        const double det = w * (a + b) + a * b;
        if (a >= 0.0 && b >= 0.0 && det > 1e-12 * dv * du) { // a floating 2-node component has det = 0
          m = u; c1 = du / det; c2 = w / det;
          if (v < u) atomicAdd(npairs, 1);
        }
      }
    }
    mate[v] = m;
    pc1[v] = c1;
    pc2[v] = c2;
    mate2[v] = -1;
    pc3[v] = 0.0;
  }
}

// ---- third member: 3x3 blocks -------------------------------------------------------------------------
// With L1 weights a stiff edge rarely comes alone: a node tied to TWO neighbours by edges 10^3..10^6 times
// stiffer than the rest is as common as an isolated stiff pair, and a 2x2 block cannot absorb it (numpy study,
// n = 20k after 29 L1 iterations: Jacobi 8 765, pairs 110, pairs + third member 39 PCG iterations).  After the
// mutual pairs are fixed, every still-single node proposes to the pair that holds its strongest neighbour
// (normalised strength >= theta3); a pair takes its strongest applicant (atomicMax on a strength|node key:
// deterministic) and the 3x3 diagonal block is inverted exactly.
__global__ void __launch_bounds__(256)
k_attach_best(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
              const int* __restrict__ sell_col, const double* __restrict__ sell_w2, const double* __restrict__ diag,
              const int* __restrict__ mate, int nslices, double theta3, unsigned long long* __restrict__ att_key,
              const int* __restrict__ npairs) {
  if (*npairs == 0) return;                               // no 2x2 block to join (mild weights: L2, Geman-McClure ...)
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int s = blockIdx.x + gridDim.x * warp; s < nslices; s += gridDim.x * wpb) {
    const int row = sell_row[s * kSellC + lane];
    if (row < 0 || mate[row] >= 0) continue;
    const double dr = diag[row];
    if (!(dr > 0.0)) continue;
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    float best = 0.f;
    int leader = -1;
    for (int j = 0; j < width; j += 4) {
      double w2[4], dc[4]; int c[4], mc[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t o = base + (int64_t)(j + q) * kSellC;
        w2[q] = sell_w2[o];
        c[q] = sell_col[o];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) { mc[q] = mate[c[q]]; dc[q] = diag[c[q]]; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (w2[q] > 0.0 && c[q] != row && mc[q] >= 0) {
          const float st = (float)(w2[q] / sqrt(dr * dc[q]));
          const int ld = c[q] < mc[q] ? c[q] : mc[q];
          if (st >= (float)theta3 && (st > best || (st == best && ld < leader))) { best = st; leader = ld; }
        }
      }
    }
    if (leader >= 0)
      atomicMax(att_key + leader, ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xffffffffu - (unsigned)row));
  }
}

// One thread per pair leader a (a < b = mate[a]) that got an applicant t: exact inverse of
//   [[da, -wab, -wat], [-wab, db, -wbt], [-wat, -wbt, dt]]
// by eliminating a, written with the "excess" e = d - (block couplings) >= 0 so that stiff couplings never
// cancel: Schur complement of (b, t) = [[w' + xb, -w'], [-w', w' + xt]], w' = wbt + wab wat / da,
// xb = eb + wab ea / da, xt = et + wat ea / da, det = w' (xb + xt) + xb xt; every entry of the inverse is a
// sum of non-negative terms.
__global__ void __launch_bounds__(256)
k_attach_block(const unsigned long long* __restrict__ att_key, const int* __restrict__ sell_pos,
               const int* __restrict__ slice_off, const int* __restrict__ slice_width, const int* __restrict__ sell_col,
               const double* __restrict__ sell_w2, const double* __restrict__ diag, const double* __restrict__ pair_w2,
               int n, int* __restrict__ mate, int* __restrict__ mate2, double* __restrict__ pc1, double* __restrict__ pc2,
               double* __restrict__ pc3, const int* __restrict__ npairs) {
  if (*npairs == 0) return;
  for (int a = blockIdx.x * blockDim.x + threadIdx.x; a < n; a += gridDim.x * blockDim.x) {
    const unsigned long long key = att_key[a];
    if (!key) continue;
    const int b = mate[a];
    if (b < 0 || b < a) continue;                        // keys are only ever posted to leaders
    const int t = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
    const int pos = sell_pos[t];
    const int sl = pos / kSellC, ln = pos % kSellC;
    const int width = slice_width[sl];
    const int64_t base = (int64_t)slice_off[sl] + ln;
    double wat = 0.0, wbt = 0.0;
    for (int j = 0; j < width; ++j) {
      const int64_t o = base + (int64_t)j * kSellC;
      const int c = sell_col[o];
      if (c == a) wat += sell_w2[o];
      else if (c == b) wbt += sell_w2[o];
    }
    const double da = diag[a], db = diag[b], dt = diag[t], wab = pair_w2[a];
    const double ea = fmax(da - wab - wat, 0.0), eb = fmax(db - wab - wbt, 0.0), et = fmax(dt - wat - wbt, 0.0);
    const double lb = wab / da, lt = wat / da;
    const double xb = eb + lb * ea, xt = et + lt * ea;
    const double wp = wbt + lb * wat;
    const double det = wp * (xb + xt) + xb * xt;
    if (!(det > 1e-12 * db * dt)) continue;              // floating 3-node component: keep the pair
    const double sbb = (wp + xt) / det, sbt = wp / det, stt = (wp + xb) / det;
    const double iab = lb * sbb + lt * sbt, iat = lb * sbt + lt * stt;
    const double iaa = 1.0 / da + lb * iab + lt * iat;
    mate2[a] = t; pc1[a] = iaa; pc2[a] = iab; pc3[a] = iat;
    mate2[b] = t; pc1[b] = sbb; pc2[b] = iab; pc3[b] = sbt;
    mate[t] = a; mate2[t] = b; pc1[t] = stt; pc2[t] = iat; pc3[t] = sbt;
  }
}

// ---- the persistent solve ----------------------------------------------------------------------
struct PcgParams {
  int n, nslices, max_iters;
  double rtol2;
  const int* sell_row; const int* slice_off; const int* slice_width;
  const int* sell_col; const double* sell_w2;
  const double4* B; const double* diag;
  double4 *X, *R, *U, *W, *P, *S;
  double* dinv;
  const int* mate; const double* pc1; const double* pc2; const int* npairs;   // 2x2 / 3x3 block-Jacobi (null: Jacobi)
  const int* mate2; const double* pc3;                                         // third member of a 3x3 block, or -1
  int debug;                                                                   // ira_options.profile == 2: print block 0's phase split
  double* partials;      // [gridDim.x][kPcgNV]
  Ctl* ctl;
};
constexpr int kPcgNV = 9;

// Block-ordered grid reduction through one grid barrier; every thread of every block returns the
// same totals (bitwise), so loop control and alpha/beta stay uniform without a broadcast.
// Needs at least kPcgNV warps per block (warp k reduces sum k).
__device__ __forceinline__ void pcg_grid_reduce(double (&v)[kPcgNV], double* partials, cg::grid_group& grid,
                                                double* red /* [kPcgNV*32] */, double* tot /* [kPcgNV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp < kPcgNV) {
    double t = lane < nw ? red[warp * 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) __stcg(&partials[blockIdx.x * kPcgNV + warp], t);
  }
  grid.sync();
  if (warp < kPcgNV) {
    double t = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&partials[b * kPcgNV + warp]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) tot[warp] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = tot[k];
}

// Chronopoulos-Gear coefficients of one iteration from the reduced totals (in shared memory, as pcg_grid_reduce
// leaves them) v = {r.u (3), u.Au (3), |r|^2 (3)}: lanes
// 0..2 of the block own one right-hand side each (three division chains side by side instead of nine in a row); every
// block computes the same numbers from the same block-ordered totals.  A column that has converged
// (|r_c| <= rtol |b_c|) is FROZEN: alpha_c = beta_c = 0 from then on, so its solution stays bit-fixed while the others
// finish (round-off in gamma / delta of a converged column cannot perturb it any more).  Raises *sc_stop when all three
// are frozen or the iteration cap is reached.  Ends with a block barrier.
__device__ __forceinline__ void pcg_coefficients(const double* v /* shared: the 9 totals */, int it, int max_iters, double rtol2, double* sc_bb,
                                                 double* sc_go, double* sc_ao, double* sc_a, double* sc_b, double* sc_rr,
                                                 int* sc_stop) {
  __shared__ int frozen[3];
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    if (it == 0) frozen[c] = !(sc_bb[c] > 0.0);                 // zero right-hand side: x_c = 0
    if (!frozen[c]) {
      sc_rr[c] = v[6 + c];
      if (v[6 + c] <= rtol2 * sc_bb[c]) frozen[c] = 1;
    }
    double alpha = 0.0, beta = 0.0;
    if (!frozen[c]) {
      const double gam = v[c], del = v[3 + c];
      double den = del;
      if (it > 0) {
        beta = sc_go[c] > 0.0 ? gam / sc_go[c] : 0.0;
        if (sc_ao[c] != 0.0) den = del - beta * gam / sc_ao[c];
      }
      alpha = den > 0.0 ? gam / den : 0.0;
      sc_go[c] = gam; sc_ao[c] = alpha;
    }
    sc_a[c] = alpha; sc_b[c] = beta;
  }
  __syncthreads();
  if (threadIdx.x == 0 && ((frozen[0] && frozen[1] && frozen[2]) || it >= max_iters)) *sc_stop = 1;
  __syncthreads();
}

template <int V, int UNR>
__global__ void __launch_bounds__(kPcgThreads, 1)
k_pcg_persistent(const PcgParams p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  const int lane = threadIdx.x & 31;
  // slices are dealt to blocks round-robin: every SM owns the same number of slices (+-1)
  const int gwarp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  double v[kPcgNV];
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;      // grid-uniform

  // ---- start: x = 0, r = b, u = M^-1 r, p = s = 0; |b|^2 ---------------------------------------
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  for (int s = gwarp; s < p.nslices; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 b = ldg256(p.B + row);
      const double d = p.diag[row];
      const double di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);   // c1 of the 2x2 block, or 1/d
      p.dinv[row] = di;
      const double4 z4 = make_double4(0, 0, 0, 0);
      st256(p.X + row, z4); st256(p.P + row, z4); st256(p.S + row, z4);
      st256(p.R + row, b);
      double4 u0 = make_double4(di * b.x, di * b.y, di * b.z, 0.0);
      if (has_pairs) {
        const int mt = p.mate[row];
        if (mt >= 0) {                                      // B is complete (written by the previous kernel)
          const double4 bm = ldg256(p.B + mt);
          const double c2 = p.pc2[row];
          u0.x += c2 * bm.x; u0.y += c2 * bm.y; u0.z += c2 * bm.z;
          const int m2 = p.mate2[row];
          if (m2 >= 0) {
            const double4 b2 = ldg256(p.B + m2);
            const double c3 = p.pc3[row];
            u0.x += c3 * b2.x; u0.y += c3 * b2.y; u0.z += c3 * b2.z;
          }
        }
      }
      st256(p.U + row, u0);
      v[0] += b.x * b.x; v[1] += b.y * b.y; v[2] += b.z * b.z;
    }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);     // also publishes U to the whole grid
  // loop-carried scalars live in shared memory (thread 0 updates them) to keep 1024 threads at 64 regs
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);   // zero right-hand side: x = 0
  }
  __syncthreads();
  int it = 0;
  // phase clocks of block 0 (every block runs the same phases between the same barriers)
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  while (!sc_stop) {
    // ---- w = A u (SpMV), gamma = r.u, delta = u.w, |r|^2 ---------------------------------------
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    for (int s = gwarp; s < p.nslices; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      const int width = p.slice_width[s];
      const int64_t base = (int64_t)p.slice_off[s] + lane;
      const double4 u = row >= 0 ? ld256(p.U + row) : make_double4(0, 0, 0, 0);
      double ax, ay, az;
      sell_row_apply<V, UNR>(p.sell_col, p.sell_w2, p.U, base, width, u, ax, ay, az);
      if (row >= 0) {
        st256(p.W + row, make_double4(ax, ay, az, 0.0));
        const double4 r = ld256(p.R + row);
        v[0] += r.x * u.x; v[1] += r.y * u.y; v[2] += r.z * u.z;
        v[3] += u.x * ax;  v[4] += u.y * ay;  v[5] += u.z * az;
        v[6] += r.x * r.x; v[7] += r.y * r.y; v[8] += r.z * r.z;
      }
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    if (timer) { const long long c = clock64(); c_spmv += c - c_mark; c_mark = c; }
    // ---- Chronopoulos-Gear coefficients (thread 0; identical in every block) ---------------------
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    // ---- p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = M^-1 r ----------------------
    for (int s = gwarp; s < p.nslices; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      if (row >= 0) {
        const double4 u = ld256(p.U + row), w = ld256(p.W + row);
        double4 pp = ld256(p.P + row), ss = ld256(p.S + row);
        pp.x = u.x + b0 * pp.x; pp.y = u.y + b1 * pp.y; pp.z = u.z + b2 * pp.z;
        ss.x = w.x + b0 * ss.x; ss.y = w.y + b1 * ss.y; ss.z = w.z + b2 * ss.z;
        st256(p.P + row, pp); st256(p.S + row, ss);
        double4 x = ld256(p.X + row), r = ld256(p.R + row);
        const double di = p.dinv[row];
        x.x += a0 * pp.x; x.y += a1 * pp.y; x.z += a2 * pp.z;
        r.x -= a0 * ss.x; r.y -= a1 * ss.y; r.z -= a2 * ss.z;
        st256(p.X + row, x); st256(p.R + row, r);
        if (!has_pairs) st256(p.U + row, make_double4(di * r.x, di * r.y, di * r.z, 0.0));
      }
    }
    ++it;
    grid.sync();                                        // new u (or new r) visible to the whole grid
    if (has_pairs) {
      // u = D^-1 r + pair correction; the mate's r was written by another thread before the barrier
      for (int s = gwarp; s < p.nslices; s += nwarps) {
        const int row = p.sell_row[s * kSellC + lane];
        if (row >= 0) {
          const double4 r = ld256(p.R + row);
          const double di = p.dinv[row];
          double4 u = make_double4(di * r.x, di * r.y, di * r.z, 0.0);
          const int mt = p.mate[row];
          if (mt >= 0) {
            const double4 rm = ld256(p.R + mt);
            const double c2 = p.pc2[row];
            u.x += c2 * rm.x; u.y += c2 * rm.y; u.z += c2 * rm.z;
            const int m2 = p.mate2[row];
            if (m2 >= 0) {
              const double4 r2 = ld256(p.R + m2);
              const double c3 = p.pc3[row];
              u.x += c3 * r2.x; u.y += c3 * r2.y; u.z += c3 * r2.z;
            }
          }
          st256(p.U + row, u);
        }
      }
      grid.sync();
    }
    if (timer) { const long long c = clock64(); c_upd += c - c_mark; c_mark = c; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
    unsigned long long ns_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
    p.ctl->cyc_spmv += c_spmv;
    p.ctl->cyc_update += c_upd;
    p.ctl->cyc_total += clock64() - c_begin;
    p.ctl->ns_total += (long long)(ns_end - ns_begin);
    p.ctl->pcg_spmv_phases += c_spmv > 0 ? it + 1 : 0;
  }
}

// ---- register-resident variant -----------------------------------------------------------------------
// When the graph has no more slices than the grid has warps (n <= 148 x 24 x 32 = 113 664 rows, i.e.
// config 3), every lane owns exactly one row for the whole solve and keeps its x, r, p, s (and the
// preconditioner coefficients) in registers: per iteration only u is written (32 B / row) and gathered;
// the 350 B / row of vector streams of the general kernel disappear, and so does its third barrier:
// a paired row recomputes its mate's new residual from the mate's (r, s) of the PREVIOUS iteration
// (ping-pong buffers RS[it & 1], written only by paired rows) and the mate's w of this iteration, all
// of which were published before the reduction barrier.
struct PcgRegParams {
  PcgParams base;
  double4* RS0r; double4* RS0s; double4* RS1r; double4* RS1s;   // ping-pong (r, s) of paired rows
  // slice of (block, warp), -1 = none: slices dealt to blocks so that every SM gathers the same number of entries
  // (null: round-robin, slice = block + grid * warp).  Degree-sorted 1024-row windows make slice widths periodic
  // (period 32 slices); round-robin over 148 blocks aliases with that period and leaves the heaviest block 14 %
  // above the mean - and every grid barrier waits for the heaviest block.
  const int* slice_map;
};

template <int V, int UNR>
__global__ void __launch_bounds__(kPcgThreads, 1)
k_pcg_persistent_reg(const PcgRegParams q) {
  const PcgParams& p = q.base;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31;
  const int slice = q.slice_map ? q.slice_map[blockIdx.x * (kPcgThreads / 32) + (threadIdx.x >> 5)]
                                : (int)(blockIdx.x + gridDim.x * (threadIdx.x >> 5));     // <= 1 slice per warp
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  int row = -1, width = 0, mt = -1, mt2 = -1;
  int64_t base = 0;
  if (slice >= 0 && slice < p.nslices) {
    row = p.sell_row[slice * kSellC + lane];
    width = p.slice_width[slice];
    base = (int64_t)p.slice_off[slice] + lane;
  }
  double v[kPcgNV];
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, s0 = 0, s1 = 0, s2 = 0;
  double u0 = 0, u1 = 0, u2 = 0, di = 0, c2 = 0, c3 = 0;
  if (row >= 0) {
    const double4 b = ldg256(p.B + row);
    const double d = p.diag[row];
    di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
    r0 = b.x; r1 = b.y; r2 = b.z;
    u0 = di * r0; u1 = di * r1; u2 = di * r2;
    if (has_pairs) {
      mt = p.mate[row];
      if (mt >= 0) {
        c2 = p.pc2[row];
        const double4 bm = ldg256(p.B + mt);
        u0 += c2 * bm.x; u1 += c2 * bm.y; u2 += c2 * bm.z;
        mt2 = p.mate2[row];
        if (mt2 >= 0) {
          c3 = p.pc3[row];
          const double4 b2 = ldg256(p.B + mt2);
          u0 += c3 * b2.x; u1 += c3 * b2.y; u2 += c3 * b2.z;
        }
        st256(q.RS1r + row, b);                                      // "previous" buffer of iteration 0
        st256(q.RS1s + row, make_double4(0, 0, 0, 0));
      }
    }
    st256(p.U + row, make_double4(u0, u1, u2, 0.0));
    v[0] = r0 * r0; v[1] = r1 * r1; v[2] = r2 * r2;
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  int it = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  while (!sc_stop) {
    double w0 = 0, w1 = 0, w2 = 0;
    if (row >= 0 || width > 0) {
      sell_row_apply<V, UNR>(p.sell_col, p.sell_w2, p.U, base, width, make_double4(u0, u1, u2, 0.0), w0, w1, w2);
    }
    if (row >= 0) {
      if (mt >= 0) st256(p.W + row, make_double4(w0, w1, w2, 0.0));   // the mate needs it after the barrier
      v[0] = r0 * u0; v[1] = r1 * u1; v[2] = r2 * u2;
      v[3] = u0 * w0; v[4] = u1 * w1; v[5] = u2 * w2;
      v[6] = r0 * r0; v[7] = r1 * r1; v[8] = r2 * r2;
    } else {
#pragma unroll
      for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    if (timer) { const long long c = clock64(); c_spmv += c - c_mark; c_mark = c; }
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    if (row >= 0) {
      p0 = u0 + b0 * p0; p1 = u1 + b1 * p1; p2 = u2 + b2 * p2;
      s0 = w0 + b0 * s0; s1 = w1 + b1 * s1; s2 = w2 + b2 * s2;
      x0 += a0 * p0; x1 += a1 * p1; x2 += a2 * p2;
      r0 -= a0 * s0; r1 -= a1 * s1; r2 -= a2 * s2;
      u0 = di * r0; u1 = di * r1; u2 = di * r2;
      if (mt >= 0) {
        double4* const curR = (it & 1) ? q.RS1r : q.RS0r;
        double4* const curS = (it & 1) ? q.RS1s : q.RS0s;
        const double4* const oldR = (it & 1) ? q.RS0r : q.RS1r;
        const double4* const oldS = (it & 1) ? q.RS0s : q.RS1s;
        const double4 rm = ld256(oldR + mt), sm = ld256(oldS + mt), wm = ld256(p.W + mt);
        const double sm0 = wm.x + b0 * sm.x, sm1 = wm.y + b1 * sm.y, sm2 = wm.z + b2 * sm.z;   // the mate's new s
        u0 += c2 * (rm.x - a0 * sm0); u1 += c2 * (rm.y - a1 * sm1); u2 += c2 * (rm.z - a2 * sm2);
        if (mt2 >= 0) {
          const double4 rn = ld256(oldR + mt2), sn = ld256(oldS + mt2), wn = ld256(p.W + mt2);
          const double t0 = wn.x + b0 * sn.x, t1 = wn.y + b1 * sn.y, t2 = wn.z + b2 * sn.z;
          u0 += c3 * (rn.x - a0 * t0); u1 += c3 * (rn.y - a1 * t1); u2 += c3 * (rn.z - a2 * t2);
        }
        st256(curR + row, make_double4(r0, r1, r2, 0.0));
        st256(curS + row, make_double4(s0, s1, s2, 0.0));
      }
      st256(p.U + row, make_double4(u0, u1, u2, 0.0));
    }
    ++it;
    grid.sync();
    if (timer) { const long long c = clock64(); c_upd += c - c_mark; c_mark = c; }
  }
  if (row >= 0) st256(p.X + row, make_double4(x0, x1, x2, 0.0));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
    unsigned long long ns_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
    p.ctl->cyc_spmv += c_spmv;
    p.ctl->cyc_update += c_upd;
    p.ctl->cyc_total += clock64() - c_begin;
    p.ctl->ns_total += (long long)(ns_end - ns_begin);
    p.ctl->pcg_spmv_phases += c_spmv > 0 ? it + 1 : 0;
  }
}

// ---- small graphs: several warps per slice ---------------------------------------------------------------
// With one slice per warp the SpMV of a PCG iteration is a latency chain - width / 4 batches of 4 gathers, one
// after the other (6 batches at degree ~22: ~5 us) - however few slices there are, and a graph of a few thousand
// nodes (config 2: 142 slices; the global rotAvg calls of config 5) leaves 95 % of the warps idle.  Here
// kMwWarps warps share a slice: warp `sub` takes batches sub, sub + kMwWarps, ... of every row, the partial sums
// meet in shared memory (added in warp order: deterministic) and warp 0 of the group - the only one that keeps
// the row's x, r, p, s in registers - does everything else exactly as k_pcg_persistent_reg.  Fewer blocks also
// make the two grid barriers cheaper.
constexpr int kMwWarps = 8;
constexpr int kMwGroups = kPcgThreads / 32 / kMwWarps;     // slices per block

__global__ void __launch_bounds__(kPcgThreads, 1)
k_pcg_persistent_reg_mw(const PcgRegParams q) {
  const PcgParams& p = q.base;
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  __shared__ double part[kMwGroups][kMwWarps][3][32];
  const int lane = threadIdx.x & 31;
  const int group = (threadIdx.x >> 5) / kMwWarps, sub = (threadIdx.x >> 5) % kMwWarps;
  const int slice = blockIdx.x * kMwGroups + group;                  // one slice per group of warps
  const bool leader = sub == 0;
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  int row = -1, width = 0, mt = -1, mt2 = -1;
  int64_t base = 0;
  if (slice < p.nslices) {
    row = p.sell_row[slice * kSellC + lane];
    width = p.slice_width[slice];
    base = (int64_t)p.slice_off[slice] + lane;
  }
  double v[kPcgNV];
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, s0 = 0, s1 = 0, s2 = 0;
  double u0 = 0, u1 = 0, u2 = 0, di = 0, c2 = 0, c3 = 0;
  const int row_all = row;                                            // every warp of the group knows the row ...
  if (!leader) row = -1;                                              // ... but only the leader owns its state
  if (row >= 0) {
    const double4 b = ldg256(p.B + row);
    const double d = p.diag[row];
    di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
    r0 = b.x; r1 = b.y; r2 = b.z;
    u0 = di * r0; u1 = di * r1; u2 = di * r2;
    if (has_pairs) {
      mt = p.mate[row];
      if (mt >= 0) {
        c2 = p.pc2[row];
        const double4 bm = ldg256(p.B + mt);
        u0 += c2 * bm.x; u1 += c2 * bm.y; u2 += c2 * bm.z;
        mt2 = p.mate2[row];
        if (mt2 >= 0) {
          c3 = p.pc3[row];
          const double4 b2 = ldg256(p.B + mt2);
          u0 += c3 * b2.x; u1 += c3 * b2.y; u2 += c3 * b2.z;
        }
        st256(q.RS1r + row, b);                                      // "previous" buffer of iteration 0
        st256(q.RS1s + row, make_double4(0, 0, 0, 0));
      }
    }
    st256(p.U + row, make_double4(u0, u1, u2, 0.0));
    v[0] = r0 * r0; v[1] = r1 * r1; v[2] = r2 * r2;
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  int it = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  long long dbg[5] = {0, 0, 0, 0, 0}, dm = 0;
  while (!sc_stop) {
    if (timer) dm = clock64();
    double w0 = 0, w1 = 0, w2 = 0;
    if (slice < p.nslices) {
      // this warp's batches of the row; the row's own u was published by the leader before the last barrier
      const double4 uo = row_all >= 0 ? ld256(p.U + row_all) : make_double4(0, 0, 0, 0);
      double ax = 0, ay = 0, az = 0;
      for (int j = sub * 4; j < width; j += 4 * kMwWarps) {
        int c[4]; double ww[4]; double4 uc[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int64_t o = base + (int64_t)(j + t) * kSellC;
          c[t] = __ldg(p.sell_col + o);
          ww[t] = __ldg(p.sell_w2 + o);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) uc[t] = ld256(p.U + c[t]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { ax += ww[t] * (uo.x - uc[t].x); ay += ww[t] * (uo.y - uc[t].y); az += ww[t] * (uo.z - uc[t].z); }
      }
      part[group][sub][0][lane] = ax; part[group][sub][1][lane] = ay; part[group][sub][2][lane] = az;
    }
    __syncthreads();
    if (timer) { const long long c = clock64(); dbg[0] += c - dm; dm = c; }
    if (leader && slice < p.nslices) {
#pragma unroll
      for (int t = 0; t < kMwWarps; ++t) { w0 += part[group][t][0][lane]; w1 += part[group][t][1][lane]; w2 += part[group][t][2][lane]; }
    }
    if (row >= 0) {
      if (mt >= 0) st256(p.W + row, make_double4(w0, w1, w2, 0.0));   // the mate needs it after the barrier
      v[0] = r0 * u0; v[1] = r1 * u1; v[2] = r2 * u2;
      v[3] = u0 * w0; v[4] = u1 * w1; v[5] = u2 * w2;
      v[6] = r0 * r0; v[7] = r1 * r1; v[8] = r2 * r2;
    } else {
#pragma unroll
      for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    if (timer) { const long long c = clock64(); c_spmv += c - c_mark; c_mark = c; dbg[1] += c - dm; dm = c; }
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (timer) { const long long c = clock64(); dbg[2] += c - dm; dm = c; }
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    if (row >= 0) {
      p0 = u0 + b0 * p0; p1 = u1 + b1 * p1; p2 = u2 + b2 * p2;
      s0 = w0 + b0 * s0; s1 = w1 + b1 * s1; s2 = w2 + b2 * s2;
      x0 += a0 * p0; x1 += a1 * p1; x2 += a2 * p2;
      r0 -= a0 * s0; r1 -= a1 * s1; r2 -= a2 * s2;
      u0 = di * r0; u1 = di * r1; u2 = di * r2;
      if (mt >= 0) {
        double4* const curR = (it & 1) ? q.RS1r : q.RS0r;
        double4* const curS = (it & 1) ? q.RS1s : q.RS0s;
        const double4* const oldR = (it & 1) ? q.RS0r : q.RS1r;
        const double4* const oldS = (it & 1) ? q.RS0s : q.RS1s;
        const double4 rm = ld256(oldR + mt), sm = ld256(oldS + mt), wm = ld256(p.W + mt);
        const double sm0 = wm.x + b0 * sm.x, sm1 = wm.y + b1 * sm.y, sm2 = wm.z + b2 * sm.z;   // the mate's new s
        u0 += c2 * (rm.x - a0 * sm0); u1 += c2 * (rm.y - a1 * sm1); u2 += c2 * (rm.z - a2 * sm2);
        if (mt2 >= 0) {
          const double4 rn = ld256(oldR + mt2), sn = ld256(oldS + mt2), wn = ld256(p.W + mt2);
          const double t0 = wn.x + b0 * sn.x, t1 = wn.y + b1 * sn.y, t2 = wn.z + b2 * sn.z;
          u0 += c3 * (rn.x - a0 * t0); u1 += c3 * (rn.y - a1 * t1); u2 += c3 * (rn.z - a2 * t2);
        }
        st256(curR + row, make_double4(r0, r1, r2, 0.0));
        st256(curS + row, make_double4(s0, s1, s2, 0.0));
      }
      st256(p.U + row, make_double4(u0, u1, u2, 0.0));
    }
    ++it;
    if (timer) { const long long c = clock64(); dbg[3] += c - dm; dm = c; }
    grid.sync();
    if (timer) { const long long c = clock64(); c_upd += c - c_mark; c_mark = c; dbg[4] += c - dm; }
  }
  if (row >= 0) st256(p.X + row, make_double4(x0, x1, x2, 0.0));
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
    unsigned long long ns_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
    p.ctl->cyc_spmv += c_spmv;
    p.ctl->cyc_update += c_upd;
    p.ctl->cyc_total += clock64() - c_begin;
    p.ctl->ns_total += (long long)(ns_end - ns_begin);
    p.ctl->pcg_spmv_phases += c_spmv > 0 ? it + 1 : 0;
    if (p.debug)
      printf("[k_pcg_persistent_reg_mw] %d blocks, iters %d: cycles/iter partial SpMV+sync %lld, reduce (grid barrier) %lld, "
             "coefficients %lld, update %lld, grid barrier %lld\n", (int)gridDim.x, it, dbg[0] / (it + 1), dbg[1] / (it + 1),
             dbg[2] / (it + 1), dbg[3] / (it > 0 ? it : 1), dbg[4] / (it > 0 ? it : 1));
  }
}

}  // namespace ira
