// irotavg::init_mst (ral/l1_irls.cpp:915-979) on the device.
//
// The reference sweeps the edge list in order, again and again, and the first edge that meets an
// unflagged node from a flagged one assigns that node's rotation: the spanning tree it builds (and
// therefore the start every later stage sees) depends on the edge ORDER.  To reproduce exactly that
// tree in parallel, every node gets the time stamp of the moment the sequential sweeps would flag it:
//
//     label(v) = (sweep << 32) | (edge index + 1),   label(0) = 0   (flags[0] = true, :922)
//
// Edge k can flag v from its other endpoint u at the first time (s, k) later than label(u): the same
// sweep when u was flagged by an earlier edge of the list, the next sweep otherwise.  label(v) is the
// minimum of that over v's edges - a monotone shortest-path problem, solved by label-correcting
// relaxation (atomicMin) to its fixed point, which is unique and equals the sequential result.
// Each thread relaxes a chunk of consecutive edges in list order, so a chain laid out in edge order
// advances a whole chunk per pass instead of one hop.
//
// With the labels final, node v's parent edge is (label & 0xffffffff) - 1 and
//     Q_v = QQ_k (x) Q_u          when v is the edge's second endpoint (:941)
//     Q_v = [QQ_k.xyz, -QQ_k.w] (x) Q_u   when v is the first (:955-958, the reference's sign)
// each computed once from the parent's final value, exactly as in the reference (rows < f_init keep
// their value, :939,953, but still propagate).  Nodes are visited in label order (radix sort), a chunk
// of consecutive positions per thread, in passes until no node waits for its parent.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ira_kernels.cuh"

namespace ira {
namespace cg = cooperative_groups;

constexpr int kMstEdgeChunk = 32;
constexpr int kMstNodeChunk = 16;
constexpr unsigned long long kMstInf = ~0ull;

struct MstCtl {
  int changed[3];     // per-pass "something moved" flags, rotated so that a slot is reset two passes ahead
  int passes_label;
  int passes_prop;
  int unreached;      // nodes the relative rotations do not span (:970-977)
};

__device__ __forceinline__ double4 ldcg256(const double4* p) {          // L2-coherent (skips L1)
  double4 v;
  asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void k_mst_init(unsigned long long* __restrict__ label, int* __restrict__ done, int* __restrict__ order,
                           int n, MstCtl* ctl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    label[i] = i == 0 ? 0ull : kMstInf;
    done[i] = i == 0 ? 1 : 0;
    order[i] = i;
  }
  if (i == 0) { ctl->changed[0] = ctl->changed[1] = ctl->changed[2] = 0; ctl->passes_label = ctl->passes_prop = 0; ctl->unreached = 0; }
}

// first time edge k fires after time stamp t
__device__ __forceinline__ unsigned long long mst_next(unsigned long long t, unsigned long long k1) {
  const unsigned long long sweep = (t >> 32) + ((t & 0xffffffffull) >= k1 ? 1ull : 0ull);
  return (sweep << 32) | k1;
}

__global__ void __launch_bounds__(256)
k_mst_labels(const int2* __restrict__ I, int64_t m, unsigned long long* label, MstCtl* ctl) {
  cg::grid_group grid = cg::this_grid();
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t nchunks = (m + kMstEdgeChunk - 1) / kMstEdgeChunk;
  int pass = 0;
  for (;; ++pass) {
    const int slot = pass % 3;
    if (gtid == 0) ctl->changed[(pass + 1) % 3] = 0;
    bool moved = false;
    for (int64_t c = gtid; c < nchunks; c += nthreads) {
      const int64_t k0 = c * kMstEdgeChunk, k1 = min(m, k0 + (int64_t)kMstEdgeChunk);
      for (int64_t k = k0; k < k1; ++k) {
        const int2 e = __ldg(I + k);
        if (e.x == e.y) continue;
        const unsigned long long tu = __ldcg(label + e.x);
        unsigned long long tv = __ldcg(label + e.y);
        if (tu != kMstInf) {                                  // flags[e1] && !flags[e2]  (:934)
          const unsigned long long cand = mst_next(tu, (unsigned long long)k + 1ull);
          if (cand < tv) { atomicMin(label + e.y, cand); tv = cand; moved = true; }
        }
        if (tv != kMstInf) {                                  // !flags[e1] && flags[e2]  (:950)
          const unsigned long long cand = mst_next(tv, (unsigned long long)k + 1ull);
          if (cand < tu) { atomicMin(label + e.x, cand); moved = true; }
        }
      }
    }
    if (__any_sync(0xffffffffu, moved) && (threadIdx.x & 31) == 0) atomicOr(&ctl->changed[slot], 1);
    grid.sync();
    if (!__ldcg(&ctl->changed[slot])) break;
  }
  if (gtid == 0) ctl->passes_label = pass + 1;
}

__global__ void __launch_bounds__(256)
k_mst_propagate(const int2* __restrict__ I, const double* __restrict__ QQ, int64_t ldqq,
                const int* __restrict__ order, const unsigned long long* __restrict__ label, int n, int f_init,
                double4* Q, int* done, MstCtl* ctl) {
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  const int nchunks = (n + kMstNodeChunk - 1) / kMstNodeChunk;
  int unreached = 0;
  int pass = 0;
  for (;; ++pass) {
    const int slot = pass % 3;
    if (gtid == 0) ctl->changed[(pass + 1) % 3] = 0;
    bool waiting = false;
    for (int c = gtid; c < nchunks; c += nthreads) {
      const int p0 = c * kMstNodeChunk, p1 = min(n, p0 + kMstNodeChunk);
      for (int pos = p0; pos < p1; ++pos) {
        const int v = order[pos];
        if (__ldcg(done + v)) continue;
        const unsigned long long lab = label[v];
        if (lab == kMstInf) { if (pass == 0) ++unreached; continue; }
        const int64_t k = (int64_t)(lab & 0xffffffffull) - 1;
        const int2 e = __ldg(I + k);
        const int u = e.x == v ? e.y : e.x;
        if (!__ldcg(done + u)) { waiting = true; continue; }
        if (v >= f_init) {                                   // do not change known rotations (:939,953)
          __threadfence();                                   // the parent's Q was published before its flag
          const double4 qu = ldcg256(Q + u);
          double4 qq = make_double4(QQ[k], QQ[ldqq + k], QQ[2 * ldqq + k], QQ[3 * ldqq + k]);
          if (e.x == v) qq.w = -qq.w;                        // QQj_inv(3) *= -1  (:956-957)
          st256(Q + v, quat_mult(qq, qu));
          __threadfence();
        }
        atomicExch(done + v, 1);
      }
    }
    if (__any_sync(0xffffffffu, waiting) && (threadIdx.x & 31) == 0) atomicOr(&ctl->changed[slot], 1);
    grid.sync();
    if (!__ldcg(&ctl->changed[slot])) break;
  }
  if (unreached) atomicAdd(&ctl->unreached, unreached);
  if (gtid == 0) ctl->passes_prop = pass + 1;
}

}  // namespace ira
