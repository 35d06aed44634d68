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
constexpr int kMstNodeChunk = 32;
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
      const int64_t k0 = c * kMstEdgeChunk;
      const int cnt = (int)min((int64_t)kMstEdgeChunk, m - k0);
      // the chunk's edges and the labels of their endpoints are fetched up front (independent loads, in flight
      // together); the sequential sweep below then only forwards what it changed itself
      int2 e[kMstEdgeChunk];
      unsigned long long lu[kMstEdgeChunk], lv[kMstEdgeChunk];
#pragma unroll 8
      for (int q = 0; q < kMstEdgeChunk; ++q) e[q] = q < cnt ? __ldg(I + k0 + q) : make_int2(0, 0);
#pragma unroll 8
      for (int q = 0; q < kMstEdgeChunk; ++q) { lu[q] = __ldcg(label + e[q].x); lv[q] = __ldcg(label + e[q].y); }
      int pn1 = -1, pn2 = -1;                                   // endpoints of the previous edge and what this
      unsigned long long pl1 = 0, pl2 = 0;                      // thread left in their labels: chains move on in registers
      for (int q = 0; q < cnt; ++q) {
        const int2 ed = e[q];
        if (ed.x == ed.y) continue;
        unsigned long long tu = lu[q], tv = lv[q];
        if (ed.x == pn1) tu = min(tu, pl1); else if (ed.x == pn2) tu = min(tu, pl2);
        if (ed.y == pn1) tv = min(tv, pl1); else if (ed.y == pn2) tv = min(tv, pl2);
        const unsigned long long k1 = (unsigned long long)(k0 + q) + 1ull;
        if (tu != kMstInf) {                                  // flags[e1] && !flags[e2]  (:934)
          const unsigned long long cand = mst_next(tu, k1);
          if (cand < tv) { atomicMin(label + ed.y, cand); tv = cand; moved = true; }
        }
        if (tv != kMstInf) {                                  // !flags[e1] && flags[e2]  (:950)
          const unsigned long long cand = mst_next(tv, k1);
          if (cand < tu) { atomicMin(label + ed.x, cand); tu = cand; moved = true; }
        }
        pn1 = ed.x; pl1 = tu; pn2 = ed.y; pl2 = tv;
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
      const int p0 = c * kMstNodeChunk;
      const int cnt = min(kMstNodeChunk, n - p0);
      // phase 1: everything that does not depend on other nodes' progress, for the whole chunk at once
      int vv[kMstNodeChunk], uu[kMstNodeChunk];              // node, parent (-1: nothing to do)
      double4 qq[kMstNodeChunk];
#pragma unroll 4
      for (int q = 0; q < kMstNodeChunk; ++q) {
        vv[q] = q < cnt ? order[p0 + q] : -1;
        uu[q] = -1;
      }
#pragma unroll 4
      for (int q = 0; q < kMstNodeChunk; ++q) {
        const int v = vv[q];
        if (v < 0 || __ldcg(done + v)) { vv[q] = -1; continue; }
        const unsigned long long lab = label[v];
        if (lab == kMstInf) { if (pass == 0) ++unreached; vv[q] = -1; continue; }
        const int64_t k = (int64_t)(lab & 0xffffffffull) - 1;
        const int2 e = __ldg(I + k);
        uu[q] = e.x == v ? e.y : e.x;
        qq[q] = make_double4(QQ[k], QQ[ldqq + k], QQ[2 * ldqq + k], QQ[3 * ldqq + k]);
        if (e.x == v) qq[q].w = -qq[q].w;                    // QQj_inv(3) *= -1  (:956-957)
      }
      // phase 2: in label order; a parent this thread has just finished is forwarded in registers, so a chain
      // that is contiguous in label order advances a whole chunk per pass
      int last = -1;
      double4 lastQ = make_double4(0, 0, 0, 1);
      unsigned int finished = 0u;
      for (int q = 0; q < cnt; ++q) {
        const int v = vv[q];
        if (v < 0) continue;
        const int u = uu[q];
        double4 qu;
        if (u == last) qu = lastQ;
        else if (__ldcg(done + u)) { __threadfence(); qu = ldcg256(Q + u); }
        else { waiting = true; continue; }
        double4 qv;
        if (v >= f_init) { qv = quat_mult(qq[q], qu); st256(Q + v, qv); }   // one product per node (:941 / :958)
        else qv = ldcg256(Q + v);                            // known rotations are kept but still propagate (:939,953)
        last = v; lastQ = qv;
        finished |= 1u << q;
      }
      if (finished) {
        __threadfence();                                     // the rotations are published before their flags
        for (int q = 0; q < cnt; ++q)
          if ((finished >> q) & 1u) atomicExch(done + vv[q], 1);
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
