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
// Chains laid out in edge order are the practical case (a SLAM sequence; the synthetic configs list their path
// first: ONE sweep flags 100 000 nodes one after the other), so the relaxation is organised for them: every warp
// owns a SEGMENT of kMstSeg consecutive edges per pass and walks it in groups of 32; the 32 edges' endpoint labels
// are loaded at once and the group is relaxed to its local fixed point IN REGISTERS (the lowest lane that can
// still improve a label applies its update, every lane patches its own copy by shuffle) - a hop costs ~50 cycles
// instead of an L2 round trip, and a chain advances a whole segment per pass (config 3: 25 passes, ~3 ms; the
// first version needed 3 126 passes of 30 us).
//
// With the labels final, node v's parent edge is (label & 0xffffffff) - 1 and
//     Q_v = QQ_k (x) Q_u          when v is the edge's second endpoint (:941)
//     Q_v = [QQ_k.xyz, -QQ_k.w] (x) Q_u   when v is the first (:955-958, the reference's sign)
// (rows < f_init keep their value, :939,953, but still propagate).  The reference evaluates these products one
// after the other down the tree - 100 000 dependent products on config 3.  Here the tree is contracted by POINTER
// JUMPING: every node holds (ancestor a, T) with Q_v = T (x) Q_a; a round replaces (a, T) by (a's ancestor,
// T (x) T_a); after ceil(log2(depth)) rounds every ancestor is a root (node 0, a given rotation, or an unreached
// node) and one product finishes the node.  Same parents, same factors, a different association of the same
// product: differences are rounding only (<= 1e-13 on the 100 000-hop chain of config 3, tested).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ira_kernels.cuh"

namespace ira {
namespace cg = cooperative_groups;

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

constexpr int kMstSeg = 4096;          // consecutive edges one warp relaxes per pass (multiple of 32)

// can edge (tu, tv) with 1-based index k1 still improve one of its endpoint labels?
__device__ __forceinline__ bool mst_can(unsigned long long tu, unsigned long long tv, unsigned long long k1) {
  return (tu != kMstInf && mst_next(tu, k1) < tv) || (tv != kMstInf && mst_next(tv, k1) < tu);
}

__global__ void __launch_bounds__(256)
k_mst_labels(const int2* __restrict__ I, int64_t m, unsigned long long* label, MstCtl* ctl) {
  cg::grid_group grid = cg::this_grid();
  const int lane = threadIdx.x & 31;
  const int64_t gwarp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t nseg = (m + kMstSeg - 1) / kMstSeg;
  int pass = 0;
  for (;; ++pass) {
    const int slot = pass % 3;
    if (gwarp == 0 && lane == 0) ctl->changed[(pass + 1) % 3] = 0;
    bool moved = false;
    for (int64_t sg = gwarp; sg < nseg; sg += nwarps) {
      const int64_t s0 = sg * kMstSeg, s1 = min(m, s0 + (int64_t)kMstSeg);
      int2 e_next = make_int2(0, 0);                           // edge endpoints are prefetched one group ahead
      if (s0 + lane < s1) e_next = __ldg(I + s0 + lane);
      for (int64_t k0 = s0; k0 < s1; k0 += 32) {
        const int64_t k = k0 + lane;
        const int2 e = k < s1 ? e_next : make_int2(0, 0);       // lanes past the segment end hold no edge
        if (k + 32 < s1) e_next = __ldg(I + k + 32);
        unsigned long long tu = kMstInf, tv = kMstInf;
        if (k < s1 && e.x != e.y) { tu = __ldcg(label + e.x); tv = __ldcg(label + e.y); }
        const unsigned long long k1 = (unsigned long long)k + 1ull;
        // Fast path for chains laid out in edge order (edge j starts where edge j-1 ended, indices increasing): the
        // time stamps of a whole run follow from its head in log2(32) steps.  next(next(t, k), k') = (sweep of
        // next(t, k), k') for k' > k, so only the SWEEP travels down the run: a segmented min-scan.
        {
          const int ey_left = __shfl_up_sync(0xffffffffu, e.y, 1);
          const bool valid = k < s1 && e.x != e.y;
          bool head = !(lane > 0 && valid && e.x == ey_left);
          unsigned int sw = (valid && tu != kMstInf) ? (unsigned int)(mst_next(tu, k1) >> 32) : 0xffffffffu;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const unsigned int swl = __shfl_up_sync(0xffffffffu, sw, d);
            const bool hl = __shfl_up_sync(0xffffffffu, (int)head, d) != 0;
            if (lane >= d && !head) { sw = min(sw, swl); head = hl; }
          }
          if (valid && sw != 0xffffffffu) {
            const unsigned long long cand = ((unsigned long long)sw << 32) | k1;
            if (cand < tv) { atomicMin(label + e.y, cand); tv = cand; moved = true; }
          }
          const unsigned long long tvl = __shfl_up_sync(0xffffffffu, tv, 1);     // the left neighbour's end = my start
          if (lane > 0 && valid && e.x == ey_left && tvl < tu) tu = tvl;
        }
        // whatever is left (backward edges, nodes shared otherwise): relax the group to its local fixed point, the
        // lowest lane that can improve a label goes first
        unsigned int can = __ballot_sync(0xffffffffu, mst_can(tu, tv, k1));
        while (can) {
          const int j = __ffs(can) - 1;
          const int ex = __shfl_sync(0xffffffffu, e.x, j), ey = __shfl_sync(0xffffffffu, e.y, j);
          unsigned long long tuj = __shfl_sync(0xffffffffu, tu, j), tvj = __shfl_sync(0xffffffffu, tv, j);
          const unsigned long long kj = (unsigned long long)(k0 + j) + 1ull;
          if (tuj != kMstInf) {                                 // flags[e1] && !flags[e2]  (:934)
            const unsigned long long cand = mst_next(tuj, kj);
            if (cand < tvj) {
              tvj = cand;
              if (lane == j) atomicMin(label + ey, cand);
              if (e.x == ey && cand < tu) tu = cand;
              if (e.y == ey && cand < tv) tv = cand;
            }
          }
          if (tvj != kMstInf) {                                 // !flags[e1] && flags[e2]  (:950)
            const unsigned long long cand = mst_next(tvj, kj);
            if (cand < tuj) {
              if (lane == j) atomicMin(label + ex, cand);
              if (e.x == ex && cand < tu) tu = cand;
              if (e.y == ex && cand < tv) tv = cand;
            }
          }
          moved = true;
          can = __ballot_sync(0xffffffffu, mst_can(tu, tv, k1));
        }
      }
    }
    if (__any_sync(0xffffffffu, moved) && lane == 0) atomicOr(&ctl->changed[slot], 1);
    grid.sync();
    if (!__ldcg(&ctl->changed[slot])) break;
  }
  if (gwarp == 0 && lane == 0) ctl->passes_label = pass + 1;
}

// Parent and factor of every node from its final label: roots (node 0, given rotations, unreached nodes) point to
// themselves with the identity.
__global__ void __launch_bounds__(256)
k_mst_parents(const int2* __restrict__ I, const double* __restrict__ QQ, int64_t ldqq,
              const unsigned long long* __restrict__ label, int n, int f_init, int* __restrict__ anc,
              double4* __restrict__ T, MstCtl* ctl) {
  int unreached = 0;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < n; v += gridDim.x * blockDim.x) {
    const unsigned long long lab = label[v];
    int a = v;
    double4 t = make_double4(0.0, 0.0, 0.0, 1.0);
    if (lab == kMstInf) ++unreached;
    else if (v >= f_init && lab != 0ull) {                   // do not change known rotations (:939,953)
      const int64_t k = (int64_t)(lab & 0xffffffffull) - 1;
      const int2 e = __ldg(I + k);
      a = e.x == v ? e.y : e.x;
      t = make_double4(QQ[k], QQ[ldqq + k], QQ[2 * ldqq + k], QQ[3 * ldqq + k]);
      if (e.x == v) t.w = -t.w;                              // QQj_inv(3) *= -1  (:956-957)
    }
    anc[v] = a;
    st256(T + v, t);
  }
  if (unreached) atomicAdd(&ctl->unreached, unreached);
}

// Pointer jumping, double-buffered: (anc, T) <- (anc[anc], T (x) T[anc]) until every ancestor is a root, then
// Q_v = T_v (x) Q_root.  Roots are the nodes with anc == self; they are never written.
__global__ void __launch_bounds__(256)
k_mst_jump(int* anc0, double4* T0, int* anc1, double4* T1, int n, double4* Q, MstCtl* ctl) {
  cg::grid_group grid = cg::this_grid();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  int* a_in = anc0; int* a_out = anc1;
  double4* t_in = T0; double4* t_out = T1;
  int round = 0;
  for (;; ++round) {
    const int slot = round % 3;
    if (gtid == 0) ctl->changed[(round + 1) % 3] = 0;
    bool again = false;
    for (int v = gtid; v < n; v += nthreads) {
      const int a = __ldcg(a_in + v);
      double4 t = ldcg256(t_in + v);
      int a2 = a;
      if (a != v) {
        const int aa = __ldcg(a_in + a);
        if (aa != a) {                                       // the ancestor is not a root yet: jump over it
          t = quat_mult(t, ldcg256(t_in + a));
          a2 = aa;
          if (__ldcg(a_in + aa) != aa) again = true;
        }
      }
      a_out[v] = a2;
      st256(t_out + v, t);
    }
    if (__any_sync(0xffffffffu, again) && (threadIdx.x & 31) == 0) atomicOr(&ctl->changed[slot], 1);
    grid.sync();
    int* ta = a_in; a_in = a_out; a_out = ta;
    double4* tt = t_in; t_in = t_out; t_out = tt;
    if (!__ldcg(&ctl->changed[slot])) break;
  }
  for (int v = gtid; v < n; v += nthreads) {
    const int a = __ldcg(a_in + v);
    if (a != v) st256(Q + v, quat_mult(ldcg256(t_in + v), ldcg256(Q + a)));
  }
  if (gtid == 0) ctl->passes_prop = round + 1;
}

}  // namespace ira
