// Multi-GPU linear solve over NVLink peer memory: ONE persistent cooperative kernel per rank runs the whole
// PCG solve; the ranks exchange vector slices, dot products and barrier flags by plain loads / stores into
// each other's HBM (one process per GPU, buffers mapped with CUDA IPC).  No NCCL call inside the solve.
//
// Partition.  The rows of A^T D^2 A (SELL slices, ira_pcg.cuh) are split into `world` contiguous slice ranges;
// rank g owns the rows of its range, i.e. walks the adjacency (the edge list) of its nodes only.  Every rank
// holds the whole graph and runs the O(m) edge kernels (residual, weights, rhs) redundantly - 3 % of the time
// at one GPU - so the only data that must cross GPUs is what the PCG iteration itself produces:
//   (1) after the SpMV: 9 partial dot products per rank (gamma = r.u, delta = u.Au, |r|^2 for 3 right-hand
//       sides) -> written into every peer's slot array, summed by everybody in rank order (bitwise identical
//       on all ranks, so alpha / beta / the stopping decision agree without a broadcast);
//   (2) after the vector update: the owner writes its rows of the new u = M^-1 r into EVERY rank's copy of u
//       (an all-gather by direct P2P stores, fused into the update loop - the transfer of row i overlaps the
//       arithmetic of row i+1), and, for rows of 2x2 preconditioner blocks whose mate lives on another rank,
//       its (r, s) and w to the mate's owner only.
// Three kernels, selected by ira_options.shard_mode (all parity-tested, bitwise identical across ranks):
//   k_pcg_peer          (shard_mode 2)  two cross-GPU barriers per PCG iteration: block barrier + one system-scope
//                                       fence per block + grid barrier + an epoch number stored into every peer's
//                                       flag array + every block polls its own rank's flags;
//   k_pcg_peer_ll       (shard_mode 1)  no barrier, no flag: every exchanged datum validates itself (below);
//   k_pcg_peer_ll_reg   (shard_mode 1, rows per rank <= 148 x 24 x 32)  the same with x, r, p, s, u in registers.
// Chronopoulos-Gear recurrences and the 2x2 / 3x3 block-Jacobi arithmetic as in k_pcg_persistent (ira_pcg.cuh).
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ira_pcg.cuh"
#include "ira_mst.cuh"      // ldcg256

namespace ira {

constexpr int kPeerMax = 8;
constexpr int kPeerThreads = 512;      // 16 warps: 128 registers per thread (768 threads spilled 320 B)

// Layout of the per-rank window (one cudaMalloc, exported with cudaIpcGetMemHandle).
struct PeerWindow {
  double4* U;        // [n]  full search-direction input vector u = M^-1 r (every owner writes its rows)
  double4* X;        // [n]  solution, all-gathered at the end of the solve
  double4* MR0;      // [2][n] ping-pong (r, s) of paired rows, written by the row's owner into the MATE's owner
  double4* MS0;      //        (buffer b = pointer + b * n: no arrays in the struct, see PeerWindowLL)
  size_t stride;
  __host__ __device__ double4* MR(int b) const { return MR0 + (size_t)b * stride; }
  __host__ __device__ double4* MS(int b) const { return MS0 + (size_t)b * stride; }
  double4* MW;       // [n]  w = A u of paired rows, same routing
  double* dots;      // [2][kPeerMax][16]
  unsigned long long* flags;   // [kPeerMax]
};
// The header (flags, dot-product slots) sits at FIXED offsets in front of the vectors, so that a later upload
// with another n moves only arrays that every solve re-initialises before reading.
constexpr size_t kPeerHdr = 8192;       // flags at 0 (64 B), dot slots at 256 (2 x 8 x 32 doubles = 4 KB)
__host__ __device__ inline size_t peer_window_bytes(int n) { return kPeerHdr + (size_t)7 * n * sizeof(double4); }
__host__ __device__ inline PeerWindow peer_window_at(unsigned char* base, int n) {
  PeerWindow w;
  double4* v = reinterpret_cast<double4*>(base + kPeerHdr);
  w.U = v; w.X = v + (size_t)n; w.MR0 = v + (size_t)2 * n; w.MS0 = v + (size_t)4 * n; w.MW = v + (size_t)6 * n;
  w.stride = (size_t)n;
  w.flags = reinterpret_cast<unsigned long long*>(base);
  w.dots = reinterpret_cast<double*>(base + 256);
  return w;
}

struct PcgPeerParams {
  PcgParams base;              // X of base is unused (the window's X is the output)
  int world, rank;
  int slice_lo, slice_hi;      // this rank's slices
  const int* sell_pos;         // row -> SELL position (owner of a row = rank whose slice range holds it)
  int slice_bound[kPeerMax + 1];
  unsigned char* win[kPeerMax];   // window base of every rank in THIS process's address space (own: local)
  unsigned long long epoch_base;  // flags hold monotonically increasing epochs across launches
  int debug;                      // print block 0's phase split (spmv_variant & 7 == 7)
  int flush;                      // 1 (default) = one system fence per warp after the u stores: pushes them onto the
                                  // link at once (NVLink store completion ~7 us here; -5 us per iteration measured),
                                  // 0 = none, 2 = one per block (spmv_variant bits 3-4, experiments)
  const int* sell_colpos;         // LL variant: SELL slot -> POSITION of its column (u lives in position order there)
  int npos;                       // padded row count (vector stride of the LL window)
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// All blocks of all ranks pass this point together.  Every thread's earlier remote stores are made visible
// system-wide first (fence), the local grid barrier collects them, block 0 then raises this rank's epoch in
// every rank's flag array, and every block waits until all ranks have raised theirs.
__device__ __forceinline__ void peer_barrier(const PcgPeerParams& q, cooperative_groups::grid_group& grid,
                                             unsigned long long epoch) {
  // one system-scope fence per block: after the block barrier thread 0 has (transitively) observed every store
  // of its block, and the fence is cumulative - 148 fences per barrier instead of 113 664
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < q.world) {
    PeerWindow w = peer_window_at(q.win[threadIdx.x], q.base.n);
    // relaxed: every block's data stores were fenced system-wide (completed at the peers) before the grid barrier
    st_relaxed_sys(w.flags + q.rank, epoch);
  }
  if (threadIdx.x < q.world) {
    PeerWindow mine = peer_window_at(q.win[q.rank], q.base.n);
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_relaxed_sys_u64(mine.flags + threadIdx.x) < epoch) {
      if ((++spins & 0xfffu) == 0u) {                      // a peer that died must not hang this GPU for ever
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) asm volatile("trap;");
      }
    }
    __threadfence_system();   // acquire at SYSTEM scope (the releasing stores came from other GPUs): also drops stale L1 lines
  }
  __syncthreads();
}

__device__ __forceinline__ int peer_owner(const PcgPeerParams& q, int row) {
  const int s = q.sell_pos[row] / kSellC;
  int g = 0;
#pragma unroll
  for (int k = 1; k < kPeerMax; ++k) g += (k < q.world && s >= q.slice_bound[k]) ? 1 : 0;
  return g;
}

template <int V, int UNR>
__global__ void __launch_bounds__(kPeerThreads, 1)
k_pcg_peer(const PcgPeerParams q) {
  namespace cgx = cooperative_groups;
  const PcgParams& p = q.base;
  cgx::grid_group grid = cgx::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const PeerWindow me = peer_window_at(q.win[q.rank], p.n);
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  unsigned long long epoch = q.epoch_base;
  double v[kPcgNV];

  // ---- start (replicated, local): u = M^-1 b for ALL rows, "previous" (r, s) = (b, 0) of paired rows;
  //      own rows: x = 0, r = b, p = s = 0.  |b|^2 over all rows is the same number on every rank.
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  for (int s = gwarp; s < p.nslices; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 b = ldg256(p.B + row);
      const double d = p.diag[row];
      const double di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
      double4 u0 = make_double4(di * b.x, di * b.y, di * b.z, 0.0);
      if (has_pairs) {
        const int mt = p.mate[row];
        if (mt >= 0) {
          const double4 bm = ldg256(p.B + mt);
          const double c2 = p.pc2[row];
          u0.x += c2 * bm.x; u0.y += c2 * bm.y; u0.z += c2 * bm.z;
          const int m2 = p.mate2[row];
          if (m2 >= 0) {
            const double4 b2 = ldg256(p.B + m2);
            const double c3 = p.pc3[row];
            u0.x += c3 * b2.x; u0.y += c3 * b2.y; u0.z += c3 * b2.z;
          }
          st256(me.MR(1) + row, b);
          st256(me.MS(1) + row, make_double4(0, 0, 0, 0));
        }
      }
      st256(me.U + row, u0);
      if (s >= q.slice_lo && s < q.slice_hi) {
        p.dinv[row] = di;
        const double4 z4 = make_double4(0, 0, 0, 0);
        st256(p.R + row, b); st256(p.P + row, z4); st256(p.S + row, z4); st256(me.X + row, z4);
      }
      v[0] += b.x * b.x; v[1] += b.y * b.y; v[2] += b.z * b.z;
    }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  // nobody may write into a window before its owner finished initialising it
  peer_barrier(q, grid, ++epoch);
  int it = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  while (!sc_stop) {
    // ---- phase A: w = A u on my rows, partial dots ----------------------------------------------
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      const int width = p.slice_width[s];
      const int64_t base = (int64_t)p.slice_off[s] + lane;
      const double4 u = row >= 0 ? ld256(me.U + row) : make_double4(0, 0, 0, 0);
      double ax, ay, az;
      sell_row_apply<V, UNR>(p.sell_col, p.sell_w2, me.U, base, width, u, ax, ay, az);
      if (row >= 0) {
        const double4 w4 = make_double4(ax, ay, az, 0.0);
        st256(p.W + row, w4);
        if (has_pairs) {
          const int mt = p.mate[row];
          if (mt >= 0) {                                                   // my block mates need my w
            const int o1 = peer_owner(q, mt);
            st256(peer_window_at(q.win[o1], p.n).MW + row, w4);
            const int m2 = p.mate2[row];
            if (m2 >= 0) { const int o2 = peer_owner(q, m2); if (o2 != o1) st256(peer_window_at(q.win[o2], p.n).MW + row, w4); }
          }
        }
        const double4 r = ld256(p.R + row);
        v[0] += r.x * u.x; v[1] += r.y * u.y; v[2] += r.z * u.z;
        v[3] += u.x * ax;  v[4] += u.y * ay;  v[5] += u.z * az;
        v[6] += r.x * r.x; v[7] += r.y * r.y; v[8] += r.z * r.z;
      }
    }
    // rank-local block-ordered sum (pcg_grid_reduce contains one grid barrier) ...
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    // ... then the rank totals go to every rank's slot array, and everybody adds the slots in rank order
    const int par = it & 1;
    if (blockIdx.x == 0 && threadIdx.x < q.world * kPcgNV) {
      const int g = threadIdx.x / kPcgNV, k = threadIdx.x % kPcgNV;
      peer_window_at(q.win[g], p.n).dots[(par * kPeerMax + q.rank) * 16 + k] = v[k];
    }
    peer_barrier(q, grid, ++epoch);
    if (threadIdx.x < kPcgNV) {
      double t = 0.0;
      for (int g = 0; g < q.world; ++g) t += ld_relaxed_sys_f64(me.dots + (par * kPeerMax + g) * 16 + threadIdx.x);
      tot[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = tot[k];
    if (timer) { const long long c = clock64(); c_spmv += c - c_mark; c_mark = c; }
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    // ---- phase B: my rows of p, s, x, r, u; u goes to every rank, (r, s) of paired rows to the mate's owner ----
    const int cur = it & 1, old = cur ^ 1;
    for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      if (row >= 0) {
        const double4 u = ld256(me.U + row), w = ld256(p.W + row);
        double4 pp = ld256(p.P + row), ss = ld256(p.S + row);
        pp.x = u.x + b0 * pp.x; pp.y = u.y + b1 * pp.y; pp.z = u.z + b2 * pp.z;
        ss.x = w.x + b0 * ss.x; ss.y = w.y + b1 * ss.y; ss.z = w.z + b2 * ss.z;
        st256(p.P + row, pp); st256(p.S + row, ss);
        double4 x = ld256(me.X + row), r = ld256(p.R + row);
        const double di = p.dinv[row];
        x.x += a0 * pp.x; x.y += a1 * pp.y; x.z += a2 * pp.z;
        r.x -= a0 * ss.x; r.y -= a1 * ss.y; r.z -= a2 * ss.z;
        st256(me.X + row, x); st256(p.R + row, r);
        double4 un = make_double4(di * r.x, di * r.y, di * r.z, 0.0);
        if (has_pairs) {
          const int mt = p.mate[row];
          if (mt >= 0) {
            // the mate's new residual from its previous (r, s) and this iteration's w - all delivered before
            // the dot-product barrier
            const double4 rm = ld256(me.MR(old) + mt), sm = ld256(me.MS(old) + mt), wm = ld256(me.MW + mt);
            const double c2 = p.pc2[row];
            const double sm0 = wm.x + b0 * sm.x, sm1 = wm.y + b1 * sm.y, sm2 = wm.z + b2 * sm.z;
            un.x += c2 * (rm.x - a0 * sm0); un.y += c2 * (rm.y - a1 * sm1); un.z += c2 * (rm.z - a2 * sm2);
            const int o1 = peer_owner(q, mt);
            const PeerWindow mw = peer_window_at(q.win[o1], p.n);
            st256(mw.MR(cur) + row, r);
            st256(mw.MS(cur) + row, ss);
            const int m2 = p.mate2[row];
            if (m2 >= 0) {
              const double4 rn = ld256(me.MR(old) + m2), sn = ld256(me.MS(old) + m2), wn = ld256(me.MW + m2);
              const double c3 = p.pc3[row];
              const double t0 = wn.x + b0 * sn.x, t1 = wn.y + b1 * sn.y, t2 = wn.z + b2 * sn.z;
              un.x += c3 * (rn.x - a0 * t0); un.y += c3 * (rn.y - a1 * t1); un.z += c3 * (rn.z - a2 * t2);
              const int o2 = peer_owner(q, m2);
              if (o2 != o1) {
                const PeerWindow mw2 = peer_window_at(q.win[o2], p.n);
                st256(mw2.MR(cur) + row, r);
                st256(mw2.MS(cur) + row, ss);
              }
            }
          }
        }
#pragma unroll
        for (int g = 0; g < kPeerMax; ++g)
          if (g < q.world) st256(reinterpret_cast<double4*>(q.win[g] + kPeerHdr) + row, un);   // U is the first vector
      }
    }
    ++it;
    peer_barrier(q, grid, ++epoch);
    if (timer) { const long long c = clock64(); c_upd += c - c_mark; c_mark = c; }
  }
  // ---- all-gather of the solution: my rows of X into every other rank's window -------------------------
  for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 x = ld256(me.X + row);
      for (int g = 0; g < q.world; ++g)
        if (g != q.rank) st256(peer_window_at(q.win[g], p.n).X + row, x);
    }
  }
  peer_barrier(q, grid, ++epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->epoch = epoch;
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

// =====================================================================================================
// Barrier-free ("LL") variant.  The barrier version above pays ~31 us of fixed cost per PCG iteration on 2
// GPUs (tools/peer_probe.py): each cross-GPU barrier needs a system fence that waits for the acknowledgement
// of every outstanding NVLink store, a grid barrier, a flag flight and a poll.  Here no barrier and no fence
// separate the iterations - every exchanged datum validates itself:
//   * u rows: the owner forces the least-significant mantissa bit of each of the 3 doubles to the PARITY of
//     the iteration that will consume them (a <= 1 ulp change, applied before anybody - the owner included -
//     uses the value, so all ranks still compute with identical numbers).  An 8-byte store is single-copy
//     atomic, so each component is individually valid or stale; a gather whose 3 parity bits do not match
//     simply re-reads (L2-coherent load) until the row has arrived - from another SM or another GPU alike,
//     which also removes the LOCAL grid barrier after the vector update;
//   * dot products: {value, epoch} in one 16-byte store (single-copy atomic), polled per value;
//   * (r, s, w) copies for 2x2 / 3x3 blocks: tagged like u, in buffers indexed by the iteration parity k & 1 and
//     carrying tag (k >> 1) & 1, so that consecutive uses of one buffer carry different tags.
// What is left per iteration: ONE local grid barrier (inside the block-ordered dot-product reduction) and one
// NVLink store->poll flight.  Write-after-read safety comes from the dot products themselves: a rank sends its
// partial sums only after all its blocks finished the SpMV (reading u), and nobody can produce the next u
// before it has received every rank's partial sums.
// =====================================================================================================
__device__ __forceinline__ double tag_lsb(double v, int p) {
  return __longlong_as_double((__double_as_longlong(v) & ~1ll) | (long long)p);
}
__device__ __forceinline__ bool has_tag(const double4& v, int p) {
  return (((__double_as_longlong(v.x) ^ p) | (__double_as_longlong(v.y) ^ p) | (__double_as_longlong(v.z) ^ p)) & 1ll) == 0;
}
__device__ __forceinline__ double4 tag4(double x, double y, double z, int p) {
  return make_double4(tag_lsb(x, p), tag_lsb(y, p), tag_lsb(z, p), 0.0);
}
// L2-coherent load of a tagged row; spins until all three components carry parity p.
__device__ __forceinline__ double4 ld_tagged(const double4* ptr, int p) {
  double4 v = ldcg256(ptr);
  if (!has_tag(v, p)) {
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    unsigned int nap = 128;                                // back off: 100 000 polling threads would saturate L2
    do {
      __nanosleep(nap);
      if (nap < 2048) nap <<= 1;
      v = ldcg256(ptr);
      if ((++spins & 0x3ffu) == 0u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) asm volatile("trap;");
      }
    } while (!has_tag(v, p));
  }
  return v;
}
// A dot-product slot is {value, epoch}.  PTX does not promise that a 16-byte vector access is single-copy atomic
// (it may be performed as two scalars), so the pair is published with release / acquire on the epoch word instead of
// one v2 store: value first (relaxed), then the epoch with st.release.sys; the reader polls the epoch with
// ld.acquire.sys and only then reads the value.  The slot is reused two iterations later, by which time every
// reader has consumed it (a rank publishes iteration k + 2 only after it has everybody's sums of k + 1).
__device__ __forceinline__ void st_dot16(double* slot, double v, unsigned long long epoch) {
  asm volatile("st.relaxed.sys.global.b64 [%0], %1;" :: "l"(slot), "l"(__double_as_longlong(v)) : "memory");
  asm volatile("st.release.sys.global.b64 [%0], %1;" :: "l"(slot + 1), "l"(epoch) : "memory");
}
__device__ __forceinline__ double ld_dot16(const double* slot, unsigned long long epoch) {
  long long v; unsigned long long e;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
  for (;;) {
    asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(e) : "l"(slot + 1) : "memory");
    if (e == epoch) break;
    if ((++spins & 0xfffu) == 0u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20000000000ull) asm volatile("trap;");
    }
  }
  asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
  return __longlong_as_double(v);
}

// one lane's row of (A^T D^2 A) u with self-validating gathers (4 in flight)
__device__ __forceinline__ void sell_row_apply_tagged(const int* __restrict__ sell_col, const double* __restrict__ sell_w2,
                                                      const double4* U, int64_t base, int width, const double4 u, int par,
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
  for (int j = 0; j < width; j += 4) {
    double4 uc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) uc[q] = ldcg256(U + c[q]);
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
      if (!has_tag(uc[q], par)) uc[q] = ld_tagged(U + c[q], par);          // not there yet: wait for this row only
      ax += w2[q] * (u.x - uc[q].x); ay += w2[q] * (u.y - uc[q].y); az += w2[q] * (u.z - uc[q].z);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) { c[q] = cn[q]; w2[q] = wn[q]; }
  }
}

// Window layout of the LL variant: U | X | MR[2] | MS[2] | MW[2] (9 n double4), dots [2][kPeerMax][16] as
// {value, epoch} pairs (2 doubles each -> [2][kPeerMax][32]), flags.
struct PeerWindowLL {          // no arrays inside: indexing a local struct at run time would put it on the stack
  double4 *U, *X, *MR0, *MS0, *MW0;   // buffer b of MR / MS / MW = pointer + b * stride
  size_t stride;
  double* dots;
  unsigned long long* flags;
  __host__ __device__ double4* MR(int b) const { return MR0 + (size_t)b * stride; }
  __host__ __device__ double4* MS(int b) const { return MS0 + (size_t)b * stride; }
  __host__ __device__ double4* MW(int b) const { return MW0 + (size_t)b * stride; }
};
__host__ __device__ inline size_t peer_window_ll_bytes(int npos) { return kPeerHdr + (size_t)9 * npos * sizeof(double4); }
__host__ __device__ inline PeerWindowLL peer_window_ll_at(unsigned char* base, int n) {
  PeerWindowLL w;
  double4* v = reinterpret_cast<double4*>(base + kPeerHdr);
  w.U = v; w.X = v + (size_t)n;
  w.MR0 = v + (size_t)2 * n; w.MS0 = v + (size_t)4 * n; w.MW0 = v + (size_t)6 * n;
  w.stride = (size_t)n;
  w.flags = reinterpret_cast<unsigned long long*>(base);
  w.dots = reinterpret_cast<double*>(base + 256);
  return w;
}

// fence-based barrier on the LL window (start and end of a solve only)
__device__ __forceinline__ void peer_barrier_ll(const PcgPeerParams& q, cooperative_groups::grid_group& grid,
                                                unsigned long long epoch) {
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < q.world)
    st_relaxed_sys(peer_window_ll_at(q.win[threadIdx.x], q.base.n).flags + q.rank, epoch);
  if (threadIdx.x < q.world) {
    const unsigned long long* f = peer_window_ll_at(q.win[q.rank], q.base.n).flags + threadIdx.x;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_relaxed_sys_u64(f) < epoch) {
      if ((++spins & 0xfffu) == 0u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) asm volatile("trap;");
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kPeerThreads, 1)
k_pcg_peer_ll(const PcgPeerParams q) {
  namespace cgx = cooperative_groups;
  const PcgParams& p = q.base;
  cgx::grid_group grid = cgx::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const PeerWindowLL me = peer_window_ll_at(q.win[q.rank], q.npos);
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  unsigned long long epoch = q.epoch_base;
  double v[kPcgNV];

  // ---- start (replicated, local): u0 = M^-1 b for ALL rows with parity 0; block members' "previous"
  //      (r, s) = (b, 0) with parity 1; own rows: x = 0, r = b, p = s = 0
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  for (int s = gwarp; s < p.nslices; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 b = ldg256(p.B + row);
      const double d = p.diag[row];
      const double di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
      double4 u0 = make_double4(di * b.x, di * b.y, di * b.z, 0.0);
      if (has_pairs) {
        const int mt = p.mate[row];
        if (mt >= 0) {
          const double4 bm = ldg256(p.B + mt);
          const double c2 = p.pc2[row];
          u0.x += c2 * bm.x; u0.y += c2 * bm.y; u0.z += c2 * bm.z;
          const int m2 = p.mate2[row];
          if (m2 >= 0) {
            const double4 b2 = ldg256(p.B + m2);
            const double c3 = p.pc3[row];
            u0.x += c3 * b2.x; u0.y += c3 * b2.y; u0.z += c3 * b2.z;
          }
          // buffer k & 1 carries tag (k >> 1) & 1 for data of iteration k: the "previous" iteration -1 lives in
          // buffer 1 with tag 1; every other buffer gets the tag its first writer will NOT use, so that stale
          // rows of an earlier solve can never be mistaken for fresh ones
          const double4 inval = tag4(0.0, 0.0, 0.0, 1);
          st256(me.MR(1) + row, tag4(b.x, b.y, b.z, 1));
          st256(me.MS(1) + row, inval);
          st256(me.MR(0) + row, inval); st256(me.MS(0) + row, inval);
          st256(me.MW(0) + row, inval); st256(me.MW(1) + row, inval);
        }
      }
      st256(me.U + (s * kSellC + lane), tag4(u0.x, u0.y, u0.z, 0));   // u lives in SELL-position order: a warp's 32 rows are 1 KB contiguous
      if (s >= q.slice_lo && s < q.slice_hi) {
        p.dinv[row] = di;
        const double4 z4 = make_double4(0, 0, 0, 0);
        st256(p.R + row, b); st256(p.P + row, z4); st256(p.S + row, z4); st256(me.X + row, z4);
      }
      v[0] += b.x * b.x; v[1] += b.y * b.y; v[2] += b.z * b.z;
    }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  peer_barrier_ll(q, grid, ++epoch);       // nobody writes into a window before its owner initialised it
  int it = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  long long d_mv = 0, d_red = 0, d_dot = 0, d_t = 0;      // debug split of phase A (q.debug)
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  while (!sc_stop) {
    const int par = it & 1, old = par ^ 1;
    const int tcur = (it >> 1) & 1, told = ((it - 1) >> 1) & 1;      // tags of this / the previous iteration's copies
    // ---- phase A: w = A u on my rows (self-validating gathers), partial dots -------------------------
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      const int width = p.slice_width[s];
      const int64_t base = (int64_t)p.slice_off[s] + lane;
      const double4 u = row >= 0 ? ld_tagged(me.U + (s * kSellC + lane), par) : make_double4(0, 0, 0, 0);
      double ax, ay, az;
      sell_row_apply_tagged(q.sell_colpos, p.sell_w2, me.U, base, width, u, par, ax, ay, az);
      if (row >= 0) {
        st256(p.W + row, make_double4(ax, ay, az, 0.0));
        if (has_pairs) {
          const int mt = p.mate[row];
          if (mt >= 0) {                                                   // my block mates need a copy of my w
            const double4 wt = tag4(ax, ay, az, tcur);
            const int o1 = peer_owner(q, mt);
            st256(peer_window_ll_at(q.win[o1], q.npos).MW(par) + row, wt);
            const int m2 = p.mate2[row];
            if (m2 >= 0) { const int o2 = peer_owner(q, m2); if (o2 != o1) st256(peer_window_ll_at(q.win[o2], q.npos).MW(par) + row, wt); }
          }
        }
        const double4 r = ld256(p.R + row);
        v[0] += r.x * u.x; v[1] += r.y * u.y; v[2] += r.z * u.z;
        v[3] += u.x * ax;  v[4] += u.y * ay;  v[5] += u.z * az;
        v[6] += r.x * r.x; v[7] += r.y * r.y; v[8] += r.z * r.z;
      }
    }
    if (timer) { d_t = clock64(); d_mv += d_t - c_mark; }
    pcg_grid_reduce(v, p.partials, grid, red, tot);        // rank-local block-ordered sum; the only grid barrier
    if (timer) { const long long c = clock64(); d_red += c - d_t; d_t = c; }
    ++epoch;
    if (blockIdx.x == 0 && threadIdx.x < q.world * kPcgNV) {
      const int g = threadIdx.x / kPcgNV, k = threadIdx.x % kPcgNV;
      st_dot16(peer_window_ll_at(q.win[g], q.npos).dots + ((par * kPeerMax + q.rank) * 16 + k) * 2, v[k], epoch);
    }
    if (threadIdx.x < kPcgNV) {
      double t = 0.0;
      for (int g = 0; g < q.world; ++g) t += ld_dot16(me.dots + ((par * kPeerMax + g) * 16 + threadIdx.x) * 2, epoch);
      tot[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = tot[k];
    if (timer) { const long long c = clock64(); c_spmv += c - c_mark; c_mark = c; d_dot += c - d_t; }
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    // ---- phase B: my rows of p, s, x, r, u; the new u (parity of the NEXT iteration) goes to every rank ----
    const int nxt = old;                                   // (it + 1) & 1
    for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      if (row >= 0) {
        const double4 u = ldcg256(me.U + (s * kSellC + lane)), w = ld256(p.W + row);   // both written by this thread's own earlier phases
        double4 pp = ld256(p.P + row), ss = ld256(p.S + row);
        pp.x = u.x + b0 * pp.x; pp.y = u.y + b1 * pp.y; pp.z = u.z + b2 * pp.z;
        ss.x = w.x + b0 * ss.x; ss.y = w.y + b1 * ss.y; ss.z = w.z + b2 * ss.z;
        st256(p.P + row, pp); st256(p.S + row, ss);
        double4 x = ld256(me.X + row), r = ld256(p.R + row);
        const double di = p.dinv[row];
        x.x += a0 * pp.x; x.y += a1 * pp.y; x.z += a2 * pp.z;
        r.x -= a0 * ss.x; r.y -= a1 * ss.y; r.z -= a2 * ss.z;
        st256(me.X + row, x); st256(p.R + row, r);
        double ux = di * r.x, uy = di * r.y, uz = di * r.z;
        if (has_pairs) {
          const int mt = p.mate[row];
          if (mt >= 0) {
            // a mate's new residual from its previous (r, s) (parity `old`) and this iteration's w (parity `par`)
            // issue every load first, validate afterwards: one L2 round trip instead of up to six in a row
            const int m2 = p.mate2[row];
            double4 rm = ldcg256(me.MR(old) + mt), sm = ldcg256(me.MS(old) + mt), wm = ldcg256(me.MW(par) + mt);
            double4 rn = make_double4(0, 0, 0, 0), sn = rn, wn = rn;
            if (m2 >= 0) { rn = ldcg256(me.MR(old) + m2); sn = ldcg256(me.MS(old) + m2); wn = ldcg256(me.MW(par) + m2); }
            if (!has_tag(rm, told)) rm = ld_tagged(me.MR(old) + mt, told);
            if (!has_tag(sm, told)) sm = ld_tagged(me.MS(old) + mt, told);
            if (!has_tag(wm, tcur)) wm = ld_tagged(me.MW(par) + mt, tcur);
            const double c2 = p.pc2[row];
            const double sm0 = wm.x + b0 * sm.x, sm1 = wm.y + b1 * sm.y, sm2 = wm.z + b2 * sm.z;
            ux += c2 * (rm.x - a0 * sm0); uy += c2 * (rm.y - a1 * sm1); uz += c2 * (rm.z - a2 * sm2);
            const double4 rt = tag4(r.x, r.y, r.z, tcur), st = tag4(ss.x, ss.y, ss.z, tcur);
            const int o1 = peer_owner(q, mt);
            const PeerWindowLL mw = peer_window_ll_at(q.win[o1], q.npos);
            st256(mw.MR(par) + row, rt);
            st256(mw.MS(par) + row, st);
            if (m2 >= 0) {
              if (!has_tag(rn, told)) rn = ld_tagged(me.MR(old) + m2, told);
              if (!has_tag(sn, told)) sn = ld_tagged(me.MS(old) + m2, told);
              if (!has_tag(wn, tcur)) wn = ld_tagged(me.MW(par) + m2, tcur);
              const double c3 = p.pc3[row];
              const double t0 = wn.x + b0 * sn.x, t1 = wn.y + b1 * sn.y, t2 = wn.z + b2 * sn.z;
              ux += c3 * (rn.x - a0 * t0); uy += c3 * (rn.y - a1 * t1); uz += c3 * (rn.z - a2 * t2);
              const int o2 = peer_owner(q, m2);
              if (o2 != o1) {
                const PeerWindowLL mw2 = peer_window_ll_at(q.win[o2], q.npos);
                st256(mw2.MR(par) + row, rt);
                st256(mw2.MS(par) + row, st);
              }
            }
          }
        }
        const double4 un = tag4(ux, uy, uz, nxt);
#pragma unroll
        for (int g = 0; g < kPeerMax; ++g)                 // 32 lanes x 32 B contiguous per destination: full NVLink packets
          if (g < q.world) st256(reinterpret_cast<double4*>(q.win[g] + kPeerHdr) + (s * kSellC + lane), un);   // U is the first vector
      }
    }
    ++it;
    if (timer) { const long long c = clock64(); c_upd += c - c_mark; c_mark = c; }
  }
  // ---- all-gather of the solution ----------------------------------------------------------------------
  for (int s = q.slice_lo + gwarp; s < q.slice_hi; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 x = ld256(me.X + row);
      for (int g = 0; g < q.world; ++g)
        if (g != q.rank) st256(peer_window_ll_at(q.win[g], q.npos).X + row, x);
    }
  }
  peer_barrier_ll(q, grid, ++epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->epoch = epoch;
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
    if (q.debug && q.rank == 0)
      printf("[k_pcg_peer_ll] iters %d: cycles/iter SpMV %lld, local reduce %lld, dot exchange %lld, update %lld\n", it,
             d_mv / (it + 1), d_red / (it + 1), d_dot / (it + 1), c_upd / (it > 0 ? it : 1));
  }
}

// the same barrier with the window bases in shared memory (indexing the kernel-parameter array at run time would
// copy the whole parameter block to every thread's stack)
__device__ __forceinline__ void peer_barrier_ll_s(unsigned char* const* s_win, int rank, int world,
                                                  cooperative_groups::grid_group& grid, unsigned long long epoch) {
  __syncthreads();
  if (threadIdx.x == 0) __threadfence_system();
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < world)
    st_relaxed_sys(reinterpret_cast<unsigned long long*>(s_win[threadIdx.x]) + rank, epoch);
  if (threadIdx.x < world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(s_win[rank]) + threadIdx.x;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_relaxed_sys_u64(f) < epoch) {
      if ((++spins & 0xfffu) == 0u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 20000000000ull) asm volatile("trap;");
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// Register-resident form of the barrier-free kernel: when a rank's slices fit one per warp (rows per rank <=
// grid x warps x 32), every lane owns ONE row for the whole solve and keeps x, r, p, s, u and its block
// coefficients in registers, like k_pcg_persistent_reg.  The vector update then touches memory only to publish
// the new u (and the block copies): the 13 x 32 B per row of L2 traffic of the HBM-resident form - which is what
// bounded its update phase (6 us for 50 000 rows, tools/peer_probe.py) - disappear.
__global__ void __launch_bounds__(kPcgThreads, 1)
k_pcg_peer_ll_reg(const PcgPeerParams q) {
  namespace cgx = cooperative_groups;
  const PcgParams& p = q.base;
  cgx::grid_group grid = cgx::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double totx[kPcgNV];                          // sums over the ranks (tot: this rank's totals)
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  __shared__ unsigned char* s_win[kPeerMax];               // kernel-parameter arrays indexed at run time would be
  if (threadIdx.x == 0) {                                  // copied to every thread's stack
#pragma unroll
    for (int g = 0; g < kPeerMax; ++g) s_win[g] = q.win[g];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  const PeerWindowLL me = peer_window_ll_at(s_win[q.rank], q.npos);
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  unsigned long long epoch = q.epoch_base;
  double v[kPcgNV];

  // ---- start, replicated and local (same as k_pcg_peer_ll): tagged u0 of ALL rows, block rows' buffers ----
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  for (int s = gwarp; s < p.nslices; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 b = ldg256(p.B + row);
      const double d = p.diag[row];
      const double di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
      double4 u0 = make_double4(di * b.x, di * b.y, di * b.z, 0.0);
      if (has_pairs) {
        const int mt = p.mate[row];
        if (mt >= 0) {
          const double4 bm = ldg256(p.B + mt);
          const double c2 = p.pc2[row];
          u0.x += c2 * bm.x; u0.y += c2 * bm.y; u0.z += c2 * bm.z;
          const int m2 = p.mate2[row];
          if (m2 >= 0) {
            const double4 b2 = ldg256(p.B + m2);
            const double c3 = p.pc3[row];
            u0.x += c3 * b2.x; u0.y += c3 * b2.y; u0.z += c3 * b2.z;
          }
          const double4 inval = tag4(0.0, 0.0, 0.0, 1);
          st256(me.MR(1) + row, tag4(b.x, b.y, b.z, 1));
          st256(me.MS(1) + row, inval);
          st256(me.MR(0) + row, inval); st256(me.MS(0) + row, inval);
          st256(me.MW(0) + row, inval); st256(me.MW(1) + row, inval);
        }
      }
      st256(me.U + (s * kSellC + lane), tag4(u0.x, u0.y, u0.z, 0));
      v[0] += b.x * b.x; v[1] += b.y * b.y; v[2] += b.z * b.z;
    }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  peer_barrier_ll_s(s_win, q.rank, q.world, grid, ++epoch);

  // ---- this lane's row -------------------------------------------------------------------------------
  const int slice = q.slice_lo + gwarp;                    // <= 1 slice per warp (checked by the host)
  int row = -1, width = 0, mt = -1, mt2 = -1, o1 = 0, o2 = 0, pos = 0;
  int64_t base = 0;
  double x0 = 0, x1 = 0, x2 = 0, r0 = 0, r1 = 0, r2 = 0, p0 = 0, p1 = 0, p2 = 0, s0 = 0, s1 = 0, s2 = 0;
  double u0 = 0, u1 = 0, u2 = 0, di = 0, c2 = 0, c3 = 0;
  if (slice < q.slice_hi) {
    pos = slice * kSellC + lane;
    row = p.sell_row[pos];
    width = p.slice_width[slice];
    base = (int64_t)p.slice_off[slice] + lane;
  }
  if (row >= 0) {
    const double4 b = ldg256(p.B + row);
    const double d = p.diag[row];
    di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
    r0 = b.x; r1 = b.y; r2 = b.z;
    const double4 ut = ldcg256(me.U + pos);                // the tagged value everybody else will gather
    u0 = ut.x; u1 = ut.y; u2 = ut.z;
    if (has_pairs) {
      mt = p.mate[row];
      if (mt >= 0) {
        c2 = p.pc2[row];
        o1 = peer_owner(q, mt);
        mt2 = p.mate2[row];
        if (mt2 >= 0) { c3 = p.pc3[row]; o2 = peer_owner(q, mt2); }
      }
    }
  }
  int it = 0;
  // only the wall time of the solve is kept (block 0): phase clocks would cost every thread ~10 registers, and
  // this kernel is already 600 B of stack over budget at 85 registers per thread
  unsigned long long ns_begin = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin));

  while (!sc_stop) {
    const int par = it & 1, old = par ^ 1;
    const int tcur = (it >> 1) & 1, told = ((it - 1) >> 1) & 1;
    // ---- phase A ----
    double w0 = 0, w1 = 0, w2 = 0;
    if (row >= 0 || width > 0)
      sell_row_apply_tagged(q.sell_colpos, p.sell_w2, me.U, base, width, make_double4(u0, u1, u2, 0.0), par, w0, w1, w2);
    if (row >= 0) {
      if (mt >= 0) {
        const double4 wt = tag4(w0, w1, w2, tcur);
        st256(peer_window_ll_at(s_win[o1], q.npos).MW(par) + row, wt);
        if (mt2 >= 0 && o2 != o1) st256(peer_window_ll_at(s_win[o2], q.npos).MW(par) + row, wt);
      }
      v[0] = r0 * u0; v[1] = r1 * u1; v[2] = r2 * u2;
      v[3] = u0 * w0; v[4] = u1 * w1; v[5] = u2 * w2;
      v[6] = r0 * r0; v[7] = r1 * r1; v[8] = r2 * r2;
    } else {
#pragma unroll
      for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    ++epoch;
    if (blockIdx.x == 0 && threadIdx.x < q.world * kPcgNV) {
      const int g = threadIdx.x / kPcgNV, k = threadIdx.x % kPcgNV;
      st_dot16(peer_window_ll_at(s_win[g], q.npos).dots + ((par * kPeerMax + q.rank) * 16 + k) * 2, tot[k], epoch);   // tot: the rank totals
    }
    if (threadIdx.x < kPcgNV) {
      double t = 0.0;
      for (int g = 0; g < q.world; ++g) t += ld_dot16(me.dots + ((par * kPeerMax + g) * 16 + threadIdx.x) * 2, epoch);
      totx[threadIdx.x] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = totx[k];
    pcg_coefficients(totx, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    // ---- phase B (registers) ----
    if (row >= 0) {
      p0 = u0 + b0 * p0; p1 = u1 + b1 * p1; p2 = u2 + b2 * p2;
      s0 = w0 + b0 * s0; s1 = w1 + b1 * s1; s2 = w2 + b2 * s2;
      x0 += a0 * p0; x1 += a1 * p1; x2 += a2 * p2;
      r0 -= a0 * s0; r1 -= a1 * s1; r2 -= a2 * s2;
      double ux = di * r0, uy = di * r1, uz = di * r2;
      if (mt >= 0) {
        {                                                  // publish my (r, s) first: the mates are waiting for them
          const double4 rt = tag4(r0, r1, r2, tcur), st = tag4(s0, s1, s2, tcur);
          const PeerWindowLL mw = peer_window_ll_at(s_win[o1], q.npos);
          st256(mw.MR(par) + row, rt);
          st256(mw.MS(par) + row, st);
          if (mt2 >= 0 && o2 != o1) {
            const PeerWindowLL mw2 = peer_window_ll_at(s_win[o2], q.npos);
            st256(mw2.MR(par) + row, rt);
            st256(mw2.MS(par) + row, st);
          }
        }
        {
          double4 rm = ldcg256(me.MR(old) + mt), sm = ldcg256(me.MS(old) + mt), wm = ldcg256(me.MW(par) + mt);
          if (!has_tag(rm, told)) rm = ld_tagged(me.MR(old) + mt, told);
          if (!has_tag(sm, told)) sm = ld_tagged(me.MS(old) + mt, told);
          if (!has_tag(wm, tcur)) wm = ld_tagged(me.MW(par) + mt, tcur);
          ux += c2 * (rm.x - a0 * (wm.x + b0 * sm.x)); uy += c2 * (rm.y - a1 * (wm.y + b1 * sm.y));
          uz += c2 * (rm.z - a2 * (wm.z + b2 * sm.z));
        }
        if (mt2 >= 0) {
          double4 rn = ldcg256(me.MR(old) + mt2), sn = ldcg256(me.MS(old) + mt2), wn = ldcg256(me.MW(par) + mt2);
          if (!has_tag(rn, told)) rn = ld_tagged(me.MR(old) + mt2, told);
          if (!has_tag(sn, told)) sn = ld_tagged(me.MS(old) + mt2, told);
          if (!has_tag(wn, tcur)) wn = ld_tagged(me.MW(par) + mt2, tcur);
          ux += c3 * (rn.x - a0 * (wn.x + b0 * sn.x)); uy += c3 * (rn.y - a1 * (wn.y + b1 * sn.y));
          uz += c3 * (rn.z - a2 * (wn.z + b2 * sn.z));
        }
      }
      const double4 un = tag4(ux, uy, uz, old);            // parity of the NEXT iteration
      u0 = un.x; u1 = un.y; u2 = un.z;
#pragma unroll
      for (int g = 0; g < kPeerMax; ++g)
        if (g < q.world) st256(reinterpret_cast<double4*>(s_win[g] + kPeerHdr) + pos, un);
    }
    if (q.flush == 1) { __syncwarp(); if (lane == 0) __threadfence_system(); }
    else if (q.flush == 2) { __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
    ++it;
  }
  if (row >= 0) {
    const double4 x = make_double4(x0, x1, x2, 0.0);
    for (int g = 0; g < q.world; ++g) st256(peer_window_ll_at(s_win[g], q.npos).X + row, x);
  }
  peer_barrier_ll_s(s_win, q.rank, q.world, grid, ++epoch);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->epoch = epoch;
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
    unsigned long long ns_end;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_end));
    p.ctl->cyc_total += (long long)(ns_end - ns_begin);     // 1 "cycle" = 1 ns here: the host only needs the ratio
    p.ctl->ns_total += (long long)(ns_end - ns_begin);
    p.ctl->pcg_spmv_phases += it + 1;
  }
}

// SELL slot -> position of its column
__global__ void k_sell_colpos(const int* __restrict__ sell_col, const int* __restrict__ sell_pos, int64_t total,
                              int* __restrict__ colpos) {
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x)
    colpos[o] = sell_pos[sell_col[o]];
}

// row -> SELL position
__global__ void k_sell_inverse(const int* __restrict__ sell_row, int npos, int* __restrict__ sell_pos) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos < npos) { const int r = sell_row[pos]; if (r >= 0) sell_pos[r] = pos; }
}

}  // namespace ira
