// Matrix-resident persistent PCG: the whole SELL copy of A^T D^2 A lives in the 148 SMs' SHARED MEMORY for the
// duration of one linear solve (replaces ls_solve, ral/l1_irls.cpp:536-556).
//
// Why (profiles/r01_ncu_prof_pcg_reg_v6.json): the SpMV phase of k_pcg_persistent_reg moves
// ~64 MB of gathered 32 B sectors of u plus ~26 MB of (col, w2) streams through L2 -> SM every PCG iteration and
// sits on the L2 sector-bandwidth ceiling (~6 TB/s): 15.5 us against a 10.4 us floor for the gathers alone.
// The matrix never changes inside a solve and one SM's share of it at config 3 (n = 100 000, m = 1 000 000:
// 21-22 slices x 32 rows x ~22 entries x 12 B ~ 180 KB) fits the 227 KB of shared memory an sm_100 block can
// own.  So each block copies its slices' (col, w2) once per solve (26 MB read from L2/HBM once, coalesced) and
// every iteration's SpMV reads them from shared memory: only the random gathers of u and the 32 B/row store of
// the new u still touch L2.  Without the software pipeline of the global streams the kernel also needs ~20 fewer
// registers (and x, p move to shared memory too): no spills at the 80 registers 22 warps per SM leave (build.log),
// where k_pcg_persistent_reg spilled 360 B.
//
// MEASURED RESULT (B200, config 3, L1, 30 IRLS iterations, profiles/r02_ab_pcg2_smem.json): 78 ms per step against
// 52 ms for k_pcg_persistent_reg - SLOWER, so this kernel is opt-in (ira_options.solver & 32) and documents the
// experiment.  Why: a block that owns 227 KB of shared memory leaves the SM ~28 KB of L1, and on sm_100 every
// outstanding global load holds an L1 line: the ~2 800 gathers a block wants in flight no longer fit, and the SpMV
// phase turns latency bound.  The 26 MB of streams it saves were worth at most ~1.5 us of the 15 us phase.
// (Also measured on the way: 22 warps per SM leave 80 registers per thread, not 88 - 6 warps per SM sub-partition
// share its 16 384 registers - so 704 threads buy nothing over 768.)
//
// Everything else is the single-reduction (Chronopoulos-Gear) PCG of ira_pcg.cuh with the exact 2x2 / 3x3 block
// preconditioner, one row per lane, x r p s in registers.  Differences:
//   * alpha/beta of the three right-hand sides are computed by three lanes in parallel (was: one thread, 9
//     dependent FP64 divisions, ~1.1 us per iteration);
//   * a column that has converged (|r_c| <= rtol |b_c|) is FROZEN: alpha_c = beta_c = 0 from then on, so its
//     solution stays bit-fixed while the other columns finish (round-off in gamma/delta of a converged column
//     can no longer perturb it).
// (Padding slots, col = row and w2 = 0, still issue their gather: predicating it put every gather behind a divergent
// branch whose other side rewrites the destination registers - the warp then waits for each load in turn; measured
// 37 us per PCG iteration instead of 22.)
// If a block's slices do not fit, every slice keeps its first `wcap` entry columns in shared memory and reads
// the rest from global memory (graphs up to ~113 664 rows per GPU; larger ones use k_pcg_persistent).
#pragma once
#include "ira_pcg.cuh"

namespace ira {

constexpr int kPcg2Threads = 704;                 // 22 warps: 148 x 22 = 3 256 slices >= 3 136 (config 3)
constexpr int kPcg2Warps = kPcg2Threads / 32;
constexpr int kPcg2StateDoubles = 6;              // x and p of every thread live in shared memory

struct Pcg2Params {
  PcgRegParams reg;
  int wcap;            // entry columns of every slice kept in shared memory (multiple of 4)
  int smem_entries;    // capacity check (entries of 12 B)
};

__global__ void __launch_bounds__(kPcg2Threads, 1)
k_pcg_smem(const Pcg2Params q2) {
  const PcgRegParams& q = q2.reg;
  const PcgParams& p = q.base;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slice = blockIdx.x + gridDim.x * warp;                   // <= 1 slice per warp
  const bool has_pairs = p.npairs != nullptr && *p.npairs > 0;
  int row = -1, width = 0, mt = -1, mt2 = -1;
  int64_t base = 0;
  if (slice < p.nslices) {
    row = p.sell_row[slice * kSellC + lane];
    width = p.slice_width[slice];
    base = (int64_t)p.slice_off[slice] + lane;
  }
  // ---- this warp's window of the shared-memory matrix ---------------------------------------------------------
  const int cw = min(width, q2.wcap);                                 // cached entry columns of this slice
  int off = 0;                                                        // entries before this warp's window
  for (int w = 0; w < warp; ++w) {
    const int s = blockIdx.x + gridDim.x * w;
    if (s < p.nslices) off += min(p.slice_width[s], q2.wcap) * kSellC;
  }
  // dynamic shared memory: x and p (6 x blockDim doubles: private to their thread, kept out of the 80-register budget
  // that 22 warps per SM leave - 6 warps per SM sub-partition x 32 lanes x 80 registers of its 16 384) | w2 | col
  double* const xs = reinterpret_cast<double*>(dyn_smem) + threadIdx.x;
  double* const ps = xs + 3 * kPcg2Threads;
  double* const w2s = reinterpret_cast<double*>(dyn_smem) + kPcg2StateDoubles * kPcg2Threads;   // [smem_entries]
  int* const cols = reinterpret_cast<int*>(w2s + q2.smem_entries);
  for (int j = 0; j < cw; ++j) {                                      // coalesced 128 B / 256 B lines
    const int64_t o = base + (int64_t)j * kSellC;
    cols[off + j * kSellC + lane] = __ldg(p.sell_col + o);
    w2s[off + j * kSellC + lane] = __ldg(p.sell_w2 + o);
  }
  __syncwarp();
  const int* const mycols = cols + off + lane;
  const double* const myw2 = w2s + off + lane;

  double v[kPcgNV];
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  double r0 = 0, r1 = 0, r2 = 0, s0 = 0, s1 = 0, s2 = 0;
  double u0 = 0, u1 = 0, u2 = 0, di = 0;
  xs[0] = 0.0; xs[kPcg2Threads] = 0.0; xs[2 * kPcg2Threads] = 0.0;
  ps[0] = 0.0; ps[kPcg2Threads] = 0.0; ps[2 * kPcg2Threads] = 0.0;
  if (row >= 0) {
    const double4 b = ldg256(p.B + row);
    const double d = p.diag[row];
    di = p.pc1 ? p.pc1[row] : (d > 0.0 ? 1.0 / d : 0.0);
    r0 = b.x; r1 = b.y; r2 = b.z;
    u0 = di * r0; u1 = di * r1; u2 = di * r2;
    if (has_pairs) {
      mt = p.mate[row];
      if (mt >= 0) {
        const double c2 = p.pc2[row];
        const double4 bm = ldg256(p.B + mt);
        u0 += c2 * bm.x; u1 += c2 * bm.y; u2 += c2 * bm.z;
        mt2 = p.mate2[row];
        if (mt2 >= 0) {
          const double c3 = p.pc3[row];
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
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; sc_a[c] = 0.0; sc_b[c] = 0.0;
  }
  if (threadIdx.x == 0) sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  __syncthreads();
  int it = 0;
  const bool timer = blockIdx.x == 0 && threadIdx.x == 0;
  long long c_spmv = 0, c_upd = 0, c_mark = 0, c_begin = 0;
  unsigned long long ns_begin = 0;
  if (timer) { c_begin = c_mark = clock64(); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ns_begin)); }

  while (!sc_stop) {
    // ---- w = A u: matrix from shared memory, u gathered through L2 -----------------------------------------
    double w0 = 0, w1 = 0, w2 = 0;
    {
      int j = 0;
      for (; j < cw; j += 4) {
        int c[4]; double ww[4]; double4 g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { c[t] = mycols[(j + t) * kSellC]; ww[t] = myw2[(j + t) * kSellC]; }
#pragma unroll
        for (int t = 0; t < 4; ++t) g[t] = ld256(p.U + c[t]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { w0 += ww[t] * (u0 - g[t].x); w1 += ww[t] * (u1 - g[t].y); w2 += ww[t] * (u2 - g[t].z); }
      }
      for (; j < width; j += 4) {                                    // entry columns that did not fit
        int c[4]; double ww[4]; double4 g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int64_t o = base + (int64_t)(j + t) * kSellC;
          c[t] = __ldg(p.sell_col + o); ww[t] = __ldg(p.sell_w2 + o);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) g[t] = ld256(p.U + c[t]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { w0 += ww[t] * (u0 - g[t].x); w1 += ww[t] * (u1 - g[t].y); w2 += ww[t] * (u2 - g[t].z); }
      }
    }
    if (row >= 0) {
      if (mt >= 0) st256(p.W + row, make_double4(w0, w1, w2, 0.0));   // the mates need it after the barrier
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
      const double p0 = u0 + b0 * ps[0], p1 = u1 + b1 * ps[kPcg2Threads], p2 = u2 + b2 * ps[2 * kPcg2Threads];
      ps[0] = p0; ps[kPcg2Threads] = p1; ps[2 * kPcg2Threads] = p2;
      s0 = w0 + b0 * s0; s1 = w1 + b1 * s1; s2 = w2 + b2 * s2;
      xs[0] += a0 * p0; xs[kPcg2Threads] += a1 * p1; xs[2 * kPcg2Threads] += a2 * p2;
      r0 -= a0 * s0; r1 -= a1 * s1; r2 -= a2 * s2;
      u0 = di * r0; u1 = di * r1; u2 = di * r2;
      if (mt >= 0) {
        const double c2 = p.pc2[row];                                 // block rows only: reloaded (L1) to save registers
        double4* const curR = (it & 1) ? q.RS1r : q.RS0r;
        double4* const curS = (it & 1) ? q.RS1s : q.RS0s;
        const double4* const oldR = (it & 1) ? q.RS0r : q.RS1r;
        const double4* const oldS = (it & 1) ? q.RS0s : q.RS1s;
        const double4 rm = ld256(oldR + mt), sm = ld256(oldS + mt), wm = ld256(p.W + mt);
        const double sm0 = wm.x + b0 * sm.x, sm1 = wm.y + b1 * sm.y, sm2 = wm.z + b2 * sm.z;   // the mate's new s
        u0 += c2 * (rm.x - a0 * sm0); u1 += c2 * (rm.y - a1 * sm1); u2 += c2 * (rm.z - a2 * sm2);
        if (mt2 >= 0) {
          const double c3 = p.pc3[row];
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
  if (row >= 0) st256(p.X + row, make_double4(xs[0], xs[kPcg2Threads], xs[2 * kPcg2Threads], 0.0));
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

}  // namespace ira
