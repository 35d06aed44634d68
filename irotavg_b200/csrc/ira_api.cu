// Host side of the C ABI (include/ira.h): context, HBM workspace, CSR build, the IRLS / PCG
// drivers, NCCL plumbing.  The loop structure restates irotavg::irls (ral/l1_irls.cpp:559-752);
// the SuiteSparseQR least-squares solve (ral/l1_irls.cpp:536-556) is replaced by Jacobi-PCG on
// the weighted normal equations.  No CPU compute path exists here.
#include "../../include/ira.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cub/device/device_radix_sort.cuh>
#include <limits>
#include <string>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "ira_kernels.cuh"
#include "ira_pcg.cuh"
#include "ira_pcg2.cuh"
#include "ira_l1ra.cuh"
#include "ira_mst.cuh"
#include "ira_small.cuh"
#include "ira_coarse.cuh"
#include "ira_peer.cuh"

using namespace ira;

// ---------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    // A buffer that has to GROW belongs to a growing graph (the rotAvg stream: every global call is a few hundred
    // views larger than the last, and re-allocating ~40 buffers costs more than the solve): double it.  First
    // allocations, and anything beyond 1 GB, get 12.5 % of headroom only.
    size_t want = bytes + bytes / 8 + 256;
    if (p && bytes < (size_t(1) << 30)) want = std::max(want, 2 * cap);
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

enum KClass { KC_RESIDUAL = 0, KC_RHS, KC_SPMV, KC_CGVEC, KC_WEIGHTS, KC_UPDATE, KC_COMM, KC_PCG, KC_N };

// NCCL resolved at run time so that the single-GPU library has no link-time dependency on it.
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load() {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && AllGather && CommDestroy && GetErrorString;
  }
};
NcclApi g_nccl;

}  // namespace

struct ira_context {
  ira_options opt;
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  std::string err;

  // problem
  bool uploaded = false;
  int64_t m = 0, m_pad = 0;
  int n = 0, f = 0, nnz = 0, lpr = 8;

  DevBuf I, QQ, weights, wres, Q, Q0, stage;
  DevBuf rowptr, ent_col, ent_eid, ent_w2, keys, vals, cubtmp;
  DevBuf X, R, Z, P, AP, B, diag, dinv, S, R2, S2;
  // SELL-32-sigma copy of the pattern (ira_pcg.cuh)
  DevBuf sell_row, slice_off, slice_width, slice_cnt, sell_col, sell_eid, sell_w2;
  DevBuf pair_key, pair_key2, pair_w2, mate, pc1, pc2, npairs, mate2, pc3, att_key;
  // l1ra (ira_l1ra.cuh): per-edge / per-node primal-dual state, allocated on first use
  DevBuf pdU, pdAX, pdL1, pdL2, pdADX, pdDU, pdDL1, pdDL2, pdEV, pdSIGX, sell_w3;
  DevBuf pdX, pdATV, pdATDV, pdW1P, pdDX, diag3, dinv3, pdctl, pdtrial;
  PdCtl* h_pdctl = nullptr;  // pinned
  DevBuf mst_label, mst_label2, mst_order, mst_order2, mst_done, mst_ctl, mst_T0, mst_T1;   // init_mst (ira_mst.cuh)
  unsigned char* small_in = nullptr;    // mapped pinned blocks of the single-block window solver (ira_small.cuh)
  unsigned char* small_out = nullptr;
  bool small_ready = false;
  int start_mode = 0;        // 0: resident calls restart from the uploaded Q0, 1: continue from the current Q
  int nslices = 0, npos = 0;
  int64_t sell_total = 0;
  bool sell_built = false;   // the SELL copy exists (it may still be unused by irls: fmt_csr)
  bool fmt_csr = false;      // multi-kernel path on the CSR sub-warp kernels (lanes_per_row set)
  bool persistent = true;    // one cooperative kernel per linear solve
  bool pairing = false;      // 2x2 block-Jacobi active in the multi-kernel path
  int pcg_blocks_per_sm = 0;
  int res_blocks_per_sm = 0;
  // matrix-in-shared-memory PCG (ira_pcg2.cuh): cached entry columns per slice, entries per block, usable flag
  std::vector<int> h_slice_width;
  // two-level PCG for small / medium graphs (ira_coarse.cuh)
  DevBuf coarse_AC, coarse_RC;
  bool coarse_ok = false;
  int coarse_nc = 0, coarse_bsz = 0, coarse_smem = 0, coarse_tri_n = 0;
  int tri_bsz = 0;                   // > 0: chain-like graph, tridiagonal coarse operator with blocks of this many rows
  bool sell_index_order = false;     // SELL pattern in row order (the TRI kernel's blocks are lane groups)
  int cur_cost = -1;         // cost of the running irls call (the L1 family keeps the block-Jacobi kernels)
  DevBuf slice_map;          // balanced slice -> (block, warp) map of k_pcg_persistent_reg (PcgRegParams::slice_map)
  bool slice_map_ok = false;
  int pcg2_wcap = 0, pcg2_entries = 0;
  bool pcg2_ok = false;
  DevBuf ctl, partials, bad, flush;
  Ctl* h_ctl = nullptr;  // pinned
  Ctl* h_hist = nullptr; // pinned, IRA_STATS_MAX_ITERS records: per-iteration control blocks of a call that never tests the score

  // comm
  ncclComm_t comm = nullptr;
  // peer-memory solve (ira_peer.cuh): this rank's window, the peers' windows mapped through CUDA IPC
  bool peer = false;
  bool replicated = false;   // world_size > 1 but the graph is too small to partition: every rank solves all of it
  DevBuf peer_win, sell_pos, ipc_stage, sell_colpos;
  void* peer_mapped[kPeerMax] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void* peer_exported = nullptr;        // the window address the current mappings were exchanged for
  int peer_n = 0;
  unsigned long long peer_epoch = 0;
  int peer_blocks_per_sm = 0;

  // profiling
  struct Span { int cls; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  double prof_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int prof_c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int launches = 0;
  int prev_cg = 0;
  int pcg_kernel = 0;        // which linear-solve driver the last solve used (ira_stats.pcg_kernel)
};

// ---------------------------------------------------------------------------------------------
#define IRA_CUDA(h, expr)                                                                    \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      char buf__[512];                                                                        \
      snprintf(buf__, sizeof buf__, "%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,      \
               cudaGetErrorString(e__));                                                      \
      (h)->err = buf__;                                                                       \
      return IRA_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define IRA_TRY(expr)                             \
  do {                                            \
    ira_status s__ = (expr);                      \
    if (s__ != IRA_OK) return s__;                \
  } while (0)

namespace {

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Two CUDA events that are destroyed on every return path (the IRA_TRY / IRA_CUDA macros return early on errors).
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&a);
    return e != cudaSuccess ? e : cudaEventCreate(&b);
  }
  ~EventPair() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

cudaEvent_t prof_event(ira_context* h) {
  if (h->ev_used == h->ev_pool.size()) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return nullptr; }   // profiling is best effort
    h->ev_pool.push_back(e);
  }
  return h->ev_pool[h->ev_used++];
}
// Resolve the recorded event pairs into the per-class accumulators (synchronises the stream).
void prof_flush(ira_context* h) {
  if (h->spans.empty()) return;
  cudaStreamSynchronize(h->stream);
  for (auto& s : h->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, s.a, s.b);
    h->prof_t[s.cls] += ms; h->prof_c[s.cls] += 1;
  }
  h->spans.clear();
  h->ev_used = 0;
}
struct ProfScope {
  ira_context* h; int idx = -1;
  ProfScope(ira_context* h_, int cls) : h(h_) {
    if (h->opt.profile) {
      if (h->spans.size() >= 8192) prof_flush(h);
      ira_context::Span s{cls, prof_event(h), prof_event(h)};
      if (s.a && s.b) {
        cudaEventRecord(s.a, h->stream);
        h->spans.push_back(s);
        idx = (int)h->spans.size() - 1;
      }
    }
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(h->spans[idx].b, h->stream); }
};

void prof_collect(ira_context* h, ira_stats* st) {
  if (!h->opt.profile) return;
  prof_flush(h);
  double t[KC_N]; int c[KC_N];
  for (int k = 0; k < KC_N; ++k) { t[k] = h->prof_t[k]; c[k] = h->prof_c[k]; h->prof_t[k] = 0.0; h->prof_c[k] = 0; }
  if (st) {
    st->t_residual_ms = t[KC_RESIDUAL]; st->n_residual = c[KC_RESIDUAL];
    st->t_rhs_ms = t[KC_RHS]; st->n_rhs = c[KC_RHS];
    st->t_spmv_ms = t[KC_SPMV]; st->n_spmv = c[KC_SPMV];
    st->t_cgvec_ms = t[KC_CGVEC]; st->n_cgvec = c[KC_CGVEC];
    st->t_weights_ms = t[KC_WEIGHTS]; st->n_weights = c[KC_WEIGHTS];
    st->t_update_ms = t[KC_UPDATE]; st->n_update = c[KC_UPDATE];
    st->t_comm_ms = t[KC_COMM]; st->n_comm = c[KC_COMM];
    st->t_pcg_ms = t[KC_PCG]; st->n_pcg = c[KC_PCG];
  }
}

int grid_nodes(const ira_context* h, int n, int threads = 256) {
  return std::max(1, std::min(cdiv(n, threads), h->sms * 8));
}
int grid_rows(const ira_context* h, int n, int lpr) {
  const int64_t warps = cdiv(n, 32 / lpr);
  return std::max(1, std::min(cdiv(warps * 32, 256), h->sms * 8));
}

ira_status launch_check(ira_context* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string(what) + " launch failed: " + cudaGetErrorString(e);
    return IRA_ERR_CUDA;
  }
  h->launches++;
  return IRA_OK;
}

int pick_lpr(const ira_context* h) {
  int l = h->opt.lanes_per_row;
  if (l >= 2 && l <= 32 && (l & (l - 1)) == 0) return l;
  const double avg = h->n > 0 ? (double)h->nnz / h->n : 0.0;
  int v = 4;
  while (v < 32 && v * 3 < avg) v *= 2;
  return v;
}

// ---- kernel launch wrappers ------------------------------------------------------------------
ira_status run_residual(ira_context* h, const double4* Qsrc, int store_theta) {
  if (h->m == 0) return IRA_OK;
  ProfScope ps(h, KC_RESIDUAL);
  const int ntiles = (int)(h->m_pad / kResTile);
  // persistent CTAs: exactly as many as are resident at once (a larger grid runs its tail as a second, half-empty wave)
  if (h->res_blocks_per_sm == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_residual, kResTile, 0) != cudaSuccess || nb < 1) { cudaGetLastError(); nb = 4; }
    h->res_blocks_per_sm = nb;
  }
  const int grid = std::min(ntiles, h->sms * h->res_blocks_per_sm);
  k_residual<<<grid, kResTile, 0, h->stream>>>(h->I.as<int2>(), h->QQ.as<double>(), h->m_pad,
                                               h->weights.as<double>(), Qsrc, h->wres.as<double4>(),
                                               h->m, ntiles, store_theta);
  return launch_check(h, "k_residual");
}

template <int LPR>
ira_status run_rhs_t(ira_context* h) {
  k_rhs_diag<LPR><<<grid_rows(h, h->n, LPR), 256, 0, h->stream>>>(
      h->rowptr.as<int>(), h->ent_eid.as<int>(), h->wres.as<double4>(), h->ent_w2.as<double>(),
      h->B.as<double4>(), h->diag.as<double>(), h->n);
  return launch_check(h, "k_rhs_diag");
}
int grid_slices(const ira_context* h) {      // a multiple of the SM count: slices are dealt round-robin
  const int per_sm = std::max(1, std::min(8, cdiv(cdiv((int64_t)h->nslices * 32, 256), h->sms)));
  return h->sms * per_sm;
}

ira_status run_rhs(ira_context* h) {
  ProfScope ps(h, KC_RHS);
  if (!h->fmt_csr) {
    k_sell_rhs<<<grid_slices(h), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->slice_off.as<int>(),
                                                      h->slice_width.as<int>(), h->sell_eid.as<int>(),
                                                      h->wres.as<double4>(), h->sell_w2.as<double>(),
                                                      h->B.as<double4>(), h->diag.as<double>(), h->nslices);
    return launch_check(h, "k_sell_rhs");
  }
  switch (h->lpr) {
    case 2: return run_rhs_t<2>(h);
    case 4: return run_rhs_t<4>(h);
    case 8: return run_rhs_t<8>(h);
    case 16: return run_rhs_t<16>(h);
    default: return run_rhs_t<32>(h);
  }
}

template <int LPR>
ira_status run_spmv_t(ira_context* h, bool fuse) {
  const int grid = grid_rows(h, h->n, LPR);
  if (fuse)
    k_spmv<LPR, true><<<grid, 256, 0, h->stream>>>(h->rowptr.as<int>(), h->ent_col.as<int>(),
                                                   h->ent_w2.as<double>(), h->P.as<double4>(),
                                                   h->AP.as<double4>(), h->n, h->ctl.as<Ctl>(),
                                                   h->partials.as<double>());
  else
    k_spmv<LPR, false><<<grid, 256, 0, h->stream>>>(h->rowptr.as<int>(), h->ent_col.as<int>(),
                                                    h->ent_w2.as<double>(), h->P.as<double4>(),
                                                    h->AP.as<double4>(), h->n, h->ctl.as<Ctl>(),
                                                    h->partials.as<double>());
  return launch_check(h, "k_spmv");
}
ira_status run_spmv(ira_context* h, bool fuse) {
  ProfScope ps(h, KC_SPMV);
  if (!h->fmt_csr) {
#define IRA_SPMV_SELL(FUSE, V, U)                                                                          \
  k_spmv_sell<FUSE, V, U><<<grid_slices(h), 256, 0, h->stream>>>(                                          \
      h->sell_row.as<int>(), h->slice_off.as<int>(), h->slice_width.as<int>(), h->sell_col.as<int>(),      \
      h->sell_w2.as<double>(), h->P.as<double4>(), h->AP.as<double4>(), h->nslices, h->ctl.as<Ctl>(),      \
      h->partials.as<double>())
    const int var = h->opt.spmv_variant;
    if (fuse) {
      switch (var) {
        case 1: IRA_SPMV_SELL(true, 1, 4); break;
        case 2: IRA_SPMV_SELL(true, 2, 4); break;
        case 3: IRA_SPMV_SELL(true, 0, 8); break;
        case 4: IRA_SPMV_SELL(true, 1, 8); break;
        case 5: IRA_SPMV_SELL(true, 2, 8); break;
        default: IRA_SPMV_SELL(true, 0, 4); break;
      }
    } else {
      switch (var) {
        case 1: case 4: IRA_SPMV_SELL(false, 1, 4); break;
        default: IRA_SPMV_SELL(false, 0, 4); break;
      }
    }
#undef IRA_SPMV_SELL
    return launch_check(h, "k_spmv_sell");
  }
  switch (h->lpr) {
    case 2: return run_spmv_t<2>(h, fuse);
    case 4: return run_spmv_t<4>(h, fuse);
    case 8: return run_spmv_t<8>(h, fuse);
    case 16: return run_spmv_t<16>(h, fuse);
    default: return run_spmv_t<32>(h, fuse);
  }
}

ira_status allreduce(ira_context* h, const void* src, void* dst, size_t count, ncclDataType_t dt, ncclRedOp_t op) {
  if (h->opt.world_size <= 1) return IRA_OK;
  if (!h->comm) { h->err = "world_size > 1 but ira_comm_init was not called"; return IRA_ERR_COMM; }
  ProfScope ps(h, KC_COMM);
  ncclResult_t r = g_nccl.AllReduce(src, dst, count, dt, op, h->comm, h->stream);
  if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return IRA_ERR_COMM; }
  return IRA_OK;
}

ira_status allreduce_f64(ira_context* h, double* buf, size_t count) {
  if (h->opt.world_size <= 1) return IRA_OK;
  if (!h->comm) { h->err = "world_size > 1 but ira_comm_init was not called"; return IRA_ERR_COMM; }
  ProfScope ps(h, KC_COMM);
  ncclResult_t r = g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, h->comm, h->stream);
  if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return IRA_ERR_COMM; }
  return IRA_OK;
}

ira_status run_cg_init(ira_context* h) {
  ProfScope ps(h, KC_CGVEC);
  k_cg_init<<<grid_nodes(h, h->n, kRedThreads), kRedThreads, 0, h->stream>>>(
      h->B.as<double4>(), h->diag.as<double>(), h->dinv.as<double>(), h->X.as<double4>(),
      h->R.as<double4>(), h->Z.as<double4>(), h->P.as<double4>(), h->n, h->ctl.as<Ctl>(),
      h->partials.as<double>(), h->pairing ? h->mate.as<int>() : nullptr, h->pairing ? h->pc1.as<double>() : nullptr,
      h->pairing ? h->pc2.as<double>() : nullptr, h->mate2.as<int>(), h->pc3.as<double>());
  return launch_check(h, "k_cg_init");
}

ira_status run_cg_iteration(ira_context* h) {
  const bool sharded = h->opt.world_size > 1;
  IRA_TRY(run_spmv(h, !sharded));
  if (sharded) {
    IRA_TRY(allreduce_f64(h, h->AP.as<double>(), (size_t)h->n * 4));
    ProfScope ps(h, KC_CGVEC);
    k_cg_dot_pap<<<grid_nodes(h, h->n, kRedThreads), kRedThreads, 0, h->stream>>>(
        h->P.as<double4>(), h->AP.as<double4>(), h->n, h->ctl.as<Ctl>(), h->partials.as<double>());
    IRA_TRY(launch_check(h, "k_cg_dot_pap"));
  }
  {
    ProfScope ps(h, KC_CGVEC);
    k_cg_update<<<grid_nodes(h, h->n, kRedThreads), kRedThreads, 0, h->stream>>>(
        h->X.as<double4>(), h->R.as<double4>(), h->Z.as<double4>(), h->P.as<double4>(),
        h->AP.as<double4>(), h->dinv.as<double>(), h->n, h->ctl.as<Ctl>(), h->partials.as<double>(),
        h->pairing ? 1 : 0);
    IRA_TRY(launch_check(h, "k_cg_update"));
    if (h->pairing) {
      k_cg_precond<<<grid_nodes(h, h->n, kRedThreads), kRedThreads, 0, h->stream>>>(
          h->R.as<double4>(), h->Z.as<double4>(), h->mate.as<int>(), h->pc1.as<double>(), h->pc2.as<double>(), h->n,
          h->ctl.as<Ctl>(), h->partials.as<double>(), h->mate2.as<int>(), h->pc3.as<double>());
      IRA_TRY(launch_check(h, "k_cg_precond"));
    }
  }
  {
    ProfScope ps(h, KC_CGVEC);
    k_cg_p<<<grid_nodes(h, h->n), 256, 0, h->stream>>>(h->P.as<double4>(), h->Z.as<double4>(), h->n,
                                                       h->ctl.as<Ctl>());
    IRA_TRY(launch_check(h, "k_cg_p"));
  }
  return IRA_OK;
}

ira_status fetch_ctl(ira_context* h) {
  IRA_CUDA(h, cudaMemcpyAsync(h->h_ctl, h->ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

// One linear step as ONE cooperative kernel (ira_pcg.cuh); nothing is read back here - the
// iteration count and residual norms stay in the device control block until the IRLS iteration's
// single host synchronisation.
// Pairwise block-Jacobi set-up for the current weights (needs the complete diagonal).
ira_status run_pairing(ira_context* h) {
  ProfScope ps(h, KC_RHS);
  IRA_CUDA(h, cudaMemsetAsync(h->npairs.p, 0, sizeof(int), h->stream));
  k_pair_best<<<grid_slices(h), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->slice_off.as<int>(),
                                                     h->slice_width.as<int>(), h->sell_col.as<int>(),
                                                     h->sell_w2.as<double>(), h->diag.as<double>(), h->nslices,
                                                     h->opt.pair_theta, h->pair_key.as<unsigned long long>(),
                                                     h->pair_w2.as<double>());
  IRA_TRY(launch_check(h, "k_pair_best"));
  if (h->opt.world_size > 1 && !h->peer && !h->replicated) {   // a node's edges live on several ranks: global strongest pick
    IRA_TRY(allreduce(h, h->pair_key.p, h->pair_key2.p, (size_t)h->n, ncclUint64, ncclMax));
    k_pair_select<<<grid_nodes(h, h->n), 256, 0, h->stream>>>(h->pair_key.as<unsigned long long>(),
                                                           h->pair_key2.as<unsigned long long>(), h->pair_w2.as<double>(), h->n);
    IRA_TRY(launch_check(h, "k_pair_select"));
    IRA_CUDA(h, cudaMemcpyAsync(h->pair_key.p, h->pair_key2.p, sizeof(unsigned long long) * (size_t)h->n,
                                cudaMemcpyDeviceToDevice, h->stream));
    IRA_TRY(allreduce(h, h->pair_w2.p, h->pair_w2.p, (size_t)h->n, ncclFloat64, ncclMax));
  }
  k_pair_mate<<<grid_nodes(h, h->n), 256, 0, h->stream>>>(h->pair_key.as<unsigned long long>(), h->pair_w2.as<double>(),
                                                         h->diag.as<double>(), h->n, h->mate.as<int>(),
                                                         h->pc1.as<double>(), h->pc2.as<double>(), h->npairs.as<int>(),
                                                         h->mate2.as<int>(), h->pc3.as<double>());
  IRA_TRY(launch_check(h, "k_pair_mate"));
  // third members (3x3 blocks); the edge-sharded NCCL path would need two more all-reduces per solve: pairs only
  if (h->opt.pair_theta3 > 0.0 && (h->opt.world_size <= 1 || h->peer || h->replicated)) {
    IRA_CUDA(h, cudaMemsetAsync(h->att_key.p, 0, sizeof(unsigned long long) * (size_t)h->n, h->stream));
    k_attach_best<<<grid_slices(h), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->slice_off.as<int>(),
                                                       h->slice_width.as<int>(), h->sell_col.as<int>(),
                                                       h->sell_w2.as<double>(), h->diag.as<double>(), h->mate.as<int>(),
                                                       h->nslices, h->opt.pair_theta3, h->att_key.as<unsigned long long>(),
                                                       h->npairs.as<int>());
    IRA_TRY(launch_check(h, "k_attach_best"));
    k_attach_block<<<grid_nodes(h, h->n), 256, 0, h->stream>>>(h->att_key.as<unsigned long long>(), h->sell_pos.as<int>(),
                                                              h->slice_off.as<int>(), h->slice_width.as<int>(),
                                                              h->sell_col.as<int>(), h->sell_w2.as<double>(),
                                                              h->diag.as<double>(), h->pair_w2.as<double>(), h->n,
                                                              h->mate.as<int>(), h->mate2.as<int>(), h->pc1.as<double>(),
                                                              h->pc2.as<double>(), h->pc3.as<double>(), h->npairs.as<int>());
    IRA_TRY(launch_check(h, "k_attach_block"));
  }
  return IRA_OK;
}

ira_status l1ra_alloc(ira_context* h);
ira_status launch_pcg_coarse(ira_context* h, const double4* rhs, double4* xout);

ira_status solve_pcg_persistent(ira_context* h) {
  IRA_TRY(run_rhs(h));
  // small / medium graphs, bounded robust weights (everything but the L1 family, whose stiff pairs need the exact
  // 2x2 / 3x3 blocks): Jacobi + coarse space, the same weights for the three coordinates
  if (h->coarse_ok && !(h->opt.solver & 4) && h->opt.spmv_variant == 0 && h->cur_cost != (int)kL1 && h->cur_cost != (int)kL15 &&
      h->cur_cost != (int)kL05) {
    IRA_TRY(l1ra_alloc(h));
    ProfScope ps(h, KC_PCG);
    k_w2_to_w3<<<grid_nodes(h, (int)std::min<int64_t>(std::max<int64_t>(h->sell_total, h->n), 1 << 30)), 256, 0, h->stream>>>(
        h->sell_w2.as<double>(), h->diag.as<double>(), h->sell_total, h->n, h->sell_w3.as<double4>(), h->diag3.as<double4>());
    IRA_TRY(launch_check(h, "k_w2_to_w3"));
    return launch_pcg_coarse(h, h->B.as<double4>(), h->X.as<double4>());
  }
  const bool pairing = h->opt.pair_theta > 0.0;
  if (pairing) IRA_TRY(run_pairing(h));
  PcgParams pp;
  pp.mate = pairing ? h->mate.as<int>() : nullptr;
  pp.pc1 = pairing ? h->pc1.as<double>() : nullptr;
  pp.pc2 = pairing ? h->pc2.as<double>() : nullptr;
  pp.npairs = pairing ? h->npairs.as<int>() : nullptr;
  pp.mate2 = pairing ? h->mate2.as<int>() : nullptr;
  pp.pc3 = pairing ? h->pc3.as<double>() : nullptr;
  pp.n = h->n; pp.nslices = h->nslices; pp.max_iters = std::max(0, h->opt.cg_max_iters);
  pp.debug = h->opt.profile == 2;
  pp.rtol2 = h->opt.cg_rtol * h->opt.cg_rtol;
  pp.sell_row = h->sell_row.as<int>(); pp.slice_off = h->slice_off.as<int>(); pp.slice_width = h->slice_width.as<int>();
  pp.sell_col = h->sell_col.as<int>(); pp.sell_w2 = h->sell_w2.as<double>();
  pp.B = h->B.as<double4>(); pp.diag = h->diag.as<double>();
  pp.X = h->X.as<double4>(); pp.R = h->R.as<double4>(); pp.U = h->Z.as<double4>(); pp.W = h->AP.as<double4>();
  pp.P = h->P.as<double4>(); pp.S = h->S.as<double4>();
  pp.dinv = h->dinv.as<double>(); pp.partials = h->partials.as<double>(); pp.ctl = h->ctl.as<Ctl>();
  // one block per SM (slices are dealt round-robin over blocks); tiny graphs use fewer blocks so that
  // the grid barrier has fewer participants
  const int grid = std::max(1, std::min(h->nslices, h->sms * h->pcg_blocks_per_sm));
  ProfScope ps(h, KC_PCG);
  void* fn = nullptr;
  if (h->nslices <= h->sms * kMwGroups && !(h->opt.solver & 4) && h->opt.spmv_variant == 0) {
    // a few thousand nodes: several warps per slice (latency chain of the SpMV cut by kMwWarps), fewer blocks
    PcgRegParams pr;
    pr.base = pp;
    pr.RS0r = h->R.as<double4>(); pr.RS0s = h->S.as<double4>(); pr.RS1r = h->R2.as<double4>(); pr.RS1s = h->S2.as<double4>();
    pr.slice_map = nullptr;
    void* rargs[] = {(void*)&pr};
    const int g2 = std::max(1, cdiv(h->nslices, kMwGroups));
    IRA_CUDA(h, cudaLaunchCooperativeKernel((void*)k_pcg_persistent_reg_mw, dim3(g2), dim3(kPcgThreads), rargs, 0, h->stream));
    h->launches++;
    h->pcg_kernel = 3;
    return IRA_OK;
  }
  if (h->pcg2_ok && (h->opt.solver & 32) && !(h->opt.solver & 4) && h->opt.spmv_variant == 0) {   // opt-in: measured slower
    // one row per lane, state in registers, the matrix in shared memory (ira_pcg2.cuh)
    Pcg2Params p2;
    p2.reg.base = pp;
    p2.reg.RS0r = h->R.as<double4>(); p2.reg.RS0s = h->S.as<double4>(); p2.reg.RS1r = h->R2.as<double4>(); p2.reg.RS1s = h->S2.as<double4>();
    p2.reg.slice_map = nullptr;
    p2.wcap = h->pcg2_wcap; p2.smem_entries = h->pcg2_entries;
    void* rargs[] = {(void*)&p2};
    IRA_CUDA(h, cudaLaunchCooperativeKernel((void*)k_pcg_smem, dim3(std::min(h->nslices, h->sms)), dim3(kPcg2Threads), rargs,
                                            (size_t)h->pcg2_entries * 12 + 8 * kPcg2StateDoubles * kPcg2Threads, h->stream));
    h->launches++;
    h->pcg_kernel = 4;
    return IRA_OK;
  }
  if (h->nslices <= grid * (kPcgThreads / 32) && !(h->opt.solver & 4)) {   // one row per lane: state in registers
    PcgRegParams pr;
    pr.base = pp;
    pr.RS0r = h->R.as<double4>(); pr.RS0s = h->S.as<double4>(); pr.RS1r = h->R2.as<double4>(); pr.RS1s = h->S2.as<double4>();
    pr.slice_map = h->slice_map_ok ? h->slice_map.as<int>() : nullptr;
    void* rargs[] = {(void*)&pr};
    switch (h->opt.spmv_variant) {
      case 1: fn = (void*)k_pcg_persistent_reg<1, 4>; break;
      case 3: fn = (void*)k_pcg_persistent_reg<0, 8>; break;
      case 4: fn = (void*)k_pcg_persistent_reg<1, 8>; break;
      default: fn = (void*)k_pcg_persistent_reg<0, 4>; break;
    }
    IRA_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPcgThreads), rargs, 0, h->stream));
    h->launches++;
    h->pcg_kernel = 2;
    return IRA_OK;
  }
  void* args[] = {(void*)&pp};
  switch (h->opt.spmv_variant) {
    case 1: fn = (void*)k_pcg_persistent<1, 4>; break;
    case 3: fn = (void*)k_pcg_persistent<0, 8>; break;
    case 4: fn = (void*)k_pcg_persistent<1, 8>; break;
    default: fn = (void*)k_pcg_persistent<0, 4>; break;
  }
  IRA_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kPcgThreads), args, 0, h->stream));
  h->launches++;
  h->pcg_kernel = 1;
  return IRA_OK;
}

// Map every rank's window into this process (CUDA IPC handles all-gathered over the NCCL communicator).
ira_status peer_setup(ira_context* h) {
  const int G = h->opt.world_size, n = std::max(h->n, 1);
  if (G > kPeerMax) { h->err = "peer-memory solve supports at most 8 ranks"; return IRA_ERR_INVALID_ARG; }
  if (G > 1 && !h->comm) { h->err = "world_size > 1 but ira_comm_init was not called"; return IRA_ERR_COMM; }
  void* before = h->peer_win.p;
  IRA_CUDA(h, h->peer_win.reserve(std::max(peer_window_bytes(n), peer_window_ll_bytes(std::max(h->npos, 1)))));
  IRA_CUDA(h, h->sell_colpos.reserve(sizeof(int) * (size_t)std::max<int64_t>(h->sell_total, 1)));
  // every rank takes the same decision: all see the same n, and a window only ever grows
  if (G == 1) {                                            // the barrier-free kernel on one GPU: no mapping to exchange
    if (h->peer_win.p != before || h->peer_exported != h->peer_win.p) {
      IRA_CUDA(h, cudaMemsetAsync(h->peer_win.p, 0, h->peer_win.cap, h->stream));
      h->peer_exported = h->peer_win.p;
      h->peer_epoch = 0;
    }
  } else if (h->peer_win.p != before || h->peer_exported != h->peer_win.p) {
    for (int g = 0; g < kPeerMax; ++g)
      if (h->peer_mapped[g]) { cudaIpcCloseMemHandle(h->peer_mapped[g]); h->peer_mapped[g] = nullptr; }
    IRA_CUDA(h, cudaMemsetAsync(h->peer_win.p, 0, h->peer_win.cap, h->stream));
    cudaIpcMemHandle_t mine;
    IRA_CUDA(h, cudaIpcGetMemHandle(&mine, h->peer_win.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    IRA_CUDA(h, h->ipc_stage.reserve(64 * (size_t)(G + 1)));
    unsigned char* st = h->ipc_stage.as<unsigned char>();
    IRA_CUDA(h, cudaMemcpyAsync(st, &mine, 64, cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = g_nccl.AllGather(st, st + 64, 64, ncclUint8, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return IRA_ERR_COMM; }
    std::vector<cudaIpcMemHandle_t> all((size_t)G);
    IRA_CUDA(h, cudaMemcpyAsync(all.data(), st + 64, 64 * (size_t)G, cudaMemcpyDeviceToHost, h->stream));
    IRA_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int g = 0; g < G; ++g) {
      if (g == h->opt.rank) continue;
      cudaError_t e = cudaIpcOpenMemHandle(&h->peer_mapped[g], all[(size_t)g], cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        h->err = std::string("cudaIpcOpenMemHandle failed (peer-memory solve needs CUDA IPC between the ranks): ") +
                 cudaGetErrorString(e);
        return IRA_ERR_COMM;
      }
    }
    h->peer_exported = h->peer_win.p;
    h->peer_epoch = 0;
    // nobody may start a solve (and write into a peer's window) before every rank has zeroed its flags
    IRA_CUDA(h, cudaMemsetAsync(st, 0, 8, h->stream));
    r = g_nccl.AllReduce(st, st, 1, ncclFloat64, ncclSum, h->comm, h->stream);
    if (r != ncclSuccess) { h->err = std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r); return IRA_ERR_COMM; }
    IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->peer_n = n;
  k_sell_colpos<<<std::max(1, std::min(cdiv(std::max<int64_t>(h->sell_total, 1), 256), h->sms * 8)), 256, 0, h->stream>>>(
      h->sell_col.as<int>(), h->sell_pos.as<int>(), h->sell_total, h->sell_colpos.as<int>());
  IRA_TRY(launch_check(h, "k_sell_colpos"));
  if (h->peer_blocks_per_sm == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pcg_peer<0, 4>, kPeerThreads, 0) != cudaSuccess || nb < 1) {
      cudaGetLastError();
      h->err = "k_pcg_peer cannot be launched cooperatively on this device";
      return IRA_ERR_CUDA;
    }
    h->peer_blocks_per_sm = nb;
  }
  return IRA_OK;
}

// One linear step, rows partitioned over the ranks, exchanged through peer memory (ira_peer.cuh).
ira_status solve_pcg_peer(ira_context* h) {
  IRA_TRY(run_rhs(h));                                   // replicated: every rank holds the whole graph
  const bool pairing = h->opt.pair_theta > 0.0;
  if (pairing) IRA_TRY(run_pairing(h));
  PcgPeerParams q;
  memset(&q, 0, sizeof q);
  PcgParams& pp = q.base;
  pp.mate = pairing ? h->mate.as<int>() : nullptr;
  pp.pc1 = pairing ? h->pc1.as<double>() : nullptr;
  pp.pc2 = pairing ? h->pc2.as<double>() : nullptr;
  pp.npairs = pairing ? h->npairs.as<int>() : nullptr;
  pp.mate2 = pairing ? h->mate2.as<int>() : nullptr;
  pp.pc3 = pairing ? h->pc3.as<double>() : nullptr;
  pp.n = h->peer_n; pp.nslices = h->nslices; pp.max_iters = std::max(0, h->opt.cg_max_iters);
  pp.rtol2 = h->opt.cg_rtol * h->opt.cg_rtol;
  pp.sell_row = h->sell_row.as<int>(); pp.slice_off = h->slice_off.as<int>(); pp.slice_width = h->slice_width.as<int>();
  pp.sell_col = h->sell_col.as<int>(); pp.sell_w2 = h->sell_w2.as<double>();
  pp.B = h->B.as<double4>(); pp.diag = h->diag.as<double>();
  pp.X = nullptr; pp.R = h->R.as<double4>(); pp.U = nullptr; pp.W = h->AP.as<double4>();
  pp.P = h->P.as<double4>(); pp.S = h->S.as<double4>();
  pp.dinv = h->dinv.as<double>(); pp.partials = h->partials.as<double>(); pp.ctl = h->ctl.as<Ctl>();
  const int G = h->opt.world_size;
  q.world = G; q.rank = h->opt.rank;
  for (int g = 0; g <= kPeerMax; ++g) q.slice_bound[g] = (int)(((int64_t)h->nslices * std::min(g, G)) / G);
  q.slice_lo = q.slice_bound[q.rank]; q.slice_hi = q.slice_bound[q.rank + 1];
  q.sell_pos = h->sell_pos.as<int>();
  for (int g = 0; g < G; ++g) q.win[g] = (unsigned char*)(g == q.rank ? h->peer_win.p : h->peer_mapped[g]);
  q.epoch_base = h->peer_epoch;
  q.debug = (h->opt.spmv_variant & 7) == 7;
  { const int fv = (h->opt.spmv_variant >> 3) & 3; q.flush = fv == 0 ? (G > 1 ? 1 : 0) : fv - 1; }   // default: one fence per warp
  q.sell_colpos = h->sell_colpos.as<int>();
  q.npos = h->npos;
  const int own = std::max(1, q.slice_hi - q.slice_lo);
  int grid = std::max(1, std::min(own, h->sms * h->peer_blocks_per_sm));
  ProfScope ps(h, KC_PCG);
  void* args[] = {(void*)&q};
  const bool ll = h->opt.shard_mode == 1 || G == 1;      // 2 = the barrier version (A/B measurements)
  void* fn = ll ? (void*)k_pcg_peer_ll : (h->opt.spmv_variant == 1 ? (void*)k_pcg_peer<1, 4> : (void*)k_pcg_peer<0, 4>);
  int threads = kPeerThreads;
  // one row per lane fits (on EVERY rank: the largest slice range decides, so all ranks take the same kernel)
  int max_own = 0;
  for (int g = 0; g < G; ++g) max_own = std::max(max_own, q.slice_bound[g + 1] - q.slice_bound[g]);
  if (ll && !(h->opt.solver & 4) && h->pcg_blocks_per_sm >= 1 && max_own <= h->sms * (kPcgThreads / 32)) {
    fn = (void*)k_pcg_peer_ll_reg;
    threads = kPcgThreads;
    grid = std::max(1, std::min(h->sms, cdiv(max_own, kPcgThreads / 32)));
    // slices are dealt round-robin over the blocks: slice = lo + block + grid * warp must cover [lo, hi)
    grid = std::max(grid, std::min(h->sms, max_own));
    if ((int64_t)grid * (kPcgThreads / 32) < max_own) { fn = (void*)k_pcg_peer_ll; threads = kPeerThreads; grid = std::max(1, std::min(own, h->sms * h->peer_blocks_per_sm)); }
  }
  IRA_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(threads), args, 0, h->stream));
  h->launches++;
  h->pcg_kernel = 5;
  const double4* xsrc = ll ? peer_window_ll_at((unsigned char*)h->peer_win.p, h->npos).X
                           : peer_window_at((unsigned char*)h->peer_win.p, h->peer_n).X;
  IRA_CUDA(h, cudaMemcpyAsync(h->X.p, xsrc, sizeof(double4) * (size_t)h->n, cudaMemcpyDeviceToDevice, h->stream));
  return IRA_OK;
}

// One linear step: Jacobi-PCG on A^T D^2 A X = A^T D^2 w, x0 = 0 (replaces ls_solve, :536-556).
ira_status solve_pcg(ira_context* h, int* iters_out, double* relres_out, int* hit_max) {
  IRA_TRY(run_rhs(h));
  IRA_TRY(allreduce_f64(h, h->B.as<double>(), (size_t)h->n * 4));
  IRA_TRY(allreduce_f64(h, h->diag.as<double>(), (size_t)h->n));
  h->pairing = h->opt.pair_theta > 0.0 && !h->fmt_csr;
  if (h->pairing) IRA_TRY(run_pairing(h));
  IRA_TRY(run_cg_init(h));
  const int cap = std::max(0, h->opt.cg_max_iters);
  const int every = std::max(1, h->opt.cg_check_every);
  int issued = 0;
  bool first = true;
  for (;;) {
    int chunk = every;
    if (first && h->prev_cg > every) chunk = h->prev_cg - every / 2;   // solves of consecutive IRLS
    first = false;                                                     // iterations need similar counts
    chunk = std::min(chunk, cap - issued);
    for (int c = 0; c < chunk; ++c) IRA_TRY(run_cg_iteration(h));
    issued += chunk;
    IRA_TRY(fetch_ctl(h));
    if (h->h_ctl->done || issued >= cap) break;
  }
  const Ctl& c = *h->h_ctl;
  double rel = 0.0;
  bool conv = true;
  for (int k = 0; k < 3; ++k) {
    if (c.bnorm2[k] > 0.0) rel = std::max(rel, sqrt(c.rnorm2[k] / c.bnorm2[k]));
    if (!(c.rnorm2[k] <= c.rtol2 * c.bnorm2[k])) conv = false;
  }
  *iters_out = c.cg_iters;
  *relres_out = rel;
  *hit_max = conv ? 0 : 1;
  h->prev_cg = c.cg_iters;
  return IRA_OK;
}

ira_status alloc_problem(ira_context* h, int64_t m, int n) {
  const int64_t m_pad = std::max<int64_t>(kResTile, ((m + kResTile - 1) / kResTile) * kResTile);
  h->m = m; h->m_pad = m_pad; h->n = n;
  IRA_CUDA(h, h->I.reserve(sizeof(int2) * m_pad));
  IRA_CUDA(h, h->QQ.reserve(sizeof(double) * 4 * m_pad));
  IRA_CUDA(h, h->weights.reserve(sizeof(double) * m_pad));
  IRA_CUDA(h, h->wres.reserve(sizeof(double4) * m_pad));
  IRA_CUDA(h, h->Q.reserve(sizeof(double4) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->Q0.reserve(sizeof(double4) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->stage.reserve(sizeof(double) * 4 * (size_t)std::max<int64_t>(std::max<int64_t>(n, m), 1)));
  IRA_CUDA(h, h->rowptr.reserve(sizeof(int) * ((size_t)n + 1)));
  IRA_CUDA(h, h->ent_col.reserve(sizeof(int) * (size_t)std::max<int64_t>(2 * m, 1)));
  IRA_CUDA(h, h->ent_eid.reserve(sizeof(int) * (size_t)std::max<int64_t>(2 * m, 1)));
  IRA_CUDA(h, h->ent_w2.reserve(sizeof(double) * (size_t)std::max<int64_t>(2 * m, 1)));
  IRA_CUDA(h, h->keys.reserve(sizeof(int) * (size_t)std::max<int64_t>(4 * m, 1)));
  IRA_CUDA(h, h->vals.reserve(sizeof(int) * (size_t)std::max<int64_t>(4 * m, 1)));
  for (DevBuf* b : {&h->X, &h->R, &h->Z, &h->P, &h->AP, &h->B, &h->S, &h->R2, &h->S2})
    IRA_CUDA(h, b->reserve(sizeof(double4) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->diag.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->dinv.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pair_key.reserve(sizeof(unsigned long long) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pair_w2.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pair_key2.reserve(sizeof(unsigned long long) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->mate.reserve(sizeof(int) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pc1.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pc2.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->npairs.reserve(sizeof(int)));
  IRA_CUDA(h, h->mate2.reserve(sizeof(int) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->pc3.reserve(sizeof(double) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->att_key.reserve(sizeof(unsigned long long) * (size_t)std::max(n, 1)));
  IRA_CUDA(h, h->sell_pos.reserve(sizeof(int) * (size_t)std::max(n, 1)));
  return IRA_OK;
}

ira_status build_csr(ira_context* h) {
  const int64_t m = h->m;
  const int n = h->n;
  IRA_CUDA(h, cudaMemsetAsync(h->bad.p, 0, sizeof(int), h->stream));
  int* keys_in = h->keys.as<int>();
  int* keys_out = keys_in + 2 * m;
  int* vals_in = h->vals.as<int>();
  int* vals_out = vals_in + 2 * m;
  if (m > 0) {
    k_csr_keys<<<std::min(cdiv(m, 256), h->sms * 16), 256, 0, h->stream>>>(h->I.as<int2>(), m, n, h->f,
                                                                         keys_in, vals_in, h->bad.as<int>());
    IRA_TRY(launch_check(h, "k_csr_keys"));
    int end_bit = 1;
    while ((1ll << end_bit) <= (int64_t)n) ++end_bit;      // keys are in [0, n]
    size_t tmp_bytes = 0;
    IRA_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                (int)(2 * m), 0, end_bit, h->stream));
    IRA_CUDA(h, h->cubtmp.reserve(tmp_bytes));
    IRA_CUDA(h, cub::DeviceRadixSort::SortPairs(h->cubtmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out,
                                                (int)(2 * m), 0, end_bit, h->stream));
    h->launches += 4;
  }
  k_csr_finalize<<<std::max(1, std::min(cdiv(2 * m + 1, 256), h->sms * 16)), 256, 0, h->stream>>>(
      keys_out, vals_out, h->I.as<int2>(), 2 * m, n, h->f, h->rowptr.as<int>(), h->ent_col.as<int>(),
      h->ent_eid.as<int>());
  IRA_TRY(launch_check(h, "k_csr_finalize"));
  int host[2] = {0, 0};
  IRA_CUDA(h, cudaMemcpyAsync(&host[0], h->rowptr.as<int>() + n, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaMemcpyAsync(&host[1], h->bad.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  if (host[1] >= 2) { h->err = "edge endpoint out of range [0, n_total)"; return IRA_ERR_INVALID_ARG; }
  // the reference's make_A turns an edge (i, i) into a single -1 entry (its second coeffRef overwrites the first,
  // ral/l1_irls.cpp:772-776): not a relative rotation between two views - refused rather than silently reinterpreted
  if (host[1] == 1) { h->err = "self-loop edge (i == j) is not a relative rotation between two views"; return IRA_ERR_INVALID_ARG; }
  h->nnz = host[0];
  h->lpr = pick_lpr(h);
  return IRA_OK;
}

// SELL-32-sigma copy of the CSR pattern (ira_pcg.cuh): window sort by degree, slice offsets, fill.
ira_status build_sell(ira_context* h) {
  const int n = h->n;
  const int nwin = std::max(1, cdiv(n, kSellSigma));
  h->npos = nwin * kSellSigma;
  h->nslices = h->npos / kSellC;
  IRA_CUDA(h, h->sell_row.reserve(sizeof(int) * (size_t)h->npos));
  IRA_CUDA(h, h->slice_width.reserve(sizeof(int) * ((size_t)h->nslices + 1)));
  IRA_CUDA(h, h->slice_cnt.reserve(sizeof(int) * ((size_t)h->nslices + 1)));
  IRA_CUDA(h, h->slice_off.reserve(sizeof(int) * ((size_t)h->nslices + 1)));
  IRA_CUDA(h, cudaMemsetAsync(h->slice_cnt.p, 0, sizeof(int) * ((size_t)h->nslices + 1), h->stream));
  k_sell_sort<<<nwin, kSellSigma, 0, h->stream>>>(h->rowptr.as<int>(), n, h->sell_row.as<int>(),
                                                 h->slice_width.as<int>(), h->slice_cnt.as<int>(), h->sell_index_order ? 1 : 0);
  IRA_TRY(launch_check(h, "k_sell_sort"));
  size_t tmp_bytes = 0;
  IRA_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, h->slice_cnt.as<int>(), h->slice_off.as<int>(),
                                            h->nslices + 1, h->stream));
  IRA_CUDA(h, h->cubtmp.reserve(tmp_bytes));
  IRA_CUDA(h, cub::DeviceScan::ExclusiveSum(h->cubtmp.p, tmp_bytes, h->slice_cnt.as<int>(), h->slice_off.as<int>(),
                                            h->nslices + 1, h->stream));
  h->launches += 2;
  int total = 0;
  IRA_CUDA(h, cudaMemcpyAsync(&total, h->slice_off.as<int>() + h->nslices, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  h->sell_total = total;
  const size_t cap = (size_t)std::max(total, 1);
  IRA_CUDA(h, h->sell_col.reserve(sizeof(int) * cap));
  IRA_CUDA(h, h->sell_eid.reserve(sizeof(int) * cap));
  IRA_CUDA(h, h->sell_w2.reserve(sizeof(double) * cap));
  k_sell_fill<<<cdiv(h->npos, 256), 256, 0, h->stream>>>(h->rowptr.as<int>(), h->ent_col.as<int>(), h->ent_eid.as<int>(),
                                                        h->sell_row.as<int>(), h->slice_off.as<int>(),
                                                        h->slice_width.as<int>(), h->npos, h->sell_col.as<int>(),
                                                        h->sell_eid.as<int>(), h->sell_w2.as<double>());
  IRA_TRY(launch_check(h, "k_sell_fill"));
  k_sell_inverse<<<cdiv(h->npos, 256), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->npos, h->sell_pos.as<int>());
  return launch_check(h, "k_sell_inverse");
}

// Balanced deal of the SELL slices to the blocks of the register-resident PCG kernel (one slice per warp, kPcgThreads / 32
// warps per block, one block per SM): longest-processing-time first - slices by decreasing width, each to the block
// that holds the fewest entries so far and still has a free warp.
ira_status plan_slice_map(ira_context* h) {
  h->slice_map_ok = false;
  const int wpb = kPcgThreads / 32;
  const int grid = std::max(1, std::min(h->nslices, h->sms * std::max(1, h->pcg_blocks_per_sm)));
  if (h->pcg_blocks_per_sm <= 0 || h->nslices <= 0 || (int64_t)grid * wpb < h->nslices || h->nslices <= h->sms * kMwGroups) return IRA_OK;
  h->h_slice_width.resize((size_t)h->nslices);
  IRA_CUDA(h, cudaMemcpyAsync(h->h_slice_width.data(), h->slice_width.p, sizeof(int) * (size_t)h->nslices,
                              cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  std::vector<int> map;
  plan::lpt_slice_map(h->h_slice_width.data(), h->nslices, grid, wpb, &map);
  IRA_CUDA(h, h->slice_map.reserve(sizeof(int) * map.size()));
  IRA_CUDA(h, cudaMemcpyAsync(h->slice_map.p, map.data(), sizeof(int) * map.size(), cudaMemcpyHostToDevice, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));      // `map` is a local
  h->slice_map_ok = true;
  return IRA_OK;
}

// Two-level PCG (ira_coarse.cuh) for graphs of up to kCoarseMaxRows nodes: partition of the node indices into
// nc <= 64 contiguous blocks; shared memory for the three nc x nc inverses.
// Host-only decision, before the SELL pattern is built: is this a chain-like graph for the tridiagonal coarse
// operator (ira_coarse.cuh, TRI)?  Blocks of 8, 16 or 32 consecutive rows (lane groups of a warp once the SELL pattern
// is in index order), at most kTriMax of them, and at most 2 % of the edges may reach beyond the adjacent block (those
// are lumped onto the diagonal of the coarse operator).  +256: dense 64-block variant only (A/B).
void plan_coarse_tri(ira_context* h, const int32_t* I_pairs) {
  h->tri_bsz = 0;
  if ((h->opt.solver & (128 | 256)) || h->opt.lanes_per_row >= 2 || h->pcg_blocks_per_sm <= 0) return;
  h->tri_bsz = plan::tri_block_rows(h->n, h->f, h->m, I_pairs);
}

ira_status plan_coarse(ira_context* h, const int32_t* I_pairs) {
  h->coarse_ok = false;
  const int n = h->n, nfree = h->n - h->f;
  if (h->pcg_blocks_per_sm <= 0 || h->nslices <= 0 || n > kCoarseMaxRows || nfree < 128) return IRA_OK;
  // Tridiagonal coarse operator (ira_coarse.cuh, TRI): decided before the SELL build (plan_coarse_tri), resources here.
  h->coarse_tri_n = 0;
  if (h->tri_bsz > 0 && h->sell_index_order) {
    const int bsz = h->tri_bsz, nc = cdiv(n, bsz);
    int N = 1;
    while (N < nc) N <<= 1;
    const int bytes = (int)sizeof(double) * 18 * N;
    int nb = 0;
    if (N <= kTriMax && cudaFuncSetAttribute(k_pcg_coarse_w3<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pcg_coarse_w3<true>, kCoarseThreads, (size_t)bytes) == cudaSuccess && nb >= 1) {
      IRA_CUDA(h, h->coarse_AC.reserve(sizeof(double) * 6 * (size_t)N));
      IRA_CUDA(h, h->coarse_RC.reserve(sizeof(double4) * (size_t)nc));
      h->coarse_nc = nc; h->coarse_bsz = bsz; h->coarse_smem = bytes; h->coarse_tri_n = N;
      h->coarse_ok = true;
      return IRA_OK;
    }
    cudaGetLastError();
  }
  int bsz = 0, nc = 0;
  if (!plan::dense_partition(n, h->f, h->m, I_pairs, &bsz, &nc)) return IRA_OK;   // chain-likeness test: ira_plan.hpp
  const int bytes = (int)sizeof(double) * (((3 * nc * nc + 3) & ~3) + 8 * nc);
  if (cudaFuncSetAttribute(k_pcg_coarse_w3<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess) { cudaGetLastError(); return IRA_OK; }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pcg_coarse_w3<false>, kCoarseThreads, (size_t)bytes) != cudaSuccess || nb < 1) {
    cudaGetLastError();
    return IRA_OK;
  }
  IRA_CUDA(h, h->coarse_AC.reserve(sizeof(double) * 3 * (size_t)nc * nc));
  IRA_CUDA(h, h->coarse_RC.reserve(sizeof(double4) * (size_t)nc));
  h->coarse_nc = nc; h->coarse_bsz = bsz; h->coarse_smem = bytes;
  h->coarse_ok = true;
  return IRA_OK;
}

// One linear solve with three weight sets by the two-level kernel: (sell_w3, diag3) x = rhs.
ira_status launch_pcg_coarse(ira_context* h, const double4* rhs, double4* xout) {
  PcgCoarseParams q;
  PcgW3Params& pp = q.w;
  pp.n = h->n; pp.nslices = h->nslices; pp.max_iters = std::max(0, h->opt.cg_max_iters);
  pp.rtol2 = h->opt.cg_rtol * h->opt.cg_rtol;
  pp.sell_row = h->sell_row.as<int>(); pp.slice_off = h->slice_off.as<int>(); pp.slice_width = h->slice_width.as<int>();
  pp.sell_col = h->sell_col.as<int>(); pp.sell_w3 = h->sell_w3.as<double4>(); pp.diag3 = h->diag3.as<double4>();
  pp.B = rhs;
  pp.X = xout; pp.R = h->R.as<double4>(); pp.U = h->Z.as<double4>(); pp.W = h->AP.as<double4>();
  pp.P = h->P.as<double4>(); pp.S = h->S.as<double4>(); pp.DINV = h->dinv3.as<double4>();
  pp.partials = h->partials.as<double>(); pp.ctl = h->ctl.as<Ctl>();
  q.sell_pos = h->sell_pos.as<int>();
  q.f = h->f; q.nc = h->coarse_nc; q.bsz = h->coarse_bsz;
  q.AC = h->coarse_AC.as<double>(); q.RC = h->coarse_RC.as<double4>();
  q.tri_n = h->coarse_tri_n;
  // one slice per warp, 12 warps per block: as few blocks as hold the slices (cheap grid barriers)
  const int grid = std::max(1, cdiv(h->nslices, kCoarseThreads / 32));
  if (grid > h->sms) { h->err = "two-level PCG: more slices than resident warps"; return IRA_ERR_INVALID_ARG; }
  void* args[] = {(void*)&q};
  IRA_CUDA(h, cudaLaunchCooperativeKernel(h->coarse_tri_n ? (void*)k_pcg_coarse_w3<true> : (void*)k_pcg_coarse_w3<false>, dim3(grid),
                                          dim3(kCoarseThreads), args, (size_t)h->coarse_smem, h->stream));
  h->launches++;
  h->pcg_kernel = h->coarse_tri_n ? 8 : 7;
  return IRA_OK;
}

// Plan of the matrix-in-shared-memory PCG kernel (ira_pcg2.cuh): one block per SM, slice s -> block s % grid,
// warp s / grid.  Picks the largest number of entry columns per slice (`wcap`, a multiple of 4) whose (col, w2)
// pairs fit every block's shared memory; slices wider than that read their tail from global memory.
ira_status plan_pcg2(ira_context* h) {
  h->pcg2_ok = false;
  if (h->pcg_blocks_per_sm <= 0 || h->nslices <= 0) return IRA_OK;
  const int grid = std::min(h->nslices, h->sms);
  if ((int64_t)grid * kPcg2Warps < h->nslices) return IRA_OK;          // more than one row per lane: k_pcg_persistent
  h->h_slice_width.resize((size_t)h->nslices);
  IRA_CUDA(h, cudaMemcpyAsync(h->h_slice_width.data(), h->slice_width.p, sizeof(int) * (size_t)h->nslices,
                              cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  int wmax = 0;
  for (int w : h->h_slice_width) wmax = std::max(wmax, w);
  int optin = 0;
  cudaFuncAttributes fa;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device) != cudaSuccess ||
      cudaFuncGetAttributes(&fa, k_pcg_smem) != cudaSuccess) { cudaGetLastError(); return IRA_OK; }
  const int state_bytes = 8 * kPcg2StateDoubles * kPcg2Threads;
  const int cap = (optin - (int)fa.sharedSizeBytes - state_bytes - 256) / 12;   // entries of (int col, double w2) beside x, p
  if (cap < 4 * kSellC) return IRA_OK;
  int wcap = wmax, need = 0;
  for (;; wcap -= 4) {
    need = 0;
    for (int b = 0; b < grid; ++b) {
      int e = 0;
      for (int s = b; s < h->nslices; s += grid) e += std::min(h->h_slice_width[(size_t)s], wcap) * kSellC;
      need = std::max(need, e);
    }
    if (need <= cap || wcap <= 0) break;
  }
  if (wcap < 4 && wmax >= 4) return IRA_OK;                            // nothing fits: not worth it
  h->pcg2_wcap = std::max(wcap, 0);
  h->pcg2_entries = std::max(need, kSellC);
  const int bytes = h->pcg2_entries * 12 + state_bytes;
  const bool dbg = getenv("IRA_DEBUG") != nullptr;
  cudaError_t ce = cudaFuncSetAttribute(k_pcg_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (ce != cudaSuccess) {
    if (dbg) fprintf(stderr, "[ira] k_pcg_smem: %d B of dynamic shared memory refused: %s\n", bytes, cudaGetErrorString(ce));
    cudaGetLastError();
    return IRA_OK;
  }
  int nb = 0;
  ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pcg_smem, kPcg2Threads, (size_t)bytes);
  if (ce != cudaSuccess || nb < 1) {
    if (dbg) fprintf(stderr, "[ira] k_pcg_smem: occupancy %d with %d B (%s)\n", nb, bytes, cudaGetErrorString(ce));
    cudaGetLastError();
    return IRA_OK;
  }
  if (dbg) fprintf(stderr, "[ira] k_pcg_smem planned: wcap %d, %d entries, %d B per block\n", h->pcg2_wcap, h->pcg2_entries, bytes);
  h->pcg2_ok = true;
  return IRA_OK;
}

ira_status upload_Q(ira_context* h, const double* Q, int64_t ld_q, DevBuf& dst) {
  const int n = h->n;
  if (n == 0) return IRA_OK;
  IRA_CUDA(h, cudaMemcpy2DAsync(h->stage.p, sizeof(double) * n, Q, sizeof(double) * ld_q, sizeof(double) * n,
                                4, cudaMemcpyHostToDevice, h->stream));
  k_colmajor_to_aos4<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->stage.as<double>(), n, dst.as<double4>(), n);
  return launch_check(h, "k_colmajor_to_aos4");
}

ira_status check_args(ira_context* h, int64_t m, int64_t n_total, int32_t f, const void* I, const void* QQ,
                      int64_t ld_qq, const void* Q, int64_t ld_q) {
  if (!h) return IRA_ERR_INVALID_ARG;
  auto bad = [&](const char* why) { h->err = why; return IRA_ERR_INVALID_ARG; };
  if (m < 0 || n_total < 0) return bad("negative size");
  if (m > (1ll << 30) - 1 || n_total > std::numeric_limits<int>::max() - 2) return bad("problem too large for int32 indices");
  if (f < 1) return bad("f < 1: at least one rotation must be fixed (ral/l1_irls.cpp:917)");
  if (f > n_total) return bad("f > n_total");
  if (m > 0 && (!I || !QQ)) return bad("null edge arrays");
  if (n_total > 0 && !Q) return bad("null Q");
  if (ld_qq < m || ld_q < n_total) return bad("leading dimension smaller than row count");
  return IRA_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int32_t ira_abi_version(void) { return IRA_ABI_VERSION; }

int32_t ira_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

const char* ira_status_string(ira_status s) {
  switch (s) {
    case IRA_OK: return "ok";
    case IRA_ERR_INVALID_ARG: return "invalid argument";
    case IRA_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU path)";
    case IRA_ERR_CUDA: return "CUDA error";
    case IRA_ERR_COMM: return "NCCL / communicator error";
    case IRA_ERR_UNKNOWN_COST: return "Unknown cost!!";
    case IRA_ERR_NOT_UPLOADED: return "no problem uploaded";
    case IRA_ERR_NONFINITE: return "non-finite score";
    case IRA_ERR_NOT_SPANNING: return "Relative rotations DO NOT SPAN all the nodes in the VIEW GRAPH";
  }
  return "unknown status";
}

ira_status ira_get_stream(ira_handle h, void** stream_out) {
  if (!h || !stream_out) return IRA_ERR_INVALID_ARG;
  *stream_out = (void*)h->stream;
  return IRA_OK;
}

const char* ira_last_error(ira_handle h) { return h ? h->err.c_str() : "null handle"; }

ira_status ira_options_default(ira_options* o) {
  if (!o) return IRA_ERR_INVALID_ARG;
  memset(o, 0, sizeof *o);
  o->device = -1;
  o->cg_max_iters = 20000;
  o->cg_rtol = 1e-10;
  o->pair_theta = 0.5;       // measured optimum on configs[2], profiles/r02_sweep_theta_{1,2,3,broad}.json
  o->pair_theta3 = 0.001;
  o->cg_check_every = 16;
  o->lanes_per_row = 0;
  o->world_size = 1;
  o->rank = 0;
  o->profile = 0;
  o->peer_min_rows = 120000;
  // A/B aid for callers that take the defaults (the C++ adapters, the CLI): IRA_SOLVER=<bits> presets `solver`
  if (const char* e = getenv("IRA_SOLVER")) o->solver = atoi(e);
  return IRA_OK;
}

ira_status ira_create(ira_handle* out, const ira_options* opt) {
  if (!out) return IRA_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return IRA_ERR_NO_DEVICE; }
  ira_context* h = new ira_context();
  if (opt) h->opt = *opt; else ira_options_default(&h->opt);
  if (h->opt.world_size < 1) h->opt.world_size = 1;
  int dev = h->opt.device;
  if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
  if (dev >= ndev) { delete h; return IRA_ERR_NO_DEVICE; }
  cudaDeviceProp prop;
  if (cudaSetDevice(dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) {
    cudaGetLastError();
    delete h;
    return IRA_ERR_NO_DEVICE;   // kernels are built for sm_100a only
  }
  h->device = dev;
  h->sms = prop.multiProcessorCount;
  bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&h->h_ctl, sizeof(Ctl)) == cudaSuccess;
  ok = ok && cudaMallocHost((void**)&h->h_hist, sizeof(Ctl) * IRA_STATS_MAX_ITERS) == cudaSuccess;
  ok = ok && h->ctl.reserve(sizeof(Ctl)) == cudaSuccess;
  ok = ok && h->partials.reserve(sizeof(double) * 8 * kRedMaxBlocks) == cudaSuccess;
  ok = ok && h->bad.reserve(sizeof(int)) == cudaSuccess;
  int coop = 0, nb = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_pcg_persistent<0, 8>, kPcgThreads, 0) == cudaSuccess)
    h->pcg_blocks_per_sm = nb;
  if (!ok) { cudaGetLastError(); ira_destroy(h); return IRA_ERR_CUDA; }
  *out = h;
  return IRA_OK;
}

ira_status ira_destroy(ira_handle h) {
  if (!h) return IRA_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int g = 0; g < kPeerMax; ++g)
    if (h->peer_mapped[g]) { cudaIpcCloseMemHandle(h->peer_mapped[g]); h->peer_mapped[g] = nullptr; }
  h->peer_win.release();
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  for (DevBuf* b : {&h->I, &h->QQ, &h->weights, &h->wres, &h->Q, &h->Q0, &h->stage, &h->rowptr, &h->ent_col,
                    &h->ent_eid, &h->ent_w2, &h->keys, &h->vals, &h->cubtmp, &h->X, &h->R, &h->Z, &h->P,
                    &h->AP, &h->B, &h->diag, &h->dinv, &h->ctl, &h->partials, &h->bad, &h->flush, &h->S, &h->sell_row,
                    &h->slice_off, &h->slice_width, &h->slice_cnt, &h->sell_col, &h->sell_eid, &h->sell_w2,
                    &h->R2, &h->S2, &h->pdU, &h->pdAX, &h->pdL1, &h->pdL2, &h->pdADX, &h->pdDU, &h->pdDL1, &h->pdDL2, &h->pdEV,
                    &h->pdSIGX, &h->sell_w3, &h->pdX, &h->pdATV, &h->pdATDV, &h->pdW1P, &h->pdDX, &h->diag3, &h->dinv3,
                    &h->pdctl, &h->pdtrial, &h->mst_label, &h->mst_label2, &h->mst_order, &h->mst_order2, &h->mst_done,
                    &h->mst_ctl, &h->mst_T0, &h->mst_T1, &h->slice_map, &h->coarse_AC, &h->coarse_RC, &h->sell_pos, &h->ipc_stage, &h->sell_colpos, &h->pair_key, &h->pair_key2, &h->pair_w2, &h->mate, &h->pc1, &h->pc2, &h->npairs, &h->mate2, &h->pc3, &h->att_key})
    b->release();
  for (auto e : h->ev_pool) cudaEventDestroy(e);
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_hist) cudaFreeHost(h->h_hist);
  if (h->h_pdctl) cudaFreeHost(h->h_pdctl);
  if (h->small_in) cudaFreeHost(h->small_in);
  if (h->small_out) cudaFreeHost(h->small_out);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return IRA_OK;
}

ira_status ira_problem_upload(ira_handle h, int64_t m, int64_t n_total, int32_t f, const int32_t* I_pairs,
                              const double* QQ, int64_t ld_qq, const double* Q0, int64_t ld_q) {
  IRA_TRY(check_args(h, m, n_total, f, I_pairs, QQ, ld_qq, Q0, ld_q));
  IRA_CUDA(h, cudaSetDevice(h->device));
  h->uploaded = false;
  h->f = f;
  IRA_TRY(alloc_problem(h, m, (int)n_total));
  const int64_t mp = h->m_pad;
  // padded tails are zero so that whole 256-edge tiles can always be bulk-copied
  IRA_CUDA(h, cudaMemsetAsync(h->I.p, 0, sizeof(int2) * mp, h->stream));
  IRA_CUDA(h, cudaMemsetAsync(h->QQ.p, 0, sizeof(double) * 4 * mp, h->stream));
  if (m > 0) {
    IRA_CUDA(h, cudaMemcpyAsync(h->I.p, I_pairs, sizeof(int2) * m, cudaMemcpyHostToDevice, h->stream));
    IRA_CUDA(h, cudaMemcpy2DAsync(h->QQ.p, sizeof(double) * mp, QQ, sizeof(double) * ld_qq, sizeof(double) * m, 4,
                                  cudaMemcpyHostToDevice, h->stream));
  }
  IRA_TRY(upload_Q(h, Q0, ld_q, h->Q0));
  IRA_TRY(build_csr(h));
  // SELL padding blow-up (heavy-tailed degrees) falls back to the CSR sub-warp kernels
  h->fmt_csr = h->opt.lanes_per_row >= 2;
  h->sell_built = false;
  h->sell_index_order = false;
  h->tri_bsz = 0;
  if (!h->fmt_csr) {
    const bool alone = h->opt.world_size <= 1 ||
                       ((h->opt.shard_mode == 1 || h->opt.shard_mode == 2) && n_total < (int64_t)h->opt.peer_min_rows);
    if (alone) plan_coarse_tri(h, I_pairs);
    h->sell_index_order = h->tri_bsz > 0;
    IRA_TRY(build_sell(h));
    h->sell_built = true;
    if (h->sell_total > 3ll * std::max(h->nnz, 1) + 64ll * kSellSigma) {
      if (h->sell_index_order) {                     // index order pads too much after all: degree-sorted windows
        h->sell_index_order = false;
        h->tri_bsz = 0;
        IRA_TRY(build_sell(h));
      }
      if (h->sell_total > 3ll * std::max(h->nnz, 1) + 64ll * kSellSigma) h->fmt_csr = true;
    }
  }
  // whole graph on every rank and too small to be worth partitioning: replicated single-GPU solves
  h->replicated = h->opt.world_size > 1 && (h->opt.shard_mode == 1 || h->opt.shard_mode == 2) &&
                  n_total < (int64_t)h->opt.peer_min_rows && !h->fmt_csr && h->pcg_blocks_per_sm > 0;
  const bool one_gpu = h->opt.world_size <= 1 || h->replicated;
  h->pcg2_ok = false;
  h->slice_map_ok = false;
  if (!h->fmt_csr && one_gpu && (h->opt.solver & 32)) IRA_TRY(plan_pcg2(h));
  if (!h->fmt_csr && one_gpu && !(h->opt.solver & 64)) IRA_TRY(plan_slice_map(h));   // +64: round-robin deal (A/B)
  h->coarse_ok = false;
  if (!h->fmt_csr && one_gpu && !(h->opt.solver & 128)) IRA_TRY(plan_coarse(h, I_pairs));   // +128: one-level kernels only (A/B)
  h->persistent = !h->fmt_csr && one_gpu && ((h->opt.solver & 3) != 1 || h->replicated) && h->pcg_blocks_per_sm > 0;
  h->peer = false;
  if ((h->opt.world_size > 1 && !h->replicated && (h->opt.shard_mode == 1 || h->opt.shard_mode == 2)) ||
      (h->opt.world_size <= 1 && (h->opt.solver & 8))) {
    if (h->fmt_csr || h->pcg_blocks_per_sm <= 0) { h->err = "peer-memory solve needs the SELL pattern and cooperative launch"; return IRA_ERR_INVALID_ARG; }
    IRA_TRY(peer_setup(h));
    h->peer = true;
  }
  h->prev_cg = 0;
  h->uploaded = true;
  h->start_mode = 0;
  if (h->n > 0) IRA_CUDA(h, cudaMemcpyAsync(h->Q.p, h->Q0.p, sizeof(double4) * h->n, cudaMemcpyDeviceToDevice, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

ira_status ira_irls_resident(ira_handle h, int32_t cost, double sigma, int32_t max_iters, double change_th,
                             int32_t* iters_out, double* runtime_s_out, ira_stats* stats) {
  if (!h) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  if (cost < 0 || cost >= kNumCosts) { h->err = "Unknown cost!!"; return IRA_ERR_UNKNOWN_COST; }
  IRA_CUDA(h, cudaSetDevice(h->device));
  const auto t0 = std::chrono::steady_clock::now();
  const int launches0 = h->launches;
  if (stats) {
    const double up = stats->t_upload_ms;
    memset(stats, 0, sizeof *stats);
    stats->t_upload_ms = up;
  }
  EventPair evp;
  IRA_CUDA(h, evp.create());
  cudaEvent_t ev0 = evp.a, ev1 = evp.b;
  IRA_CUDA(h, cudaEventRecord(ev0, h->stream));

  const int n = h->n;
  const int64_t m = h->m;
  if (n > 0 && h->start_mode == 0)
    IRA_CUDA(h, cudaMemcpyAsync(h->Q.p, h->Q0.p, sizeof(double4) * n, cudaMemcpyDeviceToDevice, h->stream));
  k_fill_f64<<<std::max(1, std::min(cdiv(h->m_pad, 256), h->sms * 8)), 256, 0, h->stream>>>(
      h->weights.as<double>(), 1.0, h->m_pad);                          // weights.setOnes() (:577)
  IRA_TRY(launch_check(h, "k_fill_f64"));
  Ctl init;
  memset(&init, 0, sizeof init);
  init.rtol2 = h->opt.cg_rtol * h->opt.cg_rtol;
  init.cg_max_iters = h->opt.cg_max_iters;
  *h->h_ctl = init;
  IRA_CUDA(h, cudaMemcpyAsync(h->ctl.p, h->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  h->prev_cg = 0;

  double score = std::numeric_limits<double>::max();                      // :574
  int iters = 0;
  ira_status rc = IRA_OK;
  h->cur_cost = cost;
  const bool deferred = change_th < 0.0 && h->persistent && !h->peer && max_iters <= IRA_STATS_MAX_ITERS;
  // one iteration's record -> stats (cg_it / rel / hit come from the control block on the persistent paths)
  auto record = [&](const Ctl& c, int k, int cg_it, double rel, int hit) {
    if (h->persistent || h->peer) {
      cg_it = c.cg_iters;
      bool conv = true;
      rel = 0.0;
      for (int q = 0; q < 3; ++q) {
        if (c.bnorm2[q] > 0.0) rel = std::max(rel, sqrt(c.rnorm2[q] / c.bnorm2[q]));
        if (!(c.rnorm2[q] <= c.rtol2 * c.bnorm2[q])) conv = false;
      }
      hit = conv ? 0 : 1;
    }
    if (stats && k < IRA_STATS_MAX_ITERS) {
      stats->score[k] = c.score;
      stats->cg_iters[k] = cg_it;
      stats->cg_relres[k] = rel;
    }
    if (stats) { stats->cg_iters_total += cg_it; stats->cg_hit_max += hit; }
  };
  while (score > change_th && iters < max_iters) {                         // :590
    IRA_TRY(run_residual(h, h->Q.as<double4>(), 0));                       // :592-593
    int cg_it = 0, hit = 0;
    double rel = 0.0;
    if (h->peer) IRA_TRY(solve_pcg_peer(h));
    else if (h->persistent) IRA_TRY(solve_pcg_persistent(h));              // :596-612
    else IRA_TRY(solve_pcg(h, &cg_it, &rel, &hit));
    if (m > 0) {
      ProfScope ps(h, KC_WEIGHTS);
      k_weights<<<std::min(cdiv(m, 256), h->sms * 16), 256, 0, h->stream>>>(
          h->I.as<int2>(), h->wres.as<double4>(), h->X.as<double4>(), h->weights.as<double>(), m, h->f,
          cost, sigma);                                                    // :614-727
      IRA_TRY(launch_check(h, "k_weights"));
    }
    {
      ProfScope ps(h, KC_UPDATE);
      k_update<<<grid_nodes(h, std::max(1, n - h->f), kRedThreads), kRedThreads, 0, h->stream>>>(
          h->Q.as<double4>(), h->X.as<double4>(), n, h->f, h->ctl.as<Ctl>(), h->partials.as<double>());  // :729-737
      IRA_TRY(launch_check(h, "k_update"));
    }
    if (deferred) {
      // change_th < 0: the loop test (:590) can only end the loop on a NaN score, so nothing has to come back per
      // iteration: the control block is copied out asynchronously and read after the last iteration
      IRA_CUDA(h, cudaMemcpyAsync(&h->h_hist[iters], h->ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
      iters++;                                                             // :739
      continue;
    }
    IRA_TRY(fetch_ctl(h));
    if (h->peer) h->peer_epoch = h->h_ctl->epoch;
    record(*h->h_ctl, iters, cg_it, rel, hit);
    score = h->h_ctl->score;
    iters++;                                                               // :739
    if (!std::isfinite(score)) { if (n - h->f > 0) rc = IRA_ERR_NONFINITE; break; }
  }
  IRA_CUDA(h, cudaEventRecord(ev1, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  if (deferred && iters > 0) {
    for (int k = 0; k < iters; ++k) {
      record(h->h_hist[k], k, 0, 0.0, 0);
      if (!std::isfinite(h->h_hist[k].score)) {          // the reference's loop ends here (NaN > th is false)
        if (n - h->f > 0) rc = IRA_ERR_NONFINITE;
        iters = k + 1;
        break;
      }
    }
    *h->h_ctl = h->h_hist[iters - 1];
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  prof_collect(h, stats);
  if (iters_out) *iters_out = iters;
  if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) {
    stats->irls_iters = iters;
    stats->pcg_kernel = (h->persistent || h->peer) ? h->pcg_kernel : 0;
    stats->t_total_ms = ms;
    stats->kernel_launches = h->launches - launches0;
    const Ctl& c = *h->h_ctl;
    if ((h->persistent || h->peer) && c.cyc_total > 0) {          // block 0's phase clocks -> milliseconds
      const double ms_per_cycle = 1e-6 * (double)c.ns_total / (double)c.cyc_total;
      stats->pcg_spmv_ms = ms_per_cycle * (double)c.cyc_spmv;
      stats->pcg_update_ms = ms_per_cycle * (double)c.cyc_update;
      stats->pcg_kernel_ms = 1e-6 * (double)c.ns_total;
      stats->pcg_spmv_phases = (int32_t)std::min<long long>(c.pcg_spmv_phases, 2147483647ll);
    }
  }
  if (rc == IRA_ERR_NONFINITE) h->err = "score became non-finite";
  return rc;
}

ira_status ira_problem_download(ira_handle h, double* Q, int64_t ld_q, double* weights) {
  if (!h) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  IRA_CUDA(h, cudaSetDevice(h->device));
  const int n = h->n;
  if (Q && n > 0) {
    if (ld_q < n) { h->err = "ld_q < n_total"; return IRA_ERR_INVALID_ARG; }
    k_aos4_to_colmajor<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->Q.as<double4>(), h->stage.as<double>(), n, n, 4);
    IRA_TRY(launch_check(h, "k_aos4_to_colmajor"));
    IRA_CUDA(h, cudaMemcpy2DAsync(Q, sizeof(double) * ld_q, h->stage.p, sizeof(double) * n, sizeof(double) * n, 4,
                                  cudaMemcpyDeviceToHost, h->stream));
  }
  if (weights && h->m > 0)
    IRA_CUDA(h, cudaMemcpyAsync(weights, h->weights.p, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

ira_status ira_irls(ira_handle h, int64_t m, int64_t n_total, int32_t f, const int32_t* I_pairs, const double* QQ,
                    int64_t ld_qq, double* Q, int64_t ld_q, int32_t cost, double sigma, int32_t max_iters,
                    double change_th, double* weights, int32_t* iters_out, double* runtime_s_out,
                    ira_stats* stats) {
  const auto t0 = std::chrono::steady_clock::now();
  IRA_TRY(check_args(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  if (m > 0 && !weights) { h->err = "null weights"; return IRA_ERR_INVALID_ARG; }
  if (cost < 0 || cost >= kNumCosts) { h->err = "Unknown cost!!"; return IRA_ERR_UNKNOWN_COST; }
  IRA_TRY(ira_problem_upload(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  const auto t1 = std::chrono::steady_clock::now();
  if (stats) stats->t_upload_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  const ira_status rc = ira_irls_resident(h, cost, sigma, max_iters, change_th, iters_out, nullptr, stats);
  if (rc != IRA_OK && rc != IRA_ERR_NONFINITE) return rc;
  const auto t2 = std::chrono::steady_clock::now();
  IRA_TRY(ira_problem_download(h, Q, ld_q, weights));
  const auto t3 = std::chrono::steady_clock::now();
  if (stats) stats->t_download_ms = std::chrono::duration<double, std::milli>(t3 - t2).count();
  if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(t3 - t0).count();
  return rc;
}

// ---- host helpers ---------------------------------------------------------------------------
ira_status ira_make_A(int64_t m, int32_t n_total, int32_t f, const int32_t* I_pairs, int32_t* col_plus,
                      int32_t* col_minus) {
  if (m < 0 || n_total < 0 || f < 0 || (m > 0 && (!I_pairs || !col_plus || !col_minus))) return IRA_ERR_INVALID_ARG;
  for (int64_t k = 0; k < m; ++k) {
    const int32_t i = I_pairs[2 * k], j = I_pairs[2 * k + 1];
    if (i < 0 || j < 0 || i >= n_total || j >= n_total) return IRA_ERR_INVALID_ARG;
    col_plus[k] = j >= f ? j - f : -1;                       // ral/l1_irls.cpp:770-772
    col_minus[k] = (j >= f && i >= f) ? i - f : -1;          // :774-776 (skipped when j is fixed)
  }
  return IRA_OK;
}

ira_status ira_quat_normalised(double* Q, int64_t n_total, int64_t ld_q, int32_t f) {
  if (!Q || ld_q < n_total || f < 0) return IRA_ERR_INVALID_ARG;
  for (int64_t i = f; i < n_total; ++i) {                    // ral/l1_irls.cpp:985-990
    double* x = Q + i; double* y = Q + ld_q + i; double* z = Q + 2 * ld_q + i; double* w = Q + 3 * ld_q + i;
    const double nrm = sqrt(*x * *x + *y * *y + *z * *z + *w * *w);
    if (nrm > 0.0) { *x /= nrm; *y /= nrm; *z /= nrm; *w /= nrm; }   // Eigen leaves a zero quaternion as is
  }
  return IRA_OK;
}

// ---- probes ---------------------------------------------------------------------------------
ira_status ira_probe_residual(ira_handle h, double* w_out) {
  if (!h || !w_out) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  IRA_CUDA(h, cudaSetDevice(h->device));
  if (h->m == 0) return IRA_OK;
  IRA_CUDA(h, cudaMemsetAsync(h->weights.p, 0, sizeof(double) * h->m_pad, h->stream));
  IRA_TRY(run_residual(h, h->Q0.as<double4>(), 1));
  k_aos4_to_colmajor<<<std::min(cdiv(h->m, 256), h->sms * 8), 256, 0, h->stream>>>(
      h->wres.as<double4>(), h->stage.as<double>(), h->m, h->m, 4);
  IRA_TRY(launch_check(h, "k_aos4_to_colmajor"));
  IRA_CUDA(h, cudaMemcpyAsync(w_out, h->stage.p, sizeof(double) * 4 * h->m, cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

static ira_status probe_prepare(ira_handle h) {   // weights = 1, residual from Q0, rhs, cg_init: a live CG state
  k_fill_f64<<<std::max(1, std::min(cdiv(h->m_pad, 256), h->sms * 8)), 256, 0, h->stream>>>(h->weights.as<double>(), 1.0, h->m_pad);
  IRA_TRY(launch_check(h, "k_fill_f64"));
  Ctl init;
  memset(&init, 0, sizeof init);
  init.rtol2 = 0.0;
  init.cg_max_iters = 1 << 30;
  *h->h_ctl = init;
  IRA_CUDA(h, cudaMemcpyAsync(h->ctl.p, h->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
  IRA_CUDA(h, cudaMemcpyAsync(h->Q.p, h->Q0.p, sizeof(double4) * h->n, cudaMemcpyDeviceToDevice, h->stream));
  IRA_TRY(run_residual(h, h->Q.as<double4>(), 0));
  IRA_TRY(run_rhs(h));
  h->pairing = false;
  IRA_TRY(run_cg_init(h));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

ira_status ira_probe_laplacian_apply(ira_handle h, const double* weights, const double* X, double* Y) {
  if (!h || !weights || !X || !Y) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  IRA_CUDA(h, cudaSetDevice(h->device));
  const int n = h->n, nf = n - h->f;
  if (nf <= 0) return IRA_OK;
  IRA_TRY(probe_prepare(h));
  if (h->m > 0) {
    IRA_CUDA(h, cudaMemcpyAsync(h->weights.p, weights, sizeof(double) * h->m, cudaMemcpyHostToDevice, h->stream));
    k_set_w2<<<std::min(cdiv(h->m, 256), h->sms * 8), 256, 0, h->stream>>>(h->wres.as<double4>(), h->weights.as<double>(), h->m);
    IRA_TRY(launch_check(h, "k_set_w2"));
  }
  IRA_TRY(run_rhs(h));
  IRA_CUDA(h, cudaMemcpyAsync(h->stage.p, X, sizeof(double) * 3 * nf, cudaMemcpyHostToDevice, h->stream));
  k_free_to_aos4<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->stage.as<double>(), nf, h->f, h->P.as<double4>(), n);
  IRA_TRY(launch_check(h, "k_free_to_aos4"));
  IRA_TRY(run_spmv(h, h->opt.world_size <= 1));
  k_aos4_to_colmajor<<<grid_nodes(h, nf), 256, 0, h->stream>>>(h->AP.as<double4>() + h->f, h->stage.as<double>(), nf, nf, 3);
  IRA_TRY(launch_check(h, "k_aos4_to_colmajor"));
  IRA_CUDA(h, cudaMemcpyAsync(Y, h->stage.p, sizeof(double) * 3 * nf, cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

ira_status ira_probe_time_kernel(ira_handle h, int32_t which, int32_t reps, int32_t flush_l2, double* avg_us_out) {
  if (!h || !avg_us_out || reps < 1 || which < 0 || which > 5) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  IRA_CUDA(h, cudaSetDevice(h->device));
  IRA_TRY(probe_prepare(h));
  const size_t flush_bytes = 512ull << 20;
  if (flush_l2) IRA_CUDA(h, h->flush.reserve(flush_bytes));
  cudaEvent_t a, b;
  IRA_CUDA(h, cudaEventCreate(&a));
  IRA_CUDA(h, cudaEventCreate(&b));
  const int saved_profile = h->opt.profile;
  h->opt.profile = 0;
  double total_ms = 0.0;
  ira_status rc = IRA_OK;
  auto one = [&]() -> ira_status {
    switch (which) {
      case 0: return run_residual(h, h->Q.as<double4>(), 0);
      case 1: return run_spmv(h, h->opt.world_size <= 1);
      case 2: return run_rhs(h);
      case 3:
        k_weights<<<std::min(cdiv(std::max<int64_t>(h->m, 1), 256), h->sms * 16), 256, 0, h->stream>>>(
            h->I.as<int2>(), h->wres.as<double4>(), h->X.as<double4>(), h->weights.as<double>(), h->m, h->f,
            (int)kL1, 0.0872664626);
        return launch_check(h, "k_weights");
      case 4:
        k_update<<<grid_nodes(h, std::max(1, h->n - h->f), kRedThreads), kRedThreads, 0, h->stream>>>(
            h->Q.as<double4>(), h->X.as<double4>(), h->n, h->f, h->ctl.as<Ctl>(), h->partials.as<double>());
        return launch_check(h, "k_update");
      default:
        k_cg_update<<<grid_nodes(h, h->n, kRedThreads), kRedThreads, 0, h->stream>>>(
            h->X.as<double4>(), h->R.as<double4>(), h->Z.as<double4>(), h->P.as<double4>(), h->AP.as<double4>(),
            h->dinv.as<double>(), h->n, h->ctl.as<Ctl>(), h->partials.as<double>(), 0);
        return launch_check(h, "k_cg_update");
    }
  };
  for (int w = 0; w < 3 && rc == IRA_OK; ++w) rc = one();   // warm-up
  if (rc == IRA_OK) {
    if (flush_l2) {
      for (int r = 0; r < reps && rc == IRA_OK; ++r) {
        cudaMemsetAsync(h->flush.p, r & 0xff, flush_bytes, h->stream);
        cudaEventRecord(a, h->stream);
        rc = one();
        cudaEventRecord(b, h->stream);
        cudaStreamSynchronize(h->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        total_ms += ms;
      }
    } else {
      cudaEventRecord(a, h->stream);
      for (int r = 0; r < reps && rc == IRA_OK; ++r) rc = one();
      cudaEventRecord(b, h->stream);
      cudaStreamSynchronize(h->stream);
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      total_ms = ms;
    }
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  h->opt.profile = saved_profile;
  if (rc != IRA_OK) return rc;
  IRA_CUDA(h, cudaGetLastError());
  *avg_us_out = total_ms * 1000.0 / reps;
  return IRA_OK;
}

// ---- communicator ---------------------------------------------------------------------------
ira_status ira_comm_unique_id(uint8_t id_out[128]) {
  if (!id_out) return IRA_ERR_INVALID_ARG;
  if (!g_nccl.load()) return IRA_ERR_COMM;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return IRA_ERR_COMM;
  memcpy(id_out, &id, 128);
  return IRA_OK;
}

ira_status ira_comm_init(ira_handle h, const uint8_t id_in[128]) {
  if (!h || !id_in) return IRA_ERR_INVALID_ARG;
  if (h->opt.world_size <= 1) return IRA_OK;
  if (h->opt.rank < 0 || h->opt.rank >= h->opt.world_size) { h->err = "rank out of range"; return IRA_ERR_INVALID_ARG; }
  if (!g_nccl.load()) { h->err = "libnccl.so.2 could not be loaded"; return IRA_ERR_COMM; }
  IRA_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId id;
  memcpy(&id, id_in, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, h->opt.world_size, id, h->opt.rank);
  if (r != ncclSuccess) { h->err = std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r); return IRA_ERR_COMM; }
  return IRA_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// irotavg::l1ra  (ral/l1_irls.hpp:98-101, ral/l1_irls.cpp:851-912) on the device, ira_l1ra.cuh
// ---------------------------------------------------------------------------------------------
namespace {

ira_status l1ra_alloc(ira_context* h) {
  const size_t e = sizeof(double4) * (size_t)h->m_pad, nn = sizeof(double4) * (size_t)std::max(h->n, 1);
  for (DevBuf* b : {&h->pdU, &h->pdAX, &h->pdL1, &h->pdL2, &h->pdADX, &h->pdDU, &h->pdDL1, &h->pdDL2, &h->pdEV, &h->pdSIGX})
    IRA_CUDA(h, b->reserve(e));
  for (DevBuf* b : {&h->pdX, &h->pdATV, &h->pdATDV, &h->pdW1P, &h->pdDX, &h->diag3, &h->dinv3})
    IRA_CUDA(h, b->reserve(nn));
  IRA_CUDA(h, h->sell_w3.reserve(sizeof(double4) * (size_t)std::max<int64_t>(h->sell_total, 1)));
  IRA_CUDA(h, h->pdctl.reserve(sizeof(PdCtl)));
  IRA_CUDA(h, h->pdtrial.reserve(sizeof(double) * 8));
  if (!h->h_pdctl) IRA_CUDA(h, cudaMallocHost((void**)&h->h_pdctl, sizeof(PdCtl)));
  return IRA_OK;
}

ira_status fetch_pdctl(ira_context* h) {
  IRA_CUDA(h, cudaMemcpyAsync(h->h_pdctl, h->pdctl.p, sizeof(PdCtl), cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  return IRA_OK;
}

int grid_edges(const ira_context* h) { return std::max(1, std::min(cdiv(std::max<int64_t>(h->m, 1), 256), h->sms * 8)); }

ira_status pd_At(ira_context* h, DevBuf& out, int want_norm) {
  k_pd_At<<<grid_slices(h), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->slice_off.as<int>(), h->slice_width.as<int>(),
                                                 h->sell_eid.as<int>(), h->pdEV.as<double4>(), out.as<double4>(), h->nslices,
                                                 want_norm, h->pdctl.as<PdCtl>(), h->partials.as<double>());
  return launch_check(h, "k_pd_At");
}

// One call of l1decode_pd for the three coordinates (pdmaxiter Newton steps); result in pdX.
ira_status l1decode_pd_device(ira_context* h, int pdmaxiter, int* newton_cg_iters, int* hit_max) {
  const int64_t m = h->m;
  const int n = h->n;
  const double4* Y = h->wres.as<double4>();
  PdCtl* ctl = h->pdctl.as<PdCtl>();
  double* part = h->partials.as<double>();
  PdCtl init;
  memset(&init, 0, sizeof init);
  init.m2 = 2.0 * (double)m;
  *h->h_pdctl = init;
  IRA_CUDA(h, cudaMemcpyAsync(h->pdctl.p, h->h_pdctl, sizeof(PdCtl), cudaMemcpyHostToDevice, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));                      // h_pdctl is reused below
  IRA_CUDA(h, cudaMemsetAsync(h->pdX.p, 0, sizeof(double4) * (size_t)std::max(n, 1), h->stream));
  const int ge = grid_edges(h);
  k_pd_absmax<<<ge, 256, 0, h->stream>>>(Y, m, ctl, part);
  IRA_TRY(launch_check(h, "k_pd_absmax"));
  k_pd_init<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                       h->pdL2.as<double4>(), h->pdEV.as<double4>(), m, ctl, part);
  IRA_TRY(launch_check(h, "k_pd_init"));
  IRA_TRY(pd_At(h, h->pdATV, 1));
  k_pd_rcent<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                        h->pdL2.as<double4>(), m, 0, pdmaxiter, ctl, part);
  IRA_TRY(launch_check(h, "k_pd_rcent"));
  for (int pd = 0; pd < pdmaxiter; ++pd) {
    k_pd_prep<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                         h->pdL2.as<double4>(), h->pdSIGX.as<double4>(), h->pdEV.as<double4>(), m, ctl);
    IRA_TRY(launch_check(h, "k_pd_prep"));
    IRA_TRY(pd_At(h, h->pdW1P, 0));
    k_sell_hweights<<<grid_slices(h), 256, 0, h->stream>>>(h->sell_row.as<int>(), h->slice_off.as<int>(),
                                                           h->slice_width.as<int>(), h->sell_eid.as<int>(),
                                                           h->pdSIGX.as<double4>(), h->sell_w3.as<double4>(),
                                                           h->diag3.as<double4>(), h->nslices);
    IRA_TRY(launch_check(h, "k_sell_hweights"));
    if (h->coarse_ok) {                                               // H dx = w1p  (:319), two-level PCG
      IRA_TRY(launch_pcg_coarse(h, h->pdW1P.as<double4>(), h->pdDX.as<double4>()));
    } else {                                                          // H dx = w1p  (:319)
      PcgW3Params pp;
      pp.n = n; pp.nslices = h->nslices; pp.max_iters = std::max(0, h->opt.cg_max_iters);
      pp.rtol2 = h->opt.cg_rtol * h->opt.cg_rtol;
      pp.sell_row = h->sell_row.as<int>(); pp.slice_off = h->slice_off.as<int>(); pp.slice_width = h->slice_width.as<int>();
      pp.sell_col = h->sell_col.as<int>(); pp.sell_w3 = h->sell_w3.as<double4>(); pp.diag3 = h->diag3.as<double4>();
      pp.B = h->pdW1P.as<double4>();
      pp.X = h->pdDX.as<double4>(); pp.R = h->R.as<double4>(); pp.U = h->Z.as<double4>(); pp.W = h->AP.as<double4>();
      pp.P = h->P.as<double4>(); pp.S = h->S.as<double4>(); pp.DINV = h->dinv3.as<double4>();
      pp.partials = h->partials.as<double>(); pp.ctl = h->ctl.as<Ctl>();
      const int grid = std::max(1, std::min(h->nslices, h->sms * h->pcg_blocks_per_sm));
      void* args[] = {(void*)&pp};
      IRA_CUDA(h, cudaLaunchCooperativeKernel((void*)k_pcg_persistent_w3, dim3(grid), dim3(kPcgThreads), args, 0, h->stream));
      h->launches++;
    }
    k_pd_direction<<<ge, 256, 0, h->stream>>>(h->I.as<int2>(), h->f, h->pdDX.as<double4>(), Y, h->pdU.as<double4>(),
                                              h->pdAX.as<double4>(), h->pdL1.as<double4>(), h->pdL2.as<double4>(),
                                              h->pdADX.as<double4>(), h->pdDU.as<double4>(), h->pdDL1.as<double4>(),
                                              h->pdDL2.as<double4>(), h->pdEV.as<double4>(), m, ctl, part);
    IRA_TRY(launch_check(h, "k_pd_direction"));
    IRA_TRY(pd_At(h, h->pdATDV, 0));
    for (int guard = 0; guard < 40; ++guard) {                        // back-tracking (:392-429), <= 33 rounds
      k_pd_trial_edges<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                                  h->pdL2.as<double4>(), h->pdADX.as<double4>(), h->pdDU.as<double4>(),
                                                  h->pdDL1.as<double4>(), h->pdDL2.as<double4>(), m, ctl, part,
                                                  h->pdtrial.as<double>());
      IRA_TRY(launch_check(h, "k_pd_trial_edges"));
      k_pd_trial_nodes<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->pdATV.as<double4>(), h->pdATDV.as<double4>(), n, ctl, part,
                                                               h->pdtrial.as<double>());
      IRA_TRY(launch_check(h, "k_pd_trial_nodes"));
      IRA_TRY(fetch_pdctl(h));
      if (!h->h_pdctl->pending) break;
    }
    k_pd_accept_edges<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                                 h->pdL2.as<double4>(), h->pdADX.as<double4>(), h->pdDU.as<double4>(),
                                                 h->pdDL1.as<double4>(), h->pdDL2.as<double4>(), m, ctl, part);
    IRA_TRY(launch_check(h, "k_pd_accept_edges"));
    k_pd_accept_nodes<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->pdX.as<double4>(), h->pdATV.as<double4>(),
                                                              h->pdDX.as<double4>(), h->pdATDV.as<double4>(), n, ctl);
    IRA_TRY(launch_check(h, "k_pd_accept_nodes"));
    k_pd_rcent<<<ge, 256, 0, h->stream>>>(Y, h->pdU.as<double4>(), h->pdAX.as<double4>(), h->pdL1.as<double4>(),
                                          h->pdL2.as<double4>(), m, 1, pdmaxiter, ctl, part);
    IRA_TRY(launch_check(h, "k_pd_rcent"));
    IRA_TRY(fetch_pdctl(h));
    IRA_TRY(fetch_ctl(h));
    *newton_cg_iters += h->h_ctl->cg_iters;
    if (getenv("IRA_DEBUG") && atoi(getenv("IRA_DEBUG")) >= 2)
      fprintf(stderr, "[ira]   newton solve: %d PCG its, |r|^2 %.17g %.17g %.17g, |b|^2 %.17g %.17g %.17g\n", h->h_ctl->cg_iters,
              h->h_ctl->rnorm2[0], h->h_ctl->rnorm2[1], h->h_ctl->rnorm2[2], h->h_ctl->bnorm2[0], h->h_ctl->bnorm2[1], h->h_ctl->bnorm2[2]);
    for (int k = 0; k < 3; ++k)
      if (!(h->h_ctl->rnorm2[k] <= h->opt.cg_rtol * h->opt.cg_rtol * h->h_ctl->bnorm2[k])) { *hit_max += 1; break; }
    const PdCtl& c = *h->h_pdctl;
    if (!c.active[0] && !c.active[1] && !c.active[2]) break;
  }
  return IRA_OK;
}

}  // namespace

extern "C" {

ira_status ira_resident_start(ira_handle h, int32_t mode) {
  if (!h || (mode != 0 && mode != 1)) return IRA_ERR_INVALID_ARG;
  h->start_mode = mode;
  return IRA_OK;
}

ira_status ira_l1ra_resident(ira_handle h, int32_t max_iters, double change_th, int32_t* iters_out,
                             double* runtime_s_out, ira_stats* stats) {
  if (!h) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  // world_size > 1: with the whole graph on every rank (shard_mode 1 / 2) each rank runs the stage itself - the
  // results are bitwise identical on all ranks (deterministic kernels); edge shards (shard_mode 0) cannot
  if (h->opt.world_size > 1 && !h->peer && !h->replicated) { h->err = "l1ra with world_size > 1 needs the whole graph on every rank (shard_mode 1)"; return IRA_ERR_INVALID_ARG; }
  if (h->pcg_blocks_per_sm <= 0) { h->err = "l1ra needs cooperative launch"; return IRA_ERR_INVALID_ARG; }
  IRA_CUDA(h, cudaSetDevice(h->device));
  // irls may have chosen the CSR kernels (lanes_per_row set, or SELL padding blow-up on hub graphs); l1ra's
  // kernels walk the SELL copy whatever its padding: build it now if the upload did not
  if (!h->sell_built) { IRA_TRY(build_sell(h)); h->sell_built = true; }
  const auto t0 = std::chrono::steady_clock::now();
  const int launches0 = h->launches;
  if (stats) memset(stats, 0, sizeof *stats);
  IRA_TRY(l1ra_alloc(h));
  const int n = h->n;
  if (n > 0 && h->start_mode == 0)
    IRA_CUDA(h, cudaMemcpyAsync(h->Q.p, h->Q0.p, sizeof(double4) * n, cudaMemcpyDeviceToDevice, h->stream));
  IRA_CUDA(h, cudaMemsetAsync(h->weights.p, 0, sizeof(double) * h->m_pad, h->stream));
  Ctl init;
  memset(&init, 0, sizeof init);
  *h->h_ctl = init;
  IRA_CUDA(h, cudaMemcpyAsync(h->ctl.p, h->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  double score = std::numeric_limits<double>::max();                       // :866
  int iter = 0;
  const int l1_step = 2;                                                   // :868 (never changes, App. A.6.7)
  ira_status rc = IRA_OK;
  while (score >= change_th && iter < max_iters) {                          // :877
    IRA_TRY(run_residual(h, h->Q.as<double4>(), 0));                        // :885-887
    int cg = 0, hit = 0;
    if (h->m > 0) IRA_TRY(l1decode_pd_device(h, l1_step, &cg, &hit));       // :889-892
    else IRA_CUDA(h, cudaMemsetAsync(h->pdX.p, 0, sizeof(double4) * (size_t)std::max(n, 1), h->stream));
    k_update<<<grid_nodes(h, std::max(1, n - h->f), kRedThreads), kRedThreads, 0, h->stream>>>(
        h->Q.as<double4>(), h->pdX.as<double4>(), n, h->f, h->ctl.as<Ctl>(), h->partials.as<double>());   // :894-902
    IRA_TRY(launch_check(h, "k_update"));
    IRA_TRY(fetch_ctl(h));
    score = h->h_ctl->score;
    if (stats && iter < IRA_STATS_MAX_ITERS) { stats->score[iter] = score; stats->cg_iters[iter] = cg; }
    if (stats) { stats->cg_iters_total += cg; stats->cg_hit_max += hit; }
    iter++;
    if (!std::isfinite(score)) { if (n - h->f > 0) rc = IRA_ERR_NONFINITE; break; }
  }
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  if (iters_out) *iters_out = iter;
  if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (stats) { stats->irls_iters = iter; stats->kernel_launches = h->launches - launches0; }
  if (rc == IRA_ERR_NONFINITE) h->err = "l1ra score became non-finite";
  return rc;
}

ira_status ira_l1ra(ira_handle h, int64_t m, int64_t n_total, int32_t f, const int32_t* I_pairs, const double* QQ,
                    int64_t ld_qq, double* Q, int64_t ld_q, int32_t max_iters, double change_th, int32_t* iters_out,
                    double* runtime_s_out, ira_stats* stats) {
  const auto t0 = std::chrono::steady_clock::now();
  IRA_TRY(check_args(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  IRA_TRY(ira_problem_upload(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  const ira_status rc = ira_l1ra_resident(h, max_iters, change_th, iters_out, nullptr, stats);
  if (rc != IRA_OK && rc != IRA_ERR_NONFINITE) return rc;
  IRA_TRY(ira_problem_download(h, Q, ld_q, nullptr));
  if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// irotavg::init_mst  (ral/l1_irls.hpp:89-90, ral/l1_irls.cpp:915-979) on the device, ira_mst.cuh
// ---------------------------------------------------------------------------------------------
extern "C" {

ira_status ira_init_mst_resident(ira_handle h, int32_t f_init, ira_mst_stats* stats) {
  if (!h) return IRA_ERR_INVALID_ARG;
  if (!h->uploaded) return IRA_ERR_NOT_UPLOADED;
  if (f_init < 1) { h->err = "init_mst: f < 1 (ral/l1_irls.cpp:917)"; return IRA_ERR_INVALID_ARG; }
  IRA_CUDA(h, cudaSetDevice(h->device));
  const int n = h->n;
  const int64_t m = h->m;
  if (stats) memset(stats, 0, sizeof *stats);
  if (n <= 1) return IRA_OK;                                              // count == n at once (:925)
  IRA_CUDA(h, h->mst_label.reserve(sizeof(unsigned long long) * (size_t)n));
  IRA_CUDA(h, h->mst_label2.reserve(sizeof(unsigned long long) * (size_t)n));
  IRA_CUDA(h, h->mst_order.reserve(sizeof(int) * (size_t)n));
  IRA_CUDA(h, h->mst_order2.reserve(sizeof(int) * (size_t)n));
  IRA_CUDA(h, h->mst_done.reserve(sizeof(int) * (size_t)n));
  IRA_CUDA(h, h->mst_T0.reserve(sizeof(double4) * (size_t)n));
  IRA_CUDA(h, h->mst_T1.reserve(sizeof(double4) * (size_t)n));
  IRA_CUDA(h, h->mst_ctl.reserve(sizeof(MstCtl)));
  EventPair evp;
  IRA_CUDA(h, evp.create());
  cudaEvent_t ev0 = evp.a, ev1 = evp.b;
  IRA_CUDA(h, cudaEventRecord(ev0, h->stream));
  k_mst_init<<<cdiv(n, 256), 256, 0, h->stream>>>(h->mst_label.as<unsigned long long>(), h->mst_done.as<int>(),
                                                 h->mst_order.as<int>(), n, h->mst_ctl.as<MstCtl>());
  IRA_TRY(launch_check(h, "k_mst_init"));
  int per_sm = 0;
  IRA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mst_labels, 256, 0));
  per_sm = std::max(1, std::min(per_sm, 4));
  {
    const int2* I = h->I.as<int2>();
    int64_t mm = m;
    unsigned long long* lab = h->mst_label.as<unsigned long long>();
    MstCtl* ctl = h->mst_ctl.as<MstCtl>();
    // one warp per segment of kMstSeg consecutive edges
    const int grid = std::max(1, std::min(cdiv(cdiv(std::max<int64_t>(m, 1), kMstSeg) * 32, 256), h->sms * per_sm));
    void* args[] = {(void*)&I, (void*)&mm, (void*)&lab, (void*)&ctl};
    IRA_CUDA(h, cudaLaunchCooperativeKernel((void*)k_mst_labels, dim3(grid), dim3(256), args, 0, h->stream));
    h->launches++;
  }
  // parents and factors from the labels, then the tree contracted by pointer jumping (O(log depth) rounds)
  k_mst_parents<<<grid_nodes(h, n), 256, 0, h->stream>>>(h->I.as<int2>(), h->QQ.as<double>(), h->m_pad,
                                                       h->mst_label.as<unsigned long long>(), n, f_init,
                                                       h->mst_order.as<int>(), h->mst_T0.as<double4>(), h->mst_ctl.as<MstCtl>());
  IRA_TRY(launch_check(h, "k_mst_parents"));
  IRA_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mst_jump, 256, 0));
  per_sm = std::max(1, std::min(per_sm, 4));
  {
    int* a0 = h->mst_order.as<int>(); int* a1 = h->mst_order2.as<int>();
    double4* t0 = h->mst_T0.as<double4>(); double4* t1 = h->mst_T1.as<double4>();
    int nn = n;
    double4* Q = h->Q0.as<double4>();
    MstCtl* ctl = h->mst_ctl.as<MstCtl>();
    const int grid = std::max(1, std::min(cdiv(n, 256), h->sms * per_sm));
    void* args[] = {(void*)&a0, (void*)&t0, (void*)&a1, (void*)&t1, (void*)&nn, (void*)&Q, (void*)&ctl};
    IRA_CUDA(h, cudaLaunchCooperativeKernel((void*)k_mst_jump, dim3(grid), dim3(256), args, 0, h->stream));
    h->launches++;
  }
  IRA_CUDA(h, cudaMemcpyAsync(h->Q.p, h->Q0.p, sizeof(double4) * (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
  IRA_CUDA(h, cudaEventRecord(ev1, h->stream));
  MstCtl host;
  IRA_CUDA(h, cudaMemcpyAsync(&host, h->mst_ctl.p, sizeof host, cudaMemcpyDeviceToHost, h->stream));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ev0, ev1);
  if (stats) {
    stats->passes_label = host.passes_label; stats->passes_propagate = host.passes_prop;
    stats->unreached = host.unreached; stats->t_ms = ms;
  }
  if (host.unreached > 0) {
    char buf[256];
    snprintf(buf, sizeof buf, "Relative rotations DO NOT SPAN all the nodes in the VIEW GRAPH\n"
             "Number of nodes in Spanning Tree = %d", n - host.unreached);
    h->err = buf;
    return IRA_ERR_NOT_SPANNING;
  }
  return IRA_OK;
}

ira_status ira_init_mst(ira_handle h, int64_t m, int64_t n_total, int32_t f_init, const int32_t* I_pairs,
                        const double* QQ, int64_t ld_qq, double* Q, int64_t ld_q, ira_mst_stats* stats) {
  IRA_TRY(check_args(h, m, n_total, f_init, I_pairs, QQ, ld_qq, Q, ld_q));
  IRA_TRY(ira_problem_upload(h, m, n_total, f_init, I_pairs, QQ, ld_qq, Q, ld_q));
  IRA_TRY(ira_init_mst_resident(h, f_init, stats));
  return ira_problem_download(h, Q, ld_q, nullptr);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// l1ra followed by irls on one upload: the call sequence of both callers
// (ral/test.cpp:295-300, src/ViewGraph.cpp:1407-1417)
// ---------------------------------------------------------------------------------------------
namespace {
// The window-sized call as one launch of one block; host buffers <-> mapped pinned blocks.
ira_status small_l1ra_irls(ira_context* h, int m, int n, int f, const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                           double* Q, int64_t ld_q, int l1_max, double l1_th, int cost, double sigma, int irls_max,
                           double irls_th, double* weights, int32_t* l1_out, int32_t* irls_out, ira_stats* st) {
  IRA_CUDA(h, cudaSetDevice(h->device));
  if (!h->small_ready) {
    IRA_CUDA(h, cudaHostAlloc((void**)&h->small_in, small_in_bytes(kSmM, kSmN), cudaHostAllocMapped));
    IRA_CUDA(h, cudaHostAlloc((void**)&h->small_out, small_out_bytes(kSmM, kSmN), cudaHostAllocMapped));
    IRA_CUDA(h, cudaFuncSetAttribute(k_small_l1ra_irls, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SmallSmem)));
    h->small_ready = true;
  }
  for (int k = 0; k < 2 * m; ++k)
    if (I_pairs[k] < 0 || I_pairs[k] >= n) { h->err = "edge endpoint out of range [0, n_total)"; return IRA_ERR_INVALID_ARG; }
  for (int k = 0; k < m; ++k)
    if (I_pairs[2 * k] == I_pairs[2 * k + 1]) { h->err = "self-loop edge (i == j) is not a relative rotation between two views"; return IRA_ERR_INVALID_ARG; }
  SmallIn hd;
  memset(&hd, 0, sizeof hd);
  hd.m = m; hd.n = n; hd.f = f; hd.cost = cost; hd.l1_max_iters = l1_max; hd.irls_max_iters = irls_max;
  hd.l1_th = l1_th; hd.irls_th = irls_th; hd.sigma = sigma;
  memcpy(h->small_in, &hd, sizeof hd);
  memcpy(h->small_in + small_in_I(), I_pairs, sizeof(int32_t) * 2 * (size_t)m);
  double* qq = reinterpret_cast<double*>(h->small_in + small_in_QQ(m));
  double* q0 = reinterpret_cast<double*>(h->small_in + small_in_Q(m));
  for (int c = 0; c < 4; ++c) {
    memcpy(qq + (size_t)c * m, QQ + (size_t)c * ld_qq, sizeof(double) * (size_t)m);
    memcpy(q0 + (size_t)c * n, Q + (size_t)c * ld_q, sizeof(double) * (size_t)n);
  }
  unsigned char *din = nullptr, *dout = nullptr;
  IRA_CUDA(h, cudaHostGetDevicePointer((void**)&din, h->small_in, 0));
  IRA_CUDA(h, cudaHostGetDevicePointer((void**)&dout, h->small_out, 0));
  k_small_l1ra_irls<<<1, kSmThreads, sizeof(SmallSmem), h->stream>>>(din, dout);
  IRA_TRY(launch_check(h, "k_small_l1ra_irls"));
  IRA_CUDA(h, cudaStreamSynchronize(h->stream));
  const SmallOut* oh = reinterpret_cast<const SmallOut*>(h->small_out);
  const double* qo = reinterpret_cast<const double*>(h->small_out + sizeof(SmallOut));
  for (int c = 0; c < 4; ++c) memcpy(Q + (size_t)c * ld_q, qo + (size_t)c * n, sizeof(double) * (size_t)n);
  if (m > 0) memcpy(weights, qo + (size_t)4 * n, sizeof(double) * (size_t)m);
  if (l1_out) *l1_out = oh->l1_iters;
  if (irls_out) *irls_out = oh->irls_iters;
  if (st) {
    memset(st, 0, sizeof *st);
    st->irls_iters = oh->irls_iters;
    st->kernel_launches = 1;
    st->pcg_kernel = 6;
    for (int k = 0; k < std::min(oh->irls_iters, kSmScores); ++k) st->score[k] = oh->irls_score[k];
  }
  if (oh->nonfinite) { h->err = "score became non-finite"; return IRA_ERR_NONFINITE; }
  return IRA_OK;
}
}  // namespace

extern "C" ira_status ira_l1ra_irls(ira_handle h, int64_t m, int64_t n_total, int32_t f, const int32_t* I_pairs,
                                    const double* QQ, int64_t ld_qq, double* Q, int64_t ld_q, int32_t l1_max_iters,
                                    double l1_change_th, int32_t cost, double sigma, int32_t irls_max_iters,
                                    double irls_change_th, double* weights, int32_t* l1_iters_out,
                                    int32_t* irls_iters_out, double* runtime_s_out, ira_stats* irls_stats) {
  const auto t0 = std::chrono::steady_clock::now();
  IRA_TRY(check_args(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  if (m > 0 && !weights) { h->err = "null weights"; return IRA_ERR_INVALID_ARG; }
  if (cost < 0 || cost >= kNumCosts) { h->err = "Unknown cost!!"; return IRA_ERR_UNKNOWN_COST; }
  if (h->opt.small_path == 0 && h->opt.world_size <= 1 && n_total <= kSmN && n_total - f >= 1 && n_total - f <= kSmNF &&
      m <= kSmM) {
    const ira_status rs = small_l1ra_irls(h, (int)m, (int)n_total, f, I_pairs, QQ, ld_qq, Q, ld_q, l1_max_iters,
                                          l1_change_th, cost, sigma, irls_max_iters, irls_change_th, weights,
                                          l1_iters_out, irls_iters_out, irls_stats);
    if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return rs;
  }
  IRA_TRY(ira_problem_upload(h, m, n_total, f, I_pairs, QQ, ld_qq, Q, ld_q));
  static thread_local ira_stats l1_stats;                  // only its non-convergence count is passed on
  ira_status rc = ira_l1ra_resident(h, l1_max_iters, l1_change_th, l1_iters_out, nullptr, &l1_stats);
  if (rc != IRA_OK && rc != IRA_ERR_NONFINITE) return rc;
  h->start_mode = 1;                                       // irls continues from l1ra's rotations
  static thread_local ira_stats own_stats;
  ira_stats* const ist = irls_stats ? irls_stats : &own_stats;
  rc = ira_irls_resident(h, cost, sigma, irls_max_iters, irls_change_th, irls_iters_out, nullptr, ist);
  h->start_mode = 0;
  if (getenv("IRA_DEBUG")) {
    fprintf(stderr, "[ira] l1ra_irls n %lld m %lld f %d: l1ra %d its, %d Newton PCG its, hit %d, score %.17g; irls %d its, %d PCG its, hit %d, "
            "kernel %d, score %.17g; %.3f ms\n",
            (long long)n_total, (long long)m, f, l1_stats.irls_iters, l1_stats.cg_iters_total, l1_stats.cg_hit_max,
            l1_stats.irls_iters > 0 ? l1_stats.score[std::min(l1_stats.irls_iters, IRA_STATS_MAX_ITERS) - 1] : 0.0, ist->irls_iters,
            ist->cg_iters_total, ist->cg_hit_max, ist->pcg_kernel,
            ist->irls_iters > 0 ? ist->score[std::min(ist->irls_iters, IRA_STATS_MAX_ITERS) - 1] : 0.0,
            1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  }
  if (irls_stats) irls_stats->cg_hit_max += l1_stats.cg_hit_max;   // unconverged Newton solves of the l1ra stage count too
  if (rc != IRA_OK && rc != IRA_ERR_NONFINITE) return rc;
  IRA_TRY(ira_problem_download(h, Q, ld_q, weights));
  if (runtime_s_out) *runtime_s_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}
