/* ira.h - C ABI of irotavg-b200: the IRLS rotation-averaging hot path of ajparra/iRotAvg on B200.
 *
 * This is the drop-in boundary.  Every entry point names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only; no C++/torch types.
 * The implementation is hand-written sm_100a CUDA (irotavg_b200/csrc); there is NO CPU fallback:
 * every compute entry point returns IRA_ERR_NO_DEVICE / IRA_ERR_CUDA when no B200-class device
 * is usable.
 *
 * Conventions (identical to the reference, ral/l1_irls.hpp:80-107, ral/test.cpp:193,221):
 *   - quaternions are [x y z w]; matrices are COLUMN-MAJOR doubles with an explicit leading
 *     dimension (Eigen::MatrixXd memory as-is: QQ is m x 4 with ld >= m, Q is n x 4 with ld >= n);
 *   - edge k = (i, j) = I[2k], I[2k+1] (std::vector<std::pair<int,int>> memory as-is) means
 *     Q_j = QQ_k (x) Q_i (ral/l1_irls.cpp:941);
 *   - the first f rows of Q are fixed, f >= 1 (ral/l1_irls.cpp:917);
 *   - `weights` are the square-root IRLS weights of the last iteration (ral/l1_irls.cpp:617-727).
 *
 * Threading: a handle is not thread-safe; distinct handles are independent.  All calls block
 * until their results are in the caller's buffers.
 */
#ifndef IRA_H_
#define IRA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRA_ABI_VERSION 2

typedef struct ira_context* ira_handle;

typedef enum ira_status {
  IRA_OK = 0,
  IRA_ERR_INVALID_ARG = 1,   /* null pointer, negative size, f < 1, edge index out of range ...  */
  IRA_ERR_NO_DEVICE = 2,     /* no CUDA device / not an sm_100 part: there is no CPU path        */
  IRA_ERR_CUDA = 3,          /* a CUDA call failed; ira_last_error() has the text                */
  IRA_ERR_COMM = 4,          /* NCCL missing or a collective failed                               */
  IRA_ERR_UNKNOWN_COST = 5,  /* "Unknown cost!!" (ral/l1_irls.cpp:723-726)                        */
  IRA_ERR_NOT_UPLOADED = 6,  /* resident call before ira_problem_upload                           */
  IRA_ERR_NONFINITE = 7,     /* NaN/Inf score: the iterate left the finite range                 */
  IRA_ERR_NOT_SPANNING = 8   /* init_mst: relative rotations do not span (ral/l1_irls.cpp:970-977)*/
} ira_status;

/* enum Cost of ral/l1_irls.hpp:56-57, same integer values. */
typedef enum ira_cost {
  IRA_COST_L2 = 0, IRA_COST_L1, IRA_COST_L15, IRA_COST_L05, IRA_COST_GEMAN_MCCLURE, IRA_COST_HUBER,
  IRA_COST_PSEUDO_HUBER, IRA_COST_ANDREWS, IRA_COST_BISQUARE, IRA_COST_CAUCHY, IRA_COST_FAIR,
  IRA_COST_LOGISTIC, IRA_COST_TALWAR, IRA_COST_WELSCH
} ira_cost;

/* The reference solves each linear step exactly (SuiteSparseQR, ral/l1_irls.cpp:550).  This
 * library solves the same weighted normal equations A^T D^2 A X = A^T D^2 w with Jacobi-PCG;
 * these knobs bound the difference.  ira_options_default() fills the values used for parity. */
typedef struct ira_options {
  int32_t device;          /* CUDA ordinal; -1 = current device                                   */
  int32_t cg_max_iters;    /* hard cap per linear solve                                           */
  double  cg_rtol;         /* stop when ||r_c|| <= cg_rtol * ||b_c|| for each of the 3 columns    */
  double  pair_theta;      /* block-Jacobi pairing threshold: nodes v,u whose edge has strength
                              w2_vu / sqrt(d_v d_u) >= pair_theta and who are each other's strongest
                              neighbour are preconditioned together (2x2 blocks, as an additive
                              coarse correction); 0 = plain Jacobi.  Default 0.5 (sweep: profiles/
                              r02_sweep_theta_*.json)                                             */
  int32_t cg_check_every;  /* host polls the device-side convergence flag every this many iters   */
  int32_t lanes_per_row;   /* 0 = SELL-32 thread-per-row kernels (default); 2..32 = CSR kernels
                              with that many lanes per row (forces solver 1)                       */
  int32_t world_size;      /* >1: edges are sharded over ranks, node vectors all-reduced (NCCL)   */
  int32_t rank;
  int32_t profile;         /* 1: record CUDA-event timings per kernel class into ira_stats; 2: also let the
                              small-graph PCG kernel print its per-iteration phase split (debugging)          */
  int32_t solver;          /* PCG driver: 0 = auto (persistent cooperative kernel on one GPU), 1 = one
                              kernel per CG step with host-polled convergence, 2 = persistent;
                              +4 = persistent kernel keeps its vectors in HBM even when one row per
                              lane would let them live in registers (A/B measurement);
                              +8 = one GPU only: the barrier-free kernel of the multi-GPU path
                              (self-validating data instead of grid barriers; measured slower, kept for A/B);
                              +128 = never use the two-level kernel for small / medium graphs (ira_coarse.cuh; A/B);
                              +256 = two-level kernel: dense 64-block coarse space only, never the tridiagonal one of
                              up to 1 024 blocks (A/B);
                              (ira_options_default presets this field from the environment variable IRA_SOLVER, so that
                              callers taking the defaults - the C++ adapters, the CLI - can be A/B-tested too)
                              +64 = deal the SELL slices to the blocks round-robin instead of balanced by entries (A/B);
                              +32 = the matrix-in-shared-memory kernel (ira_pcg2.cuh; measured SLOWER: it leaves the SM
                              28 KB of L1 and the gathers lose their memory-level parallelism; kept for A/B)   */
  int32_t spmv_variant;    /* experiment knob: gather flavour / unroll of the SELL SpMV (0 = default)  */
  int32_t small_path;      /* window-sized problems (n_total <= 64, 1 <= n_free <= 32, m <= 256) in
                              ira_l1ra_irls run as ONE single-block kernel with dense Cholesky solves
                              (irotavg_b200/csrc/ira_small.cuh): 0 = yes (default), 1 = never              */
  int32_t shard_mode;      /* world_size > 1 only.  0: every rank passes ITS EDGE SHARD; per-node partial sums are
                              all-reduced with NCCL once per PCG iteration.  1: every rank passes the WHOLE graph;
                              the rows of A^T D^2 A are partitioned over the ranks and ONE persistent kernel per
                              rank runs the solve, exchanging vector slices / dot products / barrier flags by
                              loads and stores into the peers' HBM over NVLink (CUDA IPC mappings,
                              irotavg_b200/csrc/ira_peer.cuh); weights come back whole on every rank.  1 = barrier-free
                              exchange (self-validating data), 2 = the same with two cross-GPU barriers per iteration */
  int32_t peer_min_rows;   /* world_size > 1, shard_mode 1 / 2: a graph with fewer nodes than this is NOT partitioned -
                              every rank solves the whole problem itself with the single-GPU kernels (bitwise
                              identical results, no exchange): below ~10^5 rows one PCG iteration is shorter than the
                              NVLink latencies a partitioned iteration needs.  Default 120000; 0 = always partition */
  int32_t reserved[2];
  double  pair_theta3;     /* a still-single node joins the pair holding its strongest neighbour when that edge's
                              normalised strength is >= pair_theta3 (3x3 blocks, inverted exactly); 0 = pairs only.
                              Default 0.001                                                                     */
} ira_options;

#define IRA_STATS_MAX_ITERS 256
typedef struct ira_stats {
  int32_t irls_iters;
  int32_t cg_iters_total;
  int32_t kernel_launches;                 /* kernels of this library launched by the call        */
  int32_t cg_hit_max;                      /* number of solves that stopped at cg_max_iters       */
  double  score[IRA_STATS_MAX_ITERS];      /* mean |X_i| per IRLS iteration (ral/l1_irls.cpp:729)  */
  int32_t cg_iters[IRA_STATS_MAX_ITERS];   /* PCG iterations per IRLS iteration                   */
  double  cg_relres[IRA_STATS_MAX_ITERS];  /* max_c ||r_c|| / ||b_c|| at exit                     */
  double  t_total_ms;                      /* device time of the whole call (CUDA events)         */
  double  t_upload_ms, t_download_ms;      /* host<->device + CSR build (host-buffer call only)   */
  /* profile=1 only: summed device time and launch count per kernel class                        */
  double  t_residual_ms, t_rhs_ms, t_spmv_ms, t_cgvec_ms, t_weights_ms, t_update_ms, t_comm_ms;
  int32_t n_residual, n_rhs, n_spmv, n_cgvec, n_weights, n_update, n_comm;
  /* persistent solve (one cooperative kernel per linear step): event time of those kernels
   * (profile=1), and - always - block 0's in-kernel clocks: time in the SpMV+reduction phases, in
   * the vector-update phases, whole-kernel time, number of SpMV phases executed                   */
  int32_t n_pcg, pcg_spmv_phases;
  double  t_pcg_ms, pcg_spmv_ms, pcg_update_ms, pcg_kernel_ms;
  /* which linear-solve driver ran (ABI 2): 0 one kernel per CG step, 1 k_pcg_persistent (vectors in HBM),
   * 2 k_pcg_persistent_reg, 3 k_pcg_persistent_reg_mw, 4 k_pcg_smem (matrix in shared memory), 5 peer-memory kernels,
   * 6 single-block window solver, 7 k_pcg_coarse_w3 (two-level: Jacobi + dense coarse space, ira_coarse.cuh),
   * 8 the same with the tridiagonal coarse operator (cyclic reduction, <= 1 024 blocks) */
  int32_t pcg_kernel;
  int32_t reserved_stats;
} ira_stats;

ira_status  ira_options_default(ira_options* opt);
ira_status  ira_create(ira_handle* out, const ira_options* opt /* NULL = defaults */);
ira_status  ira_destroy(ira_handle h);
const char* ira_status_string(ira_status s);
const char* ira_last_error(ira_handle h);      /* text of the last failure on this handle          */
int32_t     ira_abi_version(void);
int32_t     ira_device_count(void);            /* usable CUDA devices (0 on a CPU-only box)        */
/* The CUDA stream (cudaStream_t) every kernel of this handle is launched on, so that a harness
 * can record its own CUDA events on it. */
ira_status  ira_get_stream(ira_handle h, void** stream_out);

/* ---- irotavg::irls  (ral/l1_irls.hpp:103-106, ral/l1_irls.cpp:559-752) ----------------------
 * Host buffers in, host buffers out.  Q rows [f, n_total) are updated in place, `weights` (m)
 * is overwritten (the reference's weights.setOnes() start is implied), *iters_out / *runtime_s_out
 * are the reference's `iters` / `runtime` out-parameters (runtime is wall seconds of the call).
 * The incidence matrix A of the reference signature is a pure function of (n_total, f, I)
 * (ral/l1_irls.cpp:755-780) and is rebuilt on the device, including make_A's rule that an edge
 * whose second endpoint is fixed contributes nothing.  max_iters == 0 returns Q untouched.
 * With world_size > 1 every rank passes ITS edge shard (I, QQ, weights of m_local edges) and the
 * same full Q (shard_mode 0), or the whole graph (shard_mode 1); Q comes back identical on all ranks. */
ira_status ira_irls(ira_handle h, int64_t m, int64_t n_total, int32_t f,
                    const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                    double* Q, int64_t ld_q,
                    int32_t cost, double sigma, int32_t max_iters, double change_th,
                    double* weights, int32_t* iters_out, double* runtime_s_out,
                    ira_stats* stats /* may be NULL */);

/* ---- device-resident variant: the same loop with the graph already in HBM -------------------
 * upload:   copies (I, QQ, Q0) to the device and builds the CSR pattern of A^T A once.
 * resident: runs irls() from the uploaded Q0 (restored on the device first); nothing crosses PCIe
 *           except the per-iteration `score` word the loop test needs (ral/l1_irls.cpp:590).
 * download: copies the current Q (n_total x 4, column-major, ld_q) and weights (m) back. */
ira_status ira_problem_upload(ira_handle h, int64_t m, int64_t n_total, int32_t f,
                              const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                              const double* Q0, int64_t ld_q);
ira_status ira_irls_resident(ira_handle h, int32_t cost, double sigma, int32_t max_iters,
                             double change_th, int32_t* iters_out, double* runtime_s_out,
                             ira_stats* stats);
ira_status ira_problem_download(ira_handle h, double* Q, int64_t ld_q, double* weights);

/* ---- irotavg::l1ra  (ral/l1_irls.hpp:98-101, ral/l1_irls.cpp:851-912) -------------------------
 * The L1RA initial stage both callers run before irls (ral/test.cpp:295, src/ViewGraph.cpp:1407):
 * per outer iteration three primal-dual interior-point L1 regressions (l1decode_pd, ral/l1_irls.cpp:
 * 228-468, two Newton steps each), whose Newton systems A'^T diag(sigx) A' dx = w1p (UMFPACK LU in the
 * reference) are solved by the persistent PCG kernel.  Q rows [f, n_total) are updated in place;
 * *iters_out / *runtime_s_out are the reference's `iter` / `runtime`.  stats->score[] holds the outer
 * scores, stats->cg_iters[] the Newton-PCG iterations per outer iteration.  Works on every graph irls accepts
 * (hub graphs whose SELL padding made irls choose its CSR kernels included: the stage walks the SELL copy anyway,
 * slowly on the hub rows).  world_size > 1: every rank runs it on its whole-graph copy (shard_mode 1 / 2), with
 * bitwise identical results on all ranks; edge shards (shard_mode 0) return IRA_ERR_INVALID_ARG. */
ira_status ira_l1ra(ira_handle h, int64_t m, int64_t n_total, int32_t f,
                    const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                    double* Q, int64_t ld_q, int32_t max_iters, double change_th,
                    int32_t* iters_out, double* runtime_s_out, ira_stats* stats /* may be NULL */);
ira_status ira_l1ra_resident(ira_handle h, int32_t max_iters, double change_th,
                             int32_t* iters_out, double* runtime_s_out, ira_stats* stats);
/* mode 0 (default after an upload): ira_*_resident calls restart from the uploaded Q0;
 * mode 1: they continue from the current device Q (l1ra followed by irls, like both callers). */
ira_status ira_resident_start(ira_handle h, int32_t mode);

/* ---- l1ra then irls on ONE upload: what ViewGraph::rotAvg (src/ViewGraph.cpp:1400-1417) and the CLI
 * (ral/test.cpp:288-300) do back to back.  Same arguments as the two calls; the rotations stay on the
 * device between the stages.  *runtime_s_out is the wall time of the whole call.  Window-sized problems
 * (see ira_options.small_path) take a single launch with exact dense solves; irls_stats then carries
 * irls_iters, the first 8 scores and kernel_launches = 1.  irls_stats->cg_hit_max counts the linear solves of BOTH
 * stages that stopped at cg_max_iters without reaching cg_rtol. */
ira_status ira_l1ra_irls(ira_handle h, int64_t m, int64_t n_total, int32_t f,
                         const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                         double* Q, int64_t ld_q,
                         int32_t l1_max_iters, double l1_change_th,
                         int32_t cost, double sigma, int32_t irls_max_iters, double irls_change_th,
                         double* weights, int32_t* l1_iters_out, int32_t* irls_iters_out,
                         double* runtime_s_out, ira_stats* irls_stats /* may be NULL */);

/* ---- irotavg::init_mst  (ral/l1_irls.hpp:89-90, ral/l1_irls.cpp:915-979) ----------------------
 * Spanning-tree start: the reference sweeps the edge list in order until every node is flagged; the
 * tree (and so the result) depends on the edge order.  The device reproduces exactly that tree
 * (time-stamp relaxation, irotavg_b200/csrc/ira_mst.cuh) and the same chain of products per node, associated by
 * pointer jumping (O(log depth) rounds; differences to the sequential evaluation are rounding only).  Rows
 * [0, f_init) of Q are kept (ral/test.cpp:285-286 passes max(#given rotations, f)); the others are
 * overwritten.  IRA_ERR_NOT_SPANNING when the edges do not reach every node (:970-977; Q undefined).
 * The resident form acts on the uploaded problem: both the restart copy Q0 and the current Q get the
 * tree start, so that ira_l1ra_resident / ira_irls_resident follow without crossing PCIe. */
typedef struct ira_mst_stats {
  int32_t passes_label;       /* relaxation passes until the time stamps were final          */
  int32_t passes_propagate;   /* pointer-jumping rounds of the rotation propagation          */
  int32_t unreached;          /* nodes not spanned                                           */
  double  t_ms;               /* device time (CUDA events)                                   */
} ira_mst_stats;
ira_status ira_init_mst(ira_handle h, int64_t m, int64_t n_total, int32_t f_init,
                        const int32_t* I_pairs, const double* QQ, int64_t ld_qq,
                        double* Q, int64_t ld_q, ira_mst_stats* stats /* may be NULL */);
ira_status ira_init_mst_resident(ira_handle h, int32_t f_init, ira_mst_stats* stats /* may be NULL */);

/* ---- irotavg::make_A  (ral/l1_irls.hpp:92, ral/l1_irls.cpp:755-780) -------------------------
 * Host helper (no device): writes per edge the two column indices of row k of A, or -1:
 * col_plus[k] = j-f if j >= f else -1;  col_minus[k] = i-f if (j >= f and i >= f) else -1. */
ira_status ira_make_A(int64_t m, int32_t n_total, int32_t f, const int32_t* I_pairs,
                      int32_t* col_plus, int32_t* col_minus);

/* ---- irotavg::quat_normalised  (ral/l1_irls.hpp:112, ral/l1_irls.cpp:982-991) ---------------
 * Host helper: normalise rows [f, n_total) of the column-major Q in place. */
ira_status ira_quat_normalised(double* Q, int64_t n_total, int64_t ld_q, int32_t f);

/* ---- kernel-level probes for the parity tests and the roofline bench ------------------------
 * They act on the uploaded problem.  residual: w_out is m x 4 column-major (ld = m) =
 * log_map(delta_rel(I, QQ, Q0)) (ral/l1_irls.cpp:592-593).  laplacian_apply: Y = A^T D^2 A X for
 * host X, Y of n_free x 3 column-major (ld = n_free) with D = diag(weights) given per edge.
 * time_kernel: average device time (CUDA events on the launch stream) of `reps` back-to-back
 * launches of one kernel class on the uploaded problem; which = 0 residual, 1 spmv, 2 rhs/diag,
 * 3 weights, 4 node update, 5 cg vector update; flush_l2 != 0 writes a >L2 buffer between reps. */
ira_status ira_probe_residual(ira_handle h, double* w_out);
ira_status ira_probe_laplacian_apply(ira_handle h, const double* weights, const double* X, double* Y);
ira_status ira_probe_time_kernel(ira_handle h, int32_t which, int32_t reps, int32_t flush_l2,
                                 double* avg_us_out);

/* ---- multi-GPU (one process per GPU) --------------------------------------------------------
 * Rank 0 calls ira_comm_unique_id and broadcasts the 128 bytes by any means (torch.distributed
 * in bench.py); every rank then calls ira_comm_init on its handle (created with world_size/rank
 * set).  Collectives are NCCL all-reduces of per-node partial sums on the handle's stream. */
ira_status ira_comm_unique_id(uint8_t id_out[128]);
ira_status ira_comm_init(ira_handle h, const uint8_t id[128]);

#ifdef __cplusplus
}
#endif
#endif /* IRA_H_ */
