// L1RA initial stage on the device: irotavg::l1ra (ral/l1_irls.cpp:851-912) and its inner solver
// l1decode_pd (ral/l1_irls.cpp:228-468, l1-magic's primal-dual interior-point L1 regression), for the
// three coordinates at once.  SURVEY 8(f) "next" row #1.
//
// Per outer iteration the reference runs, for each coordinate c, l1decode_pd(x0 = 0, A, y = w[:,c],
// pdmaxiter = 2, AtA): two damped Newton steps on the perturbed KKT system.  Every vector lives here
// as one 32 B double4 per edge or node holding the three coordinates (lane .w unused); every scalar of
// the reference (tau, sdg, resnorm, step length s, back-tracking state) is kept per coordinate in a
// device control block and decided by the last block of the kernel that completes its reduction, so
// the host only polls one flag per back-tracking round.
//   * A x / A^T v use make_A's pattern (edges with a fixed second endpoint are empty rows, :770-771);
//   * the Newton matrix H = reshape(AtA * sigx) (:308-317) uses make_AtA's pattern, which keeps those
//     edges on the diagonal (:825-835) - the kEidQuirk entries of the CSR/SELL pattern;
//   * H dx = w1p (UMFPACK LU in the reference, :319, 131-184) is solved by the persistent PCG kernel
//     with one weight set per coordinate (k_pcg_persistent_w3).
#pragma once
#include "ira_pcg.cuh"

namespace ira {

constexpr double kPdTol = 1e-3, kPdAlpha = 0.01, kPdBeta = 0.5, kPdMu = 10.0;     // :231-238

struct PdCtl {
  double ymax[3];          // max |y - A x0|                                         (:253)
  double sdg[3], tau[3], resnorm[3];
  double s[3];             // current step length per coordinate
  double rd2_edges[3];     // |rdual (m part)|^2 of the last evaluated point
  double rd2_nodes[3];     // |rdual (n part)|^2 = |Atv|^2
  double rc2[3];           // |rcent|^2
  double m2;               // 2 m as a double
  int active[3];           // coordinate still iterating (not done, not stuck)
  int suff[3];             // sufficient decrease reached in the current back-tracking
  int backiter[3];
  int stuck[3];            // "Stuck backtracking, returning last iterate." (:423-428)
  int pditer;
  int pending;             // 1 while some active coordinate still back-tracks (host polls this)
  unsigned int ticket;
};

// generic grid reduction with sum / max / min over NV values; totals valid in thread 0 of the last block
template <int NV, int OP>   // OP: 0 sum, 1 max, 2 min
__device__ __forceinline__ bool grid_reduce_op(double (&v)[NV], double* partials, unsigned int* ticket, double* sm,
                                               int* sm_flag) {
  auto comb = [](double a, double b) { return OP == 0 ? a + b : (OP == 1 ? fmax(a, b) : fmin(a, b)); };
  const double ident = OP == 0 ? 0.0 : (OP == 1 ? -INFINITY : INFINITY);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
  auto block = [&](double (&x)[NV]) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x[k] = comb(x[k], __shfl_xor_sync(0xffffffffu, x[k], o));
      if (lane == 0) sm[k * 32 + warp] = x[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        double t = lane < nwarps ? sm[k * 32 + lane] : ident;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = comb(t, __shfl_xor_sync(0xffffffffu, t, o));
        x[k] = t;
      }
    }
    __syncthreads();
  };
  block(v);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) partials[blockIdx.x * NV + k] = v[k];
    __threadfence();
    *sm_flag = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  const bool last = *sm_flag != 0;
  if (last) {
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = ident;
    for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
      for (int k = 0; k < NV; ++k) acc[k] = comb(acc[k], __ldcg(&partials[b * NV + k]));
    }
    block(acc);
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = acc[k];
    if (threadIdx.x == 0) *ticket = 0u;
  }
  return last;
}

#define IRA_EDGE_LOOP(k, m) \
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < (m); k += (int64_t)gridDim.x * blockDim.x)

// ---- start of l1decode_pd (x0 = 0 => A x0 = 0) ---------------------------------------------------
__global__ void __launch_bounds__(256)
k_pd_absmax(const double4* __restrict__ Y, int64_t m, PdCtl* ctl, double* partials) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  double v[3] = {-INFINITY, -INFINITY, -INFINITY};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k);
    v[0] = fmax(v[0], fabs(y.x)); v[1] = fmax(v[1], fabs(y.y)); v[2] = fmax(v[2], fabs(y.z));
  }
  if (grid_reduce_op<3, 1>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0)
    for (int c = 0; c < 3; ++c) ctl->ymax[c] = v[c];
}

// u = 0.95|y| + 0.10 max|y|; fu1 = -y - u; fu2 = y - u; lamu = -1/fu; EV = lamu1 - lamu2 (:248-262)
__global__ void __launch_bounds__(256)
k_pd_init(const double4* __restrict__ Y, double4* __restrict__ U, double4* __restrict__ AX, double4* __restrict__ L1,
          double4* __restrict__ L2, double4* __restrict__ EV, int64_t m, PdCtl* ctl, double* partials) {
  __shared__ double sm[6 * 32];
  __shared__ int flag;
  const double m0 = ctl->ymax[0], m1 = ctl->ymax[1], m2 = ctl->ymax[2];
  double v[6] = {0, 0, 0, 0, 0, 0};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k);
    const double yy[3] = {y.x, y.y, y.z}, mx[3] = {m0, m1, m2};
    double u[3], l1[3], l2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      u[c] = 0.95 * fabs(yy[c]) + 0.10 * mx[c];
      const double f1 = -yy[c] - u[c], f2 = yy[c] - u[c];
      l1[c] = -1.0 / f1; l2[c] = -1.0 / f2;
      v[c] += -(f1 * l1[c] + f2 * l2[c]);                               // sdg (:264)
      const double rm = 1.0 - l1[c] - l2[c];                             // rdual, m part (:272-276)
      v[3 + c] += rm * rm;
    }
    st256(U + k, make_double4(u[0], u[1], u[2], 0.0));
    st256(AX + k, make_double4(0, 0, 0, 0));
    st256(L1 + k, make_double4(l1[0], l1[1], l1[2], 0.0));
    st256(L2 + k, make_double4(l2[0], l2[1], l2[2], 0.0));
    st256(EV + k, make_double4(l1[0] - l2[0], l1[1] - l2[1], l1[2] - l2[2], 0.0));
  }
  if (grid_reduce_op<6, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) {
      ctl->sdg[c] = v[c];
      ctl->tau[c] = kPdMu * ctl->m2 / v[c];                              // :265
      ctl->rd2_edges[c] = v[3 + c];
      ctl->active[c] = 1; ctl->stuck[c] = 0; ctl->suff[c] = 0; ctl->backiter[c] = 0; ctl->s[c] = 0.0;
    }
    ctl->pditer = 0;
    ctl->pending = 0;
  }
}

// OUT = A^T EV on make_A's pattern (flagged entries skipped), thread per row on the SELL layout;
// optionally |OUT|^2 per coordinate into ctl->rd2_nodes.
__global__ void __launch_bounds__(256)
k_pd_At(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
        const int* __restrict__ sell_eid, const double4* __restrict__ EV, double4* __restrict__ OUT, int nslices,
        int want_norm, PdCtl* ctl, double* partials) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double v[3] = {0, 0, 0};
  for (int s = blockIdx.x + gridDim.x * warp; s < nslices; s += gridDim.x * wpb) {
    const int row = sell_row[s * kSellC + lane];
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    double ax = 0, ay = 0, az = 0;
#pragma unroll 4
    for (int j = 0; j < width; ++j) {
      const int eid = sell_eid[base + (int64_t)j * kSellC];
      if (eid != kSellPad) {
        const bool neg = eid < 0;
        const int kk = neg ? ~eid : eid;
        if (!(kk & kEidQuirk)) {
          const double4 e = ldg256(EV + (kk & kEidMask));
          if (neg) { ax -= e.x; ay -= e.y; az -= e.z; } else { ax += e.x; ay += e.y; az += e.z; }
        }
      }
    }
    if (row >= 0) {
      st256(OUT + row, make_double4(ax, ay, az, 0.0));
      v[0] += ax * ax; v[1] += ay * ay; v[2] += az * az;
    }
  }
  if (want_norm) {
    if (grid_reduce_op<3, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0)
      for (int c = 0; c < 3; ++c) ctl->rd2_nodes[c] = v[c];
  }
}

// |rcent|^2 with the CURRENT tau, then resnorm = sqrt(|rdual|^2 + |rcent|^2)   (:267-281, :450-458);
// also closes a Newton step: pditer, done test (:460).
__global__ void __launch_bounds__(256)
k_pd_rcent(const double4* __restrict__ Y, const double4* __restrict__ U, const double4* __restrict__ AX,
           const double4* __restrict__ L1, const double4* __restrict__ L2, int64_t m, int closing, int pdmaxiter,
           PdCtl* ctl, double* partials) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  const double it0 = 1.0 / ctl->tau[0], it1 = 1.0 / ctl->tau[1], it2 = 1.0 / ctl->tau[2];
  double v[3] = {0, 0, 0};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k), u = ldg256(U + k), ax = ldg256(AX + k), l1 = ldg256(L1 + k), l2 = ldg256(L2 + k);
    const double yy[3] = {y.x, y.y, y.z}, uu[3] = {u.x, u.y, u.z}, aa[3] = {ax.x, ax.y, ax.z};
    const double a1[3] = {l1.x, l1.y, l1.z}, a2[3] = {l2.x, l2.y, l2.z}, it[3] = {it0, it1, it2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double f1 = aa[c] - yy[c] - uu[c], f2 = -aa[c] + yy[c] - uu[c];
      const double r1 = -a1[c] * f1 - it[c], r2 = -a2[c] * f2 - it[c];
      v[c] += r1 * r1 + r2 * r2;
    }
  }
  if (grid_reduce_op<3, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    if (closing) ctl->pditer += 1;
    for (int c = 0; c < 3; ++c) {
      if (!closing || ctl->active[c]) {
        ctl->rc2[c] = v[c];
        ctl->resnorm[c] = sqrt(ctl->rd2_nodes[c] + ctl->rd2_edges[c] + v[c]);
      }
      if (closing && ctl->active[c] && (ctl->stuck[c] || ctl->sdg[c] < kPdTol || ctl->pditer >= pdmaxiter))
        ctl->active[c] = 0;                                              // :460 / :423-428
    }
  }
}

// Newton set-up per edge (:293-306): sigx -> SIGX (weights of H), EV = -(1/tau)(-1/fu1 + 1/fu2) - (sig2/sig1) w2
// so that w1p = A^T EV.
__global__ void __launch_bounds__(256)
k_pd_prep(const double4* __restrict__ Y, const double4* __restrict__ U, const double4* __restrict__ AX,
          const double4* __restrict__ L1, const double4* __restrict__ L2, double4* __restrict__ SIGX,
          double4* __restrict__ EV, int64_t m, const PdCtl* ctl) {
  const double it[3] = {1.0 / ctl->tau[0], 1.0 / ctl->tau[1], 1.0 / ctl->tau[2]};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k), u = ldg256(U + k), ax = ldg256(AX + k), l1 = ldg256(L1 + k), l2 = ldg256(L2 + k);
    const double yy[3] = {y.x, y.y, y.z}, uu[3] = {u.x, u.y, u.z}, aa[3] = {ax.x, ax.y, ax.z};
    const double a1[3] = {l1.x, l1.y, l1.z}, a2[3] = {l2.x, l2.y, l2.z};
    double sx[3], ev[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double f1 = aa[c] - yy[c] - uu[c], f2 = -aa[c] + yy[c] - uu[c];
      const double w2 = -1.0 - it[c] * (1.0 / f1 + 1.0 / f2);
      const double sig1 = -a1[c] / f1 - a2[c] / f2, sig2 = a1[c] / f1 - a2[c] / f2;
      sx[c] = sig1 - sig2 * sig2 / sig1;
      ev[c] = -it[c] * (-1.0 / f1 + 1.0 / f2) - (sig2 / sig1) * w2;
    }
    st256(SIGX + k, make_double4(sx[0], sx[1], sx[2], 0.0));
    st256(EV + k, make_double4(ev[0], ev[1], ev[2], 0.0));
  }
}

// Newton matrix H = A'^T diag(sigx) A' on make_AtA's pattern: per-entry weights (3 per entry) and diagonal.
__global__ void __launch_bounds__(256)
k_sell_hweights(const int* __restrict__ sell_row, const int* __restrict__ slice_off, const int* __restrict__ slice_width,
                const int* __restrict__ sell_eid, const double4* __restrict__ SIGX, double4* __restrict__ sell_w3,
                double4* __restrict__ diag3, int nslices) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int s = blockIdx.x + gridDim.x * warp; s < nslices; s += gridDim.x * wpb) {
    const int row = sell_row[s * kSellC + lane];
    const int width = slice_width[s];
    const int64_t base = (int64_t)slice_off[s] + lane;
    double dx = 0, dy = 0, dz = 0;
#pragma unroll 4
    for (int j = 0; j < width; ++j) {
      const int64_t o = base + (int64_t)j * kSellC;
      const int eid = sell_eid[o];
      double4 w = make_double4(0, 0, 0, 0);
      if (eid != kSellPad) {
        const int kk = eid < 0 ? ~eid : eid;
        w = ldg256(SIGX + (kk & kEidMask));
        dx += w.x; dy += w.y; dz += w.z;
      }
      st256(sell_w3 + o, w);
    }
    if (row >= 0) st256(diag3 + row, make_double4(dx, dy, dz, 0.0));
  }
}

// Search direction per edge (:324-381): Adx (make_A's mask), du, dlamu1, dlamu2, EV = dlamu1 - dlamu2,
// and the largest feasible step per coordinate.
__global__ void __launch_bounds__(256)
k_pd_direction(const int2* __restrict__ I, int f, const double4* __restrict__ DX, const double4* __restrict__ Y,
               const double4* __restrict__ U, const double4* __restrict__ AX, const double4* __restrict__ L1,
               const double4* __restrict__ L2, double4* __restrict__ ADX, double4* __restrict__ DU,
               double4* __restrict__ DL1, double4* __restrict__ DL2, double4* __restrict__ EV, int64_t m, PdCtl* ctl,
               double* partials) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  const double it[3] = {1.0 / ctl->tau[0], 1.0 / ctl->tau[1], 1.0 / ctl->tau[2]};
  double v[3] = {INFINITY, INFINITY, INFINITY};
  IRA_EDGE_LOOP(k, m) {
    const int2 e = I[k];
    double adx[3] = {0, 0, 0};
    if (e.y >= f) {                                                       // row k of A is empty when j is fixed
      const double4 dj = ldg256(DX + e.y);
      adx[0] = dj.x; adx[1] = dj.y; adx[2] = dj.z;
      if (e.x >= f) { const double4 di = ldg256(DX + e.x); adx[0] -= di.x; adx[1] -= di.y; adx[2] -= di.z; }
    }
    const double4 y = ldg256(Y + k), u = ldg256(U + k), ax = ldg256(AX + k), l1 = ldg256(L1 + k), l2 = ldg256(L2 + k);
    const double yy[3] = {y.x, y.y, y.z}, uu[3] = {u.x, u.y, u.z}, aa[3] = {ax.x, ax.y, ax.z};
    const double a1[3] = {l1.x, l1.y, l1.z}, a2[3] = {l2.x, l2.y, l2.z};
    double du[3], d1[3], d2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double f1 = aa[c] - yy[c] - uu[c], f2 = -aa[c] + yy[c] - uu[c];
      const double w2 = -1.0 - it[c] * (1.0 / f1 + 1.0 / f2);
      const double sig1 = -a1[c] / f1 - a2[c] / f2, sig2 = a1[c] / f1 - a2[c] / f2;
      du[c] = (w2 - sig2 * adx[c]) / sig1;                                                       // :327
      d1[c] = -(a1[c] / f1) * (adx[c] - du[c]) - a1[c] - it[c] / f1;                             // :330-333
      d2[c] = (a2[c] / f2) * (adx[c] + du[c]) - a2[c] - it[c] / f2;                              // :336-339
      double s = v[c];
      if (d1[c] < 0.0) s = fmin(s, -a1[c] / d1[c]);                                              // :348-361
      if (d2[c] < 0.0) s = fmin(s, -a2[c] / d2[c]);
      const double p1 = adx[c] - du[c], p2 = -adx[c] - du[c];                                    // :365-380
      if (p1 > 0.0) s = fmin(s, -f1 / p1);
      if (p2 > 0.0) s = fmin(s, -f2 / p2);
      v[c] = s;
    }
    st256(ADX + k, make_double4(adx[0], adx[1], adx[2], 0.0));
    st256(DU + k, make_double4(du[0], du[1], du[2], 0.0));
    st256(DL1 + k, make_double4(d1[0], d1[1], d1[2], 0.0));
    st256(DL2 + k, make_double4(d2[0], d2[1], d2[2], 0.0));
    st256(EV + k, make_double4(d1[0] - d2[0], d1[1] - d2[1], d1[2] - d2[2], 0.0));
  }
  if (grid_reduce_op<3, 2>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) {
      ctl->s[c] = ctl->active[c] ? 0.99 * fmin(1.0, v[c]) : 0.0;          // :347, :381
      ctl->suff[c] = ctl->active[c] ? 0 : 1;
      ctl->backiter[c] = 0;
    }
    ctl->pending = 1;
  }
}

// One back-tracking evaluation, edge part (:394-416): residual norms of the trial point x + s dx.
__global__ void __launch_bounds__(256)
k_pd_trial_edges(const double4* __restrict__ Y, const double4* __restrict__ U, const double4* __restrict__ AX,
                 const double4* __restrict__ L1, const double4* __restrict__ L2, const double4* __restrict__ ADX,
                 const double4* __restrict__ DU, const double4* __restrict__ DL1, const double4* __restrict__ DL2,
                 int64_t m, PdCtl* ctl, double* partials, double* trial /* [6] */) {
  __shared__ double sm[6 * 32];
  __shared__ int flag;
  const double ss[3] = {ctl->s[0], ctl->s[1], ctl->s[2]};
  const double it[3] = {1.0 / ctl->tau[0], 1.0 / ctl->tau[1], 1.0 / ctl->tau[2]};
  double v[6] = {0, 0, 0, 0, 0, 0};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k), u = ldg256(U + k), ax = ldg256(AX + k), l1 = ldg256(L1 + k), l2 = ldg256(L2 + k);
    const double4 adx = ldg256(ADX + k), du = ldg256(DU + k), d1 = ldg256(DL1 + k), d2 = ldg256(DL2 + k);
    const double yy[3] = {y.x, y.y, y.z}, uu[3] = {u.x, u.y, u.z}, aa[3] = {ax.x, ax.y, ax.z};
    const double a1[3] = {l1.x, l1.y, l1.z}, a2[3] = {l2.x, l2.y, l2.z};
    const double ad[3] = {adx.x, adx.y, adx.z}, dd[3] = {du.x, du.y, du.z};
    const double e1[3] = {d1.x, d1.y, d1.z}, e2[3] = {d2.x, d2.y, d2.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double up = uu[c] + ss[c] * dd[c], axp = aa[c] + ss[c] * ad[c];
      const double l1p = a1[c] + ss[c] * e1[c], l2p = a2[c] + ss[c] * e2[c];
      const double f1 = axp - yy[c] - up, f2 = -axp + yy[c] - up;
      const double rm = 1.0 - l1p - l2p;                                  // rdp, m part (:408-410)
      const double r1 = -l1p * f1 - it[c], r2 = -l2p * f2 - it[c];        // rcp (:413-416)
      v[c] += rm * rm;
      v[3 + c] += r1 * r1 + r2 * r2;
    }
  }
  if (grid_reduce_op<6, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0)
    for (int c = 0; c < 6; ++c) trial[c] = v[c];
}

// Node part of the evaluation (|Atv + s Atdv|^2) and the decision (:419-428).
__global__ void __launch_bounds__(256)
k_pd_trial_nodes(const double4* __restrict__ ATV, const double4* __restrict__ ATDV, int n, PdCtl* ctl, double* partials,
                 const double* trial) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  const double ss[3] = {ctl->s[0], ctl->s[1], ctl->s[2]};
  double v[3] = {0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double4 a = ldg256(ATV + i), d = ldg256(ATDV + i);
    const double t0 = a.x + ss[0] * d.x, t1 = a.y + ss[1] * d.y, t2 = a.z + ss[2] * d.z;
    v[0] += t0 * t0; v[1] += t1 * t1; v[2] += t2 * t2;
  }
  if (grid_reduce_op<3, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    int pending = 0;
    for (int c = 0; c < 3; ++c) {
      if (!ctl->active[c] || ctl->suff[c]) continue;
      const double nrm = sqrt(v[c] + trial[c] + trial[3 + c]);
      if (nrm <= (1.0 - kPdAlpha * ctl->s[c]) * ctl->resnorm[c]) {
        ctl->suff[c] = 1;
        ctl->rd2_nodes[c] = v[c];            // rdual = rdp (:455)
        ctl->rd2_edges[c] = trial[c];
      } else {
        ctl->s[c] *= kPdBeta;
        ctl->backiter[c] += 1;
        if (ctl->backiter[c] > 32) {         // "Stuck backtracking, returning last iterate." -> xp = x
          ctl->stuck[c] = 1; ctl->suff[c] = 1; ctl->s[c] = 0.0;
        } else {
          pending = 1;
        }
      }
    }
    ctl->pending = pending;
  }
}

// Accept the step (:432-448): state += s * direction; new surrogate duality gap and tau.
__global__ void __launch_bounds__(256)
k_pd_accept_edges(const double4* __restrict__ Y, double4* __restrict__ U, double4* __restrict__ AX,
                  double4* __restrict__ L1, double4* __restrict__ L2, const double4* __restrict__ ADX,
                  const double4* __restrict__ DU, const double4* __restrict__ DL1, const double4* __restrict__ DL2,
                  int64_t m, PdCtl* ctl, double* partials) {
  __shared__ double sm[3 * 32];
  __shared__ int flag;
  const double ss[3] = {ctl->active[0] ? ctl->s[0] : 0.0, ctl->active[1] ? ctl->s[1] : 0.0, ctl->active[2] ? ctl->s[2] : 0.0};
  double v[3] = {0, 0, 0};
  IRA_EDGE_LOOP(k, m) {
    const double4 y = ldg256(Y + k);
    double4 u = ld256(U + k), ax = ld256(AX + k), l1 = ld256(L1 + k), l2 = ld256(L2 + k);
    const double4 adx = ldg256(ADX + k), du = ldg256(DU + k), d1 = ldg256(DL1 + k), d2 = ldg256(DL2 + k);
    u.x += ss[0] * du.x; u.y += ss[1] * du.y; u.z += ss[2] * du.z;
    ax.x += ss[0] * adx.x; ax.y += ss[1] * adx.y; ax.z += ss[2] * adx.z;
    l1.x += ss[0] * d1.x; l1.y += ss[1] * d1.y; l1.z += ss[2] * d1.z;
    l2.x += ss[0] * d2.x; l2.y += ss[1] * d2.y; l2.z += ss[2] * d2.z;
    st256(U + k, u); st256(AX + k, ax); st256(L1 + k, l1); st256(L2 + k, l2);
    const double yy[3] = {y.x, y.y, y.z}, uu[3] = {u.x, u.y, u.z}, aa[3] = {ax.x, ax.y, ax.z};
    const double a1[3] = {l1.x, l1.y, l1.z}, a2[3] = {l2.x, l2.y, l2.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double f1 = aa[c] - yy[c] - uu[c], f2 = -aa[c] + yy[c] - uu[c];
      v[c] += -(f1 * a1[c] + f2 * a2[c]);
    }
  }
  if (grid_reduce_op<3, 0>(v, partials, &ctl->ticket, sm, &flag) && threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) {
      if (ctl->active[c]) {
        ctl->sdg[c] = v[c];                                              // :446
        ctl->tau[c] = kPdMu * ctl->m2 / v[c];                            // :448
      }
    }
  }
}

__global__ void __launch_bounds__(256)
k_pd_accept_nodes(double4* __restrict__ X, double4* __restrict__ ATV, const double4* __restrict__ DX,
                  const double4* __restrict__ ATDV, int n, const PdCtl* ctl) {
  const double ss[3] = {ctl->active[0] ? ctl->s[0] : 0.0, ctl->active[1] ? ctl->s[1] : 0.0, ctl->active[2] ? ctl->s[2] : 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double4 x = ld256(X + i), a = ld256(ATV + i);
    const double4 dx = ldg256(DX + i), d = ldg256(ATDV + i);
    x.x += ss[0] * dx.x; x.y += ss[1] * dx.y; x.z += ss[2] * dx.z;
    a.x += ss[0] * d.x; a.y += ss[1] * d.y; a.z += ss[2] * d.z;
    st256(X + i, x); st256(ATV + i, a);
  }
}

// ---- persistent PCG with one weight set per coordinate (the Newton solves) ------------------------------
// Same Chronopoulos-Gear structure as k_pcg_persistent (HBM-resident vectors, Jacobi), but every SELL slot
// carries three weights and every row three diagonals.
struct PcgW3Params {
  int n, nslices, max_iters;
  double rtol2;
  const int* sell_row; const int* slice_off; const int* slice_width; const int* sell_col;
  const double4* sell_w3; const double4* diag3; const double4* B;
  double4 *X, *R, *U, *W, *P, *S, *DINV;
  double* partials;
  Ctl* ctl;
};

__global__ void __launch_bounds__(kPcgThreads, 1)
k_pcg_persistent_w3(const PcgW3Params p) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double red[kPcgNV * 32];
  __shared__ double tot[kPcgNV];
  __shared__ double sc_bb[3], sc_go[3], sc_ao[3], sc_a[3], sc_b[3], sc_rr[3];
  __shared__ int sc_stop;
  const int lane = threadIdx.x & 31;
  const int gwarp = blockIdx.x + gridDim.x * (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  double v[kPcgNV];
#pragma unroll
  for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
  for (int s = gwarp; s < p.nslices; s += nwarps) {
    const int row = p.sell_row[s * kSellC + lane];
    if (row >= 0) {
      const double4 b = ldg256(p.B + row), d = ldg256(p.diag3 + row);
      const double4 di = make_double4(d.x > 0.0 ? 1.0 / d.x : 0.0, d.y > 0.0 ? 1.0 / d.y : 0.0, d.z > 0.0 ? 1.0 / d.z : 0.0, 0.0);
      st256(p.DINV + row, di);
      const double4 z4 = make_double4(0, 0, 0, 0);
      st256(p.X + row, z4); st256(p.P + row, z4); st256(p.S + row, z4);
      st256(p.R + row, b);
      st256(p.U + row, make_double4(di.x * b.x, di.y * b.y, di.z * b.z, 0.0));
      v[0] += b.x * b.x; v[1] += b.y * b.y; v[2] += b.z * b.z;
    }
  }
  pcg_grid_reduce(v, p.partials, grid, red, tot);
  if (threadIdx.x == 0) {
    for (int c = 0; c < 3; ++c) { sc_bb[c] = v[c]; sc_rr[c] = v[c]; sc_go[c] = 1.0; sc_ao[c] = 1.0; }
    sc_stop = !(v[0] > 0.0 || v[1] > 0.0 || v[2] > 0.0);
  }
  __syncthreads();
  int it = 0;
  while (!sc_stop) {
#pragma unroll
    for (int k = 0; k < kPcgNV; ++k) v[k] = 0.0;
    for (int s = gwarp; s < p.nslices; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      const int width = p.slice_width[s];
      const int64_t base = (int64_t)p.slice_off[s] + lane;
      const double4 u = row >= 0 ? ld256(p.U + row) : make_double4(0, 0, 0, 0);
      double ax = 0, ay = 0, az = 0;
      for (int j = 0; j < width; j += 4) {                              // widths are multiples of 4
        int c[4]; double4 w3[4], uc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int64_t o = base + (int64_t)(j + q) * kSellC;
          c[q] = __ldg(p.sell_col + o);
          w3[q] = ldg256(p.sell_w3 + o);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) uc[q] = ld256(p.U + c[q]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ax += w3[q].x * (u.x - uc[q].x); ay += w3[q].y * (u.y - uc[q].y); az += w3[q].z * (u.z - uc[q].z);
        }
      }
      if (row >= 0) {
        st256(p.W + row, make_double4(ax, ay, az, 0.0));
        const double4 r = ld256(p.R + row);
        v[0] += r.x * u.x; v[1] += r.y * u.y; v[2] += r.z * u.z;
        v[3] += u.x * ax;  v[4] += u.y * ay;  v[5] += u.z * az;
        v[6] += r.x * r.x; v[7] += r.y * r.y; v[8] += r.z * r.z;
      }
    }
    pcg_grid_reduce(v, p.partials, grid, red, tot);
    pcg_coefficients(tot, it, p.max_iters, p.rtol2, sc_bb, sc_go, sc_ao, sc_a, sc_b, sc_rr, &sc_stop);
    if (sc_stop) break;
    const double a0 = sc_a[0], a1 = sc_a[1], a2 = sc_a[2], b0 = sc_b[0], b1 = sc_b[1], b2 = sc_b[2];
    for (int s = gwarp; s < p.nslices; s += nwarps) {
      const int row = p.sell_row[s * kSellC + lane];
      if (row >= 0) {
        const double4 u = ld256(p.U + row), w = ld256(p.W + row), di = ld256(p.DINV + row);
        double4 pp = ld256(p.P + row), ss = ld256(p.S + row);
        pp.x = u.x + b0 * pp.x; pp.y = u.y + b1 * pp.y; pp.z = u.z + b2 * pp.z;
        ss.x = w.x + b0 * ss.x; ss.y = w.y + b1 * ss.y; ss.z = w.z + b2 * ss.z;
        st256(p.P + row, pp); st256(p.S + row, ss);
        double4 x = ld256(p.X + row), r = ld256(p.R + row);
        x.x += a0 * pp.x; x.y += a1 * pp.y; x.z += a2 * pp.z;
        r.x -= a0 * ss.x; r.y -= a1 * ss.y; r.z -= a2 * ss.z;
        st256(p.X + row, x); st256(p.R + row, r);
        st256(p.U + row, make_double4(di.x * r.x, di.y * r.y, di.z * r.z, 0.0));
      }
    }
    ++it;
    grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    p.ctl->cg_iters = it;
    for (int c = 0; c < 3; ++c) { p.ctl->bnorm2[c] = sc_bb[c]; p.ctl->rnorm2[c] = sc_rr[c]; }
    p.ctl->done = 1;
  }
}

}  // namespace ira
