// Per-edge / per-node SO(3) arithmetic of the IRLS hot path, FP64, shared by every kernel.
//
// Restates (from scratch; quaternions are [x y z w]):
//   quat_mult      ral/l1_irls.cpp:99-105   Hamilton product
//   delta_rel      ral/l1_irls.cpp:109-127  p = q~_j (x) (QQ_k (x) Q_i),  q~_j = [x y z -w]
//   log_map        ral/l1_irls.cpp:498-532  theta = 2 atan2(|v|, w) wrapped to [-pi,pi), v*theta/|v|
//   exp_map        ral/l1_irls.cpp:471-492  [v sin(|v|/2)/|v|, cos(|v|/2)], non-finite -> 0
//   cost switch    ral/l1_irls.cpp:617-727  14 robust square-root weights
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace ira {

#define IRA_EPS 2.2204e-16            /* ral/l1_irls.hpp:40 */
#define IRA_PI 3.141592653589793238462643383279502884 /* EIGEN_PI */

enum Cost : int {  // ral/l1_irls.hpp:56-57
  kL2 = 0, kL1, kL15, kL05, kGemanMcClure, kHuber, kPseudoHuber, kAndrews, kBisquare, kCauchy,
  kFair, kLogistic, kTalwar, kWelsch, kNumCosts
};

// a (x) b
__host__ __device__ __forceinline__ double4 quat_mult(const double4 a, const double4 b) {
  double4 r;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  return r;
}

// Rotation-vector residual of one edge.  Returns (wx, wy, wz, theta).
__host__ __device__ __forceinline__ double4 edge_residual(const double4 qi, const double4 qq,
                                                          const double4 qj) {
  const double4 qjn = make_double4(qj.x, qj.y, qj.z, -qj.w);   // column w negated (:114-115)
  const double4 p = quat_mult(qjn, quat_mult(qq, qi));
  const double s = sqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  double theta = 2.0 * atan2(s, p.w);
  if (theta < -IRA_PI) theta += 2.0 * IRA_PI;                    // :510-517
  else if (theta >= IRA_PI) theta -= 2.0 * IRA_PI;
  double4 w;
  if (s < IRA_EPS) {                                             // :527-531
    w.x = w.y = w.z = 0.0;
  } else {
    const double aux = theta / s;
    w.x = p.x * aux; w.y = p.y * aux; w.z = p.z * aux;
  }
  w.w = theta;
  return w;
}

__host__ __device__ __forceinline__ double finite_or_zero(double v) { return isfinite(v) ? v : 0.0; }

// delta = Exp(v): [v sin(t/2)/t, cos(t/2)], t = |v|; non-finite entries -> 0 (so t = 0 gives (0,0,0,1)).
__host__ __device__ __forceinline__ double4 exp_quat(double vx, double vy, double vz, double* theta_out) {
  const double t = sqrt(vx * vx + vy * vy + vz * vz);
  double sn, cs;
  sincos(0.5 * t, &sn, &cs);
  const double k = sn / t;                                       // NaN when t == 0, as in the reference
  *theta_out = t;
  return make_double4(finite_or_zero(vx * k), finite_or_zero(vy * k), finite_or_zero(vz * k),
                      finite_or_zero(cs));
}

// New square-root weight from e2 = |E_k|^2.  `old` is the previous weight (L2 and Huber keep it).
__host__ __device__ __forceinline__ double robust_weight(int cost, double sigma, double e2, double old) {
  switch (cost) {
    case kL2: return old;                                                        // :619-620
    case kL05: { double w = 1.0 / pow(e2, 3.0 / 8.0); return w > 1e4 ? 1e4 : w; }      // :621-625
    case kL1:  { double w = 1.0 / sqrt(sqrt(e2)); return w > 1e4 ? 1e4 : w; }          // :626-630
    case kL15: { double w = 1.0 / sqrt(sqrt(sqrt(e2))); return w > 1e4 ? 1e4 : w; }    // :631-635
    case kGemanMcClure: return 1.0 / (e2 + sigma * sigma);                       // :636-642
    case kHuber: {                                                               // :643-651
      const double tun = 1.345 * sigma;
      const double e = sqrt(e2) / tun;
      return e >= 1.0 ? sqrt(1.0 / e) : old;
    }
    case kPseudoHuber: return 1.0 / sqrt(sqrt(1.0 + e2 / (sigma * sigma)));      // :652-658
    case kAndrews: {                                                             // :659-677
      const double tun = 1.339 * sigma;
      const double e = sqrt(e2) / tun;
      double w = sqrt(sin(e) / e);
      if (e >= IRA_PI) w = 0.0; else if (e < 1e-4) w = 1.0;
      if (w < 1e-4) w = 1e-4;
      return w;
    }
    case kBisquare: {                                                            // :678-684
      const double tun = 4.685 * sigma;
      const double w = 1.0 - e2 / (tun * tun);
      return w < 1e-4 ? 1e-4 : w;
    }
    case kCauchy: { const double tun = 2.385 * sigma; return 1.0 / sqrt(1.0 + e2 / (tun * tun)); }  // :685-691
    case kFair:   { const double tun = 1.400 * sigma; return 1.0 / sqrt(1.0 + sqrt(e2) / tun); }    // :692-698
    case kLogistic: {                                                            // :699-707
      const double tun = 1.205 * sigma;
      const double e = sqrt(e2) / tun;
      return e < 1e-4 ? 1.0 : sqrt(tanh(e) / e);
    }
    case kTalwar: { const double tun = 2.795 * sigma; return e2 < tun * tun ? 1.0001 : 0.0; }       // :708-714
    case kWelsch: {                                                              // :715-722
      const double tun = 2.985 * sigma;
      const double w = exp(-0.5 * e2 / (tun * tun));
      return w < 1e-4 ? 1e-4 : w;
    }
    default: return old;
  }
}

}  // namespace ira
