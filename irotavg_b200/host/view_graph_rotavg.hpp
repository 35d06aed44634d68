// view_graph_rotavg.hpp - host side of ViewGraph::rotAvg (src/ViewGraph.cpp:1263-1435) over the C ABI.
//
// The reference's containers stay as they are: `std::vector<View*> m_views`, `std::vector<bool>
// m_fixed_mask` (src/ViewGraph.hpp:128-131), `View` with its `std::map<View*, ViewConnection*>`
// adjacency (src/View.hpp:66,146) and `Pose` (cv::Matx33d rotation, src/Pose.hpp:35-59).  This header is
// a template over those types, so it compiles against the reference's own src/View.hpp + src/Pose.hpp
// (OpenCV build) and against the OpenCV-free stand-ins the tests use (tests/cpp/view_shim.hpp).
//
//   ira_b200::rot_avg(m_views, m_fixed_mask, winSize)   ==   ViewGraph::rotAvg(winSize)
//
// Steps, as in the reference: (1) collect the edges (i < j) of the last winSize views (:1282-1307);
// (2) skip when edges or vertices are fewer than the window (:1313-1321); (3) f = vertices outside the
// window + fixed views inside (:1329-1338); re-index fixed-first, each group ascending by frame id
// (:1340-1363); (4) rotations -> quaternions with rmat2quat's branch rule (:1175-1203); with f == 0 row 0
// becomes the identity and f = 1 (:1382-1386); (5) l1ra(100, 1e-3) then irls(Geman-McClure, 5 deg, 100,
// 1e-3) (:1402-1417) - ONE call, ira_l1ra_irls, the rotations stay on the device between the stages;
// (6) normalise and write the free rotations back as 3x3 matrices (:1420-1434).
#ifndef IROTAVG_B200_VIEW_GRAPH_ROTAVG_HPP_
#define IROTAVG_B200_VIEW_GRAPH_ROTAVG_HPP_

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <type_traits>
#include <vector>

#include "ira.h"

namespace ira_b200 {

// 3x3 rotation (anything indexable as R(r, c)) -> [x y z w]; branch rule of src/ViewGraph.cpp:1175-1203.
template <class Mat3T>
inline void rmat2quat(const Mat3T& R, double q[4]) {
  const double trace = R(0, 0) + R(1, 1) + R(2, 2);
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    q[3] = s * 0.5;
    s = 0.5 / s;
    q[0] = (R(2, 1) - R(1, 2)) * s;
    q[1] = (R(0, 2) - R(2, 0)) * s;
    q[2] = (R(1, 0) - R(0, 1)) * s;
  } else {
    const int i = R(0, 0) < R(1, 1) ? (R(1, 1) < R(2, 2) ? 2 : 1) : (R(0, 0) < R(2, 2) ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
    q[i] = s * 0.5;
    s = 0.5 / s;
    q[3] = (R(k, j) - R(j, k)) * s;
    q[j] = (R(j, i) + R(i, j)) * s;
    q[k] = (R(k, i) + R(i, k)) * s;
  }
}

// normalised [x y z w] -> row-major 3x3 (Eigen's Quaterniond::toRotationMatrix, :1426-1431)
inline void quat2rmat_rowmajor(const double qin[4], double R[9]) {
  const double nrm = std::sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double x = qin[0], y = qin[1], z = qin[2], w = qin[3];
  if (nrm > 0.0) { x /= nrm; y /= nrm; z /= nrm; w /= nrm; }
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// ViewGraph::savePoses (src/ViewGraph.cpp:1206-1231): one line per view, "id \t qw \t qx \t qy \t qz \t tx \t ty \t tz",
// quaternion from rmat2quat of the pose's rotation, 17 significant digits, scientific.  Returns false (after the
// reference's "Unable to save results." on std::cerr) when the file cannot be opened.
template <class ViewT>
inline bool save_poses(const std::vector<ViewT*>& views, const std::string& filename) {
  std::ofstream fs(filename);
  if (!fs.is_open()) { std::cerr << "Unable to save results." << std::endl; return false; }
  for (ViewT* view : views) {
    const auto& pose = view->pose();
    const auto& t = pose.t();
    double q[4];
    rmat2quat(pose.R(), q);
    fs << view->frame().id() << "\t";
    fs << std::setprecision(17) << std::scientific << q[3] << "\t" << q[0] << "\t" << q[1] << "\t" << q[2] << "\t";
    fs << std::setprecision(17) << std::scientific << t(0) << "\t" << t(1) << "\t" << t(2) << "\n";
  }
  return true;
}

struct RotAvgReport {
  bool solved = false;        // false: one of the reference's early returns was taken
  int vertices = 0, edges = 0, fixed = 0;
  int l1_iters = 0, irls_iters = 0;
  double seconds = 0.0;       // wall time of the library call (host buffers in, host buffers out)
};

struct RotAvgParams {         // the constants of src/ViewGraph.cpp:1402-1414
  int l1_iters = 100;
  int irls_iters = 100;
  double change_th = 0.001;
  int cost = IRA_COST_GEMAN_MCCLURE;
  double sigma = 5.0 * 3.14159265358979323846 / 180.0;
};

namespace detail_rotavg {
inline ira_handle& handle() {
  static ira_handle h = nullptr;
  if (!h) {
    ira_status s = ira_create(&h, nullptr);
    if (s != IRA_OK) {
      std::cerr << "irotavg-b200: " << ira_status_string(s) << std::endl;
      std::exit(-1);
    }
  }
  return h;
}
}  // namespace detail_rotavg

// ViewT: frame().id(), pose() -> PoseT&, connections() -> map<ViewT*, ConnT*>, ConnT::pose() -> const PoseT&.
// PoseT: R() -> Mat3 indexable (r, c); setR(Mat3); Mat3 constructible from a row-major double[9].
template <class ViewT>
inline RotAvgReport rot_avg(std::vector<ViewT*>& views, const std::vector<bool>& fixed_mask, int winSize,
                            const RotAvgParams& prm = RotAvgParams(), ira_handle h = nullptr) {
  RotAvgReport rep;
  const long nviews = (long)views.size();
  if ((long)winSize > nviews) winSize = (int)nviews;
  if (winSize < 2) return rep;                                            // nothing to optimise
  const long first_in_window = nviews - winSize;

  std::vector<int32_t> I;                                                 // (i, j) pairs, frame ids for now
  std::vector<double> qq;                                                 // 4 per edge, [x y z w]
  std::set<int> vertices;
  for (long t = first_in_window; t < nviews; ++t) {
    ViewT* view = views[t];
    const int j = view->frame().id();
    for (auto it = view->connections().begin(); it != view->connections().end(); ++it) {
      const int i = it->first->frame().id();
      if (i >= j) continue;                                               // each pair once, from its later view
      I.push_back(i);
      I.push_back(j);
      vertices.insert(i);
      vertices.insert(j);
      double q[4];
      rmat2quat(it->second->pose().R(), q);
      qq.insert(qq.end(), q, q + 4);
    }
  }
  const long m = (long)(qq.size() / 4), nv = (long)vertices.size();
  rep.edges = (int)m;
  rep.vertices = (int)nv;
  if (m < winSize || nv < winSize) return rep;                            // too few edges / unconnected

  int f = (int)nv - winSize;                                              // everything outside the window is fixed
  for (std::set<int>::const_iterator x = vertices.begin(); x != vertices.end(); ++x)
    if (*x >= first_in_window && fixed_mask[*x]) ++f;
  std::map<int, int> to_idx;
  std::vector<int> to_vertex((size_t)nv);
  int next_fixed = 0, next_free = f;
  for (std::set<int>::const_iterator x = vertices.begin(); x != vertices.end(); ++x) {
    const bool free_view = *x >= first_in_window && !fixed_mask[*x];
    const int idx = free_view ? next_free++ : next_fixed++;
    to_idx[*x] = idx;
    to_vertex[(size_t)idx] = *x;
  }
  for (size_t k = 0; k < I.size(); ++k) I[k] = to_idx[I[k]];

  std::vector<double> Q((size_t)nv * 4), QQ((size_t)m * 4), weights((size_t)m);   // column-major n x 4 / m x 4
  for (std::set<int>::const_iterator x = vertices.begin(); x != vertices.end(); ++x) {
    double q[4];
    rmat2quat(views[*x]->pose().R(), q);
    const int r = to_idx[*x];
    for (int c = 0; c < 4; ++c) Q[(size_t)c * nv + r] = q[c];
  }
  if (f == 0) {                                                           // gauge: first rotation = I
    Q[0] = 0; Q[(size_t)nv] = 0; Q[(size_t)2 * nv] = 0; Q[(size_t)3 * nv] = 1;
    f = 1;
  }
  rep.fixed = f;
  for (long k = 0; k < m; ++k)
    for (int c = 0; c < 4; ++c) QQ[(size_t)c * m + k] = qq[(size_t)4 * k + c];

  if (!h) h = detail_rotavg::handle();
  int32_t l1_out = 0, irls_out = 0;
  static ira_stats st;                                     // the reference's solves are exact: an unconverged one is reported
  ira_status s = ira_l1ra_irls(h, m, nv, f, I.data(), QQ.data(), m, Q.data(), nv, prm.l1_iters, prm.change_th,
                               prm.cost, prm.sigma, prm.irls_iters, prm.change_th, weights.data(), &l1_out,
                               &irls_out, &rep.seconds, &st);
  if (s != IRA_OK && s != IRA_ERR_NONFINITE) {
    std::cerr << "rotAvg failed: " << ira_status_string(s) << ": " << ira_last_error(h) << std::endl;
    std::exit(-1);
  }
  if (st.cg_hit_max > 0)
    std::cerr << "irotavg-b200 WARNING: rotAvg: " << st.cg_hit_max << " linear solve(s) stopped at the PCG iteration cap "
              << "without reaching cg_rtol; the step is inexact where the reference's is exact." << std::endl;
  rep.l1_iters = l1_out;
  rep.irls_iters = irls_out;
  rep.solved = true;

  typedef typename std::remove_reference<decltype(views[0]->pose())>::type PoseT;
  for (long k = f; k < nv; ++k) {                                         // write the window's poses back
    const double q[4] = {Q[(size_t)k], Q[(size_t)nv + k], Q[(size_t)2 * nv + k], Q[(size_t)3 * nv + k]};
    double R[9];
    quat2rmat_rowmajor(q, R);
    views[(size_t)to_vertex[(size_t)k]]->pose().setR(typename PoseT::Mat3(R));
  }
  return rep;
}

}  // namespace ira_b200
#endif  // IROTAVG_B200_VIEW_GRAPH_ROTAVG_HPP_
