// OpenCV-free stand-ins for the reference's containers, with the member API rotAvg touches:
//   cv::Matx33d / cv::Vec3d                       (OpenCV core types used by src/Pose.hpp:39-41)
//   irotavg::Pose                                 src/Pose.hpp:35-59   (R(), t(), setR(), setT(); identity start)
//   irotavg::Frame                                only id() (src/ViewGraph.cpp:1290 uses frame().id())
//   irotavg::View, View::ViewConnection, connect  src/View.hpp:43-147, src/ViewGraph.cpp:1438-1455
// A build with OpenCV includes the reference's own headers instead; irotavg_b200/host/view_graph_rotavg.hpp
// is a template over whichever is present.
#ifndef IROTAVG_B200_TESTS_VIEW_SHIM_HPP_
#define IROTAVG_B200_TESTS_VIEW_SHIM_HPP_

#include <map>
#include <utility>
#include <vector>

// OpenCV's core/cvdef.h macros (src/ViewGraph.cpp:1269 uses MIN)
#ifndef MIN
#define MIN(a, b) ((a) > (b) ? (b) : (a))
#endif
#ifndef MAX
#define MAX(a, b) ((a) < (b) ? (b) : (a))
#endif

namespace cv {
struct Matx33d {
  double val[9];
  Matx33d() { for (int k = 0; k < 9; ++k) val[k] = 0.0; }
  explicit Matx33d(const double* rowmajor) { for (int k = 0; k < 9; ++k) val[k] = rowmajor[k]; }
  static Matx33d eye() { Matx33d m; m.val[0] = m.val[4] = m.val[8] = 1.0; return m; }
  double operator()(int r, int c) const { return val[3 * r + c]; }
  double& operator()(int r, int c) { return val[3 * r + c]; }
};
struct Vec3d {
  double val[3];
  Vec3d(double a = 0, double b = 0, double c = 0) { val[0] = a; val[1] = b; val[2] = c; }
  double operator()(int i) const { return val[i]; }
};
struct Vec4d {
  double val[4];
  Vec4d(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double operator()(int i) const { return val[i]; }
  double& operator()(int i) { return val[i]; }
};
struct DMatch { int queryIdx, trainIdx; };
}  // namespace cv

namespace irotavg {

class Pose {
 public:
  typedef cv::Matx33d Mat3;
  typedef cv::Vec3d Vec3;
  typedef cv::Vec4d Vec4;
  Pose() : m_R(Mat3::eye()), m_t() {}
  Pose(Mat3 R, Vec3 t) : m_R(R), m_t(t) {}
  void setR(Mat3 R) { m_R = R; }
  void setT(Vec3 t) { m_t = t; }
  const Mat3& R() const { return m_R; }
  const Vec3& t() const { return m_t; }
 private:
  Mat3 m_R;
  Vec3 m_t;
};

class Frame {
 public:
  explicit Frame(int id) : m_id(id) {}
  int id() const { return m_id; }
 private:
  int m_id;
};

typedef std::vector<cv::DMatch> FeatureMatches;

class View {
 public:
  class ViewConnection {
   public:
    ViewConnection(View& v1, View& v2, FeatureMatches matches, Pose rel_pose)
        : m_v1(v1), m_v2(v2), m_matches(std::move(matches)), m_rel_pose(rel_pose) {}
    FeatureMatches& matches() { return m_matches; }
    size_t size() const { return m_matches.size(); }
    const Pose& pose() const { return m_rel_pose; }
   private:
    View& m_v1;
    View& m_v2;
    FeatureMatches m_matches;
    Pose m_rel_pose;
  };
  typedef std::map<View*, ViewConnection*> Connections;

  explicit View(Frame& frame) : m_frame(frame) {}
  Frame& frame() { return m_frame; }
  Pose& pose() { return m_pose; }
  bool isConnectedTo(const View& v) const { return m_connections.count(const_cast<View*>(&v)) > 0; }
  const Connections& connections() const { return m_connections; }

  // one connection object per unordered pair, registered in both views; false if already connected
  static bool connect(View& v1, View& v2, FeatureMatches matches, Pose rel_pose) {
    if (v1.m_connections.count(&v2) > 0) return false;
    ViewConnection* c = new ViewConnection(v1, v2, std::move(matches), rel_pose);
    v1.m_connections[&v2] = c;
    v2.m_connections[&v1] = c;
    return true;
  }

 private:
  Frame m_frame;
  Pose m_pose;
  Connections m_connections;
};

}  // namespace irotavg
#endif
