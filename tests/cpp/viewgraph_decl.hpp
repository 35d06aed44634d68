// Declarations that let the reference's OWN ViewGraph::rotAvg / rmat2quat / savePoses / fixPose source text
// (src/ViewGraph.cpp:1175-1435, extracted at build time by tools/make_dropin.py - never committed) compile
// outside the SLAM front end: the members those functions touch (src/ViewGraph.hpp:66-75,128-131) over the
// OpenCV-free containers of view_shim.hpp, and the adapter header in place of ral/l1_irls.hpp.
#ifndef IROTAVG_B200_TESTS_VIEWGRAPH_DECL_HPP_
#define IROTAVG_B200_TESTS_VIEWGRAPH_DECL_HPP_
#include <cassert>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "l1_irls.hpp"       // irotavg_b200/host: the drop-in for ral/l1_irls.hpp
#include "view_shim.hpp"

namespace irotavg {
class ViewGraph {
 public:
  void savePoses(const std::string& filename) const;
  void rotAvg(const int winSize);
  void fixPose(int idx, Pose& pose);
  bool isPoseFixed(int idx) const;
  int countFixedPoses() const;
  std::vector<View*> m_views;
  std::vector<bool> m_fixed_mask;
};
}  // namespace irotavg
using std::setprecision;
#endif
