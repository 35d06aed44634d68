// Replays an op list (oracle/rotavg_stream.py:write_ops) through the REFERENCE'S OWN ViewGraph::rotAvg source
// (compiled from the text tools/make_dropin.py extracts out of /root/reference/src/ViewGraph.cpp) on top of the
// adapter header irotavg_b200/host/l1_irls.hpp + libira.so: the literal drop-in.  Output: like rotavg_main.cpp
// (every view's rotation, row-major, 17 digits), then one line per call "window-size wall-seconds", then the
// reference's savePoses() text for the same views.
//   rotavg_refsrc ops.txt out.txt poses.txt
// The same file linked against oracle/_ref (the reference's ral/l1_irls.cpp itself, oracle/build_ref.py) instead of
// the adapter + libira.so is `oracle/_ref/rotavg_reference`: the reference's complete CPU path for config 5.
#include <chrono>
#include <cstdio>
#include <string>

#include "viewgraph_decl.hpp"

using namespace irotavg;

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: rotavg_refsrc ops.txt out.txt poses.txt\n"); return 2; }
  std::ifstream in(argv[1]);
  if (!in.is_open()) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  long nops = 0;
  in >> nops;
  ViewGraph vg;
  std::vector<int> calls;
  std::vector<double> secs;
  for (long k = 0; k < nops; ++k) {
    std::string op;
    in >> op;
    if (op == "V") {
      Frame fr((int)vg.m_views.size());
      vg.m_views.push_back(new View(fr));
      vg.m_fixed_mask.push_back(false);
    } else if (op == "E" || op == "F") {
      int a = 0, b = 0;
      in >> a;
      if (op == "E") in >> b;
      double R[9];
      for (int q = 0; q < 9; ++q) in >> R[q];
      Pose::Mat3 Rm(R);
      Pose p(Rm, Pose::Vec3(0, 0, 0));
      if (op == "E") View::connect(*vg.m_views[a], *vg.m_views[b], FeatureMatches(), p);
      else vg.fixPose(a, p);
    } else if (op == "A") {
      int win = 0;
      in >> win;
      const auto t0 = std::chrono::steady_clock::now();
      vg.rotAvg(win);
      secs.push_back(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
      calls.push_back(win);
    } else {
      std::fprintf(stderr, "bad op '%s'\n", op.c_str());
      return 2;
    }
  }
  std::ofstream out(argv[2]);
  out << std::setprecision(17);
  out << vg.m_views.size() << " " << calls.size() << "\n";
  for (size_t v = 0; v < vg.m_views.size(); ++v) {
    const Pose::Mat3& R = vg.m_views[v]->pose().R();
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out << R(r, c) << (r == 2 && c == 2 ? "\n" : " ");
  }
  for (size_t k = 0; k < calls.size(); ++k) out << calls[k] << " " << secs[k] << "\n";
  out.close();
  vg.savePoses(argv[3]);
  return 0;
}
