// Replays an op list (oracle/rotavg_stream.py:write_ops) on the reference's containers (View / Pose /
// ViewConnection, here the OpenCV-free stand-ins of view_shim.hpp) through ira_b200::rot_avg - the host
// mirror of ViewGraph::rotAvg - and writes every view's rotation (row-major, 17 digits) followed by one
// line per rotAvg call: "win solved vertices edges fixed l1_iters irls_iters seconds".
//   rotavg_main ops.txt out.txt [poses.txt]     (poses.txt: ira_b200::save_poses, the layout of ViewGraph::savePoses)
#include <chrono>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <string>

#include "view_shim.hpp"
#include "view_graph_rotavg.hpp"

using namespace irotavg;

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: rotavg_main ops.txt out.txt\n"); return 2; }
  std::ifstream in(argv[1]);
  if (!in.is_open()) { std::fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  long nops = 0;
  in >> nops;
  std::vector<View*> views;                 // ViewGraph::m_views
  std::vector<bool> fixed_mask;             // ViewGraph::m_fixed_mask
  struct Call { int win; ira_b200::RotAvgReport rep; double wall; };
  std::vector<Call> calls;
  for (long k = 0; k < nops; ++k) {
    std::string op;
    in >> op;
    if (op == "V") {
      Frame fr((int)views.size());
      views.push_back(new View(fr));
      fixed_mask.push_back(false);
    } else if (op == "E" || op == "F") {
      int a = 0, b = 0;
      in >> a;
      if (op == "E") in >> b;
      double R[9];
      for (int q = 0; q < 9; ++q) in >> R[q];
      if (op == "E") View::connect(*views[a], *views[b], FeatureMatches(), Pose(Pose::Mat3(R), Pose::Vec3()));
      else { fixed_mask[a] = true; views[a]->pose() = Pose(Pose::Mat3(R), Pose::Vec3()); }   // ViewGraph::fixPose
    } else if (op == "A") {
      int win = 0;
      in >> win;
      const auto t0 = std::chrono::steady_clock::now();
      Call c;
      c.win = win;
      c.rep = ira_b200::rot_avg(views, fixed_mask, win);
      c.wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      calls.push_back(c);
    } else {
      std::fprintf(stderr, "bad op '%s'\n", op.c_str());
      return 2;
    }
  }
  std::ofstream out(argv[2]);
  out << std::setprecision(17);
  out << views.size() << " " << calls.size() << "\n";
  for (size_t v = 0; v < views.size(); ++v) {
    const Pose::Mat3& R = views[v]->pose().R();
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) out << R(r, c) << (r == 2 && c == 2 ? "\n" : " ");
  }
  for (size_t k = 0; k < calls.size(); ++k) {
    const ira_b200::RotAvgReport& r = calls[k].rep;
    out << calls[k].win << " " << (r.solved ? 1 : 0) << " " << r.vertices << " " << r.edges << " " << r.fixed << " "
        << r.l1_iters << " " << r.irls_iters << " " << calls[k].wall << "\n";
  }
  if (argc > 3 && !ira_b200::save_poses(views, argv[3])) return 3;
  return 0;
}
