// Reads a RAL-format graph with ALL n absolute rotations given, then runs the reference's call
// sequence of ral/test.cpp:288-302 minus init_mst/l1ra through the C++ adapter
// (irotavg_b200/host/l1_irls.hpp):  make_A -> irls -> quat_normalised, and writes the CLI's output
// format (n rows "w x y z", then m weights; ral/test.cpp:314-326).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>

#include "l1_irls.hpp"

using namespace irotavg;

int main(int argc, char** argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: adapter_main in out cost sigma_rad max_iters [change_th]\n"); return 2; }
  std::ifstream in(argv[1]);
  int m, n, f;
  in >> m >> n >> f;
  I_t I; I.reserve(m);
  Mat QQ = Mat::Zero(m, 4), Q = Mat::Zero(n, 4);
  for (int k = 0; k < m; ++k) {
    int e1, e2; double w, x, y, z;
    in >> e1 >> e2 >> w >> x >> y >> z;
    I.push_back(std::make_pair(e1, e2));
    QQ(k, 0) = x; QQ(k, 1) = y; QQ(k, 2) = z; QQ(k, 3) = w;       // file [w x y z] -> [x y z w] (ral/test.cpp:193)
  }
  for (int i = 0; i < n; ++i) {
    double w, x, y, z;
    in >> w >> x >> y >> z;
    Q(i, 0) = x; Q(i, 1) = y; Q(i, 2) = z; Q(i, 3) = w;
  }
  const Cost cost = (Cost)std::atoi(argv[3]);
  const double sigma = std::atof(argv[4]);
  const int iters = std::atoi(argv[5]);
  const double th = argc > 6 ? std::atof(argv[6]) : 1e-3;
  SpMat A = make_A(n, f, I);
  Vec weights(m);
  int iters_out = 0; double runtime = 0.0;
  irls(QQ, I, A, cost, sigma, Q, f, iters, th, weights, iters_out, runtime);
  quat_normalised(Q, f);
  std::ofstream out(argv[2]);
  out << std::setprecision(17);
  out << iters_out << "\n";
  for (int i = 0; i < n; ++i) out << Q(i, 3) << " " << Q(i, 0) << " " << Q(i, 1) << " " << Q(i, 2) << "\n";
  for (int k = 0; k < m; ++k) out << weights(k) << "\n";
  return 0;
}
