// l1_irls - command-line rotation averaging on the B200 library, interface-compatible with the
// reference's `l1_irls` executable (ral/test.cpp, built by ral/CMakeLists.txt:86):
//
//   l1_irls input_file [output_file [cost [sigma_deg [irls_iters [l1_iters [change_th]]]]]]
//
// Same positional arguments and defaults (ral/test.cpp:250-272), same input format (ral/test.cpp:161-247:
// "m n f", m lines "i j w x y z", up to n lines "w x y z"; vertex ids compacted in sorted order), same call
// sequence through the adapter header (init_mst -> make_A -> l1ra -> irls -> quat_normalised,
// ral/test.cpp:285-302) and the same output file (ral/test.cpp:314-326): n rows "w x y z" then m weights,
// laid out the way Eigen's IOFormat(FullPrecision) prints a matrix - 15 significant digits, every entry
// right-aligned to the widest one, single-space separated.  IRA_CLI_PRECISION=17 in the environment
// overrides the digit count (round-trip output for parity tests).
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <string>

#include "l1_irls.hpp"
#include "ral_text_io.hpp"

namespace {

const double kPi = 3.141592653589793238462643383279502884;

void die(const std::string& msg) {
  std::cerr << msg << std::endl;
  std::exit(-1);
}

irotavg::Cost cost_from_name(const char* name) {          // names of ral/test.cpp:35-72, case-insensitive
  static const char* names[] = {"l2", "l1", "l1.5", "l0.5", "geman-mcclure", "huber", "pseudo-huber", "andrews",
                                "bisquare", "cauchy", "fair", "logistic", "talwar", "welsch"};
  std::string s(name ? name : "");
  std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  for (int c = 0; c < 14; ++c)
    if (s == names[c]) return (irotavg::Cost)c;
  die(std::string("Unknown string. ") + (name ? name : ""));
  return irotavg::Geman_McClure;
}

}  // namespace

int main(int argc, const char* argv[]) {
  using namespace irotavg;
  const int nargs = argc - 1;
  if (nargs < 1 || nargs > 7) {
    std::cerr << "usage: l1_irls input_file [output_file [cost [sigma_deg [irls_iters [l1_iters [change_th]]]]]]\n"
                 "  input_file : 'm n f', then m lines 'i j w x y z' (i<j), then >= f lines 'w x y z'\n"
                 "  output_file: n lines 'w x y z' followed by m IRLS weights (default l1_irls_out.txt)\n"
                 "  cost       : L2 L1 L1.5 L0.5 Geman-McClure Huber Pseudo-Huber Andrews Bisquare Cauchy Fair\n"
                 "               Logistic Talwar Welsch (default Geman-McClure); sigma_deg default 5;\n"
                 "  irls_iters default 50; l1_iters default 5; change_th default 0.001" << std::endl;
    return -1;
  }
  std::cout << "input file: " << argv[1] << std::endl;
  std::ifstream in(argv[1]);
  if (!in.is_open()) die(std::string("Unable to open file ") + argv[1]);

  int m = 0, n = 0, f = 0;
  in >> m >> n >> f;
  std::cout << "# rel rots ..... = " << m << "\n# abs rots ..... = " << n << "\n# fixed abs rots = " << f << std::endl;
  if (!in || m < 0 || n < 0 || f < 0) die("Corrupt input file: bad header.");

  I_t I;
  I.reserve((size_t)m);
  Mat QQ = Mat::Zero(m, 4), Q = Mat::Zero(n, 4);
  std::set<int> ids;
  for (int k = 0; k < m; ++k) {
    int a, b;
    double w, x, y, z;
    if (!(in >> a >> b >> w >> x >> y >> z)) die("Corrupt input file: inconsistent number of connections.");
    I.push_back(std::make_pair(a, b));
    ids.insert(a);
    ids.insert(b);
    QQ(k, 0) = x; QQ(k, 1) = y; QQ(k, 2) = z; QQ(k, 3) = w;        // file [w x y z] -> memory [x y z w]
  }
  std::map<int, int> compact;                                      // ids -> 0..#ids-1 in sorted order
  for (std::set<int>::const_iterator it = ids.begin(); it != ids.end(); ++it) {
    const int next = (int)compact.size();
    compact[*it] = next;
  }
  int max_second = -1;
  for (size_t k = 0; k < I.size(); ++k) {
    I[k].first = compact[I[k].first];
    I[k].second = compact[I[k].second];
    max_second = std::max(max_second, I[k].second);
  }
  int given = 0;
  while (given < n) {
    double w, x, y, z;
    if (!(in >> w >> x >> y >> z)) break;
    Q(given, 0) = x; Q(given, 1) = y; Q(given, 2) = z; Q(given, 3) = w;
    ++given;
  }
  in.close();
  if (given < f) {
    std::ostringstream o;
    o << "Insuficient number of absolute rotations. At least " << f << " must be given.";
    die(o.str());
  }
  if (n != max_second + 1) die("Corrupt input file: check abs rotations");

  const char* output_file = nargs > 1 ? argv[2] : "l1_irls_out.txt";
  const Cost cost = nargs > 2 ? cost_from_name(argv[3]) : Geman_McClure;
  const double sigma = (nargs > 3 ? std::atof(argv[4]) : 5.0) * kPi / 180.0;
  const int irls_iters = nargs > 4 ? std::atoi(argv[5]) : 50;
  const int l1_iters = nargs > 5 ? std::atoi(argv[6]) : 5;
  const double change_th = nargs > 6 ? std::atof(argv[7]) : 1e-3;
  std::cout << "output file: " << output_file << "\ncost: " << cost << "\nsigma [deg]: " << sigma * 180.0 / kPi
            << "\nIRLS max. iterations: " << irls_iters << "\nL1-RA max. iterations: " << l1_iters
            << "\nchange threshold: " << change_th << std::endl;

  if (f == 0) {                                                    // no fixed rotation: pin the first to I
    Q(0, 0) = 0; Q(0, 1) = 0; Q(0, 2) = 0; Q(0, 3) = 1;
    std::cout << "set first abs rot = I" << std::endl;
    f = 1;
  }
  std::cout << "# initial absolute rots " << given << std::endl;
  init_mst(Q, QQ, I, given > f ? given : f);
  const SpMat A = make_A(n, f, I);

  int l1_iters_out = 0, irls_iters_out = 0;
  double l1_runtime = 0.0, irls_runtime = 0.0;
  l1ra(QQ, I, A, Q, f, l1_iters, change_th, l1_iters_out, l1_runtime);
  Vec weights(m);
  irls(QQ, I, A, cost, sigma, Q, f, irls_iters, change_th, weights, irls_iters_out, irls_runtime);
  quat_normalised(Q, f);

  std::cout << "L1-RA iterations = " << l1_iters_out << "\nIRLS  iterations = " << irls_iters_out
            << "\nL1-RA runtime [s] = " << l1_runtime << "\nIRLS  runtime [s] = " << irls_runtime
            << "\ntotal runtime [s] = " << (l1_runtime + irls_runtime) << std::endl;

  std::ofstream out(output_file);
  if (!out.is_open()) {
    std::cerr << "Unable to save results." << std::endl;
    return 0;                                                      // the reference reports and still exits 0
  }
  int precision = 15;                                              // Eigen::FullPrecision for double
  if (const char* p = std::getenv("IRA_CLI_PRECISION")) precision = std::max(1, std::atoi(p));
  const int wxyz[4] = {3, 0, 1, 2};
  out << ira_b200::eigen_style(Q.data(), n, 4, n, wxyz, precision) << "\n";
  out << ira_b200::eigen_style(weights.data(), m, 1, m, nullptr, precision) << "\n";
  return 0;
}
