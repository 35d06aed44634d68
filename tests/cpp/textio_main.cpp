// Prints ira_b200::eigen_style of a fixed 3 x 2 column-major matrix (default 15 digits, then 17) - CPU only.
#include <cstdio>

#include "ral_text_io.hpp"

int main() {
  const double a[6] = {1.0, -0.5, 1.0 / 3.0, 12345.678901234567, 2e-9, -7.0};   // columns (1, -.5, 1/3), (12345.67.., 2e-9, -7)
  const int swap[2] = {1, 0};
  std::printf("%s\n--\n%s\n--\n%s\n", ira_b200::eigen_style(a, 3, 2, 3, nullptr, 15).c_str(),
              ira_b200::eigen_style(a, 3, 2, 3, swap, 15).c_str(), ira_b200::eigen_style(a, 3, 1, 3, nullptr, 17).c_str());
  return 0;
}
