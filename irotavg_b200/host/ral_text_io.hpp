// ral_text_io.hpp - the text layout of the reference CLI's output file (ral/test.cpp:314-326).
#ifndef IROTAVG_B200_RAL_TEXT_IO_HPP_
#define IROTAVG_B200_RAL_TEXT_IO_HPP_

#include <algorithm>
#include <sstream>
#include <string>
#include <vector>

namespace ira_b200 {

// Eigen's operator<< with IOFormat(precision): entries at `precision` significant digits, padded on the
// left to the width of the widest entry, columns separated by one space, rows by '\n'.
inline std::string eigen_style(const double* colmajor, long rows, long cols, long ld, const int* col_order, int precision) {
  std::vector<std::string> cell((size_t)rows * cols);
  size_t width = 0;
  for (long i = 0; i < rows; ++i)
    for (long j = 0; j < cols; ++j) {
      std::ostringstream o;
      o.precision(precision);
      o << colmajor[(size_t)(col_order ? col_order[j] : j) * ld + i];
      cell[(size_t)i * cols + j] = o.str();
      width = std::max(width, o.str().size());
    }
  std::string out;
  out.reserve((width + 1) * cell.size() + 1);
  for (long i = 0; i < rows; ++i) {
    if (i) out += '\n';
    for (long j = 0; j < cols; ++j) {
      if (j) out += ' ';
      const std::string& c = cell[(size_t)i * cols + j];
      out.append(width - c.size(), ' ');
      out += c;
    }
  }
  return out;
}


}  // namespace ira_b200
#endif
