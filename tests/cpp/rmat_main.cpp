// Host helpers of the rotAvg mirror (irotavg_b200/host/view_graph_rotavg.hpp) on stdin matrices: for each row-major
// 3x3 rotation prints rmat2quat (src/ViewGraph.cpp:1175-1203) and the matrix rebuilt by quat2rmat_rowmajor
// (Eigen's toRotationMatrix, src/ViewGraph.cpp:1426-1431).  No device call.
#include <cstdio>

#include "view_shim.hpp"
#include "view_graph_rotavg.hpp"

int main() {
  double R[9];
  std::printf("%s", "");
  while (std::scanf("%lf %lf %lf %lf %lf %lf %lf %lf %lf", R, R + 1, R + 2, R + 3, R + 4, R + 5, R + 6, R + 7, R + 8) == 9) {
    cv::Matx33d M(R);
    double q[4], B[9];
    ira_b200::rmat2quat(M, q);
    ira_b200::quat2rmat_rowmajor(q, B);
    std::printf("%.17g %.17g %.17g %.17g", q[0], q[1], q[2], q[3]);
    for (int k = 0; k < 9; ++k) std::printf(" %.17g", B[k]);
    std::printf("\n");
  }
  return 0;
}
