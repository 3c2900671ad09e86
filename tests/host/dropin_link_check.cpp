// Link check of plade_b200/libplade_dropin.so: this translation unit only DECLARES the reference's prototype
// (/root/reference/code/PLADE/plade.h:44-47), as a program built against the reference's own header would, and must link.
#include <Eigen/Dense>
#include <string>
#include <cstdio>
bool registration(Eigen::Matrix<float, 4, 4> &transformation, const std::string &target_cloud_file, const std::string &source_cloud_file);
int main(int argc, char **argv) {
  Eigen::Matrix<float, 4, 4> T;
  T.setZero();
  const bool ok = registration(T, argc > 1 ? argv[1] : "/nonexistent/target.ply", argc > 2 ? argv[2] : "/nonexistent/source.ply");
  std::printf("%d", ok ? 1 : 0);
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) std::printf(" %g", T(r, c));
  std::printf("\n");
  return 0;
}
