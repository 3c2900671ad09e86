// Compile-and-run check of the drop-in C++ header include/plade.h: the four registration() overloads of the
// reference (PLADE/plade.h:44-96) with their exact signatures, called the way PLADE/main.cpp calls them.
// Built by tests/test_oracle_cpu.py against the reference's Eigen and the Boost-free PCL shim of the oracle.
#include "plade.h"
#include <cstdio>
#include <type_traits>

typedef pcl::PointCloud<pcl::PointNormal>::Ptr CloudPtr;
// the overload set must resolve exactly like the reference's declarations
static_assert(std::is_same<decltype(registration(std::declval<Eigen::Matrix<float, 4, 4> &>(), std::declval<const std::string &>(), std::declval<const std::string &>())), bool>::value, "file overload");
static_assert(std::is_same<decltype(registration(std::declval<Eigen::Matrix<float, 4, 4> &>(), std::declval<CloudPtr>(), std::declval<CloudPtr>())), bool>::value, "cloud overload");
static_assert(std::is_same<decltype(registration(std::declval<Eigen::Matrix<float, 4, 4> &>(), std::declval<CloudPtr>(), std::declval<CloudPtr>(), std::declval<const std::vector<PLANE> &>(), std::declval<const std::vector<PLANE> &>())), bool>::value, "planes overload");
static_assert(std::is_same<decltype(registration(std::declval<Eigen::Matrix<float, 4, 4> &>(), std::declval<CloudPtr>(), std::declval<CloudPtr>(), 1, 1)), bool>::value, "min-support overload");

int main(int argc, char **argv) {
  Eigen::Matrix<float, 4, 4> T;
  T.setZero();
  CloudPtr a(new pcl::PointCloud<pcl::PointNormal>), b(new pcl::PointCloud<pcl::PointNormal>);
  for (int i = 0; i < 100; ++i) {
    pcl::PointNormal p;
    p.x = 0.01f * i; p.y = 0.02f * i; p.z = 0.f; p.normal_x = 0.f; p.normal_y = 0.f; p.normal_z = 1.f;
    a->push_back(p); b->push_back(p);
  }
  std::vector<int> idx = {0, 1, 2};
  PLANE pl(idx.begin(), idx.end());
  pl.normal = Eigen::Vector3f(0, 0, 1); pl.d = 0;
  std::vector<PLANE> planes(1, pl);
  int results = 0;
  bool ok = registration(T, std::string(argc > 1 ? argv[1] : "missing_target.ply"), std::string(argc > 2 ? argv[2] : "missing_source.ply"));
  results |= ok ? 1 : 0;
  const bool id1 = T.isApprox(Eigen::Matrix<float, 4, 4>::Identity());
  ok = registration(T, a, b);                 results |= ok ? 2 : 0;
  ok = registration(T, a, b, planes, planes); results |= ok ? 4 : 0;
  ok = registration(T, a, b, 10, 10);         results |= ok ? 8 : 0;
  const bool id2 = T.isApprox(Eigen::Matrix<float, 4, 4>::Identity());
  std::printf("results=%d identity_after_failure=%d\n", results, (id1 && id2) ? 1 : 0);
  return 0;
}
