// libplade_dropin.so: the reference's registration() entry points (PLADE/plade.h:44-96) as OUT-OF-LINE symbols with the
// reference's exact signatures, for code that only declares those prototypes (its own copy of PLADE/plade.h) and links
// against a library -- include/plade.h offers the same functions as inline wrappers.  Thin forwarding to the C ABI of
// libplade_b200.so; needs Eigen's headers to build (as the reference does), PCL's for the PointCloud overloads
// (-DPLADE_WITH_PCL).  Built by `make dropin EIGEN_INC=...` (plade_b200/csrc/Makefile).
#include <Eigen/Dense>
#include <string>
#include <vector>
#include "plade_b200.h"
#ifdef PLADE_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include "plane_extraction.h"      // PLANE
#endif

namespace {
plade_ctx *ctx() {
  struct Holder {
    plade_ctx *c;
    Holder() : c(plade_ctx_create(-1)) {}
    ~Holder() { plade_ctx_destroy(c); }
  };
  static thread_local Holder h;
  return h.c;
}
void to_eigen(const float m[16], Eigen::Matrix<float, 4, 4> &T) {
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = m[4 * r + c];      // row-major ABI -> Eigen
}
const float kIdentity[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
}  // namespace

// registration(T, target_file, source_file) -- PLADE/plade.cpp:665-707
bool registration(Eigen::Matrix<float, 4, 4> &transformation, const std::string &target_cloud_file, const std::string &source_cloud_file) {
  float m[16];
  for (int i = 0; i < 16; ++i) m[i] = kIdentity[i];
  const int ok = plade_register_files(ctx(), target_cloud_file.c_str(), source_cloud_file.c_str(), m);
  to_eigen(m, transformation);
  return ok != 0;
}

#ifdef PLADE_WITH_PCL
namespace {
std::vector<float> flatten(const pcl::PointCloud<pcl::PointNormal> &c) {
  std::vector<float> v(c.size() * 6);
  for (size_t i = 0; i < c.size(); ++i) {
    const pcl::PointNormal &p = c.points[i];
    v[6 * i] = p.x; v[6 * i + 1] = p.y; v[6 * i + 2] = p.z; v[6 * i + 3] = p.normal_x; v[6 * i + 4] = p.normal_y; v[6 * i + 5] = p.normal_z;
  }
  return v;
}
void planes_to_csr(const std::vector<PLANE> &p, std::vector<int> &off, std::vector<int> &idx, std::vector<float> &par) {
  off.assign(1, 0);
  for (const PLANE &pl : p) {
    idx.insert(idx.end(), pl.begin(), pl.end());
    off.push_back((int) idx.size());
    par.push_back(pl.normal.x()); par.push_back(pl.normal.y()); par.push_back(pl.normal.z()); par.push_back(pl.d);
  }
}
}  // namespace

// PLADE/plade.cpp:638-662
bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                  pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud) {
  std::vector<float> t = flatten(*target_cloud), s = flatten(*source_cloud);
  float m[16];
  for (int i = 0; i < 16; ++i) m[i] = kIdentity[i];
  const int ok = plade_register_clouds(ctx(), t.data(), t.size() / 6, s.data(), s.size() / 6, m);
  to_eigen(m, transformation);
  return ok != 0;
}
// PLADE/plade.cpp:31-580
bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                  pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud, const std::vector<PLANE> &target_planes,
                  const std::vector<PLANE> &source_planes) {
  std::vector<float> t = flatten(*target_cloud), s = flatten(*source_cloud);
  std::vector<int> to, ti, so, si;
  std::vector<float> tp, sp;
  planes_to_csr(target_planes, to, ti, tp);
  planes_to_csr(source_planes, so, si, sp);
  float m[16];
  for (int i = 0; i < 16; ++i) m[i] = kIdentity[i];
  const int ok = plade_register_with_planes(ctx(), t.data(), t.size() / 6, s.data(), s.size() / 6, to.data(), ti.data(), tp.data(),
                                            (int) target_planes.size(), so.data(), si.data(), sp.data(), (int) source_planes.size(), m);
  to_eigen(m, transformation);
  return ok != 0;
}
// PLADE/plade.cpp:583-599
bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                  pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud, int ransac_min_support_target, int ransac_min_support_source) {
  std::vector<float> t = flatten(*target_cloud), s = flatten(*source_cloud);
  float m[16];
  for (int i = 0; i < 16; ++i) m[i] = kIdentity[i];
  const int ok = plade_register_min_support(ctx(), t.data(), t.size() / 6, s.data(), s.size() / 6, ransac_min_support_target,
                                            ransac_min_support_source, m);
  to_eigen(m, transformation);
  return ok != 0;
}
#endif
