/*
 * Drop-in C++ header: the four registration() overloads of the reference
 * (/root/reference/code/PLADE/plade.h:44-96 — same names, argument order "target first", bool return,
 * identity left in `transformation` on failure), implemented as inline wrappers over the C ABI of
 * libplade_b200.so (include/plade_b200.h).
 *
 * Eigen is required (as in the reference).  The three PointCloud overloads are compiled when PCL is
 * available: define PLADE_WITH_PCL (the reference always builds with PCL; this repository does not
 * ship it).  Without PCL the array overload registration_arrays() gives the same functionality.
 *
 * Threading: like the reference these functions are not re-entrant per context; each host thread gets
 * its own CUDA context object (thread_local).
 *
 * The wrappers are inline, so a translation unit that includes this header needs nothing but libplade_b200.so.  Code that only
 * DECLARES the reference's prototypes (its own copy of PLADE/plade.h) and expects to link a `registration` symbol links
 * plade_b200/libplade_dropin.so instead (plade_b200/csrc/dropin.cpp: the same functions out of line).
 */
#ifndef PLADE_B200_DROPIN_H
#define PLADE_B200_DROPIN_H

#include <Eigen/Dense>
#include <string>
#include <vector>

#include "plade_b200.h"

#ifdef PLADE_WITH_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#endif

/* PLANE, /root/reference/code/PLADE/plane_extraction.h:44-50 */
#ifndef EASY3D_ALGO_POINT_CLOUD_RANSAC_H
class PLANE : public std::vector<int> {
public:
    template<class InputIt>
    PLANE(InputIt first, InputIt last) : std::vector<int>(first, last) {}
    Eigen::Vector3f normal;
    float d;
};
#endif

namespace plade_detail {
inline plade_ctx *ctx() {
    struct Holder {
        plade_ctx *c;
        Holder() : c(plade_ctx_create(-1)) {}
        ~Holder() { plade_ctx_destroy(c); }
    };
    static thread_local Holder h;
    return h.c;
}
inline void to_eigen(const float m[16], Eigen::Matrix<float, 4, 4> &T) {
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T(r, c) = m[4 * r + c];   /* row-major ABI -> Eigen */
}
inline void planes_to_csr(const std::vector<PLANE> &p, std::vector<int> &off, std::vector<int> &idx, std::vector<float> &par) {
    off.assign(1, 0);
    for (const PLANE &pl : p) {
        idx.insert(idx.end(), pl.begin(), pl.end());
        off.push_back((int) idx.size());
        par.push_back(pl.normal.x()); par.push_back(pl.normal.y()); par.push_back(pl.normal.z()); par.push_back(pl.d);
    }
}
}  // namespace plade_detail

/* registration(T, target_file, source_file) — PLADE/plade.cpp:665-707 */
inline bool registration(Eigen::Matrix<float, 4, 4> &transformation, const std::string &target_cloud_file,
                         const std::string &source_cloud_file) {
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   /* identity is what a failure leaves (PLADE/plade.cpp:696) */
    int ok = plade_register_files(plade_detail::ctx(), target_cloud_file.c_str(), source_cloud_file.c_str(), m);
    plade_detail::to_eigen(m, transformation);
    return ok != 0;
}

/* PCL-free equivalent of registration(T, target_cloud, source_cloud): interleaved x y z nx ny nz records */
inline bool registration_arrays(Eigen::Matrix<float, 4, 4> &transformation, const float *target_xyzn, size_t n_target,
                                const float *source_xyzn, size_t n_source) {
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   /* identity is what a failure leaves (PLADE/plade.cpp:696) */
    int ok = plade_register_clouds(plade_detail::ctx(), target_xyzn, n_target, source_xyzn, n_source, m);
    plade_detail::to_eigen(m, transformation);
    return ok != 0;
}

#ifdef PLADE_WITH_PCL
namespace plade_detail {
inline std::vector<float> flatten(const pcl::PointCloud<pcl::PointNormal> &c) {
    std::vector<float> v(c.size() * 6);
    for (size_t i = 0; i < c.size(); ++i) {
        const pcl::PointNormal &p = c.points[i];
        v[6 * i] = p.x; v[6 * i + 1] = p.y; v[6 * i + 2] = p.z;
        v[6 * i + 3] = p.normal_x; v[6 * i + 4] = p.normal_y; v[6 * i + 5] = p.normal_z;
    }
    return v;
}
}  // namespace plade_detail

/* PLADE/plade.cpp:638-662 */
inline bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                         pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud) {
    std::vector<float> t = plade_detail::flatten(*target_cloud), s = plade_detail::flatten(*source_cloud);
    return registration_arrays(transformation, t.data(), t.size() / 6, s.data(), s.size() / 6);
}

/* PLADE/plade.cpp:31-580 */
inline bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                         pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud, const std::vector<PLANE> &target_planes,
                         const std::vector<PLANE> &source_planes) {
    std::vector<float> t = plade_detail::flatten(*target_cloud), s = plade_detail::flatten(*source_cloud);
    std::vector<int> to, ti, so, si;
    std::vector<float> tp, sp;
    plade_detail::planes_to_csr(target_planes, to, ti, tp);
    plade_detail::planes_to_csr(source_planes, so, si, sp);
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   /* identity is what a failure leaves (PLADE/plade.cpp:696) */
    int ok = plade_register_with_planes(plade_detail::ctx(), t.data(), t.size() / 6, s.data(), s.size() / 6, to.data(), ti.data(),
                                        tp.data(), (int) target_planes.size(), so.data(), si.data(), sp.data(),
                                        (int) source_planes.size(), m);
    plade_detail::to_eigen(m, transformation);
    return ok != 0;
}

/* PLADE/plade.cpp:583-599 */
inline bool registration(Eigen::Matrix<float, 4, 4> &transformation, pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                         pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud, int ransac_min_support_target,
                         int ransac_min_support_source) {
    std::vector<float> t = plade_detail::flatten(*target_cloud), s = plade_detail::flatten(*source_cloud);
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   /* identity is what a failure leaves (PLADE/plade.cpp:696) */
    int ok = plade_register_min_support(plade_detail::ctx(), t.data(), t.size() / 6, s.data(), s.size() / 6,
                                        ransac_min_support_target, ransac_min_support_source, m);
    plade_detail::to_eigen(m, transformation);
    return ok != 0;
}
#endif  /* PLADE_WITH_PCL */

#endif  /* PLADE_B200_DROPIN_H */
