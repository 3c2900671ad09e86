// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// Boost-free restatement of the nine PCL 1.8.1 entry points that the PLADE hot path uses
// (SURVEY.md §2 row 10, §8c).  PCL itself cannot be compiled in this image (no Boost headers),
// so the reference's own P/*.cpp sources are compiled UNCHANGED against this header instead.
// Each function below restates, expression by expression, the PCL implementation it cites
// (paths relative to /root/reference/code/3rd_party/pcl-1.8.1/) and runs on the SAME Eigen 3.4.0
// and FLANN 1.8.4 headers the reference vendors, so float results are identical to a real PCL
// build with the same compiler flags.
#ifndef PLADE_PCL_SHIM_H
#define PLADE_PCL_SHIM_H

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include <algorithm>
#include <limits>
#include <functional>

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/StdVector>
#include <Eigen/SVD>
#include <Eigen/Eigenvalues>

#include <flann/flann.hpp>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_PI_2
#define M_PI_2 1.57079632679489661923
#endif

#define pcl_isfinite(x) std::isfinite(x)

namespace pcl {

// ---------------------------------------------------------------------------------------------
// Point types: common/include/pcl/impl/point_types.hpp (PointXYZ :286-300, PointNormal :820-850,
// PointXYZINormal :1000-1030).  Same 16-byte aligned float[4] unions, same constructors
// (data[3] = 1, data_n[3] = 0, curvature = 0).
// ---------------------------------------------------------------------------------------------
typedef Eigen::Map<Eigen::Array4f, Eigen::Aligned> Array4fMap;
typedef const Eigen::Map<const Eigen::Array4f, Eigen::Aligned> Array4fMapConst;
typedef Eigen::Map<Eigen::Vector3f> Vector3fMap;
typedef const Eigen::Map<const Eigen::Vector3f> Vector3fMapConst;
typedef Eigen::Map<Eigen::Vector4f, Eigen::Aligned> Vector4fMap;
typedef const Eigen::Map<const Eigen::Vector4f, Eigen::Aligned> Vector4fMapConst;

#define PLADE_SHIM_ADD_POINT4D                                                              \
  union EIGEN_ALIGN16 { float data[4]; struct { float x; float y; float z; }; };          \
  inline Vector3fMap getVector3fMap() { return Vector3fMap(data); }                         \
  inline Vector3fMapConst getVector3fMap() const { return Vector3fMapConst(data); }         \
  inline Vector4fMap getVector4fMap() { return Vector4fMap(data); }                         \
  inline Vector4fMapConst getVector4fMap() const { return Vector4fMapConst(data); }         \
  inline Array4fMap getArray4fMap() { return Array4fMap(data); }                            \
  inline Array4fMapConst getArray4fMap() const { return Array4fMapConst(data); }

#define PLADE_SHIM_ADD_NORMAL4D                                                             \
  union EIGEN_ALIGN16 { float data_n[4]; float normal[3];                                   \
    struct { float normal_x; float normal_y; float normal_z; }; };                         \
  inline Vector3fMap getNormalVector3fMap() { return Vector3fMap(data_n); }                 \
  inline Vector3fMapConst getNormalVector3fMap() const { return Vector3fMapConst(data_n); } \
  inline Vector4fMap getNormalVector4fMap() { return Vector4fMap(data_n); }                 \
  inline Vector4fMapConst getNormalVector4fMap() const { return Vector4fMapConst(data_n); }

struct EIGEN_ALIGN16 PointXYZ {
  PLADE_SHIM_ADD_POINT4D
  inline PointXYZ() { x = y = z = 0.0f; data[3] = 1.0f; }
  inline PointXYZ(float _x, float _y, float _z) { x = _x; y = _y; z = _z; data[3] = 1.0f; }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

struct EIGEN_ALIGN16 PointNormal {
  PLADE_SHIM_ADD_POINT4D
  PLADE_SHIM_ADD_NORMAL4D
  union { struct { float curvature; }; float data_c[4]; };
  inline PointNormal() {
    x = y = z = 0.0f; data[3] = 1.0f;
    normal_x = normal_y = normal_z = data_n[3] = 0.0f; curvature = 0.f;
  }
  inline PointNormal(float _x, float _y, float _z, float _nx, float _ny, float _nz) {
    x = _x; y = _y; z = _z; data[3] = 1.0f;
    normal_x = _nx; normal_y = _ny; normal_z = _nz; data_n[3] = 0.0f; curvature = 0.f;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

struct EIGEN_ALIGN16 PointXYZINormal {
  PLADE_SHIM_ADD_POINT4D
  PLADE_SHIM_ADD_NORMAL4D
  union { struct { float intensity; float curvature; }; float data_c[4]; };
  inline PointXYZINormal() {
    x = y = z = 0.0f; data[3] = 1.0f;
    normal_x = normal_y = normal_z = data_n[3] = 0.0f; intensity = 0.0f; curvature = 0;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

static_assert(sizeof(PointXYZ) == 16, "PointXYZ layout");
static_assert(sizeof(PointNormal) == 48, "PointNormal layout");
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal layout");

template <typename PointT> inline bool isFinite(const PointT &pt) {
  return pcl_isfinite(pt.x) && pcl_isfinite(pt.y) && pcl_isfinite(pt.z);
}

struct PCLHeader {
  uint32_t seq; uint64_t stamp; std::string frame_id;
  PCLHeader() : seq(0), stamp(0) {}
};

struct PointIndices {
  PCLHeader header;
  std::vector<int> indices;
  typedef std::shared_ptr<PointIndices> Ptr;
  typedef std::shared_ptr<const PointIndices> ConstPtr;
};
typedef std::vector<PointIndices> IndicesClusters;
typedef std::shared_ptr<std::vector<PointIndices> > IndicesClustersPtr;
typedef std::shared_ptr<std::vector<int> > IndicesPtr;
typedef std::shared_ptr<const std::vector<int> > IndicesConstPtr;

// ---------------------------------------------------------------------------------------------
// PointCloud: common/include/pcl/point_cloud.h:171-600 (the subset P/*.cpp touches).
// boost::shared_ptr -> std::shared_ptr; no arithmetic lives here.
// ---------------------------------------------------------------------------------------------
template <typename PointT>
class PointCloud {
public:
  typedef PointT PointType;
  typedef std::vector<PointT, Eigen::aligned_allocator<PointT> > VectorType;
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;
  typedef std::shared_ptr<const PointCloud<PointT> > ConstPtr;
  typedef typename VectorType::iterator iterator;
  typedef typename VectorType::const_iterator const_iterator;

  PointCloud() : width(0), height(0), is_dense(true),
                 sensor_origin_(Eigen::Vector4f::Zero()),
                 sensor_orientation_(Eigen::Quaternionf::Identity()) {}

  PCLHeader header;
  VectorType points;
  uint32_t width, height;
  bool is_dense;
  Eigen::Vector4f sensor_origin_;
  Eigen::Quaternionf sensor_orientation_;

  inline bool isOrganized() const { return height > 1; }
  inline iterator begin() { return points.begin(); }
  inline iterator end() { return points.end(); }
  inline const_iterator begin() const { return points.begin(); }
  inline const_iterator end() const { return points.end(); }
  inline size_t size() const { return points.size(); }
  inline void reserve(size_t n) { points.reserve(n); }
  inline bool empty() const { return points.empty(); }
  inline void resize(size_t n) {
    points.resize(n);
    if (width * height != n) { width = static_cast<uint32_t>(n); height = 1; }
  }
  inline const PointT &operator[](size_t n) const { return points[n]; }
  inline PointT &operator[](size_t n) { return points[n]; }
  inline const PointT &at(size_t n) const { return points.at(n); }
  inline PointT &at(size_t n) { return points.at(n); }
  inline const PointT &front() const { return points.front(); }
  inline const PointT &back() const { return points.back(); }
  inline void push_back(const PointT &pt) {
    points.push_back(pt);
    width = static_cast<uint32_t>(points.size()); height = 1;
  }
  inline void clear() { points.clear(); width = 0; height = 0; }
  inline Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

// ---------------------------------------------------------------------------------------------
// getMinMax3D: common/include/pcl/common/impl/common.hpp:228-262 (PointT out),
// :341-372 (indices, Vector4f out — the one VoxelGrid calls at voxel_grid.hpp:234).
// ---------------------------------------------------------------------------------------------
template <typename PointT> inline void
getMinMax3D(const PointCloud<PointT> &cloud, PointT &min_pt, PointT &max_pt) {
  Eigen::Array4f min_p, max_p;
  min_p.setConstant(FLT_MAX);
  max_p.setConstant(-FLT_MAX);
  if (cloud.is_dense) {
    for (size_t i = 0; i < cloud.points.size(); ++i) {
      Array4fMapConst pt = cloud.points[i].getArray4fMap();
      min_p = min_p.min(pt);
      max_p = max_p.max(pt);
    }
  } else {
    for (size_t i = 0; i < cloud.points.size(); ++i) {
      if (!pcl_isfinite(cloud.points[i].x) || !pcl_isfinite(cloud.points[i].y) ||
          !pcl_isfinite(cloud.points[i].z))
        continue;
      Array4fMapConst pt = cloud.points[i].getArray4fMap();
      min_p = min_p.min(pt);
      max_p = max_p.max(pt);
    }
  }
  min_pt.x = min_p[0]; min_pt.y = min_p[1]; min_pt.z = min_p[2];
  max_pt.x = max_p[0]; max_pt.y = max_p[1]; max_pt.z = max_p[2];
}

template <typename PointT> inline void
getMinMax3D(const PointCloud<PointT> &cloud, const std::vector<int> &indices,
            Eigen::Vector4f &min_pt, Eigen::Vector4f &max_pt) {
  min_pt.setConstant(FLT_MAX);
  max_pt.setConstant(-FLT_MAX);
  if (cloud.is_dense) {
    for (size_t i = 0; i < indices.size(); ++i) {
      Array4fMapConst pt = cloud.points[indices[i]].getArray4fMap();
      min_pt = min_pt.array().min(pt);
      max_pt = max_pt.array().max(pt);
    }
  } else {
    for (size_t i = 0; i < indices.size(); ++i) {
      if (!pcl_isfinite(cloud.points[indices[i]].x) || !pcl_isfinite(cloud.points[indices[i]].y) ||
          !pcl_isfinite(cloud.points[indices[i]].z))
        continue;
      Array4fMapConst pt = cloud.points[indices[i]].getArray4fMap();
      min_pt = min_pt.array().min(pt);
      max_pt = max_pt.array().max(pt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// transformPointCloud: common/include/pcl/common/impl/transforms.hpp:42-90 and the Matrix4
// forwarding overloads common/include/pcl/common/transforms.h:218-236.
// ---------------------------------------------------------------------------------------------
template <typename PointT, typename Scalar> void
transformPointCloud(const PointCloud<PointT> &cloud_in, PointCloud<PointT> &cloud_out,
                    const Eigen::Transform<Scalar, 3, Eigen::Affine> &transform,
                    bool copy_all_fields = true) {
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header;
    cloud_out.is_dense = cloud_in.is_dense;
    cloud_out.width = cloud_in.width;
    cloud_out.height = cloud_in.height;
    cloud_out.points.reserve(cloud_in.points.size());
    if (copy_all_fields)
      cloud_out.points.assign(cloud_in.points.begin(), cloud_in.points.end());
    else
      cloud_out.points.resize(cloud_in.points.size());
    cloud_out.sensor_orientation_ = cloud_in.sensor_orientation_;
    cloud_out.sensor_origin_ = cloud_in.sensor_origin_;
  }
  if (cloud_in.is_dense) {
    for (size_t i = 0; i < cloud_out.points.size(); ++i) {
      Eigen::Matrix<Scalar, 3, 1> pt(cloud_in[i].x, cloud_in[i].y, cloud_in[i].z);
      cloud_out[i].x = static_cast<float>(transform(0, 0) * pt.coeffRef(0) + transform(0, 1) * pt.coeffRef(1) + transform(0, 2) * pt.coeffRef(2) + transform(0, 3));
      cloud_out[i].y = static_cast<float>(transform(1, 0) * pt.coeffRef(0) + transform(1, 1) * pt.coeffRef(1) + transform(1, 2) * pt.coeffRef(2) + transform(1, 3));
      cloud_out[i].z = static_cast<float>(transform(2, 0) * pt.coeffRef(0) + transform(2, 1) * pt.coeffRef(1) + transform(2, 2) * pt.coeffRef(2) + transform(2, 3));
    }
  } else {
    for (size_t i = 0; i < cloud_out.points.size(); ++i) {
      if (!pcl_isfinite(cloud_in.points[i].x) || !pcl_isfinite(cloud_in.points[i].y) ||
          !pcl_isfinite(cloud_in.points[i].z))
        continue;
      Eigen::Matrix<Scalar, 3, 1> pt(cloud_in[i].x, cloud_in[i].y, cloud_in[i].z);
      cloud_out[i].x = static_cast<float>(transform(0, 0) * pt.coeffRef(0) + transform(0, 1) * pt.coeffRef(1) + transform(0, 2) * pt.coeffRef(2) + transform(0, 3));
      cloud_out[i].y = static_cast<float>(transform(1, 0) * pt.coeffRef(0) + transform(1, 1) * pt.coeffRef(1) + transform(1, 2) * pt.coeffRef(2) + transform(1, 3));
      cloud_out[i].z = static_cast<float>(transform(2, 0) * pt.coeffRef(0) + transform(2, 1) * pt.coeffRef(1) + transform(2, 2) * pt.coeffRef(2) + transform(2, 3));
    }
  }
}

template <typename PointT, typename Scalar> void
transformPointCloud(const PointCloud<PointT> &cloud_in, PointCloud<PointT> &cloud_out,
                    const Eigen::Matrix<Scalar, 4, 4> &transform, bool copy_all_fields = true) {
  Eigen::Transform<Scalar, 3, Eigen::Affine> t(transform);
  return (transformPointCloud<PointT, Scalar>(cloud_in, cloud_out, t, copy_all_fields));
}

// ---------------------------------------------------------------------------------------------
// compute3DCentroid / computeCovarianceMatrix(Normalized):
// common/include/pcl/common/impl/centroid.hpp:79-122, :180-259.
// ---------------------------------------------------------------------------------------------
template <typename PointT, typename Scalar> inline unsigned int
compute3DCentroid(const PointCloud<PointT> &cloud, Eigen::Matrix<Scalar, 4, 1> &centroid) {
  if (cloud.empty()) return (0);
  centroid.setZero();
  if (cloud.is_dense) {
    for (size_t i = 0; i < cloud.size(); ++i) {
      centroid[0] += cloud[i].x;
      centroid[1] += cloud[i].y;
      centroid[2] += cloud[i].z;
    }
    centroid /= static_cast<Scalar>(cloud.size());
    centroid[3] = 1;
    return (static_cast<unsigned int>(cloud.size()));
  } else {
    unsigned cp = 0;
    for (size_t i = 0; i < cloud.size(); ++i) {
      if (!isFinite(cloud[i])) continue;
      centroid[0] += cloud[i].x;
      centroid[1] += cloud[i].y;
      centroid[2] += cloud[i].z;
      ++cp;
    }
    centroid /= static_cast<Scalar>(cp);
    centroid[3] = 1;
    return (cp);
  }
}

template <typename PointT, typename Scalar> inline unsigned
computeCovarianceMatrix(const PointCloud<PointT> &cloud, const Eigen::Matrix<Scalar, 4, 1> &centroid,
                        Eigen::Matrix<Scalar, 3, 3> &covariance_matrix) {
  if (cloud.empty()) return (0);
  covariance_matrix.setZero();
  unsigned point_count = 0;
  for (size_t i = 0; i < cloud.size(); ++i) {
    if (!cloud.is_dense && !isFinite(cloud[i])) continue;
    Eigen::Matrix<Scalar, 4, 1> pt;
    pt[0] = cloud[i].x - centroid[0];
    pt[1] = cloud[i].y - centroid[1];
    pt[2] = cloud[i].z - centroid[2];
    covariance_matrix(1, 1) += pt.y() * pt.y();
    covariance_matrix(1, 2) += pt.y() * pt.z();
    covariance_matrix(2, 2) += pt.z() * pt.z();
    pt *= pt.x();
    covariance_matrix(0, 0) += pt.x();
    covariance_matrix(0, 1) += pt.y();
    covariance_matrix(0, 2) += pt.z();
    ++point_count;
  }
  covariance_matrix(1, 0) = covariance_matrix(0, 1);
  covariance_matrix(2, 0) = covariance_matrix(0, 2);
  covariance_matrix(2, 1) = covariance_matrix(1, 2);
  return (point_count);
}

template <typename PointT, typename Scalar> inline unsigned int
computeCovarianceMatrixNormalized(const PointCloud<PointT> &cloud, const Eigen::Matrix<Scalar, 4, 1> &centroid,
                                  Eigen::Matrix<Scalar, 3, 3> &covariance_matrix) {
  unsigned point_count = pcl::computeCovarianceMatrix(cloud, centroid, covariance_matrix);
  if (point_count != 0) covariance_matrix /= static_cast<Scalar>(point_count);
  return (point_count);
}

// common/include/pcl/common/impl/eigen.hpp:664-669
template <typename Scalar> void
getEulerAngles(const Eigen::Transform<Scalar, 3, Eigen::Affine> &t, Scalar &roll, Scalar &pitch, Scalar &yaw) {
  roll = atan2(t(2, 1), t(2, 2));
  pitch = asin(-t(2, 0));
  yaw = atan2(t(1, 0), t(0, 0));
}

// ---------------------------------------------------------------------------------------------
// KdTreeFLANN: kdtree/include/pcl/kdtree/impl/kdtree_flann.hpp:92-212 (+ convertCloudToArray
// :224-300).  Point representation = first three floats (common/include/pcl/point_representation.h
// :178-214, :318-375), validity = all finite (:101-118).  FLANN KDTreeSingleIndex, 15 pts/leaf,
// L2_Simple<float>, eps 0, sorted results.
// ---------------------------------------------------------------------------------------------
template <typename PointT>
class KdTreeFLANN {
public:
  typedef PointCloud<PointT> PointCloudT;
  typedef typename PointCloudT::ConstPtr PointCloudConstPtr;
  typedef ::flann::Index< ::flann::L2_Simple<float> > FLANNIndex;

  explicit KdTreeFLANN(bool sorted = true)
      : epsilon_(0.0f), sorted_(sorted), identity_mapping_(false), dim_(0), total_nr_points_(0),
        param_k_(::flann::SearchParams(-1, epsilon_)),
        param_radius_(::flann::SearchParams(-1, epsilon_, sorted)) {}

  void setInputCloud(const PointCloudConstPtr &cloud, const IndicesConstPtr &indices = IndicesConstPtr()) {
    index_mapping_.clear();
    flann_index_.reset();
    epsilon_ = 0.0f;
    dim_ = 3;
    input_ = cloud;
    indices_ = indices;
    if (!input_) return;
    if (indices)
      convertCloudToArray(*input_, *indices_);
    else
      convertCloudToArray(*input_);
    total_nr_points_ = static_cast<int>(index_mapping_.size());
    if (total_nr_points_ == 0) return;
    flann_index_.reset(new FLANNIndex(::flann::Matrix<float>(cloud_.get(), index_mapping_.size(), dim_),
                                      ::flann::KDTreeSingleIndexParams(15)));
    flann_index_->buildIndex();
  }

  int nearestKSearch(const PointT &point, int k, std::vector<int> &k_indices,
                     std::vector<float> &k_distances) const {
    if (k > total_nr_points_) k = total_nr_points_;
    k_indices.resize(k);
    k_distances.resize(k);
    std::vector<float> query(dim_);
    vectorize(point, &query[0]);
    ::flann::Matrix<int> k_indices_mat(&k_indices[0], 1, k);
    ::flann::Matrix<float> k_distances_mat(&k_distances[0], 1, k);
    flann_index_->knnSearch(::flann::Matrix<float>(&query[0], 1, dim_), k_indices_mat, k_distances_mat, k, param_k_);
    if (!identity_mapping_) {
      for (size_t i = 0; i < static_cast<size_t>(k); ++i) {
        int &neighbor_index = k_indices[i];
        neighbor_index = index_mapping_[neighbor_index];
      }
    }
    return (k);
  }

  int radiusSearch(const PointT &point, double radius, std::vector<int> &k_indices,
                   std::vector<float> &k_sqr_dists, unsigned int max_nn = 0) const {
    std::vector<float> query(dim_);
    vectorize(point, &query[0]);
    if (max_nn == 0 || max_nn > static_cast<unsigned int>(total_nr_points_)) max_nn = total_nr_points_;
    std::vector<std::vector<int> > indices(1);
    std::vector<std::vector<float> > dists(1);
    ::flann::SearchParams params(param_radius_);
    if (max_nn == static_cast<unsigned int>(total_nr_points_))
      params.max_neighbors = -1;
    else
      params.max_neighbors = max_nn;
    int neighbors_in_radius = flann_index_->radiusSearch(::flann::Matrix<float>(&query[0], 1, dim_), indices, dists,
                                                         static_cast<float>(radius * radius), params);
    k_indices = indices[0];
    k_sqr_dists = dists[0];
    if (!identity_mapping_) {
      for (int i = 0; i < neighbors_in_radius; ++i) {
        int &neighbor_index = k_indices[i];
        neighbor_index = index_mapping_[neighbor_index];
      }
    }
    return (neighbors_in_radius);
  }

  PointCloudConstPtr getInputCloud() const { return input_; }

private:
  static void vectorize(const PointT &p, float *out) {
    const float *ptr = reinterpret_cast<const float *>(&p);
    out[0] = ptr[0]; out[1] = ptr[1]; out[2] = ptr[2];
  }
  static bool isValid(const PointT &p) {
    const float *ptr = reinterpret_cast<const float *>(&p);
    return pcl_isfinite(ptr[0]) && pcl_isfinite(ptr[1]) && pcl_isfinite(ptr[2]);
  }
  void convertCloudToArray(const PointCloudT &cloud) {
    if (cloud.points.empty()) { cloud_.reset(); return; }
    int original_no_of_points = static_cast<int>(cloud.points.size());
    cloud_.reset(new float[original_no_of_points * dim_], std::default_delete<float[]>());
    float *cloud_ptr = cloud_.get();
    index_mapping_.reserve(original_no_of_points);
    identity_mapping_ = true;
    for (int cloud_index = 0; cloud_index < original_no_of_points; ++cloud_index) {
      if (!isValid(cloud.points[cloud_index])) { identity_mapping_ = false; continue; }
      index_mapping_.push_back(cloud_index);
      vectorize(cloud.points[cloud_index], cloud_ptr);
      cloud_ptr += dim_;
    }
  }
  void convertCloudToArray(const PointCloudT &cloud, const std::vector<int> &indices) {
    if (cloud.points.empty()) { cloud_.reset(); return; }
    int original_no_of_points = static_cast<int>(indices.size());
    cloud_.reset(new float[original_no_of_points * dim_], std::default_delete<float[]>());
    float *cloud_ptr = cloud_.get();
    index_mapping_.reserve(original_no_of_points);
    identity_mapping_ = false;
    for (std::vector<int>::const_iterator iIt = indices.begin(); iIt != indices.end(); ++iIt) {
      if (!isValid(cloud.points[*iIt])) continue;
      index_mapping_.push_back(*iIt);
      vectorize(cloud.points[*iIt], cloud_ptr);
      cloud_ptr += dim_;
    }
  }

  float epsilon_;
  bool sorted_;
  PointCloudConstPtr input_;
  IndicesConstPtr indices_;
  std::shared_ptr<FLANNIndex> flann_index_;
  std::shared_ptr<float> cloud_;
  std::vector<int> index_mapping_;
  bool identity_mapping_;
  int dim_;
  int total_nr_points_;
  ::flann::SearchParams param_k_;
  ::flann::SearchParams param_radius_;
};

namespace search {
// search/include/pcl/search/impl/kdtree.hpp:46-104 — thin forwarder, sorted = true by default
// (search/include/pcl/search/kdtree.h:93).
template <typename PointT>
class KdTree {
public:
  typedef PointCloud<PointT> PointCloudT;
  typedef typename PointCloudT::ConstPtr PointCloudConstPtr;
  typedef std::shared_ptr<KdTree<PointT> > Ptr;
  typedef std::shared_ptr<const KdTree<PointT> > ConstPtr;

  explicit KdTree(bool sorted = true) : tree_(new KdTreeFLANN<PointT>(sorted)) {}
  void setInputCloud(const PointCloudConstPtr &cloud, const IndicesConstPtr &indices = IndicesConstPtr()) {
    tree_->setInputCloud(cloud, indices);
    input_ = cloud;
    indices_ = indices;
  }
  PointCloudConstPtr getInputCloud() const { return input_; }
  int nearestKSearch(const PointT &point, int k, std::vector<int> &k_indices,
                     std::vector<float> &k_sqr_distances) const {
    return tree_->nearestKSearch(point, k, k_indices, k_sqr_distances);
  }
  int radiusSearch(const PointT &point, double radius, std::vector<int> &k_indices,
                   std::vector<float> &k_sqr_distances, unsigned int max_nn = 0) const {
    return tree_->radiusSearch(point, radius, k_indices, k_sqr_distances, max_nn);
  }
private:
  std::shared_ptr<KdTreeFLANN<PointT> > tree_;
  PointCloudConstPtr input_;
  IndicesConstPtr indices_;
};
}  // namespace search

// ---------------------------------------------------------------------------------------------
// VoxelGrid: filters/include/pcl/filters/impl/voxel_grid.hpp:214-437 (applyFilter, the
// no-filter-field branch, downsample_all_data_ = true, min_points_per_voxel_ = 0) with the
// CentroidPoint accumulators of common/include/pcl/common/impl/accumulators.hpp:65-135
// (xyz: Vector3f sum then `xyz / n`; normal: Vector4f sum then normalized(); curvature: sum / n).
// setLeafSize: filters/include/pcl/filters/voxel_grid.h:222-248.
// ---------------------------------------------------------------------------------------------
namespace detail {
template <typename PointT> struct CentroidAcc;
template <> struct CentroidAcc<PointXYZ> {
  Eigen::Vector3f xyz;
  CentroidAcc() : xyz(Eigen::Vector3f::Zero()) {}
  void add(const PointXYZ &t) { xyz += t.getVector3fMap(); }
  void get(PointXYZ &t, size_t n) const { t.getVector3fMap() = xyz / n; }
};
template <> struct CentroidAcc<PointNormal> {
  Eigen::Vector3f xyz; Eigen::Vector4f normal; float curvature;
  CentroidAcc() : xyz(Eigen::Vector3f::Zero()), normal(Eigen::Vector4f::Zero()), curvature(0) {}
  void add(const PointNormal &t) {
    xyz += t.getVector3fMap(); normal += t.getNormalVector4fMap(); curvature += t.curvature;
  }
  void get(PointNormal &t, size_t n) const {
    t.getVector3fMap() = xyz / n;
    t.getNormalVector4fMap() = normal.normalized();
    t.curvature = curvature / n;
  }
};
struct cloud_point_index_idx {
  unsigned int idx;
  unsigned int cloud_point_index;
  cloud_point_index_idx(unsigned int idx_, unsigned int cloud_point_index_)
      : idx(idx_), cloud_point_index(cloud_point_index_) {}
  bool operator<(const cloud_point_index_idx &p) const { return (idx < p.idx); }
};
}  // namespace detail

template <typename PointT>
class VoxelGrid {
public:
  typedef PointCloud<PointT> PointCloudT;
  typedef typename PointCloudT::ConstPtr PointCloudConstPtr;

  VoxelGrid() : leaf_size_(Eigen::Vector4f::Zero()), inverse_leaf_size_(Eigen::Array4f::Zero()),
                min_points_per_voxel_(0) {}

  void setLeafSize(float lx, float ly, float lz) {
    leaf_size_[0] = lx; leaf_size_[1] = ly; leaf_size_[2] = lz;
    if (leaf_size_[3] == 0) leaf_size_[3] = 1;
    inverse_leaf_size_ = Eigen::Array4f::Ones() / leaf_size_.array();
  }
  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }

  // filters/include/pcl/filters/filter.h:121-139 (no aliasing in the callers, so no temp copy)
  void filter(PointCloudT &output) {
    indices_.reset(new std::vector<int>(input_->points.size()));
    for (size_t i = 0; i < indices_->size(); ++i) (*indices_)[i] = static_cast<int>(i);
    output.header = input_->header;
    output.sensor_origin_ = input_->sensor_origin_;
    output.sensor_orientation_ = input_->sensor_orientation_;
    applyFilter(output);
  }

private:
  void applyFilter(PointCloudT &output) {
    output.height = 1;
    output.is_dense = true;
    Eigen::Vector4f min_p, max_p;
    getMinMax3D<PointT>(*input_, *indices_, min_p, max_p);

    int64_t dx = static_cast<int64_t>((max_p[0] - min_p[0]) * inverse_leaf_size_[0]) + 1;
    int64_t dy = static_cast<int64_t>((max_p[1] - min_p[1]) * inverse_leaf_size_[1]) + 1;
    int64_t dz = static_cast<int64_t>((max_p[2] - min_p[2]) * inverse_leaf_size_[2]) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
      output = *input_;
      return;
    }
    min_b_[0] = static_cast<int>(floor(min_p[0] * inverse_leaf_size_[0]));
    max_b_[0] = static_cast<int>(floor(max_p[0] * inverse_leaf_size_[0]));
    min_b_[1] = static_cast<int>(floor(min_p[1] * inverse_leaf_size_[1]));
    max_b_[1] = static_cast<int>(floor(max_p[1] * inverse_leaf_size_[1]));
    min_b_[2] = static_cast<int>(floor(min_p[2] * inverse_leaf_size_[2]));
    max_b_[2] = static_cast<int>(floor(max_p[2] * inverse_leaf_size_[2]));
    div_b_ = max_b_ - min_b_ + Eigen::Vector4i::Ones();
    div_b_[3] = 0;
    divb_mul_ = Eigen::Vector4i(1, div_b_[0], div_b_[0] * div_b_[1], 0);

    std::vector<detail::cloud_point_index_idx> index_vector;
    index_vector.reserve(indices_->size());
    for (std::vector<int>::const_iterator it = indices_->begin(); it != indices_->end(); ++it) {
      if (!input_->is_dense)
        if (!pcl_isfinite(input_->points[*it].x) || !pcl_isfinite(input_->points[*it].y) ||
            !pcl_isfinite(input_->points[*it].z))
          continue;
      int ijk0 = static_cast<int>(floor(input_->points[*it].x * inverse_leaf_size_[0]) - static_cast<float>(min_b_[0]));
      int ijk1 = static_cast<int>(floor(input_->points[*it].y * inverse_leaf_size_[1]) - static_cast<float>(min_b_[1]));
      int ijk2 = static_cast<int>(floor(input_->points[*it].z * inverse_leaf_size_[2]) - static_cast<float>(min_b_[2]));
      int idx = ijk0 * divb_mul_[0] + ijk1 * divb_mul_[1] + ijk2 * divb_mul_[2];
      index_vector.push_back(detail::cloud_point_index_idx(static_cast<unsigned int>(idx), *it));
    }
    std::sort(index_vector.begin(), index_vector.end(), std::less<detail::cloud_point_index_idx>());

    unsigned int total = 0;
    unsigned int index = 0;
    std::vector<std::pair<unsigned int, unsigned int> > first_and_last_indices_vector;
    first_and_last_indices_vector.reserve(index_vector.size());
    while (index < index_vector.size()) {
      unsigned int i = index + 1;
      while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
      if (i - index >= min_points_per_voxel_) {
        ++total;
        first_and_last_indices_vector.push_back(std::pair<unsigned int, unsigned int>(index, i));
      }
      index = i;
    }
    output.points.resize(total);
    index = 0;
    for (unsigned int cp = 0; cp < first_and_last_indices_vector.size(); ++cp) {
      unsigned int first_index = first_and_last_indices_vector[cp].first;
      unsigned int last_index = first_and_last_indices_vector[cp].second;
      detail::CentroidAcc<PointT> centroid;
      for (unsigned int li = first_index; li < last_index; ++li)
        centroid.add(input_->points[index_vector[li].cloud_point_index]);
      centroid.get(output.points[index], last_index - first_index);
      ++index;
    }
    output.width = static_cast<uint32_t>(output.points.size());
  }

  PointCloudConstPtr input_;
  IndicesPtr indices_;
  Eigen::Vector4f leaf_size_;
  Eigen::Array4f inverse_leaf_size_;
  unsigned int min_points_per_voxel_;
  Eigen::Vector4i min_b_, max_b_, div_b_, divb_mul_;
};

// ---------------------------------------------------------------------------------------------
// TransformationEstimationSVD (use_umeyama_ = true):
// registration/include/pcl/registration/impl/transformation_estimation_svd.hpp:46-62, :121-148
// -> common/include/pcl/common/impl/eigen.hpp:739-742 -> Eigen::umeyama(src, dst, false).
// ---------------------------------------------------------------------------------------------
namespace registration {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class TransformationEstimationSVD {
public:
  typedef Eigen::Matrix<Scalar, 4, 4> Matrix4;
  TransformationEstimationSVD(bool = true) {}
  inline void estimateRigidTransformation(const PointCloud<PointSource> &cloud_src,
                                          const PointCloud<PointTarget> &cloud_tgt,
                                          Matrix4 &transformation_matrix) const {
    size_t nr_points = cloud_src.points.size();
    if (cloud_tgt.points.size() != nr_points) return;
    const int npts = static_cast<int>(nr_points);
    Eigen::Matrix<Scalar, 3, Eigen::Dynamic> src(3, npts);
    Eigen::Matrix<Scalar, 3, Eigen::Dynamic> tgt(3, npts);
    for (int i = 0; i < npts; ++i) {
      src(0, i) = cloud_src[i].x; src(1, i) = cloud_src[i].y; src(2, i) = cloud_src[i].z;
      tgt(0, i) = cloud_tgt[i].x; tgt(1, i) = cloud_tgt[i].y; tgt(2, i) = cloud_tgt[i].z;
    }
    transformation_matrix = Eigen::umeyama(src, tgt, false);
  }
};
}  // namespace registration

// ---------------------------------------------------------------------------------------------
// ConditionalEuclideanClustering::segment:
// segmentation/include/pcl/segmentation/impl/conditional_euclidean_clustering.hpp:43-148
// (unorganised input -> pcl::search::KdTree; neighbour 0 skipped at :101).
// ---------------------------------------------------------------------------------------------
template <typename PointT>
class ConditionalEuclideanClustering {
public:
  typedef PointCloud<PointT> PointCloudT;
  typedef typename PointCloudT::ConstPtr PointCloudConstPtr;

  ConditionalEuclideanClustering(bool extract_removed_clusters = false)
      : condition_function_(), cluster_tolerance_(0.0f), min_cluster_size_(1),
        max_cluster_size_(std::numeric_limits<int>::max()),
        extract_removed_clusters_(extract_removed_clusters),
        small_clusters_(new IndicesClusters), large_clusters_(new IndicesClusters) {}

  void setInputCloud(const PointCloudConstPtr &cloud) { input_ = cloud; }
  void setConditionFunction(bool (*f)(const PointT &, const PointT &, float)) { condition_function_ = f; }
  void setClusterTolerance(float t) { cluster_tolerance_ = t; }
  void setMinClusterSize(int s) { min_cluster_size_ = s; }
  void setMaxClusterSize(int s) { max_cluster_size_ = s; }

  void segment(IndicesClusters &clusters) {
    clusters.clear();
    if (extract_removed_clusters_) { small_clusters_->clear(); large_clusters_->clear(); }
    if (!input_ || input_->points.empty() || !condition_function_) return;
    // PCLBase::initCompute (common/include/pcl/impl/pcl_base.hpp:138-175): fake indices 0..N-1
    IndicesPtr indices_(new std::vector<int>(input_->points.size()));
    for (size_t i = 0; i < indices_->size(); ++i) (*indices_)[i] = static_cast<int>(i);

    search::KdTree<PointT> searcher_;
    searcher_.setInputCloud(input_, indices_);
    std::vector<int> nn_indices;
    std::vector<float> nn_distances;
    std::vector<bool> processed(input_->points.size(), false);
    for (int iii = 0; iii < static_cast<int>(indices_->size()); ++iii) {
      if ((*indices_)[iii] == -1 || processed[(*indices_)[iii]]) continue;
      std::vector<int> current_cluster;
      int cii = 0;
      current_cluster.push_back((*indices_)[iii]);
      processed[(*indices_)[iii]] = true;
      while (cii < static_cast<int>(current_cluster.size())) {
        if (searcher_.radiusSearch(input_->points[current_cluster[cii]], cluster_tolerance_, nn_indices, nn_distances) < 1) {
          cii++;
          continue;
        }
        for (int nii = 1; nii < static_cast<int>(nn_indices.size()); ++nii) {
          if (nn_indices[nii] == -1 || processed[nn_indices[nii]]) continue;
          if (condition_function_(input_->points[current_cluster[cii]], input_->points[nn_indices[nii]], nn_distances[nii])) {
            current_cluster.push_back(nn_indices[nii]);
            processed[nn_indices[nii]] = true;
          }
        }
        cii++;
      }
      if (extract_removed_clusters_ ||
          (static_cast<int>(current_cluster.size()) >= min_cluster_size_ &&
           static_cast<int>(current_cluster.size()) <= max_cluster_size_)) {
        PointIndices pi;
        pi.header = input_->header;
        pi.indices.resize(current_cluster.size());
        for (int ii = 0; ii < static_cast<int>(current_cluster.size()); ++ii) pi.indices[ii] = current_cluster[ii];
        if (extract_removed_clusters_ && static_cast<int>(current_cluster.size()) < min_cluster_size_)
          small_clusters_->push_back(pi);
        else if (extract_removed_clusters_ && static_cast<int>(current_cluster.size()) > max_cluster_size_)
          large_clusters_->push_back(pi);
        else
          clusters.push_back(pi);
      }
    }
  }

private:
  PointCloudConstPtr input_;
  bool (*condition_function_)(const PointT &, const PointT &, float);
  float cluster_tolerance_;
  int min_cluster_size_, max_cluster_size_;
  bool extract_removed_clusters_;
  IndicesClustersPtr small_clusters_, large_clusters_;
};

}  // namespace pcl

#endif  // PLADE_PCL_SHIM_H
