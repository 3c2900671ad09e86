// K4a — rigid transform per descriptor match, K4b — clustering of transforms, sm_100a.
//
// K4a replaces ComputeTransformationUsingTwoVecAndOnePoint (PLADE/util.cpp:604-624), which feeds the
// three points {v1, v2, v1 x v2} -> {w1, w2, w1 x w2} to pcl TransformationEstimationSVD
// (registration/impl/transformation_estimation_svd.hpp:121-148) = Eigen::umeyama(src, dst, false):
//   sigma = (1/3) * sum (dst_i - mean_dst)(src_i - mean_src)^T ; sigma = U S V^T ;
//   R = U diag(1, 1, sign(det U det V)) V^T ;  then T = targetPoint - R * sourcePoint.
// One thread per match runs umeyama.h, a statement-for-statement float restatement of Eigen 3.4's umeyama +
// two-sided JacobiSVD<Matrix3f>, so R, T and the Euler angles equal the reference's bit for bit (the later
// stages threshold on them: clustering, the penetration filter's sample lattice, the verification ball).
//
// K4b replaces ClusterTransformation (PLADE/util.cpp:1245-1277) = pcl ConditionalEuclideanClustering
// (segmentation/impl/conditional_euclidean_clustering.hpp:43-148) over points (T, euler(R)): region
// growing over the symmetric relation
//   |Ta - Tb|^2 < float(tol*tol) (FLANN L2_Simple float)  &&  |euler_a - euler_b|^2 < angle_threshold
// whose result is the set of connected components; CEC seeds in index order, so a cluster's first
// element is its smallest index.  Here: spatial hash on T + lock-free union-find (atomicMin hooking,
// smaller index wins), label = smallest member index.
#include "kernels.h"
#include "umeyama.h"
#include "libm_flt32.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>

namespace plade {

namespace {

__global__ void transform_kernel(const MatchPairIn *__restrict__ in, int m, RigidOut *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  MatchPairIn mp = in[i];
  // PLADE/util.cpp:609-616: the third point of each triple is the cross product (Eigen float cross)
  V3 sv1(mp.sv1[0], mp.sv1[1], mp.sv1[2]), sv2(mp.sv2[0], mp.sv2[1], mp.sv2[2]);
  V3 dv1(mp.dv1[0], mp.dv1[1], mp.dv1[2]), dv2(mp.dv2[0], mp.dv2[1], mp.dv2[2]);
  V3 s3 = cross(sv1, sv2), d3 = cross(dv1, dv2);
  const float src[3][3] = {{sv1.x, sv1.y, sv1.z}, {sv2.x, sv2.y, sv2.z}, {s3.x, s3.y, s3.z}};
  const float dst[3][3] = {{dv1.x, dv1.y, dv1.z}, {dv2.x, dv2.y, dv2.z}, {d3.x, d3.y, d3.z}};
  float Rm[3][3];
  umeyama3_rotation_eigen(src, dst, Rm);
  RigidOut o;
  M3 R;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R(r, c) = Rm[r][c]; o.R[3 * r + c] = Rm[r][c]; }
  // T = targetPoint - R * sourcePoint (PLADE/util.cpp:621)
  V3 T = V3(mp.tp[0], mp.tp[1], mp.tp[2]) - mul(R, V3(mp.sp[0], mp.sp[1], mp.sp[2]));
  o.T[0] = T.x; o.T[1] = T.y; o.T[2] = T.z;
  // pcl::getEulerAngles<float> (common/impl/eigen.hpp:664-669) with the reference's libm, see libm_flt32.h
  o.euler[0] = atan2f_glibc(o.R[7], o.R[8]);
  o.euler[1] = asinf_glibc(-o.R[6]);
  o.euler[2] = atan2f_glibc(o.R[3], o.R[0]);
  o.pad = 0.f;
  out[i] = o;
}

// ---- clustering ------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cell_key3(int cx, int cy, int cz) {
  return ((unsigned long long) (unsigned int) (cx & 0x1fffff) << 42) | ((unsigned long long) (unsigned int) (cy & 0x1fffff) << 21) |
         (unsigned long long) (unsigned int) (cz & 0x1fffff);
}

__global__ void t_key_kernel(const RigidOut *__restrict__ rt, int m, float inv_cell, unsigned long long *__restrict__ keys,
                             int *__restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int cx = (int) floorf(rt[i].T[0] * inv_cell), cy = (int) floorf(rt[i].T[1] * inv_cell), cz = (int) floorf(rt[i].T[2] * inv_cell);
  keys[i] = cell_key3(cx + 0x100000, cy + 0x100000, cz + 0x100000);
  order[i] = i;
}

__device__ __forceinline__ int uf_find(int *parent, int x) {
  // parents only ever decrease, so the walk terminates; the halving store is a benign race (it
  // writes an ancestor of x, never a foreign node)
  for (;;) {
    int p = parent[x];
    if (p == x) return x;
    int gp = parent[p];
    if (gp != p) parent[x] = gp;
    x = p;
  }
}

__device__ void uf_union(int *parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    int hi = max(a, b), lo = min(a, b);
    int old = atomicMin(&parent[hi], lo);
    if (old == hi) return;
    a = old; b = lo;
  }
}

__device__ __forceinline__ int lower_bound_key(const unsigned long long *keys, int m, unsigned long long k) {
  int lo = 0, hi = m;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void cluster_edges_kernel(const RigidOut *__restrict__ rt, const unsigned long long *__restrict__ keys,
                                     const int *__restrict__ order, int m, float inv_cell, float tol2, float ang_thresh,
                                     int *__restrict__ parent) {
  int si = blockIdx.x * blockDim.x + threadIdx.x;
  if (si >= m) return;
  int a = order[si];
  RigidOut ra = rt[a];
  int cx = (int) floorf(ra.T[0] * inv_cell) + 0x100000, cy = (int) floorf(ra.T[1] * inv_cell) + 0x100000,
      cz = (int) floorf(ra.T[2] * inv_cell) + 0x100000;
  for (int dx = -1; dx <= 1; ++dx)
    for (int dy = -1; dy <= 1; ++dy) {
      // the three z-neighbours are contiguous in key order
      unsigned long long k0 = cell_key3(cx + dx, cy + dy, cz - 1), k1 = cell_key3(cx + dx, cy + dy, cz + 1);
      int j = lower_bound_key(keys, m, k0);
      for (; j < m && keys[j] <= k1; ++j) {
        int b = order[j];
        if (b <= a) continue;   // each unordered pair once
        const RigidOut &rb = rt[b];
        float ddx = __fsub_rn(ra.T[0], rb.T[0]), ddy = __fsub_rn(ra.T[1], rb.T[1]), ddz = __fsub_rn(ra.T[2], rb.T[2]);
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
        if (!(d2 < tol2)) continue;
        float e0 = __fsub_rn(ra.euler[0], rb.euler[0]), e1 = __fsub_rn(ra.euler[1], rb.euler[1]), e2 = __fsub_rn(ra.euler[2], rb.euler[2]);
        float en = __fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2));
        if (!(en < ang_thresh)) continue;
        uf_union(parent, a, b);
      }
    }
}

__global__ void iota_kernel(int *p, int m) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < m) p[i] = i; }
__global__ void flatten_kernel(int *parent, int m, int *label) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int x = i;
  while (parent[x] != x) x = parent[x];
  label[i] = x;
}

}  // namespace

void transforms_from_matches(Device &dev, HypScratch &sc, const MatchPairIn *h_in, size_t m, std::vector<RigidOut> &out) {
  out.resize(m);
  if (m == 0) return;
  cudaStream_t s = dev.stream;
  MatchPairIn *d_in = sc.in.ensure(m);
  RigidOut *d_out = sc.rt.ensure(m);
  PLADE_CUDA(cudaMemcpyAsync(d_in, h_in, sizeof(MatchPairIn) * m, cudaMemcpyHostToDevice, s));
  transform_kernel<<<div_up((long long) m, 128), 128, 0, s>>>(d_in, (int) m, d_out);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  PLADE_CUDA(cudaMemcpyAsync(out.data(), d_out, sizeof(RigidOut) * m, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
}

void cluster_transforms(Device &dev, HypScratch &sc, const std::vector<RigidOut> &rt, float dist_thresh, float ang_thresh,
                        std::vector<int> &label) {
  size_t m = rt.size();
  label.resize(m);
  if (m == 0) return;
  cudaStream_t s = dev.stream;
  RigidOut *d_rt = sc.rt.ensure(m);
  PLADE_CUDA(cudaMemcpyAsync(d_rt, rt.data(), sizeof(RigidOut) * m, cudaMemcpyHostToDevice, s));
  // radiusSearch(point, double(cluster_tolerance_)) -> float(radius * radius), strict '<'
  float tol2 = (float) ((double) dist_thresh * (double) dist_thresh);
  float cell = dist_thresh * 1.001f;
  if (!(cell > 0)) cell = 1e-6f;
  float inv_cell = 1.0f / cell;
  unsigned long long *keys = sc.key64_a.ensure(m), *keys2 = sc.key64_b.ensure(m);
  int *order = sc.cell_order.ensure(m), *order2 = sc.order_alt.ensure(m);
  int *parent = sc.label.ensure(m), *d_label = sc.misc.ensure(m);
  int blocks = div_up((long long) m, 256);
  t_key_kernel<<<blocks, 256, 0, s>>>(d_rt, (int) m, inv_cell, keys, order);
  PLADE_LAUNCH_CHECK();
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, order, order2, (int) m, 0, 63, s);
  unsigned char *tmp = sc.cub_tmp.ensure(tb);
  cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, order, order2, (int) m, 0, 63, s);
  iota_kernel<<<blocks, 256, 0, s>>>(parent, (int) m);
  cluster_edges_kernel<<<div_up((long long) m, 128), 128, 0, s>>>(d_rt, keys2, order2, (int) m, inv_cell, tol2, ang_thresh, parent);
  PLADE_LAUNCH_CHECK();
  flatten_kernel<<<blocks, 256, 0, s>>>(parent, (int) m, d_label);
  PLADE_LAUNCH_CHECK();
  dev.launches.add(12);
  PLADE_CUDA(cudaMemcpyAsync(label.data(), d_label, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
}

}  // namespace plade
