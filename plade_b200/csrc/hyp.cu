// K4a — rigid transform per descriptor match, K4b — clustering of transforms, sm_100a.
//
// K4a replaces ComputeTransformationUsingTwoVecAndOnePoint (PLADE/util.cpp:604-624), which feeds the
// three points {v1, v2, v1 x v2} -> {w1, w2, w1 x w2} to pcl TransformationEstimationSVD
// (registration/impl/transformation_estimation_svd.hpp:121-148) = Eigen::umeyama(src, dst, false):
//   sigma = (1/3) * sum (dst_i - mean_dst)(src_i - mean_src)^T ; sigma = U S V^T ;
//   R = U diag(1, 1, sign(det U det V)) V^T ;  then T = targetPoint - R * sourcePoint.
// Here R is evaluated in fp64 from the two dominant singular triplets,
//   R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T,
// which equals the umeyama rotation for every sign choice of the third singular pair (the demeaned
// triangle has rank 2), then rounded to float; T uses the reference's float expression order.
// Float-parity with Eigen's float JacobiSVD is tolerance-level (|dR| ~ 1e-7), see DESIGN.md.
//
// K4b replaces ClusterTransformation (PLADE/util.cpp:1245-1277) = pcl ConditionalEuclideanClustering
// (segmentation/impl/conditional_euclidean_clustering.hpp:43-148) over points (T, euler(R)): region
// growing over the symmetric relation
//   |Ta - Tb|^2 < float(tol*tol) (FLANN L2_Simple float)  &&  |euler_a - euler_b|^2 < angle_threshold
// whose result is the set of connected components; CEC seeds in index order, so a cluster's first
// element is its smallest index.  Here: spatial hash on T + lock-free union-find (atomicMin hooking,
// smaller index wins), label = smallest member index.
#include "kernels.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>

namespace plade {

namespace {

__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (double); V columns = eigenvectors
__device__ void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
    double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (off <= 1e-300 || off <= 1e-17 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p][q];
        if (fabs(apq) < 1e-300) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

__global__ void transform_kernel(const MatchPairIn *__restrict__ in, int m, RigidOut *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  MatchPairIn mp = in[i];
  // PLADE/util.cpp:609-616: third point = cross product, float arithmetic (Eigen cross, no FMA)
  float s3[3], d3[3];
  s3[0] = __fsub_rn(__fmul_rn(mp.sv1[1], mp.sv2[2]), __fmul_rn(mp.sv1[2], mp.sv2[1]));
  s3[1] = __fsub_rn(__fmul_rn(mp.sv1[2], mp.sv2[0]), __fmul_rn(mp.sv1[0], mp.sv2[2]));
  s3[2] = __fsub_rn(__fmul_rn(mp.sv1[0], mp.sv2[1]), __fmul_rn(mp.sv1[1], mp.sv2[0]));
  d3[0] = __fsub_rn(__fmul_rn(mp.dv1[1], mp.dv2[2]), __fmul_rn(mp.dv1[2], mp.dv2[1]));
  d3[1] = __fsub_rn(__fmul_rn(mp.dv1[2], mp.dv2[0]), __fmul_rn(mp.dv1[0], mp.dv2[2]));
  d3[2] = __fsub_rn(__fmul_rn(mp.dv1[0], mp.dv2[1]), __fmul_rn(mp.dv1[1], mp.dv2[0]));
  double S[3][3], D[3][3];   // rows = points
  for (int k = 0; k < 3; ++k) { S[0][k] = mp.sv1[k]; S[1][k] = mp.sv2[k]; S[2][k] = s3[k]; D[0][k] = mp.dv1[k]; D[1][k] = mp.dv2[k]; D[2][k] = d3[k]; }
  double ms[3], md[3];
  for (int k = 0; k < 3; ++k) { ms[k] = (S[0][k] + S[1][k] + S[2][k]) / 3.0; md[k] = (D[0][k] + D[1][k] + D[2][k]) / 3.0; }
  double sig[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
    double a = 0;
    for (int p = 0; p < 3; ++p) a += (D[p][r] - md[r]) * (S[p][c] - ms[c]);
    sig[r][c] = a / 3.0;
  }
  double A[3][3], V[3][3], w[3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r][c] = sig[0][r] * sig[0][c] + sig[1][r] * sig[1][c] + sig[2][r] * sig[2][c];
  jacobi_eig3(A, V, w);
  int i0 = 0;
  if (w[1] > w[i0]) i0 = 1;
  if (w[2] > w[i0]) i0 = 2;
  int i1 = (i0 + 1) % 3, i2 = (i0 + 2) % 3;
  if (w[i2] > w[i1]) { int t = i1; i1 = i2; i2 = t; }
  double v1[3] = {V[0][i0], V[1][i0], V[2][i0]}, v2[3] = {V[0][i1], V[1][i1], V[2][i1]};
  double u1[3], u2[3];
  for (int r = 0; r < 3; ++r) { u1[r] = sig[r][0] * v1[0] + sig[r][1] * v1[1] + sig[r][2] * v1[2]; u2[r] = sig[r][0] * v2[0] + sig[r][1] * v2[1] + sig[r][2] * v2[2]; }
  double n1 = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
  for (int r = 0; r < 3; ++r) u1[r] /= (n1 > 0 ? n1 : 1.0);
  double dp = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
  for (int r = 0; r < 3; ++r) u2[r] -= dp * u1[r];
  double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
  for (int r = 0; r < 3; ++r) u2[r] /= (n2 > 0 ? n2 : 1.0);
  double u3[3], v3[3];
  cross3(u1, u2, u3);
  cross3(v1, v2, v3);
  RigidOut o;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
    o.R[3 * r + c] = (float) (u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c]);
  // T = targetPoint - R * sourcePoint (Eigen float, row dot products left to right)
  for (int r = 0; r < 3; ++r) {
    float rs = __fadd_rn(__fadd_rn(__fmul_rn(o.R[3 * r], mp.sp[0]), __fmul_rn(o.R[3 * r + 1], mp.sp[1])), __fmul_rn(o.R[3 * r + 2], mp.sp[2]));
    o.T[r] = __fsub_rn(mp.tp[r], rs);
  }
  // pcl::getEulerAngles (common/impl/eigen.hpp:664-669), float results
  o.euler[0] = (float) atan2((double) o.R[7], (double) o.R[8]);
  o.euler[1] = (float) asin(-(double) o.R[6]);
  o.euler[2] = (float) atan2((double) o.R[3], (double) o.R[0]);
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
  PLADE_CUDA(cudaStreamSynchronize(s));
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
  static thread_local DevBuf<unsigned long long> k_a, k_b;
  unsigned long long *keys = k_a.ensure(m), *keys2 = k_b.ensure(m);
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
  PLADE_CUDA(cudaStreamSynchronize(s));
}

}  // namespace plade
