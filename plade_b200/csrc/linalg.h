// Small linear algebra + line/plane geometry shared by the host pipeline and the device kernels
// (no Eigen/OpenCV dependency).  Float expressions that feed decisions restate the reference's
// evaluation order; the library is built with -fmad=false / -ffp-contract=off so host and device
// evaluate them identically (IEEE add/mul/div/sqrt, no contraction).
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

#if defined(__CUDACC__)
#define PLADE_HD __host__ __device__ inline
#else
#define PLADE_HD inline
#endif

namespace plade {

struct V3 {
  float x, y, z;
  PLADE_HD V3() : x(0), y(0), z(0) {}
  PLADE_HD V3(float a, float b, float c) : x(a), y(b), z(c) {}
  PLADE_HD float &operator[](int i) { return (&x)[i]; }
  PLADE_HD float operator[](int i) const { return (&x)[i]; }
};
PLADE_HD V3 operator+(const V3 &a, const V3 &b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
PLADE_HD V3 operator-(const V3 &a, const V3 &b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
PLADE_HD V3 operator*(const V3 &a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
PLADE_HD V3 operator*(float s, const V3 &a) { return V3(s * a.x, s * a.y, s * a.z); }
PLADE_HD V3 operator/(const V3 &a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
PLADE_HD V3 operator-(const V3 &a) { return V3(-a.x, -a.y, -a.z); }
// Eigen 3.4 fixed-size dot / squaredNorm / sum of 3 terms: the unrolled reduction splits the range in
// halves (Core/Redux.h redux_novec_unroller), i.e. a0*b0 + (a1*b1 + a2*b2) -- checked against Eigen itself
PLADE_HD float dot(const V3 &a, const V3 &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
PLADE_HD float sqnorm(const V3 &a) { return dot(a, a); }
PLADE_HD float norm(const V3 &a) { return sqrtf(sqnorm(a)); }
PLADE_HD V3 cross(const V3 &a, const V3 &b) {
  return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen::MatrixBase::normalize(): z = squaredNorm(); if (z > 0) v /= sqrt(z)
PLADE_HD void normalize(V3 &a) {
  float z = sqnorm(a);
  if (z > 0) { float s = sqrtf(z); a.x /= s; a.y /= s; a.z /= s; }
}

struct M3 {           // row-major
  float m[9];
  PLADE_HD float operator()(int r, int c) const { return m[3 * r + c]; }
  PLADE_HD float &operator()(int r, int c) { return m[3 * r + c]; }
};
// Eigen lazy 3x3 * 3x1 coefficient product: row . vector with the same halving reduction as dot()
PLADE_HD V3 mul(const M3 &R, const V3 &v) {
  return V3(R.m[0] * v.x + (R.m[1] * v.y + R.m[2] * v.z), R.m[3] * v.x + (R.m[4] * v.y + R.m[5] * v.z),
            R.m[6] * v.x + (R.m[7] * v.y + R.m[8] * v.z));
}
// pcl::transformPointCloud formula (common/impl/transforms.hpp:69-71)
PLADE_HD V3 xform(const M3 &R, const V3 &T, const V3 &p) {
  return V3(((R.m[0] * p.x + R.m[1] * p.y) + R.m[2] * p.z) + T.x, ((R.m[3] * p.x + R.m[4] * p.y) + R.m[5] * p.z) + T.y,
            ((R.m[6] * p.x + R.m[7] * p.y) + R.m[8] * p.z) + T.z);
}
// FLANN L2_Simple (flann/algorithms/dist.h:84-90)
PLADE_HD float l2simple(const V3 &a, const V3 &b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  float r = dx * dx;
  r += dy * dy;
  r += dz * dz;
  return r;
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi in double), eigenvalues ascending, eigenvectors
// in the columns of V (unit length).  Stands in for Eigen::SelfAdjointEigenSolver<Matrix3f>
// (PLADE/util.h:199); eigenvector signs are arbitrary there as well.
inline void sym_eig3(const double Ain[3][3], double w[3], double V[3][3]) {
  double A[3][3];
  memcpy(A, Ain, sizeof(A));
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
    double diag = std::fabs(A[0][0]) + std::fabs(A[1][1]) + std::fabs(A[2][2]);
    if (off == 0.0 || off <= 1e-18 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = V[k][p], b = V[k][q]; V[k][p] = c * a - s * b; V[k][q] = s * a + c * b; }
      }
  }
  int ord[3] = {0, 1, 2};
  double ev[3] = {A[0][0], A[1][1], A[2][2]};
  std::sort(ord, ord + 3, [&](int a, int b) { return ev[a] < ev[b]; });
  double Vs[3][3];
  for (int k = 0; k < 3; ++k) { w[k] = ev[ord[k]]; for (int r = 0; r < 3; ++r) Vs[r][k] = V[r][ord[k]]; }
  memcpy(V, Vs, sizeof(Vs));
}

// Closest points of two 3-D lines, closed form in double; stands in for the 9x9 float
// cv::solve(DECOMP_SVD) of ComputeNearstTwoPointsOfTwo3DLine (PLADE/util.cpp:1167-1229).
// Returns false for (numerically) parallel lines.
PLADE_HD bool closest_points_two_lines(const V3 &d1, const V3 &p1, const V3 &d2, const V3 &p2, V3 &q1, V3 &q2) {
  double a = (double) d1.x * d2.x + (double) d1.y * d2.y + (double) d1.z * d2.z;
  double b = (double) d1.x * d1.x + (double) d1.y * d1.y + (double) d1.z * d1.z;
  double c = (double) d2.x * d2.x + (double) d2.y * d2.y + (double) d2.z * d2.z;
  double den = b * c - a * a;
  if (!(den > 1e-14 * b * c)) return false;
  double wx = (double) p2.x - p1.x, wy = (double) p2.y - p1.y, wz = (double) p2.z - p1.z;
  double w1 = wx * d1.x + wy * d1.y + wz * d1.z, w2 = wx * d2.x + wy * d2.y + wz * d2.z;
  double t1 = (c * w1 - a * w2) / den, t2 = (a * w1 - b * w2) / den;
  q1 = V3((float) (p1.x + t1 * d1.x), (float) (p1.y + t1 * d1.y), (float) (p1.z + t1 * d1.z));
  q2 = V3((float) (p2.x + t2 * d2.x), (float) (p2.y + t2 * d2.y), (float) (p2.z + t2 * d2.z));
  return true;
}

// ComputeIntersectionPointOf23DLine (PLADE/util.cpp:1461-1500): least-squares "intersection" of two
// lines (6x5 float SVD solve in the reference) = midpoint of their closest points, closed form.
PLADE_HD int line_line_point(const V3 &v1, const V3 &p1, const V3 &v2, const V3 &p2, V3 &out) {
  if (fabsf(dot(v1, v2)) > 0.9999) return -1;
  V3 q1, q2;
  if (!closest_points_two_lines(v1, p1, v2, p2, q1, q2)) return -1;
  out = V3((float) (0.5 * ((double) q1.x + q2.x)), (float) (0.5 * ((double) q1.y + q2.y)), (float) (0.5 * ((double) q1.z + q2.z)));
  return 0;
}

// ComputeIntersectionLineOfTwoPlanes (PLADE/util.cpp:626-676); planes are (n, d) with n.x + d = 0.
PLADE_HD int plane_intersection_line(const float pl1[4], const float pl2[4], V3 &lineVec, V3 &linePoint) {
  V3 p1(pl1[0], pl1[1], pl1[2]), p2(pl2[0], pl2[1], pl2[2]);
  normalize(p1);
  normalize(p2);
  if (fabsf(dot(p1, p2)) > 0.95) return -1;
  lineVec = cross(p1, p2);
  normalize(lineVec);
  const double b0 = -(double) pl1[3], b1 = -(double) pl2[3];
  // cv::Mat::inv() of a 2x2 CV_64F (closed form, opencv core lapack.cpp) followed by A^-1 * B
  double A00, A01, A10, A11;
  int which;
  if (fabsf(pl1[0] * pl2[1] - pl2[0] * pl1[1]) > 1e-6) { A00 = pl1[0]; A01 = pl1[1]; A10 = pl2[0]; A11 = pl2[1]; which = 0; }
  else if (fabsf(pl1[0] * pl2[2] - pl2[0] * pl1[2]) > 1e-6) { A00 = pl1[0]; A01 = pl1[2]; A10 = pl2[0]; A11 = pl2[2]; which = 1; }
  else if (fabsf(pl1[1] * pl2[2] - pl2[1] * pl1[2]) > 1e-6) { A00 = pl1[1]; A01 = pl1[2]; A10 = pl2[1]; A11 = pl2[2]; which = 2; }
  else return -1;
  double det = A00 * A11 - A01 * A10;
  double i00 = 0, i01 = 0, i10 = 0, i11 = 0;
  if (det != 0.) { double d = 1. / det; i11 = A00 * d; i00 = A11 * d; i01 = -A01 * d; i10 = -A10 * d; }
  double x0 = i00 * b0 + i01 * b1, x1 = i10 * b0 + i11 * b1;
  if (which == 0) linePoint = V3((float) x0, (float) x1, 0.f);
  else if (which == 1) linePoint = V3((float) x0, 0.f, (float) x1);
  else linePoint = V3(0.f, (float) x0, (float) x1);
  return 0;
}

}  // namespace plade
