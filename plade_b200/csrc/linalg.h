// Small linear algebra + line/plane geometry shared by the host pipeline and the device kernels
// (no Eigen/OpenCV dependency).  Float expressions that feed decisions restate the reference's
// evaluation order; the library is built with -fmad=false / -ffp-contract=off so host and device
// evaluate them identically (IEEE add/mul/div/sqrt, no contraction).
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>
#include <limits>

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

// Eigen::SelfAdjointEigenSolver<Matrix3f>(A, ComputeEigenvectors) restated in float, statement for statement
// (Eigen 3.4: Eigenvalues/SelfAdjointEigenSolver.h compute() :414-466, computeFromTridiagonal_impl :502-573,
// tridiagonal_qr_step :837-898; Tridiagonalization.h 3x3 real special case :464-504; Jacobi.h makeGivens
// :231-267).  The bounding-box frame of ComputeBoundingBox (PLADE/util.h:199-201) -- hence the ORDER and the
// orientation of the rectangle corners the penetration filter walks -- depends on the signs this algorithm
// happens to produce, so a generic solver is not a substitute.  Only the lower triangle of A is read.
// Eigenvalues ascending in w, eigenvectors in the columns of V.
inline void sym_eig3f_eigen(const float A[3][3], float w[3], float V[3][3]) {
  float m[3][3] = {{A[0][0], 0.f, 0.f}, {A[1][0], A[1][1], 0.f}, {A[2][0], A[2][1], A[2][2]}};
  float scale = 0.f;
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) scale = std::max(scale, std::fabs(m[r][c]));
  if (scale == 0.f) scale = 1.f;
  for (int c = 0; c < 3; ++c) for (int r = c; r < 3; ++r) m[r][c] /= scale;
  float diag[3], sub[2];
  // tridiagonalization of a real 3x3
  const float tol = std::numeric_limits<float>::min();
  diag[0] = m[0][0];
  float v1norm2 = m[2][0] * m[2][0];
  float Q[3][3];
  if (v1norm2 <= tol) {
    diag[1] = m[1][1]; diag[2] = m[2][2];
    sub[0] = m[1][0]; sub[1] = m[2][1];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Q[r][c] = r == c ? 1.f : 0.f;
  } else {
    float beta = std::sqrt(m[1][0] * m[1][0] + v1norm2);
    float invBeta = 1.f / beta;
    float m01 = m[1][0] * invBeta, m02 = m[2][0] * invBeta;
    float q = 2.f * m01 * m[2][1] + m02 * (m[2][2] - m[1][1]);
    diag[1] = m[1][1] + m02 * q;
    diag[2] = m[2][2] - m02 * q;
    sub[0] = beta;
    sub[1] = m[2][1] - m01 * q;
    const float Qi[3][3] = {{1.f, 0.f, 0.f}, {0.f, m01, m02}, {0.f, m02, -m01}};
    memcpy(Q, Qi, sizeof(Q));
  }
  // implicit symmetric QR with Wilkinson shift
  const int n = 3, maxIterations = 30;
  int end = n - 1, start = 0, iter = 0;
  const float considerAsZero = std::numeric_limits<float>::min();
  const float precision_inv = 1.f / std::numeric_limits<float>::epsilon();
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (std::fabs(sub[i]) < considerAsZero) sub[i] = 0.f;
      else {
        const float scaled = precision_inv * sub[i];
        if (scaled * scaled <= (std::fabs(diag[i]) + std::fabs(diag[i + 1]))) sub[i] = 0.f;
      }
    }
    while (end > 0 && sub[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIterations * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0.f) start--;
    // one QR step on [start, end]
    float td = (diag[end - 1] - diag[end]) * 0.5f;
    float e = sub[end - 1];
    float mu = diag[end];
    if (td == 0.f) mu -= std::fabs(e);
    else if (e != 0.f) {
      const float e2 = e * e;
      float hx = std::fabs(td), hy = std::fabs(e);            // numext::hypot -> positive_real_hypot
      float hp = std::max(hx, hy), h;
      if (hp == 0.f) h = 0.f;
      else { float qp = std::min(hy, hx) / hp; h = hp * std::sqrt(1.f + qp * qp); }
      if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
      else mu -= e2 / (td + (td > 0.f ? h : -h));
    }
    float x = diag[start] - mu, z = sub[start];
    for (int k = start; k < end && z != 0.f; ++k) {
      float c, sn;                                             // JacobiRotation::makeGivens(x, z)
      if (z == 0.f) { c = x < 0.f ? -1.f : 1.f; sn = 0.f; }
      else if (x == 0.f) { c = 0.f; sn = z < 0.f ? 1.f : -1.f; }
      else if (std::fabs(x) > std::fabs(z)) {
        float t = z / x, u = std::sqrt(1.f + t * t);
        if (x < 0.f) u = -u;
        c = 1.f / u; sn = -t * c;
      } else {
        float t = x / z, u = std::sqrt(1.f + t * t);
        if (z < 0.f) u = -u;
        sn = -1.f / u; c = -t * sn;
      }
      float sdk = sn * diag[k] + c * sub[k];
      float dkp1 = sn * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - sn * sub[k]) - sn * (c * sub[k] - sn * diag[k + 1]);
      diag[k + 1] = sn * sdk + c * dkp1;
      sub[k] = c * sdk - sn * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - sn * z;
      x = sub[k];
      if (k < end - 1) { z = -sn * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
      // Q = Q * G : applyOnTheRight(k, k+1, rot) = rotation (c, -sn) on the two columns
      if (!(c == 1.f && -sn == 0.f))
        for (int r = 0; r < 3; ++r) {
          float xi = Q[r][k], yi = Q[r][k + 1];
          Q[r][k] = c * xi + (-sn) * yi;
          Q[r][k + 1] = -(-sn) * xi + c * yi;
        }
    }
  }
  if (iter <= maxIterations * n)
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      for (int j = 1; j < n - i; ++j) if (diag[i + j] < diag[i + k]) k = j;      // minCoeff: first minimum
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k + i]);
      }
    }
  for (int i = 0; i < 3; ++i) w[i] = diag[i] * scale;
  memcpy(V, Q, sizeof(Q));
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
