// Eigen::umeyama(src, dst, /*with_scaling=*/false) for three 3-D points, restated in float for host and device.
//
// ComputeTransformationUsingTwoVecAndOnePoint (PLADE/util.cpp:604-624) sends {v1, v2, v1 x v2} -> {w1, w2, w1 x w2}
// through pcl TransformationEstimationSVD (registration/impl/transformation_estimation_svd.hpp:121-148), i.e.
// Eigen 3.4's umeyama (Geometry/Umeyama.h:93-166) with its two-sided JacobiSVD<Matrix3f> (SVD/JacobiSVD.h
// compute(), misc/RealSvd2x2.h, Jacobi/Jacobi.h), and keeps only the rotation block; T = targetPoint - R *
// sourcePoint.  Downstream stages threshold on these transforms (clustering, the penetration filter's sample
// lattice), so last-bit differences change which hypotheses survive: the restatement follows Eigen's
// statements and evaluation order exactly (no FMA; only IEEE + - * / sqrt).
#pragma once
#include <cfloat>
#include <cmath>
#include "linalg.h"

namespace plade {

struct JRot { float c, s; };     // Eigen::JacobiRotation<float>

// x' = c x + s y ; y' = -s x + c y   (internal::apply_rotation_in_the_plane, Jacobi.h:323-340, incl. its early out)
PLADE_HD void jrot_apply(float &x, float &y, float c, float s) {
  float xi = x, yi = y;
  x = c * xi + s * yi;
  y = -s * xi + c * yi;
}

// JacobiRotation::makeJacobi(x, y, z)  (Jacobi.h:91-122)
PLADE_HD JRot jrot_make_jacobi(float x, float y, float z) {
  JRot r;
  float deno = 2.f * fabsf(y);
  if (deno < FLT_MIN) { r.c = 1.f; r.s = 0.f; return r; }
  float tau = (x - z) / deno;
  float w = sqrtf(tau * tau + 1.f);
  float t = tau > 0.f ? 1.f / (tau + w) : 1.f / (tau - w);
  float sign_t = t > 0.f ? 1.f : -1.f;
  float n = 1.f / sqrtf(t * t + 1.f);
  r.s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
  r.c = n;
  return r;
}

// JacobiSVD<Matrix3f>(A, ComputeFullU | ComputeFullV): A = U diag(sv) V^T, sv descending
PLADE_HD void svd3f_eigen(const float A[3][3], float U[3][3], float V[3][3], float sv[3]) {
  const float precision = 2.f * FLT_EPSILON, considerAsZero = FLT_MIN;
  float scale = 0.f;
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) scale = fmaxf(scale, fabsf(A[r][c]));
  if (scale == 0.f) scale = 1.f;
  float W[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { W[r][c] = A[r][c] / scale; U[r][c] = V[r][c] = r == c ? 1.f : 0.f; }
  float maxDiag = fmaxf(fmaxf(fabsf(W[0][0]), fabsf(W[1][1])), fabsf(W[2][2]));
  bool finished = false;
  while (!finished) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        float threshold = fmaxf(considerAsZero, precision * maxDiag);
        if (fabsf(W[p][q]) > threshold || fabsf(W[q][p]) > threshold) {
          finished = false;
          // real_2x2_jacobi_svd (misc/RealSvd2x2.h:19-50)
          float m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          JRot rot1;
          float t = m00 + m11, d = m10 - m01;
          if (fabsf(d) < FLT_MIN) { rot1.s = 0.f; rot1.c = 1.f; }
          else {
            float u = t / d, tmp = sqrtf(1.f + u * u);
            rot1.s = 1.f / tmp;
            rot1.c = u / tmp;
          }
          if (!(rot1.c == 1.f && rot1.s == 0.f)) { jrot_apply(m00, m10, rot1.c, rot1.s); jrot_apply(m01, m11, rot1.c, rot1.s); }
          JRot jr = jrot_make_jacobi(m00, m01, m11);
          // j_left = rot1 * j_right.transpose()   (JacobiRotation::operator*, Jacobi.h:50-56)
          float oc = jr.c, os = -jr.s;
          JRot jl;
          jl.c = rot1.c * oc - rot1.s * os;
          jl.s = rot1.c * os + rot1.s * oc;
          // W.applyOnTheLeft(p, q, j_left): rows p, q
          if (!(jl.c == 1.f && jl.s == 0.f)) {
            for (int k = 0; k < 3; ++k) jrot_apply(W[p][k], W[q][k], jl.c, jl.s);
            // U.applyOnTheRight(p, q, j_left.transpose()): columns p, q, rotation j_left
            for (int k = 0; k < 3; ++k) jrot_apply(U[k][p], U[k][q], jl.c, jl.s);
          }
          // W.applyOnTheRight(p, q, j_right), V.applyOnTheRight(p, q, j_right): columns, rotation j_right^T
          float rc = jr.c, rs = -jr.s;
          if (!(rc == 1.f && rs == 0.f)) {
            for (int k = 0; k < 3; ++k) jrot_apply(W[k][p], W[k][q], rc, rs);
            for (int k = 0; k < 3; ++k) jrot_apply(V[k][p], V[k][q], rc, rs);
          }
          maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[p][p]), fabsf(W[q][q])));
        }
      }
  }
  for (int i = 0; i < 3; ++i) {
    float a = W[i][i];
    sv[i] = fabsf(a);
    if (a < 0.f) for (int k = 0; k < 3; ++k) U[k][i] = -U[k][i];
  }
  for (int i = 0; i < 3; ++i) sv[i] *= scale;
  for (int i = 0; i < 3; ++i) {
    int pos = 0;
    float best = sv[i];
    for (int k = 1; k < 3 - i; ++k) if (sv[i + k] > best) { best = sv[i + k]; pos = k; }     // maxCoeff: first maximum
    if (best == 0.f) break;
    if (pos) {
      pos += i;
      float tsv = sv[i]; sv[i] = sv[pos]; sv[pos] = tsv;
      for (int k = 0; k < 3; ++k) { float tu = U[k][pos]; U[k][pos] = U[k][i]; U[k][i] = tu; float tv = V[k][pos]; V[k][pos] = V[k][i]; V[k][i] = tv; }
    }
  }
}

// Eigen's 3x3 determinant (LU/Determinant.h bruteforce_det3_helper)
PLADE_HD float det3_eigen(const float M[3][3]) {
  float a = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]);
  float b = M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]);
  float c = M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
  return a - b + c;
}

// rotation block of umeyama(src, dst, false); src / dst as [point][coordinate], three points each
PLADE_HD void umeyama3_rotation_eigen(const float src[3][3], const float dst[3][3], float R[3][3]) {
  const float one_over_n = 1.f / 3.f;
  float sm[3], dm[3];
  for (int r = 0; r < 3; ++r) {
    sm[r] = ((src[0][r] + src[1][r]) + src[2][r]) * one_over_n;
    dm[r] = ((dst[0][r] + dst[1][r]) + dst[2][r]) * one_over_n;
  }
  float sd[3][3], dd[3][3];                      // demeaned, [coordinate][point]
  for (int r = 0; r < 3; ++r) for (int p = 0; p < 3; ++p) { sd[r][p] = src[p][r] - sm[r]; dd[r][p] = dst[p][r] - dm[r]; }
  float sigma[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    sigma[i][j] = one_over_n * ((dd[i][0] * sd[j][0] + dd[i][1] * sd[j][1]) + dd[i][2] * sd[j][2]);
  float U[3][3], V[3][3], sv[3];
  svd3f_eigen(sigma, U, V, sv);
  float S2 = (det3_eigen(U) * det3_eigen(V) < 0.f) ? -1.f : 1.f;
  // R = (U * S.asDiagonal()) * V^T, coefficient-wise product with Eigen's halving reduction
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    R[i][j] = (U[i][0] * 1.f) * V[j][0] + ((U[i][1] * 1.f) * V[j][1] + (U[i][2] * S2) * V[j][2]);
}

}  // namespace plade
