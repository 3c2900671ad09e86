// The reference's float linear solver, restated for host and device.
//
// cv::solve(A, B, X, DECOMP_SVD) on CV_32F (OpenCV 2.4.13.6 modules/core/src/lapack.cpp: solve() :1425-1447,
// JacobiSVDImpl_ :533-710, SVBkSbImpl_ :750-812) = one-sided (Hestenes) Jacobi SVD of A^T followed by a
// back-substitution.  Every operation keeps the reference's order and float/double mix; the library is built
// without FMA contraction, so device and host evaluate it identically (only IEEE + - * / sqrt are used).
// Not reproduced: the random-vector completion for an exactly singular system (needs sigma <= FLT_MIN).
#pragma once
#include <cfloat>
#include <cmath>
#include "linalg.h"

namespace plade {

// cv's own hypot template (lapack.cpp:214-229), which JacobiSVDImpl_ resolves to inside namespace cv
PLADE_HD double hypot_cv(double a, double b) {
  a = fabs(a);
  b = fabs(b);
  if (a > b) { b /= a; return a * sqrt(1 + b * b); }
  if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
  return 0;
}

// x = argmin |A x - b| for the M x N system whose TRANSPOSE is passed in At (destroyed).
template <int M, int N>
PLADE_HD void jacobi_svd_solve(float At[N][M] /* A transposed */, const float b[M], float x[N]) {
  float Vt[N][N];
  double W[N];
  const float eps = FLT_EPSILON * 2;
  for (int i = 0; i < N; ++i) {
    double sd = 0;
    for (int k = 0; k < M; ++k) { float t = At[i][k]; sd += (double) t * t; }
    W[i] = sd;
    for (int k = 0; k < N; ++k) Vt[i][k] = 0.f;
    Vt[i][i] = 1.f;
  }
  const int max_iter = M > 30 ? M : 30;
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < N - 1; ++i)
      for (int j = i + 1; j < N; ++j) {
        double a = W[i], p = 0, bb = W[j];
        for (int k = 0; k < M; ++k) p += (double) At[i][k] * At[j][k];
        if (fabs(p) <= eps * sqrt(a * bb)) continue;
        p *= 2;
        double beta = a - bb, gamma = hypot_cv(p, beta);
        float c, s;
        if (beta < 0) {
          double delta = (gamma - beta) * 0.5;
          s = (float) sqrt(delta / gamma);
          c = (float) (p / (gamma * s * 2));
        } else {
          c = (float) sqrt((gamma + beta) / (gamma * 2));
          s = (float) (p / (gamma * c * 2));
        }
        a = bb = 0;
        for (int k = 0; k < M; ++k) {
          float t0 = c * At[i][k] + s * At[j][k];
          float t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          a += (double) t0 * t0; bb += (double) t1 * t1;
        }
        W[i] = a; W[j] = bb;
        changed = true;
        for (int k = 0; k < N; ++k) {
          float t0 = c * Vt[i][k] + s * Vt[j][k];
          float t1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = t0; Vt[j][k] = t1;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < N; ++i) {
    double sd = 0;
    for (int k = 0; k < M; ++k) { float t = At[i][k]; sd += (double) t * t; }
    W[i] = sqrt(sd);
  }
  for (int i = 0; i < N - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < N; ++k) if (W[j] < W[k]) j = k;
    if (i != j) {
      double tw = W[i]; W[i] = W[j]; W[j] = tw;
      for (int k = 0; k < M; ++k) { float t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; }
      for (int k = 0; k < N; ++k) { float t = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = t; }
    }
  }
  float w[N];
  for (int i = 0; i < N; ++i) w[i] = (float) W[i];
  // left singular vectors: rows of At scaled by 1 / sigma
  for (int i = 0; i < N; ++i) {
    double sd = W[i];
    if (sd <= (double) FLT_MIN) { for (int k = 0; k < M; ++k) At[i][k] = 0.f; continue; }
    float s = (float) (1 / sd);
    for (int k = 0; k < M; ++k) At[i][k] = At[i][k] * s;
  }
  // back substitution x = V diag(1/w) U^T b  (SVBkSbImpl_, nb == 1)
  double threshold = 0;
  for (int i = 0; i < N; ++i) x[i] = 0.f;
  const int nm = M < N ? M : N;
  for (int i = 0; i < nm; ++i) threshold += w[i];
  threshold *= (float) (DBL_EPSILON * 2);
  for (int i = 0; i < nm; ++i) {
    double wi = w[i];
    if (fabs(wi) <= threshold) continue;
    wi = 1 / wi;
    double s = 0;
    for (int j = 0; j < M; ++j) { float prod = At[i][j] * b[j]; s += prod; }
    s *= wi;
    for (int j = 0; j < N; ++j) x[j] = (float) (x[j] + s * Vt[i][j]);
  }
}

// ComputeNearstTwoPointsOfTwo3DLine's system (PLADE/util.cpp:1183-1226) for normalised directions v1, v2:
//   p1' - t1 v1 = p1,  p2' - t2 v2 = p2,  p2' - p1' - t3 (v1 x v2)/|v1 x v2| = 0   ->  point1 = p1', point2 = p2'
PLADE_HD void nearest_points_cv_solve(const V3 &v1, const V3 &p1, const V3 &v2, const V3 &p2, V3 &point1, V3 &point2) {
  V3 dv = cross(v1, v2);
  normalize(dv);
  float At[9][9];
  for (int i = 0; i < 9; ++i) for (int k = 0; k < 9; ++k) At[i][k] = 0.f;
  // A(row, col) as written at util.cpp:1191-1211, stored transposed: At[col][row]
  At[0][0] = 1; At[3][0] = -v1.x;
  At[1][1] = 1; At[3][1] = -v1.y;
  At[2][2] = 1; At[3][2] = -v1.z;
  At[4][3] = 1; At[7][3] = -v2.x;
  At[5][4] = 1; At[7][4] = -v2.y;
  At[6][5] = 1; At[7][5] = -v2.z;
  At[0][6] = -1; At[4][6] = 1; At[8][6] = -dv.x;
  At[1][7] = -1; At[5][7] = 1; At[8][7] = -dv.y;
  At[2][8] = -1; At[6][8] = 1; At[8][8] = -dv.z;
  float b[9] = {p1.x, p1.y, p1.z, p2.x, p2.y, p2.z, 0.f, 0.f, 0.f};
  float x[9];
  jacobi_svd_solve<9, 9>(At, b, x);
  point1 = V3(x[0], x[1], x[2]);
  point2 = V3(x[4], x[5], x[6]);
}

// ComputeIntersectionPointOf23DLine (PLADE/util.cpp:1461-1500): least-squares "intersection" point of two lines
// from the reference's 6x5 float cv::solve(DECOMP_SVD):  x - t1 v1 = p1,  x - t2 v2 = p2.
PLADE_HD int line_line_point_cv(const V3 &v1, const V3 &p1, const V3 &v2, const V3 &p2, V3 &out) {
  if (fabsf(dot(v1, v2)) > 0.9999) return -1;      // parallel
  float At[5][6];
  for (int i = 0; i < 5; ++i) for (int k = 0; k < 6; ++k) At[i][k] = 0.f;
  At[0][0] = 1; At[3][0] = -v1.x;
  At[1][1] = 1; At[3][1] = -v1.y;
  At[2][2] = 1; At[3][2] = -v1.z;
  At[0][3] = 1; At[4][3] = -v2.x;
  At[1][4] = 1; At[4][4] = -v2.y;
  At[2][5] = 1; At[4][5] = -v2.z;
  float b[6] = {p1.x, p1.y, p1.z, p2.x, p2.y, p2.z};
  float x[5];
  jacobi_svd_solve<6, 5>(At, b, x);
  out = V3(x[0], x[1], x[2]);
  return 0;
}

}  // namespace plade
