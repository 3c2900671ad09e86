// Plane primitive arithmetic of the RANSAC acceptance chain, shared by the device kernels (ransac.cu) and by host-side
// checks (tests/host/planefit_check.cpp): the in-plane frame, the bitmap parametrisation and the least-squares refit.
// Restates R/GfxTL/HyperplaneCoordinateSystem.h:81-93, R/PlanePrimitiveShape.h:97-109, R/PlanePrimitiveShape.cpp:192-207,
// R/Plane.h:66-74 with R/GfxTL/Jacobi.h (R/ = 3rd_party/ransac/ of the reference).  Built with -fmad=false /
// -ffp-contract=off: host and device evaluate these float expressions identically.
#pragma once
#include "linalg.h"

namespace plade {

// HyperplaneCoordinateSystem::FromNormal (R/GfxTL/HyperplaneCoordinateSystem.h:81-93), AutoCAD arbitrary axis
// GfxTL's Normalize (R/GfxTL/MatrixXX.h:80-91): the squared length is summed left to right (linalg.h's normalize() follows Eigen)
PLADE_HD void gfxtl_normalize(V3 &a) {
  float s = 0.f;
  s += a.x * a.x; s += a.y * a.y; s += a.z * a.z;
  if (s == 0.f) return;
  s = sqrtf(s);
  a.x /= s; a.y /= s; a.z /= s;
}
PLADE_HD void frame_from_normal(const float n[3], float u[3], float v[3]) {
  V3 N(n[0], n[1], n[2]), a0;
  if (fabsf(n[0]) < 0.015625f && fabsf(n[1]) < 0.015625f) a0 = cross(V3(0, 1, 0), N);
  else a0 = cross(V3(0, 0, 1), N);
  gfxtl_normalize(a0);
  V3 a1 = cross(N, a0);
  gfxtl_normalize(a1);
  u[0] = a0.x; u[1] = a0.y; u[2] = a0.z;
  v[0] = a1.x; v[1] = a1.y; v[2] = a1.z;
}


// PlanePrimitiveShape::ParametersImpl (R/PlanePrimitiveShape.h:97-109): (u, v) of point p in the frame (pos, a0, a1)
PLADE_HD void plane_uv(const float p[3], const float pos[3], const float a0[3], const float a1[3], float &u, float &v) {
  const float px = p[0] - pos[0], py = p[1] - pos[1], pz = p[2] - pos[2];
  u = (px * a0[0] + py * a0[1]) + pz * a0[2];
  v = (px * a1[0] + py * a1[1]) + pz * a1[2];
}
// BitmapExtent (R/PlanePrimitiveShape.cpp:192-198) and InBitmap (:200-207): pixels per axis, pixel of a parameter
PLADE_HD long long bitmap_extent(float lo, float hi, float eps) {
  long long e = (long long) ceilf((hi - lo) / eps) + 1;
  return e < 2 ? 2 : e;
}
PLADE_HD int bitmap_pixel(float x, float lo, float eps, int extent) {
  int b = (int) floorf((x - lo) / eps);
  b = b < 0 ? 0 : b;
  return b > extent - 1 ? extent - 1 : b;
}

// Jacobi eigen-solver for symmetric 3x3 (float), the textbook cyclic algorithm started from V = I that
// R/GfxTL/Jacobi.h implements; the eigenvector of smallest |eigenvalue| is the plane normal and its
// sign is whatever the rotation sequence produces (it is never flipped afterwards).
PLADE_HD bool jacobi3f(float a[3][3], float d[3], float v[3][3]) {
  float b[3], z[3];
  for (int ip = 0; ip < 3; ++ip) { for (int iq = 0; iq < 3; ++iq) v[ip][iq] = 0.f; v[ip][ip] = 1.f; }
  for (int ip = 0; ip < 3; ++ip) { b[ip] = d[ip] = a[ip][ip]; z[ip] = 0.f; }
  for (int i = 1; i <= 200; ++i) {
    float sm = 0.f;
    for (int ip = 0; ip < 2; ++ip) for (int iq = ip + 1; iq < 3; ++iq) sm += fabsf(a[ip][iq]);
    if (sm == 0.f) return true;
    float tresh = i < 4 ? 0.2f * sm / 9.f : 0.f;
    for (int ip = 0; ip < 2; ++ip)
      for (int iq = ip + 1; iq < 3; ++iq) {
        float g = 100.f * fabsf(a[ip][iq]);
        volatile float t1 = fabsf(d[ip]) + g, t2 = fabsf(d[iq]) + g;
        if (i > 4 && t1 == fabsf(d[ip]) && t2 == fabsf(d[iq])) a[ip][iq] = 0.f;
        else if (fabsf(a[ip][iq]) > tresh) {
          float h = d[iq] - d[ip], t;
          volatile float t3 = fabsf(h) + g;
          if (t3 == fabsf(h)) t = a[ip][iq] / h;
          else {
            float theta = 0.5f * h / a[ip][iq];
            t = 1.f / (fabsf(theta) + sqrtf(1.f + theta * theta));
            if (theta < 0.f) t = -t;
          }
          float c = 1.f / sqrtf(1.f + t * t), s = t * c, tau = s / (1.f + c);
          h = t * a[ip][iq];
          z[ip] -= h; z[iq] += h; d[ip] -= h; d[iq] += h;
          a[ip][iq] = 0.f;
#define PLADE_JROT(m, i1, j1, i2, j2) { float gg = m[i1][j1], hh = m[i2][j2]; m[i1][j1] = gg - s * (hh + gg * tau); m[i2][j2] = hh + s * (gg - hh * tau); }
          for (int j = 0; j <= ip - 1; ++j) PLADE_JROT(a, j, ip, j, iq)
          for (int j = ip + 1; j <= iq - 1; ++j) PLADE_JROT(a, ip, j, j, iq)
          for (int j = iq + 1; j < 3; ++j) PLADE_JROT(a, ip, j, iq, j)
          for (int j = 0; j < 3; ++j) PLADE_JROT(v, j, ip, j, iq)
#undef PLADE_JROT
        }
      }
    for (int ip = 0; ip < 3; ++ip) { b[ip] += z[ip]; d[ip] = b[ip]; z[ip] = 0.f; }
  }
  return false;
}

// result of one full evaluation of a plane: support of the largest connected component, gaussian-weighted score,
// position sums of its members
struct Eval { long long size; double score; double sum[3]; bool ok; };

// Plane::LeastSquaresFit (R/Plane.h:66-74) from the member sums: mean, covariance about float(mean), eigenvector
// of the smallest |eigenvalue| (sign as the Jacobi rotations leave it)
PLADE_HD bool fit_plane_from_cov(const Eval &e, const double h[6], float nrm3[3], float pos3[3]) {
  float a[3][3], d[3], v[3][3];
  a[0][0] = (float) (h[0] / e.size); a[0][1] = a[1][0] = (float) (h[1] / e.size); a[0][2] = a[2][0] = (float) (h[2] / e.size);
  a[1][1] = (float) (h[3] / e.size); a[1][2] = a[2][1] = (float) (h[4] / e.size); a[2][2] = (float) (h[5] / e.size);
  if (!jacobi3f(a, d, v)) return false;
  int k = 0;
  for (int j = 1; j < 3; ++j) if (fabsf(d[j]) < fabsf(d[k])) k = j;
  nrm3[0] = v[0][k]; nrm3[1] = v[1][k]; nrm3[2] = v[2][k];
  pos3[0] = (float) (e.sum[0] / e.size); pos3[1] = (float) (e.sum[1] / e.size); pos3[2] = (float) (e.sum[2] / e.size);
  return true;
}


}  // namespace plade
