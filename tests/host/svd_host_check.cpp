// Test helper (not part of the product): runs plade_b200/csrc/svdsolve.h on the HOST so that
// tests/test_oracle_cpu.py can compare the restated float SVD solver with the compiled reference without a GPU.
// stdin (binary): int mode, int n, then n records of 12 floats (v1 p1 v2 p2 per line pair; 18 floats in mode 3).
//   mode 0: closest points of two lines (9x9 solve)  -> n*6 floats (point1, point2); directions get normalised
//   mode 1: least-squares intersection (6x5 solve)   -> n*4 floats (rc, x, y, z); directions used as given
//   mode 2: symmetric 3x3 eigen solver (linalg.h sym_eig3f_eigen); the first 9 floats of each record are A
//           (row-major) -> n*12 floats (w[3], V[9] row-major)
//   mode 3: umeyama.h, record = sv1 sv2 dv1 dv2 sp tp (18 floats) -> n*12 floats (R row-major, T) as
//           ComputeTransformationUsingTwoVecAndOnePoint (PLADE/util.cpp:604-624)
//   mode 4: libm_flt32.h against the host libm; no records; prints the number of mismatches over a fixed sweep
#include <cstdio>
#include <vector>
#include "svdsolve.h"
#include "umeyama.h"
#include "libm_flt32.h"

int main() {
  int mode = 0, n = 0;
  if (fread(&mode, sizeof(int), 1, stdin) != 1 || fread(&n, sizeof(int), 1, stdin) != 1 || n < 0) return 1;
  if (mode == 4) {
    unsigned long long st = 88172645463325252ull, bad = 0;
    auto rnd = [&] { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    for (int i = 0; i < n; ++i) {
      float a = (float) ((double) (rnd() >> 11) / 9007199254740992.0 * 2 - 1), b = (float) ((double) (rnd() >> 11) / 9007199254740992.0 * 2 - 1);
      if ((i & 7) == 0) a *= 1e-3f;
      if ((i & 15) == 1) b *= 1e-4f;
      volatile float r1 = atan2f(a, b), r2 = asinf(a), r3 = atanf(a * 4);
      bad += r1 != plade::atan2f_glibc(a, b);
      bad += r2 != plade::asinf_glibc(a);
      bad += r3 != plade::atanf_glibc(a * 4);
    }
    printf("%llu\n", bad);
    return 0;
  }
  const int width = mode == 3 ? 18 : 12;
  std::vector<float> in((size_t) n * width), out;
  if (n && fread(in.data(), sizeof(float), in.size(), stdin) != in.size()) return 1;
  for (int i = 0; i < n; ++i) {
    const float *q = &in[(size_t) i * width];
    if (mode == 3) {
      plade::V3 sv1(q[0], q[1], q[2]), sv2(q[3], q[4], q[5]), dv1(q[6], q[7], q[8]), dv2(q[9], q[10], q[11]);
      plade::V3 s3 = plade::cross(sv1, sv2), d3 = plade::cross(dv1, dv2);
      const float src[3][3] = {{sv1.x, sv1.y, sv1.z}, {sv2.x, sv2.y, sv2.z}, {s3.x, s3.y, s3.z}};
      const float dst[3][3] = {{dv1.x, dv1.y, dv1.z}, {dv2.x, dv2.y, dv2.z}, {d3.x, d3.y, d3.z}};
      plade::M3 R;
      float Rm[3][3];
      plade::umeyama3_rotation_eigen(src, dst, Rm);
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R(r, c) = Rm[r][c];
      plade::V3 T = plade::V3(q[15], q[16], q[17]) - plade::mul(R, plade::V3(q[12], q[13], q[14]));
      out.insert(out.end(), R.m, R.m + 9);
      out.insert(out.end(), {T.x, T.y, T.z});
      continue;
    }
    plade::V3 v1(q[0], q[1], q[2]), p1(q[3], q[4], q[5]), v2(q[6], q[7], q[8]), p2(q[9], q[10], q[11]), a, b;
    if (mode == 0) {
      plade::normalize(v1);
      plade::normalize(v2);
      plade::nearest_points_cv_solve(v1, p1, v2, p2, a, b);
      out.insert(out.end(), {a.x, a.y, a.z, b.x, b.y, b.z});
    } else if (mode == 2) {
      float A[3][3], w[3], V[3][3];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A[r][c] = q[3 * r + c];
      plade::sym_eig3f_eigen(A, w, V);
      out.insert(out.end(), w, w + 3);
      for (int r = 0; r < 3; ++r) out.insert(out.end(), V[r], V[r] + 3);
    } else {
      int rc = plade::line_line_point_cv(v1, p1, v2, p2, a);
      out.insert(out.end(), {(float) rc, a.x, a.y, a.z});
    }
  }
  fwrite(out.data(), sizeof(float), out.size(), stdout);
  return 0;
}
