// Test helper (not part of the product): runs plade_b200/csrc/svdsolve.h on the HOST for line pairs read from
// stdin (binary: int n, then n*12 floats) and writes n*6 floats (point1, point2) to stdout, so that
// tests/test_oracle_cpu.py can compare the restated solver with the compiled reference without a GPU.
#include <cstdio>
#include <vector>
#include "svdsolve.h"

int main() {
  int n = 0;
  if (fread(&n, sizeof(int), 1, stdin) != 1 || n < 0) return 1;
  std::vector<float> in((size_t) n * 12), out((size_t) n * 6);
  if (n && fread(in.data(), sizeof(float), in.size(), stdin) != in.size()) return 1;
  for (int i = 0; i < n; ++i) {
    const float *q = &in[(size_t) i * 12];
    plade::V3 v1(q[0], q[1], q[2]), v2(q[6], q[7], q[8]), a, b;
    plade::normalize(v1);
    plade::normalize(v2);
    plade::nearest_points_cv_solve(v1, plade::V3(q[3], q[4], q[5]), v2, plade::V3(q[9], q[10], q[11]), a, b);
    float *o = &out[(size_t) i * 6];
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z;
  }
  fwrite(out.data(), sizeof(float), out.size(), stdout);
  return 0;
}
