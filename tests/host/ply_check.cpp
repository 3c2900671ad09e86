// Host-only check of the PLY ingest (plade_b200/csrc/ply.cpp): reads the file given on the command line and prints
// "ok <n>" followed by the n x 6 floats, or "fail".  Built and driven by tests/test_oracle_cpu.py.
#include "ply.h"
#include <cstdio>
int main(int argc, char **argv) {
  if (argc < 2) return 2;
  std::vector<float> v;
  if (!plade::load_ply_xyzn(argv[1], v)) { std::printf("fail\n"); return 0; }
  std::printf("ok %zu\n", v.size() / 6);
  for (size_t i = 0; i < v.size(); ++i) std::printf("%.9g%c", v[i], (i % 6 == 5) ? '\n' : ' ');
  return 0;
}
