// CLI with the reference's two usages and result-file format (PLADE/main.cpp:30-160), on the C ABI.
#include "plade_b200.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

// Eigen's default operator<< (IOFormat(): StreamPrecision, columns separated by one space, rows by
// newline, every coefficient right-aligned to the widest one) as PLADE/main.cpp:86 prints the 4x4.
static void write_matrix(std::ostream &o, const float m[16]) {
  std::string s[16];
  size_t w = 0;
  for (int i = 0; i < 16; ++i) {
    std::ostringstream ss;
    ss.copyfmt(o);
    ss << m[i];
    s[i] = ss.str();
    w = std::max(w, s[i].size());
  }
  for (int r = 0; r < 4; ++r) {
    if (r) o << "\n";
    for (int c = 0; c < 4; ++c) {
      if (c) o << " ";
      o << std::string(w - s[4 * r + c].size(), ' ') << s[4 * r + c];
    }
  }
}

int main(int argc, char **argv) {
  if (argc != 3 && argc != 4) {
    std::cerr << "PLADE (B200) registers two point clouds dominated by planar structures.\n"
              << "Usage 1:  plade_b200_cli  target.ply  source.ply  result.txt\n"
              << "Usage 2:  plade_b200_cli  file_pairs.txt  results.txt   (two lines per pair: target, source)\n";
    return EXIT_FAILURE;
  }
  plade_ctx *ctx = plade_ctx_create(-1);
  if (!ctx) { std::cerr << "no usable CUDA device: " << plade_create_error() << std::endl; return EXIT_FAILURE; }
  float T[16];
  if (argc == 4) {
    std::ofstream output(argv[3]);
    if (!output.is_open()) { std::cerr << "failed opening the result file: " << argv[3] << std::endl; return EXIT_FAILURE; }
    int ok = plade_register_files(ctx, argv[1], argv[2], T);
    if (ok) {
      output << "target: " << argv[1] << "\nsource: " << argv[2] << "\ntransformation:\n";
      write_matrix(output, T);
      output << std::endl;
      std::cout << "the registration result has been written into file: " << argv[3] << std::endl;
      plade_ctx_destroy(ctx);
      return EXIT_SUCCESS;
    }
    output << "registration failed, an identity matrix is recorded:\n";
    write_matrix(output, T);
    output << std::endl;
    plade_ctx_destroy(ctx);
    return EXIT_FAILURE;
  }
  std::ifstream input(argv[1]);
  if (!input.is_open()) { std::cerr << "failed opening the file pairs file: " << argv[1] << std::endl; return EXIT_FAILURE; }
  std::ofstream output(argv[2]);
  if (!output.is_open()) { std::cerr << "failed opening the result file: " << argv[2] << std::endl; return EXIT_FAILURE; }
  std::string t, s;
  int success = 0, fail = 0;
  while (!input.eof()) {
    std::getline(input, t);
    if (t.empty()) continue;
    std::getline(input, s);
    if (s.empty()) continue;
    int ok = plade_register_files(ctx, t.c_str(), s.c_str(), T);
    output << "target: " << t << "\nsource: " << s << "\n";
    if (ok) { output << "transformation:\n"; ++success; }
    else { output << "registration failed, an identity matrix is recorded:\n"; ++fail; }
    write_matrix(output, T);
    output << "\n" << std::endl;
  }
  std::cout << success << " pairs succeeded, " << fail << " failed. results written into: " << argv[2] << std::endl;
  plade_ctx_destroy(ctx);
  return fail ? EXIT_FAILURE : EXIT_SUCCESS;
}
