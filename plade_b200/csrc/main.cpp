// CLI with the reference's two usages and result-file format (PLADE/main.cpp:30-160), on the C ABI.
#include "plade_b200.h"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <algorithm>
#include <sstream>
#include <string>
#include <thread>
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
  float T[16];
  if (argc == 4) {
    plade_ctx *ctx = plade_ctx_create(-1);
    if (!ctx) { std::cerr << "no usable CUDA device: " << plade_create_error() << std::endl; return EXIT_FAILURE; }
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
  if (!input.is_open()) { std::cerr << "failed opening the file containing pairs of point cloud names: " << argv[1] << std::endl; return EXIT_FAILURE; }
  std::ofstream output(argv[2]);
  if (!output.is_open()) { std::cerr << "failed opening the result file: " << argv[2] << std::endl; return EXIT_FAILURE; }
  // The pair list is read first, exactly as PLADE/main.cpp:117-134 walks it: one name per line, empty lines and
  // names that cannot be opened are skipped ("file doesn't exist"), every two surviving names form a pair, a
  // trailing single name is dropped.  The pairs are then spread over the GPUs (PLADE_DEVICES=n limits how many;
  // default: all visible) and the results written in list order with the reference's text format (:136-147).
  auto is_file = [](const std::string &filename) -> bool { std::ifstream fin(filename); return fin.is_open(); };
  std::vector<std::string> tf, sf;
  while (!input.eof()) {
    std::vector<std::string> file_pair;
    while (!input.eof() && file_pair.size() < 2) {
      std::string file_name;
      std::getline(input, file_name);
      if (!file_name.empty()) {
        if (is_file(file_name)) file_pair.push_back(file_name);
        else std::cerr << "file doesn't exist: " << file_name << std::endl;
      }
    }
    if (file_pair.size() == 2) { tf.push_back(file_pair[0]); sf.push_back(file_pair[1]); }
  }
  int n_dev = plade_device_count();
  if (n_dev < 1) { std::cerr << "no usable CUDA device (plade_b200 has no CPU fallback)" << std::endl; return EXIT_FAILURE; }
  if (const char *e = getenv("PLADE_DEVICES")) n_dev = std::max(1, std::min(n_dev, atoi(e)));
  // one registration cannot fill a B200 (it is a chain of small launches with host decisions in between), so every
  // GPU gets several workers (PLADE_WORKERS_PER_GPU, default min(4, host cores / (2 x GPUs)): a worker keeps two host threads busy)
  int per_gpu = std::max(1, std::min(4, (int) std::thread::hardware_concurrency() / (2 * n_dev)));
  if (const char *e = getenv("PLADE_WORKERS_PER_GPU")) per_gpu = std::max(1, std::min(16, atoi(e)));
  const int n = (int) tf.size();
  std::vector<int> devices;
  for (int w = 0; w < per_gpu; ++w) for (int g = 0; g < n_dev; ++g) devices.push_back(g);
  if ((int) devices.size() > std::max(n, 1)) devices.resize(std::max(n, 1));
  std::vector<const char *> tp(n), sp(n);
  for (int i = 0; i < n; ++i) { tp[i] = tf[i].c_str(); sp[i] = sf[i].c_str(); }
  std::vector<float> Ts(16 * (size_t) std::max(n, 1));
  std::vector<int> oks(std::max(n, 1), 0);
  if (plade_register_batch(devices.data(), (int) devices.size(), tp.data(), sp.data(), n, Ts.data(), oks.data()) < 0) return EXIT_FAILURE;
  int count_success = 0, count_failure = 0;
  for (int i = 0; i < n; ++i) {
    output << "target: " << tf[i] << std::endl;
    output << "source: " << sf[i] << std::endl;
    if (oks[i]) { output << "transformation:\n"; ++count_success; }
    else { output << "registration failed, an identity matrix is recorded:\n"; ++count_failure; }
    write_matrix(output, &Ts[16 * (size_t) i]);
    output << std::endl << std::endl;
  }
  if (count_success == 0) {
    std::cerr << "registration all failed (" << count_failure << " pairs)" << std::endl;
    return EXIT_FAILURE;
  }
  if (count_failure > 0)
    std::cerr << "registration of " << count_failure << " (out of " << count_failure + count_success << ") pairs failed" << std::endl;
  std::cout << "the registration result has been written into file: " << argv[2] << std::endl;
  return EXIT_SUCCESS;
}
