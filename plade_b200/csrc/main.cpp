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
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

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

// --dump-planes: the planes extract() finds in a cloud as an ASCII .vg file (the reference's save_vg, PLADE/util.cpp:1553-1616)
static void dump_planes(plade_ctx *ctx, const char *ply, const std::string &out) {
  const long long n = plade_ply_read(ply, nullptr, 0);
  if (n <= 0) { std::cerr << "--dump-planes: cannot read " << ply << std::endl; return; }
  std::vector<float> cloud((size_t) n * 6);
  if (plade_ply_read(ply, cloud.data(), (size_t) n) != n) return;
  if (plade_extract_planes(ctx, cloud.data(), (size_t) n, 10000) < 0) { std::cerr << "--dump-planes: " << plade_last_error(ctx) << std::endl; return; }
  int np = 0;
  long long ni = 0;
  plade_planes_size(ctx, &np, &ni);
  std::vector<int> off(np + 1), idx((size_t) std::max<long long>(ni, 1));
  std::vector<float> par((size_t) std::max(np, 1) * 4);
  plade_planes_get(ctx, off.data(), idx.data(), par.data());
  if (plade_dump_planes_vg(cloud.data(), (size_t) n, off.data(), idx.data(), par.data(), np, out.c_str()))
    std::cout << np << " planes of " << ply << " written into file: " << out << std::endl;
}

// --shard-verify: ONE pair on all visible GPUs (PLADE_DEVICES=n limits how many): every GPU extracts the planes and builds the
// hypotheses (replicas), the verification loop (PLADE/plade.cpp:547-564) is split over the GPUs by hypothesis and the winner
// agreed with one ncclAllReduce(ncclUint64, ncclMax) inside the library (plade_shard_init_nccl_all)
static int register_sharded(const char *tgt, const char *src, float T[16], std::string *report) {
  int n_dev = plade_device_count();
  if (const char *e = getenv("PLADE_DEVICES")) n_dev = std::max(1, std::min(n_dev, atoi(e)));
  if (n_dev < 1) { std::cerr << "no usable CUDA device" << std::endl; return 0; }
  std::vector<plade_ctx *> ctxs(n_dev, nullptr);
  for (int d = 0; d < n_dev; ++d) if (!(ctxs[d] = plade_ctx_create(d))) { std::cerr << "no usable CUDA device: " << plade_create_error() << std::endl; return 0; }
  if (n_dev > 1 && !plade_shard_init_nccl_all(ctxs.data(), n_dev)) { std::cerr << "NCCL: " << plade_create_error() << std::endl; return 0; }
  std::vector<int> oks(n_dev, 0);
  std::vector<float> Ts((size_t) n_dev * 16);
  std::vector<std::thread> th;
  for (int d = 0; d < n_dev; ++d) th.emplace_back([&, d] { oks[d] = plade_register_files(ctxs[d], tgt, src, &Ts[(size_t) d * 16]); });
  for (auto &t : th) t.join();
  for (int i = 0; i < 16; ++i) T[i] = Ts[i];
  if (report) *report = plade_last_report(ctxs[0]);
  const int ok = oks[0];
  for (int d = 0; d < n_dev; ++d) { plade_shard_finalize(ctxs[d]); plade_ctx_destroy(ctxs[d]); }
  return ok;
}

int main(int argc0, char **argv0) {
  // options (anywhere on the command line); what remains are the reference's positional arguments
  std::string report_file, dump_prefix;
  bool shard_verify = false;
  std::vector<char *> pos;
  for (int i = 0; i < argc0; ++i) {
    const std::string a = argv0[i];
    if (a == "--report" && i + 1 < argc0) report_file = argv0[++i];
    else if (a == "--dump-planes" && i + 1 < argc0) dump_prefix = argv0[++i];
    else if (a == "--shard-verify") shard_verify = true;
    else pos.push_back(argv0[i]);
  }
  const int argc = (int) pos.size();
  char **argv = pos.data();
  if (argc != 3 && argc != 4) {
    std::cerr << "PLADE (B200) registers two point clouds dominated by planar structures.\n"
              << "Usage 1:  plade_b200_cli  target.ply  source.ply  result.txt\n"
              << "Usage 2:  plade_b200_cli  file_pairs.txt  results.txt   (two lines per pair: target, source)\n"
              << "Options (usage 1):  --report file.json   planes, hypotheses, winner, inlier count / overlap ratio of the registration\n"
              << "                    --dump-planes prefix  the extracted planes as prefix_target.vg / prefix_source.vg\n"
              << "                    --shard-verify        verify the hypotheses on all visible GPUs (NCCL max-allreduce)\n";
    return EXIT_FAILURE;
  }
  float T[16];
  if (argc == 4) {
    plade_ctx *ctx = plade_ctx_create(-1);
    if (!ctx) { std::cerr << "no usable CUDA device: " << plade_create_error() << std::endl; return EXIT_FAILURE; }
    std::ofstream output(argv[3]);
    if (!output.is_open()) { std::cerr << "failed opening the result file: " << argv[3] << std::endl; return EXIT_FAILURE; }
    std::string report;
    int ok = shard_verify ? register_sharded(argv[1], argv[2], T, &report) : plade_register_files(ctx, argv[1], argv[2], T);
    if (!shard_verify) report = plade_last_report(ctx);
    if (!report_file.empty() && ok) { std::ofstream rf(report_file); rf << report << std::endl; }
    if (!dump_prefix.empty()) { dump_planes(ctx, argv[1], dump_prefix + "_target.vg"); dump_planes(ctx, argv[2], dump_prefix + "_source.vg"); }
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
  // Default: all GPUs from this one process (plade_register_batch: worker threads + contexts per GPU).  PLADE_CLI_FORK=1 splits
  // the pairs over ONE CHILD PROCESS PER GPU instead (pair p -> GPU p mod n, each child sees only its GPU); the parent then
  // never touches CUDA (fork after CUDA initialisation is not allowed).  Measured on an 8-GPU box: the first CUDA call of a
  // process costs 7-11 s there whatever it can see, so the children do not start faster than the single process does
  // (64 pairs: 14.7 s forked, 11.4 s in one process, 0.6 s in a warm process) -- the option is for isolation, not speed.
  // The GPUs this process may use, WITHOUT initialising CUDA here (on an 8-GPU box cuInit alone takes ~10 s per process when
  // all GPUs are visible): the entries of CUDA_VISIBLE_DEVICES if it is set, else the /dev/nvidia<N> device nodes; only when
  // neither says anything, a short-lived probe child asks the runtime.
  std::vector<std::string> visible;
  if (const char *cvd = getenv("CUDA_VISIBLE_DEVICES")) {
    std::stringstream ss(cvd);
    std::string tok;
    while (std::getline(ss, tok, ',')) if (!tok.empty()) visible.push_back(tok);
  } else {
    // (the minor numbers of the nodes need not start at 0 in a container; CUDA numbers the devices it can see from 0)
    int nodes = 0;
    for (int k = 0; k < 256; ++k) if (access(("/dev/nvidia" + std::to_string(k)).c_str(), F_OK) == 0) ++nodes;
    for (int k = 0; k < nodes; ++k) visible.push_back(std::to_string(k));
  }
  int n_dev = (int) visible.size();
  if (n_dev == 0) {
    const pid_t probe = fork();
    if (probe == 0) _exit(std::min(plade_device_count(), 255));
    int status = 0;
    if (probe > 0 && waitpid(probe, &status, 0) == probe && WIFEXITED(status)) n_dev = WEXITSTATUS(status);
    for (int k = 0; k < n_dev; ++k) visible.push_back(std::to_string(k));
  }
  if (n_dev < 1) { std::cerr << "no usable CUDA device (plade_b200 has no CPU fallback)" << std::endl; return EXIT_FAILURE; }
  if (const char *e = getenv("PLADE_DEVICES")) n_dev = std::max(1, std::min(n_dev, atoi(e)));
  // one registration cannot fill a B200 (it is a chain of small launches with host decisions in between), so every
  // GPU gets several workers (PLADE_WORKERS_PER_GPU, default min(4, host cores / (2 x GPUs)): a worker keeps two host threads busy)
  int per_gpu = std::max(1, std::min(4, (int) std::thread::hardware_concurrency() / (2 * n_dev)));
  if (const char *e = getenv("PLADE_WORKERS_PER_GPU")) per_gpu = std::max(1, std::min(16, atoi(e)));
  const int n = (int) tf.size();
  std::vector<const char *> tp(n), sp(n);
  for (int i = 0; i < n; ++i) { tp[i] = tf[i].c_str(); sp[i] = sf[i].c_str(); }
  std::vector<float> Ts(16 * (size_t) std::max(n, 1));
  std::vector<int> oks(std::max(n, 1), 0);
  const char *fk = getenv("PLADE_CLI_FORK");
  const bool use_fork = n_dev > 1 && n > 1 && fk && atoi(fk) != 0;
  if (!use_fork) {
    std::vector<int> devices;
    for (int w = 0; w < per_gpu; ++w) for (int g = 0; g < n_dev; ++g) devices.push_back(g);
    if ((int) devices.size() > std::max(n, 1)) devices.resize(std::max(n, 1));
    if (plade_register_batch(devices.data(), (int) devices.size(), tp.data(), sp.data(), n, Ts.data(), oks.data()) < 0) return EXIT_FAILURE;
  } else {
    for (int i = 0; i < n; ++i) for (int k = 0; k < 16; ++k) Ts[16 * (size_t) i + k] = (k % 5 == 0) ? 1.f : 0.f;
    struct Child { pid_t pid; int fd; std::vector<int> mine; };
    std::vector<Child> kids;
    for (int g = 0; g < n_dev; ++g) {
      Child c;
      for (int i = g; i < n; i += n_dev) c.mine.push_back(i);
      if (c.mine.empty()) continue;
      int fds[2];
      if (pipe(fds) != 0) { std::cerr << "pipe() failed" << std::endl; return EXIT_FAILURE; }
      c.pid = fork();
      if (c.pid == 0) {      // child: its share of the list on its GPU (the only one it sees), results back through the pipe
        close(fds[0]);
        setenv("CUDA_VISIBLE_DEVICES", visible[g].c_str(), 1);
        const int m = (int) c.mine.size();
        std::vector<const char *> ct(m), cs(m);
        for (int j = 0; j < m; ++j) { ct[j] = tp[c.mine[j]]; cs[j] = sp[c.mine[j]]; }
        std::vector<int> devices(std::min(per_gpu, m), 0);
        std::vector<float> cT(16 * (size_t) m);
        std::vector<int> cok(m, 0);
        const int rc = plade_register_batch(devices.data(), (int) devices.size(), ct.data(), cs.data(), m, cT.data(), cok.data());
        bool sent = write(fds[1], cok.data(), sizeof(int) * m) == (ssize_t) (sizeof(int) * m) &&
                    write(fds[1], cT.data(), sizeof(float) * 16 * m) == (ssize_t) (sizeof(float) * 16 * m);
        close(fds[1]);
        _exit(rc < 0 || !sent ? 1 : 0);
      }
      close(fds[1]);
      c.fd = fds[0];
      if (c.pid < 0) { std::cerr << "fork() failed" << std::endl; return EXIT_FAILURE; }
      kids.push_back(c);
    }
    auto read_all = [](int fd, void *buf, size_t nbytes) {
      char *p = static_cast<char *>(buf);
      while (nbytes) { const ssize_t r = read(fd, p, nbytes); if (r <= 0) return false; p += r; nbytes -= (size_t) r; }
      return true;
    };
    for (Child &c : kids) {
      const int m = (int) c.mine.size();
      std::vector<int> cok(m, 0);
      std::vector<float> cT(16 * (size_t) m);
      const bool got = read_all(c.fd, cok.data(), sizeof(int) * m) && read_all(c.fd, cT.data(), sizeof(float) * 16 * m);
      close(c.fd);
      int status = 0;
      waitpid(c.pid, &status, 0);
      if (!got) { std::cerr << "a GPU worker process ended without results (" << m << " pairs recorded as failed)" << std::endl; continue; }
      for (int j = 0; j < m; ++j) {
        oks[c.mine[j]] = cok[j];
        for (int k = 0; k < 16; ++k) Ts[16 * (size_t) c.mine[j] + k] = cT[16 * (size_t) j + k];
      }
    }
  }
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
