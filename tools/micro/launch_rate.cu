// Diagnostic: aggregate kernel-launch rate of T host threads, each on its own stream of ONE CUDA context,
// with a stream synchronisation every `burst` launches (the registration pipeline's pattern).
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
__global__ void tiny(int *p) { if (p && threadIdx.x == 9999) *p = 1; }
int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 20000, burst = argc > 2 ? atoi(argv[2]) : 8;
  cudaFree(0);
  for (int T : {1, 2, 4, 8, 12}) {
    std::vector<cudaStream_t> st(T);
    for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] {
      for (int i = 0; i < n; ++i) { tiny<<<148, 256, 0, st[t]>>>(nullptr); if ((i + 1) % burst == 0) cudaStreamSynchronize(st[t]); }
      cudaStreamSynchronize(st[t]);
    });
    for (auto &x : th) x.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("threads %2d: %.0f launches/s aggregate (%.2f us per launch per thread), sync every %d\n", T, T * n / dt, 1e6 * dt / n, burst);
    for (auto &s : st) cudaStreamDestroy(s);
  }
  return 0;
}
