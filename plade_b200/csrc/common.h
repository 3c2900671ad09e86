// Shared host/device plumbing for the plade_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

namespace plade {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};

#define PLADE_CUDA(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      throw ::plade::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +   \
                               __FILE__ + ":" + std::to_string(__LINE__));                     \
  } while (0)

#define PLADE_LAUNCH_CHECK() PLADE_CUDA(cudaGetLastError())

// Grow-only device buffer; lives in the context so steady-state registration allocates nothing.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { if (p) cudaFree(p); }
  T *ensure(size_t n) {
    if (n > cap) {
      if (p) PLADE_CUDA(cudaFree(p));
      p = nullptr;
      size_t want = n + n / 8 + 64;
      PLADE_CUDA(cudaMalloc(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Pinned host staging buffer (grow-only).
template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  PinBuf() = default;
  PinBuf(const PinBuf &) = delete;
  PinBuf &operator=(const PinBuf &) = delete;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  T *ensure(size_t n) {
    if (n > cap) {
      if (p) PLADE_CUDA(cudaFreeHost(p));
      p = nullptr;
      size_t want = n + n / 8 + 64;
      PLADE_CUDA(cudaMallocHost(&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
};

// Wait for a stream.  Spinning (cudaStreamSynchronize) has the lowest wake-up latency and is the default; with more host
// threads than cores (several pairs in flight per GPU x several GPUs per box) the waiting threads must yield instead:
// plade_set_param(ctx, "blocking_sync", 1 | 2) or PLADE_BLOCKING_SYNC switches every wait of the process to a blocking
// event (1: cudaEventBlockingSync) or to polling with sched_yield between the polls (2).  Defined in pipeline.cpp.
void stream_sync(cudaStream_t s);
void set_blocking_sync(int mode);      // 0 = spin (cudaStreamSynchronize), 1 = blocking event, 2 = poll + sched_yield

inline int div_up(long long a, long long b) { return (int) ((a + b - 1) / b); }

// Per-stage launch counter (bench.py reports gpu_launches from it).
struct LaunchCounter {
  long long n = 0;
  void add(int k = 1) { n += k; }
};

}  // namespace plade
