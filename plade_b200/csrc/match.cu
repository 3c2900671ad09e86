// K3c — fixed-radius descriptor matching on sm_100a.
//
// Replaces KdTreeSearchNDim<VectorXf,8>::find_neighbors(q, 0, 0.04, ...) as PLADE/util.cpp:163 calls
// it (3rd_party/ann_1.1.2/include/ANN/ANN.h:979-1029): sqRad = float(radius)*float(radius) (a float
// product widened to double); a DB point is a neighbour iff the running double sum of squared
// coordinate differences (DB and query coordinates are floats widened to double, ANN.h:918-924,
// leaf test src/kd_fix_rad_search.cpp:162-177) never exceeds sqRad; the result is returned ascending
// in distance (annkSearch with k = that count, src/kd_search.cpp:89-210).
// The kd-tree is replaced by an exact all-pairs scan: K = 8 and the decision is an fp64 threshold,
// so this is CUDA-core work, not a GEMM.  Ties in distance are ordered by DB index (the tree's tie
// order is its visit order and is not pinned by the reference).
#include "kernels.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cstring>

namespace plade {

namespace {

constexpr int kDim = 8;
constexpr int kTileDb = 256;
constexpr int kThreadsM = 128;

__global__ void __launch_bounds__(kThreadsM)
match_kernel(const float *__restrict__ db, int ndb, const float *__restrict__ q, int nq, double sq_rad,
             int db_chunk, unsigned long long *__restrict__ out_qi, unsigned long long *__restrict__ out_dist,
             unsigned int *__restrict__ counter, unsigned int capacity) {
  __shared__ double tile[kTileDb * kDim];
  const int qi = blockIdx.x * kThreadsM + threadIdx.x;
  double qv[kDim];
#pragma unroll
  for (int d = 0; d < kDim; ++d) qv[d] = (qi < nq) ? (double) q[(size_t) qi * kDim + d] : 0.0;
  const int begin = blockIdx.y * db_chunk, end = min(ndb, begin + db_chunk);
  for (int base = begin; base < end; base += kTileDb) {
    const int cnt = min(kTileDb, end - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * kDim; i += kThreadsM) tile[i] = (double) db[(size_t) base * kDim + i];
    __syncthreads();
    if (qi >= nq) continue;
    for (int j = 0; j < cnt; ++j) {
      const double *p = tile + j * kDim;
      double dist = 0.0;
      int d = 0;
#pragma unroll
      for (; d < kDim; ++d) {
        double t = qv[d] - p[d];
        dist = __dadd_rn(dist, __dmul_rn(t, t));
        if (dist > sq_rad) break;
      }
      if (d >= kDim) {
        unsigned int slot = atomicAdd(counter, 1u);
        if (slot < capacity) {
          out_qi[slot] = ((unsigned long long) (unsigned int) qi << 32) | (unsigned int) (base + j);
          out_dist[slot] = (unsigned long long) __double_as_longlong(dist);   // dist >= 0: bit order == numeric order
        }
      }
    }
  }
}

__global__ void offsets_kernel(const unsigned long long *__restrict__ qi, int m, int nq, int *__restrict__ offsets) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  int cur = (int) (qi[i] >> 32);
  int prev = (i == 0) ? -1 : (int) (qi[i - 1] >> 32);
  for (int c = prev + 1; c <= cur; ++c) offsets[c] = i;
  if (i == m - 1)
    for (int c = cur + 1; c <= nq; ++c) offsets[c] = m;
}

}  // namespace

size_t match_descriptors(Device &dev, MatchScratch &sc, const float *h_db, int ndb, const float *h_q, int nq,
                         float radius, std::vector<int> &offsets, std::vector<int> &idx, std::vector<double> &dist2) {
  offsets.assign((size_t) nq + 1, 0);
  idx.clear();
  dist2.clear();
  if (nq == 0 || ndb == 0 || radius < 0) return 0;
  cudaStream_t s = dev.stream;
  float *d_db = sc.db.ensure((size_t) ndb * kDim), *d_q = sc.q.ensure((size_t) nq * kDim);
  PLADE_CUDA(cudaMemcpyAsync(d_db, h_db, sizeof(float) * (size_t) ndb * kDim, cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemcpyAsync(d_q, h_q, sizeof(float) * (size_t) nq * kDim, cudaMemcpyHostToDevice, s));
  const float sq_f = radius * radius;           // ANN.h:987  float sqRad = radius*radius
  const double sq_rad = (double) sq_f;
  unsigned int *d_counter = reinterpret_cast<unsigned int *>(sc.counts.ensure(4));
  size_t capacity = std::max<size_t>(1u << 20, sc.out_key.cap);
  int qblocks = div_up(nq, kThreadsM);
  int nsplit = std::max(1, std::min(div_up(ndb, kTileDb), (dev.num_sms * 8 + qblocks - 1) / qblocks));
  int chunk = div_up(ndb, nsplit);
  chunk = ((chunk + kTileDb - 1) / kTileDb) * kTileDb;
  nsplit = div_up(ndb, chunk);
  unsigned int m = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    unsigned long long *qi = sc.out_key.ensure(capacity), *di = sc.out_key_alt.ensure(capacity);
    PLADE_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), s));
    match_kernel<<<dim3(qblocks, nsplit), kThreadsM, 0, s>>>(d_db, ndb, d_q, nq, sq_rad, chunk, qi, di, d_counter,
                                                           (unsigned int) capacity);
    PLADE_LAUNCH_CHECK();
    dev.launches.add();
    PLADE_CUDA(cudaMemcpyAsync(&m, d_counter, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    if (m <= capacity) break;
    capacity = (size_t) m + 1024;   // exact size known now: rerun once
  }
  if (m == 0) return 0;
  // canonical order (query asc, dist asc, db index asc): three stable LSD radix passes
  unsigned long long *qi = sc.out_key.p, *di = sc.out_key_alt.p;
  unsigned long long *qi2 = sc.sort_a.ensure(capacity), *di2 = sc.sort_b.ensure(capacity);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, qi, qi2, di, di2, (int) m, 0, 64, s);
  unsigned char *tmp = sc.cub_tmp.ensure(tb);
  cub::DeviceRadixSort::SortPairs(tmp, tb, qi, qi2, di, di2, (int) m, 0, 32, s);     // by db index
  cub::DeviceRadixSort::SortPairs(tmp, tb, di2, di, qi2, qi, (int) m, 0, 64, s);     // by distance
  cub::DeviceRadixSort::SortPairs(tmp, tb, qi, qi2, di, di2, (int) m, 32, 64, s);    // by query
  dev.launches.add(12);
  int *d_off = sc.offsets.ensure((size_t) nq + 1);
  offsets_kernel<<<div_up((int) m, 256), 256, 0, s>>>(qi2, (int) m, nq, d_off);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  std::vector<unsigned long long> h_qi(m), h_di(m);
  PLADE_CUDA(cudaMemcpyAsync(h_qi.data(), qi2, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost, s));
  PLADE_CUDA(cudaMemcpyAsync(h_di.data(), di2, sizeof(unsigned long long) * m, cudaMemcpyDeviceToHost, s));
  PLADE_CUDA(cudaMemcpyAsync(offsets.data(), d_off, sizeof(int) * ((size_t) nq + 1), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  idx.resize(m);
  dist2.resize(m);
  for (unsigned int i = 0; i < m; ++i) {
    idx[i] = (int) (h_qi[i] & 0xffffffffu);
    memcpy(&dist2[i], &h_di[i], 8);
  }
  return m;
}

}  // namespace plade
