// K1 — efficient-RANSAC plane extraction on sm_100a.
//
// Replaces PlaneExtraction::detect (PLADE/plane_extraction.cpp:61-200) -> RansacShapeDetector::Detect
// (3rd_party/ransac/RansacShapeDetector.cpp:455-907) for the only primitive PLADE registers (planes).
// What is kept from the reference, statement for statement:
//   * thresholds: eps = dist_thresh * scale, bitmapEps = bitmap_reso * scale, scale = max(dx, dy) of the
//     bounding box (the z extent is lost by the bug at plane_extraction.cpp:71-80, R/PointCloud.h:94-98);
//   * 3-point plane construction and rejection |(p2-p1)x(p3-p2)|^2 < 1e-6 (R/Plane.cpp:29-38), sample
//     verification against the normal threshold (RansacShapeDetector.cpp:143-153);
//   * the per-point compatibility predicate (R/FlatNormalThreshPointCompatibilityFunc.h:15-22 with
//     Plane::Distance R/Plane.h:31), bit-exact: |dist - n.p| < eps && |n.n_i| >= normalThresh, float,
//     dot products left to right, no FMA;
//   * acceptance of a candidate: global score at 3*eps over all unassigned points (R/Candidate.h:284-292),
//     largest 8-connected component of the inlier bitmap at bitmapEps after a cross closing
//     (R/BitmapPrimitiveShape.cpp:97-205, R/Bitmap.cpp:154,459,633), then up to three least-squares
//     refits accepted while the gaussian-weighted score improves (RansacShapeDetector.cpp:619-655,
//     R/ScoreComputer.h:10-13, R/GfxTL/Plane.h:58-95 with the Jacobi eigen-solver of R/GfxTL/Jacobi.h);
//   * the stopping rule P(overlooking a min_support plane) <= overlook_prob with
//     P = (1 - size / (n * levels * 4))^drawn (R/RansacShapeDetector.h:61-67) and the drawn-candidate
//     rescaling after each accepted shape (RansacShapeDetector.cpp:673-674).
// What is B200-native instead of translated: the octree, the nested random subsets and the lazy
// per-candidate bound refinement (R/Candidate.h:155-213) exist to save CPU work one candidate at a time.
// Here thousands of candidates are drawn per round from windows of a Morton-ordered list of the
// unassigned points (the analogue of "three points from one octree cell of a random level") and ALL of
// them are scored at once against a stratified subsample whose tiles are staged through shared memory
// by 1-D TMA bulk copies; the best one is then verified on the full cloud.  Candidate draws come from a
// counter-based generator with a fixed seed (the reference seeds rand() from time()), so detection is
// reproducible; parity with the reference is at plane level (normal, offset, support), not bit level.
#include "pipeline.h"
#include "planefit.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <mutex>

namespace plade {

namespace {

constexpr int kCandPerRound = 16384;   // candidates drawn per round
constexpr int kGenTries = 1;           // draws per candidate slot (see gen_candidates_kernel: more tries were measured and rejected)
constexpr int kStage1Points = 4096;    // stage 1: every candidate against a small stratified subsample
constexpr int kStage2Cand = 256;       // stage 2: the best stage-1 candidates ...
constexpr int kSubsample = 65536;      // ... against the large subsample
constexpr int kBandBlock = 4096;      // points per band block (see RefineArgs)
constexpr int kScoreTile = 512;       // points per TMA tile (pos + nrm = 16 KB)
constexpr int kScoreThreads = 256;    // threads per scoring block
constexpr int kScoreC1 = 2;           // candidates per thread in stage 1 (16384 candidates x 4096 points)
constexpr unsigned int kForcedKey = 2u * kStage1Points;   // stage-1 key of a carried candidate: above every real count
constexpr int kStage1KeyBits = 14;    // bits of the stage-1 keys (counts <= 4096 < kForcedKey = 8192 < 2^14)
static_assert(kForcedKey < (1u << kStage1KeyBits) && kStage1Points < (int) kForcedKey, "stage-1 sort key range");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(phase)
                 : "memory");
  } while (!ok);
}

// counter-based RNG (splitmix64 finaliser)
__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// The reference predicate.  plane = (nx, ny, nz, dist) with n.x = dist.
__device__ __forceinline__ bool compatible(const float4 &pl, const float4 &p, const float4 &nr, float eps, float nthresh) {
  float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p.x), __fmul_rn(pl.y, p.y)), __fmul_rn(pl.z, p.z));
  float dist = fabsf(__fsub_rn(pl.w, dp));
  if (!(dist < eps)) return false;
  float dn = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, nr.x), __fmul_rn(pl.y, nr.y)), __fmul_rn(pl.z, nr.z));
  return fabsf(dn) >= nthresh;
}

// compatible() without the early out (the same float operations, so the same answer): a short loop body with one candidate
// per thread keeps several points in flight only if it has no branch
__device__ __forceinline__ unsigned int compatible_bit(const float4 &pl, const float4 &p, const float4 &nr, float eps, float nthresh) {
  const float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p.x), __fmul_rn(pl.y, p.y)), __fmul_rn(pl.z, p.z));
  const float dn = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, nr.x), __fmul_rn(pl.y, nr.y)), __fmul_rn(pl.z, nr.z));
  return (unsigned int) (fabsf(__fsub_rn(pl.w, dp)) < eps) & (unsigned int) (fabsf(dn) >= nthresh);
}

// ---- Morton order -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int expand10(unsigned int v) {
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000FF;
  v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}
__global__ void morton_kernel(const float4 *__restrict__ pos, int n, float3 mn, float inv_ext, unsigned int *__restrict__ keys,
                              int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pos[i];
  unsigned int x = (unsigned int) fminf(fmaxf((p.x - mn.x) * inv_ext * 1024.f, 0.f), 1023.f);
  unsigned int y = (unsigned int) fminf(fmaxf((p.y - mn.y) * inv_ext * 1024.f, 0.f), 1023.f);
  unsigned int z = (unsigned int) fminf(fmaxf((p.z - mn.z) * inv_ext * 1024.f, 0.f), 1023.f);
  keys[i] = (expand10(x) << 2) | (expand10(y) << 1) | expand10(z);
  idx[i] = i;
}

__global__ void bbox_kernel(const float4 *__restrict__ p, int n, int *__restrict__ out6) {
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 v = p[i];
    mn[0] = fminf(mn[0], v.x); mn[1] = fminf(mn[1], v.y); mn[2] = fminf(mn[2], v.z);
    mx[0] = fmaxf(mx[0], v.x); mx[1] = fmaxf(mx[1], v.y); mx[2] = fmaxf(mx[2], v.z);
  }
  typedef cub::BlockReduce<float, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int k = 0; k < 3; ++k) {
    float a = BR(tmp).Reduce(mn[k], cub::Min());
    __syncthreads();
    float b = BR(tmp).Reduce(mx[k], cub::Max());
    __syncthreads();
    if (threadIdx.x == 0) {
      int ai = __float_as_int(a), bi = __float_as_int(b);
      atomicMin(out6 + k, ai >= 0 ? ai : ai ^ 0x7fffffff);
      atomicMax(out6 + 3 + k, bi >= 0 ? bi : bi ^ 0x7fffffff);
    }
  }
}

// ---- candidate generation --------------------------------------------------------------------------------
__global__ void gen_candidates_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm,
                                      const int *__restrict__ order, int m, int nlevels, float nthresh,
                                      unsigned long long seed, float4 *__restrict__ cand, int *__restrict__ n_valid,
                                      unsigned int *__restrict__ valid_bits /* kCandPerRound / 32 words: bit c % 32 of word c / 32 = slot c holds a plane */) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  float4 out = make_float4(0.f, 0.f, 0.f, 3.0e38f);
  if (c < kCandPerRound && m >= 3) {
    unsigned long long s = mix64(seed * 0x100000001B3ull + (unsigned long long) c);
    // Most raw draws are useless: triangles of the small windows are below Plane::Init's area threshold, those of the large
    // ones span several surfaces (measured on the 2 M-point scene: 12 % of the draws pass).  Drawing again until the slot holds
    // a verified plane (kGenTries = 16) was measured: the rounds do not get fewer (the stopping rule counts verified
    // candidates and is rescaled after every accepted shape), the per-pair time went UP 12 % and the end-to-end success
    // rates of the sample pairs went down (the retried draws favour mid-size windows).  One draw per slot it stays.
    for (int attempt = 0; attempt < kGenTries && !ok; ++attempt) {
    int r0 = (int) (s % (unsigned long long) m);
    s = mix64(s);
    int level = (int) (s % (unsigned long long) nlevels);
    long long w = 4ll << level;                       // half window in Morton order
    s = mix64(s);
    long long span = 2 * w + 1;
    long long r1 = r0 - w + (long long) (s % (unsigned long long) span);
    s = mix64(s);
    long long r2 = r0 - w + (long long) (s % (unsigned long long) span);
    s = mix64(s);
    r1 = min(max(r1, 0ll), (long long) m - 1);
    r2 = min(max(r2, 0ll), (long long) m - 1);
    if (r1 != r0 && r2 != r0 && r1 != r2) {
      int i0 = order[r0], i1 = order[(int) r1], i2 = order[(int) r2];
      float4 p1 = pos[i0], p2 = pos[i1], p3 = pos[i2];
      // Plane::Init(p1, p2, p3): normal = (p2 - p1) x (p3 - p2)
      float ax = __fsub_rn(p2.x, p1.x), ay = __fsub_rn(p2.y, p1.y), az = __fsub_rn(p2.z, p1.z);
      float bx = __fsub_rn(p3.x, p2.x), by = __fsub_rn(p3.y, p2.y), bz = __fsub_rn(p3.z, p2.z);
      float nx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
      float ny = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
      float nz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
      float sq = __fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz));
      if (sq >= 1e-6f) {
        float l = sqrtf(sq);
        nx = __fdiv_rn(nx, l); ny = __fdiv_rn(ny, l); nz = __fdiv_rn(nz, l);
        float dist = __fadd_rn(__fadd_rn(__fmul_rn(p1.x, nx), __fmul_rn(p1.y, ny)), __fmul_rn(p1.z, nz));
        float4 pl = make_float4(nx, ny, nz, dist);
        // sample verification: normals of all three samples must agree with the plane
        float4 n1 = nrm[i0], n2 = nrm[i1], n3 = nrm[i2];
        float d1 = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(nx, n1.x), __fmul_rn(ny, n1.y)), __fmul_rn(nz, n1.z)));
        float d2 = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(nx, n2.x), __fmul_rn(ny, n2.y)), __fmul_rn(nz, n2.z)));
        float d3 = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(nx, n3.x), __fmul_rn(ny, n3.y)), __fmul_rn(nz, n3.z)));
        if (d1 >= nthresh && d2 >= nthresh && d3 >= nthresh) { out = pl; ok = true; }
      }
    }
    }
  }
  if (c < kCandPerRound) cand[c] = out;
  unsigned int mk = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) {
    if (mk) atomicAdd(n_valid, __popc(mk));
    if (c < kCandPerRound) valid_bits[c >> 5] = mk;      // (the grid is exactly kCandPerRound threads: every word is written)
  }
}

// Slots of the round's candidate array that hold a plane -- the first n_forced (carried pool entries, copied over the draws
// of those slots) and the draws that passed Plane::Init and the sample verification (gen_candidates_kernel's valid_bits;
// measured: 12 % of the draws) -- in ascending slot order.  One block of 256 threads, two 32-slot words per thread.
__device__ void compact_live_slots(const unsigned int *__restrict__ valid_bits, int n_forced, int *__restrict__ live, int *__restrict__ n_live) {
  static_assert(kCandPerRound == 256 * 64, "two 32-slot words per thread");
  typedef cub::BlockScan<int, 256> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  const int tid = threadIdx.x;
  unsigned int w[2];
  for (int h = 0; h < 2; ++h) {
    const int first = 64 * tid + 32 * h;                   // slot of bit 0
    const int f = min(max(n_forced - first, 0), 32);       // forced slots in this word
    w[h] = valid_bits[2 * tid + h] | (f >= 32 ? 0xffffffffu : ((1u << f) - 1u));
  }
  int off, total;
  Scan(scan_tmp).ExclusiveSum(__popc(w[0]) + __popc(w[1]), off, total);
  for (int h = 0; h < 2; ++h)
    while (w[h]) { const int b = __ffs(w[h]) - 1; w[h] &= w[h] - 1; live[off++] = 64 * tid + 32 * h + b; }
  if (tid == 0) *n_live = total;
}

// (both subsamples of a round in one launch: threads [0, Sa) build `sub_a` with seed_a, threads [Sa, Sa + Sb) build `sub_b`;
//  with live != null the launch carries one more block, the last, which lists the live candidate slots)
__global__ void __launch_bounds__(256)
gather_sub_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, const int *__restrict__ order,
                  int m, int Sa, unsigned long long seed_a, float4 *__restrict__ sub_a, int Sb, unsigned long long seed_b,
                  float4 *__restrict__ sub_b /* [2*S]: pos | nrm interleaved per tile */,
                  const unsigned int *__restrict__ valid_bits, int n_forced, int *__restrict__ live, int *__restrict__ n_live) {
  if (live != nullptr && blockIdx.x == gridDim.x - 1) { compact_live_slots(valid_bits, n_forced, live, n_live); return; }
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Sa + Sb) return;
  const bool first = k < Sa;
  const int S = first ? Sa : Sb;
  const unsigned long long seed = first ? seed_a : seed_b;
  float4 *sub = first ? sub_a : sub_b;
  if (!first) k -= Sa;
  // stratified: one point from each of S equal strata of the Morton-ordered list
  long long b = (long long) k * m / S, e = (long long) (k + 1) * m / S;
  if (e <= b) e = b + 1;
  unsigned long long s = mix64(seed ^ (0xABCDEFull + (unsigned long long) k));
  int r = (int) (b + (long long) (s % (unsigned long long) (e - b)));
  int i = order[min(r, m - 1)];
  int tile = k / kScoreTile, j = k - tile * kScoreTile;
  sub[(size_t) tile * 2 * kScoreTile + j] = pos[i];
  sub[(size_t) tile * 2 * kScoreTile + kScoreTile + j] = nrm[i];
}

// K1a: every candidate of the round against the subsample.  grid = (tile groups, candidate groups).
// C candidates per thread: the two shared-memory loads of a point (broadcast to the warp) are amortised over C
// plane tests, and the C independent dependency chains keep the FP pipe busy with few warps per SM.
// sel != null: only the slots sel[0 .. n_cand) are scored -- stage 1: the live slots (compact_live_slots; *n_sel of them,
// blocks beyond that leave at once), counts in counts[sel[c]], i.e. where scoring every slot would have put them;
// stage 2 (by_rank): the kStage2Cand slots select_top_kernel picked, counts in counts[c] (c = the rank).
template <int C>
__global__ void __launch_bounds__(kScoreThreads)
score_candidates_kernel(const float4 *__restrict__ sub, int S, const float4 *__restrict__ cand, const int *__restrict__ sel,
                        const int *__restrict__ n_sel, int by_rank, int n_cand, float eps, float nthresh, int tiles_per_block,
                        unsigned int *__restrict__ counts) {
  __shared__ __align__(128) float4 buf[2][2 * kScoreTile];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  if (n_sel) n_cand = min(n_cand, *n_sel);
  if ((int) blockIdx.y * (kScoreThreads * C) >= n_cand) return;      // uniform over the block
  const int c0 = blockIdx.y * (kScoreThreads * C) + tid;          // this thread's candidates: c0 + k * kScoreThreads
  float4 pl[C];
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const int c = c0 + k * kScoreThreads;
    pl[k] = (c < n_cand) ? cand[sel ? sel[c] : c] : make_float4(0.f, 0.f, 0.f, 3.0e38f);
  }
  const int n_tiles = (S + kScoreTile - 1) / kScoreTile;
  const int t0 = blockIdx.x * tiles_per_block, t1 = min(n_tiles, t0 + tiles_per_block);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t, int b) {
    int np = min(kScoreTile, S - t * kScoreTile);
    // a partial last tile still carries pos at [0, np) and nrm at [kScoreTile, kScoreTile + np)
    uint32_t bytes = (np == kScoreTile) ? 2u * kScoreTile * 16u : (uint32_t) (kScoreTile + np) * 16u;
    mbar_expect_tx(&bar[b], bytes);
    tma_load_1d(buf[b], sub + (size_t) t * 2 * kScoreTile, bytes, &bar[b]);
  };
  uint32_t phase[2] = {0, 0};
  if (tid == 0 && t0 < t1) issue(t0, 0);
  unsigned int cnt[C];
#pragma unroll
  for (int k = 0; k < C; ++k) cnt[k] = 0;
  for (int t = t0; t < t1; ++t) {
    const int b = (t - t0) & 1;
    if (tid == 0 && t + 1 < t1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(t + 1, b ^ 1);
    }
    mbar_wait(&bar[b], phase[b]);
    phase[b] ^= 1;
    const int np = min(kScoreTile, S - t * kScoreTile);
    const float4 *P = buf[b], *N = buf[b] + kScoreTile;
    if constexpr (C == 1) {
#pragma unroll 8
      for (int j = 0; j < np; ++j) cnt[0] += compatible_bit(pl[0], P[j], N[j], eps, nthresh);
    } else {
#pragma unroll 2
      for (int j = 0; j < np; ++j) {
        const float4 p = P[j], nr = N[j];
#pragma unroll
        for (int k = 0; k < C; ++k) cnt[k] += compatible(pl[k], p, nr, eps, nthresh) ? 1u : 0u;
      }
    }
    __syncthreads();   // everyone done with buf[b] before it is refilled two iterations later
  }
#pragma unroll
  for (int k = 0; k < C; ++k) {
    const int c = c0 + k * kScoreThreads;
    if (c < n_cand && cnt[k]) atomicAdd(&counts[(sel && !by_rank) ? sel[c] : c], cnt[k]);
  }
}

// Stage 2 (kStage2Cand candidates x 65536 points) has too few candidates for one-candidate-per-thread to fill the
// GPU, so the roles are swapped: one thread per subsample point (65536 threads), the candidate planes broadcast from
// shared memory, and the per-candidate counts formed from warp ballots.  Same predicate, same integer counts.
__global__ void __launch_bounds__(256)
score_points_kernel(const float4 *__restrict__ sub, int S, const float4 *__restrict__ cand, const int *__restrict__ sel,
                    float eps, float nthresh, unsigned int *__restrict__ counts) {
  __shared__ float4 pl_s[kStage2Cand];
  __shared__ unsigned int cnt_s[kStage2Cand];
  const int tid = threadIdx.x, lane = tid & 31;
  for (int c = tid; c < kStage2Cand; c += blockDim.x) { pl_s[c] = cand[sel[c]]; cnt_s[c] = 0u; }
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + tid;
  const bool live = k < S;
  const int kk = live ? k : S - 1;
  const int tile = kk / kScoreTile, j = kk - tile * kScoreTile;
  const float4 p = sub[(size_t) tile * 2 * kScoreTile + j], nr = sub[(size_t) tile * 2 * kScoreTile + kScoreTile + j];
#pragma unroll 4
  for (int c = 0; c < kStage2Cand; ++c) {
    const float4 pl = pl_s[c];
    const float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p.x), __fmul_rn(pl.y, p.y)), __fmul_rn(pl.z, p.z));
    const bool near = live && fabsf(__fsub_rn(pl.w, dp)) < eps;
    if (__any_sync(0xffffffffu, near)) {            // warp-uniform branch
      const float dn = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, nr.x), __fmul_rn(pl.y, nr.y)), __fmul_rn(pl.z, nr.z));
      const unsigned int m = __ballot_sync(0xffffffffu, near && fabsf(dn) >= nthresh);
      if (lane == 0 && m) atomicAdd(&cnt_s[c], (unsigned int) __popc(m));
    }
  }
  __syncthreads();
  for (int c = tid; c < kStage2Cand; c += blockDim.x) if (cnt_s[c]) atomicAdd(&counts[c], cnt_s[c]);
}

// Stage-1 selection: the kStage2Cand candidates a stable descending sort of the stage-1 counts would put first
// (carried pool candidates forced to the front, ties by ascending candidate index), in that order, without
// sorting all kCandPerRound of them: one block builds the histogram of the counts (<= kStage1Points), finds the
// threshold count by a suffix scan, compacts the candidates above it plus the lowest-index ties, and bitonic-sorts
// the kStage2Cand survivors.  Replaces four library sort kernels per round by one 1-SM kernel.
constexpr int kSelThreads = 1024;
// layout of RansacScratch::round_buf in 32-bit words: one memset clears [0, kRoundTop), one copy fetches [kRoundCounts2, kRoundEnd)
constexpr int kRoundCounts2 = kCandPerRound, kRoundNValid = kRoundCounts2 + kStage2Cand, kRoundTop = kRoundNValid + 4,
              kRoundEnd = kRoundTop + 4 * kStage2Cand;
static_assert((kRoundTop * 4) % 16 == 0, "selected planes must be float4-aligned");
__global__ void __launch_bounds__(kSelThreads)
select_top_kernel(const unsigned int *__restrict__ counts, int n, int n_forced, int *__restrict__ out /* kStage2Cand */,
                  const float4 *__restrict__ cand, float4 *__restrict__ cand_top /* the selected planes, in order */) {
  constexpr int K = kStage2Cand, kBins = kStage1Points + 1, kPerThread = (kBins + kSelThreads - 1) / kSelThreads;
  __shared__ unsigned int hist[kSelThreads * kPerThread];
  __shared__ unsigned long long keys[K];
  __shared__ int s_thr, s_above, s_fill;
  typedef cub::BlockScan<int, kSelThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  const int tid = threadIdx.x;
  n_forced = min(n_forced, K);
  const int need = K - n_forced;                         // slots left for the ranked candidates
  for (int b = tid; b < kSelThreads * kPerThread; b += kSelThreads) hist[b] = 0u;
  if (tid < K) keys[tid] = 0ull;
  if (tid == 0) { s_thr = -1; s_above = 0; s_fill = 0; }
  __syncthreads();
  for (int i = n_forced + tid; i < n; i += kSelThreads) atomicAdd(&hist[min(counts[i], (unsigned int) kStage1Points)], 1u);
  __syncthreads();
  // suffix scan over the bins, highest count first: thread t owns bins top - t*kPerThread - j
  {
    const int top = kSelThreads * kPerThread - 1;
    int mine = 0;
    for (int j = 0; j < kPerThread; ++j) mine += (int) hist[top - (tid * kPerThread + j)];
    int incl;
    Scan(scan_tmp).InclusiveSum(mine, incl);
    const int excl = incl - mine;                        // candidates in bins above this thread's
    if (excl < need && incl >= need) {                   // the threshold bin is one of mine (exactly one thread)
      int run = excl;
      for (int j = 0; j < kPerThread; ++j) {
        const int b = top - (tid * kPerThread + j), c = (int) hist[b];
        if (run + c >= need) { s_thr = b; s_above = run; break; }
        run += c;
      }
    }
  }
  __syncthreads();
  const int thr = s_thr, above = s_above;                // thr < 0: fewer than `need` candidates in total -> take all
  const int quota = thr < 0 ? 0 : need - above;          // ties at the threshold to take, lowest index first
  // ordered compaction: every thread walks a contiguous index range
  const int per = (n - n_forced + kSelThreads - 1) / kSelThreads;
  const int i0 = n_forced + tid * per, i1 = min(n, i0 + per);
  int ties = 0;
  for (int i = i0; i < i1; ++i) ties += (thr >= 0 && (int) min(counts[i], (unsigned int) kStage1Points) == thr) ? 1 : 0;
  int tie_rank;
  __syncthreads();
  Scan(scan_tmp).ExclusiveSum(ties, tie_rank);
  for (int i = i0; i < i1; ++i) {
    const int c = (int) min(counts[i], (unsigned int) kStage1Points);
    bool take = thr < 0 || c > thr;
    if (thr >= 0 && c == thr) { take = tie_rank < quota; ++tie_rank; }
    if (take) {
      const int slot = atomicAdd(&s_fill, 1);
      if (slot < need) keys[n_forced + slot] = ((unsigned long long) (unsigned int) c << 32) | (unsigned long long) (0xFFFFFFFFu - (unsigned int) i);
    }
  }
  if (tid < n_forced) keys[tid] = ((unsigned long long) kForcedKey << 32) | (unsigned long long) (0xFFFFFFFFu - (unsigned int) tid);
  __syncthreads();
  // bitonic sort, descending: (count desc, index asc) == the order of the stable descending radix sort it replaces
  for (int k = 2; k <= K; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (tid < K) {
        const int ixj = tid ^ j;
        if (ixj > tid) {
          const unsigned long long a = keys[tid], b = keys[ixj];
          const bool desc = (tid & k) == 0;
          if (desc ? a < b : a > b) { keys[tid] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  // unfilled slots (fewer than K candidates) keep key 0 -> index 0xFFFFFFFF: point them at candidate 0 as the sort did not exist for them
  if (tid < K) {
    const unsigned long long kx = keys[tid];
    const int sel = kx ? (int) (0xFFFFFFFFu - (unsigned int) (kx & 0xFFFFFFFFull)) : 0;
    out[tid] = sel;
    cand_top[tid] = cand[sel];
  }
}

// K1a/K1b stage API: n_planes planes against the whole cloud (assigned[i] == -1 only).
__global__ void score_planes_full_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm,
                                         const int *__restrict__ assigned, int n, const float4 *__restrict__ planes,
                                         int n_planes, float eps, float nthresh, unsigned int *__restrict__ counts,
                                         unsigned char *__restrict__ mask0) {
  extern __shared__ float4 spl[];
  for (int i = threadIdx.x; i < n_planes; i += blockDim.x) spl[i] = planes[i];
  __syncthreads();
  for (int p = 0; p < n_planes; ++p) {
    unsigned int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      bool in = (assigned == nullptr || assigned[i] == -1) && compatible(spl[p], pos[i], nrm[i], eps, nthresh);
      if (p == 0 && mask0) mask0[i] = in ? 1 : 0;
      c += in ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&counts[p], c);
  }
}

// ---- refinement kernels --------------------------------------------------------------------------------
struct PlaneFrame {      // plane + in-plane frame (PlanePrimitiveShape: m_plane, m_hcs)
  float4 pl;             // n, dist
  float3 pos, u, v;
};

__device__ __forceinline__ int f2o(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }

// pass A: inlier flags at eps3 + (u, v) bounding box.  `assigned` may be null (all points free) and the point
// count may come from device memory (d_count), as for the compacted band of a candidate (see band_compact_kernel).
__global__ void flag_uv_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, const int *__restrict__ assigned,
                               int n, const int *__restrict__ d_count, PlaneFrame f, float eps3, float nthresh, unsigned char *__restrict__ flag,
                               int *__restrict__ uvbox /* 4 ordered ints: umin vmin umax vmax */) {
  if (d_count) n = *d_count;
  float umin = 3.4e38f, vmin = 3.4e38f, umax = -3.4e38f, vmax = -3.4e38f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = pos[i];
    bool in = (assigned == nullptr || assigned[i] == -1) && compatible(f.pl, p, nrm[i], eps3, nthresh);
    flag[i] = in ? 1 : 0;
    if (in) {
      // PlanePrimitiveShape::ParametersImpl (R/PlanePrimitiveShape.h:97-109)
      float px = __fsub_rn(p.x, f.pos.x), py = __fsub_rn(p.y, f.pos.y), pz = __fsub_rn(p.z, f.pos.z);
      float u = __fadd_rn(__fadd_rn(__fmul_rn(px, f.u.x), __fmul_rn(py, f.u.y)), __fmul_rn(pz, f.u.z));
      float v = __fadd_rn(__fadd_rn(__fmul_rn(px, f.v.x), __fmul_rn(py, f.v.y)), __fmul_rn(pz, f.v.z));
      umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
  }
  typedef cub::BlockReduce<float, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  float a = BR(tmp).Reduce(umin, cub::Min()); __syncthreads();
  float b = BR(tmp).Reduce(vmin, cub::Min()); __syncthreads();
  float c = BR(tmp).Reduce(umax, cub::Max()); __syncthreads();
  float d = BR(tmp).Reduce(vmax, cub::Max());
  if (threadIdx.x == 0) {
    atomicMin(uvbox + 0, f2o(a)); atomicMin(uvbox + 1, f2o(b));
    atomicMax(uvbox + 2, f2o(c)); atomicMax(uvbox + 3, f2o(d));
  }
}

// Band of a candidate: the still-unassigned points within `band` of the candidate's plane (band = +inf: all of
// them), compacted into contiguous copies.  Every evaluation of the candidate and of its least-squares refits
// then touches the band only (a few % of the cloud) instead of the whole cloud; the host proves before each
// evaluation that no point outside the band can be an inlier of the plane being evaluated (see band_covers).
__global__ void band_compact_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ nrm, const int *__restrict__ assigned, int n,
                                    float4 pl, float band, float4 *__restrict__ posB, float4 *__restrict__ nrmB, int *__restrict__ idxB,
                                    int *__restrict__ d_nb /* [segs] */, int seg_cap, int segs) {
  // HBM streaming pass (20 B per point: assigned + pos): four independent loads in flight per thread, the warp's
  // survivors are appended with one atomic per 4 x 32 points.  Output layout: see RefineArgs (block b of kBandBlock
  // points -> segment b % segs; a warp's 128 points never straddle a block).
  constexpr int kU = 4;
  static_assert(kBandBlock % (32 * kU) == 0, "a warp chunk stays inside one band block");
  const unsigned lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (long long base = (long long) warp * (32 * kU); base < n; base += (long long) n_warps * (32 * kU)) {
    int a[kU];
    float4 p[kU];
    bool in[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = base + u * 32 + lane;
      a[u] = i < n ? __ldg(assigned + i) : 0;
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = base + u * 32 + lane;
      p[u] = (i < n && a[u] == -1) ? __ldg(pos + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned bal[kU];
    int total = 0;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const long long i = base + u * 32 + lane;
      float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p[u].x), __fmul_rn(pl.y, p[u].y)), __fmul_rn(pl.z, p[u].z));
      in[u] = i < n && a[u] == -1 && fabsf(__fsub_rn(pl.w, dp)) < band;
      bal[u] = __ballot_sync(0xffffffffu, in[u]);
      total += __popc(bal[u]);
    }
    if (!total) continue;
    const int seg = (int) ((base / kBandBlock) % segs);
    int start = 0;
    if (lane == 0) start = atomicAdd(d_nb + seg, total);
    start = __shfl_sync(0xffffffffu, start, 0) + seg * seg_cap;
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (in[u]) {
        const int i = (int) (base + u * 32 + lane);
        const int k = start + __popc(bal[u] & ((1u << lane) - 1));
        posB[k] = p[u]; nrmB[k] = __ldg(nrm + i); idxB[k] = i;
      }
      start += __popc(bal[u]);
    }
  }
}

// pass C: rasterise (BuildBitmap / InBitmap, R/BitmapPrimitiveShape.h:101-150, R/PlanePrimitiveShape.cpp:200-207)
__global__ void raster_kernel(const float4 *__restrict__ pos, const unsigned char *__restrict__ flag, int n, PlaneFrame f,
                              float umin, float vmin, float bmp_eps, int uext, int vext, int *__restrict__ pix,
                              unsigned char *__restrict__ bitmap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!flag[i]) return;
  float4 p = pos[i];
  float px = __fsub_rn(p.x, f.pos.x), py = __fsub_rn(p.y, f.pos.y), pz = __fsub_rn(p.z, f.pos.z);
  float u = __fadd_rn(__fadd_rn(__fmul_rn(px, f.u.x), __fmul_rn(py, f.u.y)), __fmul_rn(pz, f.u.z));
  float v = __fadd_rn(__fadd_rn(__fmul_rn(px, f.v.x), __fmul_rn(py, f.v.y)), __fmul_rn(pz, f.v.z));
  int bu = (int) floorf(__fdiv_rn(__fsub_rn(u, umin), bmp_eps));
  int bv = (int) floorf(__fdiv_rn(__fsub_rn(v, vmin), bmp_eps));
  bu = min(max(bu, 0), uext - 1);
  bv = min(max(bv, 0), vext - 1);
  int id = bu + bv * uext;
  pix[i] = id;
  bitmap[id] = 1;
}

// ---- device-side bitmap pipeline (no host round trip between the passes of one candidate evaluation) ----
struct BmpInfo { float umin, vmin; int ue, ve; int ok; int overflow; int pad0, pad1; };
constexpr int kBmpCap = 1 << 20;       // pixels; larger bitmaps take the host path
constexpr float kBandMul = 3.f;        // half-width of a candidate's band in units of eps3 (= 3 eps)

__device__ __forceinline__ BmpInfo bmp_info_from_box(const int *uvbox, float bmp_eps) {
  int a[4];
  float uv[4];
  for (int k = 0; k < 4; ++k) { a[k] = uvbox[k]; a[k] = a[k] >= 0 ? a[k] : a[k] ^ 0x7fffffff; uv[k] = __int_as_float(a[k]); }
  BmpInfo r;
  r.umin = uv[0]; r.vmin = uv[1]; r.ue = r.ve = 0; r.ok = 0; r.overflow = 0; r.pad0 = r.pad1 = 0;
  if (uv[0] <= uv[2]) {
    // BitmapExtent (R/PlanePrimitiveShape.cpp:192-198)
    long long ue = (long long) ceilf(__fdiv_rn(__fsub_rn(uv[2], uv[0]), bmp_eps)) + 1, ve = (long long) ceilf(__fdiv_rn(__fsub_rn(uv[3], uv[1]), bmp_eps)) + 1;
    if (ue < 2) ue = 2;
    if (ve < 2) ve = 2;
    if (ue * ve <= kBmpCap) { r.ue = (int) ue; r.ve = (int) ve; r.ok = 1; } else r.overflow = 1;
  }
  return r;
}

// (the bitmap geometry is derived from the (u, v) box by every thread; block 0 publishes it for the later passes)
__global__ void raster_dev_kernel(const float4 *__restrict__ pos, const unsigned char *__restrict__ flag, const int *__restrict__ d_count,
                                  PlaneFrame f, float bmp_eps, const int *__restrict__ uvbox, BmpInfo *__restrict__ info, int *__restrict__ pix,
                                  unsigned char *__restrict__ bitmap) {
  const BmpInfo I = bmp_info_from_box(uvbox, bmp_eps);
  if (blockIdx.x == 0 && threadIdx.x == 0) *info = I;
  if (!I.ok) return;
  const int n = *d_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!flag[i]) continue;
    float4 p = pos[i];
    float px = __fsub_rn(p.x, f.pos.x), py = __fsub_rn(p.y, f.pos.y), pz = __fsub_rn(p.z, f.pos.z);
    float u = __fadd_rn(__fadd_rn(__fmul_rn(px, f.u.x), __fmul_rn(py, f.u.y)), __fmul_rn(pz, f.u.z));
    float v = __fadd_rn(__fadd_rn(__fmul_rn(px, f.v.x), __fmul_rn(py, f.v.y)), __fmul_rn(pz, f.v.z));
    int bu = (int) floorf(__fdiv_rn(__fsub_rn(u, I.umin), bmp_eps));
    int bv = (int) floorf(__fdiv_rn(__fsub_rn(v, I.vmin), bmp_eps));
    bu = min(max(bu, 0), I.ue - 1);
    bv = min(max(bv, 0), I.ve - 1);
    int id = bu + bv * I.ue;
    pix[i] = id;
    bitmap[id] = 1;
  }
}

// One block: cross closing (dilate, erode; R/Bitmap.cpp:154,459), 8-connected labelling by min-label
// propagation with pointer jumping (a union-find with L2 atomics was tried and is 3x slower on the nearly full
// bitmaps of real planes), largest component by PIXEL count (first in raster order on ties,
// R/BitmapPrimitiveShape.cpp:170-173) -> mask.  Leaves the bitmap cleared for the next evaluation.
__device__ void cc_block(unsigned char *bitmap, unsigned char *tmp, int *lab, int *cnt, const BmpInfo I, unsigned char *mask) {
  __shared__ int changed;
  __shared__ unsigned long long best;
  if (!I.ok) return;
  const int ue = I.ue, ve = I.ve, P = ue * ve, tid = threadIdx.x, nt = blockDim.x;
  for (int p = tid; p < P; p += nt) {
    int x = p % ue, y = p / ue;
    tmp[p] = bitmap[p] | (x > 0 ? bitmap[p - 1] : 0) | (x + 1 < ue ? bitmap[p + 1] : 0) | (y > 0 ? bitmap[p - ue] : 0) | (y + 1 < ve ? bitmap[p + ue] : 0);
  }
  __syncthreads();
  for (int p = tid; p < P; p += nt) {
    int x = p % ue, y = p / ue;
    unsigned char e = tmp[p] & (x > 0 ? tmp[p - 1] : 1) & (x + 1 < ue ? tmp[p + 1] : 1) & (y > 0 ? tmp[p - ue] : 1) & (y + 1 < ve ? tmp[p + ue] : 1);
    lab[p] = e ? p : -1;
    cnt[p] = 0;
  }
  do {
    __syncthreads();
    if (tid == 0) changed = 0;
    __syncthreads();
    for (int p = tid; p < P; p += nt) {
      int l = lab[p];
      if (l < 0) continue;
      int x = p % ue, y = p / ue, m = l;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          int xx = x + dx, yy = y + dy;
          if (xx < 0 || yy < 0 || xx >= ue || yy >= ve) continue;
          int lq = lab[yy * ue + xx];
          if (lq >= 0 && lq < m) m = lq;
        }
      if (m < l) { lab[p] = m; changed = 1; }
    }
    __syncthreads();
    for (int p = tid; p < P; p += nt) {
      int l = lab[p];
      if (l < 0) continue;
      int r = l;
      for (int it = 0; it < 64; ++it) { int nr = lab[r]; if (nr < 0 || nr >= r) break; r = nr; }
      if (r < l) { lab[p] = r; changed = 1; }
    }
    __syncthreads();
  } while (changed);
  for (int p = tid; p < P; p += nt) { int l = lab[p]; if (l >= 0) atomicAdd(&cnt[l], 1); }
  if (tid == 0) best = 0ull;
  __syncthreads();
  for (int p = tid; p < P; p += nt)
    if (lab[p] == p) atomicMax(&best, ((unsigned long long) (unsigned int) cnt[p] << 32) | (unsigned long long) (0xFFFFFFFFu - (unsigned int) p));
  __syncthreads();
  const int bl = best ? (int) (0xFFFFFFFFu - (unsigned int) (best & 0xFFFFFFFFull)) : -2;
  for (int p = tid; p < P; p += nt) { mask[p] = (lab[p] == bl) ? 1 : 0; bitmap[p] = 0; }
}
__global__ void __launch_bounds__(1024)
cc_kernel(unsigned char *bitmap, unsigned char *tmp, int *lab, int *cnt, const BmpInfo *__restrict__ info, unsigned char *mask) {
  cc_block(bitmap, tmp, lab, cnt, *info, mask);
}

__global__ void cov_dev_kernel(const float4 *__restrict__ pos, const int *__restrict__ idx, const unsigned char *__restrict__ member,
                               const int *__restrict__ d_count, const double *__restrict__ acc, double *__restrict__ acc6) {
  const int n = *d_count;
  // float mean of the members (GfxTL::Mean) from the sums of the select pass: acc = {count, sx, sy, sz}
  const double cnt = acc[0];
  if (!(cnt > 0)) return;
  const float mx = (float) (acc[1] / cnt), my = (float) (acc[2] / cnt), mz = (float) (acc[3] / cnt);
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!member[idx[i]]) continue;
    float4 p = pos[i];
    double dx = (double) (p.x - mx), dy = (double) (p.y - my), dz = (double) (p.z - mz);
    c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int k = 0; k < 6; ++k) {
    double r = BR(tmp).Sum(c[k]);
    __syncthreads();
    if (threadIdx.x == 0 && r != 0) atomicAdd(acc6 + k, r);
  }
}

// select with the mask produced on the device (no-op when the bitmap stage was skipped)
__global__ void select_dev_kernel(const float4 *__restrict__ pos, const int *__restrict__ idx, const unsigned char *__restrict__ flag,
                                  const int *__restrict__ pix, const unsigned char *__restrict__ comp_mask, const BmpInfo *__restrict__ info,
                                  const int *__restrict__ d_count, float4 pl, float eps3, unsigned char *__restrict__ member /* by point index */,
                                  double *__restrict__ acc) {
  const int ok = info->ok, n = *d_count;
  double cnt = 0, sx = 0, sy = 0, sz = 0, sc = 0;
  const float denom = 2.f / 9.f * eps3 * eps3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    bool mem = ok && flag[i] && comp_mask[pix[i]];
    member[idx[i]] = mem ? 1 : 0;
    if (mem) {
      float4 p = pos[i];
      float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p.x), __fmul_rn(pl.y, p.y)), __fmul_rn(pl.z, p.z));
      float d = fabsf(__fsub_rn(pl.w, dp));
      cnt += 1; sx += p.x; sy += p.y; sz += p.z;
      sc += (double) expf(-d * d / denom);
    }
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  double r0 = BR(tmp).Sum(cnt); __syncthreads();
  double r1 = BR(tmp).Sum(sx); __syncthreads();
  double r2 = BR(tmp).Sum(sy); __syncthreads();
  double r3 = BR(tmp).Sum(sz); __syncthreads();
  double r4 = BR(tmp).Sum(sc);
  if (threadIdx.x == 0 && r0 > 0) {
    atomicAdd(acc + 0, r0); atomicAdd(acc + 1, r1); atomicAdd(acc + 2, r2); atomicAdd(acc + 3, r3); atomicAdd(acc + 4, r4);
  }
}

// end of a candidate: accept (assign the members of `member` to shape_id) and/or clear both membership maps
// over the band, restoring the all-zero invariant they have between candidates
__global__ void band_finish_kernel(const int *__restrict__ idx, const int *__restrict__ d_count, unsigned char *__restrict__ member,
                                   unsigned char *__restrict__ other, int shape_id, int *__restrict__ assigned) {
  const int n = *d_count;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int g = idx[i];
    if (shape_id >= 0 && member[g]) assigned[g] = shape_id;
    member[g] = 0;
    other[g] = 0;
  }
}

// accept after refine_cluster_kernel: its membership map is indexed by band position (segmented layout, see RefineArgs)
__global__ void band_assign_kernel(const int *__restrict__ idx, const int *__restrict__ d_nb, int seg_cap, int segs,
                                   const unsigned char *__restrict__ member_band, int shape_id, int *__restrict__ assigned) {
  for (int sg = 0; sg < segs; ++sg) {
    const int n = d_nb[sg], sb = sg * seg_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      if (member_band[sb + i]) assigned[idx[sb + i]] = shape_id;
  }
}

// pass E: members of the largest component; count, sum of positions, gaussian-weighted score
__global__ void select_kernel(const float4 *__restrict__ pos, const unsigned char *__restrict__ flag, const int *__restrict__ pix,
                              const unsigned char *__restrict__ comp_mask, int n, float4 pl, float eps3,
                              unsigned char *__restrict__ member, double *__restrict__ acc /* cnt, sx, sy, sz, score */) {
  double cnt = 0, sx = 0, sy = 0, sz = 0, sc = 0;
  const float denom = 2.f / 9.f * eps3 * eps3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    bool mem = flag[i] && comp_mask[pix[i]];
    member[i] = mem ? 1 : 0;
    if (mem) {
      float4 p = pos[i];
      float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p.x), __fmul_rn(pl.y, p.y)), __fmul_rn(pl.z, p.z));
      float d = fabsf(__fsub_rn(pl.w, dp));
      cnt += 1; sx += p.x; sy += p.y; sz += p.z;
      sc += (double) expf(-d * d / denom);
    }
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  double r0 = BR(tmp).Sum(cnt); __syncthreads();
  double r1 = BR(tmp).Sum(sx); __syncthreads();
  double r2 = BR(tmp).Sum(sy); __syncthreads();
  double r3 = BR(tmp).Sum(sz); __syncthreads();
  double r4 = BR(tmp).Sum(sc);
  if (threadIdx.x == 0 && r0 > 0) {
    atomicAdd(acc + 0, r0); atomicAdd(acc + 1, r1); atomicAdd(acc + 2, r2); atomicAdd(acc + 3, r3); atomicAdd(acc + 4, r4);
  }
}

// pass G: covariance about `mean` over the members (K1d)
__global__ void cov_kernel(const float4 *__restrict__ pos, const unsigned char *__restrict__ member, int n, float3 mean,
                           double *__restrict__ acc6) {
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (!member[i]) continue;
    float4 p = pos[i];
    double dx = (double) (p.x - mean.x), dy = (double) (p.y - mean.y), dz = (double) (p.z - mean.z);
    c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
  }
  typedef cub::BlockReduce<double, 256> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int k = 0; k < 6; ++k) {
    double r = BR(tmp).Sum(c[k]);
    __syncthreads();
    if (threadIdx.x == 0 && r != 0) atomicAdd(acc6 + k, r);
  }
}

__global__ void mark_kernel(const unsigned char *__restrict__ member, int n, int shape_id, int *__restrict__ assigned) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && member[i]) assigned[i] = shape_id;
}

struct IsUnassigned {
  const int *assigned;
  __device__ bool operator()(const int &i) const { return assigned[i] == -1; }
};

// ---- host pieces ----------------------------------------------------------------------------------------
// cross closing (DilateCross then ErodeCross, no wrapping) + 8-connected labelling; returns the
// per-pixel mask of the component with the most pixels (first one on ties), R/Bitmap.cpp:154,459,633.
void largest_component(std::vector<unsigned char> &bmp, int ue, int ve, std::vector<unsigned char> &mask) {
  std::vector<unsigned char> dil((size_t) ue * ve), ero((size_t) ue * ve);
  auto at = [&](const std::vector<unsigned char> &b, int x, int y, unsigned char oob) -> unsigned char {
    if (x < 0 || y < 0 || x >= ue || y >= ve) return oob;
    return b[(size_t) y * ue + x];
  };
  for (int y = 0; y < ve; ++y)
    for (int x = 0; x < ue; ++x)
      dil[(size_t) y * ue + x] = at(bmp, x, y, 0) | at(bmp, x - 1, y, 0) | at(bmp, x + 1, y, 0) | at(bmp, x, y - 1, 0) | at(bmp, x, y + 1, 0);
  for (int y = 0; y < ve; ++y)
    for (int x = 0; x < ue; ++x)
      ero[(size_t) y * ue + x] = at(dil, x, y, 1) & at(dil, x - 1, y, 1) & at(dil, x + 1, y, 1) & at(dil, x, y - 1, 1) & at(dil, x, y + 1, 1);
  bmp = ero;
  // 8-connected components by flood fill in raster order (component ids in order of first pixel)
  std::vector<int> lab((size_t) ue * ve, 0);
  std::vector<size_t> sizes(1, 0);
  std::vector<int> stack;
  int cur = 0;
  for (int y0 = 0; y0 < ve; ++y0)
    for (int x0 = 0; x0 < ue; ++x0) {
      size_t id0 = (size_t) y0 * ue + x0;
      if (!bmp[id0] || lab[id0]) continue;
      ++cur;
      sizes.push_back(0);
      stack.push_back((int) id0);
      lab[id0] = cur;
      while (!stack.empty()) {
        int id = stack.back();
        stack.pop_back();
        sizes[cur]++;
        int x = id % ue, y = id / ue;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            int xx = x + dx, yy = y + dy;
            if (xx < 0 || yy < 0 || xx >= ue || yy >= ve) continue;
            size_t j = (size_t) yy * ue + xx;
            if (bmp[j] && !lab[j]) { lab[j] = cur; stack.push_back((int) j); }
          }
      }
    }
  mask.assign((size_t) ue * ve, 0);
  if (cur == 0) return;
  int best = 1;
  for (int c = 2; c <= cur; ++c) if (sizes[best] < sizes[c]) best = c;
  for (size_t i = 0; i < lab.size(); ++i) mask[i] = lab[i] == best;
}

// PlanePrimitiveShape(normal, position): plane + in-plane frame
__host__ __device__ PlaneFrame make_plane_frame(const float nrm3[3], const float pos3[3]) {
  PlaneFrame f;
  float u[3], v[3];
  frame_from_normal(nrm3, u, v);
  float dist = (pos3[0] * nrm3[0] + pos3[1] * nrm3[1]) + pos3[2] * nrm3[2];   // m_pos.dot(m_normal)
  f.pl = make_float4(nrm3[0], nrm3[1], nrm3[2], dist);
  f.pos = make_float3(pos3[0], pos3[1], pos3[2]);
  f.u = make_float3(u[0], u[1], u[2]);
  f.v = make_float3(v[0], v[1], v[2]);
  return f;
}

// Can a point outside the band (|d_band(x)| < halfwidth) be within eps3 of plane `pl`?  d_pl(x) - d_band(x) is affine
// in x, so over the bounding box of the cloud its extreme values sit at the corners; if they stay below
// halfwidth - eps3 (minus float slack) every inlier of `pl` lies inside the band.  The fitted normal may be flipped.
__host__ __device__ bool band_covers_plane(const float4 &band_pl, float halfwidth, const float4 &pl, const float mn[3], const float mx[3],
                                           float ext, float eps3) {
  double worst[2] = {0, 0};
  for (int k = 0; k < 8; ++k) {
    const double x = (k & 1) ? mx[0] : mn[0], y = (k & 2) ? mx[1] : mn[1], z = (k & 4) ? mx[2] : mn[2];
    const double db = band_pl.w - (band_pl.x * x + band_pl.y * y + band_pl.z * z);
    const double dp = pl.w - (pl.x * x + pl.y * y + pl.z * z);
    worst[0] = fmax(worst[0], fabs(dp - db));
    worst[1] = fmax(worst[1], fabs(-dp - db));
  }
  const double slack = 1e-5 * (double) ext + 1e-3 * eps3;
  return fmin(worst[0], worst[1]) + eps3 + slack <= (double) halfwidth;
}

// ---- candidate refinement as ONE thread-block-cluster kernel --------------------------------------------------
// The acceptance test of a candidate (RansacShapeDetector.cpp:619-655) is a chain of up to four full evaluations
// -- inlier flags + (u, v) box -> bitmap -> closing + connected components -> members + score -> covariance ->
// least-squares refit -> next evaluation -- each a handful of passes over the candidate's band (10^5 points) with
// a tiny serial decision in between.  As separate launches that is ~20 kernels and 3-4 host round trips per
// candidate, all latency.  Here one cluster of kRefCluster CTAs (hardware co-scheduled, barrier.cluster between
// the passes) runs the whole chain, the refit and the accept logic included, on 16 of the 148 SMs, so the other
// lane / other contexts keep the rest of the GPU busy and the host synchronises once per candidate.
constexpr int kRefThreads = 1024;
constexpr int kRefUnroll = 4;        // band points per thread and batch
constexpr int kRefUnroll1 = 2;       // ... in the first pass (position + normal per point)
constexpr int kRefSmemPix = 1 << 16;  // bitmaps up to this many pixels are rasterised / looked up through shared-memory bits

struct RefineCtl {        // written by the boss thread between evaluations, read by every thread after the barrier
  PlaneFrame frame;
  int go;                 // 1: evaluate `frame` into membership map `member_sel`
  int member_sel;
  int pad0, pad1;
};
struct RefineOut {
  long long acc_size;     // support of the accepted (possibly refitted) plane; 0 if the candidate has no inliers
  float acc_n[3], acc_p[3];
  int acc_sel;            // membership map that holds its members (0 = A, 1 = B)
  int status;             // 0 = done, 1 = bitmap larger than the device cap (host path), 2 = a refit left the band
  int evals;
  int n_band;             // points of the candidate's band (algorithmic bytes of the launch = 28 B x n_band x evals)
  unsigned int phase_ns[6];   // device time per phase summed over the evaluations (flags, raster, components, select, covariance, decision)
  double acc_score;           // gaussian-weighted score (Candidate::WeightedScore) of the accepted plane
};
struct RefineArgs {
  // The band of a candidate is stored in `segs` segments of seg_cap points each (one per CTA of the refinement cluster):
  // points [4096 b, 4096 (b + 1)) of the cloud go to segment b % segs, so a segment never holds more than seg_cap points
  // and every CTA refines the part of the band it (or band_compact_kernel on its behalf) wrote -- band data never
  // crosses CTAs, only the reductions do.  d_nb[s] = points in segment s.
  float4 *posB, *nrmB;
  int *idxB, *d_nb;
  int seg_cap, segs;
  unsigned char *flag;
  int *pix;
  unsigned char *bmp, *btmp, *bmask;
  int *lab, *ccnt;
  unsigned char *member_a, *member_b;     // membership of the band points, by BAND position (two maps: candidate / clone)
  int *uvbox;
  double *acc;
  RefineCtl *ctl;
  RefineOut *out;
  float4 cand_pl, band_pl;
  float cand_pos[3];                      // a point on the candidate plane (the frame origin of its first evaluation)
  float eps3, nthresh, bmp_eps, band_halfwidth, ext;
  float mn[3], mx[3];
  int min_support, band_full;
};
__host__ __device__ inline int band_seg_cap(long long n, int segs) { return (int) ((n / kBandBlock / segs + 1) * kBandBlock); }

// state of the accept loop (detect_planes_dev's pool walk) while accept_loop_kernel runs it on the device
struct BatchState {
  long long m;                            // unassigned points
  double drawn;                           // drawn candidates (rescaled after every accepted shape)
  double prob;
  int n_found, stop, nlevels, min_support;
};
enum { kVerdictSkipped = 0, kVerdictNotEligible = 1, kVerdictFallback = 2, kVerdictRejected = 3, kVerdictAccepted = 5 };
struct Verdict { int kind, shape_id; long long size; float n[3], p[3]; int evals, n_band; unsigned int phase_ns[6]; unsigned int band_ns, pad; };
struct PoolCand { float4 pl; double est; double pad; };
struct AcceptCtl { int go, accept, shape_id, pad; };      // boss -> cluster, per candidate
struct AcceptArgs {
  RefineArgs r;                           // buffers and thresholds; the candidate plane is filled in per pool entry
  const float4 *pos, *nrm;
  int *assigned;
  int n, k;
  const PoolCand *pool;
  BatchState *state;
  Verdict *verdict;
  AcceptCtl *actl;
  float band;
};

// N sums at once: two block barriers in total instead of two per value; results valid in thread 0
template <int N>
__device__ __forceinline__ void block_sum_n(double (&v)[N], double *sh /* N * 32 */) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k)
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) sh[k * 32 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      double x = l < nw ? sh[k * 32 + l] : 0.0;
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      v[k] = x;
    }
  }
}
// (umin, vmin, umax, vmax) at once; results valid in thread 0
__device__ __forceinline__ void block_box4(float (&v)[4], float *sh /* 4 * 32 */) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_down_sync(0xffffffffu, v[k], o); v[k] = k < 2 ? fminf(v[k], t) : fmaxf(v[k], t); }
  __syncthreads();
  if (l == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) sh[k * 32 + w] = v[k];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x = l < nw ? sh[k * 32 + l] : (k < 2 ? 3.4e38f : -3.4e38f);
      for (int o = 16; o > 0; o >>= 1) { const float t = __shfl_down_sync(0xffffffffu, x, o); x = k < 2 ? fminf(x, t) : fmaxf(x, t); }
      v[k] = x;
    }
  }
}
__device__ __forceinline__ unsigned int cluster_rank() { unsigned int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned int cluster_size() { unsigned int r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// barrier over all threads of the cluster; release/acquire at cluster scope orders the global-memory traffic of the passes
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// The acceptance chain of ONE candidate on its (segmented) band, executed by every thread of a cluster; the verdict lands
// in *a.out.  cand_pl / cand_pos: the candidate (a.cand_pl for the host-driven launch, a pool entry in accept_loop_kernel).
__device__ void refine_candidate_dev(const RefineArgs &a, const float4 cand_pl, const float cand_pos0, const float cand_pos1, const float cand_pos2,
                                     const float4 band_pl) {
  __shared__ double sh_d[6 * 32];
  __shared__ float sh_f[4 * 32];
  __shared__ unsigned int sbits[kRefSmemPix / 32];     // per-CTA bit image of the bitmap (pass 2) / of the component mask (pass 4)
  const int rank = (int) cluster_rank(), nthr = (int) blockDim.x;
  const int tid = (int) threadIdx.x;
  const bool boss = rank == 0 && tid == 0;
  const int n = __ldcg(a.d_nb + rank);      // points of this CTA's band segment
  const int sb = rank * a.seg_cap;          // ... which starts here
  const float denom = 2.f / 9.f * a.eps3 * a.eps3;
  volatile RefineCtl *ctl = a.ctl;

  // boss-only state of the refinement loop (RansacShapeDetector.cpp:619-655; the host version is in detect_planes_dev),
  // kept in shared memory so that it does not occupy registers of the 16K threads that never touch it
  struct Boss {
    Eval clone;
    double cov_clone[6], newScore, oldScore;
    float acc_n[3], acc_p[3], fn[3], fp[3];
    long long acc_size;
    double acc_score;
    int acc_sel, work_sel, iter, status, evals;
    unsigned long long t_prev;
    unsigned int phase_ns[6];
  };
  __shared__ Boss B;
  auto lap = [&](int phase) {       // boss: %globaltimer at the phase boundaries (PLADE_TIMING prints the sums)
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    B.phase_ns[phase] += (unsigned int) (t - B.t_prev);
    B.t_prev = t;
  };
  if (boss) {
    B.clone = Eval{0, 0, {0, 0, 0}, false};
    for (int k = 0; k < 6; ++k) B.cov_clone[k] = 0;
    B.newScore = B.oldScore = 0;
    B.acc_n[0] = cand_pl.x; B.acc_n[1] = cand_pl.y; B.acc_n[2] = cand_pl.z;
    // position of the 3-point plane: any point on it (the reference keeps the first sample)
    B.acc_p[0] = cand_pos0; B.acc_p[1] = cand_pos1; B.acc_p[2] = cand_pos2;
    for (int k = 0; k < 3; ++k) B.fn[k] = B.fp[k] = 0;
    B.acc_size = 0;
    B.acc_score = 0;
    B.acc_sel = 0; B.work_sel = 1; B.iter = 0; B.status = 0; B.evals = 0;
    for (int k = 0; k < 6; ++k) B.phase_ns[k] = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(B.t_prev));
  }

  auto publish = [&](const PlaneFrame &f, int sel, int go) {     // boss: next evaluation + clean accumulators
    RefineCtl c;
    c.frame = f; c.go = go; c.member_sel = sel; c.pad0 = c.pad1 = 0;
    *a.ctl = c;
    a.uvbox[0] = a.uvbox[1] = f2o(3.4e38f);
    a.uvbox[2] = a.uvbox[3] = f2o(-3.4e38f);
    for (int k = 0; k < 16; ++k) a.acc[k] = 0.0;
    // (made visible to the other CTAs by the release / acquire of the cluster barrier that follows)
  };
  // boss: top of the do-loop body -- refit the clone and, if the refit stays inside the band, evaluate it next
  auto try_next = [&]() -> bool {
    ++B.iter;
    B.oldScore = B.newScore;
    if (!fit_plane_from_cov(B.clone, B.cov_clone, B.fn, B.fp)) return false;
    PlaneFrame f2 = make_plane_frame(B.fn, B.fp);
    if (!a.band_full && !band_covers_plane(band_pl, a.band_halfwidth, f2.pl, a.mn, a.mx, a.ext, a.eps3)) { B.status = 2; return false; }
    publish(f2, B.work_sel, 1);
    return true;
  };
  if (boss) {
    PlaneFrame f0 = make_plane_frame(B.acc_n, B.acc_p);
    f0.pl.w = cand_pl.w;
    publish(f0, B.acc_sel, 1);
  }
  cluster_barrier();

  for (int ev = 0; ev < 4; ++ev) {
    if (!ctl->go) break;                       // uniform over the cluster
    PlaneFrame f;
    f.pl = make_float4(ctl->frame.pl.x, ctl->frame.pl.y, ctl->frame.pl.z, ctl->frame.pl.w);
    f.pos = make_float3(ctl->frame.pos.x, ctl->frame.pos.y, ctl->frame.pos.z);
    f.u = make_float3(ctl->frame.u.x, ctl->frame.u.y, ctl->frame.u.z);
    f.v = make_float3(ctl->frame.v.x, ctl->frame.v.y, ctl->frame.v.z);
    unsigned char *member = ctl->member_sel ? a.member_b : a.member_a;

    // Every pass walks the band in batches of kRefUnroll points per thread with all loads of a batch issued before
    // the first use: a pass is a dependent-latency chain (a dozen points per thread), not a bandwidth problem.
    // pass 1: inlier flags at eps3 + (u, v) box (flag_uv_kernel)
    {
      float umin = 3.4e38f, vmin = 3.4e38f, umax = -3.4e38f, vmax = -3.4e38f;
      for (int base = tid; base < n; base += kRefUnroll1 * nthr) {
        float4 p[kRefUnroll1], nr[kRefUnroll1];
#pragma unroll
        for (int u = 0; u < kRefUnroll1; ++u) {
          const int ic = sb + min(base + u * nthr, n - 1);    // clamped: the loads are unconditional (no partially
          p[u] = a.posB[ic]; nr[u] = a.nrmB[ic];              // defined arrays, which would live in local memory)
        }
#pragma unroll
        for (int u = 0; u < kRefUnroll1; ++u) {
          const int i = base + u * nthr;
          if (i >= n) continue;
          const bool in = compatible(f.pl, p[u], nr[u], a.eps3, a.nthresh);
          a.flag[sb + i] = in ? 1 : 0;
          if (in) {
            float px = __fsub_rn(p[u].x, f.pos.x), py = __fsub_rn(p[u].y, f.pos.y), pz = __fsub_rn(p[u].z, f.pos.z);
            float uu = __fadd_rn(__fadd_rn(__fmul_rn(px, f.u.x), __fmul_rn(py, f.u.y)), __fmul_rn(pz, f.u.z));
            float vv = __fadd_rn(__fadd_rn(__fmul_rn(px, f.v.x), __fmul_rn(py, f.v.y)), __fmul_rn(pz, f.v.z));
            umin = fminf(umin, uu); umax = fmaxf(umax, uu); vmin = fminf(vmin, vv); vmax = fmaxf(vmax, vv);
          }
        }
      }
      float box4[4] = {umin, vmin, umax, vmax};
      block_box4(box4, sh_f);
      if (threadIdx.x == 0) {
        atomicMin(a.uvbox + 0, f2o(box4[0])); atomicMin(a.uvbox + 1, f2o(box4[1]));
        atomicMax(a.uvbox + 2, f2o(box4[2])); atomicMax(a.uvbox + 3, f2o(box4[3]));
      }
    }
    cluster_barrier();
    if (boss) lap(0);
    // pass 2: bitmap geometry + rasterisation (raster_dev_kernel)
    int box[4];
    for (int k = 0; k < 4; ++k) box[k] = __ldcg(a.uvbox + k);
    const BmpInfo I = bmp_info_from_box(box, a.bmp_eps);
    const int P = I.ok ? I.ue * I.ve : 0;
    const bool small = P <= kRefSmemPix;     // (thousands of band points fall on each pixel: setting bits in shared memory
    if (I.ok) {                              //  and flushing once per CTA avoids ~10^5 same-address stores in L2)
      if (small) {
        for (int w = threadIdx.x; w < (P + 31) / 32; w += nthr) sbits[w] = 0u;
        __syncthreads();
      }
      for (int base = tid; base < n; base += kRefUnroll * nthr) {
        float4 p[kRefUnroll];
        unsigned int fl[kRefUnroll];
#pragma unroll
        for (int u = 0; u < kRefUnroll; ++u) {
          const int i = base + u * nthr, ic = sb + min(i, n - 1);
          fl[u] = a.flag[ic]; p[u] = a.posB[ic];
          if (i >= n) fl[u] = 0;
        }
#pragma unroll
        for (int u = 0; u < kRefUnroll; ++u) {
          const int i = base + u * nthr;
          if (!fl[u]) continue;
          float px = __fsub_rn(p[u].x, f.pos.x), py = __fsub_rn(p[u].y, f.pos.y), pz = __fsub_rn(p[u].z, f.pos.z);
          float uu = __fadd_rn(__fadd_rn(__fmul_rn(px, f.u.x), __fmul_rn(py, f.u.y)), __fmul_rn(pz, f.u.z));
          float vv = __fadd_rn(__fadd_rn(__fmul_rn(px, f.v.x), __fmul_rn(py, f.v.y)), __fmul_rn(pz, f.v.z));
          int bu = (int) floorf(__fdiv_rn(__fsub_rn(uu, I.umin), a.bmp_eps));
          int bv = (int) floorf(__fdiv_rn(__fsub_rn(vv, I.vmin), a.bmp_eps));
          bu = min(max(bu, 0), I.ue - 1);
          bv = min(max(bv, 0), I.ve - 1);
          const int id = bu + bv * I.ue;
          a.pix[sb + i] = id;
          if (small) atomicOr(&sbits[id >> 5], 1u << (id & 31));
          else a.bmp[id] = 1;
        }
      }
      if (small) {
        __syncthreads();
        for (int w = threadIdx.x; w < (P + 31) / 32; w += nthr) {
          unsigned int bits = sbits[w];
          while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; a.bmp[32 * w + b] = 1; }
        }
      }
    }
    cluster_barrier();
    if (boss) lap(1);
    // pass 3: closing + connected components on CTA 0 (cc_kernel); leaves the bitmap cleared
    if (rank == 0) cc_block(a.bmp, a.btmp, a.lab, a.ccnt, I, a.bmask);
    cluster_barrier();
    if (boss) lap(2);
    // pass 4: members of the largest component, their count / position sums / weighted score (select_dev_kernel);
    // the member bit is also kept next to the flag (bit 1) for pass 5
    {
      if (I.ok && small) {      // component mask -> shared-memory bits (one L2 read per pixel and CTA instead of one per band point)
        for (int w = threadIdx.x; w < (P + 31) / 32; w += nthr) sbits[w] = 0u;
        __syncthreads();
        for (int q = threadIdx.x; q < P; q += nthr) if (__ldcg(a.bmask + q)) atomicOr(&sbits[q >> 5], 1u << (q & 31));
        __syncthreads();
      }
      double cnt = 0, sx = 0, sy = 0, sz = 0, sc = 0;
      for (int base = tid; base < n; base += kRefUnroll * nthr) {
        float4 p[kRefUnroll];
        int px[kRefUnroll];
        unsigned int fl[kRefUnroll], mk[kRefUnroll];
#pragma unroll
        for (int u = 0; u < kRefUnroll; ++u) {
          const int i = base + u * nthr, ic = sb + min(i, n - 1);
          fl[u] = a.flag[ic]; px[u] = a.pix[ic];
          p[u] = a.posB[ic];
          if (i >= n || !I.ok) fl[u] = 0;
        }
#pragma unroll
        for (int u = 0; u < kRefUnroll; ++u)       // pix is only defined where flagged
          mk[u] = !fl[u] ? 0u : small ? ((sbits[px[u] >> 5] >> (px[u] & 31)) & 1u) : (unsigned int) __ldcg(a.bmask + px[u]);
#pragma unroll
        for (int u = 0; u < kRefUnroll; ++u) {
          const int i = base + u * nthr;
          if (i >= n) continue;
          const bool mem = fl[u] && mk[u];
          member[sb + i] = mem ? 1 : 0;
          a.flag[sb + i] = (unsigned char) ((fl[u] & 1) | (mem ? 2 : 0));
          if (mem) {
            float dp = __fadd_rn(__fadd_rn(__fmul_rn(f.pl.x, p[u].x), __fmul_rn(f.pl.y, p[u].y)), __fmul_rn(f.pl.z, p[u].z));
            float d = fabsf(__fsub_rn(f.pl.w, dp));
            cnt += 1; sx += p[u].x; sy += p[u].y; sz += p[u].z;
            sc += (double) expf(-d * d / denom);
          }
        }
      }
      double r5[5] = {cnt, sx, sy, sz, sc};
      block_sum_n<5>(r5, sh_d);
      if (threadIdx.x == 0 && r5[0] > 0) {
        atomicAdd(a.acc + 0, r5[0]); atomicAdd(a.acc + 1, r5[1]); atomicAdd(a.acc + 2, r5[2]); atomicAdd(a.acc + 3, r5[3]); atomicAdd(a.acc + 4, r5[4]);
      }
    }
    cluster_barrier();
    if (boss) lap(3);
    // pass 5: covariance about float(mean) of the members (cov_dev_kernel)
    {
      const double cnt = __ldcg(a.acc + 0);
      if (cnt > 0) {
        const float mx = (float) (__ldcg(a.acc + 1) / cnt), my = (float) (__ldcg(a.acc + 2) / cnt), mz = (float) (__ldcg(a.acc + 3) / cnt);
        double c[6] = {0, 0, 0, 0, 0, 0};
        for (int base = tid; base < n; base += kRefUnroll * nthr) {
          float4 p[kRefUnroll];
          unsigned int fl[kRefUnroll];
#pragma unroll
          for (int u = 0; u < kRefUnroll; ++u) {
            const int i = base + u * nthr, ic = sb + min(i, n - 1);
            fl[u] = a.flag[ic]; p[u] = a.posB[ic];
            if (i >= n) fl[u] = 0;
          }
#pragma unroll
          for (int u = 0; u < kRefUnroll; ++u) {
            if (!(fl[u] & 2)) continue;
            double dx = (double) (p[u].x - mx), dy = (double) (p[u].y - my), dz = (double) (p[u].z - mz);
            c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
          }
        }
        block_sum_n<6>(c, sh_d);
        if (threadIdx.x == 0) for (int k = 0; k < 6; ++k) if (c[k] != 0) atomicAdd(a.acc + 8 + k, c[k]);
      }
    }
    cluster_barrier();
    // decision (boss thread): the loop of RansacShapeDetector.cpp:619-655
    if (boss) {
      lap(4);
      ++B.evals;
      Eval e;
      double cov[6];
      e.size = (long long) __ldcg(a.acc + 0);
      e.sum[0] = __ldcg(a.acc + 1); e.sum[1] = __ldcg(a.acc + 2); e.sum[2] = __ldcg(a.acc + 3);
      e.score = __ldcg(a.acc + 4);
      e.ok = e.size > 0;
      for (int k = 0; k < 6; ++k) cov[k] = __ldcg(a.acc + 8 + k);
      bool go = false;
      if (ev == 0) {
        if (I.overflow) B.status = 1;
        else {
          B.acc_size = e.ok ? e.size : 0;
          B.acc_score = e.ok ? e.score : 0;
          if (e.ok) {
            B.clone = e;
            for (int k = 0; k < 6; ++k) B.cov_clone[k] = cov[k];
            B.newScore = B.clone.score;
            B.iter = 0;
            go = try_next();
          }
        }
      } else if (!I.overflow) {
        B.newScore = e.score;
        if (e.ok) {
          B.clone = e;
          for (int k = 0; k < 6; ++k) B.cov_clone[k] = cov[k];
          if (B.newScore > B.oldScore && e.size > a.min_support) {
            for (int k = 0; k < 3; ++k) { B.acc_n[k] = B.fn[k]; B.acc_p[k] = B.fp[k]; }
            B.acc_size = e.size;
            B.acc_score = e.score;
            const int t = B.acc_sel; B.acc_sel = B.work_sel; B.work_sel = t;      // the B.clone becomes the candidate
          }
          if (B.newScore > B.oldScore && B.iter < 3) go = try_next();
        }
      }
      if (!go) a.ctl->go = 0;
    }
    cluster_barrier();
    if (boss) lap(5);
  }
  if (boss) {
    RefineOut o;
    o.acc_size = B.acc_size;
    for (int k = 0; k < 3; ++k) { o.acc_n[k] = B.acc_n[k]; o.acc_p[k] = B.acc_p[k]; }
    o.acc_sel = B.acc_sel; o.status = B.status; o.evals = B.evals; o.acc_score = B.acc_score;
    o.n_band = 0;
    for (int q = 0; q < a.segs; ++q) o.n_band += __ldcg(a.d_nb + q);
    for (int k = 0; k < 6; ++k) o.phase_ns[k] = B.phase_ns[k];
    *a.out = o;
  }
}

// host-driven launch: one candidate whose band band_compact_kernel has already built
__global__ void __launch_bounds__(kRefThreads, 1) refine_cluster_kernel(const __grid_constant__ RefineArgs a) {
  refine_candidate_dev(a, a.cand_pl, a.cand_pos[0], a.cand_pos[1], a.cand_pos[2], a.band_pl);
}

// CandidateFailureProbability, R/RansacShapeDetector.h:61-67 (reqSamples = 3)
__host__ __device__ inline double failure_probability_hd(double size, double n, double drawn, double levels) {
  const double p = pow(1.0 - size / (n * levels * 4.0), drawn);
  return p < 1.0 ? p : 1.0;
}
// drawn candidates that stay valid after a shape of `size` points left a cloud of m (RansacShapeDetector.cpp:673-674);
// the cube is spelled out so that host and device round alike
__host__ __device__ inline double rescale_drawn(double drawn, long long size, long long m) {
  const float x = 1.f - ((float) size / (float) m);
  return (double) ((x * x) * x) * drawn;
}

// The accept loop of one scoring round (detect_planes_dev's pool walk, RansacShapeDetector.cpp:583-675) on the device:
// ONE cluster walks the pool in order -- eligibility test, band of the candidate (a streaming pass of the cluster over
// the cloud), acceptance chain (refine_candidate_dev), accept / reject decision, removal of the accepted points -- with
// exactly the sequential semantics of the host loop and no host round trip between candidates.  It stops at the first
// candidate that is no longer eligible, or that needs the host (bitmap beyond the device cap, refit leaving its band).
__global__ void __launch_bounds__(kRefThreads, 1) accept_loop_kernel(const __grid_constant__ AcceptArgs A) {
  __shared__ int s_nb;
  const RefineArgs &a = A.r;
  const int rank = (int) cluster_rank(), segs = (int) cluster_size(), nthr = (int) blockDim.x, tid = (int) threadIdx.x;
  const bool boss = rank == 0 && tid == 0;
  const unsigned lane = tid & 31;
  volatile BatchState *st = A.state;
  volatile AcceptCtl *ac = A.actl;
  const int sb = rank * a.seg_cap;
  const long long n_blocks = ((long long) A.n + kBandBlock - 1) / kBandBlock;
  for (int j = 0; j < A.k; ++j) {
    const float4 pl = A.pool[j].pl;
    unsigned long long t0 = 0;
    // ---- eligible?  (entry 0 was tested by the host with the same state)
    if (boss) {
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      const long long m = st->m;
      const double est = A.pool[j].est;
      const bool ok = !st->stop && (j == 0 || (est >= (double) st->min_support && failure_probability_hd(est, (double) m, st->drawn, (double) st->nlevels) <= st->prob));
      if (!ok && !st->stop) { Verdict v = {}; v.kind = kVerdictNotEligible; A.verdict[j] = v; st->stop = 1; }
      ac->go = ok ? 1 : 0;
    }
    if (tid == 0) s_nb = 0;
    cluster_barrier();
    if (!ac->go) break;                        // uniform over the cluster
    // ---- band: the unassigned points within A.band of the candidate plane, block b of the cloud -> segment b % segs
    for (long long b = rank; b < n_blocks; b += segs) {
      constexpr int kU = kBandBlock / kRefThreads;      // 4 loads in flight per thread
      int as[kU];
      float4 p[kU];
      const long long base = b * kBandBlock + tid;
#pragma unroll
      for (int u = 0; u < kU; ++u) { const long long i = base + u * kRefThreads; as[u] = i < A.n ? __ldcg(A.assigned + i) : 0; }
#pragma unroll
      for (int u = 0; u < kU; ++u) { const long long i = base + u * kRefThreads; p[u] = (i < A.n && as[u] == -1) ? __ldg(A.pos + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const long long i = base + u * kRefThreads;
        const float dp = __fadd_rn(__fadd_rn(__fmul_rn(pl.x, p[u].x), __fmul_rn(pl.y, p[u].y)), __fmul_rn(pl.z, p[u].z));
        const bool in = i < A.n && as[u] == -1 && fabsf(__fsub_rn(pl.w, dp)) < A.band;
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        if (!bal) continue;
        int start = 0;
        if (lane == 0) start = atomicAdd(&s_nb, __popc(bal));
        start = __shfl_sync(0xffffffffu, start, 0);
        if (in) {
          const int q = sb + start + __popc(bal & ((1u << lane) - 1));
          a.posB[q] = p[u]; a.nrmB[q] = __ldg(A.nrm + i); a.idxB[q] = (int) i;
        }
      }
    }
    __syncthreads();
    const int n_seg = s_nb;                    // (read before thread 0 can reset it for the next candidate)
    if (tid == 0) a.d_nb[rank] = n_seg;
    cluster_barrier();
    unsigned int band_ns = 0;
    if (boss) { unsigned long long t1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); band_ns = (unsigned int) (t1 - t0); }
    // ---- acceptance chain; n * dist is a point on the 3-point plane
    refine_candidate_dev(a, pl, pl.x * pl.w, pl.y * pl.w, pl.z * pl.w, pl);
    // ---- decision (RansacShapeDetector.cpp:297,423-430,659-675)
    if (boss) {
      const volatile RefineOut *o = a.out;
      Verdict v = {};
      v.evals = o->evals; v.n_band = o->n_band; v.band_ns = band_ns;
      for (int k = 0; k < 6; ++k) v.phase_ns[k] = o->phase_ns[k];
      const long long m = st->m, acc_size = o->acc_size;
      int accept = 0;
      if (o->status != 0) { v.kind = kVerdictFallback; st->stop = 1; }
      else if (acc_size < (long long) st->min_support) v.kind = kVerdictRejected;
      else {
        v.kind = kVerdictAccepted;
        v.shape_id = st->n_found;
        v.size = acc_size;
        for (int k = 0; k < 3; ++k) { v.n[k] = o->acc_n[k]; v.p[k] = o->acc_p[k]; }
        st->n_found = v.shape_id + 1;
        st->drawn = rescale_drawn(st->drawn, acc_size, m);
        st->m = m - acc_size;
        if (m - acc_size < (long long) st->min_support || m - acc_size < 3) st->stop = 1;
        accept = 1;
      }
      A.verdict[j] = v;
      ac->accept = accept; ac->shape_id = v.shape_id;
      ac->pad = o->acc_sel;
    }
    cluster_barrier();
    // ---- accept: the members leave the cloud
    if (ac->accept) {
      const unsigned char *member = ac->pad ? a.member_b : a.member_a;
      const int shape_id = ac->shape_id;
      for (int i = tid; i < n_seg; i += nthr) if (member[sb + i]) A.assigned[a.idxB[sb + i]] = shape_id;
    }
    // (the barrier at the top of the next iteration orders these writes before the next band pass reads `assigned`)
  }
  // entries the loop never reached keep kind == kVerdictSkipped (the host zeroes the verdicts before the launch)
}

double failure_probability(double size, double n, double drawn, double levels) { return failure_probability_hd(size, n, drawn, levels); }

struct FoundPlane { float n[3]; float pos[3]; long long size; };

struct RansacScratch {
  DevBuf<unsigned int> keys, keys_alt, counts, counts_sorted, counts2;
  DevBuf<int> order, order_alt, assigned, pix, misc;
  DevBuf<float4> cand, sub, sub1, cand_top, band_pos, band_nrm;
  DevBuf<int> band_idx;
  DevBuf<int> cidx, cidx_sorted;
  DevBuf<unsigned char> flag, member, member2, bitmap, mask, cub_tmp, bmp_dev, bmp_tmp, mask_dev;
  DevBuf<int> cc_lab, cc_cnt, remap;
  DevBuf<float> mean3;
  bool bmp_dev_clean = false;
  DevBuf<double> acc;
  DevBuf<double> refine_mem;      // RefineCtl at +0, RefineOut at +128 bytes
  DevBuf<int> cand_live;           // slots of `cand` that hold a plane this round (compact_live_slots)
  DevBuf<unsigned int> cand_bits;  // one bit per slot: the draw passed (gen_candidates_kernel)
  DevBuf<unsigned int> round_buf;  // one scoring round: counts | counts2 | n_valid | selected planes (see kRound* offsets)
  PinBuf<unsigned int> round_host; // page-locked landing zone of the round's results and of the cluster kernel's verdict
  PinBuf<float4> pool_host;
  DevBuf<unsigned char> memb_a, memb_b;   // band-local membership maps of refine_cluster_kernel
  // state of the last detection on this lane, so that extract() can CONTINUE it with a halved min_support instead of
  // starting over (the planes found so far are exactly the ones a fresh run would find first)
  struct Resume {
    bool valid = false;
    const void *cloud = nullptr;
    int n = 0, m = 0, cur_is_order2 = 1;
    double drawn = 0;
    unsigned long long round_seed = 0;
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
    std::vector<FoundPlane> found;
  } resume;
  DevBuf<unsigned char> batch_mem;        // accept_loop_kernel: BatchState | AcceptCtl | pool entries | verdicts (see kAcc* offsets)
  PinBuf<unsigned char> batch_host;       // page-locked mirror: upload source and the verdicts' landing zone
};
constexpr int kAcceptMax = 64;            // pool entries one accept_loop_kernel launch can walk (= the pool size)
constexpr int kAccState = 0, kAccCtl = 64, kAccPool = 128, kAccVerdict = kAccPool + kAcceptMax * (int) sizeof(PoolCand),
              kAccBytes = kAccVerdict + kAcceptMax * (int) sizeof(Verdict);
static_assert(sizeof(BatchState) <= 64 && sizeof(AcceptCtl) <= 64 && sizeof(PoolCand) == 32 && sizeof(Verdict) % 8 == 0, "batch_mem layout");

// cluster size of refine_cluster_kernel on this device: 16 (non-portable) when the GPU can co-schedule it, else 8
int refine_block_threads() {
  static const int t = [] { const char *e = getenv("PLADE_REFINE_THREADS"); int v = e ? atoi(e) : kRefThreads; return (v == 256 || v == 512 || v == 1024) ? v : kRefThreads; }();
  return t;
}
// Function attributes and occupancy are per device: plade_register_batch and the CLI run contexts on several GPUs in one
// process, so the non-portable cluster opt-in and the occupancy query are made once per DEVICE (not once per process).
int refine_cluster_size(int device) {
  constexpr int kMaxDev = 64;
  static std::mutex mu;
  static int cached[kMaxDev];
  static bool known[kMaxDev] = {};
  std::lock_guard<std::mutex> lock(mu);       // (the two lanes of a context ask concurrently)
  if (device < 0 || device >= kMaxDev) return 0;
  if (known[device]) return cached[device];
  int size = 0;
  if (!getenv("PLADE_NO_CLUSTER_REFINE")) {
    const char *e = getenv("PLADE_REFINE_CLUSTER");
    const int first = e ? atoi(e) : 16;
    int cur = -1;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    for (int want : {first, 8}) {
      if (want < 1 || want > 16) continue;
      if (want > 8 && (cudaFuncSetAttribute(refine_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                       cudaFuncSetAttribute(accept_loop_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) { cudaGetLastError(); continue; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(want); cfg.blockDim = dim3(refine_block_threads());
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = want; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n_clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&n_clusters, refine_cluster_kernel, &cfg) == cudaSuccess && n_clusters >= 1) { size = want; break; }
      cudaGetLastError();
    }
    if (cur != device && cur >= 0) cudaSetDevice(cur);
  }
  cached[device] = size;
  known[device] = true;
  return size;
}

}  // namespace

// one scratch set per context and lane: lane 0 = the context's own stream, lane 1 = the helper stream used to
// extract the planes of the second cloud concurrently
static RansacScratch &scratch_of(Registrar &r, int lane) {
  if (!r.ransac_scratch[lane]) r.ransac_scratch[lane] = new RansacScratch;
  return *static_cast<RansacScratch *>(r.ransac_scratch[lane]);
}
void free_ransac_scratch(Registrar &r) {
  for (int lane = 0; lane < 2; ++lane) { delete static_cast<RansacScratch *>(r.ransac_scratch[lane]); r.ransac_scratch[lane] = nullptr; }
}

// (group may be the same array as assigned: the top-40 cut of extract() remaps in place)
__global__ void remap_group_kernel(const int *assigned, int n, const int *__restrict__ remap, int n_remap, int *group) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = assigned[i];
  group[i] = (a >= 0 && a < n_remap) ? remap[a] : -1;
}

std::vector<PlaneParam> Registrar::detect_planes_dev(const CloudDev &c, int min_support, DevBuf<int> &group_out, int lane, bool resume) {
  std::vector<PlaneParam> result;
  Device &dev = lane == 0 ? this->dev : this->dev2;          // shadows the member: everything below runs on the lane's stream
  auto mark = [&](const char *name) { if (lane == 0) this->mark(name); };
  mark("ransac_begin");
  const int n = (int) c.n;
  if (n < 3) { std::cerr << "point set has less than 3 points" << std::endl; return result; }
  cudaStream_t s = dev.stream;
  RansacScratch &rs = scratch_of(*this, lane);
  const int blocks_n = std::min(div_up(n, 256), dev.num_sms * 8);

  // bounding box -> scale (bug-compatible: z ignored)
  int *d_misc = rs.misc.ensure(128);
  RansacScratch::Resume &rz = rs.resume;
  const bool cont = resume && rz.valid && rz.cloud == (const void *) c.pos.p && rz.n == n;      // continue the previous detection of this cloud
  float mn[3], mx[3];
  if (cont) { for (int k = 0; k < 3; ++k) { mn[k] = rz.mn[k]; mx[k] = rz.mx[k]; } }
  else {
    int init[6];
    float big = 3.4e38f, nbig = -3.4e38f;
    int bi, nbi;
    memcpy(&bi, &big, 4); memcpy(&nbi, &nbig, 4);
    nbi ^= 0x7fffffff;
    init[0] = init[1] = init[2] = bi; init[3] = init[4] = init[5] = nbi;
    PLADE_CUDA(cudaMemcpyAsync(d_misc, init, sizeof(init), cudaMemcpyHostToDevice, s));
    bbox_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, n, d_misc);
    PLADE_LAUNCH_CHECK();
    int h6[6];
    PLADE_CUDA(cudaMemcpyAsync(h6, d_misc, sizeof(h6), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    for (int k = 0; k < 3; ++k) {
      int a = h6[k], b = h6[3 + k];
      a = a >= 0 ? a : a ^ 0x7fffffff; b = b >= 0 ? b : b ^ 0x7fffffff;
      memcpy(&mn[k], &a, 4); memcpy(&mx[k], &b, 4);
    }
  }
  rz.valid = false;
  const float scale = std::max(mx[0] - mn[0], mx[1] - mn[1]);
  const float eps = params.ransac_dist_thresh * scale;
  const float bmp_eps = params.ransac_bitmap_reso * scale;
  const float nthresh = params.ransac_normal_thresh;
  const float eps3 = 3 * eps;
  const double prob = params.ransac_prob;

  // Morton order of all points
  unsigned int *keys = rs.keys.ensure(n), *keys2 = rs.keys_alt.ensure(n);
  int *order = rs.order.ensure(n), *order2 = rs.order_alt.ensure(n);
  float ext = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
  unsigned char *tmp = nullptr;
  int *assigned = rs.assigned.ensure(n);
  int *cur_order = order2, *alt_order = order;
  if (cont) {
    if (!rz.cur_is_order2) std::swap(cur_order, alt_order);
  } else {
    morton_kernel<<<div_up(n, 256), 256, 0, s>>>(c.pos.p, n, make_float3(mn[0], mn[1], mn[2]), ext > 0 ? 1.f / ext : 0.f, keys, order);
    PLADE_LAUNCH_CHECK();
    size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys2, order, order2, n, 0, 30, s);
    IsUnassigned pred{nullptr};
    cub::DeviceSelect::If(nullptr, tb2, order, order2, d_misc, n, pred, s);
    tmp = rs.cub_tmp.ensure(std::max(tb, tb2));
    cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys2, order, order2, n, 0, 30, s);
    dev.launches.add(8);
    PLADE_CUDA(cudaMemsetAsync(assigned, 0xff, sizeof(int) * n, s));
  }
  mark("ransac_morton");

  // band-indexed scratch: a segmented band (RefineArgs) spans up to seg_cap x 16 >= n positions
  const size_t band_cap_max = std::max<size_t>((size_t) n, (size_t) band_seg_cap(n, 16) * 16);
  unsigned char *flag = rs.flag.ensure(band_cap_max), *member_a = rs.member.ensure(n), *member_b = rs.member2.ensure(n);
  PLADE_CUDA(cudaMemsetAsync(member_a, 0, n, s));       // membership maps over all points; all-zero between candidates
  PLADE_CUDA(cudaMemsetAsync(member_b, 0, n, s));
  int *pix = rs.pix.ensure(band_cap_max);
  float4 *posB = rs.band_pos.ensure(band_cap_max), *nrmB = rs.band_nrm.ensure(band_cap_max);
  int *idxB = rs.band_idx.ensure(band_cap_max);
  rs.memb_a.ensure(band_cap_max); rs.memb_b.ensure(band_cap_max);
  const int blocks_b = std::min(blocks_n, dev.num_sms * 4);
  int n_band_builds = 0, n_band_full = 0, n_cluster_fallbacks = 0, refine_evals = 0, n_batches = 0, n_batch_cands = 0;
  double band_phase_ns = 0;
  bool force_single = false;
  double refine_phase_ns[6] = {0, 0, 0, 0, 0, 0};
  float4 *cand = rs.cand.ensure(kCandPerRound);
  double *acc = rs.acc.ensure(16);
  int *d_uvbox = d_misc + 16, *d_nsel = d_misc + 24, *d_nb = d_misc + 64;      // d_nb: one count per band segment (<= 16)

  std::vector<FoundPlane> found;
  int m = n;                       // unassigned points
  double drawn = 0;
  unsigned long long round_seed = params.seed;
  // (the drawn-candidate count starts again: the candidates of the previous pass were only kept when they promised >= its min_support)
  if (cont) { found = rz.found; m = rz.m; round_seed = rz.round_seed; }
  float4 best_pl = make_float4(0, 0, 0, 0);
  double best_est = 0;
  const int max_rounds = 4000;
  int rejects = 0, dry_rounds = 0;
  struct PoolEntry { double est; float4 pl; };
  constexpr int kPoolSize = 64;
  std::vector<PoolEntry> pool;
  std::vector<float4> banned;
  for (int round = 0; round < max_rounds && m >= min_support && m >= 3; ++round) {
    int nlevels = 1;
    while ((8ll << nlevels) < m) ++nlevels;
    // --- candidates + two-stage scores: all candidates vs a small subsample, the best of them vs the large one
    const int S = std::min(m, kSubsample), S1 = std::min(m, kStage1Points);
    float4 *sub = rs.sub.ensure((size_t) div_up(S, kScoreTile) * 2 * kScoreTile);
    float4 *sub1 = rs.sub1.ensure((size_t) div_up(S1, kScoreTile) * 2 * kScoreTile);
    unsigned int *rb = rs.round_buf.ensure(kRoundEnd);
    unsigned int *counts = rb, *counts2 = rb + kRoundCounts2;
    int *d_nvalid = reinterpret_cast<int *>(rb + kRoundNValid);
    float4 *cand_top = reinterpret_cast<float4 *>(rb + kRoundTop);
    int *cidx_sorted = rs.cidx_sorted.ensure(kStage2Cand);
    unsigned int *h_round = rs.round_host.ensure(kRoundEnd - kRoundCounts2 + 64);
    PLADE_CUDA(cudaMemsetAsync(rb, 0, sizeof(unsigned int) * kRoundTop, s));
    round_seed = mix64(round_seed + 1);
    static_assert(kCandPerRound % 128 == 0, "gen_candidates_kernel: one thread per slot, whole warps");
    unsigned int *valid_bits = rs.cand_bits.ensure(kCandPerRound / 32);
    gen_candidates_kernel<<<kCandPerRound / 128, 128, 0, s>>>(c.pos.p, c.nrm.p, cur_order, m, nlevels, nthresh, round_seed, cand, d_nvalid, valid_bits);
    if (!pool.empty()) {   // carried candidates occupy the first slots
      float4 *hp = rs.pool_host.ensure(kPoolSize);
      for (size_t q = 0; q < pool.size(); ++q) hp[q] = pool[q].pl;
      PLADE_CUDA(cudaMemcpyAsync(cand, hp, sizeof(float4) * pool.size(), cudaMemcpyHostToDevice, s));
    }
    // Stage 1 scores the live slots only (12 % of the draws hold a plane): the same counts in the same places as scoring
    // all kCandPerRound slots, on an eighth of the blocks, one candidate per thread instead of two (score_live = 0: all slots)
    // Stage 2 (score_live = 2) is the same kernel, one of the 256 selected candidates per thread and one 512-point tile of
    // the large subsample per block, instead of score_points_kernel (one point per thread, a ballot per candidate)
    int *live = params.score_live ? rs.cand_live.ensure(kCandPerRound) : nullptr;
    gather_sub_kernel<<<div_up(S1 + S, 256) + (live ? 1 : 0), 256, 0, s>>>(c.pos.p, c.nrm.p, cur_order, m, S1, round_seed ^ 0x5bd1e995ull, sub1, S,
                                                                          round_seed, sub, valid_bits, (int) pool.size(), live, d_nvalid + 1);
    {
      const int n_tiles = div_up(S1, kScoreTile);
      dev.clock.begin(KernelClock::kScoreCandidates, 28.0 * S1, s);     // 28 B per point per pass
      if (live)
        score_candidates_kernel<1><<<dim3(n_tiles, kCandPerRound / kScoreThreads), kScoreThreads, 0, s>>>(sub1, S1, cand, live, d_nvalid + 1, 0, kCandPerRound,
                                                                                                         eps, nthresh, 1, counts);
      else
        score_candidates_kernel<kScoreC1><<<dim3(n_tiles, kCandPerRound / (kScoreThreads * kScoreC1)), kScoreThreads, 0, s>>>(sub1, S1, cand, nullptr, nullptr, 0,
                                                                                                                             kCandPerRound, eps, nthresh, 1, counts);
      dev.clock.end(s);
    }
    select_top_kernel<<<1, kSelThreads, 0, s>>>(counts, kCandPerRound, (int) pool.size(), cidx_sorted, cand, cand_top);
    dev.clock.begin(KernelClock::kScoreCandidates, 28.0 * S, s);
    static_assert(kStage2Cand == kScoreThreads, "stage 2 by candidate: one block row");
    if (params.score_live >= 2)
      score_candidates_kernel<1><<<dim3(div_up(S, kScoreTile), 1), kScoreThreads, 0, s>>>(sub, S, cand, cidx_sorted, nullptr, 1, kStage2Cand, eps, nthresh, 1, counts2);
    else
      score_points_kernel<<<div_up(S, 256), 256, 0, s>>>(sub, S, cand, cidx_sorted, eps, nthresh, counts2);
    dev.clock.end(s);
    PLADE_LAUNCH_CHECK();
    dev.launches.add(5);
    // one copy into page-locked memory (a copy into pageable memory would block the host once per call)
    PLADE_CUDA(cudaMemcpyAsync(h_round, counts2, sizeof(unsigned int) * (kRoundEnd - kRoundCounts2), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    const unsigned int *h_counts = h_round;
    const int n_valid = (int) h_round[kRoundNValid - kRoundCounts2];
    const float4 *h_cand = reinterpret_cast<const float4 *>(h_round + (kRoundTop - kRoundCounts2));
    mark("ransac_score_round");
    drawn += n_valid;
    // Candidate pool (the reference keeps its candidate list across iterations, RansacShapeDetector.cpp:
    // 548-855): the best distinct planes of this round are re-scored next round in the first slots, so a
    // plane that lost against a bigger one is still there after the bigger one has been removed.
    {
      std::vector<int> ord(kStage2Cand);
      for (int k = 0; k < kStage2Cand; ++k) ord[k] = k;
      std::sort(ord.begin(), ord.end(), [&](int a, int b) { return h_counts[a] != h_counts[b] ? h_counts[a] > h_counts[b] : a < b; });
      pool.clear();
      for (int q = 0; q < kStage2Cand && (int) pool.size() < kPoolSize; ++q) {
        const int k = ord[q];
        if (h_counts[k] == 0) break;
        const float4 &a = h_cand[k];
        bool dup = false;
        for (const PoolEntry &e : pool) {
          float dn = a.x * e.pl.x + a.y * e.pl.y + a.z * e.pl.z;
          float dd = dn >= 0 ? std::fabs(a.w - e.pl.w) : std::fabs(a.w + e.pl.w);
          if (std::fabs(dn) > 0.995f && dd < 2 * eps) { dup = true; break; }
        }
        for (const float4 &e : banned) {
          if (dup) break;
          float dn = a.x * e.x + a.y * e.y + a.z * e.z;
          float dd = dn >= 0 ? std::fabs(a.w - e.w) : std::fabs(a.w + e.w);
          if (std::fabs(dn) > 0.995f && dd < 2 * eps) dup = true;
        }
        if (!dup) pool.push_back({(double) h_counts[k] * m / S, a});
      }
    }
    mark("ransac_pool");
    // Walk the pool best-first.  After an accepted plane the remaining entries keep their (now slightly
    // stale, never too small) estimates: distinct planes share almost no points, and every candidate is
    // fully re-evaluated on the current cloud before it is accepted, so no re-scoring round is needed.
    bool stop_all = false, fresh = true;
    for (;;) {
    best_est = pool.empty() ? 0.0 : pool[0].est;
    if (!pool.empty()) best_pl = pool[0].pl;
    const bool enough_for_min = failure_probability(min_support, m, drawn, nlevels) <= prob;
    const bool best_ok = best_est >= min_support && failure_probability(best_est, m, drawn, nlevels) <= prob;
    if (!best_ok) {
      // nothing of min_support size left (w.h.p.): require the evidence in three consecutive scoring rounds
      if (fresh) { if (enough_for_min && best_est < min_support) { if (++dry_rounds >= 3) stop_all = true; } else dry_rounds = 0; }
      break;
    }
    fresh = false;
    dry_rounds = 0;
    // ---- the pool walk on the device: accept_loop_kernel takes the eligible pool entries in order (band, acceptance chain,
    // decision, removal of the accepted points), ONE launch and one host round trip for all of them.  The one-candidate
    // path below remains for the candidate the kernel hands back (bitmap beyond the device cap, refit that left its
    // band), for ransac_batch = 1 and for devices without the cluster kernel.
    {
      const int csize = refine_cluster_size(dev.id);
      int K = 0;
      if (csize && !force_single && params.ransac_batch > 1) {
        const int kmax = std::min<int>({kAcceptMax, params.ransac_batch, (int) pool.size()});
        K = 1;       // entry 0 is eligible (best_ok); the kernel re-tests the others with the state it has by then
        while (K < kmax && pool[K].est >= min_support) ++K;
      }
      force_single = false;
      if (K >= 1) {
        unsigned char *bm = rs.batch_mem.ensure(kAccBytes);
        unsigned char *bh = rs.batch_host.ensure(2 * kAccBytes);
        memset(bh, 0, kAccBytes);
        BatchState st0;
        st0.m = m; st0.drawn = drawn; st0.prob = prob; st0.n_found = (int) found.size(); st0.stop = 0; st0.nlevels = nlevels; st0.min_support = min_support;
        memcpy(bh + kAccState, &st0, sizeof(st0));
        for (int j = 0; j < K; ++j) { PoolCand pc; pc.pl = pool[j].pl; pc.est = pool[j].est; pc.pad = 0; memcpy(bh + kAccPool + j * sizeof(PoolCand), &pc, sizeof(pc)); }
        PLADE_CUDA(cudaMemcpyAsync(bm, bh, kAccBytes, cudaMemcpyHostToDevice, s));
        AcceptArgs aa;
        RefineArgs &ra = aa.r;
        ra.seg_cap = band_seg_cap(n, csize); ra.segs = csize;
        ra.posB = posB; ra.nrmB = nrmB; ra.idxB = idxB; ra.d_nb = d_nb;
        ra.flag = flag; ra.pix = pix;
        ra.bmp = rs.bmp_dev.ensure(kBmpCap); ra.btmp = rs.bmp_tmp.ensure(kBmpCap); ra.bmask = rs.mask_dev.ensure(kBmpCap);
        ra.lab = rs.cc_lab.ensure(kBmpCap); ra.ccnt = rs.cc_cnt.ensure(kBmpCap);
        ra.member_a = rs.memb_a.ensure(band_cap_max); ra.member_b = rs.memb_b.ensure(band_cap_max);
        ra.uvbox = d_uvbox; ra.acc = acc;
        unsigned char *rm = reinterpret_cast<unsigned char *>(rs.refine_mem.ensure(64));
        ra.ctl = reinterpret_cast<RefineCtl *>(rm); ra.out = reinterpret_cast<RefineOut *>(rm + 128);
        ra.cand_pl = pool[0].pl; ra.band_pl = pool[0].pl;
        for (int k = 0; k < 3; ++k) ra.cand_pos[k] = 0.f;
        ra.eps3 = eps3; ra.nthresh = nthresh; ra.bmp_eps = bmp_eps; ra.band_halfwidth = kBandMul * eps3; ra.ext = ext;
        for (int k = 0; k < 3; ++k) { ra.mn[k] = mn[k]; ra.mx[k] = mx[k]; }
        ra.min_support = min_support; ra.band_full = 0;
        if (!rs.bmp_dev_clean) { PLADE_CUDA(cudaMemsetAsync(ra.bmp, 0, kBmpCap, s)); rs.bmp_dev_clean = true; }
        aa.pos = c.pos.p; aa.nrm = c.nrm.p; aa.assigned = assigned; aa.n = n; aa.k = K;
        aa.pool = reinterpret_cast<const PoolCand *>(bm + kAccPool);
        aa.state = reinterpret_cast<BatchState *>(bm + kAccState);
        aa.verdict = reinterpret_cast<Verdict *>(bm + kAccVerdict);
        aa.actl = reinterpret_cast<AcceptCtl *>(bm + kAccCtl);
        aa.band = kBandMul * eps3;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(csize); cfg.blockDim = dim3(kRefThreads); cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        dev.clock.begin(KernelClock::kRefineCluster, 0.0, s);
        PLADE_CUDA(cudaLaunchKernelEx(&cfg, accept_loop_kernel, aa));
        dev.clock.end(s);
        dev.launches.add();
        unsigned char *back = bh + kAccBytes;
        PLADE_CUDA(cudaMemcpyAsync(back, bm, kAccBytes, cudaMemcpyDeviceToHost, s));
        stream_sync(s);
        BatchState st1;
        memcpy(&st1, back + kAccState, sizeof(st1));
        const int m_before = m;
        int n_acc = 0;
        bool walk_on = true;
        for (int j = 0; j < K && walk_on; ++j) {
          Verdict v;
          memcpy(&v, back + kAccVerdict + j * sizeof(Verdict), sizeof(v));
          if (v.kind == kVerdictAccepted || v.kind == kVerdictRejected || v.kind == kVerdictFallback) {
            // algorithmic bytes: 20 B per point for the band pass + 28 B per band point per evaluation
            if (dev.clock.enabled) dev.clock.bytes[KernelClock::kRefineCluster] += 20.0 * n + 28.0 * v.n_band * v.evals;
            for (int k = 0; k < 6; ++k) refine_phase_ns[k] += v.phase_ns[k];
            band_phase_ns += v.band_ns;
            refine_evals += v.evals;
            ++n_band_builds;
          }
          switch (v.kind) {
            case kVerdictAccepted: {
              FoundPlane fp;
              memcpy(fp.n, v.n, sizeof(fp.n)); memcpy(fp.pos, v.p, sizeof(fp.pos));
              fp.size = v.size;
              found.push_back(fp);
              pool.erase(pool.begin());
              ++n_acc;
              break;
            }
            case kVerdictRejected:
              // its score can only shrink from here on: never look at this plane (or a duplicate of it) again
              banned.push_back(pool[0].pl);
              pool.erase(pool.begin());
              if (++rejects >= 256) { stop_all = true; walk_on = false; }
              break;
            case kVerdictFallback: force_single = true; ++n_cluster_fallbacks; walk_on = false; break;    // pool[0] goes through the host-driven path
            default: walk_on = false; break;                                                                // not eligible any more / not reached
          }
        }
        ++n_batches;
        n_batch_cands += K;
        m = (int) st1.m;
        drawn = st1.drawn;
        if (n_acc) {
          // compact the Morton list to the still-unassigned points, once per launch
          IsUnassigned pr{assigned};
          size_t tb3 = 0;
          cub::DeviceSelect::If(nullptr, tb3, cur_order, alt_order, d_nsel, m_before, pr, s);
          tmp = rs.cub_tmp.ensure(tb3);
          cub::DeviceSelect::If(tmp, tb3, cur_order, alt_order, d_nsel, m_before, pr, s);
          dev.launches.add(3);
          std::swap(cur_order, alt_order);
        }
        mark("ransac_accept_loop");
        if (stop_all) break;
        if (m < min_support || m < 3) break;
        continue;
      }
    }
    // --- refine the best candidate on the full cloud -------------------------------------------------------
    PlaneFrame fr;
    auto make_frame = [&](const float nrm3[3], const float pos3[3]) { return make_plane_frame(nrm3, pos3); };
    auto evaluate = [&](const PlaneFrame &f, unsigned char *member) -> Eval {
      Eval e{0, 0, {0, 0, 0}, false};
      int init[4];
      { float big = 3.4e38f, nb = -3.4e38f; int a, b; memcpy(&a, &big, 4); memcpy(&b, &nb, 4); b ^= 0x7fffffff; init[0] = init[1] = a; init[2] = init[3] = b; }
      PLADE_CUDA(cudaMemcpyAsync(d_uvbox, init, sizeof(init), cudaMemcpyHostToDevice, s));
      flag_uv_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, c.nrm.p, assigned, n, nullptr, f, eps3, nthresh, flag, d_uvbox);
      PLADE_LAUNCH_CHECK();
      dev.launches.add();
      int hb[4];
      PLADE_CUDA(cudaMemcpyAsync(hb, d_uvbox, sizeof(hb), cudaMemcpyDeviceToHost, s));
      stream_sync(s);
      float uv[4];
      for (int k = 0; k < 4; ++k) { int a = hb[k]; a = a >= 0 ? a : a ^ 0x7fffffff; memcpy(&uv[k], &a, 4); }
      if (uv[0] > uv[2]) return e;     // no inliers
      // BitmapExtent (R/PlanePrimitiveShape.cpp:192-198)
      size_t ue = (size_t) std::ceil((uv[2] - uv[0]) / bmp_eps) + 1, ve = (size_t) std::ceil((uv[3] - uv[1]) / bmp_eps) + 1;
      if (ue < 2) ue = 2;
      if (ve < 2) ve = 2;
      if (ue * ve > (size_t) 1 << 26) return e;
      unsigned char *bitmap = rs.bitmap.ensure(ue * ve), *mask = rs.mask.ensure(ue * ve);
      PLADE_CUDA(cudaMemsetAsync(bitmap, 0, ue * ve, s));
      raster_kernel<<<div_up(n, 256), 256, 0, s>>>(c.pos.p, flag, n, f, uv[0], uv[1], bmp_eps, (int) ue, (int) ve, pix, bitmap);
      PLADE_LAUNCH_CHECK();
      std::vector<unsigned char> hbmp(ue * ve), hmask;
      PLADE_CUDA(cudaMemcpyAsync(hbmp.data(), bitmap, ue * ve, cudaMemcpyDeviceToHost, s));
      stream_sync(s);
      largest_component(hbmp, (int) ue, (int) ve, hmask);
      PLADE_CUDA(cudaMemcpyAsync(mask, hmask.data(), ue * ve, cudaMemcpyHostToDevice, s));
      PLADE_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 16, s));
      select_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, flag, pix, mask, n, f.pl, eps3, member, acc);
      PLADE_LAUNCH_CHECK();
      dev.launches.add(2);
      double h[5];
      PLADE_CUDA(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, s));
      stream_sync(s);
      e.size = (long long) h[0]; e.sum[0] = h[1]; e.sum[1] = h[2]; e.sum[2] = h[3]; e.score = h[4]; e.ok = e.size > 0;
      return e;
    };
    // clears both membership maps over the current band; with shape_id >= 0 first assigns the members of `acc`
    auto band_finish = [&](int shape_id, unsigned char *acc_map) {
      band_finish_kernel<<<blocks_b, 256, 0, s>>>(idxB, d_nb, acc_map ? acc_map : member_a, acc_map == member_a ? member_b : member_a,
                                                 acc_map ? shape_id : -1, assigned);
      PLADE_LAUNCH_CHECK();
      dev.launches.add();
    };
    // ---- band of this candidate (see band_compact_kernel) -------------------------------------------------------
    float4 band_pl = make_float4(0, 0, 0, 0);
    bool band_full = false;
    int band_segs = 1, band_cap_seg = n;
    auto build_band = [&](const float4 &pl, bool full, int segs = 1) {      // segs > 1: the layout of the cluster kernel (RefineArgs)
      band_segs = segs; band_cap_seg = segs > 1 ? band_seg_cap(n, segs) : n;
      posB = rs.band_pos.ensure((size_t) band_cap_seg * segs); nrmB = rs.band_nrm.ensure((size_t) band_cap_seg * segs); idxB = rs.band_idx.ensure((size_t) band_cap_seg * segs);
      PLADE_CUDA(cudaMemsetAsync(d_nb, 0, sizeof(int) * 16, s));
      dev.clock.begin(KernelClock::kBandCompact, 20.0 * n, s);       // assigned (4 B) + position (16 B) of every point, once
      band_compact_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, c.nrm.p, assigned, n, pl, full ? INFINITY : kBandMul * eps3, posB, nrmB, idxB, d_nb, band_cap_seg, segs);
      dev.clock.end(s);
      PLADE_LAUNCH_CHECK();
      dev.launches.add();
      band_pl = pl;
      band_full = full;
      ++n_band_builds;
    };
    auto band_covers = [&](const float4 &pl) -> bool {
      return band_full || band_covers_plane(band_pl, kBandMul * eps3, pl, mn, mx, ext, eps3);
    };
    // fused evaluation on the band: flag/bbox -> raster -> closing + components -> select -> mean -> covariance,
    // one sync.  cov6 receives the covariance sums about float(mean) of the selected members (input of the LS
    // refit); member is a map over ALL points (zero outside the band).
    auto evaluate_fused = [&](const PlaneFrame &f, unsigned char *member, double cov6[6], bool &overflow) -> Eval {
      Eval e{0, 0, {0, 0, 0}, false};
      unsigned char *bmp = rs.bmp_dev.ensure(kBmpCap), *btmp = rs.bmp_tmp.ensure(kBmpCap), *bmask = rs.mask_dev.ensure(kBmpCap);
      int *lab = rs.cc_lab.ensure(kBmpCap), *ccnt = rs.cc_cnt.ensure(kBmpCap);
      BmpInfo *info = reinterpret_cast<BmpInfo *>(d_misc + 32);
      if (!rs.bmp_dev_clean) { PLADE_CUDA(cudaMemsetAsync(bmp, 0, kBmpCap, s)); rs.bmp_dev_clean = true; }
      // widening keeps both membership maps: they are indexed by point, and the full band is a superset of the old one
      if (!band_covers(f.pl)) { build_band(f.pl, true); ++n_band_full; }
      int init[4];
      { float big = 3.4e38f, nb = -3.4e38f; int a, b; memcpy(&a, &big, 4); memcpy(&b, &nb, 4); b ^= 0x7fffffff; init[0] = init[1] = a; init[2] = init[3] = b; }
      PLADE_CUDA(cudaMemcpyAsync(d_uvbox, init, sizeof(init), cudaMemcpyHostToDevice, s));
      PLADE_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 16, s));
      flag_uv_kernel<<<blocks_b, 256, 0, s>>>(posB, nrmB, nullptr, 0, d_nb, f, eps3, nthresh, flag, d_uvbox);
      raster_dev_kernel<<<blocks_b, 256, 0, s>>>(posB, flag, d_nb, f, bmp_eps, d_uvbox, info, pix, bmp);
      cc_kernel<<<1, 1024, 0, s>>>(bmp, btmp, lab, ccnt, info, bmask);
      select_dev_kernel<<<blocks_b, 256, 0, s>>>(posB, idxB, flag, pix, bmask, info, d_nb, f.pl, eps3, member, acc);
      cov_dev_kernel<<<blocks_b, 256, 0, s>>>(posB, idxB, member, d_nb, acc, acc + 8);
      PLADE_LAUNCH_CHECK();
      dev.launches.add(5);
      double h[16];
      BmpInfo hi;
      PLADE_CUDA(cudaMemcpyAsync(h, acc, sizeof(h), cudaMemcpyDeviceToHost, s));
      PLADE_CUDA(cudaMemcpyAsync(&hi, info, sizeof(hi), cudaMemcpyDeviceToHost, s));
      stream_sync(s);
      overflow = hi.overflow != 0;
      e.size = (long long) h[0]; e.sum[0] = h[1]; e.sum[1] = h[2]; e.sum[2] = h[3]; e.score = h[4]; e.ok = e.size > 0;
      for (int k = 0; k < 6; ++k) cov6[k] = h[8 + k];
      return e;
    };
    auto fit_from_cov = [&](const Eval &e, const double h[6], float nrm3[3], float pos3[3]) -> bool { return fit_plane_from_cov(e, h, nrm3, pos3); };
    auto ls_fit = [&](const Eval &e, const unsigned char *member, float nrm3[3], float pos3[3]) -> bool {
      // Plane::LeastSquaresFit (R/Plane.h:66-74): mean, covariance, eigenvector of smallest |lambda|
      float3 mean = make_float3((float) (e.sum[0] / e.size), (float) (e.sum[1] / e.size), (float) (e.sum[2] / e.size));
      PLADE_CUDA(cudaMemsetAsync(acc + 8, 0, sizeof(double) * 6, s));
      cov_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, member, n, mean, acc + 8);
      PLADE_LAUNCH_CHECK();
      dev.launches.add();
      double h[6];
      PLADE_CUDA(cudaMemcpyAsync(h, acc + 8, sizeof(h), cudaMemcpyDeviceToHost, s));
      stream_sync(s);
      float a[3][3], d[3], v[3][3];
      a[0][0] = (float) (h[0] / e.size); a[0][1] = a[1][0] = (float) (h[1] / e.size); a[0][2] = a[2][0] = (float) (h[2] / e.size);
      a[1][1] = (float) (h[3] / e.size); a[1][2] = a[2][1] = (float) (h[4] / e.size); a[2][2] = (float) (h[5] / e.size);
      if (!jacobi3f(a, d, v)) return false;
      int k = 0;
      for (int j = 1; j < 3; ++j) if (std::fabs(d[j]) < std::fabs(d[k])) k = j;
      nrm3[0] = v[0][k]; nrm3[1] = v[1][k]; nrm3[2] = v[2][k];
      pos3[0] = mean.x; pos3[1] = mean.y; pos3[2] = mean.z;
      return true;
    };

    float cn[3] = {best_pl.x, best_pl.y, best_pl.z};
    // position of the 3-point plane: any point on it (the reference keeps the first sample); n * dist lies on the plane
    float cp[3] = {best_pl.x * best_pl.w, best_pl.y * best_pl.w, best_pl.z * best_pl.w};
    fr = make_frame(cn, cp);
    fr.pl.w = best_pl.w;
    unsigned char *acc_member = member_a, *work_member = member_b;
    float acc_n[3] = {cn[0], cn[1], cn[2]}, acc_p[3] = {cp[0], cp[1], cp[2]};
    long long acc_size = 0;
    bool host_path = false, refined = false;
    const unsigned char *band_member = nullptr;      // members of the accepted plane by band position (cluster path)
    const int csize = refine_cluster_size(dev.id);
    build_band(fr.pl, false, csize ? csize : 1);
    // ---- the whole acceptance chain in one cluster kernel, one host round trip (refine_cluster_kernel) -----------
    if (csize) {
      unsigned char *rm = reinterpret_cast<unsigned char *>(rs.refine_mem.ensure(64));
      RefineArgs ra;
      ra.posB = posB; ra.nrmB = nrmB; ra.idxB = idxB; ra.d_nb = d_nb;
      ra.seg_cap = band_cap_seg; ra.segs = csize;
      for (int k = 0; k < 3; ++k) ra.cand_pos[k] = cp[k];
      ra.flag = flag; ra.pix = pix;
      ra.bmp = rs.bmp_dev.ensure(kBmpCap); ra.btmp = rs.bmp_tmp.ensure(kBmpCap); ra.bmask = rs.mask_dev.ensure(kBmpCap);
      ra.lab = rs.cc_lab.ensure(kBmpCap); ra.ccnt = rs.cc_cnt.ensure(kBmpCap);
      ra.member_a = rs.memb_a.ensure(band_cap_max); ra.member_b = rs.memb_b.ensure(band_cap_max);
      ra.uvbox = d_uvbox; ra.acc = acc;
      ra.ctl = reinterpret_cast<RefineCtl *>(rm); ra.out = reinterpret_cast<RefineOut *>(rm + 128);
      ra.cand_pl = best_pl; ra.band_pl = band_pl;
      ra.eps3 = eps3; ra.nthresh = nthresh; ra.bmp_eps = bmp_eps; ra.band_halfwidth = kBandMul * eps3; ra.ext = ext;
      for (int k = 0; k < 3; ++k) { ra.mn[k] = mn[k]; ra.mx[k] = mx[k]; }
      ra.min_support = min_support; ra.band_full = band_full ? 1 : 0;
      if (!rs.bmp_dev_clean) { PLADE_CUDA(cudaMemsetAsync(ra.bmp, 0, kBmpCap, s)); rs.bmp_dev_clean = true; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(csize); cfg.blockDim = dim3(refine_block_threads()); cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      dev.clock.begin(KernelClock::kRefineCluster, 0.0, s);
      PLADE_CUDA(cudaLaunchKernelEx(&cfg, refine_cluster_kernel, ra));
      dev.clock.end(s);
      dev.launches.add();
      static_assert(sizeof(RefineOut) <= 64 * sizeof(unsigned int) && ((kRoundEnd - kRoundCounts2) * 4) % 8 == 0, "verdict slot of round_host");
      RefineOut *h_ro = reinterpret_cast<RefineOut *>(rs.round_host.ensure(kRoundEnd - kRoundCounts2 + 64) + (kRoundEnd - kRoundCounts2));
      PLADE_CUDA(cudaMemcpyAsync(h_ro, ra.out, sizeof(RefineOut), cudaMemcpyDeviceToHost, s));     // page-locked: no implicit host block
      stream_sync(s);
      const RefineOut ro = *h_ro;
      if (dev.clock.enabled) dev.clock.bytes[KernelClock::kRefineCluster] += 28.0 * ro.n_band * ro.evals;     // 28 B per point per pass
      for (int k = 0; k < 6; ++k) refine_phase_ns[k] += ro.phase_ns[k];
      refine_evals += ro.evals;
      if (ro.status == 0) {
        refined = true;
        acc_size = ro.acc_size;
        memcpy(acc_n, ro.acc_n, sizeof(acc_n)); memcpy(acc_p, ro.acc_p, sizeof(acc_p));
        band_member = ro.acc_sel ? ra.member_b : ra.member_a;
      } else {
        ++n_cluster_fallbacks;         // (the cluster kernel never touches the per-point maps: they are still all zero)
      }
    }
    double cov_cur[6];
    bool ovf = false;
    Eval cur{0, 0, {0, 0, 0}, false};
    if (!refined) {
      build_band(fr.pl, false);
      cur = evaluate_fused(fr, acc_member, cov_cur, ovf);      // GlobalScore + ConnectedComponent (+ weighted score of the clone)
      host_path = ovf;                                         // bitmap larger than the device cap: host labelling
      if (host_path) cur = evaluate(fr, acc_member);
      acc_size = cur.ok ? cur.size : 0;
    }
    if (!refined && cur.ok) {
      Eval clone = cur;
      double cov_clone[6];
      memcpy(cov_clone, cov_cur, sizeof(cov_cur));
      const unsigned char *clone_member = acc_member;
      double newScore = clone.score, oldScore;
      int iter = 0;
      do {
        ++iter;
        oldScore = newScore;
        float fn[3], fp[3];
        if (host_path) { if (!ls_fit(clone, clone_member, fn, fp)) break; }
        else if (!fit_from_cov(clone, cov_clone, fn, fp)) break;
        PlaneFrame f2 = make_frame(fn, fp);
        double cov2[6];
        bool ovf2 = false;
        Eval e2 = host_path ? evaluate(f2, work_member) : evaluate_fused(f2, work_member, cov2, ovf2);
        if (!host_path && ovf2) break;
        newScore = e2.score;
        if (!e2.ok) break;
        clone = e2;
        memcpy(cov_clone, cov2, sizeof(cov2));
        clone_member = work_member;
        if (newScore > oldScore && e2.size > min_support) {
          memcpy(acc_n, fn, sizeof(fn)); memcpy(acc_p, fp, sizeof(fp));
          acc_size = e2.size;
          std::swap(acc_member, work_member);      // the clone becomes the candidate
          clone_member = acc_member;
        }
      } while (newScore > oldScore && iter < 3);
    }
    best_est = 0;    // the pool is invalid once points are removed (or the candidate failed)
    mark("ransac_refine");
    // the reference only accepts a candidate whose fully evaluated (connected-component) support
    // reaches min_support (FindBestCandidate, RansacShapeDetector.cpp:297,423-430)
    if (acc_size < min_support) {
      if (host_path) { PLADE_CUDA(cudaMemsetAsync(member_a, 0, n, s)); PLADE_CUDA(cudaMemsetAsync(member_b, 0, n, s)); }
      else if (!refined) band_finish(-1, nullptr);
      // its score can only shrink from here on: never look at this plane (or a duplicate of it) again
      banned.push_back(best_pl);
      if (!pool.empty()) pool.erase(pool.begin());
      if (++rejects >= 256) { stop_all = true; break; }
      continue;
    }
    unsigned char *member = acc_member;
    // --- accept: remove the points (RansacShapeDetector.cpp:659-675) ---------------------------------------------
    if (host_path) {
      mark_kernel<<<div_up(n, 256), 256, 0, s>>>(member, n, (int) found.size(), assigned);
      PLADE_LAUNCH_CHECK();
      PLADE_CUDA(cudaMemsetAsync(member_a, 0, n, s));
      PLADE_CUDA(cudaMemsetAsync(member_b, 0, n, s));
    } else if (refined) {
      band_assign_kernel<<<blocks_b, 256, 0, s>>>(idxB, d_nb, band_cap_seg, band_segs, band_member, (int) found.size(), assigned);
      PLADE_LAUNCH_CHECK();
      dev.launches.add();
    } else band_finish((int) found.size(), member);
    FoundPlane fp;
    memcpy(fp.n, acc_n, sizeof(acc_n)); memcpy(fp.pos, acc_p, sizeof(acc_p));
    fp.size = acc_size;
    found.push_back(fp);
    drawn = rescale_drawn(drawn, acc_size, m);
    // compact the Morton list to the still-unassigned points
    IsUnassigned pr{assigned};
    size_t tb3 = 0;
    cub::DeviceSelect::If(nullptr, tb3, cur_order, alt_order, d_nsel, m, pr, s);
    tmp = rs.cub_tmp.ensure(tb3);
    cub::DeviceSelect::If(tmp, tb3, cur_order, alt_order, d_nsel, m, pr, s);
    dev.launches.add(3);
    // the accepted members are exactly acc_size previously unassigned points: no round trip for the new count
    std::swap(cur_order, alt_order);
    m = (int) (m - acc_size);
    if (!pool.empty()) pool.erase(pool.begin());
    mark("ransac_accept");
    if (m < min_support || m < 3) break;
    }   // pool walk
    if (stop_all) break;
  }

  // what extract() needs to continue this detection with a lower min_support
  rz.cloud = (const void *) c.pos.p; rz.n = n; rz.m = m; rz.drawn = drawn; rz.round_seed = round_seed;
  rz.cur_is_order2 = cur_order == order2 ? 1 : 0;
  for (int k = 0; k < 3; ++k) { rz.mn[k] = mn[k]; rz.mx[k] = mx[k]; }
  rz.found = found;
  rz.valid = true;
  // ---- output (PLADE/plane_extraction.cpp:115-160): drop shapes below min_support, unit normal, d = -n.p --------
  // membership stays on the device: group_out[i] = index of the returned plane of point i, or -1
  std::vector<int> remap(std::max<size_t>(found.size(), 1), -1);
  for (size_t k = 0; k < found.size(); ++k) {
    if (found[k].size < (long long) min_support) continue;
    remap[k] = (int) result.size();
    PlaneParam pr;
    V3 nn(found[k].n[0], found[k].n[1], found[k].n[2]);
    float l = std::sqrt(nn.x * nn.x + nn.y * nn.y + nn.z * nn.z);     // Vec3f::normalize
    if (l > 0) { nn.x /= l; nn.y /= l; nn.z /= l; }
    pr.n[0] = nn.x; pr.n[1] = nn.y; pr.n[2] = nn.z;
    pr.d = -(nn.x * found[k].pos[0] + nn.y * found[k].pos[1] + nn.z * found[k].pos[2]);
    pr.size = found[k].size;
    result.push_back(pr);
  }
  int *d_remap = rs.order_alt.p == cur_order ? rs.order.p : rs.order_alt.p;   // the inactive order buffer is free now
  PLADE_CUDA(cudaMemcpyAsync(d_remap, remap.data(), sizeof(int) * remap.size(), cudaMemcpyHostToDevice, s));
  int *g = group_out.ensure(n);
  remap_group_kernel<<<div_up(n, 256), 256, 0, s>>>(assigned, n, d_remap, (int) found.size(), g);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  stream_sync(s);
  dev.clock.collect();
  mark("ransac_output");
  if (getenv("PLADE_TIMING")) fprintf(stderr, "[plade ransac lane %d] planes %zu, accept-loop launches %d (%d pool entries offered), band pass %.0f us; candidates evaluated on a band: %d, widened to all points: %d, cluster-kernel fallbacks: %d; cluster kernel: %d evaluations, us per phase: flags %.0f raster %.0f components %.0f select %.0f covariance %.0f decision %.0f\n",
                                   lane, found.size(), n_batches, n_batch_cands, band_phase_ns / 1e3, n_band_builds - n_band_full, n_band_full, n_cluster_fallbacks, refine_evals, refine_phase_ns[0] / 1e3, refine_phase_ns[1] / 1e3,
                                   refine_phase_ns[2] / 1e3, refine_phase_ns[3] / 1e3, refine_phase_ns[4] / 1e3, refine_phase_ns[5] / 1e3);
  return result;
}

// extract(), PLADE/plade.cpp:602-635.
// One documented, parameter-gated deviation: the reference stops halving the support as soon as its
// detector RETURNS min_planes planes.  Its detector finds a plane whose support is below ~1.25 x min_support in a minority
// of the runs and only planes from ~1.6-2 x min_support with certainty (measured detection frequency over seeds,
// tests/golden/seed_sweep_ref.json), so its loop effectively runs until enough planes are comfortably above the threshold.
// The detector here finds every plane >= min_support; counting all of them stops one halving earlier than the reference does
// on scans that hold exactly min_planes planes near the threshold -- its own room pair, where the 180-degree-symmetric
// hypothesis then wins.  Planes count towards min_planes when their support is >= detect_margin x the support of that pass
// (default 2: the planes the reference finds with certainty); ALL planes >= min_support are returned.  detect_margin = 1 is
// the literal rule.
std::vector<PlaneParam> Registrar::extract_planes_dev(const CloudDev &c, int init_min_support, DevBuf<int> &group_out, int lane) {
  Device &dev = lane == 0 ? this->dev : this->dev2;
  const int min_num = params.min_planes, max_num = params.max_planes, min_allowed_support = params.min_allowed_support;
  auto counted = [&](const std::vector<PlaneParam> &pl, int support) {
    int k = 0;
    for (const PlaneParam &q : pl) if ((double) q.size >= params.detect_margin * support) ++k;
    return k;
  };
  std::vector<PlaneParam> planes = detect_planes_dev(c, init_min_support, group_out, lane);
  if (counted(planes, init_min_support) >= min_num && (int) planes.size() <= max_num) return planes;
  if ((int) planes.size() > max_num) {
    // the reference sorts with a (non-strict) `>=` comparator; a stable descending sort is the defined equivalent
    std::vector<int> ord(planes.size());
    for (size_t i = 0; i < ord.size(); ++i) ord[i] = (int) i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return planes[a].size > planes[b].size; });
    std::vector<int> remap(planes.size(), -1);
    std::vector<PlaneParam> kept;
    for (int r = 0; r < max_num; ++r) { remap[ord[r]] = r; kept.push_back(planes[ord[r]]); }
    int *dr = scratch_of(*this, lane).remap.ensure(remap.size());
    PLADE_CUDA(cudaMemcpyAsync(dr, remap.data(), sizeof(int) * remap.size(), cudaMemcpyHostToDevice, dev.stream));
    remap_group_kernel<<<div_up((long long) c.n, 256), 256, 0, dev.stream>>>(group_out.p, (int) c.n, dr, (int) remap.size(), group_out.p);
    PLADE_LAUNCH_CHECK();
    stream_sync(dev.stream);
    std::cout << kept.size() << " of the " << planes.size() << " extracted planes will be used for registration" << std::endl;
    return kept;
  }
  const int max_trials = params.max_trials;
  int min_support = init_min_support / 2;
  int trials = 1;
  int used_support = init_min_support;
  while (counted(planes, used_support) < min_num && trials < max_trials && min_support >= min_allowed_support) {
    planes = detect_planes_dev(c, min_support, group_out, lane, params.detect_resume != 0);
    used_support = min_support;
    min_support /= 2;
    ++trials;
  }
  if (trials > 1)
    std::cout << "min_support = " << min_support << " used for extracting the " << planes.size() << " planes from point cloud" << std::endl;
  return planes;
}

// host view of a device-resident plane set (stage API): index lists in ascending point order
std::vector<PlaneRec> Registrar::planes_to_host(const CloudDev &c, const std::vector<PlaneParam> &pp, const DevBuf<int> &group) {
  std::vector<int> h(c.n);
  if (c.n) PLADE_CUDA(cudaMemcpyAsync(h.data(), group.p, sizeof(int) * c.n, cudaMemcpyDeviceToHost, dev.stream));
  stream_sync(dev.stream);
  std::vector<PlaneRec> out(pp.size());
  for (size_t k = 0; k < pp.size(); ++k) {
    out[k].idx.reserve((size_t) pp[k].size);
    for (int a = 0; a < 3; ++a) out[k].n[a] = pp[k].n[a];
    out[k].d = pp[k].d;
  }
  for (size_t i = 0; i < c.n; ++i) if (h[i] >= 0 && (size_t) h[i] < out.size()) out[h[i]].idx.push_back((int) i);
  return out;
}

std::vector<PlaneRec> Registrar::detect_planes(const CloudDev &c, int min_support) {
  DevBuf<int> &g = stage_group;
  std::vector<PlaneParam> pp = detect_planes_dev(c, min_support, g);
  return planes_to_host(c, pp, g);
}

std::vector<PlaneRec> Registrar::extract_planes(const CloudDev &c, int init_min_support) {
  DevBuf<int> &g = stage_group;
  std::vector<PlaneParam> pp = extract_planes_dev(c, init_min_support, g);
  return planes_to_host(c, pp, g);
}

// stage API: the acceptance chain of ONE candidate plane (RansacShapeDetector.cpp:613-655) on the device -- band_compact_kernel
// + refine_cluster_kernel exactly as the detection runs them -- for the parity test against the reference's own Candidate /
// BitmapPrimitiveShape / Plane classes.  assigned (may be null): points with assigned[i] != -1 are not available.
bool Registrar::refine_candidate_stage(const CloudDev &c, const int *h_assigned, const float nrm3[3], const float pos3[3], int min_support,
                                       float out_n[3], float out_p[3], unsigned char *h_member, long long *out_size, int *out_evals, double *out_score) {
  const int n = (int) c.n;
  cudaStream_t s = dev.stream;
  RansacScratch &rs = scratch_of(*this, 0);
  const int csize = refine_cluster_size(dev.id);
  if (!csize) throw std::runtime_error("refine_candidate_stage: the cluster kernel is not available on this device");
  if (n < 3) return false;
  const int blocks_n = std::min(div_up(n, 256), dev.num_sms * 8);
  int *d_misc = rs.misc.ensure(128);
  {
    int init[6];
    float big = 3.4e38f, nbig = -3.4e38f;
    int bi, nbi;
    memcpy(&bi, &big, 4); memcpy(&nbi, &nbig, 4);
    nbi ^= 0x7fffffff;
    init[0] = init[1] = init[2] = bi; init[3] = init[4] = init[5] = nbi;
    PLADE_CUDA(cudaMemcpyAsync(d_misc, init, sizeof(init), cudaMemcpyHostToDevice, s));
  }
  bbox_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, n, d_misc);
  PLADE_LAUNCH_CHECK();
  int h6[6];
  PLADE_CUDA(cudaMemcpyAsync(h6, d_misc, sizeof(h6), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  float mn[3], mx[3];
  for (int k = 0; k < 3; ++k) {
    int a = h6[k], b = h6[3 + k];
    a = a >= 0 ? a : a ^ 0x7fffffff; b = b >= 0 ? b : b ^ 0x7fffffff;
    memcpy(&mn[k], &a, 4); memcpy(&mx[k], &b, 4);
  }
  const float scale = std::max(mx[0] - mn[0], mx[1] - mn[1]);       // (bug-compatible: z ignored)
  const float eps3 = 3 * params.ransac_dist_thresh * scale, bmp_eps = params.ransac_bitmap_reso * scale;
  const float ext = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
  int *assigned = rs.assigned.ensure(n);
  if (h_assigned) PLADE_CUDA(cudaMemcpyAsync(assigned, h_assigned, sizeof(int) * n, cudaMemcpyHostToDevice, s));
  else PLADE_CUDA(cudaMemsetAsync(assigned, 0xff, sizeof(int) * n, s));
  const int seg_cap = band_seg_cap(n, csize);
  const size_t cap = (size_t) seg_cap * csize;
  RefineArgs ra;
  ra.posB = rs.band_pos.ensure(cap); ra.nrmB = rs.band_nrm.ensure(cap); ra.idxB = rs.band_idx.ensure(cap);
  ra.d_nb = d_misc + 64; ra.seg_cap = seg_cap; ra.segs = csize;
  ra.flag = rs.flag.ensure(cap); ra.pix = rs.pix.ensure(cap);
  ra.bmp = rs.bmp_dev.ensure(kBmpCap); ra.btmp = rs.bmp_tmp.ensure(kBmpCap); ra.bmask = rs.mask_dev.ensure(kBmpCap);
  ra.lab = rs.cc_lab.ensure(kBmpCap); ra.ccnt = rs.cc_cnt.ensure(kBmpCap);
  ra.member_a = rs.memb_a.ensure(cap); ra.member_b = rs.memb_b.ensure(cap);
  ra.uvbox = d_misc + 16; ra.acc = rs.acc.ensure(16);
  unsigned char *rm = reinterpret_cast<unsigned char *>(rs.refine_mem.ensure(64));
  ra.ctl = reinterpret_cast<RefineCtl *>(rm); ra.out = reinterpret_cast<RefineOut *>(rm + 128);
  const float dist = (pos3[0] * nrm3[0] + pos3[1] * nrm3[1]) + pos3[2] * nrm3[2];       // m_pos.dot(m_normal), R/Plane.cpp:13-23
  ra.cand_pl = make_float4(nrm3[0], nrm3[1], nrm3[2], dist); ra.band_pl = ra.cand_pl;
  for (int k = 0; k < 3; ++k) { ra.cand_pos[k] = pos3[k]; ra.mn[k] = mn[k]; ra.mx[k] = mx[k]; }
  ra.eps3 = eps3; ra.nthresh = params.ransac_normal_thresh; ra.bmp_eps = bmp_eps; ra.band_halfwidth = kBandMul * eps3; ra.ext = ext;
  ra.min_support = min_support; ra.band_full = 0;
  if (!rs.bmp_dev_clean) { PLADE_CUDA(cudaMemsetAsync(ra.bmp, 0, kBmpCap, s)); rs.bmp_dev_clean = true; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize); cfg.blockDim = dim3(refine_block_threads()); cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  RefineOut ro;
  for (int attempt = 0; attempt < 2; ++attempt) {      // a refit that leaves the band is redone on all unassigned points, as the detection does
    ra.band_full = attempt;
    PLADE_CUDA(cudaMemsetAsync(ra.d_nb, 0, sizeof(int) * 16, s));
    band_compact_kernel<<<blocks_n, 256, 0, s>>>(c.pos.p, c.nrm.p, assigned, n, ra.cand_pl, attempt ? INFINITY : kBandMul * eps3, ra.posB, ra.nrmB, ra.idxB, ra.d_nb, seg_cap, csize);
    PLADE_LAUNCH_CHECK();
    PLADE_CUDA(cudaLaunchKernelEx(&cfg, refine_cluster_kernel, ra));
    dev.launches.add(3);
    PLADE_CUDA(cudaMemcpyAsync(&ro, ra.out, sizeof(ro), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    if (ro.status != 2) break;
  }
  if (ro.status != 0) { last_error = "refine_candidate_stage: candidate needs the host path (status " + std::to_string(ro.status) + ")"; return false; }
  // members: band-local map -> per-point mask (through a scratch shape index)
  int *mark = rs.order_alt.ensure(n);
  PLADE_CUDA(cudaMemsetAsync(mark, 0, sizeof(int) * n, s));
  band_assign_kernel<<<std::min(blocks_n, dev.num_sms * 4), 256, 0, s>>>(ra.idxB, ra.d_nb, seg_cap, csize, ro.acc_sel ? ra.member_b : ra.member_a, 1, mark);
  PLADE_LAUNCH_CHECK();
  std::vector<int> hm(n);
  PLADE_CUDA(cudaMemcpyAsync(hm.data(), mark, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  for (int i = 0; i < n; ++i) h_member[i] = hm[i] ? 1 : 0;
  memcpy(out_n, ro.acc_n, sizeof(float) * 3); memcpy(out_p, ro.acc_p, sizeof(float) * 3);
  *out_size = ro.acc_size; *out_evals = ro.evals; *out_score = ro.acc_score;
  return true;
}

// stage API: closing + largest 8-connected component of a ue x ve bitmap on the device (cc_kernel), as the RANSAC
// acceptance test runs it (BitmapPrimitiveShape::ConnectedComponent, R/BitmapPrimitiveShape.cpp:155-205)
void largest_component_device(Device &dev, const unsigned char *h_bitmap, int ue, int ve, unsigned char *h_mask) {
  const size_t P = (size_t) ue * ve;
  if (P == 0 || P > (size_t) kBmpCap) throw std::runtime_error("largest_component_device: bitmap size out of range");
  static thread_local DevBuf<unsigned char> bmp, tmp, mask;
  static thread_local DevBuf<int> lab, cnt, info;
  unsigned char *d_bmp = bmp.ensure(P), *d_tmp = tmp.ensure(P), *d_mask = mask.ensure(P);
  int *d_lab = lab.ensure(P), *d_cnt = cnt.ensure(P);
  BmpInfo hi{0.f, 0.f, ue, ve, 1, 0, 0, 0};
  BmpInfo *d_info = reinterpret_cast<BmpInfo *>(info.ensure(8));
  cudaStream_t s = dev.stream;
  PLADE_CUDA(cudaMemcpyAsync(d_bmp, h_bitmap, P, cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemcpyAsync(d_info, &hi, sizeof(hi), cudaMemcpyHostToDevice, s));
  cc_kernel<<<1, 1024, 0, s>>>(d_bmp, d_tmp, d_lab, d_cnt, d_info, d_mask);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  PLADE_CUDA(cudaMemcpyAsync(h_mask, d_mask, P, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
}

// stage API: plane consensus counts over the whole cloud (K1a/K1b predicate)
void score_planes_full(Device &dev, const float4 *pos, const float4 *nrm, const int *assigned, size_t n, const float4 *d_planes,
                       int n_planes, float eps, float nthresh, unsigned int *d_counts, unsigned char *d_mask0) {
  PLADE_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * n_planes, dev.stream));
  if (n == 0 || n_planes == 0) return;
  int blocks = std::min(div_up((long long) n, 256), dev.num_sms * 8);
  score_planes_full_kernel<<<blocks, 256, sizeof(float4) * n_planes, dev.stream>>>(pos, nrm, assigned, (int) n, d_planes, n_planes, eps,
                                                                                 nthresh, d_counts, d_mask0);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
}

}  // namespace plade
