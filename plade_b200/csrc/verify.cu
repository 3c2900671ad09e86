// K5 — hypothesis verification on sm_100a.
//
// Replaces the reference loop PLADE/plade.cpp:547-564: per hypothesis, transform the down-sampled
// source cloud (pcl::transformPointCloud, common/impl/transforms.hpp:64-72), restrict the target to
// the ball around the transformed source centre and count source points that have a target point
// within the inlier distance (ComputeOverlap, PLADE/util.h:612-647; FLANN L2_Simple, dist.h:84-90;
// strict '<' against float(r*r), kdtree_flann.hpp:193 / result_set.h:479,582).
//
// Design: the reference builds two kd-trees per hypothesis; here the target ds cloud is binned ONCE
// into a uniform grid (cell edge just above the inlier distance) and one persistent kernel walks
// (source tile x hypothesis chunk) work items.  A tile of source points is staged into shared
// memory with one 1-D TMA bulk copy (cp.async.bulk + mbarrier) and reused for every hypothesis of
// the chunk, so the source cloud streams from HBM/L2 once per chunk instead of once per hypothesis.
// All float arithmetic uses explicit round-to-nearest mul/add (the library is also built with
// -fmad=false) so the inlier decision is bit-identical to the scalar reference.
#include "kernels.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>

namespace plade {

namespace {

constexpr int kTile = 512;           // source points per shared-memory tile (8 KB)
constexpr int kThreads = 256;
constexpr int kHypChunk = 32;        // hypotheses per work item (the tile is reused for all of them)
constexpr int kHypGroup = 16;        // hypotheses per filter/search pass; the queue holds every (hypothesis, point) of a pass
static_assert(kTile <= 1024 && kHypChunk <= 64, "queue entries pack (hypothesis << 10 | point) into 16 bits");

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!ok);
}

// the nine (dy, dz) rows of the 3x3x3 neighbourhood, nearest first: centre, four edge-adjacent, four corners
__device__ constexpr int kRowDy[9] = {0, -1, 1, 0, 0, -1, 1, -1, 1};
__device__ constexpr int kRowDz[9] = {0, 0, 0, -1, 1, -1, -1, 1, 1};

struct GridView {
  const float4 *pts;
  const int *cell_start;
  const unsigned int *raw;     // occupancy: bit set <=> the cell holds a target point (grid extended by one empty cell per side)
  const unsigned int *dil;     // dilated occupancy: bit set <=> some target point lies in the 27 cells around this one
  int ey, ewords;
  float minx, miny, minz, inv_cell;
  int nx, ny, nz;
};

// FLANN L2_Simple: result = 0; result += d*d for x, y, z in order (no FMA contraction).
__device__ __forceinline__ float dist2_l2simple(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

// pcl::transformPointCloud: x' = ((m00*x + m01*y) + m02*z) + m03, plain float, left to right.
__device__ __forceinline__ void transform_point(const HypParams &hp, float sx, float sy, float sz, float &x, float &y, float &z) {
  x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hp.R[0], sx), __fmul_rn(hp.R[1], sy)), __fmul_rn(hp.R[2], sz)), hp.T[0]);
  y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hp.R[3], sx), __fmul_rn(hp.R[4], sy)), __fmul_rn(hp.R[5], sz)), hp.T[1]);
  z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(hp.R[6], sx), __fmul_rn(hp.R[7], sy)), __fmul_rn(hp.R[8], sz)), hp.T[2]);
}

// Filter pass: can the transformed point have a target point within the inlier distance at all?  One bit test of
// the dilated occupancy rejects every point whose 27-cell neighbourhood is empty (most points of a wrong hypothesis).
__device__ __forceinline__ bool point_may_have_inlier(const GridView &g, const HypParams &hp, float sx, float sy, float sz) {
  float x, y, z;
  transform_point(hp, sx, sy, sz, x, y, z);
  float fx = (x - g.minx) * g.inv_cell, fy = (y - g.miny) * g.inv_cell, fz = (z - g.minz) * g.inv_cell;
  // outside the grid by more than one cell (or NaN) => no target point can be within the inlier radius
  if (!(fx >= -1.0f && fy >= -1.0f && fz >= -1.0f && fx < (float) (g.nx + 1) && fy < (float) (g.ny + 1) &&
        fz < (float) (g.nz + 1)))
    return false;
  int cx = (int) floorf(fx), cy = (int) floorf(fy), cz = (int) floorf(fz);
  // (the bitmap has at most 2^26 / 32 words: 32-bit index arithmetic)
  unsigned int w = __ldg(g.dil + (unsigned int) (((cz + 1) * g.ey + (cy + 1)) * g.ewords + ((cx + 1) >> 5)));
  return (w >> ((cx + 1) & 31)) & 1u;
}

// Search pass, for a point that passed the filter: exact test against the target points of the 27 cells around it.
// Two steps so that the lanes of a warp stay together: (1) a fixed nine-iteration loop gathers the occupancy bits
// of the nine rows into one 27-bit mask (rows outside the grid, and corner rows that cannot hold a point within
// the inlier distance, contribute nothing) — the 1.2 MB bitmap answers "empty" without touching the 4-byte-per-
// cell CSR table; (2) a loop over the NON-EMPTY rows only, nearest row first (a true inlier is usually found in
// the centre row), with early exit.
__device__ __forceinline__ bool point_has_inlier(const GridView &g, const HypParams &hp, float rball2, float rin2,
                                                 float sx, float sy, float sz) {
  float x, y, z;
  transform_point(hp, sx, sy, sz, x, y, z);
  float fx = (x - g.minx) * g.inv_cell, fy = (y - g.miny) * g.inv_cell, fz = (z - g.minz) * g.inv_cell;
  int cx = (int) floorf(fx), cy = (int) floorf(fy), cz = (int) floorf(fz);
  int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.nx - 1);
  int y0 = max(cy - 1, 0), y1 = min(cy + 1, g.ny - 1);
  int z0 = max(cz - 1, 0), z1 = min(cz + 1, g.nz - 1);
  if (x0 > x1) return false;
  const unsigned int xmask = (2u << (x1 - x0)) - 1u;     // x1 - x0 + 1 bits
  const int xe0 = x0 + 1, xw = xe0 >> 5, xs = xe0 & 31;
  const bool two_words = xs + (x1 - x0) > 31;
  // a corner row (dy != 0 and dz != 0) can only hold a point within one cell edge (>= the inlier distance: the cell
  // is 2^-10 larger) of the query if the query's in-cell offsets towards it satisfy oy^2 + oz^2 < 1 (in cell
  // units; 1.01 leaves room for rounding)
  const float ry = fy - (float) cy, rz = fz - (float) cz;          // in-cell position, [0, 1)
  unsigned int m27 = 0;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int dy = kRowDy[k], dz = kRowDz[k];
    const int yy = cy + dy, zz = cz + dz;
    bool ok = yy >= y0 && yy <= y1 && zz >= z0 && zz <= z1;
    if (k >= 5) {
      const float oy = dy < 0 ? ry : 1.0f - ry, oz = dz < 0 ? rz : 1.0f - rz;
      ok = ok && !(oy * oy + oz * oz > 1.01f);
    }
    if (ok) {
      const unsigned int *orow = g.raw + (unsigned int) (((zz + 1) * g.ey + (yy + 1)) * g.ewords + xw);
      const unsigned int lo = __ldg(orow), hi = two_words ? __ldg(orow + 1) : 0u;
      m27 |= (__funnelshift_r(lo, hi, xs) & xmask) << (3 * k);
    }
  }
  while (m27) {
    const int k = (__ffs(m27) - 1) / 3;
    const unsigned int occ = (m27 >> (3 * k)) & 7u;
    m27 &= ~(7u << (3 * k));
    // dy + 1 and dz + 1 of row k, two bits each, packed (same order as kRowDy / kRowDz)
    const int dy = (int) ((0x22161u >> (2 * k)) & 3u) - 1, dz = (int) ((0x28215u >> (2 * k)) & 3u) - 1;
    const int row = ((cz + dz) * g.ny + (cy + dy)) * g.nx + x0;
    const int b = __ldg(g.cell_start + row + (__ffs(occ) - 1)), e = __ldg(g.cell_start + row + (32 - __clz(occ)));
    for (int i = b; i < e; ++i) {
      float4 t = __ldg(g.pts + i);
      if (dist2_l2simple(x, y, z, t.x, t.y, t.z) < rin2) {
        // coarse-overlap ball: target point must be inside ball(c, ball_radius)
        if (dist2_l2simple(hp.c[0], hp.c[1], hp.c[2], t.x, t.y, t.z) < rball2) return true;
      }
    }
  }
  return false;
}

// Persistent kernel over (source tile, hypothesis chunk) work items.  Per group of kHypGroup hypotheses the block
// runs two passes over the tile: a FILTER pass (transform + one occupancy bit) in which all lanes do the same
// cheap work and the survivors are appended to a shared-memory queue, then a SEARCH pass in which consecutive
// lanes take consecutive queue entries — so the expensive, data-dependent grid walk runs with full warps instead
// of the ~1 lane in 5 that survives the filter.
__global__ void __launch_bounds__(kThreads)
verify_kernel(const float4 *__restrict__ src, int ns, GridView g, const HypParams *__restrict__ hyps, int H,
              float rball2, float rin2, unsigned int *__restrict__ counts, int n_tiles, int n_chunks) {
  __shared__ __align__(128) float4 tile[kTile];
  __shared__ __align__(16) HypParams hp_s[kHypChunk];
  __shared__ unsigned short queue[kHypGroup * kTile];
  __shared__ unsigned int cnt_s[kHypChunk];
  __shared__ unsigned int q_n[2];
  __shared__ __align__(8) uint64_t bar;

  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0;
  const int n_work = n_tiles * n_chunks;
  for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
    const int chunk = w / n_tiles, t = w - chunk * n_tiles;
    const int p0 = t * kTile;
    const int np = min(kTile, ns - p0);
    const int h0 = chunk * kHypChunk;
    const int nh = min(kHypChunk, H - h0);
    if (tid == 0) {
      // order prior generic-proxy reads of `tile` before the async-proxy overwrite
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar, (uint32_t) np * 16u);
      tma_load_1d(tile, src + p0, (uint32_t) np * 16u, &bar);
      q_n[0] = 0;
    }
    for (int i = tid; i < nh * 16; i += kThreads)
      reinterpret_cast<float *>(hp_s)[i] = __ldg(reinterpret_cast<const float *>(hyps + h0) + i);
    if (tid < kHypChunk) cnt_s[tid] = 0;
    mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();
    for (int hg = 0, par = 0; hg < nh; hg += kHypGroup, par ^= 1) {
      const int hg_end = min(hg + kHypGroup, nh);
      // ---- filter pass
      for (int h = hg; h < hg_end; ++h) {
        const HypParams hp = hp_s[h];
        for (int i0 = 0; i0 < np; i0 += kThreads) {
          const int i = i0 + tid;
          bool pass = false;
          if (i < np) {
            const float4 s = tile[i];
            pass = point_may_have_inlier(g, hp, s.x, s.y, s.z);
          }
          const unsigned int m = __ballot_sync(0xffffffffu, pass);
          if (m) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&q_n[par], (unsigned int) __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) queue[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short) (((h - hg) << 10) | i);
          }
        }
      }
      __syncthreads();
      // ---- search pass
      const int qn = (int) q_n[par];
      if (tid == 0) q_n[par ^ 1] = 0;       // the next group's counter: nobody touches it before the barrier below
      for (int j = tid; j < qn; j += kThreads) {
        const unsigned int e = queue[j];
        const int h = hg + (int) (e >> 10);
        const float4 s = tile[e & 1023u];
        if (point_has_inlier(g, hp_s[h], rball2, rin2, s.x, s.y, s.z)) atomicAdd(&cnt_s[h], 1u);
      }
      __syncthreads();
    }
    if (tid < nh && cnt_s[tid]) atomicAdd(&counts[h0 + tid], cnt_s[tid]);
    __syncthreads();
  }
}

// ---- grid build ----------------------------------------------------------------------------------
__global__ void minmax_kernel(const float4 *__restrict__ p, int n, float *__restrict__ out6) {
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
      // float atomics via int ordering tricks are avoided: two-stage through ordered-int encoding
      int ai = __float_as_int(a), bi = __float_as_int(b);
      ai = ai >= 0 ? ai : ai ^ 0x7fffffff;
      bi = bi >= 0 ? bi : bi ^ 0x7fffffff;
      atomicMin(reinterpret_cast<int *>(out6) + k, ai);
      atomicMax(reinterpret_cast<int *>(out6) + 3 + k, bi);
    }
  }
}

__global__ void cell_key_kernel(const float4 *__restrict__ p, int n, float minx, float miny, float minz,
                                float inv_cell, int nx, int ny, int nz, unsigned int *__restrict__ keys,
                                int *__restrict__ order) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = p[i];
  int cx = min(max((int) floorf((v.x - minx) * inv_cell), 0), nx - 1);
  int cy = min(max((int) floorf((v.y - miny) * inv_cell), 0), ny - 1);
  int cz = min(max((int) floorf((v.z - minz) * inv_cell), 0), nz - 1);
  keys[i] = (unsigned int) ((cz * ny + cy) * nx + cx);
  order[i] = i;
}

// points into cell order + per-cell point counts (the CSR table is the exclusive prefix sum of the counts)
__global__ void gather_cells_kernel(const float4 *__restrict__ p, const int *__restrict__ order,
                                    const unsigned int *__restrict__ keys, int n, float4 *__restrict__ out,
                                    int *__restrict__ cell_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = p[order[i]];
  v.w = 0.f;
  out[i] = v;
  atomicAdd(cell_count + keys[i], 1);
}

// raw occupancy bits over the grid extended by one empty cell on every side (bit xe = x + 1 of row (y + 1, z + 1))
__global__ void occ_raw_kernel(const int *__restrict__ cell_start, int nx, int ny, int nz, int ey, int ewords,
                               unsigned int *__restrict__ raw) {
  long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long) nz * ny * ewords;
  if (t >= total) return;
  int w = (int) (t % ewords);
  int y = (int) ((t / ewords) % ny), z = (int) (t / ((long long) ewords * ny));
  unsigned int bits = 0;
  const int *row = cell_start + ((size_t) z * ny + y) * nx;
  for (int b = 0; b < 32; ++b) {
    int x = 32 * w + b - 1;
    if (x >= 0 && x < nx && row[x + 1] > row[x]) bits |= 1u << b;
  }
  raw[((size_t) (z + 1) * ey + (y + 1)) * ewords + w] = bits;
}

__global__ void occ_dilate_kernel(const unsigned int *__restrict__ raw, int ey, int ez, int ewords, unsigned int *__restrict__ dil) {
  long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long) ez * ey * ewords;
  if (t >= total) return;
  int w = (int) (t % ewords);
  int y = (int) ((t / ewords) % ey), z = (int) (t / ((long long) ewords * ey));
  unsigned int acc = 0;
  for (int dz = -1; dz <= 1; ++dz)
    for (int dy = -1; dy <= 1; ++dy) {
      int yy = y + dy, zz = z + dz;
      if (yy < 0 || yy >= ey || zz < 0 || zz >= ez) continue;
      const unsigned int *r = raw + ((size_t) zz * ey + yy) * ewords;
      unsigned int c = r[w], l = w > 0 ? r[w - 1] : 0u, rr = w + 1 < ewords ? r[w + 1] : 0u;
      acc |= c | (c << 1) | (c >> 1) | (l >> 31) | (rr << 31);
    }
  dil[t] = acc;
}

inline float ordered_int_to_float(int v) {
  v = v >= 0 ? v : v ^ 0x7fffffff;
  float f;
  memcpy(&f, &v, 4);
  return f;
}

}  // namespace

void build_target_grid(Device &dev, const float4 *d_tgt, size_t n, float inlier_dist, TargetGrid &grid) {
  grid.n = n;
  if (n == 0) { grid.nx = grid.ny = grid.nz = 0; return; }
  cudaStream_t s = dev.stream;
  // bounding box
  float *d6 = grid.mm.ensure(6);
  int init[6];
  {
    float big = 3.4e38f, nbig = -3.4e38f;
    int bi, nbi;
    memcpy(&bi, &big, 4); memcpy(&nbi, &nbig, 4);
    nbi = nbi ^ 0x7fffffff;
    init[0] = init[1] = init[2] = bi;
    init[3] = init[4] = init[5] = nbi;
  }
  PLADE_CUDA(cudaMemcpyAsync(d6, init, sizeof(init), cudaMemcpyHostToDevice, s));
  int blocks = std::min(div_up((long long) n, 256), dev.num_sms * 8);
  minmax_kernel<<<blocks, 256, 0, s>>>(d_tgt, (int) n, d6);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  int h6[6];
  PLADE_CUDA(cudaMemcpyAsync(h6, d6, sizeof(h6), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  float mn[3], mx[3];
  for (int k = 0; k < 3; ++k) { mn[k] = ordered_int_to_float(h6[k]); mx[k] = ordered_int_to_float(h6[3 + k]); }

  // Cell edge slightly above the inlier distance so that |dx| < r always lands within +-1 cell even
  // after float rounding of the cell coordinate; grown further if the grid would exceed 2^26 cells.
  double cell = (double) inlier_dist * (1.0 + 1.0 / 1024.0);
  if (!(cell > 0)) cell = 1e-6;
  double ex = (double) mx[0] - mn[0], ey = (double) mx[1] - mn[1], ez = (double) mx[2] - mn[2];
  // a non-finite coordinate (reachable from an untrusted PLY) would make the cell-growing loop below spin for ever
  if (!(std::isfinite(ex) && std::isfinite(ey) && std::isfinite(ez) && ex >= 0 && ey >= 0 && ez >= 0))
    throw std::runtime_error("build_target_grid: non-finite coordinates in the down-sampled target cloud");
  for (int it = 0;; ++it) {
    double cells = (std::floor(ex / cell) + 1) * (std::floor(ey / cell) + 1) * (std::floor(ez / cell) + 1);
    if (cells <= 67108864.0) break;
    if (it > 400) throw std::runtime_error("build_target_grid: grid extent out of range");
    cell *= 1.26;
  }
  grid.cell = (float) cell;
  grid.inv_cell = (float) (1.0 / cell);
  grid.minx = mn[0]; grid.miny = mn[1]; grid.minz = mn[2];
  grid.nx = (int) std::floor(ex / cell) + 1;
  grid.ny = (int) std::floor(ey / cell) + 1;
  grid.nz = (int) std::floor(ez / cell) + 1;
  int ncells = grid.nx * grid.ny * grid.nz;

  unsigned int *keys = grid.keys.ensure(n), *keys2 = grid.keys_alt.ensure(n);
  int *ord = grid.order.ensure(n), *ord2 = grid.order_alt.ensure(n);
  cell_key_kernel<<<div_up((long long) n, 256), 256, 0, s>>>(d_tgt, (int) n, grid.minx, grid.miny, grid.minz,
                                                            grid.inv_cell, grid.nx, grid.ny, grid.nz, keys, ord);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  size_t tmp_bytes = 0;
  int end_bit = 1;
  while ((1ll << end_bit) < ncells) ++end_bit;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ord, ord2, (int) n, 0, end_bit, s);
  unsigned char *tmp = grid.cub_tmp.ensure(tmp_bytes);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ord, ord2, (int) n, 0, end_bit, s);
  dev.launches.add(3);
  float4 *pts = grid.pts.ensure(n);
  int *cs = grid.cell_start.ensure((size_t) ncells + 1);
  // CSR table: zero, count the points of every cell, exclusive prefix sum in place (streams the table twice instead
  // of having one thread per point fill the runs of empty cells in front of it)
  PLADE_CUDA(cudaMemsetAsync(cs, 0, sizeof(int) * ((size_t) ncells + 1), s));
  gather_cells_kernel<<<div_up((long long) n, 256), 256, 0, s>>>(d_tgt, ord2, keys2, (int) n, pts, cs);
  PLADE_LAUNCH_CHECK();
  size_t scan_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cs, cs, ncells + 1, s);
  tmp = grid.cub_tmp.ensure(std::max(scan_bytes, tmp_bytes));
  cub::DeviceScan::ExclusiveSum(tmp, scan_bytes, cs, cs, ncells + 1, s);
  dev.launches.add(3);
  // dilated occupancy bitmap
  const int gex = grid.nx + 2, gey = grid.ny + 2, gez = grid.nz + 2;
  grid.ey = gey;
  grid.ewords = (gex + 31) / 32;
  const size_t words = (size_t) gez * gey * grid.ewords;
  unsigned int *raw = grid.occ_raw.ensure(words), *dil = grid.occ_dil.ensure(words);
  PLADE_CUDA(cudaMemsetAsync(raw, 0, sizeof(unsigned int) * words, s));
  occ_raw_kernel<<<div_up((long long) grid.nz * grid.ny * grid.ewords, 256), 256, 0, s>>>(cs, grid.nx, grid.ny, grid.nz, gey, grid.ewords, raw);
  occ_dilate_kernel<<<div_up((long long) words, 256), 256, 0, s>>>(raw, gey, gez, grid.ewords, dil);
  PLADE_LAUNCH_CHECK();
  dev.launches.add(2);
}

void verify_hypotheses(Device &dev, const float4 *d_src, size_t ns, const TargetGrid &grid,
                       const HypParams *d_hyp, int H, float ball_radius, float inlier_dist,
                       unsigned int *d_counts) {
  cudaStream_t s = dev.stream;
  PLADE_CUDA(cudaMemsetAsync(d_counts, 0, sizeof(unsigned int) * (size_t) H, s));
  if (H == 0 || ns == 0 || grid.n == 0) return;
  // static_cast<float>(radius * radius) with radius a double (kdtree_flann.hpp:193)
  float rball2 = (float) ((double) ball_radius * (double) ball_radius);
  float rin2 = (float) ((double) inlier_dist * (double) inlier_dist);
  GridView g{grid.pts.p, grid.cell_start.p, grid.occ_raw.p, grid.occ_dil.p, grid.ey, grid.ewords, grid.minx, grid.miny, grid.minz, grid.inv_cell, grid.nx, grid.ny, grid.nz};
  int n_tiles = div_up((long long) ns, kTile);
  int n_chunks = div_up(H, kHypChunk);
  long long n_work = (long long) n_tiles * n_chunks;
  // persistent grid: every SM filled to the occupancy the kernel's registers / shared memory allow
  static thread_local int per_sm = 0;
  if (per_sm == 0) {
    PLADE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, verify_kernel, kThreads, 0));
    per_sm = std::max(per_sm, 1);
  }
  int blocks = (int) std::min<long long>(n_work, (long long) dev.num_sms * per_sm);
  verify_kernel<<<blocks, kThreads, 0, s>>>(d_src, (int) ns, g, d_hyp, H, rball2, rin2, d_counts, n_tiles, n_chunks);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
}

// ---- sharded verification: the best hypothesis of this rank's shard as one packed u64 ---------------------------
// key = (float bits of the score) << 32 | (0xFFFFFFFF - global hypothesis index): a MAX over keys -- here over the shard,
// then over the ranks by ncclAllReduce -- picks the highest score, ties -> the lowest index.  Entry i of the shard is the
// hypothesis rank + i * world.  mode 0: the registration score of PLADE/plade.cpp:558-562 with ComputeOverlap's ratio
// (PLADE/util.h:644), the same float / double mix as the host path (matched planes ride in HypParams::pad as an int);
// mode 1: the inlier count itself.
__global__ void shard_key_kernel(const unsigned int *__restrict__ counts, const HypParams *__restrict__ hyp, int n_mine, int rank, int world,
                                 double denom, double n_src_planes, int mode, unsigned long long *__restrict__ key) {
  unsigned long long best = 0ull;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_mine; i += gridDim.x * blockDim.x) {
    const unsigned int cnt = counts[i];
    unsigned int bits;
    if (mode == 0) {
      const float overlap = (float) ((double) cnt / denom);
      const double planes = (double) __float_as_int(hyp[i].pad);
      const float score = (float) (0.2 * (planes / n_src_planes) + 0.8 * overlap);
      bits = __float_as_uint(score);
    } else bits = cnt;
    const unsigned int h = (unsigned int) rank + (unsigned int) i * (unsigned int) world;
    const unsigned long long k = ((unsigned long long) bits << 32) | (unsigned long long) (0xFFFFFFFFu - h);
    best = k > best ? k : best;
  }
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long t = __shfl_down_sync(0xffffffffu, best, o); best = t > best ? t : best; }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(key, best);
}

void shard_best_key(Device &dev, const unsigned int *d_counts, const HypParams *d_hyp, int n_mine, int rank, int world, double denom,
                    double n_src_planes, int mode, unsigned long long *d_key) {
  PLADE_CUDA(cudaMemsetAsync(d_key, 0, sizeof(unsigned long long), dev.stream));
  if (n_mine <= 0) return;
  shard_key_kernel<<<std::min(div_up(n_mine, 256), dev.num_sms), 256, 0, dev.stream>>>(d_counts, d_hyp, n_mine, rank, world, denom, n_src_planes, mode, d_key);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
}

}  // namespace plade
