// K2b — voxel-grid down-sampling and K2a — k-nearest squared distances, sm_100a.
//
// voxel_downsample*: pcl::VoxelGrid::applyFilter (3rd_party/pcl-1.8.1/filters/include/pcl/filters/
// impl/voxel_grid.hpp:214-437) as PLADE calls it through DownSamplePointCloud (PLADE/util.h:162-184):
//   inverse_leaf = 1.0f / leaf;  min_b = floor(min_p * inverse_leaf);  div_b = max_b - min_b + 1
//   ijk = int(floor(x * inverse_leaf) - float(min_b));  idx = i + j*div_b0 + k*div_b0*div_b1
//   output order = ascending idx;  centroid = (sequential float sum of members) / float(count)
//   (AccumulatorXYZ, common/impl/accumulators.hpp:65-84).
// The reference sorts (idx, point) with an UNSTABLE std::sort, so the order of the float additions
// inside one voxel is unspecified there; here it is canonical: ascending original point index
// (stable radix sort), which makes the kernel deterministic and restatable (oracle/restate.c).
//
// knn_sqdist: the arithmetic of average_spacing (PLADE/util.cpp:1619-1648): FLANN L2_Simple float
// squared distances (dist.h:84-90), the k smallest per query, ascending.  Brute force over all
// points — exact, so it reproduces the kd-tree's result set.
#include "kernels.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace plade {

namespace {

constexpr int kMaxGroups = 64;

struct GroupParams {
  float inv_leaf;
  int min_b[3];
  int div0, div01;
  int valid;     // 0 => leaf too small for this group (PCL returns the input unchanged)
};

__device__ __forceinline__ int f2o(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
inline float o2f(int v) { v = v >= 0 ? v : v ^ 0x7fffffff; float f; memcpy(&f, &v, 4); return f; }

__global__ void group_bbox_kernel(const float4 *__restrict__ p, int n, const int *__restrict__ group,
                                  int *__restrict__ bbox /* ngroups*6 ordered ints */) {
  // per-block boxes in shared memory first (global atomics on <= 64 x 6 addresses from every thread
  // serialise in L2: 11.7 ms for 2 x 2M points in the first profile), one flush per block at the end
  __shared__ int sb[kMaxGroups * 6];
  for (int k = threadIdx.x; k < kMaxGroups * 6; k += blockDim.x) sb[k] = (k % 6 < 3) ? 0x7f7fffff : (int) (0xff7fffff ^ 0x7fffffff);
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int g = group ? group[i] : 0;
    if (g < 0 || g >= kMaxGroups) continue;
    float4 v = p[i];
    int *b = sb + 6 * g;
    int ox = f2o(v.x), oy = f2o(v.y), oz = f2o(v.z);
    if (ox < b[0]) atomicMin(b + 0, ox);
    if (oy < b[1]) atomicMin(b + 1, oy);
    if (oz < b[2]) atomicMin(b + 2, oz);
    if (ox > b[3]) atomicMax(b + 3, ox);
    if (oy > b[4]) atomicMax(b + 4, oy);
    if (oz > b[5]) atomicMax(b + 5, oz);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < kMaxGroups * 6; k += blockDim.x) {
    if (k % 6 < 3) { if (sb[k] != 0x7f7fffff) atomicMin(bbox + k, sb[k]); }
    else { if (sb[k] != (int) (0xff7fffff ^ 0x7fffffff)) atomicMax(bbox + k, sb[k]); }
  }
}

__global__ void single_bbox_kernel(const float4 *__restrict__ p, int n, int *__restrict__ bbox) {
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
    if (threadIdx.x == 0) { atomicMin(bbox + k, f2o(a)); atomicMax(bbox + 3 + k, f2o(b)); }
  }
}

__global__ void voxel_key_kernel(const float4 *__restrict__ p, int n, const int *__restrict__ group,
                                 const GroupParams *__restrict__ gp, int vbits, unsigned long long unselected_key,
                                 unsigned long long *__restrict__ keys, int *__restrict__ idx, int *__restrict__ n_selected) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool sel = false;
  if (i < n) {
    int g = group ? group[i] : 0;
    unsigned long long key = unselected_key;      // sorts behind every (group, voxel) key
    if (g >= 0) {
      GroupParams q = gp[g];
      float4 v = p[i];
      int i0 = (int) (floorf(__fmul_rn(v.x, q.inv_leaf)) - (float) q.min_b[0]);
      int i1 = (int) (floorf(__fmul_rn(v.y, q.inv_leaf)) - (float) q.min_b[1]);
      int i2 = (int) (floorf(__fmul_rn(v.z, q.inv_leaf)) - (float) q.min_b[2]);
      int id = i0 + i1 * q.div0 + i2 * q.div01;
      // a group whose grid would overflow int32 keeps every point: one "voxel" per point
      if (!q.valid) id = i;
      key = ((unsigned long long) (unsigned int) g << vbits) | (unsigned int) id;     // only vbits + gbits key bits are sorted
      sel = true;
    }
    keys[i] = key;
    idx[i] = i;
  }
  unsigned int m = __ballot_sync(0xffffffffu, sel);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_selected, __popc(m));
}

__global__ void head_flag_kernel(const unsigned long long *__restrict__ keys, int m, int *__restrict__ flags) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// flags holds the inclusive scan (voxel id + 1).  seg_start[v] = first sorted position of voxel v.
__global__ void seg_start_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ scan, int m, int vbits,
                                 int *__restrict__ seg_start, int *__restrict__ group_first_voxel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  bool head = (i == 0) || keys[i] != keys[i - 1];
  if (head) {
    int v = scan[i] - 1;
    seg_start[v] = i;
    int g = (int) (keys[i] >> vbits);
    if (i == 0 || (int) (keys[i - 1] >> vbits) != g) group_first_voxel[g] = v;
  }
}

__global__ void centroid_kernel(const float4 *__restrict__ p, const int *__restrict__ idx,
                                const int *__restrict__ seg_start, int nvox, int m, float4 *__restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  int b = seg_start[v], e = (v + 1 < nvox) ? seg_start[v + 1] : m;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int i = b; i < e; ++i) {
    float4 q = __ldg(p + idx[i]);
    sx = __fadd_rn(sx, q.x); sy = __fadd_rn(sy, q.y); sz = __fadd_rn(sz, q.z);
  }
  float c = (float) (e - b);
  out[v] = make_float4(__fdiv_rn(sx, c), __fdiv_rn(sy, c), __fdiv_rn(sz, c), 0.f);
}

size_t voxel_impl(Device &dev, VoxelScratch &sc, const float4 *d_pts, size_t n, const int *d_group, int ngroups,
                  float leaf, DevBuf<float4> &out, std::vector<int> *group_start) {
  if (ngroups > kMaxGroups) throw std::runtime_error("voxel_downsample: too many groups");
  cudaStream_t s = dev.stream;
  if (group_start) group_start->assign(ngroups + 1, 0);
  if (n == 0) return 0;
  // 1. per-group bounding boxes
  int *d_bbox = reinterpret_cast<int *>(sc.minmax.ensure(6 * kMaxGroups + 8));
  std::vector<int> h_bbox(6 * ngroups);
  for (int g = 0; g < ngroups; ++g)
    for (int k = 0; k < 3; ++k) { h_bbox[6 * g + k] = 0x7f7fffff; h_bbox[6 * g + 3 + k] = (int) 0xff7fffff ^ 0x7fffffff; }
  PLADE_CUDA(cudaMemcpyAsync(d_bbox, h_bbox.data(), sizeof(int) * h_bbox.size(), cudaMemcpyHostToDevice, s));
  int blocks = std::min(div_up((long long) n, 256), dev.num_sms * 8);
  if (d_group) group_bbox_kernel<<<blocks, 256, 0, s>>>(d_pts, (int) n, d_group, d_bbox);
  else single_bbox_kernel<<<blocks, 256, 0, s>>>(d_pts, (int) n, d_bbox);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  PLADE_CUDA(cudaMemcpyAsync(h_bbox.data(), d_bbox, sizeof(int) * h_bbox.size(), cudaMemcpyDeviceToHost, s));
  stream_sync(s);

  // 2. per-group grid parameters, exactly as voxel_grid.hpp:237-262
  std::vector<GroupParams> gp(ngroups);
  long long max_id = 0;        // largest voxel index any group can produce -> number of key bits to sort
  const float inv = 1.0f / leaf;
  for (int g = 0; g < ngroups; ++g) {
    GroupParams &q = gp[g];
    q.inv_leaf = inv;
    float mn[3], mx[3];
    for (int k = 0; k < 3; ++k) { mn[k] = o2f(h_bbox[6 * g + k]); mx[k] = o2f(h_bbox[6 * g + 3 + k]); }
    q.valid = 1;
    if (mn[0] > mx[0]) { q.min_b[0] = q.min_b[1] = q.min_b[2] = 0; q.div0 = q.div01 = 1; continue; }  // empty group
    long long dx = (long long) ((mx[0] - mn[0]) * inv) + 1;
    long long dy = (long long) ((mx[1] - mn[1]) * inv) + 1;
    long long dz = (long long) ((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > 2147483647ll) q.valid = 0;
    int maxb[3];
    for (int k = 0; k < 3; ++k) {
      q.min_b[k] = (int) std::floor(mn[k] * inv);
      maxb[k] = (int) std::floor(mx[k] * inv);
    }
    int d0 = maxb[0] - q.min_b[0] + 1, d1 = maxb[1] - q.min_b[1] + 1, d2 = maxb[2] - q.min_b[2] + 1;
    q.div0 = d0;
    q.div01 = d0 * d1;
    max_id = std::max(max_id, q.valid ? (long long) d0 * d1 * d2 - 1 : (long long) n - 1);
  }
  int vbits = 1;
  while (vbits < 32 && (1ll << vbits) <= max_id) ++vbits;
  static_assert(sizeof(GroupParams) == 28, "GroupParams layout");
  // layout of sc.counter: [0] selected-point counter | [16..) GroupParams[kMaxGroups] | first voxel per group
  int *d_counter = sc.counter.ensure(16 + kMaxGroups * 8 + kMaxGroups);
  GroupParams *d_params = reinterpret_cast<GroupParams *>(d_counter + 16);
  int *d_group_first = d_counter + 16 + kMaxGroups * 7;
  PLADE_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(int) * 16, s));
  PLADE_CUDA(cudaMemcpyAsync(d_params, gp.data(), sizeof(GroupParams) * ngroups, cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemsetAsync(d_group_first, 0xff, sizeof(int) * kMaxGroups, s));

  // 3. keys, 4. stable sort
  unsigned long long *keys = sc.keys.ensure(n), *keys2 = sc.keys_alt.ensure(n);
  int *idx = sc.idx.ensure(n), *idx2 = sc.idx_alt.ensure(n);
  int gbits = 1;
  while ((1 << gbits) < ngroups + 1) ++gbits;
  // key = group << vbits | voxel; unselected points carry the (unused) group 2^gbits - 1 and sort to the end.
  // Only the vbits + gbits significant bits are sorted (32 instead of 64 for the per-plane grids of a 2 M cloud).
  const unsigned long long unselected_key = (((1ull << gbits) - 1) << vbits) | ((1ull << vbits) - 1);
  voxel_key_kernel<<<div_up((long long) n, 256), 256, 0, s>>>(d_pts, (int) n, d_group, d_params, vbits, unselected_key, keys, idx, d_counter);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  size_t tmp_bytes = 0;
  int end_bit = vbits + gbits;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, idx, idx2, (int) n, 0, end_bit, s);
  unsigned char *tmp = sc.cub_tmp.ensure(tmp_bytes);
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, idx, idx2, (int) n, 0, end_bit, s);
  dev.launches.add(d_group ? 9 : 5);
  int m = 0;
  PLADE_CUDA(cudaMemcpyAsync(&m, d_counter, sizeof(int), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  if (m == 0) return 0;

  // 5. voxel ids
  int *flags = sc.flags.ensure(n);
  head_flag_kernel<<<div_up(m, 256), 256, 0, s>>>(keys2, m, flags);
  PLADE_LAUNCH_CHECK();
  size_t scan_bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, flags, flags, m, s);
  tmp = sc.cub_tmp.ensure(std::max(scan_bytes, tmp_bytes));
  cub::DeviceScan::InclusiveSum(tmp, scan_bytes, flags, flags, m, s);
  dev.launches.add(3);
  int nvox = 0;
  PLADE_CUDA(cudaMemcpyAsync(&nvox, flags + (m - 1), sizeof(int), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  int *seg = sc.seg_start.ensure((size_t) nvox + 1);
  seg_start_kernel<<<div_up(m, 256), 256, 0, s>>>(keys2, flags, m, vbits, seg, d_group_first);
  PLADE_LAUNCH_CHECK();
  // 6. centroids
  float4 *o = out.ensure(nvox);
  centroid_kernel<<<div_up(nvox, 128), 128, 0, s>>>(d_pts, idx2, seg, nvox, m, o);
  PLADE_LAUNCH_CHECK();
  dev.launches.add(2);
  if (group_start) {
    std::vector<int> first(kMaxGroups);
    PLADE_CUDA(cudaMemcpyAsync(first.data(), d_group_first, sizeof(int) * kMaxGroups, cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    (*group_start)[ngroups] = nvox;
    for (int g = ngroups - 1; g >= 0; --g) (*group_start)[g] = first[g] >= 0 ? first[g] : (*group_start)[g + 1];
  }
  return (size_t) nvox;
}

// ---- brute-force k nearest squared distances ------------------------------------------------------
constexpr int kKnnK = 8;          // capacity of the per-thread sorted list (k <= 8)
constexpr int kKnnTile = 1024;
constexpr int kKnnThreads = 128;

__global__ void __launch_bounds__(kKnnThreads)
knn_partial_kernel(const float4 *__restrict__ pts, int n, const int *__restrict__ qidx, int nq, int k,
                   int chunk_len, float *__restrict__ partial /* [nsplit][nq][k] */) {
  __shared__ float4 tile[kKnnTile];
  const int q = blockIdx.x * kKnnThreads + threadIdx.x;
  const int split = blockIdx.y;
  const int p_begin = split * chunk_len, p_end = min(n, p_begin + chunk_len);
  float4 qp = make_float4(0, 0, 0, 0);
  if (q < nq) qp = pts[qidx[q]];
  float best[kKnnK];
#pragma unroll
  for (int j = 0; j < kKnnK; ++j) best[j] = 3.4e38f;
  for (int base = p_begin; base < p_end; base += kKnnTile) {
    int cnt = min(kKnnTile, p_end - base);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kKnnThreads) tile[i] = pts[base + i];
    __syncthreads();
    if (q < nq) {
      for (int i = 0; i < cnt; ++i) {
        float4 t = tile[i];
        float dx = __fsub_rn(qp.x, t.x), dy = __fsub_rn(qp.y, t.y), dz = __fsub_rn(qp.z, t.z);
        float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d < best[kKnnK - 1]) {
          best[kKnnK - 1] = d;
#pragma unroll
          for (int j = kKnnK - 1; j > 0; --j) {
            if (best[j] < best[j - 1]) { float tsw = best[j]; best[j] = best[j - 1]; best[j - 1] = tsw; }
          }
        }
      }
    }
  }
  if (q < nq)
    for (int j = 0; j < k; ++j) partial[((size_t) split * nq + q) * k + j] = best[j];
}

__global__ void knn_merge_kernel(const float *__restrict__ partial, int nsplit, int nq, int k, float *__restrict__ out) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  float best[kKnnK];
#pragma unroll
  for (int j = 0; j < kKnnK; ++j) best[j] = 3.4e38f;
  for (int s = 0; s < nsplit; ++s)
    for (int j = 0; j < k; ++j) {
      float d = partial[((size_t) s * nq + q) * k + j];
      if (d < best[kKnnK - 1]) {
        best[kKnnK - 1] = d;
#pragma unroll
        for (int jj = kKnnK - 1; jj > 0; --jj)
          if (best[jj] < best[jj - 1]) { float t = best[jj]; best[jj] = best[jj - 1]; best[jj - 1] = t; }
      }
    }
  for (int j = 0; j < k; ++j) out[(size_t) q * k + j] = best[j];
}

// ---- grid-accelerated exact k nearest squared distances ------------------------------------------------
__global__ void knn_key_kernel(const float4 *__restrict__ p, int n, float minx, float miny, float minz, float inv_c, int nx, int ny,
                               int nz, unsigned int *__restrict__ keys, int *__restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 v = p[i];
  int cx = min(max((int) floorf((v.x - minx) * inv_c), 0), nx - 1);
  int cy = min(max((int) floorf((v.y - miny) * inv_c), 0), ny - 1);
  int cz = min(max((int) floorf((v.z - minz) * inv_c), 0), nz - 1);
  keys[i] = ((unsigned int) cz << 20) | ((unsigned int) cy << 10) | (unsigned int) cx;
  idx[i] = i;
}

__global__ void knn_gather_kernel(const float4 *__restrict__ p, const int *__restrict__ idx, int n, float4 *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = p[idx[i]];
}

__device__ __forceinline__ int lower_bound_u32(const unsigned int *a, int n, unsigned int key) {
  int lo = 0, hi = n;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid; }
  return lo;
}

// One thread per query: scan the cube of cells of radius rho around the query's cell, keep the k smallest
// FLANN-L2_Simple distances; every point outside the cube is farther than rho * cell, so the list is exact
// as soon as its k-th entry is below that bound (otherwise the cube grows).
__global__ void knn_grid_kernel(const float4 *__restrict__ pts, const float4 *__restrict__ spts, const unsigned int *__restrict__ keys,
                                int n, const int *__restrict__ qidx, int nq, int k, float minx, float miny, float minz, float cell,
                                float inv_c, int nx, int ny, int nz, float *__restrict__ out) {
  int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const float4 q = pts[qidx[qi]];
  const int cx = min(max((int) floorf((q.x - minx) * inv_c), 0), nx - 1);
  const int cy = min(max((int) floorf((q.y - miny) * inv_c), 0), ny - 1);
  const int cz = min(max((int) floorf((q.z - minz) * inv_c), 0), nz - 1);
  float best[kKnnK];
  for (int rho = 1;; ++rho) {
#pragma unroll
    for (int j = 0; j < kKnnK; ++j) best[j] = 3.4e38f;
    const bool all = rho > 12 || (cx - rho <= 0 && cy - rho <= 0 && cz - rho <= 0 && cx + rho >= nx - 1 && cy + rho >= ny - 1 && cz + rho >= nz - 1);
    const int z0 = all ? 0 : max(cz - rho, 0), z1 = all ? nz - 1 : min(cz + rho, nz - 1);
    const int y0 = all ? 0 : max(cy - rho, 0), y1 = all ? ny - 1 : min(cy + rho, ny - 1);
    const int x0 = all ? 0 : max(cx - rho, 0), x1 = all ? nx - 1 : min(cx + rho, nx - 1);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        unsigned int base = ((unsigned int) z << 20) | ((unsigned int) y << 10);
        int lo = lower_bound_u32(keys, n, base | (unsigned int) x0), hi = lower_bound_u32(keys, n, (base | (unsigned int) x1) + 1u);
        for (int i = lo; i < hi; ++i) {
          float4 t = __ldg(spts + i);
          float dx = __fsub_rn(q.x, t.x), dy = __fsub_rn(q.y, t.y), dz = __fsub_rn(q.z, t.z);
          float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          if (d < best[kKnnK - 1]) {
            best[kKnnK - 1] = d;
#pragma unroll
            for (int j = kKnnK - 1; j > 0; --j)
              if (best[j] < best[j - 1]) { float tsw = best[j]; best[j] = best[j - 1]; best[j - 1] = tsw; }
          }
        }
      }
    if (all) break;
    float lim = (float) rho * cell * 0.999f;
    if (best[k - 1] <= lim * lim) break;
  }
  for (int j = 0; j < k; ++j) out[(size_t) qi * k + j] = best[j];
}

}  // namespace

size_t voxel_downsample(Device &dev, VoxelScratch &sc, const float4 *d_pts, size_t n, float leaf,
                        DevBuf<float4> &out) {
  return voxel_impl(dev, sc, d_pts, n, nullptr, 1, leaf, out, nullptr);
}

size_t voxel_downsample_groups(Device &dev, VoxelScratch &sc, const float4 *d_pts, size_t n,
                               const int *d_group, int ngroups, float leaf, DevBuf<float4> &out,
                               std::vector<int> &out_group_start) {
  return voxel_impl(dev, sc, d_pts, n, d_group, ngroups, leaf, out, &out_group_start);
}

void knn_sqdist(Device &dev, KnnScratch &ks, const float4 *d_pts, size_t n, const int *d_query_idx, int nq, int k, float *d_out) {
  if (k > kKnnK) throw std::runtime_error("knn_sqdist: k > 8");
  if (nq == 0) return;
  cudaStream_t s = dev.stream;
  if (n <= 8192) {
    // small clouds: plain brute force
    int qblocks = div_up(nq, kKnnThreads);
    int nsplit = std::max(1, std::min(64, (dev.num_sms * 4 + qblocks - 1) / qblocks));
    int chunk = div_up((long long) n, nsplit);
    chunk = ((chunk + kKnnTile - 1) / kKnnTile) * kKnnTile;
    nsplit = div_up((long long) n, chunk);
    float *d_partial = ks.partial.ensure((size_t) nsplit * nq * k);
    knn_partial_kernel<<<dim3(qblocks, nsplit), kKnnThreads, 0, s>>>(d_pts, (int) n, d_query_idx, nq, k, chunk, d_partial);
    PLADE_LAUNCH_CHECK();
    knn_merge_kernel<<<div_up(nq, 128), 128, 0, s>>>(d_partial, nsplit, nq, k, d_out);
    PLADE_LAUNCH_CHECK();
    dev.launches.add(2);
    return;
  }
  // uniform grid sized from a surface-density estimate (cell ~ 2.5 x expected spacing, <= 1023 cells per axis)
  DevBuf<int> &bbox_buf = ks.bbox_buf, &idx_a = ks.idx_a, &idx_b = ks.idx_b;
  DevBuf<unsigned int> &key_a = ks.key_a, &key_b = ks.key_b;
  DevBuf<float4> &sorted_pts = ks.sorted_pts;
  DevBuf<unsigned char> &cub_tmp = ks.cub_tmp;
  int *d_bbox = bbox_buf.ensure(8);
  int h_bbox[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int) (0xff7fffff ^ 0x7fffffff), (int) (0xff7fffff ^ 0x7fffffff), (int) (0xff7fffff ^ 0x7fffffff)};
  PLADE_CUDA(cudaMemcpyAsync(d_bbox, h_bbox, sizeof(h_bbox), cudaMemcpyHostToDevice, s));
  single_bbox_kernel<<<std::min(div_up((long long) n, 256), dev.num_sms * 8), 256, 0, s>>>(d_pts, (int) n, d_bbox);
  PLADE_LAUNCH_CHECK();
  PLADE_CUDA(cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  float mn[3], mx[3];
  for (int a = 0; a < 3; ++a) { mn[a] = o2f(h_bbox[a]); mx[a] = o2f(h_bbox[3 + a]); }
  double ex = (double) mx[0] - mn[0], ey = (double) mx[1] - mn[1], ez = (double) mx[2] - mn[2];
  double area = 2.0 * (ex * ey + ey * ez + ez * ex);
  double maxext = std::max(ex, std::max(ey, ez));
  double cell = 2.5 * std::sqrt(std::max(area, 1e-30) / (double) n);
  if (!(cell > 0) || !std::isfinite(cell)) cell = 1e-6;
  cell = std::max(cell, maxext / 1000.0);
  if (!(cell > 0)) cell = 1e-6;
  int nx = std::min(1023, (int) std::floor(ex / cell)) + 1, ny = std::min(1023, (int) std::floor(ey / cell)) + 1,
      nz = std::min(1023, (int) std::floor(ez / cell)) + 1;
  float fcell = (float) cell, inv_c = (float) (1.0 / cell);
  unsigned int *ka = key_a.ensure(n), *kb = key_b.ensure(n);
  int *ia = idx_a.ensure(n), *ib = idx_b.ensure(n);
  knn_key_kernel<<<div_up((long long) n, 256), 256, 0, s>>>(d_pts, (int) n, mn[0], mn[1], mn[2], inv_c, nx, ny, nz, ka, ia);
  PLADE_LAUNCH_CHECK();
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, ka, kb, ia, ib, (int) n, 0, 30, s);
  unsigned char *tmp = cub_tmp.ensure(tb);
  cub::DeviceRadixSort::SortPairs(tmp, tb, ka, kb, ia, ib, (int) n, 0, 30, s);
  float4 *sp = sorted_pts.ensure(n);
  knn_gather_kernel<<<div_up((long long) n, 256), 256, 0, s>>>(d_pts, ib, (int) n, sp);
  knn_grid_kernel<<<div_up(nq, 64), 64, 0, s>>>(d_pts, sp, kb, (int) n, d_query_idx, nq, k, mn[0], mn[1], mn[2], fcell, inv_c, nx, ny, nz, d_out);
  PLADE_LAUNCH_CHECK();
  dev.launches.add(9);
}

}  // namespace plade
