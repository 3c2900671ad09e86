// Oriented bounding boxes of the down-sampled clouds (whole cloud + one per plane) on sm_100a.
//
// Replaces ComputeBoundingBox (PLADE/util.h:187-248): centroid (pcl::compute3DCentroid, centroid.hpp:79-122),
// normalised covariance (centroid.hpp:180-259), eigenvectors, extent of the points in the eigen frame
// (pcl::transformPointCloud + getMinMax3D), centre and the eight corners.  All segments (both clouds, all
// planes) go through three launches: moment sums, covariance sums, min/max in the rotated frame; the 3x3
// eigen-decomposition in between runs on the host.
// The reference accumulates the moments sequentially in float; here they are reduced in fp64 and rounded
// once (the result differs from the reference by the reference's own float accumulation error, ~1e-6
// relative; tests/test_gpu_parity.py::test_bounding_box_vs_reference holds both to 2e-5).
#include "kernels.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <cmath>
#include <cstring>

namespace plade {

namespace {

constexpr int kObbChunk = 4096;
constexpr int kObbThreads = 256;

struct BlockDesc { int seg, begin, end, pad; };

__device__ __forceinline__ int f2o(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
inline float o2f(int v) { v = v >= 0 ? v : v ^ 0x7fffffff; float f; memcpy(&f, &v, 4); return f; }

__global__ void __launch_bounds__(kObbThreads)
obb_sum_kernel(const ObbSeg *__restrict__ segs, const BlockDesc *__restrict__ blocks, double *__restrict__ acc /* seg x 16 */) {
  const BlockDesc b = blocks[blockIdx.x];
  const float4 *p = segs[b.seg].p;
  double sx = 0, sy = 0, sz = 0;
  for (int i = b.begin + threadIdx.x; i < b.end; i += kObbThreads) { float4 v = p[i]; sx += v.x; sy += v.y; sz += v.z; }
  typedef cub::BlockReduce<double, kObbThreads> BR;
  __shared__ typename BR::TempStorage tmp;
  double r0 = BR(tmp).Sum(sx); __syncthreads();
  double r1 = BR(tmp).Sum(sy); __syncthreads();
  double r2 = BR(tmp).Sum(sz);
  if (threadIdx.x == 0) { atomicAdd(acc + 16 * b.seg + 0, r0); atomicAdd(acc + 16 * b.seg + 1, r1); atomicAdd(acc + 16 * b.seg + 2, r2); }
}

__global__ void __launch_bounds__(kObbThreads)
obb_cov_kernel(const ObbSeg *__restrict__ segs, const BlockDesc *__restrict__ blocks, double *__restrict__ acc) {
  const BlockDesc b = blocks[blockIdx.x];
  const ObbSeg sg = segs[b.seg];
  const float fn = (float) sg.n;
  // centroid = float(sum) / float(n)  (compute3DCentroid: float accumulator, then `centroid /= n`)
  const float cx = __fdiv_rn((float) acc[16 * b.seg + 0], fn), cy = __fdiv_rn((float) acc[16 * b.seg + 1], fn), cz = __fdiv_rn((float) acc[16 * b.seg + 2], fn);
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = b.begin + threadIdx.x; i < b.end; i += kObbThreads) {
    float4 v = sg.p[i];
    double px = (double) __fsub_rn(v.x, cx), py = (double) __fsub_rn(v.y, cy), pz = (double) __fsub_rn(v.z, cz);
    c[0] += px * px; c[1] += py * px; c[2] += pz * px; c[3] += py * py; c[4] += py * pz; c[5] += pz * pz;
  }
  typedef cub::BlockReduce<double, kObbThreads> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int k = 0; k < 6; ++k) {
    double r = BR(tmp).Sum(c[k]);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(acc + 16 * b.seg + 4 + k, r);
  }
}

__global__ void __launch_bounds__(kObbThreads)
obb_minmax_kernel(const ObbSeg *__restrict__ segs, const BlockDesc *__restrict__ blocks, const float *__restrict__ frames /* seg x 12: Rt, t */,
                  int *__restrict__ mm /* seg x 6 ordered ints */) {
  const BlockDesc b = blocks[blockIdx.x];
  const float4 *p = segs[b.seg].p;
  M3 Rt;
  for (int k = 0; k < 9; ++k) Rt.m[k] = frames[12 * b.seg + k];
  const V3 t(frames[12 * b.seg + 9], frames[12 * b.seg + 10], frames[12 * b.seg + 11]);
  float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  for (int i = b.begin + threadIdx.x; i < b.end; i += kObbThreads) {
    float4 v = p[i];
    V3 q = xform(Rt, t, V3(v.x, v.y, v.z));
    mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
    mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
  }
  typedef cub::BlockReduce<float, kObbThreads> BR;
  __shared__ typename BR::TempStorage tmp;
  for (int k = 0; k < 3; ++k) {
    float a = BR(tmp).Reduce(mn[k], cub::Min());
    __syncthreads();
    float c = BR(tmp).Reduce(mx[k], cub::Max());
    __syncthreads();
    if (threadIdx.x == 0) { atomicMin(mm + 6 * b.seg + k, f2o(a)); atomicMax(mm + 6 * b.seg + 3 + k, f2o(c)); }
  }
}

}  // namespace

void obb_segments(Device &dev, ObbScratch &sc, const std::vector<ObbSeg> &segs, std::vector<ObbResult> &out) {
  const int S = (int) segs.size();
  out.assign(S, ObbResult());
  for (int i = 0; i < S; ++i) out[i].rc = segs[i].n > 0 ? 0 : -1;
  std::vector<BlockDesc> blocks;
  for (int i = 0; i < S; ++i)
    for (int b = 0; b < segs[i].n; b += kObbChunk) blocks.push_back({i, b, std::min(segs[i].n, b + kObbChunk), 0});
  if (blocks.empty()) return;
  cudaStream_t s = dev.stream;
  ObbSeg *d_segs = reinterpret_cast<ObbSeg *>(sc.segs.ensure(sizeof(ObbSeg) * S));
  BlockDesc *d_blocks = reinterpret_cast<BlockDesc *>(sc.blocks.ensure(sizeof(BlockDesc) * blocks.size()));
  double *d_acc = sc.acc.ensure((size_t) 16 * S);
  float *d_frames = sc.frames.ensure((size_t) 12 * S);
  int *d_mm = sc.mm.ensure((size_t) 6 * S);
  PLADE_CUDA(cudaMemcpyAsync(d_segs, segs.data(), sizeof(ObbSeg) * S, cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemcpyAsync(d_blocks, blocks.data(), sizeof(BlockDesc) * blocks.size(), cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double) * 16 * S, s));
  const int nb = (int) blocks.size();
  obb_sum_kernel<<<nb, kObbThreads, 0, s>>>(d_segs, d_blocks, d_acc);
  obb_cov_kernel<<<nb, kObbThreads, 0, s>>>(d_segs, d_blocks, d_acc);
  PLADE_LAUNCH_CHECK();
  std::vector<double> acc((size_t) 16 * S);
  PLADE_CUDA(cudaMemcpyAsync(acc.data(), d_acc, sizeof(double) * 16 * S, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  // host: eigen frames (same statements as the tail of ComputeBoundingBox)
  std::vector<float> frames((size_t) 12 * S, 0.f);
  std::vector<M3> Es(S);
  std::vector<V3> ctrs(S);
  std::vector<int> mm_init((size_t) 6 * S);
  for (int i = 0; i < S; ++i) {
    for (int k = 0; k < 3; ++k) { mm_init[6 * i + k] = 0x7f7fffff; mm_init[6 * i + 3 + k] = (int) (0xff7fffff ^ 0x7fffffff); }
    if (segs[i].n <= 0) continue;
    const float fn = (float) segs[i].n;
    const double *a = &acc[16 * i];
    V3 ctr((float) a[0] / fn, (float) a[1] / fn, (float) a[2] / fn);
    float c00 = (float) a[4] / fn, c01 = (float) a[5] / fn, c02 = (float) a[6] / fn, c11 = (float) a[7] / fn, c12 = (float) a[8] / fn, c22 = (float) a[9] / fn;
    // same solver, statement for statement, as Eigen::SelfAdjointEigenSolver<Matrix3f>: the signs of the frame
    // axes decide the order / orientation of the corner loop the penetration filter later walks
    float A[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}}, w[3], V[3][3];
    sym_eig3f_eigen(A, w, V);
    V3 e0(V[0][0], V[1][0], V[2][0]), e1(V[0][1], V[1][1], V[2][1]);
    V3 e2 = cross(e0, e1);
    M3 E, Rt;
    for (int r = 0; r < 3; ++r) { E(r, 0) = e0[r]; E(r, 1) = e1[r]; E(r, 2) = e2[r]; }
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rt(r, c) = E(c, r);
    V3 t = -1.f * mul(Rt, ctr);
    memcpy(&frames[12 * i], Rt.m, sizeof(float) * 9);
    frames[12 * i + 9] = t.x; frames[12 * i + 10] = t.y; frames[12 * i + 11] = t.z;
    Es[i] = E;
    ctrs[i] = ctr;
  }
  PLADE_CUDA(cudaMemcpyAsync(d_frames, frames.data(), sizeof(float) * 12 * S, cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaMemcpyAsync(d_mm, mm_init.data(), sizeof(int) * 6 * S, cudaMemcpyHostToDevice, s));
  obb_minmax_kernel<<<nb, kObbThreads, 0, s>>>(d_segs, d_blocks, d_frames, d_mm);
  PLADE_LAUNCH_CHECK();
  dev.launches.add(3);
  std::vector<int> mm((size_t) 6 * S);
  PLADE_CUDA(cudaMemcpyAsync(mm.data(), d_mm, sizeof(int) * 6 * S, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  for (int i = 0; i < S; ++i) {
    if (segs[i].n <= 0) continue;
    ObbResult &o = out[i];
    V3 mn(o2f(mm[6 * i]), o2f(mm[6 * i + 1]), o2f(mm[6 * i + 2])), mx(o2f(mm[6 * i + 3]), o2f(mm[6 * i + 4]), o2f(mm[6 * i + 5]));
    V3 mean_diag = 0.5f * (mx + mn);
    o.center = mul(Es[i], mean_diag) + ctrs[i];
    o.width = mx.x - mn.x;
    o.depth = mx.y - mn.y;
    o.height = mx.z - mn.z;
    float x = mn.x, y = mn.y, z = mn.z;
    V3 loc[8] = {mn,
                 V3(x, (float) (y + o.depth), z),
                 V3(x, (float) (y + o.depth), (float) (z + o.height)),
                 V3(x, y, (float) (z + o.height)),
                 V3((float) (x + o.width), y, (float) (z + o.height)),
                 V3((float) (x + o.width), (float) (y + o.depth), z),
                 V3((float) (x + o.width), y, z),
                 V3((float) (x + o.width), (float) (y + o.depth), (float) (z + o.height))};
    for (int k = 0; k < 8; ++k) o.corners[k] = xform(Es[i], ctrs[i], loc[k]);
  }
}

}  // namespace plade
