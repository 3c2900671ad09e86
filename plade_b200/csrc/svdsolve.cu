// K3a — closest points of all line pairs, bit-compatible with the reference's float SVD solve, on sm_100a.
//
// The reference finds the closest points of two 3-D lines with a 9x9 cv::solve(DECOMP_SVD) in FLOAT
// (PLADE/util.cpp:1183-1226) whose rounding error (up to 4e-4 in the points, 5 % of the descriptor matching
// radius) is part of its observable behaviour: a closed-form solution does not reproduce its descriptors.
// svdsolve.h restates that solver; here one thread solves one line pair (the system lives in local memory,
// 728 B/thread; a registration needs a few thousand pairs, one launch).
#include "kernels.h"
#include "svdsolve.h"

namespace plade {

namespace {

// in: v1[3] p1[3] v2[3] p2[3] (directions already normalised by the caller, as the reference does in place)
// out: point1[3] point2[3]
__global__ void nearest_points_kernel(const float *__restrict__ in, int n, float *__restrict__ out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float *q = in + 12 * (size_t) t;
  V3 a, b;
  nearest_points_cv_solve(V3(q[0], q[1], q[2]), V3(q[3], q[4], q[5]), V3(q[6], q[7], q[8]), V3(q[9], q[10], q[11]), a, b);
  float *o = out + 6 * (size_t) t;
  o[0] = a.x; o[1] = a.y; o[2] = a.z;
  o[3] = b.x; o[4] = b.y; o[5] = b.z;
}

}  // namespace

void nearest_points_batch(Device &dev, SvdScratch &sc, const float *h_in12, int n, float *h_out6) {
  if (n <= 0) return;
  cudaStream_t s = dev.stream;
  float *d_in = sc.in.ensure((size_t) 12 * n), *d_out = sc.out.ensure((size_t) 6 * n);
  PLADE_CUDA(cudaMemcpyAsync(d_in, h_in12, sizeof(float) * 12 * (size_t) n, cudaMemcpyHostToDevice, s));
  nearest_points_kernel<<<div_up(n, 64), 64, 0, s>>>(d_in, n, d_out);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  PLADE_CUDA(cudaMemcpyAsync(h_out6, d_out, sizeof(float) * 6 * (size_t) n, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
}

}  // namespace plade
