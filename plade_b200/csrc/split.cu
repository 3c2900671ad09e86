// Input layout conversion: interleaved PLY vertex records (x y z nx ny nz, the C-ABI cloud format)
// -> two float4 streams (position | normal), the layout every point kernel reads with 16-byte loads.
#include "kernels.h"

namespace plade {

namespace {
__global__ void split_kernel(const float *__restrict__ in, int n, float4 *__restrict__ pos, float4 *__restrict__ nrm) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 *p = reinterpret_cast<const float2 *>(in + (size_t) i * 6);   // records are 8-byte aligned
  float2 a = p[0], b = p[1], c = p[2];
  pos[i] = make_float4(a.x, a.y, b.x, 0.f);
  nrm[i] = make_float4(b.y, c.x, c.y, 0.f);
}
}  // namespace

void split_cloud(Device &dev, const float *d_xyzn, size_t n, float4 *pos, float4 *nrm) {
  if (n == 0) return;
  split_kernel<<<div_up((long long) n, 256), 256, 0, dev.stream>>>(d_xyzn, (int) n, pos, nrm);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
}

}  // namespace plade
