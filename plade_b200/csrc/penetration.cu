// K4d — penetration filter on sm_100a.
//
// Replaces the loop PLADE/util.cpp:450-519 and AreTwoPlanesPenetrable (PLADE/util.cpp:1279-1458): a
// candidate transform is dropped when some transformed source plane "penetrates" a non-coincident
// target plane, i.e. points of both planes lie on both sides of the other plane along the segment
// where their bounding rectangles cross.  The reference transforms every source plane cloud, builds
// a FLANN kd-tree per (hypothesis, plane) and walks the segment with radius searches; its result
// per (hypothesis, source plane, target plane) triple has no side effects, so here all triples are
// evaluated independently and OR-reduced per hypothesis:
//   kernel 1 (one thread per triple): the scalar geometry — coincidence gate, plane/plane line,
//     clipping against both rectangles (the same host/device functions of linalg.h), segment;
//   kernel 2 (one block per surviving triple): the two sampling passes.  The radius searches become
//     exact tests of every plane point against the few segment samples its projection can reach
//     (FLANN L2_Simple float arithmetic, strict '<' against float(r*r)), with the "fresh point"
//     bookkeeping replaced by evaluating each probe point once.
#include "kernels.h"
#include "linalg.h"
#include "svdsolve.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace plade {

namespace {

constexpr int kMaxSteps = 4096;
constexpr int kPenThreads = 256;

struct Triple {
  int h, i1, j1, nsteps;
  float start[3], direc[3];
  float plane1[4];
};

struct PenArgs {
  const float4 *planes_s, *planes_t;          // (n, d) per plane
  const float *corners_s, *corners_t;         // P x 4 x 3
  const float *center_s, *center_t;           // P x 3
  const int *off_s, *off_t;                   // P + 1
  const float4 *pts_s, *pts_t;                // per-plane ds points
  const float *hyp;                           // H x 12 (R row-major, T)
  const float *dist_table;                    // dist_k = k-fold float accumulation of searchRadius
  int H, Ps, Pt;
  float lengthThreshold, angleThreshold, searchRadius, minDistance;
  int minPoints;
};

__device__ __forceinline__ V3 ld3(const float *p) { return V3(p[0], p[1], p[2]); }

// geometry part of AreTwoPlanesPenetrable (PLADE/util.cpp:1295-1373), eight lanes per (hypothesis, plane, plane)
// triple: lane e of the group intersects the plane/plane line with rectangle edge e (0-3: first rectangle, 4-7:
// second) -- the expensive part, a restated 9x9 SVD solve per edge -- then the leader lane finishes the scalar
// logic exactly as the one-thread version did.  Returns 1 (in every lane) when the sampling passes are needed.
__device__ int segment_of_pair8(const float plane1[4], const float plane2[4], const V3 c1[4], const V3 c2[4], int e, unsigned int gshift,
                                V3 &start, V3 &direc, float &length) {
  V3 lineVec, linePoint;
  const int have_line = 0 == plane_intersection_line(plane1, plane2, lineVec, linePoint);   // same inputs in all 8 lanes
  bool hit = false;
  V3 ip;
  if (have_line) {
    const V3 *c = e < 4 ? c1 : c2;
    const int i = (e & 3) + 1;                              // the reference's loop index, 1..4
    V3 tl = c[i % 4] - c[(i - 1) % 4];
    normalize(tl);
    if (0 == line_line_point_cv(lineVec, linePoint, tl, c[i - 1], ip)) hit = !(dot(c[(i - 1) % 4] - ip, c[i % 4] - ip) > 0);
  }
  const unsigned int all = __ballot_sync(0xffffffffu, hit);
  const unsigned int m1 = (all >> gshift) & 0xFu, m2 = (all >> (gshift + 4)) & 0xFu;
  // the two intersection points of each rectangle, in edge order (ip1[0], ip1[1], ip2[0], ip2[1] of the reference)
  const int a0 = __ffs(m1) - 1, a1 = __ffs(m1 & (m1 - 1)) - 1, b0 = __ffs(m2) - 1, b1 = __ffs(m2 & (m2 - 1)) - 1;
  const int src_lane[4] = {(int) gshift + max(a0, 0), (int) gshift + max(a1, 0), (int) gshift + 4 + max(b0, 0), (int) gshift + 4 + max(b1, 0)};
  V3 inter[4];
  for (int k = 0; k < 4; ++k) {
    inter[k].x = __shfl_sync(0xffffffffu, ip.x, src_lane[k]);
    inter[k].y = __shfl_sync(0xffffffffu, ip.y, src_lane[k]);
    inter[k].z = __shfl_sync(0xffffffffu, ip.z, src_lane[k]);
  }
  if (!have_line) return 0;
  if (__popc(m1) != 2 || __popc(m2) != 2) return 0;      // empty -> not penetrable; any other count -> the reference returns -1
  direc = inter[1] - inter[0];
  normalize(direc);
  float len[4];
  int idx[4];
  for (int i = 0; i < 4; ++i) { len[i] = dot(inter[i] - inter[0], direc); idx[i] = i; }
  // std::sort of 4 LENGTHINDEX with myCompareLess == insertion sort (stable for ties)
  for (int i = 1; i < 4; ++i) {
    float l = len[i];
    int id = idx[i], j = i - 1;
    while (j >= 0 && l < len[j]) { len[j + 1] = len[j]; idx[j + 1] = idx[j]; --j; }
    len[j + 1] = l;
    idx[j + 1] = id;
  }
  if (0 == (idx[0] / 2 - idx[1] / 2)) return 0;
  start = inter[idx[1]];
  V3 endp = inter[idx[2]];
  length = norm(endp - start);
  return 1;
}

__global__ void __launch_bounds__(128) pen_geometry_kernel(PenArgs a, Triple *__restrict__ out, int *__restrict__ n_out, int *__restrict__ overflow) {
  const long long gtid = (long long) blockIdx.x * blockDim.x + threadIdx.x;
  const long long t = gtid >> 3;                          // triple of this 8-lane group
  const int e = (int) (gtid & 7);
  const unsigned int gshift = (threadIdx.x & 31) & ~7u;
  const long long total = (long long) a.H * a.Ps * a.Pt;
  // (no early return before the warp-wide ballot / shuffles: out-of-range and gated-out groups stay in the warp, idle)
  bool active = t < total;
  int j1 = 0, i1 = 0, h = 0;
  if (active) {
    j1 = (int) (t % a.Pt); i1 = (int) ((t / a.Pt) % a.Ps); h = (int) (t / ((long long) a.Pt * a.Ps));
    if (a.off_s[i1 + 1] == a.off_s[i1] || a.off_t[j1 + 1] == a.off_t[j1]) active = false;   // no corner points -> returns -1
  }
  M3 R;
  V3 T;
  float plane1[4] = {0.f, 0.f, 1.f, 0.f}, plane2[4] = {0.f, 0.f, 1.f, 0.f};
  V3 c1[4], c2[4];
  if (active) {
    for (int k = 0; k < 9; ++k) R.m[k] = a.hyp[12 * h + k];
    T = V3(a.hyp[12 * h + 9], a.hyp[12 * h + 10], a.hyp[12 * h + 11]);
    float4 ps = a.planes_s[i1], pt = a.planes_t[j1];
    V3 pn = mul(R, V3(ps.x, ps.y, ps.z));
    plane1[0] = pn.x; plane1[1] = pn.y; plane1[2] = pn.z; plane1[3] = -(-ps.w + dot(pn, T));
    plane2[0] = pt.x; plane2[1] = pt.y; plane2[2] = pt.z; plane2[3] = pt.w;
    V3 cc = mul(R, ld3(a.center_s + 3 * i1)) + T;
    V3 plane_A(pt.x, pt.y, pt.z);
    // PLADE/util.cpp:487-492 (note: the dot product is compared with the ANGLE threshold there)
    double c2p = (fabsf(dot(plane_A, cc) + pt.w) + fabsf(dot(pn, ld3(a.center_t + 3 * j1)) + plane1[3])) / 2;
    if (c2p < a.lengthThreshold && dot(pn, plane_A) > a.angleThreshold) active = false;
  }
  if (active)
    for (int k = 0; k < 4; ++k) { c1[k] = xform(R, T, ld3(a.corners_s + 12 * i1 + 3 * k)); c2[k] = ld3(a.corners_t + 12 * j1 + 3 * k); }
  V3 start, direc;
  float length = 0.f;
  // inactive groups run the edge step on a dummy pair of parallel planes (no intersection line: a cheap early out)
  const int need = segment_of_pair8(plane1, plane2, c1, c2, e, gshift, start, direc, length);
  if (!active || !need || e != 0) return;
  int nsteps = 0;
  while (nsteps < kMaxSteps && a.dist_table[nsteps] < length) ++nsteps;   // for (dist = 0; dist < length; dist += searchRadius)
  if (nsteps == 0) return;                // no sample: both counts stay 0 -> not penetrable
  if (nsteps >= kMaxSteps && a.dist_table[kMaxSteps - 1] < length) { atomicExch(overflow, 1); }
  int slot = atomicAdd(n_out, 1);
  Triple tr;
  tr.h = h; tr.i1 = i1; tr.j1 = j1; tr.nsteps = nsteps;
  tr.start[0] = start.x; tr.start[1] = start.y; tr.start[2] = start.z;
  tr.direc[0] = direc.x; tr.direc[1] = direc.y; tr.direc[2] = direc.z;
  for (int k = 0; k < 4; ++k) tr.plane1[k] = plane1[k];
  out[slot] = tr;
}

__global__ void __launch_bounds__(kPenThreads)
pen_sample_kernel(PenArgs a, const Triple *__restrict__ triples, int n_triples, int *__restrict__ pen) {
  __shared__ int cnt[kMaxSteps];
  __shared__ int s_pos, s_neg, s_skip;
  for (int w = blockIdx.x; w < n_triples; w += gridDim.x) {
    const Triple tr = triples[w];
    // hypothesis already dropped by another pair? (read once and broadcast so the skip is block-uniform)
    __syncthreads();
    if (threadIdx.x == 0) s_skip = pen[tr.h];
    __syncthreads();
    if (s_skip) continue;
    M3 R;
    for (int k = 0; k < 9; ++k) R.m[k] = a.hyp[12 * tr.h + k];
    const V3 T(a.hyp[12 * tr.h + 9], a.hyp[12 * tr.h + 10], a.hyp[12 * tr.h + 11]);
    const V3 start(tr.start[0], tr.start[1], tr.start[2]), direc(tr.direc[0], tr.direc[1], tr.direc[2]);
    const float4 pt = a.planes_t[tr.j1];
    const float r = a.searchRadius, rh = a.searchRadius / 2;
    const float r2 = (float) ((double) r * (double) r), rh2 = (float) ((double) rh * (double) rh);
    const float inv_r = 1.0f / r;
    const int sb = a.off_s[tr.i1], se = a.off_s[tr.i1 + 1], tb = a.off_t[tr.j1], te = a.off_t[tr.j1 + 1];
    // Conservative pre-test: a sample point start + d * direc can only be within radius rr of p if p is within rr of
    // the LINE.  The squared distance to the line is evaluated in double (no cancellation at the 1e-5 scale of rr^2)
    // and compared with a 1 % margin, far above the float rounding of the sample positions, so every point the
    // exact float tests below would accept still reaches them; most plane points are rejected here.
    const double sdx = start.x, sdy = start.y, sdz = start.z, ddx = direc.x, ddy = direc.y, ddz = direc.z;
    const double inv_dd = 1.0 / (ddx * ddx + ddy * ddy + ddz * ddz);
    auto far_from_line = [&](const V3 &p, double lim) {
      const double vx = (double) p.x - sdx, vy = (double) p.y - sdy, vz = (double) p.z - sdz;
      const double t = vx * ddx + vy * ddy + vz * ddz;
      return (vx * vx + vy * vy + vz * vz) - t * t * inv_dd > lim;
    };
    const double lim_gate = 1.01 * (double) rh2, lim_probe = 1.01 * (double) r2;
    bool penetrable = true;
    for (int pass = 0; pass < 2 && penetrable; ++pass) {
      // pass 0: gate = target plane cloud, probe = transformed source plane cloud, classified against plane2
      // pass 1: roles swapped, classified against plane1
      __syncthreads();
      for (int k = threadIdx.x; k < tr.nsteps; k += kPenThreads) cnt[k] = 0;
      if (threadIdx.x == 0) { s_pos = 0; s_neg = 0; }
      __syncthreads();
      const bool gate_is_tgt = pass == 0;
      const int gb = gate_is_tgt ? tb : sb, ge = gate_is_tgt ? te : se;
      for (int i = gb + threadIdx.x; i < ge; i += kPenThreads) {
        float4 q = gate_is_tgt ? a.pts_t[i] : a.pts_s[i];
        V3 p(q.x, q.y, q.z);
        if (!gate_is_tgt) p = xform(R, T, p);
        if (far_from_line(p, lim_gate)) continue;
        int k0 = (int) floorf(dot(p - start, direc) * inv_r);
        for (int k = k0 - 1; k <= k0 + 2; ++k) {
          if (k < 0 || k >= tr.nsteps) continue;
          V3 sp = start + a.dist_table[k] * direc;
          if (l2simple(sp, p) < rh2) atomicAdd(&cnt[k], 1);
        }
      }
      __syncthreads();
      const int qb = gate_is_tgt ? sb : tb, qe = gate_is_tgt ? se : te;
      const float pl0 = gate_is_tgt ? pt.x : tr.plane1[0], pl1 = gate_is_tgt ? pt.y : tr.plane1[1],
                  pl2 = gate_is_tgt ? pt.z : tr.plane1[2], pl3 = gate_is_tgt ? pt.w : tr.plane1[3];
      int pos = 0, neg = 0;
      for (int i = qb + threadIdx.x; i < qe; i += kPenThreads) {
        float4 q = gate_is_tgt ? a.pts_s[i] : a.pts_t[i];
        V3 p(q.x, q.y, q.z);
        if (gate_is_tgt) p = xform(R, T, p);
        if (far_from_line(p, lim_probe)) continue;
        int k0 = (int) floorf(dot(p - start, direc) * inv_r);
        bool hit = false;
        for (int k = k0 - 2; k <= k0 + 3 && !hit; ++k) {
          if (k < 0 || k >= tr.nsteps || cnt[k] < 2) continue;
          V3 sp = start + a.dist_table[k] * direc;
          hit = l2simple(sp, p) < r2;
        }
        if (hit) {
          float td = pl0 * p.x + pl1 * p.y + pl2 * p.z + pl3;
          if (fabsf(td) > a.minDistance) { if (td >= 0) ++pos; else ++neg; }
        }
      }
      pos = __reduce_add_sync(0xffffffffu, pos);
      neg = __reduce_add_sync(0xffffffffu, neg);
      if ((threadIdx.x & 31) == 0) { if (pos) atomicAdd(&s_pos, pos); if (neg) atomicAdd(&s_neg, neg); }
      __syncthreads();
      const int P = s_pos, N = s_neg;
      if (pass == 0) { if (P < a.minPoints || N < a.minPoints) penetrable = false; }
      else { if (P < a.minPoints && N < a.minPoints) penetrable = false; }
      if (penetrable && (double) max(P, N) / min(P, N + 1) > 5) penetrable = false;
    }
    if (penetrable && threadIdx.x == 0) atomicOr(&pen[tr.h], 1);
  }
}

}  // namespace

void penetration_filter(Device &dev, PenScratch &sc, const PenSide &src, const PenSide &tgt, const float *h_hyp12, int H,
                        float lengthThreshold, float angleThreshold, std::vector<unsigned char> &pen_out) {
  pen_out.assign(H, 0);
  const int Ps = (int) src.planes.size(), Pt = (int) tgt.planes.size();
  if (H == 0 || Ps == 0 || Pt == 0) return;
  cudaStream_t s = dev.stream;
  // --- small tables -> one upload
  const float searchRadius = lengthThreshold;                       // PLADE/util.cpp:494
  const float minDistance = (float) ((double) lengthThreshold / 2);
  std::vector<float> tab(kMaxSteps);
  {
    float d = 0;
    for (int k = 0; k < kMaxSteps; ++k) { tab[k] = d; d += searchRadius; }
  }
  size_t nf = 0;
  auto reserve = [&](size_t n) { size_t o = nf; nf += (n + 3) & ~size_t(3); return o; };
  size_t o_ps = reserve(4 * Ps), o_pt = reserve(4 * Pt), o_cs = reserve(12 * Ps), o_ct = reserve(12 * Pt), o_es = reserve(3 * Ps),
         o_et = reserve(3 * Pt), o_os = reserve(Ps + 1), o_ot = reserve(Pt + 1), o_h = reserve(12 * (size_t) H), o_tab = reserve(kMaxSteps);
  std::vector<float> host(nf, 0.f);
  for (int i = 0; i < Ps; ++i) {
    memcpy(&host[o_ps + 4 * i], src.planes[i].data(), 16);
    memcpy(&host[o_cs + 12 * i], src.corners4[i].data(), 48);
    memcpy(&host[o_es + 3 * i], &src.center[i], 12);
  }
  for (int j = 0; j < Pt; ++j) {
    memcpy(&host[o_pt + 4 * j], tgt.planes[j].data(), 16);
    memcpy(&host[o_ct + 12 * j], tgt.corners4[j].data(), 48);
    memcpy(&host[o_et + 3 * j], &tgt.center[j], 12);
  }
  memcpy(&host[o_os], src.ds_start.data(), sizeof(int) * (Ps + 1));
  memcpy(&host[o_ot], tgt.ds_start.data(), sizeof(int) * (Pt + 1));
  memcpy(&host[o_h], h_hyp12, sizeof(float) * 12 * (size_t) H);
  memcpy(&host[o_tab], tab.data(), sizeof(float) * kMaxSteps);
  float *d = sc.tables.ensure(nf);
  PLADE_CUDA(cudaMemcpyAsync(d, host.data(), sizeof(float) * nf, cudaMemcpyHostToDevice, s));
  PenArgs a;
  a.planes_s = reinterpret_cast<const float4 *>(d + o_ps); a.planes_t = reinterpret_cast<const float4 *>(d + o_pt);
  a.corners_s = d + o_cs; a.corners_t = d + o_ct; a.center_s = d + o_es; a.center_t = d + o_et;
  a.off_s = reinterpret_cast<const int *>(d + o_os); a.off_t = reinterpret_cast<const int *>(d + o_ot);
  a.pts_s = src.d_pts; a.pts_t = tgt.d_pts;
  a.hyp = d + o_h; a.dist_table = d + o_tab;
  a.H = H; a.Ps = Ps; a.Pt = Pt;
  a.lengthThreshold = lengthThreshold; a.angleThreshold = angleThreshold; a.searchRadius = searchRadius; a.minDistance = minDistance;
  a.minPoints = 10;
  const long long total = (long long) H * Ps * Pt;
  Triple *d_tr = reinterpret_cast<Triple *>(sc.triples.ensure((size_t) total * sizeof(Triple)));
  int *d_flags = sc.flags.ensure((size_t) H + 8);
  PLADE_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * ((size_t) H + 8), s));
  int *d_n = d_flags + H, *d_ovf = d_flags + H + 1;
  pen_geometry_kernel<<<div_up(total * 8, 128), 128, 0, s>>>(a, d_tr, d_n, d_ovf);
  PLADE_LAUNCH_CHECK();
  dev.launches.add();
  int h2[2];
  PLADE_CUDA(cudaMemcpyAsync(h2, d_n, sizeof(h2), cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  if (h2[1]) throw std::runtime_error("penetration filter: segment longer than 4096 search radii");
  if (getenv("PLADE_TIMING")) fprintf(stderr, "[plade penetration] hypotheses %d, planes %d x %d, (hypothesis, plane, plane) triples to sample: %d\n", H, Ps, Pt, h2[0]);
  if (h2[0] > 0) {
    int blocks = std::min(h2[0], dev.num_sms * 8);
    pen_sample_kernel<<<blocks, kPenThreads, 0, s>>>(a, d_tr, h2[0], d_flags);
    PLADE_LAUNCH_CHECK();
    dev.launches.add();
  }
  std::vector<int> flags(H);
  PLADE_CUDA(cudaMemcpyAsync(flags.data(), d_flags, sizeof(int) * H, cudaMemcpyDeviceToHost, s));
  stream_sync(s);
  for (int h = 0; h < H; ++h) pen_out[h] = flags[h] != 0;
}

}  // namespace plade
