// Internal C++ interface between the host pipeline and the sm_100a kernels.
// Every stage cites the reference loop it replaces (paths relative to /root/reference/code/).
#pragma once
#include "common.h"
#include "linalg.h"
#include <array>

namespace plade {

// CUDA-event stopwatch around selected launches of one stream.  begin/end only record events (no host wait);
// collect() is called once the stream is idle and adds the elapsed device times to the totals that bench.py
// reads through plade_kernel_times (the roofline figure of the dominant kernel is measured live this way).
struct KernelClock {
  enum { kScoreCandidates = 0, kRefineCluster = 1, kBandCompact = 2, kKinds = 3 };
  struct Span { cudaEvent_t a = nullptr, b = nullptr; int kind = 0; };
  std::vector<Span> spans;
  size_t used = 0;
  double ms[kKinds] = {0}, bytes[kKinds] = {0};
  long long launches[kKinds] = {0};
  // off by default: two event records per timed launch are two more driver calls on a path that is bound by the
  // number of driver calls (plade_set_param("kernel_clock", 1) turns the clock on for the calls that want the figures)
  bool enabled = false;
  void begin(int kind, double algorithmic_bytes, cudaStream_t s) {
    if (!enabled) return;
    if (used == spans.size()) {
      Span sp;
      PLADE_CUDA(cudaEventCreate(&sp.a));
      PLADE_CUDA(cudaEventCreate(&sp.b));
      spans.push_back(sp);
    }
    spans[used].kind = kind;
    bytes[kind] += algorithmic_bytes;
    PLADE_CUDA(cudaEventRecord(spans[used].a, s));
  }
  void end(cudaStream_t s) { if (!enabled) return; PLADE_CUDA(cudaEventRecord(spans[used].b, s)); ++used; }
  void collect() {          // the stream must be idle
    for (size_t i = 0; i < used; ++i) {
      float t = 0;
      if (cudaEventElapsedTime(&t, spans[i].a, spans[i].b) == cudaSuccess) { ms[spans[i].kind] += t; ++launches[spans[i].kind]; }
    }
    used = 0;
  }
  void reset() { used = 0; for (int k = 0; k < kKinds; ++k) { ms[k] = bytes[k] = 0; launches[k] = 0; } }
  void destroy() { for (Span &sp : spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); } spans.clear(); used = 0; }
};

struct Device {
  int id = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  LaunchCounter launches;
  KernelClock clock;
};

// ------------------------------------------------------------------------------------------------
// K5 — hypothesis verification (PLADE/plade.cpp:547-564 + ComputeOverlap PLADE/util.h:612-647)
// ------------------------------------------------------------------------------------------------
struct HypParams {       // 64 B, one per hypothesis
  float R[9];            // row-major
  float T[3];
  float c[3];            // ball centre = R * source_bbox_centre + T, computed by the caller
  float pad;
};

struct TargetGrid {
  DevBuf<float4> pts;        // target ds points sorted by cell
  DevBuf<int> cell_start;    // ncells + 1
  DevBuf<unsigned int> keys, keys_alt;
  DevBuf<int> order, order_alt;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<float> mm;          // 6 ordered-int encoded bbox values
  DevBuf<unsigned int> occ_raw, occ_dil;   // occupancy bits over the grid extended by one cell, and its 3x3x3 dilation
  int ey = 0, ewords = 0;
  float minx = 0, miny = 0, minz = 0, inv_cell = 0, cell = 0;
  int nx = 0, ny = 0, nz = 0;
  size_t n = 0;
};

// Build the uniform grid (cell edge >= inlier_dist) over n target points (device float4, w ignored).
void build_target_grid(Device &dev, const float4 *d_tgt, size_t n, float inlier_dist, TargetGrid &grid);

// counts[h] = #{ s : exists t in ball(c_h, ball_radius) with dist2(transform_h(s), t) < inlier_dist^2 }
// radii follow kdtree_flann.hpp:193: r2 = float(double(r) * double(r)), strict '<'.
void verify_hypotheses(Device &dev, const float4 *d_src, size_t ns, const TargetGrid &grid,
                       const HypParams *d_hyp, int H, float ball_radius, float inlier_dist,
                       unsigned int *d_counts);

// best hypothesis of a shard as a packed u64 key in d_key[0] (see shard_key_kernel, verify.cu); entry i = hypothesis rank + i * world
void shard_best_key(Device &dev, const unsigned int *d_counts, const HypParams *d_hyp, int n_mine, int rank, int world, double denom,
                    double n_src_planes, int mode, unsigned long long *d_key);

// ------------------------------------------------------------------------------------------------
// K2b — voxel-grid down-sampling (pcl VoxelGrid::applyFilter, filters/impl/voxel_grid.hpp:214-437)
// K2a — average spacing (PLADE/util.cpp:1619-1648)
// ------------------------------------------------------------------------------------------------
struct VoxelScratch {
  DevBuf<unsigned long long> keys, keys_alt;
  DevBuf<int> idx, idx_alt;
  DevBuf<int> seg_start;
  DevBuf<int> flags;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<float> minmax;   // 6 floats
  DevBuf<int> counter;
};

// Down-sample `n` points (device float4 xyz_, optional gather list `d_sel` of m indices, or nullptr
// for all n) with leaf size `leaf`.  Output centroids (float4, w = 0) in PCL's output order
// (voxel index ascending, i.e. z-major / y / x-minor).  Returns the number of voxels.
// If group != nullptr, group[i] (>= 0) partitions the selected points into independent clouds that
// are voxelised separately in one pass; out_group_start receives ngroups+1 offsets (host).
size_t voxel_downsample(Device &dev, VoxelScratch &sc, const float4 *d_pts, size_t n, float leaf,
                        DevBuf<float4> &out);
size_t voxel_downsample_groups(Device &dev, VoxelScratch &sc, const float4 *d_pts, size_t n,
                               const int *d_group, int ngroups, float leaf, DevBuf<float4> &out,
                               std::vector<int> &out_group_start);

// k nearest squared distances (ascending, float, FLANN L2_Simple arithmetic) of nq query points
// (indices into d_pts) against all n points.  d_out: nq * k floats.
struct KnnScratch {      // grow-only scratch of knn_sqdist (owned by the context: nothing is allocated per call or per thread)
  DevBuf<float> partial;
  DevBuf<int> bbox_buf, idx_a, idx_b;
  DevBuf<unsigned int> key_a, key_b;
  DevBuf<float4> sorted_pts;
  DevBuf<unsigned char> cub_tmp;
};
void knn_sqdist(Device &dev, KnnScratch &ks, const float4 *d_pts, size_t n, const int *d_query_idx, int nq, int k,
                float *d_out);

// ------------------------------------------------------------------------------------------------
// K3c — descriptor radius matching (KdTreeSearchNDim<.,8>::find_neighbors, ANN.h:979-1029)
// ------------------------------------------------------------------------------------------------
struct MatchScratch {
  DevBuf<unsigned long long> sort_a, sort_b;
  DevBuf<float> db, q;
  DevBuf<int> counts, offsets;
  DevBuf<int> out_idx, out_idx_alt;
  DevBuf<unsigned long long> out_key, out_key_alt;
  DevBuf<int> out_q, out_q_alt;
  DevBuf<unsigned char> cub_tmp;
};
// Host in / host out.  offsets[nq+1]; idx/dist2 sorted per query by (dist2 asc, idx asc).
size_t match_descriptors(Device &dev, MatchScratch &sc, const float *h_db, int ndb, const float *h_q,
                         int nq, float radius, std::vector<int> &offsets, std::vector<int> &idx,
                         std::vector<double> &dist2);

// ------------------------------------------------------------------------------------------------
// K4a/b/c — hypothesis generation
// ------------------------------------------------------------------------------------------------
struct MatchPairIn {    // one (query pair, db pair) match, PLADE/util.cpp:316-322
  float sv1[3], sv2[3], dv1[3], dv2[3], sp[3], tp[3];
};
struct RigidOut { float R[9]; float T[3]; float euler[3]; float pad; };

struct HypScratch {
  DevBuf<unsigned long long> key64_a, key64_b;
  DevBuf<MatchPairIn> in;
  DevBuf<RigidOut> rt;
  DevBuf<int> label, cell_key, cell_order, cell_start;
  DevBuf<unsigned int> keys, keys_alt;
  DevBuf<int> order_alt;
  DevBuf<unsigned char> cub_tmp;
  DevBuf<int> misc;
};
// K4a: R from umeyama({v1,v2,v1xv2} -> {w1,w2,w1xw2}), T = tp - R*sp, euler as pcl::getEulerAngles.
void transforms_from_matches(Device &dev, HypScratch &sc, const MatchPairIn *h_in, size_t m,
                             std::vector<RigidOut> &out);
// K4b: connected components of {(a,b): |Ta-Tb|^2 < dist_thresh^2 (float), |euler_a-euler_b|^2 < ang_thresh}.
// label[i] = smallest member index of i's component.
void cluster_transforms(Device &dev, HypScratch &sc, const std::vector<RigidOut> &rt, float dist_thresh,
                        float ang_thresh, std::vector<int> &label);

// ------------------------------------------------------------------------------------------------
// K4d — penetration filter (PLADE/util.cpp:450-519, AreTwoPlanesPenetrable PLADE/util.cpp:1279-1458)
// ------------------------------------------------------------------------------------------------
struct PenSide {                                   // per-plane data of one cloud
  std::vector<std::array<float, 4>> planes;        // (n, d)
  std::vector<std::array<V3, 4>> corners4;         // bounding rectangle projected on the plane
  std::vector<V3> center;
  std::vector<int> ds_start;                       // P + 1 offsets into d_pts
  const float4 *d_pts = nullptr;                   // device: per-plane down-sampled points
};
struct PenScratch {
  DevBuf<float> tables;
  DevBuf<unsigned char> triples;
  DevBuf<int> flags;
};
// pen_out[h] = 1 iff hypothesis h (12 floats: R row-major, T) makes some source plane penetrate a
// non-coincident target plane.
void penetration_filter(Device &dev, PenScratch &sc, const PenSide &src, const PenSide &tgt, const float *h_hyp12, int H,
                        float lengthThreshold, float angleThreshold, std::vector<unsigned char> &pen_out);


// ------------------------------------------------------------------------------------------------
// Oriented bounding boxes (ComputeBoundingBox, PLADE/util.h:187-248) of many point segments at once
// ------------------------------------------------------------------------------------------------
struct ObbSeg { const float4 *p; int n; int pad; };
struct ObbResult {
  int rc = -1;
  V3 center;
  double width = 0, height = 0, depth = 0;
  V3 corners[8];
};
struct ObbScratch {
  DevBuf<unsigned char> segs, blocks;
  DevBuf<double> acc;
  DevBuf<float> frames;
  DevBuf<int> mm;
};
void obb_segments(Device &dev, ObbScratch &sc, const std::vector<ObbSeg> &segs, std::vector<ObbResult> &out);

// ------------------------------------------------------------------------------------------------
// Closest points of line pairs through the reference's own float Jacobi-SVD solve
// (ComputeNearstTwoPointsOfTwo3DLine, PLADE/util.cpp:1167-1229), one thread per pair.
// h_in12: v1 p1 v2 p2 per pair (directions normalised); h_out6: point1 point2 per pair.
// ------------------------------------------------------------------------------------------------
struct SvdScratch { DevBuf<float> in, out; };
void nearest_points_batch(Device &dev, SvdScratch &sc, const float *h_in12, int n, float *h_out6);

}  // namespace plade
