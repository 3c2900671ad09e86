// Host orchestrator: restates the control flow of registration() (PLADE/plade.cpp:31-580) and
// MatchingLines() (PLADE/util.cpp:31-520) on flat arrays, with the data-parallel stages on the GPU.
// Order-defining host logic (match order, cluster seed order, the two std::sort calls with their
// comparators, the candidate budget) follows the reference statement by statement because it
// decides WHICH hypothesis wins; file:line citations are relative to /root/reference/code/.
#include "pipeline.h"
#include <atomic>
#include <sched.h>
#include "nccl_shard.h"
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <numeric>
#include <condition_variable>
#include <mutex>
#include <sstream>
#include <thread>
#include <exception>

namespace plade {

namespace {

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Line {            // INTERSECTION_LINE, PLADE/util.h:70-78
  V3 vec, pt;
  int p0, p1;
};
struct NearPts {         // NearstPointsTwoLine, PLADE/util.h:61-68
  V3 a, b;
  double length;
};
struct LenIdx {          // LENGTHINDEX, PLADE/util.h:347-350
  float length;
  int index;
};
inline bool cmp_greater(const LenIdx &a, const LenIdx &b) { return a.length > b.length; }    // util.h:360

struct Side {            // MatchInformation, PLADE/util.h:80-102
  size_t n_ds = 0;
  std::vector<float4> ds;
  V3 center;
  double radius = 0;
  std::vector<std::array<float, 4>> planes;
  std::vector<int> plane_ds_start;          // P+1
  std::vector<float4> plane_ds;
  std::vector<std::array<V3, 4>> corners4;
  std::vector<V3> plane_center;
  std::vector<float> plane_radius;
  std::vector<Line> lines;
};

// ComputeNearstTwoPointsOfTwo3DLine (PLADE/util.cpp:1167-1229), split in two: the host part normalises both
// directions IN PLACE and rejects identical directions (util.cpp:1170-1176); the 9x9 float SVD solve of all
// accepted pairs then runs as one batch on the GPU (svdsolve.cu), and length = |p1 - p2| (float norm).
struct PairSolve {
  std::vector<float> in, out;    // 12 / 6 floats per accepted pair
  int count = 0;
  int add(V3 &v1, const V3 &p1, V3 &v2, const V3 &p2) {
    normalize(v1);
    normalize(v2);
    if (v1.x == v2.x && v1.y == v2.y && v1.z == v2.z) return -1;
    const float q[12] = {v1.x, v1.y, v1.z, p1.x, p1.y, p1.z, v2.x, v2.y, v2.z, p2.x, p2.y, p2.z};
    in.insert(in.end(), q, q + 12);
    return count++;
  }
  void get(int k, V3 &a, V3 &b, double &len) const {
    if (k < 0) { len = -1; return; }
    const float *o = &out[6 * (size_t) k];
    a = V3(o[0], o[1], o[2]);
    b = V3(o[3], o[4], o[5]);
    len = norm(a - b);
  }
};

// ComputeDescriptorVectorForPairLines, method22 (PLADE/util.cpp:533-602)
void pair_descriptor(const V3 &l1, const V3 &l2, const V3 &l1sp1, const V3 &l1sp2, const V3 &l2sp1, const V3 &l2sp2,
                     float *d, V3 &newLine1, V3 &newLine2) {
  float angle1 = std::fabs(dot(l1, l2sp1)), angle2 = std::fabs(dot(l1, l2sp2));
  V3 n2a, n2b;
  if (angle1 <= angle2) { n2a = l2sp1; n2b = l2sp2; } else { n2a = l2sp2; n2b = l2sp1; }
  newLine2 = cross(n2a, n2b);
  V3 n1a, n1b;
  angle1 = std::fabs(dot(l2, l1sp1));
  angle2 = std::fabs(dot(l2, l1sp2));
  if (angle1 <= angle2) { n1a = l1sp1; n1b = l1sp2; } else { n1a = l1sp2; n1b = l1sp1; }
  newLine1 = cross(n1a, n1b);
  d[1] = dot(newLine1, newLine2);
  d[2] = dot(n1a, n1b);
  d[3] = dot(n2a, n2b);
  d[4] = dot(newLine1, n2a);
  d[5] = dot(newLine1, n2b);
  d[6] = dot(newLine2, n1a);
  d[7] = dot(newLine2, n1b);
}

}  // namespace

namespace {
std::atomic<int> g_blocking_sync{[] { const char *e = getenv("PLADE_BLOCKING_SYNC"); const int v = e ? atoi(e) : 0; return v < 0 ? 0 : v > 2 ? 2 : v; }()};
}
void set_blocking_sync(int mode) { g_blocking_sync.store(mode < 0 ? 0 : mode > 2 ? 2 : mode); }
void stream_sync(cudaStream_t s) {
  const int mode = g_blocking_sync.load(std::memory_order_relaxed);
  if (mode == 0) { PLADE_CUDA(cudaStreamSynchronize(s)); return; }
  if (mode == 2) {
    // poll and yield: the waiting thread gives its core to any other runnable thread between two polls (the waits of this
    // pipeline are 0.1-0.5 ms: too short for a sleeping wait to be cheap, too many for a spinning wait to be fair)
    for (;;) {
      const cudaError_t q = cudaStreamQuery(s);
      if (q == cudaSuccess) return;
      if (q != cudaErrorNotReady) PLADE_CUDA(q);
      sched_yield();
    }
  }
  thread_local cudaEvent_t ev[64] = {};
  int d = 0;
  PLADE_CUDA(cudaGetDevice(&d));
  if (d < 0 || d >= 64) { PLADE_CUDA(cudaStreamSynchronize(s)); return; }
  if (!ev[d]) PLADE_CUDA(cudaEventCreateWithFlags(&ev[d], cudaEventBlockingSync | cudaEventDisableTiming));
  PLADE_CUDA(cudaEventRecord(ev[d], s));
  PLADE_CUDA(cudaEventSynchronize(ev[d]));
}

Registrar::Registrar(int device) {
  if (device >= 0) PLADE_CUDA(cudaSetDevice(device));
  PLADE_CUDA(cudaGetDevice(&dev.id));
  cudaDeviceProp prop;
  PLADE_CUDA(cudaGetDeviceProperties(&prop, dev.id));
  dev.num_sms = prop.multiProcessorCount;
  PLADE_CUDA(cudaStreamCreateWithFlags(&dev.stream, cudaStreamNonBlocking));
  dev2 = dev;
  dev2.launches.n = 0;
  PLADE_CUDA(cudaStreamCreateWithFlags(&dev2.stream, cudaStreamNonBlocking));
  PLADE_CUDA(cudaEventCreate(&ev0));
  PLADE_CUDA(cudaEventCreate(&ev1));
  PLADE_CUDA(cudaEventCreate(&ev_user0));
  PLADE_CUDA(cudaEventCreate(&ev_user1));
}

void Registrar::mark(const char *name) {
  static const bool on = getenv("PLADE_TIMING") != nullptr;
  if (on) marks.emplace_back(name, now_s());
}

void Registrar::print_marks() {
  if (marks.size() < 2) { marks.clear(); return; }
  std::map<std::string, double> acc;
  std::vector<std::string> order;
  for (size_t i = 1; i < marks.size(); ++i) {
    if (!acc.count(marks[i].first)) order.push_back(marks[i].first);
    acc[marks[i].first] += marks[i].second - marks[i - 1].second;
  }
  fprintf(stderr, "[plade timing, ms]");
  for (const std::string &k : order) fprintf(stderr, " %s=%.2f", k.c_str(), acc[k] * 1e3);
  fprintf(stderr, "\n");
  marks.clear();
}

Registrar::~Registrar() {
  if (nccl) { nccl_comm_destroy(nccl); nccl = nullptr; }
  free_ransac_scratch(*this);
  for (cudaEvent_t e : {ev0, ev1, ev_user0, ev_user1}) if (e) cudaEventDestroy(e);
  dev.clock.destroy();
  dev2.clock.destroy();
  if (dev.stream) cudaStreamDestroy(dev.stream);
  if (dev2.stream) cudaStreamDestroy(dev2.stream);
}

// defined in split.cu
void split_cloud(Device &dev, const float *d_xyzn, size_t n, float4 *pos, float4 *nrm);

void Registrar::prepare_side(int side, const CloudDev &C, const std::vector<PlaneParam> &P, const int *d_group, float leaf, SidePrep &out) {
  Device &d = side == 0 ? dev : dev2;
  VoxelScratch &vs = side == 0 ? vox : vox2;
  DevBuf<float4> &ds_all = side == 0 ? ds_tgt : ds_src, &ds_pl = side == 0 ? ds_planes_t : ds_planes_s;
  out.ready = false;
  out.leaf = leaf;
  out.n_ds = voxel_downsample(d, vs, C.pos.p, C.n, leaf, ds_all);
  if (side == 0) mark("voxel_full");
  out.n_plane_ds = voxel_downsample_groups(d, vs, C.pos.p, C.n, d_group, (int) P.size(), leaf, ds_pl, out.plane_ds_start);
  if (side == 0) mark("voxel_planes");
  // oriented bounding boxes of the cloud and of every plane: one batch
  std::vector<ObbSeg> segs;
  segs.push_back({ds_all.p, (int) out.n_ds, 0});
  for (size_t i = 0; i < P.size(); ++i) segs.push_back({ds_pl.p + out.plane_ds_start[i], out.plane_ds_start[i + 1] - out.plane_ds_start[i], 0});
  obb_segments(d, side == 0 ? obb_sc : obb_sc2, segs, out.obb);
  if (side == 0) mark("obb");
  out.ready = true;
}

void Registrar::upload(const float *xyzn, size_t n, CloudDev &out, int lane, bool wait) {
  Device &d = lane == 0 ? dev : dev2;
  out.n = n;
  if (n == 0) return;
  float *d_in = upload_stage[lane].ensure(n * 6);
  PLADE_CUDA(cudaMemcpyAsync(d_in, xyzn, sizeof(float) * n * 6, cudaMemcpyHostToDevice, d.stream));
  split_cloud(d, d_in, n, out.pos.ensure(n), out.nrm.ensure(n));
  if (wait) stream_sync(d.stream);
}

// average_spacing (PLADE/util.cpp:1619-1648): k = 6, ~10000 strided samples
float Registrar::average_spacing(const CloudDev &c, int lane) {
  Device &dev = lane == 0 ? this->dev : this->dev2;      // shadows the member: everything below runs on the lane's stream
  const int k = 6, samples = 10000;
  size_t num = c.n;
  if (num == 0) return 0.f;
  size_t step = 1;
  if (num > (size_t) samples) step = num / samples;
  std::vector<int> q;
  for (size_t i = 0; i < num; i += step) q.push_back((int) i);
  int kk = (int) std::min<size_t>(k, num);
  int *d_q = qidx.ensure(q.size());
  float *d_o = knn_out.ensure(q.size() * kk);
  PLADE_CUDA(cudaMemcpyAsync(d_q, q.data(), sizeof(int) * q.size(), cudaMemcpyHostToDevice, dev.stream));
  knn_sqdist(dev, knn_sc, c.pos.p, num, d_q, (int) q.size(), kk, d_o);
  std::vector<float> h(q.size() * kk);
  PLADE_CUDA(cudaMemcpyAsync(h.data(), d_o, sizeof(float) * h.size(), cudaMemcpyDeviceToHost, dev.stream));
  stream_sync(dev.stream);
  double total = 0.0;
  size_t total_count = 0;
  for (size_t s = 0; s < q.size(); ++s) {
    int nbs = kk;
    if (nbs <= 1) continue;
    double avg = 0.0;
    for (int i = 1; i < nbs; ++i) avg += std::sqrt(h[s * kk + i]);   // float sqrt, double accumulate
    total += (avg / nbs);
    ++total_count;
  }
  return static_cast<float>(total / total_count);
}

bool Registrar::register_host_clouds(const float *t, size_t nt, const float *s, size_t ns, CloudDev &dt, CloudDev &ds, float out16[16]) {
  // each lane uploads its own cloud on its own stream and goes straight on to extract its planes: the two
  // host-to-device copies share the link, and neither lane waits for the other's cloud
  pending_host[0] = t; pending_n[0] = nt;
  pending_host[1] = s; pending_n[1] = ns;
  dt.n = nt; ds.n = ns;
  struct Clear { Registrar &r; ~Clear() { r.pending_host[0] = r.pending_host[1] = nullptr; } } clear{*this};
  return register_clouds(dt, ds, out16);
}

bool Registrar::register_clouds(const CloudDev &tgt, const CloudDev &src, float out16[16]) {
  double t0 = now_s();
  dev.clock.reset();
  dev2.clock.reset();
  std::cout << "extracting planes for both point clouds...\n";
  // the two clouds are independent: the source's planes are extracted on a helper thread + stream while
  // this thread does the target's (the kernels are small, so the two lanes overlap on the GPU and the
  // host-side round trips of one lane hide behind the other's)
  std::vector<PlaneParam> tp, sp;
  std::exception_ptr helper_error;
  std::mutex spacing_mutex;
  std::condition_variable spacing_cv;
  bool spacing_ready = false, helper_failed = false;
  // the per-lane set-up is valid for this call only, whatever way it ends
  struct PrepReset {
    Registrar &r;
    explicit PrepReset(Registrar &reg) : r(reg) { clear(); }
    ~PrepReset() { clear(); }
    void clear() { r.prep[0].ready = r.prep[1].ready = false; r.lane_spacing = 0; }
  } prep_reset(*this);
  std::thread helper([&] {
    try {
      PLADE_CUDA(cudaSetDevice(dev.id));
      if (pending_host[1]) upload(pending_host[1], pending_n[1], const_cast<CloudDev &>(src), 1, false);
      // the leaf size of both voxel grids is 4 x the average spacing of the SOURCE cloud (PLADE/plade.cpp:41-45): this lane
      // computes it first, so that each lane can down-sample its own cloud as soon as its planes are known
      const float sp_avg = src.n ? average_spacing(src, 1) : 0.f;
      { std::lock_guard<std::mutex> l(spacing_mutex); lane_spacing = sp_avg; spacing_ready = true; }
      spacing_cv.notify_all();
      sp = extract_planes_dev(src, params.init_min_support, group_s, 1);
      if ((int) sp.size() >= params.min_planes && sp_avg * 4 > 0) prepare_side(1, src, sp, group_s.p, sp_avg * 4, prep[1]);
    } catch (...) {
      helper_error = std::current_exception();
      { std::lock_guard<std::mutex> l(spacing_mutex); helper_failed = true; }
      spacing_cv.notify_all();
    }
  });
  try {
    if (pending_host[0]) upload(pending_host[0], pending_n[0], const_cast<CloudDev &>(tgt), 0, false);
    tp = extract_planes_dev(tgt, params.init_min_support, group_t, 0);
    float sp_avg = 0;
    bool failed = false;
    {
      std::unique_lock<std::mutex> l(spacing_mutex);
      spacing_cv.wait(l, [&] { return spacing_ready || helper_failed; });
      sp_avg = lane_spacing;
      failed = helper_failed;       // read under the lock: the helper thread may still be writing it
    }
    if (!failed && (int) tp.size() >= params.min_planes && sp_avg * 4 > 0) prepare_side(0, tgt, tp, group_t.p, sp_avg * 4, prep[0]);
  } catch (...) { helper.join(); throw; }
  helper.join();
  dev.launches.n += dev2.launches.n;
  dev2.launches.n = 0;
  for (int k = 0; k < KernelClock::kKinds; ++k) {     // fold the helper lane's kernel times into the context's
    dev.clock.ms[k] += dev2.clock.ms[k]; dev.clock.bytes[k] += dev2.clock.bytes[k]; dev.clock.launches[k] += dev2.clock.launches[k];
  }
  if (helper_error) std::rethrow_exception(helper_error);
  if ((int) tp.size() < params.min_planes) {
    std::cerr << "too few (only " << tp.size() << ") planes extracted from the target point cloud" << std::endl;
    last_error = "too few planes in target";
    return false;
  }
  if ((int) sp.size() < params.min_planes) {
    std::cerr << "two few (only " << sp.size() << ") planes extracted from the source point cloud" << std::endl;
    last_error = "too few planes in source";
    return false;
  }
  times.planes = now_s() - t0;
  return register_core(tgt, src, tp, sp, group_t.p, group_s.p, out16);
}

bool Registrar::register_min_support(const CloudDev &tgt, const CloudDev &src, int ms_t, int ms_s, float out16[16]) {
  double t0 = now_s();
  prep[0].ready = prep[1].ready = false;
  lane_spacing = 0;
  std::vector<PlaneParam> tp = detect_planes_dev(tgt, ms_t, group_t);
  std::vector<PlaneParam> sp = detect_planes_dev(src, ms_s, group_s);
  times.planes = now_s() - t0;
  return register_core(tgt, src, tp, sp, group_t.p, group_s.p, out16);
}

bool Registrar::register_with_planes(const CloudDev &tgt, const CloudDev &src, const std::vector<PlaneRec> &tplanes,
                                     const std::vector<PlaneRec> &splanes, float out16[16]) {
  // caller-supplied planes: host index lists -> plane index per point, uploaded once
  const CloudDev *clouds[2] = {&tgt, &src};
  const std::vector<PlaneRec> *pin[2] = {&tplanes, &splanes};
  DevBuf<int> *gbuf[2] = {&group_t, &group_s};
  std::vector<PlaneParam> pp[2];
  for (int side = 0; side < 2; ++side) {
    const size_t n = clouds[side]->n;
    std::vector<int> grp(n, -1);
    for (size_t i = 0; i < pin[side]->size(); ++i) {
      const PlaneRec &p = (*pin[side])[i];
      for (int id : p.idx) if (id >= 0 && (size_t) id < n) grp[id] = (int) i;
      pp[side].push_back({{p.n[0], p.n[1], p.n[2]}, p.d, (long long) p.idx.size()});
    }
    int *d = gbuf[side]->ensure(std::max<size_t>(n, 1));
    if (n) PLADE_CUDA(cudaMemcpyAsync(d, grp.data(), sizeof(int) * n, cudaMemcpyHostToDevice, dev.stream));
    stream_sync(dev.stream);
  }
  times.planes = 0;
  prep[0].ready = prep[1].ready = false;
  lane_spacing = 0;
  return register_core(tgt, src, pp[0], pp[1], group_t.p, group_s.p, out16);
}

bool Registrar::register_core(const CloudDev &tgt, const CloudDev &src, const std::vector<PlaneParam> &tplanes,
                              const std::vector<PlaneParam> &splanes, const int *d_group_t, const int *d_group_s, float out16[16]) {
  const double t_begin = now_s();
  mark("core_begin");
  for (int i = 0; i < 16; ++i) out16[i] = (i % 5 == 0) ? 1.f : 0.f;
  cudaStream_t s = dev.stream;
  if (tgt.n == 0 || src.n == 0) { last_error = "empty cloud"; return false; }

  // ---- constants, PLADE/plade.cpp:41-56 -----------------------------------------------------------
  double t0 = now_s();
  const float average_space = lane_spacing > 0 ? lane_spacing : average_spacing(src);
  times.spacing = now_s() - t0;
  mark("spacing");
  float downSampleDistance = average_space * 4;
  float lengthThreshold = average_space * 5;
  float angleThreshold = (float) (5.0 / 180 * M_PI);
  float cosAngleThreshold = std::cos(angleThreshold);
  const int maxCandidateResultNum = params.max_candidates;
  float scale = (float) (lengthThreshold / std::cos(M_PI_2 - angleThreshold));
  if (debug) {
    put("average_space", std::vector<float>(1, average_space));
    put("downsample_distance", std::vector<float>(1, downSampleDistance));
  }
  if (!(downSampleDistance > 0)) { last_error = "degenerate spacing"; return false; }

  // ---- per-side set-up, PLADE/plade.cpp:74-122 / :287-335 -----------------------------------------
  t0 = now_s();
  Side S[2];   // 0 = target ("main"), 1 = source ("current")
  const CloudDev *clouds[2] = {&tgt, &src};
  const std::vector<PlaneParam> *planes_in[2] = {&tplanes, &splanes};
  const int *d_groups[2] = {d_group_t, d_group_s};
  DevBuf<float4> *ds_dev[2] = {&ds_tgt, &ds_src};
  DevBuf<float4> *ds_pl[2] = {&ds_planes_t, &ds_planes_s};
  for (int side = 0; side < 2; ++side) {
    Side &A = S[side];
    const CloudDev &C = *clouds[side];
    const std::vector<PlaneParam> &P = *planes_in[side];
    // done already inside the cloud's extraction lane (register_clouds), unless the planes came from the caller
    if (!(prep[side].ready && prep[side].leaf == downSampleDistance)) {
      if (side == 1) stream_sync(dev.stream);     // lane 1 reads nothing lane 0 still writes, but keep the order simple
      prepare_side(side, C, P, d_groups[side], downSampleDistance, prep[side]);
      if (side == 1) stream_sync(dev2.stream);
    }
    A.n_ds = prep[side].n_ds;
    A.plane_ds_start = prep[side].plane_ds_start;
    const size_t nv = prep[side].n_plane_ds;
    if (debug) {     // host copies are only needed for the stage dumps
      A.ds.resize(A.n_ds);
      A.plane_ds.resize(nv);
      if (A.n_ds) PLADE_CUDA(cudaMemcpyAsync(A.ds.data(), ds_dev[side]->p, sizeof(float4) * A.n_ds, cudaMemcpyDeviceToHost, s));
      if (nv) PLADE_CUDA(cudaMemcpyAsync(A.plane_ds.data(), ds_pl[side]->p, sizeof(float4) * nv, cudaMemcpyDeviceToHost, s));
      stream_sync(s);
    }
  }
  {
    std::vector<ObbResult> obb;
    for (int side = 0; side < 2; ++side) obb.insert(obb.end(), prep[side].obb.begin(), prep[side].obb.end());
    size_t k = 0;
    for (int side = 0; side < 2; ++side) {
      Side &A = S[side];
      const std::vector<PlaneParam> &P = *planes_in[side];
      const ObbResult &full = obb[k++];
      if (full.rc != 0) { last_error = "empty down-sampled cloud"; return false; }
      A.center = full.center;
      A.radius = std::max(std::max(full.width, full.height), full.depth) / 2;
      A.planes.resize(P.size());
      A.corners4.resize(P.size());
      A.plane_center.resize(P.size());
      A.plane_radius.resize(P.size());
      for (size_t i = 0; i < P.size(); ++i) {
        A.planes[i] = {P[i].n[0], P[i].n[1], P[i].n[2], P[i].d};
        const ObbResult &o = obb[k++];
        for (int c = 0; c < 4; ++c) A.corners4[i][c] = V3();
        A.plane_center[i] = V3();
        A.plane_radius[i] = 0;
        if (o.rc != 0) continue;
        // ProjectPoints2Plane on corners [0,4) (PLADE/util.h:293-329)
        float Aa = P[i].n[0], B = P[i].n[1], Cc = P[i].n[2], D = P[i].d;
        for (int c = 0; c < 4; ++c) {
          const V3 &q = o.corners[c];
          float kk = -(Aa * q.x + B * q.y + Cc * q.z + D) / (Aa * Aa + B * B + Cc * Cc);
          A.corners4[i][c] = V3(q.x + kk * Aa, q.y + kk * B, q.z + kk * Cc);
        }
        A.plane_center[i] = (A.corners4[i][0] + A.corners4[i][2]) / 2.f;
        A.plane_radius[i] = norm(A.corners4[i][0] - A.corners4[i][2]) / 2;
      }
    }
    mark("obb");
  }
  times.downsample = now_s() - t0;

  // ---- intersection lines, PLADE/plade.cpp:124-172 / :337-381 ---------------------------------------
  t0 = now_s();
  for (int side = 0; side < 2; ++side) {
    Side &A = S[side];
    size_t np = A.planes.size();
    for (size_t i = 0; i < np; ++i)
      for (size_t j = i + 1; j < np; ++j) {
        Line L;
        if (0 != plane_intersection_line(A.planes[i].data(), A.planes[j].data(), L.vec, L.pt)) continue;
        V3 tv = L.pt - A.center;
        double distance = std::sqrt((double) sqnorm(tv) - std::pow((double) dot(tv, L.vec), 2));
        if (distance > A.radius) continue;
        // ComputeMeanDistanceOfLine2Plane (util.h:390-426) only re-normalises lineVec (its result is unused);
        // it fails iff the plane has no bounding-box corners, i.e. an empty down-sampled plane
        if (A.plane_ds_start[i + 1] == A.plane_ds_start[i]) continue;
        normalize(L.vec);
        if (A.plane_ds_start[j + 1] == A.plane_ds_start[j]) continue;
        normalize(L.vec);
        L.p0 = (int) i;
        L.p1 = (int) j;
        A.lines.push_back(L);
      }
  }
  Side &M = S[0], &Cu = S[1];
  const size_t mainLinesNum = M.lines.size(), currentLinesNum = Cu.lines.size();
  times.lines = now_s() - t0;
  if (debug) {
    for (int side = 0; side < 2; ++side) {
      std::string p = side == 0 ? "tgt_" : "src_";
      std::vector<float> lv, ds, pc, pr, c4;
      std::vector<int> lp;
      for (const Line &L : S[side].lines) {
        lv.insert(lv.end(), {L.vec.x, L.vec.y, L.vec.z, L.pt.x, L.pt.y, L.pt.z});
        lp.push_back(L.p0); lp.push_back(L.p1);
      }
      for (const float4 &q : S[side].ds) ds.insert(ds.end(), {q.x, q.y, q.z});
      put(p + "lines", lv); put(p + "line_planes", lp); put(p + "ds", ds);
      put(p + "center", std::vector<float>{S[side].center.x, S[side].center.y, S[side].center.z});
      put(p + "radius", std::vector<double>(1, S[side].radius));
      std::vector<float> pds;
      for (const float4 &q : S[side].plane_ds) pds.insert(pds.end(), {q.x, q.y, q.z});
      put(p + "plane_ds", pds); put(p + "plane_ds_offsets", S[side].plane_ds_start);
      for (size_t i = 0; i < S[side].planes.size(); ++i) {
        for (int k = 0; k < 4; ++k) c4.insert(c4.end(), {S[side].corners4[i][k].x, S[side].corners4[i][k].y, S[side].corners4[i][k].z});
        pc.insert(pc.end(), {S[side].plane_center[i].x, S[side].plane_center[i].y, S[side].plane_center[i].z});
        pr.push_back(S[side].plane_radius[i]);
      }
      put(p + "plane_corners4", c4); put(p + "plane_center", pc); put(p + "plane_radius", pr);
    }
  }
  if (mainLinesNum == 0 || currentLinesNum == 0) {
    std::cerr << "registration failed: no matched result found" << std::endl;
    last_error = "no intersection lines";
    return false;
  }

  // ---- target descriptor table "22", ConstructPairLinesKdTree (PLADE/util.cpp:706-1165) -------------
  t0 = now_s();
  const float angleThresh10 = (float) std::cos(10.0 / 180 * M_PI);
  // pass 1: play both sides' in-place normalisations in the reference's order and collect every line pair
  // whose closest points the reference solves for (all i < j), then solve them in one launch
  PairSolve solve;
  std::vector<int> t_slot(mainLinesNum * mainLinesNum, -1), s_slot(currentLinesNum * currentLinesNum, -1);
  {
    std::vector<Line> keep = M.lines;
    for (size_t i = 0; i < mainLinesNum; ++i)
      for (size_t j = i + 1; j < mainLinesNum; ++j)
        t_slot[i * mainLinesNum + j] = solve.add(M.lines[i].vec, M.lines[i].pt, M.lines[j].vec, M.lines[j].pt);
    M.lines = keep;                  // pass 2 replays the normalisations interleaved with the descriptor tests
    for (size_t i = 0; i < currentLinesNum; ++i)
      for (size_t j = i + 1; j < currentLinesNum; ++j)
        s_slot[i * currentLinesNum + j] = solve.add(Cu.lines[i].vec, Cu.lines[i].pt, Cu.lines[j].vec, Cu.lines[j].pt);
    solve.out.resize((size_t) 6 * solve.count);
    nearest_points_batch(dev, svd_sc, solve.in.data(), solve.count, solve.out.data());
  }
  std::vector<float> db_desc;                 // 8 per entry
  struct DbRec { V3 v1, v2, p1; };
  std::vector<DbRec> db;
  std::vector<int> db_pair;
  {
    std::vector<V3> sp(M.planes.size());
    for (size_t i = 0; i < sp.size(); ++i) sp[i] = V3(M.planes[i][0], M.planes[i][1], M.planes[i][2]);
    std::vector<NearPts> tab(mainLinesNum * mainLinesNum);
    for (size_t i = 0; i < mainLinesNum; ++i) {
      Line &l1 = M.lines[i];
      for (size_t j = 0; j < mainLinesNum; ++j) {
        if (i == j) continue;
        Line &l2 = M.lines[j];
        NearPts &e = tab[i * mainLinesNum + j];
        if (i > j) {
          const NearPts &o = tab[j * mainLinesNum + i];
          e.a = o.b; e.b = o.a; e.length = o.length;
        } else {
          normalize(l1.vec);          // replays the in-place normalisation of pass 1
          normalize(l2.vec);
          solve.get(t_slot[i * mainLinesNum + j], e.a, e.b, e.length);
          e.length = e.length / scale;
        }
        if (std::fabs(dot(l1.vec, l2.vec)) > angleThresh10) continue;
        float d[8];
        V3 nl1, nl2;
        pair_descriptor(l1.vec, l2.vec, sp[l1.p0], sp[l1.p1], sp[l2.p0], sp[l2.p1], d, nl1, nl2);
        d[0] = (float) e.length;
        db_desc.insert(db_desc.end(), d, d + 8);
        db.push_back({nl1, nl2, e.a});
        db_pair.push_back((int) i); db_pair.push_back((int) j);
      }
    }
  }
  // ---- source pair table + query descriptors, PLADE/plade.cpp:453-521, PLADE/util.cpp:133-168 ---------
  std::vector<float> q_desc;
  struct QRec { V3 v1, v2, p1; int i, j; };
  std::vector<QRec> qrec;
  {
    std::vector<V3> sp(Cu.planes.size());
    for (size_t i = 0; i < sp.size(); ++i) sp[i] = V3(Cu.planes[i][0], Cu.planes[i][1], Cu.planes[i][2]);
    std::vector<NearPts> tab(currentLinesNum * currentLinesNum);
    for (size_t i = 0; i < currentLinesNum; ++i)
      for (size_t j = i + 1; j < currentLinesNum; ++j) {
        NearPts &e = tab[i * currentLinesNum + j];
        solve.get(s_slot[i * currentLinesNum + j], e.a, e.b, e.length);
        e.length = e.length / scale;
      }
    for (size_t i = 0; i < currentLinesNum; ++i)
      for (size_t j = i + 1; j < currentLinesNum; ++j) {
        const Line &l1 = Cu.lines[i], &l2 = Cu.lines[j];
        if (std::fabs(dot(l1.vec, l2.vec)) > angleThresh10) continue;
        const NearPts &e = tab[i * currentLinesNum + j];
        float d[8];
        QRec r;
        pair_descriptor(l1.vec, l2.vec, sp[l1.p0], sp[l1.p1], sp[l2.p0], sp[l2.p1], d, r.v1, r.v2);
        d[0] = (float) e.length;
        r.p1 = e.a; r.i = (int) i; r.j = (int) j;
        q_desc.insert(q_desc.end(), d, d + 8);
        qrec.push_back(r);
      }
  }
  times.descriptors = now_s() - t0;
  mark("lines+descriptors");
  if (debug) {
    put("tgt_db_desc", db_desc); put("tgt_db_pair", db_pair); put("src_q_desc", q_desc);
    std::vector<int> qp;
    for (const QRec &r : qrec) { qp.push_back(r.i); qp.push_back(r.j); }
    put("lines_to_match", qp);
  }

  // ---- K3c: radius matching, PLADE/util.cpp:163 -----------------------------------------------------
  t0 = now_s();
  std::vector<int> m_off, m_idx;
  std::vector<double> m_d2;
  size_t n_match = match_descriptors(dev, match_sc, db_desc.data(), (int) (db_desc.size() / 8), q_desc.data(),
                                     (int) (q_desc.size() / 8), (float) params.descriptor_radius, m_off, m_idx, m_d2);
  times.match = now_s() - t0;
  mark("match");
  if (debug) { put("match_offsets", m_off); put("match_idx", m_idx); put("match_dist2", m_d2); }

  // ---- K4a: one rigid transform per match, PLADE/util.cpp:303-327 -------------------------------------
  t0 = now_s();
  std::vector<MatchPairIn> mp(n_match);
  {
    size_t k = 0;
    for (size_t qi = 0; qi < qrec.size(); ++qi)
      for (int e = m_off[qi]; e < m_off[qi + 1]; ++e, ++k) {
        const QRec &q = qrec[qi];
        const DbRec &d = db[m_idx[e]];
        MatchPairIn &o = mp[k];
        for (int c = 0; c < 3; ++c) { o.sv1[c] = q.v1[c]; o.sv2[c] = q.v2[c]; o.dv1[c] = d.v1[c]; o.dv2[c] = d.v2[c]; o.sp[c] = q.p1[c]; o.tp[c] = d.p1[c]; }
      }
  }
  std::vector<RigidOut> rt;
  transforms_from_matches(dev, hyp_sc, mp.data(), n_match, rt);
  if (n_match == 0) {
    std::cerr << "registration failed: no matched result found" << std::endl;
    last_error = "no descriptor matches";
    return false;
  }
  // ---- K4b: ClusterTransformation(len/2, ang/2), PLADE/util.cpp:331 ------------------------------------
  std::vector<int> label;
  cluster_transforms(dev, hyp_sc, rt, (float) ((double) lengthThreshold / 2), (float) ((double) angleThreshold / 2), label);
  // clusters in CEC emission order = ascending smallest member; size per cluster
  std::vector<int> rep;                         // representative (= label value) per cluster, ascending
  std::vector<int> csize;
  {
    std::vector<int> cid(n_match, -1);
    for (size_t i = 0; i < n_match; ++i)
      if (label[i] == (int) i) { cid[i] = (int) rep.size(); rep.push_back((int) i); csize.push_back(0); }
    for (size_t i = 0; i < n_match; ++i) csize[cid[label[i]]]++;
  }
  if (debug) {
    std::vector<float> R, T;
    for (const RigidOut &r : rt) { R.insert(R.end(), r.R, r.R + 9); T.insert(T.end(), r.T, r.T + 3); }
    put("init_R", R); put("init_T", T); put("cluster_label", label);
    std::vector<float> in18;
    for (const MatchPairIn &m : mp) in18.insert(in18.end(), m.sv1, m.sv1 + 18);
    put("match_in18", in18);
    put("cluster_params", std::vector<float>{(float) ((double) lengthThreshold / 2), (float) ((double) angleThreshold / 2)});
  }
  // sort by cluster size, same struct / comparator / algorithm as PLADE/util.cpp:335-347
  std::vector<LenIdx> sortVec(rep.size());
  for (size_t i = 0; i < sortVec.size(); ++i) { sortVec[i].index = (int) i; sortVec[i].length = (float) csize[i]; }
  std::sort(sortVec.begin(), sortVec.end(), cmp_greater);

  // ---- K4c: centre gate + plane consistency, PLADE/util.cpp:352-401 -----------------------------------
  const size_t currentPlanesNum = Cu.planes.size(), mainPlanesNum = M.planes.size();
  const float maxRadius = (float) M.radius;
  std::vector<std::vector<std::pair<int, int>>> matches;
  std::vector<int> cand_rt;     // index into rt
  for (size_t ii = 0; ii < sortVec.size(); ++ii) {
    int k = rep[sortVec[ii].index];
    M3 R; memcpy(R.m, rt[k].R, sizeof(R.m));
    V3 T(rt[k].T[0], rt[k].T[1], rt[k].T[2]);
    V3 tc = mul(R, Cu.center) + T;
    if (norm(tc - M.center) > maxRadius) continue;
    std::vector<std::pair<int, int>> cur;
    for (size_t i1 = 0; i1 < currentPlanesNum; ++i1) {
      V3 plane1 = mul(R, V3(Cu.planes[i1][0], Cu.planes[i1][1], Cu.planes[i1][2]));
      float d = -(-Cu.planes[i1][3] + dot(plane1, T));
      V3 srcCenter2Dest = mul(R, Cu.plane_center[i1]) + T;
      for (size_t j1 = 0; j1 < mainPlanesNum; ++j1) {
        V3 plane_A(M.planes[j1][0], M.planes[j1][1], M.planes[j1][2]);
        if (dot(plane1, plane_A) < cosAngleThreshold) continue;
        double center2PlaneDistance = (std::fabs(dot(plane_A, srcCenter2Dest) + M.planes[j1][3]) + std::fabs(dot(plane1, M.plane_center[j1]) + d)) / 2;
        if (center2PlaneDistance > lengthThreshold) continue;
        double distance = norm(srcCenter2Dest - M.plane_center[j1]);
        if (distance / (Cu.plane_radius[i1] + M.plane_radius[j1]) > 1) continue;
        cur.push_back(std::make_pair((int) i1, (int) j1));
        break;
      }
    }
    matches.push_back(cur);
    cand_rt.push_back(k);
  }
  // ---- candidate budget, PLADE/util.cpp:403-445 ----------------------------------------------------------
  size_t maxMatchNum = 0;
  for (auto &m : matches) maxMatchNum = std::max(maxMatchNum, m.size());
  std::vector<std::vector<int>> matchedPlanes;
  if (maxMatchNum > 0) {
    int matchedCount = 0;
    for (size_t i = maxMatchNum; i >= 2; i--) {
      std::vector<int> tmp;
      for (size_t j = 0; j < matches.size(); ++j)
        if (i == matches[j].size()) { tmp.push_back((int) j); matchedCount++; }
      matchedPlanes.push_back(tmp);
      if (matchedCount >= maxCandidateResultNum) break;
    }
  }
  times.hypotheses = now_s() - t0;
  mark("hypotheses");

  // ---- K4d: penetration filter, PLADE/util.cpp:447-519 ------------------------------------------------------
  t0 = now_s();
  std::vector<MatchedHyp> results;
  {
    // the candidates the reference would look at, in its order (count++ > maxCandidateResultNum stops the walk)
    std::vector<int> cand;
    int count = 0;
    bool stop = false;
    for (size_t m = 0; m < matchedPlanes.size() && !stop; ++m)
      for (size_t i = 0; i < matchedPlanes[m].size(); ++i) {
        if (count++ > maxCandidateResultNum) { stop = true; break; }
        cand.push_back(matchedPlanes[m][i]);
      }
    std::vector<float> hyp12(cand.size() * 12);
    for (size_t c = 0; c < cand.size(); ++c) {
      const RigidOut &r = rt[cand_rt[cand[c]]];
      memcpy(&hyp12[12 * c], r.R, sizeof(float) * 9);
      memcpy(&hyp12[12 * c + 9], r.T, sizeof(float) * 3);
    }
    PenSide ps, pt;
    ps.planes = Cu.planes; ps.corners4 = Cu.corners4; ps.center = Cu.plane_center; ps.ds_start = Cu.plane_ds_start; ps.d_pts = ds_planes_s.p;
    pt.planes = M.planes; pt.corners4 = M.corners4; pt.center = M.plane_center; pt.ds_start = M.plane_ds_start; pt.d_pts = ds_planes_t.p;
    std::vector<unsigned char> pen;
    penetration_filter(dev, pen_sc, ps, pt, hyp12.data(), (int) cand.size(), lengthThreshold, angleThreshold, pen);
    if (debug) {
      std::vector<int> c_rt, c_np, flag(pen.begin(), pen.end());
      for (int index : cand) { c_rt.push_back(cand_rt[index]); c_np.push_back((int) matches[index].size()); }
      put("cand_rt", c_rt); put("cand_nplanes", c_np); put("cand_pen", flag);
    }
    for (size_t c = 0; c < cand.size(); ++c) {
      if (pen[c]) continue;
      const int index = cand[c], k = cand_rt[index];
      MatchedHyp r;
      memcpy(r.R.m, rt[k].R, sizeof(r.R.m));
      r.T = V3(rt[k].T[0], rt[k].T[1], rt[k].T[2]);
      r.planes = matches[index];
      results.push_back(r);
    }
  }
  times.penetration = now_s() - t0;
  mark("penetration");
  if (debug) {
    std::vector<float> R, T;
    std::vector<int> np;
    for (const MatchedHyp &r : results) {
      R.insert(R.end(), r.R.m, r.R.m + 9);
      T.insert(T.end(), {r.T.x, r.T.y, r.T.z});
      np.push_back((int) r.planes.size());
    }
    put("mr_R", R); put("mr_T", T); put("mr_nplanes", np);
  }
  if (results.empty()) {
    std::cerr << "registration failed: no matched result found" << std::endl;
    last_error = "no matched result";
    return false;
  }

  // ---- K5: verification, PLADE/plade.cpp:545-575 ------------------------------------------------------------
  t0 = now_s();
  const int H = (int) results.size();
  std::vector<HypParams> hp(H);
  for (int i = 0; i < H; ++i) {
    memcpy(hp[i].R, results[i].R.m, sizeof(float) * 9);
    hp[i].T[0] = results[i].T.x; hp[i].T[1] = results[i].T.y; hp[i].T[2] = results[i].T.z;
    V3 cc = mul(results[i].R, Cu.center) + results[i].T;
    hp[i].c[0] = cc.x; hp[i].c[1] = cc.y; hp[i].c[2] = cc.z;
    hp[i].pad = 0;
  }
  const bool nccl_path = shard_world > 1 && nccl != nullptr;
  for (int i = 0; i < H; ++i) { const int np = (int) results[i].planes.size(); memcpy(&hp[i].pad, &np, sizeof(int)); }   // rides along for shard_key_kernel
  bool list_is_local = true;
  if (nccl_path) {
    // Every rank ran its own plane extraction and matching; their hypothesis lists agree unless a float sum went the other
    // way somewhere.  Rank 0's list is THE list: its length, then its records, are broadcast (two small ncclBroadcasts).
    unsigned long long *d_key = d_shard_key.ensure(4), *h_key = h_shard_key.ensure(4);
    h_key[0] = (unsigned long long) H;
    PLADE_CUDA(cudaMemcpyAsync(d_key, h_key, sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
    nccl_broadcast_bytes(nccl, d_key, sizeof(unsigned long long), 0, s);
    PLADE_CUDA(cudaMemcpyAsync(h_key, d_key, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    const int H0 = (int) h_key[0];
    if (H0 <= 0) { last_error = "sharded verification: rank 0 has no hypothesis"; return false; }
    HypParams *d_all = d_hyp_all.ensure((size_t) H0);
    if (shard_rank == 0) PLADE_CUDA(cudaMemcpyAsync(d_all, hp.data(), sizeof(HypParams) * H0, cudaMemcpyHostToDevice, s));
    nccl_broadcast_bytes(nccl, d_all, sizeof(HypParams) * (size_t) H0, 0, s);
    std::vector<HypParams> hp0((size_t) H0);
    PLADE_CUDA(cudaMemcpyAsync(hp0.data(), d_all, sizeof(HypParams) * H0, cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    list_is_local = H0 == H && memcmp(hp0.data(), hp.data(), sizeof(HypParams) * (size_t) H0) == 0;
    hp.swap(hp0);
  }
  const int H_all = (int) hp.size();
  // hypothesis shard of this rank (all ranks hold replicas of the clouds)
  std::vector<int> mine;
  for (int i = 0; i < H_all; ++i) if (i % shard_world == shard_rank) mine.push_back(i);
  std::vector<HypParams> hp_mine(mine.size());
  for (size_t i = 0; i < mine.size(); ++i) hp_mine[i] = hp[mine[i]];
  build_target_grid(dev, ds_tgt.p, M.n_ds, downSampleDistance, grid);
  HypParams *d_h = d_hyp.ensure(std::max<size_t>(1, hp_mine.size()));
  unsigned int *d_c = d_counts.ensure(std::max<size_t>(1, hp_mine.size()));
  if (!hp_mine.empty()) PLADE_CUDA(cudaMemcpyAsync(d_h, hp_mine.data(), sizeof(HypParams) * hp_mine.size(), cudaMemcpyHostToDevice, s));
  PLADE_CUDA(cudaEventRecord(ev0, s));
  verify_hypotheses(dev, ds_src.p, Cu.n_ds, grid, d_h, (int) hp_mine.size(), (float) Cu.radius, downSampleDistance, d_c);
  PLADE_CUDA(cudaEventRecord(ev1, s));
  std::vector<unsigned int> counts_mine(hp_mine.size());
  // Every rank runs its own RANSAC and matching: the shards are only comparable if all ranks hold the SAME hypothesis
  // list.  A hash of the list travels with the key (MAX of h and of ~h: both equal the local values iff all ranks agree).
  unsigned long long list_hash = 1469598103934665603ull;
  if (shard_world > 1) {
    auto mix = [&](const void *p, size_t nbytes) { const unsigned char *b = static_cast<const unsigned char *>(p); for (size_t k = 0; k < nbytes; ++k) { list_hash ^= b[k]; list_hash *= 1099511628211ull; } };
    mix(&H_all, sizeof(H_all));
    for (int i = 0; i < H_all; ++i) mix(hp[i].R, sizeof(float) * 16);
  }
  unsigned long long reduced[3] = {0, 0, 0};
  if (nccl_path) {
    // device-side argmax of the shard + ONE ncclAllReduce(ncclUint64, ncclMax) on this stream; no count leaves the device
    unsigned long long *d_key = d_shard_key.ensure(4), *h_key = h_shard_key.ensure(4);
    shard_best_key(dev, d_c, d_h, (int) hp_mine.size(), shard_rank, shard_world, (double) std::min(Cu.n_ds, M.n_ds), (double) currentPlanesNum, 0, d_key);
    h_key[1] = list_hash; h_key[2] = ~list_hash;
    PLADE_CUDA(cudaMemcpyAsync(d_key + 1, h_key + 1, sizeof(unsigned long long) * 2, cudaMemcpyHostToDevice, s));
    nccl_allreduce_max_u64(nccl, d_key, 3, s);
    PLADE_CUDA(cudaMemcpyAsync(h_key, d_key, sizeof(unsigned long long) * 3, cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    for (int k = 0; k < 3; ++k) reduced[k] = h_key[k];
  } else {
    if (!hp_mine.empty()) PLADE_CUDA(cudaMemcpyAsync(counts_mine.data(), d_c, sizeof(unsigned int) * hp_mine.size(), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
  }
  {
    float ms = 0;
    PLADE_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    times.verify_kernel_ms = ms;
    times.verify_h = (double) hp_mine.size(); times.verify_ns = (double) Cu.n_ds; times.verify_nt = (double) M.n_ds;
  }
  const size_t denom = std::min(Cu.n_ds, M.n_ds);
  auto score_of = [&](int i, unsigned int cnt) {
    float overlap = (float) (double(cnt) / denom);                                   // util.h:644
    return (float) (0.2 * (results[i].planes.size() / double(currentPlanesNum)) + 0.8 * overlap);   // plade.cpp:561-562
  };
  int best = -1;
  if (shard_world == 1) {
    std::vector<LenIdx> ov(H);
    for (int i = 0; i < H; ++i) { ov[i].index = i; ov[i].length = score_of(i, counts_mine[i]); }
    if (debug) {
      std::vector<float> sc; std::vector<int> cn;
      for (int i = 0; i < H; ++i) { sc.push_back(ov[i].length); cn.push_back((int) counts_mine[i]); }
      put("ver_score", sc); put("ver_count", cn);
      std::vector<float> cc;
      for (int i = 0; i < H; ++i) cc.insert(cc.end(), hp[i].c, hp[i].c + 3);
      put("ver_center", cc);
    }
    std::sort(ov.begin(), ov.end(), cmp_greater);        // plade.cpp:565 (same struct, comparator, algorithm)
    best = ov[0].index;
  } else {
    // packed key {score bits, ~index}: max picks the highest score, ties -> lowest index
    if (!nccl_path) {
      if (!allreduce) { last_error = "sharded verification without a reducer (plade_shard_init_nccl or plade_set_shard with a callback)"; return false; }
      unsigned long long key = 0;
      for (size_t i = 0; i < mine.size(); ++i) {
        float sc = score_of(mine[i], counts_mine[i]);
        unsigned int bits;
        memcpy(&bits, &sc, 4);
        unsigned long long k = ((unsigned long long) bits << 32) | (0xFFFFFFFFu - (unsigned int) mine[i]);
        key = std::max(key, k);
      }
      reduced[0] = key; reduced[1] = list_hash; reduced[2] = ~list_hash;
      for (int k = 0; k < 3; ++k) allreduce(&reduced[k], allreduce_user);
    }
    if (reduced[1] != list_hash || reduced[2] != ~list_hash) {
      std::cerr << "sharded verification: the ranks built different hypothesis lists" << std::endl;
      last_error = "sharded verification: the ranks built different hypothesis lists";
      return false;
    }
    if (reduced[0] == 0) { last_error = "sharded verification failed: no hypothesis was verified"; return false; }
    best = (int) (0xFFFFFFFFu - (unsigned int) (reduced[0] & 0xFFFFFFFFu));
    if (best < 0 || best >= H_all) { last_error = "sharded verification failed"; return false; }
  }
  times.verify = now_s() - t0;

  MatchedHyp W;
  if (list_is_local) W = results[best];
  else { memcpy(W.R.m, hp[best].R, sizeof(float) * 9); W.T = V3(hp[best].T[0], hp[best].T[1], hp[best].T[2]); }     // rank 0's record
  {
    // headless stand-in for the reference's ResultViewer: what the winner rests on (plade_last_report, CLI --report)
    std::ostringstream js;
    js.precision(9);
    js << "{\"target_points\": " << tgt.n << ", \"source_points\": " << src.n << ", \"target_planes\": " << M.planes.size() << ", \"source_planes\": "
       << Cu.planes.size() << ", \"average_spacing\": " << average_space << ", \"target_ds_points\": " << M.n_ds << ", \"source_ds_points\": " << Cu.n_ds
       << ", \"hypotheses_verified\": " << H << ", \"shard_world\": " << shard_world << ", \"winner\": " << best << ", \"winner_matched_planes\": [";
    for (size_t k = 0; k < W.planes.size(); ++k) js << (k ? ", " : "") << "[" << W.planes[k].first << ", " << W.planes[k].second << "]";
    js << "]";
    if (shard_world == 1) {
      const double ov = double(counts_mine[best]) / double(denom);
      js << ", \"winner_inliers\": " << counts_mine[best] << ", \"overlap_ratio\": " << ov << ", \"score\": " << score_of(best, counts_mine[best]);
    }
    js << ", \"target_plane_supports\": [";
    for (size_t k = 0; k < tplanes.size(); ++k) js << (k ? ", " : "") << tplanes[k].size;
    js << "], \"source_plane_supports\": [";
    for (size_t k = 0; k < splanes.size(); ++k) js << (k ? ", " : "") << splanes[k].size;
    js << "]}";
    report = js.str();
  }
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) out16[4 * r + c] = W.R(r, c);
    out16[4 * r + 3] = W.T[r];
  }
  out16[12] = out16[13] = out16[14] = 0.f;
  out16[15] = 1.f;
  times.total = now_s() - t_begin;
  mark("verify");
  print_marks();
  return true;
}

}  // namespace plade
