// ORACLE / TEST INFRASTRUCTURE ONLY — never linked into the product library.
//
// C-ABI driver around the reference's own, unmodified sources (compiled where they lie under
// /root/reference by oracle/Makefile).  It adds no algorithm: every entry point calls the
// reference function it names (file:line relative to /root/reference/code/) and copies plain
// arrays in and out so that tests/ and bench.py's cpu_baseline / --impl reference arm can drive
// it through ctypes.  Intermediate results are published as named blobs (ref_blob).
//
//   * time() is interposed (link flag -Bsymbolic-functions) so that the time-seeded RANSAC
//     (3rd_party/ransac/RansacShapeDetector.cpp:463-464) is reproducible: ref_set_seed(s>=0).
//   * P/plade.cpp is compiled a second time with MatchingLines/registration/extract renamed on
//     the command line; the renamed MatchingLines call lands in the hook below, which records
//     what registration() (PLADE/plade.cpp:31-536) built and forwards to the real function.
//   * cv::fitLine (imgproc) is referenced by PLADE/util.cpp:684 (Fit3DLine) but that path is
//     dead — boundary lines are never produced (PLADE/plade.cpp:175-178,383-384) — so it is
//     stubbed here instead of compiling OpenCV imgproc.
#include "util.h"
#include "plade.h"
#include "plane_extraction.h"
#include <opencv2/imgproc/imgproc.hpp>
#include <MiscLib/Random.h>
#include <RansacShapeDetector.h>
#include <PlanePrimitiveShape.h>
#include <ScorePrimitiveShapeVisitor.h>
#include <FlatNormalThreshPointCompatibilityFunc.h>
#include <Candidate.h>
#include <algorithm>

#include <map>
#include <string>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <chrono>

// ---- reference symbols without a header declaration --------------------------------------------
std::vector<PLANE> extract(pcl::PointCloud<pcl::PointNormal>::Ptr cloud, int init_min_support);  // PLADE/plade.cpp:602
bool plade_oracle_registration_hooked(Eigen::Matrix<float, 4, 4> &transformation,
                                      pcl::PointCloud<pcl::PointNormal>::Ptr target_cloud,
                                      pcl::PointCloud<pcl::PointNormal>::Ptr source_cloud,
                                      const std::vector<PLANE> &target_planes,
                                      const std::vector<PLANE> &source_planes);

namespace cv {
void fitLine(InputArray, OutputArray, int, double, double, double) {
  fprintf(stderr, "oracle: cv::fitLine stub reached (dead path in the reference)\n");
  abort();
}
}  // namespace cv

// ---- deterministic time() ----------------------------------------------------------------------
static long g_seed = -1;
extern "C" time_t time(time_t *t) {
  time_t v;
  if (g_seed >= 0) v = (time_t) g_seed;
  else v = (time_t) std::chrono::duration_cast<std::chrono::seconds>(
               std::chrono::system_clock::now().time_since_epoch()).count();
  if (t) *t = v;
  return v;
}

// ---- blob store --------------------------------------------------------------------------------
static std::map<std::string, std::vector<char> > g_blobs;
template <typename T> static void put(const std::string &name, const std::vector<T> &v) {
  std::vector<char> &b = g_blobs[name];
  b.resize(v.size() * sizeof(T));
  if (!v.empty()) memcpy(b.data(), v.data(), b.size());
}
template <typename T> static void put1(const std::string &name, T v) { put(name, std::vector<T>(1, v)); }

typedef pcl::PointCloud<pcl::PointNormal> CloudPN;
typedef pcl::PointCloud<pcl::PointXYZ> CloudXYZ;

static CloudPN::Ptr make_cloud(const float *xyzn, size_t n) {
  CloudPN::Ptr c(new CloudPN);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float *p = xyzn + 6 * i;
    c->at(i) = pcl::PointNormal(p[0], p[1], p[2], p[3], p[4], p[5]);
  }
  return c;
}
static CloudXYZ::Ptr make_xyz(const float *xyz, size_t n) {
  CloudXYZ::Ptr c(new CloudXYZ);
  c->resize(n);
  for (size_t i = 0; i < n; ++i) c->at(i) = pcl::PointXYZ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
  return c;
}
static void put_xyz(const std::string &name, const CloudXYZ &c) {
  std::vector<float> v(c.size() * 3);
  for (size_t i = 0; i < c.size(); ++i) { v[3 * i] = c[i].x; v[3 * i + 1] = c[i].y; v[3 * i + 2] = c[i].z; }
  put(name, v);
}
static void to_rowmajor(const Eigen::Matrix4f &T, float *out16) {
  for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) out16[4 * r + c] = T(r, c);
}
static std::vector<PLANE> make_planes(const int *offsets, const int *indices, const float *params, int np) {
  std::vector<PLANE> planes;
  for (int i = 0; i < np; ++i) {
    PLANE pl(indices + offsets[i], indices + offsets[i + 1]);
    pl.normal = Eigen::Vector3f(params[4 * i], params[4 * i + 1], params[4 * i + 2]);
    pl.d = params[4 * i + 3];
    planes.push_back(pl);
  }
  return planes;
}
static void put_planes(const std::string &prefix, const std::vector<PLANE> &planes) {
  std::vector<int> off(1, 0), idx;
  std::vector<float> par;
  for (size_t i = 0; i < planes.size(); ++i) {
    idx.insert(idx.end(), planes[i].begin(), planes[i].end());
    off.push_back((int) idx.size());
    par.push_back(planes[i].normal.x()); par.push_back(planes[i].normal.y());
    par.push_back(planes[i].normal.z()); par.push_back(planes[i].d);
  }
  put(prefix + "plane_offsets", off);
  put(prefix + "plane_indices", idx);
  put(prefix + "plane_params", par);
}

// ---- the MatchingLines hook (called from the renamed copy of PLADE/plade.cpp:536) -----------------
static bool g_dump = false;
static void dump_side(const std::string &p, MatchInformation &m) {
  put_xyz(p + "ds", *m.points);
  std::vector<float> c(m.boundingCenter.data(), m.boundingCenter.data() + 3);
  put(p + "center", c);
  const std::vector<INTERSECTION_LINE> &lines = *m.pIntersectionLine;
  std::vector<float> lv; std::vector<int> lp;
  for (size_t i = 0; i < lines.size(); ++i) {
    for (int k = 0; k < 3; ++k) lv.push_back(lines[i].lineVec[k]);
    for (int k = 0; k < 3; ++k) lv.push_back(lines[i].linePoint[k]);
    lp.push_back(lines[i].supportPlanes.size() > 0 ? lines[i].supportPlanes[0] : -1);
    lp.push_back(lines[i].supportPlanes.size() > 1 ? lines[i].supportPlanes[1] : -1);
  }
  put(p + "lines", lv); put(p + "line_planes", lp);
  const std::vector<std::vector<NearstPointsTwoLine> > &tab = *m.pNearstPointsTwoLines;
  size_t L = tab.size();
  std::vector<double> len(L * L, 0.0); std::vector<float> p1(L * L * 3, 0.f), p2(L * L * 3, 0.f);
  for (size_t i = 0; i < L; ++i) for (size_t j = 0; j < L; ++j) {
    if (i == j) continue;
    len[i * L + j] = tab[i][j].length;
    for (int k = 0; k < 3; ++k) { p1[(i * L + j) * 3 + k] = tab[i][j].points1[k]; p2[(i * L + j) * 3 + k] = tab[i][j].points2[k]; }
  }
  put(p + "pair_len", len); put(p + "pair_p1", p1); put(p + "pair_p2", p2);
  std::vector<float> planes;
  for (size_t i = 0; i < m.pPlanes->size(); ++i) for (int k = 0; k < 4; ++k) planes.push_back((*m.pPlanes)[i][k]);
  put(p + "planes", planes);
  std::vector<int> off(1, 0); std::vector<float> pds, corners, centers, radii;
  for (size_t i = 0; i < m.pDownSamplePlanePoints->size(); ++i) {
    const CloudXYZ &c2 = *(*m.pDownSamplePlanePoints)[i];
    for (size_t j = 0; j < c2.size(); ++j) { pds.push_back(c2[j].x); pds.push_back(c2[j].y); pds.push_back(c2[j].z); }
    off.push_back((int) (pds.size() / 3));
    const std::vector<Eigen::Vector3f> &fc = (*m.pBoundingBoxFourCornerPoints)[i];
    for (size_t j = 0; j < 4; ++j) for (int k = 0; k < 3; ++k) corners.push_back(j < fc.size() ? fc[j][k] : 0.f);
    for (int k = 0; k < 3; ++k) centers.push_back((*m.pBoundingBoxCenterForEachPlane)[i][k]);
    radii.push_back((*m.pBoundingBoxRadiusForEachPlane)[i]);
  }
  put(p + "plane_ds_offsets", off); put(p + "plane_ds", pds); put(p + "plane_corners4", corners);
  put(p + "plane_center", centers); put(p + "plane_radius", radii);
}

static std::vector<MatchedResult> g_last_results;
static MatchInformation g_last_current, g_last_main;

void plade_oracle_MatchingLines_hook(MatchInformation current, MatchInformation main,
                                     std::vector<std::pair<int, int> > linesTobeMatched,
                                     std::vector<std::vector<int> > coarseMatches,
                                     std::vector<MatchedResult> &pMatchedResult, Parameter parameter) {
  if (g_dump) {
    dump_side("src_", current);
    dump_side("tgt_", main);
    std::vector<int> q;
    for (size_t i = 0; i < linesTobeMatched.size(); ++i) { q.push_back(linesTobeMatched[i].first); q.push_back(linesTobeMatched[i].second); }
    put("lines_to_match", q);
    const std::vector<PAIRLINE> &lf22 = (*parameter.mainLinesInformation)[0];
    std::vector<float> desc, v1, v2, p1, p2; std::vector<int> pr;
    for (size_t i = 0; i < lf22.size(); ++i) {
      for (int k = 0; k < 8; ++k) desc.push_back(lf22[i].descriptor[k]);
      for (int k = 0; k < 3; ++k) { v1.push_back(lf22[i].lineVec1[k]); v2.push_back(lf22[i].lineVec2[k]); p1.push_back(lf22[i].linePoints1[k]); p2.push_back(lf22[i].linePoints2[k]); }
      pr.push_back(lf22[i].originalIndex1); pr.push_back(lf22[i].originalIndex2);
    }
    put("tgt_db_desc", desc); put("tgt_db_vec1", v1); put("tgt_db_vec2", v2); put("tgt_db_p1", p1); put("tgt_db_p2", p2);
    put("tgt_db_pair", pr);
    std::vector<double> par;
    par.push_back(parameter.lengthThreshold); par.push_back(parameter.angleThreshold);
    par.push_back(parameter.cosAngleThreshold); par.push_back(parameter.maxCandidateResultNum);
    par.push_back(parameter.maxNeighbor); par.push_back(parameter.maxRadius);
    put("match_params", par);
  }
  MatchingLines(current, main, linesTobeMatched, coarseMatches, pMatchedResult, parameter);  // PLADE/util.cpp:31
  g_last_results = pMatchedResult;
  g_last_current = current;  // holds shared_ptrs to the ds clouds (kept alive)
  g_last_main = main;
  if (g_dump) {
    std::vector<float> R, T; std::vector<int> np, pl, off(1, 0);
    for (size_t i = 0; i < pMatchedResult.size(); ++i) {
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R.push_back(pMatchedResult[i].R(r, c));
      for (int r = 0; r < 3; ++r) T.push_back(pMatchedResult[i].T(r));
      np.push_back((int) pMatchedResult[i].matchedPlanes.size());
      for (size_t k = 0; k < pMatchedResult[i].matchedPlanes.size(); ++k) { pl.push_back(pMatchedResult[i].matchedPlanes[k].first); pl.push_back(pMatchedResult[i].matchedPlanes[k].second); }
      off.push_back((int) pl.size() / 2);
    }
    put("mr_R", R); put("mr_T", T); put("mr_nplanes", np); put("mr_planes", pl); put("mr_plane_offsets", off);
  }
}

// Verification of the recorded hypotheses with the reference's own ComputeOverlap
// (PLADE/util.h:612-647), driven exactly like PLADE/plade.cpp:545-564.
static void dump_verification(CloudPN::Ptr source_cloud, int n_src_planes) {
  const float average_space = average_spacing(source_cloud, 6);              // PLADE/plade.cpp:41
  float downSampleDistance = average_space * 4;                              // :46
  double w, h, d; Eigen::Vector3f center;
  CloudXYZ::Ptr src_ds = g_last_current.points, tgt_ds = g_last_main.points;
  ComputeBoundingBox<pcl::PointXYZ>(src_ds, center, w, h, d);               // :295
  double currentRadius = MAX(MAX(w, h), d) / 2;                              // :299
  put1<float>("average_space", average_space);
  put1<float>("downsample_distance", downSampleDistance);
  put1<double>("src_radius", currentRadius);
  pcl::search::KdTree<pcl::PointXYZ>::Ptr tgt_tree(new pcl::search::KdTree<pcl::PointXYZ>);
  tgt_tree->setInputCloud(tgt_ds);
  std::vector<float> overlap, score, centers;
  for (size_t i = 0; i < g_last_results.size(); ++i) {
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    T.block(0, 0, 3, 3) = g_last_results[i].R;
    T.block(0, 3, 3, 1) = g_last_results[i].T;
    CloudXYZ::Ptr trans(new CloudXYZ);
    pcl::transformPointCloud(*src_ds, *trans, T);
    Eigen::Vector3f cc = g_last_results[i].R * g_last_current.boundingCenter + g_last_results[i].T;
    pcl::search::KdTree<pcl::PointXYZ>::Ptr kd(new pcl::search::KdTree<pcl::PointXYZ>);
    kd->setInputCloud(trans);
    float ov;
    ComputeOverlap<pcl::PointXYZ>(kd, tgt_tree, cc, currentRadius, downSampleDistance, ov);
    overlap.push_back(ov);
    float sc = 0.2 * (g_last_results[i].matchedPlanes.size() / double(n_src_planes)) + 0.8 * ov;
    score.push_back(sc);
    for (int k = 0; k < 3; ++k) centers.push_back(cc[k]);
  }
  put("ver_overlap", overlap); put("ver_score", score); put("ver_center", centers);
}

extern "C" {

// Fixed-seed mode also rewinds the RANSAC library's lagged-Fibonacci cursor: rn_setseed()
// (3rd_party/ransac/MiscLib/Random.cpp:25-58) refills the buffer but leaves MiscLib::rn_point where the
// previous Detect() stopped, so without this a call's result depends on the calls made before it.
static void rewind_rng() { if (g_seed >= 0) MiscLib::rn_point = MiscLib_RN_BUFSIZE; }
void ref_set_seed(long seed) { g_seed = seed; rewind_rng(); }
void ref_clear_blobs() { g_blobs.clear(); }
const void *ref_blob(const char *name, size_t *nbytes) {
  std::map<std::string, std::vector<char> >::iterator it = g_blobs.find(name);
  if (it == g_blobs.end()) { *nbytes = 0; return 0; }
  *nbytes = it->second.size();
  return it->second.data();
}
int ref_blob_names(char *buf, size_t cap) {
  std::string s;
  for (std::map<std::string, std::vector<char> >::iterator it = g_blobs.begin(); it != g_blobs.end(); ++it) { s += it->first; s += "\n"; }
  if (s.size() + 1 > cap) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int) s.size();
}

// bool registration(T, target_file, source_file)  PLADE/plade.cpp:665 (swap rule included)
int ref_registration_files(const char *tgt, const char *src, float *out16) {
  rewind_rng();
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  bool ok = registration(T, std::string(tgt), std::string(src));
  to_rowmajor(T, out16);
  return ok ? 1 : 0;
}

// bool registration(T, target_cloud, source_cloud)  PLADE/plade.cpp:638 (no swap at this level)
int ref_registration_clouds(const float *tgt, size_t nt, const float *src, size_t ns, float *out16) {
  rewind_rng();
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  bool ok = registration(T, make_cloud(tgt, nt), make_cloud(src, ns));
  to_rowmajor(T, out16);
  return ok ? 1 : 0;
}

// bool registration(T, target_cloud, source_cloud, target_planes, source_planes)  PLADE/plade.cpp:31
// dump != 0: run the hooked copy and publish stage blobs (+ verification scores).
int ref_registration_planes(const float *tgt, size_t nt, const float *src, size_t ns,
                            const int *t_off, const int *t_idx, const float *t_par, int t_np,
                            const int *s_off, const int *s_idx, const float *s_par, int s_np,
                            float *out16, int dump) {
  Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
  CloudPN::Ptr tc = make_cloud(tgt, nt), sc = make_cloud(src, ns);
  std::vector<PLANE> tp = make_planes(t_off, t_idx, t_par, t_np), sp = make_planes(s_off, s_idx, s_par, s_np);
  bool ok;
  if (dump) {
    g_dump = true;
    g_last_results.clear();
    ok = plade_oracle_registration_hooked(T, tc, sc, tp, sp);
    g_dump = false;
    if (!g_last_results.empty()) dump_verification(sc, s_np);
    std::vector<float> t16(16); to_rowmajor(T, t16.data()); put("final_T", t16);
    g_last_current = MatchInformation(); g_last_main = MatchInformation();
  } else {
    ok = registration(T, tc, sc, tp, sp);
  }
  to_rowmajor(T, out16);
  return ok ? 1 : 0;
}

// std::vector<PLANE> extract(cloud, init_min_support)  PLADE/plade.cpp:602  -> blobs <prefix>plane_*
int ref_extract(const float *xyzn, size_t n, int init_min_support, const char *prefix) {
  rewind_rng();
  std::vector<PLANE> planes = extract(make_cloud(xyzn, n), init_min_support);
  put_planes(prefix, planes);
  return (int) planes.size();
}

// PlaneExtraction::detect(cloud, min_support, .005, .02, .8, .001)  PLADE/plane_extraction.cpp:173
int ref_detect(const float *xyzn, size_t n, int min_support, float dist_thresh, float bitmap_reso,
               float normal_thresh, float overlook_prob, const char *prefix) {
  rewind_rng();
  CloudPN::Ptr c = make_cloud(xyzn, n);
  std::vector<PLANE> planes = PlaneExtraction::detect(*c, min_support, dist_thresh, bitmap_reso, normal_thresh, overlook_prob);
  put_planes(prefix, planes);
  return (int) planes.size();
}

// float average_spacing(cloud, 6)  PLADE/util.cpp:1619
float ref_average_spacing(const float *xyzn, size_t n) { return average_spacing(make_cloud(xyzn, n), 6); }

// DownSamplePointCloud<PointXYZ|PointNormal>  PLADE/util.h:162-184 -> blob `name` (float xyz)
int ref_voxel_downsample(const float *pts, size_t n, int stride, float leaf, const char *name) {
  CloudXYZ::Ptr out(new CloudXYZ);
  int rc;
  if (stride == 6) { CloudPN::Ptr c = make_cloud(pts, n); rc = DownSamplePointCloud<pcl::PointNormal>(c, out, leaf, leaf, leaf); }
  else { CloudXYZ::Ptr c = make_xyz(pts, n); rc = DownSamplePointCloud<pcl::PointXYZ>(c, out, leaf, leaf, leaf); }
  if (rc != 0) return -1;
  put_xyz(name, *out);
  return (int) out->size();
}

// ComputeBoundingBox<PointXYZ>  PLADE/util.h:187-248 ; out: center[3], whd[3] (width,height,depth), corners[24]
int ref_bounding_box(const float *xyz, size_t n, float *center, double *whd, float *corners) {
  CloudXYZ::Ptr c = make_xyz(xyz, n);
  Eigen::Vector3f ctr; CloudXYZ cp;
  int rc = ComputeBoundingBox<pcl::PointXYZ>(c, ctr, whd[0], whd[1], whd[2], &cp);
  if (rc != 0) return rc;
  for (int k = 0; k < 3; ++k) center[k] = ctr[k];
  for (size_t i = 0; i < cp.size() && i < 8; ++i) { corners[3 * i] = cp[i].x; corners[3 * i + 1] = cp[i].y; corners[3 * i + 2] = cp[i].z; }
  return 0;
}

// ComputeIntersectionLineOfTwoPlanes  PLADE/util.cpp:626-676
int ref_plane_intersection(const float *p1, const float *p2, float *vec, float *pt) {
  Eigen::Vector4f a(p1[0], p1[1], p1[2], p1[3]), b(p2[0], p2[1], p2[2], p2[3]);
  Eigen::Vector3f v, p;
  int rc = ComputeIntersectionLineOfTwoPlanes(a, b, v, p);
  for (int k = 0; k < 3; ++k) { vec[k] = v[k]; pt[k] = p[k]; }
  return rc;
}

// ComputeNearstTwoPointsOfTwo3DLine  PLADE/util.cpp:1167-1229 (9x9 float cv::solve, DECOMP_SVD)
int ref_nearest_points_two_lines(const float *v1, const float *p1, const float *v2, const float *p2,
                                 float *q1, float *q2, double *len) {
  Eigen::Vector3f a(v1[0], v1[1], v1[2]), b(p1[0], p1[1], p1[2]), c(v2[0], v2[1], v2[2]), d(p2[0], p2[1], p2[2]), o1, o2;
  int rc = ComputeNearstTwoPointsOfTwo3DLine(a, b, c, d, o1, o2, *len);
  for (int k = 0; k < 3; ++k) { q1[k] = o1[k]; q2[k] = o2[k]; }
  return rc;
}

// ComputeIntersectionPointOf23DLine  PLADE/util.cpp:1461-1500 (6x5 float cv::solve)
int ref_line_line_intersection(const float *v1, const float *p1, const float *v2, const float *p2, float *out) {
  Eigen::Vector3f a(v1[0], v1[1], v1[2]), b(p1[0], p1[1], p1[2]), c(v2[0], v2[1], v2[2]), d(p2[0], p2[1], p2[2]), o;
  int rc = ComputeIntersectionPointOf23DLine(a, b, c, d, o);
  for (int k = 0; k < 3; ++k) out[k] = o[k];
  return rc;
}

// KdTreeSearchNDim<VectorXf,8>::find_neighbors(q, 0, 0.04, ...)  3rd_party/ann_1.1.2/include/ANN/ANN.h:979-1029
// exactly as PLADE/util.cpp:163 uses it.  Output CSR: offsets[nq+1] in blob "<name>_offsets",
// db indices "<name>_idx", ANN squared distances (double->float as the wrapper does) "<name>_dist".
int ref_match_descriptors(const float *db, int ndb, const float *q, int nq, double radius, const char *name) {
  std::vector<Eigen::VectorXf> pts(ndb, Eigen::VectorXf(8));
  KdTreeSearchNDim<Eigen::VectorXf, 8> tree;
  tree.begin();
  for (int i = 0; i < ndb; ++i) { for (int k = 0; k < 8; ++k) pts[i][k] = db[8 * i + k]; tree.add_point(&pts[i]); }
  tree.end();
  std::vector<int> off(1, 0), idx; std::vector<float> dist;
  std::vector<int> nb; std::vector<float> nd;
  for (int i = 0; i < nq; ++i) {
    Eigen::VectorXf v(8);
    for (int k = 0; k < 8; ++k) v[k] = q[8 * i + k];
    tree.find_neighbors(v, 0, radius, nb, nd);
    idx.insert(idx.end(), nb.begin(), nb.end());
    dist.insert(dist.end(), nd.begin(), nd.end());
    off.push_back((int) idx.size());
  }
  std::string n(name);
  put(n + "_offsets", off); put(n + "_idx", idx); put(n + "_dist", dist);
  return (int) idx.size();
}

// ComputeDescriptorVectorForPairLines(method22)  PLADE/util.cpp:533-602 ; desc[1..7], newLine1/2
void ref_pair_descriptor(const float *l1vec, const float *l2vec, const float *l1sp1, const float *l1sp2,
                         const float *l2sp1, const float *l2sp2, float *desc8, float *nl1, float *nl2) {
  INTERSECTION_LINE a, b;
  a.lineVec = Eigen::Vector3f(l1vec[0], l1vec[1], l1vec[2]);
  b.lineVec = Eigen::Vector3f(l2vec[0], l2vec[1], l2vec[2]);
  Eigen::Vector3f s11(l1sp1[0], l1sp1[1], l1sp1[2]), s12(l1sp2[0], l1sp2[1], l1sp2[2]);
  Eigen::Vector3f s21(l2sp1[0], l2sp1[1], l2sp1[2]), s22(l2sp2[0], l2sp2[1], l2sp2[2]);
  Eigen::VectorXf d(8); d.setZero();
  Eigen::Vector3f n1, n2;
  ComputeDescriptorVectorForPairLines(a, b, s11, s12, s21, s22, d, n1, n2, method22);
  for (int k = 0; k < 8; ++k) desc8[k] = d[k];
  for (int k = 0; k < 3; ++k) { nl1[k] = n1[k]; nl2[k] = n2[k]; }
}

// ComputeTransformationUsingTwoVecAndOnePoint  PLADE/util.cpp:604-624 (Eigen::umeyama, 3 points), batched.
// in: per item sv1,sv2,dv1,dv2,sp,tp (18 floats) ; out: R row-major 9 + T 3
void ref_transform_from_two_vecs(const float *in, int n, float *R9, float *T3) {
  for (int i = 0; i < n; ++i) {
    const float *p = in + 18 * i;
    Eigen::Vector3f a(p[0], p[1], p[2]), b(p[3], p[4], p[5]), c(p[6], p[7], p[8]), d(p[9], p[10], p[11]);
    Eigen::Vector3f sp(p[12], p[13], p[14]), tp(p[15], p[16], p[17]);
    Eigen::Matrix3f R; Eigen::Vector3f T;
    ComputeTransformationUsingTwoVecAndOnePoint(a, b, c, d, sp, tp, R, T);
    for (int r = 0; r < 3; ++r) for (int cc = 0; cc < 3; ++cc) R9[9 * i + 3 * r + cc] = R(r, cc);
    for (int r = 0; r < 3; ++r) T3[3 * i + r] = T(r);
  }
}

// ClusterTransformation  PLADE/util.cpp:1245-1277 ; labels[n] = cluster id in CEC emission order
int ref_cluster_transformations(const float *R9, const float *T3, int n, float dist_thresh, float ang_thresh, int *labels) {
  std::vector<Eigen::Matrix3f> Rs(n); std::vector<Eigen::Vector3f> Ts(n);
  for (int i = 0; i < n; ++i) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Rs[i](r, c) = R9[9 * i + 3 * r + c];
    for (int r = 0; r < 3; ++r) Ts[i](r) = T3[3 * i + r];
  }
  pcl::IndicesClusters clusters;
  ClusterTransformation(Rs, Ts, dist_thresh, ang_thresh, clusters);
  for (size_t c = 0; c < clusters.size(); ++c)
    for (size_t k = 0; k < clusters[c].indices.size(); ++k) labels[clusters[c].indices[k]] = (int) c;
  return (int) clusters.size();
}

// Eigen::SelfAdjointEigenSolver<Eigen::Matrix3f>, the solver ComputeBoundingBox uses (PLADE/util.h:199-200).
void ref_self_adjoint_eig3(const float *A9, float *w3, float *V9) {
  Eigen::Matrix3f A;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) A(r, c) = A9[3 * r + c];
  Eigen::SelfAdjointEigenSolver<Eigen::Matrix3f> es(A, Eigen::ComputeEigenvectors);
  Eigen::Matrix3f V = es.eigenvectors();
  for (int r = 0; r < 3; ++r) { w3[r] = es.eigenvalues()(r); for (int c = 0; c < 3; ++c) V9[3 * r + c] = V(r, c); }
}

// Penetration filter of MatchingLines, PLADE/util.cpp:466-511, for H hypotheses (12 floats each: R row-major,
// T): the loop over (source plane i1, target plane j1) as written there -- same statements, same order --
// around the reference's own AreTwoPlanesPenetrable (PLADE/util.cpp:1279-1458), pcl::transformPointCloud and
// pcl::search::KdTree.  flags[h] = 1 iff the reference would drop hypothesis h.
void ref_penetration_filter(const float *s_planes4, int Ps, const float *s_corners12, const float *s_centers3,
                            const float *s_pts, const int *s_off,
                            const float *t_planes4, int Pt, const float *t_corners12, const float *t_centers3,
                            const float *t_pts, const int *t_off,
                            const float *hyp12, int H, float lengthThreshold, float angleThreshold, unsigned char *flags) {
  std::vector<Eigen::Vector4f> planes(Ps), mainPlanes(Pt);
  std::vector<std::vector<Eigen::Vector3f> > sc(Ps), tc(Pt);
  std::vector<Eigen::Vector3f> scen(Ps), tcen(Pt);
  std::vector<CloudXYZ::Ptr> sp(Ps), tp(Pt);
  std::vector<pcl::search::KdTree<pcl::PointXYZ>::Ptr> ttree(Pt);
  for (int i = 0; i < Ps; ++i) {
    planes[i] = Eigen::Vector4f(s_planes4[4 * i], s_planes4[4 * i + 1], s_planes4[4 * i + 2], s_planes4[4 * i + 3]);
    for (int k = 0; k < 4; ++k) sc[i].push_back(Eigen::Vector3f(s_corners12[12 * i + 3 * k], s_corners12[12 * i + 3 * k + 1], s_corners12[12 * i + 3 * k + 2]));
    scen[i] = Eigen::Vector3f(s_centers3[3 * i], s_centers3[3 * i + 1], s_centers3[3 * i + 2]);
    sp[i] = make_xyz(s_pts + 3 * (size_t) s_off[i], s_off[i + 1] - s_off[i]);
  }
  for (int j = 0; j < Pt; ++j) {
    mainPlanes[j] = Eigen::Vector4f(t_planes4[4 * j], t_planes4[4 * j + 1], t_planes4[4 * j + 2], t_planes4[4 * j + 3]);
    for (int k = 0; k < 4; ++k) tc[j].push_back(Eigen::Vector3f(t_corners12[12 * j + 3 * k], t_corners12[12 * j + 3 * k + 1], t_corners12[12 * j + 3 * k + 2]));
    tcen[j] = Eigen::Vector3f(t_centers3[3 * j], t_centers3[3 * j + 1], t_centers3[3 * j + 2]);
    tp[j] = make_xyz(t_pts + 3 * (size_t) t_off[j], t_off[j + 1] - t_off[j]);
    ttree[j].reset(new pcl::search::KdTree<pcl::PointXYZ>);
    ttree[j]->setInputCloud(tp[j]);
  }
  for (int h = 0; h < H; ++h) {
    Eigen::Matrix3f R;
    Eigen::Vector3f T;
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R(r, c) = hyp12[12 * h + 3 * r + c]; T(r) = hyp12[12 * h + 9 + r]; }
    Eigen::Matrix4f transformation = Eigen::Matrix4f::Identity();
    transformation.block(0, 0, 3, 3) = R;
    transformation.block(0, 3, 3, 1) = T;
    bool isPentrable = false;
    pcl::search::KdTree<pcl::PointXYZ>::Ptr tempKdtree(new pcl::search::KdTree<pcl::PointXYZ>);
    for (int i1 = 0; i1 < Ps; i1++) {
      isPentrable = false;
      Eigen::Vector4f plane1;
      plane1.block(0, 0, 3, 1) = R * planes[i1].block(0, 0, 3, 1);
      plane1(3) = -(-planes[i1](3) + (plane1.block(0, 0, 3, 1).transpose() * T)(0));
      CloudXYZ::Ptr tempTransPoints(new CloudXYZ);
      pcl::transformPointCloud(*sp[i1], *tempTransPoints, transformation);
      pcl::PointCloud<pcl::PointXYZ> tempPclFourCorner;
      ExchnageBetweentPCLPointXYZwithEigenVector3f(tempPclFourCorner, sc[i1]);
      pcl::transformPointCloud(tempPclFourCorner, tempPclFourCorner, transformation);
      std::vector<Eigen::Vector3f> tempEigenFourCorner;
      ExchnageBetweentPCLPointXYZwithEigenVector3f(tempPclFourCorner, tempEigenFourCorner);
      tempKdtree->setInputCloud(tempTransPoints);
      Eigen::Vector3f currentCenter2Main = R * scen[i1] + T;
      for (int j1 = 0; j1 < Pt; j1++) {
        Eigen::Vector3f plane_A = mainPlanes[j1].block(0, 0, 3, 1);
        Eigen::Vector3f plane_B = plane1.block(0, 0, 3, 1);
        double center2PlaneDistance = (abs(plane_A.dot(currentCenter2Main) + mainPlanes[j1](3)) + abs(plane_B.dot(tcen[j1]) + plane1(3))) / 2;
        if (center2PlaneDistance < lengthThreshold && plane_B.dot(plane_A) > angleThreshold) continue;
        if (0 != AreTwoPlanesPenetrable(plane1, mainPlanes[j1], tempEigenFourCorner, tc[j1], tempKdtree, ttree[j1], isPentrable,
                                        lengthThreshold, 10, lengthThreshold / 2)) continue;
        if (isPentrable) break;
      }
      if (isPentrable) break;
    }
    flags[h] = isPentrable ? 1 : 0;
  }
}

// Verification body PLADE/plade.cpp:547-560 + ComputeOverlap PLADE/util.h:612-647 for H hypotheses.
// centers[3H] are the ball centres (R*c+T) as the caller computed them; out: overlap ratio (float) and
// the integer inlier count recovered from it is NOT exposed by the reference, so counts[] is the
// numerator re-counted with the same two radiusSearch calls.
void ref_compute_overlap(const float *src_ds, size_t ns, const float *tgt_ds, size_t nt,
                         const float *R9, const float *T3, const float *centers, int H,
                         float query_radius, float inlier_distance, float *overlap, int *counts) {
  CloudXYZ::Ptr s = make_xyz(src_ds, ns), t = make_xyz(tgt_ds, nt);
  pcl::search::KdTree<pcl::PointXYZ>::Ptr tgt_tree(new pcl::search::KdTree<pcl::PointXYZ>);
  tgt_tree->setInputCloud(t);
  for (int i = 0; i < H; ++i) {
    Eigen::Matrix4f T = Eigen::Matrix4f::Identity();
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T(r, c) = R9[9 * i + 3 * r + c]; T(r, 3) = T3[3 * i + r]; }
    CloudXYZ::Ptr trans(new CloudXYZ);
    pcl::transformPointCloud(*s, *trans, T);
    Eigen::Vector3f cc(centers[3 * i], centers[3 * i + 1], centers[3 * i + 2]);
    pcl::search::KdTree<pcl::PointXYZ>::Ptr kd(new pcl::search::KdTree<pcl::PointXYZ>);
    kd->setInputCloud(trans);
    ComputeOverlap<pcl::PointXYZ>(kd, tgt_tree, cc, query_radius, inlier_distance, overlap[i]);
    if (counts) {
      std::vector<int> nb; std::vector<float> nd;
      int cnt = 0;
      if (tgt_tree->radiusSearch(pcl::PointXYZ(cc(0), cc(1), cc(2)), query_radius, nb, nd) > 0) {
        CloudXYZ::Ptr ball(new CloudXYZ);
        ball->resize(nb.size());
        for (size_t k = 0; k < nb.size(); ++k) (*ball)[k] = (*t)[nb[k]];
        pcl::search::KdTree<pcl::PointXYZ> bt;
        bt.setInputCloud(ball);
        for (size_t k = 0; k < trans->size(); ++k)
          if (bt.radiusSearch(trans->at(k), inlier_distance, nb, nd, 1) > 0) cnt++;
      }
      counts[i] = cnt;
    }
  }
}

// ==== RANSAC building blocks (rows a5-a8 of SURVEY.md 8a): the reference's own classes, one call each ==================
// A RANSAC-library PointCloud as PlaneExtraction::detect builds it (PLADE/plane_extraction.cpp:186-197)
static void make_ransac_cloud(const float *xyzn, size_t n, ::PointCloud &pc) {
  pc.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const float *p = xyzn + 6 * i;
    pc[i] = Point(Vec3f(p[0], p[1], p[2]), Vec3f(p[3], p[4], p[5]));
    pc[i].index = i;
  }
}
static PlanePrimitiveShape *make_plane_shape(const float nrm[3], const float pos[3]) {
  return new PlanePrimitiveShape(Plane(Vec3f(pos[0], pos[1], pos[2]), Vec3f(nrm[0], nrm[1], nrm[2])));   // R/Plane.cpp:13-23
}

// PlanePrimitiveShape::Parameters (R/PlanePrimitiveShape.h:97-109) with the in-plane frame of
// HyperplaneCoordinateSystem::FromNormal (R/GfxTL/HyperplaneCoordinateSystem.h:81-93): uv[2n]; frame6 = u[3] v[3]
void ref_plane_parameters(const float nrm[3], const float pos[3], const float *xyz, size_t n, float *uv, float *frame6) {
  PlanePrimitiveShape *sh = make_plane_shape(nrm, pos);
  for (size_t i = 0; i < n; ++i) {
    std::pair<float, float> q;
    sh->Parameters(Vec3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), &q);
    uv[2 * i] = q.first; uv[2 * i + 1] = q.second;
  }
  Vec3f u = sh->getXDim(), v = sh->getYDim();
  for (int k = 0; k < 3; ++k) { frame6[k] = u[k]; frame6[3 + k] = v[k]; }
  sh->Release();
}

// BitmapPrimitiveShape::ConnectedComponent (R/BitmapPrimitiveShape.cpp:155-265; BuildBitmap R/BitmapPrimitiveShape.h:101-150,
// BitmapExtent / InBitmap R/PlanePrimitiveShape.cpp:192-207, closing + Components R/Bitmap.cpp:154,459,633) on the points
// idx[0..m) of the cloud: returns the size of the largest component and writes its members (ascending) to out_idx
long long ref_connected_component(const float *xyzn, size_t n, const float nrm[3], const float pos[3], const int *idx, size_t m,
                                  float bitmap_eps, int do_filtering, int *out_idx) {
  ::PointCloud pc;
  make_ransac_cloud(xyzn, n, pc);
  PlanePrimitiveShape *sh = make_plane_shape(nrm, pos);
  MiscLib::Vector<size_t> ind(m);
  for (size_t i = 0; i < m; ++i) ind[i] = (size_t) idx[i];
  size_t k = sh->ConnectedComponent(pc, bitmap_eps, &ind, do_filtering != 0);
  std::vector<int> mem(k);
  for (size_t i = 0; i < k; ++i) mem[i] = (int) ind[i];
  std::sort(mem.begin(), mem.end());
  for (size_t i = 0; i < k; ++i) out_idx[i] = mem[i];
  sh->Release();
  return (long long) k;
}

// Plane::LeastSquaresFit (R/Plane.h:66-74, R/Plane.cpp:169-176 -> GfxTL::Mean, CovarianceMatrix, Jacobi, R/GfxTL/Plane.h:58-95)
int ref_plane_ls_fit(const float *xyzn, size_t n, const int *idx, size_t m, float out_nrm[3], float out_pos[3]) {
  ::PointCloud pc;
  make_ransac_cloud(xyzn, n, pc);
  MiscLib::Vector<size_t> ind(m);
  for (size_t i = 0; i < m; ++i) ind[i] = (size_t) idx[i];
  Plane pl(Vec3f(0, 0, 0), Vec3f(0, 0, 1));
  bool ok = pl.LeastSquaresFit(pc, ind.begin(), ind.end());
  for (int k = 0; k < 3; ++k) { out_nrm[k] = pl.getNormal()[k]; out_pos[k] = pl.getPosition()[k]; }
  return ok ? 1 : 0;
}

// Candidate::WeightedScore (R/Candidate.cpp:77-87, weigh() R/ScoreComputer.h:10-13) over idx[0..m)
float ref_weighted_score(const float *xyzn, size_t n, const float nrm[3], const float pos[3], const int *idx, size_t m, float epsilon,
                         float normal_thresh) {
  ::PointCloud pc;
  make_ransac_cloud(xyzn, n, pc);
  PlanePrimitiveShape *sh = make_plane_shape(nrm, pos);
  Candidate c(sh, 0);
  c.Indices(new MiscLib::RefCounted<MiscLib::Vector<size_t> >);
  c.Indices()->Release();
  sh->Release();
  for (size_t i = 0; i < m; ++i) c.Indices()->push_back((size_t) idx[i]);
  return c.WeightedScore(pc, epsilon, normal_thresh);
}

// The acceptance chain of one candidate, RansacShapeDetector::Detect R/RansacShapeDetector.cpp:613-655, driven with the
// reference's own Candidate / octree / visitor classes on a cloud whose points are all unassigned (assigned[i] != -1
// marks points that already belong to a shape): GlobalScore at 3 eps (R/Candidate.h:284-292) -> ConnectedComponent at
// bitmapEps -> up to three LS refits accepted while GlobalWeightedScore improves.  Detect's Fit() is private; it is
// exactly initialShape.LSFit(...) for LS_FITTING (R/RansacShapeDetector.cpp:956-969), called here directly.
// Returns the support of the accepted plane; out_nrm / out_pos = its Plane, out_idx = its members (ascending),
// trace[0] = number of refits accepted, trace[1] = number of refits tried.
long long ref_refine_candidate(const float *xyzn, size_t n, const int *assigned, const float nrm[3], const float pos[3], float epsilon,
                               float normal_thresh, float bitmap_eps, unsigned int min_support, float out_nrm[3], float out_pos[3],
                               int *out_idx, int *trace) {
  ::PointCloud pc;
  make_ransac_cloud(xyzn, n, pc);
  GfxTL::AACube<GfxTL::Vector3Df> bcube;
  bcube.Bound(pc.begin(), pc.end());
  MiscLib::Vector<size_t> globalOctreeIndices(n);
  for (size_t i = 0; i < n; ++i) globalOctreeIndices[i] = i;
  IndexedOctreeType globalOctree;
  globalOctree.MaxBucketSize() = 20;
  globalOctree.MaxSubdivisionLevel() = 10;
  globalOctree.IndexedData(globalOctreeIndices.begin(), globalOctreeIndices.end(), pc.begin());
  globalOctree.Build(bcube);
  ScorePrimitiveShapeVisitor<FlatNormalThreshPointCompatibilityFunc, IndexedOctreeType> globalScoreVisitor(3 * epsilon, normal_thresh);
  MiscLib::Vector<int> shapeIndex(n, -1);
  if (assigned) for (size_t i = 0; i < n; ++i) shapeIndex[i] = assigned[i];
  globalScoreVisitor.SetShapeIndex(shapeIndex);

  PlanePrimitiveShape *sh = make_plane_shape(nrm, pos);
  Candidate cand(sh, 0);
  cand.Indices(new MiscLib::RefCounted<MiscLib::Vector<size_t> >);
  cand.Indices()->Release();
  sh->Release();
  cand.GlobalScore(globalScoreVisitor, globalOctree);
  cand.ConnectedComponent(pc, bitmap_eps);
  Candidate clone;
  cand.Clone(&clone);
  float oldScore, newScore;
  newScore = clone.GlobalWeightedScore(globalScoreVisitor, globalOctree, pc, 3 * epsilon, normal_thresh, bitmap_eps);
  size_t fittingIter = 0;
  int accepted = 0, tried = 0;
  do {
    ++fittingIter;
    oldScore = newScore;
    std::pair<size_t, float> score;
    PrimitiveShape *shape = clone.Shape()->LSFit(pc, epsilon, normal_thresh, clone.Indices()->begin(), clone.Indices()->end(), &score);
    if (shape) {
      ++tried;
      clone.Shape(shape);
      newScore = clone.GlobalWeightedScore(globalScoreVisitor, globalOctree, pc, 3 * epsilon, normal_thresh, bitmap_eps);
      size_t newSize = clone.Size();
      shape->Release();
      if (newScore > oldScore && newSize > min_support) { clone.Clone(&cand); ++accepted; }
    }
  } while (newScore > oldScore && fittingIter < 3);
  const Plane &pl = dynamic_cast<const PlanePrimitiveShape *>(cand.Shape())->Internal();
  for (int k = 0; k < 3; ++k) { out_nrm[k] = pl.getNormal()[k]; out_pos[k] = pl.getPosition()[k]; }
  std::vector<int> mem(cand.Indices()->size());
  for (size_t i = 0; i < mem.size(); ++i) mem[i] = (int) (*cand.Indices())[i];
  std::sort(mem.begin(), mem.end());
  for (size_t i = 0; i < mem.size(); ++i) out_idx[i] = mem[i];
  if (trace) { trace[0] = accepted; trace[1] = tried; }
  return (long long) mem.size();
}

// load_ply_cloud (PLADE/util.cpp:1505-1546 over PlyReader, PLADE/ply_reader.cpp:46-148): blob `name` = interleaved x y z nx ny nz
long long ref_load_ply(const char *path, const char *name) {
  CloudPN::Ptr c(new CloudPN);
  if (!load_ply_cloud(path, *c)) return -1;
  std::vector<float> v(c->size() * 6);
  for (size_t i = 0; i < c->size(); ++i) {
    const pcl::PointNormal &p = c->at(i);
    v[6 * i] = p.x; v[6 * i + 1] = p.y; v[6 * i + 2] = p.z; v[6 * i + 3] = p.normal_x; v[6 * i + 4] = p.normal_y; v[6 * i + 5] = p.normal_z;
  }
  put(name, v);
  return (long long) c->size();
}

}  // extern "C"
