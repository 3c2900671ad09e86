/*
 * ORACLE / TEST INFRASTRUCTURE ONLY — never linked into, imported by or called from the product.
 *
 * Plain-C restatement ("port") of the data-parallel loops of the PLADE hot path, one function per
 * loop, each citing the reference code it follows (paths relative to /root/reference/code/).  These
 * are the brute-force definitions the CUDA kernels are checked against; they are themselves pinned
 * against the reference's own compiled sources (oracle/_ref, built by oracle/Makefile from
 * /root/reference) in tests/test_oracle_cpu.py, and through them against the reference's golden
 * output sample_data/file_pairs_results.txt.
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction — the canonical arithmetic).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* FLANN L2_Simple<float> (3rd_party/flann/algorithms/dist.h:84-90): result += diff*diff, x y z */
static float l2simple(const float *a, const float *b) {
  float r = 0.f;
  for (int k = 0; k < 3; ++k) { float d = a[k] - b[k]; r += d * d; }
  return r;
}

/* Verification, PLADE/plade.cpp:547-560 + ComputeOverlap PLADE/util.h:612-647.
 * transform: 3rd_party/pcl-1.8.1/common/include/pcl/common/impl/transforms.hpp:69-71
 * radii:     float(double(r)*double(r)), kdtree_flann.hpp:193; strict '<' (result_set.h:479,582)
 * counts[h] = #{ s : exists t, dist2(t, c_h) < rball2 and dist2(R_h s + T_h, t) < rin2 } */
void oracle_verify_counts(const float *src, size_t ns, const float *tgt, size_t nt, const float *R9, const float *T3,
                          const float *centers, int H, float ball_radius, float inlier_dist, uint32_t *counts) {
  const float rball2 = (float) ((double) ball_radius * (double) ball_radius);
  const float rin2 = (float) ((double) inlier_dist * (double) inlier_dist);
  unsigned char *in_ball = (unsigned char *) malloc(nt ? nt : 1);
  for (int h = 0; h < H; ++h) {
    const float *R = R9 + 9 * h, *T = T3 + 3 * h, *c = centers + 3 * h;
    size_t nb = 0;
    for (size_t j = 0; j < nt; ++j) { in_ball[j] = l2simple(c, tgt + 3 * j) < rball2; nb += in_ball[j]; }
    uint32_t cnt = 0;
    if (nb) {
      for (size_t i = 0; i < ns; ++i) {
        const float *s = src + 3 * i;
        float p[3];
        p[0] = R[0] * s[0] + R[1] * s[1] + R[2] * s[2] + T[0];
        p[1] = R[3] * s[0] + R[4] * s[1] + R[5] * s[2] + T[1];
        p[2] = R[6] * s[0] + R[7] * s[1] + R[8] * s[2] + T[2];
        for (size_t j = 0; j < nt; ++j)
          if (in_ball[j] && l2simple(p, tgt + 3 * j) < rin2) { ++cnt; break; }
      }
    }
    counts[h] = cnt;
  }
  free(in_ball);
}

/* pcl::VoxelGrid::applyFilter, 3rd_party/pcl-1.8.1/filters/include/pcl/filters/impl/voxel_grid.hpp:214-437
 * (xyz accumulator: common/impl/accumulators.hpp:65-84).  The reference orders the members of a voxel
 * with an unstable std::sort; the canonical order here (and in the CUDA kernel) is ascending point
 * index.  pts: n x stride floats; out: up to n x 3; returns the number of voxels, -1 on bad input,
 * -2 when the leaf is too small (the reference then returns the input cloud unchanged). */
typedef struct { uint32_t key; uint32_t idx; } key_idx;
static int cmp_key_idx(const void *a, const void *b) {
  const key_idx *x = (const key_idx *) a, *y = (const key_idx *) b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
long long oracle_voxel_downsample(const float *pts, size_t n, int stride, float leaf, float *out) {
  if (n == 0 || !(leaf > 0)) return -1;
  float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) { float v = pts[i * stride + k]; if (v < mn[k]) mn[k] = v; if (v > mx[k]) mx[k] = v; }
  const float inv = 1.0f / leaf;
  int64_t dx = (int64_t) ((mx[0] - mn[0]) * inv) + 1, dy = (int64_t) ((mx[1] - mn[1]) * inv) + 1, dz = (int64_t) ((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t) 2147483647) return -2;
  int minb[3], maxb[3];
  for (int k = 0; k < 3; ++k) { minb[k] = (int) floorf(mn[k] * inv); maxb[k] = (int) floorf(mx[k] * inv); }
  int div0 = maxb[0] - minb[0] + 1, div1 = maxb[1] - minb[1] + 1;
  key_idx *kv = (key_idx *) malloc(sizeof(key_idx) * n);
  for (size_t i = 0; i < n; ++i) {
    int i0 = (int) (floorf(pts[i * stride] * inv) - (float) minb[0]);
    int i1 = (int) (floorf(pts[i * stride + 1] * inv) - (float) minb[1]);
    int i2 = (int) (floorf(pts[i * stride + 2] * inv) - (float) minb[2]);
    kv[i].key = (uint32_t) (i0 + i1 * div0 + i2 * div0 * div1);
    kv[i].idx = (uint32_t) i;
  }
  qsort(kv, n, sizeof(key_idx), cmp_key_idx);
  long long nv = 0;
  size_t i = 0;
  while (i < n) {
    size_t j = i;
    float s[3] = {0.f, 0.f, 0.f};
    while (j < n && kv[j].key == kv[i].key) {
      for (int k = 0; k < 3; ++k) s[k] += pts[(size_t) kv[j].idx * stride + k];
      ++j;
    }
    float c = (float) (j - i);
    for (int k = 0; k < 3; ++k) out[3 * nv + k] = s[k] / c;
    ++nv;
    i = j;
  }
  free(kv);
  return nv;
}

/* k smallest FLANN L2_Simple squared distances per query (the arithmetic behind average_spacing,
 * PLADE/util.cpp:1619-1648); out: nq x k ascending. */
void oracle_knn_sqdist(const float *pts, size_t n, int stride, const int *qidx, int nq, int k, float *out) {
  float *best = (float *) malloc(sizeof(float) * (k + 1));
  for (int q = 0; q < nq; ++q) {
    const float *qp = pts + (size_t) qidx[q] * stride;
    for (int j = 0; j < k; ++j) best[j] = 3.402823466e38f;
    for (size_t i = 0; i < n; ++i) {
      float d = l2simple(qp, pts + i * stride);
      if (d < best[k - 1]) {
        int j = k - 1;
        while (j > 0 && best[j - 1] > d) { best[j] = best[j - 1]; --j; }
        best[j] = d;
      }
    }
    memcpy(out + (size_t) q * k, best, sizeof(float) * k);
  }
  free(best);
}

/* average_spacing(cloud, 6), PLADE/util.cpp:1619-1648 (samples = 10000, not accurate) */
float oracle_average_spacing(const float *pts, size_t num, int stride) {
  const int k = 6, samples = 10000;
  if (num == 0) return 0.f;
  size_t step = 1;
  if (num > (size_t) samples) step = num / samples;
  int kk = num < (size_t) k ? (int) num : k;
  double total = 0.0;
  size_t total_count = 0;
  float d2[8];
  for (size_t i = 0; i < num; i += step) {
    int qi = (int) i;
    oracle_knn_sqdist(pts, num, stride, &qi, 1, kk, d2);
    if (kk <= 1) continue;
    double avg = 0.0;
    for (int j = 1; j < kk; ++j) avg += sqrtf(d2[j]);
    total += (avg / kk);
    ++total_count;
  }
  return (float) (total / total_count);
}

/* Descriptor radius search: KdTreeSearchNDim<.,8>::find_neighbors(q, 0, radius)
 * (3rd_party/ann_1.1.2/include/ANN/ANN.h:979-1029; leaf test src/kd_fix_rad_search.cpp:162-177).
 * sqRad = float(radius)*float(radius) widened to double; running double sum of squared differences
 * of float coordinates widened to double must never exceed sqRad.  Output per query ascending in
 * (dist, db index).  offsets[nq+1]; idx/dist2 sized by a first call with idx == NULL. */
typedef struct { double d; int i; } dist_idx;
static int cmp_dist_idx(const void *a, const void *b) {
  const dist_idx *x = (const dist_idx *) a, *y = (const dist_idx *) b;
  if (x->d != y->d) return x->d < y->d ? -1 : 1;
  return x->i < y->i ? -1 : (x->i > y->i);
}
long long oracle_match_descriptors(const float *db, int ndb, const float *q, int nq, float radius, int *offsets, int *idx,
                                   double *dist2) {
  const float sqf = radius * radius;
  const double sq = (double) sqf;
  dist_idx *tmp = (dist_idx *) malloc(sizeof(dist_idx) * (ndb ? ndb : 1));
  long long total = 0;
  offsets[0] = 0;
  for (int a = 0; a < nq; ++a) {
    int m = 0;
    for (int b = 0; b < ndb; ++b) {
      double dist = 0.0;
      int d = 0;
      for (; d < 8; ++d) {
        double t = (double) q[8 * a + d] - (double) db[8 * b + d];
        dist = dist + t * t;
        if (dist > sq) break;
      }
      if (d >= 8) { tmp[m].d = dist; tmp[m].i = b; ++m; }
    }
    qsort(tmp, m, sizeof(dist_idx), cmp_dist_idx);
    if (idx) for (int k = 0; k < m; ++k) { idx[total + k] = tmp[k].i; dist2[total + k] = tmp[k].d; }
    total += m;
    offsets[a + 1] = (int) total;
  }
  free(tmp);
  return total;
}

/* RANSAC point/plane compatibility (3rd_party/ransac/FlatNormalThreshPointCompatibilityFunc.h:15-22,
 * Plane::Distance 3rd_party/ransac/Plane.h:31, Vec3f::dot 3rd_party/ransac/basic.h:80-86):
 * unassigned && fabs(dist - n.p) < eps && fabs(n.n_i) >= normal_thresh.  xyzn: n x 6. */
void oracle_score_planes(const float *xyzn, size_t n, const int *assigned, const float *planes4, int n_planes, float eps,
                         float normal_thresh, uint32_t *counts, unsigned char *mask0) {
  for (int p = 0; p < n_planes; ++p) {
    const float *pl = planes4 + 4 * p;
    uint32_t c = 0;
    for (size_t i = 0; i < n; ++i) {
      const float *v = xyzn + 6 * i;
      int in = 0;
      if (!assigned || assigned[i] == -1) {
        float dp = pl[0] * v[0] + pl[1] * v[1] + pl[2] * v[2];
        float dist = fabsf(pl[3] - dp);
        if (dist < eps) {
          float dn = pl[0] * v[3] + pl[1] * v[4] + pl[2] * v[5];
          in = fabsf(dn) >= normal_thresh;
        }
      }
      if (p == 0 && mask0) mask0[i] = (unsigned char) in;
      c += in;
    }
    counts[p] = c;
  }
}

/* ClusterTransformation (PLADE/util.cpp:1245-1277) = pcl ConditionalEuclideanClustering
 * (segmentation/impl/conditional_euclidean_clustering.hpp:43-148): connected components of
 * |Ta-Tb|^2 < float(tol*tol) (L2_Simple) && |euler_a - euler_b|^2 < ang_thresh; label = smallest
 * member index.  euler = pcl::getEulerAngles (common/impl/eigen.hpp:664-669).  O(n^2), small n only. */
static int uf_find(int *p, int x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
void oracle_cluster_transforms(const float *R9, const float *T3, int n, float dist_thresh, float ang_thresh, int *labels) {
  const float tol2 = (float) ((double) dist_thresh * (double) dist_thresh);
  float *eu = (float *) malloc(sizeof(float) * 3 * (n ? n : 1));
  int *par = (int *) malloc(sizeof(int) * (n ? n : 1));
  for (int i = 0; i < n; ++i) {
    const float *R = R9 + 9 * i;
    eu[3 * i] = atan2f(R[7], R[8]);
    eu[3 * i + 1] = asinf(-R[6]);
    eu[3 * i + 2] = atan2f(R[3], R[0]);
    par[i] = i;
  }
  for (int a = 0; a < n; ++a)
    for (int b = a + 1; b < n; ++b) {
      if (!(l2simple(T3 + 3 * a, T3 + 3 * b) < tol2)) continue;
      float e0 = eu[3 * a] - eu[3 * b], e1 = eu[3 * a + 1] - eu[3 * b + 1], e2 = eu[3 * a + 2] - eu[3 * b + 2];
      float en = e0 * e0 + e1 * e1 + e2 * e2;
      if (!(en < ang_thresh)) continue;
      int ra = uf_find(par, a), rb = uf_find(par, b);
      if (ra != rb) { if (ra < rb) par[rb] = ra; else par[ra] = rb; }
    }
  for (int i = 0; i < n; ++i) labels[i] = uf_find(par, i);
  free(eu);
  free(par);
}
