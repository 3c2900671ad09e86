/*
 * plade_b200 — C ABI of the B200-native PLADE registration hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  The C++ overloads of
 * the reference (`bool registration(Eigen::Matrix4f&, ...)`, /root/reference/code/PLADE/plade.h:44-96)
 * are thin inline wrappers over these entry points (include/plade.h).
 *
 * Conventions
 *   - clouds: interleaved float records  x y z nx ny nz  (exactly the vertex record of the PLY files
 *     the reference reads, PLADE/ply_reader.cpp:46-148 -> PLADE/util.cpp:1505-1546).
 *   - TARGET always comes first, as in the reference (PLADE/plade.h:44-47).
 *   - out16: row-major 4x4 that maps SOURCE onto TARGET; set to identity first, overwritten only on
 *     success (PLADE/plade.cpp:696, :569-575).
 *   - return value: 1 = success, 0 = failure (the reference's bool); the message that the reference
 *     prints on std::cerr is also available through plade_last_error().  Nothing aborts or throws
 *     across the boundary.  The library needs a CUDA device: there is no CPU fallback.
 *   - a context is single-threaded and owns its CUDA stream and scratch memory (the reference is not
 *     re-entrant at all); use one context per host thread / per GPU.
 *   - planes: CSR offsets[np+1] + point indices + params[4*np] = (nx, ny, nz, d), n.x + d = 0
 *     (PLANE, PLADE/plane_extraction.h:44-50).
 */
#ifndef PLADE_B200_H
#define PLADE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plade_ctx plade_ctx;
typedef struct plade_cloud plade_cloud;   /* a cloud resident in HBM */
typedef void (*plade_allreduce_max_u64)(unsigned long long *value, void *user);

/* ---- context ---------------------------------------------------------------------------------- */
plade_ctx *plade_ctx_create(int device);          /* device < 0: current device; NULL on failure */
void plade_ctx_destroy(plade_ctx *ctx);
const char *plade_last_error(plade_ctx *ctx);
const char *plade_create_error(void);             /* why the last plade_ctx_create returned NULL */
int plade_device_count(void);                     /* visible CUDA devices (0 when there is none) */
/* tunables; names = fields of plade::Params (defaults are the reference's literals). 1 if known. */
int plade_set_param(plade_ctx *ctx, const char *name, double value);
/* hypothesis sharding for multi-GPU verification (PLADE/plade.cpp:547-564 iterations are independent):
 * this context verifies hypotheses h with h % world == rank and `reduce` (may be NULL when world == 1)
 * must all-reduce(MAX) one u64 across the ranks (NCCL in bench.py). */
void plade_set_shard(plade_ctx *ctx, int rank, int world, plade_allreduce_max_u64 reduce, void *user);
/* The same sharding with the collective inside the library (PLADE/plade.cpp:547-564 is the loop that is split):
 * every rank builds the same hypothesis list, rank r verifies h % world == r, the best (score bits << 32 | ~index) of each shard
 * is formed on the device and ONE ncclAllReduce(ncclUint64, ncclMax) on the context's stream picks the winner -- no inlier count
 * leaves the device.  A hash of the hypothesis list rides along: ranks that disagree on the list fail loudly.  libnccl.so.2 is
 * opened at run time (no link-time dependency; inside a torch process it is the NCCL torch already loaded).
 *   plade_nccl_unique_id     ncclGetUniqueId on rank 0; ship the 128 bytes to the other ranks out of band (MPI, torch, a file)
 *   plade_shard_init_nccl    ncclCommInitRank on the context's device (one process per GPU)
 *   plade_shard_init_nccl_all  one process, n contexts on n different GPUs: ncclCommInitAll, rank i = ctxs[i]; the n registrations
 *                              must then run concurrently (one host thread per context), as any collective does
 *   plade_shard_finalize     destroys the communicator and returns the context to unsharded operation
 * All return 1 on success, 0 on failure (plade_last_error / plade_create_error). */
int plade_nccl_unique_id(char out128[128]);
int plade_shard_init_nccl(plade_ctx *ctx, const char id128[128], int rank, int world);
int plade_shard_init_nccl_all(plade_ctx **ctxs, int n);
void plade_shard_finalize(plade_ctx *ctx);
long long plade_launch_count(plade_ctx *ctx);     /* kernels launched by this context so far */
/* seconds of the last registration: upload, planes, spacing, downsample, lines, descriptors, match,
 * hypotheses, penetration, verify, total, then the verification kernel of that call as timed with
 * CUDA events on the context stream: kernel_ms, hypotheses, source ds points, target ds points (15 values) */
int plade_stage_times(plade_ctx *ctx, double *out, int n);
/* Device time of one kernel family during the last plade_register_* call, measured with CUDA events recorded on
 * the launching stream around every launch: out = { total ms, launches, algorithmic bytes }.
 * kernel = "score_candidates" (K1a, 28 B per subsample point per pass), "refine_cluster" (K1b-d, 28 B per band
 * point per evaluation), "band_compact" (20 B per point) or "verify" (K5).  1 if known. */
int plade_kernel_times(plade_ctx *ctx, const char *kernel, double out[3]);
/* CUDA-event stopwatch on the context's own stream (the stream every kernel of this context is
 * launched on): start records an event, stop records a second one, waits for it and returns the
 * elapsed device time in milliseconds. */
int plade_timer_start(plade_ctx *ctx);
float plade_timer_stop_ms(plade_ctx *ctx);

/* ---- registration() — the four reference overloads -------------------------------------------- */
/* registration(T, target_file, source_file)                         PLADE/plade.cpp:665-707 */
int plade_register_files(plade_ctx *ctx, const char *target_ply, const char *source_ply, float out16[16]);
/* registration(T, target_cloud, source_cloud)                       PLADE/plade.cpp:638-662 */
int plade_register_clouds(plade_ctx *ctx, const float *tgt_xyzn, size_t nt, const float *src_xyzn, size_t ns,
                          float out16[16]);
/* registration(T, target_cloud, source_cloud, planes, planes)       PLADE/plade.cpp:31-580 */
int plade_register_with_planes(plade_ctx *ctx, const float *tgt_xyzn, size_t nt, const float *src_xyzn, size_t ns,
                               const int *t_off, const int *t_idx, const float *t_par, int t_np,
                               const int *s_off, const int *s_idx, const float *s_par, int s_np, float out16[16]);
/* registration(T, target_cloud, source_cloud, min_support, min_support)   PLADE/plade.cpp:583-599 */
int plade_register_min_support(plade_ctx *ctx, const float *tgt_xyzn, size_t nt, const float *src_xyzn, size_t ns,
                               int min_support_target, int min_support_source, float out16[16]);

/* Batch mode of the reference CLI (PLADE/main.cpp:97-159: a list of (target, source) file pairs registered one
 * after the other) spread over the GPUs of one box.  Pairs are independent: one worker thread +
 * context per entry of devices[] (NULL => 0..n_devices-1) takes the next unclaimed pair, and parses the
 * following pair's PLY files into page-locked memory while the current one is on the GPU.  No collective.
 * out16 = n_pairs row-major 4x4 (identity where ok[p] == 0), same per-pair semantics as plade_register_files.
 * Returns the number of successful pairs, or -1 when no device could be used. */
int plade_register_batch(const int *devices, int n_devices, const char *const *target_files,
                         const char *const *source_files, int n_pairs, float *out16, int *ok);
/* The worker contexts (streams, device scratch, page-locked PLY buffers) of plade_register_batch are kept for the
 * next call; this frees them.  Optional: process exit releases everything. */
void plade_batch_release(void);

/* HBM-resident variants (inputs uploaded once; used for the device-resident throughput figure) */
plade_cloud *plade_cloud_upload(plade_ctx *ctx, const float *xyzn, size_t n);
void plade_cloud_free(plade_ctx *ctx, plade_cloud *cloud);
size_t plade_cloud_size(const plade_cloud *cloud);
int plade_register_resident(plade_ctx *ctx, const plade_cloud *tgt, const plade_cloud *src, float out16[16]);
int plade_register_resident_with_planes(plade_ctx *ctx, const plade_cloud *tgt, const plade_cloud *src,
                                        const int *t_off, const int *t_idx, const float *t_par, int t_np,
                                        const int *s_off, const int *s_idx, const float *s_par, int s_np,
                                        float out16[16]);

/* ---- stage entry points (what the parity tests call) ------------------------------------------ */
/* extract(cloud, init_min_support) PLADE/plade.cpp:602-635 / PlaneExtraction::detect
 * PLADE/plane_extraction.cpp:173-200.  Returns the number of planes (>= 0) or -1; fetch them with
 * plade_planes_size / plade_planes_get. */
int plade_extract_planes(plade_ctx *ctx, const float *xyzn, size_t n, int init_min_support);
int plade_detect_planes(plade_ctx *ctx, const float *xyzn, size_t n, int min_support);
int plade_planes_size(plade_ctx *ctx, int *n_planes, long long *n_indices);
int plade_planes_get(plade_ctx *ctx, int *offsets, int *indices, float *params);
/* Plane consensus score (the RANSAC predicate, 3rd_party/ransac/FlatNormalThreshPointCompatibilityFunc.h
 * :15-22 with Plane::Distance 3rd_party/ransac/Plane.h:31): for each of n_planes (nx,ny,nz,dist with
 * n.x = dist) counts[p] = #{ i : assigned[i] == -1 && |dist - n.p_i| < eps && |n.n_i| >= normal_thresh };
 * if inlier_mask != NULL (n bytes) it receives the mask of plane 0. */
int plade_score_planes(plade_ctx *ctx, const float *xyzn, size_t n, const int *assigned, const float *planes4,
                       int n_planes, float eps, float normal_thresh, unsigned int *counts, unsigned char *inlier_mask);
/* Connected-component step of the RANSAC acceptance test (BitmapPrimitiveShape::ConnectedComponent,
 * 3rd_party/ransac/BitmapPrimitiveShape.cpp:155-205 with Bitmap.cpp:154,459,633): cross closing (dilate, erode) of
 * the ue x ve bitmap (row-major, ue fastest, non-zero = set), 8-connected labelling, mask (0/1) of the component
 * with the most pixels (first in raster order on ties).  ue * ve <= 2^20.  Returns 1, or 0 on error. */
int plade_largest_component(plade_ctx *ctx, const unsigned char *bitmap, int ue, int ve, unsigned char *mask);
/* The acceptance chain of ONE RANSAC candidate plane (RansacShapeDetector::Detect, 3rd_party/ransac/RansacShapeDetector.cpp:613-655:
 * Candidate::GlobalScore at 3 eps -> ConnectedComponent at bitmapEps -> up to three least-squares refits accepted while
 * Candidate::GlobalWeightedScore improves and the support stays above min_support), run on the device exactly as the detection
 * runs it (band_compact_kernel + refine_cluster_kernel).  normal / position: the candidate (Plane(pos, normal), R/Plane.cpp:13-23);
 * assigned may be NULL (all points free).  Out: the accepted plane, member_mask[n] (0/1), its support, the number of
 * evaluations and its gaussian-weighted score (Candidate::WeightedScore, R/Candidate.cpp:77-87).  Returns 1, or 0 on error. */
int plade_refine_candidate(plade_ctx *ctx, const float *xyzn, size_t n, const int *assigned, const float normal[3], const float position[3],
                           int min_support, float out_normal[3], float out_position[3], unsigned char *member_mask, long long *size,
                           int *evaluations, double *weighted_score);
/* Host-side restatements shared with the kernels (plade_b200/csrc/planefit.h; no device needed):
 * plade_plane_parameters  = HyperplaneCoordinateSystem::FromNormal (R/GfxTL/HyperplaneCoordinateSystem.h:81-93) + PlanePrimitiveShape::
 *                           Parameters (R/PlanePrimitiveShape.h:97-109): uv[2n], frame6 = u[3] v[3];
 * plade_plane_ls_fit      = Plane::LeastSquaresFit (R/Plane.h:66-74; member sums in double, see DESIGN.md deviations), xyz packed;
 * plade_bitmap_layout     = BitmapExtent / InBitmap (R/PlanePrimitiveShape.cpp:192-207): extent2 = (uextent, vextent), pixels[n]. */
void plade_plane_parameters(const float normal[3], const float position[3], const float *xyz, size_t n, float *uv, float frame6[6]);
int plade_plane_ls_fit(const float *xyz, size_t n, float out_normal[3], float out_position[3]);
void plade_bitmap_layout(float umin, float umax, float vmin, float vmax, float bitmap_eps, const float *uv, size_t n, long long extent2[2], int *pixels);
/* average_spacing(cloud, 6) PLADE/util.cpp:1619-1648 */
float plade_average_spacing(plade_ctx *ctx, const float *xyzn, size_t n);
/* DownSamplePointCloud / pcl::VoxelGrid  PLADE/util.h:162-184; xyz has `stride` floats per point
 * (3 or 6); out_xyz needs room for n*3 floats; returns the number of voxels or -1. */
long long plade_voxel_downsample(plade_ctx *ctx, const float *pts, size_t n, int stride, float leaf, float *out_xyz);
/* ComputeBoundingBox PLADE/util.h:187-248: center[3], whd[3] = width, height, depth, corners[24] */
int plade_bounding_box(plade_ctx *ctx, const float *xyz, size_t n, float *center, double *whd, float *corners);
/* ComputeNearstTwoPointsOfTwo3DLine PLADE/util.cpp:1167-1229 for n line pairs at once: lines12 holds
 * v1[3] p1[3] v2[3] p2[3] per pair; directions are normalised as the reference does, the 9x9 float
 * cv::solve(DECOMP_SVD) is restated on the GPU.  points6 = point1[3] point2[3]; length[i] = |point1-point2|
 * or -1 where the reference returns -1 (identical directions).  Returns 0 or -1. */
int plade_nearest_points_two_lines(plade_ctx *ctx, const float *lines12, int n, float *points6, double *length);
/* Penetration filter of MatchingLines, PLADE/util.cpp:466-511 around AreTwoPlanesPenetrable PLADE/util.cpp:
 * 1279-1458, for n_hyp candidate transforms (hyp12: R row-major, T).  Per side: planes4 (n, d), corners12 = the
 * 4 bounding-rectangle corners of each plane in the reference's order, centers3, the per-plane down-sampled
 * points xyz with offsets[P+1].  flags[h] = 1 iff the reference would drop hypothesis h.  Returns 0 or -1. */
int plade_penetration_filter(plade_ctx *ctx,
                             const float *src_planes4, int n_src_planes, const float *src_corners12, const float *src_centers3,
                             const float *src_xyz, const int *src_offsets,
                             const float *tgt_planes4, int n_tgt_planes, const float *tgt_corners12, const float *tgt_centers3,
                             const float *tgt_xyz, const int *tgt_offsets,
                             const float *hyp12, int n_hyp, float length_threshold, float angle_threshold, unsigned char *flags);
/* descriptor radius search, KdTreeSearchNDim<.,8>::find_neighbors(q, 0, radius) ANN.h:979-1029.
 * offsets[nq+1]; returns the total number of matches M (or -1); plade_match_results copies the M
 * (db index, squared distance) pairs, per query ascending in (distance, index). */
long long plade_match_descriptors(plade_ctx *ctx, const float *db8, int ndb, const float *q8, int nq, float radius,
                                  int *offsets);
int plade_match_results(plade_ctx *ctx, int *idx, double *dist2);
/* ComputeTransformationUsingTwoVecAndOnePoint PLADE/util.cpp:604-624, batched: in = n x 18 floats
 * (sv1 sv2 dv1 dv2 sourcePoint targetPoint); out R9 row-major, T3. */
int plade_transforms_from_matches(plade_ctx *ctx, const float *in18, int n, float *R9, float *T3);
/* ClusterTransformation PLADE/util.cpp:1245-1277: labels[i] = smallest index of i's cluster. */
int plade_cluster_transforms(plade_ctx *ctx, const float *R9, const float *T3, int n, float dist_thresh,
                             float ang_thresh, int *labels);
/* Verification, PLADE/plade.cpp:547-560 + ComputeOverlap PLADE/util.h:612-647: counts[h] = number of
 * source ds points with a target ds point (inside ball(centers[h], ball_radius)) closer than
 * inlier_dist after applying (R[h], T[h]).  xyz arrays are packed 3 floats per point. */
int plade_verify_hypotheses(plade_ctx *ctx, const float *src_ds_xyz, size_t ns, const float *tgt_ds_xyz, size_t nt,
                            const float *R9, const float *T3, const float *centers3, int H, float ball_radius,
                            float inlier_dist, unsigned int *counts);
/* Same with both ds clouds already resident (uploaded once by plade_verify_upload); times only the
 * kernel path.  Used by bench.py for the roofline of the dominant kernel. */
int plade_verify_upload(plade_ctx *ctx, const float *src_ds_xyz, size_t ns, const float *tgt_ds_xyz, size_t nt,
                        float inlier_dist);
int plade_verify_resident(plade_ctx *ctx, const float *R9, const float *T3, const float *centers3, int H,
                          float ball_radius, float inlier_dist, unsigned int *counts, float *kernel_ms);

/* BASELINE config 4: H given hypotheses against the clouds made resident by plade_verify_upload (on every rank), sharded over the
 * ranks of the context's NCCL communicator (world 1 without one): this rank verifies h % world == rank; best_index / best_count =
 * the hypothesis with the most inliers over ALL ranks (ties: lowest index), agreed by ncclAllReduce(ncclUint64, ncclMax);
 * device_ms = CUDA-event time of kernel + key + collective on the context stream.  Returns 1 or 0. */
int plade_verify_sharded(plade_ctx *ctx, const float *R9, const float *T3, const float *centers3, int H, float ball_radius,
                         float inlier_dist, int *best_index, unsigned int *best_count, float *device_ms);

/* ---- result / CLI extras ----------------------------------------------------------------
 * plade_last_report: JSON text describing the last registration of this context -- planes and supports per cloud, down-sampled
 * sizes, hypotheses verified, the winner, its matched plane pairs, its inlier count / overlap ratio / score: the numbers the
 * reference's ResultViewer lets a user judge by eye.  Valid until the next registration on the context.
 * plade_dump_planes_vg: the reference's save_vg (PLADE/util.cpp:1553-1616) -- an ASCII .vg vertex-group file of a cloud and its planes
 * (Mapple / Easy3D read it); group_parameters hold (nx, ny, nz, d) instead of the reference's zeros.  Host only.  1 on success. */
/* plade_ply_read: the PLY reader of the file overload (load_ply_cloud, PLADE/util.cpp:1505-1546): interleaved x y z nx ny nz into
 * out_xyzn (room for capacity_points records); out_xyzn == NULL only returns the number of points; -1 on failure.  Host only. */
long long plade_ply_read(const char *path, float *out_xyzn, size_t capacity_points);
const char *plade_last_report(plade_ctx *ctx);
int plade_dump_planes_vg(const float *xyzn, size_t n, const int *offsets, const int *indices, const float *params4, int n_planes, const char *path);

/* ---- debugging / parity dumps ------------------------------------------------------------------- */
void plade_set_debug(plade_ctx *ctx, int on);     /* record named stage blobs during registration */
const void *plade_debug_blob(plade_ctx *ctx, const char *name, size_t *nbytes);

#ifdef __cplusplus
}
#endif
#endif /* PLADE_B200_H */
