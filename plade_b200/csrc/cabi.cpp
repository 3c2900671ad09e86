// extern "C" boundary of libplade_b200.so (declared in include/plade_b200.h).  Every entry point
// catches C++ exceptions (CUDA errors included), records the message and returns the failure value:
// nothing throws or aborts across the ABI.
#include "../../include/plade_b200.h"
#include "pipeline.h"
#include "nccl_shard.h"
#include "planefit.h"
#include "ply.h"
#include "libm_flt32.h"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <iostream>
#include <memory>
#include <mutex>
#include <thread>

using namespace plade;

namespace plade {
void largest_component_device(Device &dev, const unsigned char *h_bitmap, int ue, int ve, unsigned char *h_mask);
void score_planes_full(Device &dev, const float4 *pos, const float4 *nrm, const int *assigned, size_t n, const float4 *d_planes,
                       int n_planes, float eps, float nthresh, unsigned int *d_counts, unsigned char *d_mask0);
}

struct plade_cloud {
  CloudDev c;
};

struct plade_ctx {
  std::unique_ptr<Registrar> reg;
  std::vector<PlaneRec> planes;          // result of the last extract/detect stage call
  std::vector<int> match_idx;
  std::vector<double> match_d2;
  CloudDev tmp_t, tmp_s;
  PinBuf<float> pin_t, pin_s;            // PLY records of the file overload (page-locked)
  // resident buffers for plade_verify_upload / plade_verify_resident
  DevBuf<float4> v_src, v_tgt;
  size_t v_ns = 0, v_nt = 0;
  TargetGrid v_grid;
  DevBuf<HypParams> v_hyp;
  DevBuf<unsigned int> v_counts;
  DevBuf<float4> stage_a, stage_b;
  DevBuf<int> stage_i;
  DevBuf<unsigned int> stage_u;
  DevBuf<unsigned char> stage_m;
  std::string err;
};

static std::string g_create_error;

// (a context may be driven from any host thread: make its device current first -- a new thread starts on device 0)
#define PLADE_TRY(ctx, fail, ...)                                                      \
  if (!(ctx)) return fail;                                                             \
  try { PLADE_CUDA(cudaSetDevice((ctx)->reg->dev.id)); __VA_ARGS__ } catch (const std::exception &e) {                                     \
    (ctx)->err = e.what();                                                             \
    (ctx)->reg->last_error = e.what();                                                 \
    std::cerr << "plade_b200: " << e.what() << std::endl;                              \
    return fail;                                                                       \
  }

static std::vector<PlaneRec> planes_from_csr(const int *off, const int *idx, const float *par, int np) {
  std::vector<PlaneRec> v(np > 0 ? np : 0);
  for (int i = 0; i < np; ++i) {
    v[i].idx.assign(idx + off[i], idx + off[i + 1]);
    v[i].n[0] = par[4 * i]; v[i].n[1] = par[4 * i + 1]; v[i].n[2] = par[4 * i + 2];
    v[i].d = par[4 * i + 3];
  }
  return v;
}

static void identity16(float *o) { for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? 1.f : 0.f; }

// host xyz (stride floats per point) -> device float4
static void upload_xyz(Registrar &r, const float *pts, size_t n, int stride, DevBuf<float4> &out) {
  std::vector<float4> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = make_float4(pts[i * stride], pts[i * stride + 1], pts[i * stride + 2], 0.f);
  float4 *d = out.ensure(std::max<size_t>(n, 1));
  if (n) PLADE_CUDA(cudaMemcpyAsync(d, h.data(), sizeof(float4) * n, cudaMemcpyHostToDevice, r.dev.stream));
  stream_sync(r.dev.stream);
}

extern "C" {

plade_ctx *plade_ctx_create(int device) {
  try {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
      g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (plade_b200 has no CPU fallback)";
      std::cerr << "plade_b200: " << g_create_error << std::endl;
      return nullptr;
    }
    plade_ctx *c = new plade_ctx;
    c->reg.reset(new Registrar(device));
    return c;
  } catch (const std::exception &e) {
    g_create_error = e.what();
    std::cerr << "plade_b200: " << e.what() << std::endl;
    return nullptr;
  }
}
void plade_ctx_destroy(plade_ctx *ctx) { delete ctx; }
const char *plade_last_error(plade_ctx *ctx) { return ctx ? (ctx->err.empty() ? ctx->reg->last_error.c_str() : ctx->err.c_str()) : ""; }
const char *plade_create_error(void) { return g_create_error.c_str(); }
int plade_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  return count;
}

int plade_set_param(plade_ctx *ctx, const char *name, double v) {
  if (!ctx) return 0;
  Params &p = ctx->reg->params;
  std::string n(name);
  if (n == "ransac_dist_thresh") p.ransac_dist_thresh = (float) v;
  else if (n == "ransac_bitmap_reso") p.ransac_bitmap_reso = (float) v;
  else if (n == "ransac_normal_thresh") p.ransac_normal_thresh = (float) v;
  else if (n == "ransac_prob") p.ransac_prob = (float) v;
  else if (n == "init_min_support") p.init_min_support = (int) v;
  else if (n == "min_planes") p.min_planes = (int) v;
  else if (n == "max_planes") p.max_planes = (int) v;
  else if (n == "min_allowed_support") p.min_allowed_support = (int) v;
  else if (n == "max_trials") p.max_trials = (int) v;
  else if (n == "detect_margin") p.detect_margin = v;
  else if (n == "ransac_batch") p.ransac_batch = (int) v;
  else if (n == "score_live") p.score_live = (int) v;
  else if (n == "detect_resume") p.detect_resume = (int) v;
  else if (n == "blocking_sync") set_blocking_sync((int) v);       // process-wide, see stream_sync
  else if (n == "kernel_clock") { ctx->reg->dev.clock.enabled = ctx->reg->dev2.clock.enabled = v != 0; }
  else if (n == "max_candidates") p.max_candidates = (int) v;
  else if (n == "descriptor_radius") p.descriptor_radius = v;
  else if (n == "seed") p.seed = (unsigned long long) v;
  else return 0;
  return 1;
}

void plade_set_shard(plade_ctx *ctx, int rank, int world, plade_allreduce_max_u64 reduce, void *user) {
  if (!ctx) return;
  if (world > 1 && !reduce && !ctx->reg->nccl) {
    // a shard without a reducer would silently return the best of the local shard only
    ctx->err = "plade_set_shard: world > 1 needs a reducer (or plade_shard_init_nccl)";
    std::cerr << "plade_b200: " << ctx->err << std::endl;
    world = 1; rank = 0;
  }
  ctx->reg->shard_rank = rank;
  ctx->reg->shard_world = world < 1 ? 1 : world;
  ctx->reg->allreduce = reduce;
  ctx->reg->allreduce_user = user;
}

static void fill_hyp(std::vector<HypParams> &hp, const float *R9, const float *T3, const float *c3, int H);

// PLY ingest of the file overload as a stage entry (ply.cpp; PLADE/util.cpp:1505-1546): out_xyzn == NULL -> only the point count
long long plade_ply_read(const char *path, float *out_xyzn, size_t capacity_points) {
  if (!path) return -1;
  std::vector<float> v;
  if (!load_ply_xyzn(path, v)) return -1;
  const size_t n = v.size() / 6;
  if (out_xyzn) {
    if (n > capacity_points) return -1;
    memcpy(out_xyzn, v.data(), v.size() * sizeof(float));
  }
  return (long long) n;
}

const char *plade_last_report(plade_ctx *ctx) { return ctx ? ctx->reg->report.c_str() : ""; }

// save_vg of the reference (PLADE/util.cpp:1553-1616), same text layout; group_parameters carry the plane (nx, ny, nz, d) where the
// reference writes zeros, and the colours are a fixed function of the plane index instead of rand()
int plade_dump_planes_vg(const float *xyzn, size_t n, const int *offsets, const int *indices, const float *params4, int n_planes, const char *path) {
  if (!path || (n && !xyzn) || n_planes < 0) return 0;
  FILE *f = fopen(path, "w");
  if (!f) { std::cerr << "could not open file: " << path << std::endl; return 0; }
  fprintf(f, "num_points: %zu\n", n);
  for (size_t i = 0; i < n; ++i) fprintf(f, "%.9g %.9g %.9g ", xyzn[6 * i], xyzn[6 * i + 1], xyzn[6 * i + 2]);
  fprintf(f, "\nnum_colors: 0\nnum_normals: %zu\n", n);
  for (size_t i = 0; i < n; ++i) fprintf(f, "%.9g %.9g %.9g ", xyzn[6 * i + 3], xyzn[6 * i + 4], xyzn[6 * i + 5]);
  fprintf(f, "\nnum_groups: %d\n", n_planes);
  for (int k = 0; k < n_planes; ++k) {
    fprintf(f, "group_type: 0\nnum_group_parameters: 4\ngroup_parameters: %.9g %.9g %.9g %.9g \ngroup_label: unknown\n", params4[4 * k], params4[4 * k + 1],
            params4[4 * k + 2], params4[4 * k + 3]);
    const unsigned h = 2654435761u * (unsigned) (k + 1);
    fprintf(f, "group_color: %.3f %.3f %.3f\n", 0.3 + 0.7 * ((h >> 8) & 255) / 255.0, 0.3 + 0.7 * ((h >> 16) & 255) / 255.0, 0.3 + 0.7 * ((h >> 24) & 255) / 255.0);
    fprintf(f, "group_num_point: %d\n", offsets[k + 1] - offsets[k]);
    for (int i = offsets[k]; i < offsets[k + 1]; ++i) fprintf(f, "%d ", indices[i]);
    fprintf(f, "\nnum_children: 0\n");
  }
  const bool ok = !ferror(f);
  fclose(f);
  return ok ? 1 : 0;
}

int plade_nccl_unique_id(char out128[128]) {
  try { nccl_unique_id(out128); return 1; } catch (const std::exception &e) { g_create_error = e.what(); std::cerr << "plade_b200: " << e.what() << std::endl; return 0; }
}
int plade_shard_init_nccl(plade_ctx *ctx, const char id128[128], int rank, int world) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    if (r.nccl) { nccl_comm_destroy(r.nccl); r.nccl = nullptr; }
    if (world < 1 || rank < 0 || rank >= world) throw std::runtime_error("plade_shard_init_nccl: bad rank / world");
    r.nccl = nccl_comm_init_rank(id128, rank, world);
    r.shard_rank = rank; r.shard_world = world;
    r.allreduce = nullptr; r.allreduce_user = nullptr;
    return 1;
  })
}
int plade_shard_init_nccl_all(plade_ctx **ctxs, int n) {
  if (!ctxs || n < 1) return 0;
  try {
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) { if (!ctxs[i]) return 0; devs[i] = ctxs[i]->reg->dev.id; }
    std::vector<ShardComm *> comms(n, nullptr);
    nccl_comm_init_all(comms.data(), devs.data(), n);
    for (int i = 0; i < n; ++i) {
      Registrar &r = *ctxs[i]->reg;
      if (r.nccl) nccl_comm_destroy(r.nccl);
      r.nccl = comms[i];
      r.shard_rank = i; r.shard_world = n;
      r.allreduce = nullptr; r.allreduce_user = nullptr;
    }
    return 1;
  } catch (const std::exception &e) {
    g_create_error = e.what();
    std::cerr << "plade_b200: " << e.what() << std::endl;
    return 0;
  }
}
void plade_shard_finalize(plade_ctx *ctx) {
  if (!ctx) return;
  Registrar &r = *ctx->reg;
  if (r.nccl) { cudaSetDevice(r.dev.id); nccl_comm_destroy(r.nccl); r.nccl = nullptr; }
  r.shard_rank = 0; r.shard_world = 1; r.allreduce = nullptr; r.allreduce_user = nullptr;
}

// Config 4: H given hypotheses against the resident down-sampled clouds (plade_verify_upload on every rank),
// sharded over the ranks of the context's NCCL communicator: this rank verifies h % world == rank, the winner (highest
// inlier count, ties -> lowest index) is agreed with one ncclAllReduce(ncclUint64, ncclMax).  No count leaves the device.
int plade_verify_sharded(plade_ctx *ctx, const float *R9, const float *T3, const float *centers3, int H, float ball_radius, float inlier_dist,
                         int *best_index, unsigned int *best_count, float *device_ms) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    cudaStream_t s = r.dev.stream;
    const int world = r.shard_world, rank = r.shard_rank;
    if (world > 1 && !r.nccl) throw std::runtime_error("plade_verify_sharded: no NCCL communicator (plade_shard_init_nccl)");
    std::vector<HypParams> hp;
    fill_hyp(hp, R9, T3, centers3, H);
    std::vector<HypParams> mine;
    for (int h = rank; h < H; h += world) mine.push_back(hp[h]);
    const int nm = (int) mine.size();
    HypParams *d_h = ctx->v_hyp.ensure(std::max(nm, 1));
    unsigned int *d_c = ctx->v_counts.ensure(std::max(nm, 1));
    unsigned long long *d_key = r.d_shard_key.ensure(4), *h_key = r.h_shard_key.ensure(4);
    if (nm) PLADE_CUDA(cudaMemcpyAsync(d_h, mine.data(), sizeof(HypParams) * nm, cudaMemcpyHostToDevice, s));
    PLADE_CUDA(cudaEventRecord(r.ev_user0, s));
    verify_hypotheses(r.dev, ctx->v_src.p, ctx->v_ns, ctx->v_grid, d_h, nm, ball_radius, inlier_dist, d_c);
    shard_best_key(r.dev, d_c, d_h, nm, rank, world, 1.0, 1.0, 1, d_key);
    if (world > 1) nccl_allreduce_max_u64(r.nccl, d_key, 1, s);
    PLADE_CUDA(cudaEventRecord(r.ev_user1, s));
    PLADE_CUDA(cudaMemcpyAsync(h_key, d_key, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    float ms = 0;
    PLADE_CUDA(cudaEventElapsedTime(&ms, r.ev_user0, r.ev_user1));
    if (device_ms) *device_ms = ms;
    const unsigned long long key = h_key[0];
    if (key == 0 && H > 0) {      // every count is zero: the lowest index wins
      if (best_index) *best_index = 0;
      if (best_count) *best_count = 0;
      return 1;
    }
    if (best_index) *best_index = H > 0 ? (int) (0xFFFFFFFFu - (unsigned int) (key & 0xFFFFFFFFull)) : -1;
    if (best_count) *best_count = (unsigned int) (key >> 32);
    return 1;
  })
}
long long plade_launch_count(plade_ctx *ctx) { return ctx ? ctx->reg->dev.launches.n : 0; }
int plade_stage_times(plade_ctx *ctx, double *out, int n) {
  if (!ctx) return 0;
  const StageTimes &t = ctx->reg->times;
  double v[15] = {t.upload, t.planes, t.spacing, t.downsample, t.lines, t.descriptors, t.match, t.hypotheses, t.penetration, t.verify, t.total,
                  t.verify_kernel_ms, t.verify_h, t.verify_ns, t.verify_nt};
  for (int i = 0; i < n && i < 15; ++i) out[i] = v[i];
  return 15;
}
int plade_kernel_times(plade_ctx *ctx, const char *kernel, double out[3]) {
  if (!ctx || !kernel || !out) return 0;
  out[0] = out[1] = out[2] = 0;
  const std::string k(kernel);
  const Registrar &r = *ctx->reg;
  const int kind = k == "score_candidates" ? KernelClock::kScoreCandidates : k == "refine_cluster" ? KernelClock::kRefineCluster
                   : k == "band_compact" ? KernelClock::kBandCompact : -1;
  if (kind >= 0) {
    out[0] = r.dev.clock.ms[kind];
    out[1] = (double) r.dev.clock.launches[kind];
    out[2] = r.dev.clock.bytes[kind];
    return 1;
  }
  if (k == "verify") {
    out[0] = r.times.verify_kernel_ms;
    out[1] = 1;
    out[2] = r.times.verify_h * 16.0 * r.times.verify_ns + 16.0 * r.times.verify_nt;
    return 1;
  }
  return 0;
}
int plade_timer_start(plade_ctx *ctx) {
  PLADE_TRY(ctx, 0, { PLADE_CUDA(cudaEventRecord(ctx->reg->ev_user0, ctx->reg->dev.stream)); return 1; })
}
float plade_timer_stop_ms(plade_ctx *ctx) {
  PLADE_TRY(ctx, -1.f, {
    Registrar &r = *ctx->reg;
    PLADE_CUDA(cudaEventRecord(r.ev_user1, r.dev.stream));
    PLADE_CUDA(cudaEventSynchronize(r.ev_user1));
    float ms = 0;
    PLADE_CUDA(cudaEventElapsedTime(&ms, r.ev_user0, r.ev_user1));
    return ms;
  })
}
void plade_set_debug(plade_ctx *ctx, int on) { if (ctx) { ctx->reg->debug = on != 0; if (on) ctx->reg->blobs.clear(); } }
const void *plade_debug_blob(plade_ctx *ctx, const char *name, size_t *nbytes) {
  *nbytes = 0;
  if (!ctx) return nullptr;
  auto it = ctx->reg->blobs.find(name);
  if (it == ctx->reg->blobs.end()) return nullptr;
  *nbytes = it->second.size();
  return it->second.data();
}

// ---- registration ---------------------------------------------------------------------------------------
int plade_register_clouds(plade_ctx *ctx, const float *tgt, size_t nt, const float *src, size_t ns, float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    Registrar &r = *ctx->reg;
    return r.register_host_clouds(tgt, nt, src, ns, ctx->tmp_t, ctx->tmp_s, out16) ? 1 : 0;
  })
}

// swap rule + registration + inverse-on-swap of the file overload (PLADE/plade.cpp:689-704) on loaded clouds
static int register_loaded(plade_ctx *ctx, const float *t, size_t nt, const float *s, size_t ns, float out16[16], bool announce) {
  bool switched = false;
  if (ns >= nt * 1.2f) {
    std::swap(t, s);
    std::swap(nt, ns);
    switched = true;
    if (announce) std::cout << "---->>> ATTENTION: target and source have been switched for efficiency <<<----" << std::endl;
  }
  Registrar &r = *ctx->reg;
  if (!r.register_host_clouds(t, nt, s, ns, ctx->tmp_t, ctx->tmp_s, out16)) { std::cerr << "registration failed" << std::endl; return 0; }
  if (switched) {
    // the reference calls Matrix4f::inverse() (general 4x4 inverse); here the general inverse of the 3x3 block
    // via its adjugate in double (R need not be exactly orthonormal in float), equal up to rounding
    double R[9], T[3];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[3 * i + j] = out16[4 * i + j]; T[i] = out16[4 * i + 3]; }
    double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    double inv[9] = {(R[4] * R[8] - R[5] * R[7]) / det, (R[2] * R[7] - R[1] * R[8]) / det, (R[1] * R[5] - R[2] * R[4]) / det,
                     (R[5] * R[6] - R[3] * R[8]) / det, (R[0] * R[8] - R[2] * R[6]) / det, (R[2] * R[3] - R[0] * R[5]) / det,
                     (R[3] * R[7] - R[4] * R[6]) / det, (R[1] * R[6] - R[0] * R[7]) / det, (R[0] * R[4] - R[1] * R[3]) / det};
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) out16[4 * i + j] = (float) inv[3 * i + j];
      out16[4 * i + 3] = (float) -(inv[3 * i] * T[0] + inv[3 * i + 1] * T[1] + inv[3 * i + 2] * T[2]);
    }
  }
  return 1;
}

int plade_register_files(plade_ctx *ctx, const char *target_ply, const char *source_ply, float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    std::cout << "target file: " << target_ply << std::endl;
    std::cout << "source file: " << source_ply << std::endl;
    if (file_extension(target_ply) != "ply" || file_extension(source_ply) != "ply") {
      std::cerr << "only PLY format is accepted" << std::endl;
      ctx->err = "only PLY format is accepted";
      return 0;
    }
    // the records are read straight into page-locked memory so the H2D copy is one DMA
    struct Pin {
      static float *alloc(size_t n, void *user) { return static_cast<PinBuf<float> *>(user)->ensure(std::max<size_t>(n, 1)); }
    };
    size_t nt = 0, ns = 0;
    if (!load_ply_xyzn_into(target_ply, &Pin::alloc, &ctx->pin_t, nt)) { std::cerr << "loading target point cloud failed" << std::endl; ctx->err = "loading target point cloud failed"; return 0; }
    if (!load_ply_xyzn_into(source_ply, &Pin::alloc, &ctx->pin_s, ns)) { std::cerr << "loading source point cloud failed" << std::endl; ctx->err = "loading source point cloud failed"; return 0; }
    return register_loaded(ctx, ctx->pin_t.p, nt, ctx->pin_s.p, ns, out16, true);
  })
}

// Batch mode of the reference CLI (PLADE/main.cpp:97-159) over the GPUs of one box: pairs are independent, so
// worker g (one host thread + one context on devices[g]) takes the next unclaimed pair from a shared counter;
// each worker has a loader thread that parses the NEXT pair's PLY files into its second pair of pinned
// buffers while the current pair is on the GPU.  No data-path collective; results are written by pair index.
// Worker state of plade_register_batch, kept across calls: creating a context, growing its scratch (~100 device
// allocations) and pinning the PLY buffers costs a few hundred milliseconds and synchronises the whole device, so a
// worker takes an idle state of its device from this pool and returns it when the batch is done.
namespace {
struct BatchSlot {
  PinBuf<float> t, s;
  size_t nt = 0, ns = 0;
  int pair = -1;          // -1: end of work
  bool loaded = false;    // both files parsed
  std::string why;
};
struct BatchWorkerState {
  int device = 0;
  plade_ctx *ctx = nullptr;
  BatchSlot slots[2];
};
std::mutex g_batch_pool_mutex;
std::vector<BatchWorkerState *> g_batch_pool;     // idle states (never destroyed implicitly: see plade_batch_release)
BatchWorkerState *batch_state_acquire(int device) {
  {
    std::lock_guard<std::mutex> l(g_batch_pool_mutex);
    for (size_t i = 0; i < g_batch_pool.size(); ++i)
      if (g_batch_pool[i]->device == device) {
        BatchWorkerState *st = g_batch_pool[i];
        g_batch_pool.erase(g_batch_pool.begin() + i);
        return st;
      }
  }
  plade_ctx *ctx = plade_ctx_create(device);
  if (!ctx) return nullptr;
  BatchWorkerState *st = new BatchWorkerState;
  st->device = device;
  st->ctx = ctx;
  return st;
}
void batch_state_release(BatchWorkerState *st) {
  std::lock_guard<std::mutex> l(g_batch_pool_mutex);
  g_batch_pool.push_back(st);
}
}  // namespace

void plade_batch_release(void) {
  std::vector<BatchWorkerState *> all;
  { std::lock_guard<std::mutex> l(g_batch_pool_mutex); all.swap(g_batch_pool); }
  for (BatchWorkerState *st : all) {
    cudaSetDevice(st->device);
    plade_ctx_destroy(st->ctx);
    delete st;
  }
}

int plade_register_batch(const int *devices, int n_devices, const char *const *target_files, const char *const *source_files,
                         int n_pairs, float *out16, int *ok) {
  if (n_pairs < 0 || n_devices < 1 || !target_files || !source_files || !out16 || !ok) return -1;
  for (int p = 0; p < n_pairs; ++p) { identity16(out16 + 16 * p); ok[p] = 0; }
  typedef BatchSlot Slot;
  struct Pin {
    static float *alloc(size_t n, void *user) { return static_cast<PinBuf<float> *>(user)->ensure(std::max<size_t>(n, 1)); }
  };
  std::atomic<int> next(0), successes(0), workers_up(0);
  std::mutex io;     // one pair's messages at a time
  const bool timing = getenv("PLADE_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  auto worker = [&](int g) {
    const double t_begin = now();
    double t_load = 0, t_reg = 0, t_wait = 0;
    int n_done = 0;
    int device = devices ? devices[g] : g;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return; }
    BatchWorkerState *state = batch_state_acquire(device);
    if (!state) return;
    plade_ctx *ctx = state->ctx;
    workers_up.fetch_add(1);
    Slot *slots = state->slots;
    std::mutex m;
    std::condition_variable cv;
    int ready[2] = {0, 0};    // 0 = free for the loader, 1 = filled for the worker
    std::thread loader([&] {
      cudaSetDevice(device);
      for (int k = 0;; k ^= 1) {
        { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return ready[k] == 0; }); }
        Slot &sl = slots[k];
        sl.pair = next.fetch_add(1);
        sl.loaded = false;
        sl.why.clear();
        if (sl.pair >= n_pairs) sl.pair = -1;
        else {
          const char *tf = target_files[sl.pair], *sf = source_files[sl.pair];
          const double tl0 = now();
          try {
            if (file_extension(tf) != "ply" || file_extension(sf) != "ply") sl.why = "only PLY format is accepted";
            else if (!load_ply_xyzn_into(tf, &Pin::alloc, &sl.t, sl.nt)) sl.why = "loading target point cloud failed";
            else if (!load_ply_xyzn_into(sf, &Pin::alloc, &sl.s, sl.ns)) sl.why = "loading source point cloud failed";
            else sl.loaded = true;
          } catch (const std::exception &e) { sl.why = e.what(); }
          t_load += now() - tl0;
        }
        { std::lock_guard<std::mutex> l(m); ready[k] = 1; }
        cv.notify_all();
        if (sl.pair < 0) break;
      }
    });
    const double t_setup = now() - t_begin;
    for (int k = 0;; k ^= 1) {
      const double tw0 = now();
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return ready[k] == 1; }); }
      t_wait += now() - tw0;
      Slot &sl = slots[k];
      if (sl.pair < 0) break;
      int good = 0;
      const double tr0 = now();
      if (!sl.loaded) {
        std::lock_guard<std::mutex> l(io);
        std::cerr << sl.why << " (pair " << sl.pair << ")" << std::endl;
      } else {
        try {
          ctx->err.clear();
          good = register_loaded(ctx, sl.t.p, sl.nt, sl.s.p, sl.ns, out16 + 16 * sl.pair, false);
        } catch (const std::exception &e) {
          std::lock_guard<std::mutex> l(io);
          std::cerr << "plade_b200: " << e.what() << " (pair " << sl.pair << ")" << std::endl;
          identity16(out16 + 16 * sl.pair);
          good = 0;
        }
      }
      t_reg += now() - tr0;
      ++n_done;
      ok[sl.pair] = good;
      if (good) successes.fetch_add(1);
      { std::lock_guard<std::mutex> l(m); ready[k] = 0; }
      cv.notify_all();
    }
    loader.join();
    const double t_work = now() - t_begin;
    batch_state_release(state);
    if (timing) {
      std::lock_guard<std::mutex> l(io);
      fprintf(stderr, "[plade batch worker %d on device %d] %d pairs; ms: context %.1f, waiting for the loader %.1f, registering %.1f, loader busy %.1f, teardown %.1f\n",
              g, device, n_done, 1e3 * t_setup, 1e3 * t_wait, 1e3 * t_reg, 1e3 * t_load, 1e3 * (now() - t_begin - t_work));
    }
  };
  std::vector<std::thread> th;
  for (int g = 0; g < n_devices; ++g) th.emplace_back(worker, g);
  for (std::thread &t : th) t.join();
  if (workers_up.load() == 0) {
    std::cerr << "plade_b200: no usable CUDA device for the batch: " << g_create_error << std::endl;
    return -1;
  }
  return successes.load();
}

int plade_register_with_planes(plade_ctx *ctx, const float *tgt, size_t nt, const float *src, size_t ns, const int *t_off,
                               const int *t_idx, const float *t_par, int t_np, const int *s_off, const int *s_idx,
                               const float *s_par, int s_np, float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    Registrar &r = *ctx->reg;
    r.upload(tgt, nt, ctx->tmp_t);
    r.upload(src, ns, ctx->tmp_s);
    return r.register_with_planes(ctx->tmp_t, ctx->tmp_s, planes_from_csr(t_off, t_idx, t_par, t_np),
                                  planes_from_csr(s_off, s_idx, s_par, s_np), out16) ? 1 : 0;
  })
}

int plade_register_min_support(plade_ctx *ctx, const float *tgt, size_t nt, const float *src, size_t ns, int ms_t, int ms_s,
                               float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    Registrar &r = *ctx->reg;
    r.upload(tgt, nt, ctx->tmp_t);
    r.upload(src, ns, ctx->tmp_s);
    return r.register_min_support(ctx->tmp_t, ctx->tmp_s, ms_t, ms_s, out16) ? 1 : 0;
  })
}

plade_cloud *plade_cloud_upload(plade_ctx *ctx, const float *xyzn, size_t n) {
  PLADE_TRY(ctx, nullptr, {
    plade_cloud *c = new plade_cloud;
    ctx->reg->upload(xyzn, n, c->c);
    return c;
  })
}
void plade_cloud_free(plade_ctx *, plade_cloud *cloud) { delete cloud; }
size_t plade_cloud_size(const plade_cloud *cloud) { return cloud ? cloud->c.n : 0; }

int plade_register_resident(plade_ctx *ctx, const plade_cloud *tgt, const plade_cloud *src, float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    if (!tgt || !src) return 0;
    return ctx->reg->register_clouds(tgt->c, src->c, out16) ? 1 : 0;
  })
}

int plade_register_resident_with_planes(plade_ctx *ctx, const plade_cloud *tgt, const plade_cloud *src, const int *t_off,
                                        const int *t_idx, const float *t_par, int t_np, const int *s_off, const int *s_idx,
                                        const float *s_par, int s_np, float out16[16]) {
  identity16(out16);
  PLADE_TRY(ctx, 0, {
    ctx->err.clear();
    if (!tgt || !src) return 0;
    return ctx->reg->register_with_planes(tgt->c, src->c, planes_from_csr(t_off, t_idx, t_par, t_np),
                                          planes_from_csr(s_off, s_idx, s_par, s_np), out16) ? 1 : 0;
  })
}

// ---- stages -------------------------------------------------------------------------------------------------
int plade_extract_planes(plade_ctx *ctx, const float *xyzn, size_t n, int init_min_support) {
  PLADE_TRY(ctx, -1, {
    ctx->reg->upload(xyzn, n, ctx->tmp_t);
    ctx->planes = ctx->reg->extract_planes(ctx->tmp_t, init_min_support);
    return (int) ctx->planes.size();
  })
}
int plade_detect_planes(plade_ctx *ctx, const float *xyzn, size_t n, int min_support) {
  PLADE_TRY(ctx, -1, {
    ctx->reg->upload(xyzn, n, ctx->tmp_t);
    ctx->planes = ctx->reg->detect_planes(ctx->tmp_t, min_support);
    return (int) ctx->planes.size();
  })
}
int plade_planes_size(plade_ctx *ctx, int *n_planes, long long *n_indices) {
  if (!ctx) return 0;
  *n_planes = (int) ctx->planes.size();
  long long t = 0;
  for (auto &p : ctx->planes) t += (long long) p.idx.size();
  *n_indices = t;
  return 1;
}
int plade_planes_get(plade_ctx *ctx, int *offsets, int *indices, float *params) {
  if (!ctx) return 0;
  int o = 0;
  offsets[0] = 0;
  for (size_t i = 0; i < ctx->planes.size(); ++i) {
    const PlaneRec &p = ctx->planes[i];
    memcpy(indices + o, p.idx.data(), sizeof(int) * p.idx.size());
    o += (int) p.idx.size();
    offsets[i + 1] = o;
    params[4 * i] = p.n[0]; params[4 * i + 1] = p.n[1]; params[4 * i + 2] = p.n[2]; params[4 * i + 3] = p.d;
  }
  return 1;
}

int plade_largest_component(plade_ctx *ctx, const unsigned char *bitmap, int ue, int ve, unsigned char *mask) {
  PLADE_TRY(ctx, 0, {
    largest_component_device(ctx->reg->dev, bitmap, ue, ve, mask);
    return 1;
  })
}

int plade_refine_candidate(plade_ctx *ctx, const float *xyzn, size_t n, const int *assigned, const float normal[3], const float position[3],
                           int min_support, float out_normal[3], float out_position[3], unsigned char *member_mask, long long *size,
                           int *evaluations, double *weighted_score) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    r.upload(xyzn, n, ctx->tmp_t);
    long long sz = 0;
    int ev = 0;
    double sc = 0;
    if (!r.refine_candidate_stage(ctx->tmp_t, assigned, normal, position, min_support, out_normal, out_position, member_mask, &sz, &ev, &sc)) {
      ctx->err = r.last_error;
      return 0;
    }
    if (size) *size = sz;
    if (evaluations) *evaluations = ev;
    if (weighted_score) *weighted_score = sc;
    return 1;
  })
}

// ---- host-side plane arithmetic (planefit.h, the same inline functions the kernels call) ------------------------------
void plade_plane_parameters(const float normal[3], const float position[3], const float *xyz, size_t n, float *uv, float frame6[6]) {
  float u[3], v[3];
  frame_from_normal(normal, u, v);
  for (int k = 0; k < 3; ++k) { frame6[k] = u[k]; frame6[3 + k] = v[k]; }
  for (size_t i = 0; i < n; ++i) plane_uv(xyz + 3 * i, position, u, v, uv[2 * i], uv[2 * i + 1]);
}
int plade_plane_ls_fit(const float *xyz, size_t n, float out_normal[3], float out_position[3]) {
  if (n < 1) return 0;
  // member sums in double, mean rounded to float, covariance about float(mean) -- as refine_candidate_dev forms them
  Eval e{(long long) n, 0, {0, 0, 0}, true};
  for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) e.sum[k] += (double) xyz[3 * i + k];
  const float mx = (float) (e.sum[0] / e.size), my = (float) (e.sum[1] / e.size), mz = (float) (e.sum[2] / e.size);
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    const double dx = (double) (xyz[3 * i] - mx), dy = (double) (xyz[3 * i + 1] - my), dz = (double) (xyz[3 * i + 2] - mz);
    c[0] += dx * dx; c[1] += dx * dy; c[2] += dx * dz; c[3] += dy * dy; c[4] += dy * dz; c[5] += dz * dz;
  }
  return fit_plane_from_cov(e, c, out_normal, out_position) ? 1 : 0;
}
void plade_bitmap_layout(float umin, float umax, float vmin, float vmax, float bitmap_eps, const float *uv, size_t n, long long extent2[2], int *pixels) {
  extent2[0] = bitmap_extent(umin, umax, bitmap_eps);
  extent2[1] = bitmap_extent(vmin, vmax, bitmap_eps);
  for (size_t i = 0; i < n; ++i)
    pixels[i] = bitmap_pixel(uv[2 * i], umin, bitmap_eps, (int) extent2[0]) + bitmap_pixel(uv[2 * i + 1], vmin, bitmap_eps, (int) extent2[1]) * (int) extent2[0];
}

int plade_score_planes(plade_ctx *ctx, const float *xyzn, size_t n, const int *assigned, const float *planes4, int n_planes,
                       float eps, float normal_thresh, unsigned int *counts, unsigned char *inlier_mask) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    cudaStream_t s = r.dev.stream;
    r.upload(xyzn, n, ctx->tmp_t);
    int *d_as = nullptr;
    if (assigned) {
      d_as = ctx->stage_i.ensure(std::max<size_t>(n, 1));
      PLADE_CUDA(cudaMemcpyAsync(d_as, assigned, sizeof(int) * n, cudaMemcpyHostToDevice, s));
    }
    float4 *d_pl = ctx->stage_a.ensure(std::max(n_planes, 1));
    PLADE_CUDA(cudaMemcpyAsync(d_pl, planes4, sizeof(float4) * n_planes, cudaMemcpyHostToDevice, s));
    unsigned int *d_c = ctx->stage_u.ensure(std::max(n_planes, 1));
    unsigned char *d_m = inlier_mask ? ctx->stage_m.ensure(std::max<size_t>(n, 1)) : nullptr;
    score_planes_full(r.dev, ctx->tmp_t.pos.p, ctx->tmp_t.nrm.p, d_as, n, d_pl, n_planes, eps, normal_thresh, d_c, d_m);
    PLADE_CUDA(cudaMemcpyAsync(counts, d_c, sizeof(unsigned int) * n_planes, cudaMemcpyDeviceToHost, s));
    if (inlier_mask && n) PLADE_CUDA(cudaMemcpyAsync(inlier_mask, d_m, n, cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    return 1;
  })
}

float plade_average_spacing(plade_ctx *ctx, const float *xyzn, size_t n) {
  PLADE_TRY(ctx, -1.f, {
    ctx->reg->upload(xyzn, n, ctx->tmp_t);
    return ctx->reg->average_spacing(ctx->tmp_t);
  })
}

long long plade_voxel_downsample(plade_ctx *ctx, const float *pts, size_t n, int stride, float leaf, float *out_xyz) {
  PLADE_TRY(ctx, -1, {
    Registrar &r = *ctx->reg;
    if (n == 0 || !(leaf > 0)) return -1;     // DownSamplePointCloud returns -1 (PLADE/util.h:165-167)
    upload_xyz(r, pts, n, stride, ctx->stage_a);
    size_t nv = voxel_downsample(r.dev, r.vox, ctx->stage_a.p, n, leaf, ctx->stage_b);
    std::vector<float4> h(nv);
    if (nv) PLADE_CUDA(cudaMemcpyAsync(h.data(), ctx->stage_b.p, sizeof(float4) * nv, cudaMemcpyDeviceToHost, r.dev.stream));
    stream_sync(r.dev.stream);
    for (size_t i = 0; i < nv; ++i) { out_xyz[3 * i] = h[i].x; out_xyz[3 * i + 1] = h[i].y; out_xyz[3 * i + 2] = h[i].z; }
    return (long long) nv;
  })
}

int plade_bounding_box(plade_ctx *ctx, const float *xyz, size_t n, float *center, double *whd, float *corners) {
  PLADE_TRY(ctx, -1, {
    Registrar &r = *ctx->reg;
    if (n == 0) return -1;
    upload_xyz(r, xyz, n, 3, ctx->stage_a);
    std::vector<ObbSeg> segs(1, ObbSeg{ctx->stage_a.p, (int) n, 0});
    std::vector<ObbResult> out;
    obb_segments(r.dev, r.obb_sc, segs, out);
    if (out[0].rc != 0) return out[0].rc;
    center[0] = out[0].center.x; center[1] = out[0].center.y; center[2] = out[0].center.z;
    whd[0] = out[0].width; whd[1] = out[0].height; whd[2] = out[0].depth;
    for (int k = 0; k < 8; ++k) { corners[3 * k] = out[0].corners[k].x; corners[3 * k + 1] = out[0].corners[k].y; corners[3 * k + 2] = out[0].corners[k].z; }
    return 0;
  })
}

int plade_nearest_points_two_lines(plade_ctx *ctx, const float *lines12, int n, float *points6, double *length) {
  PLADE_TRY(ctx, -1, {
    Registrar &r = *ctx->reg;
    std::vector<float> in, out;
    std::vector<int> slot(n, -1);
    int m = 0;
    for (int i = 0; i < n; ++i) {
      const float *q = lines12 + 12 * (size_t) i;
      V3 v1(q[0], q[1], q[2]), v2(q[6], q[7], q[8]);
      normalize(v1);
      normalize(v2);
      if (v1.x == v2.x && v1.y == v2.y && v1.z == v2.z) continue;
      const float e[12] = {v1.x, v1.y, v1.z, q[3], q[4], q[5], v2.x, v2.y, v2.z, q[9], q[10], q[11]};
      in.insert(in.end(), e, e + 12);
      slot[i] = m++;
    }
    out.resize((size_t) 6 * m);
    nearest_points_batch(r.dev, r.svd_sc, in.data(), m, out.data());
    for (int i = 0; i < n; ++i) {
      float *o = points6 + 6 * (size_t) i;
      if (slot[i] < 0) { for (int k = 0; k < 6; ++k) o[k] = 0.f; length[i] = -1; continue; }
      memcpy(o, &out[6 * (size_t) slot[i]], sizeof(float) * 6);
      length[i] = norm(V3(o[0], o[1], o[2]) - V3(o[3], o[4], o[5]));
    }
    return 0;
  })
}

static void fill_pen_side(Registrar &r, PenSide &ps, const float *planes4, int P, const float *corners12, const float *centers3,
                          const float *xyz, const int *off, DevBuf<float4> &stage) {
  ps.planes.resize(P); ps.corners4.resize(P); ps.center.resize(P);
  for (int i = 0; i < P; ++i) {
    ps.planes[i] = {planes4[4 * i], planes4[4 * i + 1], planes4[4 * i + 2], planes4[4 * i + 3]};
    for (int k = 0; k < 4; ++k) ps.corners4[i][k] = V3(corners12[12 * i + 3 * k], corners12[12 * i + 3 * k + 1], corners12[12 * i + 3 * k + 2]);
    ps.center[i] = V3(centers3[3 * i], centers3[3 * i + 1], centers3[3 * i + 2]);
  }
  ps.ds_start.assign(off, off + P + 1);
  upload_xyz(r, xyz, (size_t) off[P], 3, stage);
  ps.d_pts = stage.p;
}

int plade_penetration_filter(plade_ctx *ctx,
                             const float *src_planes4, int n_src_planes, const float *src_corners12, const float *src_centers3,
                             const float *src_xyz, const int *src_offsets,
                             const float *tgt_planes4, int n_tgt_planes, const float *tgt_corners12, const float *tgt_centers3,
                             const float *tgt_xyz, const int *tgt_offsets,
                             const float *hyp12, int n_hyp, float length_threshold, float angle_threshold, unsigned char *flags) {
  PLADE_TRY(ctx, -1, {
    Registrar &r = *ctx->reg;
    PenSide ps, pt;
    fill_pen_side(r, ps, src_planes4, n_src_planes, src_corners12, src_centers3, src_xyz, src_offsets, ctx->stage_a);
    fill_pen_side(r, pt, tgt_planes4, n_tgt_planes, tgt_corners12, tgt_centers3, tgt_xyz, tgt_offsets, ctx->stage_b);
    std::vector<unsigned char> pen;
    penetration_filter(r.dev, r.pen_sc, ps, pt, hyp12, n_hyp, length_threshold, angle_threshold, pen);
    for (int h = 0; h < n_hyp; ++h) flags[h] = pen[h];
    return 0;
  })
}

long long plade_match_descriptors(plade_ctx *ctx, const float *db8, int ndb, const float *q8, int nq, float radius, int *offsets) {
  PLADE_TRY(ctx, -1, {
    Registrar &r = *ctx->reg;
    std::vector<int> off;
    size_t m = match_descriptors(r.dev, r.match_sc, db8, ndb, q8, nq, radius, off, ctx->match_idx, ctx->match_d2);
    memcpy(offsets, off.data(), sizeof(int) * off.size());
    return (long long) m;
  })
}
int plade_match_results(plade_ctx *ctx, int *idx, double *dist2) {
  if (!ctx) return 0;
  if (!ctx->match_idx.empty()) {
    memcpy(idx, ctx->match_idx.data(), sizeof(int) * ctx->match_idx.size());
    memcpy(dist2, ctx->match_d2.data(), sizeof(double) * ctx->match_d2.size());
  }
  return 1;
}

int plade_transforms_from_matches(plade_ctx *ctx, const float *in18, int n, float *R9, float *T3) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    static_assert(sizeof(MatchPairIn) == 18 * sizeof(float), "MatchPairIn layout");
    std::vector<RigidOut> out;
    transforms_from_matches(r.dev, r.hyp_sc, reinterpret_cast<const MatchPairIn *>(in18), (size_t) n, out);
    for (int i = 0; i < n; ++i) { memcpy(R9 + 9 * i, out[i].R, sizeof(float) * 9); memcpy(T3 + 3 * i, out[i].T, sizeof(float) * 3); }
    return 1;
  })
}

int plade_cluster_transforms(plade_ctx *ctx, const float *R9, const float *T3, int n, float dist_thresh, float ang_thresh, int *labels) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    std::vector<RigidOut> rt(n);
    for (int i = 0; i < n; ++i) {
      memcpy(rt[i].R, R9 + 9 * i, sizeof(float) * 9);
      memcpy(rt[i].T, T3 + 3 * i, sizeof(float) * 3);
      // pcl::getEulerAngles on the float rotation (same expression as the K4a kernel)
      rt[i].euler[0] = atan2f_glibc(rt[i].R[7], rt[i].R[8]);
      rt[i].euler[1] = asinf_glibc(-rt[i].R[6]);
      rt[i].euler[2] = atan2f_glibc(rt[i].R[3], rt[i].R[0]);
      rt[i].pad = 0;
    }
    std::vector<int> label;
    cluster_transforms(r.dev, r.hyp_sc, rt, dist_thresh, ang_thresh, label);
    if (n) memcpy(labels, label.data(), sizeof(int) * n);
    return 1;
  })
}

static void fill_hyp(std::vector<HypParams> &hp, const float *R9, const float *T3, const float *c3, int H) {
  hp.resize(H);
  for (int i = 0; i < H; ++i) {
    memcpy(hp[i].R, R9 + 9 * i, sizeof(float) * 9);
    memcpy(hp[i].T, T3 + 3 * i, sizeof(float) * 3);
    memcpy(hp[i].c, c3 + 3 * i, sizeof(float) * 3);
    hp[i].pad = 0;
  }
}

int plade_verify_upload(plade_ctx *ctx, const float *src_ds_xyz, size_t ns, const float *tgt_ds_xyz, size_t nt, float inlier_dist) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    upload_xyz(r, src_ds_xyz, ns, 3, ctx->v_src);
    upload_xyz(r, tgt_ds_xyz, nt, 3, ctx->v_tgt);
    ctx->v_ns = ns; ctx->v_nt = nt;
    build_target_grid(r.dev, ctx->v_tgt.p, nt, inlier_dist, ctx->v_grid);
    stream_sync(r.dev.stream);
    return 1;
  })
}

int plade_verify_resident(plade_ctx *ctx, const float *R9, const float *T3, const float *centers3, int H, float ball_radius,
                          float inlier_dist, unsigned int *counts, float *kernel_ms) {
  PLADE_TRY(ctx, 0, {
    Registrar &r = *ctx->reg;
    cudaStream_t s = r.dev.stream;
    std::vector<HypParams> hp;
    fill_hyp(hp, R9, T3, centers3, H);
    HypParams *d_h = ctx->v_hyp.ensure(std::max(H, 1));
    unsigned int *d_c = ctx->v_counts.ensure(std::max(H, 1));
    if (H) PLADE_CUDA(cudaMemcpyAsync(d_h, hp.data(), sizeof(HypParams) * H, cudaMemcpyHostToDevice, s));
    cudaEvent_t e0, e1;
    PLADE_CUDA(cudaEventCreate(&e0));
    PLADE_CUDA(cudaEventCreate(&e1));
    PLADE_CUDA(cudaEventRecord(e0, s));
    verify_hypotheses(r.dev, ctx->v_src.p, ctx->v_ns, ctx->v_grid, d_h, H, ball_radius, inlier_dist, d_c);
    PLADE_CUDA(cudaEventRecord(e1, s));
    if (H) PLADE_CUDA(cudaMemcpyAsync(counts, d_c, sizeof(unsigned int) * H, cudaMemcpyDeviceToHost, s));
    stream_sync(s);
    float ms = 0;
    PLADE_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms) *kernel_ms = ms;
    return 1;
  })
}

int plade_verify_hypotheses(plade_ctx *ctx, const float *src_ds_xyz, size_t ns, const float *tgt_ds_xyz, size_t nt, const float *R9,
                            const float *T3, const float *centers3, int H, float ball_radius, float inlier_dist, unsigned int *counts) {
  if (!plade_verify_upload(ctx, src_ds_xyz, ns, tgt_ds_xyz, nt, inlier_dist)) return 0;
  return plade_verify_resident(ctx, R9, T3, centers3, H, ball_radius, inlier_dist, counts, nullptr);
}

}  // extern "C"
