// s2m_api.cu -- context, per-frame orchestration and the C ABI of include/s2m.h.
//
// Host-side rows of the reference kept here (all tiny, FP64, no contraction):
//   A  transformAssociateToMap   laserMapping.cpp:143-147
//   B  centre cube + window shift :313-508 (a window in WORLD cube coordinates;
//      shifting the reference's pointer arrays == moving that window, clearing a
//      recycled slab == dropping store entries whose cube left the window)
//   C  valid block                :510-530
//   U  transformUpdate            :149-153
// Everything else runs on the device (s2m_kernels.cu) on one stream, with a
// single device->host read-back (poses + counters) per call.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <sched.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/s2m.h"
#include "s2m_internal.h"

using namespace s2m;

namespace {

struct SlotHost {
  int cen[3] = {10, 10, 5};                 // laserCloudCenWidth/Height/Depth (:74-76)
  double q_wmap_wodom[4] = {0, 0, 0, 1};    // :116-117
  double t_wmap_wodom[3] = {0, 0, 0};
  double pose[7] = {0, 0, 0, 1, 0, 0, 0};   // parameters[7] (:110)
  int n_store[2] = {0, 0};                  // entries per class (after the last call)
  int val_lo[3] = {0, 0, 0}, val_hi[3] = {-1, -1, -1};
  bool force_pending_check = false;         // after an upload every entry is raw
  bool idx_valid = false;                   // the device cell index matches the store and the valid block
  int idx_built_n[2] = {0, 0};              // local-map points per class when the index was last built (table sizing)
  unsigned long long seq[2] = {0, 0};
  long long frames = 0;
};

struct HostTables {  // one pinned block, copied to the device in one go
  FrameDesc desc[kMaxBatch];
  int in_off[2 * kMaxBatch + 1];
  int lp_off[2 * kMaxBatch + 1];
  int so_off[2 * kMaxBatch + 1];
  int hash_off[2 * kMaxBatch + 1];   // static: the cell-table region of every segment (sized by cap_map_*)
  int idx_list[4 * kMaxBatch];       // segments whose cell index is rebuilt this frame, then their new table masks (at + G)
  int idx_poff[2 * kMaxBatch + 1];   // packed offsets of the local points of those segments
};

}  // namespace

// ---- NCCL, resolved at run time --------------------------------------------------------------
// Only the sharded-map mode needs it. dlopen("libnccl.so.2") returns the copy the process
// already loaded (torch's), so no second NCCL is ever linked in.
namespace nccl {
typedef struct { char internal[128]; } UniqueId;
typedef void* Comm;
enum { kSum = 0, kInt32 = 2, kFloat64 = 8 };  // ncclRedOp_t / ncclDataType_t values (stable ABI)
static int (*GetUniqueId)(UniqueId*) = nullptr;
static int (*CommInitRank)(Comm*, int, UniqueId, int) = nullptr;
static int (*AllReduce)(const void*, void*, size_t, int, int, Comm, cudaStream_t) = nullptr;
static int (*CommDestroy)(Comm) = nullptr;
static const char* (*GetErrorString)(int) = nullptr;
static bool load() {
  if (AllReduce) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  GetUniqueId = (int (*)(UniqueId*))dlsym(h, "ncclGetUniqueId");
  CommInitRank = (int (*)(Comm*, int, UniqueId, int))dlsym(h, "ncclCommInitRank");
  AllReduce = (int (*)(const void*, void*, size_t, int, int, Comm, cudaStream_t))dlsym(h, "ncclAllReduce");
  CommDestroy = (int (*)(Comm))dlsym(h, "ncclCommDestroy");
  GetErrorString = (const char* (*)(int))dlsym(h, "ncclGetErrorString");
  return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
}
}  // namespace nccl

// One host thread per lane of a multi-lane context (see s2m_params.lanes); jobs run in order.
struct LaneWorker {
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::deque<std::function<void()>> q;
  long long submitted = 0, completed = 0;
  bool quit = false;
  void start(int device) {
    th = std::thread([this, device] {
      cudaSetDevice(device);
      for (;;) {
        std::function<void()> j;
        {
          std::unique_lock<std::mutex> lk(m);
          cv.wait(lk, [&] { return !q.empty() || quit; });
          if (q.empty()) return;
          j = std::move(q.front());
          q.pop_front();
        }
        j();
        {
          std::lock_guard<std::mutex> lk(m);
          ++completed;
        }
        cv.notify_all();
      }
    });
  }
  long long submit(std::function<void()> j) {
    std::lock_guard<std::mutex> lk(m);
    q.push_back(std::move(j));
    cv.notify_all();
    return ++submitted;
  }
  void wait(long long ticket) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return completed >= ticket; });
  }
  void stop() {
    { std::lock_guard<std::mutex> lk(m); quit = true; }
    cv.notify_all();
    if (th.joinable()) th.join();
  }
};

// one frame in flight on a multi-lane context
struct PendingFrame {
  std::vector<std::vector<int>> co, so;  // per lane, rebased offsets
  std::vector<double> q, t;
  std::vector<int> active;
  bool has_active = false;
  double *q_out = nullptr, *t_out = nullptr;
  s2m_stats* stats = nullptr;
  int* status = nullptr;
  std::vector<int> rcs;
  std::vector<long long> tickets;
  int evslot = 0;
};

struct s2m_ctx {
  s2m_params P;
  // multi-lane context: the slots are split over `children` (each a complete context with its own
  // stream, buffers and host thread); the fields below this block are unused in the parent
  std::vector<s2m_ctx*> children;
  std::vector<LaneWorker*> workers;
  std::vector<cudaEvent_t> lane_done;   // [2 frames in flight][lanes]
  cudaEvent_t lane_start[2] = {nullptr, nullptr};
  cudaEvent_t ev_done = nullptr;        // blocking-sync stand-in for cudaStreamSynchronize (host_waits_block)
  std::deque<std::unique_ptr<PendingFrame>> in_flight;
  int frame_seq = 0;
  int lane_batch = 0;
  // leaf: double-buffered incoming clouds so the copy of the next frame overlaps this frame's work
  float4* in_buf[2] = {nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr};
  int stage_next = 0;
  Dev d;
  int cur = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::vector<SlotHost> slots;
  // scan-to-scan odometry contexts (s2m_odom_create): per-slot state of laserOdometry.cpp:67, :90-95
  struct OdomHost { bool inited = false; double q_w[4] = {0, 0, 0, 1}, t_w[3] = {0, 0, 0}, para[7] = {0, 0, 0, 1, 0, 0, 0}; };
  std::vector<OdomHost> odom;
  bool is_odom = false;
  long long od_cap = 0;  // points d.od_last can hold
  void* od_tmp = nullptr;
  size_t od_tmp_bytes = 0;
  HostTables* ht = nullptr;      // pinned
  HostTables* d_ht = nullptr;    // device copy (desc/in_off/lp_off/hash_off point into it)
  SlotOut* h_out = nullptr;      // pinned
  int* h_err = nullptr;          // pinned
  int* h_dsoff = nullptr;        // pinned [G+1]: down-sampled counts read back mid-frame
  cudaEvent_t ev_ds = nullptr;
  uint32_t* h_bbox = nullptr;    // pinned [G][6]: boxes of the incoming clouds (ordered-uint encoding)
  int* h_lpcnt = nullptr;        // pinned [2G]: local-map sizes of this frame, then how many of those points are still raw
  cudaEvent_t ev_bbox = nullptr;
  LmState* lm_trace = nullptr;   // [2][B] device
  LmState* h_lm = nullptr;       // pinned [2][B]
  std::vector<void*> allocs;
  // debug: guard bands around every device allocation (S2M_GUARD_BYTES > 0 at create), checked after every call
  size_t guard_bytes = 0;
  std::vector<GuardDesc> guards;
  std::string err;
  long long launches = 0;
  int hash_cap_total = 0;
  size_t bkt_total = 0;
  // profiling: CUDA events at phase boundaries of every frame (on the launching stream)
  bool profiling = false;
  bool count_candidates = false;  // profiling level 2: also count the map points in every query's 27 cells (K4's algorithmic bytes)
  std::vector<cudaEvent_t> ev_pool;
  std::vector<int> ev_phase;      // phase id that ENDS at this event (-1: frame start)
  size_t ev_used = 0;
  double phase_ms[S2M_N_PHASES] = {0};
  double k4_bytes = 0, k4_cand27 = 0;
  long long k4_launches = 0;
  int sm_count = 148;
  // sharded-map mode
  nccl::Comm comm = nullptr;
  std::vector<cudaEvent_t> ar_pool;
  size_t ar_used = 0;
  double ar_ms = 0;
  long long ar_count = 0;
};

// How the host threads wait for the device (three short waits per frame and lane): S2M_SYNC = spin (default: the
// driver's own policy, lowest latency; measured fine down to 4 cores for 6 lanes, gpurun_out/g10_sync.log) | block
// (the waits sleep on blocking-sync events: for hosts with fewer cores than ranks).
static thread_local int tl_parent_lanes = 0;  // lanes of the multi-lane context whose sub-contexts are being created
static bool host_waits_block(int lanes) {
  (void)lanes;
  const char* e = getenv("S2M_SYNC");
  return e && !strcmp(e, "block");
}
// wait for everything enqueued on the context's stream so far
static cudaError_t wait_stream(s2m_ctx* ctx) {
  if (!ctx->ev_done) return cudaStreamSynchronize(ctx->stream);
  cudaError_t e = cudaEventRecord(ctx->ev_done, ctx->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->ev_done);
}

static int prof_mark(s2m_ctx* ctx, int phase) {
  if (!ctx->profiling) return 0;
  if (ctx->ev_used == ctx->ev_pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return -1;
    ctx->ev_pool.push_back(e);
    ctx->ev_phase.push_back(0);
  }
  ctx->ev_phase[ctx->ev_used] = phase;
  cudaEventRecord(ctx->ev_pool[ctx->ev_used], ctx->stream);
  ctx->ev_used++;
  return 0;
}
static void prof_resolve(s2m_ctx* ctx) {
  cudaStreamSynchronize(ctx->stream);
  for (size_t i = 1; i < ctx->ev_used; ++i) {
    if (ctx->ev_phase[i] < 0) continue;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev_pool[i - 1], ctx->ev_pool[i]) == cudaSuccess) ctx->phase_ms[ctx->ev_phase[i]] += ms;
    if (ctx->ev_phase[i] == S2M_PHASE_ASSOCIATE) ctx->k4_launches++;
  }
  ctx->ev_used = 0;
}

// multi-lane contexts forward per-slot calls to the lane that holds the slot
#define ROUTE_SLOT(expr)                                                          \
  if (ctx && !ctx->children.empty()) {                                            \
    if (slot < 0 || slot >= ctx->P.batch) return S2M_ERR_ARG;                     \
    s2m_ctx* ch = ctx->children[slot / ctx->lane_batch];                          \
    slot = slot % ctx->lane_batch;                                                \
    const int rc_ = (expr);                                                       \
    if (rc_ < 0) ctx->err = ch->err;                                              \
    return rc_;                                                                   \
  }

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
      return S2M_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

template <typename T>
static int dev_alloc(s2m_ctx* ctx, T** p, size_t n) {
  void* q = nullptr;
  const size_t bytes = (std::max<size_t>(n, 1) * sizeof(T) + 15) / 16 * 16, g = ctx->guard_bytes;
  CK(cudaMalloc(&q, bytes + 2 * g));
  ctx->allocs.push_back(q);
  if (g) {  // pattern before and after the payload: an out-of-bounds store of any kernel shows up in s2m_debug_guard_check
    CK(cudaMemset(q, 0xA5, g));
    CK(cudaMemset((char*)q + g + bytes, 0xA5, g));
    ctx->guards.push_back(GuardDesc{(const uint32_t*)q, (const uint32_t*)((char*)q + g + bytes), (unsigned)(g / 4)});
  }
  *p = (T*)((char*)q + g);
  return 0;
}

static long long next_pow2(long long v) {
  long long p = 1;
  while (p < v) p <<= 1;
  return p;
}
// slots of a segment's cell table: a power of two >= 4 x points (<= 3 entries per point: cells + virtual x-neighbours), >= 1024
#ifndef S2M_HASH_MULT
#define S2M_HASH_MULT 4
#endif
static long long table_size(long long n_points) { return next_pow2(std::max<long long>(1024, S2M_HASH_MULT * n_points)); }

extern "C" void s2m_default_params(s2m_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->line_res = 0.4f;   // laserMapping.cpp:915
  p->plane_res = 0.8f;  // :916
  p->device = 0;
  p->batch = 1;
  p->cap_corner_in = 16384;
  p->cap_surf_in = 131072;
  p->cap_map_corner = 1 << 20;
  p->cap_map_surf = 1 << 21;
  p->shard_world = 1;
  p->lanes = 0;
}

extern "C" const char* s2m_strerror(int code) {
  switch (code) {
    case S2M_OK: return "ok";
    case S2M_MAP_TOO_SMALL: return "map corner and surf num are not enough (pose = odometry guess)";
    case S2M_ERR_CUDA: return "CUDA error";
    case S2M_ERR_ARG: return "bad argument";
    case S2M_ERR_CAPACITY: return "capacity exceeded";
    case S2M_ERR_RANGE: return "coordinates outside the supported lattice range";
    case S2M_ERR_NCCL: return "NCCL error";
    case S2M_ERR_IO: return "file could not be read or written";
    case S2M_ERR_INTERNAL: return "internal consistency check failed";
    default: return "unknown";
  }
}
extern "C" const char* s2m_last_error(s2m_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

static int parent_wait(s2m_ctx* ctx);
extern "C" void s2m_destroy(s2m_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->P.device);
  while (!ctx->in_flight.empty()) parent_wait(ctx);
  for (LaneWorker* w : ctx->workers) { w->stop(); delete w; }
  for (s2m_ctx* ch : ctx->children) s2m_destroy(ch);
  for (cudaEvent_t e : ctx->lane_done) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (ctx->lane_start[i]) cudaEventDestroy(ctx->lane_start[i]);
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
  }
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (void* p : ctx->allocs) cudaFree(p);
  if (ctx->ht) cudaFreeHost(ctx->ht);
  if (ctx->h_out) cudaFreeHost(ctx->h_out);
  if (ctx->h_err) cudaFreeHost(ctx->h_err);
  if (ctx->h_dsoff) cudaFreeHost(ctx->h_dsoff);
  if (ctx->ev_ds) cudaEventDestroy(ctx->ev_ds);
  if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
  if (ctx->h_bbox) cudaFreeHost(ctx->h_bbox);
  if (ctx->h_lpcnt) cudaFreeHost(ctx->h_lpcnt);
  if (ctx->ev_bbox) cudaEventDestroy(ctx->ev_bbox);
  if (ctx->h_lm) cudaFreeHost(ctx->h_lm);
  for (auto& e : ctx->ev_pool) cudaEventDestroy(e);
  for (auto& e : ctx->ar_pool) cudaEventDestroy(e);
  if (ctx->comm && nccl::CommDestroy) nccl::CommDestroy(ctx->comm);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

static int create_impl(s2m_ctx* ctx) {
  const s2m_params& P = ctx->P;
  CK(cudaSetDevice(P.device));
  CK(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, P.device));
  CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) CK(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
  ctx->stream = ctx->own_stream;
  if (const char* gb = getenv("S2M_GUARD_BYTES")) ctx->guard_bytes = (size_t)std::max(0, atoi(gb)) / 16 * 16;
  Dev& d = ctx->d;
  std::memset(&d, 0, sizeof(d));
  const int B = P.batch, G = 2 * B;
  d.B = B; d.G = G;
  {
    const float min_leaf = std::min(P.line_res, P.plane_res);
    int nvox = (int)(50.0f / min_leaf) + 3, b = 1;
    while ((1 << b) < nvox) ++b;
    d.vox_bits = std::min(b, 11);
    d.delta_pbits = std::max(3 * d.vox_bits, 18);
  }
  d.shard_world = P.shard_world;
  s2m_shard_slab(P.shard_rank, P.shard_world, &d.shard_lo, &d.shard_hi);
  d.inv_leaf[0] = 1.0f / P.line_res;   // pcl::VoxelGrid::setLeafSize: inverse_leaf_size = 1 / leaf (float)
  d.inv_leaf[1] = 1.0f / P.plane_res;
  const long long cap_in = (long long)B * ((long long)P.cap_corner_in + P.cap_surf_in);
  const long long cap_lp = (long long)B * ((long long)P.cap_map_corner + P.cap_map_surf);
  if (cap_in + cap_lp >= (1ll << 30)) { ctx->err = "capacities too large (packed index space is 30 bits)"; return S2M_ERR_ARG; }
  // cell-index tags hold a voxel coordinate inside its cube in 8 bits and a bucket number in 24
  if ((int)(50.0f / std::min(P.line_res, P.plane_res)) + 3 > 256) { ctx->err = "mapping resolutions below 0.2 m are not supported (voxel coordinate inside a cube > 8 bits)"; return S2M_ERR_ARG; }
  if (P.cap_map_corner >= (1 << 24) || P.cap_map_surf >= (1 << 24)) { ctx->err = "cap_map_* must be below 2^24 per slot"; return S2M_ERR_ARG; }
  // the per-frame arrival number of a raw point is a field of delta_pbits bits in the map-update sort key
  if (P.cap_corner_in >= (1 << d.delta_pbits) || P.cap_surf_in >= (1 << d.delta_pbits)) {
    ctx->err = "cap_corner_in / cap_surf_in must be below 2^" + std::to_string(d.delta_pbits) + " for these leaf sizes";
    return S2M_ERR_ARG;
  }
  d.cap_in = (int)cap_in; d.cap_lp = (int)cap_lp; d.cap_sort = (int)(cap_in + cap_lp);
  d.max_tiles = (P.cap_corner_in + P.cap_surf_in + 31) / 32 + 1;  // partial rows per slot: one per 32-query unit

  CK(cudaMallocHost((void**)&ctx->ht, sizeof(HostTables)));
  CK(cudaMallocHost((void**)&ctx->h_out, sizeof(SlotOut) * B));
  CK(cudaMallocHost((void**)&ctx->h_err, sizeof(int)));
  CK(cudaMallocHost((void**)&ctx->h_dsoff, sizeof(int) * (2 * kMaxBatch + 1)));
  const unsigned wait_flags = cudaEventDisableTiming | (host_waits_block(std::max(1, tl_parent_lanes)) ? cudaEventBlockingSync : 0);
  CK(cudaEventCreateWithFlags(&ctx->ev_ds, wait_flags));
  if (wait_flags & cudaEventBlockingSync) CK(cudaEventCreateWithFlags(&ctx->ev_done, wait_flags));
  CK(cudaMallocHost((void**)&ctx->h_bbox, sizeof(uint32_t) * 6 * 2 * kMaxBatch));
  CK(cudaMallocHost((void**)&ctx->h_lpcnt, sizeof(int) * 4 * kMaxBatch));
  CK(cudaEventCreateWithFlags(&ctx->ev_bbox, wait_flags));
  CK(cudaMallocHost((void**)&ctx->h_lm, sizeof(LmState) * 2 * B));
  std::memset(ctx->ht, 0, sizeof(HostTables));
  if (dev_alloc(ctx, &ctx->d_ht, 1)) return S2M_ERR_CUDA;
  d.desc = ctx->d_ht->desc; d.in_off = ctx->d_ht->in_off; d.lp_off = ctx->d_ht->lp_off; d.so_off = ctx->d_ht->so_off; d.hash_off = ctx->d_ht->hash_off;
  d.idx_list = ctx->d_ht->idx_list; d.idx_poff = ctx->d_ht->idx_poff;

  int rc = 0;
  rc |= dev_alloc(ctx, &d.st_base, G); rc |= dev_alloc(ctx, &d.st_cap, G);
  rc |= dev_alloc(ctx, &ctx->in_buf[0], d.cap_in); rc |= dev_alloc(ctx, &ctx->in_buf[1], d.cap_in);
  d.in_pts = ctx->in_buf[0];
  rc |= dev_alloc(ctx, &d.vkey, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.vkey2, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.vval, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.vval2, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.flag, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.scan, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.bbox, 6 * G);
  rc |= dev_alloc(ctx, &d.ds_pts, d.cap_in); rc |= dev_alloc(ctx, &d.ds_off, G + 1);
  for (int b = 0; b < 2; ++b) { rc |= dev_alloc(ctx, &d.st_key[b], d.cap_lp); rc |= dev_alloc(ctx, &d.st_pt[b], d.cap_lp); }
  rc |= dev_alloc(ctx, &d.st_n, G); rc |= dev_alloc(ctx, &d.st_n_new, G);
  rc |= dev_alloc(ctx, &d.rng_start, G * kCols); rc |= dev_alloc(ctx, &d.loc_off, G * (kCols + 1)); rc |= dev_alloc(ctx, &d.lp_cnt, 2 * G);
  rc |= dev_alloc(ctx, &d.knn_ticket, 1); rc |= dev_alloc(ctx, &d.nbr, (size_t)d.cap_in * 6);
  rc |= dev_alloc(ctx, &d.qs_key, d.cap_in); rc |= dev_alloc(ctx, &d.qs_key2, d.cap_in);
  rc |= dev_alloc(ctx, &d.qs_val2, d.cap_in);
  rc |= dev_alloc(ctx, &d.cand27, d.cap_in); rc |= dev_alloc(ctx, &d.knn_stats, 4);
  // persistent cell index: per segment a table of a power of two >= 4 x cap_map slots (load <= 1/4) and a pool of
  // cap_map four-entry buckets
  long long hcap = 0;
  for (int g = 0; g < G; ++g) { ctx->ht->hash_off[g] = (int)hcap; hcap += table_size(g < B ? P.cap_map_corner : P.cap_map_surf); }
  if (hcap >= (1ll << 31)) { ctx->err = "cell tables too large (4 entries per map point)"; return S2M_ERR_ARG; }
  ctx->ht->hash_off[G] = (int)hcap;
  ctx->hash_cap_total = (int)hcap;
  rc |= dev_alloc(ctx, &d.hash_tab, (size_t)hcap); rc |= dev_alloc(ctx, &d.hmask, G);
  // (a quarter more buckets than points: cells that two threads create at the same moment leak one until the next rebuild)
  std::vector<int> boff(G + 1, 0);
  for (int g = 0; g < G; ++g) {
    const long long c = g < B ? P.cap_map_corner : P.cap_map_surf;
    boff[g + 1] = boff[g] + (int)std::min<long long>(c + c / 4 + 1024, (1 << 24) - 1);  // (a sparse map has one cell, hence one bucket, per point)
  }
  ctx->bkt_total = (size_t)boff[G];
  rc |= dev_alloc(ctx, &d.bkt, ctx->bkt_total * kBktE); rc |= dev_alloc(ctx, &d.bnext, ctx->bkt_total);
  rc |= dev_alloc(ctx, &d.bkt_off, G + 1); rc |= dev_alloc(ctx, &d.bcnt, G); rc |= dev_alloc(ctx, &d.idx_soff, G + 1);
  rc |= dev_alloc(ctx, &d.rec, (size_t)d.cap_in * 6); rc |= dev_alloc(ctx, &d.rec_valid, d.cap_in);
  rc |= dev_alloc(ctx, &d.partials, (size_t)B * d.max_tiles * kPartial);
  rc |= dev_alloc(ctx, &d.lm, B); rc |= dev_alloc(ctx, &d.out, B); rc |= dev_alloc(ctx, &d.err_flag, 1);
  rc |= dev_alloc(ctx, &d.ticket, B);
  rc |= dev_alloc(ctx, &d.shard_sums, (size_t)B * kPartial); rc |= dev_alloc(ctx, &d.shard_counts, G);
  rc |= dev_alloc(ctx, &ctx->lm_trace, 2 * B);
  if (P.trace) {
    rc |= dev_alloc(ctx, &d.tr_idx, (size_t)2 * d.cap_in * 5); rc |= dev_alloc(ctx, &d.tr_d2, (size_t)2 * d.cap_in * 5);
    rc |= dev_alloc(ctx, &d.tr_used, (size_t)2 * d.cap_in);
  }
  rc |= dev_alloc(ctx, &d.dl_pt, d.cap_in);
  rc |= dev_alloc(ctx, &d.ins_key, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.ins_ckey, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.ins_pt, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.ins_cpt, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.run_off, G + 1);
  rc |= dev_alloc(ctx, &d.dead_n, G); rc |= dev_alloc(ctx, &d.dead_lo, (size_t)G * kValidCubes); rc |= dev_alloc(ctx, &d.dead_cum, (size_t)G * (kValidCubes + 1));
  rc |= dev_alloc(ctx, &d.upd_pos, d.cap_sort + 1);
  rc |= dev_alloc(ctx, &d.aflag, d.cap_sort + 1); rc |= dev_alloc(ctx, &d.ascan, d.cap_sort + 1);
  d.cub_tmp_bytes = cub_temp_bytes(d.cap_sort + 1, d.cap_lp + 1);
  rc |= dev_alloc(ctx, (char**)&d.cub_tmp, d.cub_tmp_bytes);
  if (rc) return S2M_ERR_CUDA;

  std::vector<int> base(G), cap(G);
  int acc = 0;
  for (int g = 0; g < G; ++g) {
    base[g] = acc;
    cap[g] = g < B ? P.cap_map_corner : P.cap_map_surf;
    acc += cap[g];
  }
  CK(cudaMemcpy(d.st_base, base.data(), sizeof(int) * G, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d.st_cap, cap.data(), sizeof(int) * G, cudaMemcpyHostToDevice));
  CK(cudaMemset(d.st_n, 0, sizeof(int) * G));
  CK(cudaMemset(d.st_n_new, 0, sizeof(int) * G));
  CK(cudaMemset(d.err_flag, 0, sizeof(int)));
  CK(cudaMemset(d.ticket, 0, sizeof(int) * B));
  CK(cudaMemset(d.knn_ticket, 0, sizeof(int)));
  CK(cudaMemset(d.knn_stats, 0, 4 * sizeof(unsigned long long)));
  CK(cudaMemset(d.cand27, 0, sizeof(int) * (size_t)d.cap_in));
  d.count_cand = 0;
  CK(cudaMemset(d.out, 0, sizeof(SlotOut) * B));
  CK(cudaMemset(d.lm, 0, sizeof(LmState) * B));
  CK(cudaMemset(d.ds_off, 0, sizeof(int) * (G + 1)));
  CK(cudaMemcpy(d.bkt_off, boff.data(), sizeof(int) * (G + 1), cudaMemcpyHostToDevice));
  CK(cudaMemset(d.hash_tab, 0xFF, sizeof(unsigned long long) * (size_t)ctx->hash_cap_total));
  CK(cudaMemset(d.bkt, 0xFF, sizeof(float4) * ctx->bkt_total * kBktE));
  CK(cudaMemset(d.bnext, 0xFF, sizeof(uint32_t) * ctx->bkt_total));
  CK(cudaMemset(d.bcnt, 0, sizeof(int) * G));
  CK(cudaMemset(d.hmask, 0, sizeof(int) * G));
  ctx->slots.assign(B, SlotHost());
  return S2M_OK;
}

// debug: number of guard words that no longer hold the pattern (0 = no out-of-bounds store so far)
extern "C" long long s2m_debug_knn_fallbacks(s2m_ctx* ctx) {
  if (!ctx) return S2M_ERR_ARG;
  long long n = 0;
  for (s2m_ctx* ch : ctx->children) {
    const long long c = s2m_debug_knn_fallbacks(ch);
    if (c < 0) return c;
    n += c;
  }
  if (!ctx->d.knn_stats) return n;
  if (cudaSetDevice(ctx->P.device) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return S2M_ERR_CUDA;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, ctx->d.knn_stats, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return S2M_ERR_CUDA;
  return n + (long long)v;
}
extern "C" int s2m_debug_guard_check(s2m_ctx* ctx) {
  if (!ctx) return S2M_ERR_ARG;
  int bad = 0;
  for (s2m_ctx* ch : ctx->children) {
    const int b = s2m_debug_guard_check(ch);
    if (b < 0) return b;
    bad += b;
  }
  if (ctx->guards.empty()) return bad;
  CK(cudaSetDevice(ctx->P.device));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t n = ctx->guards.size();
  GuardDesc* dg = nullptr;
  int* dbad = nullptr;
  CK(cudaMalloc((void**)&dg, sizeof(GuardDesc) * n));
  CK(cudaMalloc((void**)&dbad, sizeof(int)));
  CK(cudaMemcpy(dg, ctx->guards.data(), sizeof(GuardDesc) * n, cudaMemcpyHostToDevice));
  CK(cudaMemset(dbad, 0, sizeof(int)));
  launch_guard_check(dg, (int)n, dbad, ctx->stream);
  int h = 0;
  CK(cudaMemcpyAsync(&h, dbad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  cudaFree(dg);
  cudaFree(dbad);
  return bad + h;
}

extern "C" int s2m_create(const s2m_params* p, s2m_ctx** out) {
  if (!p || !out) return S2M_ERR_ARG;
  *out = nullptr;
  if (p->batch < 1 || p->lanes < 0 || p->lanes > 64 || p->batch > 64 * kMaxBatch || !(p->line_res > 0.03f) || !(p->plane_res > 0.03f) ||
      p->cap_corner_in < 1 || p->cap_surf_in < 1 || p->cap_map_corner < 1 || p->cap_map_surf < 1)
    return S2M_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || p->device < 0 || p->device >= ndev) return S2M_ERR_CUDA;
  s2m_ctx* ctx = new s2m_ctx();
  ctx->P = *p;
  if (ctx->P.shard_world < 1) ctx->P.shard_world = 1;
  {  // lanes: slots split over independent sub-contexts that run concurrently inside one call
    int lanes = std::max(0, p->lanes);
    if (p->batch > kMaxBatch) lanes = std::max(lanes, (p->batch + kMaxBatch - 1) / kMaxBatch);
    lanes = std::min(lanes, p->batch);
    if (lanes >= 1) {
      ctx->lane_batch = (p->batch + lanes - 1) / lanes;
      int rc = S2M_OK;
      for (int i = 0; i * ctx->lane_batch < p->batch && rc == S2M_OK; ++i) {
        s2m_params cp = *p;
        tl_parent_lanes = p->lanes;
        cp.lanes = 0;
        cp.batch = std::min(ctx->lane_batch, p->batch - i * ctx->lane_batch);
        s2m_ctx* ch = nullptr;
        rc = s2m_create(&cp, &ch);
        if (rc == S2M_OK) ctx->children.push_back(ch);
      }
      tl_parent_lanes = 0;
      if (rc == S2M_OK && (cudaSetDevice(p->device) != cudaSuccess ||
                           cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
                           cudaEventCreateWithFlags(&ctx->lane_start[0], cudaEventDisableTiming) != cudaSuccess ||
                           cudaEventCreateWithFlags(&ctx->lane_start[1], cudaEventDisableTiming) != cudaSuccess))
        rc = S2M_ERR_CUDA;
      if (rc != S2M_OK) { s2m_destroy(ctx); return rc; }
      ctx->stream = ctx->own_stream;
      for (size_t i = 0; i < 2 * ctx->children.size(); ++i) {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        ctx->lane_done.push_back(e);
      }
      for (size_t i = 0; i < ctx->children.size(); ++i) {
        LaneWorker* w = new LaneWorker();
        w->start(p->device);
        ctx->workers.push_back(w);
      }
      *out = ctx;
      return S2M_OK;
    }
  }
  int rc = create_impl(ctx);
  if (rc != S2M_OK) {
    fprintf(stderr, "s2m_create: %s\n", ctx->err.c_str());
    s2m_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return S2M_OK;
}

extern "C" int s2m_set_stream(s2m_ctx* ctx, void* cuda_stream) {
  if (!ctx) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return S2M_OK;
}

extern "C" long long s2m_launch_count(s2m_ctx* ctx) {
  if (!ctx) return 0;
  long long n = ctx->launches;
  for (s2m_ctx* ch : ctx->children) n += ch->launches;
  return n;
}

// ---- sharded map: allreduce of the per-rank sums (latency-bound: 32 doubles per slot) ----------
static int shard_allreduce(s2m_ctx* ctx, void* buf, size_t count, int dtype) {
  if (ctx->P.shard_world <= 1 || !ctx->comm) return S2M_OK;  // comm-less = single-rank debugging of the filters
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->profiling) {
    while (ctx->ar_pool.size() < ctx->ar_used + 2) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      ctx->ar_pool.push_back(e);
    }
    e0 = ctx->ar_pool[ctx->ar_used++]; e1 = ctx->ar_pool[ctx->ar_used++];
    CK(cudaEventRecord(e0, ctx->stream));
  }
  const int rc = nccl::AllReduce(buf, buf, count, dtype, nccl::kSum, ctx->comm, ctx->stream);
  if (rc != 0) { ctx->err = std::string("ncclAllReduce: ") + (nccl::GetErrorString ? nccl::GetErrorString(rc) : "?"); return S2M_ERR_NCCL; }
  if (ctx->profiling) CK(cudaEventRecord(e1, ctx->stream));
  return S2M_OK;
}

// ---- host rows A, B, C ------------------------------------------------------
static void row_A(SlotHost& s, const double q_wodom[4], const double t_wodom[3]) {
  quat_mul_exact(s.q_wmap_wodom, q_wodom, s.pose);
  double r[3];
  quat_rotate_exact(s.q_wmap_wodom, t_wodom[0], t_wodom[1], t_wodom[2], r);
  for (int i = 0; i < 3; ++i) s.pose[4 + i] = xdadd(r[i], s.t_wmap_wodom[i]);
}
static void row_U(SlotHost& s, const double q_wodom[4], const double t_wodom[3]) {
  double qi[4], r[3];
  quat_inverse_exact(q_wodom, qi);
  quat_mul_exact(s.pose, qi, s.q_wmap_wodom);
  quat_rotate_exact(s.q_wmap_wodom, t_wodom[0], t_wodom[1], t_wodom[2], r);
  for (int i = 0; i < 3; ++i) s.t_wmap_wodom[i] = xdsub(s.pose[4 + i], r[i]);
}
// rows B + C for a sensor at t: moves the window, fills the descriptor boxes
static void rows_BC(SlotHost& s, const double t[3], FrameDesc& fd) {
  const int W[3] = {kWinI, kWinJ, kWinK};
  int wc[3];
  for (int a = 0; a < 3; ++a) {
    wc[a] = cube_of(t[a]);
    int c = wc[a] + s.cen[a];
    while (c < 3) { c++; s.cen[a]++; }
    while (c >= W[a] - 3) { c--; s.cen[a]--; }
  }
  const int half[3] = {2, 2, 1};
  for (int a = 0; a < 3; ++a) {
    fd.win_lo[a] = -s.cen[a];
    fd.win_hi[a] = W[a] - 1 - s.cen[a];
    fd.val_lo[a] = std::max(wc[a] - half[a], fd.win_lo[a]);
    fd.val_hi[a] = std::min(wc[a] + half[a], fd.win_hi[a]);
    fd.origin[a] = 50 * fd.val_lo[a] - 25;
  }
}

// offsets that depend on the (host-known) store sizes
static void fill_store_tables(s2m_ctx* ctx, int* total_lp) {
  const int B = ctx->d.B, G = ctx->d.G;
  HostTables& T = *ctx->ht;
  int acc = 0;
  for (int g = 0; g < G; ++g) {
    const int n = ctx->slots[g < B ? g : g - B].n_store[g >= B];
    T.lp_off[g] = acc - g; T.so_off[g] = acc; acc += n + 1;  // merge index space: one extra position per segment (inserts behind the last entry)
  }
  T.lp_off[G] = acc - G; T.so_off[G] = acc;
  *total_lp = acc;
}
// Lists the segments of the slots flagged in `rebuild` for a bulk rebuild of their cell index: their local-map
// sizes are in ctx->h_lpcnt (read back).  Returns the number of segments; *points = their local points.
static int plan_index_rebuild(s2m_ctx* ctx, const std::vector<char>& rebuild, int* points) {
  const int B = ctx->d.B, G = ctx->d.G;
  HostTables& T = *ctx->ht;
  int n_seg = 0, acc = 0;
  for (int g = 0; g < G; ++g) {
    const int b = g < B ? g : g - B;
    if (!rebuild[b]) continue;
    const int n = ctx->h_lpcnt[g];
    const long long region = (long long)T.hash_off[g + 1] - T.hash_off[g];
    T.idx_list[n_seg] = g;
    T.idx_list[G + n_seg] = (int)(std::min(table_size(n), region) - 1);
    T.idx_poff[n_seg] = acc;
    acc += n;
    ++n_seg;
    ctx->slots[b].idx_built_n[g >= B] = n;
  }
  T.idx_poff[n_seg] = acc;
  *points = acc;
  return n_seg;
}

static int finish_call(s2m_ctx* ctx) {
  CK(cudaMemcpyAsync(ctx->h_out, ctx->d.out, sizeof(SlotOut) * ctx->d.B, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(ctx->h_err, ctx->d.err_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(wait_stream(ctx));
  CK(cudaGetLastError());
  if (*ctx->h_err != 0) {
    int e = *ctx->h_err;
    cudaMemsetAsync(ctx->d.err_flag, 0, sizeof(int), ctx->stream);
    ctx->err = std::string("device reported: ") + s2m_strerror(e);
    return e;
  }
  if (ctx->guard_bytes) {
    const int bad = s2m_debug_guard_check(ctx);
    if (bad != 0) { ctx->err = "guard band overwritten: " + std::to_string(bad) + " words (an out-of-bounds store)"; return S2M_ERR_INTERNAL; }
  }
  return S2M_OK;
}

// one frame for all active slots; inputs already packed in d.in_pts
static int run_frame(s2m_ctx* ctx, const int* corner_off, const int* surf_off, const double* q_wodom,
                     const double* t_wodom, const int* active, double* q_out, double* t_out, s2m_stats* stats,
                     int* status) {
  Dev& d = ctx->d;
  const int B = d.B, G = d.G;
  HostTables& T = *ctx->ht;
  cudaStream_t s = ctx->stream;
  const int total_in = T.in_off[G];
  bool check_pending = false, window_shift = false;
  int tiles = 0;
  // A frame is atomic: the device map is double buffered and only swapped at the end, the host-side slot
  // state (window centre, valid block, pending-check flag) is restored if anything below fails.
  struct Rollback {
    std::vector<SlotHost>& live;
    std::vector<SlotHost> saved;
    bool armed = true;
    explicit Rollback(std::vector<SlotHost>& l) : live(l), saved(l) {}
    ~Rollback() {
      if (!armed) return;
      live = saved;
      for (SlotHost& sh : live) sh.idx_valid = false;  // (the failed frame may have touched the cell index)
    }
  } rollback(ctx->slots);
  std::vector<char> idx_rebuild(B, 0), idx_follow(B, 0);
  for (int b = 0; b < B; ++b) {
    SlotHost& sh = ctx->slots[b];
    FrameDesc& fd = T.desc[b];
    fd.active = active ? (active[b] != 0) : 1;
    fd.allow_opt = !ctx->P.skip_optimization;
    if (!fd.active) continue;
    row_A(sh, q_wodom + 4 * b, t_wodom + 3 * b);
    std::memcpy(fd.pose, sh.pose, sizeof(fd.pose));
    const int cen_before[3] = {sh.cen[0], sh.cen[1], sh.cen[2]};
    rows_BC(sh, sh.pose + 4, fd);
    fd.seq_base[0] = sh.seq[0]; fd.seq_base[1] = sh.seq[1];
    bool moved = sh.force_pending_check;
    // entries can only leave the store (cubes that left the window) when the window moved; an uploaded or restored
    // store is treated the same way
    window_shift = window_shift || sh.force_pending_check || sh.cen[0] != cen_before[0] || sh.cen[1] != cen_before[1] ||
                   sh.cen[2] != cen_before[2];
    for (int a = 0; a < 3; ++a) moved = moved || fd.val_lo[a] != sh.val_lo[a] || fd.val_hi[a] != sh.val_hi[a];
    check_pending = check_pending || moved;
    // cell index: bulk rebuild when the valid block changed (or the index is stale); otherwise the map update keeps
    // it in step.  A frame that merges raw points of newly valid cubes is followed by one more rebuild.
    idx_rebuild[b] = moved || !sh.idx_valid || ctx->P.shard_world > 1;
    idx_follow[b] = !moved && ctx->P.shard_world <= 1;
    for (int a = 0; a < 3; ++a) { sh.val_lo[a] = fd.val_lo[a]; sh.val_hi[a] = fd.val_hi[a]; }
    sh.force_pending_check = false;
    const int nq = (corner_off[b + 1] - corner_off[b]) + (surf_off[b + 1] - surf_off[b]);
    tiles = std::max(tiles, (nq + kTile - 1) / kTile);
  }
  int total_lp;
  fill_store_tables(ctx, &total_lp);
  const int total_store = total_lp;
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, s));

  long long k = 0;
  prof_mark(ctx, S2M_PHASE_INPUT);
  // Two small results come back before the bulk of the frame is enqueued: the boxes of the incoming
  // clouds (exact width of PCL's voxel index = radix passes of the scan filter) and the sizes of the
  // local maps (25 ranges of the sorted store: the guard :555, the staging of raw points, the sizing of a
  // rebuilt cell index).  The round trip is hidden by the other lanes sharing the GPU.
  int longest_in = 0;
  for (int g = 0; g < G; ++g) longest_in = std::max(longest_in, T.in_off[g + 1] - T.in_off[g]);
  k += launch_voxel_bbox(d, total_in, longest_in, s);
  k += launch_local_ranges(d, ctx->cur, s);
  CK(cudaMemcpyAsync(ctx->h_bbox, d.bbox, sizeof(uint32_t) * 6 * G, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(ctx->h_lpcnt, d.lp_cnt, sizeof(int) * 2 * G, cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(ctx->ev_bbox, s));
  CK(cudaEventSynchronize(ctx->ev_bbox));
  total_lp = 0;  // the staging area of the raw points the merge absorbs: sized by their exact number
  for (int g = 0; g < G; ++g) { T.lp_off[g] = total_lp; total_lp += ctx->h_lpcnt[G + g]; }
  T.lp_off[G] = total_lp;
  for (int b = 0; b < B; ++b) {
    if (!T.desc[b].active) continue;
    const SlotHost& sh = ctx->slots[b];
    for (int c2 = 0; c2 < 2; ++c2)  // the table was sized for the map it was built from
      if (2 * (long long)ctx->h_lpcnt[c2 * B + b] > 3 * (long long)sh.idx_built_n[c2] + 2048) idx_rebuild[b] = 1;
    T.desc[b].idx_flags = (idx_rebuild[b] ? 1 : 0) | (idx_follow[b] ? 2 : 0);
  }
  for (int b = 0; b < B; ++b) if (!T.desc[b].active) { idx_rebuild[b] = 0; T.desc[b].idx_flags = 0; }
  int idx_points = 0;
  const int idx_segs = plan_index_rebuild(ctx, idx_rebuild, &idx_points);
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, s));
  k += launch_index_rebuild(d, ctx->cur, idx_segs, idx_points, s);
  const bool sharded = ctx->P.shard_world > 1;
  if (sharded) { int rs = shard_allreduce(ctx, d.shard_counts, (size_t)G, nccl::kInt32); if (rs != S2M_OK) return rs; }
  k += launch_guard(d, s);
  prof_mark(ctx, S2M_PHASE_INDEX);
  int key_bits = 1;
  for (int g = 0; g < G && total_in > 0; ++g) {
    const int cnt = T.in_off[g + 1] - T.in_off[g];
    if (cnt == 0) continue;
    auto ord2f = [](uint32_t u) { uint32_t b2 = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u; float f; std::memcpy(&f, &b2, 4); return f; };
    const uint32_t* bb = ctx->h_bbox + 6 * g;
    const float inv = d.inv_leaf[g >= B];
    double cells = 1.0;
    long long lin = 1;
    for (int a2 = 0; a2 < 3; ++a2) {
      const float mn = ord2f(bb[a2]), mx = ord2f(bb[3 + a2]);
      cells *= (double)((long long)xfmul(xfsub(mx, mn), inv) + 1);             // PCL's overflow test
      lin *= (long long)((int)floorf(xfmul(mx, inv)) - (int)floorf(xfmul(mn, inv)) + 1);  // div_b product
    }
    const long long span = cells > 2147483647.0 ? (long long)cnt : std::min<long long>(lin, 2147483647LL);
    int bits = 1;
    while ((1ll << bits) < span) ++bits;
    key_bits = std::max(key_bits, bits);
  }
  key_bits = std::min(key_bits, 31);
  k += launch_voxel_filter(d, total_in, key_bits, s);
  // exact down-sampled counts size the association / evaluation grids and the map update; the short
  // wait is hidden by the other contexts sharing the GPU
  CK(cudaMemcpyAsync(ctx->h_dsoff, d.ds_off, sizeof(int) * (G + 1), cudaMemcpyDeviceToHost, s));
  CK(cudaEventRecord(ctx->ev_ds, s));
  prof_mark(ctx, S2M_PHASE_VOXEL);
  CK(cudaEventSynchronize(ctx->ev_ds));
  const int n_ds = ctx->h_dsoff[G];
  tiles = 0;
  int chunks = 0;  // 32-query work units of knn_kernel over all slots
  for (int b = 0; b < B; ++b) {
    const int nq = (ctx->h_dsoff[b + 1] - ctx->h_dsoff[b]) + (ctx->h_dsoff[B + b + 1] - ctx->h_dsoff[B + b]);
    tiles = std::max(tiles, (nq + kTile - 1) / kTile);
    chunks += (nq + 31) / 32;
  }
  // K4a: one resident wave of persistent warps sharing a work ticket; K4b: one block per 128 points of a slot
#if S2M_KNN_GROUP
  chunks = (n_ds + 31) / 32;  // ... of the grouped search: 32 consecutive points of the block-sorted order
  const int knn_resident = S2M_K4G_MINB;
#else
  const int knn_resident = S2M_K4A_MINB;
#endif
  const int knn_blocks = std::max(1, std::min((chunks + kTile / 32 - 1) / (kTile / 32), knn_resident * ctx->sm_count));
  const int fit_blocks = std::max(1, tiles);
  const int eval_blocks = std::max(1, (tiles + kEvalTilesPerBlock - 1) / kEvalTilesPerBlock);
  d.count_cand = ctx->count_candidates ? 1 : 0;
  if (ctx->count_candidates) k += launch_count27(d, n_ds, s);  // the byte count of the roofline: outside the K4 bracket
  for (int outer = 0; outer < 2; ++outer) {  // laserMapping.cpp:563
    CK(cudaMemsetAsync(d.knn_ticket, 0, sizeof(int), s));  // the association's work ticket
    prof_mark(ctx, S2M_PHASE_READBACK);  // host wait for the down-sampled counts (outer 0); the K4 bracket starts here
    if (outer == 0) k += launch_query_sort(d, n_ds, s);  // (the grouped search's query order: part of K4's time)
    k += launch_associate(d, outer, ctx->cur, knn_blocks, fit_blocks, n_ds, ctx->P.trace != 0, s);
    prof_mark(ctx, S2M_PHASE_ASSOCIATE);
    if (!sharded) {
      k += launch_solve(d, outer, true, s);  // Ceres solve, max_num_iterations = 4 (:713-721)
    } else {
      // one 32-double block per slot across the ranks per evaluation, then the LM step on every rank
      k += launch_reduce_units(d, s);
      int rs = shard_allreduce(ctx, d.shard_sums, (size_t)B * kPartial, nccl::kFloat64);
      if (rs != S2M_OK) return rs;
      k += launch_lm_shard(d, outer, 0, s);
      for (int it = 0; it < 4; ++it) {  // options.max_num_iterations = 4 (:716)
        k += launch_evaluate(d, outer, eval_blocks, s);
        rs = shard_allreduce(ctx, d.shard_sums, (size_t)B * kPartial, nccl::kFloat64);
        if (rs != S2M_OK) return rs;
        k += launch_lm_shard(d, outer, 1, s);
      }
    }
    prof_mark(ctx, S2M_PHASE_SOLVE);
    if (ctx->P.trace)
      CK(cudaMemcpyAsync(ctx->lm_trace + (size_t)outer * B, d.lm, sizeof(LmState) * B, cudaMemcpyDeviceToDevice, s));
  }
  k += launch_finish_pose(d, s);
  k += launch_map_update(d, ctx->cur, n_ds, total_lp, total_store, check_pending, window_shift, false, s);
  prof_mark(ctx, S2M_PHASE_UPDATE);
  ctx->launches += k;
  int rc = finish_call(ctx);
  prof_mark(ctx, S2M_PHASE_READBACK);
  if (rc != S2M_OK) return rc;
  rollback.armed = false;
  // swap store buffers
  ctx->cur ^= 1;
  std::swap(d.st_n, d.st_n_new);

  for (int b = 0; b < B; ++b) {
    if (!T.desc[b].active) { if (status) status[b] = S2M_OK; continue; }
    SlotHost& sh = ctx->slots[b];
    const SlotOut& o = ctx->h_out[b];
    std::memcpy(sh.pose, o.pose, sizeof(sh.pose));
    row_U(sh, q_wodom + 4 * b, t_wodom + 3 * b);  // :735
    sh.n_store[0] = o.n_store[0]; sh.n_store[1] = o.n_store[1];
    sh.seq[0] += (unsigned long long)o.n_ds[0]; sh.seq[1] += (unsigned long long)o.n_ds[1];
    sh.frames++;
    sh.idx_valid = idx_follow[b] != 0;  // the map update kept the cell index in step (else: rebuilt next frame)
    for (int i = 0; i < 4; ++i) q_out[4 * b + i] = sh.pose[i];
    for (int i = 0; i < 3; ++i) t_out[3 * b + i] = sh.pose[4 + i];
    if (status) status[b] = o.optimized ? S2M_OK : S2M_MAP_TOO_SMALL;
    if (stats) {
      s2m_stats& st = stats[b];
      st.n_corner_in = corner_off[b + 1] - corner_off[b]; st.n_surf_in = surf_off[b + 1] - surf_off[b];
      st.n_corner_ds = o.n_ds[0]; st.n_surf_ds = o.n_ds[1];
      st.n_map_corner = o.n_local[0]; st.n_map_surf = o.n_local[1];
      for (int i = 0; i < 2; ++i) {
        st.n_edge[i] = o.n_edge[i]; st.n_plane[i] = o.n_plane[i];
        st.lm_iters[i] = o.lm_iters[i]; st.lm_term[i] = o.lm_term[i];
        st.cost_initial[i] = o.cost_initial[i]; st.cost_final[i] = o.cost_final[i];
      }
      st.optimized = o.optimized;
    }
  }
  // algorithmic bytes of the two association launches of this frame:
  // (Nc+Ns)(16 + 27*8) + 16 * map points in the 27 cells of the queries + 48 * accepted correspondences
  if (ctx->profiling) {
    for (int outer = 0; outer < 2; ++outer) {
      double bytes = 0;
      for (int b = 0; b < B; ++b) {
        const SlotOut& o = ctx->h_out[b];
        if (!T.desc[b].active || !o.optimized) continue;
        bytes += (double)(o.n_ds[0] + o.n_ds[1]) * (16.0 + 27.0 * 8.0) + 16.0 * (o.cand[0] + o.cand[1]) +
                 48.0 * (o.n_edge[outer] + o.n_plane[outer]);
      }
      ctx->k4_bytes += bytes;
    }
    for (int b = 0; b < B; ++b)
      if (T.desc[b].active && ctx->h_out[b].optimized) ctx->k4_cand27 += ctx->h_out[b].cand[0] + ctx->h_out[b].cand[1];
  }
  return S2M_OK;
}

static int check_offsets(s2m_ctx* ctx, const int* corner_off, const int* surf_off) {
  const int B = ctx->d.B;
  HostTables& T = *ctx->ht;
  if (corner_off[0] != 0 || surf_off[0] != 0) { ctx->err = "offset arrays must start at 0"; return S2M_ERR_ARG; }
  for (int b = 0; b < B; ++b) {
    const int nc = corner_off[b + 1] - corner_off[b], ns = surf_off[b + 1] - surf_off[b];
    if (nc < 0 || ns < 0) { ctx->err = "offsets must be non-decreasing"; return S2M_ERR_ARG; }
    if (nc > ctx->P.cap_corner_in || ns > ctx->P.cap_surf_in) { ctx->err = "incoming cloud larger than cap_*_in"; return S2M_ERR_CAPACITY; }
  }
  // class-major segments: [corner of slot 0..B-1][surf of slot 0..B-1]
  const int NC = corner_off[B];
  for (int b = 0; b <= B; ++b) T.in_off[b] = corner_off[b];
  for (int b = 0; b <= B; ++b) T.in_off[B + b] = NC + surf_off[b];
  return S2M_OK;
}

// leaf: enqueue the copy of one frame's clouds into the next staging buffer (copy stream)
static int leaf_stage(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf, const int* surf_off,
                      cudaMemcpyKind kind, cudaStream_t cs, cudaEvent_t after, int* stage_out) {
  const int B = ctx->d.B;
  const long long NC = corner_off[B], NS = surf_off[B];
  if (corner_off[0] != 0 || surf_off[0] != 0 || NC < 0 || NS < 0) { ctx->err = "offset arrays must start at 0"; return S2M_ERR_ARG; }
  if (NC + NS > ctx->d.cap_in) { ctx->err = "incoming clouds larger than the context's input capacity"; return S2M_ERR_CAPACITY; }
  const int st = ctx->stage_next;
  ctx->stage_next ^= 1;
  if (after) CK(cudaStreamWaitEvent(cs, after, 0));
  if (NC > 0) CK(cudaMemcpyAsync(ctx->in_buf[st], corner, sizeof(float4) * (size_t)NC, kind, cs));
  if (NS > 0) CK(cudaMemcpyAsync(ctx->in_buf[st] + NC, surf, sizeof(float4) * (size_t)NS, kind, cs));
  CK(cudaEventRecord(ctx->ev_in[st], cs));
  *stage_out = st;
  return S2M_OK;
}
// leaf: run one frame on the clouds staged in buffer `st`
static int leaf_run(s2m_ctx* ctx, int st, const int* corner_off, const int* surf_off, const double* q_wodom,
                    const double* t_wodom, const int* active, double* q_out, double* t_out, s2m_stats* stats, int* status) {
  CK(cudaSetDevice(ctx->P.device));
  int rc = check_offsets(ctx, corner_off, surf_off);
  if (rc != S2M_OK) return rc;
  ctx->d.in_pts = ctx->in_buf[st];
  prof_mark(ctx, -1);
  CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[st], 0));
  return run_frame(ctx, corner_off, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats, status);
}

static int parent_wait(s2m_ctx* ctx) {
  if (ctx->in_flight.empty()) { ctx->err = "no frame in flight"; return S2M_ERR_ARG; }
  std::unique_ptr<PendingFrame> pf = std::move(ctx->in_flight.front());
  ctx->in_flight.pop_front();
  const int L = (int)ctx->children.size();
  int rc_all = S2M_OK;
  for (int i = 0; i < L; ++i) {
    ctx->workers[i]->wait(pf->tickets[i]);
    cudaStreamWaitEvent(ctx->stream, ctx->lane_done[pf->evslot * L + i], 0);
    if (pf->rcs[i] < 0 && rc_all == S2M_OK) { rc_all = pf->rcs[i]; ctx->err = ctx->children[i]->err; }
  }
  return rc_all;
}

// every lane advances its slots on its own stream and host thread; the caller's stream is fenced
// on both sides so events recorded around submit ... wait bracket all the work
static int parent_submit(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf, const int* surf_off,
                         const double* q_wodom, const double* t_wodom, const int* active, double* q_out, double* t_out,
                         s2m_stats* stats, int* status, cudaMemcpyKind kind) {
  if (ctx->in_flight.size() >= 2) { ctx->err = "at most two frames may be in flight: call s2m_register_batch_wait first"; return S2M_ERR_ARG; }
  const int L = (int)ctx->children.size(), Bl = ctx->lane_batch, B = ctx->P.batch;
  std::unique_ptr<PendingFrame> pf(new PendingFrame());
  pf->evslot = ctx->frame_seq++ & 1;
  pf->co.resize(L); pf->so.resize(L); pf->rcs.assign(L, S2M_OK); pf->tickets.assign(L, 0);
  pf->q.assign(q_wodom, q_wodom + 4 * (size_t)B); pf->t.assign(t_wodom, t_wodom + 3 * (size_t)B);
  pf->has_active = active != nullptr;
  if (active) pf->active.assign(active, active + B);
  pf->q_out = q_out; pf->t_out = t_out; pf->stats = stats; pf->status = status;
  CK(cudaEventRecord(ctx->lane_start[pf->evslot], ctx->stream));
  PendingFrame* P = pf.get();
  for (int i = 0; i < L; ++i) {
    s2m_ctx* ch = ctx->children[i];
    const int b0 = i * Bl, nb = ch->P.batch;
    P->co[i].resize(nb + 1); P->so[i].resize(nb + 1);
    for (int b = 0; b <= nb; ++b) { P->co[i][b] = corner_off[b0 + b] - corner_off[b0]; P->so[i][b] = surf_off[b0 + b] - surf_off[b0]; }
    int st = 0;
    int rc = leaf_stage(ch, corner ? corner + 4 * (size_t)corner_off[b0] : nullptr, P->co[i].data(),
                        surf ? surf + 4 * (size_t)surf_off[b0] : nullptr, P->so[i].data(), kind, ch->copy_stream,
                        ctx->lane_start[P->evslot], &st);
    if (rc != S2M_OK) P->rcs[i] = rc;
    cudaEvent_t done = ctx->lane_done[P->evslot * L + i];
    P->tickets[i] = ctx->workers[i]->submit([=]() {
      if (P->rcs[i] == S2M_OK)
        P->rcs[i] = leaf_run(ch, st, P->co[i].data(), P->so[i].data(), P->q.data() + 4 * b0, P->t.data() + 3 * b0,
                             P->has_active ? P->active.data() + b0 : nullptr, P->q_out + 4 * b0, P->t_out + 3 * b0,
                             P->stats ? P->stats + b0 : nullptr, P->status ? P->status + b0 : nullptr);
      cudaEventRecord(done, ch->stream);
    });
  }
  ctx->in_flight.push_back(std::move(pf));
  return S2M_OK;
}

static int register_batch_impl(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf,
                               const int* surf_off, const double* q_wodom, const double* t_wodom, const int* active,
                               double* q_out, double* t_out, s2m_stats* stats, int* status, cudaMemcpyKind kind) {
  if (!ctx || !corner_off || !surf_off || !q_wodom || !t_wodom || !q_out || !t_out) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  if (!ctx->children.empty()) {
    if (!ctx->in_flight.empty()) { ctx->err = "frames still in flight: wait for them before a synchronous call"; return S2M_ERR_ARG; }
    int rc = parent_submit(ctx, corner, corner_off, surf, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats, status, kind);
    if (rc != S2M_OK) return rc;
    return parent_wait(ctx);
  }
  int st = 0;
  int rc = leaf_stage(ctx, corner, corner_off, surf, surf_off, kind, ctx->stream, nullptr, &st);
  if (rc != S2M_OK) return rc;
  return leaf_run(ctx, st, corner_off, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats, status);
}

extern "C" int s2m_register_batch_submit(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf,
                                         const int* surf_off, const double* q_wodom, const double* t_wodom,
                                         const int* active, double* q_out, double* t_out, s2m_stats* stats, int* status,
                                         int device_ptrs) {
  if (!ctx || !corner_off || !surf_off || !q_wodom || !t_wodom || !q_out || !t_out) return S2M_ERR_ARG;
  if (ctx->children.empty()) { ctx->err = "the asynchronous pair needs a multi-lane context (s2m_params.lanes >= 1)"; return S2M_ERR_ARG; }
  CK(cudaSetDevice(ctx->P.device));
  return parent_submit(ctx, corner, corner_off, surf, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats, status,
                       device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice);
}
extern "C" int s2m_register_batch_wait(s2m_ctx* ctx) {
  if (!ctx || ctx->children.empty()) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  return parent_wait(ctx);
}

extern "C" int s2m_register_batch(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf,
                                  const int* surf_off, const double* q_wodom, const double* t_wodom,
                                  const int* active, double* q_out, double* t_out, s2m_stats* stats, int* status) {
  return register_batch_impl(ctx, corner, corner_off, surf, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats,
                             status, cudaMemcpyHostToDevice);
}
extern "C" int s2m_register_batch_dev(s2m_ctx* ctx, const float* corner, const int* corner_off, const float* surf,
                                      const int* surf_off, const double* q_wodom, const double* t_wodom,
                                      const int* active, double* q_out, double* t_out, s2m_stats* stats,
                                      int* status) {
  return register_batch_impl(ctx, corner, corner_off, surf, surf_off, q_wodom, t_wodom, active, q_out, t_out, stats,
                             status, cudaMemcpyDeviceToDevice);
}

extern "C" int s2m_register(s2m_ctx* ctx, const float* corner, int nc, const float* surf, int ns,
                            const double q_wodom[4], const double t_wodom[3], double q_out[4], double t_out[3],
                            s2m_stats* stats) {
  if (!ctx || nc < 0 || ns < 0) return S2M_ERR_ARG;
  const int B = ctx->P.batch;  // (a multi-lane context goes through the batch call like any other: lane worker, in-flight check)
  std::vector<int> co(B + 1, nc), so(B + 1, ns), act(B, 0);
  co[0] = so[0] = 0;
  act[0] = 1;
  std::vector<double> q(4 * B, 0.0), t(3 * B, 0.0), qo(4 * B), to(3 * B);
  std::vector<s2m_stats> st(B);
  std::vector<int> status(B, 0);
  std::memcpy(q.data(), q_wodom, 32); std::memcpy(t.data(), t_wodom, 24);
  int rc = s2m_register_batch(ctx, corner, co.data(), surf, so.data(), q.data(), t.data(), act.data(), qo.data(),
                              to.data(), st.data(), status.data());
  if (rc != S2M_OK) return rc;
  std::memcpy(q_out, qo.data(), 32); std::memcpy(t_out, to.data(), 24);
  if (stats) *stats = st[0];
  return status[0];
}

extern "C" int s2m_get_correction(s2m_ctx* ctx, int slot, double q[4], double t[3]) {
  ROUTE_SLOT(s2m_get_correction(ch, slot, q, t));
  if (!ctx || slot < 0 || slot >= ctx->d.B) return S2M_ERR_ARG;
  std::memcpy(q, ctx->slots[slot].q_wmap_wodom, 32);
  std::memcpy(t, ctx->slots[slot].t_wmap_wodom, 24);
  return S2M_OK;
}
extern "C" int s2m_get_window(s2m_ctx* ctx, int slot, int cen[3]) {
  ROUTE_SLOT(s2m_get_window(ch, slot, cen));
  if (!ctx || slot < 0 || slot >= ctx->d.B) return S2M_ERR_ARG;
  std::memcpy(cen, ctx->slots[slot].cen, 12);
  return S2M_OK;
}

extern "C" int s2m_transform_cloud(s2m_ctx* ctx, int slot, const float* in, int n, float* out) {
  ROUTE_SLOT(s2m_transform_cloud(ch, slot, in, n, out));
  if (!ctx || slot < 0 || slot >= ctx->d.B || n < 0) return S2M_ERR_ARG;
  if (n > ctx->d.cap_sort) return S2M_ERR_CAPACITY;
  if (n == 0) return S2M_OK;
  CK(cudaSetDevice(ctx->P.device));
  // scratch: ins_pt / ins_cpt (float4, cap_sort+1) and the first 7 doubles of rec
  CK(cudaMemcpyAsync(ctx->d.ins_pt, in, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->d.partials, ctx->slots[slot].pose, 56, cudaMemcpyHostToDevice, ctx->stream));
  ctx->launches += launch_transform_cloud(ctx->d.partials, ctx->d.ins_pt, ctx->d.ins_cpt, n, ctx->stream);
  CK(cudaMemcpyAsync(out, ctx->d.ins_cpt, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return S2M_OK;
}

// ---- map access ---------------------------------------------------------------
extern "C" int s2m_map_upload(s2m_ctx* ctx, int slot, const float* corner, int nc, const float* surf, int ns) {
  ROUTE_SLOT(s2m_map_upload(ch, slot, corner, nc, surf, ns));
  if (!ctx || slot < 0 || slot >= ctx->d.B || nc < 0 || ns < 0) return S2M_ERR_ARG;
  if (nc > ctx->P.cap_map_corner || ns > ctx->P.cap_map_surf) return S2M_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->P.device));
  Dev& d = ctx->d;
  const int B = d.B, G = d.G;
  HostTables& T = *ctx->ht;
  cudaStream_t s = ctx->stream;
  SlotHost& sh = ctx->slots[slot];
  // forget the slot's current entries
  int zero = 0;
  CK(cudaMemcpyAsync(d.st_n + slot, &zero, 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.st_n + B + slot, &zero, 4, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));
  sh.n_store[0] = sh.n_store[1] = 0;
  sh.seq[0] = sh.seq[1] = 0;
  // The map goes through the frame staging buffers (sized by cap_*_in, far smaller than the map), so it is
  // pushed in chunks; the arrival numbers continue from chunk to chunk, which keeps the upload order.
  const int step_c = ctx->P.cap_corner_in, step_s = ctx->P.cap_surf_in;
  int kept = 0;
  for (int c0 = 0, s0 = 0; c0 < nc || s0 < ns || (nc + ns == 0 && c0 == 0 && s0 == 0);) {
    const int cn = std::min(step_c, nc - c0), sn = std::min(step_s, ns - s0);
    std::vector<int> dsoff(G + 1, 0);
    for (int g = 0; g <= G; ++g) dsoff[g] = (g > slot ? cn : 0) + (g > B + slot ? sn : 0);
    for (int b = 0; b < B; ++b) {
      FrameDesc& fd = T.desc[b];
      fd.active = (b == slot);
      fd.allow_opt = 0;
      fd.idx_flags = 0;
      if (b != slot) continue;
      const int W[3] = {kWinI, kWinJ, kWinK};
      for (int a = 0; a < 3; ++a) {
        fd.win_lo[a] = -sh.cen[a]; fd.win_hi[a] = W[a] - 1 - sh.cen[a];
        fd.val_lo[a] = 1; fd.val_hi[a] = 0;  // nothing is valid: every point stays raw
        fd.origin[a] = 0;
      }
      fd.seq_base[0] = sh.seq[0]; fd.seq_base[1] = sh.seq[1];
    }
    int total_lp;
    fill_store_tables(ctx, &total_lp);
    CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d.ds_off, dsoff.data(), sizeof(int) * (G + 1), cudaMemcpyHostToDevice, s));
    if (cn) CK(cudaMemcpyAsync(d.ds_pts, corner + 4 * (size_t)c0, sizeof(float4) * (size_t)cn, cudaMemcpyHostToDevice, s));
    if (sn) CK(cudaMemcpyAsync(d.ds_pts + cn, surf + 4 * (size_t)s0, sizeof(float4) * (size_t)sn, cudaMemcpyHostToDevice, s));
    ctx->launches += launch_map_update(d, ctx->cur, cn + sn, total_lp, total_lp, false, false, true, s);
    int rc = finish_call(ctx);
    if (rc != S2M_OK) return rc;
    ctx->cur ^= 1;
    std::swap(d.st_n, d.st_n_new);
    const SlotOut& o = ctx->h_out[slot];
    sh.n_store[0] = o.n_store[0]; sh.n_store[1] = o.n_store[1];
    sh.seq[0] += (unsigned long long)cn; sh.seq[1] += (unsigned long long)sn;
    c0 += cn; s0 += sn;
    kept = o.n_store[0] + o.n_store[1];
    if (nc + ns == 0) break;
  }
  sh.force_pending_check = true;
  sh.idx_valid = false;
  sh.val_lo[0] = 1; sh.val_hi[0] = 0;
  return (nc + ns) - kept;
}

static bool host_entry_dead(const SlotHost& sh, uint64_t key) {
  int ci, cj, ck;
  unpack_cube(key_cube(key), ci, cj, ck);
  const int W[3] = {kWinI, kWinJ, kWinK};
  const int c[3] = {ci, cj, ck};
  for (int a = 0; a < 3; ++a)
    if (c[a] < -sh.cen[a] || c[a] > W[a] - 1 - sh.cen[a]) return true;
  return false;
}

static int download_store(s2m_ctx* ctx, int slot, int cls, std::vector<uint64_t>& keys, std::vector<float>& pts) {
  Dev& d = ctx->d;
  const int g = cls * d.B + slot;
  const int n = ctx->slots[slot].n_store[cls];
  keys.resize(n); pts.resize((size_t)4 * n);
  if (n == 0) return S2M_OK;
  int base = 0;
  for (int h = 0; h < g; ++h) base += (h < d.B ? ctx->P.cap_map_corner : ctx->P.cap_map_surf);
  CK(cudaMemcpyAsync(keys.data(), d.st_key[ctx->cur] + base, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(pts.data(), d.st_pt[ctx->cur] + base, sizeof(float4) * n, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return S2M_OK;
}

extern "C" int s2m_map_download(s2m_ctx* ctx, int slot, int cls, float* out, int cap) {
  ROUTE_SLOT(s2m_map_download(ch, slot, cls, out, cap));
  if (!ctx || slot < 0 || slot >= ctx->d.B || cls < 0 || cls > 1) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  std::vector<uint64_t> keys;
  std::vector<float> pts;
  int rc = download_store(ctx, slot, cls, keys, pts);
  if (rc != S2M_OK) return rc;
  int n = 0;
  for (size_t i = 0; i < keys.size(); ++i) {
    if (host_entry_dead(ctx->slots[slot], keys[i])) continue;  // evicted lazily at the next update
    if (out && n < cap) std::memcpy(out + 4 * (size_t)n, &pts[4 * i], 16);
    ++n;
  }
  return n;
}

// ---- PCD v0.7 binary, FIELDS x y z intensity (float32): the layout pcl::io::savePCDFileBinary
// writes for PointXYZI (laserPosegraphOptimization.cpp:695) and the reference ships under
// utils/sample_data/*/Scans. Host-only helpers; no device is touched.
extern "C" int s2m_pcd_write(const char* path, const float* xyzi, int n) {
  if (!path || n < 0 || (n > 0 && !xyzi)) return S2M_ERR_ARG;
  FILE* f = std::fopen(path, "wb");
  if (!f) return S2M_ERR_IO;
  std::fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\n"
                  "TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n", n, n);
  const size_t w = n ? std::fwrite(xyzi, 16, (size_t)n, f) : 0;
  const int bad = std::fclose(f);
  return (w == (size_t)n && !bad) ? S2M_OK : S2M_ERR_IO;
}
// returns the number of points in the file (copies min(n, cap) of them), or an error
extern "C" int s2m_pcd_read(const char* path, float* xyzi, int cap) {
  if (!path || cap < 0) return S2M_ERR_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return S2M_ERR_IO;
  char line[256];
  long long points = -1;
  bool fields_ok = false, sizes_ok = false, types_ok = false, binary = false;
  while (std::fgets(line, sizeof line, f)) {
    if (!std::strncmp(line, "FIELDS", 6)) fields_ok = !std::strncmp(line, "FIELDS x y z intensity", 22);
    else if (!std::strncmp(line, "SIZE", 4)) sizes_ok = !std::strncmp(line, "SIZE 4 4 4 4", 12);
    else if (!std::strncmp(line, "TYPE", 4)) types_ok = !std::strncmp(line, "TYPE F F F F", 12);
    else if (!std::strncmp(line, "POINTS", 6)) points = std::atoll(line + 6);
    else if (!std::strncmp(line, "DATA", 4)) { binary = !std::strncmp(line, "DATA binary", 11) && line[11] != '_'; break; }
  }
  int rc;
  if (!fields_ok || !sizes_ok || !types_ok || !binary || points < 0 || points > 0x7fffffffLL) rc = S2M_ERR_IO;
  else {
    const size_t want = (size_t)std::min<long long>(points, cap);
    rc = (want == 0 || (xyzi && std::fread(xyzi, 16, want, f) == want)) ? (int)points : S2M_ERR_IO;
  }
  std::fclose(f);
  return rc;
}

// Checkpoint of one slot: <prefix>.corner.pcd, <prefix>.surf.pcd (the map, rows I/W state) and
// <prefix>.state (window centre :74-76, wmap<-wodom correction :116-117, as hex floats).
extern "C" int s2m_checkpoint_save(s2m_ctx* ctx, int slot, const char* prefix) {
  if (!ctx || !prefix) return S2M_ERR_ARG;
  const std::string base(prefix);
  for (int cls = 0; cls < 2; ++cls) {
    const int n = s2m_map_download(ctx, slot, cls, nullptr, 0);
    if (n < 0) return n;
    std::vector<float> pts(4 * (size_t)std::max(n, 1));
    int rc = s2m_map_download(ctx, slot, cls, pts.data(), n);
    if (rc < 0) return rc;
    rc = s2m_pcd_write((base + (cls ? ".surf.pcd" : ".corner.pcd")).c_str(), pts.data(), n);
    if (rc != S2M_OK) { ctx->err = "cannot write the checkpoint PCD"; return rc; }
  }
  int cen[3];
  double q[4], t[3];
  int rc = s2m_get_window(ctx, slot, cen);
  if (rc == S2M_OK) rc = s2m_get_correction(ctx, slot, q, t);
  if (rc != S2M_OK) return rc;
  FILE* f = std::fopen((base + ".state").c_str(), "w");
  if (!f) { ctx->err = "cannot write the checkpoint state"; return S2M_ERR_IO; }
  std::fprintf(f, "s2m-checkpoint 1\ncen %d %d %d\nq_wmap_wodom %a %a %a %a\nt_wmap_wodom %a %a %a\n", cen[0], cen[1], cen[2],
               q[0], q[1], q[2], q[3], t[0], t[1], t[2]);
  return std::fclose(f) ? S2M_ERR_IO : S2M_OK;
}
static int restore_state(s2m_ctx* ctx, int slot, const int cen[3], const double q[4], const double t[3]) {
  ROUTE_SLOT(restore_state(ch, slot, cen, q, t));
  if (!ctx || slot < 0 || slot >= ctx->d.B) return S2M_ERR_ARG;
  SlotHost& sh = ctx->slots[slot];
  std::memcpy(sh.cen, cen, 12);
  std::memcpy(sh.q_wmap_wodom, q, 32);
  std::memcpy(sh.t_wmap_wodom, t, 24);
  return S2M_OK;
}
// returns the number of points that fell outside the restored window (0 for a checkpoint of this library)
extern "C" int s2m_checkpoint_load(s2m_ctx* ctx, int slot, const char* prefix) {
  if (!ctx || !prefix) return S2M_ERR_ARG;
  const std::string base(prefix);
  FILE* f = std::fopen((base + ".state").c_str(), "r");
  if (!f) { ctx->err = "cannot read the checkpoint state"; return S2M_ERR_IO; }
  int ver = 0, cen[3];
  double q[4], t[3];
  const int got = std::fscanf(f, "s2m-checkpoint %d cen %d %d %d q_wmap_wodom %la %la %la %la t_wmap_wodom %la %la %la", &ver, &cen[0],
                              &cen[1], &cen[2], &q[0], &q[1], &q[2], &q[3], &t[0], &t[1], &t[2]);
  std::fclose(f);
  if (got != 11 || ver != 1) { ctx->err = "malformed checkpoint state"; return S2M_ERR_IO; }
  std::vector<float> pts[2];
  int n[2];
  for (int cls = 0; cls < 2; ++cls) {
    const std::string path = base + (cls ? ".surf.pcd" : ".corner.pcd");
    n[cls] = s2m_pcd_read(path.c_str(), nullptr, 0);
    if (n[cls] < 0) { ctx->err = "cannot read the checkpoint PCD"; return n[cls]; }
    pts[cls].resize(4 * (size_t)std::max(n[cls], 1));
    if (s2m_pcd_read(path.c_str(), pts[cls].data(), n[cls]) != n[cls]) { ctx->err = "cannot read the checkpoint PCD"; return S2M_ERR_IO; }
  }
  // everything that can be checked is checked before the slot is touched
  if (slot < 0 || slot >= ctx->P.batch) return S2M_ERR_ARG;
  if (n[0] > ctx->P.cap_map_corner || n[1] > ctx->P.cap_map_surf) { ctx->err = "checkpoint larger than cap_map_*"; return S2M_ERR_CAPACITY; }
  int rc = restore_state(ctx, slot, cen, q, t);
  if (rc != S2M_OK) return rc;
  return s2m_map_upload(ctx, slot, pts[0].data(), n[0], pts[1].data(), n[1]);
}

// prepares the descriptor of `slot` for a sensor at centre_t (rows B, C), uploads the tables and computes the ranges of
// the local map; with_index: also (re)builds the slot's cell index for that valid block (the live index of the slot
// is stale afterwards and is rebuilt by the next frame)
static int prepare_local(s2m_ctx* ctx, int slot, const double centre_t[3], bool with_index) {
  HostTables& T = *ctx->ht;
  const int B = ctx->d.B, G = ctx->d.G;
  for (int b = 0; b < B; ++b) { T.desc[b].active = (b == slot); T.desc[b].idx_flags = 0; }
  FrameDesc& fd = T.desc[slot];
  SlotHost probe = ctx->slots[slot];  // a getter must not move the live window
  rows_BC(probe, centre_t, fd);
  int total_lp;
  fill_store_tables(ctx, &total_lp);
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, ctx->stream));
  ctx->launches += launch_local_ranges(ctx->d, ctx->cur, ctx->stream);
  if (!with_index) return S2M_OK;
  CK(cudaMemcpyAsync(ctx->h_lpcnt, ctx->d.lp_cnt, sizeof(int) * G, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::vector<char> rebuild(B, 0);
  rebuild[slot] = 1;
  int points = 0;
  const int saved_n[2] = {ctx->slots[slot].idx_built_n[0], ctx->slots[slot].idx_built_n[1]};
  const int n_seg = plan_index_rebuild(ctx, rebuild, &points);
  ctx->slots[slot].idx_built_n[0] = saved_n[0]; ctx->slots[slot].idx_built_n[1] = saved_n[1];
  ctx->slots[slot].idx_valid = false;
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, ctx->stream));
  ctx->launches += launch_index_rebuild(ctx->d, ctx->cur, n_seg, points, ctx->stream);
  return S2M_OK;
}

extern "C" int s2m_get_local_map(s2m_ctx* ctx, int slot, int cls, const double centre_t[3], float* out, int cap) {
  ROUTE_SLOT(s2m_get_local_map(ch, slot, cls, centre_t, out, cap));
  if (!ctx || slot < 0 || slot >= ctx->d.B || cls < 0 || cls > 1 || !centre_t) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  int rc = prepare_local(ctx, slot, centre_t, false);
  if (rc != S2M_OK) return rc;
  const int g = cls * ctx->d.B + slot;
  ctx->launches += launch_gather_local(ctx->d, ctx->cur, g, ctx->d.ins_pt, ctx->stream);
  rc = finish_call(ctx);
  if (rc != S2M_OK) return rc;
  const int n = ctx->h_out[slot].n_local[cls];
  if (out && n > 0) {
    CK(cudaMemcpyAsync(out, ctx->d.ins_pt, sizeof(float4) * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

extern "C" int s2m_debug_knn(s2m_ctx* ctx, int slot, int cls, const double centre_t[3], const float* q_xyz, int n,
                             int32_t* idx5, float* d2_5) {
  ROUTE_SLOT(s2m_debug_knn(ch, slot, cls, centre_t, q_xyz, n, idx5, d2_5));
  if (!ctx || slot < 0 || slot >= ctx->d.B || cls < 0 || cls > 1 || !centre_t || n < 0) return S2M_ERR_ARG;
  if ((size_t)n * 5 > (size_t)ctx->d.cap_sort) return S2M_ERR_CAPACITY;
  CK(cudaSetDevice(ctx->P.device));
  int rc = prepare_local(ctx, slot, centre_t, true);
  if (rc != S2M_OK) return rc;
  Dev& d = ctx->d;
  cudaStream_t s = ctx->stream;
  // scratch: queries in dl_pt/ins_pt area, results in vval (int32) and flag (float bits)
  float* dq = (float*)d.ins_pt;
  int32_t* didx = (int32_t*)d.vval;
  float* dd2 = (float*)d.flag;
  if (n) CK(cudaMemcpyAsync(dq, q_xyz, sizeof(float) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
  ctx->launches += launch_knn_debug(d, ctx->cur, slot, cls, dq, n, didx, dd2, s);
  rc = finish_call(ctx);
  if (rc != S2M_OK) return rc;
  if (n) {
    CK(cudaMemcpyAsync(idx5, didx, sizeof(int32_t) * 5 * (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(d2_5, dd2, sizeof(float) * 5 * (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  return ctx->h_out[slot].n_local[cls];
}

extern "C" int s2m_get_surround(s2m_ctx* ctx, int slot, float* out, int cap) {
  ROUTE_SLOT(s2m_get_surround(ch, slot, out, cap));
  if (!ctx || slot < 0 || slot >= ctx->d.B || cap < 0) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  // corner then surf of each valid cube, cubes in gather order (laserMapping.cpp:810-815), gathered
  // on the device from the valid block of the last registration; only the result crosses PCIe
  Dev& d = ctx->d;
  const SlotHost& sh = ctx->slots[slot];
  HostTables& T = *ctx->ht;
  FrameDesc& fd = T.desc[slot];
  for (int a2 = 0; a2 < 3; ++a2) { fd.val_lo[a2] = sh.val_lo[a2]; fd.val_hi[a2] = sh.val_hi[a2]; }
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, ctx->stream));
  const int room = std::min(cap, d.cap_sort);
  ctx->launches += launch_surround(d, ctx->cur, slot, d.ins_pt, room, d.ticket, ctx->stream);
  int n = 0;
  CK(cudaMemcpyAsync(&n, d.ticket, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaMemsetAsync(d.ticket, 0, sizeof(int), ctx->stream));
  if (out && n > 0) {
    CK(cudaMemcpyAsync(out, d.ins_pt, sizeof(float4) * (size_t)std::min(n, room), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}

// ---- trace ----------------------------------------------------------------------
static int seg_ds_range(s2m_ctx* ctx, int slot, int cls, int* off, int* n) {
  std::vector<int> dsoff(ctx->d.G + 1);
  CK(cudaMemcpyAsync(dsoff.data(), ctx->d.ds_off, sizeof(int) * (ctx->d.G + 1), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const int g = cls * ctx->d.B + slot;
  *off = dsoff[g]; *n = dsoff[g + 1] - dsoff[g];
  return S2M_OK;
}
extern "C" int s2m_trace_cloud(s2m_ctx* ctx, int slot, int cls, float* out, int cap) {
  ROUTE_SLOT(s2m_trace_cloud(ch, slot, cls, out, cap));
  if (!ctx || slot < 0 || slot >= ctx->d.B || cls < 0 || cls > 1) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  int off, n;
  int rc = seg_ds_range(ctx, slot, cls, &off, &n);
  if (rc != S2M_OK) return rc;
  if (out && n > 0) {
    CK(cudaMemcpyAsync(out, ctx->d.ds_pts + off, sizeof(float4) * (size_t)std::min(n, cap), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}
extern "C" int s2m_trace_knn(s2m_ctx* ctx, int slot, int outer, int cls, int32_t* idx5, float* d2_5, uint8_t* used,
                             int cap) {
  ROUTE_SLOT(s2m_trace_knn(ch, slot, outer, cls, idx5, d2_5, used, cap));
  if (!ctx || !ctx->P.trace || slot < 0 || slot >= ctx->d.B || cls < 0 || cls > 1 || outer < 0 || outer > 1) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  int off, n;
  int rc = seg_ds_range(ctx, slot, cls, &off, &n);
  if (rc != S2M_OK) return rc;
  const int m = std::min(n, cap);
  if (m > 0) {
    const size_t o = (size_t)outer * ctx->d.cap_in + off;
    CK(cudaMemcpyAsync(idx5, ctx->d.tr_idx + 5 * o, sizeof(int32_t) * 5 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(d2_5, ctx->d.tr_d2 + 5 * o, sizeof(float) * 5 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(used, ctx->d.tr_used + o, (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return n;
}
extern "C" int s2m_trace_lm(s2m_ctx* ctx, int slot, int outer, double pose7[7], double sums28[28], double iters24[24],
                            int* n_iter, int* termination) {
  ROUTE_SLOT(s2m_trace_lm(ch, slot, outer, pose7, sums28, iters24, n_iter, termination));
  if (!ctx || !ctx->P.trace || slot < 0 || slot >= ctx->d.B || outer < 0 || outer > 1) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  LmState* h = ctx->h_lm;
  CK(cudaMemcpyAsync(h, ctx->lm_trace + (size_t)outer * ctx->d.B + slot, sizeof(LmState), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  std::memcpy(pose7, h->x, 56);
  std::memcpy(sums28, h->init_sums, 28 * 8);
  std::memcpy(iters24, h->it_log, 24 * 8);
  *n_iter = h->iteration;
  *termination = h->termination;
  return S2M_OK;
}

// ---- scan-to-scan odometry (laserOdometry.cpp:220-591) ----------------------
extern "C" int s2m_odom_create(int device, int batch, int cap_sharp, int cap_flat, int cap_less_sharp, int cap_less_flat,
                               int trace, s2m_ctx** out) {
  if (!out || batch < 1 || batch > kMaxBatch || cap_sharp < 1 || cap_flat < 1 || cap_less_sharp < 1 || cap_less_flat < 1 ||
      cap_less_sharp >= (1 << 24) || cap_less_flat >= (1 << 24) || cap_sharp + cap_flat >= (1 << 20))
    return S2M_ERR_ARG;
  s2m_params P;
  s2m_default_params(&P);
  P.device = device; P.batch = batch; P.trace = trace;
  P.cap_corner_in = cap_sharp; P.cap_surf_in = cap_flat;
  P.cap_map_corner = 1024; P.cap_map_surf = 1024;  // the map store of a mapping context is not used here
  s2m_ctx* ctx = nullptr;
  int rc = s2m_create(&P, &ctx);
  if (rc != S2M_OK) return rc;
  ctx->is_odom = true;
  ctx->odom.resize(batch);
  ctx->od_cap = (long long)batch * ((long long)cap_less_sharp + cap_less_flat);
  const size_t oc = (size_t)ctx->od_cap;
  bool bad = dev_alloc(ctx, &ctx->d.od_sorted, oc) || dev_alloc(ctx, &ctx->d.od_key, oc) || dev_alloc(ctx, &ctx->d.od_key2, oc) ||
             dev_alloc(ctx, &ctx->d.od_val, oc) || dev_alloc(ctx, &ctx->d.od_val2, oc) || dev_alloc(ctx, &ctx->d.od_ckey, oc) ||
             dev_alloc(ctx, &ctx->d.od_corr, (size_t)ctx->d.cap_in) || dev_alloc(ctx, &ctx->d.od_bestd, (size_t)ctx->d.cap_in) ||
             dev_alloc(ctx, &ctx->d.od_fb_list, (size_t)ctx->d.cap_in) || dev_alloc(ctx, &ctx->d.od_fb_cnt, 1) ||
             dev_alloc(ctx, &ctx->d.od_first_ge, (size_t)2 * batch * 257) || dev_alloc(ctx, &ctx->d.od_last_le, (size_t)2 * batch * 257);
  if (!bad) {
    ctx->od_tmp_bytes = odom_sort_temp_bytes(ctx->d, (int)oc);
    char* tmp = nullptr;
    bad = dev_alloc(ctx, &tmp, ctx->od_tmp_bytes) != 0;
    ctx->od_tmp = tmp;
  }
  bad = bad || dev_alloc(ctx, &ctx->d.od_last, oc) || dev_alloc(ctx, &ctx->d.od_last_off, 2 * batch + 1) ||
                   dev_alloc(ctx, &ctx->d.od_meta, 2 * (oc / 32 + 2 * (size_t)batch + 1)) ||
                   dev_alloc(ctx, &ctx->d.od_chunk_off, 2 * batch + 1) ||
                   cudaMemset(ctx->d.od_chunk_off, 0, sizeof(int) * (2 * batch + 1)) != cudaSuccess;
  if (bad || cudaMemset(ctx->d.od_last_off, 0, sizeof(int) * (2 * batch + 1)) != cudaSuccess) {
    s2m_destroy(ctx);
    return S2M_ERR_CUDA;
  }
  *out = ctx;
  return S2M_OK;
}

// One sweep of every slot: sharp / flat are the queries, less_sharp / less_flat become the next call's
// targets (:556-566).  Outputs per slot: q_w_curr, t_w_curr (:504-505), optionally para (q_last_curr,
// t_last_curr) and counts[4] = corner correspondences of pass 0, 1, plane correspondences of pass 0, 1.
extern "C" int s2m_odom_step_batch(s2m_ctx* ctx, const float* sharp, const int* sharp_off, const float* flat,
                                   const int* flat_off, const float* less_sharp, const int* ls_off,
                                   const float* less_flat, const int* lf_off, int device_ptrs, double* q_w_out,
                                   double* t_w_out, double* para_out, int* counts_out) {
  if (!ctx || !ctx->is_odom || !sharp_off || !flat_off || !ls_off || !lf_off || !q_w_out || !t_w_out) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  Dev& d = ctx->d;
  const int B = d.B, G = d.G;
  HostTables& T = *ctx->ht;
  cudaStream_t s = ctx->stream;
  const cudaMemcpyKind kind = device_ptrs ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  int rc = check_offsets(ctx, sharp_off, flat_off);
  if (rc != S2M_OK) return rc;
  if (ls_off[0] != 0 || lf_off[0] != 0 || (long long)ls_off[B] + lf_off[B] > ctx->od_cap) {
    ctx->err = "less-sharp / less-flat clouds exceed the context capacity";
    return S2M_ERR_CAPACITY;
  }
  const int NC = sharp_off[B], NS = flat_off[B];
  int tiles = 0;
  bool any = false;
  for (int b = 0; b < B; ++b) {
    FrameDesc& fd = T.desc[b];
    fd.active = ctx->odom[b].inited ? 1 : 0;  // the first sweep only initialises (:267-271)
    fd.allow_opt = 1;
    std::memcpy(fd.pose, ctx->odom[b].para, sizeof(fd.pose));
    any = any || fd.active;
    const int nq = (sharp_off[b + 1] - sharp_off[b]) + (flat_off[b + 1] - flat_off[b]);
    tiles = std::max(tiles, (nq + kTile - 1) / kTile);
  }
  long long k = 0;
  CK(cudaMemcpyAsync(ctx->d_ht, ctx->ht, sizeof(HostTables), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.ds_off, T.in_off, sizeof(int) * (G + 1), cudaMemcpyHostToDevice, s));
  if (NC > 0) CK(cudaMemcpyAsync(d.ds_pts, sharp, sizeof(float4) * (size_t)NC, kind, s));
  if (NS > 0) CK(cudaMemcpyAsync(d.ds_pts + NC, flat, sizeof(float4) * (size_t)NS, kind, s));
  k += launch_odom_guard(d, s);
  if (any) {
    const int eval_blocks = std::max(1, (tiles + kEvalTilesPerBlock - 1) / kEvalTilesPerBlock);
    for (int outer = 0; outer < 2; ++outer) {  // :277
      k += launch_odom_associate(d, outer, std::max(tiles, 1), S2M_OD_FB * ctx->sm_count, ctx->P.trace != 0, s);  // latency-bound warps: many in flight
      for (int it = 0; it < 4; ++it) k += launch_evaluate(d, outer, eval_blocks, s);  // max_num_iterations = 4 (:497)
    }
  }
  k += launch_finish_pose(d, s);
  ctx->launches += k;
  rc = finish_call(ctx);
  if (rc != S2M_OK) return rc;
  for (int b = 0; b < B; ++b) {
    s2m_ctx::OdomHost& o = ctx->odom[b];
    const SlotOut& so = ctx->h_out[b];
    if (o.inited) {
      std::memcpy(o.para, so.pose, sizeof(o.para));
      double r[3], qn[4];
      quat_rotate_exact(o.q_w, o.para[4], o.para[5], o.para[6], r);  // t_w_curr = t_w_curr + q_w_curr * t_last_curr
      for (int i = 0; i < 3; ++i) o.t_w[i] = xdadd(o.t_w[i], r[i]);
      quat_mul_exact(o.q_w, o.para, qn);                              // q_w_curr = q_w_curr * q_last_curr
      std::memcpy(o.q_w, qn, sizeof(qn));
    }
    o.inited = true;
    std::memcpy(q_w_out + 4 * b, o.q_w, 32);
    std::memcpy(t_w_out + 3 * b, o.t_w, 24);
    if (para_out) std::memcpy(para_out + 7 * b, o.para, 56);
    if (counts_out) {
      counts_out[4 * b] = so.n_edge[0]; counts_out[4 * b + 1] = so.n_edge[1];
      counts_out[4 * b + 2] = so.n_plane[0]; counts_out[4 * b + 3] = so.n_plane[1];
    }
  }
  // this sweep's less-sharp / less-flat clouds are the next call's targets (:556-562)
  std::vector<int> lo(2 * B + 1), co(2 * B + 1, 0);
  for (int b = 0; b <= B; ++b) { lo[b] = ls_off[b]; lo[B + b] = ls_off[B] + lf_off[b]; }
  for (int g = 0; g < 2 * B; ++g) co[g + 1] = co[g] + (lo[g + 1] - lo[g] + 31) / 32;  // 32-point chunks per cloud
  CK(cudaMemcpyAsync(d.od_last_off, lo.data(), sizeof(int) * (2 * B + 1), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d.od_chunk_off, co.data(), sizeof(int) * (2 * B + 1), cudaMemcpyHostToDevice, s));
  if (ls_off[B] > 0) CK(cudaMemcpyAsync(d.od_last, less_sharp, sizeof(float4) * (size_t)ls_off[B], kind, s));
  if (lf_off[B] > 0) CK(cudaMemcpyAsync(d.od_last + ls_off[B], less_flat, sizeof(float4) * (size_t)lf_off[B], kind, s));
  ctx->launches += launch_odom_sort(d, ls_off[B] + lf_off[B], ctx->od_tmp, ctx->od_tmp_bytes, s);
  ctx->launches += launch_odom_meta(d, co[2 * B], s);
  CK(cudaMemcpyAsync(ctx->h_err, d.err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));  // indexing may reject the clouds
  CK(cudaStreamSynchronize(s));
  if (*ctx->h_err != 0) {
    const int e = *ctx->h_err;
    cudaMemsetAsync(d.err_flag, 0, sizeof(int), s);
    cudaMemsetAsync(d.od_last_off, 0, sizeof(int) * (2 * B + 1), s);  // the rejected clouds are not kept
    cudaMemsetAsync(d.od_chunk_off, 0, sizeof(int) * (2 * B + 1), s);
    ctx->err = "previous-sweep clouds outside the supported domain (ring numbers 0..255, coordinates within 256 m)";
    return e;
  }
  return S2M_OK;
}

// ---- profiling ------------------------------------------------------------------
extern "C" int s2m_set_profiling(s2m_ctx* ctx, int on) {
  if (!ctx) return S2M_ERR_ARG;
  for (s2m_ctx* ch : ctx->children) s2m_set_profiling(ch, on);
  ctx->profiling = on != 0;
  ctx->count_candidates = on >= 2;
  return S2M_OK;
}
extern "C" int s2m_k4_profile(s2m_ctx* ctx, int reset, double* ms_total, long long* launches, double* alg_bytes) {
  if (!ctx) return S2M_ERR_ARG;
  if (!ctx->children.empty()) {  // sums over the lanes
    double ms = 0, by = 0;
    long long n = 0;
    for (s2m_ctx* ch : ctx->children) {
      double m1 = 0, b1 = 0;
      long long n1 = 0;
      int rc = s2m_k4_profile(ch, reset, &m1, &n1, &b1);
      if (rc != S2M_OK) return rc;
      ms += m1; by += b1; n += n1;
    }
    if (ms_total) *ms_total = ms;
    if (launches) *launches = n;
    if (alg_bytes) *alg_bytes = by;
    return S2M_OK;
  }
  CK(cudaSetDevice(ctx->P.device));
  prof_resolve(ctx);
  if (ms_total) *ms_total = ctx->phase_ms[S2M_PHASE_ASSOCIATE];
  if (launches) *launches = ctx->k4_launches;
  if (alg_bytes) *alg_bytes = ctx->k4_bytes;
  if (reset) {
    for (double& v : ctx->phase_ms) v = 0;
    ctx->k4_launches = 0; ctx->k4_bytes = 0;
  }
  return S2M_OK;
}
extern "C" int s2m_phase_profile(s2m_ctx* ctx, int reset, double ms[S2M_N_PHASES]) {
  if (!ctx || !ms) return S2M_ERR_ARG;
  if (!ctx->children.empty()) {  // sums over the lanes
    for (int i = 0; i < S2M_N_PHASES; ++i) ms[i] = 0;
    for (s2m_ctx* ch : ctx->children) {
      double m1[S2M_N_PHASES];
      int rc = s2m_phase_profile(ch, reset, m1);
      if (rc != S2M_OK) return rc;
      for (int i = 0; i < S2M_N_PHASES; ++i) ms[i] += m1[i];
    }
    return S2M_OK;
  }
  CK(cudaSetDevice(ctx->P.device));
  prof_resolve(ctx);
  for (int i = 0; i < S2M_N_PHASES; ++i) ms[i] = ctx->phase_ms[i];
  if (reset) {
    for (double& v : ctx->phase_ms) v = 0;
    ctx->k4_launches = 0; ctx->k4_bytes = 0;
  }
  return S2M_OK;
}

// ---- sharded-map mode: not wired in this build ------------------------------------
// Slab of world x owned by `rank`: contiguous cube columns of the initial window (world cube
// columns -10..10, laserMapping.cpp:74-79), the first and last slab open-ended. No GPU needed.
extern "C" int s2m_shard_slab(int rank, int world, float* x_lo, float* x_hi) {
  if (world < 1 || rank < 0 || rank >= world || !x_lo || !x_hi) return S2M_ERR_ARG;
  *x_lo = -INFINITY; *x_hi = INFINITY;
  if (world > 1) {
    const int lo_col = -10 + (kWinI * rank) / world;
    const int hi_col = -10 + (kWinI * (rank + 1)) / world;
    if (rank > 0) *x_lo = (float)(50.0 * lo_col - 25.0);
    if (rank < world - 1) *x_hi = (float)(50.0 * hi_col - 25.0);
  }
  return S2M_OK;
}
extern "C" int s2m_shard_unique_id(void* id128) {
  if (!id128) return S2M_ERR_ARG;
  if (!nccl::load()) return S2M_ERR_NCCL;
  nccl::UniqueId id;
  if (nccl::GetUniqueId(&id) != 0) return S2M_ERR_NCCL;
  std::memcpy(id128, &id, sizeof(id));
  return S2M_OK;
}
extern "C" int s2m_shard_init(s2m_ctx* ctx, const void* id128) {
  if (!ctx || !id128 || ctx->P.shard_world <= 1 || !ctx->children.empty()) return S2M_ERR_ARG;
  if (!nccl::load()) { ctx->err = "libnccl.so.2 not found"; return S2M_ERR_NCCL; }
  CK(cudaSetDevice(ctx->P.device));
  nccl::UniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  const int rc = nccl::CommInitRank(&ctx->comm, ctx->P.shard_world, id, ctx->P.shard_rank);
  if (rc != 0) { ctx->err = std::string("ncclCommInitRank: ") + (nccl::GetErrorString ? nccl::GetErrorString(rc) : "?"); ctx->comm = nullptr; return S2M_ERR_NCCL; }
  return S2M_OK;
}
extern "C" int s2m_shard_profile(s2m_ctx* ctx, int reset, double* allreduce_ms_total, long long* count) {
  if (!ctx) return S2M_ERR_ARG;
  CK(cudaSetDevice(ctx->P.device));
  CK(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i + 1 < ctx->ar_used; i += 2) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ar_pool[i], ctx->ar_pool[i + 1]) == cudaSuccess) { ctx->ar_ms += ms; ctx->ar_count++; }
  }
  ctx->ar_used = 0;
  if (allreduce_ms_total) *allreduce_ms_total = ctx->ar_ms;
  if (count) *count = ctx->ar_count;
  if (reset) { ctx->ar_ms = 0; ctx->ar_count = 0; }
  return S2M_OK;
}
