/* s2m.h -- C ABI of the B200-native scan-to-map registration engine.
 *
 * Drop-in boundary for the hot path of SC-A-LOAM's laserMapping node:
 * /root/reference/src/laserMapping.cpp:310-802 (process(), from
 * transformAssociateToMap() to the per-cube re-filter).  The reference has no
 * plugin / FFI interface for this path (alaserMapping is an executable with
 * file-scope state, laserMapping.cpp:66-140); these entry points are what a ROS
 * shim that keeps the node's topics and its two parameters would bind
 * (INTEGRATION.md shows that shim).  Plain pointers and sizes only.
 * Further down, the rows SURVEY.md 8f ranks next: map checkpoint (s2m_pcd_*, s2m_checkpoint_*),
 * scan-to-scan odometry (s2m_odom_*, laserOdometry.cpp:220-591) and feature extraction
 * (s2m_fx_*, scanRegistration.cpp:116-454).
 *
 * Conventions
 *   - point clouds are packed float[4*n]: x, y, z, intensity (pcl::PointXYZI
 *     payload, include/aloam_velodyne/common.h:43);
 *   - quaternions are double[4] in Eigen coefficient order x, y, z, w, as in
 *     `double parameters[7]` (laserMapping.cpp:110-112); translations double[3];
 *   - a context holds `batch` independent sequences ("slots"), each with its own
 *     map, window and odometry correction; the batch calls advance every active
 *     slot by one frame inside the same kernel launches;
 *   - a context is single-caller; distinct contexts are independent;
 *   - return value: 0 ok, >0 soft status, <0 error (never throws, never aborts).
 */
#ifndef S2M_H_
#define S2M_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2M_OK 0
#define S2M_MAP_TOO_SMALL 1   /* guard laserMapping.cpp:555 failed: pose = odometry guess (ROS_WARN :733) */
#define S2M_ERR_CUDA (-1)
#define S2M_ERR_ARG (-2)
#define S2M_ERR_CAPACITY (-3) /* an input or the map exceeded the capacities given at create */
#define S2M_ERR_RANGE (-4)    /* coordinates outside the supported lattice range */
#define S2M_ERR_NCCL (-5)
#define S2M_ERR_IO (-6)       /* a checkpoint / PCD file could not be read or written */
#define S2M_ERR_INTERNAL (-7) /* debug builds / S2M_GUARD_BYTES: a guard band around a device buffer was overwritten */

typedef struct s2m_ctx s2m_ctx;

/* Parameters. The first two are the node's ROS parameters, read as float like
 * laserMapping.cpp:913-919; the constants of the reference (21x21x11 cubes of
 * 50 m, 5x5x3 valid block, k=5, d2<1, ratio 3, plane tol 0.2, Huber 0.1,
 * 2 outer x 4 LM iterations) are compiled in. */
typedef struct s2m_params {
  float line_res;        /* mapping_line_resolution  (default 0.4) */
  float plane_res;       /* mapping_plane_resolution (default 0.8) */
  int device;            /* CUDA device ordinal */
  int batch;             /* independent sequences held by this context (>=1) */
  int cap_corner_in;     /* per slot: max points of one incoming corner cloud */
  int cap_surf_in;       /* per slot: max points of one incoming surf cloud */
  int cap_map_corner;    /* per slot: max corner points in the 21x21x11 window */
  int cap_map_surf;      /* per slot: max surf points in the window */
  int skip_optimization; /* debug: rows K..S skipped, pose = guess (map-parity tests) */
  int trace;             /* debug: keep per-query kNN results of the last call */
  int shard_rank;        /* sharded-map mode: this GPU's rank, 0 when unsharded */
  int shard_world;       /* sharded-map mode: number of GPUs, 1 when unsharded */
  int lanes;             /* 0: plain context (<= 64 slots, work issued on the caller's thread).
                            >= 1: the slots are split over this many concurrent lanes (own stream + host
                            thread each, <= 64 slots per lane) and s2m_register_batch_submit/_wait work */
} s2m_params;

/* Per-registration statistics (the reference's commented-out printf lines,
 * laserMapping.cpp:553-729, carry the same quantities). */
typedef struct s2m_stats {
  int n_corner_in, n_surf_in;   /* incoming */
  int n_corner_ds, n_surf_ds;   /* after the scan voxel filter (:543-551) */
  int n_map_corner, n_map_surf; /* points in the valid 5x5x3 block (:539-540) */
  int n_edge[2], n_plane[2];    /* residual blocks per outer iteration (:621, :686) */
  int optimized;                /* 1 if guard :555 passed */
  int lm_iters[2];              /* LM iterations executed per solve (<=4) */
  int lm_term[2];               /* 0 max iterations, 1 gradient, 2 parameter, 3 function tolerance, 4 radius, 5 no residuals, 6 invalid steps */
  double cost_initial[2], cost_final[2];
} s2m_stats;

void s2m_default_params(s2m_params* p);
int s2m_create(const s2m_params* p, s2m_ctx** out);
void s2m_destroy(s2m_ctx* ctx);
const char* s2m_strerror(int code);
const char* s2m_last_error(s2m_ctx* ctx); /* text of the last CUDA/argument error */

/* Run all work of this context on `cuda_stream` (a cudaStream_t) instead of the
 * context's own stream; lets a caller bracket calls with its own events. */
int s2m_set_stream(s2m_ctx* ctx, void* cuda_stream);

/* One frame of one sequence (slot 0): rows A,B,C,V,G,(K,E,F,R,L,Q,S)x2,U,I,W of
 * laserMapping.cpp:310-802.  Inputs are what the node receives on
 * /laser_cloud_corner_last, /laser_cloud_surf_last and /laser_odom_to_init
 * (laserMapping.cpp:279-298); outputs are q_w_curr / t_w_curr as published on
 * /aft_mapped_to_init (:861-872).  Host buffers. */
int s2m_register(s2m_ctx* ctx, const float* corner_xyzi, int n_corner, const float* surf_xyzi,
                 int n_surf, const double q_wodom[4], const double t_wodom[3], double q_w_out[4],
                 double t_w_out[3], s2m_stats* stats);

/* One frame of every slot.  Slot s reads points
 * [corner_off[s], corner_off[s+1]) of the packed corner array (same for surf),
 * pose q_wodom[4*s..], t_wodom[3*s..]; writes q_w_out[4*s..], t_w_out[3*s..],
 * stats[s] (may be NULL) and status[s] (S2M_OK / S2M_MAP_TOO_SMALL).
 * active[s]==0 skips a slot (NULL = all active).  Host buffers. */
int s2m_register_batch(s2m_ctx* ctx, const float* corner_xyzi, const int* corner_off,
                       const float* surf_xyzi, const int* surf_off, const double* q_wodom,
                       const double* t_wodom, const int* active, double* q_w_out, double* t_w_out,
                       s2m_stats* stats, int* status);

/* Same, with the two packed clouds already resident in device memory
 * (offset / pose / output arrays stay on the host). */
int s2m_register_batch_dev(s2m_ctx* ctx, const float* d_corner_xyzi, const int* corner_off,
                           const float* d_surf_xyzi, const int* surf_off, const double* q_wodom,
                           const double* t_wodom, const int* active, double* q_w_out,
                           double* t_w_out, s2m_stats* stats, int* status);

/* Asynchronous pair of the batch call, multi-lane contexts only (params.lanes >= 1).
 * submit enqueues one frame of every slot and returns at once; wait completes the
 * OLDEST frame in flight (returns its error code; its outputs are then written).
 * At most two frames may be in flight: the host->device copy of frame f+1 then overlaps
 * the registration of frame f, which is how the node's callback queue
 * (laserMapping.cpp:232-306 pops the next message while the last one is still being
 * published) maps onto a copy engine.  Offsets, poses and `active` are copied at submit;
 * the clouds and the output arrays must stay valid until the matching wait.
 * device_ptrs != 0: the clouds are device pointers (as s2m_register_batch_dev). */
int s2m_register_batch_submit(s2m_ctx* ctx, const float* corner_xyzi, const int* corner_off,
                              const float* surf_xyzi, const int* surf_off, const double* q_wodom,
                              const double* t_wodom, const int* active, double* q_w_out,
                              double* t_w_out, s2m_stats* stats, int* status, int device_ptrs);
int s2m_register_batch_wait(s2m_ctx* ctx);

/* --- map checkpoint ---------------------------------------------------
 * PCD v0.7 binary, FIELDS x y z intensity, float32: the layout pcl::io::savePCDFileBinary
 * produces for PointXYZI (laserPosegraphOptimization.cpp:695) and the reference ships under
 * utils/sample_data/<seq>/Scans/ (188-byte header for 5-digit counts, 16 B per point).  Host-only helpers.
 * s2m_pcd_read returns the number of points in the file and copies min(n, cap) of them. */
int s2m_pcd_write(const char* path, const float* xyzi, int n);
int s2m_pcd_read(const char* path, float* xyzi_out, int cap);
/* The reference keeps its map only in RAM; these two make the node resumable.  save writes
 * <prefix>.corner.pcd, <prefix>.surf.pcd and <prefix>.state (window centre laserMapping.cpp:74-76
 * and the wmap<-wodom correction :116-117, hex floats).  load restores them into `slot`; a context
 * that continues from a checkpoint produces the bits of the uninterrupted run.  load returns the
 * number of points that fell outside the restored window (0 for this library's own files). */
int s2m_checkpoint_save(s2m_ctx* ctx, int slot, const char* prefix);
int s2m_checkpoint_load(s2m_ctx* ctx, int slot, const char* prefix);

/* wmap<-wodom correction kept by transformUpdate() (laserMapping.cpp:149-153);
 * the shim's high-rate odometry relay (:198-230) composes with it on the host. */
int s2m_get_correction(s2m_ctx* ctx, int slot, double q_wmap_wodom[4], double t_wmap_wodom[3]);

/* Row X (laserMapping.cpp:845-849): transform a full-resolution cloud with the
 * slot's current pose, in double, rounded to float. Host buffers. */
int s2m_transform_cloud(s2m_ctx* ctx, int slot, const float* in_xyzi, int n, float* out_xyzi);

/* ---- map access (test hooks, checkpoint/resume) ---- */

/* Replace the slot's map: every point is pushed into its cube raw, in upload
 * order, as laserMapping.cpp:753-759 would; returns the number dropped for lying
 * outside the 21x21x11 window (>=0) or an error. */
int s2m_map_upload(s2m_ctx* ctx, int slot, const float* corner_xyzi, int n_corner,
                   const float* surf_xyzi, int n_surf);
/* Whole-window map of one class (0 corner, 1 surf): cubes in the order of the
 * reference's gather loops (i, j, k nested, laserMapping.cpp:513-517), each cube
 * in its stored order. Returns the count. */
int s2m_map_download(s2m_ctx* ctx, int slot, int cls, float* out_xyzi, int cap);
/* Local map as gathered for a sensor at centre_t (rows B, C), in gather order
 * (laserMapping.cpp:513-538) -- this order defines the kNN index space. */
int s2m_get_local_map(s2m_ctx* ctx, int slot, int cls, const double centre_t[3], float* out_xyzi,
                      int cap);
/* Surround cloud published every 5th frame (laserMapping.cpp:807-822): corner
 * then surf of each valid cube. */
int s2m_get_surround(s2m_ctx* ctx, int slot, float* out_xyzi, int cap);
/* Window centre laserCloudCenWidth/Height/Depth (laserMapping.cpp:74-76). */
int s2m_get_window(s2m_ctx* ctx, int slot, int cen[3]);

/* Exact bounded kNN(5) (row K) of world-frame float queries against the local map
 * around centre_t.  idx5 are indices into s2m_get_local_map()'s order, d2_5 the
 * float squared distances ascending, ties by lower index.  The search is bounded
 * by the reference's gate d2[4] < 1.0 (laserMapping.cpp:585, :653): a query whose
 * 5th neighbour is not within 1 m gets idx -1 / d2 +inf in every slot it could
 * not fill from its 27 one-metre cells. */
int s2m_debug_knn(s2m_ctx* ctx, int slot, int cls, const double centre_t[3], const float* q_xyz,
                  int n, int32_t* idx5, float* d2_5);

/* Trace of the last register call (params.trace=1): per outer iteration and
 * class, kNN idx/d2 (5 per down-sampled scan point) and the used flag. */
int s2m_trace_cloud(s2m_ctx* ctx, int slot, int cls, float* out_xyzi, int cap); /* down-sampled scan */
int s2m_trace_knn(s2m_ctx* ctx, int slot, int outer, int cls, int32_t* idx5, float* d2_5,
                  uint8_t* used, int cap);
/* pose after the solve (7), the 28 reduced sums at the first evaluation (21
 * upper-triangular JtJ, 6 Jtr, cost), per-iteration [cost, cost_change, radius,
 * step_norm, model_change, accepted] x 4 */
int s2m_trace_lm(s2m_ctx* ctx, int slot, int outer, double pose7[7], double sums28[28],
                 double iters24[24], int* n_iter, int* termination);

/* Debug aid (compute-sanitizer stand-in): with the environment variable S2M_GUARD_BYTES=<n> set when a context is
 * created, every device buffer is allocated with n pattern bytes before and after it, checked after every call
 * (an overwritten band fails the call with S2M_ERR_INTERNAL).  Returns the number of overwritten guard words so
 * far, 0 when the mode is off. */
int s2m_debug_guard_check(s2m_ctx* ctx);

/* Queries the grouped kNN kernel handed to the thread-per-query search so far (candidate box larger than the
 * shared-memory pool, or too many ties at the fifth distance); both searches are exact, this is a tuning counter. */
long long s2m_debug_knn_fallbacks(s2m_ctx* ctx);

/* Kernel launches issued by this context so far (bench.py's gpu_launches). */
long long s2m_launch_count(s2m_ctx* ctx);
/* CUDA-event time (ms) of the association (K4) launches since the last call with reset!=0, their count, and the
 * algorithmic bytes they processed (SURVEY.md 8d's B_K4). s2m_set_profiling: 0 off, 1 events at the phase boundaries
 * of every frame, 2 additionally one small kernel per frame (outside the K4 bracket) that counts the map points in
 * every query's 27 cells -- the candidate term of B_K4; s2m_k4_profile's alg_bytes needs level 2. */
int s2m_set_profiling(s2m_ctx* ctx, int on);
int s2m_k4_profile(s2m_ctx* ctx, int reset, double* ms_total, long long* launches,
                   double* alg_bytes);
/* CUDA-event time (ms) per phase of the frames run with profiling on -- the
 * reference's stopwatch phases (laserMapping.cpp:308-859 t_shift, t_tree, t_data,
 * t_solver, t_add, t_filter) regrouped by kernel family. */
#define S2M_PHASE_INPUT 0     /* copy of the incoming clouds into the packed device buffer */
#define S2M_PHASE_VOXEL 1     /* scan voxel-grid filter (row V) */
#define S2M_PHASE_INDEX 2     /* local map + 1 m cell index (rows C, T) */
#define S2M_PHASE_ASSOCIATE 3 /* fused association kernel, both outer iterations (rows P..Q) */
#define S2M_PHASE_SOLVE 4     /* LM evaluations + steps (row S) */
#define S2M_PHASE_UPDATE 5    /* map insert / re-filter / evict (rows I, W, B) */
#define S2M_PHASE_READBACK 6  /* pose + counter read-back and host synchronisation */
#define S2M_N_PHASES 7
int s2m_phase_profile(s2m_ctx* ctx, int reset, double ms[S2M_N_PHASES]);

/* Sharded-map mode (BASELINE config 5): x-slabs of cubes per GPU, one allreduce
 * of the 28-double normal-equation block per evaluation.  nccl_unique_id is the
 * 128-byte ncclUniqueId obtained on rank 0 with s2m_shard_unique_id. */
int s2m_shard_unique_id(void* id128);
/* world-x interval [lo, hi) a rank owns (needs no GPU; used by the CPU multi-process tests) */
int s2m_shard_slab(int rank, int world, float* x_lo, float* x_hi);
int s2m_shard_init(s2m_ctx* ctx, const void* id128);
int s2m_shard_profile(s2m_ctx* ctx, int reset, double* allreduce_ms_total, long long* count);

/* --- scan-to-scan odometry --------------------------------------------
 * laserOdometry.cpp:220-591 with DISTORTION 0: per sweep, the sharp / flat points are matched against
 * the previous sweep's less-sharp / less-flat clouds (exact nearest neighbour :303/:392, ring-
 * constrained second / third neighbours :313-357/:401-452), LidarEdgeFactor + LidarPlaneFactor
 * (lidarFactor.hpp:12-104), HuberLoss(0.1), two passes of a 4-iteration Ceres solve (:277, :495-500),
 * then q_w_curr / t_w_curr are advanced (:504-505).  `batch` independent sequences per context; the
 * four clouds of each are packed with B+1 offsets, host pointers or (device_ptrs != 0) device pointers,
 * e.g. straight from s2m_fx_device_cloud.  The first call of a slot only initialises (:267-271).
 * Domain: int(intensity) of the less-sharp / less-flat points (the ring number, :310) in 0..255 and
 * coordinates within +-256 m, else S2M_ERR_RANGE; cap_sharp + cap_flat < 2^20, cap_less_* < 2^24.
 * The context is an s2m_ctx: s2m_destroy, s2m_last_error, s2m_launch_count, s2m_trace_knn apply. */
int s2m_odom_create(int device, int batch, int cap_sharp, int cap_flat, int cap_less_sharp,
                    int cap_less_flat, int trace, s2m_ctx** out);
int s2m_odom_step_batch(s2m_ctx* ctx, const float* sharp, const int* sharp_off, const float* flat,
                        const int* flat_off, const float* less_sharp, const int* less_sharp_off,
                        const float* less_flat, const int* less_flat_off, int device_ptrs,
                        double* q_w_out, double* t_w_out, double* para_out, int* counts_out);

/* --- feature extraction ----------------------------------------------
 * scanRegistration.cpp:116-454 (laserCloudHandler) for a batch of raw sweeps: the five clouds
 * the node publishes -- /velodyne_cloud_2 (ring-major, intensity = ring + 0.1 * relative time),
 * /laser_cloud_sharp, /laser_cloud_less_sharp, /laser_cloud_flat, /laser_cloud_less_flat
 * (:415-446).  laserMapping consumes less_sharp / less_flat / full (relayed unchanged by
 * laserOdometry.cpp:574-590).  The results stay on the device until the next extract, so they
 * can be handed to s2m_register_batch_dev without leaving HBM. */
#define S2M_SENSOR_HDL64 0 /* LIDAR_TYPE / N_SCANS pairs of scanRegistration.cpp:172-208 */
#define S2M_SENSOR_VLP16 1
#define S2M_SENSOR_OS1_64 2
#define S2M_SENSOR_HDL32 3
#define S2M_FX_FULL 0
#define S2M_FX_SHARP 1
#define S2M_FX_LESS_SHARP 2
#define S2M_FX_FLAT 3
#define S2M_FX_LESS_FLAT 4
typedef struct s2m_fx s2m_fx;
typedef struct s2m_fx_params {
  int device;
  int batch;            /* sweeps per call */
  int cap_points;       /* max raw points of one sweep */
  int sensor;           /* S2M_SENSOR_* */
  double minimum_range; /* the node's minimum_range parameter (:458) */
} s2m_fx_params;
int s2m_fx_create(const s2m_fx_params* p, s2m_fx** out);
void s2m_fx_destroy(s2m_fx* fx);
const char* s2m_fx_last_error(s2m_fx* fx);
/* xyz: packed float[3*n] raw points of the B sweeps in arrival order (host, or device if
 * device_input != 0); off[B+1]: sweep offsets in points (host). */
int s2m_fx_extract(s2m_fx* fx, const float* xyz, const int* off, int device_input);
/* per-sweep offsets (B+1, in points) of output `which` of the last extract */
int s2m_fx_offsets(s2m_fx* fx, int which, int* off_out);
/* packed xyzi of output `which` over all sweeps -> host; returns the number of points
 * (out may be NULL to query) */
int s2m_fx_download(s2m_fx* fx, int which, float* out_xyzi, int cap_points_total);
/* device pointer to the same packed xyzi (valid until the next extract / destroy) */
const float* s2m_fx_device_cloud(s2m_fx* fx, int which);
long long s2m_fx_launch_count(s2m_fx* fx);

#ifdef __cplusplus
}
#endif
#endif /* S2M_H_ */
