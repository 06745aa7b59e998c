// s2m_internal.h -- device data layout shared by the kernels and the C-ABI layer.
//
// Vocabulary (follows the reference, laserMapping.cpp):
//   slot     one independent sequence: own map, window, odometry correction
//   class    0 = corner, 1 = surf
//   segment  g = 2*slot + class; every per-point array is packed by segment
//   cube     50 m map cube; (ci,cj,ck) are WORLD cube coordinates, i.e. the
//            reference's array index minus laserCloudCenWidth/Height/Depth
//   window   the 21x21x11 cubes the reference keeps (:77-79)
//   valid    the 5x5x3 block around the sensor, clipped to the window (:513-530)
//   store    the device map of one segment: entries sorted by a 64-bit key
//            [cube:30][pending:1][payload:33]; filtered entries (payload = voxel
//            coordinate inside the cube, one per cube x voxel) precede pending
//            entries (payload = arrival number; raw points pushed into cubes
//            outside the valid block, :753-759 vs :789-791).  Key order ==
//            the reference's gather order (i, j, k loops, each cube's cloud in
//            VoxelGrid output order followed by pushed-back raw points).
//   local    concatenation of the valid cubes of a store, in key order; the
//            position in it is the kNN index the reference would report
//   cell     1 m lattice cell floor(p); candidates of a query are the points of
//            its 27 neighbouring cells (exact for the reference's d2[4] < 1 gate)
//   cell index  per segment, PERSISTENT: an open-addressing table cell -> (count, bucket) and a
//            pool of small fixed-size buckets (chained) holding (x, y, z, tag) of every local-map point.
//            Built in bulk when the valid block changes, otherwise updated by the map update with
//            the few thousand points a frame changes.  tag = [cube of the valid block:7]
//            [pending:1][voxel z,y,x:3x8 | arrival rank:24] orders like the reference's gather
//            order, so (d2, tag) ties break exactly like (d2, index)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "s2m_math.cuh"

namespace s2m {

#ifndef S2M_OD_FB
#define S2M_OD_FB 32  // blocks per SM of odom_fallback_kernel (one warp per listed query, latency-bound)
#endif
#ifndef S2M_OD_MINB
#define S2M_OD_MINB 6  // resident blocks per SM of odom_associate_kernel: latency-bound walks, more warps win (3.9 -> 3.0 ms)
#endif
#ifndef S2M_K4A_MINB
#define S2M_K4A_MINB 8  // resident blocks per SM the kNN kernel is compiled for (<= 64 registers)
#endif
#ifndef S2M_KNN_GROUP
#define S2M_KNN_GROUP 1  // K4a: 1 = grouped search over shared-memory candidate pools, 0 = thread per query
#endif
#ifndef S2M_K4G_MINB
#define S2M_K4G_MINB 4  // resident blocks per SM of the grouped kNN kernel (shared-memory bound)
#endif
#ifndef S2M_K4B_MINB
#define S2M_K4B_MINB 4  // ... and the fit / residual kernel (128 registers, FP64)
#endif
constexpr int kEvalTilesPerBlock = 8;  // ... one evaluate_kernel block
constexpr int kTile = 128;       // queries per block of the association / evaluation kernels
constexpr int kPartial = 32;     // doubles per block partial: 28 sums, n_edge, n_plane, cand_corner, cand_surf
constexpr int kCols = 25;        // (i,j) columns of the valid block
constexpr int kValidCubes = 75;
constexpr int kMaxBatch = 64;    // 7 segment bits in the 32-bit cell sort key
constexpr int kWinI = 21, kWinJ = 21, kWinK = 11;
constexpr uint64_t kSentinel64 = ~0ull;
constexpr uint32_t kSentinel32 = ~0u;

// store key helpers -----------------------------------------------------------
constexpr int kCubeBiasIJ = 2048, kCubeBiasK = 32;
S2M_HD uint32_t pack_cube(int ci, int cj, int ck) {
  return ((uint32_t)(ci + kCubeBiasIJ) << 18) | ((uint32_t)(cj + kCubeBiasIJ) << 6) | (uint32_t)(ck + kCubeBiasK);
}
S2M_HD bool cube_in_range(int ci, int cj, int ck) {
  return ci > -kCubeBiasIJ && ci < kCubeBiasIJ - 1 && cj > -kCubeBiasIJ && cj < kCubeBiasIJ - 1 &&
         ck > -kCubeBiasK && ck < kCubeBiasK - 1;
}
S2M_HD void unpack_cube(uint32_t c, int& ci, int& cj, int& ck) {
  ci = (int)(c >> 18) - kCubeBiasIJ; cj = (int)((c >> 6) & 0xFFF) - kCubeBiasIJ; ck = (int)(c & 0x3F) - kCubeBiasK;
}
S2M_HD uint64_t store_key(uint32_t cube, uint32_t pending, uint64_t payload) {
  return ((uint64_t)cube << 34) | ((uint64_t)pending << 33) | (payload & 0x1FFFFFFFFull);
}
S2M_HD uint32_t key_cube(uint64_t k) { return (uint32_t)(k >> 34); }
S2M_HD uint32_t key_pending(uint64_t k) { return (uint32_t)((k >> 33) & 1); }
S2M_HD uint64_t key_payload(uint64_t k) { return k & 0x1FFFFFFFFull; }

// voxel coordinate of p relative to the first voxel touching the cube, per axis
S2M_HD int voxel_rel(float p, int cube, float inv_leaf) {
  float lo = (float)(50.0 * (double)cube - 25.0);
  return voxel_coord(p, inv_leaf) - voxel_coord(lo, inv_leaf);
}

struct FrameDesc {          // per slot, written by the host every call
  double pose[7];           // initial guess q_w_curr, t_w_curr (row A)
  int win_lo[3], win_hi[3]; // window, world cube coords, inclusive
  int val_lo[3], val_hi[3]; // valid block clipped to the window, inclusive (lo > hi: none)
  int origin[3];            // first 1 m cell of the valid block: 50*val_lo - 25
  int active;               // slot takes part in this call
  int allow_opt;            // 0: skip rows K..S
  int idx_flags;            // bit 0: rebuild the cell index of this slot now; bit 1: keep it up to date in the map update
  unsigned long long seq_base[2];
};

struct SlotOut {            // per slot, read back by the host after every call
  double pose[7];
  int n_ds[2], n_local[2], n_store[2];
  int n_edge[2], n_plane[2];
  int optimized, lm_iters[2], lm_term[2];
  int pad;
  double cost_initial[2], cost_final[2];
  double cand[2];           // candidate points visited by corner / surf queries (outer 0)
};

// All device pointers of one context (sizes are host-known capacities).
struct Dev {
  int B, G;                     // slots, segments
  int vox_bits;                 // bits of a voxel coordinate inside a cube (delta sort key)
  int delta_pbits;              // payload bits of the delta sort key
  int shard_world;              // >1: spatially sharded map (x-slabs per rank)
  float shard_lo, shard_hi;     // this rank's slab in world x, [lo, hi)
  double* shard_sums;           // [B][kPartial] per-rank sums awaiting the allreduce
  int* shard_counts;            // [G] map points this rank owns in the valid block (allreduced for the guard)
  float inv_leaf[2];
  // ---- per call small tables (device copies of host arrays)
  FrameDesc* desc;              // [B]
  int* in_off;                  // [G+1] packed offsets of the incoming clouds
  int* lp_off;                  // [G+1] packed offsets of the raw local-map points a merge absorbs (exact: counts are read back)
  int* so_off;                  // [G+1] packed offsets of the whole stores (merge index space)
  int* st_base;                 // [G]   base of each segment in the store arrays
  int* st_cap;                  // [G]
  int* hash_off;                // [G+1] base of each segment's cell table
  // ---- incoming + down-sampled scan
  float4* in_pts;               // [cap_in] packed by segment
  uint64_t *vkey, *vkey2;       // [cap_sort]
  uint32_t *vval, *vval2;       // [cap_sort]
  uint32_t *flag, *scan;        // [cap_sort]
  uint32_t* bbox;               // [G][6] min xyz / max xyz, order-preserving uint encoding
  float4* ds_pts;               // [cap_in] packed, voxel-filtered scan (sensor frame)
  int* ds_off;                  // [G+1] device-computed
  // ---- store (double buffered)
  uint64_t* st_key[2];
  float4* st_pt[2];
  int* st_n;                    // [G] entries in the current buffer
  int* st_n_new;                // [G]
  // ---- local map + cell index
  int* rng_start;               // [G][25]
  int* loc_off;                 // [G][26]
  int* lp_cnt;                  // [2G] points of the local map (read back: sizes the index exactly), then how many of them are raw
  int* nbr;                     // [cap_in][6] K4a -> K4b: (n << 1 | gate), five entry numbers in d.bkt
  int* knn_ticket;              // next 32-query work unit of knn_kernel
  uint32_t *qs_key, *qs_key2;   // [cap_in] (segment, 2 m block) key of every scan point: in scan order / bucketed by block
  uint32_t* qs_val2;            // [cap_in] position in ds_pts of the bucketed points
  int* cand27;                  // [cap_in] map points in the 27 cells of a query (profiling only)
  int count_cand;               // fit_kernel adds cand27 into the slot's candidate counters
  unsigned long long* knn_stats;  // [4] queries the grouped search handed to the per-thread search
  float4* od_last;              // odometry: less-sharp / less-flat clouds of the previous sweep, class-major
  int* od_last_off;             // [2B+1]
  float4* od_sorted;            // the same clouds ordered by 1 m cell inside each segment, .w = index | ring << 24
  uint32_t* od_ckey;            // cell key of every entry of od_sorted
  int4* od_corr;                // [cap_in] per query: closest, second, third index (-1 none), pending flags
  float* od_bestd;              // [cap_in] distance of the closest point
  int* od_fb_list;              // [cap_in] queries the cell search could not settle
  int* od_fb_cnt;
  int *od_first_ge, *od_last_le;  // [2B][257] ring tables of the previous clouds (first index with ring >= r, last with ring < r)
  unsigned long long *od_key, *od_key2;
  uint32_t *od_val, *od_val2;
  float4* od_meta;              // [chunks][2] box + ring range of every 32-point chunk of od_last
  int* od_chunk_off;            // [2B+1]
  // ---- persistent cell index (see the vocabulary above)
  unsigned long long* hash_tab; // [hash_off[G]] per segment: [cell:24][count:16][bucket:24], empty = all ones
  int* hmask;                   // [G] slots of the segment's table in use - 1 (set at every rebuild)
  float4* bkt;                  // [4 * bkt_off[G]] buckets of four (x, y, z, tag); tag -1 = unused / removed
  uint32_t* bnext;              // [bkt_off[G]] next bucket of a cell's chain, ~0 none
  int* bkt_off;                 // [G+1] first bucket of each segment's pool (static)
  int* bcnt;                    // [G] buckets handed out
  int* idx_list;                // [G] segments to rebuild this frame (compacted), count in idx_n
  int* idx_poff;                // [G+1] packed offsets of their local points
  int* idx_soff;                // [G+1] first position of each listed segment in the cell-sorted order (bulk build)
  // ---- association / solve
  double* rec;                  // [cap_in][6] cached correspondences
  uint8_t* rec_valid;           // [cap_in]
  double* partials;             // [B][max_tiles][kPartial]: one row per 32-query unit (association) / working block (evaluation)
  int max_tiles;                // rows per slot
  int* ticket;                  // [B] last-block election counters
  LmState* lm;                  // [B]
  SlotOut* out;                 // [B]
  int* err_flag;                // device error code (0 ok)
  // ---- trace (optional)
  int32_t* tr_idx;              // [2][cap_in][5]
  float* tr_d2;                 // [2][cap_in][5]
  uint8_t* tr_used;             // [2][cap_in]
  // ---- map update
  float4* dl_pt;                // [cap_in] transformed scan points (world, float)
  uint64_t *ins_key, *ins_ckey; // [cap_sort]
  float4 *ins_pt, *ins_cpt;     // [cap_sort]
  int* upd_pos;                 // [cap_sort] per delta run: merge-space index of the store entry it re-centroids, -1 none
  int* run_off;                 // [G+1]
  int *dead_n, *dead_lo, *dead_cum; // [G], [G][75], [G][76]: store ranges of the raw points a frame absorbs (pending_gather_kernel)
  uint32_t *aflag, *ascan;      // [cap_lp + 1]
  // ---- cub temp
  void* cub_tmp;
  size_t cub_tmp_bytes;
  int cap_in, cap_sort, cap_lp;
};

struct GuardDesc {  // debug guard bands of one device allocation (S2M_GUARD_BYTES)
  const uint32_t *front, *back;
  unsigned words;
};
int launch_guard_check(const GuardDesc* g, int n, int* bad, cudaStream_t s);

// launchers (s2m_kernels.cu); every one returns the number of kernels it launched
size_t cub_temp_bytes(int cap_sort, int cap_lp);
int launch_voxel_bbox(const Dev& d, int total_in, int longest_segment, cudaStream_t s);
int launch_voxel_filter(const Dev& d, int total_in, int key_bits, cudaStream_t s);
int launch_local_ranges(const Dev& d, int cur, cudaStream_t s);
int launch_index_rebuild(const Dev& d, int cur, int n_seg, int total_points, cudaStream_t s);
int launch_guard(const Dev& d, cudaStream_t s);
int launch_associate(const Dev& d, int outer, int cur, int knn_blocks, int fit_blocks, int n_ds, bool trace, cudaStream_t s);
int launch_query_sort(const Dev& d, int n_ds, cudaStream_t s);
int launch_count27(const Dev& d, int n_ds, cudaStream_t s);
int launch_solve(const Dev& d, int outer, bool from_units, cudaStream_t s);
int launch_reduce_units(const Dev& d, cudaStream_t s);
int launch_evaluate(const Dev& d, int outer, int blocks_per_slot, cudaStream_t s);
int launch_lm_shard(const Dev& d, int outer, int after, cudaStream_t s);
int launch_odom_sort(const Dev& d, int n, void* tmp, size_t tmp_bytes, cudaStream_t s);
size_t odom_sort_temp_bytes(const Dev& d, int n);
int launch_odom_meta(const Dev& d, int nchunks, cudaStream_t s);
int launch_odom_guard(const Dev& d, cudaStream_t s);
int launch_odom_associate(const Dev& d, int outer, int tiles, int fallback_blocks, bool trace, cudaStream_t s);
int launch_finish_pose(const Dev& d, cudaStream_t s);
int launch_map_update(const Dev& d, int cur, int n_ds, int total_lp, int total_store, bool check_pending, bool window_shift,
                      bool identity_pose, cudaStream_t s);
#ifndef S2M_KNN_PRED
#define S2M_KNN_PRED 0   // load only the entries a bucket holds instead of always four
#endif
#ifndef S2M_KNN_BATCH9
#define S2M_KNN_BATCH9 0  // issue the 27 cell probes nine at a time instead of three at a time
#endif
#ifndef S2M_BKT_E
#define S2M_BKT_E 4
#endif
constexpr int kBktE = S2M_BKT_E;               // entries per bucket (a multiple of 4: 64 or 128 bytes)
constexpr uint32_t kNoBkt = 0xFFFFFFu;         // 24-bit "none" inside a table entry
constexpr unsigned long long kCellCount1 = 1ull << 24;
int launch_knn_debug(const Dev& d, int cur, int slot, int cls, const float* d_q, int n, int32_t* d_idx, float* d_d2,
                     cudaStream_t s);
int launch_transform_cloud(const double* d_pose7, const float4* in, float4* out, int n, cudaStream_t s);
int launch_gather_local(const Dev& d, int cur, int g, float4* out, cudaStream_t s);
int launch_surround(const Dev& d, int cur, int slot, float4* out, int cap, int* n_out, cudaStream_t s);

}  // namespace s2m
