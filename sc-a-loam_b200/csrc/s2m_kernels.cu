// s2m_kernels.cu -- sm_100a kernels of the scan-to-map registration engine.
//
// Kernel map (SURVEY.md section 2.3 numbering; reference rows of section 8a):
//   K1  vox_*            scan voxel-grid filter                 row V  laserMapping.cpp:543-551
//   K3  range/local_key/cand_build  local map + 1 m cell index  rows C,T :510-540, :559-560
//   K4  associate_kernel transform + exact kNN5 + edge PCA / plane QR + residual,
//                        Jacobian, Huber + block reduction      rows P,K,E,F,R,L,Q :578-706
//   K5  evaluate_kernel  re-evaluation at LM trial poses from cached correspondences
//   K6  lm_*_kernel      6x6 trust-region LM step               row S  :713-721
//   K2  delta_*/merge_*  map insert, per-voxel re-centroid, window evict   rows I,W,B :737-802
//   K7  transform_cloud_kernel  full-resolution cloud transform  row X  :845-849
// Sorting / prefix sums use CUB device primitives (plumbing); everything on the
// registration hot path (K4, K5, K6) is hand-written.
#include <cooperative_groups.h>
#include <cub/cub.cuh>

#include "s2m_internal.h"

namespace s2m {

// ----------------------------------------------------------------------------
// small device helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ int find_seg(const int* __restrict__ off, int nseg, int i) {
  int lo = 0, hi = nseg;  // largest g in [0,nseg) with off[g] <= i
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound_u64(const uint64_t* __restrict__ a, int n, uint64_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int lower_bound_u32(const uint32_t* __restrict__ a, int n, uint32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int seg_cls(const Dev& d, int g) { return g >= d.B; }
__device__ __forceinline__ int seg_slot(const Dev& d, int g) { return g >= d.B ? g - d.B : g; }
__device__ __forceinline__ void set_err(const Dev& d, int code) { atomicCAS(d.err_flag, 0, code); }
__device__ __forceinline__ uint32_t cell_hash(uint32_t k) {
  k *= 0x9E3779B1u;
  return k ^ (k >> 15);
}
// First slot probed for a cell: always an even one, so a 16-byte load fetches the first two slots of the probe sequence.
__device__ __forceinline__ uint32_t cell_home(uint32_t k, uint32_t mask) { return cell_hash(k) & mask & ~1u; }

// ----------------------------------------------------------------------------
// K1: scan voxel-grid filter (pcl::VoxelGrid semantics)
// ----------------------------------------------------------------------------
// order-preserving float <-> uint so atomicMin/atomicMax work on floats
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
// bbox[g][0..2] = min (init 0xFFFFFFFF), bbox[g][3..5] = max (init 0), ordered-uint encoded.
// grid (chunks of the longest segment, segments): a block covers kBoxPts consecutive points of ONE segment
// (kBoxPts / 256 per thread, strided so the loads coalesce) and adds its box with six atomics.
constexpr int kBoxPts = 2048;
__global__ void __launch_bounds__(256) vox_bbox_kernel(Dev d) {
  const int g = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p0 = d.in_off[g] + blockIdx.x * kBoxPts, p1 = min(p0 + kBoxPts, d.in_off[g + 1]);
  if (p0 >= p1) return;
  __shared__ uint32_t sm[6][8];
  uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0u, 0u, 0u};
  for (int i = p0 + (int)threadIdx.x; i < p1; i += 256) {
    const float4 p = d.in_pts[i];
    const uint32_t ox = f2ord(p.x), oy = f2ord(p.y), oz = f2ord(p.z);
    mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
    mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
      mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
    }
  if (lane == 0) for (int k = 0; k < 3; ++k) { sm[k][wid] = mn[k]; sm[3 + k][wid] = mx[k]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    uint32_t v = sm[threadIdx.x][0];
    for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? min(v, sm[threadIdx.x][w]) : max(v, sm[threadIdx.x][w]);
    if (threadIdx.x < 3) atomicMin(d.bbox + 6 * g + threadIdx.x, v);
    else atomicMax(d.bbox + 6 * g + threadIdx.x, v);
  }
}
__global__ void vox_bbox_init_kernel(Dev d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 6 * d.G) d.bbox[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;
}

template <bool kNarrow>  // kNarrow: segment + voxel index fit 32 bits -> 4-byte sort keys (one radix pass and a third of the bytes less)
__global__ void vox_key_kernel(Dev d, int n, int key_bits) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  // the block's points almost always belong to one segment: two searches per block instead of one per point
  __shared__ int gs[2];
  if (threadIdx.x == 0) {
    gs[0] = find_seg(d.in_off, d.G, min(i, n - 1));
    gs[1] = find_seg(d.in_off, d.G, min(i + (int)blockDim.x - 1, n - 1));
  }
  __syncthreads();
  if (i >= n) return;
  const int g = gs[0] == gs[1] ? gs[0] : find_seg(d.in_off, d.G, i);
  const float inv = d.inv_leaf[seg_cls(d, g)];
  float bb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) bb[k] = ord2f(d.bbox[6 * g + k]);
  // "Leaf size is too small for the input dataset": PCL warns and passes the input through
  const long long dx = (long long)xfmul(xfsub(bb[3], bb[0]), inv) + 1;
  const long long dy = (long long)xfmul(xfsub(bb[4], bb[1]), inv) + 1;
  const long long dz = (long long)xfmul(xfsub(bb[5], bb[2]), inv) + 1;
  uint64_t key;
  if (dx * dy * dz > 2147483647LL) {
    key = (uint64_t)(i - d.in_off[g]);
  } else {
    // pcl::VoxelGrid: idx = ijk0 + ijk1*div_x + ijk2*div_x*div_y, an int (< 2^31 by the test above)
    const float4 p = d.in_pts[i];
    const int m0 = (int)floorf(xfmul(bb[0], inv)), m1 = (int)floorf(xfmul(bb[1], inv)), m2 = (int)floorf(xfmul(bb[2], inv));
    const int d0 = (int)floorf(xfmul(bb[3], inv)) - m0 + 1, d1 = (int)floorf(xfmul(bb[4], inv)) - m1 + 1;
    const int i0 = (int)(floorf(xfmul(p.x, inv)) - (float)m0);
    const int i1 = (int)(floorf(xfmul(p.y, inv)) - (float)m1);
    const int i2 = (int)(floorf(xfmul(p.z, inv)) - (float)m2);
    key = (uint64_t)((long long)i0 + (long long)i1 * d0 + (long long)i2 * d0 * d1) & 0x7FFFFFFFull;
  }
  if (key >> key_bits) set_err(d, -4);  // the host derived key_bits from the same boxes
  if (kNarrow) reinterpret_cast<uint32_t*>(d.vkey)[i] = ((uint32_t)g << key_bits) | (uint32_t)key;
  else d.vkey[i] = ((uint64_t)g << key_bits) | key;
  d.vval[i] = (uint32_t)i;
}

// flag[p] = 1 where a run of equal keys starts; flag[n] = 0 so scan[n] = number of runs
template <typename K>
__global__ void head_flag_kernel(const K* __restrict__ keys, uint32_t* __restrict__ flag, int n) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  flag[p] = (p < n) && (p == 0 || keys[p] != keys[p - 1]);
}

// positions of the run heads, compacted (rank r -> first sorted position of voxel r)
__global__ void vox_heads_kernel(Dev d, int n) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n || !d.flag[p]) return;
  d.vval[d.scan[p]] = (uint32_t)p;  // vval (the unsorted value buffer) is free after the sort
}
// one thread per voxel: sequential float sums in (stable) sorted order, like pcl::VoxelGrid
__global__ void vox_centroid_kernel(Dev d, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_runs = (int)d.scan[n];
  if (r >= n_runs) return;
  const int p = (int)d.vval[r];
  const int e = r + 1 < n_runs ? (int)d.vval[r + 1] : n;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  for (int j = p; j < e; ++j) {
    const float4 q = d.in_pts[d.vval2[j]];
    sx = xfadd(sx, q.x); sy = xfadd(sy, q.y); sz = xfadd(sz, q.z); si = xfadd(si, q.w);
  }
  const float c = (float)(e - p);
  d.ds_pts[r] = make_float4(xfdiv(sx, c), xfdiv(sy, c), xfdiv(sz, c), xfdiv(si, c));
}

__global__ void ds_off_kernel(Dev d, int n) {
  const int g = threadIdx.x + blockIdx.x * blockDim.x;
  if (g > d.G) return;
  const int a = d.in_off[g];
  d.ds_off[g] = d.scan[a < n ? a : n];
  if (g < d.G) d.out[seg_slot(d, g)].n_ds[seg_cls(d, g)] = (int)d.scan[min(d.in_off[g + 1], n)] - (int)d.scan[min(a, n)];
}

// ----------------------------------------------------------------------------
// K3a: local map = the valid cubes of the store, in gather order (25 contiguous key ranges)
// ----------------------------------------------------------------------------
__global__ void range_kernel(Dev d, int cur) {
  const int g = blockIdx.x, c = threadIdx.x;
  const FrameDesc& fd = d.desc[seg_slot(d, g)];
  const uint64_t* keys = d.st_key[cur] + d.st_base[g];
  const int n = d.st_n[g];
  int lo = 0, len = 0, raw = 0;  // raw: points of the column's cubes that are still unfiltered (absorbed by this frame's merge)
  if (c < kCols && fd.active) {
    const int wi = fd.val_lo[0] + c / 5, wj = fd.val_lo[1] + c % 5;
    if (wi <= fd.val_hi[0] && wj <= fd.val_hi[1] && fd.val_lo[2] <= fd.val_hi[2]) {
      lo = lower_bound_u64(keys, n, store_key(pack_cube(wi, wj, fd.val_lo[2]), 0, 0));
      const int hi = lower_bound_u64(keys, n, store_key(pack_cube(wi, wj, fd.val_hi[2]) + 1, 0, 0));
      len = hi - lo;
      for (int wk = fd.val_lo[2]; wk <= fd.val_hi[2] && len > 0; ++wk) {
        const uint32_t cube = pack_cube(wi, wj, wk);
        raw += lower_bound_u64(keys + lo, len, store_key(cube + 1, 0, 0)) - lower_bound_u64(keys + lo, len, store_key(cube, 1, 0));
      }
    }
  }
  int incl = len;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (c >= o) incl += v;
  }
  for (int o = 16; o > 0; o >>= 1) raw += __shfl_xor_sync(0xffffffffu, raw, o);
  if (c == 0) d.lp_cnt[d.G + g] = raw;
  if (c < kCols) {
    d.rng_start[g * kCols + c] = lo;
    d.loc_off[g * (kCols + 1) + c] = incl - len;
  }
  if (c == kCols - 1) {
    d.loc_off[g * (kCols + 1) + kCols] = incl;
    d.lp_cnt[g] = incl;
    d.out[seg_slot(d, g)].n_local[seg_cls(d, g)] = incl;
  }
}

// store index of local index l of segment g
__device__ __forceinline__ int local_to_store(const Dev& d, int g, int l) {
  const int* lo = d.loc_off + g * (kCols + 1);
  int a = 0, b = kCols;  // largest col with lo[col] <= l
  while (b - a > 1) {
    int m = (a + b) >> 1;
    if (lo[m] <= l) a = m; else b = m;
  }
  return d.rng_start[g * kCols + a] + (l - lo[a]);
}

// ----------------------------------------------------------------------------
// K3: the persistent 1 m cell index of the local map (rows C, T: laserMapping.cpp:510-540, :559-560)
//
// The reference gathers the valid cubes and rebuilds two KD-trees every frame.  Here every segment keeps
// a voxel-hash style index alive across frames: an open-addressing table cell -> [cell:24][count:16]
// [bucket:24] and a pool of 4-entry buckets (64 B, chained) holding (x, y, z, tag) of every point of the
// local map.  It is built in bulk only when the valid block changes (launch_index_rebuild: the points are
// sorted by cell once, so a cell's buckets start out consecutive); between those frames the map update puts the few thousand points a frame re-centroids or adds
// into it (idx_apply, called by delta_reduce_kernel): O(changed points) per frame, no sort.
// tag = [cube of the valid block, gather order:7][pending:1][voxel z, y, x:3x8 | arrival rank:24]
// orders exactly like the position in the reference's gathered cloud, so the kNN tie rule (d2, index)
// is (d2, tag); the index itself is only materialised for the trace / debug outputs (tag_to_local).
// ----------------------------------------------------------------------------
constexpr unsigned long long kCellEmpty = ~0ull;
__device__ __forceinline__ uint32_t cell_key_of(const Dev& d, const FrameDesc& fd, float x, float y, float z) {
  const int rx = (int)floorf(x) - fd.origin[0], ry = (int)floorf(y) - fd.origin[1], rz = (int)floorf(z) - fd.origin[2];
  if ((unsigned)rx > 255u || (unsigned)ry > 255u || (unsigned)rz > 255u) set_err(d, -4);
  return ((uint32_t)(rz & 255) << 16) | ((uint32_t)(ry & 255) << 8) | (uint32_t)(rx & 255);
}
// tag of a store entry of segment g lying in the valid block of fd
__device__ __forceinline__ uint32_t cube_rel(const FrameDesc& fd, int ci, int cj, int ck) {
  return (uint32_t)(((ci - fd.val_lo[0]) * 5 + (cj - fd.val_lo[1])) * 3 + (ck - fd.val_lo[2]));
}
__device__ __forceinline__ uint32_t tag_filtered(const FrameDesc& fd, int ci, int cj, int ck, uint32_t vz, uint32_t vy, uint32_t vx) {
  return (cube_rel(fd, ci, cj, ck) << 25) | (vz << 16) | (vy << 8) | vx;
}
__device__ __forceinline__ int bucket_alloc(const Dev& d, int g) {
  const int b = atomicAdd(d.bcnt + g, 1);
  if (b >= d.bkt_off[g + 1] - d.bkt_off[g] || b >= (int)kNoBkt) { set_err(d, -3); return -1; }
  return b;
}
// entry number (index into d.bkt) of slot `slot` of the chain that starts at bucket b; grows the chain when asked
__device__ __forceinline__ long long chain_slot(const Dev& d, int g, int b, int slot, bool grow) {
  const int base = d.bkt_off[g];
  for (int c = slot / kBktE; c > 0; --c) {
    uint32_t nb = d.bnext[base + b];
    if (nb == kSentinel32) {
      if (!grow) return -1;
      const int fresh = bucket_alloc(d, g);
      if (fresh < 0) return -1;
      const uint32_t prev = atomicCAS(d.bnext + base + b, kSentinel32, (uint32_t)fresh);
      nb = prev == kSentinel32 ? (uint32_t)fresh : prev;  // (a lost race leaks one bucket until the next rebuild)
    }
    b = (int)nb;
  }
  return (long long)(base + b) * kBktE + slot % kBktE;
}
__device__ __forceinline__ void idx_insert(const Dev& d, int g, uint32_t k24, const float4 e) {
  unsigned long long* tab = d.hash_tab + d.hash_off[g];
  const uint32_t mask = (uint32_t)d.hmask[g];
  uint32_t s = cell_home(k24, mask);
  for (int guard = 0; guard <= (int)mask; ++guard) {
    unsigned long long old = tab[s];
    if (old == kCellEmpty) {
      const int b = bucket_alloc(d, g);
      if (b < 0) return;
      const unsigned long long entry = ((unsigned long long)k24 << 40) | kCellCount1 | (unsigned long long)b;
      old = atomicCAS(tab + s, kCellEmpty, entry);
      if (old == kCellEmpty) { d.bkt[(size_t)(d.bkt_off[g] + b) * kBktE] = e; return; }
    }
    if ((uint32_t)(old >> 40) == k24) {
      const unsigned long long prev = atomicAdd(tab + s, kCellCount1);
      const int slot = (int)((prev >> 24) & 0xFFFFu);
      if (slot >= 0xFFFF) { set_err(d, -3); return; }
      const long long at = chain_slot(d, g, (int)(prev & kNoBkt), slot, true);
      if (at >= 0) d.bkt[at] = e;
      return;
    }
    s = (s + 1) & mask;
  }
  set_err(d, -3);  // table full
}
// entry number of the point with this tag in cell k24, -1 if it is not there
__device__ __forceinline__ long long idx_find(const Dev& d, int g, uint32_t k24, uint32_t tag) {
  const unsigned long long* tab = d.hash_tab + d.hash_off[g];
  const uint32_t mask = (uint32_t)d.hmask[g];
  uint32_t s = cell_home(k24, mask);
  unsigned long long e = tab[s];
  while (e != kCellEmpty && (uint32_t)(e >> 40) != k24) { s = (s + 1) & mask; e = tab[s]; }
  if (e == kCellEmpty) return -1;
  const int cnt = (int)((e >> 24) & 0xFFFFu), base = d.bkt_off[g];
  int b = (int)(e & kNoBkt);
  for (int i = 0; i < cnt; ++i) {
    if (i && i % kBktE == 0) {
      const uint32_t nb = d.bnext[base + b];
      if (nb == kSentinel32) return -1;
      b = (int)nb;
    }
    const long long at = (long long)(base + b) * kBktE + i % kBktE;
    if (__float_as_uint(d.bkt[at].w) == tag) return at;
  }
  return -1;
}
// The map update re-centroided (had_old) or created the filtered entry `tag`: keep the index in step.
__device__ __forceinline__ void idx_apply(const Dev& d, int g, const FrameDesc& fd, uint32_t tag, bool had_old, const float4 o,
                                          const float4 c) {
  const float4 e = make_float4(c.x, c.y, c.z, __uint_as_float(tag));
  const uint32_t kn = cell_key_of(d, fd, c.x, c.y, c.z);
  if (had_old) {
    const uint32_t ko = cell_key_of(d, fd, o.x, o.y, o.z);
    const long long at = idx_find(d, g, ko, tag);
    if (at < 0) { set_err(d, -7); return; }
    if (ko == kn) { d.bkt[at] = e; return; }
    d.bkt[at] = make_float4(0.f, 0.f, 0.f, __uint_as_float(kSentinel32));  // the centroid left this cell
  }
  idx_insert(d, g, kn, e);
}

// ---- bulk (re)build of the segments listed in d.idx_list -----------------------------------------------
// grid (chunks, listed segments): clear the table slots that will be used and the buckets that were
__global__ void idx_reset_kernel(Dev d, const int* __restrict__ new_mask) {
  const int g = d.idx_list[blockIdx.y];
  unsigned long long* tab = d.hash_tab + d.hash_off[g];
  const int slots = new_mask[blockIdx.y] + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < slots; i += gridDim.x * blockDim.x) tab[i] = kCellEmpty;
  const int used = min(d.bcnt[g], d.bkt_off[g + 1] - d.bkt_off[g]);
  float4* bk = d.bkt + (size_t)d.bkt_off[g] * kBktE;
  const float4 none = make_float4(0.f, 0.f, 0.f, __uint_as_float(kSentinel32));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < used * kBktE; i += gridDim.x * blockDim.x) bk[i] = none;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < used; i += gridDim.x * blockDim.x) d.bnext[d.bkt_off[g] + i] = kSentinel32;
}
__global__ void idx_arm_kernel(Dev d, const int* __restrict__ new_mask, int n_seg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seg) return;
  const int g = d.idx_list[i];
  d.bcnt[g] = 0;
  d.hmask[g] = new_mask[i];
  if (d.shard_world > 1) d.shard_counts[g] = 0;
}
// The build sorts the points by cell once, so a cell's buckets are consecutive and need no atomics:
// 1. one thread per local-map point of the listed segments (packed by d.idx_poff): sort key [list position:7][cell:24]
__global__ void idx_key_kernel(Dev d, int cur, int n_seg, int total, uint32_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int g = -1;
  bool owned = false;  // sharded map: this rank owns the point (it counts towards the guard :555)
  if (i < total) {
    const int si = find_seg(d.idx_poff, n_seg, i);
    g = d.idx_list[si];
    const int l = i - d.idx_poff[si];
    uint32_t key = (uint32_t)n_seg << 24;  // (not a local point: behind every segment)
    if (l < d.loc_off[g * (kCols + 1) + kCols]) {
      const float4 p = d.st_pt[cur][d.st_base[g] + local_to_store(d, g, l)];
      key = ((uint32_t)si << 24) | cell_key_of(d, d.desc[seg_slot(d, g)], p.x, p.y, p.z);
      owned = d.shard_world > 1 && p.x >= d.shard_lo && p.x < d.shard_hi;
    }
    keys[i] = key;
    d.vval[i] = (uint32_t)l;
  }
  if (d.shard_world > 1) {  // one atomic per warp and segment instead of one per point
    const unsigned peers = __match_any_sync(0xffffffffu, owned ? g : -1);
    if (owned && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(d.shard_counts + g, __popc(peers));
  }
}
// 2. (after the sort) first sorted position of every listed segment
__global__ void idx_segs_kernel(Dev d, int n_seg, int total, const uint32_t* __restrict__ keys) {
  const int si = threadIdx.x + blockIdx.x * blockDim.x;
  if (si > n_seg) return;
  d.idx_soff[si] = lower_bound_u32(keys, total, (uint32_t)si << 24);
}
// 3. one thread per sorted point.  Cell number c of a segment (d.scan = cells that start before a position), whose
// first point is the segment's sorted point number s, owns the buckets from c + s / 4 on: consecutive cells never
// overlap ((s + n) / 4 >= s / 4 + n / 4) and at most one bucket per cell stays unused.  The last point of a cell
// publishes the cell in the table.
__global__ void idx_fill_kernel(Dev d, int cur, int n_seg, int total, const uint32_t* __restrict__ keys) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const uint32_t key = keys[p];
  const int si = (int)(key >> 24);
  if (si >= n_seg) return;
  const int g = d.idx_list[si], first = d.idx_soff[si];
  int p0 = p;
  while (p0 > first && keys[p0 - 1] == key) --p0;
  const int j = p - p0, pool = d.bkt_off[g + 1] - d.bkt_off[g];
  const int b0 = (int)(d.scan[p0] - d.scan[first]) + (p0 - first) / kBktE, b = b0 + j / kBktE;
  if (b >= pool || b >= (int)kNoBkt || j >= 0xFFFF) { set_err(d, -3); return; }
  const int pos = local_to_store(d, g, (int)d.vval2[p]);
  const uint64_t* skeys = d.st_key[cur] + d.st_base[g];
  const uint64_t skey = skeys[pos];
  const float4 pt = d.st_pt[cur][d.st_base[g] + pos];
  const FrameDesc& fd = d.desc[seg_slot(d, g)];
  int ci, cj, ck;
  unpack_cube(key_cube(skey), ci, cj, ck);
  uint32_t tag;
  if (key_pending(skey)) {  // raw point of a cube that has just become valid: rank among the cube's raw points
    const int pfirst = lower_bound_u64(skeys, d.st_n[g], store_key(key_cube(skey), 1, 0));
    tag = (cube_rel(fd, ci, cj, ck) << 25) | (1u << 24) | (uint32_t)min(pos - pfirst, 0xFFFFFF);
  } else {
    const uint64_t pl = key_payload(skey);
    const uint32_t vz = (uint32_t)(pl >> 22) & 0x7FFu, vy = (uint32_t)(pl >> 11) & 0x7FFu, vx = (uint32_t)pl & 0x7FFu;
    if ((vz | vy | vx) > 255u) set_err(d, -4);
    tag = tag_filtered(fd, ci, cj, ck, vz, vy, vx);
  }
  const int base = d.bkt_off[g];
  d.bkt[(size_t)(base + b) * kBktE + j % kBktE] = make_float4(pt.x, pt.y, pt.z, __uint_as_float(tag));
  if (j > 0 && j % kBktE == 0) d.bnext[base + b - 1] = (uint32_t)b;
  if (p + 1 < total && keys[p + 1] == key) return;
  // ---- last point of its cell ----
  if (p + 1 >= total || (int)(keys[p + 1] >> 24) != si) d.bcnt[g] = b + 1;  // ... and of its segment: buckets handed out
  unsigned long long* tab = d.hash_tab + d.hash_off[g];
  const uint32_t mask = (uint32_t)d.hmask[g];
  const unsigned long long entry = ((unsigned long long)(key & 0xFFFFFFu) << 40) | ((unsigned long long)(j + 1) << 24) | (unsigned long long)b0;
  uint32_t s = cell_home(key & 0xFFFFFFu, mask);
  for (uint32_t guard = 0; guard <= mask; ++guard) {
    if (atomicCAS(tab + s, kCellEmpty, entry) == kCellEmpty) return;
    s = (s + 1) & mask;
  }
  set_err(d, -3);  // table full
}
// position in the gathered local map (the reference's kNN index) of the entry with this tag: trace / debug only
__device__ __forceinline__ int tag_to_local(const Dev& d, int cur, int g, uint32_t tag) {
  const FrameDesc& fd = d.desc[seg_slot(d, g)];
  const int rel = (int)(tag >> 25);
  const int ci = fd.val_lo[0] + rel / 15, cj = fd.val_lo[1] + (rel / 3) % 5, ck = fd.val_lo[2] + rel % 3;
  const uint64_t* keys = d.st_key[cur] + d.st_base[g];
  const uint32_t cube = pack_cube(ci, cj, ck);
  int pos;
  if ((tag >> 24) & 1u) {
    pos = lower_bound_u64(keys, d.st_n[g], store_key(cube, 1, 0)) + (int)(tag & 0xFFFFFFu);
  } else {
    const uint64_t pl = ((uint64_t)((tag >> 16) & 255u) << 22) | ((uint64_t)((tag >> 8) & 255u) << 11) | (uint64_t)(tag & 255u);
    pos = lower_bound_u64(keys, d.st_n[g], store_key(cube, 0, pl));
  }
  const int col = (ci - fd.val_lo[0]) * 5 + (cj - fd.val_lo[1]);
  return d.loc_off[g * (kCols + 1) + col] + (pos - d.rng_start[g * kCols + col]);
}

// ----------------------------------------------------------------------------
// Row K: exact bounded kNN(5), thread per query
// ----------------------------------------------------------------------------
constexpr unsigned kFull = 0xffffffffu;
// Exact bounded kNN(5) of one query over the 27 cells around it.  Distances are
// the reference's float ((dx*dx)+(dy*dy))+(dz*dz); order is (d2, tag) == (d2, local index):
// both live in one 64-bit key (float bits of a non-negative d2 are monotonic).
struct Knn5 {
  unsigned long long key[5];  // d2 bits << 32 | tag, ascending
  uint32_t at[5];             // entry number in d.bkt
};
__device__ __forceinline__ float knn_d2(const Knn5& r, int k) { return __uint_as_float((uint32_t)(r.key[k] >> 32)); }
__device__ __forceinline__ uint32_t knn_tag(const Knn5& r, int k) { return (uint32_t)r.key[k]; }
// "none" marker: the search starts from the reference's gate (d2 = 1.0f, tag 0), see knn5_cells
constexpr unsigned long long kKnnInit = (unsigned long long)0x3F800000u << 32;

__device__ __forceinline__ void knn_offer(Knn5& r, float qx, float qy, float qz, const float4 c, uint32_t at) {
  const float dd = dist2(qx, qy, qz, c.x, c.y, c.z);
  const uint32_t tag = __float_as_uint(c.w);
  if (tag == kSentinel32 || dd > knn_d2(r, 4)) return;  // removed entry; cheap float test first, ties go through the exact 64-bit compare
  const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | tag;
  if (key < r.key[4]) {  // replace the 5th, then bubble it up (compare-exchange chain)
    r.key[4] = key;
    r.at[4] = at;
#pragma unroll
    for (int i = 4; i > 0; --i) {
      const unsigned long long lo = r.key[i - 1], hi = r.key[i];
      const uint32_t alo = r.at[i - 1], ahi = r.at[i];
      const bool sw = hi < lo;
      r.key[i - 1] = sw ? hi : lo;
      r.key[i] = sw ? lo : hi;
      r.at[i - 1] = sw ? ahi : alo;
      r.at[i] = sw ? alo : ahi;
    }
  }
}
// all points of one cell: `cnt` entries of the bucket chain that starts at bucket b, four at a time (four
// independent 16-byte loads of one 64-byte line); the pointer to the next bucket is followed once per kBktE entries
__device__ __forceinline__ void knn_scan_cell(const Dev& d, int g, int b, int cnt, float qx, float qy, float qz, Knn5& r) {
  const int base = d.bkt_off[g];
  int sub = 0;  // entry offset inside the bucket
#pragma unroll 1
  while (cnt > 0) {
    const uint32_t at0 = (uint32_t)(base + b) * kBktE + (uint32_t)sub;
    const float4* __restrict__ p = d.bkt + at0;
#if S2M_KNN_PRED
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 c0 = __ldg(p), c1 = cnt > 1 ? __ldg(p + 1) : zero4, c2 = cnt > 2 ? __ldg(p + 2) : zero4, c3 = cnt > 3 ? __ldg(p + 3) : zero4;
#else
    const float4 c0 = __ldg(p), c1 = __ldg(p + 1), c2 = __ldg(p + 2), c3 = __ldg(p + 3);
#endif
    uint32_t nb = (uint32_t)b;
    if (cnt > 4 && sub + 4 == kBktE) nb = __ldg(d.bnext + base + b);
    knn_offer(r, qx, qy, qz, c0, at0);
    if (cnt > 1) knn_offer(r, qx, qy, qz, c1, at0 + 1);
    if (cnt > 2) knn_offer(r, qx, qy, qz, c2, at0 + 2);
    if (cnt > 3) knn_offer(r, qx, qy, qz, c3, at0 + 3);
    cnt -= 4;
    sub = (sub + 4) % kBktE;
    if (nb == kSentinel32) break;
    b = (int)nb;
  }
}

// per-thread staging of the 27 cell probes (one column per lane of a warp): count << 24 | bucket, 0 count = empty
struct KnnStage {
  uint32_t cell[27][32];
};

// Rows (dy,dz) of three x-adjacent cells are visited near to far; a row is skipped when
// a lower bound of the FLOAT distance to any point in it exceeds the current 5th
// distance.  The bound is built with the same rounding steps as dist2() ((0 + by*by) +
// bz*bz with by, bz the exact distances to the row's boundary planes) and rounding is
// monotonic, so no point that could enter the result is ever skipped.
// Returns the number of map points in the 27 cells (all nine rows, pruned or not).
__device__ __forceinline__ int knn5_cells(const Dev& d, int g, const int origin[3], float qx, float qy, float qz,
                                          Knn5& r, KnnStage& st) {
  // The result is only used when the 5th distance is < 1.0 (laserMapping.cpp:585, :653), so the
  // search starts from that bound: key (1.0f, tag 0) is larger than every (d2 < 1, any tag)
  // and not larger than any (d2 >= 1, .) -- candidates at 1 m or more never enter.
#pragma unroll
  for (int k = 0; k < 5; ++k) { r.key[k] = kKnnInit; r.at[k] = 0u; }
  const float fly = floorf(qy), flz = floorf(qz);
  const int cx = (int)floorf(qx) - origin[0], cy = (int)fly - origin[1], cz = (int)flz - origin[2];
  const uint32_t mask = (uint32_t)d.hmask[g];
  const unsigned long long* __restrict__ tab = d.hash_tab + d.hash_off[g];
  const int t = threadIdx.x & 31;
  // exact distances from the query to the cell's boundary planes (fractional parts are exact)
  const float fy = xfsub(qy, fly), fz = xfsub(qz, flz);
  const float gy = xfsub(1.0f, fy), gz = xfsub(1.0f, fz);
  const int sy = fy < 0.5f ? -1 : 1, sz = fz < 0.5f ? -1 : 1;  // side of the nearer boundary
  int total = 0;
#if S2M_KNN_BATCH9
#pragma unroll 1
  for (int ob = 0; ob < 9; ob += 3) {  // phase 1: the 27 cell probes, nine independent loads (three rows) at a time
    uint32_t k24[9], sl[9];
    unsigned long long e[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      const int o = ob + j / 3, a = j % 3;
      const int dy = (o == 1 || o == 3 || o == 7) ? sy : ((o == 4 || o == 6 || o == 8) ? -sy : 0);
      const int dz = (o == 2 || o == 3 || o == 6) ? sz : ((o == 5 || o == 7 || o == 8) ? -sz : 0);
      const int z = cz + dz, y = cy + dy, x = cx + a - 1;
      const bool ok = (unsigned)z <= 255u && (unsigned)y <= 255u && (unsigned)x <= 255u;
      k24[j] = ((uint32_t)(z & 255) << 16) | ((uint32_t)(y & 255) << 8) | (uint32_t)(x & 255);
      sl[j] = cell_home(k24[j], mask);
      e[j] = ok ? tab[sl[j]] : kCellEmpty;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
      while (e[j] != kCellEmpty && (uint32_t)(e[j] >> 40) != k24[j]) {
        sl[j] = (sl[j] + 1) & mask;
        e[j] = tab[sl[j]];
      }
      uint32_t rec = 0u;
      if (e[j] != kCellEmpty) {
        const uint32_t c = (uint32_t)(e[j] >> 24) & 0xFFFFu;
        total += (int)c;
        rec = (min(c, 255u) << 24) | (uint32_t)(e[j] & kNoBkt);
      }
      st.cell[3 * ob + j][t] = rec;
    }
  }
#else
#pragma unroll 1
  for (int o = 0; o < 9; ++o) {  // phase 1: the 27 cell probes, three rows (nine independent loads) at a time
    // row order: centre, (sy,0), (0,sz), (sy,sz), (-sy,0), (0,-sz), (-sy,sz), (sy,-sz), (-sy,-sz)
    const int dy = (o == 1 || o == 3 || o == 7) ? sy : ((o == 4 || o == 6 || o == 8) ? -sy : 0);
    const int dz = (o == 2 || o == 3 || o == 6) ? sz : ((o == 5 || o == 7 || o == 8) ? -sz : 0);
    const int z = cz + dz, y = cy + dy;
    const bool row_ok = (unsigned)z <= 255u && (unsigned)y <= 255u;
    uint32_t k24[3], sl[3];
    unsigned long long e[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int x = cx + a - 1;
      k24[a] = ((uint32_t)(z & 255) << 16) | ((uint32_t)(y & 255) << 8) | (uint32_t)(x & 255);
      sl[a] = cell_home(k24[a], mask);
      e[a] = (row_ok && (unsigned)x <= 255u) ? tab[sl[a]] : kCellEmpty;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      while (e[a] != kCellEmpty && (uint32_t)(e[a] >> 40) != k24[a]) {
        sl[a] = (sl[a] + 1) & mask;
        e[a] = tab[sl[a]];
      }
      uint32_t rec = 0u;
      if (e[a] != kCellEmpty) {
        const uint32_t c = (uint32_t)(e[a] >> 24) & 0xFFFFu;
        total += (int)c;
        rec = (min(c, 255u) << 24) | (uint32_t)(e[a] & kNoBkt);
      }
      st.cell[3 * o + a][t] = rec;
    }
  }
#endif
  // phase 2: rows near to far
#pragma unroll 1
  for (int o = 0; o < 9; ++o) {
    const uint32_t r0 = st.cell[3 * o][t], r1 = st.cell[3 * o + 1][t], r2 = st.cell[3 * o + 2][t];
    if ((r0 | r1 | r2) >> 24 == 0u) continue;
    const float by = (o == 0 || o == 2 || o == 5) ? 0.0f : ((o == 1 || o == 3 || o == 7) ? (sy < 0 ? fy : gy) : (sy < 0 ? gy : fy));
    const float bz = (o == 0 || o == 1 || o == 4) ? 0.0f : ((o == 2 || o == 3 || o == 6) ? (sz < 0 ? fz : gz) : (sz < 0 ? gz : fz));
    if (xfadd(xfmul(by, by), xfmul(bz, bz)) > knn_d2(r, 4)) continue;  // strict: a tie at the 5th distance may still win on the tag
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const uint32_t rec = a == 0 ? r1 : (a == 1 ? r0 : r2);  // the query's own column first
      int cnt = (int)(rec >> 24);
      if (cnt == 0) continue;
      if (cnt == 255) {  // saturated staging field: the exact count is in the table
        const int dy = (o == 1 || o == 3 || o == 7) ? sy : ((o == 4 || o == 6 || o == 8) ? -sy : 0);
        const int dz = (o == 2 || o == 3 || o == 6) ? sz : ((o == 5 || o == 7 || o == 8) ? -sz : 0);
        const int x = cx + (a == 0 ? 0 : (a == 1 ? -1 : 1));
        const uint32_t kk = ((uint32_t)((cz + dz) & 255) << 16) | ((uint32_t)((cy + dy) & 255) << 8) | (uint32_t)(x & 255);
        uint32_t s = cell_home(kk, mask);
        unsigned long long ee = tab[s];
        while (ee != kCellEmpty && (uint32_t)(ee >> 40) != kk) { s = (s + 1) & mask; ee = tab[s]; }
        cnt = ee == kCellEmpty ? 0 : (int)((ee >> 24) & 0xFFFFu);
      }
      knn_scan_cell(d, g, (int)(rec & kNoBkt), cnt, qx, qy, qz, r);
    }
  }
  return total;
}

// ----------------------------------------------------------------------------
// Block-level accumulation.  Each thread adds the 28 sums of its own queries into
// its column of a shared array; the block reduces the columns once, in a fixed
// tree (deterministic), and writes one partial per block.  The block that
// finishes last for a slot (ticket counter) sums the partials in block order and
// runs the LM step -- no separate reduction / solver launch.
// ----------------------------------------------------------------------------
constexpr int kAccCols = kTile / 4;  // four neighbouring lanes share one accumulator column
struct BlockAcc {
  double v[kPartial][kAccCols];
};
__device__ __forceinline__ void acc_zero(BlockAcc& A) {
  if (threadIdx.x < kAccCols) {
#pragma unroll
    for (int k = 0; k < kPartial; ++k) A.v[k][threadIdx.x] = 0.0;
  }
}
// every lane of the warp must call this (zeros where it has nothing to add)
__device__ __forceinline__ void acc_add(BlockAcc& A, const Sums28& S, double n_edge, double n_plane, double cand_c,
                                        double cand_s) {
  const int col = threadIdx.x >> 2;
  const bool lead = (threadIdx.x & 3) == 0;
#pragma unroll
  for (int k = 0; k < kPartial; ++k) {
    double v = k < 28 ? S.v[k < 28 ? k : 0] : (k == 28 ? n_edge : (k == 29 ? n_plane : (k == 30 ? cand_c : cand_s)));
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (lead) A.v[k][col] += v;
  }
}
__device__ __forceinline__ void acc_store(BlockAcc& A, double* __restrict__ dst) {
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int k = w; k < kPartial; k += kTile / 32) {
    double v = A.v[k][l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (l == 0) dst[k] = v;
  }
}
// true in exactly one block per (launch, slot): the last of `nwork` blocks to get here
__device__ __forceinline__ bool block_is_last(int* ticket, int nwork) {
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    last = (t == nwork - 1);
    if (last) *ticket = 0;  // self-cleaning for the next launch
  }
  __syncthreads();
  return last != 0;
}
__device__ __forceinline__ void sum_partials(const Dev& d, int slot, int nwork, double* sm /*kPartial, shared*/) {
  if (threadIdx.x < kPartial) {
    const double* p = d.partials + (size_t)slot * d.max_tiles * kPartial;
    double v = 0.0;
    for (int b = 0; b < nwork; ++b) v += __ldcg(p + (size_t)b * kPartial + threadIdx.x);  // fixed order
    sm[threadIdx.x] = v;
  }
  __syncthreads();
}
// The LM state lives in global memory; the serial solver step runs on a shared-memory
// copy (one coalesced load/store by the whole block instead of ~100 dependent L2 round trips).
static_assert(sizeof(LmState) % 8 == 0, "LmState is copied as 8-byte words");
__device__ __forceinline__ void lm_load(LmState* sh, const LmState* gl) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gl);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(sh);
  for (int i = threadIdx.x; i < (int)(sizeof(LmState) / 8); i += blockDim.x) dst[i] = __ldcg(src + i);
  __syncthreads();
}
__device__ __forceinline__ void lm_store(LmState* gl, const LmState* sh) {
  __syncthreads();
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(sh);
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(gl);
  for (int i = threadIdx.x; i < (int)(sizeof(LmState) / 8); i += blockDim.x) dst[i] = src[i];
}
__device__ __forceinline__ void slot_counts(const Dev& d, int slot, int& dc0, int& nc, int& ds0, int& nq) {
  dc0 = d.ds_off[slot]; nc = d.ds_off[slot + 1] - dc0;
  ds0 = d.ds_off[d.B + slot]; nq = nc + d.ds_off[d.B + slot + 1] - ds0;
}

// ----------------------------------------------------------------------------
// K4: fused association.  One thread per down-sampled scan point; a block walks
// tiles of kTile points of one slot (grid.x blocks per slot, grid.y = slots).
// ----------------------------------------------------------------------------
// LM tails: run by the last block of a slot (unsharded) or by lm_shard_kernel after the
// allreduce of the 32 sums (sharded map).  `red` = reduced sums in shared memory.
__device__ __forceinline__ void lm_tail_begin(const Dev& d, int slot, int outer, const double* red, LmState* Ls) {
  lm_load(Ls, d.lm + slot);
  if (threadIdx.x == 0) {
    Sums28 S;
    for (int i = 0; i < 28; ++i) S.v[i] = red[i];
    SlotOut& o = d.out[slot];
    o.n_edge[outer] = (int)red[28]; o.n_plane[outer] = (int)red[29];
    if (outer == 0) { o.cand[0] = red[30]; o.cand[1] = red[31]; }
    double x0[7];
    for (int i = 0; i < 7; ++i) x0[i] = Ls->x[i];
    lm_begin(*Ls, x0, S, (int)red[28] + (int)red[29], 4);
    o.lm_iters[outer] = Ls->iteration; o.lm_term[outer] = Ls->termination;
    o.cost_initial[outer] = Ls->initial_cost; o.cost_final[outer] = Ls->final_cost;
  }
  lm_store(d.lm + slot, Ls);
}
__device__ __forceinline__ void lm_tail_after(const Dev& d, int slot, int outer, const double* red, LmState* Ls) {
  lm_load(Ls, d.lm + slot);
  if (threadIdx.x == 0) {
    Sums28 S;
    for (int i = 0; i < 28; ++i) S.v[i] = red[i];
    lm_after_eval(*Ls, S, 4);
    SlotOut& o = d.out[slot];
    o.lm_iters[outer] = Ls->iteration; o.lm_term[outer] = Ls->termination;
    o.cost_final[outer] = Ls->final_cost;
  }
  lm_store(d.lm + slot, Ls);
}
// sharded map: the per-rank sums wait in d.shard_sums for the allreduce
__device__ __forceinline__ void shard_publish(const Dev& d, int slot, const double* red) {
  if (threadIdx.x < kPartial) d.shard_sums[(size_t)slot * kPartial + threadIdx.x] = red[threadIdx.x];
}
__global__ void lm_shard_kernel(Dev d, int outer, int after) {
  const int slot = blockIdx.x;
  if (!d.out[slot].optimized) return;
  __shared__ double red[kPartial];
  __shared__ LmState Ls;
  if (after) {
    const LmState& L = d.lm[slot];
    if (L.done || !L.have_candidate) return;
  }
  if (threadIdx.x < kPartial) red[threadIdx.x] = d.shard_sums[(size_t)slot * kPartial + threadIdx.x];
  __syncthreads();
  if (after) lm_tail_after(d, slot, outer, red, &Ls);
  else lm_tail_begin(d, slot, outer, red, &Ls);
}

// ----------------------------------------------------------------------------
// K4 = the association of one outer iteration (laserMapping.cpp:578-706), two back-to-back kernels:
//   K4a knn_kernel : pointAssociateToMap (FP64, exact) + the exact bounded kNN(5) + the 1 m gate
//                    (rows P, T, K).  Integer / float only, <= 64 registers, 8 blocks per SM.  Warps
//                    take units of 32 consecutive scan points of a slot from a device-wide ticket.
//   K4b fit_kernel : thread per gated point: edge PCA / plane QR (FP64), residual + tangent-space
//                    Jacobian + Huber, 48-byte correspondence record (rows E, F, R, L, Q); the 28
//                    sums of every 32-point unit are transposed through shared memory and added in
//                    lane order: one 32-double partial per unit, no block barrier anywhere.
// The hand-off is 24 bytes per point (gate + n, five entry numbers of the cell index).  The solver
// (solve_kernel) adds the unit partials of a slot in unit order, so every bit of the pose is
// independent of what else shares the launch.
// ----------------------------------------------------------------------------
template <bool kTrace>
__global__ void __launch_bounds__(kTile, S2M_K4A_MINB) knn_kernel(Dev d, int outer) {
  // Work unit = 32 consecutive queries of one slot, taken by a WARP from a device-wide ticket:
  // the trip counts of the search vary a lot between queries, so static tiles leave most of a
  // block (and the tail of the grid) idle. The results do not depend on who computes them.
  __shared__ KnnStage stage_w[kTile / 32];
  KnnStage& stage = stage_w[threadIdx.x >> 5];
  __shared__ int chunk_off[kMaxBatch + 1];
  const int B = d.B, lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {  // chunks per slot (none for a slot that is not optimised this frame), prefix-summed
    int c[2], incl[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int sl = 32 * h + lane;
      c[h] = 0;
      if (sl < B && d.out[sl].optimized) {
        int dc0, nc, ds0, nq;
        slot_counts(d, sl, dc0, nc, ds0, nq);
        c[h] = (nq + 31) >> 5;
      }
      incl[h] = c[h];
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl[h], o);
        if (lane >= o) incl[h] += v;
      }
    }
    const int first = __shfl_sync(0xffffffffu, incl[0], 31);
    if (lane == 0) chunk_off[0] = 0;
    chunk_off[1 + lane] = incl[0];
    if (32 + lane < kMaxBatch) chunk_off[33 + lane] = first + incl[1];
  }
  __syncthreads();
  const int total = chunk_off[B];
  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(d.knn_ticket, 1);
    chunk = __shfl_sync(0xffffffffu, chunk, 0);
    if (chunk >= total) break;
    int slot = 0, hi = B;  // chunk_off[slot] <= chunk < chunk_off[hi]
    while (hi - slot > 1) {
      const int mid = (slot + hi) >> 1;
      if (chunk_off[mid] <= chunk) slot = mid; else hi = mid;
    }
    int dc0, nc, ds0, nq;
    slot_counts(d, slot, dc0, nc, ds0, nq);
    const int q = ((chunk - chunk_off[slot]) << 5) + lane;
    int visited = 0, cls = 0;
    if (q < nq) {
      cls = q >= nc;
      const int pos_q = cls ? ds0 + (q - nc) : dc0 + q;  // position in the packed query list
      const float4 p = d.ds_pts[pos_q];
      float w[3];
      xf_point(d.lm[slot].x, p.x, p.y, p.z, w);
      if (d.shard_world > 1 && !(w[0] >= d.shard_lo && w[0] < d.shard_hi)) {
        // sharded map: this query is answered by the rank whose x-slab holds it
        d.nbr[6 * (size_t)pos_q] = 0;
      } else {
        const int origin[3] = {d.desc[slot].origin[0], d.desc[slot].origin[1], d.desc[slot].origin[2]};
        Knn5 r;
        visited = knn5_cells(d, cls ? B + slot : slot, origin, w[0], w[1], w[2], r, stage);
        const bool gate = knn_d2(r, 4) < 1.0f;  // laserMapping.cpp:585 / :653
        int* nb = d.nbr + 6 * (size_t)pos_q;
        nb[0] = (visited << 1) | (gate ? 1 : 0);  // visited = map points in the query's 27 cells
        if (gate) {  // entry numbers in d.bkt of the five neighbours, nearest first
#pragma unroll
          for (int kk = 0; kk < 5; ++kk) nb[1 + kk] = (int)r.at[kk];
        }
      }
    }
  }
}

// ----------------------------------------------------------------------------
// K4a, grouped form (the default): queries that share a 2 m block of cells search together.
//
// The scan points are bucketed once per frame by (segment, 2 m block of the cell under the initial guess)
// (qgroup_kernel), so the 32 queries a warp takes contain whole blocks.  Lanes whose ACTUAL cells (this outer
// iteration's pose) lie in the same block form a group (match.any); the group's candidate set is every map
// point in the box of cells [lowest cell - 1, highest cell + 1] (3..4 cells per axis), a superset of each
// member's 27 cells -- points outside a query's 27 cells are at least 1 m away in the reference's float
// arithmetic and can never pass the d2[4] < 1 gate, so the result is unchanged.
//   1. the warp probes the cells of all its groups' boxes, four (group, cell) items per lane and step, each
//      probe one 16-byte load of the first two slots of the cell's probe sequence; occupied cells go to a hit
//      list, their point counts to per-group totals (shared-memory atomics);
//   2. the groups' candidate ranges are laid out in the warp's shared-memory pool (prefix sum; groups that
//      do not fit wait for the next round) and the lanes copy the hit cells' bucket chains into it;
//   3. every lane runs over ITS GROUP's candidates -- one broadcast 16-byte shared load per candidate, the
//      reference's float distance, its low bits replaced by the candidate's position in the range, and a
//      branch-free six-deep min/max network on those keys (seeded with the gate, 1.0f);
//   4. if the fifth and sixth key differ above the position bits the five kept candidates are the five nearest
//      and are put into the reference's (d2, index) order as 64-bit keys (exact d2 bits, tag); otherwise the
//      lane runs over the range once more and orders everything up to the fifth key exactly.
// No loop depends on a lane's own data except through its group, so the lanes of a warp stay together.
// Lanes whose group is too large for the pool (or a warp with more occupied cells than the hit list holds)
// fall back to the thread-per-query search (knn5_cells); d.knn_stats counts them.
// ----------------------------------------------------------------------------
#ifndef S2M_KNN_POOL
#define S2M_KNN_POOL 512
#endif
constexpr int kGPool = S2M_KNN_POOL;  // candidates a warp stages per round (16 bytes each)
constexpr int kGHits = 256;           // occupied cells a warp may stage per chunk
constexpr int kGIdxBits = 10;         // a candidate's position in its group's range rides in the low bits of its distance
static_assert(kGPool <= (1 << kGIdxBits) && kGPool * sizeof(float4) >= sizeof(KnnStage), "positions are 10 bits; the fallback stages in the pool");
struct KnnGroupSmem {
  float4 pool[kGPool];        // (x, y, z, entry number in d.bkt) of the staged candidates
  uint2 hits[kGHits];         // .x = first bucket | group << 24, .y = count << 16 | offset in the group's range
  int gtot[32];               // per group: map points in its box
  int gstart[32];             // first pool entry of the group this round, -1: not in this round
  int gibase[33];             // first probe item of the group
  int4 gbox[32];              // lowest cell of the box (x, y, z), dims dx | dy << 8 | dz << 16
  int gtab[32], gmask[32], gbkt[32];  // hash_off, hmask, bkt_off of the group's segment
};

// Block key of a cell (relative to the valid block's origin): cells -1..256 -> blocks 0..129
__device__ __forceinline__ bool cell_near_block(int cx, int cy, int cz) {
  return (unsigned)(cx + 1) <= 257u && (unsigned)(cy + 1) <= 257u && (unsigned)(cz + 1) <= 257u;
}
__device__ __forceinline__ uint32_t block_key(int g, int cx, int cy, int cz) {
  return ((uint32_t)g << 24) | ((uint32_t)((cz + 2) >> 1) << 16) | ((uint32_t)((cy + 2) >> 1) << 8) | (uint32_t)((cx + 2) >> 1);
}
// Once per frame: the scan points of every segment bucketed by the 2 m block they fall into under the pose the frame
// starts from (a counting sort over a hash of the block, one thread block per segment, histogram and cursors in
// shared memory).  All the grouped search needs is that the points of a block end up next to each other -- the
// order inside a bucket is whatever the atomics make it, and no result depends on it.
constexpr int kQBins = 2048;
__global__ void __launch_bounds__(1024) qgroup_kernel(Dev d) {
  __shared__ int hist[kQBins];
  __shared__ int wsum[32];
  const int g = blockIdx.x, slot = seg_slot(d, g), t = threadIdx.x;
  const int base = d.ds_off[g], n = d.ds_off[g + 1] - base;
  if (n <= 0) return;
  if (!d.out[slot].optimized) {  // not searched this frame
    for (int i = t; i < n; i += 1024) d.qs_key2[base + i] = kSentinel32;
    return;
  }
  for (int i = t; i < kQBins; i += 1024) hist[i] = 0;
  __syncthreads();
  double pose[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = d.lm[slot].x[i];
  const int o0 = d.desc[slot].origin[0], o1 = d.desc[slot].origin[1], o2 = d.desc[slot].origin[2];
  for (int i = t; i < n; i += 1024) {
    const float4 p = d.ds_pts[base + i];
    float w[3];
    xf_point(pose, p.x, p.y, p.z, w);
    const float fx = floorf(w[0]) - (float)o0, fy = floorf(w[1]) - (float)o1, fz = floorf(w[2]) - (float)o2;
    uint32_t key = ((uint32_t)g << 24) | 0xFFFFFFu;
    if (fabsf(fx) < 1e6f && fabsf(fy) < 1e6f && fabsf(fz) < 1e6f && cell_near_block((int)fx, (int)fy, (int)fz))
      key = block_key(g, (int)fx, (int)fy, (int)fz);
    d.qs_key[base + i] = key;
    atomicAdd(&hist[cell_hash(key) & (kQBins - 1)], 1);
  }
  __syncthreads();
  {  // exclusive prefix sum of the histogram: two bins per thread
    const int a = hist[2 * t], b = hist[2 * t + 1];
    int v = a + b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if ((t & 31) >= o) v += u;
    }
    if ((t & 31) == 31) wsum[t >> 5] = v;
    __syncthreads();
    if (t < 32) {
      int s = wsum[t];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, s, o);
        if (t >= o) s += u;
      }
      wsum[t] = s;
    }
    __syncthreads();
    const int excl = v - (a + b) + (t >= 32 ? wsum[(t >> 5) - 1] : 0);
    hist[2 * t] = excl;
    hist[2 * t + 1] = excl + a;
  }
  __syncthreads();
  for (int i = t; i < n; i += 1024) {
    const uint32_t key = d.qs_key[base + i];
    const int pos = atomicAdd(&hist[cell_hash(key) & (kQBins - 1)], 1);
    d.qs_key2[base + pos] = key;
    d.qs_val2[base + pos] = (uint32_t)(base + i);
  }
}
// profiling only: map points in the 27 cells of every query (the C-bar of SURVEY 8d's byte formula)
__global__ void count27_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = find_seg(d.ds_off, d.G, i), slot = seg_slot(d, g);
  int total = 0;
  if (d.out[slot].optimized) {
    const float4 p = d.ds_pts[i];
    float w[3];
    xf_point(d.lm[slot].x, p.x, p.y, p.z, w);
    const int* o = d.desc[slot].origin;
    const float fx = floorf(w[0]) - (float)o[0], fy = floorf(w[1]) - (float)o[1], fz = floorf(w[2]) - (float)o[2];
    if (fabsf(fx) < 1e6f && fabsf(fy) < 1e6f && fabsf(fz) < 1e6f && cell_near_block((int)fx, (int)fy, (int)fz) &&
        !(d.shard_world > 1 && !(w[0] >= d.shard_lo && w[0] < d.shard_hi))) {
      const unsigned long long* __restrict__ tab = d.hash_tab + d.hash_off[g];
      const uint32_t mask = (uint32_t)d.hmask[g];
      for (int j = 0; j < 27; ++j) {
        const int x = (int)fx + j % 3 - 1, y = (int)fy + (j / 3) % 3 - 1, z = (int)fz + j / 9 - 1;
        if ((unsigned)x > 255u || (unsigned)y > 255u || (unsigned)z > 255u) continue;
        const uint32_t k24 = ((uint32_t)z << 16) | ((uint32_t)y << 8) | (uint32_t)x;
        uint32_t sl = cell_home(k24, mask);
        unsigned long long e = tab[sl];
        while (e != kCellEmpty && (uint32_t)(e >> 40) != k24) { sl = (sl + 1) & mask; e = tab[sl]; }
        if (e != kCellEmpty) total += (int)((e >> 24) & 0xFFFFu);
      }
    }
  }
  d.cand27[i] = total;
}

// The grouped search of one warp: lane = one query (valid or not) at world point (qx, qy, qz) of segment g whose
// cell relative to the valid block's origin is (cx, cy, cz).  Returns the gate (laserMapping.cpp:585 / :653) and,
// if it holds, the five neighbours in r.  Every lane of the warp must call it.
__device__ __forceinline__ bool knn5_group(const Dev& d, KnnGroupSmem& S, bool valid, int g, const int origin[3], int cx, int cy,
                                           int cz, float qx, float qy, float qz, Knn5& r) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  // ---- groups
  const uint32_t bkey = valid ? block_key(g, cx, cy, cz) : (0x80000000u | (uint32_t)lane);
  const unsigned gm = __match_any_sync(kFull, bkey);
  const int leader = __ffs(gm) - 1;
  const bool is_leader = lane == leader;
  const unsigned leaders = __ballot_sync(kFull, is_leader);
  const int gnum = __popc(leaders & ((1u << leader) - 1u)), ng = __popc(leaders);
  // The members' cells are the block's low (2k-2) or high (2k-1) cell per axis: the box is one cell wider on
  // each side of the cells that occur (3 or 4 cells per axis).
  const int px = (cx + 2) & 1, py = (cy + 2) & 1, pz = (cz + 2) & 1;
  const unsigned ox = __ballot_sync(kFull, px != 0) & gm, oy = __ballot_sync(kFull, py != 0) & gm, oz = __ballot_sync(kFull, pz != 0) & gm;
  const int lox = cx - px - (ox != gm ? 1 : 0), loy = cy - py - (oy != gm ? 1 : 0), loz = cz - pz - (oz != gm ? 1 : 0);
  const int dx = 3 + ((ox != gm && ox != 0u) ? 1 : 0), dy = 3 + ((oy != gm && oy != 0u) ? 1 : 0), dz = 3 + ((oz != gm && oz != 0u) ? 1 : 0);
  const int ncell = valid ? dx * dy * dz : 0;
  int incl = is_leader ? ncell : 0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += v;
  }
  const int items = __shfl_sync(kFull, incl, 31);
  if (is_leader) {
    S.gibase[gnum] = incl - ncell;
    S.gbox[gnum] = make_int4(lox, loy, loz, dx | (dy << 8) | (dz << 16));
    S.gtot[gnum] = 0;
    S.gtab[gnum] = valid ? d.hash_off[g] : 0;
    S.gmask[gnum] = valid ? d.hmask[g] : 0;
    S.gbkt[gnum] = valid ? d.bkt_off[g] : 0;
  }
  if (lane == 0) S.gibase[ng] = items;
  __syncwarp();
  // ---- 1. probes: item = (group, cell of its box); every lane takes four items per step, their table loads (the
  // first two slots of each probe sequence in one 16-byte load) in flight together
  int nh = 0, gp = 0;
#pragma unroll 1
  for (int it0 = 0; it0 < items; it0 += 128) {
    ulonglong2 e2[4];
    uint32_t k24[4], sl[4];
    int gq[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int it = it0 + 32 * j + lane;
      e2[j] = make_ulonglong2(kCellEmpty, kCellEmpty); k24[j] = 0u; sl[j] = 0u;
      gq[j] = gp;
      if (it0 + 32 * j >= items) continue;  // (whole warp)
      if (it < items) {
        while (S.gibase[gp + 1] <= it) ++gp;
        const int ci = it - S.gibase[gp];
        const int4 box = S.gbox[gp];
        const int bx = box.w & 255, by = (box.w >> 8) & 255;
        // ci = (iz * by + iy) * bx + ix with bx, by in {3, 4}
        const int row = bx == 4 ? ci >> 2 : (ci * 43) >> 7;
        const int ix = ci - row * bx;
        const int iz = by == 4 ? row >> 2 : (row * 43) >> 7;
        const int iy = row - iz * by;
        const int x = box.x + ix, y = box.y + iy, z = box.z + iz;
        if ((unsigned)x <= 255u && (unsigned)y <= 255u && (unsigned)z <= 255u) {
          k24[j] = ((uint32_t)z << 16) | ((uint32_t)y << 8) | (uint32_t)x;
          sl[j] = cell_home(k24[j], (uint32_t)S.gmask[gp]);
          e2[j] = *reinterpret_cast<const ulonglong2*>(d.hash_tab + S.gtab[gp] + sl[j]);
        }
      }
      gq[j] = gp;
    }
    unsigned long long e[4];
    bool open[4];  // neither of the two slots settled the probe
    bool any_open = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool m0 = e2[j].x == kCellEmpty || (uint32_t)(e2[j].x >> 40) == k24[j];
      const bool m1 = e2[j].y == kCellEmpty || (uint32_t)(e2[j].y >> 40) == k24[j];
      e[j] = m0 ? e2[j].x : e2[j].y;
      open[j] = !m0 && !m1;
      any_open = any_open || open[j];
    }
    if (__any_sync(kFull, any_open)) {  // (both slots taken by other cells: linear probing, the four sequences together)
      for (;;) {
        bool again = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (open[j]) {
            const uint32_t mask = (uint32_t)S.gmask[gq[j]];
            sl[j] = (sl[j] + 2) & mask;
            e2[j] = *reinterpret_cast<const ulonglong2*>(d.hash_tab + S.gtab[gq[j]] + sl[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (open[j]) {
            const bool m0 = e2[j].x == kCellEmpty || (uint32_t)(e2[j].x >> 40) == k24[j];
            const bool m1 = e2[j].y == kCellEmpty || (uint32_t)(e2[j].y >> 40) == k24[j];
            e[j] = m0 ? e2[j].x : e2[j].y;
            open[j] = !m0 && !m1;
            again = again || open[j];
          }
        }
        if (!__any_sync(kFull, again)) break;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (it0 + 32 * j >= items) continue;  // (whole warp)
      const uint32_t cnt = e[j] != kCellEmpty ? (uint32_t)(e[j] >> 24) & 0xFFFFu : 0u;
      const bool hit = cnt > 0u;
      uint32_t off = 0;
      if (hit) off = (uint32_t)atomicAdd(&S.gtot[gq[j]], (int)cnt);
      const unsigned hm = __ballot_sync(kFull, hit);
      if (hit) {
        const int h = nh + __popc(hm & lt);
        if (h < kGHits) S.hits[h] = make_uint2((uint32_t)(e[j] & kNoBkt) | ((uint32_t)gq[j] << 24), (cnt << 16) | min(off, 0xFFFFu));
      }
      nh += __popc(hm);
    }
  }
  bool fb = false, done = !valid, gate = false;
  unsigned done_groups = 0u;
  if (nh > kGHits) {  // (more occupied cells than the hit list holds: the whole warp searches per thread)
    fb = valid;
    done = true;
    done_groups = kFull;
  }
  __syncwarp();
  // ---- rounds: as many groups as fit the pool at a time
  const float4 far4 = make_float4(3e18f, 3e18f, 3e18f, 0.f);
#pragma unroll 1
  for (;;) {
    const bool pending = lane < ng && !((done_groups >> lane) & 1u);
    const int v = pending ? S.gtot[lane] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += u;
    }
    const unsigned pend_mask = __ballot_sync(kFull, pending);
    if (pend_mask == 0u) break;
    const bool fits = pending && inc <= kGPool;
    const unsigned fit_mask = __ballot_sync(kFull, fits);
    if (fit_mask == 0u) {  // the first pending group alone exceeds the pool
      const int first = __ffs(pend_mask) - 1;
      if (!done && gnum == first) { fb = true; done = true; }
      done_groups |= 1u << first;
      continue;
    }
    if (lane < ng) S.gstart[lane] = fits ? inc - v : -1;
    __syncwarp();
    // ---- 2. copy the hit cells of this round's groups into the pool: two cells per lane and step, the loads of their
    // first buckets in flight together
#pragma unroll 1
    for (int h0 = lane; h0 < nh; h0 += 64) {
      uint2 hr[2];
      int st[2], cnt[2], base[2];
      float4 c[2][kBktE];
      uint32_t nb[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int h = h0 + 32 * u;
        st[u] = -1; cnt[u] = 0; base[u] = 0; nb[u] = kSentinel32;
        hr[u] = make_uint2(0u, 0u);
        if (h < nh) {
          hr[u] = S.hits[h];
          const int gq = (int)(hr[u].x >> 24);
          st[u] = S.gstart[gq];
          base[u] = S.gbkt[gq];
          cnt[u] = st[u] >= 0 ? (int)(hr[u].y >> 16) : 0;
        }
        const uint32_t b0 = hr[u].x & kNoBkt;
        const float4* __restrict__ p = d.bkt + (size_t)(base[u] + (int)b0) * kBktE;
#pragma unroll
        for (int e = 0; e < kBktE; ++e)
          if (e < cnt[u]) c[u][e] = __ldg(p + e);
        if (cnt[u] > kBktE) nb[u] = __ldg(d.bnext + base[u] + b0);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (cnt[u] == 0) continue;
        float4* __restrict__ dst = S.pool + st[u] + (int)(hr[u].y & 0xFFFFu);
        uint32_t b = hr[u].x & kNoBkt;
        const uint32_t at0 = (uint32_t)(base[u] + (int)b) * kBktE;
#pragma unroll
        for (int e = 0; e < kBktE; ++e) {
          if (e < cnt[u]) {
            float4 v = c[u][e];
            if (__float_as_uint(v.w) == kSentinel32) v = far4;  // removed entry
            v.w = __uint_as_float(at0 + (uint32_t)e);
            dst[e] = v;
          }
        }
        int j = min(cnt[u], kBktE);
        b = nb[u];
#pragma unroll 1
        while (j < cnt[u] && b != kSentinel32) {  // the rest of a longer chain
          const uint32_t at1 = (uint32_t)(base[u] + (int)b) * kBktE;
          const float4* __restrict__ p = d.bkt + at1;
          const int m = min(cnt[u] - j, kBktE);
          uint32_t nx = kSentinel32;
          if (cnt[u] - j > kBktE) nx = __ldg(d.bnext + base[u] + b);
#pragma unroll
          for (int e = 0; e < kBktE; ++e) {
            if (e < m) {
              float4 v = __ldg(p + e);
              if (__float_as_uint(v.w) == kSentinel32) v = far4;
              v.w = __uint_as_float(at1 + (uint32_t)e);
              dst[j + e] = v;
            }
          }
          j += m;
          b = nx;
        }
        for (; j < cnt[u]; ++j) dst[j] = far4;  // (a chain shorter than its count: never seen; keeps the pool defined)
      }
    }
    __syncwarp();
    // ---- 3. one pass over the group's candidates: the six smallest (truncated distance | position) keys
    const int gs = done ? -1 : S.gstart[gnum];
    if (gs >= 0) {
      const float4* __restrict__ P = S.pool + gs;
      const int U = S.gtot[gnum];
      // position bits riding in a key: 8 for the usual group, kGIdxBits for the large ones (fewer near-ties)
      const uint32_t kIdxMask = U <= 256 ? 0xFFu : (1u << kGIdxBits) - 1u;
      float a0 = 1.0f, a1 = 1.0f, a2 = 1.0f, a3 = 1.0f, a4 = 1.0f, a5 = 1.0f;  // the search starts from the gate
#pragma unroll 4
      for (int i = 0; i < U; ++i) {
        const float4 c = P[i];
        const float dd = dist2(qx, qy, qz, c.x, c.y, c.z);
        float t = __uint_as_float((__float_as_uint(dd) & ~kIdxMask) | (uint32_t)i), m;
        m = fminf(a0, t); t = fmaxf(a0, t); a0 = m;
        m = fminf(a1, t); t = fmaxf(a1, t); a1 = m;
        m = fminf(a2, t); t = fmaxf(a2, t); a2 = m;
        m = fminf(a3, t); t = fmaxf(a3, t); a3 = m;
        m = fminf(a4, t); t = fmaxf(a4, t); a4 = m;
        a5 = fminf(a5, t);
      }
      // Keys order like (d2 with its low bits cleared, position).  If the fifth and sixth key differ above the
      // position bits, every candidate that was not kept is strictly farther than the five kept ones, so they ARE
      // the five nearest; their exact order (d2, tag) is settled below.  Otherwise (a near-tie at the fifth place)
      // the lane runs over the candidates once more and orders everything up to the fifth key's truncated distance
      // exactly.
      const uint32_t k4 = __float_as_uint(a4), k5 = __float_as_uint(a5);
      if (k4 < 0x3F800000u) {  // five candidates below 1.0f: laserMapping.cpp:585 / :653
        gate = true;
#pragma unroll
        for (int k = 0; k < 5; ++k) { r.key[k] = kKnnInit; r.at[k] = 0u; }
        if ((k4 & ~kIdxMask) == (k5 & ~kIdxMask)) {
          const uint32_t t4 = k4 & ~kIdxMask;
#pragma unroll 1
          for (int i = 0; i < U; ++i) {
            const float4 c = P[i];
            const float dd = dist2(qx, qy, qz, c.x, c.y, c.z);
            if ((__float_as_uint(dd) & ~kIdxMask) > t4) continue;
            const uint32_t at = __float_as_uint(c.w);
            const unsigned long long key = ((unsigned long long)__float_as_uint(dd) << 32) | __float_as_uint(__ldg(&d.bkt[at].w));
            if (key < r.key[4]) {
              r.key[4] = key;
              r.at[4] = at;
#pragma unroll
              for (int q = 4; q > 0; --q) {
                const unsigned long long lo = r.key[q - 1], hi = r.key[q];
                const uint32_t alo = r.at[q - 1], ahi = r.at[q];
                const bool sw = hi < lo;
                r.key[q - 1] = sw ? hi : lo;
                r.key[q] = sw ? lo : hi;
                r.at[q - 1] = sw ? ahi : alo;
                r.at[q] = sw ? alo : ahi;
              }
            }
          }
          atomicAdd(d.knn_stats + 1, 1ull);
        } else {
          const float ak[5] = {a0, a1, a2, a3, a4};
          uint32_t db[5], atk[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const float4 c = P[__float_as_uint(ak[j]) & kIdxMask];
            atk[j] = __float_as_uint(c.w);
            db[j] = __float_as_uint(dist2(qx, qy, qz, c.x, c.y, c.z));
          }
          bool tie = false;  // the tags only matter between equal distances
#pragma unroll
          for (int j = 1; j < 5; ++j)
#pragma unroll
            for (int i = 0; i < j; ++i) tie = tie || db[i] == db[j];
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const uint32_t at = atk[j];
            const uint32_t tag = tie ? __float_as_uint(__ldg(&d.bkt[at].w)) : 0u;
            const unsigned long long key = ((unsigned long long)db[j] << 32) | tag;
            r.key[4] = key;  // (always below the seed in slot 4: fewer than five real keys so far)
            r.at[4] = at;
#pragma unroll
            for (int i = 4; i > 0; --i) {
              const unsigned long long lo = r.key[i - 1], hi = r.key[i];
              const uint32_t alo = r.at[i - 1], ahi = r.at[i];
              const bool sw = hi < lo;
              r.key[i - 1] = sw ? hi : lo;
              r.key[i] = sw ? lo : hi;
              r.at[i - 1] = sw ? ahi : alo;
              r.at[i] = sw ? alo : ahi;
            }
          }
        }
      }
      done = true;
    }
    done_groups |= fit_mask;
    __syncwarp();  // the pool is reused by the next round
  }
  // ---- lanes the grouped search could not serve
  if (__any_sync(kFull, fb)) {
    if (fb) {
      KnnStage& st = *reinterpret_cast<KnnStage*>(S.pool);
      knn5_cells(d, g, origin, qx, qy, qz, r, st);
      gate = knn_d2(r, 4) < 1.0f;
      atomicAdd(d.knn_stats, 1ull);
    }
    __syncwarp();
  }
  return gate;
}

// s2m_debug_knn: world-frame queries against segment g.  Runs the grouped search (the production path; groups form
// among whatever queries share a warp) AND the thread-per-query search; the two exact searches must agree bit for
// bit, else the call fails with S2M_ERR_INTERNAL.
__global__ void __launch_bounds__(kTile) knn_debug_kernel(Dev d, int cur, int g, const float* __restrict__ q, int n,
                                                          int32_t* __restrict__ idx, float* __restrict__ d2) {
  extern __shared__ __align__(16) unsigned char knn_group_smem[];
  KnnGroupSmem& S = reinterpret_cast<KnnGroupSmem*>(knn_group_smem)[threadIdx.x >> 5];
  __shared__ KnnStage st[kTile / 32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int* o = d.desc[seg_slot(d, g)].origin;
  const int origin[3] = {o[0], o[1], o[2]};
  bool valid = i < n;
  float qx = 0.f, qy = 0.f, qz = 0.f;
  int cx = 0, cy = 0, cz = 0;
  if (valid) {
    qx = q[3 * i]; qy = q[3 * i + 1]; qz = q[3 * i + 2];
    const float fx = floorf(qx) - (float)origin[0], fy = floorf(qy) - (float)origin[1], fz = floorf(qz) - (float)origin[2];
    valid = fabsf(fx) < 1e6f && fabsf(fy) < 1e6f && fabsf(fz) < 1e6f && cell_near_block((int)fx, (int)fy, (int)fz);
    if (valid) { cx = (int)fx; cy = (int)fy; cz = (int)fz; }
  }
  Knn5 r, r2;
  const bool ok = knn5_group(d, S, valid, g, origin, cx, cy, cz, qx, qy, qz, r);
  if (i >= n) return;
  knn5_cells(d, g, origin, qx, qy, qz, r2, st[threadIdx.x >> 5]);
  const bool ok2 = knn_d2(r2, 4) < 1.0f;  // the reference's gate; beyond it the bounded search is not exact
  bool same = ok == ok2;
  if (ok && ok2) {
#pragma unroll
    for (int k = 0; k < 5; ++k) same = same && knn_d2(r, k) == knn_d2(r2, k) && r.at[k] == r2.at[k];
  }
  if (!same) set_err(d, -7);
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    idx[5 * i + k] = ok ? tag_to_local(d, cur, g, __float_as_uint(d.bkt[r.at[k]].w)) : -1;
    d2[5 * i + k] = ok ? knn_d2(r, k) : INFINITY;
  }
}

__global__ void __launch_bounds__(kTile, S2M_K4G_MINB) knn_group_kernel(Dev d, int outer, int n_sorted) {
  extern __shared__ __align__(16) unsigned char knn_group_smem[];
  KnnGroupSmem& S = reinterpret_cast<KnnGroupSmem*>(knn_group_smem)[threadIdx.x >> 5];
  const int B = d.B, lane = threadIdx.x & 31;
  const int total = (n_sorted + 31) >> 5;
  // The next work unit (32 consecutive points of the block-sorted order, taken from a device-wide ticket) is
  // fetched while the current one is searched.
  int chunk = 0;
  if (lane == 0) chunk = atomicAdd(d.knn_ticket, 1);
  chunk = __shfl_sync(kFull, chunk, 0);
  uint32_t skey = kSentinel32, spos = 0u;
  if (chunk < total && (chunk << 5) + lane < n_sorted) { skey = d.qs_key2[(chunk << 5) + lane]; spos = d.qs_val2[(chunk << 5) + lane]; }
  float4 sp = d.ds_pts[spos];
  while (chunk < total) {
    int nchunk = 0;
    if (lane == 0) nchunk = atomicAdd(d.knn_ticket, 1);
    nchunk = __shfl_sync(kFull, nchunk, 0);
    uint32_t nkey = kSentinel32, npos = 0u;
    if (nchunk < total && (nchunk << 5) + lane < n_sorted) { nkey = d.qs_key2[(nchunk << 5) + lane]; npos = d.qs_val2[(nchunk << 5) + lane]; }
    const float4 np = d.ds_pts[npos];
    bool valid = skey != kSentinel32;
    int g = 0, slot = 0, cx = 0, cy = 0, cz = 0;
    const int pos_q = (int)spos;
    float w[3] = {0.f, 0.f, 0.f};
    int origin[3] = {0, 0, 0};
    if (valid) {
      g = (int)(skey >> 24);
      slot = g >= B ? g - B : g;
      const float4 p = sp;
      xf_point(d.lm[slot].x, p.x, p.y, p.z, w);
      origin[0] = d.desc[slot].origin[0]; origin[1] = d.desc[slot].origin[1]; origin[2] = d.desc[slot].origin[2];
      if (d.shard_world > 1 && !(w[0] >= d.shard_lo && w[0] < d.shard_hi)) {
        d.nbr[6 * (size_t)pos_q] = 0;  // sharded map: this query is answered by the rank whose x-slab holds it
        valid = false;
      } else {
        const float fx = floorf(w[0]) - (float)origin[0], fy = floorf(w[1]) - (float)origin[1], fz = floorf(w[2]) - (float)origin[2];
        const bool near = fabsf(fx) < 1e6f && fabsf(fy) < 1e6f && fabsf(fz) < 1e6f && cell_near_block((int)fx, (int)fy, (int)fz);
        if (!near) {
          d.nbr[6 * (size_t)pos_q] = 0;  // no map cell within reach: the gate fails
          valid = false;
        } else {
          cx = (int)fx; cy = (int)fy; cz = (int)fz;
        }
      }
    }
    Knn5 r;
    const bool gate = knn5_group(d, S, valid, g, origin, cx, cy, cz, w[0], w[1], w[2], r);
    if (valid) {
      int* nb = d.nbr + 6 * (size_t)pos_q;
      nb[0] = gate ? 1 : 0;
      if (gate) {  // entry numbers in d.bkt of the five neighbours, nearest first
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) nb[1 + kk] = (int)r.at[kk];
      }
    }
    chunk = nchunk; skey = nkey; spos = npos; sp = np;
  }
}

constexpr int kSumRows = 30;  // 28 sums + n_edge + n_plane, transposed through shared memory
template <bool kTrace>
__global__ void __launch_bounds__(kTile, S2M_K4B_MINB) fit_kernel(Dev d, int outer, int cur) {
  __shared__ double Tbuf[kTile / 32][kSumRows * 33];
  const int slot = blockIdx.y;
  if (!d.out[slot].optimized) return;
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int unit = blockIdx.x * (kTile / 32) + wid;
  if (unit * 32 >= nq) return;  // whole warp; nothing below synchronises the block
  const int q = unit * 32 + lane;
  const bool live = q < nq;
  const int cls = q >= nc;
  const int di = cls ? ds0 + (q - nc) : dc0 + q;
  double pose[7];  // accepted pose of the slot (guard_kernel / the previous solve)
#pragma unroll
  for (int i = 0; i < 7; ++i) pose[i] = d.lm[slot].x[i];
  bool used = false, gate = false;
  int ncand = 0;
  double rec[6] = {0, 0, 0, 0, 0, 0};
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) {
    const int* nbp = d.nbr + 6 * (size_t)di;
    const int head = nbp[0];
    gate = head & 1;
    ncand = d.count_cand ? d.cand27[di] : 0;
    p = d.ds_pts[di];
    if (gate) {
      float nb[5][3];
      int nbi[5];
#pragma unroll
      for (int k = 0; k < 5; ++k) {
        const float4 c = __ldg(d.bkt + (uint32_t)nbp[1 + k]);
        nb[k][0] = c.x; nb[k][1] = c.y; nb[k][2] = c.z;
        nbi[k] = kTrace ? tag_to_local(d, cur, cls ? d.B + slot : slot, __float_as_uint(c.w)) : 0;
      }
      if (kTrace) {
        float w[3];
        xf_point(pose, p.x, p.y, p.z, w);
        const size_t o = (size_t)outer * d.cap_in + di;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          d.tr_idx[5 * o + k] = nbi[k];
          d.tr_d2[5 * o + k] = dist2(w[0], w[1], w[2], nb[k][0], nb[k][1], nb[k][2]);
        }
      }
      used = cls == 0 ? edge_fit(nb, rec, rec + 3) : plane_fit(nb, rec, rec[3]);
      if (used) {
        double2* ro = reinterpret_cast<double2*>(d.rec + 6 * (size_t)di);
        ro[0] = make_double2(rec[0], rec[1]);
        ro[1] = make_double2(rec[2], rec[3]);
        ro[2] = make_double2(rec[4], rec[5]);
      }
    } else if (kTrace) {
      const size_t o = (size_t)outer * d.cap_in + di;
#pragma unroll
      for (int k = 0; k < 5; ++k) { d.tr_idx[5 * o + k] = -1; d.tr_d2[5 * o + k] = INFINITY; }
    }
    d.rec_valid[di] = used ? 1 : 0;
    if (kTrace) d.tr_used[(size_t)outer * d.cap_in + di] = used ? 1 : 0;
  }
  Sums28 Sm;
  Sm.zero();
  double ne = 0.0, np = 0.0;
  if (used) {
    const double cp[3] = {(double)p.x, (double)p.y, (double)p.z};
    if (cls == 0) { accum_edge(Sm, pose, cp, rec, rec + 3); ne = 1.0; }
    else { accum_plane(Sm, pose, cp, rec, rec[3]); np = 1.0; }
  }
  // ---- the unit's sums: transpose, then lane k adds row k in lane order ----
  double* T = Tbuf[wid];
#pragma unroll
  for (int k = 0; k < 28; ++k) T[k * 33 + lane] = Sm.v[k];
  T[28 * 33 + lane] = ne;
  T[29 * 33 + lane] = np;
  __syncwarp();
  double v = 0.0;
  if (lane < kSumRows) {
    const double* row = T + lane * 33;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) v += row[i];
  }
  // candidates in the 27 cells of the unit's corner / surf queries (the C-bar of SURVEY 8d's byte formula)
  int cc = cls ? 0 : ncand, cs = cls ? ncand : 0;
  for (int o = 16; o > 0; o >>= 1) {
    cc += __shfl_down_sync(kFull, cc, o);
    cs += __shfl_down_sync(kFull, cs, o);
  }
  cc = __shfl_sync(kFull, cc, 0);
  cs = __shfl_sync(kFull, cs, 0);
  if (lane == 30) v = (double)cc;
  if (lane == 31) v = (double)cs;
  d.partials[((size_t)slot * d.max_tiles + unit) * kPartial + lane] = v;
}

// Sum of the unit partials of a slot in unit order (fixed: deterministic).  Whole warp; lane k returns sum k.
__device__ __forceinline__ double sum_unit_partials(const Dev& d, int slot, int lane) {
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int units = (nq + 31) >> 5;
  const double* __restrict__ p = d.partials + (size_t)slot * d.max_tiles * kPartial + lane;
  double v = 0.0;
#pragma unroll 8
  for (int u = 0; u < units; ++u) v += __ldcg(p + (size_t)u * kPartial);
  return v;
}
// sharded map: the per-rank sums of the association wait in d.shard_sums for the allreduce
__global__ void reduce_units_kernel(Dev d) {
  const int slot = blockIdx.x;
  if (!d.out[slot].optimized) return;
  d.shard_sums[(size_t)slot * kPartial + threadIdx.x] = sum_unit_partials(d, slot, threadIdx.x);
}

// ----------------------------------------------------------------------------
// K5 + K6 solve_kernel: one Ceres solve (laserMapping.cpp:713-721, <= 4 LM iterations) in ONE
// launch.  A thread-block CLUSTER of four CTAs owns a slot: CTA 0 adds the association's unit
// partials and starts the trust-region loop; per iteration every CTA evaluates the candidate
// pose over its quarter of the cached correspondences (FP64 residuals / Jacobians / Huber),
// CTAs 1..3 hand their 32 sums to CTA 0 through distributed shared memory, CTA 0 takes the
// accept / reject decision and the next step and publishes the next candidate to all four.
// The summation order depends only on the slot's own size.
// ----------------------------------------------------------------------------
constexpr int kSolveThreads = 256;  // 128 registers each: a CTA takes half an SM, so solves of one lane interleave with other lanes' kernels
constexpr int kSolveCtas = 4;       // CTAs per cluster = per slot
constexpr int kSolveWin = 8 * kSolveThreads;  // queries whose accepted correspondences are compacted at a time
struct SolveShared {
  double red[kSolveCtas][kPartial];  // [0] this CTA's sums, [r] those of CTA r (used in CTA 0 only)
  double pose[7];            // candidate under evaluation
  int go;                    // another evaluation is wanted
  double col[kSumRows][kSolveThreads / 4 + 1];
  int list[kSolveWin];       // queries of the current window that carry a correspondence, in query order
  int woff[8 * (kSolveThreads / 32) + 1];  // per (pass, warp) of the compaction: count, then offset; [last] = total
};
__device__ __forceinline__ void solve_block_reduce(SolveShared& sh, const Sums28& Sm, double ne, double np) {
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
#pragma unroll
  for (int k = 0; k < kSumRows; ++k) {
    double v = k < 28 ? Sm.v[k < 28 ? k : 0] : (k == 28 ? ne : np);
    v += __shfl_xor_sync(kFull, v, 1);
    v += __shfl_xor_sync(kFull, v, 2);
    if ((t & 3) == 0) sh.col[k][t >> 2] = v;
  }
  __syncthreads();
  for (int k = wid; k < kSumRows; k += kSolveThreads / 32) {
    double v = 0.0;
#pragma unroll
    for (int c = 0; c < kSolveThreads / 4; c += 32) v += sh.col[k][c + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
    if (lane == 0) sh.red[0][k] = v;
  }
  __syncthreads();
}
__global__ void __cluster_dims__(kSolveCtas, 1, 1) __launch_bounds__(kSolveThreads, 2) solve_kernel(Dev d, int outer, int from_units) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int slot = blockIdx.x / kSolveCtas;
  const int half = (int)cluster.block_rank();  // which part of the slot's correspondences this CTA evaluates
  if (blockIdx.x == 0 && threadIdx.x == 0) *d.knn_ticket = 0;  // the association finished: re-arm its work ticket
  if (!d.out[slot].optimized) return;  // every CTA of the cluster
  __shared__ SolveShared sh;
  __shared__ LmState Ls;
  const int t = threadIdx.x;
  SolveShared* peer0 = cluster.map_shared_rank(&sh, 0);
  // CTA 0, threads 0..7: the next candidate (or the end of the solve) to every CTA
  auto publish = [&]() {
    if (t < 8) {
      const int go = !Ls.done && Ls.have_candidate;
      for (int r = 0; r < kSolveCtas; ++r) {
        SolveShared* pr = cluster.map_shared_rank(&sh, r);
        if (t < 7) pr->pose[t] = Ls.xc[t]; else pr->go = go;
      }
    }
  };
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int per_cta = (nq + kSolveCtas - 1) / kSolveCtas, q_lo = half * per_cta, q_hi = min(nq, q_lo + per_cta);
  // ---- start of the solve: sums of the first evaluation, lm_begin ----
  if (half == 0) {
    lm_load(&Ls, d.lm + slot);
    if (from_units) {
      if (t < kPartial) sh.red[1][t] = sum_unit_partials(d, slot, t);
      __syncthreads();
      if (t == 0) {
        Sums28 S0;
        for (int i = 0; i < 28; ++i) S0.v[i] = sh.red[1][i];
        SlotOut& o = d.out[slot];
        o.n_edge[outer] = (int)sh.red[1][28]; o.n_plane[outer] = (int)sh.red[1][29];
        if (outer == 0) { o.cand[0] = sh.red[1][30]; o.cand[1] = sh.red[1][31]; }
        double x0[7];
        for (int i = 0; i < 7; ++i) x0[i] = Ls.x[i];
        lm_begin(Ls, x0, S0, (int)sh.red[1][28] + (int)sh.red[1][29], 4);
      }
    }
    __syncthreads();
    publish();
  }
  cluster.sync();
  for (int it = 0; it < 4 && sh.go; ++it) {  // options.max_num_iterations = 4 (:716)
    Sums28 Sm;
    Sm.zero();
    double ne = 0.0, np = 0.0;
    // This CTA's quarter of the slot's queries, a window at a time: about a third carry a correspondence, so they
    // are compacted first (ballot + prefix, query order) and the evaluation runs on dense lanes.
    for (int w0 = q_lo; w0 < q_hi; w0 += kSolveWin) {
      const int lane = t & 31, wid = t >> 5;
      unsigned have[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int q = w0 + k * kSolveThreads + t;
        const bool v = q < q_hi && d.rec_valid[q < nc ? dc0 + q : ds0 + (q - nc)];
        have[k] = __ballot_sync(kFull, v);
        if (lane == 0) sh.woff[k * (kSolveThreads / 32) + wid] = __popc(have[k]);
      }
      __syncthreads();
      if (wid == 0) {  // exclusive prefix over the 64 counts, two per lane
        const int a = sh.woff[2 * lane], b = sh.woff[2 * lane + 1];
        int incl = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += v;
        }
        sh.woff[2 * lane] = incl - a - b;
        sh.woff[2 * lane + 1] = incl - b;
        if (lane == 31) sh.woff[8 * (kSolveThreads / 32)] = incl;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if ((have[k] >> lane) & 1u)
          sh.list[sh.woff[k * (kSolveThreads / 32) + wid] + __popc(have[k] & ((1u << lane) - 1u))] = w0 + k * kSolveThreads + t;
      __syncthreads();
      const int n_list = sh.woff[8 * (kSolveThreads / 32)];
      for (int e = t; e < n_list; e += kSolveThreads) {
        const int q = sh.list[e];
        const int di = q < nc ? dc0 + q : ds0 + (q - nc);
        const float4 p = d.ds_pts[di];
        const double cp[3] = {(double)p.x, (double)p.y, (double)p.z};
        const double2* ri = reinterpret_cast<const double2*>(d.rec + 6 * (size_t)di);
        const double2 a = ri[0], b = ri[1], c = ri[2];
        const double rec[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
        if (q < nc) { accum_edge(Sm, sh.pose, cp, rec, rec + 3); ne += 1.0; }
        else { accum_plane(Sm, sh.pose, cp, rec, rec[3]); np += 1.0; }
      }
      __syncthreads();  // (the list is rebuilt by the next window)
    }
    solve_block_reduce(sh, Sm, ne, np);
    if (half != 0 && t < kPartial) peer0->red[half][t] = sh.red[0][t];
    cluster.sync();
    if (half == 0) {
      if (t < kPartial) {
        double v = sh.red[0][t];
        for (int r = 1; r < kSolveCtas; ++r) v += sh.red[r][t];  // fixed order
        sh.red[0][t] = v;
      }
      __syncthreads();
      if (t == 0) {
        Sums28 S1;
        for (int i = 0; i < 28; ++i) S1.v[i] = sh.red[0][i];
        lm_after_eval(Ls, S1, 4);
      }
      __syncthreads();
      publish();
    }
    cluster.sync();
  }
  if (half == 0) {
    if (t == 0) {
      SlotOut& o = d.out[slot];
      o.lm_iters[outer] = Ls.iteration; o.lm_term[outer] = Ls.termination;
      o.cost_initial[outer] = Ls.initial_cost; o.cost_final[outer] = Ls.final_cost;
    }
    lm_store(d.lm + slot, &Ls);
  }
}

// ----------------------------------------------------------------------------
// K5 + K6: evaluation at the LM candidate pose from the cached correspondences,
// then (last block) the accept / reject decision and the next trust-region step.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(kTile, 4) evaluate_kernel(Dev d, int outer) {
  const int slot = blockIdx.y;
  if (!d.out[slot].optimized) return;
  LmState& L = d.lm[slot];
  if (L.done || !L.have_candidate) return;
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int ntiles = (nq + kTile - 1) / kTile;
  const int nwork = max(1, (ntiles + kEvalTilesPerBlock - 1) / kEvalTilesPerBlock);  // see fit_kernel
  if ((int)blockIdx.x >= nwork) return;
  __shared__ double pose[7];
  __shared__ BlockAcc A;
  __shared__ double red[kPartial];
  if (threadIdx.x < 7) pose[threadIdx.x] = L.xc[threadIdx.x];
  acc_zero(A);
  __syncthreads();
  // each thread keeps the sums of all its records in registers; one reduction per block
  Sums28 S;
  S.zero();
  double ne = 0.0, np = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += nwork) {
    const int q = tile * kTile + threadIdx.x;
    const int di = q < nc ? dc0 + q : ds0 + (q - nc);
    if (q < nq && d.rec_valid[di]) {
      const float4 p = d.ds_pts[di];
      const double cp[3] = {(double)p.x, (double)p.y, (double)p.z};
      const double2* ri = reinterpret_cast<const double2*>(d.rec + 6 * (size_t)di);
      const double2 a = ri[0], b = ri[1], c = ri[2];
      const double rec[6] = {a.x, a.y, b.x, b.y, c.x, c.y};
      if (q < nc) { accum_edge(S, pose, cp, rec, rec + 3); ne += 1.0; }
      else { accum_plane(S, pose, cp, rec, rec[3]); np += 1.0; }
    }
  }
  acc_add(A, S, ne, np, 0.0, 0.0);
  acc_store(A, d.partials + ((size_t)slot * d.max_tiles + blockIdx.x) * kPartial);
  if (!block_is_last(d.ticket + slot, nwork)) return;
  sum_partials(d, slot, nwork, red);
  if (d.shard_world > 1) { shard_publish(d, slot, red); return; }
  __shared__ LmState Ls;
  lm_tail_after(d, slot, outer, red, &Ls);
}

// ----------------------------------------------------------------------------
// Scan-to-scan odometry (laserOdometry.cpp:277-505): the queries are the
// sharp / flat points of the current sweep, the targets the less-sharp / less-flat clouds of
// the previous one (d.od_last, ring-major as the feature extraction left them).  Per query:
// TransformToStart (:108-126 with s = 1: q*p + t), exact nearest neighbour (order (d2, index)), the ring-constrained
// second (and third) neighbour walks of :313-357 / :401-452, then the same residual / Jacobian /
// Huber accumulation, block reduction and LM start as the mapping association: an edge factor
// (lp-a) x (lp-b) / |a-b| is (lp - a) x u with u = (a-b)/|a-b|; the three-point plane factor
// (lidarFactor.hpp:57-104) is n.lp + d with n = normalize((j-l) x (j-m)), d = -n.j.
// ----------------------------------------------------------------------------
// lower bound of the float squared distance from `sel` to any point inside the box of a chunk
__device__ __forceinline__ float chunk_bound(const float4 ma, const float4 mb, const float sel[3]) {
  const float bx = fmaxf(0.0f, fmaxf(xfsub(ma.x, sel[0]), xfsub(sel[0], ma.w)));
  const float by = fmaxf(0.0f, fmaxf(xfsub(ma.y, sel[1]), xfsub(sel[1], mb.x)));
  const float bz = fmaxf(0.0f, fmaxf(xfsub(ma.z, sel[2]), xfsub(sel[2], mb.y)));
  return xfadd(xfadd(xfmul(bx, bx), xfmul(by, by)), xfmul(bz, bz));
}
__device__ __forceinline__ float odom_sq(const float4 p, const float sel[3]) {  // the float expression of :322-327
  const float dx = xfsub(p.x, sel[0]), dy = xfsub(p.y, sel[1]), dz = xfsub(p.z, sel[2]);
  return xfadd(xfadd(xfmul(dx, dx), xfmul(dy, dy)), xfmul(dz, dz));
}
// Correspondence search in three kernels:
//   odom_search_kernel    every query: 27 half-metre cells around the transformed point through a
//                         cell-ordered copy of the previous cloud.  A nearest neighbour found closer than
//                         0.5 m is THE nearest neighbour (everything outside the 27 cells is at least 0.5 m
//                         away); likewise a second / third neighbour closer than 0.5 m.  Anything not
//                         settled that way is appended to a list.
//   odom_fallback_kernel  the listed queries, one warp each: exhaustive nearest neighbour with box bounds
//                         per 32-point chunk, the reference's two walks 32 points per step.
//   odom_fit_kernel       factors, Huber, the 28 sums, block reduction, start of the LM solve.
constexpr int kOdNeedNN = 1, kOdNeedWalk = 2;
constexpr int kOdSid = 256;  // ring numbers 0..255 (checked when the previous cloud is indexed)
constexpr float kOdSettled = 0.25f;  // (cell size)^2: everything outside the 27 cells is at least one cell away
constexpr int kOdBias = 512;  // cell coordinates are floor(2 x) + 512 in 0..1023: half-metre cells, +-256 m

__device__ __forceinline__ int od_cell(float v) { return (int)floorf(xfmul(v, 2.0f)) + kOdBias; }  // 2 v is exact
struct OdCloud {  // the previous cloud of one (class, slot)
  int l0, ln;
  const float4* last;      // ring-major, as received
  const float4* meta;      // box + ring range per 32-point chunk
  int nch;
  const float4* sorted;    // cell-ordered copy, .w = index | ring << 24
  const uint32_t* ckey;    // cell key of every entry of `sorted`
  const int* first_ge;     // [kOdSid + 1] first index whose ring number is >= r (ln if none)
  const int* last_le;      // [kOdSid + 1] last index whose ring number is <= r - 1 (-1 if none); entry r = rings < r
};
__device__ __forceinline__ OdCloud od_cloud(const Dev& d, int seg) {
  OdCloud c;
  c.l0 = d.od_last_off[seg]; c.ln = d.od_last_off[seg + 1] - c.l0;
  c.last = d.od_last + c.l0;
  const int c0 = d.od_chunk_off[seg];
  c.meta = d.od_meta + 2 * (size_t)c0; c.nch = d.od_chunk_off[seg + 1] - c0;
  c.sorted = d.od_sorted + c.l0; c.ckey = d.od_ckey + c.l0;
  c.first_ge = d.od_first_ge + (size_t)seg * (kOdSid + 1); c.last_le = d.od_last_le + (size_t)seg * (kOdSid + 1);
  return c;
}
// class of a candidate of the ring walks: 2 = "second" neighbour, 3 = "third" (surf only), 0 = none
__device__ __forceinline__ int od_class(int cls, bool fwd, int sid, int id) {
  if (cls == 0) return (fwd ? sid > id : sid < id) ? 2 : 0;
  return (fwd ? sid <= id : sid >= id) ? 2 : 3;
}
// first index that stops each walk of the reference (:317-318, :340-341, :404-405, :430-431)
__device__ __forceinline__ void od_breaks(const OdCloud& C, int best_i, int id, int& jbeg, int& jend) {
  const int hi = id + 3, lo = id - 3;  // (double)sid > id + 2.5  <=>  sid >= id + 3 (NEARBY_SCAN, :64)
  // tables: the first index anywhere with ring >= hi is the break if it lies behind best_i (likewise the last
  // index with ring <= lo before it); only a cloud that is not ring-major there needs the search below
  const int fg = hi > kOdSid ? C.ln : C.first_ge[max(hi, 0)];
  const int ll = lo < 0 ? -1 : C.last_le[min(lo + 1, kOdSid)];
  if (fg > best_i && ll < best_i) { jend = fg; jbeg = ll; return; }
  jend = C.ln; jbeg = -1;
  for (int j = best_i + 1; j < C.ln;) {
    const int c = j >> 5;
    if (__float_as_int(C.meta[2 * c + 1].w) < hi) { j = 32 * c + 32; continue; }
    const int je = min(32 * c + 32, C.ln);
    for (; j < je; ++j)
      if ((int)C.last[j].w >= hi) { jend = j; break; }
    if (jend != C.ln) break;
  }
  for (int j = best_i - 1; j >= 0;) {
    const int c = j >> 5;
    if (__float_as_int(C.meta[2 * c + 1].z) > lo) { j = 32 * c - 1; continue; }
    for (; j >= 32 * c; --j)
      if ((int)C.last[j].w <= lo) { jbeg = j; break; }
    if (jbeg != -1) break;
  }
}
__global__ void __launch_bounds__(kTile, S2M_OD_MINB) odom_search_kernel(Dev d) {
  const int slot = blockIdx.y;
  if (!d.out[slot].optimized) return;
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int q = blockIdx.x * kTile + threadIdx.x;
  if (q >= nq) return;
  const int cls = q >= nc;
  const int di = cls ? ds0 + (q - nc) : dc0 + q;
  const float4 p = d.ds_pts[di];
  float sel[3];
  xf_point(d.lm[slot].x, p.x, p.y, p.z, sel);
  const OdCloud C = od_cloud(d, cls * d.B + slot);
  int best_i = -1, second = -1, third = -1, flags = 0;
  float best_d = INFINITY;
  uint32_t best_w = 0;
  const int cx = od_cell(sel[0]), cy = od_cell(sel[1]), cz = od_cell(sel[2]);
  int rlo[9];
  if (cx < 1 || cx > 1022 || cy < 1 || cy > 1022 || cz < 1 || cz > 1022) {
    flags = kOdNeedNN;
  } else {
    // ---- nearest neighbour among the 27 cells (nine rows of three x-adjacent cells) ----
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const uint32_t klo = ((uint32_t)(cz + r / 3 - 1) << 20) | ((uint32_t)(cy + r % 3 - 1) << 10) | (uint32_t)(cx - 1);
      int a = 0, b = C.ln;
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (C.ckey[mid] < klo) a = mid + 1; else b = mid;
      }
      rlo[r] = a;
      // the row's run of entries, four at a time (their loads in flight together; the run ends at the first key beyond the row)
      for (int e = a; e < C.ln; e += 4) {
        uint32_t ck[4];
        float4 c4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int eu = min(e + u, C.ln - 1);
          ck[u] = C.ckey[eu];
          c4[u] = C.sorted[eu];
        }
        bool more = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          more = more && e + u < C.ln && ck[u] <= klo + 2u;
          if (more) {
            const float dd = dist2(sel[0], sel[1], sel[2], c4[u].x, c4[u].y, c4[u].z);
            const uint32_t w = __float_as_uint(c4[u].w);
            const int idx = (int)(w & 0xFFFFFFu);
            if (dd < best_d || (dd == best_d && idx < best_i)) { best_d = dd; best_i = idx; best_w = w; }
          }
        }
        if (!more) break;
      }
    }
    if (!(best_i >= 0 && best_d < kOdSettled)) flags = kOdNeedNN;  // not settled: something outside the cells may be nearer
  }
  if (flags == 0) {  // (best_d < 1 implies the 25 gate of :306 / :394)
    // ---- second / third neighbour among the same cells ----
    const int id = (int)(best_w >> 24);
    int jbeg, jend;
    od_breaks(C, best_i, id, jbeg, jend);
    float m2 = 25.0f, m3 = 25.0f;
    int k2 = INT_MAX, k3 = INT_MAX;  // position in the reference's walk order of the current minima
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const uint32_t khi = (((uint32_t)(cz + r / 3 - 1) << 20) | ((uint32_t)(cy + r % 3 - 1) << 10) | (uint32_t)(cx - 1)) + 2u;
      for (int e0 = rlo[r]; e0 < C.ln; e0 += 4) {
        uint32_t ck[4];
        float4 c44[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int eu = min(e0 + u, C.ln - 1);
          ck[u] = C.ckey[eu];
          c44[u] = C.sorted[eu];
        }
        bool more = true;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          more = more && e0 + u < C.ln && ck[u] <= khi;
          if (!more) continue;
          const float4 c4 = c44[u];
          const uint32_t w = __float_as_uint(c4.w);
          const int j = (int)(w & 0xFFFFFFu);
          if (j <= jbeg || j >= jend || j == best_i) continue;
          const bool fwd = j > best_i;
          const int k = od_class(cls, fwd, (int)(w >> 24), id);
          if (k == 0) continue;
          const float dd = odom_sq(c4, sel);
          const int rank = fwd ? j - best_i : C.ln + best_i - j;
          if (k == 2) { if (dd < m2 || (dd == m2 && rank < k2)) { m2 = dd; k2 = rank; second = j; } }
          else if (dd < m3 || (dd == m3 && rank < k3)) { m3 = dd; k3 = rank; third = j; }
        }
        if (!more) break;
      }
    }
    if (!(m2 < kOdSettled) || (cls == 1 && !(m3 < kOdSettled))) { flags = kOdNeedWalk; second = third = -1; }
  }
  d.od_corr[di] = make_int4(best_i, second, third, flags);
  d.od_bestd[di] = best_d;
  if (flags) d.od_fb_list[atomicAdd(d.od_fb_cnt, 1)] = (slot << 20) | q;
}

// One WARP per listed query: the chains of a thread-per-query fallback (tens of thousands of dependent
// steps) would set the duration of the whole step.  Nearest neighbour: lanes test 32 chunk boxes at a time,
// chunks that can still win are scanned 32 points per step.  Walks: the two ranges of the reference's walks
// (up to the exact break positions), 32 consecutive points per step; minima reduced with the walk's tie rule
// (forwards the lowest index, backwards the highest, and the backward one only if strictly closer).
__device__ __forceinline__ float warp_min_f(float v) {  // v >= 0: the bit patterns order like the values
  return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
__global__ void __launch_bounds__(kTile) odom_fallback_kernel(Dev d) {
  const int n = *d.od_fb_cnt;
  const int lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += (gridDim.x * blockDim.x) >> 5) {
    const int item = d.od_fb_list[i];
    const int slot = item >> 20, q = item & 0xFFFFF;
    int dc0, nc, ds0, nq;
    slot_counts(d, slot, dc0, nc, ds0, nq);
    const int cls = q >= nc;
    const int di = cls ? ds0 + (q - nc) : dc0 + q;
    const float4 p = d.ds_pts[di];
    float sel[3];
    xf_point(d.lm[slot].x, p.x, p.y, p.z, sel);
    const OdCloud C = od_cloud(d, cls * d.B + slot);
    const int4 co = d.od_corr[di];
    float best_d = d.od_bestd[di];
    int best_i = co.x;
    // second attempt through the cells: 5 x 5 x 5 half-metre cells = everything within 1 m, one row of five
    // x-adjacent cells per lane
    const int cx = od_cell(sel[0]), cy = od_cell(sel[1]), cz = od_cell(sel[2]);
    const bool wide = cx >= 2 && cx <= 1021 && cy >= 2 && cy <= 1021 && cz >= 2 && cz <= 1021;
    uint32_t row_hi = 0;
    int row_lo = 0;
    if (wide && lane < 25) {
      const uint32_t klo = ((uint32_t)(cz + lane / 5 - 2) << 20) | ((uint32_t)(cy + lane % 5 - 2) << 10) | (uint32_t)(cx - 2);
      int a = 0, b = C.ln;
      while (a < b) {
        const int mid = (a + b) >> 1;
        if (C.ckey[mid] < klo) a = mid + 1; else b = mid;
      }
      row_lo = a;
      row_hi = klo + 4u;
    }
    const bool have_row = wide && lane < 25;
    bool need_nn = (co.w & kOdNeedNN) != 0;
    if (need_nn && wide) {
      float ld = INFINITY;
      int li = INT_MAX;
      if (have_row)
        for (int e = row_lo; e < C.ln && C.ckey[e] <= row_hi; ++e) {
          const float4 c4 = C.sorted[e];
          const float dd = dist2(sel[0], sel[1], sel[2], c4.x, c4.y, c4.z);
          const int idx = (int)(__float_as_uint(c4.w) & 0xFFFFFFu);
          if (dd < ld || (dd == ld && idx < li)) { ld = dd; li = idx; }
        }
      const float m = warp_min_f(ld);
      if (m < 1.0f) {  // everything outside the block is at least 1 m away
        best_d = m;
        best_i = __reduce_min_sync(full, ld == m ? li : INT_MAX);
        need_nn = false;
      }
    }
    if (need_nn) {  // exhaustive: every chunk whose box can still hold the minimum
      const float bound = fminf(best_d, 25.0f);
      best_d = INFINITY; best_i = -1;
      for (int cb = 0; cb < C.nch; cb += 32) {
        const int c = cb + lane;
        const float lb = c < C.nch ? chunk_bound(C.meta[2 * c], C.meta[2 * c + 1], sel) : INFINITY;
        for (unsigned need = __ballot_sync(full, lb <= fminf(bound, best_d)); need; need &= need - 1) {
          const int k = __ffs(need) - 1;
          if (__shfl_sync(full, lb, k) > fminf(bound, best_d)) continue;
          const int j = 32 * (cb + k) + lane;
          float dd = INFINITY;
          if (j < C.ln) { const float4 c4 = C.last[j]; dd = dist2(sel[0], sel[1], sel[2], c4.x, c4.y, c4.z); }
          const float m = warp_min_f(dd);
          if (m < best_d) {  // chunks come in index order: a later equal distance never replaces
            best_d = m;
            best_i = __reduce_min_sync(full, dd == m ? j : INT_MAX);
          }
        }
      }
    }
    int second = -1, third = -1;
    if (best_i >= 0 && (double)best_d < 25.0) {  // DISTANCE_SQ_THRESHOLD (:63)
      const int id = (int)C.last[best_i].w;
      int jbeg, jend;
      od_breaks(C, best_i, id, jbeg, jend);
      // second / third neighbour among the same 125 cells: settled if closer than 1 m
      bool settled = false;
      if (wide) {
        float l2 = 25.0f, l3 = 25.0f;
        int r2 = INT_MAX, r3 = INT_MAX, j2 = -1, j3 = -1;
        if (have_row)
          for (int e = row_lo; e < C.ln && C.ckey[e] <= row_hi; ++e) {
            const float4 c4 = C.sorted[e];
            const uint32_t w = __float_as_uint(c4.w);
            const int j = (int)(w & 0xFFFFFFu);
            if (j <= jbeg || j >= jend || j == best_i) continue;
            const bool fwd = j > best_i;
            const int k = od_class(cls, fwd, (int)(w >> 24), id);
            if (k == 0) continue;
            const float dd = odom_sq(c4, sel);
            const int rank = fwd ? j - best_i : C.ln + best_i - j;  // position in the reference's walk order
            if (k == 2) { if (dd < l2 || (dd == l2 && rank < r2)) { l2 = dd; r2 = rank; j2 = j; } }
            else if (dd < l3 || (dd == l3 && rank < r3)) { l3 = dd; r3 = rank; j3 = j; }
          }
        const float m2 = warp_min_f(l2), m3 = warp_min_f(l3);
        if (m2 < 1.0f && (cls == 0 || m3 < 1.0f)) {
          const int q2 = __reduce_min_sync(full, l2 == m2 ? r2 : INT_MAX);
          second = __reduce_max_sync(full, (l2 == m2 && r2 == q2) ? j2 : -1);
          if (cls == 1) {
            const int q3 = __reduce_min_sync(full, l3 == m3 ? r3 : INT_MAX);
            third = __reduce_max_sync(full, (l3 == m3 && r3 == q3) ? j3 : -1);
          }
          settled = true;
        }
      }
      float f2 = 25.0f, f3 = 25.0f, b2 = 25.0f, b3 = 25.0f;  // minPointSqDis2 / 3, forward and backward parts
      int fi2 = -1, fi3 = -1, bi2 = -1, bi3 = -1;                // (warp-uniform)
      // chunks of the two ranges, 32 boxes per step; a chunk is scanned only if its box bound is below the
      // running minimum of a class its ring range can hold (a candidate has to be strictly closer to win)
      for (int dir = 0; dir < 2 && !settled; ++dir) {
        const int ca = dir == 0 ? (best_i + 1) >> 5 : (jbeg + 1) >> 5;
        const int cz = dir == 0 ? (jend - 1) >> 5 : (best_i - 1) >> 5;
        if ((dir == 0 && best_i + 1 >= jend) || (dir == 1 && best_i - 1 <= jbeg)) continue;
        for (int step = 0; step * 32 <= cz - ca; ++step) {
          const int c = dir == 0 ? ca + step * 32 + lane : cz - step * 32 - lane;  // walk order
          float lb = INFINITY;
          bool may2 = false, may3 = false;
          if (c >= ca && c <= cz) {
            const float4 ma = C.meta[2 * c], mb = C.meta[2 * c + 1];
            lb = chunk_bound(ma, mb, sel);
            const int smin = __float_as_int(mb.z), smax = __float_as_int(mb.w);
            if (cls == 0) may2 = dir == 0 ? smax > id : smin < id;
            else { may2 = dir == 0 ? smin <= id : smax >= id; may3 = dir == 0 ? smax > id : smin < id; }
          }
          const float t2 = dir == 0 ? f2 : fminf(f2, b2), t3 = dir == 0 ? f3 : fminf(f3, b3);
          for (unsigned need = __ballot_sync(full, (may2 && lb < t2) || (may3 && lb < t3)); need; need &= need - 1) {
            const int k = __ffs(need) - 1;
            const int ck = __shfl_sync(full, c, k);
            const float lbk = __shfl_sync(full, lb, k);
            const float u2 = dir == 0 ? f2 : fminf(f2, b2), u3 = dir == 0 ? f3 : fminf(f3, b3);
            if (!(lbk < u2) && !(lbk < u3)) continue;  // the minima moved since the ballot
            const int j = 32 * ck + lane;
            int kl = 0;
            float dd = INFINITY;
            if (j < C.ln && (dir == 0 ? (j > best_i && j < jend) : (j < best_i && j > jbeg))) {
              const float4 c4 = C.last[j];
              kl = od_class(cls, dir == 0, (int)c4.w, id);
              dd = odom_sq(c4, sel);
            }
            const float m2 = warp_min_f(kl == 2 ? dd : INFINITY), m3 = warp_min_f(kl == 3 ? dd : INFINITY);
            if (dir == 0) {  // first met = lowest index
              if (m2 < f2) { f2 = m2; fi2 = __reduce_min_sync(full, (kl == 2 && dd == m2) ? j : INT_MAX); }
              if (m3 < f3) { f3 = m3; fi3 = __reduce_min_sync(full, (kl == 3 && dd == m3) ? j : INT_MAX); }
            } else {         // first met = highest index
              if (m2 < b2) { b2 = m2; bi2 = __reduce_max_sync(full, (kl == 2 && dd == m2) ? j : -1); }
              if (m3 < b3) { b3 = m3; bi3 = __reduce_max_sync(full, (kl == 3 && dd == m3) ? j : -1); }
            }
          }
        }
      }
      if (!settled) {
        second = (bi2 >= 0 && b2 < f2) ? bi2 : fi2;
        third = (bi3 >= 0 && b3 < f3) ? bi3 : fi3;
      }
    } else {
      best_i = -1;
    }
    if (lane == 0) {
      d.od_corr[di] = make_int4(best_i, second, third, 0);
      d.od_bestd[di] = best_d;
    }
  }
}

template <bool kTrace>
__global__ void __launch_bounds__(kTile, 4) odom_fit_kernel(Dev d, int outer) {
  const int slot = blockIdx.y;
  if (!d.out[slot].optimized) return;
  int dc0, nc, ds0, nq;
  slot_counts(d, slot, dc0, nc, ds0, nq);
  const int nwork = max(1, (nq + kTile - 1) / kTile);  // one tile per block; depends only on the slot's own size
  if ((int)blockIdx.x >= nwork) return;
  __shared__ double pose[7];
  __shared__ BlockAcc A;
  __shared__ double red[kPartial];
  if (threadIdx.x < 7) pose[threadIdx.x] = d.lm[slot].x[threadIdx.x];
  acc_zero(A);
  __syncthreads();
  const int q = blockIdx.x * kTile + threadIdx.x;
  const bool live = q < nq;
  const int cls = q >= nc;
  const int di = cls ? ds0 + (q - nc) : dc0 + q;
  Sums28 S;
  S.zero();
  double ne = 0.0, np = 0.0;
  if (live) {
    const int4 co = d.od_corr[di];
    const float4 p = d.ds_pts[di];
    const float4* __restrict__ last = d.od_last + d.od_last_off[cls * d.B + slot];
    bool used = false;
    double rec[6] = {0, 0, 0, 0, 0, 0};
    const double cp[3] = {(double)p.x, (double)p.y, (double)p.z};
    if (co.x >= 0 && cls == 0 && co.y >= 0) {
      const float4 a = last[co.x], b = last[co.y];
      const double e[3] = {(double)a.x - (double)b.x, (double)a.y - (double)b.y, (double)a.z - (double)b.z};
      const double inv = 1.0 / sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      rec[0] = a.x; rec[1] = a.y; rec[2] = a.z;
      rec[3] = e[0] * inv; rec[4] = e[1] * inv; rec[5] = e[2] * inv;
      accum_edge(S, pose, cp, rec, rec + 3);
      ne = 1.0;
      used = true;
    } else if (co.x >= 0 && cls == 1 && co.y >= 0 && co.z >= 0) {
      const float4 a = last[co.x], l = last[co.y], m = last[co.z];
      const double u[3] = {(double)a.x - (double)l.x, (double)a.y - (double)l.y, (double)a.z - (double)l.z};
      const double v[3] = {(double)a.x - (double)m.x, (double)a.y - (double)m.y, (double)a.z - (double)m.z};
      double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
      const double z = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
      if (z > 0.0) { const double nn = sqrt(z); n[0] /= nn; n[1] /= nn; n[2] /= nn; }
      rec[0] = n[0]; rec[1] = n[1]; rec[2] = n[2];
      rec[3] = -(n[0] * (double)a.x + n[1] * (double)a.y + n[2] * (double)a.z);
      accum_plane(S, pose, cp, rec, rec[3]);
      np = 1.0;
      used = true;
    }
    if (used) {
      double2* ro = reinterpret_cast<double2*>(d.rec + 6 * (size_t)di);
      ro[0] = make_double2(rec[0], rec[1]);
      ro[1] = make_double2(rec[2], rec[3]);
      ro[2] = make_double2(rec[4], rec[5]);
    }
    d.rec_valid[di] = used ? 1 : 0;
    if (kTrace) {
      const size_t o = (size_t)outer * d.cap_in + di;
      d.tr_idx[5 * o] = co.x; d.tr_idx[5 * o + 1] = co.y; d.tr_idx[5 * o + 2] = co.z;
      d.tr_idx[5 * o + 3] = d.tr_idx[5 * o + 4] = -1;
      d.tr_d2[5 * o] = d.od_bestd[di];
      d.tr_used[o] = used;
    }
  }
  acc_add(A, S, ne, np, 0.0, 0.0);
  acc_store(A, d.partials + ((size_t)slot * d.max_tiles + blockIdx.x) * kPartial);
  if (!block_is_last(d.ticket + slot, nwork)) return;
  sum_partials(d, slot, nwork, red);
  __shared__ LmState Ls;
  lm_tail_begin(d, slot, outer, red, &Ls);
}
// Previous clouds a second time, ordered by half-metre cell (x fastest) inside each cloud:
// key = [segment][cz][cy][cx] with biased coordinates; .w of the copy = index | ring << 24.
__global__ void odom_sort_key_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = find_seg(d.od_last_off, 2 * d.B, i);
  const float4 p = d.od_last[i];
  const int cx = od_cell(p.x), cy = od_cell(p.y), cz = od_cell(p.z);
  const int sid = (int)p.w;
  if ((unsigned)cx > 1023u || (unsigned)cy > 1023u || (unsigned)cz > 1023u || (unsigned)sid > 255u) set_err(d, -4);  // S2M_ERR_RANGE
  d.od_key[i] = ((unsigned long long)g << 30) | ((unsigned long long)(cz & 1023) << 20) | ((unsigned long long)(cy & 1023) << 10) |
                (unsigned long long)(cx & 1023);
  d.od_val[i] = (uint32_t)(i - d.od_last_off[g]);
}
__global__ void odom_gather_sorted_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = (int)(d.od_key2[i] >> 30);
  const uint32_t l = d.od_val2[i];
  const float4 p = d.od_last[d.od_last_off[g] + l];
  d.od_sorted[i] = make_float4(p.x, p.y, p.z, __uint_as_float((l & 0xFFFFFFu) | ((uint32_t)((int)p.w & 255) << 24)));
  d.od_ckey[i] = (uint32_t)(d.od_key2[i] & 0x3FFFFFFFull);
}
// ring tables of every previous cloud: od_first_ge[g][r] = first index with ring >= r, od_last_le[g][r] = last
// index with ring < r.  Built from the first / last index of every ring number.
__global__ void odom_ring_init_kernel(Dev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * d.B * (kOdSid + 1)) return;
  d.od_first_ge[t] = INT_MAX;
  d.od_last_le[t] = -1;
}
__global__ void odom_ring_mark_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = find_seg(d.od_last_off, 2 * d.B, i);
  const int l = i - d.od_last_off[g];
  const int sid = min(max((int)d.od_last[i].w, 0), kOdSid - 1);
  atomicMin(d.od_first_ge + (size_t)g * (kOdSid + 1) + sid, l);      // for now: first index OF ring sid
  atomicMax(d.od_last_le + (size_t)g * (kOdSid + 1) + sid + 1, l);   // for now: last index OF ring sid, at [sid + 1]
}
__global__ void odom_ring_scan_kernel(Dev d) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 2 * d.B) return;
  const int ln = d.od_last_off[g + 1] - d.od_last_off[g];
  int* fg = d.od_first_ge + (size_t)g * (kOdSid + 1);
  int* ll = d.od_last_le + (size_t)g * (kOdSid + 1);
  int run = ln;
  for (int r = kOdSid; r >= 0; --r) {  // suffix minimum
    run = min(run, fg[r] == INT_MAX ? ln : fg[r]);
    fg[r] = run;
  }
  int best = -1;
  for (int r = 0; r <= kOdSid; ++r) {  // prefix maximum: entry r covers rings < r
    best = max(best, ll[r]);
    ll[r] = best;
  }
}
// box + ring-number range of every 32-point chunk of the previous clouds (one warp per chunk):
// meta[2c] = (min x, min y, min z, max x), meta[2c+1] = (max y, max z, bits of min ring, bits of max ring)
__global__ void odom_meta_kernel(Dev d, int nchunks) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= nchunks) return;
  const int g = find_seg(d.od_chunk_off, 2 * d.B, c);
  const int j = 32 * (c - d.od_chunk_off[g]) + lane;
  const int l0 = d.od_last_off[g], ln = d.od_last_off[g + 1] - l0;
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  int smin = INT_MAX, smax = INT_MIN;
  if (j < ln) {
    const float4 p = d.od_last[l0 + j];
    mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z;
    smin = smax = (int)p.w;
  }
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
    smin = min(smin, __shfl_xor_sync(0xffffffffu, smin, o));
    smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, o));
  }
  if (lane == 0) {
    d.od_meta[2 * (size_t)c] = make_float4(mn[0], mn[1], mn[2], mx[0]);
    d.od_meta[2 * (size_t)c + 1] = make_float4(mx[1], mx[2], __int_as_float(smin), __int_as_float(smax));
  }
}
// start of an odometry step: which slots solve, LM state from the previous relative motion (para_q, para_t)
__global__ void odom_guard_kernel(Dev d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.B) return;
  const FrameDesc& fd = d.desc[s];
  SlotOut& o = d.out[s];
  o.optimized = fd.active && fd.allow_opt;
  for (int k = 0; k < 2; ++k) {
    o.n_edge[k] = o.n_plane[k] = 0; o.lm_iters[k] = 0; o.lm_term[k] = 0;
    o.cost_initial[k] = o.cost_final[k] = 0.0; o.cand[k] = 0.0;
  }
  LmState& L = d.lm[s];
  for (int i = 0; i < 7; ++i) L.x[i] = L.xc[i] = fd.pose[i];
  L.done = !o.optimized; L.have_candidate = 0; L.iteration = 0;
  d.ticket[s] = 0;
}

__global__ void guard_kernel(Dev d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.B) return;
  const FrameDesc& fd = d.desc[s];
  SlotOut& o = d.out[s];
  // laserMapping.cpp:555
  // sharded map: the guard looks at the whole map = sum over ranks of the points each rank owns
  const int mc = d.shard_world > 1 ? d.shard_counts[s] : o.n_local[0];
  const int ms = d.shard_world > 1 ? d.shard_counts[d.B + s] : o.n_local[1];
  const int opt = fd.active && fd.allow_opt && mc > 10 && ms > 50;
  o.optimized = opt;
  for (int k = 0; k < 2; ++k) {
    o.n_edge[k] = o.n_plane[k] = 0; o.lm_iters[k] = 0; o.lm_term[k] = 0;
    o.cost_initial[k] = o.cost_final[k] = 0.0; o.cand[k] = 0.0;
  }
  LmState& L = d.lm[s];
  for (int i = 0; i < 7; ++i) L.x[i] = L.xc[i] = fd.pose[i];
  L.done = !opt; L.have_candidate = 0; L.iteration = 0;
  d.ticket[s] = 0;
}

__global__ void finish_pose_kernel(Dev d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.B) return;
  SlotOut& o = d.out[s];
  for (int i = 0; i < 7; ++i) o.pose[i] = o.optimized ? d.lm[s].x[i] : d.desc[s].pose[i];
}

// ----------------------------------------------------------------------------
// K2: map update -- insert (row I), per-voxel re-centroid of the valid cubes
// (row W), eviction of cubes that left the window (row B), as one sorted merge.
// delta sort key: [segment:7][window-relative cube:13][pending:1][payload:33]
// ----------------------------------------------------------------------------
__device__ __forceinline__ bool in_box(int ci, int cj, int ck, const int lo[3], const int hi[3]) {
  return ci >= lo[0] && ci <= hi[0] && cj >= lo[1] && cj <= hi[1] && ck >= lo[2] && ck <= hi[2];
}
__device__ __forceinline__ uint64_t delta_key(const Dev& d, int g, const FrameDesc& fd, int ci, int cj, int ck, uint32_t pend,
                                              uint64_t payload) {
  const uint32_t rel = (uint32_t)(((ci - fd.win_lo[0]) * kWinJ + (cj - fd.win_lo[1])) * kWinK + (ck - fd.win_lo[2]));
  const int P = d.delta_pbits;
  return ((uint64_t)g << (14 + P)) | ((uint64_t)rel << (1 + P)) | ((uint64_t)pend << P) | (payload & ((1ull << P) - 1ull));
}
__device__ __forceinline__ uint64_t voxel_payload(const Dev& d, int cls, float x, float y, float z, int ci, int cj,
                                                  int ck) {
  const float inv = d.inv_leaf[cls];
  const int vx = voxel_rel(x, ci, inv), vy = voxel_rel(y, cj, inv), vz = voxel_rel(z, ck, inv);
  const int b = d.vox_bits;
  if ((unsigned)vx >= (1u << b) || (unsigned)vy >= (1u << b) || (unsigned)vz >= (1u << b)) set_err(d, -4);
  return ((uint64_t)vz << (2 * b)) | ((uint64_t)vy << b) | (uint64_t)vx;  // compact (sort key only)
}

__global__ void delta_key_kernel(Dev d, int front, int n_delta, bool identity_pose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_delta) return;
  uint64_t key = kSentinel64;
  uint32_t val = 0;
  if (i >= front) {
    const int di = i - front;
    if (di < d.ds_off[d.G]) {
      const int g = find_seg(d.ds_off, d.G, di);
      const FrameDesc& fd = d.desc[seg_slot(d, g)];
      const float4 p = d.ds_pts[di];
      float w[3];
      if (identity_pose) { w[0] = p.x; w[1] = p.y; w[2] = p.z; }
      else xf_point(d.out[seg_slot(d, g)].pose, p.x, p.y, p.z, w);
      d.dl_pt[di] = make_float4(w[0], w[1], w[2], p.w);
      const int ci = cube_of((double)w[0]), cj = cube_of((double)w[1]), ck = cube_of((double)w[2]);
      bool mine = true;
      if (d.shard_world > 1) {
        // membership is decided per VOXEL (all points of a voxel go to the same ranks, so centroids
        // agree across ranks): keep it if its x-extent touches [lo - 1 m, hi + 1 m]
        const float leaf = 1.0f / d.inv_leaf[seg_cls(d, g)];
        const float vx = floorf(xfmul(w[0], d.inv_leaf[seg_cls(d, g)]));
        mine = (vx + 1.0f) * leaf >= d.shard_lo - 1.01f && vx * leaf <= d.shard_hi + 1.01f;
      }
      if (mine && fd.active && in_box(ci, cj, ck, fd.win_lo, fd.win_hi)) {  // laserMapping.cpp:753-755
        if (!cube_in_range(ci, cj, ck)) set_err(d, -4);
        if (in_box(ci, cj, ck, fd.val_lo, fd.val_hi))
          key = delta_key(d, g, fd, ci, cj, ck, 0, voxel_payload(d, seg_cls(d, g), w[0], w[1], w[2], ci, cj, ck));
        else
          key = delta_key(d, g, fd, ci, cj, ck, 1, (unsigned long long)(di - d.ds_off[g]));  // arrival offset inside this frame
        val = (uint32_t)di;
      }
    }
    d.vkey[i] = key;
    d.vval[i] = val;
  }
  // the front region is pre-filled with the sentinel by a memset and then
  // populated by pending_gather_kernel
}

// raw points already sitting in cubes that are valid now: they join this frame's re-filter
// grid (segments, cubes of the valid block): block (g, c) stages the raw points of cube c; block (g, 0) also
// publishes where they lie in the store -- they die in this frame's merge (d.dead_*: non-empty ranges in key order)
__global__ void pending_gather_kernel(Dev d, int cur) {
  const int g = blockIdx.x, c = blockIdx.y;
  const FrameDesc& fd = d.desc[seg_slot(d, g)];
  const int t = threadIdx.x;
  if (!fd.active) {
    if (c == 0 && t == 0) { d.dead_n[g] = 0; d.dead_cum[g * (kValidCubes + 1)] = 0; }
    return;
  }
  __shared__ int lo_s[kValidCubes], off_s[kValidCubes + 1];
  const uint64_t* keys = d.st_key[cur] + d.st_base[g];
  const int n = d.st_n[g];
  if (t < kValidCubes) {
    const int ci = fd.val_lo[0] + t / 15, cj = fd.val_lo[1] + (t / 3) % 5, ck = fd.val_lo[2] + t % 3;
    int lo = 0, len = 0;
    if (ci <= fd.val_hi[0] && cj <= fd.val_hi[1] && ck <= fd.val_hi[2]) {
      const uint32_t cube = pack_cube(ci, cj, ck);
      lo = lower_bound_u64(keys, n, store_key(cube, 1, 0));
      len = lower_bound_u64(keys, n, store_key(cube + 1, 0, 0)) - lo;
    }
    lo_s[t] = lo;
    off_s[t + 1] = len;
  }
  __syncthreads();
  if (t == 0) {
    off_s[0] = 0;
    for (int k = 0; k < kValidCubes; ++k) off_s[k + 1] += off_s[k];
    if (c == 0) {
      int m = 0;
      int* dl = d.dead_lo + g * kValidCubes;
      int* dc = d.dead_cum + g * (kValidCubes + 1);
      for (int k = 0; k < kValidCubes; ++k)
        if (off_s[k + 1] > off_s[k]) { dl[m] = lo_s[k]; dc[m] = off_s[k]; ++m; }
      dc[m] = off_s[kValidCubes];
      d.dead_n[g] = m;
    }
  }
  __syncthreads();
  const int len = off_s[c + 1] - off_s[c];
  const int ci = fd.val_lo[0] + c / 15, cj = fd.val_lo[1] + (c / 3) % 5, ck = fd.val_lo[2] + c % 3;
  for (int j = t; j < len; j += blockDim.x) {
    const int src = d.st_base[g] + lo_s[c] + j;
    const float4 p = d.st_pt[cur][src];
    const int pos = d.lp_off[g] + off_s[c] + j;
    d.vkey[pos] = delta_key(d, g, fd, ci, cj, ck, 0, voxel_payload(d, seg_cls(d, g), p.x, p.y, p.z, ci, cj, ck));
    d.vval[pos] = 0x80000000u | (uint32_t)src;
  }
}

__device__ __forceinline__ float4 delta_point(const Dev& d, int cur, uint32_t v) {
  return (v & 0x80000000u) ? d.st_pt[cur][v & 0x7FFFFFFFu] : d.dl_pt[v];
}

// one thread per run of equal delta keys: re-centroid (old centroid first, then
// raw points in arrival order -- the order a stable sort gives pcl::VoxelGrid)
__global__ void delta_reduce_kernel(Dev d, int cur, int n_delta) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_delta || !d.flag[p]) return;
  const uint64_t key = d.vkey2[p];
  if (key == kSentinel64) return;
  const int r = (int)d.scan[p];
  const int P = d.delta_pbits, vb = d.vox_bits;
  const int g = (int)(key >> (14 + P));
  const FrameDesc& fd = d.desc[seg_slot(d, g)];
  const int rel = (int)((key >> (1 + P)) & 0x1FFF);
  const uint32_t pend = (uint32_t)((key >> P) & 1);
  const uint64_t pl = key & ((1ull << P) - 1ull);
  // store payload: arrival number (raw) or the 3 x 11-bit voxel coordinate inside the cube (filtered)
  const uint64_t vm = (1ull << vb) - 1ull;
  const uint64_t store_pl = pend ? fd.seq_base[seg_cls(d, g)] + pl
                                 : (((pl >> (2 * vb)) & vm) << 22) | (((pl >> vb) & vm) << 11) | (pl & vm);
  const int ci = fd.win_lo[0] + rel / (kWinJ * kWinK), cj = fd.win_lo[1] + (rel / kWinK) % kWinJ,
            ck = fd.win_lo[2] + rel % kWinK;
  const uint64_t wkey = store_key(pack_cube(ci, cj, ck), pend, store_pl);
  if (pend) {  // raw point for a cube outside the valid block: stays raw
    d.ins_key[r] = wkey;
    d.ins_pt[r] = delta_point(d, cur, d.vval2[p]);
    return;
  }
  const uint64_t* keys = d.st_key[cur] + d.st_base[g];
  const int n = d.st_n[g];
  const int pos = lower_bound_u64(keys, n, wkey);
  const bool exists = pos < n && keys[pos] == wkey;
  float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
  int cnt = 0;
  if (exists) {
    const float4 o = d.st_pt[cur][d.st_base[g] + pos];
    sx = xfadd(sx, o.x); sy = xfadd(sy, o.y); sz = xfadd(sz, o.z); si = xfadd(si, o.w);
    cnt = 1;
  }
  for (int e = p; e < n_delta && d.vkey2[e] == key; ++e) {
    const float4 q = delta_point(d, cur, d.vval2[e]);
    sx = xfadd(sx, q.x); sy = xfadd(sy, q.y); sz = xfadd(sz, q.z); si = xfadd(si, q.w);
    ++cnt;
  }
  const float c = (float)cnt;
  const float4 cen = make_float4(xfdiv(sx, c), xfdiv(sy, c), xfdiv(sz, c), xfdiv(si, c));
  if (fd.idx_flags & 2) {  // the cell index of this slot follows the map (it is not rebuilt next frame)
    const uint32_t vx = (uint32_t)(pl & vm), vy = (uint32_t)((pl >> vb) & vm), vz = (uint32_t)((pl >> (2 * vb)) & vm);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (exists) o = d.st_pt[cur][d.st_base[g] + pos];
    idx_apply(d, g, fd, tag_filtered(fd, ci, cj, ck, vz, vy, vx), exists, o, cen);
  }
  if (exists) {
    // key unchanged: the new centroid replaces the entry in the NEXT store buffer (upd_apply_kernel, after
    // the merge copy).  The current buffer is never written, so a frame that fails later -- capacity, range --
    // leaves the map exactly as it was.
    d.upd_pos[r] = d.so_off[g] + pos;
    d.ins_pt[r] = cen;
  } else {
    d.ins_key[r] = wkey;
    d.ins_pt[r] = cen;
  }
}
// re-centroided entries: same key, new point, at the position merge_old_kernel gave the old entry
__global__ void upd_apply_kernel(Dev d, int cur, int n_runs_max) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs_max) return;
  const int i = d.upd_pos[r];
  if (i < 0 || !d.aflag[i]) return;
  const int g = find_seg(d.so_off, d.G, i);
  const int pos = (int)(d.ascan[i] - d.ascan[d.so_off[g]]) + (int)d.flag[i] - 1;
  if (pos >= d.st_cap[g]) return;  // merge_old_kernel raised the error
  d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.ins_pt[r];
}

__global__ void ins_flag_kernel(const uint64_t* __restrict__ ins_key, uint32_t* __restrict__ flag, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  flag[r] = (r < n) && ins_key[r] != kSentinel64;
}
__global__ void ins_compact_kernel(Dev d, int n) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || d.ins_key[r] == kSentinel64) return;
  const int o = (int)d.ascan[r];
  d.ins_ckey[o] = d.ins_key[r];
  d.ins_cpt[o] = d.ins_pt[r];
}
// run_off[g] = compacted-insert offset of segment g (reuses d.run_off)
__global__ void ins_off_kernel(Dev d, int n_delta) {
  const int g = threadIdx.x + blockIdx.x * blockDim.x;
  if (g > d.G) return;
  const int pos = lower_bound_u64(d.vkey2, n_delta, (uint64_t)g << (14 + d.delta_pbits));
  const int run = (int)d.scan[pos];  // scan has n_delta+1 entries
  d.run_off[g] = (int)d.ascan[run];  // ascan here = exclusive scan of insert flags over runs
}

// d.vval2 holds, per insert of a segment (sorted), the old position it goes in front of
__device__ __forceinline__ int inserts_up_to(const uint32_t* __restrict__ lb, int a, int b, uint32_t l) {  // first j in [a, b) with lb[j] > l
  while (a < b) {
    const int m = (a + b) >> 1;
    if (lb[m] <= l) a = m + 1; else b = m;
  }
  return a;
}
__device__ __forceinline__ bool entry_dead(const FrameDesc& fd, uint64_t key) {
  int ci, cj, ck;
  unpack_cube(key_cube(key), ci, cj, ck);
  if (!in_box(ci, cj, ck, fd.win_lo, fd.win_hi)) return true;           // cube left the window (:346-347 ...)
  return key_pending(key) && in_box(ci, cj, ck, fd.val_lo, fd.val_hi);  // raw point merged by this re-filter
}
// The merge of the old store with the frame's sorted inserts, without a search per old entry: every insert marks
// the old position it goes in front of (ins_mark_kernel: one binary search per INSERT), one prefix sum over
// "alive + inserts in front" then gives every survivor and every insert its place.  The index space (d.so_off) has
// one extra position per segment for the inserts that go behind its last entry.
__global__ void ins_mark_kernel(Dev d, int cur, int n_max, bool count) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_max || j >= d.run_off[d.G]) return;
  const int g = find_seg(d.run_off, d.G, j);
  const int lb = lower_bound_u64(d.st_key[cur] + d.st_base[g], d.st_n[g], d.ins_ckey[j]);
  d.vval2[j] = (uint32_t)lb;
  if (count) atomicAdd(d.flag + d.so_off[g] + lb, 1u);
}
// d.aflag = the entry survives; d.flag += that (d.flag holds the inserts in front of the position)
__global__ void alive_flag_kernel(Dev d, int cur, int total_store, bool any_dead) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > total_store) return;
  uint32_t f = 0;
  if (i < total_store) {
    const int g = find_seg(d.so_off, d.G, i);
    const int l = i - d.so_off[g];
    if (l < d.st_n[g]) {
      const FrameDesc& fd = d.desc[seg_slot(d, g)];
      f = (any_dead && fd.active) ? !entry_dead(fd, d.st_key[cur][d.st_base[g] + l]) : 1u;  // (nothing dies unless the valid block moved)
    }
    d.flag[i] += f;
  }
  d.aflag[i] = f;
}
__global__ void merge_old_kernel(Dev d, int cur, int total_store) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_store || !d.aflag[i]) return;
  const int g = find_seg(d.so_off, d.G, i);
  const int src = d.st_base[g] + (i - d.so_off[g]);
  const int pos = (int)(d.ascan[i] - d.ascan[d.so_off[g]]) + (int)d.flag[i] - 1;  // survivors + inserts before it, inserts in front of it
  if (pos >= d.st_cap[g]) { set_err(d, -3); return; }
  d.st_key[cur ^ 1][d.st_base[g] + pos] = d.st_key[cur][src];
  d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.st_pt[cur][src];
}
__global__ void merge_new_kernel(Dev d, int cur, int n_max) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_max || j >= d.run_off[d.G]) return;
  const int g = find_seg(d.run_off, d.G, j);
  const uint32_t lb = d.vval2[j];
  // inserts that go in front of the same old entry are consecutive (sorted by key); a cube whose raw points have just
  // been filtered puts thousands in front of one entry
  const int first = lb == 0u ? d.run_off[g] : inserts_up_to(d.vval2, d.run_off[g], j, lb - 1u);
  const int pos = (int)(d.ascan[d.so_off[g] + (int)lb] - d.ascan[d.so_off[g]]) + (j - first);
  if (pos >= d.st_cap[g]) { set_err(d, -3); return; }
  d.st_key[cur ^ 1][d.st_base[g] + pos] = d.ins_ckey[j];
  d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.ins_cpt[j];
}
// ---- the same merge when the window did not shift: the only entries that can die are the raw points of the valid
// cubes (<= 75 contiguous key ranges per segment, published by pending_gather_kernel; none when the valid block did not
// move).  Every survivor then moves by (inserts that go in front of it or of an earlier entry) - (dead entries before
// it): two short binary searches, no flags and no prefix sum over the store.
// dead entries of segment g in front of store position l; *dead = position l itself dies
__device__ __forceinline__ int dead_before(const Dev& d, int g, int l, bool* dead) {
  const int* lo = d.dead_lo + g * kValidCubes;
  const int* cum = d.dead_cum + g * (kValidCubes + 1);
  int a = 0, b = d.dead_n[g];  // number of ranges that start at or before l
  while (a < b) {
    const int m = (a + b) >> 1;
    if (lo[m] <= l) a = m + 1; else b = m;
  }
  *dead = false;
  if (a == 0) return 0;
  const int len = cum[a] - cum[a - 1], in = l - lo[a - 1];
  *dead = in < len;
  return cum[a - 1] + min(in, len);
}
__global__ void shift_old_kernel(Dev d, int cur, int total_store, bool deaths) {
  // per block: its segment(s) and the inserts that can fall among its entries (when it lies inside one segment)
  __shared__ int seg[2], win[2];
  const int i0 = blockIdx.x * blockDim.x, i = i0 + threadIdx.x, i1 = min(i0 + (int)blockDim.x, total_store) - 1;
  if (threadIdx.x < 2) {
    const int g = find_seg(d.so_off, d.G, threadIdx.x == 0 ? i0 : i1);
    seg[threadIdx.x] = g;
    const int l = (threadIdx.x == 0 ? i0 - 1 : i1) - d.so_off[g];  // inserts in front of entries before the block / up to its last entry
    win[threadIdx.x] = l < 0 ? d.run_off[g] : inserts_up_to(d.vval2, d.run_off[g], d.run_off[g + 1], (uint32_t)l);
  }
  __syncthreads();
  if (i >= total_store) return;
  const int g0 = seg[0], g1 = seg[1];
  const int g = g0 == g1 ? g0 : find_seg(d.so_off, d.G, i);
  const int l = i - d.so_off[g];
  if (l >= d.st_n[g]) return;  // (the extra position behind the segment's last entry)
  const int a = g0 == g1 ? win[0] : d.run_off[g], b = g0 == g1 ? win[1] : d.run_off[g + 1];
  int pos = l + inserts_up_to(d.vval2, a, b, (uint32_t)l) - d.run_off[g];
  if (deaths) {
    bool dead;
    pos -= dead_before(d, g, l, &dead);
    if (dead) return;
  }
  if (pos >= d.st_cap[g]) { set_err(d, -3); return; }
  const int src = d.st_base[g] + l;
  d.st_key[cur ^ 1][d.st_base[g] + pos] = d.st_key[cur][src];
  d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.st_pt[cur][src];
}
// inserts and re-centroided entries of the same case; also the new sizes
__global__ void shift_new_kernel(Dev d, int cur, int n_max, bool deaths) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  bool dead;
  if (j < d.G) {
    const int g = j;
    const int n = d.st_n[g] + d.run_off[g + 1] - d.run_off[g] - (deaths ? d.dead_cum[g * (kValidCubes + 1) + d.dead_n[g]] : 0);
    if (n > d.st_cap[g]) set_err(d, -3);
    d.st_n_new[g] = min(n, d.st_cap[g]);
    d.out[seg_slot(d, g)].n_store[seg_cls(d, g)] = n;
  }
  if (j < n_max && j < d.run_off[d.G]) {
    const int g = find_seg(d.run_off, d.G, j);
    const int lb = (int)d.vval2[j];
    const int pos = lb + (j - d.run_off[g]) - (deaths ? dead_before(d, g, lb, &dead) : 0);
    if (pos < d.st_cap[g]) {
      d.st_key[cur ^ 1][d.st_base[g] + pos] = d.ins_ckey[j];
      d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.ins_cpt[j];
    }
  }
  if (j < n_max) {
    const int i = d.upd_pos[j];
    if (i >= 0) {
      const int g = find_seg(d.so_off, d.G, i);
      const int l = i - d.so_off[g];
      const int pos = l + inserts_up_to(d.vval2, d.run_off[g], d.run_off[g + 1], (uint32_t)l) - d.run_off[g] -
                      (deaths ? dead_before(d, g, l, &dead) : 0);
      if (pos < d.st_cap[g]) d.st_pt[cur ^ 1][d.st_base[g] + pos] = d.ins_pt[j];
    }
  }
}
__global__ void store_count_kernel(Dev d) {
  const int g = threadIdx.x + blockIdx.x * blockDim.x;
  if (g >= d.G) return;
  const int n = (int)(d.ascan[d.so_off[g + 1]] - d.ascan[d.so_off[g]]);  // survivors + inserts
  if (n > d.st_cap[g]) set_err(d, -3);
  d.st_n_new[g] = min(n, d.st_cap[g]);
  d.out[seg_slot(d, g)].n_store[seg_cls(d, g)] = n;
}

// ----------------------------------------------------------------------------
// K7 and utilities
// ----------------------------------------------------------------------------
__global__ void transform_cloud_kernel(const double* __restrict__ pose7, const float4* __restrict__ in,
                                       float4* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __shared__ double pose[7];
  if (threadIdx.x < 7) pose[threadIdx.x] = pose7[threadIdx.x];
  __syncthreads();
  const float4 p = in[i];
  float w[3];
  xf_point(pose, p.x, p.y, p.z, w);
  out[i] = make_float4(w[0], w[1], w[2], p.w);
}
__global__ void gather_local_kernel(Dev d, int cur, int g, float4* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= d.loc_off[g * (kCols + 1) + kCols]) return;
  out[l] = d.st_pt[cur][d.st_base[g] + local_to_store(d, g, l)];
}

// Row X, /laser_cloud_surround (laserMapping.cpp:807-815): corner then surf cloud of every valid
// cube, cubes in gather order. One block: 150 (cube, class) ranges by binary search, a scan, a copy.
__global__ void surround_kernel(Dev d, int cur, int slot, float4* __restrict__ out, int cap, int* __restrict__ n_out) {
  __shared__ int lo_s[2 * kValidCubes], off_s[2 * kValidCubes + 1];
  const FrameDesc& fd = d.desc[slot];
  const int t = threadIdx.x;
  if (t < 2 * kValidCubes) {
    const int c = t >> 1, cls = t & 1, g = cls * d.B + slot;
    const int ci = fd.val_lo[0] + c / 15, cj = fd.val_lo[1] + (c / 3) % 5, ck = fd.val_lo[2] + c % 3;
    int lo = 0, len = 0;
    if (ci <= fd.val_hi[0] && cj <= fd.val_hi[1] && ck <= fd.val_hi[2]) {
      const uint64_t* keys = d.st_key[cur] + d.st_base[g];
      const uint32_t cube = pack_cube(ci, cj, ck);
      lo = lower_bound_u64(keys, d.st_n[g], store_key(cube, 0, 0));
      len = lower_bound_u64(keys, d.st_n[g], store_key(cube + 1, 0, 0)) - lo;
    }
    lo_s[t] = lo;
    off_s[t + 1] = len;
  }
  __syncthreads();
  if (t == 0) {
    off_s[0] = 0;
    for (int k = 0; k < 2 * kValidCubes; ++k) off_s[k + 1] += off_s[k];
    *n_out = off_s[2 * kValidCubes];
  }
  __syncthreads();
  for (int r = 0; r < 2 * kValidCubes; ++r) {
    const int g = (r & 1) * d.B + slot, len = off_s[r + 1] - off_s[r];
    for (int j = t; j < len; j += blockDim.x)
      if (off_s[r] + j < cap) out[off_s[r] + j] = d.st_pt[cur][d.st_base[g] + lo_s[r] + j];
  }
}

// debug: guard bands around every allocation must still hold their pattern
__global__ void guard_check_kernel(const GuardDesc* __restrict__ g, int* __restrict__ bad) {
  const GuardDesc gd = g[blockIdx.x];
  const uint32_t* p = blockIdx.y ? gd.back : gd.front;
  int n = 0;
  for (unsigned i = threadIdx.x; i < gd.words; i += blockDim.x) n += p[i] != 0xA5A5A5A5u;
  if (n) atomicAdd(bad, n);
}
int launch_guard_check(const GuardDesc* g, int n, int* bad, cudaStream_t s) {
  if (n <= 0) return 0;
  guard_check_kernel<<<dim3(n, 2), 256, 0, s>>>(g, bad);
  return 1;
}

// ----------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

size_t cub_temp_bytes(int cap_sort, int cap_lp) {
  size_t a = 0, b = 0, c = 0, e = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (uint64_t*)nullptr, (uint64_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, cap_sort);
  cub::DeviceRadixSort::SortPairs(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, cap_lp);
  cub::DeviceScan::ExclusiveSum(nullptr, c, (uint32_t*)nullptr, (uint32_t*)nullptr, cap_sort + 1);
  cub::DeviceScan::ExclusiveSum(nullptr, e, (uint32_t*)nullptr, (uint32_t*)nullptr, cap_lp + 1);
  return std::max(std::max(a, b), std::max(c, e)) + 256;
}

int launch_voxel_bbox(const Dev& d, int n, int longest_segment, cudaStream_t s) {
  int k = 0;
  if (n > 0) {
    vox_bbox_init_kernel<<<cdiv(6 * d.G, 256), 256, 0, s>>>(d); ++k;
    vox_bbox_kernel<<<dim3(cdiv(std::max(longest_segment, 1), kBoxPts), d.G), 256, 0, s>>>(d); ++k;
  }
  return k;
}
// key_bits: width of PCL's linear voxel index over all segments (<= 31), known from the boxes;
// the segment number sits right above it, so the sort covers key_bits + log2(segments) bits
int launch_voxel_filter(const Dev& d, int n, int key_bits, cudaStream_t s) {
  int k = 0;
  int gbits = 1;
  while ((1 << gbits) < d.G) ++gbits;
  const bool narrow = key_bits + gbits <= 32;
  uint32_t *k32 = reinterpret_cast<uint32_t*>(d.vkey), *k32b = reinterpret_cast<uint32_t*>(d.vkey2);
  if (n > 0) {
    size_t tb = d.cub_tmp_bytes;
    if (narrow) {
      vox_key_kernel<true><<<cdiv(n, 256), 256, 0, s>>>(d, n, key_bits); ++k;
      cub::DeviceRadixSort::SortPairs(d.cub_tmp, tb, k32, k32b, d.vval, d.vval2, n, 0, key_bits + gbits, s);
    } else {
      vox_key_kernel<false><<<cdiv(n, 256), 256, 0, s>>>(d, n, key_bits); ++k;
      cub::DeviceRadixSort::SortPairs(d.cub_tmp, tb, d.vkey, d.vkey2, d.vval, d.vval2, n, 0, key_bits + gbits, s);
    }
  }
  if (narrow) head_flag_kernel<uint32_t><<<cdiv(n + 1, 256), 256, 0, s>>>(k32b, d.flag, n);
  else head_flag_kernel<uint64_t><<<cdiv(n + 1, 256), 256, 0, s>>>(d.vkey2, d.flag, n);
  ++k;
  size_t tb = d.cub_tmp_bytes;
  cub::DeviceScan::ExclusiveSum(d.cub_tmp, tb, d.flag, d.scan, n + 1, s);
  if (n > 0) {
    vox_heads_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, n); ++k;
    vox_centroid_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, n); ++k;  // n >= number of voxels; surplus threads exit
  }
  ds_off_kernel<<<cdiv(d.G + 1, 128), 128, 0, s>>>(d, n); ++k;
  return k;
}

int launch_local_ranges(const Dev& d, int cur, cudaStream_t s) {
  range_kernel<<<d.G, 32, 0, s>>>(d, cur);
  return 1;
}
// Bulk (re)build of the cell index of the n_seg segments listed in d.idx_list (their new table masks follow the
// list in the same device array); total_points = their local-map points (d.idx_poff).  Needs launch_local_ranges.
int launch_index_rebuild(const Dev& d, int cur, int n_seg, int total_points, cudaStream_t s) {
  if (n_seg <= 0) return 0;
  const int* new_mask = d.idx_list + d.G;
  idx_reset_kernel<<<dim3(64, n_seg), 256, 0, s>>>(d, new_mask);
  idx_arm_kernel<<<cdiv(n_seg, 128), 128, 0, s>>>(d, new_mask, n_seg);
  int k = 2;
  if (total_points > 0) {
    // (the sort buffers of the scan voxel filter and of the map update are idle at this point of a frame)
    uint32_t *ka = reinterpret_cast<uint32_t*>(d.vkey), *kb = reinterpret_cast<uint32_t*>(d.vkey2);
    idx_key_kernel<<<cdiv(total_points, 256), 256, 0, s>>>(d, cur, n_seg, total_points, ka); ++k;
    int sbits = 1;
    while ((1 << sbits) <= n_seg) ++sbits;
    size_t tb = d.cub_tmp_bytes;
    cub::DeviceRadixSort::SortPairs(d.cub_tmp, tb, ka, kb, d.vval, d.vval2, total_points, 0, 24 + sbits, s);
    head_flag_kernel<uint32_t><<<cdiv(total_points + 1, 256), 256, 0, s>>>(kb, d.flag, total_points); ++k;
    tb = d.cub_tmp_bytes;
    cub::DeviceScan::ExclusiveSum(d.cub_tmp, tb, d.flag, d.scan, total_points + 1, s);
    idx_segs_kernel<<<cdiv(n_seg + 1, 128), 128, 0, s>>>(d, n_seg, total_points, kb); ++k;
    idx_fill_kernel<<<cdiv(total_points, 256), 256, 0, s>>>(d, cur, n_seg, total_points, kb); ++k;
  }
  return k;
}
int launch_guard(const Dev& d, cudaStream_t s) {
  guard_kernel<<<cdiv(d.B, 64), 64, 0, s>>>(d);
  return 1;
}
// One association (rows P..Q of one outer iteration) for every optimised slot: K4a with `knn_blocks`
// persistent blocks sharing a device-wide tile ticket (the caller zeroes d.knn_ticket), then K4b.
int launch_associate(const Dev& d, int outer, int cur, int knn_blocks, int fit_blocks, int n_ds, bool trace, cudaStream_t s) {
  if (knn_blocks <= 0 || fit_blocks <= 0) return 0;
  dim3 gb(fit_blocks, d.B);
#if S2M_KNN_GROUP
  constexpr size_t smem = sizeof(KnnGroupSmem) * (kTile / 32);
  static bool armed = false;  // (one device per process)
  if (!armed) {
    if (cudaFuncSetAttribute(knn_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    armed = true;
  }
  knn_group_kernel<<<knn_blocks, kTile, smem, s>>>(d, outer, n_ds);
#else
  knn_kernel<false><<<knn_blocks, kTile, 0, s>>>(d, outer);
#endif
  if (trace) fit_kernel<true><<<gb, kTile, 0, s>>>(d, outer, cur);
  else fit_kernel<false><<<gb, kTile, 0, s>>>(d, outer, cur);
  return 2;
}
// profiling: the map points in every query's 27 cells (the roofline's byte count)
int launch_count27(const Dev& d, int n_ds, cudaStream_t s) {
  if (n_ds <= 0) return 0;
  count27_kernel<<<cdiv(n_ds, 256), 256, 0, s>>>(d, n_ds);
  return 1;
}
// Once per frame, before the first association: the scan points bucketed by (segment, 2 m block) for the grouped search.
int launch_query_sort(const Dev& d, int n_ds, cudaStream_t s) {
  if (n_ds <= 0) return 0;
  int k = 0;
#if S2M_KNN_GROUP
  qgroup_kernel<<<d.G, 1024, 0, s>>>(d);
  k += 1;
#endif
  return k;
}
// The LM solve that follows it (<= 4 iterations) in one launch: a cluster of four CTAs per slot.
int launch_solve(const Dev& d, int outer, bool from_units, cudaStream_t s) {
  solve_kernel<<<kSolveCtas * d.B, kSolveThreads, 0, s>>>(d, outer, from_units ? 1 : 0);
  return 1;
}
int launch_reduce_units(const Dev& d, cudaStream_t s) {
  reduce_units_kernel<<<d.B, kPartial, 0, s>>>(d);
  return 1;
}
int launch_lm_shard(const Dev& d, int outer, int after, cudaStream_t s) {
  lm_shard_kernel<<<d.B, kTile, 0, s>>>(d, outer, after);
  return 1;
}
int launch_evaluate(const Dev& d, int outer, int blocks_per_slot, cudaStream_t s) {
  if (blocks_per_slot <= 0) return 0;
  dim3 grid(blocks_per_slot, d.B);
  evaluate_kernel<<<grid, kTile, 0, s>>>(d, outer);
  return 1;
}
int launch_odom_sort(const Dev& d, int n, void* tmp, size_t tmp_bytes, cudaStream_t s) {
  if (n <= 0) return 0;
  odom_sort_key_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, n);
  int bits = 1;
  while ((1 << bits) < 2 * d.B) ++bits;
  cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d.od_key, d.od_key2, d.od_val, d.od_val2, n, 0, 30 + bits, s);
  odom_gather_sorted_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, n);
  odom_ring_init_kernel<<<cdiv(2 * d.B * (kOdSid + 1), 256), 256, 0, s>>>(d);
  odom_ring_mark_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, n);
  odom_ring_scan_kernel<<<cdiv(2 * d.B, 64), 64, 0, s>>>(d);
  return 5;
}
size_t odom_sort_temp_bytes(const Dev& d, int n) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, d.od_key, d.od_key2, d.od_val, d.od_val2, n, 0, 40, (cudaStream_t)0);
  return tb;
}
int launch_odom_meta(const Dev& d, int nchunks, cudaStream_t s) {
  if (nchunks <= 0) return 0;
  odom_meta_kernel<<<cdiv(nchunks, 8), 256, 0, s>>>(d, nchunks);
  return 1;
}
int launch_odom_guard(const Dev& d, cudaStream_t s) {
  odom_guard_kernel<<<cdiv(d.B, 64), 64, 0, s>>>(d);
  return 1;
}
int launch_odom_associate(const Dev& d, int outer, int tiles, int fallback_blocks, bool trace, cudaStream_t s) {
  if (tiles <= 0) return 0;
  dim3 grid(tiles, d.B);
  cudaMemsetAsync(d.od_fb_cnt, 0, sizeof(int), s);
  odom_search_kernel<<<grid, kTile, 0, s>>>(d);
  odom_fallback_kernel<<<fallback_blocks, kTile, 0, s>>>(d);
  if (trace) odom_fit_kernel<true><<<grid, kTile, 0, s>>>(d, outer);
  else odom_fit_kernel<false><<<grid, kTile, 0, s>>>(d, outer);
  return 3;
}
int launch_finish_pose(const Dev& d, cudaStream_t s) {
  finish_pose_kernel<<<cdiv(d.B, 64), 64, 0, s>>>(d);
  return 1;
}

// total_lp: points of the local maps (d.lp_off spacing, the pending-point staging area);
// total_store: entries of the whole stores (d.so_off spacing, the merge)
// check_pending: some valid block moved (raw points of its cubes are absorbed); window_shift: some window moved or a
// store was replaced (entries anywhere may have to go) -- implies check_pending
int launch_map_update(const Dev& d, int cur, int n_ds, int total_lp, int total_store, bool check_pending,
                      bool window_shift, bool identity_pose, cudaStream_t s) {
  const int total_in = n_ds;  // exact number of down-sampled points (host read it back)
  int k = 0;
  const int front = check_pending ? total_lp : 0;
  const int n_delta = front + total_in;
  if (front > 0) {
    cudaMemsetAsync(d.vkey, 0xFF, sizeof(uint64_t) * (size_t)front, s);
    cudaMemsetAsync(d.vval, 0, sizeof(uint32_t) * (size_t)front, s);
    pending_gather_kernel<<<dim3(d.G, kValidCubes), 128, 0, s>>>(d, cur); ++k;
  }
  if (n_delta > 0) {
    delta_key_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, front, n_delta, identity_pose); ++k;
    size_t tb = d.cub_tmp_bytes;
    int gbits = 1;
    while ((1 << gbits) < d.G) ++gbits;
    // the sentinel is all ones inside the sorted bit range too, so it stays at the end
    cub::DeviceRadixSort::SortPairs(d.cub_tmp, tb, d.vkey, d.vkey2, d.vval, d.vval2, n_delta, 0, 14 + d.delta_pbits + gbits, s);
  }
  head_flag_kernel<uint64_t><<<cdiv(n_delta + 1, 256), 256, 0, s>>>(d.vkey2, d.flag, n_delta); ++k;
  size_t tb = d.cub_tmp_bytes;
  cub::DeviceScan::ExclusiveSum(d.cub_tmp, tb, d.flag, d.scan, n_delta + 1, s);
  cudaMemsetAsync(d.ins_key, 0xFF, sizeof(uint64_t) * (size_t)(n_delta + 1), s);
  cudaMemsetAsync(d.upd_pos, 0xFF, sizeof(int) * (size_t)(n_delta + 1), s);
  if (n_delta > 0) { delta_reduce_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, cur, n_delta); ++k; }
  // compact the inserts (runs that created a new store entry)
  ins_flag_kernel<<<cdiv(n_delta + 1, 256), 256, 0, s>>>(d.ins_key, d.aflag, n_delta); ++k;
  tb = d.cub_tmp_bytes;
  cub::DeviceScan::ExclusiveSum(d.cub_tmp, tb, d.aflag, d.ascan, n_delta + 1, s);
  if (n_delta > 0) { ins_compact_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, n_delta); ++k; }
  ins_off_kernel<<<cdiv(d.G + 1, 128), 128, 0, s>>>(d, n_delta); ++k;
  if (!window_shift) {  // only the raw points of the valid cubes can die: old entries shift by two counts
    if (n_delta > 0) { ins_mark_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, cur, n_delta, false); ++k; }
    const bool deaths = front > 0;  // (pending_gather_kernel ran and published the ranges)
    if (total_store > 0) { shift_old_kernel<<<cdiv(total_store, 256), 256, 0, s>>>(d, cur, total_store, deaths); ++k; }
    shift_new_kernel<<<cdiv(max(n_delta, d.G), 256), 256, 0, s>>>(d, cur, n_delta, deaths); ++k;
    return k;
  }
  // survivors of the old store and the places of the inserts between them
  cudaMemsetAsync(d.flag, 0, sizeof(uint32_t) * (size_t)(total_store + 1), s);
  if (n_delta > 0) { ins_mark_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, cur, n_delta, true); ++k; }
  alive_flag_kernel<<<cdiv(total_store + 1, 256), 256, 0, s>>>(d, cur, total_store, check_pending); ++k;
  tb = d.cub_tmp_bytes;
  cub::DeviceScan::ExclusiveSum(d.cub_tmp, tb, d.flag, d.ascan, total_store + 1, s);
  if (total_store > 0) { merge_old_kernel<<<cdiv(total_store, 256), 256, 0, s>>>(d, cur, total_store); ++k; }
  if (n_delta > 0) { merge_new_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, cur, n_delta); ++k; }
  if (n_delta > 0 && total_store > 0) { upd_apply_kernel<<<cdiv(n_delta, 256), 256, 0, s>>>(d, cur, n_delta); ++k; }
  store_count_kernel<<<cdiv(d.G, 128), 128, 0, s>>>(d); ++k;
  return k;
}

int launch_knn_debug(const Dev& d, int cur, int slot, int cls, const float* d_q, int n, int32_t* d_idx, float* d_d2,
                     cudaStream_t s) {
  if (n <= 0) return 0;
  constexpr size_t smem = sizeof(KnnGroupSmem) * (kTile / 32);
  if (cudaFuncSetAttribute(knn_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
  knn_debug_kernel<<<cdiv(n, kTile), kTile, smem, s>>>(d, cur, cls * d.B + slot, d_q, n, d_idx, d_d2);
  return 1;
}
int launch_transform_cloud(const double* d_pose7, const float4* in, float4* out, int n, cudaStream_t s) {
  if (n <= 0) return 0;
  transform_cloud_kernel<<<cdiv(n, 256), 256, 0, s>>>(d_pose7, in, out, n);
  return 1;
}
int launch_surround(const Dev& d, int cur, int slot, float4* out, int cap, int* n_out, cudaStream_t s) {
  surround_kernel<<<1, 512, 0, s>>>(d, cur, slot, out, cap, n_out);
  return 1;
}
int launch_gather_local(const Dev& d, int cur, int g, float4* out, cudaStream_t s) {
  // upper bound on the local size is the store size; the kernel bounds itself
  const int n = d.cap_lp;
  gather_local_kernel<<<cdiv(n, 256), 256, 0, s>>>(d, cur, g, out);
  return 1;
}

}  // namespace s2m
