// s2m_fx.cu -- feature extraction for a batch of lidar sweeps on the device.
//
// /root/reference/src/scanRegistration.cpp:116-454 (laserCloudHandler):
// clean-up, ring number + relative time per point, ring-major re-ordering, 11-tap curvature,
// six sectors per ring with 2 sharp / 20 less-sharp / 4 flat picks and neighbour suppression,
// and the per-ring 0.2 m VoxelGrid of the less-flat points.  Results are bit-identical to the
// CPU restatement oracle/scan_registration.cpp (same arithmetic order, no FMA contraction).
//
// Pipeline per call (B sweeps, one stream):
//   valid flags -> scan -> compaction -> per-sweep start/end orientation
//   ring numbers (+ first point past the half turn, atomicMin) -> intensities and sort keys
//   stable radix sort by (sweep, ring) -> ring offsets -> ring-major cloud
//   curvature + "gap to the previous point" flags
//   one WARP per (sweep, ring): every pick is a warp-wide arg-max / arg-min of (curvature, index) over
//   the not yet suppressed points of the sector; the warp compacts the less-flat points and their box
//   prefix sums of the pick counts -> packed sharp / less-sharp / flat clouds
//   per-ring voxel keys -> stable radix sort -> one thread per voxel sums its run in order
#include <cub/cub.cuh>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/s2m.h"
#include "s2m_math.cuh"

namespace s2m {
namespace fx {

constexpr int kRings = 64;        // ring slots per sweep (N_SCANS <= 64, scanRegistration.cpp:470-480)
constexpr int kSectors = 6;       // :295
constexpr int kSecPerScan = kRings * kSectors;
constexpr int kLess = 20, kSharp = 2, kFlat = 4;  // :317-327, :366

struct Dev {
  int B, sensor;
  float thres2;
  const float* xyz;        // [n][3] raw sweeps, packed
  int* in_off;             // [B+1]
  uint32_t *flag, *fscan;  // [cap+1]
  int* vidx;               // [cap] kept-after-cleaning -> raw index
  int* voff;               // [B+1]
  float* ori;              // [B][2] start / end orientation (:143-156)
  int* jstar;              // [B] first cleaned point that saw ori - start > pi (:232-235)
  int* ok;                 // [B] sweep has >= 12 cleaned points
  int* ring;               // [cap]
  float4* pts;             // [cap] cleaned order, intensity filled
  uint32_t *key, *key2, *val, *val2;
  int* ring_off;           // [B*64+1] into the ring-major cloud
  float4* cloud;           // [cap] ring-major (= the "full" output)
  float* curv;             // [cap]
  unsigned char* gapf;     // [cap] distance^2 to the previous point of the cloud > 0.05
  unsigned char* state;    // [cap] pick state of every cloud point (select_kernel)
  int *n_sharp, *n_less, *n_flat;        // [B*384+1]
  int *o_sharp, *o_less, *o_flat;        // exclusive scans
  int *i_sharp, *i_less, *i_flat;        // pick lists (cloud indices)
  float4 *out_sharp, *out_less, *out_flat;
  int *lf_cnt, *lf_off;    // [B*64+1]
  float* rbox;             // [B*64][6]
  float4* lf_stage;        // [cap] less-flat points staged at their ring's offset
  float4* lf_pts;          // [cap] packed
  unsigned long long *vkey, *vkey2;
  uint32_t *vval, *vval2;
  float4* out_lf;          // [cap] voxel centroids
  int* hoff;               // [5][B+1] per-sweep offsets of the five outputs
  int* err;
};

__device__ __forceinline__ int find_off(const int* __restrict__ off, int n, int i) {
  int lo = 0, hi = n;  // largest b in [0,n) with off[b] <= i
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- cleaning (:138-139, :97-103) --------------------------------------------------------
__global__ void valid_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  uint32_t f = 0;
  if (i < n) {
    const float x = d.xyz[3 * (size_t)i], y = d.xyz[3 * (size_t)i + 1], z = d.xyz[3 * (size_t)i + 2];
    const float r2 = xfadd(xfadd(xfmul(x, x), xfmul(y, y)), xfmul(z, z));
    f = (isfinite(x) && isfinite(y) && isfinite(z) && !(r2 < d.thres2)) ? 1u : 0u;
  }
  d.flag[i] = f;
}
__global__ void compact_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && d.flag[i]) d.vidx[d.fscan[i]] = i;
}
// scanRegistration.cpp:56 `using std::atan2;`: the reference calls the float overload; the correctly rounded float is used
// here and in the oracle (assumption A7 of oracle/scan_registration.cpp: any libm's atan2f is within one ulp of it)
__device__ __forceinline__ float neg_atan2f(float y, float x) { return (float)(-atan2((double)y, (double)x)); }
constexpr double kPi = 3.14159265358979323846;
// per-sweep start / end orientation (:143-156)
__global__ void sweep_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > d.B) return;
  const int v0 = (int)d.fscan[d.in_off[b]];
  d.voff[b] = v0;
  if (b == d.B) return;
  const int v1 = (int)d.fscan[d.in_off[b + 1]];
  d.jstar[b] = INT_MAX;
  d.ok[b] = (v1 - v0) >= 12;
  if (v1 - v0 < 12) return;
  const int a = d.vidx[v0], z = d.vidx[v1 - 1];
  const float start = neg_atan2f(d.xyz[3 * (size_t)a + 1], d.xyz[3 * (size_t)a]);
  float end = (float)xdadd((double)neg_atan2f(d.xyz[3 * (size_t)z + 1], d.xyz[3 * (size_t)z]), 2 * kPi);  // float atan2, then + 2 pi in double (:144)
  if ((double)xfsub(end, start) > 3 * kPi) end = (float)xdsub((double)end, 2 * kPi);
  else if ((double)xfsub(end, start) < kPi) end = (float)xdadd((double)end, 2 * kPi);
  d.ori[2 * b] = start;
  d.ori[2 * b + 1] = end;
}

// ring number (:168-213); -1 = dropped
__device__ __forceinline__ int ring_of(int sensor, float x, float y, float z) {
  const float planar = xfadd(xfmul(x, x), xfmul(y, y));
  const float angle = (float)xddiv(xdmul(atan(xddiv((double)z, sqrt((double)planar))), 180.0), kPi);
  if (!(angle == angle)) return -1;  // x = y = z = 0 with minimum_range 0: int(NaN) is INT_MIN on x86, i.e. dropped
  int id;
  if (sensor == S2M_SENSOR_VLP16) {
    id = (int)xdadd((double)xfmul(xfadd(angle, 15.0f), 0.5f), 0.5);
    if (id > 15 || id < 0) return -1;
  } else if (sensor == S2M_SENSOR_HDL32) {
    id = (int)xddiv(xdmul(xdadd((double)angle, 92.0 / 3.0), 3.0), 4.0);
    if (id > 31 || id < 0) return -1;
  } else if (sensor == S2M_SENSOR_HDL64) {
    if ((double)angle >= -8.83) id = (int)xdadd(xdmul((double)xfsub(2.0f, angle), 3.0), 0.5);
    else id = 32 + (int)xdadd(xdmul(xdsub(-8.83, (double)angle), 2.0), 0.5);
    if ((double)angle > 2.0 || (double)angle < -24.33 || id > 50 || id < 0) return -1;
  } else {
    id = (int)xdadd(xddiv(xdadd((double)angle, 22.5), 2.0), 0.5);
    if (id > 63 || id < 0) return -1;
  }
  return id;
}
// orientation of a point before the half turn was seen (:218-236)
__device__ __forceinline__ float ori_first_half(float y, float x, float start) {
  float ori = neg_atan2f(y, x);
  if ((double)ori < xdsub((double)start, kPi / 2)) ori = (float)xdadd((double)ori, 2 * kPi);
  else if ((double)ori > xdadd((double)start, kPi * 3 / 2)) ori = (float)xdsub((double)ori, 2 * kPi);
  return ori;
}
// FP32 screening of the two decisions of ring_kernel.  The exact arithmetic above is FP64
// (atan, sqrt, two divisions, atan2 per point) and would make this kernel FP64-pipe bound; the
// decisions only depend on which side of a threshold a value falls, so a float estimate settles
// every point that is not within a (generous) margin of a threshold and the exact path runs for
// the rest.  Float error of the estimates is < 5e-5 in the units compared; margins are 2e-3 / 1e-4.
__device__ __forceinline__ bool ring_fast(int sensor, float x, float y, float z, int& id) {
  const float planar = x * x + y * y;
  if (!(planar > 0.0f)) return false;
  const float a = atanf(z / sqrtf(planar)) * 57.29577951f;
  const float m = 2e-3f;
  float u;
  int lim;
  if (sensor == S2M_SENSOR_VLP16) { u = (a + 15.0f) * 0.5f + 0.5f; lim = 15; }
  else if (sensor == S2M_SENSOR_HDL32) { u = (a + 30.66666667f) * 0.75f; lim = 31; }
  else if (sensor == S2M_SENSOR_HDL64) {
    if (fabsf(a + 8.83f) < m || fabsf(a - 2.0f) < m || fabsf(a + 24.33f) < m) return false;
    if (a > 2.0f || a < -24.33f) { id = -1; return true; }
    u = a >= -8.83f ? (2.0f - a) * 3.0f + 0.5f : 32.0f + truncf((-8.83f - a) * 2.0f + 0.5f) + 0.5f;
    if (a < -8.83f) {  // the integer part comes from the inner expression
      const float v = (-8.83f - a) * 2.0f + 0.5f;
      if (fabsf(v - rintf(v)) < m) return false;
    }
    lim = 50;
  } else { u = (a + 22.5f) * 0.5f + 0.5f; lim = 63; }
  if (!(fabsf(u) < 1e6f) || fabsf(u - rintf(u)) < m) return false;
  const int t = (int)u;
  id = (t > lim || t < 0) ? -1 : t;
  return true;
}
__device__ __forceinline__ bool half_fast(float y, float x, float start, bool& past) {
  const float pi = 3.14159265f, m = 1e-4f;
  float o = -atan2f(y, x);
  const float lo = start - 0.5f * pi, hi = start + 1.5f * pi;
  if (fabsf(o - lo) < m || fabsf(o - hi) < m) return false;
  if (o < lo) o += 2.0f * pi;
  else if (o > hi) o -= 2.0f * pi;
  const float dlt = o - start;
  if (fabsf(dlt - pi) < m) return false;
  past = dlt > pi;
  return true;
}
__global__ void ring_kernel(Dev d, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = j < n && j < d.voff[d.B];
  int b = -1, cand = INT_MAX;  // cand: this point is past the half turn (:232-235)
  if (live) {
    b = find_off(d.voff, d.B, j);
    int id = -1;
    if (d.ok[b]) {
      const size_t s = 3 * (size_t)d.vidx[j];
      const float x = d.xyz[s], y = d.xyz[s + 1], z = d.xyz[s + 2];
      if (!ring_fast(d.sensor, x, y, z, id)) id = ring_of(d.sensor, x, y, z);
      if (id >= 0) {
        const float start = d.ori[2 * b];
        bool past;
        if (!half_fast(y, x, start, past)) past = (double)xfsub(ori_first_half(y, x, start), start) > kPi;
        if (past) cand = j;
      }
    }
    d.ring[j] = id;
  }
  // Half of every sweep qualifies, all aimed at one word per sweep: reduce inside the warp first
  // (lanes of the sweep of lane 0; a warp straddling two sweeps lets the others go alone) and look
  // at the current value before touching the atomic.
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  const int m = __reduce_min_sync(0xffffffffu, b == b0 ? cand : INT_MAX);
  if ((threadIdx.x & 31) == 0 && b0 >= 0 && m < *(volatile int*)(d.jstar + b0)) atomicMin(d.jstar + b0, m);
  if (live && b != b0 && cand < *(volatile int*)(d.jstar + b)) atomicMin(d.jstar + b, cand);
}
// relative time -> intensity (:218-253); sort key = sweep * 64 + ring
__global__ void time_kernel(Dev d, int n) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t key = (uint32_t)d.B * kRings;  // sorts behind every kept point
  if (j < d.voff[d.B]) {
    const int id = d.ring[j];
    if (id >= 0) {
      const int b = find_off(d.voff, d.B, j);
      const size_t s = 3 * (size_t)d.vidx[j];
      const float x = d.xyz[s], y = d.xyz[s + 1], z = d.xyz[s + 2];
      const float start = d.ori[2 * b], end = d.ori[2 * b + 1];
      float ori;
      if (j <= d.jstar[b]) {
        ori = ori_first_half(y, x, start);
      } else {
        ori = (float)xdadd((double)neg_atan2f(y, x), 2 * kPi);
        if ((double)ori < xdsub((double)end, kPi * 3 / 2)) ori = (float)xdadd((double)ori, 2 * kPi);
        else if ((double)ori > xdadd((double)end, kPi / 2)) ori = (float)xdsub((double)ori, 2 * kPi);
      }
      const float rel = xfdiv(xfsub(ori, start), xfsub(end, start));
      d.pts[j] = make_float4(x, y, z, (float)xdadd((double)id, xdmul(0.1, (double)rel)));
      key = (uint32_t)(b * kRings + id);
    }
  }
  d.key[j] = key;
  d.val[j] = (uint32_t)j;
}
__global__ void ring_off_kernel(Dev d, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > d.B * kRings) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (d.key2[mid] < (uint32_t)t) lo = mid + 1; else hi = mid;
  }
  d.ring_off[t] = lo;
}
__global__ void gather_kernel(Dev d, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && k < d.ring_off[d.B * kRings]) d.cloud[k] = d.pts[d.val2[k]];
}
// curvature (:271-281) and the gap flags the neighbour suppression tests (:332-339, :345-352)
__global__ void curv_kernel(Dev d, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || k >= d.ring_off[d.B * kRings]) return;
  const int b = (int)(d.key2[k] >> 6);
  const int base = d.ring_off[b * kRings], end = d.ring_off[(b + 1) * kRings];
  const int i = k - base, nb = end - base;
  float c = 0.0f;
  if (i >= 5 && i < nb - 5) {
    float s[3];
    {
      const float4 p = d.cloud[k - 5];
      s[0] = p.x; s[1] = p.y; s[2] = p.z;
    }
#pragma unroll
    for (int o = -4; o <= 5; ++o) {
      const float4 p = d.cloud[k + o];
      if (o == 0) {
        s[0] = xfsub(s[0], xfmul(10.0f, p.x)); s[1] = xfsub(s[1], xfmul(10.0f, p.y)); s[2] = xfsub(s[2], xfmul(10.0f, p.z));
      } else {
        s[0] = xfadd(s[0], p.x); s[1] = xfadd(s[1], p.y); s[2] = xfadd(s[2], p.z);
      }
    }
    c = xfadd(xfadd(xfmul(s[0], s[0]), xfmul(s[1], s[1])), xfmul(s[2], s[2]));
  }
  d.curv[k] = c;
  unsigned char g = 1;
  if (i > 0) {
    const float4 p = d.cloud[k], q = d.cloud[k - 1];
    const float dx = xfsub(p.x, q.x), dy = xfsub(p.y, q.y), dz = xfsub(p.z, q.z);
    g = (double)xfadd(xfadd(xfmul(dx, dx), xfmul(dy, dy)), xfmul(dz, dz)) > 0.05;
  }
  d.gapf[k] = g;
}

// ---- the picks (:292-413): one warp per (sweep, ring) ------------------------------------
// state byte per cloud point (d.state): bit0 picked, bit1 gap-to-previous, bits 2-3 label (1: less sharp,
// 2: sharp, 3: flat).  The reference sorts a sector by curvature and walks it from the top (then from the
// bottom), skipping points a previous pick suppressed: the next point it takes is always the largest
// (smallest) not-yet-suppressed one, so each pick is a warp-wide arg-max (arg-min) over the sector with
// the key (curvature, index) -- the order of A8 -- and no sort and no size limit are needed.
__device__ __forceinline__ void suppress(unsigned char* st, int li) {
  for (int l = 1; l <= 5; ++l) {
    if (st[li + l] & 2) break;
    st[li + l] |= 1;
  }
  for (int l = -1; l >= -5; --l) {
    if (st[li + l + 1] & 2) break;
    st[li + l] |= 1;
  }
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  return v;
}
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u < v ? u : v;
  }
  return v;
}
// The picks of one sector (:305-394) on a private state array `ps` (ps[0] = the sector's first point, five points of
// margin on either side; bit0 picked, bit1 gap-to-previous, bits 2-3 label).  One warp.
__device__ __forceinline__ void select_sector(const Dev& d, unsigned char* ps, const float* __restrict__ cv, int n, int sec, int base,
                                              int lane) {
  int n_less = 0, n_sharp = 0, n_flat = 0;
  for (int pick = 0; pick < kLess; ++pick) {  // :305-356
    unsigned long long best = 0ull;
    for (int t = lane; t < n; t += 32) {
      const float c = cv[t];
      if ((double)c > 0.1 && !(ps[t] & 1)) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | (uint32_t)t;
        best = key > best ? key : best;
      }
    }
    best = warp_max_u64(best);
    if (best == 0ull) break;
    if (lane == 0) {
      const int li = (int)(uint32_t)best;
      if (pick < kSharp) {
        ps[li] |= 2 << 2;
        d.i_sharp[sec * kSharp + n_sharp++] = base + li;
      } else {
        ps[li] |= 1 << 2;
      }
      d.i_less[sec * kLess + n_less++] = base + li;
      ps[li] |= 1;
      suppress(ps, li);
    }
    __syncwarp();
  }
  for (int pick = 0; pick < kFlat; ++pick) {  // :358-394
    unsigned long long best = ~0ull;
    for (int t = lane; t < n; t += 32) {
      const float c = cv[t];
      if ((double)c < 0.1 && !(ps[t] & 1)) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(c) << 32) | (uint32_t)t;
        best = key < best ? key : best;
      }
    }
    best = warp_min_u64(best);
    if (best == ~0ull) break;
    if (lane == 0) {
      const int li = (int)(uint32_t)best;
      ps[li] |= 3 << 2;
      d.i_flat[sec * kFlat + n_flat++] = base + li;
      if (pick < kFlat - 1) {  // the fourth flat point ends the walk before it is marked (:367-370)
        ps[li] |= 1;
        suppress(ps, li);
      }
    }
    __syncwarp();
  }
  if (lane == 0) { d.n_sharp[sec] = n_sharp; d.n_less[sec] = n_less; d.n_flat[sec] = n_flat; }
}
// The same picks with the sector's curvatures held in registers (lane l owns points l, l + 32, ...; KMAX per lane)
// and the "already picked / suppressed" state as two bit masks per lane: a pick is two warp-wide redux operations
// (largest curvature, then the position among equals), and the neighbour suppression is computed by every lane from
// the gap flags (`gap`, shared memory, gap[t] <-> point t of the sector, t in [-5, n + 5)) instead of by one lane.
// in_marks: marks an earlier sector left on points 0..4.  Returns through om / pk the marks left on the next
// sector's first five points and which of the own first five points were picked.
template <int KMAX>
__device__ __forceinline__ void select_sector_reg(const Dev& d, const unsigned char* gap, const float* __restrict__ cv, int n, int sec,
                                                  int base, int lane, unsigned in_marks, unsigned& om, unsigned& pk, int& n_sharp_out,
                                                  int& n_less_out, int& n_flat_out) {
  uint32_t cb[KMAX];
  uint32_t av_sharp = 0u, av_flat = 0u;  // bit k: point lane + 32 k is a corner / flat candidate and still free
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int t = lane + 32 * k;
    cb[k] = 0u;
    if (t < n) {
      const float c = cv[t];
      cb[k] = __float_as_uint(c);  // curvatures are sums of squares: non-negative, the bit pattern orders like the value
      const bool free_pt = !(t < 5 && ((in_marks >> t) & 1u));
      if (free_pt && (double)c > 0.1) av_sharp |= 1u << k;
      if (free_pt && (double)c < 0.1) av_flat |= 1u << k;
    }
  }
  om = 0u; pk = 0u;
  int n_less = 0, n_sharp = 0, n_flat = 0;
  // the neighbours a pick at li suppresses (:332-352): the run [lo, hi] around li that no gap interrupts
  auto exclude = [&](int li) {
    uint32_t win = 0u;  // bit i: a gap in front of point li - 5 + i (eleven independent shared-memory loads)
#pragma unroll
    for (int i = 0; i < 11; ++i) win |= (uint32_t)((gap[li - 5 + i] >> 1) & 1) << i;
    const uint32_t fw = (win >> 6) & 31u;  // gaps in front of li+1 .. li+5: the forward walk stops at the first
    const uint32_t bw = (win >> 1) & 31u;  // gaps in front of li-4 .. li: the backward walk stops at the first from li down
    const int hi = li + (fw ? __ffs((int)fw) - 1 : 5);
    const int lo = li - (bw ? __clz((int)bw) - 27 : 5);
    const int x = lo + ((lane - lo) & 31);  // the one point of [lo, hi] this lane owns, if any
    if (x <= hi && x >= 0 && x < n) { av_sharp &= ~(1u << (x >> 5)); av_flat &= ~(1u << (x >> 5)); }
    for (int p2 = n; p2 <= hi; ++p2) om |= 1u << (p2 - n);
  };
  for (int pick = 0; pick < kLess; ++pick) {  // :305-356: largest (curvature, index) first
    uint32_t bv = 0u;
    int bk = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (((av_sharp >> k) & 1u) && cb[k] >= bv) { bv = cb[k]; bk = k; }
    const uint32_t gmax = __reduce_max_sync(0xffffffffu, bv);
    if (gmax == 0u) break;
    const int li = __reduce_max_sync(0xffffffffu, (bv == gmax && ((av_sharp >> bk) & 1u)) ? lane + 32 * bk : -1);
    if (lane == 0) {
      if (pick < kSharp) d.i_sharp[sec * kSharp + n_sharp] = base + li;
      d.i_less[sec * kLess + n_less] = base + li;
    }
    if (pick < kSharp) ++n_sharp;
    ++n_less;
    if (li < 5) pk |= 1u << li;
    exclude(li);
  }
  for (int pick = 0; pick < kFlat; ++pick) {  // :358-394: smallest (curvature, index) first
    uint32_t bv = 0xFFFFFFFFu;
    int bk = 0;
#pragma unroll
    for (int k = KMAX - 1; k >= 0; --k)
      if (((av_flat >> k) & 1u) && cb[k] <= bv) { bv = cb[k]; bk = k; }
    const bool have = av_flat != 0u;
    const uint32_t gmin = __reduce_min_sync(0xffffffffu, have ? bv : 0xFFFFFFFFu);
    if (__ballot_sync(0xffffffffu, have) == 0u) break;
    const int li = __reduce_min_sync(0xffffffffu, (have && bv == gmin) ? lane + 32 * bk : 0x7FFFFFFF);
    if (lane == 0) d.i_flat[sec * kFlat + n_flat] = base + li;
    ++n_flat;
    if (li < 5) pk |= 1u << li;
    if (pick < kFlat - 1) exclude(li);  // the fourth flat point ends the walk before it is marked (:367-370)
  }
  if (lane == 0) { d.n_sharp[sec] = n_sharp; d.n_less[sec] = n_less; d.n_flat[sec] = n_flat; }
  n_sharp_out = n_sharp; n_less_out = n_less; n_flat_out = n_flat;
}
// One block per (sweep, ring), one warp per sector.  The reference walks the six sectors of a ring in order, and
// the only thing a sector hands to the next one is the neighbour-suppression marks its picks leave on the next
// sector's first five points (:332-339).  The six warps therefore run their sectors AT THE SAME TIME, assuming no
// incoming marks; afterwards the sectors are validated in order: if the marks sector j really leaves do not touch
// a point sector j+1 picked, excluding those points from j+1's arg-max/arg-min walks would not have changed any
// pick, so its result stands -- otherwise sector j+1 is redone with the marks set (and the check moves on with its
// new outgoing marks).  Rings whose sectors exceed the staging (more than kSecCap points) or are shorter than the
// reach of a mark take the serial walk on the global state array.
constexpr int kSecCap = 1000, kSecPad = 5;
__global__ void __launch_bounds__(32 * kSectors) select_kernel(Dev d) {
  __shared__ unsigned char priv[kSectors][kSecCap + 2 * kSecPad + 6];
  __shared__ unsigned out_marks[kSectors], picked5[kSectors];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rid = blockIdx.x;
  const int rs = d.ring_off[rid], re = d.ring_off[rid + 1], len = re - rs;
  const int first = rs + 5, last = re - 6;
  if (last - first < 6) {  // :294
    if (w == 0) {
      if (lane < kSectors) d.n_sharp[rid * kSectors + lane] = d.n_less[rid * kSectors + lane] = d.n_flat[rid * kSectors + lane] = 0;
      if (lane == 0) d.lf_cnt[rid] = 0;
    }
    return;
  }
  unsigned char* st = d.state + rs;
  const int sp = first + (last - first) * w / 6, ep = first + (last - first) * (w + 1) / 6 - 1;
  const int n = ep - sp + 1;
  // every sector of the ring fits the staging and has at least five points (marks reach one sector ahead at most)
  const bool staged = (last - first + 5) / 6 + 1 <= kSecCap && last - first >= 36;
  if (staged) {
    unsigned char* gap = priv[w] + kSecPad;  // gap[t] <-> cloud point sp + t, t in [-5, n + 5)
    const int sec = rid * kSectors + w;
    for (int t = lane - kSecPad; t < n + kSecPad; t += 32) gap[t] = d.gapf[sp + t] ? 2 : 0;
    __syncwarp();
    int n_sharp = 0, n_less = 0, n_flat = 0;
    for (unsigned in_marks = 0u, pass = 0u;; ++pass) {
      // pass 0: every sector, speculatively without incoming marks; pass j = 1..5: sector j again if it has to
      bool run = pass == 0u;
      if (pass > 0u) {
        __syncthreads();
        if (pass >= (unsigned)kSectors) break;
        in_marks = out_marks[pass - 1];
        run = (unsigned)w == pass && (in_marks & picked5[w]) != 0u;
      }
      if (run) {
        unsigned om, pk;
        if (n <= 256) select_sector_reg<8>(d, gap, d.curv + sp, n, sec, sp, lane, in_marks, om, pk, n_sharp, n_less, n_flat);
        else if (n <= 512) select_sector_reg<16>(d, gap, d.curv + sp, n, sec, sp, lane, in_marks, om, pk, n_sharp, n_less, n_flat);
        else select_sector_reg<32>(d, gap, d.curv + sp, n, sec, sp, lane, in_marks, om, pk, n_sharp, n_less, n_flat);
        if (lane == 0) { out_marks[w] = om; picked5[w] = pk; }
      }
    }
    // labels of the sector's points for the less-flat pass below (1 less sharp, 2 sharp, 3 flat), from the final lists
    for (int t = lane; t < n; t += 32) st[sp - rs + t] = 0;
    __syncwarp();
    if (lane < n_less) st[d.i_less[sec * kLess + lane] - rs] = (lane < n_sharp ? 2 : 1) << 2;
    if (lane < n_flat) st[d.i_flat[sec * kFlat + lane] - rs] = 3 << 2;
  } else if (w == 0) {
    for (int t = lane; t < len; t += 32) st[t] = d.gapf[rs + t] ? 2 : 0;
    __syncwarp();
    for (int j = 0; j < kSectors; ++j) {
      const int sp2 = first + (last - first) * j / 6, ep2 = first + (last - first) * (j + 1) / 6 - 1;
      select_sector(d, st + (sp2 - rs), d.curv + sp2, ep2 - sp2 + 1, rid * kSectors + j, sp2, lane);
      __syncwarp();
    }
  }
  __syncthreads();
  // less-flat points of the ring (:396-402): label <= 0 over first .. last-1, in position order; every warp packs
  // its own sector behind the sectors before it
  __shared__ int lf_n[kSectors];
  __shared__ float lf_box[kSectors][6];
  int mine = 0;
  for (int base = sp; base <= ep; base += 32) {
    const int k = base + lane;
    bool keep = false;
    if (k <= ep) {
      const int lab = (st[k - rs] >> 2) & 3;
      keep = lab == 0 || lab == 3;
    }
    mine += __popc(__ballot_sync(0xffffffffu, keep));
  }
  if (lane == 0) lf_n[w] = mine;
  __syncthreads();
  int running = 0;
  for (int j = 0; j < w; ++j) running += lf_n[j];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int base = sp; base <= ep; base += 32) {
    const int k = base + lane;
    bool keep = false;
    if (k <= ep) {
      const int lab = (st[k - rs] >> 2) & 3;
      keep = lab == 0 || lab == 3;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const float4 p = d.cloud[k];
      d.lf_stage[rs + running + __popc(bal & ((1u << lane) - 1u))] = p;
      mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
      mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
    running += __popc(bal);
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  if (lane == 0) for (int a = 0; a < 3; ++a) { lf_box[w][a] = mn[a]; lf_box[w][3 + a] = mx[a]; }
  __syncthreads();
  if (w == 0 && lane == 0) {
    int total = 0;
    for (int j = 0; j < kSectors; ++j) total += lf_n[j];
    d.lf_cnt[rid] = total;
    for (int a = 0; a < 3; ++a) {
      float lo = lf_box[0][a], hi = lf_box[0][3 + a];
      for (int j = 1; j < kSectors; ++j) { lo = fminf(lo, lf_box[j][a]); hi = fmaxf(hi, lf_box[j][3 + a]); }
      d.rbox[6 * rid + a] = lo; d.rbox[6 * rid + 3 + a] = hi;
    }
  }
}

// per-sweep offsets of the outputs, for the host
__global__ void pack_off_kernel(Dev d) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > d.B) return;
  const int P = d.B + 1;
  d.hoff[0 * P + b] = d.ring_off[b * kRings];
  d.hoff[1 * P + b] = d.o_sharp[b * kSecPerScan];
  d.hoff[2 * P + b] = d.o_less[b * kSecPerScan];
  d.hoff[3 * P + b] = d.o_flat[b * kSecPerScan];
  d.hoff[4 * P + b] = d.lf_off[b * kRings];  // less-flat points BEFORE the voxel filter
}
__global__ void emit_picks_kernel(Dev d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.B * kSecPerScan * kLess) return;
  const int sec = t / kLess, q = t % kLess;
  if (q < d.n_less[sec]) d.out_less[d.o_less[sec] + q] = d.cloud[d.i_less[sec * kLess + q]];
  if (q < kSharp && q < d.n_sharp[sec]) d.out_sharp[d.o_sharp[sec] + q] = d.cloud[d.i_sharp[sec * kSharp + q]];
  if (q < kFlat && q < d.n_flat[sec]) d.out_flat[d.o_flat[sec] + q] = d.cloud[d.i_flat[sec * kFlat + q]];
}

// ---- per-ring VoxelGrid(0.2) (:404-411; PCL 1.8 semantics, oracle assumption A1) ------------
__global__ void lf_key_kernel(Dev d, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || k >= d.ring_off[d.B * kRings]) return;
  const int rid = (int)d.key2[k];
  const int pos = k - d.ring_off[rid];
  if (pos >= d.lf_cnt[rid]) return;
  const float4 p = d.lf_stage[k];
  const float inv = xfdiv(1.0f, 0.2f);
  const float* bx = d.rbox + 6 * rid;
  int lo[3], dv[3];
  double cells = 1.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = (int)floorf(xfmul(bx[a], inv));
    dv[a] = (int)floorf(xfmul(bx[3 + a], inv)) - lo[a] + 1;
    cells *= (double)((long long)xfmul(xfsub(bx[3 + a], bx[a]), inv) + 1);
  }
  if (cells > 2147483647.0) atomicCAS(d.err, 0, S2M_ERR_RANGE);  // PCL would pass the ring through unfiltered
  const int i0 = (int)xfsub(floorf(xfmul(p.x, inv)), (float)lo[0]);
  const int i1 = (int)xfsub(floorf(xfmul(p.y, inv)), (float)lo[1]);
  const int i2 = (int)xfsub(floorf(xfmul(p.z, inv)), (float)lo[2]);
  const int vk = i0 + i1 * dv[0] + i2 * dv[0] * dv[1];
  const int dest = d.lf_off[rid] + pos;
  d.lf_pts[dest] = p;
  d.vkey[dest] = ((unsigned long long)rid << 32) | (uint32_t)vk;
  d.vval[dest] = (uint32_t)dest;
}
__global__ void lf_head_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  d.flag[i] = (i < n && (i == 0 || d.vkey2[i] != d.vkey2[i - 1])) ? 1u : 0u;
}
__global__ void lf_centroid_kernel(Dev d, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !d.flag[i]) return;
  const unsigned long long key = d.vkey2[i];
  float s[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  int cnt = 0;
  for (int t = i; t < n && d.vkey2[t] == key; ++t, ++cnt) {
    const float4 p = d.lf_pts[d.vval2[t]];
    s[0] = xfadd(s[0], p.x); s[1] = xfadd(s[1], p.y); s[2] = xfadd(s[2], p.z); s[3] = xfadd(s[3], p.w);
  }
  const float c = (float)cnt;
  d.out_lf[d.fscan[i]] = make_float4(xfdiv(s[0], c), xfdiv(s[1], c), xfdiv(s[2], c), xfdiv(s[3], c));
}
__global__ void lf_out_off_kernel(Dev d, int n) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > d.B) return;
  const unsigned long long key = (unsigned long long)(b * kRings) << 32;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (d.vkey2[mid] < key) lo = mid + 1; else hi = mid;
  }
  d.hoff[4 * (d.B + 1) + b] = (int)d.fscan[lo];  // fscan has n+1 entries: voxels before position lo
}

}  // namespace fx
}  // namespace s2m

using namespace s2m::fx;

struct s2m_fx {
  s2m_fx_params P;
  Dev d;
  std::vector<void*> allocs;
  cudaStream_t stream = nullptr;
  float* d_xyz = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  int* h_off = nullptr;  // pinned [5][B+1]
  int* h_err = nullptr;
  long long cap = 0;
  long long launches = 0;
  std::string err;
  bool have = false;
};

#define FXCK(call)                                                                 \
  do {                                                                             \
    cudaError_t e_ = (call);                                                       \
    if (e_ != cudaSuccess) {                                                       \
      if (fx) fx->err = std::string(#call) + ": " + cudaGetErrorString(e_);        \
      return S2M_ERR_CUDA;                                                         \
    }                                                                              \
  } while (0)

template <typename T>
static int fx_alloc(s2m_fx* fx, T** p, size_t n) {
  void* q = nullptr;
  FXCK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
  fx->allocs.push_back(q);
  *p = (T*)q;
  return 0;
}
static inline int cdivi(long long a, int b) { return (int)((a + b - 1) / b); }

extern "C" void s2m_fx_destroy(s2m_fx* fx) {
  if (!fx) return;
  cudaSetDevice(fx->P.device);
  if (fx->stream) { cudaStreamSynchronize(fx->stream); cudaStreamDestroy(fx->stream); }
  for (void* p : fx->allocs) cudaFree(p);
  if (fx->h_off) cudaFreeHost(fx->h_off);
  if (fx->h_err) cudaFreeHost(fx->h_err);
  delete fx;
}

static int fx_create_impl(s2m_fx* fx) {
  const s2m_fx_params& P = fx->P;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { fx->err = "no CUDA device"; return S2M_ERR_CUDA; }
  FXCK(cudaSetDevice(P.device));
  FXCK(cudaStreamCreateWithFlags(&fx->stream, cudaStreamNonBlocking));
  const int B = P.batch;
  const size_t cap = (size_t)B * (size_t)P.cap_points;
  fx->cap = (long long)cap;
  Dev& d = fx->d;
  std::memset(&d, 0, sizeof d);
  d.B = B;
  d.sensor = P.sensor;
  const float thres = (float)P.minimum_range;
  d.thres2 = thres * thres;
  int rc = 0;
  rc |= fx_alloc(fx, &fx->d_xyz, cap * 3);
  rc |= fx_alloc(fx, &d.in_off, B + 1); rc |= fx_alloc(fx, &d.flag, cap + 1); rc |= fx_alloc(fx, &d.fscan, cap + 1);
  rc |= fx_alloc(fx, &d.vidx, cap); rc |= fx_alloc(fx, &d.voff, B + 1); rc |= fx_alloc(fx, &d.ori, 2 * B);
  rc |= fx_alloc(fx, &d.jstar, B); rc |= fx_alloc(fx, &d.ok, B); rc |= fx_alloc(fx, &d.ring, cap);
  rc |= fx_alloc(fx, &d.pts, cap); rc |= fx_alloc(fx, &d.key, cap); rc |= fx_alloc(fx, &d.key2, cap);
  rc |= fx_alloc(fx, &d.val, cap); rc |= fx_alloc(fx, &d.val2, cap); rc |= fx_alloc(fx, &d.ring_off, B * kRings + 1);
  rc |= fx_alloc(fx, &d.cloud, cap); rc |= fx_alloc(fx, &d.curv, cap); rc |= fx_alloc(fx, &d.gapf, cap); rc |= fx_alloc(fx, &d.state, cap);
  const size_t ns = (size_t)B * kSecPerScan;
  rc |= fx_alloc(fx, &d.n_sharp, ns + 1); rc |= fx_alloc(fx, &d.n_less, ns + 1); rc |= fx_alloc(fx, &d.n_flat, ns + 1);
  rc |= fx_alloc(fx, &d.o_sharp, ns + 1); rc |= fx_alloc(fx, &d.o_less, ns + 1); rc |= fx_alloc(fx, &d.o_flat, ns + 1);
  rc |= fx_alloc(fx, &d.i_sharp, ns * kSharp); rc |= fx_alloc(fx, &d.i_less, ns * kLess); rc |= fx_alloc(fx, &d.i_flat, ns * kFlat);
  rc |= fx_alloc(fx, &d.out_sharp, ns * kSharp); rc |= fx_alloc(fx, &d.out_less, ns * kLess); rc |= fx_alloc(fx, &d.out_flat, ns * kFlat);
  rc |= fx_alloc(fx, &d.lf_cnt, B * kRings + 1); rc |= fx_alloc(fx, &d.lf_off, B * kRings + 1); rc |= fx_alloc(fx, &d.rbox, (size_t)B * kRings * 6);
  rc |= fx_alloc(fx, &d.lf_stage, cap); rc |= fx_alloc(fx, &d.lf_pts, cap);
  rc |= fx_alloc(fx, &d.vkey, cap); rc |= fx_alloc(fx, &d.vkey2, cap); rc |= fx_alloc(fx, &d.vval, cap); rc |= fx_alloc(fx, &d.vval2, cap);
  rc |= fx_alloc(fx, &d.out_lf, cap); rc |= fx_alloc(fx, &d.hoff, 5 * (B + 1)); rc |= fx_alloc(fx, &d.err, 1);
  if (rc) return S2M_ERR_CUDA;
  size_t t1 = 0, t2 = 0, t3 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, d.key, d.key2, d.val, d.val2, (int)cap, 0, 32, fx->stream);
  cub::DeviceRadixSort::SortPairs(nullptr, t2, d.vkey, d.vkey2, d.vval, d.vval2, (int)cap, 0, 64, fx->stream);
  cub::DeviceScan::ExclusiveSum(nullptr, t3, d.flag, d.fscan, (int)cap + 1, fx->stream);
  fx->cub_bytes = std::max(t1, std::max(t2, t3));
  FXCK(cudaMalloc(&fx->cub_tmp, fx->cub_bytes));
  fx->allocs.push_back(fx->cub_tmp);
  FXCK(cudaMallocHost((void**)&fx->h_off, sizeof(int) * 5 * (B + 1)));
  FXCK(cudaMallocHost((void**)&fx->h_err, sizeof(int)));
  FXCK(cudaMemset(d.err, 0, sizeof(int)));
  FXCK(cudaMemset(d.n_sharp, 0, sizeof(int) * (ns + 1)));
  FXCK(cudaMemset(d.n_less, 0, sizeof(int) * (ns + 1)));
  FXCK(cudaMemset(d.n_flat, 0, sizeof(int) * (ns + 1)));
  FXCK(cudaMemset(d.lf_cnt, 0, sizeof(int) * (B * kRings + 1)));
  return S2M_OK;
}

extern "C" int s2m_fx_create(const s2m_fx_params* p, s2m_fx** out) {
  if (!p || !out) return S2M_ERR_ARG;
  *out = nullptr;
  if (p->batch < 1 || p->batch > 1024 || p->cap_points < 16 || p->sensor < 0 || p->sensor > 3 ||
      (long long)p->batch * p->cap_points > 0x7fffff00LL || !(p->minimum_range >= 0.0))
    return S2M_ERR_ARG;
  s2m_fx* fx = new s2m_fx();
  fx->P = *p;
  const int rc = fx_create_impl(fx);
  if (rc != S2M_OK) { s2m_fx_destroy(fx); return rc; }
  *out = fx;
  return S2M_OK;
}

extern "C" const char* s2m_fx_last_error(s2m_fx* fx) { return fx ? fx->err.c_str() : "null context"; }
extern "C" long long s2m_fx_launch_count(s2m_fx* fx) { return fx ? fx->launches : 0; }

extern "C" int s2m_fx_extract(s2m_fx* fx, const float* xyz, const int* off, int device_input) {
  if (!fx || !xyz || !off) return S2M_ERR_ARG;
  FXCK(cudaSetDevice(fx->P.device));
  Dev& d = fx->d;
  const int B = d.B;
  if (off[0] != 0) { fx->err = "offsets must start at 0"; return S2M_ERR_ARG; }
  for (int b = 0; b < B; ++b)
    if (off[b + 1] < off[b] || off[b + 1] - off[b] > fx->P.cap_points) { fx->err = "sweep larger than cap_points"; return S2M_ERR_CAPACITY; }
  const int n = off[B];
  cudaStream_t s = fx->stream;
  fx->have = false;
  FXCK(cudaMemcpyAsync(d.in_off, off, sizeof(int) * (B + 1), cudaMemcpyHostToDevice, s));
  if (n > 0) FXCK(cudaMemcpyAsync(fx->d_xyz, xyz, sizeof(float) * 3 * (size_t)n, device_input ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
  d.xyz = fx->d_xyz;
  const int T = 256;
  long long k = 0;
  size_t tb;
  valid_kernel<<<cdivi(n + 1, T), T, 0, s>>>(d, n); ++k;
  tb = fx->cub_bytes;
  cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.flag, d.fscan, n + 1, s);
  if (n > 0) { compact_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k; }
  sweep_kernel<<<cdivi(B + 1, 64), 64, 0, s>>>(d); ++k;
  if (n > 0) {
    ring_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k;
    time_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k;
    int bits = 1;
    while ((1 << bits) <= B * kRings) ++bits;
    tb = fx->cub_bytes;
    cub::DeviceRadixSort::SortPairs(fx->cub_tmp, tb, d.key, d.key2, d.val, d.val2, n, 0, bits, s);
  }
  ring_off_kernel<<<cdivi(B * kRings + 1, T), T, 0, s>>>(d, n); ++k;
  if (n > 0) {
    gather_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k;
    curv_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k;
  }
  select_kernel<<<B * kRings, 32 * kSectors, 0, s>>>(d); ++k;
  const int ns = B * kSecPerScan;
  tb = fx->cub_bytes; cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.n_sharp, d.o_sharp, ns + 1, s);
  tb = fx->cub_bytes; cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.n_less, d.o_less, ns + 1, s);
  tb = fx->cub_bytes; cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.n_flat, d.o_flat, ns + 1, s);
  tb = fx->cub_bytes; cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.lf_cnt, d.lf_off, B * kRings + 1, s);
  pack_off_kernel<<<cdivi(B + 1, 64), 64, 0, s>>>(d); ++k;
  FXCK(cudaMemcpyAsync(fx->h_off, d.hoff, sizeof(int) * 5 * (B + 1), cudaMemcpyDeviceToHost, s));
  emit_picks_kernel<<<cdivi((long long)ns * kLess, T), T, 0, s>>>(d); ++k;
  FXCK(cudaStreamSynchronize(s));  // the number of less-flat points sizes the voxel sort
  const int n_lf = fx->h_off[4 * (B + 1) + B];
  if (n > 0) { lf_key_kernel<<<cdivi(n, T), T, 0, s>>>(d, n); ++k; }
  if (n_lf > 0) {
    int bits = 1;
    while ((1 << bits) < B * kRings) ++bits;
    tb = fx->cub_bytes;
    cub::DeviceRadixSort::SortPairs(fx->cub_tmp, tb, d.vkey, d.vkey2, d.vval, d.vval2, n_lf, 0, 32 + bits, s);
  }
  lf_head_kernel<<<cdivi(n_lf + 1, T), T, 0, s>>>(d, n_lf); ++k;
  tb = fx->cub_bytes;
  cub::DeviceScan::ExclusiveSum(fx->cub_tmp, tb, d.flag, d.fscan, n_lf + 1, s);
  if (n_lf > 0) { lf_centroid_kernel<<<cdivi(n_lf, T), T, 0, s>>>(d, n_lf); ++k; }
  lf_out_off_kernel<<<cdivi(B + 1, 64), 64, 0, s>>>(d, n_lf); ++k;
  FXCK(cudaMemcpyAsync(fx->h_off + 4 * (B + 1), d.hoff + 4 * (B + 1), sizeof(int) * (B + 1), cudaMemcpyDeviceToHost, s));
  FXCK(cudaMemcpyAsync(fx->h_err, d.err, sizeof(int), cudaMemcpyDeviceToHost, s));
  FXCK(cudaStreamSynchronize(s));
  FXCK(cudaGetLastError());
  fx->launches += k;
  if (*fx->h_err != 0) {
    const int e = *fx->h_err;
    cudaMemsetAsync(d.err, 0, sizeof(int), s);
    fx->err = "voxel lattice of a ring overflows 31 bits";
    return e;
  }
  fx->have = true;
  return S2M_OK;
}

static const float4* fx_cloud_ptr(const s2m_fx* fx, int which) {
  switch (which) {
    case S2M_FX_FULL: return fx->d.cloud;
    case S2M_FX_SHARP: return fx->d.out_sharp;
    case S2M_FX_LESS_SHARP: return fx->d.out_less;
    case S2M_FX_FLAT: return fx->d.out_flat;
    case S2M_FX_LESS_FLAT: return fx->d.out_lf;
    default: return nullptr;
  }
}
extern "C" int s2m_fx_offsets(s2m_fx* fx, int which, int* off_out) {
  if (!fx || !off_out || which < 0 || which > 4) return S2M_ERR_ARG;
  if (!fx->have) { fx->err = "no extracted sweep batch"; return S2M_ERR_ARG; }
  const int P = fx->d.B + 1;
  const int base = fx->h_off[which * P];  // the full cloud starts at ring_off[0] = 0; the others at 0 too
  for (int b = 0; b < P; ++b) off_out[b] = fx->h_off[which * P + b] - base;
  return S2M_OK;
}
extern "C" const float* s2m_fx_device_cloud(s2m_fx* fx, int which) {
  if (!fx || !fx->have) return nullptr;
  return reinterpret_cast<const float*>(fx_cloud_ptr(fx, which));
}
extern "C" int s2m_fx_download(s2m_fx* fx, int which, float* out, int cap) {
  if (!fx || which < 0 || which > 4 || cap < 0) return S2M_ERR_ARG;
  if (!fx->have) { fx->err = "no extracted sweep batch"; return S2M_ERR_ARG; }
  FXCK(cudaSetDevice(fx->P.device));
  const int P = fx->d.B + 1;
  const int n = fx->h_off[which * P + fx->d.B] - fx->h_off[which * P];
  if (out && n > 0) {
    if (n > cap) return S2M_ERR_CAPACITY;
    FXCK(cudaMemcpyAsync(out, fx_cloud_ptr(fx, which), sizeof(float4) * (size_t)n, cudaMemcpyDeviceToHost, fx->stream));
    FXCK(cudaStreamSynchronize(fx->stream));
  }
  return n;
}
