// oracle/scan_registration.cpp -- CPU restatement of SC-A-LOAM's feature extraction
// (SURVEY.md 8f row N2): /root/reference/src/scanRegistration.cpp:116-454, laserCloudHandler.
//
// TEST INFRASTRUCTURE ONLY (same rules as s2m_oracle.cpp): nothing under sc-a-loam_b200/
// links, loads or calls this file.
//
// PARITY UNPINNED: the reference has no tests or vectors for this path and needs ROS + PCL
// to build.  One real-data check exists: the KAIST03 scans the reference ships
// carry, in their intensity channel, the ring number this node assigned (integer part); the
// ring rule below reproduces it for every point (tests/test_scan_registration.py, authoring
// container only).
//
// Stage map (file = scanRegistration.cpp):
//   removeNaNFromPointCloud + removeClosedPointCloud   :138-139, :87-114  -> clean()
//   start / end orientation                             :143-156          -> sweep_bounds()
//   ring number per lidar type, relative time           :165-255          -> ring_of(), the loop in extract()
//   ring-major concatenation, scanStartInd/scanEndInd   :261-267
//   11-tap curvature                                    :271-281          -> curvature()
//   six sectors per ring: sort, 2 sharp / 20 less sharp / 4 flat, neighbour suppression
//                                                       :292-394          -> select_sector(), suppress()
//   less-flat points of the ring + VoxelGrid(0.2)       :396-413          -> voxel_thin() (assumption A1 of s2m_oracle.cpp)
//
// Assumptions (numbered on from s2m_oracle.cpp):
//   A7 atan2 at :143 / :144 / :221: the file declares `using std::atan2;` (:56), so with float arguments the
//      reference calls the FLOAT overload std::atan2(float, float) = atan2f (round-1 advisor finding; the earlier
//      reading "resolves to the C double function" was wrong for atan2 -- it holds for the `atan` of :168 / :182 /
//      :195 / :206, whose argument is a float expression but which is called unqualified on the C name).  atan2f is
//      not correctly rounded in every libm: glibc 2.39 (this image) differs from the correctly rounded result in
//      16 % of random arguments (measured, by one ulp), glibc >= 2.41 (CORE-MATH) never.  The restatement and the
//      kernel both use the CORRECTLY ROUNDED float, (float)atan2((double)y, (double)x) -- the value every libm is
//      within one ulp of and the newest ones return.  Effect of that ulp: the low bit of the relative-time fraction
//      stored in the intensity channel, and the half-turn decision for a point within one ulp of the threshold;
//      ring numbers are not affected (they come from `atan`, above).  Parity with a given reference BUILD is
//      therefore within one ulp of `ori`, not bit-exact, unless that build's libm rounds atan2f correctly.
//   A8 std::sort (:301) is not stable: equal curvatures inside a sector come out in an
//      implementation-defined order.  Canonical order here: (curvature, index) ascending
//      (tie_rule 0); tie_rule 1 runs std::sort with the reference's comparator so tests can show
//      both agree on the inputs used.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace {

struct Pt { float x, y, z, i; };

enum Sensor { HDL64 = 0, VLP16 = 1, OS1_64 = 2, HDL32 = 3 };

int scan_lines(int sensor) { return sensor == VLP16 ? 16 : (sensor == HDL32 ? 32 : 64); }

// :138-139 (PCL drops a point when x, y or z is not finite), :97-103
void clean(const float* xyz, int n, float thres, std::vector<Pt>& out) {
  out.clear();
  for (int i = 0; i < n; ++i) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
    if (x * x + y * y + z * z < thres * thres) continue;
    out.push_back({x, y, z, 0.0f});
  }
}

// :143-156
void sweep_bounds(const std::vector<Pt>& c, float& start, float& end) {
  start = (float)(-::atan2((double)c.front().y, (double)c.front().x));
  end = (float)((double)(float)(-::atan2((double)c.back().y, (double)c.back().x)) + 2 * M_PI);  // float atan2f, then + 2 pi in double (:144)
  if (end - start > 3 * M_PI) end = (float)(end - 2 * M_PI);
  else if (end - start < M_PI) end = (float)(end + 2 * M_PI);
}

// :168-213; returns -1 for a point the reference drops
int ring_of(int sensor, const Pt& p) {
  const float planar = p.x * p.x + p.y * p.y;
  const float angle = (float)(::atan((double)p.z / ::sqrt((double)planar)) * 180 / M_PI);
  const int lines = scan_lines(sensor);
  if (!(angle == angle)) return -1;  // x = y = z = 0 (only with minimum_range 0): int(NaN) is INT_MIN on x86, i.e. dropped
  int id;
  if (sensor == VLP16) {
    id = int((angle + 15) / 2 + 0.5);
    if (id > lines - 1 || id < 0) return -1;
  } else if (sensor == HDL32) {
    id = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
    if (id > lines - 1 || id < 0) return -1;
  } else if (sensor == HDL64) {
    if (angle >= -8.83) id = int((2 - angle) * 3.0 + 0.5);
    else id = lines / 2 + int((-8.83 - angle) * 2.0 + 0.5);
    if (angle > 2 || angle < -24.33 || id > 50 || id < 0) return -1;
  } else {
    id = int((angle + 22.5) / 2 + 0.5);
    if (id > lines - 1 || id < 0) return -1;
  }
  return id;
}

// :271-281
void curvature(const std::vector<Pt>& c, std::vector<float>& curv) {
  const int n = (int)c.size();
  curv.assign(n, 0.0f);
  for (int i = 5; i < n - 5; ++i) {
    float d[3];
    for (int a = 0; a < 3; ++a) {
      auto v = [&](int k) { return a == 0 ? c[k].x : (a == 1 ? c[k].y : c[k].z); };
      d[a] = v(i - 5) + v(i - 4) + v(i - 3) + v(i - 2) + v(i - 1) - 10 * v(i) + v(i + 1) + v(i + 2) + v(i + 3) + v(i + 4) + v(i + 5);
    }
    curv[i] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  }
}

// :330-355 (and the identical block :371-394)
void suppress(const std::vector<Pt>& c, std::vector<int>& picked, int ind) {
  auto gap2 = [&](int a, int b) {
    const float dx = c[a].x - c[b].x, dy = c[a].y - c[b].y, dz = c[a].z - c[b].z;
    return dx * dx + dy * dy + dz * dz;
  };
  for (int l = 1; l <= 5; ++l) {
    if (gap2(ind + l, ind + l - 1) > 0.05) break;
    picked[ind + l] = 1;
  }
  for (int l = -1; l >= -5; --l) {
    if (gap2(ind + l, ind + l + 1) > 0.05) break;
    picked[ind + l] = 1;
  }
}

// PCL 1.8 VoxelGrid on one ring's less-flat points (A1)
void voxel_thin(const std::vector<Pt>& in, float leaf, std::vector<Pt>& out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;
  float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
  for (const Pt& p : in) {
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  int lo[3], hi[3];
  for (int a = 0; a < 3; ++a) { lo[a] = (int)std::floor(mn[a] * inv); hi[a] = (int)std::floor(mx[a] * inv); }
  const int64_t dx = hi[0] - lo[0] + 1, dy = hi[1] - lo[1] + 1;
  std::vector<std::pair<int64_t, int>> keyed(in.size());
  for (size_t k = 0; k < in.size(); ++k) {
    const int i0 = (int)(std::floor(in[k].x * inv) - (float)lo[0]);
    const int i1 = (int)(std::floor(in[k].y * inv) - (float)lo[1]);
    const int i2 = (int)(std::floor(in[k].z * inv) - (float)lo[2]);
    keyed[k] = {i0 + i1 * dx + i2 * dx * dy, (int)k};
  }
  std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  for (size_t a = 0; a < keyed.size();) {
    size_t b = a;
    float s[4] = {0, 0, 0, 0};
    for (; b < keyed.size() && keyed[b].first == keyed[a].first; ++b) {
      const Pt& p = in[keyed[b].second];
      s[0] += p.x; s[1] += p.y; s[2] += p.z; s[3] += p.i;
    }
    const float cnt = (float)(b - a);
    out.push_back({s[0] / cnt, s[1] / cnt, s[2] / cnt, s[3] / cnt});
    a = b;
  }
}

struct Features {
  std::vector<Pt> full, sharp, less_sharp, flat, less_flat;
};

void extract(int sensor, double minimum_range, const float* xyz, int n, int tie_rule, Features& F) {
  const double scan_period = 0.1;  // :62
  const int lines = scan_lines(sensor);
  std::vector<Pt> in;
  clean(xyz, n, (float)minimum_range, in);
  F = Features();
  if (in.size() < 12) return;  // the reference would index out of range; callers never send this
  float start, end;
  sweep_bounds(in, start, end);
  bool half = false;
  std::vector<std::vector<Pt>> rings(lines);
  for (const Pt& src : in) {  // :161-256
    Pt p = src;
    const int id = ring_of(sensor, p);
    if (id < 0) continue;
    float ori = (float)(-::atan2((double)p.y, (double)p.x));
    if (!half) {
      if (ori < start - M_PI / 2) ori = (float)(ori + 2 * M_PI);
      else if (ori > start + M_PI * 3 / 2) ori = (float)(ori - 2 * M_PI);
      if (ori - start > M_PI) half = true;
    } else {
      ori = (float)(ori + 2 * M_PI);
      if (ori < end - M_PI * 3 / 2) ori = (float)(ori + 2 * M_PI);
      else if (ori > end + M_PI / 2) ori = (float)(ori - 2 * M_PI);
    }
    const float rel = (ori - start) / (end - start);
    p.i = (float)(id + scan_period * rel);
    rings[id].push_back(p);
  }
  std::vector<int> first(lines), last(lines);
  for (int r = 0; r < lines; ++r) {  // :261-267
    first[r] = (int)F.full.size() + 5;
    F.full.insert(F.full.end(), rings[r].begin(), rings[r].end());
    last[r] = (int)F.full.size() - 6;
  }
  const std::vector<Pt>& c = F.full;
  std::vector<float> curv;
  curvature(c, curv);
  std::vector<int> order(c.size()), picked(c.size(), 0), label(c.size(), 0);
  for (size_t i = 0; i < c.size(); ++i) order[i] = (int)i;
  std::vector<Pt> ring_less, thin;
  for (int r = 0; r < lines; ++r) {  // :292-413
    if (last[r] - first[r] < 6) continue;
    ring_less.clear();
    for (int j = 0; j < 6; ++j) {
      const int sp = first[r] + (last[r] - first[r]) * j / 6;
      const int ep = first[r] + (last[r] - first[r]) * (j + 1) / 6 - 1;
      if (tie_rule == 1)
        std::sort(order.begin() + sp, order.begin() + ep + 1, [&](int a, int b) { return curv[a] < curv[b]; });
      else
        std::sort(order.begin() + sp, order.begin() + ep + 1,
                  [&](int a, int b) { return curv[a] < curv[b] || (curv[a] == curv[b] && a < b); });
      int largest = 0;
      for (int k = ep; k >= sp; --k) {
        const int ind = order[k];
        if (picked[ind] != 0 || !(curv[ind] > 0.1)) continue;
        ++largest;
        if (largest <= 2) { label[ind] = 2; F.sharp.push_back(c[ind]); F.less_sharp.push_back(c[ind]); }
        else if (largest <= 20) { label[ind] = 1; F.less_sharp.push_back(c[ind]); }
        else break;
        picked[ind] = 1;
        suppress(c, picked, ind);
      }
      int smallest = 0;
      for (int k = sp; k <= ep; ++k) {
        const int ind = order[k];
        if (picked[ind] != 0 || !(curv[ind] < 0.1)) continue;
        label[ind] = -1;
        F.flat.push_back(c[ind]);
        if (++smallest >= 4) break;
        picked[ind] = 1;
        suppress(c, picked, ind);
      }
      for (int k = sp; k <= ep; ++k)
        if (label[k] <= 0) ring_less.push_back(c[k]);
    }
    voxel_thin(ring_less, 0.2f, thin);
    F.less_flat.insert(F.less_flat.end(), thin.begin(), thin.end());
  }
}

int emit(const std::vector<Pt>& v, float* out, int cap, int* n) {
  *n = (int)v.size();
  if ((int)v.size() > cap) return -1;
  if (out && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(Pt));
  return 0;
}

}  // namespace

extern "C" {

// sensor: 0 HDL64, 1 VLP16, 2 OS1-64, 3 HDL32.  Every output holds up to `cap` xyzi points.
// Returns 0, or -1 if an output did not fit (the counts are still written).
int orc_scan_registration(int sensor, double minimum_range, const float* xyz, int n, int tie_rule, int cap,
                          float* full, int* n_full, float* sharp, int* n_sharp, float* less_sharp, int* n_less_sharp,
                          float* flat, int* n_flat, float* less_flat, int* n_less_flat) {
  Features F;
  extract(sensor, minimum_range, xyz, n, tie_rule, F);
  int rc = 0;
  rc |= emit(F.full, full, cap, n_full);
  rc |= emit(F.sharp, sharp, cap, n_sharp);
  rc |= emit(F.less_sharp, less_sharp, cap, n_less_sharp);
  rc |= emit(F.flat, flat, cap, n_flat);
  rc |= emit(F.less_flat, less_flat, cap, n_less_flat);
  return rc;
}

// ring number of each point by the rule of :168-213 (-1 = dropped); used by the KAIST03 check
void orc_ring_of(int sensor, const float* xyz, int n, int* ring) {
  for (int i = 0; i < n; ++i) ring[i] = ring_of(sensor, {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0f});
}

}  // extern "C"
