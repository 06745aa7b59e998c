// oracle/s2m_oracle.cpp -- CPU restatement of SC-A-LOAM's scan-to-map path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under sc-a-loam_b200/ links, loads or calls
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs do, and only as the checker or the timed CPU baseline.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
// this path, and its arithmetic lives in PCL 1.8 / FLANN /
// Eigen 3.3 / Ceres 1.12-2.1, none of which is installed here.
// What IS pinned: the kNN part is checked against the KD-tree the reference
// vendors (include/scancontext/nanoflann.hpp, built into oracle/_ref by
// oracle/Makefile) and against brute force.  One end-to-end anchor on the
// reference's OWN OUTPUT exists, at centimetre (not bit) level: the real OS1-64
// scans and saved poses under utils/sample_data/KAIST03 -- this restatement,
// given the map those poses produce, pulls guesses perturbed by 0.2 m / 1 deg
// back to within 6 cm / 0.45 deg of the pose the reference saved
// (tests/golden/make_kaist03.py, tests/test_golden.py).
//
// What follows the reference line by line (file = /root/reference/src/laserMapping.cpp):
//   A  transformAssociateToMap            :143-147      -> Mapper::associate_to_map
//   U  transformUpdate                    :149-153      -> Mapper::transform_update
//   P  pointAssociateToMap                :155-164      -> Mapper::point_to_map
//   B  centre cube + window shift         :313-508      -> Mapper::shift_window
//   C  valid set + gather                 :510-540      -> Mapper::gather
//   V  VoxelGrid scan / cube filters      :543-551,:788-802 -> voxel_grid (assumption A1)
//   G  guard                              :555
//   T,K KdTreeFLANN build + nearestKSearch :559-560,:583,:649 -> KdTree (A2,A3)
//   E  edge fit                           :585-622      -> eig3_selfadjoint (A4)
//   F  plane fit                          :651-687      -> colpiv_qr_solve_5x3 (A4)
//   R1,R2 lidarFactor.hpp:12-55,:106-138  -> EdgeFactor/PlaneNormFactor on Jet<7>
//   L,Q HuberLoss(0.1), EigenQuaternionParameterization :566-573 (A5)
//   S  ceres::Solve DENSE_QR, 4 iterations :713-721     -> solve_trust_region (A6)
//   O  2 outer iterations                 :563
//   I  map insertion                      :737-784
//   W  re-filter of the valid cubes       :788-802
// Also here: laserOdometry.cpp:220-591 with lidarFactor.hpp:57-104 -> class Odometer
// (three-point PlaneFactor, exact 1-NN, ring-constrained second / third neighbours, same LM).
//
// Numbered assumptions about upstream libraries:
//   A1 pcl::VoxelGrid 1.8: float inverse leaf, floor(p*inv) lattice, key order x
//      fastest, centroid of x,y,z,intensity as sequential float sums / count.
//      std::sort there is unstable; here a STABLE sort defines the summation
//      order (input order inside a voxel).
//   A2 KdTreeFLANN -> flann::KDTreeSingleIndex<L2_Simple<float>>, leaf 15, split
//      rule middleSplit_.
//   A3 nearestKSearch: eps 0, sorted, simple insertion result set, ties kept in
//      visit order. Canonical definition used for parity: brute force ordered by
//      (d2, index), identical whenever the six smallest distances are distinct.
//   A4 Eigen 3.3: Quaternion*Vector3, slerp, SelfAdjointEigenSolver<Matrix3d>
//      (tridiagonal + implicit QR), ColPivHouseholderQR least squares.
//   A5 Ceres HuberLoss / Corrector / EigenQuaternionParameterization.
//   A6 Ceres TrustRegionMinimizer + LevenbergMarquardtStrategy + DenseQRSolver
//      defaults (jacobi scaling, radius 1e4, tolerances 1e-6/1e-10/1e-8).
//
// What checks each assumption today (tests/, "pin" = against code or data of the reference itself):
//   A1  numpy restatement of the lattice / key order / centroid (test_voxel_grid_matches_numpy, edge cases); map bits of
//       CUDA path == oracle over replays.  Not pinned against PCL itself (not installable); the KAIST03 anchor bounds
//       its effect at centimetre level.
//   A2,A3 pin: the reference's vendored nanoflann.hpp compiled from where it lies (oracle/_ref) returns the same
//       neighbours and float distances as the restated FLANN tree and as brute force (test_knn_brute_kdtree_nanoflann_agree,
//       tests/golden/knn_nanoflann.npz); ties: test_knn_ties_are_canonical, test_knn_ties_on_a_lattice_map (GPU).
//   A4  numpy.linalg.eigh / lstsq (test_eig3_matches_numpy, test_plane_qr_matches_lstsq); accepted sets identical between
//       the CUDA path's closed-form / Householder fits and these (test_registration_matches_oracle_on_uploaded_map).
//   A5  Jets through the literal functors vs closed forms (test_factor_autodiff_matches_closed_form); Huber per block:
//       test_lm_matches_an_independent_minimiser_on_a_mixed_robust_problem.
//   A6  schedule (radius, step quality, tolerances, iteration count) identical between this QR-based loop and the
//       product's normal-equation loop (test_lm_schedule_matches_oracle); minimiser vs scipy on plane-only and on mixed
//       edge + plane robust problems (test_lm_converges_like_scipy_on_plane_only_problem, ..._mixed_robust_problem).
//       Not pinned against a Ceres build (not installable).
//   A7,A8 (feature extraction, oracle/scan_registration.cpp): pin for the ring rule on the 765 919 points of the reference's
//       own scans; A7 (float atan2) was corrected in round 2 from exactly that data.
//   End to end pin: the 21 real KAIST03 keyframes and the poses the reference saved for them (tests/test_golden.py,
//       tests/test_odometry.py), at centimetre level.
//
// Build: see oracle/Makefile  (g++ -O3 -ffp-contract=off, no -march, no fast-math,
// like the reference's CMakeLists.txt:7).

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

namespace orc {

struct Pt { float x, y, z, i; };  // pcl::PointXYZI payload, common.h:43
typedef std::vector<Pt> Cloud;

// =========================================================================
// A1. pcl::VoxelGrid<PointXYZI>::filter
// =========================================================================
void voxel_grid(const Cloud& in, float leaf, Cloud& out) {
  out.clear();
  if (in.empty()) return;
  const float inv = 1.0f / leaf;
  float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
  for (const Pt& p : in) {
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {  // "Leaf size is too small": pass-through
    out = in;
    return;
  }
  int minb[3], maxb[3], div[3];
  for (int k = 0; k < 3; ++k) {
    minb[k] = (int)std::floor(mn[k] * inv);
    maxb[k] = (int)std::floor(mx[k] * inv);
    div[k] = maxb[k] - minb[k] + 1;
  }
  const int m1 = div[0], m2 = div[0] * div[1];
  std::vector<std::pair<int, int>> keyed(in.size());
  for (size_t n = 0; n < in.size(); ++n) {
    int i0 = (int)(std::floor(in[n].x * inv) - (float)minb[0]);
    int i1 = (int)(std::floor(in[n].y * inv) - (float)minb[1]);
    int i2 = (int)(std::floor(in[n].z * inv) - (float)minb[2]);
    keyed[n] = {i0 + i1 * m1 + i2 * m2, (int)n};
  }
  std::stable_sort(keyed.begin(), keyed.end(),
                   [](const std::pair<int, int>& a, const std::pair<int, int>& b) {
                     return a.first < b.first;
                   });
  size_t a = 0;
  while (a < keyed.size()) {
    size_t b = a;
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
    while (b < keyed.size() && keyed[b].first == keyed[a].first) {
      const Pt& p = in[keyed[b].second];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      ++b;
    }
    const float c = (float)(b - a);
    out.push_back({sx / c, sy / c, sz / c, si / c});
    a = b;
  }
}

// =========================================================================
// A2/A3. flann::KDTreeSingleIndex<L2_Simple<float>> (leaf 15) + knnSearch(5)
// =========================================================================
inline float l2_simple(const float* a, const float* b) {
  float r = 0.f;
  for (int k = 0; k < 3; ++k) {
    float d = a[k] - b[k];
    r += d * d;
  }
  return r;
}

class KdTree {
 public:
  void build(const Cloud& cloud) {
    n_ = (int)cloud.size();
    pts_.resize((size_t)n_ * 3);
    for (int i = 0; i < n_; ++i) {
      pts_[3 * i] = cloud[i].x; pts_[3 * i + 1] = cloud[i].y; pts_[3 * i + 2] = cloud[i].z;
    }
    vind_.resize(n_);
    for (int i = 0; i < n_; ++i) vind_[i] = i;
    nodes_.clear();
    nodes_.reserve(n_ / 4 + 16);
    if (n_ == 0) return;
    for (int k = 0; k < 3; ++k) root_lo_[k] = root_hi_[k] = pts_[k];
    for (int i = 1; i < n_; ++i)
      for (int k = 0; k < 3; ++k) {
        float v = pts_[3 * i + k];
        if (v < root_lo_[k]) root_lo_[k] = v;
        if (v > root_hi_[k]) root_hi_[k] = v;
      }
    float lo[3] = {root_lo_[0], root_lo_[1], root_lo_[2]};
    float hi[3] = {root_hi_[0], root_hi_[1], root_hi_[2]};
    root_ = divide(0, n_, lo, hi);
    // reorder = true: leaf scans read points contiguously
    ordered_.resize((size_t)n_ * 3);
    for (int i = 0; i < n_; ++i)
      for (int k = 0; k < 3; ++k) ordered_[3 * i + k] = pts_[3 * vind_[i] + k];
  }

  // k nearest of q; returns the number found (min(k, n)). d2 ascending.
  int knn(const float q[3], int k, int* idx, float* d2) const {
    if (n_ == 0) return 0;
    k = std::min(k, n_);
    Result rs{idx, d2, k, 0, std::numeric_limits<float>::max()};
    float dists[3] = {0.f, 0.f, 0.f};
    float distsq = 0.f;
    for (int i = 0; i < 3; ++i) {
      if (q[i] < root_lo_[i]) { dists[i] = (q[i] - root_lo_[i]) * (q[i] - root_lo_[i]); distsq += dists[i]; }
      if (q[i] > root_hi_[i]) { dists[i] = (q[i] - root_hi_[i]) * (q[i] - root_hi_[i]); distsq += dists[i]; }
    }
    search(rs, q, root_, distsq, dists);
    return rs.count;
  }
  int size() const { return n_; }

 private:
  struct Node { int child1, child2, left, right, divfeat; float divlow, divhigh; };
  struct Result {
    int* idx; float* d2; int cap; int count; float worst;
    void add(float dist, int index) {
      if (dist >= worst) return;
      if (count < cap) ++count;
      int i;
      for (i = count - 1; i > 0; --i) {
        if (d2[i - 1] > dist) { d2[i] = d2[i - 1]; idx[i] = idx[i - 1]; }
        else break;
      }
      d2[i] = dist; idx[i] = index;
      if (count == cap) worst = d2[cap - 1];
    }
  };

  void minmax(const int* ind, int count, int dim, float& mn, float& mx) const {
    mn = mx = pts_[3 * ind[0] + dim];
    for (int i = 1; i < count; ++i) {
      float v = pts_[3 * ind[i] + dim];
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
  }
  void plane_split(int* ind, int count, int dim, float cutval, int& lim1, int& lim2) const {
    int left = 0, right = count - 1;
    for (;;) {
      while (left <= right && pts_[3 * ind[left] + dim] < cutval) ++left;
      while (left <= right && pts_[3 * ind[right] + dim] >= cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]); ++left; --right;
    }
    lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && pts_[3 * ind[left] + dim] <= cutval) ++left;
      while (left <= right && pts_[3 * ind[right] + dim] > cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]); ++left; --right;
    }
    lim2 = left;
  }
  void middle_split(int* ind, int count, int& index, int& cutfeat, float& cutval,
                    const float* lo, const float* hi) const {
    const float EPS = 0.00001f;
    float max_span = hi[0] - lo[0];
    for (int i = 1; i < 3; ++i) max_span = std::max(max_span, hi[i] - lo[i]);
    float max_spread = -1.f;
    cutfeat = 0;
    for (int i = 0; i < 3; ++i) {
      float span = hi[i] - lo[i];
      if (span > (1 - EPS) * max_span) {
        float mn, mx;
        minmax(ind, count, i, mn, mx);
        float spread = mx - mn;
        if (spread > max_spread) { cutfeat = i; max_spread = spread; }
      }
    }
    float split_val = (lo[cutfeat] + hi[cutfeat]) / 2;
    float mn, mx;
    minmax(ind, count, cutfeat, mn, mx);
    if (split_val < mn) cutval = mn;
    else if (split_val > mx) cutval = mx;
    else cutval = split_val;
    int lim1, lim2;
    plane_split(ind, count, cutfeat, cutval, lim1, lim2);
    if (lim1 > count / 2) index = lim1;
    else if (lim2 < count / 2) index = lim2;
    else index = count / 2;
  }
  int divide(int left, int right, float* lo, float* hi) {
    int id = (int)nodes_.size();
    nodes_.push_back(Node{-1, -1, 0, 0, 0, 0.f, 0.f});
    if (right - left <= kLeaf) {
      nodes_[id].left = left; nodes_[id].right = right;
      for (int k = 0; k < 3; ++k) lo[k] = hi[k] = pts_[3 * vind_[left] + k];
      for (int i = left + 1; i < right; ++i)
        for (int k = 0; k < 3; ++k) {
          float v = pts_[3 * vind_[i] + k];
          if (lo[k] > v) lo[k] = v;
          if (hi[k] < v) hi[k] = v;
        }
      return id;
    }
    int idx, cutfeat;
    float cutval;
    middle_split(&vind_[left], right - left, idx, cutfeat, cutval, lo, hi);
    float llo[3] = {lo[0], lo[1], lo[2]}, lhi[3] = {hi[0], hi[1], hi[2]};
    lhi[cutfeat] = cutval;
    int c1 = divide(left, left + idx, llo, lhi);
    float rlo[3] = {lo[0], lo[1], lo[2]}, rhi[3] = {hi[0], hi[1], hi[2]};
    rlo[cutfeat] = cutval;
    int c2 = divide(left + idx, right, rlo, rhi);
    Node& nd = nodes_[id];
    nd.child1 = c1; nd.child2 = c2; nd.divfeat = cutfeat;
    nd.divlow = lhi[cutfeat]; nd.divhigh = rlo[cutfeat];
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(llo[k], rlo[k]);
      hi[k] = std::max(lhi[k], rhi[k]);
    }
    return id;
  }
  void search(Result& rs, const float* q, int node, float mindistsq, float* dists) const {
    const Node& nd = nodes_[node];
    if (nd.child1 < 0) {
      float worst = rs.worst;
      for (int i = nd.left; i < nd.right; ++i) {
        float d = l2_simple(q, &ordered_[3 * i]);
        if (d < worst) rs.add(d, vind_[i]);
      }
      return;
    }
    int f = nd.divfeat;
    float val = q[f];
    float diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
    int best, other;
    float cut;
    if (diff1 + diff2 < 0) { best = nd.child1; other = nd.child2; cut = diff2 * diff2; }
    else { best = nd.child2; other = nd.child1; cut = diff1 * diff1; }
    search(rs, q, best, mindistsq, dists);
    float dst = dists[f];
    mindistsq = mindistsq + cut - dst;
    dists[f] = cut;
    if (mindistsq <= rs.worst) search(rs, q, other, mindistsq, dists);
    dists[f] = dst;
  }

  static constexpr int kLeaf = 15;
  int n_ = 0, root_ = 0;
  std::vector<float> pts_, ordered_;
  std::vector<int> vind_;
  std::vector<Node> nodes_;
  float root_lo_[3], root_hi_[3];
};

// Canonical kNN: every float distance, ordered by (d2, index).
int knn_brute(const Cloud& cloud, const float q[3], int k, int* idx, float* d2) {
  int n = (int)cloud.size();
  k = std::min(k, n);
  int count = 0;
  for (int i = 0; i < n; ++i) {
    float p[3] = {cloud[i].x, cloud[i].y, cloud[i].z};
    float d = l2_simple(q, p);
    if (count == k && !(d < d2[k - 1])) continue;  // later equal index never displaces
    int j = (count < k) ? count++ : k - 1;
    for (; j > 0 && d2[j - 1] > d; --j) { d2[j] = d2[j - 1]; idx[j] = idx[j - 1]; }
    d2[j] = d; idx[j] = i;
  }
  return count;
}

// =========================================================================
// A4. Eigen pieces
// =========================================================================
struct Quat { double x, y, z, w; };

// Eigen 3.3 QuaternionBase::_transformVector
template <typename T>
inline void quat_rotate(const T q[4] /*x,y,z,w*/, const T v[3], T out[3]) {
  T uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
  T c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) out[i] = v[i] + q[3] * uv[i] + c[i];
}
inline Quat quat_mul(const Quat& a, const Quat& b) {  // Eigen quat_product (scalar path)
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Quat quat_inverse(const Quat& q) {  // Eigen: conjugate / squaredNorm
  double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
  if (n2 > 0) return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
  return {0, 0, 0, 0};
}

inline void givens(double p, double q, double& c, double& s) {  // JacobiRotation::makeGivens
  if (q == 0) { c = p < 0 ? -1 : 1; s = 0; }
  else if (p == 0) { c = 0; s = q < 0 ? 1 : -1; }
  else if (std::fabs(p) > std::fabs(q)) {
    double t = q / p, u = std::sqrt(1 + t * t);
    if (p < 0) u = -u;
    c = 1 / u; s = -t * c;
  } else {
    double t = p / q, u = std::sqrt(1 + t * t);
    if (q < 0) u = -u;
    s = -1 / u; c = -t * s;
  }
}

// SelfAdjointEigenSolver<Matrix3d>::compute: scale, 3x3 tridiagonalisation,
// implicit symmetric QR with Wilkinson shift, ascending sort. Reads the lower
// triangle of m (row-major 3x3). evecs column-major-by-column: evecs[r][c].
bool eig3_selfadjoint(const double m[9], double evals[3], double evecs[3][3]) {
  double a[3][3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) a[r][c] = (c <= r) ? m[3 * r + c] : 0.0;
  double scale = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c <= r; ++c) scale = std::max(scale, std::fabs(a[r][c]));
  if (scale == 0) scale = 1;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c <= r; ++c) a[r][c] /= scale;
  double diag[3], sub[2], Q[3][3];
  const double tol = DBL_MIN;
  diag[0] = a[0][0];
  double v1norm2 = a[2][0] * a[2][0];
  if (v1norm2 <= tol) {
    diag[1] = a[1][1]; diag[2] = a[2][2]; sub[0] = a[1][0]; sub[1] = a[2][1];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Q[r][c] = (r == c);
  } else {
    double beta = std::sqrt(a[1][0] * a[1][0] + v1norm2);
    double invBeta = 1.0 / beta;
    double m01 = a[1][0] * invBeta, m02 = a[2][0] * invBeta;
    double q = 2.0 * m01 * a[2][1] + m02 * (a[2][2] - a[1][1]);
    diag[1] = a[1][1] + m02 * q;
    diag[2] = a[2][2] - m02 * q;
    sub[0] = beta;
    sub[1] = a[2][1] - m01 * q;
    double Q0[3][3] = {{1, 0, 0}, {0, m01, m02}, {0, m02, -m01}};
    std::memcpy(Q, Q0, sizeof(Q));
  }
  const int n = 3, maxIter = 30;
  int end = n - 1, start = 0, iter = 0;
  const double precision = 2.0 * DBL_EPSILON;
  while (end > 0) {
    for (int i = start; i < end; ++i)
      if (std::fabs(sub[i]) <= (std::fabs(diag[i]) + std::fabs(diag[i + 1])) * precision ||
          std::fabs(sub[i]) <= DBL_MIN)
        sub[i] = 0;
    while (end > 0 && sub[end - 1] == 0) end--;
    if (end <= 0) break;
    iter++;
    if (iter > maxIter * n) break;
    start = end - 1;
    while (start > 0 && sub[start - 1] != 0) start--;
    // tridiagonal_qr_step
    double td = (diag[end - 1] - diag[end]) * 0.5;
    double e = sub[end - 1];
    double mu = diag[end];
    if (td == 0) mu -= std::fabs(e);
    else if (e != 0) {
      double e2 = e * e, h = std::hypot(td, e);
      if (e2 == 0) mu -= e / ((td + (td > 0 ? h : -h)) / e);
      else mu -= e2 / (td + (td > 0 ? h : -h));
    }
    double x = diag[start] - mu, z = sub[start];
    for (int k = start; k < end && z != 0; ++k) {
      double c, s;
      givens(x, z, c, s);
      double sdk = s * diag[k] + c * sub[k];
      double dkp1 = s * sub[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * sub[k]) - s * (c * sub[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      sub[k] = c * sdk - s * dkp1;
      if (k > start) sub[k - 1] = c * sub[k - 1] - s * z;
      x = sub[k];
      if (k < end - 1) { z = -s * sub[k + 1]; sub[k + 1] = c * sub[k + 1]; }
      for (int r = 0; r < 3; ++r) {  // Q = Q * G on columns k,k+1
        double xi = Q[r][k], yi = Q[r][k + 1];
        Q[r][k] = c * xi - s * yi;
        Q[r][k + 1] = s * xi + c * yi;
      }
    }
  }
  bool ok = iter <= maxIter * n;
  if (ok)
    for (int i = 0; i < n - 1; ++i) {
      int k = 0;
      for (int j = 1; j < n - i; ++j) if (diag[i + j] < diag[i + k]) k = j;
      if (k > 0) {
        std::swap(diag[i], diag[k + i]);
        for (int r = 0; r < 3; ++r) std::swap(Q[r][i], Q[r][k + i]);
      }
    }
  for (int i = 0; i < 3; ++i) evals[i] = diag[i] * scale;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) evecs[r][c] = Q[r][c];
  return ok;
}

// makeHouseholderInPlace on v[0..n): returns tau, beta; v[1..) becomes essential.
inline void make_householder(double* v, int n, double& tau, double& beta) {
  double tail2 = 0;
  for (int i = 1; i < n; ++i) tail2 += v[i] * v[i];
  double c0 = v[0];
  if (tail2 <= DBL_MIN) {
    tau = 0; beta = c0;
    for (int i = 1; i < n; ++i) v[i] = 0;
  } else {
    beta = std::sqrt(c0 * c0 + tail2);
    if (c0 >= 0) beta = -beta;
    for (int i = 1; i < n; ++i) v[i] = v[i] / (c0 - beta);
    tau = (beta - c0) / beta;
  }
}

// Eigen::Matrix<double,5,3>::colPivHouseholderQr().solve(b): least squares with
// the LAPACK-style norm down-dating of Eigen 3.3. A row-major 5x3.
void colpiv_qr_solve_5x3(const double A_in[15], const double b_in[5], double x[3]) {
  const int rows = 5, cols = 3, size = 3;
  double qr[5][3];
  for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c) qr[r][c] = A_in[3 * r + c];
  double hcoef[3], normsU[3], normsD[3];
  int transp[3];
  for (int k = 0; k < cols; ++k) {
    double s = 0;
    for (int r = 0; r < rows; ++r) s += qr[r][k] * qr[r][k];
    normsD[k] = normsU[k] = std::sqrt(s);
  }
  double maxn = std::max(normsU[0], std::max(normsU[1], normsU[2]));
  double th = maxn * DBL_EPSILON;
  const double threshold_helper = th * th / (double)rows;
  const double downdate_thr = std::sqrt(DBL_EPSILON);
  int nonzero_pivots = size;
  for (int k = 0; k < size; ++k) {
    int big = k;
    for (int j = k + 1; j < cols; ++j) if (normsU[j] > normsU[big]) big = j;
    double big2 = normsU[big] * normsU[big];
    if (nonzero_pivots == size && big2 < threshold_helper * (double)(rows - k)) nonzero_pivots = k;
    transp[k] = big;
    if (k != big) {
      for (int r = 0; r < rows; ++r) std::swap(qr[r][k], qr[r][big]);
      std::swap(normsU[k], normsU[big]);
      std::swap(normsD[k], normsD[big]);
    }
    double v[5], tau, beta;
    int len = rows - k;
    for (int r = 0; r < len; ++r) v[r] = qr[k + r][k];
    make_householder(v, len, tau, beta);
    for (int r = 1; r < len; ++r) qr[k + r][k] = v[r];
    qr[k][k] = beta;
    hcoef[k] = tau;
    for (int j = k + 1; j < cols; ++j) {  // applyHouseholderOnTheLeft
      double tmp = 0;
      for (int r = 1; r < len; ++r) tmp += v[r] * qr[k + r][j];
      tmp += qr[k][j];
      qr[k][j] -= tau * tmp;
      for (int r = 1; r < len; ++r) qr[k + r][j] -= tau * v[r] * tmp;
    }
    for (int j = k + 1; j < cols; ++j) {
      if (normsU[j] != 0) {
        double temp = std::fabs(qr[k][j]) / normsU[j];
        temp = (1.0 + temp) * (1.0 - temp);
        temp = temp < 0 ? 0 : temp;
        double r2 = normsU[j] / normsD[j];
        double temp2 = temp * r2 * r2;
        if (temp2 <= downdate_thr) {
          double s = 0;
          for (int r = k + 1; r < rows; ++r) s += qr[r][j] * qr[r][j];
          normsD[j] = std::sqrt(s);
          normsU[j] = normsD[j];
        } else {
          normsU[j] *= std::sqrt(temp);
        }
      }
    }
  }
  // permutation indices from the transpositions
  int perm[3] = {0, 1, 2};
  for (int k = 0; k < size; ++k) std::swap(perm[k], perm[transp[k]]);
  x[0] = x[1] = x[2] = 0;
  if (nonzero_pivots == 0) return;
  double c[5];
  for (int r = 0; r < rows; ++r) c[r] = b_in[r];
  for (int k = 0; k < nonzero_pivots; ++k) {  // c = H_k ... H_0 c
    double tmp = c[k];
    for (int r = k + 1; r < rows; ++r) tmp += qr[r][k] * c[r];
    c[k] -= hcoef[k] * tmp;
    for (int r = k + 1; r < rows; ++r) c[r] -= hcoef[k] * qr[r][k] * tmp;
  }
  for (int i = nonzero_pivots - 1; i >= 0; --i) {
    double s = c[i];
    for (int j = i + 1; j < nonzero_pivots; ++j) s -= qr[i][j] * c[j];
    c[i] = s / qr[i][i];
  }
  for (int i = 0; i < nonzero_pivots; ++i) x[perm[i]] = c[i];
}

// =========================================================================
// Ceres Jet<double,7> and the two cost functors (lidarFactor.hpp)
// =========================================================================
struct Jet {
  double a;
  double v[7];
  Jet() : a(0) { for (double& d : v) d = 0; }
  explicit Jet(double s) : a(s) { for (double& d : v) d = 0; }
  Jet(double s, int k) : a(s) { for (double& d : v) d = 0; v[k] = 1.0; }
};
inline Jet operator+(const Jet& f, const Jet& g) { Jet h; h.a = f.a + g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
inline Jet operator-(const Jet& f, const Jet& g) { Jet h; h.a = f.a - g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
inline Jet operator-(const Jet& f) { Jet h; h.a = -f.a; for (int i = 0; i < 7; ++i) h.v[i] = -f.v[i]; return h; }
inline Jet operator*(const Jet& f, const Jet& g) { Jet h; h.a = f.a * g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
inline Jet operator/(const Jet& f, const Jet& g) {
  Jet h;
  const double ginv = 1.0 / g.a, fg = f.a * ginv;
  h.a = fg;
  for (int i = 0; i < 7; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * ginv;
  return h;
}
inline Jet jsqrt(const Jet& f) { Jet h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] * t; return h; }
inline Jet jacos(const Jet& f) { Jet h; h.a = std::acos(f.a); const double t = -1.0 / std::sqrt(1.0 - f.a * f.a); for (int i = 0; i < 7; ++i) h.v[i] = t * f.v[i]; return h; }
inline Jet jsin(const Jet& f) { Jet h; h.a = std::sin(f.a); const double t = std::cos(f.a); for (int i = 0; i < 7; ++i) h.v[i] = t * f.v[i]; return h; }
inline Jet jabs(const Jet& f) { return f.a < 0 ? -f : f; }

// scalar overloads so the functor templates read like lidarFactor.hpp
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jacos(double x) { return std::acos(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jabs(double x) { return std::fabs(x); }
inline double scalar_of(double x) { return x; }
inline double scalar_of(const Jet& x) { return x.a; }

// Eigen 3.3 QuaternionBase::slerp, coefficients stored x,y,z,w.
template <typename T>
inline void quat_slerp(const T a[4], const T& t, const T b[4], T out[4]) {
  const T one = T(1.0) - T(DBL_EPSILON);
  T d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
  T absD = jabs(d);
  T scale0, scale1;
  if (scalar_of(absD) >= scalar_of(one)) {
    scale0 = T(1.0) - t;
    scale1 = t;
  } else {
    T theta = jacos(absD);
    T sinTheta = jsin(theta);
    scale0 = jsin((T(1.0) - t) * theta) / sinTheta;
    scale1 = jsin(t * theta) / sinTheta;
  }
  if (scalar_of(d) < 0) scale1 = -scale1;
  for (int i = 0; i < 4; ++i) out[i] = scale0 * a[i] + scale1 * b[i];
}

struct EdgeFactor {  // lidarFactor.hpp:12-55
  double cp[3], lpa[3], lpb[3], s;
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    T cpT[3] = {T(cp[0]), T(cp[1]), T(cp[2])};
    T a[3] = {T(lpa[0]), T(lpa[1]), T(lpa[2])};
    T b[3] = {T(lpb[0]), T(lpb[1]), T(lpb[2])};
    T q_last_curr[4] = {q[0], q[1], q[2], q[3]};
    T q_identity[4] = {T(0.0), T(0.0), T(0.0), T(1.0)};
    T qs[4];
    quat_slerp(q_identity, T(s), q_last_curr, qs);
    T tl[3] = {T(s) * t[0], T(s) * t[1], T(s) * t[2]};
    T lp[3];
    quat_rotate(qs, cpT, lp);
    for (int i = 0; i < 3; ++i) lp[i] = lp[i] + tl[i];
    T u[3] = {lp[0] - a[0], lp[1] - a[1], lp[2] - a[2]};
    T w[3] = {lp[0] - b[0], lp[1] - b[1], lp[2] - b[2]};
    T nu[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    T de[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    T den = jsqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
    residual[0] = nu[0] / den;
    residual[1] = nu[1] / den;
    residual[2] = nu[2] / den;
    return true;
  }
};

struct PlaneNormFactor {  // lidarFactor.hpp:106-138
  double cp[3], n[3], d;
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    T cpT[3] = {T(cp[0]), T(cp[1]), T(cp[2])};
    T pw[3];
    quat_rotate(q, cpT, pw);
    for (int i = 0; i < 3; ++i) pw[i] = pw[i] + t[i];
    residual[0] = T(n[0]) * pw[0] + T(n[1]) * pw[1] + T(n[2]) * pw[2] + T(d);
    return true;
  }
};

struct PlaneFactor {  // lidarFactor.hpp:57-104 (laserOdometry's plane factor: three points of the last sweep)
  double cp[3], lpj[3], ljm[3], s;
  void set(const double cp_[3], const double j[3], const double l[3], const double m[3], double s_) {
    const double a[3] = {j[0] - l[0], j[1] - l[1], j[2] - l[2]}, b[3] = {j[0] - m[0], j[1] - m[1], j[2] - m[2]};
    double n[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};  // :64
    const double z = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];  // Eigen normalize(): divide by the norm when squaredNorm > 0
    if (z > 0.0) { const double inv = std::sqrt(z); for (int k = 0; k < 3; ++k) n[k] /= inv; }
    for (int k = 0; k < 3; ++k) { cp[k] = cp_[k]; lpj[k] = j[k]; ljm[k] = n[k]; }
    s = s_;
  }
  template <typename T>
  bool operator()(const T* q, const T* t, T* residual) const {
    T cpT[3] = {T(cp[0]), T(cp[1]), T(cp[2])};
    T q_last_curr[4] = {q[0], q[1], q[2], q[3]};
    T q_identity[4] = {T(0.0), T(0.0), T(0.0), T(1.0)};
    T qs[4];
    quat_slerp(q_identity, T(s), q_last_curr, qs);
    T lp[3];
    quat_rotate(qs, cpT, lp);
    for (int i = 0; i < 3; ++i) lp[i] = lp[i] + T(s) * t[i];
    residual[0] = (lp[0] - T(lpj[0])) * T(ljm[0]) + (lp[1] - T(lpj[1])) * T(ljm[1]) + (lp[2] - T(lpj[2])) * T(ljm[2]);
    return true;
  }
};

// =========================================================================
// A5/A6. Problem, evaluation, trust-region LM with dense QR
// =========================================================================
struct Block {
  int kind;  // 0 edge (3 residuals), 1 plane-norm (1 residual), 2 three-point plane (1 residual)
  EdgeFactor e;
  PlaneNormFactor p;
  PlaneFactor p3;
};

struct IterLog { double cost, cost_change, radius, step_norm, model_change; int accepted; };
struct SolveLog {
  int n_iter = 0;
  int termination = 0;  // 0 max iterations, 1 gradient, 2 parameter tol, 3 function tol, 4 radius, 5 no residuals
  double initial_cost = 0, final_cost = 0;
  double init_sums[28];  // upper-tri JtJ (21) + Jtr (6) + cost at x0, unscaled tangent space
  IterLog it[8];
};

inline void huber(double s, double rho[3]) {  // ceres::HuberLoss(0.1)
  const double a = 0.1, b = a * a;
  if (s > b) {
    const double r = std::sqrt(s);
    rho[0] = 2.0 * a * r - b;
    rho[1] = std::max(DBL_MIN, a / r);
    rho[2] = -rho[1] / (2.0 * s);
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

inline void quat_plus(const double x[4], const double delta[3], double out[4]) {
  const double nd = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    const double sbd = std::sin(nd) / nd;
    Quat dq{sbd * delta[0], sbd * delta[1], sbd * delta[2], std::cos(nd)};
    Quat r = quat_mul(dq, Quat{x[0], x[1], x[2], x[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
  } else { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; }
}
inline void plus7(const double x[7], const double d[6], double out[7]) {
  quat_plus(x, d, out);
  for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i];
}

// Evaluate cost and optionally robustified residuals + tangent-space Jacobian
// (row-major N x 6) + gradient, in block insertion order like a 1-thread Ceres.
double evaluate(const std::vector<Block>& blocks, const double x[7], std::vector<double>* res,
                std::vector<double>* jac, double* grad) {
  const bool want = (res != nullptr);
  int nres = 0;
  for (const Block& b : blocks) nres += (b.kind == 0) ? 3 : 1;
  if (want) {
    res->assign(nres, 0.0);
    if (jac) jac->assign((size_t)nres * 6, 0.0);
    if (grad) for (int i = 0; i < 6; ++i) grad[i] = 0;
  }
  // EigenQuaternionParameterization::ComputeJacobian (4x3 row-major)
  const double P[12] = {x[3], x[2], -x[1], -x[2], x[3], x[0], x[1], -x[0], x[3], -x[0], -x[1], -x[2]};
  double cost = 0;
  int row = 0;
  for (const Block& b : blocks) {
    const int nr = (b.kind == 0) ? 3 : 1;
    double r[3];
    double J7[3][7];
    if (want && jac) {
      Jet q[4], t[3], rj[3];
      for (int i = 0; i < 4; ++i) q[i] = Jet(x[i], i);
      for (int i = 0; i < 3; ++i) t[i] = Jet(x[4 + i], 4 + i);
      if (b.kind == 0) b.e(q, t, rj); else if (b.kind == 1) b.p(q, t, rj); else b.p3(q, t, rj);
      for (int k = 0; k < nr; ++k) { r[k] = rj[k].a; for (int i = 0; i < 7; ++i) J7[k][i] = rj[k].v[i]; }
    } else {
      if (b.kind == 0) b.e(x, x + 4, r); else if (b.kind == 1) b.p(x, x + 4, r); else b.p3(x, x + 4, r);
    }
    double s = 0;
    for (int k = 0; k < nr; ++k) s += r[k] * r[k];
    double rho[3];
    huber(s, rho);
    cost += 0.5 * rho[0];
    if (want) {
      const double sq = std::sqrt(rho[1]);  // Corrector with rho'' <= 0
      for (int k = 0; k < nr; ++k) {
        if (jac) {
          double* Jr = &(*jac)[(size_t)(row + k) * 6];
          for (int c = 0; c < 3; ++c) {
            double acc = 0;
            for (int i = 0; i < 4; ++i) acc += J7[k][i] * P[3 * i + c];
            Jr[c] = acc * sq;
          }
          for (int c = 0; c < 3; ++c) Jr[3 + c] = J7[k][4 + c] * sq;
        }
        (*res)[row + k] = r[k] * sq;
      }
      if (jac && grad)
        for (int k = 0; k < nr; ++k)
          for (int c = 0; c < 6; ++c) grad[c] += (*jac)[(size_t)(row + k) * 6 + c] * (*res)[row + k];
    }
    row += nr;
  }
  return cost;
}

// Householder QR least squares (Eigen householderQr().solve) of the (N+6)x6 stack.
void dense_qr_solve(std::vector<double>& A, int rows, std::vector<double>& rhs, double y[6]) {
  const int cols = 6;
  std::vector<double> v(rows);
  for (int k = 0; k < cols; ++k) {
    int len = rows - k;
    for (int r = 0; r < len; ++r) v[r] = A[(size_t)(k + r) * cols + k];
    double tau, beta;
    make_householder(v.data(), len, tau, beta);
    A[(size_t)k * cols + k] = beta;
    for (int r = 1; r < len; ++r) A[(size_t)(k + r) * cols + k] = v[r];
    for (int j = k + 1; j < cols; ++j) {
      double tmp = 0;
      for (int r = 1; r < len; ++r) tmp += v[r] * A[(size_t)(k + r) * cols + j];
      tmp += A[(size_t)k * cols + j];
      A[(size_t)k * cols + j] -= tau * tmp;
      for (int r = 1; r < len; ++r) A[(size_t)(k + r) * cols + j] -= tau * v[r] * tmp;
    }
    double tmp = 0;
    for (int r = 1; r < len; ++r) tmp += v[r] * rhs[k + r];
    tmp += rhs[k];
    rhs[k] -= tau * tmp;
    for (int r = 1; r < len; ++r) rhs[k + r] -= tau * v[r] * tmp;
  }
  for (int i = cols - 1; i >= 0; --i) {
    double s = rhs[i];
    for (int j = i + 1; j < cols; ++j) s -= A[(size_t)i * cols + j] * y[j];
    y[i] = s / A[(size_t)i * cols + i];
  }
}

void solve_trust_region(const std::vector<Block>& blocks, double x[7], int max_iterations,
                        SolveLog* log) {
  SolveLog local;
  SolveLog& L = log ? *log : local;
  L = SolveLog();
  if (blocks.empty()) { L.termination = 5; return; }
  std::vector<double> res, jac, model(0);
  double grad[6];
  double cost = evaluate(blocks, x, &res, &jac, grad);
  const int N = (int)res.size();
  L.initial_cost = cost;
  {  // record the unscaled normal equations at x0 (what the GPU path reduces)
    int k = 0;
    for (int a = 0; a < 6; ++a)
      for (int b = a; b < 6; ++b) {
        double s = 0;
        for (int r = 0; r < N; ++r) s += jac[(size_t)r * 6 + a] * jac[(size_t)r * 6 + b];
        L.init_sums[k++] = s;
      }
    for (int a = 0; a < 6; ++a) L.init_sums[21 + a] = grad[a];
    L.init_sums[27] = cost;
  }
  double scale[6];
  for (int c = 0; c < 6; ++c) {
    double s = 0;
    for (int r = 0; r < N; ++r) s += jac[(size_t)r * 6 + c] * jac[(size_t)r * 6 + c];
    scale[c] = 1.0 / (1.0 + std::sqrt(s));
  }
  auto scale_jac = [&]() { for (int r = 0; r < N; ++r) for (int c = 0; c < 6; ++c) jac[(size_t)r * 6 + c] *= scale[c]; };
  auto grad_max_norm = [&]() {
    double neg[6], xp[7], m = 0;
    for (int i = 0; i < 6; ++i) neg[i] = -grad[i];
    plus7(x, neg, xp);
    for (int i = 0; i < 7; ++i) m = std::max(m, std::fabs(x[i] - xp[i]));
    return m;
  };
  scale_jac();
  double gmax = grad_max_norm();
  double x_norm = 0;
  for (int i = 0; i < 7; ++i) x_norm += x[i] * x[i];
  x_norm = std::sqrt(x_norm);
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diag = false;
  double diag[6];
  int iteration = 0, invalid_run = 0;
  L.final_cost = cost;
  for (;;) {
    if (iteration >= max_iterations) { L.termination = 0; break; }
    if (gmax <= 1e-10) { L.termination = 1; break; }
    if (radius <= 1e-32) { L.termination = 4; break; }
    ++iteration;
    IterLog& IL = L.it[std::min(iteration - 1, 7)];
    IL = IterLog{cost, 0, radius, 0, 0, 0};
    L.n_iter = iteration;
    // LevenbergMarquardtStrategy::ComputeStep
    if (!reuse_diag)
      for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (int r = 0; r < N; ++r) s += jac[(size_t)r * 6 + c] * jac[(size_t)r * 6 + c];
        diag[c] = std::min(std::max(s, 1e-6), 1e32);
      }
    std::vector<double> A((size_t)(N + 6) * 6, 0.0), rhs(N + 6, 0.0);
    std::memcpy(A.data(), jac.data(), sizeof(double) * (size_t)N * 6);
    for (int c = 0; c < 6; ++c) A[(size_t)(N + c) * 6 + c] = std::sqrt(diag[c] / radius);
    std::memcpy(rhs.data(), res.data(), sizeof(double) * N);
    double step[6];
    dense_qr_solve(A, N + 6, rhs, step);
    bool finite = true;
    for (int c = 0; c < 6; ++c) { step[c] = -step[c]; finite = finite && std::isfinite(step[c]); }
    reuse_diag = true;
    // model cost change
    double mcc = 0;
    if (finite) {
      for (int r = 0; r < N; ++r) {
        double m = 0;
        for (int c = 0; c < 6; ++c) m += jac[(size_t)r * 6 + c] * step[c];
        mcc += m * (res[r] + m / 2.0);
      }
      mcc = -mcc;
    }
    IL.model_change = mcc;
    if (!finite || !(mcc > 0.0)) {  // invalid step
      if (++invalid_run >= 5) { L.termination = 6; break; }
      radius = radius / decrease_factor;  // StepIsInvalid() == StepRejected(0)
      decrease_factor *= 2.0;
      reuse_diag = true;
      continue;
    }
    invalid_run = 0;
    double delta[6], xc[7];
    for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];
    plus7(x, delta, xc);
    double cand_cost = evaluate(blocks, xc, nullptr, nullptr, nullptr);
    double sn = 0;
    for (int i = 0; i < 7; ++i) sn += (x[i] - xc[i]) * (x[i] - xc[i]);
    sn = std::sqrt(sn);
    IL.step_norm = sn;
    if (sn <= 1e-8 * (x_norm + 1e-8)) { L.termination = 2; break; }
    IL.cost_change = cost - cand_cost;
    if (std::fabs(IL.cost_change) <= 1e-6 * cost) { L.termination = 3; break; }
    double rel = IL.cost_change / mcc;
    if (rel > 1e-3) {
      IL.accepted = 1;
      for (int i = 0; i < 7; ++i) x[i] = xc[i];
      x_norm = 0;
      for (int i = 0; i < 7; ++i) x_norm += x[i] * x[i];
      x_norm = std::sqrt(x_norm);
      cost = evaluate(blocks, x, &res, &jac, grad);
      scale_jac();
      gmax = grad_max_norm();
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(1e16, radius);
      decrease_factor = 2.0;
      reuse_diag = false;
      L.final_cost = cost;
    } else {
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diag = true;
    }
  }
}

// =========================================================================
// The mapper: process() rows A..W
// =========================================================================
struct Stats {
  int n_corner_in, n_surf_in, n_corner_ds, n_surf_ds, n_map_corner, n_map_surf;
  int n_edge[2], n_plane[2];
  int optimized;       // guard G passed
  int lm_iters[2];
  int lm_term[2];
  double cost_initial[2], cost_final[2];
  double cand_corner, cand_surf;  // mean map points in the 27 one-metre cells of a query (outer 0)
  double t_ms[8];                 // shift, tree, data, solver, add, filter, whole, (spare)
};

struct Trace {  // optional per-registration trace for parity tests
  bool on = false;
  Cloud corner_ds, surf_ds, corner_map, surf_map;
  std::vector<int> idx[2][2];    // [outer][class] 5 per query
  std::vector<float> d2[2][2];
  std::vector<unsigned char> used[2][2];
  double pose_after[2][7];
  SolveLog lm[2];
};

class Mapper {
 public:
  static constexpr int W = 21, H = 21, D = 11, NUM = W * H * D;
  Mapper(float line_res, float plane_res) : line_res_(line_res), plane_res_(plane_res),
      corner_(NUM), surf_(NUM) {}

  float line_res_, plane_res_;
  int cenW = 10, cenH = 10, cenD = 5;
  std::vector<Cloud> corner_, surf_;
  double par[7] = {0, 0, 0, 1, 0, 0, 0};  // q_w_curr (x,y,z,w), t_w_curr
  Quat q_wmap_wodom{0, 0, 0, 1};
  double t_wmap_wodom[3] = {0, 0, 0};
  Quat q_wodom{0, 0, 0, 1};
  double t_wodom[3] = {0, 0, 0};
  int valid_[125], n_valid_ = 0;
  bool use_kdtree = true, skip_opt = false;
  Trace trace;
  int frame = 0;

  void associate_to_map() {  // :143-147
    Quat q = quat_mul(q_wmap_wodom, q_wodom);
    par[0] = q.x; par[1] = q.y; par[2] = q.z; par[3] = q.w;
    double qq[4] = {q_wmap_wodom.x, q_wmap_wodom.y, q_wmap_wodom.z, q_wmap_wodom.w}, r[3];
    quat_rotate(qq, t_wodom, r);
    for (int i = 0; i < 3; ++i) par[4 + i] = r[i] + t_wmap_wodom[i];
  }
  void transform_update() {  // :149-153
    Quat qw{par[0], par[1], par[2], par[3]};
    q_wmap_wodom = quat_mul(qw, quat_inverse(q_wodom));
    double qq[4] = {q_wmap_wodom.x, q_wmap_wodom.y, q_wmap_wodom.z, q_wmap_wodom.w}, r[3];
    quat_rotate(qq, t_wodom, r);
    for (int i = 0; i < 3; ++i) t_wmap_wodom[i] = par[4 + i] - r[i];
  }
  void point_to_map(const Pt& pi, Pt& po) const {  // :155-164
    double p[3] = {pi.x, pi.y, pi.z}, r[3];
    quat_rotate(par, p, r);
    po.x = (float)(r[0] + par[4]);
    po.y = (float)(r[1] + par[5]);
    po.z = (float)(r[2] + par[6]);
    po.i = pi.i;
  }
  static int cube_of(double v, int cen) {  // :313-322 / :742-751
    int c = int((v + 25.0) / 50.0) + cen;
    if (v + 25.0 < 0) c--;
    return c;
  }
  int ind(int i, int j, int k) const { return i + W * j + W * H * k; }

  void shift_axis(int axis, int dir) {
    // dir=+1: contents move towards higher index (reference "centerCube < 3" branch)
    const int n[3] = {W, H, D};
    for (int a = 0; a < n[(axis + 1) % 3]; ++a)
      for (int b = 0; b < n[(axis + 2) % 3]; ++b) {
        auto at = [&](int t) {
          int c[3];
          c[axis] = t; c[(axis + 1) % 3] = a; c[(axis + 2) % 3] = b;
          return ind(c[0], c[1], c[2]);
        };
        const int last = n[axis] - 1;
        if (dir > 0) {
          Cloud hc; hc.swap(corner_[at(last)]);
          Cloud hs; hs.swap(surf_[at(last)]);
          for (int t = last; t >= 1; --t) { corner_[at(t)].swap(corner_[at(t - 1)]); surf_[at(t)].swap(surf_[at(t - 1)]); }
          corner_[at(0)].clear(); surf_[at(0)].clear();
        } else {
          for (int t = 0; t < last; ++t) { corner_[at(t)].swap(corner_[at(t + 1)]); surf_[at(t)].swap(surf_[at(t + 1)]); }
          corner_[at(last)].clear(); surf_[at(last)].clear();
        }
      }
  }
  void shift_window(int& ci, int& cj, int& ck) {  // :313-508
    ci = cube_of(par[4], cenW); cj = cube_of(par[5], cenH); ck = cube_of(par[6], cenD);
    while (ci < 3) { shift_axis(0, +1); ci++; cenW++; }
    while (ci >= W - 3) { shift_axis(0, -1); ci--; cenW--; }
    while (cj < 3) { shift_axis(1, +1); cj++; cenH++; }
    while (cj >= H - 3) { shift_axis(1, -1); cj--; cenH--; }
    while (ck < 3) { shift_axis(2, +1); ck++; cenD++; }
    while (ck >= D - 3) { shift_axis(2, -1); ck--; cenD--; }
  }
  void gather(int ci, int cj, int ck, Cloud& cm, Cloud& sm) {  // :510-540
    n_valid_ = 0;
    for (int i = ci - 2; i <= ci + 2; i++)
      for (int j = cj - 2; j <= cj + 2; j++)
        for (int k = ck - 1; k <= ck + 1; k++)
          if (i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D) valid_[n_valid_++] = ind(i, j, k);
    cm.clear(); sm.clear();
    for (int v = 0; v < n_valid_; ++v) {
      cm.insert(cm.end(), corner_[valid_[v]].begin(), corner_[valid_[v]].end());
      sm.insert(sm.end(), surf_[valid_[v]].begin(), surf_[valid_[v]].end());
    }
  }
  void insert(const Cloud& stack, std::vector<Cloud>& arr) {  // :738-784
    for (const Pt& p : stack) {
      Pt s;
      point_to_map(p, s);
      int ci = cube_of((double)s.x, cenW), cj = cube_of((double)s.y, cenH), ck = cube_of((double)s.z, cenD);
      if (ci >= 0 && ci < W && cj >= 0 && cj < H && ck >= 0 && ck < D) arr[ind(ci, cj, ck)].push_back(s);
    }
  }

  int process(const Cloud& corner_last, const Cloud& surf_last, const double q_odom[4],
              const double t_odom[3], Stats* st);
};

static double now_ms() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static double mean_candidates(const Cloud& map, const Cloud& queries_world) {
  // mean number of map points inside the 27 one-metre cells around each query
  // (SURVEY 8d's C-bar): counted with a sorted cell-key list.
  if (map.empty() || queries_world.empty()) return 0.0;
  auto cell = [](float v) { return (int64_t)std::floor(v) + (1 << 20); };
  auto key = [](int64_t cx, int64_t cy, int64_t cz) { return (cz << 42) | (cy << 21) | cx; };
  std::vector<int64_t> keys(map.size());
  for (size_t i = 0; i < map.size(); ++i) keys[i] = key(cell(map[i].x), cell(map[i].y), cell(map[i].z));
  std::sort(keys.begin(), keys.end());
  double total = 0;
  for (const Pt& q : queries_world) {
    int64_t cx = cell(q.x), cy = cell(q.y), cz = cell(q.z);
    for (int dz = -1; dz <= 1; ++dz)
      for (int dy = -1; dy <= 1; ++dy) {
        // the three x-adjacent cells are consecutive keys
        auto lo = std::lower_bound(keys.begin(), keys.end(), key(cx - 1, cy + dy, cz + dz));
        auto hi = std::upper_bound(keys.begin(), keys.end(), key(cx + 1, cy + dy, cz + dz));
        total += (double)(hi - lo);
      }
  }
  return total / (double)queries_world.size();
}

int Mapper::process(const Cloud& corner_last, const Cloud& surf_last, const double q_odom[4],
                    const double t_odom[3], Stats* st) {
  Stats S;
  std::memset(&S, 0, sizeof(S));
  const double t_whole = now_ms();
  q_wodom = Quat{q_odom[0], q_odom[1], q_odom[2], q_odom[3]};
  for (int i = 0; i < 3; ++i) t_wodom[i] = t_odom[i];
  associate_to_map();  // :310

  double t0 = now_ms();
  int ci, cj, ck;
  shift_window(ci, cj, ck);
  Cloud corner_map, surf_map;
  gather(ci, cj, ck, corner_map, surf_map);
  Cloud corner_ds, surf_ds;
  voxel_grid(corner_last, line_res_, corner_ds);   // :543-546
  voxel_grid(surf_last, plane_res_, surf_ds);      // :548-551
  S.t_ms[0] = now_ms() - t0;
  S.n_corner_in = (int)corner_last.size(); S.n_surf_in = (int)surf_last.size();
  S.n_corner_ds = (int)corner_ds.size(); S.n_surf_ds = (int)surf_ds.size();
  S.n_map_corner = (int)corner_map.size(); S.n_map_surf = (int)surf_map.size();
  if (trace.on) {
    trace.corner_ds = corner_ds; trace.surf_ds = surf_ds;
    trace.corner_map = corner_map; trace.surf_map = surf_map;
    for (int o = 0; o < 2; ++o)
      for (int c = 0; c < 2; ++c) { trace.idx[o][c].clear(); trace.d2[o][c].clear(); trace.used[o][c].clear(); }
  }

  int status = 1;  // map too small
  if (!skip_opt && corner_map.size() > 10 && surf_map.size() > 50) {  // :555
    status = 0;
    S.optimized = 1;
    t0 = now_ms();
    KdTree kd_corner, kd_surf;
    if (use_kdtree) { kd_corner.build(corner_map); kd_surf.build(surf_map); }  // :559-560
    S.t_ms[1] = now_ms() - t0;
    for (int iter = 0; iter < 2; ++iter) {  // :563
      t0 = now_ms();
      std::vector<Block> blocks;
      blocks.reserve(corner_ds.size() + surf_ds.size());
      int idx[5];
      float d2[5];
      if (iter == 0 && st) {
        Cloud qc(corner_ds.size()), qs(surf_ds.size());
        for (size_t i = 0; i < corner_ds.size(); ++i) point_to_map(corner_ds[i], qc[i]);
        for (size_t i = 0; i < surf_ds.size(); ++i) point_to_map(surf_ds[i], qs[i]);
        S.cand_corner = mean_candidates(corner_map, qc);
        S.cand_surf = mean_candidates(surf_map, qs);
        t0 = now_ms();
      }
      for (size_t i = 0; i < corner_ds.size(); ++i) {  // :578-641
        const Pt& ori = corner_ds[i];
        Pt sel;
        point_to_map(ori, sel);
        float q[3] = {sel.x, sel.y, sel.z};
        int found = use_kdtree ? kd_corner.knn(q, 5, idx, d2) : knn_brute(corner_map, q, 5, idx, d2);
        bool used = false;
        if (found == 5 && d2[4] < 1.0) {
          double near[5][3], center[3] = {0, 0, 0};
          for (int j = 0; j < 5; ++j) {
            near[j][0] = corner_map[idx[j]].x; near[j][1] = corner_map[idx[j]].y; near[j][2] = corner_map[idx[j]].z;
            for (int k = 0; k < 3; ++k) center[k] = center[k] + near[j][k];
          }
          for (int k = 0; k < 3; ++k) center[k] = center[k] / 5.0;
          double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
          for (int j = 0; j < 5; ++j) {
            double zm[3] = {near[j][0] - center[0], near[j][1] - center[1], near[j][2] - center[2]};
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[3 * r + c] = cov[3 * r + c] + zm[r] * zm[c];
          }
          double ev[3], evec[3][3];
          eig3_selfadjoint(cov, ev, evec);
          if (ev[2] > 3 * ev[1]) {  // :612
            Block b;
            b.kind = 0;
            b.e.cp[0] = ori.x; b.e.cp[1] = ori.y; b.e.cp[2] = ori.z;
            for (int k = 0; k < 3; ++k) {
              b.e.lpa[k] = 0.1 * evec[k][2] + center[k];
              b.e.lpb[k] = -0.1 * evec[k][2] + center[k];
            }
            b.e.s = 1.0;
            blocks.push_back(b);
            used = true;
            S.n_edge[iter]++;
          }
        }
        if (trace.on) {
          for (int j = 0; j < 5; ++j) { trace.idx[iter][0].push_back(j < found ? idx[j] : -1); trace.d2[iter][0].push_back(j < found ? d2[j] : INFINITY); }
          trace.used[iter][0].push_back(used);
        }
      }
      for (size_t i = 0; i < surf_ds.size(); ++i) {  // :644-706
        const Pt& ori = surf_ds[i];
        Pt sel;
        point_to_map(ori, sel);
        float q[3] = {sel.x, sel.y, sel.z};
        int found = use_kdtree ? kd_surf.knn(q, 5, idx, d2) : knn_brute(surf_map, q, 5, idx, d2);
        bool used = false;
        if (found == 5 && d2[4] < 1.0) {
          double A[15], b5[5] = {-1, -1, -1, -1, -1};
          for (int j = 0; j < 5; ++j) {
            A[3 * j] = surf_map[idx[j]].x; A[3 * j + 1] = surf_map[idx[j]].y; A[3 * j + 2] = surf_map[idx[j]].z;
          }
          double norm[3];
          colpiv_qr_solve_5x3(A, b5, norm);
          double nn = std::sqrt(norm[0] * norm[0] + norm[1] * norm[1] + norm[2] * norm[2]);
          double negative_OA_dot_norm = 1 / nn;
          for (int k = 0; k < 3; ++k) norm[k] = norm[k] / nn;  // Eigen normalize(): v /= norm
          bool planeValid = true;
          for (int j = 0; j < 5; ++j)
            if (std::fabs(norm[0] * surf_map[idx[j]].x + norm[1] * surf_map[idx[j]].y +
                          norm[2] * surf_map[idx[j]].z + negative_OA_dot_norm) > 0.2) {
              planeValid = false;
              break;
            }
          if (planeValid) {
            Block b;
            b.kind = 1;
            b.p.cp[0] = ori.x; b.p.cp[1] = ori.y; b.p.cp[2] = ori.z;
            for (int k = 0; k < 3; ++k) b.p.n[k] = norm[k];
            b.p.d = negative_OA_dot_norm;
            blocks.push_back(b);
            used = true;
            S.n_plane[iter]++;
          }
        }
        if (trace.on) {
          for (int j = 0; j < 5; ++j) { trace.idx[iter][1].push_back(j < found ? idx[j] : -1); trace.d2[iter][1].push_back(j < found ? d2[j] : INFINITY); }
          trace.used[iter][1].push_back(used);
        }
      }
      S.t_ms[2] += now_ms() - t0;
      t0 = now_ms();
      SolveLog lg;
      solve_trust_region(blocks, par, 4, &lg);  // :713-721
      S.t_ms[3] += now_ms() - t0;
      S.lm_iters[iter] = lg.n_iter; S.lm_term[iter] = lg.termination;
      S.cost_initial[iter] = lg.initial_cost; S.cost_final[iter] = lg.final_cost;
      if (trace.on) { trace.lm[iter] = lg; std::memcpy(trace.pose_after[iter], par, sizeof(par)); }
    }
  }
  transform_update();  // :735

  t0 = now_ms();
  insert(corner_ds, corner_);
  insert(surf_ds, surf_);
  S.t_ms[4] = now_ms() - t0;
  t0 = now_ms();
  for (int v = 0; v < n_valid_; ++v) {  // :788-802
    Cloud tmp;
    voxel_grid(corner_[valid_[v]], line_res_, tmp);
    corner_[valid_[v]].swap(tmp);
    voxel_grid(surf_[valid_[v]], plane_res_, tmp);
    surf_[valid_[v]].swap(tmp);
  }
  S.t_ms[5] = now_ms() - t0;
  S.t_ms[6] = now_ms() - t_whole;
  frame++;
  if (st) *st = S;
  return status;
}

// =========================================================================
// laserOdometry.cpp:220-591 restated (scan-to-scan odometry).
// DISTORTION = 0 (:57), so s = 1 everywhere and TransformToStart (:108-126) is q*p + t.
// nearestKSearch(., 1, ...) (:303, :392): exact nearest neighbour; canonical tie rule (d2, index).
// =========================================================================
class Odometer {
 public:
  double para[7] = {0, 0, 0, 1, 0, 0, 0};   // para_q (x,y,z,w), para_t (:94-95): q_last_curr, t_last_curr
  Quat q_w{0, 0, 0, 1};                      // q_w_curr, t_w_curr (:90-91)
  double t_w[3] = {0, 0, 0};
  bool inited = false;                       // systemInited (:67)
  Cloud corner_last, surf_last;              // laserCloudCornerLast / SurfLast (:83-84)
  int n_corner[2] = {0, 0}, n_plane[2] = {0, 0};
  std::vector<int> tr_edge[2], tr_plane[2];  // per optimisation pass: closest / second / third index per query (-1 none)
  SolveLog lm[2];

  static int nearest(const Cloud& c, const float q[3], float* d2) {
    int best = -1;
    float bd = INFINITY;
    for (size_t i = 0; i < c.size(); ++i) {
      const float dx = q[0] - c[i].x, dy = q[1] - c[i].y, dz = q[2] - c[i].z;
      const float d = dx * dx + dy * dy + dz * dz;
      if (d < bd) { bd = d; best = (int)i; }
    }
    *d2 = bd;
    return best;
  }
  void to_start(const Pt& p, float out[3]) const {  // :108-126 with s = 1
    const double v[3] = {(double)p.x, (double)p.y, (double)p.z};
    double r[3];
    quat_rotate(para, v, r);
    for (int k = 0; k < 3; ++k) out[k] = (float)(r[k] + para[4 + k]);
  }
  static float sq(const Pt& p, const float s[3]) {  // :322-327: float expression
    return (p.x - s[0]) * (p.x - s[0]) + (p.y - s[1]) * (p.y - s[1]) + (p.z - s[2]) * (p.z - s[2]);
  }
  void step(const Cloud& sharp, const Cloud& flat, const Cloud& less_sharp, const Cloud& less_flat) {
    const double kDistSq = 25, kNearby = 2.5;  // :63-64
    n_corner[0] = n_corner[1] = n_plane[0] = n_plane[1] = 0;
    for (int o = 0; o < 2; ++o) { tr_edge[o].clear(); tr_plane[o].clear(); lm[o] = SolveLog(); }
    if (!inited) {
      inited = true;  // :267-271
    } else {
      for (int opt = 0; opt < 2; ++opt) {  // :277
        std::vector<Block> blocks;
        for (const Pt& p : sharp) {  // :300-385
          float sel[3], d2;
          to_start(p, sel);
          const int c = corner_last.empty() ? -1 : nearest(corner_last, sel, &d2);
          int closest = -1, second = -1;
          if (c >= 0 && d2 < kDistSq) {
            closest = c;
            const int id = int(corner_last[c].i);
            double best = kDistSq;
            for (int j = c + 1; j < (int)corner_last.size(); ++j) {
              if (int(corner_last[j].i) <= id) continue;
              if (int(corner_last[j].i) > id + kNearby) break;
              const double d = sq(corner_last[j], sel);
              if (d < best) { best = d; second = j; }
            }
            for (int j = c - 1; j >= 0; --j) {
              if (int(corner_last[j].i) >= id) continue;
              if (int(corner_last[j].i) < id - kNearby) break;
              const double d = sq(corner_last[j], sel);
              if (d < best) { best = d; second = j; }
            }
          }
          tr_edge[opt].push_back(closest); tr_edge[opt].push_back(second);
          if (second >= 0) {
            Block b;
            b.kind = 0;
            const Pt &a = corner_last[closest], &bb = corner_last[second];
            b.e.cp[0] = p.x; b.e.cp[1] = p.y; b.e.cp[2] = p.z;
            b.e.lpa[0] = a.x; b.e.lpa[1] = a.y; b.e.lpa[2] = a.z;
            b.e.lpb[0] = bb.x; b.e.lpb[1] = bb.y; b.e.lpb[2] = bb.z;
            b.e.s = 1.0;
            blocks.push_back(b);
            n_corner[opt]++;
          }
        }
        for (const Pt& p : flat) {  // :388-486
          float sel[3], d2;
          to_start(p, sel);
          const int c = surf_last.empty() ? -1 : nearest(surf_last, sel, &d2);
          int closest = -1, second = -1, third = -1;
          if (c >= 0 && d2 < kDistSq) {
            closest = c;
            const int id = int(surf_last[c].i);
            double best2 = kDistSq, best3 = kDistSq;
            for (int j = c + 1; j < (int)surf_last.size(); ++j) {
              if (int(surf_last[j].i) > id + kNearby) break;
              const double d = sq(surf_last[j], sel);
              if (int(surf_last[j].i) <= id && d < best2) { best2 = d; second = j; }
              else if (int(surf_last[j].i) > id && d < best3) { best3 = d; third = j; }
            }
            for (int j = c - 1; j >= 0; --j) {
              if (int(surf_last[j].i) < id - kNearby) break;
              const double d = sq(surf_last[j], sel);
              if (int(surf_last[j].i) >= id && d < best2) { best2 = d; second = j; }
              else if (int(surf_last[j].i) < id && d < best3) { best3 = d; third = j; }
            }
          }
          tr_plane[opt].push_back(closest); tr_plane[opt].push_back(second); tr_plane[opt].push_back(third);
          if (second >= 0 && third >= 0) {
            Block b;
            b.kind = 2;
            const double cp[3] = {p.x, p.y, p.z};
            const Pt &a = surf_last[closest], &l = surf_last[second], &m = surf_last[third];
            const double pj[3] = {a.x, a.y, a.z}, pl[3] = {l.x, l.y, l.z}, pm[3] = {m.x, m.y, m.z};
            b.p3.set(cp, pj, pl, pm, 1.0);
            blocks.push_back(b);
            n_plane[opt]++;
          }
        }
        solve_trust_region(blocks, para, 4, &lm[opt]);  // :495-500
      }
      // :504-505
      double r[3];
      const double qq[4] = {q_w.x, q_w.y, q_w.z, q_w.w};
      quat_rotate(qq, para + 4, r);
      for (int k = 0; k < 3; ++k) t_w[k] = t_w[k] + r[k];
      q_w = quat_mul(q_w, Quat{para[0], para[1], para[2], para[3]});
    }
    corner_last = less_sharp;  // :556-562
    surf_last = less_flat;
  }
};

}  // namespace orc

extern "C" {

typedef struct orc_stats {
  int n_corner_in, n_surf_in, n_corner_ds, n_surf_ds, n_map_corner, n_map_surf;
  int n_edge[2], n_plane[2];
  int optimized;
  int lm_iters[2];
  int lm_term[2];
  double cost_initial[2], cost_final[2];
  double cand_corner, cand_surf;
  double t_ms[8];
} orc_stats;

static_assert(sizeof(orc_stats) == sizeof(orc::Stats), "stats layout");

void* orc_create(float line_res, float plane_res) { return new orc::Mapper(line_res, plane_res); }
void orc_destroy(void* h) { delete (orc::Mapper*)h; }
void orc_set_options(void* h, int use_kdtree, int skip_optimization, int trace) {
  auto* m = (orc::Mapper*)h;
  m->use_kdtree = use_kdtree != 0;
  m->skip_opt = skip_optimization != 0;
  m->trace.on = trace != 0;
}

static orc::Cloud to_cloud(const float* xyzi, int n) {
  orc::Cloud c(n);
  if (n) std::memcpy(c.data(), xyzi, sizeof(orc::Pt) * (size_t)n);
  return c;
}

int orc_register(void* h, const float* corner, int nc, const float* surf, int ns,
                 const double q_wodom[4], const double t_wodom[3], double q_out[4],
                 double t_out[3], orc_stats* st) {
  auto* m = (orc::Mapper*)h;
  int rc = m->process(to_cloud(corner, nc), to_cloud(surf, ns), q_wodom, t_wodom, (orc::Stats*)st);
  for (int i = 0; i < 4; ++i) q_out[i] = m->par[i];
  for (int i = 0; i < 3; ++i) t_out[i] = m->par[4 + i];
  return rc;
}

void orc_get_correction(void* h, double q[4], double t[3]) {
  auto* m = (orc::Mapper*)h;
  q[0] = m->q_wmap_wodom.x; q[1] = m->q_wmap_wodom.y; q[2] = m->q_wmap_wodom.z; q[3] = m->q_wmap_wodom.w;
  for (int i = 0; i < 3; ++i) t[i] = m->t_wmap_wodom[i];
}

// Replace the map: every point is pushed into its cube in upload order, raw
// (as after laserMapping.cpp:753-759), window centred at its initial position.
int orc_map_upload(void* h, const float* corner, int nc, const float* surf, int ns) {
  auto* m = (orc::Mapper*)h;
  for (auto& c : m->corner_) c.clear();
  for (auto& c : m->surf_) c.clear();
  auto push = [&](const float* p, int n, std::vector<orc::Cloud>& arr) {
    int dropped = 0;
    for (int i = 0; i < n; ++i) {
      orc::Pt s{p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]};
      int ci = orc::Mapper::cube_of((double)s.x, m->cenW), cj = orc::Mapper::cube_of((double)s.y, m->cenH),
          ck = orc::Mapper::cube_of((double)s.z, m->cenD);
      if (ci >= 0 && ci < 21 && cj >= 0 && cj < 21 && ck >= 0 && ck < 11) arr[m->ind(ci, cj, ck)].push_back(s);
      else dropped++;
    }
    return dropped;
  };
  int d = push(corner, nc, m->corner_);
  d += push(surf, ns, m->surf_);
  return d;
}

// Local map as process() would gather it for a sensor at centre_t (rows B, C).
// cls 0 corner, 1 surf. Returns the count (may exceed cap; only cap are copied).
int orc_get_local_map(void* h, int cls, const double centre_t[3], float* out, int cap) {
  auto* m = (orc::Mapper*)h;
  double save[3] = {m->par[4], m->par[5], m->par[6]};
  for (int i = 0; i < 3; ++i) m->par[4 + i] = centre_t[i];
  int ci, cj, ck;
  m->shift_window(ci, cj, ck);
  orc::Cloud cm, sm;
  m->gather(ci, cj, ck, cm, sm);
  for (int i = 0; i < 3; ++i) m->par[4 + i] = save[i];
  const orc::Cloud& c = cls == 0 ? cm : sm;
  int n = (int)c.size();
  if (out && n) std::memcpy(out, c.data(), sizeof(orc::Pt) * (size_t)std::min(n, cap));
  return n;
}

// Whole-window map in the order of the gather loops (i, j, k nested,
// laserMapping.cpp:513-517), each cube in its stored order; for bit-exact map comparisons.
int orc_get_map(void* h, int cls, float* out, int cap) {
  auto* m = (orc::Mapper*)h;
  auto& arr = cls == 0 ? m->corner_ : m->surf_;
  int n = 0;
  for (int i = 0; i < 21; ++i)
    for (int j = 0; j < 21; ++j)
      for (int k = 0; k < 11; ++k)
        for (auto& p : arr[m->ind(i, j, k)]) {
          if (out && n < cap) std::memcpy(out + 4 * (size_t)n, &p, 16);
          ++n;
        }
  return n;
}
// /laser_cloud_surround as published every 5th frame (laserMapping.cpp:807-815): corner then
// surf cloud of every valid cube of the last process() call.
int orc_get_surround(void* h, float* out, int cap) {
  auto* m = (orc::Mapper*)h;
  int n = 0;
  for (int v = 0; v < m->n_valid_; ++v)
    for (int c = 0; c < 2; ++c)
      for (auto& p : (c == 0 ? m->corner_ : m->surf_)[m->valid_[v]]) {
        if (out && n < cap) std::memcpy(out + 4 * (size_t)n, &p, 16);
        ++n;
      }
  return n;
}
void orc_get_window(void* h, int cen[3]) {
  auto* m = (orc::Mapper*)h;
  cen[0] = m->cenW; cen[1] = m->cenH; cen[2] = m->cenD;
}

// kNN(5) of world-frame float queries against the local map gathered around centre_t.
// method 0 = canonical brute force (d2, idx); 1 = KD-tree restatement (A2/A3).
int orc_debug_knn(void* h, int cls, const double centre_t[3], const float* q_xyz, int nq,
                  int method, int32_t* idx5, float* d2_5) {
  auto* m = (orc::Mapper*)h;
  double save[3] = {m->par[4], m->par[5], m->par[6]};
  for (int i = 0; i < 3; ++i) m->par[4 + i] = centre_t[i];
  int ci, cj, ck;
  m->shift_window(ci, cj, ck);
  orc::Cloud cm, sm;
  m->gather(ci, cj, ck, cm, sm);
  for (int i = 0; i < 3; ++i) m->par[4 + i] = save[i];
  const orc::Cloud& c = cls == 0 ? cm : sm;
  orc::KdTree kd;
  if (method == 1) kd.build(c);
  for (int i = 0; i < nq; ++i) {
    int idx[5];
    float d2[5];
    int found = method == 1 ? kd.knn(q_xyz + 3 * i, 5, idx, d2) : orc::knn_brute(c, q_xyz + 3 * i, 5, idx, d2);
    for (int j = 0; j < 5; ++j) {
      idx5[5 * i + j] = j < found ? idx[j] : -1;
      d2_5[5 * i + j] = j < found ? d2[j] : INFINITY;
    }
  }
  return (int)c.size();
}

// ----- stand-alone pieces for unit tests -----
int orc_voxel_grid(const float* in, int n, float leaf, float* out) {
  orc::Cloud o;
  orc::voxel_grid(to_cloud(in, n), leaf, o);
  if (!o.empty()) std::memcpy(out, o.data(), sizeof(orc::Pt) * o.size());
  return (int)o.size();
}
int orc_knn(const float* map_xyzi, int m, const float* q_xyz, int nq, int method, int32_t* idx5,
            float* d2_5) {
  orc::Cloud c = to_cloud(map_xyzi, m);
  orc::KdTree kd;
  if (method == 1) kd.build(c);
  for (int i = 0; i < nq; ++i) {
    int idx[5];
    float d2[5];
    int found = method == 1 ? kd.knn(q_xyz + 3 * i, 5, idx, d2) : orc::knn_brute(c, q_xyz + 3 * i, 5, idx, d2);
    for (int j = 0; j < 5; ++j) {
      idx5[5 * i + j] = j < found ? idx[j] : -1;
      d2_5[5 * i + j] = j < found ? d2[j] : INFINITY;
    }
  }
  return 0;
}
int orc_eig3(const double m[9], double evals[3], double evecs[9]) {
  double e[3][3];
  bool ok = orc::eig3_selfadjoint(m, evals, e);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) evecs[3 * r + c] = e[r][c];
  return ok ? 0 : 1;
}
void orc_plane_qr(const double A[15], double n[3]) {
  const double b[5] = {-1, -1, -1, -1, -1};
  orc::colpiv_qr_solve_5x3(A, b, n);
}
// residual + 3x7 / 1x7 autodiff Jacobian of the two factors at pose x (7).
void orc_edge_factor(const double cp[3], const double a[3], const double b[3], const double x[7],
                     double r[3], double J[21]) {
  orc::EdgeFactor f;
  for (int i = 0; i < 3; ++i) { f.cp[i] = cp[i]; f.lpa[i] = a[i]; f.lpb[i] = b[i]; }
  f.s = 1.0;
  orc::Jet q[4], t[3], rj[3];
  for (int i = 0; i < 4; ++i) q[i] = orc::Jet(x[i], i);
  for (int i = 0; i < 3; ++i) t[i] = orc::Jet(x[4 + i], 4 + i);
  f(q, t, rj);
  for (int k = 0; k < 3; ++k) { r[k] = rj[k].a; for (int i = 0; i < 7; ++i) J[7 * k + i] = rj[k].v[i]; }
}
void orc_plane_factor(const double cp[3], const double n[3], double d, const double x[7], double r[1],
                      double J[7]) {
  orc::PlaneNormFactor f;
  for (int i = 0; i < 3; ++i) { f.cp[i] = cp[i]; f.n[i] = n[i]; }
  f.d = d;
  orc::Jet q[4], t[3], rj[1];
  for (int i = 0; i < 4; ++i) q[i] = orc::Jet(x[i], i);
  for (int i = 0; i < 3; ++i) t[i] = orc::Jet(x[4 + i], 4 + i);
  f(q, t, rj);
  r[0] = rj[0].a;
  for (int i = 0; i < 7; ++i) J[i] = rj[0].v[i];
}
// Solve a hand-made problem: blocks given as kind + 10 doubles
// (edge: cp[3], a[3], b[3], -; plane: cp[3], n[3], d, -, -, -).
int orc_solve(const int* kinds, const double* data10, int nb, double x[7], int max_iter,
              double* log_cost /*[1+2*max_iter]*/, int* n_iter, int* termination) {
  std::vector<orc::Block> blocks(nb);
  for (int i = 0; i < nb; ++i) {
    const double* d = data10 + 10 * i;
    blocks[i].kind = kinds[i];
    if (kinds[i] == 0) {
      for (int k = 0; k < 3; ++k) { blocks[i].e.cp[k] = d[k]; blocks[i].e.lpa[k] = d[3 + k]; blocks[i].e.lpb[k] = d[6 + k]; }
      blocks[i].e.s = 1.0;
    } else {
      for (int k = 0; k < 3; ++k) { blocks[i].p.cp[k] = d[k]; blocks[i].p.n[k] = d[3 + k]; }
      blocks[i].p.d = d[6];
    }
  }
  orc::SolveLog lg;
  orc::solve_trust_region(blocks, x, max_iter, &lg);
  if (log_cost) {
    log_cost[0] = lg.initial_cost;
    for (int i = 0; i < lg.n_iter && i < max_iter; ++i) { log_cost[1 + 2 * i] = lg.it[i].cost_change; log_cost[2 + 2 * i] = lg.it[i].radius; }
  }
  if (n_iter) *n_iter = lg.n_iter;
  if (termination) *termination = lg.termination;
  return 0;
}

// ----- trace of the last orc_register (trace option on) -----
int orc_trace_sizes(void* h, int sizes[4]) {
  auto* m = (orc::Mapper*)h;
  sizes[0] = (int)m->trace.corner_ds.size(); sizes[1] = (int)m->trace.surf_ds.size();
  sizes[2] = (int)m->trace.corner_map.size(); sizes[3] = (int)m->trace.surf_map.size();
  return 0;
}
// which: 0 corner_ds, 1 surf_ds, 2 corner_map, 3 surf_map
int orc_trace_cloud(void* h, int which, float* out) {
  auto* m = (orc::Mapper*)h;
  const orc::Cloud* c[4] = {&m->trace.corner_ds, &m->trace.surf_ds, &m->trace.corner_map, &m->trace.surf_map};
  if (!c[which]->empty()) std::memcpy(out, c[which]->data(), sizeof(orc::Pt) * c[which]->size());
  return (int)c[which]->size();
}
int orc_trace_knn(void* h, int outer, int cls, int32_t* idx5, float* d2_5, unsigned char* used) {
  auto* m = (orc::Mapper*)h;
  auto& I = m->trace.idx[outer][cls];
  auto& Dd = m->trace.d2[outer][cls];
  auto& U = m->trace.used[outer][cls];
  if (!I.empty()) { std::memcpy(idx5, I.data(), 4 * I.size()); std::memcpy(d2_5, Dd.data(), 4 * Dd.size()); }
  if (!U.empty()) std::memcpy(used, U.data(), U.size());
  return (int)U.size();
}
// per outer iteration: pose after the solve (7), the 28 initial sums, then per
// LM iteration [cost, cost_change, radius, step_norm, model_change, accepted] x 4
int orc_trace_lm(void* h, int outer, double pose7[7], double sums28[28], double iters[24],
                 int* n_iter, int* termination) {
  auto* m = (orc::Mapper*)h;
  std::memcpy(pose7, m->trace.pose_after[outer], 56);
  std::memcpy(sums28, m->trace.lm[outer].init_sums, 28 * 8);
  for (int i = 0; i < 4; ++i) {
    const orc::IterLog& il = m->trace.lm[outer].it[i];
    double* o = iters + 6 * i;
    o[0] = il.cost; o[1] = il.cost_change; o[2] = il.radius; o[3] = il.step_norm; o[4] = il.model_change; o[5] = il.accepted;
  }
  *n_iter = m->trace.lm[outer].n_iter;
  *termination = m->trace.lm[outer].termination;
  return 0;
}


// ----- scan-to-scan odometry (row N3) -----
void* orc_odom_create() { return new orc::Odometer(); }
void orc_odom_destroy(void* h) { delete (orc::Odometer*)h; }
// one sweep's four feature clouds -> q_w_curr (x,y,z,w), t_w_curr, para (q_last_curr, t_last_curr), counts[4]
void orc_odom_step(void* h, const float* sharp, int n_sharp, const float* flat, int n_flat, const float* less_sharp,
                   int n_ls, const float* less_flat, int n_lf, double q_w[4], double t_w[3], double para[7], int counts[4]) {
  orc::Odometer& O = *(orc::Odometer*)h;
  O.step(to_cloud(sharp, n_sharp), to_cloud(flat, n_flat), to_cloud(less_sharp, n_ls), to_cloud(less_flat, n_lf));
  q_w[0] = O.q_w.x; q_w[1] = O.q_w.y; q_w[2] = O.q_w.z; q_w[3] = O.q_w.w;
  for (int k = 0; k < 3; ++k) t_w[k] = O.t_w[k];
  for (int k = 0; k < 7; ++k) para[k] = O.para[k];
  counts[0] = O.n_corner[0]; counts[1] = O.n_corner[1]; counts[2] = O.n_plane[0]; counts[3] = O.n_plane[1];
}
// correspondences of optimisation pass `opt` of the last step: edge[2*n_sharp], plane[3*n_flat]
void orc_odom_trace(void* h, int opt, int32_t* edge, int32_t* plane) {
  orc::Odometer& O = *(orc::Odometer*)h;
  std::memcpy(edge, O.tr_edge[opt].data(), O.tr_edge[opt].size() * sizeof(int32_t));
  std::memcpy(plane, O.tr_plane[opt].data(), O.tr_plane[opt].size() * sizeof(int32_t));
}
}  // extern "C"
