// harness/synth.cpp -- seeded synthetic inputs for the scan-to-map path.
//
// This is INPUT GENERATION, not the product path and not the oracle.  It makes
// what the reference's laserMapping node receives on its four topics
// (/root/reference/src/laserMapping.cpp:921-927):
//   * a raw LiDAR sweep of a procedural street-grid world (SURVEY.md section 8d:
//     grid yawed 17 deg, undulating ground, facades with recesses, poles, cars),
//   * the upstream feature split of that sweep -- a restatement of
//     /root/reference/src/scanRegistration.cpp:142-420 (ring assignment,
//     11-tap curvature, per-sextant sharp/flat picking, 0.2 m per-ring voxel
//     filter) -- giving the "less sharp" corner cloud and "less flat" surf cloud
//     that laserOdometry republishes unchanged as /laser_cloud_corner_last and
//     /laser_cloud_surf_last (/root/reference/src/laserOdometry.cpp:570-591),
//   * a drifting odometry pose chain standing in for /laser_odom_to_init.
//
// Build: g++ -O2 -fopenmp -shared -fPIC harness/synth.cpp -o harness/libs2m_harness.so
// Plain C ABI so tests/bench load it with ctypes.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------- hashing RNG
inline uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t hash4(uint64_t seed, int64_t a, int64_t b, int64_t c) {
  uint64_t h = mix64(seed);
  h = mix64(h ^ (uint64_t)a * 0xD6E8FEB86659FD93ull);
  h = mix64(h ^ (uint64_t)b * 0xA5A5A5A5A5A5A5A5ull);
  h = mix64(h ^ (uint64_t)c * 0xC2B2AE3D27D4EB4Full);
  return h;
}
inline double u01(uint64_t h) { return (double)(h >> 11) * (1.0 / 9007199254740992.0); }
inline double gauss(uint64_t h) {
  double u1 = u01(h), u2 = u01(mix64(h));
  if (u1 < 1e-300) u1 = 1e-300;
  return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * kPi * u2);
}

// ---------------------------------------------------------------- small math
struct V3 { double x, y, z; };
struct Q4 { double x, y, z, w; };
inline Q4 qmul(const Q4& a, const Q4& b) {
  return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
          a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w,
          a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Q4 qconj(const Q4& q) { return {-q.x, -q.y, -q.z, q.w}; }
inline Q4 qnorm(const Q4& q) {
  double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
  return {q.x / n, q.y / n, q.z / n, q.w / n};
}
inline V3 qrot(const Q4& q, const V3& v) {
  V3 u{q.x, q.y, q.z};
  V3 uv{u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x};
  uv = {uv.x + uv.x, uv.y + uv.y, uv.z + uv.z};
  V3 c{u.y * uv.z - u.z * uv.y, u.z * uv.x - u.x * uv.z, u.x * uv.y - u.y * uv.x};
  return {v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z};
}
inline Q4 q_from_rpy(double roll, double pitch, double yaw) {
  Q4 qz{0, 0, std::sin(yaw / 2), std::cos(yaw / 2)};
  Q4 qy{0, std::sin(pitch / 2), 0, std::cos(pitch / 2)};
  Q4 qx{std::sin(roll / 2), 0, 0, std::cos(roll / 2)};
  return qmul(qmul(qz, qy), qx);
}
inline Q4 q_from_rotvec(double rx, double ry, double rz) {
  double a = std::sqrt(rx * rx + ry * ry + rz * rz);
  if (a < 1e-15) return {0, 0, 0, 1};
  double s = std::sin(a / 2) / a;
  return {rx * s, ry * s, rz * s, std::cos(a / 2)};
}

// ---------------------------------------------------------------- the world
// Street frame = world frame rotated by -kGridYaw about z (grid is
// yawed off the axes so voxel lattices do not align with the facades).
constexpr double kGridYaw = 17.0 * kPi / 180.0;
constexpr double kPitch = 80.0;        // street centre-line spacing (m)
constexpr double kSensorH = 1.73;      // sensor height above ground
constexpr double kGroundAmp = 0.3;     // +-0.3 m undulation
constexpr double kGroundWave = 40.0;   // ... over 40 m

struct World {
  uint64_t seed;
  double cy, sy;
  explicit World(uint64_t s) : seed(s), cy(std::cos(kGridYaw)), sy(std::sin(kGridYaw)) {}
  void to_street(double x, double y, double& xs, double& ys) const {
    xs = cy * x + sy * y;
    ys = -sy * x + cy * y;
  }
  void to_world(double xs, double ys, double& x, double& y) const {
    x = cy * xs - sy * ys;
    y = sy * xs + cy * ys;
  }
  double ground_s(double xs, double ys) const {
    return kGroundAmp * std::sin(2 * kPi * xs / kGroundWave) * std::cos(2 * kPi * ys / kGroundWave);
  }
  double ground(double x, double y) const {
    double xs, ys;
    to_street(x, y, xs, ys);
    return ground_s(xs, ys);
  }
};

// A vertical primitive in the street frame.
struct Prim {
  int kind;              // 0 = wall segment, 1 = pole (circle), 2 = horizontal rect (car roof)
  double ax, ay, bx, by; // segment end points | circle centre (ax,ay), radius bx | rect min (ax,ay) max (bx,by)
  double zlo, zhi;       // vertical extent (zhi = roof height for kind 2)
};

void add_wall_with_recesses(std::vector<Prim>& out, double x0, double y0, double x1, double y1,
                            double nx, double ny, double zhi) {
  // Wall from (x0,y0) to (x1,y1); (nx,ny) points INTO the building. Every 12 m a
  // 3 m wide, 0.5 m deep recess.
  double L = std::hypot(x1 - x0, y1 - y0);
  double tx = (x1 - x0) / L, ty = (y1 - y0) / L;
  double s = 0.0;
  auto P = [&](double along, double depth, double& px, double& py) {
    px = x0 + tx * along + nx * depth;
    py = y0 + ty * along + ny * depth;
  };
  auto seg = [&](double s0, double d0, double s1, double d1) {
    Prim p;
    p.kind = 0;
    P(s0, d0, p.ax, p.ay);
    P(s1, d1, p.bx, p.by);
    p.zlo = -5.0;
    p.zhi = zhi;
    out.push_back(p);
  };
  for (double r0 = 6.0; r0 + 3.0 < L - 3.0; r0 += 12.0) {
    seg(s, 0.0, r0, 0.0);
    seg(r0, 0.0, r0, 0.5);
    seg(r0, 0.5, r0 + 3.0, 0.5);
    seg(r0 + 3.0, 0.5, r0 + 3.0, 0.0);
    s = r0 + 3.0;
  }
  seg(s, 0.0, L, 0.0);
}

void collect_prims(const World& w, double sx, double sy, double range, std::vector<Prim>& out) {
  out.clear();
  int ca = (int)std::floor(sx / kPitch), cb = (int)std::floor(sy / kPitch);
  int R = (int)std::ceil(range / kPitch) + 1;
  for (int a = ca - R; a <= ca + R; ++a)
    for (int b = cb - R; b <= cb + R; ++b) {
      // building block inside cell (a,b); set-backs 10..15 m from the street centre lines
      double wW = 10.0 + 5.0 * u01(hash4(w.seed, a, b, 1));
      double wE = 10.0 + 5.0 * u01(hash4(w.seed, a, b, 2));
      double wS = 10.0 + 5.0 * u01(hash4(w.seed, a, b, 3));
      double wN = 10.0 + 5.0 * u01(hash4(w.seed, a, b, 4));
      double H = 8.0 + 17.0 * u01(hash4(w.seed, a, b, 5));
      double x0 = a * kPitch + wW, x1 = (a + 1) * kPitch - wE;
      double y0 = b * kPitch + wS, y1 = (b + 1) * kPitch - wN;
      // cull whole building by distance to its rectangle
      double dx = std::max({x0 - sx, 0.0, sx - x1}), dy = std::max({y0 - sy, 0.0, sy - y1});
      if (dx * dx + dy * dy > range * range) continue;
      add_wall_with_recesses(out, x0, y0, x1, y0, 0, 1, H);   // south face
      add_wall_with_recesses(out, x1, y0, x1, y1, -1, 0, H);  // east
      add_wall_with_recesses(out, x1, y1, x0, y1, 0, -1, H);  // north
      add_wall_with_recesses(out, x0, y1, x0, y0, 1, 0, H);   // west
    }
  // street furniture along both street families
  int j0 = (int)std::floor((sx - range) / 10.0), j1 = (int)std::ceil((sx + range) / 10.0);
  int k0 = (int)std::floor((sy - range) / 10.0), k1 = (int)std::ceil((sy + range) / 10.0);
  auto near_crossing = [&](double v) {
    double r = v - kPitch * std::floor(v / kPitch + 0.5);
    return std::fabs(r) < 16.0;
  };
  auto add_pole = [&](double px, double py) {
    if ((px - sx) * (px - sx) + (py - sy) * (py - sy) > range * range) return;
    Prim p;
    p.kind = 1;
    p.ax = px; p.ay = py; p.bx = 0.15; p.by = 0;
    p.zlo = -5.0;
    p.zhi = w.ground_s(px, py) + 6.0;
    out.push_back(p);
  };
  auto add_car = [&](double cx, double cy, bool along_x) {
    if ((cx - sx) * (cx - sx) + (cy - sy) * (cy - sy) > range * range) return;
    double hx = along_x ? 2.1 : 0.9, hy = along_x ? 0.9 : 2.1;
    double top = w.ground_s(cx, cy) + 1.5;
    double X0 = cx - hx, X1 = cx + hx, Y0 = cy - hy, Y1 = cy + hy;
    Prim s;
    s.kind = 0; s.zlo = -5.0; s.zhi = top;
    s.ax = X0; s.ay = Y0; s.bx = X1; s.by = Y0; out.push_back(s);
    s.ax = X1; s.ay = Y0; s.bx = X1; s.by = Y1; out.push_back(s);
    s.ax = X1; s.ay = Y1; s.bx = X0; s.by = Y1; out.push_back(s);
    s.ax = X0; s.ay = Y1; s.bx = X0; s.by = Y0; out.push_back(s);
    Prim r;
    r.kind = 2; r.ax = X0; r.ay = Y0; r.bx = X1; r.by = Y1; r.zlo = top; r.zhi = top;
    out.push_back(r);
  };
  int b0 = (int)std::floor((sy - range) / kPitch), b1 = (int)std::ceil((sy + range) / kPitch);
  for (int b = b0; b <= b1; ++b)
    for (int j = j0; j <= j1; ++j) {
      double x = 10.0 * j + 5.0;
      if (near_crossing(x)) continue;
      for (int side = -1; side <= 1; side += 2) {
        add_pole(x, b * kPitch + side * 6.5);
        if (u01(hash4(w.seed, j, b * 2 + (side > 0), 11)) < 0.35)
          add_car(x - 5.0, b * kPitch + side * 4.0, true);
      }
    }
  int a0 = (int)std::floor((sx - range) / kPitch), a1 = (int)std::ceil((sx + range) / kPitch);
  for (int a = a0; a <= a1; ++a)
    for (int k = k0; k <= k1; ++k) {
      double y = 10.0 * k + 5.0;
      if (near_crossing(y)) continue;
      for (int side = -1; side <= 1; side += 2) {
        add_pole(a * kPitch + side * 6.5, y);
        if (u01(hash4(w.seed, k, a * 2 + (side > 0), 12)) < 0.35)
          add_car(a * kPitch + side * 4.0, y - 5.0, false);
      }
    }
}

// ---------------------------------------------------------------- sensors
struct Sensor {
  int n_beams, n_az;
  double elev_deg[128];
  double min_range, max_range, dropout;
};
Sensor make_sensor(int kind) {
  Sensor s{};
  if (kind == 0) {  // HDL-64: rule of scanRegistration.cpp:190-203
    s.n_beams = 64; s.n_az = 1900;
    for (int i = 0; i < 32; ++i) s.elev_deg[i] = 2.0 - i / 3.0;
    for (int i = 0; i < 32; ++i) s.elev_deg[32 + i] = -8.83 - i / 2.0;
    s.min_range = 5.0; s.max_range = 120.0; s.dropout = 0.0;
  } else if (kind == 1) {  // VLP-16: scanRegistration.cpp:171-179
    s.n_beams = 16; s.n_az = 1800;
    for (int i = 0; i < 16; ++i) s.elev_deg[i] = -15.0 + 2.0 * i;
    s.min_range = 0.1; s.max_range = 100.0; s.dropout = 0.0;
  } else {  // OS1-64 (MulRan shape): +-16.6 deg, 1024 columns, ~44 % no-returns
    s.n_beams = 64; s.n_az = 1024;
    for (int i = 0; i < 64; ++i) s.elev_deg[i] = -16.6 + 33.2 * i / 63.0;
    s.min_range = 0.5; s.max_range = 120.0; s.dropout = 0.30;
  }
  return s;
}

constexpr int kBins = 2048;

}  // namespace

#ifdef _OPENMP
#include <omp.h>
#endif
extern "C" {

// OpenMP threads of the ray caster (a launcher such as torchrun exports OMP_NUM_THREADS=1 to every rank)
void synth_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// Closed-block trajectory at 1 m/frame (SURVEY 8d "KITTI-05 shape"): straight
// runs along street centre lines and 90-degree left turns at 3 deg/frame.
// poses: 7 doubles per frame [qx qy qz qw tx ty tz], sensor->world.
int synth_trajectory(uint64_t seed, int n_frames, double step_m, double* poses) {
  World w(seed);
  const double yaw_rate = 3.0 * kPi / 180.0 * step_m;  // per frame
  const int turn_frames = (int)std::lround((kPi / 2) / yaw_rate);
  const double radius = step_m / yaw_rate;
  const int straight_frames = (int)std::lround((kPitch - 2 * radius) / step_m);
  // start on the centre line y'=0 heading +x', just after a crossing's arc
  double xs = radius, ys = 0.0, heading = 0.0;
  int phase = 0, left = straight_frames;
  for (int f = 0; f < n_frames; ++f) {
    double x, y;
    w.to_world(xs, ys, x, y);
    double z = w.ground(x, y) + kSensorH;
    // terrain-following attitude plus a slow wobble so all 6 dof are exercised
    double e = 0.5;
    double hx = std::cos(heading + kGridYaw), hy = std::sin(heading + kGridYaw);
    double gf = (w.ground(x + e * hx, y + e * hy) - w.ground(x - e * hx, y - e * hy)) / (2 * e);
    double gl = (w.ground(x - e * hy, y + e * hx) - w.ground(x + e * hy, y - e * hx)) / (2 * e);
    double pitch = -std::atan(gf) + 0.004 * std::sin(0.21 * f);
    double roll = std::atan(gl) + 0.004 * std::cos(0.17 * f);
    Q4 q = q_from_rpy(roll, pitch, heading + kGridYaw);
    double* p = poses + 7 * f;
    p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; p[4] = x; p[5] = y; p[6] = z;
    // advance
    double dyaw = (phase == 1) ? yaw_rate : 0.0;
    double hmid = heading + 0.5 * dyaw;
    xs += step_m * std::cos(hmid);
    ys += step_m * std::sin(hmid);
    heading += dyaw;
    if (--left == 0) {
      phase ^= 1;
      left = phase ? turn_frames : straight_frames;
    }
  }
  return 0;
}

// Odometry chain: true relative motion corrupted by a small per-frame error
// (random walk), standing in for laserOdometry's /laser_odom_to_init.
int synth_odometry(uint64_t seed, int n_frames, const double* true_poses, double sigma_t,
                   double sigma_r_rad, double* odom_poses) {
  Q4 qo{0, 0, 0, 1};
  V3 to{0, 0, 0};
  for (int f = 0; f < n_frames; ++f) {
    const double* p = true_poses + 7 * f;
    if (f == 0) {
      qo = {p[0], p[1], p[2], p[3]};
      to = {p[4], p[5], p[6]};
    } else {
      const double* pp = true_poses + 7 * (f - 1);
      Q4 q0{pp[0], pp[1], pp[2], pp[3]}, q1{p[0], p[1], p[2], p[3]};
      Q4 dq = qmul(qconj(q0), q1);
      V3 dt = qrot(qconj(q0), V3{p[4] - pp[4], p[5] - pp[5], p[6] - pp[6]});
      Q4 nq = q_from_rotvec(sigma_r_rad * gauss(hash4(seed, f, 1, 77)),
                            sigma_r_rad * gauss(hash4(seed, f, 2, 77)),
                            sigma_r_rad * gauss(hash4(seed, f, 3, 77)));
      dq = qnorm(qmul(dq, nq));
      dt.x += sigma_t * gauss(hash4(seed, f, 4, 77));
      dt.y += sigma_t * gauss(hash4(seed, f, 5, 77));
      dt.z += sigma_t * gauss(hash4(seed, f, 6, 77));
      V3 r = qrot(qo, dt);
      to = {to.x + r.x, to.y + r.y, to.z + r.z};
      qo = qnorm(qmul(qo, dq));
    }
    double* o = odom_poses + 7 * f;
    o[0] = qo.x; o[1] = qo.y; o[2] = qo.z; o[3] = qo.w; o[4] = to.x; o[5] = to.y; o[6] = to.z;
  }
  return 0;
}

// One sweep. sensor: 0 HDL-64 (64x1900), 1 VLP-16 (16x1800), 2 OS1-64 (64x1024).
// Output: xyz floats in the SENSOR frame, beam-major, each beam one clockwise
// sweep (the order scanRegistration's curvature taps assume). Returns n points.
int synth_scan(uint64_t seed, int sensor_kind, const double* pose7, int frame, double range_sigma,
               float* xyz_out, int cap_points) {
  World w(seed);
  Sensor S = make_sensor(sensor_kind);
  Q4 q{pose7[0], pose7[1], pose7[2], pose7[3]};
  V3 o{pose7[4], pose7[5], pose7[6]};
  double osx, osy;
  w.to_street(o.x, o.y, osx, osy);
  // orientation expressed in the street frame
  Q4 qs = qmul(Q4{0, 0, std::sin(-kGridYaw / 2), std::cos(-kGridYaw / 2)}, q);

  std::vector<Prim> prims;
  collect_prims(w, osx, osy, S.max_range + 5.0, prims);
  std::vector<std::vector<int>> bins(kBins);
  auto bin_of = [](double ang) {
    int b = (int)std::floor((ang + kPi) / (2 * kPi) * kBins);
    return std::min(std::max(b, 0), kBins - 1);
  };
  for (int i = 0; i < (int)prims.size(); ++i) {
    const Prim& p = prims[i];
    double lo, hi;
    if (p.kind == 1) {
      double d = std::hypot(p.ax - osx, p.ay - osy);
      if (d <= p.bx + 1e-6) continue;
      double c = std::atan2(p.ay - osy, p.ax - osx), h = std::asin(p.bx / d) + 1e-3;
      lo = c - h; hi = c + h;
    } else {
      double xs[4], ys[4];
      int n = 2;
      xs[0] = p.ax; ys[0] = p.ay; xs[1] = p.bx; ys[1] = p.by;
      if (p.kind == 2) { n = 4; xs[2] = p.ax; ys[2] = p.by; xs[3] = p.bx; ys[3] = p.ay; }
      double c = std::atan2(ys[0] - osy, xs[0] - osx);
      lo = hi = 0;
      for (int k = 1; k < n; ++k) {
        double a = std::atan2(ys[k] - osy, xs[k] - osx) - c;
        while (a > kPi) a -= 2 * kPi;
        while (a < -kPi) a += 2 * kPi;
        lo = std::min(lo, a); hi = std::max(hi, a);
      }
      lo += c - 1e-3; hi += c + 1e-3;
    }
    int nb = (int)std::ceil((hi - lo) / (2 * kPi) * kBins) + 1;
    double a = lo;
    for (int k = 0; k <= nb; ++k, a += 2 * kPi / kBins) {
      double aa = std::min(a, hi);
      while (aa >= kPi) aa -= 2 * kPi;
      while (aa < -kPi) aa += 2 * kPi;
      int b = bin_of(aa);
      if (bins[b].empty() || bins[b].back() != i) bins[b].push_back(i);
    }
  }

  const int n_rays = S.n_beams * S.n_az;
  std::vector<float> pts((size_t)n_rays * 3);
  std::vector<unsigned char> ok((size_t)n_rays, 0);
#pragma omp parallel for schedule(static, 256)
  for (int ray = 0; ray < n_rays; ++ray) {
    int beam = ray / S.n_az, col = ray % S.n_az;
    double el = S.elev_deg[beam] * kPi / 180.0;
    double az = -2.0 * kPi * (col + 0.5) / S.n_az;  // clockwise
    V3 ds{std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
    V3 d = qrot(qs, ds);  // street-frame direction
    double hn = std::hypot(d.x, d.y);
    if (hn < 1e-6) continue;
    double ux = d.x / hn, uy = d.y / hn, m = d.z / hn;
    double rho_max = S.max_range * hn, best = 1e30;
    // ground: march between the heights the terrain can take, then bisect
    if (m < -1e-9) {
      double lo = std::max(0.0, (o.z - kGroundAmp) / (-m) - 1.0);
      double hi = std::min(rho_max, (o.z + kGroundAmp) / (-m) + 1.0);
      double prev = lo;
      auto f = [&](double r) { return o.z + m * r - w.ground_s(osx + ux * r, osy + uy * r); };
      if (f(lo) > 0) {
        for (double r = lo + 1.0;; r += 1.0) {
          double rr = std::min(r, hi);
          if (f(rr) <= 0) {
            double a = prev, b = rr;
            for (int it = 0; it < 40; ++it) {
              double c = 0.5 * (a + b);
              (f(c) > 0 ? a : b) = c;
            }
            best = 0.5 * (a + b);
            break;
          }
          prev = rr;
          if (rr >= hi) break;
        }
      }
    }
    int b = bin_of(std::atan2(uy, ux));
    for (int pi : bins[b]) {
      const Prim& p = prims[pi];
      double rho = -1;
      if (p.kind == 0) {
        double ex = p.bx - p.ax, ey = p.by - p.ay;
        double den = ux * ey - uy * ex;
        if (std::fabs(den) < 1e-12) continue;
        double wx = p.ax - osx, wy = p.ay - osy;
        double t = (wx * ey - wy * ex) / den;
        double s = (wx * uy - wy * ux) / den;
        if (t <= 0 || s < 0 || s > 1) continue;
        rho = t;
      } else if (p.kind == 1) {
        double wx = p.ax - osx, wy = p.ay - osy;
        double tc = wx * ux + wy * uy;
        double d2 = wx * wx + wy * wy - tc * tc;
        double r2 = p.bx * p.bx;
        if (d2 > r2 || tc <= 0) continue;
        rho = tc - std::sqrt(r2 - d2);
        if (rho <= 0) continue;
      } else {
        if (m > -1e-9) continue;
        double r = (p.zhi - o.z) / m;
        if (r <= 0) continue;
        double hx = osx + ux * r, hy = osy + uy * r;
        if (hx < p.ax || hx > p.bx || hy < p.ay || hy > p.by) continue;
        rho = r;
      }
      if (rho >= best) continue;
      if (p.kind != 2) {
        double z = o.z + m * rho;
        if (z < p.zlo || z > p.zhi) continue;
      }
      best = rho;
    }
    if (best > rho_max) continue;
    double r3 = best / hn;
    uint64_t h = hash4(seed ^ 0x5ca1ab1eull, frame, ray, 31);
    if (S.dropout > 0 && u01(mix64(h ^ 0x77)) < S.dropout) continue;
    r3 += range_sigma * gauss(h);
    if (r3 < 0.3) continue;
    pts[(size_t)ray * 3 + 0] = (float)(ds.x * r3);
    pts[(size_t)ray * 3 + 1] = (float)(ds.y * r3);
    pts[(size_t)ray * 3 + 2] = (float)(ds.z * r3);
    ok[ray] = 1;
  }
  int n = 0;
  for (int ray = 0; ray < n_rays; ++ray)
    if (ok[ray]) {
      if (n >= cap_points) return -1;
      std::memcpy(xyz_out + (size_t)n * 3, &pts[(size_t)ray * 3], 12);
      ++n;
    }
  return n;
}

}  // extern "C"

// ------------------------------------------------------------------------
// Upstream feature split: restatement of scanRegistration.cpp:142-420.
namespace {

struct P4 { float x, y, z, i; };

// PCL-1.8 VoxelGrid semantics with a stable sort; used here
// only for the 0.2 m per-ring "less flat" thinning (scanRegistration.cpp:414-420).
void voxel_thin(const std::vector<P4>& in, float leaf, std::vector<P4>& out) {
  out.clear();
  if (in.empty()) return;
  float inv = 1.0f / leaf;
  float mn[3] = {in[0].x, in[0].y, in[0].z}, mx[3] = {in[0].x, in[0].y, in[0].z};
  for (const P4& p : in) {
    mn[0] = std::min(mn[0], p.x); mx[0] = std::max(mx[0], p.x);
    mn[1] = std::min(mn[1], p.y); mx[1] = std::max(mx[1], p.y);
    mn[2] = std::min(mn[2], p.z); mx[2] = std::max(mx[2], p.z);
  }
  int minb[3], maxb[3];
  for (int k = 0; k < 3; ++k) {
    minb[k] = (int)std::floor(mn[k] * inv);
    maxb[k] = (int)std::floor(mx[k] * inv);
  }
  int64_t d0 = maxb[0] - minb[0] + 1, d1 = maxb[1] - minb[1] + 1;
  std::vector<std::pair<int64_t, int>> keys(in.size());
  for (size_t n = 0; n < in.size(); ++n) {
    int i0 = (int)(std::floor(in[n].x * inv) - (float)minb[0]);
    int i1 = (int)(std::floor(in[n].y * inv) - (float)minb[1]);
    int i2 = (int)(std::floor(in[n].z * inv) - (float)minb[2]);
    keys[n] = {i0 + i1 * d0 + i2 * d0 * d1, (int)n};
  }
  std::stable_sort(keys.begin(), keys.end(),
                   [](const auto& a, const auto& b) { return a.first < b.first; });
  size_t a = 0;
  while (a < keys.size()) {
    size_t b = a;
    float sx = 0, sy = 0, sz = 0, si = 0;
    while (b < keys.size() && keys[b].first == keys[a].first) {
      const P4& p = in[keys[b].second];
      sx += p.x; sy += p.y; sz += p.z; si += p.i;
      ++b;
    }
    float c = (float)(b - a);
    out.push_back({sx / c, sy / c, sz / c, si / c});
    a = b;
  }
}

}  // namespace

extern "C" {

// raw sweep (xyz, sensor frame) -> less-sharp corners, less-flat surfs, full
// ring-ordered cloud. Each output is xyzi floats (intensity = scanID + 0.1*relTime).
// Returns 0, or -1 when an output capacity is too small.
int synth_features(int sensor_kind, double minimum_range, const float* xyz, int n_in,
                   float* corner, int cap_corner, int* n_corner, float* surf, int cap_surf,
                   int* n_surf, float* full, int cap_full, int* n_full) {
  const int N_SCANS = (sensor_kind == 1) ? 16 : 64;
  const double scanPeriod = 0.1;
  std::vector<P4> in;
  in.reserve(n_in);
  const float thres = (float)minimum_range;
  for (int i = 0; i < n_in; ++i) {
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
    if (x * x + y * y + z * z < thres * thres) continue;
    in.push_back({x, y, z, 0.f});
  }
  int cloudSize = (int)in.size();
  *n_corner = *n_surf = *n_full = 0;
  if (cloudSize < 12) return 0;
  float startOri = -std::atan2(in[0].y, in[0].x);
  float endOri = -std::atan2(in[cloudSize - 1].y, in[cloudSize - 1].x) + 2 * (float)kPi;
  if (endOri - startOri > 3 * kPi) endOri -= 2 * kPi;
  else if (endOri - startOri < kPi) endOri += 2 * kPi;

  bool halfPassed = false;
  std::vector<std::vector<P4>> rings(N_SCANS);
  int count = cloudSize;
  for (int i = 0; i < cloudSize; ++i) {
    P4 p = in[i];
    float angle = std::atan(p.z / std::sqrt(p.x * p.x + p.y * p.y)) * 180 / kPi;
    int scanID = 0;
    if (sensor_kind == 1) {
      scanID = int((angle + 15) / 2 + 0.5);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    } else if (sensor_kind == 0) {
      if (angle >= -8.83) scanID = int((2 - angle) * 3.0 + 0.5);
      else scanID = N_SCANS / 2 + int((-8.83 - angle) * 2.0 + 0.5);
      if (angle > 2 || angle < -24.33 || scanID > 50 || scanID < 0) { count--; continue; }
    } else {
      scanID = int((angle + 22.5) / 2 + 0.5);
      if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
    }
    float ori = -std::atan2(p.y, p.x);
    if (!halfPassed) {
      if (ori < startOri - kPi / 2) ori += 2 * kPi;
      else if (ori > startOri + kPi * 3 / 2) ori -= 2 * kPi;
      if (ori - startOri > kPi) halfPassed = true;
    } else {
      ori += 2 * kPi;
      if (ori < endOri - kPi * 3 / 2) ori += 2 * kPi;
      else if (ori > endOri + kPi / 2) ori -= 2 * kPi;
    }
    float relTime = (ori - startOri) / (endOri - startOri);
    p.i = scanID + scanPeriod * relTime;
    rings[scanID].push_back(p);
  }
  cloudSize = count;
  std::vector<P4> cloud;
  cloud.reserve(cloudSize);
  std::vector<int> scanStart(N_SCANS), scanEnd(N_SCANS);
  for (int i = 0; i < N_SCANS; ++i) {
    scanStart[i] = (int)cloud.size() + 5;
    cloud.insert(cloud.end(), rings[i].begin(), rings[i].end());
    scanEnd[i] = (int)cloud.size() - 6;
  }
  std::vector<float> curv(cloudSize, 0.f);
  std::vector<int> sortInd(cloudSize), picked(cloudSize, 0), label(cloudSize, 0);
  for (int i = 5; i < cloudSize - 5; ++i) {
    float dX = cloud[i - 5].x + cloud[i - 4].x + cloud[i - 3].x + cloud[i - 2].x + cloud[i - 1].x -
               10 * cloud[i].x + cloud[i + 1].x + cloud[i + 2].x + cloud[i + 3].x + cloud[i + 4].x +
               cloud[i + 5].x;
    float dY = cloud[i - 5].y + cloud[i - 4].y + cloud[i - 3].y + cloud[i - 2].y + cloud[i - 1].y -
               10 * cloud[i].y + cloud[i + 1].y + cloud[i + 2].y + cloud[i + 3].y + cloud[i + 4].y +
               cloud[i + 5].y;
    float dZ = cloud[i - 5].z + cloud[i - 4].z + cloud[i - 3].z + cloud[i - 2].z + cloud[i - 1].z -
               10 * cloud[i].z + cloud[i + 1].z + cloud[i + 2].z + cloud[i + 3].z + cloud[i + 4].z +
               cloud[i + 5].z;
    curv[i] = dX * dX + dY * dY + dZ * dZ;
    sortInd[i] = i;
  }
  auto gap = [&](int a, int b) {
    float dx = cloud[a].x - cloud[b].x, dy = cloud[a].y - cloud[b].y, dz = cloud[a].z - cloud[b].z;
    return dx * dx + dy * dy + dz * dz;
  };
  auto mark_neighbours = [&](int ind) {
    for (int l = 1; l <= 5; ++l) {
      if (gap(ind + l, ind + l - 1) > 0.05) break;
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; --l) {
      if (gap(ind + l, ind + l + 1) > 0.05) break;
      picked[ind + l] = 1;
    }
  };
  std::vector<P4> cornerLess, surfLess, ringLess, ringDS;
  for (int i = 0; i < N_SCANS; ++i) {
    if (scanEnd[i] - scanStart[i] < 6) continue;
    ringLess.clear();
    for (int j = 0; j < 6; ++j) {
      int sp = scanStart[i] + (scanEnd[i] - scanStart[i]) * j / 6;
      int ep = scanStart[i] + (scanEnd[i] - scanStart[i]) * (j + 1) / 6 - 1;
      std::sort(sortInd.begin() + sp, sortInd.begin() + ep + 1,
                [&](int a, int b) { return curv[a] < curv[b]; });
      int largest = 0;
      for (int k = ep; k >= sp; --k) {
        int ind = sortInd[k];
        if (picked[ind] == 0 && curv[ind] > 0.1) {
          largest++;
          if (largest <= 2) { label[ind] = 2; cornerLess.push_back(cloud[ind]); }
          else if (largest <= 20) { label[ind] = 1; cornerLess.push_back(cloud[ind]); }
          else break;
          picked[ind] = 1;
          mark_neighbours(ind);
        }
      }
      int smallest = 0;
      for (int k = sp; k <= ep; ++k) {
        int ind = sortInd[k];
        if (picked[ind] == 0 && curv[ind] < 0.1) {
          label[ind] = -1;
          smallest++;
          if (smallest >= 4) break;
          picked[ind] = 1;
          mark_neighbours(ind);
        }
      }
      for (int k = sp; k <= ep; ++k)
        if (label[k] <= 0) ringLess.push_back(cloud[k]);
    }
    voxel_thin(ringLess, 0.2f, ringDS);
    surfLess.insert(surfLess.end(), ringDS.begin(), ringDS.end());
  }
  if ((int)cornerLess.size() > cap_corner || (int)surfLess.size() > cap_surf ||
      (full && (int)cloud.size() > cap_full))
    return -1;
  std::memcpy(corner, cornerLess.data(), cornerLess.size() * sizeof(P4));
  std::memcpy(surf, surfLess.data(), surfLess.size() * sizeof(P4));
  if (full) std::memcpy(full, cloud.data(), cloud.size() * sizeof(P4));
  *n_corner = (int)cornerLess.size();
  *n_surf = (int)surfLess.size();
  *n_full = full ? (int)cloud.size() : 0;
  return 0;
}

// Every world surface inside the square |x|, |y| <= half (world frame) sampled directly -- the "synthesise
// directly" variant of SURVEY 8d config 3 (a saturated 21 x 21 x 11 cube window without driving it):
//   surf   : the ground on a jittered lattice of `surf_step`, facades / car sides on a lattice of the same step
//   corner : poles and the vertical edges of facades (building corners, recess edges) every `corner_step` in z
// Points are xyzi (intensity 0), world frame, with `sigma` metres of Gaussian noise.  Returns -1 when a
// buffer is too small; counts in n_corner / n_surf.
int synth_surfaces(uint64_t seed, double half, double surf_step, double corner_step, double sigma, float* corner,
                   int cap_corner, int* n_corner, float* surf, int cap_surf, int* n_surf) {
  World w(seed);
  std::vector<float> C, S;
  uint64_t ctr = 0;
  auto emit = [&](std::vector<float>& out, double xs, double ys, double z) {
    double x, y;
    w.to_world(xs, ys, x, y);
    if (std::fabs(x) > half || std::fabs(y) > half) return;
    const uint64_t h = hash4(seed ^ 0x5AFE, (int64_t)ctr, 7, 9);
    ++ctr;
    out.push_back((float)(x + sigma * gauss(h)));
    out.push_back((float)(y + sigma * gauss(mix64(h ^ 1))));
    out.push_back((float)(z + sigma * gauss(mix64(h ^ 2))));
    out.push_back(0.0f);
  };
  const double R = half * 1.5;  // the street frame is yawed: cover the rotated square
  // ground
  const int ng = (int)std::ceil(2 * R / surf_step);
  for (int i = 0; i < ng; ++i)
    for (int j = 0; j < ng; ++j) {
      const uint64_t h = hash4(seed ^ 0x6E0D, i, j, 3);
      const double xs = -R + (i + u01(h)) * surf_step, ys = -R + (j + u01(mix64(h))) * surf_step;
      emit(S, xs, ys, w.ground_s(xs, ys));
    }
  // vertical primitives
  std::vector<Prim> prims;
  collect_prims(w, 0.0, 0.0, R * 1.5, prims);
  for (const Prim& p : prims) {
    if (p.kind == 1) {  // pole: a vertical line of corner points
      const double g = w.ground_s(p.ax, p.ay);
      for (double z = g + 0.1; z < p.zhi; z += corner_step) emit(C, p.ax + p.bx, p.ay, z);
    } else if (p.kind == 0) {  // wall segment: a lattice of surf points, its two end edges as corner lines
      const double L = std::hypot(p.bx - p.ax, p.by - p.ay);
      const double g0 = w.ground_s(p.ax, p.ay), g1 = w.ground_s(p.bx, p.by);
      const int na = std::max(1, (int)std::floor(L / surf_step));
      for (int a = 0; a <= na; ++a) {
        const double f = (double)a / na, xs = p.ax + f * (p.bx - p.ax), ys = p.ay + f * (p.by - p.ay);
        const double g = w.ground_s(xs, ys);
        for (double z = g + 0.5 * surf_step; z < p.zhi; z += surf_step) emit(S, xs, ys, z);
      }
      for (double z = g0 + 0.1; z < p.zhi; z += corner_step) emit(C, p.ax, p.ay, z);
      for (double z = g1 + 0.1; z < p.zhi; z += corner_step) emit(C, p.bx, p.by, z);
    } else {  // car roof
      for (double xs = p.ax; xs <= p.bx; xs += surf_step)
        for (double ys = p.ay; ys <= p.by; ys += surf_step) emit(S, xs, ys, p.zhi);
    }
  }
  if ((int)(C.size() / 4) > cap_corner || (int)(S.size() / 4) > cap_surf) return -1;
  std::memcpy(corner, C.data(), C.size() * sizeof(float));
  std::memcpy(surf, S.data(), S.size() * sizeof(float));
  *n_corner = (int)(C.size() / 4);
  *n_surf = (int)(S.size() / 4);
  return 0;
}

}  // extern "C"
