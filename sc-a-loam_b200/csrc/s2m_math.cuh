// s2m_math.cuh -- per-point and per-solve arithmetic of the scan-to-map path.
//
// Everything here is __host__ __device__ so the same code runs inside the CUDA
// kernels and inside the host-side unit tests (tests/test_host_math.py drive it
// through csrc/s2m_hostmath.cpp and compare with the oracle).
//
// Exactness classes
//   EXACT (bit parity with the reference arithmetic; no FMA contraction):
//     xf_point()   pointAssociateToMap, laserMapping.cpp:155-164 (double, then float)
//     cube_of()    cube index rule, laserMapping.cpp:313-322 / :742-751
//     dist2()      FLANN L2_Simple float distance used by nearestKSearch (:583, :649)
//     voxel_coord() pcl::VoxelGrid lattice floor(p * inverse_leaf)
//   TOLERANCE (FP64, any order, FMA allowed; parity 1e-4 m / 1e-5 rad on poses):
//     edge_fit / plane_fit (:585-622, :651-687), factor residuals + Jacobians
//     (lidarFactor.hpp:12-55, :106-138 with EigenQuaternionParameterization),
//     HuberLoss(0.1) (:566), the Ceres trust-region step (:713-721).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define S2M_HD __host__ __device__ __forceinline__
#else
#define S2M_HD inline
#endif

namespace s2m {

// ---- rounding-exact primitives -------------------------------------------
#if defined(__CUDA_ARCH__)
S2M_HD double xdmul(double a, double b) { return __dmul_rn(a, b); }
S2M_HD double xdadd(double a, double b) { return __dadd_rn(a, b); }
S2M_HD double xdsub(double a, double b) { return __dsub_rn(a, b); }
S2M_HD double xddiv(double a, double b) { return __ddiv_rn(a, b); }
S2M_HD float xfmul(float a, float b) { return __fmul_rn(a, b); }
S2M_HD float xfadd(float a, float b) { return __fadd_rn(a, b); }
S2M_HD float xfsub(float a, float b) { return __fsub_rn(a, b); }
S2M_HD float xfdiv(float a, float b) { return __fdiv_rn(a, b); }
#else  // host objects are built with -ffp-contract=off
S2M_HD double xdmul(double a, double b) { return a * b; }
S2M_HD double xdadd(double a, double b) { return a + b; }
S2M_HD double xdsub(double a, double b) { return a - b; }
S2M_HD double xddiv(double a, double b) { return a / b; }
S2M_HD float xfmul(float a, float b) { return a * b; }
S2M_HD float xfadd(float a, float b) { return a + b; }
S2M_HD float xfsub(float a, float b) { return a - b; }
S2M_HD float xfdiv(float a, float b) { return a / b; }
#endif

// Eigen 3.3 Quaternion * Vector3 (uv = 2 q.vec x v; v + w uv + q.vec x uv), no contraction.
S2M_HD void quat_rotate_exact(const double q[4], double vx, double vy, double vz, double out[3]) {
  double ux = xdsub(xdmul(q[1], vz), xdmul(q[2], vy));
  double uy = xdsub(xdmul(q[2], vx), xdmul(q[0], vz));
  double uz = xdsub(xdmul(q[0], vy), xdmul(q[1], vx));
  ux = xdadd(ux, ux); uy = xdadd(uy, uy); uz = xdadd(uz, uz);
  double cx = xdsub(xdmul(q[1], uz), xdmul(q[2], uy));
  double cy = xdsub(xdmul(q[2], ux), xdmul(q[0], uz));
  double cz = xdsub(xdmul(q[0], uy), xdmul(q[1], ux));
  out[0] = xdadd(xdadd(vx, xdmul(q[3], ux)), cx);
  out[1] = xdadd(xdadd(vy, xdmul(q[3], uy)), cy);
  out[2] = xdadd(xdadd(vz, xdmul(q[3], uz)), cz);
}
// pointAssociateToMap: q*p + t in double, rounded to float.
S2M_HD void xf_point(const double pose[7], float x, float y, float z, float out[3]) {
  double r[3];
  quat_rotate_exact(pose, (double)x, (double)y, (double)z, r);
  out[0] = (float)xdadd(r[0], pose[4]);
  out[1] = (float)xdadd(r[1], pose[5]);
  out[2] = (float)xdadd(r[2], pose[6]);
}
// World cube coordinate (the reference's index minus laserCloudCen*):
// int((v + 25.0) / 50.0), one less when v + 25.0 < 0 (trunc-then-decrement).
S2M_HD int cube_of(double v) {
  double s = xdadd(v, 25.0);
  int c = (int)xddiv(s, 50.0);
  if (s < 0) c--;
  return c;
}
S2M_HD float dist2(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = xfsub(qx, px), dy = xfsub(qy, py), dz = xfsub(qz, pz);
  return xfadd(xfadd(xfmul(dx, dx), xfmul(dy, dy)), xfmul(dz, dz));
}
S2M_HD int voxel_coord(float v, float inv_leaf) { return (int)floorf(xfmul(v, inv_leaf)); }

// Eigen quaternion product (scalar path), no contraction -- rows A and U run on
// the host with exactly this.
S2M_HD void quat_mul_exact(const double a[4], const double b[4], double o[4]) {
  double x = xdsub(xdadd(xdadd(xdmul(a[3], b[0]), xdmul(a[0], b[3])), xdmul(a[1], b[2])), xdmul(a[2], b[1]));
  double y = xdsub(xdadd(xdadd(xdmul(a[3], b[1]), xdmul(a[1], b[3])), xdmul(a[2], b[0])), xdmul(a[0], b[2]));
  double z = xdsub(xdadd(xdadd(xdmul(a[3], b[2]), xdmul(a[2], b[3])), xdmul(a[0], b[1])), xdmul(a[1], b[0]));
  double w = xdsub(xdsub(xdsub(xdmul(a[3], b[3]), xdmul(a[0], b[0])), xdmul(a[1], b[1])), xdmul(a[2], b[2]));
  o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}
S2M_HD void quat_inverse_exact(const double q[4], double o[4]) {
  double n2 = xdadd(xdadd(xdadd(xdmul(q[0], q[0]), xdmul(q[1], q[1])), xdmul(q[2], q[2])), xdmul(q[3], q[3]));
  if (n2 > 0) { o[0] = xddiv(-q[0], n2); o[1] = xddiv(-q[1], n2); o[2] = xddiv(-q[2], n2); o[3] = xddiv(q[3], n2); }
  else { o[0] = o[1] = o[2] = o[3] = 0; }
}

// ---- tolerance-class FP64 ---------------------------------------------------
S2M_HD void quat_rotate(const double q[4], const double v[3], double out[3]) {
  double ux = q[1] * v[2] - q[2] * v[1], uy = q[2] * v[0] - q[0] * v[2], uz = q[0] * v[1] - q[1] * v[0];
  ux += ux; uy += uy; uz += uz;
  out[0] = v[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
  out[1] = v[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
  out[2] = v[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}

// The two largest eigenvalues of a symmetric 3x3 (cyclic Jacobi on the six
// unique entries, FP64) and, only if asked, the unit eigenvector of the largest.
// a = {xx, xy, xz, yy, yz, zz}.  The vector is the best-conditioned cross product
// of two rows of (A - lambda_max I): exact to rounding when lambda_max is isolated,
// which the caller guarantees (it is used only when lambda_max > 3 lambda_mid).
S2M_HD void eig3_top(const double a_in[6], double& lam_mid, double& lam_max) {
  double a00 = a_in[0], a01 = a_in[1], a02 = a_in[2], a11 = a_in[3], a12 = a_in[4], a22 = a_in[5];
  const double total = a00 * a00 + a11 * a11 + a22 * a22 + 2 * (a01 * a01 + a02 * a02 + a12 * a12);
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = a01 * a01 + a02 * a02 + a12 * a12;
    if (off <= 1e-22 * total || off == 0.0) break;  // eigenvalue error ~ off^2 / gap: far below the ratio test's resolution
#define S2M_JROT(app, aqq, apq, arp, arq)                               \
  if (apq != 0.0) {                                                     \
    double theta = (aqq - app) / (2.0 * apq);                           \
    double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0)); \
    double c = 1.0 / sqrt(t * t + 1.0), s = t * c;                      \
    app -= t * apq; aqq += t * apq; apq = 0.0;                          \
    double rp = arp, rq = arq;                                          \
    arp = c * rp - s * rq; arq = s * rp + c * rq;                       \
  }
    S2M_JROT(a00, a11, a01, a02, a12)
    S2M_JROT(a00, a22, a02, a01, a12)
    S2M_JROT(a11, a22, a12, a01, a02)
#undef S2M_JROT
  }
  const double hi = fmax(a00, fmax(a11, a22)), lo = fmin(a00, fmin(a11, a22));
  lam_max = hi;
  lam_mid = (a00 + a11 + a22) - hi - lo;
}
// The same two eigenvalues in closed form (Smith's trigonometric solution of the characteristic cubic):
// a handful of multiplications, one acos and two cos instead of ~15 Jacobi rotations.  Absolute error a few
// ulp of lam_max -- enough to DECIDE lam_max > 3 lam_mid except in a thin band around equality, where the caller
// falls back to the iterative solver.
S2M_HD void eig3_top_closed(const double a[6], double& lam_mid, double& lam_max) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
  const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
  const double q = (a00 + a11 + a22) / 3.0;
  const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
  const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
  if (!(p2 > 0.0)) { lam_mid = lam_max = q; return; }
  const double p = sqrt(p2 / 6.0);
  const double det = b00 * (b11 * b22 - a12 * a12) - a01 * (a01 * b22 - a12 * a02) + a02 * (a01 * a12 - b11 * a02);
  const double r = fmin(1.0, fmax(-1.0, det / (2.0 * p * p * p)));
  const double phi = acos(r) / 3.0;
  const double e_max = q + 2.0 * p * cos(phi);
  const double e_min = q + 2.0 * p * cos(phi + 2.0943951023931953);  // + 2 pi / 3
  lam_max = e_max;
  lam_mid = 3.0 * q - e_max - e_min;
}
S2M_HD void eig3_vector(const double a[6], double lam, double dir[3]) {
  const double r0[3] = {a[0] - lam, a[1], a[2]}, r1[3] = {a[1], a[3] - lam, a[4]}, r2[3] = {a[2], a[4], a[5] - lam};
  double c0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
  double c1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
  double c2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
  double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
  double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  double n2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
  if (n1 > n0) { c0[0] = c1[0]; c0[1] = c1[1]; c0[2] = c1[2]; n0 = n1; }
  if (n2 > n0) { c0[0] = c2[0]; c0[1] = c2[1]; c0[2] = c2[2]; n0 = n2; }
  const double inv = 1.0 / sqrt(n0);
  dir[0] = c0[0] * inv; dir[1] = c0[1] * inv; dir[2] = c0[2] * inv;
}

// Edge fit (laserMapping.cpp:585-622): 5 neighbours -> centre and unit direction;
// accepted iff lambda_max > 3 * lambda_mid.
S2M_HD bool edge_fit(const float nb[5][3], double c[3], double u[3]) {
  c[0] = c[1] = c[2] = 0;
  for (int j = 0; j < 5; ++j)
    for (int k = 0; k < 3; ++k) c[k] = c[k] + (double)nb[j][k];
  for (int k = 0; k < 3; ++k) c[k] = c[k] / 5.0;
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (int j = 0; j < 5; ++j) {
    double x = (double)nb[j][0] - c[0], y = (double)nb[j][1] - c[1], z = (double)nb[j][2] - c[2];
    a[0] += x * x; a[1] += x * y; a[2] += x * z; a[3] += y * y; a[4] += y * z; a[5] += z * z;
  }
  double lmid, lmax;
  eig3_top_closed(a, lmid, lmax);
  // within 1e-9 of the acceptance threshold (:612) the closed form does not decide: iterate
  if (fabs(lmax - 3 * lmid) <= 1e-9 * lmax) eig3_top(a, lmid, lmax);
  if (!(lmax > 3 * lmid)) { u[0] = u[1] = u[2] = 0; return false; }
  eig3_vector(a, lmax, u);
  return true;
}

// Plane fit (laserMapping.cpp:651-687): least squares A n = -1 by column-pivoted
// Householder QR in FP64, n normalised, valid iff all five |n.p + d| <= 0.2.
S2M_HD bool plane_fit(const float nb[5][3], double n[3], double& d) {
  // Columns are kept as three separate arrays and every index below is a compile-time constant
  // (the pivot choice moves data with selects), so on the device the whole factorisation lives in
  // registers: no local-memory array.
  double C[3][5], b[5] = {-1, -1, -1, -1, -1};
#pragma unroll
  for (int j = 0; j < 5; ++j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) C[k][j] = (double)nb[j][k];
  }
  int perm[3] = {0, 1, 2};
  double cn[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) s += C[k][r] * C[k][r];
    cn[k] = s;
  }
  const double cmax = fmax(cn[0], fmax(cn[1], cn[2]));
  const double tiny = cmax * (DBL_EPSILON * DBL_EPSILON);  // rank threshold on squared norms
  int rank = 3;
  double diag[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int big = k;
    double cbig = cn[k];
#pragma unroll
    for (int j = k + 1; j < 3; ++j) if (cn[j] > cbig) { big = j; cbig = cn[j]; }
    if (rank == 3 && cbig < tiny * (double)(5 - k) / 5.0) rank = k;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      if (big == j) {  // swap columns k and j
#pragma unroll
        for (int r = 0; r < 5; ++r) { const double t = C[k][r]; C[k][r] = C[j][r]; C[j][r] = t; }
        const double t = cn[k]; cn[k] = cn[j]; cn[j] = t;
        const int ti = perm[k]; perm[k] = perm[j]; perm[j] = ti;
      }
    }
    double tail2 = 0;
#pragma unroll
    for (int r = k + 1; r < 5; ++r) tail2 += C[k][r] * C[k][r];
    const double c0 = C[k][k];
    double tau, beta;
    if (tail2 <= DBL_MIN) {
      tau = 0; beta = c0;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) C[k][r] = 0;
    } else {
      beta = sqrt(c0 * c0 + tail2);
      if (c0 >= 0) beta = -beta;
      const double inv = 1.0 / (c0 - beta);
#pragma unroll
      for (int r = k + 1; r < 5; ++r) C[k][r] *= inv;
      tau = (beta - c0) / beta;
    }
    diag[k] = beta;
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {
      double tmp = C[j][k];
#pragma unroll
      for (int r = k + 1; r < 5; ++r) tmp += C[k][r] * C[j][r];
      tmp *= tau;
      C[j][k] -= tmp;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) C[j][r] -= C[k][r] * tmp;
    }
    {
      double tmp = b[k];
#pragma unroll
      for (int r = k + 1; r < 5; ++r) tmp += C[k][r] * b[r];
      tmp *= tau;
      b[k] -= tmp;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) b[r] -= C[k][r] * tmp;
    }
#pragma unroll
    for (int j = k + 1; j < 3; ++j) {  // exact recomputation of the trailing column norms
      double s = 0;
#pragma unroll
      for (int r = k + 1; r < 5; ++r) s += C[j][r] * C[j][r];
      cn[j] = s;
    }
  }
  // back substitution on R (upper triangle: R[i][j] = C[j][i], diagonal in diag), unknowns beyond the rank are 0
  double y[3] = {0, 0, 0};
#pragma unroll
  for (int i = 2; i >= 0; --i) {
    if (i < rank) {
      double s = b[i];
#pragma unroll
      for (int j = i + 1; j < 3; ++j) if (j < rank) s -= C[j][i] * y[j];
      y[i] = s / diag[i];
    }
  }
  double x[3] = {0, 0, 0};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (i < rank) {
#pragma unroll
      for (int a = 0; a < 3; ++a) if (perm[i] == a) x[a] = y[i];
    }
  }
  double nn = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  d = 1.0 / nn;
  n[0] = x[0] / nn; n[1] = x[1] / nn; n[2] = x[2] / nn;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 5; ++j)
    if (!(fabs(n[0] * (double)nb[j][0] + n[1] * (double)nb[j][1] + n[2] * (double)nb[j][2] + d) <= 0.2)) ok = false;
  return ok;
}

// 28 accumulators of one evaluation: upper-triangular JtJ (21), Jtr (6), cost.
struct Sums28 {
  double v[28];
  S2M_HD void zero() { for (int i = 0; i < 28; ++i) v[i] = 0; }
};

S2M_HD void huber_scale(double s, double& w, double& half_rho) {  // HuberLoss(0.1) + Corrector
  if (s > 0.01) {
    double r = sqrt(s);
    half_rho = 0.5 * (0.2 * r - 0.01);
    w = sqrt(fmax(DBL_MIN, 0.1 / r));
  } else { half_rho = 0.5 * s; w = 1.0; }
}
S2M_HD void add_row(Sums28& S, const double J[6], double r) {
  int k = 0;
  for (int a = 0; a < 6; ++a) {
    for (int b = a; b < 6; ++b) S.v[k++] += J[a] * J[b];
    S.v[21 + a] += J[a] * r;
  }
}
// LidarEdgeFactor (s=1): r = (lp - c) x u ; dr/dlp = -[u]x ; tangent Jacobian
// [dr/dlp * (-2 [Rp]x) , dr/dlp] (EigenQuaternionParameterization, delta = half angle).
S2M_HD void accum_edge(Sums28& S, const double pose[7], const double cp[3], const double c[3],
                       const double u[3]) {
  double Rp[3];
  quat_rotate(pose, cp, Rp);
  double v[3] = {Rp[0] + pose[4] - c[0], Rp[1] + pose[5] - c[1], Rp[2] + pose[6] - c[2]};
  double r[3] = {v[1] * u[2] - v[2] * u[1], v[2] * u[0] - v[0] * u[2], v[0] * u[1] - v[1] * u[0]};
  double s = r[0] * r[0] + r[1] * r[1] + r[2] * r[2], w, hr;
  huber_scale(s, w, hr);
  S.v[27] += hr;
  // rows of M = -[u]x :  m0 = (0, u2, -u1), m1 = (-u2, 0, u0), m2 = (u1, -u0, 0)
  const double M[3][3] = {{0, u[2], -u[1]}, {-u[2], 0, u[0]}, {u[1], -u[0], 0}};
  for (int k = 0; k < 3; ++k) {
    const double* m = M[k];
    double J[6];
    J[0] = -2.0 * (m[1] * Rp[2] - m[2] * Rp[1]) * w;
    J[1] = -2.0 * (m[2] * Rp[0] - m[0] * Rp[2]) * w;
    J[2] = -2.0 * (m[0] * Rp[1] - m[1] * Rp[0]) * w;
    J[3] = m[0] * w; J[4] = m[1] * w; J[5] = m[2] * w;
    add_row(S, J, r[k] * w);
  }
}
// LidarPlaneNormFactor: r = n . lp + d
S2M_HD void accum_plane(Sums28& S, const double pose[7], const double cp[3], const double n[3], double d) {
  double Rp[3];
  quat_rotate(pose, cp, Rp);
  double r = n[0] * (Rp[0] + pose[4]) + n[1] * (Rp[1] + pose[5]) + n[2] * (Rp[2] + pose[6]) + d;
  double w, hr;
  huber_scale(r * r, w, hr);
  S.v[27] += hr;
  double J[6];
  J[0] = -2.0 * (n[1] * Rp[2] - n[2] * Rp[1]) * w;
  J[1] = -2.0 * (n[2] * Rp[0] - n[0] * Rp[2]) * w;
  J[2] = -2.0 * (n[0] * Rp[1] - n[1] * Rp[0]) * w;
  J[3] = n[0] * w; J[4] = n[1] * w; J[5] = n[2] * w;
  add_row(S, J, r * w);
}

// ---- Ceres trust-region Levenberg-Marquardt on the reduced 6x6 system ---------
struct LmState {
  double x[7];        // accepted pose (q xyzw, t)
  double xc[7];       // candidate pose being evaluated
  double cost;        // cost at x
  double H[21], g[6]; // unscaled tangent-space normal equations at x
  double scale[6];    // jacobi scaling, fixed at iteration 0
  double diag[6];
  double step[6];     // scaled step of the pending candidate
  double radius, decrease_factor, model_change, x_norm;
  int reuse_diag, iteration, done, termination, invalid_run, have_candidate, n_res;
  // log (mirrors the oracle's SolveLog)
  double init_sums[28];
  double it_log[4][6];
  double initial_cost, final_cost;
};

S2M_HD void quat_plus(const double x[4], const double dl[3], double o[4]) {
  double nd = sqrt(dl[0] * dl[0] + dl[1] * dl[1] + dl[2] * dl[2]);
  if (nd > 0.0) {
    double sbd = sin(nd) / nd;
    double dq[4] = {sbd * dl[0], sbd * dl[1], sbd * dl[2], cos(nd)};
    o[0] = dq[3] * x[0] + dq[0] * x[3] + dq[1] * x[2] - dq[2] * x[1];
    o[1] = dq[3] * x[1] + dq[1] * x[3] + dq[2] * x[0] - dq[0] * x[2];
    o[2] = dq[3] * x[2] + dq[2] * x[3] + dq[0] * x[1] - dq[1] * x[0];
    o[3] = dq[3] * x[3] - dq[0] * x[0] - dq[1] * x[1] - dq[2] * x[2];
  } else { o[0] = x[0]; o[1] = x[1]; o[2] = x[2]; o[3] = x[3]; }
}
S2M_HD void pose_plus(const double x[7], const double d[6], double o[7]) {
  quat_plus(x, d, o);
  o[4] = x[4] + d[3]; o[5] = x[5] + d[4]; o[6] = x[6] + d[5];
}
S2M_HD int tri(int a, int b) { return a <= b ? a * 6 - a * (a - 1) / 2 + (b - a) : b * 6 - b * (b - 1) / 2 + (a - b); }

S2M_HD double grad_max_norm(const double x[7], const double g[6]) {
  double neg[6], xp[7], m = 0;
  for (int i = 0; i < 6; ++i) neg[i] = -g[i];
  pose_plus(x, neg, xp);
  for (int i = 0; i < 7; ++i) m = fmax(m, fabs(x[i] - xp[i]));
  return m;
}

// Solve (S H S + diag/radius) y = S g by Cholesky; step = -y; fills model_change.
// Every loop is fully unrolled so the 6x6 work stays in registers (this runs on a
// single thread at the tail of the evaluation kernel: its latency is on the
// critical path of every LM iteration).
S2M_HD bool lm_compute_step(LmState& L) {
  double A[21], gs[6];  // upper triangle of S H S, row-major packed like L.H
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    gs[a] = L.g[a] * L.scale[a];
#pragma unroll
    for (int b = a; b < 6; ++b) A[tri(a, b)] = L.H[tri(a, b)] * L.scale[a] * L.scale[b];
  }
  if (!L.reuse_diag) {
#pragma unroll
    for (int a = 0; a < 6; ++a) L.diag[a] = fmin(fmax(A[tri(a, a)], 1e-6), 1e32);
  }
  L.reuse_diag = 1;
  double C[21];  // Cholesky factor, C[tri(j,i)] = L_ij for i >= j (stored by column index first)
#pragma unroll
  for (int k = 0; k < 21; ++k) C[k] = A[k];
#pragma unroll
  for (int a = 0; a < 6; ++a) C[tri(a, a)] += L.diag[a] / L.radius;
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = C[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; ++k) s -= C[tri(k, j)] * C[tri(k, j)];
    if (!(s > 0)) ok = false;
    const double dj = sqrt(ok ? s : 1.0);
    C[tri(j, j)] = dj;
    const double inv = 1.0 / dj;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double t = C[tri(j, i)];
#pragma unroll
      for (int k = 0; k < j; ++k) t -= C[tri(k, i)] * C[tri(k, j)];
      C[tri(j, i)] = t * inv;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double s = gs[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= C[tri(k, i)] * y[k];
    y[i] = s / C[tri(i, i)];
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= C[tri(i, k)] * y[k];
    y[i] = s / C[tri(i, i)];
  }
  double mcc = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    L.step[a] = ok ? -y[a] : 0.0;
    if (!(L.step[a] == L.step[a]) || fabs(L.step[a]) > 1e300) ok = false;
  }
  if (ok) {
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      double hs = 0;
#pragma unroll
      for (int b = 0; b < 6; ++b) hs += A[tri(a, b)] * L.step[b];
      mcc += L.step[a] * (gs[a] + 0.5 * hs);
    }
    mcc = -mcc;
  }
  L.model_change = mcc;
  return ok && mcc > 0.0;
}

// Advance the minimizer until it needs a cost/Jacobian evaluation at L.xc
// (have_candidate=1) or terminates (done=1).  Mirrors the loop head of
// TrustRegionMinimizer::Minimize, including invalid-step handling.
S2M_HD void lm_next_candidate(LmState& L, int max_iterations) {
  L.have_candidate = 0;
  for (;;) {
    if (L.iteration >= max_iterations) { L.done = 1; L.termination = 0; return; }
    if (grad_max_norm(L.x, L.g) <= 1e-10) { L.done = 1; L.termination = 1; return; }
    if (L.radius <= 1e-32) { L.done = 1; L.termination = 4; return; }
    L.iteration++;
    double* lg = L.it_log[L.iteration - 1 < 4 ? L.iteration - 1 : 3];
    lg[0] = L.cost; lg[1] = 0; lg[2] = L.radius; lg[3] = 0; lg[4] = 0; lg[5] = 0;
    bool valid = lm_compute_step(L);
    lg[4] = L.model_change;
    if (!valid) {
      if (++L.invalid_run >= 5) { L.done = 1; L.termination = 6; return; }
      L.radius = L.radius / L.decrease_factor;
      L.decrease_factor *= 2.0;
      L.reuse_diag = 1;
      continue;
    }
    L.invalid_run = 0;
    double delta[6];
    for (int a = 0; a < 6; ++a) delta[a] = L.step[a] * L.scale[a];
    pose_plus(L.x, delta, L.xc);
    L.have_candidate = 1;
    return;
  }
}

// After the first evaluation at x0.
S2M_HD void lm_begin(LmState& L, const double x0[7], const Sums28& S, int n_res, int max_iterations) {
  for (int i = 0; i < 7; ++i) L.x[i] = L.xc[i] = x0[i];
  for (int i = 0; i < 28; ++i) L.init_sums[i] = S.v[i];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 6; ++j) L.it_log[i][j] = 0;
  L.n_res = n_res;
  L.iteration = 0; L.done = 0; L.termination = 0; L.invalid_run = 0; L.have_candidate = 0;
  L.initial_cost = L.final_cost = S.v[27];
  L.cost = S.v[27];
  for (int i = 0; i < 21; ++i) L.H[i] = S.v[i];
  for (int i = 0; i < 6; ++i) L.g[i] = S.v[21 + i];
  if (n_res == 0) { L.done = 1; L.termination = 5; return; }
  for (int a = 0; a < 6; ++a) L.scale[a] = 1.0 / (1.0 + sqrt(L.H[tri(a, a)]));
  L.radius = 1e4; L.decrease_factor = 2.0; L.reuse_diag = 0;
  double xn = 0;
  for (int i = 0; i < 7; ++i) xn += L.x[i] * L.x[i];
  L.x_norm = sqrt(xn);
  lm_next_candidate(L, max_iterations);
}

// After the evaluation at the candidate L.xc.
S2M_HD void lm_after_eval(LmState& L, const Sums28& S, int max_iterations) {
  if (L.done || !L.have_candidate) return;
  double* lg = L.it_log[L.iteration - 1 < 4 ? L.iteration - 1 : 3];
  const double cand_cost = S.v[27];
  double sn = 0;
  for (int i = 0; i < 7; ++i) sn += (L.x[i] - L.xc[i]) * (L.x[i] - L.xc[i]);
  sn = sqrt(sn);
  lg[3] = sn;
  L.have_candidate = 0;
  if (sn <= 1e-8 * (L.x_norm + 1e-8)) { L.done = 1; L.termination = 2; return; }
  const double cost_change = L.cost - cand_cost;
  lg[1] = cost_change;
  if (fabs(cost_change) <= 1e-6 * L.cost) { L.done = 1; L.termination = 3; return; }
  const double rel = cost_change / L.model_change;
  if (rel > 1e-3) {
    lg[5] = 1;
    for (int i = 0; i < 7; ++i) L.x[i] = L.xc[i];
    double xn = 0;
    for (int i = 0; i < 7; ++i) xn += L.x[i] * L.x[i];
    L.x_norm = sqrt(xn);
    L.cost = cand_cost;
    L.final_cost = cand_cost;
    for (int i = 0; i < 21; ++i) L.H[i] = S.v[i];
    for (int i = 0; i < 6; ++i) L.g[i] = S.v[21 + i];
    const double tq = 2.0 * rel - 1.0;
    double dnm = 1.0 - tq * tq * tq;  // Ceres: 1 - pow(2 rho - 1, 3)
    L.radius = L.radius / fmax(1.0 / 3.0, dnm);
    L.radius = fmin(1e16, L.radius);
    L.decrease_factor = 2.0;
    L.reuse_diag = 0;
  } else {
    L.radius = L.radius / L.decrease_factor;
    L.decrease_factor *= 2.0;
    L.reuse_diag = 1;
  }
  lm_next_candidate(L, max_iterations);
}

}  // namespace s2m
