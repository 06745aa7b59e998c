// s2m_hostmath.cpp -- host build of csrc/s2m_math.cuh for CPU unit tests.
// The functions below are the SAME source the kernels compile (s2m_math.cuh is
// __host__ __device__); this file only gives them a C ABI so pytest can compare
// them with the oracle without a GPU.  Not part of libs2m.so.
#include <cstring>
#include <vector>

#include "s2m_internal_host.h"

using namespace s2m;

extern "C" {

void hm_xf_point(const double pose[7], const float* in_xyz, int n, float* out_xyz) {
  for (int i = 0; i < n; ++i) xf_point(pose, in_xyz[3 * i], in_xyz[3 * i + 1], in_xyz[3 * i + 2], out_xyz + 3 * i);
}
int hm_cube_of(double v) { return cube_of(v); }
float hm_dist2(const float q[3], const float p[3]) { return dist2(q[0], q[1], q[2], p[0], p[1], p[2]); }
int hm_edge_fit(const float nb[15], double c[3], double u[3]) {
  float a[5][3];
  std::memcpy(a, nb, sizeof(a));
  return edge_fit(a, c, u) ? 1 : 0;
}
int hm_plane_fit(const float nb[15], double n[3], double* d) {
  float a[5][3];
  std::memcpy(a, nb, sizeof(a));
  return plane_fit(a, n, *d) ? 1 : 0;
}
// 28 sums of a problem given as kinds + 10 doubles per block:
// edge: cp[3], c[3], u[3]; plane: cp[3], n[3], d
void hm_sums(const int* kinds, const double* data10, int nb, const double pose[7], double out28[28]) {
  Sums28 S;
  S.zero();
  for (int i = 0; i < nb; ++i) {
    const double* d = data10 + 10 * i;
    if (kinds[i] == 0) accum_edge(S, pose, d, d + 3, d + 6);
    else accum_plane(S, pose, d, d + 3, d[6]);
  }
  std::memcpy(out28, S.v, sizeof(S.v));
}
// the whole LM schedule on the host, evaluation = hm_sums; mirrors what the
// lm_begin / evaluate / lm_after kernels do on the device
int hm_solve(const int* kinds, const double* data10, int nb, double x[7], int max_iter, double* iters24,
             int* n_iter, int* termination) {
  LmState L;
  std::memset(&L, 0, sizeof(L));
  Sums28 S;
  hm_sums(kinds, data10, nb, x, S.v);
  lm_begin(L, x, S, nb, max_iter);
  while (!L.done && L.have_candidate) {
    hm_sums(kinds, data10, nb, L.xc, S.v);
    lm_after_eval(L, S, max_iter);
  }
  std::memcpy(x, L.x, 56);
  if (iters24) std::memcpy(iters24, L.it_log, 24 * 8);
  if (n_iter) *n_iter = L.iteration;
  if (termination) *termination = L.termination;
  return 0;
}
unsigned long long hm_store_key(int ci, int cj, int ck, unsigned pending, unsigned long long payload) {
  return store_key(pack_cube(ci, cj, ck), pending, payload);
}
int hm_voxel_rel(float p, int cube, float inv_leaf) { return voxel_rel(p, cube, inv_leaf); }

}  // extern "C"
