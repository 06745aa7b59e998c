// host-only view of the shared arithmetic and key helpers (no CUDA headers)
#pragma once
#include "s2m_math.cuh"
namespace s2m {
constexpr int kCubeBiasIJ = 2048, kCubeBiasK = 32;
inline uint32_t pack_cube(int ci, int cj, int ck) {
  return ((uint32_t)(ci + kCubeBiasIJ) << 18) | ((uint32_t)(cj + kCubeBiasIJ) << 6) | (uint32_t)(ck + kCubeBiasK);
}
inline uint64_t store_key(uint32_t cube, uint32_t pending, uint64_t payload) {
  return ((uint64_t)cube << 34) | ((uint64_t)pending << 33) | (payload & 0x1FFFFFFFFull);
}
inline int voxel_rel(float p, int cube, float inv_leaf) {
  float lo = (float)(50.0 * (double)cube - 25.0);
  return voxel_coord(p, inv_leaf) - voxel_coord(lo, inv_leaf);
}
}  // namespace s2m
