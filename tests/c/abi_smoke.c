/* Plain-C consumer of include/s2m.h: what a maintainer's binding sees (compiled and run by
 * tests/test_abi.py; host-only entry points, so it runs without a GPU). */
#include <stdio.h>
#include <string.h>

#include "s2m.h"

int main(int argc, char** argv) {
  s2m_params p;
  memset(&p, 0xAB, sizeof p);
  s2m_default_params(&p);
  if (!(p.line_res > 0.39f && p.line_res < 0.41f && p.plane_res > 0.79f && p.plane_res < 0.81f)) return 1; /* laserMapping.cpp:913-916 */
  if (p.batch != 1 || p.shard_world != 1 || p.lanes != 0) return 2;
  if (strcmp(s2m_strerror(S2M_OK), "ok") != 0 || strlen(s2m_strerror(S2M_ERR_IO)) == 0) return 3;
  float lo = 0.f, hi = 0.f;
  if (s2m_shard_slab(0, 2, &lo, &hi) != S2M_OK || !(lo < hi)) return 4;
  if (argc > 1) { /* PCD round trip */
    float pts[8] = {1.f, 2.f, 3.f, 0.5f, -4.f, 5.5f, 6.25f, 17.1f}, back[8];
    if (s2m_pcd_write(argv[1], pts, 2) != S2M_OK) return 5;
    if (s2m_pcd_read(argv[1], back, 2) != 2 || memcmp(pts, back, sizeof pts) != 0) return 6;
  }
  {
    s2m_ctx* ctx = NULL; /* bad arguments are rejected before any device is touched */
    if (s2m_create(NULL, &ctx) != S2M_ERR_ARG || ctx != NULL) return 7;
    if (s2m_fx_create(NULL, NULL) != S2M_ERR_ARG) return 8;
  }
  printf("abi smoke ok: s2m_params %zu bytes, s2m_stats %zu bytes\n", sizeof(s2m_params), sizeof(s2m_stats));
  return 0;
}
