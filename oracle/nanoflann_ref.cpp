// oracle/nanoflann_ref.cpp -- thin adaptor around the KD-tree the reference
// vendors (include/scancontext/nanoflann.hpp, v1.3.2, FLANN-derived single
// index with the same L2_Simple loop and insertion result set as FLANN).
//
// TEST INFRASTRUCTURE ONLY. The header is NOT copied into this repo: it is
// compiled from where it lies under /root/reference by oracle/Makefile, and the
// output goes to oracle/_ref/ (git-ignored). It pins the oracle's kNN
// (oracle/s2m_oracle.cpp KdTree / knn_brute) to code from the reference tree.
#include <cstdint>
#include <cmath>
#include <vector>
#include "nanoflann.hpp"

namespace {
struct Cloud3 {
  const float* p;  // xyzi, stride 4
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return p[4 * idx + dim]; }
  template <class BBOX> bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Cloud3>, Cloud3, 3, int> Tree;
}  // namespace

extern "C" int ref_nanoflann_knn5(const float* map_xyzi, int m, const float* q_xyz, int nq,
                                  int32_t* idx5, float* d2_5) {
  Cloud3 c{map_xyzi, (size_t)m};
  Tree tree(3, c, nanoflann::KDTreeSingleIndexAdaptorParams(15));
  tree.buildIndex();
  for (int i = 0; i < nq; ++i) {
    int idx[5];
    float d2[5];
    size_t found = tree.knnSearch(q_xyz + 3 * i, 5, idx, d2);
    for (int j = 0; j < 5; ++j) {
      idx5[5 * i + j] = j < (int)found ? idx[j] : -1;
      d2_5[5 * i + j] = j < (int)found ? d2[j] : INFINITY;
    }
  }
  return 0;
}
