#pragma once
#include <cstddef>
#include <vector>
namespace pcl {
template <typename P> struct PointCloud {
  std::vector<P> points;
  std::size_t size() const { return points.size(); }
  void resize(std::size_t n) { points.resize(n); }
  P& operator[](std::size_t i) { return points[i]; }
  const P& operator[](std::size_t i) const { return points[i]; }
};
}
