#pragma once
// stub of the slice of roscpp the mapping node uses
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include <ros/time.h>
#define ROS_FATAL(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_WARN(...) std::fprintf(stderr, __VA_ARGS__)
namespace ros {
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void spin() {}
inline void shutdown() {}
struct Subscriber {};
struct Publisher { template <typename M> void publish(const M&) const {} };
struct NodeHandle {
  template <typename T> bool param(const std::string&, T& out, const T& def) const { out = def; return false; }
  template <typename M, typename F> Subscriber subscribe(const std::string&, unsigned, F cb) {
    std::function<void(const std::shared_ptr<const M>&)> f = cb;  // the callback must accept M::ConstPtr
    (void)f;
    return Subscriber();
  }
  template <typename M> Publisher advertise(const std::string&, unsigned) { return Publisher(); }
};
}
