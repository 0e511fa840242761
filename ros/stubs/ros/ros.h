// Stand-in for <ros/ros.h>: ONLY what the shim sources under ros/ use, so that `make -C ros check` can
// syntax-check them in a container without ROS. Not part of the product; a real build uses roscpp.
#pragma once
#include <cstdio>
#include <string>
namespace ros {
inline void init(int&, char**, const std::string&) {}
struct ServiceServer {};
struct Subscriber {};
struct Publisher {
  template <typename M> void publish(const M&) const {}
};
struct NodeHandle {
  template <typename T> bool param(const std::string&, T& v, const T& d) const { v = d; return false; }
  template <typename Req, typename Res>
  ServiceServer advertiseService(const std::string&, bool (*)(Req&, Res&)) { return ServiceServer(); }
  template <typename M> Subscriber subscribe(const std::string&, unsigned, void (*)(const M&)) { return Subscriber(); }
  template <typename M> Publisher advertise(const std::string&, unsigned) { return Publisher(); }
  bool ok() const { return false; }
};
inline void spin() {}
inline void spinOnce() {}
}  // namespace ros
#define ROS_INFO(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define ROS_FATAL(...) std::fprintf(stderr, __VA_ARGS__)
