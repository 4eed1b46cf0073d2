// oracle/shim/ros/ros.h — stand-in for the roscpp facilities the reference's hot-path sources use:
// NodeHandle::param(name, var, default) and the ROS_INFO / ROS_WARN / ROS_ERROR macros.
// TEST INFRASTRUCTURE (oracle/_ref build only).  Values come from a process-wide table filled by
// the ref_*capi.cpp wrappers (the ROS parameter server's role).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

namespace ros {

inline std::map<std::string, double>& shim_param_table() {
  static std::map<std::string, double> t;
  return t;
}
inline std::map<std::string, std::string>& shim_string_table() {
  static std::map<std::string, std::string> t;
  return t;
}

class NodeHandle {
 public:
  template <class T>
  bool param(const std::string& name, T& var, const T& def) const {
    auto it = shim_param_table().find(name);
    if (it == shim_param_table().end()) {
      var = def;
      return false;
    }
    var = static_cast<T>(it->second);
    return true;
  }
  bool param(const std::string& name, std::string& var, const std::string& def) const {
    auto it = shim_string_table().find(name);
    if (it == shim_string_table().end()) {
      var = def;
      return false;
    }
    var = it->second;
    return true;
  }
};

}  // namespace ros

#ifndef ROS_INFO
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { if (std::getenv("FCVM_REF_VERBOSE")) { std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#endif
