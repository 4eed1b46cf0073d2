// oracle/shim/ros/ros.h — stand-in for the one roscpp facility the reference's perception_utils.cpp
// uses: NodeHandle::param(name, var, default).  TEST INFRASTRUCTURE (oracle/_ref build only).
// Values come from a process-wide table filled by ref_capi.cpp (the ROS parameter server's role).
#pragma once

#include <map>
#include <string>

namespace ros {

inline std::map<std::string, double>& shim_param_table() {
  static std::map<std::string, double> t;
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
};

}  // namespace ros
