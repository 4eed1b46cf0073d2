// Stand-in for <ros/ros.h>: NodeHandle::param over an in-process table.  TEST INFRASTRUCTURE.
#pragma once
#include <map>
#include <string>
namespace ros {
class NodeHandle {
 public:
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;
  template <class T>
  bool param(const std::string& name, T& var, const T& def) const {
    auto it = num.find(name);
    if (it == num.end()) { var = def; return false; }
    var = static_cast<T>(it->second);
    return true;
  }
  bool param(const std::string& name, std::string& var, const std::string& def) const {
    auto it = str.find(name);
    if (it == str.end()) { var = def; return false; }
    var = it->second;
    return true;
  }
};
}  // namespace ros
