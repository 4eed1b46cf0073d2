// Stand-in for <pcl/point_cloud.h>.  TEST INFRASTRUCTURE.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
template <class T>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;   // boost::shared_ptr in PCL 1.10
  std::vector<T> points;
  void push_back(const T& p) { points.push_back(p); }
  void clear() { points.clear(); }
  std::size_t size() const { return points.size(); }
};
}  // namespace pcl
