// Stand-in for <pcl/point_cloud.h>.  TEST INFRASTRUCTURE.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

#include <pcl/point_types.h>
namespace pcl {
template <class T>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<T>> Ptr;        // boost::shared_ptr in PCL 1.10
  typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
  std::vector<T> points;
  void push_back(const T& p) { points.push_back(p); }
  void clear() { points.clear(); }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
};
}  // namespace pcl
