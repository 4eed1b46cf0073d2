// Stand-in for pcl::KdTreeFLANN (exact search over FLANN's L2_Simple<float>): brute force with the
// documented semantics of SURVEY.md A.8 — squared distance accumulated sequentially in float;
// nearestKSearch(k = 1) = first minimum; radiusSearch accepts d2 < (float)(r*r) and returns
// (distance, index)-sorted results; an empty tree throws like PCL.  TEST INFRASTRUCTURE.
// FLANN / PCL sources are not available in this image: tie order is the stated assumption.
#pragma once
#include <algorithm>
#include <cfloat>
#include <stdexcept>
#include <utility>
#include <vector>

#include <pcl/point_cloud.h>
namespace pcl {
template <class T>
class KdTreeFLANN {
 public:
  void setInputCloud(const typename PointCloud<T>::Ptr& c) { cloud_ = c; }
  int nearestKSearch(const T& q, int k, std::vector<int>& idx, std::vector<float>& d2) const {
    if (!cloud_ || cloud_->points.empty()) throw std::runtime_error("KdTreeFLANN: empty input cloud");
    (void)k;   // only k = 1 occurs (viewpoint_manager.cpp:923)
    int best = -1;
    float bd = FLT_MAX;
    for (std::size_t i = 0; i < cloud_->points.size(); ++i) {
      const float d = dist2(q, cloud_->points[i]);
      if (d < bd) { bd = d; best = (int)i; }
    }
    idx.assign(1, best);
    d2.assign(1, bd);
    return 1;
  }
  int radiusSearch(const T& q, double radius, std::vector<int>& idx, std::vector<float>& d2) const {
    idx.clear();
    d2.clear();
    if (!cloud_) return 0;
    const float r2 = static_cast<float>(radius * radius);
    std::vector<std::pair<float, int>> found;
    for (std::size_t i = 0; i < cloud_->points.size(); ++i) {
      const float d = dist2(q, cloud_->points[i]);
      if (d < r2) found.emplace_back(d, (int)i);
    }
    std::sort(found.begin(), found.end());
    for (auto& f : found) { d2.push_back(f.first); idx.push_back(f.second); }
    return (int)idx.size();
  }

 private:
  static float dist2(const T& a, const T& b) {
    float result = 0.0f, diff;
    diff = a.x - b.x; result += diff * diff;
    diff = a.y - b.y; result += diff * diff;
    diff = a.z - b.z; result += diff * diff;
    return result;
  }
  typename PointCloud<T>::Ptr cloud_;
};
}  // namespace pcl
