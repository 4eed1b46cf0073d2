// Stand-in for pcl::RandomSample.  PCL seeds it from time() (wall clock) and its exact algorithm is
// not available in this image, so — exactly like the oracle and the CUDA library — the sample is
// drawn by selection sampling (Knuth's Algorithm S) driven by minstd_rand0 seeded with
// seed + 0x9e3779b9 + call#, where `seed` is the injected seed (ref_vm_capi.cpp sets it).
// TEST INFRASTRUCTURE.
#pragma once
#include <cstdint>
#include <random>
#include <vector>

#include <pcl/point_cloud.h>
namespace pcl {
struct RandomSampleShimState {
  static uint64_t& seed() { static uint64_t s = 0; return s; }
  static uint64_t& calls() { static uint64_t c = 0; return c; }
};
template <class T>
class RandomSample {
 public:
  void setInputCloud(const typename PointCloud<T>::Ptr& c) { in_ = c; }
  void setSample(unsigned n) { sample_ = n; }
  void filter(PointCloud<T>& out) {
    std::minstd_rand0 g((std::minstd_rand0::result_type)(RandomSampleShimState::seed() + 0x9e3779b9ull + RandomSampleShimState::calls()++));
    const std::vector<T> in = in_->points;   // copy: `out` may alias the input (viewpoint_manager.cpp:252-254)
    std::vector<T> res;
    std::size_t N = in.size(), n = sample_, index = 0;
    while (n > 0 && index < in.size()) {
      const float U = (float)((double)g() / 2147483646.0);
      if ((float)N * U <= (float)n) { res.push_back(in[index]); --n; }
      ++index; --N;
    }
    out.points = res;
  }

 private:
  typename PointCloud<T>::Ptr in_;
  unsigned sample_ = 0;
};
}  // namespace pcl
