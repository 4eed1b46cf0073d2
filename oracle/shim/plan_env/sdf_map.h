// Stand-in for <plan_env/sdf_map.h>: the HC subset of predrecon::SDFMap that viewpoint_manager.cpp
// uses (hcmp_->resolution_, hcmp_->map_origin_, initHCMap, occCheck, safety_check, indexToPos_hc),
// with the reference's signatures (sdf_map.h:69-71, 115, 122-125, 132) and the grid of the CPU
// oracle behind it.  TEST INFRASTRUCTURE (oracle/_ref build only).
//
// The reference's own sdf_map.{h,cpp} cannot be compiled here: the header pulls in map_ros.h
// (message_filters, cv_bridge, pcl_conversions, nav_msgs, visualization_msgs, boost) and the
// source carries the whole legacy ESDF / fusion code.  So the occupancy grid in this build is the
// oracle's restatement orc::HCMap (oracle/fcvm_oracle.hpp); what the build pins is
// viewpoint_manager.cpp itself.
#ifndef _SDF_MAP_H
#define _SDF_MAP_H

#include <ros/ros.h>
#include <Eigen/Eigen>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <memory>
#include <vector>

#include "../../fcvm_oracle.hpp"

using namespace std;

namespace predrecon {

struct HCMapParam {
  double resolution_, resolution_inv_, size_inflate;
  Eigen::Vector3d map_origin_, map_size_;
  Eigen::Vector3i map_voxel_num_;
};

class SDFMap {
 public:
  unique_ptr<HCMapParam> hcmp_;

  void initHCMap(ros::NodeHandle& nh, pcl::PointCloud<pcl::PointXYZ>::Ptr& model) {
    hcmp_.reset(new HCMapParam);
    orc::Params p = orc::Params();
    nh.param("hcmap/resolution", p.resolution, -1.0);
    nh.param("viewpoint_manager/visible_range", p.visible_range, -1.0);
    std::vector<float> xyz;
    xyz.reserve(3 * model->points.size());
    for (const auto& q : model->points) { xyz.push_back(q.x); xyz.push_back(q.y); xyz.push_back(q.z); }
    map_.initHCMap(p, xyz.data(), (int64_t)model->points.size());
    hcmp_->resolution_ = map_.resolution_;
    hcmp_->resolution_inv_ = map_.resolution_inv_;
    hcmp_->size_inflate = map_.size_inflate;
    hcmp_->map_origin_ = Eigen::Vector3d(map_.map_origin_.x, map_.map_origin_.y, map_.map_origin_.z);
    hcmp_->map_size_ = Eigen::Vector3d(map_.map_size_.x, map_.map_size_.y, map_.map_size_.z);
    hcmp_->map_voxel_num_ = Eigen::Vector3i(map_.map_voxel_num_.x, map_.map_voxel_num_.y, map_.map_voxel_num_.z);
  }
  void indexToPos_hc(const Eigen::Vector3i& id, Eigen::Vector3d& pos) {
    orc::V3 p;
    map_.indexToPos_hc(orc::V3i(id(0), id(1), id(2)), p);
    pos = Eigen::Vector3d(p.x, p.y, p.z);
  }
  bool occCheck(Eigen::Vector3i& id) { ++occ_calls; return map_.occCheck(orc::V3i(id(0), id(1), id(2))); }
  bool safety_check(Eigen::Vector3d& pos) { ++safety_calls; return map_.safety_check(orc::V3(pos(0), pos(1), pos(2))); }
  void publishMap() {}

  long long occ_calls = 0, safety_calls = 0;
  orc::HCMap map_;
};

}  // namespace predrecon
#endif
