// Stand-in for <pcl/point_types.h>: the two point types the ViewpointManager interface uses.
// TEST INFRASTRUCTURE (compile check of fc-planner_b200/shim without PCL installed).
#pragma once
namespace pcl {
struct PointXYZ { float x, y, z; };
struct PointNormal { float x, y, z, normal_x, normal_y, normal_z, curvature; };
}  // namespace pcl
