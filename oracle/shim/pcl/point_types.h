// Stand-in for <pcl/point_types.h>: float point types with PCL's field names.  TEST INFRASTRUCTURE.
#pragma once
namespace pcl {
struct PointXYZ { float x = 0, y = 0, z = 0; };
struct Normal { float normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
struct PointNormal { float x = 0, y = 0, z = 0, normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
}  // namespace pcl
