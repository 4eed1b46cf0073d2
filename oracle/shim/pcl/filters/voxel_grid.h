// stand-in for <pcl/filters/voxel_grid.h>: viewpoint_manager.cpp includes it but uses nothing from it.  TEST INFRASTRUCTURE.
#pragma once
#include <pcl/point_cloud.h>
