// stand-in for <pcl/common/common.h>: viewpoint_manager.cpp includes it but uses nothing from it.  TEST INFRASTRUCTURE.
#pragma once
#include <pcl/point_cloud.h>
