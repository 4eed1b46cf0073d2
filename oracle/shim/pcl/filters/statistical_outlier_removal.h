// stand-in for <pcl/filters/statistical_outlier_removal.h>: viewpoint_manager.cpp includes it but uses nothing from it.  TEST INFRASTRUCTURE.
#pragma once
#include <pcl/point_cloud.h>
