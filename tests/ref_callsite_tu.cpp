// ref_callsite_tu.cpp — TEST INFRASTRUCTURE.  Compiles the reference's OWN call blocks of predrecon::ViewpointManager,
// byte for byte, against the drop-in header (fc-planner_b200/shim/viewpoint_manager/viewpoint_manager.h).  The three blocks
// are not stored in this repository: tests/test_cpp_callsites.py cuts them out of /root/reference at test time into
// tests/_build/callsite_*.inc (git-ignored), and this file only supplies the variables each block finds in scope at its
// original location (their declarations are restated from the reference headers cited below).
//   callsite_hcplanner.inc   hierarchical_coverage_planner/src/hcplanner.cpp:708-716          (scope: hcplanner.h:206, PR = PlanResults)
//   callsite_skeleton.inc    viewpoint_manager/src/skeleton_viewpoint.cpp:488-496             (same names)
//   callsite_node.inc        viewpoint_manager/src/viewpoint_manager_node.cpp:128-136         (scope: viewpoint_manager_node.h)
#include <viewpoint_manager/viewpoint_manager.h>

#include <cstdio>
#include <map>
#include <memory>
#include <vector>

using namespace std;
using namespace predrecon;

struct PlanResults {                                   // the members the blocks touch (hcplanner.h: struct PlannerResults)
  pcl::PointCloud<pcl::PointXYZ>::Ptr occ_model, model;
  map<Eigen::Vector3d, Eigen::Vector3d, Vector3dCompare0> pt_normal_pairs;
};

static PlanResults PR;
static shared_ptr<ViewpointManager> viewpoint_manager_;
static pcl::PointCloud<pcl::PointNormal>::Ptr all_safe_normal_vps;
static pcl::PointCloud<pcl::PointXYZ> uncovered_area;

static size_t block_hcplanner() {
#include "callsite_hcplanner.inc"
  return updated_vps.size();
}
static size_t block_skeleton() {
#include "callsite_skeleton.inc"
  return updated_vps.size();
}
// viewpoint_manager_node.cpp keeps its clouds in file-scope variables of these names
static pcl::PointCloud<pcl::PointXYZ>::Ptr model_, downsampled_model_;
static pcl::PointCloud<pcl::PointNormal>::Ptr vps;
static map<Eigen::Vector3d, Eigen::Vector3d, Vector3dCompare0> tmp_pt_normal_pairs;
static size_t block_node() {
#include "callsite_node.inc"
  return updated_vps.size();
}

int main() {
  ros::NodeHandle nh;
  viewpoint_manager_.reset(new ViewpointManager);
  viewpoint_manager_->init(nh);
  PR.occ_model.reset(new pcl::PointCloud<pcl::PointXYZ>);
  PR.model.reset(new pcl::PointCloud<pcl::PointXYZ>);
  all_safe_normal_vps.reset(new pcl::PointCloud<pcl::PointNormal>);
  model_ = PR.occ_model; downsampled_model_ = PR.model; vps = all_safe_normal_vps;
  // empty inputs: every call logs an error and returns (the reference's convention), nothing is evaluated
  const size_t n = block_hcplanner() + block_skeleton() + block_node();
  std::printf("callsites ran: %zu viewpoints\n", n);
  return 0;
}
