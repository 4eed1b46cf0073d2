// fcvm_planner_hooks.h — reference-side binding of the widened rows: drop-in bodies for three
// member functions / loops of `hierarchical_coverage_planner` and `HCSolver`, written with the reference's own types
// (pcl::PointCloud, Eigen::VectorXd, Eigen::Vector3d) and forwarding to the C ABI (include/fcvm.h).
//
//   predrecon::fcvm_hooks::CoverageEvaluator   holds the fcvm_cov_t of (camera, Fullmodel)
//     .CoverageEvaluation(poseSet, visible_group)        body of hcplanner.cpp:1032-1126
//     .PathCoverage(path, CoverState)                    the PinHoleCamera loop of hcplanner.cpp:253-261
//   predrecon::fcvm_hooks::PlannerGrid         holds the fcvm_grid_t of the planner's SDFMap buffers
//     .lightStateDet(candidates, rosaCloud, det)         hcplanner.cpp:805-863 for a batch
//     .safetyBox(viewpoints, safe_radius, ok)            hcplanner.cpp:658-676 for a batch
//     .straightLines(positions, safe, dist)              search_Path's straight-line test, hcsolver.cpp:556-583, all pairs
//
// Errors follow the reference's convention: logged (ROS_ERROR), never thrown; outputs stay empty.
#pragma once

#include <cstdint>
#include <vector>

#include <Eigen/Eigen>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <ros/ros.h>

#include "fcvm.h"

#ifndef ROS_ERROR
#include <cstdio>
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#endif

namespace predrecon {
namespace fcvm_hooks {

class CoverageEvaluator {
 public:
  CoverageEvaluator() = default;
  CoverageEvaluator(const CoverageEvaluator&) = delete;
  CoverageEvaluator& operator=(const CoverageEvaluator&) = delete;
  ~CoverageEvaluator() { fcvm_cov_destroy(h_); }

  // params: perception_utils/{top,left,right}_angle, max_dist (+ device); camera: hcplanner/{fx,fy,cx,cy}, fov_h, fov_w
  bool init(const fcvm_params& params, const fcvm_camera& camera) {
    fcvm_cov_destroy(h_);
    h_ = fcvm_cov_create(&params, &camera);
    return ok(h_ ? 0 : -1);
  }
  // `*Fullmodel = *PR.occ_model` (hcplanner.cpp:252): the cloud every pose is checked against
  bool setCloud(const pcl::PointCloud<pcl::PointXYZ>::Ptr& Fullmodel) {
    if (!h_ || !Fullmodel) return false;
    cloud_ = Fullmodel;
    std::vector<float> xyz;
    xyz.reserve(3 * Fullmodel->points.size());
    for (const auto& p : Fullmodel->points) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
    return ok(fcvm_cov_set_cloud(h_, xyz.data(), (int64_t)Fullmodel->points.size()));
  }
  // hcplanner.cpp:1032-1126: returns the coverage rate, appends one cloud per pose to visible_group
  double CoverageEvaluation(std::vector<Eigen::VectorXd>& poseSet, std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr>& visible_group) {
    if (!h_ || !cloud_) return 0.0;
    std::vector<double> p5 = pack(poseSet);
    std::vector<int64_t> offs(poseSet.size() + 1, 0);
    double rate = 0.0;
    const int64_t total = fcvm_cov_evaluate(h_, p5.data(), (int64_t)poseSet.size(), &rate, offs.data(), nullptr, 0, nullptr);
    if (!ok(total < 0 ? (int)total : 0)) return 0.0;
    std::vector<int32_t> ids((size_t)total);
    fcvm_cov_fetch_ids(h_, ids.data(), total);
    for (size_t k = 0; k < poseSet.size(); ++k) {
      pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>);
      for (int64_t e = offs[k]; e < offs[k + 1]; ++e) c->points.push_back(cloud_->points[(size_t)ids[(size_t)e]]);
      visible_group.push_back(c);
    }
    return rate;
  }
  // hcplanner.cpp:253-270: CoverState[x] = true for every id any pose of the path covers; returns the coverage
  double PathCoverage(std::vector<Eigen::VectorXd>& path, std::vector<bool>& CoverState) {
    if (!h_ || !cloud_) return 0.0;
    std::vector<double> p5 = pack(path);
    std::vector<uint8_t> covered(cloud_->points.size(), 0);
    const int64_t rc = fcvm_cov_pinhole(h_, p5.data(), (int64_t)path.size(), nullptr, nullptr, 0, covered.data());
    if (!ok(rc < 0 ? (int)rc : 0)) return 0.0;
    CoverState.assign(covered.size(), false);
    int numTrue = 0;
    for (size_t s = 0; s < covered.size(); ++s)
      if (covered[s]) { CoverState[s] = true; numTrue++; }
    return (double)numTrue / (double)covered.size();
  }

 private:
  static std::vector<double> pack(const std::vector<Eigen::VectorXd>& poses) {
    std::vector<double> p5;
    p5.reserve(5 * poses.size());
    for (const auto& v : poses)
      for (int c = 0; c < 5; ++c) p5.push_back(v(c));
    return p5;
  }
  static bool ok(int rc) {
    if (rc < 0) { ROS_ERROR("[CoverageAnalyzer] %s", fcvm_last_error()); return false; }
    return true;
  }
  fcvm_cov_t* h_ = nullptr;
  pcl::PointCloud<pcl::PointXYZ>::Ptr cloud_;
};

class PlannerGrid {
 public:
  PlannerGrid() = default;
  PlannerGrid(const PlannerGrid&) = delete;
  PlannerGrid& operator=(const PlannerGrid&) = delete;
  ~PlannerGrid() { fcvm_grid_destroy(h_); }

  // the planner's SDFMap after InternalSpace / OuterCheck: HCMap->hcmp_->{map_origin_, resolution_, map_voxel_num_} and
  // HCMap->hcmd_->{occupancy_buffer_hc_, occupancy_inflate_buffer_hc_, occupancy_buffer_internal_} (vector<char>)
  bool init(int device, const Eigen::Vector3d& map_origin, double resolution, const int voxel_num[3], const std::vector<char>& occupancy_hc,
            const std::vector<char>& occupancy_inflate, const std::vector<char>& occupancy_internal) {
    fcvm_grid_destroy(h_);
    const double org[3] = {map_origin(0), map_origin(1), map_origin(2)};
    const int32_t dims[3] = {voxel_num[0], voxel_num[1], voxel_num[2]};
    auto ptr = [](const std::vector<char>& v) { return v.empty() ? nullptr : reinterpret_cast<const int8_t*>(v.data()); };
    h_ = fcvm_grid_create(device, org, resolution, dims, ptr(occupancy_hc), ptr(occupancy_inflate), ptr(occupancy_internal));
    return ok(h_ ? 0 : -1);
  }
  // hcplanner.cpp:805-863 for a batch of candidates; det[i] = true: outside
  bool lightStateDet(const std::vector<Eigen::Vector3d>& candidates, const pcl::PointCloud<pcl::PointXYZ>& rosaCloud, std::vector<bool>& det) {
    det.clear();
    if (!h_ || candidates.empty()) return false;
    std::vector<double> c;
    for (const auto& v : candidates) { c.push_back(v(0)); c.push_back(v(1)); c.push_back(v(2)); }
    std::vector<float> r;
    for (const auto& p : rosaCloud.points) { r.push_back(p.x); r.push_back(p.y); r.push_back(p.z); }
    std::vector<uint8_t> d(candidates.size(), 0);
    if (!ok(fcvm_grid_light_state(h_, c.data(), (int64_t)candidates.size(), r.data(), (int64_t)rosaCloud.points.size(), d.data(), nullptr, nullptr)))
      return false;
    det.assign(d.begin(), d.end());
    return true;
  }
  // hcplanner.cpp:658-676 for a batch of sampled viewpoints
  bool safetyBox(const pcl::PointCloud<pcl::PointNormal>& viewpoints, double safe_radius_vp, std::vector<bool>& safe) {
    safe.clear();
    if (!h_ || viewpoints.points.empty()) return false;
    std::vector<float> v;
    for (const auto& p : viewpoints.points) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); }
    std::vector<uint8_t> s(viewpoints.points.size(), 0);
    if (!ok(fcvm_grid_safety_box(h_, v.data(), (int64_t)viewpoints.points.size(), safe_radius_vp, s.data()))) return false;
    safe.assign(s.begin(), s.end());
    return true;
  }
  // the straight-line part of search_Path (hcsolver.cpp:556-583) for every ordered pair: safe[i * n + j], dist[i * n + j]
  bool straightLines(const std::vector<Eigen::Vector3d>& positions, std::vector<uint8_t>& safe, std::vector<double>& dist) {
    const size_t n = positions.size();
    safe.assign(n * n, 0);
    dist.assign(n * n, 0.0);
    if (!h_ || n == 0) return false;
    std::vector<double> p;
    for (const auto& v : positions) { p.push_back(v(0)); p.push_back(v(1)); p.push_back(v(2)); }
    return ok(fcvm_grid_line_of_sight(h_, p.data(), (int64_t)n, safe.data(), dist.data()));
  }

 private:
  static bool ok(int rc) {
    if (rc < 0) { ROS_ERROR("[Planner] %s", fcvm_last_error()); return false; }
    return true;
  }
  fcvm_grid_t* h_ = nullptr;
};

}  // namespace fcvm_hooks
}  // namespace predrecon
