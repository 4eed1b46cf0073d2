// viewpoint_manager/viewpoint_manager.h — drop-in replacement header for FC-Planner's
// `predrecon::ViewpointManager`, forwarding every public method to the C ABI of libfcvm.so
// (include/fcvm.h).  Put this directory ahead of the reference's
// viewpoint_manager/include on the include path and link libfcvm.so instead of
// libviewpoint_manager.so: the three call sites (hcplanner.cpp:47, 709-716;
// skeleton_viewpoint.cpp:44, 489-496; viewpoint_manager_node.cpp:130-136, 239) compile unchanged.
//
// Public interface mirrored (reference: viewpoint_manager/include/viewpoint_manager/viewpoint_manager.h):
//   :53-63  Vector3dCompare0      :65-77  VectorHashEigenVector3d      :79-84  SingleViewpoint
//   :180-193 ViewpointManager(), ~ViewpointManager(), init, reset, setModel, setMapPointCloud,
//            setNormals, setInitViewpoints, updateViewpoints, getUpdatedViewpoints,
//            getUncoveredModelCloud
// Error convention as the reference: void methods, ROS_ERROR + return on bad input; nothing throws.
// A missing CUDA device is reported the same way — the library has no CPU fallback.
//
// Multi-GPU switch: the launch parameter `fcvm/gpus` (default 1).  With N > 1 the manager holds N library handles, one per
// CUDA device fcvm/device .. fcvm/device + N - 1, joined by the library's own NCCL communicator (fcvm_set_shard_nccl), and
// every public call is made on all of them at once, one host thread per handle: each GPU evaluates its block of every
// viewpoint batch, all end with the same (bit-identical) state, results are read from the first.  The planner's code does
// not change.
#ifndef _VIEWPOINT_MANAGER_H_
#define _VIEWPOINT_MANAGER_H_

#include <ros/ros.h>
#include <Eigen/Eigen>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "fcvm.h"

#ifndef ROS_ERROR
#include <cstdio>
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#endif

namespace predrecon
{
  struct Vector3dCompare0
  {
    bool operator()(const Eigen::Vector3d &a, const Eigen::Vector3d &b) const
    {
      for (int k = 0; k < 2; ++k)
        if (a(k) != b(k))
          return a(k) < b(k);
      return a(2) < b(2);
    }
  };

  struct VectorHashEigenVector3d
  {
    std::size_t operator()(const Eigen::Vector3d &v) const
    {
      std::size_t seed = 0;
      for (int i = 0; i < 3; ++i)
        seed ^= std::hash<double>()(v(i)) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
      return seed;
    }
  };

  struct SingleViewpoint
  {
    int vp_id;
    Eigen::VectorXd pose;   // (x, y, z, pitch, yaw)
    int vox_count;
  };

  class ViewpointManager
  {
  public:
    ViewpointManager() : h_(nullptr) {}
    ~ViewpointManager() { destroy_all(); }
    ViewpointManager(const ViewpointManager &) = delete;
    ViewpointManager &operator=(const ViewpointManager &) = delete;

    // viewpoint_manager.cpp:32-62 + perception_utils.cpp:32-37 + sdf_map.cpp:128-136: the same
    // parameter names and defaults, read once here and handed to the library as a plain struct.
    void init(ros::NodeHandle &nh)
    {
      fcvm_params p = fcvm_params();
      std::string attitude;
      bool zflag = false, pose_update = true;
      nh.param("viewpoint_manager/visible_range", p.visible_range, -1.0);
      nh.param("viewpoint_manager/viewpoints_distance", p.viewpoints_distance, -1.0);
      nh.param("viewpoint_manager/fov_h", p.fov_h, -1.0);
      nh.param("viewpoint_manager/fov_w", p.fov_w, -1.0);
      nh.param("viewpoint_manager/pitch_upper", p.pitch_upper, -1.0);
      nh.param("viewpoint_manager/pitch_lower", p.pitch_lower, -1.0);
      nh.param("viewpoint_manager/zGround", zflag, false);
      nh.param("viewpoint_manager/GroundPos", p.ground_pos, -1.0);
      nh.param("viewpoint_manager/safeHeight", p.safe_height, -1.0);
      nh.param("viewpoint_manager/safe_radius", p.safe_radius, -1.0);
      nh.param("viewpoint_manager/attitude_type", attitude, std::string("null"));
      nh.param("viewpoint_manager/max_iter_num", p.max_iter_num, 0);            // reference default (cpp:46)
      nh.param("viewpoint_manager/pose_update", pose_update, true);         // reference default (cpp:47)
      nh.param("perception_utils/top_angle", p.top_angle, -1.0);
      nh.param("perception_utils/left_angle", p.left_angle, -1.0);
      nh.param("perception_utils/right_angle", p.right_angle, -1.0);
      nh.param("perception_utils/max_dist", p.max_dist, -1.0);
      nh.param("perception_utils/vis_dist", p.vis_dist, -1.0);
      nh.param("hcmap/resolution", p.resolution, -1.0);
      int seed = 0, device = 0, threads = 0, gpus = 1;
      nh.param("fcvm/seed", seed, 0);             // replaces the reference's wall-clock seeds (cpp:962-977, 249-255)
      nh.param("fcvm/device", device, 0);
      nh.param("fcvm/host_threads", threads, 0);
      nh.param("fcvm/gpus", gpus, 1);
      p.z_ground = zflag ? 1 : 0;
      p.pose_update = pose_update ? 1 : 0;
      p.attitude = attitude == "yaw" ? FCVM_ATTITUDE_YAW : (attitude == "all" ? FCVM_ATTITUDE_ALL : FCVM_ATTITUDE_OTHER);
      p.seed = (uint64_t)seed;
      p.host_threads = threads;
      destroy_all();
      if (gpus < 1) gpus = 1;
      if (gpus > 1 && threads <= 0)               // the handles of one process share its cores
        p.host_threads = (int)(std::thread::hardware_concurrency() / (unsigned)gpus > 0 ? std::thread::hardware_concurrency() / (unsigned)gpus : 1);
      for (int k = 0; k < gpus; ++k)
      {
        p.device = device + k;
        fcvm_t *h = fcvm_create(&p);
        if (!h) { ROS_ERROR("[ViewpointManager] %s", fcvm_last_error()); destroy_all(); return; }
        hs_.push_back(h);
      }
      h_ = hs_[0];
      if (gpus > 1)
      {
        uint8_t id[128];
        if (fcvm_nccl_unique_id(id) != FCVM_OK) { ROS_ERROR("[ViewpointManager] %s", fcvm_last_error()); destroy_all(); return; }
        const int n = gpus;
        if (all([&](int k, fcvm_t *h) { return fcvm_set_shard_nccl(h, id, k, n); }) != FCVM_OK) destroy_all();
      }
    }

    void reset() { all([](int, fcvm_t *h) { return fcvm_reset(h); }); }

    void setModel(pcl::PointCloud<pcl::PointXYZ>::Ptr input_model)
    {
      std::vector<float> xyz;
      pack(*input_model, xyz);
      all([&](int, fcvm_t *h) { return fcvm_set_model(h, xyz.data(), (int64_t)(xyz.size() / 3)); });
    }

    void setMapPointCloud(pcl::PointCloud<pcl::PointXYZ>::Ptr input_map)
    {
      std::vector<float> xyz;
      pack(*input_map, xyz);
      all([&](int, fcvm_t *h) { return fcvm_set_map_cloud(h, xyz.data(), (int64_t)(xyz.size() / 3)); });
    }

    void setNormals(std::map<Eigen::Vector3d, Eigen::Vector3d, Vector3dCompare0> &pt_normal_pairs)
    {
      std::vector<double> k, v;
      k.reserve(3 * pt_normal_pairs.size());
      v.reserve(3 * pt_normal_pairs.size());
      for (const auto &kv : pt_normal_pairs)
        for (int i = 0; i < 3; ++i)
        {
          k.push_back(kv.first(i));
          v.push_back(kv.second(i));
        }
      all([&](int, fcvm_t *h) { return fcvm_set_normals(h, k.data(), v.data(), (int64_t)pt_normal_pairs.size()); });
    }

    void setInitViewpoints(pcl::PointCloud<pcl::PointNormal>::Ptr input_normal_vps)
    {
      std::vector<float> rows;
      rows.reserve(6 * input_normal_vps->points.size());
      for (const auto &q : input_normal_vps->points)
      {
        const float r[6] = {q.x, q.y, q.z, q.normal_x, q.normal_y, q.normal_z};
        rows.insert(rows.end(), r, r + 6);
      }
      all([&](int, fcvm_t *h) { return fcvm_set_init_viewpoints(h, rows.data(), (int64_t)(rows.size() / 6)); });
    }

    void updateViewpoints() { all([](int, fcvm_t *h) { return fcvm_update(h); }); }

    void getUpdatedViewpoints(std::vector<SingleViewpoint> &viewpoints)
    {
      const int64_t n = h_ ? fcvm_get_viewpoints(h_, nullptr, nullptr, nullptr, 0) : 0;
      std::vector<int32_t> id((size_t)(n > 0 ? n : 0)), vox(id.size());
      std::vector<double> pose(5 * id.size());
      if (n > 0) fcvm_get_viewpoints(h_, id.data(), pose.data(), vox.data(), n);
      viewpoints.clear();        // `viewpoints = VPP.final_vps_` (viewpoint_manager.cpp:1135)
      for (size_t i = 0; i < id.size(); ++i)
      {
        SingleViewpoint s;
        s.vp_id = id[i];
        s.pose.resize(5);
        for (int c = 0; c < 5; ++c) s.pose(c) = pose[5 * i + c];
        s.vox_count = vox[i];
        viewpoints.push_back(s);
      }
    }

    void getUncoveredModelCloud(pcl::PointCloud<pcl::PointXYZ> &cloud)
    {
      const int64_t n = h_ ? fcvm_get_uncovered(h_, nullptr, 0) : 0;
      std::vector<float> xyz(3 * (size_t)(n > 0 ? n : 0));
      if (n > 0) fcvm_get_uncovered(h_, xyz.data(), n);
      cloud.clear();                                   // viewpoint_manager.cpp:1140
      for (size_t i = 0; i + 2 < xyz.size(); i += 3)
      {
        pcl::PointXYZ q;
        q.x = xyz[i]; q.y = xyz[i + 1]; q.z = xyz[i + 2];
        cloud.push_back(q);
      }
    }

    fcvm_t *handle() { return h_; }   // evaluator-level entry points and parity taps of fcvm.h

  private:
    static void pack(const pcl::PointCloud<pcl::PointXYZ> &c, std::vector<float> &xyz)
    {
      xyz.clear();
      xyz.reserve(3 * c.points.size());
      for (const auto &q : c.points)
      {
        xyz.push_back(q.x); xyz.push_back(q.y); xyz.push_back(q.z);
      }
    }
    // f(rank, handle) on every handle at once (one thread per extra handle: the sharded calls are collective); the first
    // failure is logged like the reference logs bad input, and its code returned
    template <class F>
    int all(F f)
    {
      if (hs_.empty()) { ROS_ERROR("[ViewpointManager] init() was not called (code %d)", FCVM_ENOTREADY); return FCVM_ENOTREADY; }
      std::vector<int> rc(hs_.size(), FCVM_OK);
      std::vector<std::string> msg(hs_.size());
      auto run = [&](size_t k) { rc[k] = f((int)k, hs_[k]); if (rc[k] != FCVM_OK) msg[k] = fcvm_last_error(); };
      std::vector<std::thread> th;
      for (size_t k = 1; k < hs_.size(); ++k) th.emplace_back(run, k);
      run(0);
      for (auto &t : th) t.join();
      for (size_t k = 0; k < hs_.size(); ++k)
        if (rc[k] != FCVM_OK) { ROS_ERROR("[ViewpointManager] %s (code %d)", msg[k].c_str(), rc[k]); return rc[k]; }
      return FCVM_OK;
    }
    void destroy_all()
    {
      for (fcvm_t *h : hs_) fcvm_destroy(h);
      hs_.clear();
      h_ = nullptr;
    }
    fcvm_t *h_;                    // hs_[0]: where results are read from
    std::vector<fcvm_t *> hs_;     // one handle per GPU
  };
} // namespace predrecon

#endif
