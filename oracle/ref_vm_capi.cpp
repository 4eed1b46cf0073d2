// ref_vm_capi.cpp — C entry points around the REFERENCE'S OWN viewpoint_manager.cpp (together with
// its raycast.cpp and perception_utils.cpp), compiled from where they lie under /root/reference
// against the stand-in headers in shim/ (Eigen, PCL point cloud / kd-tree / RandomSample, ROS
// NodeHandle, plan_env/sdf_map.h).  TEST INFRASTRUCTURE: built into oracle/_ref/libfcvm_refvm.so,
// loaded only by tests/test_oracle_vs_ref_vm.py to validate the oracle restatement end to end.
//
// From the reference: every statement of ViewpointManager (init, reset, setMapPointCloud, setModel,
// setNormals, setInitViewpoints, updateViewpoints, getPose, updateViewpointInfo, viewpointsPrune,
// updatePoseGravitation, viewpointSafetyCheck, nearCheck, uncNeighVps, the getters), RayCaster and
// PerceptionUtils, running on the real libstdc++ containers (unordered_map order, std::sort) and
// the real glibc libm.  Not from the reference: Eigen, the PCL containers/search, RandomSample,
// the occupancy grid (see shim/plan_env/sdf_map.h), and the two wall-clock seeds, replaced by the
// injected-seed scheme (shim/clock_shim.h, shim/pcl/filters/random_sample.h).
#include <viewpoint_manager/viewpoint_manager.h>

#include <cstdint>

namespace {
void set_params(const orc::Params& p) {
  auto& t = ros::shim_param_table();
  t["viewpoint_manager/visible_range"] = p.visible_range;
  t["viewpoint_manager/viewpoints_distance"] = p.viewpoints_distance;
  t["viewpoint_manager/fov_h"] = p.fov_h;
  t["viewpoint_manager/fov_w"] = p.fov_w;
  t["viewpoint_manager/pitch_upper"] = p.pitch_upper;
  t["viewpoint_manager/pitch_lower"] = p.pitch_lower;
  t["viewpoint_manager/zGround"] = p.z_ground;
  t["viewpoint_manager/GroundPos"] = p.ground_pos;
  t["viewpoint_manager/safeHeight"] = p.safe_height;
  t["viewpoint_manager/safe_radius"] = p.safe_radius;
  t["viewpoint_manager/max_iter_num"] = p.max_iter_num;
  t["viewpoint_manager/pose_update"] = p.pose_update;
  t["perception_utils/top_angle"] = p.top_angle;
  t["perception_utils/left_angle"] = p.left_angle;
  t["perception_utils/right_angle"] = p.right_angle;
  t["perception_utils/max_dist"] = p.max_dist;
  t["perception_utils/vis_dist"] = p.vis_dist;
  t["hcmap/resolution"] = p.resolution;
  ros::shim_string_table()["viewpoint_manager/attitude_type"] = p.attitude == orc::ATT_YAW ? "yaw" : (p.attitude == orc::ATT_ALL ? "all" : "null");
}
}  // namespace

struct refvm_handle {
  predrecon::ViewpointManager vm;
  orc::Params prm;
};

extern "C" {

refvm_handle* refvm_create(const orc::Params* p) {
  refvm_handle* h = new refvm_handle();
  h->prm = *p;
  set_params(*p);
  ros::NodeHandle nh;
  h->vm.init(nh);
  std::chrono::system_clock::base() = p->seed;     // (renamed to the shim clock by clock_shim.h)
  std::chrono::system_clock::calls() = 0;
  pcl::RandomSampleShimState::seed() = p->seed;
  pcl::RandomSampleShimState::calls() = 0;
  return h;
}
void refvm_destroy(refvm_handle* h) { delete h; }
// reset() of the reference + the seed counters (the oracle's reset() rewinds its call counters too)
int refvm_reset(refvm_handle* h) {
  h->vm.reset();
  std::chrono::system_clock::calls() = 0;
  pcl::RandomSampleShimState::calls() = 0;
  return 0;
}
int refvm_set_map_cloud(refvm_handle* h, const float* xyz, int64_t n) {
  set_params(h->prm);
  pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>);
  for (int64_t i = 0; i < n; ++i) { pcl::PointXYZ p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; c->push_back(p); }
  h->vm.setMapPointCloud(c);
  return 0;
}
int refvm_set_model(refvm_handle* h, const float* xyz, int64_t n) {
  pcl::PointCloud<pcl::PointXYZ>::Ptr c(new pcl::PointCloud<pcl::PointXYZ>);
  for (int64_t i = 0; i < n; ++i) { pcl::PointXYZ p; p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; c->push_back(p); }
  h->vm.setModel(c);
  return 0;
}
int refvm_set_normals(refvm_handle* h, const double* pt, const double* nrm, int64_t n) {
  std::map<Eigen::Vector3d, Eigen::Vector3d, predrecon::Vector3dCompare0> m;
  for (int64_t i = 0; i < n; ++i)
    m.insert(std::make_pair(Eigen::Vector3d(pt[3 * i], pt[3 * i + 1], pt[3 * i + 2]), Eigen::Vector3d(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2])));
  h->vm.setNormals(m);
  return 0;
}
int refvm_set_init_viewpoints(refvm_handle* h, const float* pos_dir, int64_t n) {
  pcl::PointCloud<pcl::PointNormal>::Ptr c(new pcl::PointCloud<pcl::PointNormal>);
  for (int64_t i = 0; i < n; ++i) {
    pcl::PointNormal p;
    p.x = pos_dir[6 * i]; p.y = pos_dir[6 * i + 1]; p.z = pos_dir[6 * i + 2];
    p.normal_x = pos_dir[6 * i + 3]; p.normal_y = pos_dir[6 * i + 4]; p.normal_z = pos_dir[6 * i + 5];
    c->push_back(p);
  }
  h->vm.setInitViewpoints(c);
  return 0;
}
int refvm_update(refvm_handle* h) { h->vm.updateViewpoints(); return 0; }
int64_t refvm_get_viewpoints(refvm_handle* h, int32_t* vp_id, double* pose5, int32_t* vox, int64_t cap) {
  std::vector<predrecon::SingleViewpoint> v;
  h->vm.getUpdatedViewpoints(v);
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) {
    if (vp_id) vp_id[i] = v[(size_t)i].vp_id;
    if (pose5) for (int c = 0; c < 5; ++c) pose5[5 * i + c] = v[(size_t)i].pose(c);
    if (vox) vox[i] = v[(size_t)i].vox_count;
  }
  return (int64_t)v.size();
}
int64_t refvm_get_uncovered(refvm_handle* h, float* xyz, int64_t cap) {
  pcl::PointCloud<pcl::PointXYZ> c;
  h->vm.getUncoveredModelCloud(c);
  for (int64_t i = 0; i < (int64_t)c.size() && i < cap; ++i) {
    xyz[3 * i] = c.points[(size_t)i].x; xyz[3 * i + 1] = c.points[(size_t)i].y; xyz[3 * i + 2] = c.points[(size_t)i].z;
  }
  return (int64_t)c.size();
}
// private state (the class is compiled with `private` opened in this TU only; layout is unaffected)
int64_t refvm_get_vps_pose(refvm_handle* h, double* pose5, int64_t cap_rows) {
  const Eigen::MatrixXd& m = h->vm.VPI.vps_pose;
  for (int64_t i = 0; i < m.rows() && i < cap_rows; ++i)
    for (int c = 0; c < 5; ++c) pose5[5 * i + c] = m((int)i, c);
  return m.rows();
}
int64_t refvm_get_vps_voxcount(refvm_handle* h, int32_t* out, int64_t cap) {
  const auto& v = h->vm.VPI.vps_voxcount;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}
int64_t refvm_get_vps_contri(refvm_handle* h, int32_t* out, int64_t cap) {
  const auto& v = h->vm.VPI.vps_contri;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}
int64_t refvm_get_cover_state(refvm_handle* h, uint8_t* covered, int32_t* contrib, int32_t* id, int64_t cap) {
  const auto& m = h->vm.MCI;
  for (int64_t i = 0; i < (int64_t)m.cover_state.size() && i < cap; ++i) {
    if (covered) covered[i] = m.cover_state[(size_t)i] ? 1 : 0;
    if (contrib) contrib[i] = m.cover_contrib[(size_t)i];
    if (id) id[i] = m.contrib_id[(size_t)i];
  }
  return (int64_t)m.cover_state.size();
}
int64_t refvm_lookups(refvm_handle* h, int64_t* safety) {
  if (safety) *safety = h->vm.HCMap ? h->vm.HCMap->safety_calls : 0;
  return h->vm.HCMap ? h->vm.HCMap->occ_calls : 0;
}

}  // extern "C"
