// hooks_harness.cpp — the planner-side loops of SURVEY.md 8(f) written against fc-planner_b200/shim/
// hierarchical_coverage_planner/fcvm_planner_hooks.h with the reference's types.  TEST INFRASTRUCTURE: built by
// tests/test_cpp_hooks.py with g++ against the stand-in ROS/PCL/Eigen headers in tests/shim_stubs, linked to libfcvm.so.
//
//   hooks_harness <input.bin> <output.bin>
// input.bin : double fov4[top,left,right,max_dist]; double cam6[fx,fy,cx,cy,fov_h,fov_w]; int64 n_cloud, n_pose, n_rosa, n_pos;
//             float cloud[3*n_cloud]; double poses[5*n_pose]; double origin[3], res; int32 dims[3]; int8 hc[G], inflate[G];
//             float rosa[3*n_rosa]; double pos[3*n_pos]
// output.bin: double rate; int64 group sizes[n_pose]; float group points...; double path_cov; uint8 det[n_pos]; uint8 box[n_pos];
//             uint8 safe[n_pos^2]; double dist[n_pos^2]
#include <hierarchical_coverage_planner/fcvm_planner_hooks.h>

#include <cstring>
#include <fstream>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  double fov4[4], cam6[6];
  int64_t n[4];
  in.read((char*)fov4, sizeof(fov4)); in.read((char*)cam6, sizeof(cam6)); in.read((char*)n, sizeof(n));
  fcvm_params prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.top_angle = fov4[0]; prm.left_angle = fov4[1]; prm.right_angle = fov4[2]; prm.max_dist = fov4[3];
  fcvm_camera cam = {cam6[0], cam6[1], cam6[2], cam6[3], cam6[4], cam6[5]};
  pcl::PointCloud<pcl::PointXYZ>::Ptr Fullmodel(new pcl::PointCloud<pcl::PointXYZ>);
  for (int64_t i = 0; i < n[0]; ++i) { pcl::PointXYZ p; in.read((char*)&p.x, 12); Fullmodel->push_back(p); }
  std::vector<Eigen::VectorXd> TrajPose((size_t)n[1]);
  for (auto& v : TrajPose) { v.resize(5); for (int c = 0; c < 5; ++c) in.read((char*)&v(c), 8); }
  double org[3], res; int32_t dims[3];
  in.read((char*)org, 24); in.read((char*)&res, 8); in.read((char*)dims, 12);
  const size_t G = (size_t)dims[0] * dims[1] * dims[2];
  std::vector<char> hc(G), inflate(G), internal;
  in.read(hc.data(), (std::streamsize)G); in.read(inflate.data(), (std::streamsize)G);
  pcl::PointCloud<pcl::PointXYZ> rosaCloud;
  for (int64_t i = 0; i < n[2]; ++i) { pcl::PointXYZ p; in.read((char*)&p.x, 12); rosaCloud.push_back(p); }
  std::vector<Eigen::Vector3d> pos((size_t)n[3]);
  pcl::PointCloud<pcl::PointNormal> seg_vps;
  for (auto& v : pos) {
    in.read((char*)&v(0), 8); in.read((char*)&v(1), 8); in.read((char*)&v(2), 8);
    pcl::PointNormal q; q.x = (float)v(0); q.y = (float)v(1); q.z = (float)v(2); q.normal_x = q.normal_y = q.normal_z = 0; q.curvature = 0;
    seg_vps.push_back(q);
  }

  std::ofstream o(argv[2], std::ios::binary);
  // ---- hcopp.cpp:108: double coverageRate = PathPlanner_->CoverageEvaluation(TrajGen_->TrajPose); ----
  predrecon::fcvm_hooks::CoverageEvaluator ce;
  std::vector<pcl::PointCloud<pcl::PointXYZ>::Ptr> visible_group;
  double coverageRate = 0.0, coverage = 0.0;
  std::vector<bool> CoverState;
  if (ce.init(prm, cam) && ce.setCloud(Fullmodel)) {
    coverageRate = ce.CoverageEvaluation(TrajPose, visible_group);
    coverage = ce.PathCoverage(TrajPose, CoverState);               // hcplanner.cpp:250-270
  }
  visible_group.resize(TrajPose.size());
  o.write((char*)&coverageRate, 8);
  for (auto& g : visible_group) { int64_t k = g ? (int64_t)g->points.size() : 0; o.write((char*)&k, 8); }
  for (auto& g : visible_group) if (g) for (auto& p : g->points) o.write((char*)&p.x, 12);
  o.write((char*)&coverage, 8);

  // ---- the planner's own map: lightStateDet, safety box, search_Path straight lines ----
  predrecon::fcvm_hooks::PlannerGrid pg;
  std::vector<bool> det, box;
  std::vector<uint8_t> safe; std::vector<double> dist;
  const int vn[3] = {dims[0], dims[1], dims[2]};
  if (pg.init(0, Eigen::Vector3d(org[0], org[1], org[2]), res, vn, hc, inflate, internal)) {
    pg.lightStateDet(pos, rosaCloud, det);
    pg.safetyBox(seg_vps, 1.0, box);
    pg.straightLines(pos, safe, dist);
  }
  det.resize(pos.size()); box.resize(pos.size()); safe.resize(pos.size() * pos.size()); dist.resize(pos.size() * pos.size());
  for (bool b : det) { char c = b; o.write(&c, 1); }
  for (bool b : box) { char c = b; o.write(&c, 1); }
  o.write((char*)safe.data(), (std::streamsize)safe.size());
  o.write((char*)dist.data(), (std::streamsize)(8 * dist.size()));
  return 0;
}
