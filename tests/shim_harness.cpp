// shim_harness.cpp — the reference's 8-call sequence (hcplanner.cpp:709-716) written against the
// drop-in header fc-planner_b200/shim/viewpoint_manager/viewpoint_manager.h, exactly as the
// planner writes it.  TEST INFRASTRUCTURE: built by tests/test_cpp_shim.py with g++ against the
// stand-in ROS/PCL/Eigen headers in tests/shim_stubs and linked to libfcvm.so.
//
//   shim_harness <params.txt> <input.bin> <output.bin>
// input.bin : int64 n_map, n_model, n_vps; float map[3*n_map], model[3*n_model]; double normals[3*n_model];
//             float vps[6*n_vps]
// output.bin: int64 n_final; {int32 id; double pose[5]; int32 vox} * n_final; int64 n_unc; float unc[3*n_unc]
#include <viewpoint_manager/viewpoint_manager.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  ros::NodeHandle nh;
  {
    std::ifstream f(argv[1]);
    std::string k, v;
    while (f >> k >> v) {
      if (k == "viewpoint_manager/attitude_type") nh.str[k] = v;
      else nh.num[k] = std::atof(v.c_str());
    }
  }
  std::ifstream in(argv[2], std::ios::binary);
  int64_t n[3] = {0, 0, 0};
  in.read((char*)n, sizeof(n));
  pcl::PointCloud<pcl::PointXYZ>::Ptr map(new pcl::PointCloud<pcl::PointXYZ>), model(new pcl::PointCloud<pcl::PointXYZ>);
  pcl::PointCloud<pcl::PointNormal>::Ptr vps(new pcl::PointCloud<pcl::PointNormal>);
  auto read_xyz = [&](pcl::PointCloud<pcl::PointXYZ>& c, int64_t cnt) {
    for (int64_t i = 0; i < cnt; ++i) { pcl::PointXYZ p; in.read((char*)&p.x, 12); c.push_back(p); }
  };
  read_xyz(*map, n[0]);
  read_xyz(*model, n[1]);
  std::map<Eigen::Vector3d, Eigen::Vector3d, predrecon::Vector3dCompare0> normals;
  for (int64_t i = 0; i < n[1]; ++i) {
    double d[3];
    in.read((char*)d, 24);
    const pcl::PointXYZ& p = model->points[(size_t)i];
    normals[Eigen::Vector3d(p.x, p.y, p.z)] = Eigen::Vector3d(d[0], d[1], d[2]);
  }
  for (int64_t i = 0; i < n[2]; ++i) { pcl::PointNormal p; in.read((char*)&p.x, 24); p.curvature = 0; vps->push_back(p); }

  predrecon::ViewpointManager viewpoint_manager_;
  viewpoint_manager_.init(nh);
  // ---- the call sequence of hcplanner.cpp:709-716 ----
  viewpoint_manager_.reset();
  viewpoint_manager_.setMapPointCloud(map);
  viewpoint_manager_.setModel(model);
  viewpoint_manager_.setNormals(normals);
  if (n[2] > 0) viewpoint_manager_.setInitViewpoints(vps);
  viewpoint_manager_.updateViewpoints();
  std::vector<predrecon::SingleViewpoint> out;
  viewpoint_manager_.getUpdatedViewpoints(out);
  pcl::PointCloud<pcl::PointXYZ> unc;
  viewpoint_manager_.getUncoveredModelCloud(unc);

  std::ofstream o(argv[3], std::ios::binary);
  int64_t nf = (int64_t)out.size();
  o.write((char*)&nf, 8);
  for (auto& s : out) {
    int32_t id = s.vp_id, vox = s.vox_count;
    o.write((char*)&id, 4);
    for (int c = 0; c < 5; ++c) { double v = s.pose(c); o.write((char*)&v, 8); }
    o.write((char*)&vox, 4);
  }
  int64_t nu = (int64_t)unc.size();
  o.write((char*)&nu, 8);
  for (auto& p : unc.points) o.write((char*)&p.x, 12);
  std::printf("shim_harness: %lld final viewpoints, %lld uncovered points\n", (long long)nf, (long long)nu);
  return 0;
}
