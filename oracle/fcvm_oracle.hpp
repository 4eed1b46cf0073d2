// fcvm_oracle.hpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A dependency-free (no ROS / PCL / Eigen / FLANN) C++14 restatement of FC-Planner's
// viewpoint visibility + coverage-pruning path, written to follow the reference statement by
// statement so that the CUDA path can be checked against it.  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use it, and only as the
// checker.  The product library (fc-planner_b200/csrc) never includes, links or calls it.
//
// PARITY PIN.  The reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
// section 4) and its own build (catkin + ROS + PCL + Eigen + FLANN) is impossible in this image.
// What pins this file instead is the reference's own SOURCE, compiled by oracle/Makefile from where
// it lies under /root/reference against stand-in third-party headers (oracle/shim/):
//   oracle/_ref/libfcvm_ref.so    raycast.cpp + perception_utils.cpp      (tests/test_oracle_vs_ref.py)
//   oracle/_ref/libfcvm_refvm.so  + the whole viewpoint_manager.cpp       (tests/test_oracle_vs_ref_vm.py)
// The oracle equals both bit for bit: marcher voxel sequences, FOV verdicts, E1-E4 masks, and for
// the full call sequence the selected viewpoints, all poses, counts, cover state, uncovered cloud
// and lookup counts, on synthetic scenes and on the three shipped scenes (= tests/golden/*.npz).
// STILL UNPINNED (third-party internals that exist neither here nor under /root/reference): Eigen's
// SSE2 packet reduction order for 3-vectors (assumed (a0*b0 + a1*b1) + a2*b2),
// FLANN's tie order, pcl::RandomSample's algorithm and the wall-clock seeds (replaced by an
// injected seed on all sides), and sdf_map.cpp (needs PCL/ROS/OpenCV; restated in HCMap below).
//
// All paths are relative to /root/reference/FC-Planner/src/.
//
// Build like the reference (-O3 -std=c++14, no -march, no -ffast-math; CMakeLists.txt:4-8 of
// viewpoint_manager / plan_env) plus -ffp-contract=off so that "no FMA" is explicit.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

// ----------------------------------------------------------------------------------------
// Minimal stand-ins for Eigen::Vector3d / Vector3i with Eigen's evaluation order.
// with SSE2 Eigen 3.3/3.4 reduces a 3-vector as (x0*y0 + x1*y1) + x2*y2
// (one Packet2d + scalar tail).  ORC_ALT_REDUX selects x0*y0 + (x1*y1 + x2*y2) for the
// sensitivity count reported by tests/test_reduction_order.py.
// ----------------------------------------------------------------------------------------
struct V3 {
  double x, y, z;
  V3() : x(0), y(0), z(0) {}
  V3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct V3i {
  int x, y, z;
  V3i() : x(0), y(0), z(0) {}
  V3i(int a, int b, int c) : x(a), y(b), z(c) {}
  int& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  int operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
inline V3 operator/(const V3& a, double s) { return V3(a.x / s, a.y / s, a.z / s); }  // true division
inline double redux3(double a, double b, double c) {
#ifdef ORC_ALT_REDUX
  return a + (b + c);
#else
  return (a + b) + c;
#endif
}
inline double dot(const V3& a, const V3& b) { return redux3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline double squaredNorm(const V3& a) { return dot(a, a); }
inline double norm(const V3& a) { return std::sqrt(squaredNorm(a)); }
// MatrixBase::normalize()/normalized() (Eigen >= 3.3): divide by sqrt(z) only when z > 0
inline V3 normalized(const V3& a) {
  double z = squaredNorm(a);
  if (z > 0.0) return a / std::sqrt(z);
  return a;
}

// ----------------------------------------------------------------------------------------
// Parameters: same layout as fcvm_params in include/fcvm.h (kept separate on purpose: the
// oracle does not include product headers).
// ----------------------------------------------------------------------------------------
struct Params {
  double visible_range, viewpoints_distance, fov_h, fov_w, pitch_upper, pitch_lower;
  double ground_pos, safe_height, safe_radius;
  double top_angle, left_angle, right_angle, max_dist, vis_dist;
  double resolution;
  int32_t z_ground, attitude, max_iter_num, pose_update;
  uint64_t seed;
  int32_t device, host_threads;
};
enum { ATT_OTHER = 0, ATT_YAW = 1, ATT_ALL = 2 };

struct Counters {
  int64_t pair_tests = 0;      // insideFOV calls
  int64_t rays = 0;            // biInput calls (BiRC invocations)
  int64_t lookups = 0;         // occCheck calls
  int64_t walk_lookups = 0;    // occCheck calls inside the surface-offset walk
  int64_t cap_trips = 0;       // marcher step cap fired (reference would not terminate)
  int64_t evals = 0;           // evaluator passes (one viewpoint x all points)
  int64_t safety_lookups = 0;  // safety_check calls
  int64_t steps = 0;           // biNextId / nextId calls
};

// ----------------------------------------------------------------------------------------
// plan_env/src/raycast.cpp:26-43
// ----------------------------------------------------------------------------------------
inline int signum(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }
inline double mod(double value, double modulus) { return std::fmod(std::fmod(value, modulus) + modulus, modulus); }
inline double intbound(double s, double ds) {
  if (ds < 0) {
    return intbound(-s, -ds);
  } else {
    s = mod(s, 1);
    return (1 - s) / ds;
  }
}

// ----------------------------------------------------------------------------------------
// RayCaster: plan_env/include/plan_env/raycast.h:39-106, plan_env/src/raycast.cpp:341-559
// ----------------------------------------------------------------------------------------
class RayCaster {
 public:
  // raycast.cpp:341-345
  void setParams(const double& res, const V3& origin) {
    resolution_ = res;
    half_ = V3(0.5, 0.5, 0.5);
    offset_ = half_ - origin / resolution_;
  }
  // raycast.cpp:347-390
  bool input(const V3& start, const V3& end) {
    start_ = start / resolution_;
    end_ = end / resolution_;
    x_ = (int)std::floor(start_.x);
    y_ = (int)std::floor(start_.y);
    z_ = (int)std::floor(start_.z);
    endX_ = (int)std::floor(end_.x);
    endY_ = (int)std::floor(end_.y);
    endZ_ = (int)std::floor(end_.z);
    dx_ = endX_ - x_;
    dy_ = endY_ - y_;
    dz_ = endZ_ - z_;
    stepX_ = signum((int)dx_);
    stepY_ = signum((int)dy_);
    stepZ_ = signum((int)dz_);
    tMaxX_ = intbound(start_.x, dx_);
    tMaxY_ = intbound(start_.y, dy_);
    tMaxZ_ = intbound(start_.z, dz_);
    tDeltaX_ = ((double)stepX_) / dx_;
    tDeltaY_ = ((double)stepY_) / dy_;
    tDeltaZ_ = ((double)stepZ_) / dz_;
    cap_ = (int64_t)std::abs(endX_ - x_) + std::abs(endY_ - y_) + std::abs(endZ_ - z_) + 8;
    calls_ = 0;
    return !(stepX_ == 0 && stepY_ == 0 && stepZ_ == 0);
  }
  // raycast.cpp:392-444
  bool biInput(const V3& start, const V3& end) {
    start_ = start / resolution_;
    inter_ = (0.5 * (start + end)) / resolution_;
    end_ = end / resolution_;
    x_ = (int)std::floor(start_.x);
    y_ = (int)std::floor(start_.y);
    z_ = (int)std::floor(start_.z);
    interX_ = (int)std::floor(inter_.x);
    interY_ = (int)std::floor(inter_.y);
    interZ_ = (int)std::floor(inter_.z);
    endX_ = (int)std::floor(end_.x);
    endY_ = (int)std::floor(end_.y);
    endZ_ = (int)std::floor(end_.z);
    dx_p = interX_ - x_; dy_p = interY_ - y_; dz_p = interZ_ - z_;
    dx_n = interX_ - endX_; dy_n = interY_ - endY_; dz_n = interZ_ - endZ_;
    stepXp = signum((int)dx_p); stepYp = signum((int)dy_p); stepZp = signum((int)dz_p);
    stepXn = signum((int)dx_n); stepYn = signum((int)dy_n); stepZn = signum((int)dz_n);
    tMaxXp = intbound(start_.x, dx_p);
    tMaxYp = intbound(start_.y, dy_p);
    tMaxZp = intbound(start_.z, dz_p);
    tMaxXn = intbound(end_.x, dx_n);
    tMaxYn = intbound(end_.y, dy_n);
    tMaxZn = intbound(end_.z, dz_n);
    tDeltaXp = ((double)stepXp) / dx_p;
    tDeltaYp = ((double)stepYp) / dy_p;
    tDeltaZp = ((double)stepZp) / dz_p;
    tDeltaXn = ((double)stepXn) / dx_n;
    tDeltaYn = ((double)stepYn) / dy_n;
    tDeltaZn = ((double)stepZn) / dz_n;
    int64_t mp = (int64_t)std::fabs(dx_p) + (int64_t)std::fabs(dy_p) + (int64_t)std::fabs(dz_p);
    int64_t mn = (int64_t)std::fabs(dx_n) + (int64_t)std::fabs(dy_n) + (int64_t)std::fabs(dz_n);
    cap_ = std::max(mp, mn) + 8;
    calls_ = 0;
    return !(stepXp == 0 && stepYp == 0 && stepZp == 0);
  }
  // raycast.cpp:446-479.  The reference has no step cap; `capped` reports the call at which it
  // would have marched past |dx|+|dy|+|dz|+8 voxels.
  bool nextId(V3i& idx) {
    idx = toIndex(x_, y_, z_);
    if (x_ == endX_ && y_ == endY_ && z_ == endZ_) return false;
    if (++calls_ > cap_) { capped = true; return false; }
    if (tMaxX_ < tMaxY_) {
      if (tMaxX_ < tMaxZ_) { x_ += stepX_; tMaxX_ += tDeltaX_; }
      else                 { z_ += stepZ_; tMaxZ_ += tDeltaZ_; }
    } else {
      if (tMaxY_ < tMaxZ_) { y_ += stepY_; tMaxY_ += tDeltaY_; }
      else                 { z_ += stepZ_; tMaxZ_ += tDeltaZ_; }
    }
    return true;
  }
  // raycast.cpp:481-559
  bool biNextId(V3i& idx, V3i& idxR) {
    bool pflag = true, nflag = true;
    idx = toIndex(x_, y_, z_);
    if (x_ == interX_ && y_ == interY_ && z_ == interZ_) pflag = false;
    if (pflag == true) {
      if (tMaxXp < tMaxYp) {
        if (tMaxXp < tMaxZp) { x_ += stepXp; tMaxXp += tDeltaXp; }
        else                 { z_ += stepZp; tMaxZp += tDeltaZp; }
      } else {
        if (tMaxYp < tMaxZp) { y_ += stepYp; tMaxYp += tDeltaYp; }
        else                 { z_ += stepZp; tMaxZp += tDeltaZp; }
      }
    }
    idxR = toIndex(endX_, endY_, endZ_);
    if (endX_ == interX_ && endY_ == interY_ && endZ_ == interZ_) nflag = false;
    if (nflag == true) {
      if (tMaxXn < tMaxYn) {
        if (tMaxXn < tMaxZn) { endX_ += stepXn; tMaxXn += tDeltaXn; }
        else                 { endZ_ += stepZn; tMaxZn += tDeltaZn; }
      } else {
        if (tMaxYn < tMaxZn) { endY_ += stepYn; tMaxYn += tDeltaYn; }
        else                 { endZ_ += stepZn; tMaxZn += tDeltaZn; }
      }
    }
    if (pflag == false && nflag == false) return false;
    if (++calls_ > cap_) { capped = true; return false; }
    return true;
  }
  bool capped = false;

 private:
  // `idx = (tmp + offset_).cast<int>()` (raycast.cpp:447-448, 485-486, 520-521): int -> double,
  // one add, truncation toward zero.
  V3i toIndex(int x, int y, int z) const {
    return V3i((int)((double)x + offset_.x), (int)((double)y + offset_.y), (int)((double)z + offset_.z));
  }
  V3 start_, end_, inter_, half_, offset_;
  int x_ = 0, y_ = 0, z_ = 0, endX_ = 0, endY_ = 0, endZ_ = 0, interX_ = 0, interY_ = 0, interZ_ = 0;
  double dx_ = 0, dy_ = 0, dz_ = 0;
  int stepX_ = 0, stepY_ = 0, stepZ_ = 0;
  double tMaxX_ = 0, tMaxY_ = 0, tMaxZ_ = 0, tDeltaX_ = 0, tDeltaY_ = 0, tDeltaZ_ = 0;
  double dx_p = 0, dy_p = 0, dz_p = 0, dx_n = 0, dy_n = 0, dz_n = 0;
  int stepXp = 0, stepYp = 0, stepZp = 0, stepXn = 0, stepYn = 0, stepZn = 0;
  double tMaxXp = 0, tMaxYp = 0, tMaxZp = 0, tMaxXn = 0, tMaxYn = 0, tMaxZn = 0;
  double tDeltaXp = 0, tDeltaYp = 0, tDeltaZp = 0, tDeltaXn = 0, tDeltaYn = 0, tDeltaZn = 0;
  double resolution_ = 1.0;
  int64_t cap_ = 0, calls_ = 0;
};

// ----------------------------------------------------------------------------------------
// SDFMap, HC subset: plan_env/src/sdf_map.cpp:122-188, plan_env/include/plan_env/sdf_map.h:
// 183-193 (HCMapParam), 217-227 (HCMapData), 234-237, 244-247, 266-273, 367-373, 375-395, 397-427.
// ----------------------------------------------------------------------------------------
class HCMap {
 public:
  double resolution_ = 0, resolution_inv_ = 0, size_inflate = 0;
  V3 map_origin_, map_size_;
  V3i map_voxel_num_, map_origin_idx_;
  std::vector<char> occupancy_buffer_hc_;
  std::vector<char> occupancy_buffer_internal_;  // all zero inside ViewpointManager's private map

  // sdf_map.cpp:122-188 (model = the map cloud; float xyz triples)
  void initHCMap(const Params& p, const float* xyz, int64_t n) {
    resolution_ = p.resolution;
    size_inflate = p.visible_range;
    // pcl::getMinMax3D: float min / max per component (sdf_map.cpp:142-144)
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        float v = xyz[3 * i + c];
        mn[c] = std::min(mn[c], v);
        mx[c] = std::max(mx[c], v);
      }
    double x_size = 4 * size_inflate + mx[0] - mn[0];
    double y_size = 4 * size_inflate + mx[1] - mn[1];
    double z_size = 4 * size_inflate + mx[2] - mn[2];
    resolution_inv_ = 1 / resolution_;
    map_origin_ = V3(mn[0] - 2 * size_inflate, mn[1] - 2 * size_inflate, mn[2] - 2 * size_inflate);
    map_origin_idx_ = V3i(0, 0, 0);          // posToIndex_hc(origin) == (0,0,0), sdf_map.cpp:154
    map_size_ = V3(x_size, y_size, z_size);
    for (int i = 0; i < 3; ++i) map_voxel_num_[i] = (int)std::ceil(map_size_[i] / resolution_);
    int64_t buffer_size = (int64_t)map_voxel_num_.x * map_voxel_num_.y * map_voxel_num_.z;
    occupancy_buffer_hc_.assign((size_t)buffer_size, 0);
    occupancy_buffer_internal_.clear();      // never written on this path; kept implicit (all 0)
    for (int64_t i = 0; i < n; ++i) {
      V3 pt_occ(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      V3i id_occ;
      posToIndex_hc(pt_occ, id_occ);
      int64_t adr = toAddress_hc(id_occ);
      if (adr >= 0 && adr < buffer_size)     // reference: unchecked
        occupancy_buffer_hc_[(size_t)adr] = 1;
    }
  }
  // sdf_map.h:234-237
  void posToIndex_hc(const V3& pos, V3i& id) const {
    for (int i = 0; i < 3; ++i) id[i] = (int)std::floor((pos[i] - map_origin_[i]) * resolution_inv_);
  }
  // sdf_map.h:244-247
  void indexToPos_hc(const V3i& id, V3& pos) const {
    for (int i = 0; i < 3; ++i) pos[i] = (id[i] + 0.5) * resolution_ + map_origin_[i];
  }
  // sdf_map.h:266-273 (int in the reference; widened here so the synthetic scene cannot wrap)
  int64_t toAddress_hc(const V3i& id) const {
    int64_t xi = id.x - map_origin_idx_.x, yi = id.y - map_origin_idx_.y, zi = id.z - map_origin_idx_.z;
    return xi * map_voxel_num_.y * map_voxel_num_.z + yi * map_voxel_num_.z + zi;
  }
  // sdf_map.h:367-373
  bool isInMap_hc(const V3i& idx) const {
    if (idx.x < 0 || idx.y < 0 || idx.z < 0) return false;
    if (idx.x > map_voxel_num_.x - 1 || idx.y > map_voxel_num_.y - 1 || idx.z > map_voxel_num_.z - 1) return false;
    return true;
  }
  // sdf_map.h:375-383
  bool safety_check(const V3i& id) const {
    if (!isInMap_hc(id)) return false;
    if (occupancy_buffer_hc_[(size_t)toAddress_hc(id)] == 1) return false;  // internal buffer is all zero
    return true;
  }
  // sdf_map.h:385-395
  bool safety_check(const V3& pos) const {
    V3i id;
    posToIndex_hc(pos, id);
    return safety_check(id);
  }
  // sdf_map.h:397-420 — true = in map and NOT occupied
  bool occCheck(const V3i& id) const {
    if (!isInMap_hc(id)) return false;
    if (occupancy_buffer_hc_[(size_t)toAddress_hc(id)] == 1) return false;
    return true;
  }
};

// ----------------------------------------------------------------------------------------
// PerceptionUtils: active_perception/src/perception_utils.cpp:29-74 (init), 76-96 (setPose_PY),
// 115-124 (insideFOV).
// ----------------------------------------------------------------------------------------
class PerceptionUtils {
 public:
  void init(const Params& p) {
    top_angle_ = p.top_angle; left_angle_ = p.left_angle; right_angle_ = p.right_angle;
    max_dist_ = p.max_dist;
    n_top_ = V3(0.0, std::sin(M_PI / 2 - top_angle_), std::cos(M_PI / 2 - top_angle_));
    n_bottom_ = V3(0.0, -std::sin(M_PI / 2 - top_angle_), std::cos(M_PI / 2 - top_angle_));
    n_left_ = V3(std::sin(M_PI / 2 - left_angle_), 0.0, std::cos(M_PI / 2 - left_angle_));
    n_right_ = V3(-std::sin(M_PI / 2 - right_angle_), 0.0, std::cos(M_PI / 2 - right_angle_));
    // T_cb_ << 0,-1,0,0, 0,0,1,0, 1,0,0,0, 0,0,0,1;  T_bc_ = T_cb_.inverse() — a signed permutation,
    // so the inverse is its transpose, exactly.
    const double Tcb[4][4] = {{0, -1, 0, 0}, {0, 0, 1, 0}, {1, 0, 0, 0}, {0, 0, 0, 1}};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) T_bc_[i][j] = Tcb[j][i];
  }
  void setPose_PY(const V3& pos, const double& pitch, const double& yaw) {
    pos_ = pos; pitch_ = pitch; yaw_ = yaw;
    const double Ry[3][3] = {{std::cos(yaw_), -std::sin(yaw_), 0.0}, {std::sin(yaw_), std::cos(yaw_), 0.0}, {0.0, 0.0, 1.0}};
    const double Rp[3][3] = {{std::cos(pitch_), 0.0, -std::sin(pitch_)}, {0.0, 1.0, 0.0}, {std::sin(pitch_), 0.0, std::cos(pitch_)}};
    double Twb[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Twb[i][j] = redux3(Ry[i][0] * Rp[0][j], Ry[i][1] * Rp[1][j], Ry[i][2] * Rp[2][j]);
    Twb[0][3] = pos.x; Twb[1][3] = pos.y; Twb[2][3] = pos.z;
    double Rwc[3][3];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        // 4-term inner product; every term but one is an exact zero, so the order is immaterial
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += Twb[i][k] * T_bc_[k][j];
        Rwc[i][j] = s;
      }
    const V3 cam[4] = {n_top_, n_bottom_, n_left_, n_right_};
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < 3; ++i) normals_[q][i] = redux3(Rwc[i][0] * cam[q].x, Rwc[i][1] * cam[q].y, Rwc[i][2] * cam[q].z);
  }
  bool insideFOV(const V3& point) const {
    V3 dir = point - pos_;
    if (norm(dir) > max_dist_) return false;
    dir = normalized(dir);
    for (int q = 0; q < 4; ++q)
      if (dot(dir, normals_[q]) < 0.0) return false;
    return true;
  }
  V3 normals_[4];
  V3 pos_;

 private:
  double pitch_ = 0, yaw_ = 0;
  double left_angle_ = 0, right_angle_ = 0, top_angle_ = 0, max_dist_ = 0;
  V3 n_top_, n_bottom_, n_left_, n_right_;
  double T_bc_[4][4];
};

// ----------------------------------------------------------------------------------------
// Stand-in for pcl::KdTreeFLANN<PointXYZ>: exact search, L2_Simple<float>
// (sequential float accumulation), radius acceptance d2 < r2 with float r2, results sorted by
// (distance, index).  Brute force: the oracle only has to be right.
// ----------------------------------------------------------------------------------------
struct FloatCloudSearch {
  std::vector<float> pts;  // xyz triples
  void setInputCloud(const std::vector<float>& p) { pts = p; }
  static float dist2(const float* a, const float* b) {
    float result = 0.0f, diff;
    for (int i = 0; i < 3; ++i) { diff = a[i] - b[i]; result += diff * diff; }
    return result;
  }
  // returns false when the cloud is empty (PCL throws; reference catches and returns false)
  bool nearestKSearch1(const float* q, int& id, float& d2) const {
    size_t n = pts.size() / 3;
    if (n == 0) return false;
    id = -1; d2 = FLT_MAX;
    for (size_t i = 0; i < n; ++i) {
      float d = dist2(q, &pts[3 * i]);
      if (d < d2) { d2 = d; id = (int)i; }
    }
    return id >= 0;
  }
  void radiusSearch(const float* q, double radius, std::vector<int>& ids, std::vector<float>& d2s) const {
    float r2 = static_cast<float>(radius * radius);
    std::vector<std::pair<float, int>> found;
    size_t n = pts.size() / 3;
    for (size_t i = 0; i < n; ++i) {
      float d = dist2(q, &pts[3 * i]);
      if (d < r2) found.emplace_back(d, (int)i);
    }
    std::sort(found.begin(), found.end());
    ids.clear(); d2s.clear();
    for (auto& f : found) { d2s.push_back(f.first); ids.push_back(f.second); }
  }
};

// ----------------------------------------------------------------------------------------
// The stateful scratch objects of ViewpointManager (raycaster_, percep_utils_; viewpoint_manager.h
// :214-216) plus the two ray idioms the four evaluators share.  One instance lives in the manager;
// the multi-threaded CPU baseline gives each thread its own over the shared read-only map.
// ----------------------------------------------------------------------------------------
struct EvalContext {
  RayCaster raycaster_;
  PerceptionUtils percep_utils_;
  const HCMap* map = nullptr;
  Counters ctr;

  bool occ(const V3i& idx) { ctr.lookups++; return map->occCheck(idx); }

  // BiRC as used at viewpoint_manager.cpp:386-396 / 482-492 / 844-853 (discard_first) and
  // 572-580 (no discard).  SURVEY.md A.5.
  bool biRC(const V3& a, const V3& b, bool discard_first) {
    V3i idx, idx_rev;
    bool visible_flag = true, visible_flag_rev = true;
    ctr.rays++;
    raycaster_.capped = false;
    raycaster_.biInput(a, b);
    if (discard_first) { raycaster_.biNextId(idx, idx_rev); ctr.steps++; }
    while (raycaster_.biNextId(idx, idx_rev)) {
      ctr.steps++;
      visible_flag = occ(idx);
      visible_flag_rev = occ(idx_rev);
      if (visible_flag == false || visible_flag_rev == false) break;
    }
    visible_flag = occ(idx);
    // A capped march has run past its midpoint voxel: the reference would walk on until an occCheck fails (an occupied voxel
    // or the map's edge, where occCheck is false) and so return "not visible" (viewpoint_manager.cpp:388-396).
    if (raycaster_.capped) { ctr.cap_trips++; return false; }
    return visible_flag == true && visible_flag_rev == true;
  }
  // surface-offset walk, viewpoint_manager.cpp:463-477 / 554-568
  V3 offsetTarget(const V3& pos1, const V3& pos2) {
    V3i idx;
    V3 free_pose = pos1;
    raycaster_.capped = false;
    raycaster_.input(pos1, pos2);
    int count = 0;
    while (raycaster_.nextId(idx)) {
      ctr.steps++;
      ctr.walk_lookups++;
      if (occ(idx)) {
        count++;
        if (count == 2) break;
      }
    }
    if (raycaster_.capped) ctr.cap_trips++;
    map->indexToPos_hc(idx, free_pose);
    return free_pose;
  }
  bool insideFOV(const V3& p) { ctr.pair_tests++; return percep_utils_.insideFOV(p); }

  // One evaluator pass for one pose: the visibility sets of getPose pass 1 (E1),
  // pass 2 (E2, restricted to `restrict`), updatePoseGravitation (E3), updateViewpointInfo (E4),
  // without their pose/ownership side effects.  Used by orc_evaluate (parity of masks for
  // arbitrary poses) and by the CPU baseline.
  void evalPass(int stage, const double pose[5], const float* model_xyz, int64_t P, double visible_range,
                const uint32_t* restrict_row, std::vector<int>& visible) {
    visible.clear();
    ctr.evals++;
    V3 vp(pose[0], pose[1], pose[2]);
    percep_utils_.setPose_PY(vp, pose[3], pose[4]);
    for (int64_t i = 0; i < P; ++i) {
      if (stage == 2 && !(restrict_row[i >> 5] >> (i & 31) & 1u)) continue;
      V3 pt(model_xyz[3 * i], model_xyz[3 * i + 1], model_xyz[3 * i + 2]);
      if (!insideFOV(pt)) continue;
      bool vis;
      if (stage == 1) {
        vis = biRC(vp, pt, true) && norm(pt - vp) <= visible_range;
      } else if (stage == 2) {
        vis = biRC(vp, offsetTarget(pt, vp), true);
      } else if (stage == 3) {
        vis = biRC(vp, pt, true);
      } else {
        vis = biRC(vp, offsetTarget(pt, vp), false);
      }
      if (vis) visible.push_back((int)i);
    }
  }
};

struct PointXYZ { float x, y, z; };
struct PointNormal { float x, y, z, normal_x, normal_y, normal_z; };
struct SingleViewpoint { int vp_id; double pose[5]; int vox_count; };

// viewpoint_manager.h:53-63
struct Vector3dCompare0 {
  bool operator()(const V3& v1, const V3& v2) const {
    if (v1.x != v2.x) return v1.x < v2.x;
    if (v1.y != v2.y) return v1.y < v2.y;
    return v1.z < v2.z;
  }
};
// viewpoint_manager.h:65-77
struct VectorHashEigenVector3d {
  std::size_t operator()(const V3& vec) const {
    std::size_t seed = 0;
    for (int i = 0; i < 3; ++i) {
      std::hash<double> hasher;
      seed ^= hasher(vec[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    }
    return seed;
  }
};
struct V3Eq {
  bool operator()(const V3& a, const V3& b) const { return a.x == b.x && a.y == b.y && a.z == b.z; }
};

struct Pose5 { double v[5]; };

// one evaluator event, for mask-level parity
struct TraceEvent {
  int stage;
  int vp_id;
  double pose[5];
  std::vector<int> visible;  // ascending model-point indices
};

// ----------------------------------------------------------------------------------------
// ViewpointManager: viewpoint_manager/src/viewpoint_manager.cpp (all of it)
// ----------------------------------------------------------------------------------------
class ViewpointManager {
 public:
  Counters& ctr() { return ec_.ctr; }
  bool trace_on = false;
  std::vector<TraceEvent> trace;
  std::vector<int> last_prune_order;  // prune_order_ of the most recent viewpointsPrune
  // per-stage wall-clock accumulators for the CPU baseline
  double t_eval_s[5] = {0, 0, 0, 0, 0};

  // viewpoint_manager.cpp:32-62
  void init(const Params& p) {
    prm = p;
    visible_range = p.visible_range;
    dist_vp = p.viewpoints_distance;
    fov_h = p.fov_h; fov_w = p.fov_w;
    pitch_upper = p.pitch_upper; pitch_lower = p.pitch_lower;
    zFlag = p.z_ground != 0;
    GroundZPos = p.ground_pos; safeHeight = p.safe_height; safe_radius_vp = p.safe_radius;
    attitude = p.attitude; max_iter_num = p.max_iter_num; pose_update = p.pose_update != 0;
    if (attitude == ATT_YAW) { pitch_upper = 0.0; pitch_lower = 0.0; }
    fov_base = std::min(fov_h, fov_w) * M_PI / 360.0;
    ec_.percep_utils_.init(p);
    ec_.map = &HCMap_;
    unc_calls_ = 0; sample_calls_ = 0;
  }
  // viewpoint_manager.cpp:64-72
  void reset() {
    TS = TaskState();
    VPI = ViewpointsInfo();
    MCI = ModelCloudInfo();
    // VPP.reset() clear()s the maps (viewpoint_manager.h:169-177): their bucket arrays, hence the
    // iteration order of later insertions, survive a reset
    VPP.inverse_idx_.clear(); VPP.idx_viewpoints_.clear(); VPP.idx_ctrl_voxels_.clear();
    VPP.idx_live_state_.clear(); VPP.idx_query_state_.clear(); VPP.final_vps_.clear();
    map_cloud_kdtree_ = FloatCloudSearch();
    unc_calls_ = 0; sample_calls_ = 0;
  }
  // viewpoint_manager.cpp:74-89
  void setMapPointCloud(const float* xyz, int64_t n) {
    std::vector<float> new_model(xyz, xyz + 3 * n);
    HCMap_.initHCMap(prm, new_model.data(), n);
    map_cloud_kdtree_.setInputCloud(new_model);
    ec_.raycaster_ = RayCaster();
    ec_.raycaster_.setParams(HCMap_.resolution_, HCMap_.map_origin_);
    ec_.map = &HCMap_;
    TS.map_flag_ = true;
  }
  // viewpoint_manager.cpp:91-108
  void setModel(const float* xyz, int64_t n) {
    if (n == 0) return;
    model_.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) model_[(size_t)i] = PointXYZ{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    MCI.cover_state.resize(model_.size(), false);
    MCI.cover_contrib.resize(model_.size(), 0);
    MCI.contrib_id.resize(model_.size(), -1);
    TS.model_flag_ = true;
  }
  // viewpoint_manager.cpp:110-120
  void setNormals(const double* pt, const double* nrm, int64_t n) {
    if (n == 0) return;
    TS.normal_flag_ = true;
    MCI.pt_normal_pairs.clear();
    for (int64_t i = 0; i < n; ++i)
      MCI.pt_normal_pairs.insert(std::make_pair(V3(pt[3 * i], pt[3 * i + 1], pt[3 * i + 2]), V3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2])));
  }
  // viewpoint_manager.cpp:122-141
  void setInitViewpoints(const std::vector<PointNormal>& input_normal_vps) {
    if (input_normal_vps.size() == 0) return;
    // reference: vps_voxcount.resize(V); SURVEY.md A.9 — the later resize(P) is indexed by
    // viewpoint id, which overflows when V > P; here every count array is sized max(V, P).
    VPI.vps_voxcount.resize(input_normal_vps.size(), 0);
    VPI.vps_contri.resize(input_normal_vps.size(), 0);
    VPI.vps_pose.resize(input_normal_vps.size());
    for (int i = 0; i < (int)input_normal_vps.size(); i++) {
      PointNormal nor_vp = input_normal_vps[(size_t)i];
      V3 pos(nor_vp.x, nor_vp.y, nor_vp.z);
      V3 dir(nor_vp.normal_x, nor_vp.normal_y, nor_vp.normal_z);
      VPI.vps_pose[(size_t)i] = getPose(i, pos, dir);
    }
    TS.viewpoints_flag_ = true;
  }
  // viewpoint_manager.cpp:143-330
  void updateViewpoints() {
    if (!TS.viewpoints_flag_) {
      std::vector<PointNormal> vps;
      for (int i = 0; i < (int)model_.size(); i++) {
        PointXYZ pt = model_[(size_t)i], vp;
        V3 pt_vec(pt.x, pt.y, pt.z);
        auto it = MCI.pt_normal_pairs.find(pt_vec);
        if (it == MCI.pt_normal_pairs.end()) continue;  // reference: unchecked dereference (UB)
        V3 normal_dir = it->second;
        vp.x = (float)(pt.x + dist_vp * normal_dir[0]);
        vp.y = (float)(pt.y + dist_vp * normal_dir[1]);
        vp.z = (float)(pt.z + dist_vp * normal_dir[2]);
        if (!viewpointSafetyCheck(vp)) continue;
        PointNormal v;
        v.x = vp.x; v.y = vp.y; v.z = vp.z;
        v.normal_x = (float)(-normal_dir[0]);
        v.normal_y = (float)(-normal_dir[1]);
        v.normal_z = (float)(-normal_dir[2]);
        vps.push_back(v);
      }
      setInitViewpoints(vps);
    }
    if (!TS.isReady()) return;

    std::vector<int> temp_vps_voxcount;
    MCI.cover_state.assign(model_.size(), false);
    MCI.cover_contrib.assign(model_.size(), 0);
    MCI.contrib_id.assign(model_.size(), -1);
    temp_vps_voxcount = VPI.vps_voxcount;
    VPI.vps_voxcount.assign(std::max(model_.size(), VPI.vps_pose.size()), 0);  // reference: resize(P), A.9
    viewpointsPrune(VPI.vps_pose, temp_vps_voxcount);

    std::vector<int> uncoveredID;
    for (int i = 0; i < (int)MCI.cover_state.size(); ++i)
      if (MCI.cover_state[(size_t)i] == false) uncoveredID.push_back(i);

    std::unordered_map<int, std::vector<PointNormal>> uncVpsLib;
    for (int i = 0; i < (int)uncoveredID.size(); ++i) {
      int p = uncoveredID[(size_t)i];
      V3 pt_vec(model_[(size_t)p].x, model_[(size_t)p].y, model_[(size_t)p].z);
      auto it = MCI.pt_normal_pairs.find(pt_vec);
      V3 normal_dir = (it == MCI.pt_normal_pairs.end()) ? V3(0, 0, 1) : it->second;
      float negiNormal[3] = {(float)normal_dir[0], (float)normal_dir[1], (float)normal_dir[2]};
      uncVpsLib[p] = uncNeighVps(model_[(size_t)p], negiNormal);
    }

    int unc_iter = 0, last_unc_num = -1;
    int iter_max = max_iter_num;
    while ((int)uncoveredID.size() > 0 && unc_iter < iter_max && last_unc_num != (int)uncoveredID.size()) {
      last_unc_num = (int)uncoveredID.size();
      std::vector<PointNormal> UncoveredVps;
      for (auto p : uncoveredID)
        for (auto& uvp : uncVpsLib[p]) UncoveredVps.push_back(uvp);
      int candiNum = 100;
      if ((int)UncoveredVps.size() > candiNum) UncoveredVps = randomSample(UncoveredVps, candiNum);

      std::vector<V3> tempVps, tempVpsDir;
      for (int i = 0; i < (int)UncoveredVps.size(); i++) {
        tempVps.push_back(V3(UncoveredVps[(size_t)i].x, UncoveredVps[(size_t)i].y, UncoveredVps[(size_t)i].z));
        tempVpsDir.push_back(V3(UncoveredVps[(size_t)i].normal_x, UncoveredVps[(size_t)i].normal_y, UncoveredVps[(size_t)i].normal_z));
      }
      std::vector<bool> tempCoverState = MCI.cover_state;
      std::vector<int> tempCoverContrib = MCI.cover_contrib;
      std::vector<int> tempContribId = MCI.contrib_id;

      int oriVpNum = (int)VPI.vps_pose.size();
      VPI.last_vps_num = oriVpNum;
      std::vector<Pose5> pose_set_unc(tempVps.size());
      VPI.vps_voxcount.insert(VPI.vps_voxcount.end(), tempVps.size(), 0);
      VPI.vps_contri.insert(VPI.vps_contri.end(), tempVps.size(), 0);
      std::vector<int> before_vps_voxcount = VPI.vps_voxcount;

      for (int i = 0; i < (int)tempVps.size(); ++i) {
        int vp_id = oriVpNum + i;
        pose_set_unc[(size_t)i] = getPose(vp_id, tempVps[(size_t)i], tempVpsDir[(size_t)i]);
      }
      VPI.vps_pose.insert(VPI.vps_pose.end(), pose_set_unc.begin(), pose_set_unc.end());

      MCI.cover_state = tempCoverState;
      MCI.cover_contrib = tempCoverContrib;
      MCI.contrib_id = tempContribId;
      temp_vps_voxcount = VPI.vps_voxcount;
      VPI.vps_voxcount = before_vps_voxcount;
      viewpointsPrune(VPI.vps_pose, temp_vps_voxcount);

      uncoveredID.clear();
      for (int i = 0; i < (int)MCI.cover_state.size(); ++i)
        if (MCI.cover_state[(size_t)i] == false) uncoveredID.push_back(i);
      unc_iter++;
    }
  }
  // viewpoint_manager.cpp:1133-1136
  const std::vector<SingleViewpoint>& getUpdatedViewpoints() const { return VPP.final_vps_; }
  // viewpoint_manager.cpp:1138-1144
  void getUncoveredModelCloud(std::vector<PointXYZ>& cloud) const {
    cloud.clear();
    for (int i = 0; i < (int)MCI.cover_state.size(); ++i)
      if (MCI.cover_state[(size_t)i] == false) cloud.push_back(model_[(size_t)i]);
  }

  // -------------------------------------------------------------------- state (public for taps)
  struct TaskState {
    bool map_flag_ = false, model_flag_ = false, viewpoints_flag_ = false, normal_flag_ = false;
    bool isReady() const { return map_flag_ && model_flag_ && viewpoints_flag_ && normal_flag_; }
  };
  struct ModelCloudInfo {
    std::vector<bool> cover_state;
    std::vector<int> cover_contrib, contrib_id;
    std::map<V3, V3, Vector3dCompare0> pt_normal_pairs;
  };
  struct ViewpointsInfo {
    std::vector<Pose5> vps_pose;
    std::vector<int> vps_voxcount, vps_contri;
    int last_vps_num = 0;
  };
  struct ViewpointsPrune {
    std::unordered_map<V3, int, VectorHashEigenVector3d, V3Eq> inverse_idx_;
    std::unordered_map<int, Pose5> idx_viewpoints_;
    std::unordered_map<int, int> idx_ctrl_voxels_;
    std::unordered_map<int, bool> idx_live_state_;
    std::unordered_map<int, bool> idx_query_state_;
    std::vector<SingleViewpoint> final_vps_;
  };
  Params prm;
  HCMap HCMap_;
  EvalContext ec_;   // raycaster_ + percep_utils_ of the reference
  FloatCloudSearch map_cloud_kdtree_;
  std::vector<PointXYZ> model_;
  ModelCloudInfo MCI;
  ViewpointsInfo VPI;
  ViewpointsPrune VPP;
  TaskState TS;

  bool biRC(const V3& a, const V3& b, bool d) { return ec_.biRC(a, b, d); }
  V3 offsetTarget(const V3& a, const V3& b) { return ec_.offsetTarget(a, b); }

  // viewpoint_manager.cpp:332-531
  Pose5 getPose(int vp_id, const V3& pos, const V3& dir) {
    std::vector<int> index_set;
    V3 viewpoint = pos, visible_candidate;
    V3 fov_sample_ray = normalized(dir);
    V3 fov_yaw_ray = normalized(V3(fov_sample_ray[0], fov_sample_ray[1], 0.0));
    double init_py[2];
    PitchYaw(fov_sample_ray, init_py);

    if (!pose_update) {
      Pose5 pose_res = {{viewpoint[0], viewpoint[1], viewpoint[2], init_py[0], init_py[1]}};
      updateViewpointInfo(pose_res, vp_id);
      return pose_res;
    }
    auto t0 = std::chrono::steady_clock::now();
    ec_.ctr.evals++;
    ec_.percep_utils_.setPose_PY(viewpoint, init_py[0], init_py[1]);
    for (int i = 0; i < (int)model_.size(); ++i) {
      V3 pt_vec(model_[(size_t)i].x, model_[(size_t)i].y, model_[(size_t)i].z);
      if (ec_.insideFOV(pt_vec) == true) index_set.push_back(i);
    }
    V3 ray_dir, ray_dir_yaw;
    double ray_length;
    double avg_pitch = 0.0, avg_yaw = 0.0;
    double max_avg_yaw = -1.0;
    int free_num = 0;
    std::unordered_map<int, V3> free_vox;
    std::vector<int> e1_visible;
    Pose5 pose_result;
    for (int i = 0; i < (int)index_set.size(); ++i) {
      int pi = index_set[(size_t)i];
      visible_candidate = V3(model_[(size_t)pi].x, model_[(size_t)pi].y, model_[(size_t)pi].z);
      if (biRC(viewpoint, visible_candidate, true)) {
        ray_dir = normalized(visible_candidate - viewpoint);
        ray_dir_yaw = visible_candidate - viewpoint;
        ray_dir_yaw[2] = 0.0;   // `ray_dir_yaw.normalized();` at cpp:401 is a discarded no-op (A.9)
        ray_length = norm(visible_candidate - viewpoint);
        if (ray_length <= visible_range) {
          double pitchyaw[2];
          PitchYaw(ray_dir, pitchyaw);
          warpAngle(pitchyaw[0]);
          warpAngle(pitchyaw[1]);
          avg_pitch += pitchyaw[0];
          double yaw_vec = std::min(std::max(dot(fov_sample_ray, ray_dir), -1.0), 1.0);
          double yawoff = std::acos(yaw_vec);
          if (yawoff > max_avg_yaw) max_avg_yaw = yawoff;
          // ray_dir_yaw.cross(fov_yaw_ray)[2]
          if (ray_dir_yaw[0] * fov_yaw_ray[1] - ray_dir_yaw[1] * fov_yaw_ray[0] < 0) yawoff = -yawoff;
          avg_yaw += yawoff;
          free_num++;
          free_vox[pi] = visible_candidate;
          e1_visible.push_back(pi);
        }
      }
    }
    t_eval_s[1] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (trace_on) record(1, vp_id, Pose5{{viewpoint[0], viewpoint[1], viewpoint[2], init_py[0], init_py[1]}}, e1_visible);

    if (free_num == 0) {
      avg_pitch = init_py[0];
      avg_yaw = init_py[1];
      pose_result = Pose5{{viewpoint[0], viewpoint[1], viewpoint[2], avg_pitch, avg_yaw}};
      ec_.percep_utils_.setPose_PY(viewpoint, avg_pitch, avg_yaw);
      return pose_result;
    }
    avg_pitch = avg_pitch / free_num;
    avg_yaw = avg_yaw / free_num;
    avg_yaw = init_py[1] + avg_yaw;
    warpAngle(avg_pitch);
    warpAngle(avg_yaw);
    if (avg_pitch > pitch_upper * M_PI / 180.0) avg_pitch = pitch_upper * M_PI / 180.0;
    if (avg_pitch < pitch_lower * M_PI / 180.0) avg_pitch = pitch_lower * M_PI / 180.0;

    std::map<int, V3> valid_vox;
    int contribute_voxel_num = 0;
    pose_result = Pose5{{viewpoint[0], viewpoint[1], viewpoint[2], avg_pitch, avg_yaw}};
    auto t1 = std::chrono::steady_clock::now();
    ec_.ctr.evals++;
    ec_.percep_utils_.setPose_PY(viewpoint, avg_pitch, avg_yaw);
    for (const auto& id_vox : free_vox) {
      bool inside_flag = ec_.insideFOV(id_vox.second);
      if (inside_flag == true) {
        V3 free_pose = offsetTarget(id_vox.second, viewpoint);
        if (biRC(viewpoint, free_pose, true)) {
          contribute_voxel_num++;
          valid_vox[id_vox.first] = id_vox.second;
        }
      }
    }
    t_eval_s[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
    if (trace_on) {
      std::vector<int> v;
      for (auto& kv : valid_vox) v.push_back(kv.first);
      record(2, vp_id, pose_result, v);
    }
    if (contribute_voxel_num > 0) {
      std::vector<int> ids;
      for (const auto& kv : valid_vox) ids.push_back(kv.first);
      own(ids, vp_id, contribute_voxel_num);
    }
    return pose_result;
  }

  // viewpoint_manager.cpp:533-618
  void updateViewpointInfo(const Pose5& pose, const int& vp_id) {
    auto t0 = std::chrono::steady_clock::now();
    ec_.ctr.evals++;
    int contribute_voxel_num = 0;
    std::vector<int> valid_vox;
    V3 position_(pose.v[0], pose.v[1], pose.v[2]);
    double pitch_ = pose.v[3], yaw_ = pose.v[4];
    ec_.percep_utils_.setPose_PY(position_, pitch_, yaw_);
    for (int i = 0; i < (int)model_.size(); ++i) {
      V3 pt_vec(model_[(size_t)i].x, model_[(size_t)i].y, model_[(size_t)i].z);
      if (ec_.insideFOV(pt_vec)) {
        V3 free_pose = offsetTarget(pt_vec, position_);
        if (biRC(position_, free_pose, false)) {
          contribute_voxel_num++;
          valid_vox.push_back(i);
        }
      }
    }
    t_eval_s[4] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (trace_on) record(4, vp_id, pose, valid_vox);
    if (contribute_voxel_num == 0) return;
    own(valid_vox, vp_id, contribute_voxel_num);
  }

  // ownership rule, viewpoint_manager.cpp:500-528 and 589-617
  void own(const std::vector<int>& valid, int vp_id, int contribute_voxel_num) {
    VPI.vps_contri[(size_t)vp_id] = contribute_voxel_num;
    for (int j : valid) {
      if (MCI.cover_state[(size_t)j] == false) {
        MCI.cover_state[(size_t)j] = true;
        MCI.cover_contrib[(size_t)j] = contribute_voxel_num;
        MCI.contrib_id[(size_t)j] = vp_id;
        VPI.vps_voxcount[(size_t)vp_id] += 1;
      } else if (contribute_voxel_num > MCI.cover_contrib[(size_t)j]) {
        int before_vp_id = MCI.contrib_id[(size_t)j];
        if (before_vp_id >= VPI.last_vps_num) {
          MCI.cover_contrib[(size_t)j] = contribute_voxel_num;
          MCI.contrib_id[(size_t)j] = vp_id;
          VPI.vps_voxcount[(size_t)before_vp_id] -= 1;
          VPI.vps_voxcount[(size_t)vp_id] += 1;
        }
      }
    }
  }

  // viewpoint_manager.cpp:620-710
  void viewpointsPrune(std::vector<Pose5> vps_pose, std::vector<int> vps_voxcount) {
    VPP.inverse_idx_.clear();
    VPP.idx_viewpoints_.clear();
    VPP.idx_ctrl_voxels_.clear();
    VPP.idx_live_state_.clear();
    VPP.idx_query_state_.clear();
    std::vector<int> prune_order_;
    last_prune_order.clear();

    PointXYZ vp_pt_;
    std::vector<float> vp_cloud_;
    for (int i = VPI.last_vps_num; i < (int)vps_pose.size(); ++i) {
      if (vps_voxcount[(size_t)i] > 0) {
        Pose5 qualified_vp_ = vps_pose[(size_t)i];
        int index_ = i;
        vp_pt_.x = (float)qualified_vp_.v[0];
        vp_pt_.y = (float)qualified_vp_.v[1];
        vp_pt_.z = (float)qualified_vp_.v[2];
        V3 tmp_pos(vp_pt_.x, vp_pt_.y, vp_pt_.z);
        vp_cloud_.push_back(vp_pt_.x); vp_cloud_.push_back(vp_pt_.y); vp_cloud_.push_back(vp_pt_.z);
        VPP.inverse_idx_[tmp_pos] = index_;
        VPP.idx_viewpoints_[index_] = qualified_vp_;
        VPP.idx_ctrl_voxels_[index_] = vps_voxcount[(size_t)i];
        VPP.idx_live_state_[index_] = true;
        VPP.idx_query_state_[index_] = false;
      }
    }
    if (vp_cloud_.size() == 0) return;

    std::vector<std::pair<int, int>> ctrlVector(VPP.idx_ctrl_voxels_.begin(), VPP.idx_ctrl_voxels_.end());
    std::sort(ctrlVector.begin(), ctrlVector.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });
    for (const auto& p : ctrlVector) prune_order_.push_back(p.first);
    last_prune_order = prune_order_;

    FloatCloudSearch vps_tree_;
    vps_tree_.setInputCloud(vp_cloud_);

    double bubble_radius = 0.55 * dist_vp * std::tan(fov_base);
    for (auto prune_idx : prune_order_) {
      Pose5& pp = VPP.idx_viewpoints_.find(prune_idx)->second;
      V3 prune_position(pp.v[0], pp.v[1], pp.v[2]);
      bool psv = nearCheck(prune_position, VPP.idx_ctrl_voxels_.find(prune_idx)->second);
      if (psv == false) {
        VPP.idx_live_state_.find(prune_idx)->second = false;
        continue;
      }
      std::vector<int> index_set;
      std::vector<float> radius_set;
      vp_pt_.x = (float)pp.v[0];
      vp_pt_.y = (float)pp.v[1];
      vp_pt_.z = (float)pp.v[2];
      float q[3] = {vp_pt_.x, vp_pt_.y, vp_pt_.z};
      vps_tree_.radiusSearch(q, bubble_radius, index_set, radius_set);
      if (VPP.idx_live_state_.find(prune_idx)->second == true && (int)index_set.size() > 1)
        updatePoseGravitation(vp_cloud_, index_set, prune_idx);
    }
    for (auto& pair : VPP.idx_viewpoints_)
      if (VPP.idx_live_state_.find(pair.first)->second == true) updateViewpointInfo(pair.second, pair.first);

    SingleViewpoint struct_vp;
    for (auto& pair : VPP.idx_viewpoints_) {
      if (VPP.idx_live_state_.find(pair.first)->second == true && VPI.vps_voxcount[(size_t)pair.first] > 0) {
        struct_vp.vp_id = pair.first;
        for (int k = 0; k < 5; ++k) struct_vp.pose[k] = pair.second.v[k];
        struct_vp.vox_count = VPI.vps_voxcount[(size_t)pair.first];
        VPP.final_vps_.push_back(struct_vp);
      }
    }
  }

  // viewpoint_manager.cpp:712-887
  void updatePoseGravitation(const std::vector<float>& cloud, std::vector<int>& inner_ids, const int& cur_idx) {
    VPP.idx_query_state_.find(cur_idx)->second = true;
    double pitch_bound = 0.3 * fov_h * M_PI / 180.0;
    double yaw_bound = 0.3 * fov_w * M_PI / 180.0;

    Pose5 cur_pose = VPP.idx_viewpoints_.find(cur_idx)->second;
    V3 cur_position(cur_pose.v[0], cur_pose.v[1], cur_pose.v[2]);
    double cur_pitch = cur_pose.v[3];
    double cur_yaw = cur_pose.v[4];
    int cur_vox = VPP.idx_ctrl_voxels_.find(cur_idx)->second;

    std::vector<int> updated_indexer;
    for (auto id_ : inner_ids) {
      bool update_flag = false;
      V3 vp_finder(cloud[3 * (size_t)id_], cloud[3 * (size_t)id_ + 1], cloud[3 * (size_t)id_ + 2]);
      int index_finder = VPP.inverse_idx_.find(vp_finder)->second;
      Pose5 inner_pose = VPP.idx_viewpoints_.find(index_finder)->second;
      int inner_vox = VPP.idx_ctrl_voxels_.find(index_finder)->second;
      if (inner_vox > cur_vox) continue;
      double diff_yaw = std::fabs(cur_pose.v[4] - inner_pose.v[4]);
      warpAngle(diff_yaw);
      double diff_pitch = std::fabs(cur_pose.v[3] - inner_pose.v[3]);
      warpAngle(diff_pitch);
      if (std::fabs(diff_pitch) < pitch_bound && std::fabs(diff_yaw) < yaw_bound) update_flag = true;
      if (update_flag == true && VPP.idx_query_state_.find(index_finder)->second == false) {
        VPP.idx_live_state_.find(index_finder)->second = false;
        updated_indexer.push_back(index_finder);
      }
    }
    V3 updated_position = cur_position;
    double updated_pitch = cur_pitch;
    double updated_yaw = cur_yaw;
    for (auto ui_ : updated_indexer) {
      double grav_factor = (double)VPP.idx_ctrl_voxels_.find(ui_)->second / (double)cur_vox;
      const Pose5& up = VPP.idx_viewpoints_.find(ui_)->second;
      V3 up_position(up.v[0], up.v[1], up.v[2]);
      V3 pull_vec = grav_factor * (up_position - cur_position);
      updated_position = updated_position + pull_vec;
      if (attitude == ATT_ALL) {
        double up_pitch = up.v[3];
        double pitch_diff = std::abs(up_pitch - cur_pitch) > M_PI ? (2 * M_PI - std::abs(up_pitch - cur_pitch)) : std::abs(up_pitch - cur_pitch);
        int cal_flag_pitch = up_pitch - cur_pitch > 0 ? 1 : -1;
        int cal_dir_pitch = std::abs(up_pitch - cur_pitch) > M_PI ? -1 : 1;
        double pull_pitch = grav_factor * cal_flag_pitch * cal_dir_pitch * pitch_diff;
        updated_pitch = updated_pitch + pull_pitch;
      }
      double up_yaw = up.v[4];
      double yaw_diff = std::abs(up_yaw - cur_yaw) > M_PI ? (2 * M_PI - std::abs(up_yaw - cur_yaw)) : std::abs(up_yaw - cur_yaw);
      int cal_flag_yaw = up_yaw - cur_yaw > 0 ? 1 : -1;
      int cal_dir_yaw = std::abs(up_yaw - cur_yaw) > M_PI ? -1 : 1;
      double pull_yaw = grav_factor * cal_flag_yaw * cal_dir_yaw * yaw_diff;
      updated_yaw = updated_yaw + pull_yaw;
    }
    if (zFlag == true && updated_position[2] < GroundZPos + safeHeight) updated_position[2] = GroundZPos + safeHeight + 0.1;

    PointXYZ searchPoint;
    searchPoint.x = (float)updated_position[0];
    searchPoint.y = (float)updated_position[1];
    searchPoint.z = (float)updated_position[2];
    Pose5& cur_ref = VPP.idx_viewpoints_.find(cur_idx)->second;
    if (viewpointSafetyCheck(searchPoint) == true) {
      cur_ref.v[0] = updated_position[0]; cur_ref.v[1] = updated_position[1]; cur_ref.v[2] = updated_position[2];
    } else {
      cur_ref.v[0] = cur_position[0]; cur_ref.v[1] = cur_position[1]; cur_ref.v[2] = cur_position[2];
    }
    if (!pose_update) {
      cur_ref.v[3] = updated_pitch;
      cur_ref.v[4] = updated_yaw;
    } else {
      auto t0 = std::chrono::steady_clock::now();
      ec_.ctr.evals++;
      std::vector<int> inner_idx;
      ec_.percep_utils_.setPose_PY(updated_position, updated_pitch, updated_yaw);
      for (int i = 0; i < (int)model_.size(); ++i) {
        V3 pt_vec(model_[(size_t)i].x, model_[(size_t)i].y, model_[(size_t)i].z);
        if (ec_.insideFOV(pt_vec) == true) inner_idx.push_back(i);
      }
      V3 visible_candidate, ray_dir;
      double avg_pitch = updated_pitch, avg_yaw = updated_yaw;
      std::vector<int> e3_visible;
      for (int i = 0; i < (int)inner_idx.size(); ++i) {
        int pi = inner_idx[(size_t)i];
        visible_candidate = V3(model_[(size_t)pi].x, model_[(size_t)pi].y, model_[(size_t)pi].z);
        if (biRC(updated_position, visible_candidate, true)) {
          e3_visible.push_back(pi);
          ray_dir = normalized(visible_candidate - updated_position);
          double pitchyaw[2];
          PitchYaw(ray_dir, pitchyaw);
          warpAngle(pitchyaw[0]);
          warpAngle(pitchyaw[1]);
          double pitch_off = std::abs(pitchyaw[0] - updated_pitch) > M_PI ? (2 * M_PI - std::abs(pitchyaw[0] - updated_pitch)) : std::abs(pitchyaw[0] - updated_pitch);
          int flag_pitch = pitchyaw[0] - updated_pitch > 0 ? 1 : -1;
          int dir_pitch = std::abs(pitchyaw[0] - updated_pitch) > M_PI ? -1 : 1;
          avg_pitch = avg_pitch + flag_pitch * dir_pitch * pitch_off;
          double yaw_off = std::abs(pitchyaw[1] - updated_yaw) > M_PI ? (2 * M_PI - std::abs(pitchyaw[1] - updated_yaw)) : std::abs(pitchyaw[1] - updated_yaw);
          int flag_yaw = pitchyaw[1] - updated_yaw > 0 ? 1 : -1;
          int dir_yaw = std::abs(pitchyaw[1] - updated_yaw) > M_PI ? -1 : 1;
          avg_yaw = avg_yaw + flag_yaw * dir_yaw * yaw_off;
        }
      }
      t_eval_s[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (trace_on) record(3, cur_idx, Pose5{{updated_position[0], updated_position[1], updated_position[2], updated_pitch, updated_yaw}}, e3_visible);
      updated_pitch = avg_pitch;
      updated_yaw = avg_yaw;
      if (updated_pitch > pitch_upper * M_PI / 180.0) updated_pitch = pitch_upper * M_PI / 180.0;
      if (updated_pitch < pitch_lower * M_PI / 180.0) updated_pitch = pitch_lower * M_PI / 180.0;
      warpAngle(updated_pitch);
      warpAngle(updated_yaw);
      cur_ref.v[3] = updated_pitch;
      cur_ref.v[4] = updated_yaw;
    }
  }

  // viewpoint_manager.cpp:889-914.  vp.x etc. are floats mixed into double expressions.
  bool viewpointSafetyCheck(PointXYZ& vp) {
    if (zFlag && vp.z < GroundZPos + safeHeight) return false;
    V3 safe_check_pt;
    int vox_num = 2;
    double res = HCMap_.resolution_;
    int safe_voxel = (int)std::floor(0.5 * safe_radius_vp / res);
    if (safe_voxel <= 0) return false;  // reference: infinite loop; refuse instead
    for (int i = 0; vp.x - safe_radius_vp + i * res < vp.x + safe_radius_vp; i = i + safe_voxel)
      for (int j = 0; vp.y - safe_radius_vp + j * res < vp.y + safe_radius_vp; j = j + safe_voxel)
        for (int k = 0; vp.z - safe_radius_vp + k * res < vp.z + safe_radius_vp; k = k + safe_voxel) {
          safe_check_pt[0] = vp.x - safe_radius_vp + i * res;
          safe_check_pt[1] = vp.y - safe_radius_vp + j * res;
          safe_check_pt[2] = vp.z - safe_radius_vp + k * res;
          ec_.ctr.safety_lookups++;
          if (!HCMap_.safety_check(safe_check_pt)) return false;
        }
    V3 vp_vec(vp.x, vp.y, vp.z);
    if (!nearCheck(vp_vec, vox_num)) return false;
    return true;
  }
  // viewpoint_manager.cpp:916-946
  bool nearCheck(const V3& cur_position, const int& vox_num) {
    if (!(std::isfinite(cur_position[0]) && std::isfinite(cur_position[1]) && std::isfinite(cur_position[2]))) return false;
    bool preserve_ = true;
    float q[3] = {(float)cur_position[0], (float)cur_position[1], (float)cur_position[2]};
    int nid; float nd2;
    if (!map_cloud_kdtree_.nearestKSearch1(q, nid, nd2)) return false;
    double dist = std::sqrt(nd2);   // sqrt(float) promoted: std::sqrt(float) -> float, then to double
    if (dist < safe_radius_vp || vox_num <= 1) preserve_ = false;
    return preserve_;
  }
  // viewpoint_manager.cpp:948-1131.  The reference seeds 8 default_random_engines from the wall
  // clock; here engine k of call c is seeded with seed + 8*c + k (k = 0..3 pitch, 4..7 yaw).
  std::vector<PointNormal> uncNeighVps(PointXYZ oript, const float normal[3]) {
    double max_pitch_range = 40.0, max_yaw_range = 40.0;
    if (attitude == ATT_YAW) max_pitch_range = 0;
    std::vector<PointNormal> unc_vps;
    V3 oriNormal(normal[0], normal[1], normal[2]);
    double oriPY[2];
    PitchYaw(oriNormal, oriPY);
    uint64_t base = prm.seed + 8ull * (uint64_t)unc_calls_++;
    double rand_p[4], rand_y[4];
    std::uniform_real_distribution<double> dist_p(0.0, max_pitch_range * M_PI / 180.0);
    for (int k = 0; k < 4; ++k) {
      std::default_random_engine g;
      g.seed((std::default_random_engine::result_type)(base + (uint64_t)k));
      rand_p[k] = dist_p(g);
    }
    std::uniform_real_distribution<double> dist_y(0.0, max_yaw_range * M_PI / 180.0);
    for (int k = 0; k < 4; ++k) {
      std::default_random_engine g;
      g.seed((std::default_random_engine::result_type)(base + 4ull + (uint64_t)k));
      rand_y[k] = dist_y(g);
    }
    // py_1..py_8 (cpp:984-999): signs of (pitch, yaw) offsets and which random pair they use
    static const int sp[8] = {+1, -1, +1, -1, +1, -1, +1, -1};
    static const int sy[8] = {+1, -1, +1, -1, -1, +1, -1, +1};
    static const int rk[8] = {0, 1, 2, 3, 0, 1, 2, 3};
    for (int q = 0; q < 8; ++q) {
      double py[2];
      py[0] = sp[q] > 0 ? oriPY[0] + rand_p[rk[q]] : oriPY[0] - rand_p[rk[q]];
      py[1] = sy[q] > 0 ? oriPY[1] + rand_y[rk[q]] : oriPY[1] - rand_y[rk[q]];
      V3 vec = pyToVec(py);
      PointNormal vp;
      vp.x = (float)(oript.x + dist_vp * vec[0]);
      vp.y = (float)(oript.y + dist_vp * vec[1]);
      vp.z = (float)(oript.z + dist_vp * vec[2]);
      vp.normal_x = (float)(-vec[0]);
      vp.normal_y = (float)(-vec[1]);
      vp.normal_z = (float)(-vec[2]);
      vps8_[q] = vp;
    }
    for (int q = 0; q < 8; ++q) {
      PointXYZ xyz{vps8_[q].x, vps8_[q].y, vps8_[q].z};
      if (viewpointSafetyCheck(xyz) == true) unc_vps.push_back(vps8_[q]);
    }
    return unc_vps;
  }
  // Stand-in for pcl::RandomSample (cpp:249-255, wall-clock seeded in the reference): selection
  // sampling (Knuth's Algorithm S) driven by minstd_rand0 seeded with seed + 0x9e3779b9 + call#.
  std::vector<PointNormal> randomSample(const std::vector<PointNormal>& in, int sample) {
    std::minstd_rand0 g((std::minstd_rand0::result_type)(prm.seed + 0x9e3779b9ull + (uint64_t)sample_calls_++));
    std::vector<PointNormal> out;
    size_t N = in.size(), n = (size_t)sample, index = 0;
    while (n > 0 && index < in.size()) {
      const float U = (float)((double)g() / 2147483646.0);
      if ((float)N * U <= (float)n) { out.push_back(in[index]); --n; }
      ++index; --N;
    }
    return out;
  }

  // viewpoint_manager.h:238-267
  static void PitchYaw(const V3& vec, double PY[2]) {
    PY[0] = std::asin(vec.z / (norm(vec) + 1e-3));
    PY[1] = std::atan2(vec.y, vec.x);
  }
  static V3 pyToVec(const double py[2]) {
    return V3(std::cos(py[0]) * std::cos(py[1]), std::cos(py[0]) * std::sin(py[1]), std::sin(py[0]));
  }
  static void warpAngle(double& angle) {
    while (angle < -M_PI) angle += 2 * M_PI;
    while (angle > M_PI) angle -= 2 * M_PI;
  }

 private:
  void record(int stage, int vp_id, const Pose5& pose, const std::vector<int>& vis) {
    TraceEvent e;
    e.stage = stage; e.vp_id = vp_id;
    for (int k = 0; k < 5; ++k) e.pose[k] = pose.v[k];
    e.visible = vis;
    std::sort(e.visible.begin(), e.visible.end());
    trace.push_back(std::move(e));
  }
  double visible_range = 0, pitch_upper = 0, pitch_lower = 0, fov_h = 0, fov_w = 0;
  bool zFlag = false;
  double GroundZPos = 0, safeHeight = 0, safe_radius_vp = 0;
  bool pose_update = true;
  double dist_vp = 0, fov_base = 0;
  int attitude = ATT_ALL, max_iter_num = 0;
  int64_t unc_calls_ = 0, sample_calls_ = 0;
  PointNormal vps8_[8];
};

}  // namespace orc
