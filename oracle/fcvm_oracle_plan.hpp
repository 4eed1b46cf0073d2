// fcvm_oracle_plan.hpp — CPU ORACLE of three ray / probe loops of the planner that run on the planner's own SDFMap
// (SURVEY.md 8(f) ranks 2 and 4).  TEST INFRASTRUCTURE, NOT PRODUCT CODE (rules of fcvm_oracle.hpp).
//
// Restated, statement by statement, from FC-Planner/src/hierarchical_coverage_planner/src/:
//   hcplanner.cpp:805-863   lightStateDet: BiRC from a candidate viewpoint to its nearest ROSA vertex, parity of the
//                           number of occupied voxels crossed (odd = the candidate is outside the structure)
//   hcplanner.cpp:658-692   the safety box around a candidate (lattice of SDFMap::safety_check probes, stride 2)
//   hcsolver.cpp:556-583    search_Path's straight-line test: one-directional march p1 -> p2 against the inflated and
//                           internal buffers; on success the cost is the Euclidean distance
// and from FC-Planner/src/plan_env/src/sdf_map.cpp:
//   :263-395   SDFMap::InternalSpace: per skeleton segment, mark the segment cloud's voxels (hc + inflate, with the
//              (2 inflateVoxel + 1)^3 dilation of the inflate buffer), then cast rays from the projection points on the
//              segment to the cloud points of each cutting plane - and from the nearest projection point to every point no
//              plane caught - marking the voxels crossed in occupancy_buffer_internal_
//   :440-455   SDFMap::SetOcc;   :457-473   SDFMap::OuterCheck (clears internal marks along outward rays)
//   :475-491   SDFMap::points_in_plane
// (sdf_map.cpp itself cannot be compiled here: its header pulls in map_ros.h -> message_filters / cv_bridge / OpenCV.)
// The marchers are RayCaster of fcvm_oracle.hpp (pinned to the reference's raycast.cpp, tests/test_oracle_vs_ref.py);
// the three call-site loops are additionally pinned by driving the reference's own RayCaster object with them
// (oracle/ref_capi.cpp: ref_light_state, ref_line_of_sight).  The kd-tree query of lightStateDet is an exact k = 1
// search with FLANN's L2_Simple<float>; the lowest index wins ties (FLANN's tie order is not pinned).
#pragma once

#include "fcvm_oracle.hpp"

namespace orc {

// The planner's SDFMap as these loops see it: geometry + three char buffers (sdf_map.h:217-227), x-major (sdf_map.h:266-269)
struct PlanGrid {
  HCMap map;                       // geometry, occupancy_buffer_hc_, occupancy_buffer_internal_
  std::vector<char> inflate;       // occupancy_inflate_buffer_hc_

  void set(const double origin[3], double res, const int dims[3], const int8_t* hc, const int8_t* inf, const int8_t* internal) {
    map.resolution_ = res;
    map.resolution_inv_ = 1 / res;
    map.map_origin_ = V3(origin[0], origin[1], origin[2]);
    map.map_origin_idx_ = V3i(0, 0, 0);
    map.map_voxel_num_ = V3i(dims[0], dims[1], dims[2]);
    const size_t G = (size_t)dims[0] * dims[1] * dims[2];
    map.occupancy_buffer_hc_.assign(G, 0);
    map.occupancy_buffer_internal_.assign(G, 0);
    inflate.assign(G, 0);
    for (size_t i = 0; i < G; ++i) {
      if (hc) map.occupancy_buffer_hc_[i] = (char)hc[i];
      if (internal) map.occupancy_buffer_internal_[i] = (char)internal[i];
      if (inf) inflate[i] = (char)inf[i];
    }
  }
  // SDFMap::safety_check (sdf_map.h:375-395) with the internal buffer in play
  bool safety_check(const V3& pos) const {
    V3i id;
    map.posToIndex_hc(pos, id);
    if (!map.isInMap_hc(id)) return false;
    const size_t a = (size_t)map.toAddress_hc(id);
    if (map.occupancy_buffer_hc_[a] == 1 || map.occupancy_buffer_internal_[a] == 1) return false;
    return true;
  }
};

struct PlanCounters {
  int64_t lookups = 0;      // occCheck / buffer reads / safety_check probes
  int64_t cap_trips = 0;    // marches the reference would not have terminated
};

class PlanOps {
 public:
  PlanGrid g;
  RayCaster raycaster_;
  PlanCounters ctr;

  void setGrid(const double origin[3], double res, const int dims[3], const int8_t* hc, const int8_t* inf, const int8_t* internal) {
    g.set(origin, res, dims, hc, inf, internal);
    raycaster_.setParams(res, g.map.map_origin_);          // hcplanner.cpp:89, hcsolver.cpp:40
  }

  // hcplanner.cpp:805-863 with K = 1.  rosa: the ROSA vertex cloud (float).  Returns det; crossCount and the vertex id out.
  bool lightStateDet(const V3& vpCandi, const V3& vpCandiNormal, const FloatCloudSearch& ROSATree, int& crossCount, int& nearest) {
    const V3 vpTest = vpCandi + (0.0 * g.map.resolution_) * vpCandiNormal;                     // :809
    const float q[3] = {(float)vpCandi.x, (float)vpCandi.y, (float)vpCandi.z};                 // pcl::PointXYZ searchVP
    float d2;
    crossCount = 0;
    nearest = -1;
    if (!ROSATree.nearestKSearch1(q, nearest, d2)) return false;
    const V3 nearestROSAVertex(ROSATree.pts[3 * (size_t)nearest], ROSATree.pts[3 * (size_t)nearest + 1], ROSATree.pts[3 * (size_t)nearest + 2]);
    V3i idxFwd, idxRev;
    bool fwdFlag, revFlag;
    raycaster_.capped = false;
    raycaster_.biInput(vpTest, nearestROSAVertex);
    raycaster_.biNextId(idxFwd, idxRev);
    while (raycaster_.biNextId(idxFwd, idxRev)) {
      fwdFlag = occ(idxFwd);
      revFlag = occ(idxRev);
      if (fwdFlag == false) crossCount++;
      if (revFlag == false) crossCount++;
    }
    fwdFlag = occ(idxFwd);
    if (fwdFlag == false) crossCount++;
    if (raycaster_.capped) ctr.cap_trips++;
    return crossCount % 2 != 0;                                                                  // :846-849, detNum == K
  }

  // hcplanner.cpp:658-676: note that `break` only leaves the k loop, so safe_flag ends up as the verdict of the LAST
  // (i, j) column that was walked - restated literally.
  bool safetyBox(const PointNormal& vp, double safe_radius_vp) {
    const double res = g.map.resolution_;
    bool safe_flag = true;
    for (int i = 0; vp.x - safe_radius_vp + i * res < vp.x + safe_radius_vp; i = i + 2)
      for (int j = 0; vp.y - safe_radius_vp + j * res < vp.y + safe_radius_vp; j = j + 2)
        for (int k = 0; vp.z - safe_radius_vp + k * res < vp.z + safe_radius_vp; k = k + 2) {
          const V3 safe_check_pt(vp.x - safe_radius_vp + i * res, vp.y - safe_radius_vp + j * res, vp.z - safe_radius_vp + k * res);
          ctr.lookups++;
          safe_flag = g.safety_check(safe_check_pt);
          if (safe_flag == false) break;
        }
    return safe_flag;
  }

  // hcsolver.cpp:556-583: true = the straight segment is free; dist = (p1 - p2).norm()
  bool straightLine(const V3& p1, const V3& p2, double& dist) {
    bool safe = true;
    V3i idx;
    raycaster_.capped = false;
    raycaster_.input(p1, p2);
    while (raycaster_.nextId(idx)) {
      ctr.lookups++;
      // the reference reads both buffers before isInMap_hc (out-of-map reads are undefined behaviour there)
      bool blocked = !g.map.isInMap_hc(idx);
      if (!blocked) {
        const size_t a = (size_t)g.map.toAddress_hc(idx);
        blocked = g.inflate[a] == 1 || g.map.occupancy_buffer_internal_[a] == 1;
      }
      if (blocked) { safe = false; break; }
    }
    if (raycaster_.capped) { ctr.cap_trips++; safe = false; }     // the reference would never return; refuse the segment
    dist = norm(p1 - p2);
    return safe;
  }

  // ---- SDFMap::SetOcc (sdf_map.cpp:440-455) / the marking loop of initHCMap (:178-187)
  void setOcc(const float* xyz, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
      const V3 pt_w(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
      V3i idx;
      g.map.posToIndex_hc(pt_w, idx);
      if (g.map.isInMap_hc(idx)) g.map.occupancy_buffer_hc_[(size_t)g.map.toAddress_hc(idx)] = 1;
    }
  }

  // one `internal_cast_->input(a, b); nextId(idx); while (nextId(idx)) buffer[idx] = value` loop (sdf_map.cpp:332-337, 379-384,
  // 465-471).  Voxels outside the map are skipped and counted (the reference writes out of bounds there).
  template <class F>
  void castMark(const V3& a, const V3& b, F&& touch) {
    V3i idx;
    raycaster_.capped = false;
    raycaster_.input(a, b);
    raycaster_.nextId(idx);
    while (raycaster_.nextId(idx)) {
      ctr.lookups++;
      if (g.map.isInMap_hc(idx)) touch((size_t)g.map.toAddress_hc(idx));
      else oob_writes++;
    }
    if (raycaster_.capped) ctr.cap_trips++;
  }
  int64_t oob_writes = 0;

  // SDFMap::points_in_plane (sdf_map.cpp:475-491): indices instead of copies; marks seg_occ_visited_buffer_
  std::vector<int> pointsInPlane(const float* cloud, int64_t n, const V3& point_on_plane, const V3& plane_normal, double thickness, std::vector<char>& visited) {
    std::vector<int> ids;
    for (int64_t i = 0; i < n; ++i) {
      const V3 pt_vec(cloud[3 * i], cloud[3 * i + 1], cloud[3 * i + 2]);
      const double distance_from_plane = dot(pt_vec - point_on_plane, plane_normal);
      if (std::abs(distance_from_plane) <= thickness) { ids.push_back((int)i); visited[(size_t)i] = 1; }
    }
    return ids;
  }

  // SDFMap::InternalSpace (sdf_map.cpp:263-395) for ONE segment: cloud = its points (float xyz), seg_start / seg_end = the two
  // skeleton vertices.  The reference iterates a std::map over the segments; every write is "= 1", so the order is immaterial.
  void internalSpaceSegment(const float* cloud, int64_t n, const V3& seg_start, const V3& seg_end, int inflate_num, double proj_interval, double thickness) {
    const V3 seg_dir = normalized(seg_end - seg_start);
    const double seg_length = norm(seg_end - seg_start);
    for (int64_t p = 0; p < n; ++p) {
      const V3 pt_w(cloud[3 * p], cloud[3 * p + 1], cloud[3 * p + 2]);
      V3i idx;
      g.map.posToIndex_hc(pt_w, idx);
      if (g.map.isInMap_hc(idx)) {
        const size_t a = (size_t)g.map.toAddress_hc(idx);
        g.map.occupancy_buffer_hc_[a] = 1;
        g.inflate[a] = 1;
      } else oob_writes++;
      for (int i = -inflate_num; i < inflate_num + 1; i = i + 1)
        for (int j = -inflate_num; j < inflate_num + 1; j = j + 1)
          for (int k = -inflate_num; k < inflate_num + 1; k = k + 1) {
            if (i == 0 && j == 0 && k == 0) continue;
            const V3i neighbor(idx.x + i, idx.y + j, idx.z + k);
            if (g.map.isInMap_hc(neighbor)) g.inflate[(size_t)g.map.toAddress_hc(neighbor)] = 1;
            else oob_writes++;
          }
    }
    /* cutting plane method */
    std::vector<char> visited((size_t)n, 0);
    std::vector<V3> proj_points;
    for (int i = 0; proj_interval * i < seg_length; ++i) proj_points.push_back(seg_start + (proj_interval * i) * seg_dir);
    proj_points.push_back(seg_end);
    auto mark = [&](size_t a) { g.map.occupancy_buffer_internal_[a] = 1; };
    for (const V3& proj_pt : proj_points) {
      const std::vector<int> cut_plane = pointsInPlane(cloud, n, proj_pt, seg_dir, thickness, visited);
      for (int id : cut_plane) castMark(proj_pt, V3(cloud[3 * (size_t)id], cloud[3 * (size_t)id + 1], cloud[3 * (size_t)id + 2]), mark);
    }
    for (int64_t j = 0; j < n; ++j) {
      if (visited[(size_t)j]) continue;
      const V3 remain_pt(cloud[3 * j], cloud[3 * j + 1], cloud[3 * j + 2]);
      double dist = 0.0, dist_min = 100000.0;
      int distri_id = -1;
      for (int k = 0; k < (int)proj_points.size(); ++k) {
        dist = norm(proj_points[(size_t)k] - remain_pt);
        if (dist < dist_min) { dist_min = dist; distri_id = k; }
      }
      if (distri_id < 0) continue;          // the reference indexes proj_points[-1] here (every distance >= 100000)
      castMark(proj_points[(size_t)distri_id], remain_pt, mark);
    }
  }

  // SDFMap::OuterCheck (sdf_map.cpp:457-473): outers = n x (position, direction)
  void outerCheck(const double* outers6, int64_t n, double checkScale) {
    for (int64_t i = 0; i < n; ++i) {
      const V3 start(outers6[6 * i], outers6[6 * i + 1], outers6[6 * i + 2]);
      const V3 end = start + checkScale * V3(outers6[6 * i + 3], outers6[6 * i + 4], outers6[6 * i + 5]);
      castMark(start, end, [&](size_t a) { if (g.map.occupancy_buffer_internal_[a] == 1) g.map.occupancy_buffer_internal_[a] = 0; });
    }
  }

 private:
  bool occ(const V3i& idx) { ctr.lookups++; return g.map.occCheck(idx); }
};

}  // namespace orc
