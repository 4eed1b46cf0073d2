// fcvm_oracle_plan.hpp — CPU ORACLE of three ray / probe loops of the planner that run on the planner's own SDFMap
// (SURVEY.md 8(f) ranks 2 and 4).  TEST INFRASTRUCTURE, NOT PRODUCT CODE (rules of fcvm_oracle.hpp).
//
// Restated, statement by statement, from FC-Planner/src/hierarchical_coverage_planner/src/:
//   hcplanner.cpp:805-863   lightStateDet: BiRC from a candidate viewpoint to its nearest ROSA vertex, parity of the
//                           number of occupied voxels crossed (odd = the candidate is outside the structure)
//   hcplanner.cpp:658-692   the safety box around a candidate (lattice of SDFMap::safety_check probes, stride 2)
//   hcsolver.cpp:556-583    search_Path's straight-line test: one-directional march p1 -> p2 against the inflated and
//                           internal buffers; on success the cost is the Euclidean distance
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

 private:
  bool occ(const V3i& idx) { ctr.lookups++; return g.map.occCheck(idx); }
};

}  // namespace orc
