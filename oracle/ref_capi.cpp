// ref_capi.cpp — C entry points around the REFERENCE'S OWN raycast.cpp and perception_utils.cpp,
// compiled from where they lie under /root/reference (never copied) against the stand-in headers
// in shim/.  TEST INFRASTRUCTURE: built into oracle/_ref/libfcvm_ref.so, loaded only by tests/ (to
// validate the oracle restatement) — never by the product library.
//
// What comes from the reference here: every statement of RayCaster (setParams, input, biInput,
// nextId, biNextId), of the free functions signum / mod / intbound, and of PerceptionUtils (init,
// setPose_PY, insideFOV).  What does not: Eigen (shim/Eigen/Eigen), the ROS parameter server
// (shim/ros/ros.h), the occupancy grid (orc::HCMap of the oracle; sdf_map.cpp needs PCL) and the
// four call-site loops of viewpoint_manager.cpp that drive those objects, which ref_evaluate
// restates next to their line numbers (viewpoint_manager.cpp needs PCL + ROS).
#include <cstdint>
#include <iostream>
#include <memory>
#include <array>
#include <atomic>
#include <thread>
#include <vector>

#define private public   // normals_ of PerceptionUtils is read by ref_fov_normals
#include <active_perception/perception_utils.h>
#include <plan_env/raycast.h>
#undef private

#include "fcvm_oracle.hpp"
#include "fcvm_oracle_plan.hpp"

int signum(int x);   // raycast.cpp:26 defines this overload; raycast.h:29 only declares signum(double)

namespace {

void set_params(const orc::Params& p) {
  auto& t = ros::shim_param_table();
  t["perception_utils/top_angle"] = p.top_angle;
  t["perception_utils/left_angle"] = p.left_angle;
  t["perception_utils/right_angle"] = p.right_angle;
  t["perception_utils/max_dist"] = p.max_dist;
  t["perception_utils/vis_dist"] = p.vis_dist;
}

Eigen::Vector3d ev(const orc::V3& v) { return Eigen::Vector3d(v.x, v.y, v.z); }

struct RefEval {
  RayCaster raycaster_;
  predrecon::PerceptionUtils percep_utils_;
  const orc::HCMap* map = nullptr;
  int64_t lookups = 0, rays = 0, pairs = 0, walk = 0;

  bool occ(const Eigen::Vector3i& idx) {
    lookups++;
    return map->occCheck(orc::V3i(idx(0), idx(1), idx(2)));
  }
  // viewpoint_manager.cpp:380-396 (E1), 479-492 (E2), 841-853 (E3): one discarded biNextId;
  // 569-580 (E4): none
  bool biRC(const Eigen::Vector3d& a, const Eigen::Vector3d& b, bool discard_first) {
    Eigen::Vector3i idx, idx_rev;
    bool visible_flag = true, visible_flag_rev = true;
    rays++;
    raycaster_.biInput(a, b);
    if (discard_first) raycaster_.biNextId(idx, idx_rev);
    while (raycaster_.biNextId(idx, idx_rev)) {
      visible_flag = occ(idx);
      visible_flag_rev = occ(idx_rev);
      if (visible_flag == false || visible_flag_rev == false) break;
    }
    visible_flag = occ(idx);
    return visible_flag == true && visible_flag_rev == true;
  }
  // viewpoint_manager.cpp:463-477 / 554-568
  Eigen::Vector3d offsetTarget(const Eigen::Vector3d& pos1, const Eigen::Vector3d& pos2) {
    Eigen::Vector3i idx;
    raycaster_.input(pos1, pos2);
    int count = 0;
    while (raycaster_.nextId(idx)) {
      walk++;
      if (occ(idx)) {
        count++;
        if (count == 2) break;
      }
    }
    orc::V3 fp;
    map->indexToPos_hc(orc::V3i(idx(0), idx(1), idx(2)), fp);
    return ev(fp);
  }
};

}  // namespace

struct ref_handle {
  orc::Params prm;
  orc::HCMap map;
  std::vector<float> model;
};

extern "C" {

int ref_signum(int x) { return signum(x); }
double ref_mod(double v, double m) { return mod(v, m); }
double ref_intbound(double s, double ds) { return intbound(s, ds); }

// same contracts as orc_ray_ids / orc_biray_ids (oracle_capi.cpp)
int64_t ref_ray_ids(double res, const double* origin3, const double* start3, const double* end3, int32_t* out, int64_t cap_calls, int64_t max_calls) {
  RayCaster rc;
  rc.setParams(res, Eigen::Vector3d(origin3[0], origin3[1], origin3[2]));
  rc.input(Eigen::Vector3d(start3[0], start3[1], start3[2]), Eigen::Vector3d(end3[0], end3[1], end3[2]));
  int64_t k = 0;
  Eigen::Vector3i idx;
  for (;;) {
    bool more = rc.nextId(idx);
    if (k < cap_calls) { out[3 * k] = idx(0); out[3 * k + 1] = idx(1); out[3 * k + 2] = idx(2); }
    ++k;
    if (!more || k >= max_calls) break;   // the reference has no step cap: bound the harness instead
  }
  return k;
}
int64_t ref_biray_ids(double res, const double* origin3, const double* start3, const double* end3, int32_t* out, int64_t cap_calls, int64_t max_calls) {
  RayCaster rc;
  rc.setParams(res, Eigen::Vector3d(origin3[0], origin3[1], origin3[2]));
  rc.biInput(Eigen::Vector3d(start3[0], start3[1], start3[2]), Eigen::Vector3d(end3[0], end3[1], end3[2]));
  int64_t k = 0;
  Eigen::Vector3i a, b;
  for (;;) {
    bool more = rc.biNextId(a, b);
    if (k < cap_calls) {
      out[6 * k] = a(0); out[6 * k + 1] = a(1); out[6 * k + 2] = a(2);
      out[6 * k + 3] = b(0); out[6 * k + 4] = b(1); out[6 * k + 5] = b(2);
    }
    ++k;
    if (!more || k >= max_calls) break;
  }
  return k;
}

// The planner's call-site loops (hcplanner.cpp:828-849 lightStateDet, hcsolver.cpp:562-576 search_Path) driving the
// REFERENCE'S RayCaster; the grid is the oracle's PlanGrid, the nearest vertex is given.  max_calls bounds a marcher
// that would not terminate (reported through *capped).
int ref_light_state_cross(const orc::PlanGrid* g, const double* vp3, const double* vertex3, int64_t max_calls, int32_t* capped) {
  RayCaster rc;
  rc.setParams(g->map.resolution_, ev(g->map.map_origin_));
  Eigen::Vector3i idxFwd, idxRev;
  bool fwdFlag, revFlag;
  auto occ = [&](const Eigen::Vector3i& i) { return g->map.occCheck(orc::V3i(i(0), i(1), i(2))); };
  rc.biInput(Eigen::Vector3d(vp3[0], vp3[1], vp3[2]), Eigen::Vector3d(vertex3[0], vertex3[1], vertex3[2]));
  rc.biNextId(idxFwd, idxRev);
  int crossCount = 0;
  int64_t calls = 0;
  *capped = 0;
  while (rc.biNextId(idxFwd, idxRev)) {
    if (++calls > max_calls) { *capped = 1; break; }
    fwdFlag = occ(idxFwd);
    revFlag = occ(idxRev);
    if (fwdFlag == false) crossCount++;
    if (revFlag == false) crossCount++;
  }
  fwdFlag = occ(idxFwd);
  if (fwdFlag == false) crossCount++;
  return crossCount;
}
int ref_straight_line(const orc::PlanGrid* g, const double* p1, const double* p2, int64_t max_calls, int32_t* capped) {
  RayCaster rc;
  rc.setParams(g->map.resolution_, ev(g->map.map_origin_));
  bool safe = true;
  Eigen::Vector3i idx;
  int64_t calls = 0;
  *capped = 0;
  rc.input(Eigen::Vector3d(p1[0], p1[1], p1[2]), Eigen::Vector3d(p2[0], p2[1], p2[2]));
  while (rc.nextId(idx)) {
    if (++calls > max_calls) { *capped = 1; safe = false; break; }
    const orc::V3i id(idx(0), idx(1), idx(2));
    bool blocked = !g->map.isInMap_hc(id);
    if (!blocked) {
      const size_t a = (size_t)g->map.toAddress_hc(id);
      blocked = g->inflate[a] == 1 || g->map.occupancy_buffer_internal_[a] == 1;
    }
    if (blocked) { safe = false; break; }
  }
  return safe ? 1 : 0;
}
// The marking loop of SDFMap::InternalSpace / OuterCheck (sdf_map.cpp:332-337, 379-384, 465-471) driving the reference's RayCaster:
//   internal_cast_->input(a, b); internal_cast_->nextId(idx); while (internal_cast_->nextId(idx)) <touch idx>;
// returns the number of voxels touched; the first `cap` of them as map indices in out[3 * k ..]
int64_t ref_cast_marks(const orc::PlanGrid* g, const double* a, const double* b, int32_t* out, int64_t cap, int64_t max_calls, int32_t* capped) {
  RayCaster rc;
  rc.setParams(g->map.resolution_, ev(g->map.map_origin_));
  Eigen::Vector3i idx;
  int64_t k = 0, calls = 0;
  *capped = 0;
  rc.input(Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Vector3d(b[0], b[1], b[2]));
  rc.nextId(idx);
  while (rc.nextId(idx)) {
    if (++calls > max_calls) { *capped = 1; break; }
    if (k < cap) { out[3 * k] = idx(0); out[3 * k + 1] = idx(1); out[3 * k + 2] = idx(2); }
    ++k;
  }
  return k;
}
orc::PlanGrid* ref_plan_grid_create(const double* origin3, double res, const int32_t* dims3, const int8_t* hc, const int8_t* inflate, const int8_t* internal) {
  orc::PlanGrid* g = new orc::PlanGrid();
  const int d[3] = {dims3[0], dims3[1], dims3[2]};
  g->set(origin3, res, d, hc, inflate, internal);
  return g;
}
void ref_plan_grid_destroy(orc::PlanGrid* g) { delete g; }

int ref_fov_normals(const orc::Params* p, double pitch, double yaw, double* out12) {
  set_params(*p);
  ros::NodeHandle nh;
  predrecon::PerceptionUtils pu;
  pu.init(nh);
  pu.setPose_PY(Eigen::Vector3d(0.0, 0.0, 0.0), pitch, yaw);
  for (int q = 0; q < 4; ++q)
    for (int i = 0; i < 3; ++i) out12[3 * q + i] = pu.normals_[q](i);
  return 0;
}
int ref_inside_fov(const orc::Params* p, const double* pose5, const double* pts, int64_t n, uint8_t* out) {
  set_params(*p);
  ros::NodeHandle nh;
  predrecon::PerceptionUtils pu;
  pu.init(nh);
  pu.setPose_PY(Eigen::Vector3d(pose5[0], pose5[1], pose5[2]), pose5[3], pose5[4]);
  for (int64_t i = 0; i < n; ++i) out[i] = pu.insideFOV(Eigen::Vector3d(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])) ? 1 : 0;
  return 0;
}

ref_handle* ref_create(const orc::Params* p) {
  ref_handle* h = new ref_handle();
  h->prm = *p;
  return h;
}
void ref_destroy(ref_handle* h) { delete h; }
int ref_set_map_cloud(ref_handle* h, const float* xyz, int64_t n) { h->map.initHCMap(h->prm, xyz, n); return 0; }
int ref_set_model(ref_handle* h, const float* xyz, int64_t n) { h->model.assign(xyz, xyz + 3 * n); return 0; }

// One evaluator pass per pose with the reference's RayCaster and PerceptionUtils objects, driven
// exactly as viewpoint_manager.cpp drives them (E1 358-425, E2 457-498, E3 822-875, E4 546-586).
// Same contract as orc_evaluate (single thread).  counters4 = pair tests, rays, lookups, walk lookups.
static void ref_eval_range(ref_handle* h, RefEval& e, int32_t stage, const double* pose5, int64_t v0, int64_t v1, const uint32_t* restrict_rows, int32_t* count,
                           uint32_t* rows) {
  const int64_t P = (int64_t)h->model.size() / 3, words = (P + 31) / 32;
  for (int64_t v = v0; v < v1; ++v) {
    const double* q = pose5 + 5 * v;
    const Eigen::Vector3d vp(q[0], q[1], q[2]);
    e.percep_utils_.setPose_PY(vp, q[3], q[4]);
    uint32_t* row = rows ? rows + v * words : nullptr;
    if (row) for (int64_t w = 0; w < words; ++w) row[w] = 0u;
    int c = 0;
    for (int64_t i = 0; i < P; ++i) {
      if (stage == 2 && !((restrict_rows[v * words + (i >> 5)] >> (i & 31)) & 1u)) continue;
      const Eigen::Vector3d pt(h->model[3 * i], h->model[3 * i + 1], h->model[3 * i + 2]);
      e.pairs++;
      if (!e.percep_utils_.insideFOV(pt)) continue;
      bool vis;
      if (stage == 1) {
        const bool free_ray = e.biRC(vp, pt, true);
        const double ray_length = (pt - vp).norm();                      // viewpoint_manager.cpp:398-403
        vis = free_ray && ray_length <= h->prm.visible_range;
      } else if (stage == 2) {
        vis = e.biRC(vp, e.offsetTarget(pt, vp), true);
      } else if (stage == 3) {
        vis = e.biRC(vp, pt, true);
      } else {
        vis = e.biRC(vp, e.offsetTarget(pt, vp), false);
      }
      if (vis) {
        ++c;
        if (row) row[i >> 5] |= 1u << (i & 31);
      }
    }
    if (count) count[v] = c;
  }
}

int ref_evaluate(ref_handle* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows, int32_t* count,
                 uint32_t* rows, int64_t* counters4) {
  if (stage < 1 || stage > 4) return -1;
  set_params(h->prm);
  ros::NodeHandle nh;
  RefEval e;
  e.percep_utils_.init(nh);
  e.raycaster_.setParams(h->map.resolution_, ev(h->map.map_origin_));     // viewpoint_manager.cpp:84-85
  e.map = &h->map;
  ref_eval_range(h, e, stage, pose5, 0, n, restrict_rows, count, rows);
  if (counters4) { counters4[0] = e.pairs; counters4[1] = e.rays; counters4[2] = e.lookups; counters4[3] = e.walk; }
  return 0;
}

// The same pass with the viewpoints handed out to `threads` host threads, each with its own RayCaster / PerceptionUtils objects
// (the reference itself is single-threaded on this path; this is what "all host cores" means for bench.py --impl reference).
int ref_evaluate_mt(ref_handle* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows, int32_t* count, uint32_t* rows,
                    int64_t* counters4, int32_t threads) {
  if (stage < 1 || stage > 4) return -1;
  if (threads < 1) threads = 1;
  set_params(h->prm);
  std::atomic<int64_t> next(0);
  std::vector<std::array<int64_t, 4>> ctr((size_t)threads, std::array<int64_t, 4>{0, 0, 0, 0});
  auto work = [&](int t) {
    ros::NodeHandle nh;
    RefEval e;
    e.percep_utils_.init(nh);
    e.raycaster_.setParams(h->map.resolution_, ev(h->map.map_origin_));
    e.map = &h->map;
    for (;;) {
      const int64_t v = next.fetch_add(1);
      if (v >= n) break;
      ref_eval_range(h, e, stage, pose5, v, v + 1, restrict_rows, count, rows);
    }
    ctr[(size_t)t] = {e.pairs, e.rays, e.lookups, e.walk};
  };
  std::vector<std::thread> th;
  for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& t : th) t.join();
  if (counters4) {
    for (int k = 0; k < 4; ++k) counters4[k] = 0;
    for (auto& c : ctr) for (int k = 0; k < 4; ++k) counters4[k] += c[(size_t)k];
  }
  return 0;
}

}  // extern "C"
