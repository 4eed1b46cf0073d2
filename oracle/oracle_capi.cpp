// oracle_capi.cpp — C entry points of the CPU ORACLE for ctypes.  TEST INFRASTRUCTURE.
// Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke()
// load this library, as the checker.  See fcvm_oracle.hpp for the restatement and its citations.
#include "fcvm_oracle.hpp"
#include "fcvm_oracle_cov.hpp"
#include "fcvm_oracle_plan.hpp"

#include <atomic>
#include <thread>

using namespace orc;

struct orc_handle {
  ViewpointManager vm;
  std::vector<float> model_xyz;  // copy of the model cloud for evalPass
};

static void rows_from(const std::vector<int>& vis, uint32_t* row, int64_t words) {
  for (int64_t w = 0; w < words; ++w) row[w] = 0u;
  for (int j : vis) row[j >> 5] |= 1u << (j & 31);
}

extern "C" {

orc_handle* orc_create(const Params* p) {
  orc_handle* h = new orc_handle();
  h->vm.init(*p);
  return h;
}
void orc_destroy(orc_handle* h) { delete h; }
int orc_reset(orc_handle* h) { h->vm.reset(); h->model_xyz.clear(); return 0; }
int orc_set_map_cloud(orc_handle* h, const float* xyz, int64_t n) { h->vm.setMapPointCloud(xyz, n); return 0; }
int orc_set_model(orc_handle* h, const float* xyz, int64_t n) {
  if (n == 0) return -1;
  h->vm.setModel(xyz, n);
  h->model_xyz.assign(xyz, xyz + 3 * n);
  return 0;
}
int orc_set_normals(orc_handle* h, const double* pt, const double* nrm, int64_t n) {
  if (n == 0) return -1;
  h->vm.setNormals(pt, nrm, n);
  return 0;
}
int orc_set_init_viewpoints(orc_handle* h, const float* pos_dir, int64_t n) {
  if (n == 0) return -1;
  std::vector<PointNormal> v((size_t)n);
  for (int64_t i = 0; i < n; ++i)
    v[(size_t)i] = PointNormal{pos_dir[6 * i], pos_dir[6 * i + 1], pos_dir[6 * i + 2], pos_dir[6 * i + 3], pos_dir[6 * i + 4], pos_dir[6 * i + 5]};
  h->vm.setInitViewpoints(v);
  return 0;
}
int orc_update(orc_handle* h) {
  if (!h->vm.TS.map_flag_ || !h->vm.TS.model_flag_ || !h->vm.TS.normal_flag_) return -2;
  h->vm.updateViewpoints();
  return h->vm.TS.isReady() ? 0 : -2;
}
int64_t orc_get_viewpoints(orc_handle* h, int32_t* vp_id, double* pose5, int32_t* vox, int64_t cap) {
  const auto& f = h->vm.getUpdatedViewpoints();
  for (int64_t i = 0; i < (int64_t)f.size() && i < cap; ++i) {
    if (vp_id) vp_id[i] = f[(size_t)i].vp_id;
    if (pose5) for (int k = 0; k < 5; ++k) pose5[5 * i + k] = f[(size_t)i].pose[k];
    if (vox) vox[i] = f[(size_t)i].vox_count;
  }
  return (int64_t)f.size();
}
int64_t orc_get_uncovered(orc_handle* h, float* xyz, int64_t cap) {
  std::vector<PointXYZ> c;
  h->vm.getUncoveredModelCloud(c);
  for (int64_t i = 0; i < (int64_t)c.size() && i < cap; ++i) {
    xyz[3 * i] = c[(size_t)i].x; xyz[3 * i + 1] = c[(size_t)i].y; xyz[3 * i + 2] = c[(size_t)i].z;
  }
  return (int64_t)c.size();
}
int64_t orc_get_vps_pose(orc_handle* h, double* pose5, int64_t cap_rows) {
  const auto& v = h->vm.VPI.vps_pose;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap_rows; ++i)
    for (int k = 0; k < 5; ++k) pose5[5 * i + k] = v[(size_t)i].v[k];
  return (int64_t)v.size();
}
int64_t orc_get_vps_voxcount(orc_handle* h, int32_t* out, int64_t cap) {
  const auto& v = h->vm.VPI.vps_voxcount;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}
int64_t orc_get_vps_contri(orc_handle* h, int32_t* out, int64_t cap) {
  const auto& v = h->vm.VPI.vps_contri;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}
int64_t orc_get_cover_state(orc_handle* h, uint8_t* covered, int32_t* cover_contrib, int32_t* contrib_id, int64_t cap) {
  const auto& m = h->vm.MCI;
  for (int64_t i = 0; i < (int64_t)m.cover_state.size() && i < cap; ++i) {
    if (covered) covered[i] = m.cover_state[(size_t)i] ? 1 : 0;
    if (cover_contrib) cover_contrib[i] = m.cover_contrib[(size_t)i];
    if (contrib_id) contrib_id[i] = m.contrib_id[(size_t)i];
  }
  return (int64_t)m.cover_state.size();
}
int orc_get_grid(orc_handle* h, double* origin3, int32_t* voxel_num3, uint8_t* bits, int64_t cap_bytes) {
  const HCMap& m = h->vm.HCMap_;
  if (origin3) { origin3[0] = m.map_origin_.x; origin3[1] = m.map_origin_.y; origin3[2] = m.map_origin_.z; }
  if (voxel_num3) { voxel_num3[0] = m.map_voxel_num_.x; voxel_num3[1] = m.map_voxel_num_.y; voxel_num3[2] = m.map_voxel_num_.z; }
  if (bits) {
    int64_t g = (int64_t)m.occupancy_buffer_hc_.size();
    if (cap_bytes < (g + 7) / 8) return -5;
    std::memset(bits, 0, (size_t)((g + 7) / 8));
    for (int64_t i = 0; i < g; ++i)
      if (m.occupancy_buffer_hc_[(size_t)i] == 1) bits[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
  return 0;
}
int orc_trace_enable(orc_handle* h, int32_t on) { h->vm.trace_on = on != 0; if (!on) h->vm.trace.clear(); return 0; }
int64_t orc_trace_count(orc_handle* h) { return (int64_t)h->vm.trace.size(); }
int64_t orc_trace_words(orc_handle* h) { return ((int64_t)h->vm.model_.size() + 31) / 32; }
int orc_trace_get(orc_handle* h, int64_t i, int32_t* stage, int32_t* vp_id, double* pose5, uint32_t* row) {
  if (i < 0 || i >= (int64_t)h->vm.trace.size()) return -1;
  const TraceEvent& e = h->vm.trace[(size_t)i];
  if (stage) *stage = e.stage;
  if (vp_id) *vp_id = e.vp_id;
  if (pose5) for (int k = 0; k < 5; ++k) pose5[k] = e.pose[k];
  if (row) rows_from(e.visible, row, orc_trace_words(h));
  return 0;
}
int64_t orc_prune_order(orc_handle* h, int32_t* out, int64_t cap) {
  const auto& v = h->vm.last_prune_order;
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) out[i] = v[(size_t)i];
  return (int64_t)v.size();
}
// [0] pair tests [1] rays [2] voxel lookups [3] walk lookups [4] cap trips [5] marcher calls
// [6] viewpoint evaluations [7] safety_check lookups
int orc_get_counters(orc_handle* h, int64_t* out8) {
  const Counters& c = h->vm.ctr();
  out8[0] = c.pair_tests; out8[1] = c.rays; out8[2] = c.lookups; out8[3] = c.walk_lookups;
  out8[4] = c.cap_trips; out8[5] = c.steps; out8[6] = c.evals; out8[7] = c.safety_lookups;
  return 0;
}
int orc_get_stage_seconds(orc_handle* h, double* out5) {
  for (int k = 0; k < 5; ++k) out5[k] = h->vm.t_eval_s[k];
  return 0;
}

// Evaluator pass over a batch of poses (no pose / ownership side effects), optionally sharded
// over `threads` std::threads (the reference itself is single-threaded on this path).
// counters8 (may be NULL) receives the counters of this call only.
int orc_evaluate(orc_handle* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows,
                 int32_t* count, uint32_t* rows, int32_t threads, int64_t* counters8) {
  if (stage < 1 || stage > 4) return -1;
  const int64_t P = (int64_t)h->vm.model_.size();
  const int64_t words = (P + 31) / 32;
  if (threads < 1) threads = 1;
  if (threads > n) threads = (int32_t)std::max<int64_t>(1, n);
  std::vector<Counters> ctrs((size_t)threads);
  std::atomic<int64_t> next(0);
  auto work = [&](int t) {
    EvalContext ec;
    ec.percep_utils_.init(h->vm.prm);
    ec.raycaster_.setParams(h->vm.HCMap_.resolution_, h->vm.HCMap_.map_origin_);
    ec.map = &h->vm.HCMap_;
    std::vector<int> vis;
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= n) break;
      ec.evalPass(stage, pose5 + 5 * i, h->model_xyz.data(), P, h->vm.prm.visible_range,
                  restrict_rows ? restrict_rows + i * words : nullptr, vis);
      if (count) count[i] = (int32_t)vis.size();
      if (rows) rows_from(vis, rows + i * words, words);
    }
    ctrs[(size_t)t] = ec.ctr;
  };
  if (threads == 1) {
    work(0);
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  if (counters8) {
    for (int k = 0; k < 8; ++k) counters8[k] = 0;
    for (auto& c : ctrs) {
      counters8[0] += c.pair_tests; counters8[1] += c.rays; counters8[2] += c.lookups; counters8[3] += c.walk_lookups;
      counters8[4] += c.cap_trips; counters8[5] += c.steps; counters8[6] += c.evals; counters8[7] += c.safety_lookups;
    }
  }
  return 0;
}

// ---- micro known-answer taps (tests/test_oracle_kat.py, tests/test_oracle_vs_ref.py) ----
double orc_intbound(double s, double ds) { return intbound(s, ds); }
double orc_mod(double v, double m) { return mod(v, m); }

// March start->end with the one-directional caster; out = visited map indices (3 ints per call),
// including the final (false) call.  Returns the number of nextId calls made.
int64_t orc_ray_ids(double res, const double* origin3, const double* start3, const double* end3, int32_t* out, int64_t cap_calls) {
  RayCaster rc;
  rc.setParams(res, V3(origin3[0], origin3[1], origin3[2]));
  rc.input(V3(start3[0], start3[1], start3[2]), V3(end3[0], end3[1], end3[2]));
  int64_t k = 0;
  V3i idx;
  for (;;) {
    bool more = rc.nextId(idx);
    if (k < cap_calls) { out[3 * k] = idx.x; out[3 * k + 1] = idx.y; out[3 * k + 2] = idx.z; }
    ++k;
    if (!more) break;
  }
  return k;
}
// Same for the bidirectional caster: out = (idx, idxR) pairs, 6 ints per call.
int64_t orc_biray_ids(double res, const double* origin3, const double* start3, const double* end3, int32_t* out, int64_t cap_calls) {
  RayCaster rc;
  rc.setParams(res, V3(origin3[0], origin3[1], origin3[2]));
  rc.biInput(V3(start3[0], start3[1], start3[2]), V3(end3[0], end3[1], end3[2]));
  int64_t k = 0;
  V3i a, b;
  for (;;) {
    bool more = rc.biNextId(a, b);
    if (k < cap_calls) {
      out[6 * k] = a.x; out[6 * k + 1] = a.y; out[6 * k + 2] = a.z;
      out[6 * k + 3] = b.x; out[6 * k + 4] = b.y; out[6 * k + 5] = b.z;
    }
    ++k;
    if (!more) break;
  }
  return k;
}
int orc_fov_normals(const Params* p, double pitch, double yaw, double* out12) {
  PerceptionUtils pu;
  pu.init(*p);
  pu.setPose_PY(V3(0, 0, 0), pitch, yaw);
  for (int q = 0; q < 4; ++q) for (int i = 0; i < 3; ++i) out12[3 * q + i] = pu.normals_[q][i];
  return 0;
}
int orc_inside_fov(const Params* p, const double* pose5, const double* pts, int64_t n, uint8_t* out) {
  PerceptionUtils pu;
  pu.init(*p);
  pu.setPose_PY(V3(pose5[0], pose5[1], pose5[2]), pose5[3], pose5[4]);
  for (int64_t i = 0; i < n; ++i) out[i] = pu.insideFOV(V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])) ? 1 : 0;
  return 0;
}
int orc_safety_check(orc_handle* h, const float* xyz, int64_t n, uint8_t* out) {
  for (int64_t i = 0; i < n; ++i) {
    PointXYZ p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    out[i] = h->vm.viewpointSafetyCheck(p) ? 1 : 0;
  }
  return 0;
}
int orc_pitch_yaw(const double* v3, double* py2) { ViewpointManager::PitchYaw(V3(v3[0], v3[1], v3[2]), py2); return 0; }

// ---- pin-hole coverage evaluation (fcvm_oracle_cov.hpp; hcplanner.cpp:1032-1211) ----
struct orc_cov_handle { CoverageEvaluator ce; };

orc_cov_handle* orc_cov_create(const Params* p, const Camera* c) {
  orc_cov_handle* h = new orc_cov_handle();
  h->ce.init(*p, *c);
  return h;
}
void orc_cov_destroy(orc_cov_handle* h) { delete h; }
int orc_cov_set_cloud(orc_cov_handle* h, const float* xyz, int64_t n) { h->ce.setCloud(xyz, n); return 0; }
// PinHoleCamera for m poses (optionally over `threads` std::threads; poses are independent):
// offsets[m + 1], ids (ascending per pose, at most cap), covered[n] = union.  Returns the total count.
int64_t orc_cov_pinhole(orc_cov_handle* h, const double* pose5, int64_t m, int64_t* offsets, int32_t* ids, int64_t cap, uint8_t* covered,
                        int32_t threads, int64_t* counters3) {
  std::vector<std::vector<int>> vis((size_t)m);
  if (threads < 1) threads = 1;
  std::vector<CovCounters> ctrs((size_t)threads);
  std::atomic<int64_t> next(0);
  auto work = [&](int t) {
    CoverageEvaluator ce = h->ce;      // own scratch tables, shared read-only cloud copy
    ce.ctr = CovCounters();
    for (;;) {
      const int64_t k = next.fetch_add(1);
      if (k >= m) break;
      ce.PinHoleCamera(pose5 + 5 * k, vis[(size_t)k]);
    }
    ctrs[(size_t)t] = ce.ctr;
  };
  if (threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  int64_t total = 0;
  if (covered) for (int64_t i = 0; i < h->ce.size(); ++i) covered[i] = 0;
  for (int64_t k = 0; k < m; ++k) {
    if (offsets) offsets[k] = total;
    for (int id : vis[(size_t)k]) {
      if (ids && total < cap) ids[total] = id;
      if (covered) covered[id] = 1;
      ++total;
    }
  }
  if (offsets) offsets[m] = total;
  if (counters3) {
    counters3[0] = counters3[1] = counters3[2] = 0;
    for (auto& c : ctrs) { counters3[0] += c.pair_tests; counters3[1] += c.projections; counters3[2] += c.out_of_image; }
  }
  return total;
}
// CoverageEvaluation: returns the rate through *rate; groups as CSR; ever[n] = points visible in any pose
int64_t orc_cov_evaluate(orc_cov_handle* h, const double* pose5, int64_t m, double* rate, int64_t* group_offsets, int32_t* group_ids, int64_t cap,
                         uint8_t* ever) {
  std::vector<std::vector<int>> groups;
  std::vector<char> ev;
  const double r = h->ce.CoverageEvaluation(pose5, m, groups, ev);
  if (rate) *rate = r;
  int64_t total = 0;
  for (int64_t k = 0; k < m; ++k) {
    if (group_offsets) group_offsets[k] = total;
    for (int id : groups[(size_t)k]) {
      if (group_ids && total < cap) group_ids[total] = id;
      ++total;
    }
  }
  if (group_offsets) group_offsets[m] = total;
  if (ever) for (size_t i = 0; i < ev.size(); ++i) ever[i] = (uint8_t)ev[i];
  return total;
}
// CameraModelProjection for single points (known-answer tests): out = distance, xID, yID per point
int orc_cov_project(orc_cov_handle* h, const double* pose5, const double* pts, int64_t n, double* dist, int32_t* xy) {
  double M[3][3];
  CoverageEvaluator::poseMatrix(pose5, M);
  for (int64_t i = 0; i < n; ++i) {
    int cod[2];
    h->ce.project(pose5, M, V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]), dist[i], cod);
    xy[2 * i] = cod[0]; xy[2 * i + 1] = cod[1];
  }
  return 0;
}
int orc_cov_image(orc_cov_handle* h, int32_t* wh, double* res2) {
  wh[0] = h->ce.xPixel; wh[1] = h->ce.yPixel; res2[0] = h->ce.xRes; res2[1] = h->ce.yRes;
  return 0;
}

// ---- planner-grid loops (fcvm_oracle_plan.hpp; hcplanner.cpp:658-692, 805-863; hcsolver.cpp:556-583) ----
struct orc_plan_handle { PlanOps ops; };

orc_plan_handle* orc_plan_create(const double* origin3, double res, const int32_t* dims3, const int8_t* hc, const int8_t* inflate, const int8_t* internal) {
  orc_plan_handle* h = new orc_plan_handle();
  const int d[3] = {dims3[0], dims3[1], dims3[2]};
  h->ops.setGrid(origin3, res, d, hc, inflate, internal);
  return h;
}
void orc_plan_destroy(orc_plan_handle* h) { delete h; }
int orc_plan_light_state(orc_plan_handle* h, const double* cand, int64_t n, const float* rosa, int64_t m, uint8_t* det, int32_t* cross, int32_t* nearest) {
  FloatCloudSearch tree;
  tree.setInputCloud(std::vector<float>(rosa, rosa + 3 * m));
  for (int64_t i = 0; i < n; ++i) {
    int c = 0, id = -1;
    const bool d = h->ops.lightStateDet(V3(cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]), V3(0, 0, 1), tree, c, id);
    det[i] = d ? 1 : 0;
    if (cross) cross[i] = c;
    if (nearest) nearest[i] = id;
  }
  return 0;
}
int orc_plan_safety_box(orc_plan_handle* h, const float* vp, int64_t n, double safe_radius, uint8_t* ok) {
  for (int64_t i = 0; i < n; ++i) {
    PointNormal p{vp[3 * i], vp[3 * i + 1], vp[3 * i + 2], 0, 0, 0};
    ok[i] = h->ops.safetyBox(p, safe_radius) ? 1 : 0;
  }
  return 0;
}
// all ordered pairs (i, j): safe[i * n + j], dist[i * n + j]; `threads` std::threads over rows
int orc_plan_line_of_sight(orc_plan_handle* h, const double* p, int64_t n, uint8_t* safe, double* dist, int32_t threads) {
  if (threads < 1) threads = 1;
  std::atomic<int64_t> next(0);
  std::vector<PlanCounters> ctrs((size_t)threads);
  auto work = [&](int t) {
    PlanOps ops = h->ops;      // own RayCaster scratch; the grid copy is shared-nothing but read-only in use
    ops.ctr = PlanCounters();
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n) break;
      for (int64_t j = 0; j < n; ++j) {
        double d;
        const bool s = ops.straightLine(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), V3(p[3 * j], p[3 * j + 1], p[3 * j + 2]), d);
        safe[i * n + j] = s ? 1 : 0;
        if (dist) dist[i * n + j] = d;
      }
    }
    ctrs[(size_t)t] = ops.ctr;
  };
  if (threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  for (auto& c : ctrs) { h->ops.ctr.lookups += c.lookups; h->ops.ctr.cap_trips += c.cap_trips; }
  return 0;
}
// SDFMap::SetOcc / InternalSpace / OuterCheck (sdf_map.cpp:263-473) on the handle's three buffers
int orc_plan_set_occ(orc_plan_handle* h, const float* xyz, int64_t n) { h->ops.setOcc(xyz, n); return 0; }
int orc_plan_internal_space(orc_plan_handle* h, const float* seg_xyz, const int64_t* seg_offsets, int64_t nseg, const double* seg_ends6, int32_t inflate_num,
                            double proj_interval, double thickness) {
  for (int64_t s = 0; s < nseg; ++s)
    h->ops.internalSpaceSegment(seg_xyz + 3 * seg_offsets[s], seg_offsets[s + 1] - seg_offsets[s], V3(seg_ends6[6 * s], seg_ends6[6 * s + 1], seg_ends6[6 * s + 2]),
                                V3(seg_ends6[6 * s + 3], seg_ends6[6 * s + 4], seg_ends6[6 * s + 5]), inflate_num, proj_interval, thickness);
  return 0;
}
int orc_plan_outer_check(orc_plan_handle* h, const double* outers6, int64_t n, double check_scale) { h->ops.outerCheck(outers6, n, check_scale); return 0; }
int orc_plan_get_buffers(orc_plan_handle* h, int8_t* hc, int8_t* inflate, int8_t* internal) {
  const size_t G = h->ops.g.inflate.size();
  for (size_t i = 0; i < G; ++i) {
    if (hc) hc[i] = (int8_t)h->ops.g.map.occupancy_buffer_hc_[i];
    if (inflate) inflate[i] = (int8_t)h->ops.g.inflate[i];
    if (internal) internal[i] = (int8_t)h->ops.g.map.occupancy_buffer_internal_[i];
  }
  return 0;
}
int64_t orc_plan_oob_writes(orc_plan_handle* h) { return h->ops.oob_writes; }
int orc_plan_counters(orc_plan_handle* h, int64_t* out2) { out2[0] = h->ops.ctr.lookups; out2[1] = h->ops.ctr.cap_trips; return 0; }

}  // extern "C"
