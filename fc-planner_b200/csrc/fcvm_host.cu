// fcvm_host.cu — the manager behind the C ABI (include/fcvm.h): a batched, device-resident
// re-design of predrecon::ViewpointManager (FC-Planner/src/viewpoint_manager/src/viewpoint_manager.cpp).
//
// The reference evaluates one viewpoint at a time with a stateful ray caster.  Here each stage of
// the call sequence becomes a few launches over ALL of the stage's viewpoints:
//
//   setMapPointCloud (cpp:74-89 + SDFMap::initHCMap)
//       K_cloud_minmax -> host: grid geometry in FP64 -> K_voxelise; cell index of the map cloud (fcvm_map.cu)
//   setInitViewpoints / round appends (getPose, cpp:332-531)
//       host: PitchYaw(dir) -> FOV normals                (libm: host, same as the reference)
//       K_E1 over the batch -> bitset rows -> CSR          (FOV + BiRC + range; index lists copied out per sub-batch)
//       host: ordered pitch / yaw sums over the visible rays (asin / acos: host; sharded over the ranks of a multi-GPU run)
//       K_E2 restricted to the E1 rows -> rows, popc       (FOV' + offset walk + BiRC)
//       K_own sweep in id order                            (ownership, cpp:500-528)
//   viewpointsPrune (cpp:620-710)
//       host: qualify, std::sort over the real unordered_map order; K_safety (nearCheck of every candidate); radius
//             search; the sequential absorb flag chain and gravitation pulls (cpp:712-810); K_safety
//             (viewpointSafetyCheck of the pulled positions)
//       K_E3 over all gravitated viewpoints at once -> host ordered offset sums (cpp:817-886)
//       K_E4 over the live viewpoints in unordered_map order -> popc -> K_own sweep
//   uncovered-area rounds (cpp:206-330)
//       host: 8 candidates per uncovered point (RNG + libm) -> K_safety over all of them -> getPose batch -> prune
//
// There is no CPU implementation of any K_* stage in this library: without a CUDA device every
// evaluating call fails with FCVM_ENODEVICE.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <thread>
#include <sched.h>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/fcvm.h"
#include "fcvm_comm.h"
#include "fcvm_geom.h"
#include "fcvm_hostutil.h"
#include "fcvm_kernels.h"

namespace fcvm {

static thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
const char* last_error_cstr() { return g_err.c_str(); }

struct Pose5 { double v[5]; };
// visible sets of a batch as ascending index lists: row i = idx[offs[i] .. offs[i+1])
struct CsrOut {
  std::vector<long long> offs;
  const uint32_t* idx = nullptr;
  // optional caller buffer (ideally page-locked, fcvm_host_alloc): the streamed path copies the index lists straight into it
  uint32_t* user_idx = nullptr;
  long long user_cap = 0;
  bool in_user = false;      // the first min(total, user_cap) indices are in user_idx, nothing in idx
  // rows [lo, hi) only (hi < 0: all): a rank of a sharded run extracts the visible sets of its own viewpoints;
  // offs then has hi - lo + 1 entries, row lo first
  long long lo = 0, hi = -1;
};
struct FinalVp { int vp_id; Pose5 pose; int vox_count; };
struct TraceEv { int stage; int vp_id; Pose5 pose; std::vector<uint32_t> row; };
struct PtNormal { float x, y, z, nx, ny, nz; };

struct D3Less {   // Vector3dCompare0, viewpoint_manager.h:53-63
  bool operator()(const D3& a, const D3& b) const {
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z < b.z;
  }
};
struct D3Hash {   // VectorHashEigenVector3d, viewpoint_manager.h:65-77
  std::size_t operator()(const D3& v) const {
    std::size_t seed = 0;
    const double c[3] = {v.x, v.y, v.z};
    for (int i = 0; i < 3; ++i) seed ^= std::hash<double>()(c[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};
struct D3Eq {
  bool operator()(const D3& a, const D3& b) const { return a.x == b.x && a.y == b.y && a.z == b.z; }
};

}  // namespace fcvm

using namespace fcvm;

struct fcvm_handle {
  fcvm_params prm;
  // derived at init (viewpoint_manager.cpp:49-55)
  double pitch_upper, pitch_lower, fov_base;
  Camera cam;
  int threads = 1;
  bool host_sums = false, own_sweep = false;   // A/B switches, taken from the environment when the handle is created

  // TaskState (viewpoint_manager.h:102-125)
  bool map_flag = false, model_flag = false, viewpoints_flag = false, normal_flag = false;

  // map
  HostGrid grid;                               // geometry; the bricks live on the device (downloaded on demand for the parity tap)
  DevBuf<unsigned long long> d_bricks;          // [guard | bricks | guard]; the guard words read 'occupied' (see occ_fast)
  DevBuf<int> d_snap_vox;                       // vps_voxcount before an uncovered-area round (update_impl)
  DevBuf<int> d_sat;                            // summed-area table over 8^3-voxel cells (GridDesc::sat)
  bool sat_built = false;                       // false: the grid is too large for the table's 32-bit offsets
  bool use_sat = true;                          // FCVM_NO_PFREE=1 switches the half-ray certificate off (A/B)
  size_t brick_guard = 0;                       // guard words on each side
  int guard_steps = 0;                          // longest half-ray (voxel steps) the fast path accepts
  // map cloud on the device + its uniform cell index (nearCheck), built by fcvm_map.cu
  DevBuf<float> d_map_xyz;
  DevBuf<float4> d_map_sorted;
  DevBuf<long long> d_cell_start;
  DevBuf<int> d_cell_count;
  DevBuf<float> d_minmax;
  CellGeom cells;
  int64_t n_map = 0;
  // batched safety checks
  DevBuf<double> d_gather;
  DevBuf<float> d_cand;
  DevBuf<uint8_t> d_ok;
  DevBuf<unsigned long long> d_looks;
  // model
  std::vector<float> model;   // xyz triples
  int64_t P = 0, W = 0;       // W = row stride in words (multiple of 4)
  DevBuf<float4> d_points;
  // normals
  std::map<D3, D3, D3Less> pt_normal_pairs;

  // VPI / MCI (viewpoint_manager.h:128-158)
  std::vector<Pose5> vps_pose;
  std::vector<int> vps_voxcount, vps_contri;   // host mirrors; voxcount is authoritative on the device during a stage
  int last_vps_num = 0;
  DevBuf<int> d_cover_contrib, d_contrib_id, d_voxcount;
  DevBuf<int> d_snap_contrib, d_snap_id;       // MCI snapshot of an uncovered-area round (cpp:272-274)
  size_t voxcount_n = 0;      // logical length of vps_voxcount
  std::vector<FinalVp> final_vps;
  // ViewpointsPrune maps (viewpoint_manager.h:160-178).  They are members that are only ever
  // clear()ed, never reconstructed (viewpoint_manager.cpp:622-626, h:169-177), so their bucket
  // arrays - and with them the iteration order that decides the prune order among equal counts,
  // the E4 processing order and the order of final_vps_ - carry over from call to
  // call.  Only the key order of idx_viewpoints_ is observable, so it maps id -> slot here.
  std::unordered_map<int, int> idx_slot;

  // device scratch for a batch
  DevBuf<VpRec> d_recs;
  DevBuf<RayEnd> d_vp_ends, d_pt_ends;         // record positions / model points divided by the resolution (k_ray_ends)
  DevBuf<uint32_t> d_rows1, d_rows2;
  DevBuf<int> d_count;
  DevBuf<int4> d_sched;
  DevBuf<unsigned long long> d_counters;
  DevBuf<long long> d_offsets;
  DevBuf<uint32_t> d_indices;
  PinBuf<uint32_t> h_indices;
  PinBuf<long long> h_offsets;
  PinBuf<uint32_t> h_rows;
  PinBuf<int> h_count;
  PinBuf<VpRec> h_recs;
  PinBuf<int> h_own_ids;
  PinBuf<int4> h_own_sched;
  int64_t resident_n = 0;
  int64_t resident_W = 0;
  bool resident_e1 = false;                    // the first rows buffer holds the E1 rows of the resident batch
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;          // D2H of finished sub-batches while the next one is evaluated
  std::vector<cudaStream_t> sub_streams;       // sub-batch k of an evaluator pass runs on sub_streams[k]: priorities fall with k, so the block
                                               // scheduler finishes the sub-batches in order while the tail of one overlaps the head of the next
  cudaEvent_t sub_start = nullptr;             // "the batch's inputs are on the device" (recorded on the main stream)
  int64_t rows_cap_bytes = 4ll << 30;          // manager-level rows buffer cap (ensure_device: from the device's memory size)
  int pri_lo = 0, pri_hi = 0;                  // cudaDeviceGetStreamPriorityRange: least, greatest
  cudaStream_t aux_stream = nullptr;           // pose-sum pipeline of a finished E1 sub-batch while the next one is evaluated (high priority)
  std::vector<cudaEvent_t> pipe_ev;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_kernel_ms = 0.0;
  bool timed = false;
  bool device_ok = false;
  // device time per evaluator stage: event pairs recorded around every k_eval launch, read after the next stream sync
  struct StageEv { cudaEvent_t a, b; int stage; };
  std::vector<StageEv> stage_ev_pending, stage_ev_free;
  StageEv span_ev{nullptr, nullptr, 0};         // the pair of a pass made of overlapping sub-batches (launch_stage, span)
  double stage_ms[4] = {0, 0, 0, 0};

  // ownership as a per-point arg-max (k_own_best / k_own_apply)
  DevBuf<unsigned long long> d_best;
  DevBuf<int> d_ids;
  // pitch / yaw sums of getPose on the device (fcvm_pose.cu)
  DevBuf<PoseAux> d_aux;
  PinBuf<PoseAux> h_aux;
  AsyncBuf<double> d_valp, d_valy, d_argp, d_argy;     // stream-ordered: they grow while the next E1 sub-batch is running
  AsyncBuf<unsigned long long> d_keyp, d_keyy;
  AsyncBuf<uint32_t> d_pidx;
  DevBuf<double> d_sums;
  DevBuf<unsigned long long> d_nflag;
  PinBuf<double> h_sums;
  PinBuf<unsigned long long> h_nflag;
  static constexpr int kRing = 4;                      // page-locked staging ring for the flagged arguments / their libm values
  static constexpr size_t kPiece = 1u << 20;           // doubles per slot (8 MB)
  PinBuf<double> h_ring[kRing];
  cudaEvent_t ring_d2h[kRing] = {nullptr, nullptr, nullptr, nullptr}, ring_h2d[kRing] = {nullptr, nullptr, nullptr, nullptr};
  std::unique_ptr<WorkerPool> pool;
  int64_t pose_rays = 0, pose_flagged = 0;     // rays summed on the device / of which sent to the host's libm

  // wall-clock split of the manager-level calls (FCVM_PROFILE=1 prints it after fcvm_update; fcvm_get_timers)
  double t_eval = 0, t_own = 0, t_init = 0, t_update = 0, t_unc = 0, t_sums = 0, t_chain = 0, t_safety = 0, t_pose = 0, t_comm = 0;
  // taps
  bool trace_on = false;
  std::vector<TraceEv> trace;
  int64_t counters[FCVM_NCOUNTERS] = {0, 0, 0, 0, 0, 0, 0, 0};
  int64_t unc_calls = 0, sample_calls = 0;
  int64_t bounds_violations = 0;

  // sharding
  int rank = 0, world = 1;
  fcvm_comm_fn comm = nullptr;
  void* comm_user = nullptr;
  NcclComm* nccl = nullptr;                    // owned when set by fcvm_set_shard_nccl
  int64_t comm_stats[4] = {0, 0, 0, 0};        // all-gathers, bytes gathered, all-reduces, bytes reduced
};

namespace fcvm {

// f(i) for i in [0, n) on the handle's persistent worker pool (created on first use); n <= block runs on the caller.
// (parallel_for of fcvm_hostutil.h spawns its threads per call, ~0.3 ms for 16: most of a shipped scene's per-call budget.)
template <class F>
static void pfor(fcvm_handle* h, int64_t n, int64_t block, F&& f) {
  if (n <= 0) return;
  if (h->threads <= 1 || n <= block) {
    for (int64_t i = 0; i < n; ++i) f(i);
    return;
  }
  if (!h->pool) h->pool.reset(new WorkerPool(h->threads));
  h->pool->run(n, block, f);
}

// ------------------------------------------------------------------------------------------------
// device bring-up (lazy: creating a handle and the host-only helpers work without a GPU)
// ------------------------------------------------------------------------------------------------
// Every public entry that touches CUDA passes through here first: the handle's device becomes the calling thread's
// current device (a process may hold handles on several devices, driven from several threads).
static int ensure_device(fcvm_handle* h) {
  if (h->device_ok) {
    CU(cudaSetDevice(h->prm.device));
    return FCVM_OK;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(FCVM_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                    " (this library has no CPU fallback)");
  CU(cudaSetDevice(h->prm.device));
  CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    int lo_pri = 0, hi_pri = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    // the copy stream also runs the compaction kernels of finished sub-batches: they must not queue behind the evaluator's CTAs
    CU(cudaStreamCreateWithPriority(&h->copy_stream, cudaStreamNonBlocking, hi_pri));
    CU(cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, hi_pri));
    h->pri_lo = lo_pri; h->pri_hi = hi_pri;
    CU(cudaEventCreateWithFlags(&h->sub_start, cudaEventDisableTiming));
  }
  {
    // the stream-ordered allocator keeps what the pose-sum pipeline frees (default: returned to the OS at the next sync)
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, h->prm.device));
    uint64_t keep = ~0ull;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    h->rows_cap_bytes = std::max<int64_t>(4ll << 30, std::min<int64_t>(32ll << 30, (int64_t)(total_b / 6)));
  }
  CU(cudaEventCreate(&h->ev0));
  CU(cudaEventCreate(&h->ev1));
  CU(h->d_counters.reserve(8));
  CU(cudaMemset(h->d_counters.p, 0, 8 * sizeof(unsigned long long)));
  {
    // Once per device and process: load the path's kernels now (the driver loads a kernel at its first launch otherwise, and a
    // first updateViewpoints launches ~25 different ones), and let the stream-ordered pool own a first block of memory so that
    // the handle's buffers are carved out of it instead of costing a driver allocation each.
    static std::atomic<bool> warmed[64];
    const int dev = h->prm.device;
    if (dev < 0 || dev >= 64 || !warmed[dev].exchange(true)) {
      CU(preload_eval_kernels());
      CU(preload_pose_kernels());
      CU(preload_map_kernels());
      void* block = nullptr;
      if (cudaMallocAsync(&block, (size_t)64 << 20, h->stream) == cudaSuccess) CU(cudaFreeAsync(block, h->stream));
      else (void)cudaGetLastError();
      CU(cudaStreamSynchronize(h->stream));
    }
  }
  if (!h->pool) h->pool.reset(new WorkerPool(h->threads));
  h->device_ok = true;
  return FCVM_OK;
}

// Stream of sub-batch k: one below the pose-sum pipeline's priority for k = 0, falling with k (numerically rising) down to the
// device's least priority.  Kernels of several sub-batches are resident at once; the block scheduler dispatches the CTAs of the
// highest-priority kernel first, so sub-batch k completes before k + 1 although nothing serialises them, and the SMs that the
// tail of k leaves idle are taken by k + 1 at once.
static int sub_stream(fcvm_handle* h, size_t k, cudaStream_t& out) {
  while (h->sub_streams.size() <= k) {
    const int pri = std::min(h->pri_lo, h->pri_hi + 1 + (int)h->sub_streams.size());
    cudaStream_t st;
    CU(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, pri));
    h->sub_streams.push_back(st);
  }
  out = h->sub_streams[k];
  return FCVM_OK;
}

static void fill_grid_desc(const fcvm_handle* h, GridDesc& g) {
  const HostGrid& hg = h->grid;
  g.bricks = h->d_bricks.p + h->brick_guard;
  g.guard_steps = h->guard_steps;
  g.sat = (h->use_sat && h->sat_built) ? h->d_sat.p : nullptr;
  g.sat_ny = hg.ns[1] / FCVM_SAT_BRICKS + 1; g.sat_nz = hg.ns[2] / FCVM_SAT_BRICKS + 1;
  g.nx = hg.dims[0]; g.ny = hg.dims[1]; g.nz = hg.dims[2];
  g.nbx = hg.ns[0]; g.nby = hg.ns[1]; g.nbz = hg.ns[2];
  g.nly = hg.ns[1] >> 1; g.nlz = hg.ns[2] >> 2;
  g.nbricks = (int)hg.nbricks();
  g.res = hg.res;
  g.res_rcp = 1.0 / hg.res;
  // RayCaster::setParams (raycast.cpp:341-345): offset_ = half_ - origin / resolution_
  g.offx = 0.5 - hg.origin[0] / hg.res;
  g.offy = 0.5 - hg.origin[1] / hg.res;
  g.offz = 0.5 - hg.origin[2] / hg.res;
  g.orgx = hg.origin[0]; g.orgy = hg.origin[1]; g.orgz = hg.origin[2];
}

static void fill_fov_const(const fcvm_handle* h, FovConst& f) {
  f.max_dist = h->prm.max_dist;
  const double md2 = h->prm.max_dist * h->prm.max_dist;
  f.md2_lo = md2 * (1.0 - 1e-12);
  f.md2_hi = md2 * (1.0 + 1e-12);
  f.plane_eps = 1e-10 * std::max(h->prm.max_dist, 1.0);
  f.visible_range = h->prm.visible_range;
  const double vr2 = h->prm.visible_range * h->prm.visible_range;
  f.vr2_lo = vr2 * (1.0 - 1e-12);
  f.vr2_hi = vr2 * (1.0 + 1e-12);
}

// choose the CTA decomposition: tiles of points x chunks of viewpoints, aiming at >= 4 waves of
// 148 SMs x 2 resident CTAs when the problem is large enough and at enough CTAs to fill the chip
// when it is small.
static void choose_tiling(int64_t P, int64_t V, int& tile_pts, int& vp_per_cta) {
  tile_pts = P >= 8 * EVAL_MAX_TILE ? EVAL_MAX_TILE : (P >= 4096 ? 1024 : 512);
  const int64_t tiles = (P + tile_pts - 1) / tile_pts;
  const int64_t target_ctas = 148 * EVAL_MIN_CTAS * EVAL_WAVES;   // many short CTAs: tiles far from a viewpoint chunk finish early
  int64_t chunks = std::max<int64_t>(1, std::min<int64_t>((target_ctas + tiles - 1) / tiles, (V + EVAL_WARPS - 1) / EVAL_WARPS));
  int64_t per = (V + chunks - 1) / chunks;
  per = std::max<int64_t>(per, 1);
  while ((V + per - 1) / per > 65535) per *= 2;
  vp_per_cta = (int)per;
}

// One evaluator launch over recs[0..n) already on the device -> rows (zeroed here).  Adds the
// kernel's counters to the handle when `count_stats`.
// span: 0 = a launch of its own; 1 / 2 / 3 = first / middle / last sub-batch of one pass whose kernels overlap on priority
// streams - the per-stage device time (fcvm_get_timers) then runs from the first kernel's start to the last one's end.
static int launch_stage(fcvm_handle* h, int stage, const VpRec* d_recs, int64_t n, uint32_t* d_rows, const uint32_t* d_restrict,
                        cudaStream_t s, bool count_stats, int span = 0) {
  EvalArgs a;
  fill_grid_desc(h, a.grid);
  fill_fov_const(h, a.fov);
  a.points = h->d_points.p;
  a.P = h->P;
  a.recs = d_recs;
  a.V = n;
  // the records' positions / resolution, once per viewpoint instead of once per ray (same slot as the record in its buffer)
  const size_t rec0 = (size_t)(d_recs - h->d_recs.p);
  if (d_recs < h->d_recs.p || rec0 + (size_t)n > h->d_recs.cap) return fail(FCVM_EINTERNAL, "records outside the handle's buffer");
  CU(h->d_vp_ends.reserve(h->d_recs.cap));
  CU(launch_ray_ends(reinterpret_cast<const double*>(d_recs), (int)(sizeof(VpRec) / sizeof(double)), nullptr, n, h->grid.res, h->d_vp_ends.p + rec0, s));
  a.vp_ends = h->d_vp_ends.p + rec0;
  a.pt_ends = h->d_pt_ends.p;
  a.rows = d_rows;
  a.restrict_rows = d_restrict;
  a.words = h->W;
  choose_tiling(h->P, n, a.tile_pts, a.vp_per_cta);
  a.counters = count_stats ? h->d_counters.p : nullptr;
  CU(cudaMemsetAsync(d_rows, 0, (size_t)n * h->W * sizeof(uint32_t), s));
  CU(cudaEventRecord(h->ev0, s));      // ev0..ev1 bracket the visibility kernel alone (roofline timing)
  fcvm_handle::StageEv se{nullptr, nullptr, stage};
  if (count_stats && span <= 1) {      // per-stage device time of the manager-level calls (fcvm_get_timers)
    if (!h->stage_ev_free.empty()) { se = h->stage_ev_free.back(); h->stage_ev_free.pop_back(); se.stage = stage; }
    else { CU(cudaEventCreate(&se.a)); CU(cudaEventCreate(&se.b)); }
    CU(cudaEventRecord(se.a, s));
    if (span == 1) h->span_ev = se;
  }
  CU(launch_eval(stage, a, s));
  if (count_stats && span == 0) { CU(cudaEventRecord(se.b, s)); h->stage_ev_pending.push_back(se); }
  if (count_stats && span == 3) { CU(cudaEventRecord(h->span_ev.b, s)); h->stage_ev_pending.push_back(h->span_ev); }
  CU(cudaEventRecord(h->ev1, s));
  h->timed = true;
  h->counters[5] += 2;                 // k_ray_ends + k_eval
  h->counters[6] += n;
  return FCVM_OK;
}

static int safety_batch(fcvm_handle* h, const float* xyz, int64_t n, int mode, std::vector<uint8_t>& ok);
static inline int safety_batch(fcvm_handle* h, const std::vector<float>& xyz, int mode, std::vector<uint8_t>& ok) {
  return safety_batch(h, xyz.data(), (int64_t)xyz.size() / 3, mode, ok);
}

static int fetch_counters(fcvm_handle* h) {
  unsigned long long c[8];
  CU(cudaMemcpyAsync(c, h->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemsetAsync(h->d_counters.p, 0, sizeof(c), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  for (int k = 0; k < 5; ++k) h->counters[k] += (int64_t)c[k];
  h->bounds_violations += (int64_t)c[5];   // only FCVM_DEBUG_BOUNDS builds ever write slot 5
  for (auto& se : h->stage_ev_pending) {
    float ms = 0.f;
    if (cudaEventQuery(se.b) == cudaSuccess && cudaEventElapsedTime(&ms, se.a, se.b) == cudaSuccess) h->stage_ms[se.stage - 1] += ms;
    else (void)cudaGetLastError();       // launched on a caller's stream that has not finished: not accounted
    h->stage_ev_free.push_back(se);
  }
  h->stage_ev_pending.clear();
  return FCVM_OK;
}

// one collective on the handle's stream (fcvm_comm.h); n = bytes per rank (all-gather) or int64 elements (all-reduce)
static int do_comm(fcvm_handle* h, int32_t op, void* buf_dev, int64_t n) {
  Stopwatch sw(h->t_comm);
  if (!h->comm) return fail(FCVM_EINVAL, "world > 1 but no collectives set (fcvm_set_shard / fcvm_set_shard_nccl)");
  if (h->comm(h->comm_user, op, buf_dev, n, h->stream) != 0) return fail(FCVM_EINTERNAL, std::string("collective failed: ") + last_error_cstr());
  if (op == FCVM_COMM_ALLGATHER) { h->comm_stats[0] += 1; h->comm_stats[1] += n * h->world; }
  else { h->comm_stats[2] += 1; h->comm_stats[3] += n * 8; }
  if (env().profile >= 2) CU(cudaStreamSynchronize(h->stream));    // attribute the wait to the collective, not to whoever syncs next
  return FCVM_OK;
}

constexpr int64_t kPipeMinRows = 1024;   // smallest sub-batch worth a launch of its own

// rows per batch chunk.  The evaluator-level entries (rows / index lists go to the host) keep one rows buffer <= 4 GiB, which
// bounds their page-locked staging; the manager-level batches (rows never leave the device) use a sixth of the device's memory
// per buffer, at most 32 GiB - two such buffers live on the device - so that configs[4] (200 000 viewpoints x 125 KB) is ONE
// chunk on a 180 GB B200: one ownership pass, no pipeline refill between chunks, and a sharded run's GPUs each get a large block.
static int64_t chunk_rows(const fcvm_handle* h, int64_t n, bool manager = false) {
  const int64_t cap = manager ? h->rows_cap_bytes : (int64_t)(4ll << 30);
  const int64_t per = std::max<int64_t>(1, cap / (h->W * 4));
  return std::min(n, per);
}

// Evaluate `stage` for recs (host) [0..n): upload, launch, optional exchange across ranks, download
// rows into rows_out (n x W, host) and/or counts.  d_rows_keep: leave rows on the device in this
// buffer (must hold n x W words).
// gather_rows (sharded runs only): true = every rank ends up with all rows and counts (the evaluator-level API); false = rows
// and counts of the other ranks' blocks stay zero (the manager: nothing it does needs foreign rows, see ownership()).
// f(j) for every set bit j of a bitset row, ascending
template <class F>
static void for_each_bit(const uint32_t* row, int64_t words, F&& f) {
  for (int64_t w = 0; w < words; ++w) {
    uint32_t x = row[w];
    while (x) {
      const int b = __builtin_ctz(x);
      x &= x - 1;
      f((int)(w * 32 + b));
    }
  }
}

constexpr int64_t kSmallRowsWords = 16384;      // 64 KB of bitset rows: below this a batch is summed / listed on the host (one wait)
static int eval_batch(fcvm_handle* h, int stage, const VpRec* recs, int64_t n, const uint32_t* restrict_host, uint32_t* rows_out,
                      int* count_out, DevBuf<uint32_t>& d_rows, bool restrict_resident = false, CsrOut* csr = nullptr, bool gather_rows = true) {
  int rc = ensure_device(h);
  if (rc) return rc;
  if (n == 0) return FCVM_OK;
  Stopwatch sw(h->t_eval);
  struct Verbose {     // FCVM_PROFILE=2: one line per evaluator batch
    int stage; int64_t n; std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~Verbose() {
      if (env().profile >= 2) std::fprintf(stderr, "[fcvm]   eval_batch E%d n=%lld: %.2f ms\n", stage, (long long)n,
                           std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
  } verbose{stage, n};
  if (env().profile >= 3) { cudaStreamSynchronize(h->stream); }
  const int64_t W = h->W;
  // pad the batch to a multiple of the world size so that shards are equal (NCCL all-gather)
  const int64_t per_rank = (n + h->world - 1) / h->world;
  const int64_t n_pad = per_rank * h->world;
  CU(h->d_recs.reserve((size_t)n_pad));
  CU(d_rows.reserve((size_t)n_pad * W));
  CU(h->d_count.reserve((size_t)n_pad));
  CU(cudaMemcpyAsync(h->d_recs.p, recs, (size_t)n * sizeof(VpRec), cudaMemcpyHostToDevice, h->stream));
  const uint32_t* d_restrict = nullptr;
  if (stage == FCVM_STAGE_E2) {
    DevBuf<uint32_t>& other = (&d_rows == &h->d_rows1) ? h->d_rows2 : h->d_rows1;
    if (restrict_resident) {
      // the E1 rows of this very batch are still in the other rows buffer
      if (other.cap < (size_t)n_pad * W) return fail(FCVM_EINTERNAL, "resident E1 rows missing");
    } else {
      if (!restrict_host) return fail(FCVM_EINVAL, "E2 needs the E1 rows to restrict to");
      CU(other.reserve((size_t)n_pad * W));
      CU(cudaMemcpyAsync(other.p, restrict_host, (size_t)n * W * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    }
    d_restrict = other.p;
  }
  const int64_t lo = std::min(n, per_rank * h->rank), hi = std::min(n, per_rank * (h->rank + 1));
  h->timed = false;
  const bool sharded = h->world > 1;
  if (sharded && gather_rows) CU(cudaMemsetAsync(d_rows.p, 0, (size_t)n_pad * W * sizeof(uint32_t), h->stream));
  // Single GPU with rows wanted on the host: evaluate in sub-batches (FCVM_PIPE_SUBS of them) and copy each finished
  // one out on the copy stream while the next is being evaluated (the rows dominate the transfer).
  bool rows_streamed = false;
  // Single GPU, visible sets wanted as CSR (what the manager consumes): evaluate in sub-batches; while sub-batch k + 1 is
  // being evaluated, the rows of sub-batch k are counted, scanned, compacted to index lists and copied out on the copy
  // stream.  The host only waits for the (tiny) offsets of a finished sub-batch, with the next kernel already queued.
  bool csr_streamed = false;
  if (csr && !rows_out && h->world == 1 && n >= 2 * kPipeMinRows) {
    // Sub-batches of decreasing size (55 / 30 / 15 %): what cannot overlap is the compaction and copy-out of the LAST one's
    // lists, while every extra launch costs a load-balance tail of its own.  FCVM_PIPE_SUBS=k: k equal sub-batches instead.
    std::vector<std::pair<int64_t, int64_t>> subs;      // (first row, rows)
    if (env().pipe_subs) {
      const int64_t sub = std::max<int64_t>(kPipeMinRows, (n + env().pipe_subs - 1) / env().pipe_subs);
      for (int64_t s0 = 0; s0 < n; s0 += sub) subs.emplace_back(s0, std::min(sub, n - s0));
    } else if (n >= 8 * kPipeMinRows) {
      static const std::vector<int> kFracs = {55, 30, 15};
      const std::vector<int>& fr = env().pipe_fracs.empty() ? kFracs : env().pipe_fracs;
      int64_t s0 = 0;
      for (size_t i = 0; i < fr.size() && s0 < n; ++i) {
        const int64_t sn = i + 1 == fr.size() ? n - s0 : std::min<int64_t>(n - s0, (int64_t)((n * fr[i] / 100 + 31) & ~31ll));
        if (sn > 0) subs.emplace_back(s0, sn);
        s0 += sn;
      }
      if (s0 < n) subs.emplace_back(s0, n - s0);
    } else {
      const int64_t a = (n * 6 / 10 + 31) & ~31ll;
      subs = {{0, a}, {a, n - a}};
    }
    const size_t nsub = subs.size();
    CU(h->d_offsets.reserve((size_t)n + nsub));
    CU(h->h_offsets.reserve((size_t)n + nsub));
    while (h->pipe_ev.size() < 2 * nsub) {
      cudaEvent_t e;
      CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->pipe_ev.push_back(e);
    }
    size_t k = 0;
    CU(cudaEventRecord(h->sub_start, h->stream));
    for (; k < nsub; ++k) {      // everything of the evaluation side is queued up front, one priority stream per sub-batch
      const int64_t s0 = subs[k].first, sn = subs[k].second;
      cudaStream_t ss;
      rc = sub_stream(h, k, ss);
      if (rc) return rc;
      CU(cudaStreamWaitEvent(ss, h->sub_start, 0));
      if ((int)k >= std::max(1, h->pri_lo - h->pri_hi)) CU(cudaStreamWaitEvent(ss, h->pipe_ev[2 * (k - 1)], 0));   // out of priority levels
      rc = launch_stage(h, stage, h->d_recs.p + s0, sn, d_rows.p + s0 * W, d_restrict ? d_restrict + s0 * W : nullptr, ss, true,
                        nsub == 1 ? 0 : (k == 0 ? 1 : (k + 1 == nsub ? 3 : 2)));
      if (rc) return rc;
      CU(launch_popc_rows(d_rows.p + s0 * W, sn, W, h->d_count.p + s0, ss));
      CU(launch_exscan(h->d_count.p + s0, sn, h->d_offsets.p + s0 + (int64_t)k, ss));
      h->counters[5] += 2;
      CU(cudaEventRecord(h->pipe_ev[2 * k], ss));
    }
    for (k = 0; k < nsub; ++k) CU(cudaStreamWaitEvent(h->stream, h->pipe_ev[2 * k], 0));    // the main stream continues after all of them
    const bool vprof = env().profile >= 3;
    auto tmark = [&](const char* what) { if (vprof) std::fprintf(stderr, "[fcvm]     %s at %.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - verbose.t0).count()); };
    tmark("kernels queued");
    csr->offs.assign((size_t)n + 1, 0);
    long long base = 0;
    for (k = 0; k < nsub; ++k) {
      const int64_t s0 = subs[k].first, sn = subs[k].second;
      long long* h_off = h->h_offsets.p + s0 + (int64_t)k;
      CU(cudaStreamWaitEvent(h->copy_stream, h->pipe_ev[2 * k], 0));
      CU(cudaMemcpyAsync(h_off, h->d_offsets.p + s0 + (int64_t)k, ((size_t)sn + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->copy_stream));
      CU(cudaEventRecord(h->pipe_ev[2 * k + 1], h->copy_stream));
      CU(cudaEventSynchronize(h->pipe_ev[2 * k + 1]));
      const long long ct = h_off[sn];
      tmark("sub-batch offsets on host");
      for (int64_t i = 0; i < sn; ++i) csr->offs[(size_t)(s0 + i + 1)] = base + h_off[i + 1];
      if (ct > 0) {
        const size_t need = (size_t)(base + ct);
        if (need > h->d_indices.cap || (!csr->user_idx && need > h->h_indices.cap)) {
          // grow once, from the density seen so far (a later shortfall would have to move what is already there)
          const size_t est = (size_t)((double)need * (double)n / (double)(s0 + sn) * 1.15) + 1024;
          if (base == 0) {
            CU(h->d_indices.reserve(est));
            if (!csr->user_idx) CU(h->h_indices.reserve(est));
          } else {
            CU(cudaStreamSynchronize(h->copy_stream));
            DevBuf<uint32_t> nd;
            CU(nd.reserve(est));
            h->d_indices = std::move(nd);
            if (!csr->user_idx) {
              PinBuf<uint32_t> nh;
              CU(nh.reserve(est));
              std::memcpy(nh.p, h->h_indices.p, (size_t)base * sizeof(uint32_t));
              h->h_indices = std::move(nh);
            }
          }
        }
        tmark("index buffers ready");
        CU(launch_compact_rows(d_rows.p + s0 * W, sn, W, h->d_offsets.p + s0 + (int64_t)k, h->d_indices.p + base, h->copy_stream));
        h->counters[5] += 1;
        if (csr->user_idx) {
          const long long fit = std::min<long long>(ct, csr->user_cap - base);   // a caller buffer that is too small keeps a valid prefix
          if (fit > 0)
            CU(cudaMemcpyAsync(csr->user_idx + base, h->d_indices.p + base, (size_t)fit * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->copy_stream));
        } else {
          CU(cudaMemcpyAsync(h->h_indices.p + base, h->d_indices.p + base, (size_t)ct * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->copy_stream));
        }
      }
      base += ct;
    }
    csr_streamed = true;
    csr->in_user = csr->user_idx != nullptr;
    h->timed = false;
  } else if (rows_out && h->world == 1 && n >= 2 * kPipeMinRows) {
    const int pipe_subs = env().pipe_subs ? env().pipe_subs : 4;
    const int64_t sub = std::max<int64_t>(kPipeMinRows, (n + pipe_subs - 1) / pipe_subs);
    size_t k = 0;
    for (int64_t s0 = 0; s0 < n; s0 += sub, ++k) {
      const int64_t sn = std::min(sub, n - s0);
      rc = launch_stage(h, stage, h->d_recs.p + s0, sn, d_rows.p + s0 * W, d_restrict ? d_restrict + s0 * W : nullptr, h->stream, true);
      if (rc) return rc;
      if (h->pipe_ev.size() <= k) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pipe_ev.push_back(e);
      }
      CU(cudaEventRecord(h->pipe_ev[k], h->stream));
      CU(cudaStreamWaitEvent(h->copy_stream, h->pipe_ev[k], 0));
      CU(cudaMemcpyAsync(rows_out + s0 * W, d_rows.p + s0 * W, (size_t)sn * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->copy_stream));
    }
    rows_streamed = true;
    h->timed = false;    // ev0/ev1 bracket only the last sub-batch
  } else if (hi > lo) {
    rc = launch_stage(h, stage, h->d_recs.p + lo, hi - lo, d_rows.p + lo * W, d_restrict ? d_restrict + lo * W : nullptr, h->stream, true);
    if (rc) return rc;
  }
  if (sharded && gather_rows) {
    rc = do_comm(h, FCVM_COMM_ALLGATHER, d_rows.p, per_rank * W * (int64_t)sizeof(uint32_t));
    if (rc) return rc;
  }
  if (csr_streamed) {
    // the sub-batches already produced counts, offsets and index lists (see above)
    if (rows_out) CU(cudaMemcpyAsync(rows_out, d_rows.p, (size_t)n * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (count_out) CU(cudaMemcpyAsync(count_out, h->d_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    rc = fetch_counters(h);   // synchronises the stream
    if (rc) return rc;
    if (env().profile >= 3) std::fprintf(stderr, "[fcvm]     main stream drained at %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - verbose.t0).count());
    CU(cudaStreamSynchronize(h->copy_stream));
    csr->idx = h->h_indices.p;
    if (!csr->in_user) h->h_indices.note_use();
    return FCVM_OK;
  }
  if (sharded && !gather_rows) {
    // only this rank's block was evaluated: the other rows of the buffer are stale, their counts read 0
    CU(cudaMemsetAsync(h->d_count.p, 0, (size_t)n_pad * sizeof(int), h->stream));
    CU(launch_popc_rows(d_rows.p + lo * W, hi - lo, W, h->d_count.p + lo, h->stream));
  } else {
    CU(launch_popc_rows(d_rows.p, n, W, h->d_count.p, h->stream));
  }
  h->counters[5] += 1;
  const int64_t c_lo = csr ? std::max<long long>(0, csr->lo) : 0, c_hi = csr ? (csr->hi < 0 ? n : std::min<long long>(n, csr->hi)) : 0;
  const int64_t c_n = std::max<int64_t>(0, c_hi - c_lo);
  if (csr && !rows_out && !sharded && n * W <= kSmallRowsWords) {
    // a small batch (E3 of a handful of gravitated viewpoints): the rows themselves go to the host with the counts, one wait
    // instead of "offsets, wait, lists, wait", and the index lists are read off the rows there
    CU(h->h_rows.reserve((size_t)(n * W)));
    CU(cudaMemcpyAsync(h->h_rows.p, d_rows.p, (size_t)(n * W) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    if (count_out) CU(cudaMemcpyAsync(count_out, h->d_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    rc = fetch_counters(h);   // synchronises the stream
    if (rc) return rc;
    h->h_rows.note_use();
    csr->offs.assign((size_t)c_n + 1, 0);
    long long total = 0;
    for (int64_t i = 0; i < c_n; ++i) {
      const uint32_t* row = h->h_rows.p + (c_lo + i) * W;
      for (int64_t w = 0; w < W; ++w) total += __builtin_popcount(row[w]);
      csr->offs[(size_t)i + 1] = total;
    }
    CU(h->h_indices.reserve((size_t)std::max<long long>(total, 1)));
    uint32_t* out = h->h_indices.p;
    for (int64_t i = 0; i < c_n; ++i) for_each_bit(h->h_rows.p + (c_lo + i) * W, W, [&](int j) { *out++ = (uint32_t)j; });
    csr->idx = h->h_indices.p;
    if (h->timed) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
      h->last_kernel_ms = ms;
    }
    return FCVM_OK;
  }
  if (csr) {
    CU(h->d_offsets.reserve((size_t)c_n + 1));
    CU(h->h_offsets.reserve((size_t)c_n + 1));
    CU(launch_exscan(h->d_count.p + c_lo, c_n, h->d_offsets.p, h->stream));
    h->counters[5] += 1;
    CU(cudaMemcpyAsync(h->h_offsets.p, h->d_offsets.p, ((size_t)c_n + 1) * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  }
  if (rows_out && !rows_streamed) CU(cudaMemcpyAsync(rows_out, d_rows.p, (size_t)n * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  if (count_out) CU(cudaMemcpyAsync(count_out, h->d_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  auto vmark = [&](const char* what) {
    if (env().profile >= 3)
      std::fprintf(stderr, "[fcvm]     E%d n=%lld: %s at %.2f ms\n", stage, (long long)n, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - verbose.t0).count());
  };
  vmark("queued");
  rc = fetch_counters(h);   // synchronises the stream
  if (rc) return rc;
  vmark("kernels done");
  if (rows_streamed) CU(cudaStreamSynchronize(h->copy_stream));
  if (csr) {
    csr->offs.assign(h->h_offsets.p, h->h_offsets.p + c_n + 1);
    const long long total = csr->offs[(size_t)c_n];
    CU(h->d_indices.reserve((size_t)std::max<long long>(total, 1)));
    vmark("device index buffer ready");
    CU(h->h_indices.reserve((size_t)std::max<long long>(total, 1)));
    vmark("host index buffer ready");
    if (total > 0) {
      CU(launch_compact_rows(d_rows.p + c_lo * W, c_n, W, h->d_offsets.p, h->d_indices.p, h->stream));
      h->counters[5] += 1;
      CU(cudaMemcpyAsync(h->h_indices.p, h->d_indices.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
      CU(cudaStreamSynchronize(h->stream));
    }
    vmark("index lists on the host");
    csr->idx = h->h_indices.p;
    h->h_indices.note_use();
    vmark("note_use");
  }
  if (h->timed) {
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_ms = ms;
  }
  return FCVM_OK;
}

// The ownership rule (viewpoint_manager.cpp:500-528, 589-617) for one stage: rows [0, n) of d_rows stand for viewpoint ids
// vp_ids[row]; `order` lists the row indices in processing order.  Of the rows, [lo, hi) are resident on this GPU with their
// popcounts in d_count[lo..hi) (a single GPU holds them all).  Default: per-point arg-max (k_own_best over the local rows,
// one all-reduce MAX of the P keys when sharded, k_own_apply); FCVM_OWN_SWEEP=1 on a single GPU: the ordered row sweep.
// count_host (n ints) receives every row's contribute_voxel_num.
static int ownership(fcvm_handle* h, const DevBuf<uint32_t>& d_rows, int64_t n, int64_t lo, int64_t hi, const std::vector<int>& order,
                     const std::vector<int>& vp_ids, int* count_host) {
  Stopwatch sw(h->t_own);
  if (n <= 0) return FCVM_OK;
  int rc;
  const int64_t per = (n + h->world - 1) / h->world, n_pad = per * h->world;
  if (h->world > 1) {
    // every rank needs every row's contribute_voxel_num (vps_contri, the keys' decode): 4 bytes per viewpoint
    rc = do_comm(h, FCVM_COMM_ALLGATHER, h->d_count.p, per * (int64_t)sizeof(int));
    if (rc) return rc;
  }
  (void)n_pad;
  CU(cudaMemcpyAsync(count_host, h->d_count.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  int max_id = -1;
  bool any = false;
  for (size_t t = 0; t < order.size(); ++t) {
    const int r = order[t];
    if (count_host[r] <= 0) continue;     // `if (contribute_voxel_num > 0)` cpp:500 / `== 0 return` cpp:586
    any = true;
    h->vps_contri[(size_t)vp_ids[(size_t)r]] = count_host[r];
    max_id = std::max(max_id, vp_ids[(size_t)r]);
  }
  if (!any) return FCVM_OK;
  if ((size_t)max_id >= h->voxcount_n) return fail(FCVM_EINTERNAL, "vps_voxcount shorter than a viewpoint id");

  if (h->own_sweep && h->world == 1) {
    std::vector<int4> sched;
    for (int r : order)
      if (count_host[r] > 0) sched.push_back(make_int4(r, vp_ids[(size_t)r], count_host[r], 0));
    CU(h->d_sched.reserve(sched.size()));
    CU(cudaMemcpyAsync(h->d_sched.p, sched.data(), sched.size() * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    CU(launch_own_sweep(d_rows.p, h->W, h->d_sched.p, (int)sched.size(), h->last_vps_num, h->P, h->d_cover_contrib.p, h->d_contrib_id.p,
                        h->d_voxcount.p, h->stream));
    h->counters[5] += 1;
    CU(cudaStreamSynchronize(h->stream));
    return FCVM_OK;
  }

  // id_at_pos[t] = viewpoint id of the t-th row in processing order; the local rows with a positive count, sorted by
  // key = (contri, ~position) descending, are what k_own_best walks
  // (page-locked staging owned by the handle: the copies below are still in flight when this function returns)
  CU(h->h_own_ids.reserve(order.size()));
  CU(h->h_own_sched.reserve(order.size()));
  int* id_at_pos = h->h_own_ids.p;
  int4* sched_buf = h->h_own_sched.p;
  size_t nsched = 0;
  for (size_t t = 0; t < order.size(); ++t) {
    const int r = order[t];
    id_at_pos[t] = vp_ids[(size_t)r];
    if (r >= lo && r < hi && count_host[r] > 0) sched_buf[nsched++] = make_int4(r, count_host[r], (int)~(uint32_t)t, 0);
  }
  std::sort(sched_buf, sched_buf + nsched, [](const int4& a, const int4& b) {
    return a.y != b.y ? a.y > b.y : (uint32_t)a.z > (uint32_t)b.z;
  });
  CU(h->d_ids.reserve(order.size()));
  CU(cudaMemcpyAsync(h->d_ids.p, id_at_pos, order.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  CU(h->d_sched.reserve(std::max<size_t>(nsched, 1)));
  if (nsched) CU(cudaMemcpyAsync(h->d_sched.p, sched_buf, nsched * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
  CU(h->d_best.reserve((size_t)h->P));
  CU(cudaMemsetAsync(h->d_best.p, 0, (size_t)h->P * sizeof(unsigned long long), h->stream));
  CU(launch_own_best(d_rows.p, h->W, h->d_sched.p, (int)nsched, h->P, h->d_best.p, h->stream));
  h->counters[5] += 1;
  if (h->world > 1) {
    rc = do_comm(h, FCVM_COMM_ALLREDUCE_MAX, h->d_best.p, h->P);     // keys are < 2^63: a signed MAX orders them like the unsigned one
    if (rc) return rc;
  }
  CU(launch_own_apply(h->d_best.p, h->d_ids.p, h->last_vps_num, h->P, h->d_cover_contrib.p, h->d_contrib_id.p, h->d_voxcount.p, h->stream));
  h->counters[5] += 1;
  return FCVM_OK;       // not synchronised: everything that reads the cover state is ordered on the same stream
}

// vps_voxcount lives on the device while stages run; these move it
static int voxcount_resize(fcvm_handle* h, size_t n) {   // grow with zeros, keep contents
  if (n <= h->voxcount_n) { h->voxcount_n = n; return FCVM_OK; }
  if (n > h->d_voxcount.cap) {
    DevBuf<int> nb;
    CU(nb.reserve(std::max(n, h->d_voxcount.cap * 2)));
    if (h->voxcount_n) CU(cudaMemcpyAsync(nb.p, h->d_voxcount.p, h->voxcount_n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));      // the old buffer is released next
    h->d_voxcount = std::move(nb);
  }
  CU(cudaMemsetAsync(h->d_voxcount.p + h->voxcount_n, 0, (n - h->voxcount_n) * sizeof(int), h->stream));
  h->voxcount_n = n;
  return FCVM_OK;
}
static int voxcount_download(fcvm_handle* h, std::vector<int>& out) {
  out.assign(h->voxcount_n, 0);
  if (h->voxcount_n) {
    CU(cudaMemcpyAsync(out.data(), h->d_voxcount.p, h->voxcount_n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
  }
  return FCVM_OK;
}
static int cover_reset(fcvm_handle* h) {   // MCI reset, cpp:178-183
  CU(h->d_cover_contrib.reserve((size_t)h->P));
  CU(h->d_contrib_id.reserve((size_t)h->P));
  CU(cudaMemsetAsync(h->d_cover_contrib.p, 0, (size_t)h->P * sizeof(int), h->stream));
  CU(cudaMemsetAsync(h->d_contrib_id.p, 0xff, (size_t)h->P * sizeof(int), h->stream));   // -1
  return FCVM_OK;
}

static void record(fcvm_handle* h, int stage, int vp_id, const Pose5& pose, const uint32_t* row) {
  TraceEv e;
  e.stage = stage; e.vp_id = vp_id; e.pose = pose;
  e.row.assign(row, row + h->W);
  h->trace.push_back(std::move(e));
}

// ------------------------------------------------------------------------------------------------
// viewpointSafetyCheck / nearCheck (viewpoint_manager.cpp:889-946)
// ------------------------------------------------------------------------------------------------
// Both run on the device for whole batches of candidate positions (k_safety, fcvm_map.cu): mode 0 = viewpointSafetyCheck
// (ground test, lattice of safety_check probes with the reference's early exit, nearCheck with vox_num = 2), mode 1 =
// nearCheck alone.  counters[7] advances by the number of safety_check probes the reference would have made.
static int safety_batch(fcvm_handle* h, const float* xyz, int64_t n, int mode, std::vector<uint8_t>& ok) {
  Stopwatch sw_safety(h->t_safety);
  ok.assign((size_t)n, 0);
  if (n == 0) return FCVM_OK;
  const fcvm_params& p = h->prm;
  CU(h->d_cand.reserve((size_t)3 * n));
  CU(h->d_ok.reserve((size_t)n));
  CU(h->d_looks.reserve(1));
  cudaStream_t s = h->stream;
  CU(cudaMemcpyAsync(h->d_cand.p, xyz, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->d_looks.p, 0, sizeof(unsigned long long), s));
  SafetyArgs a;
  for (int c = 0; c < 3; ++c) { a.geom.org[c] = h->grid.origin[c]; a.geom.dims[c] = h->grid.dims[c]; a.geom.nb[c] = h->grid.ns[c]; }
  a.geom.res_inv = h->grid.res_inv;
  a.bricks = h->d_bricks.p + h->brick_guard;
  a.sat = (h->use_sat && h->sat_built) ? h->d_sat.p : nullptr;
  a.sat_ny = h->grid.ns[1] / FCVM_SAT_BRICKS + 1; a.sat_nz = h->grid.ns[2] / FCVM_SAT_BRICKS + 1;
  a.cells = h->cells;
  a.cell_start = h->d_cell_start.p;
  a.sorted = h->d_map_sorted.p;
  a.n_map = h->n_map;
  a.sr = p.safe_radius;
  a.res = h->grid.res;
  a.safe_voxel = (int)std::floor(0.5 * p.safe_radius / h->grid.res);
  a.z_ground = p.z_ground;
  a.ground_limit = p.ground_pos + p.safe_height;
  a.mode = mode;
  a.cand = h->d_cand.p;
  a.n = n;
  a.ok = h->d_ok.p;
  a.lookups = h->d_looks.p;
  CU(launch_safety(a, s));
  h->counters[5] += 1;
  unsigned long long looks = 0;
  CU(cudaMemcpyAsync(ok.data(), h->d_ok.p, (size_t)n, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(&looks, h->d_looks.p, sizeof(looks), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  h->counters[7] += (int64_t)looks;
  return FCVM_OK;
}

// Sharded runs: the per-viewpoint host reductions (libm over the visible rays) are the expensive host part, and the ranks
// of one box share its cores.  Each rank reduces only the viewpoints of its own shard and the results (k doubles per
// viewpoint) are exchanged with the same all-gather callback the bitset rows use.  vals holds n x k doubles; on return
// every rank has all of them.  Every value is produced by exactly one rank with the same code, so the outcome is
// bit-identical to the unsharded run.
static void shard_of(const fcvm_handle* h, int64_t n, int64_t& lo, int64_t& hi) {
  const int64_t per = (n + h->world - 1) / h->world;
  lo = std::min(n, per * h->rank);
  hi = std::min(n, per * (h->rank + 1));
}
// vals = n records of `bytes` each, of which this rank has filled its own block [lo, hi); on return every rank has all
static int gather_records(fcvm_handle* h, void* vals, int64_t n, int64_t bytes) {
  if (h->world <= 1 || n == 0) return FCVM_OK;
  const int64_t per = (n + h->world - 1) / h->world, n_pad = per * h->world;
  int64_t lo, hi;
  shard_of(h, n, lo, hi);
  const size_t words = (size_t)((n_pad * bytes + 7) / 8);
  CU(h->d_gather.reserve(words));
  char* dev = reinterpret_cast<char*>(h->d_gather.p);
  char* host = static_cast<char*>(vals);
  if (hi > lo) CU(cudaMemcpyAsync(dev + lo * bytes, host + lo * bytes, (size_t)((hi - lo) * bytes), cudaMemcpyHostToDevice, h->stream));
  {
    int rc = do_comm(h, FCVM_COMM_ALLGATHER, dev, per * bytes);
    if (rc) return rc;
  }
  CU(cudaMemcpyAsync(host, dev, (size_t)(n * bytes), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return FCVM_OK;
}
static int gather_doubles(fcvm_handle* h, double* vals, int64_t n, int k) { return gather_records(h, vals, n, (int64_t)k * (int64_t)sizeof(double)); }

// ------------------------------------------------------------------------------------------------
// getPose over a batch (viewpoint_manager.cpp:332-531).  ids[i] = viewpoint id of candidate i.
// poses_out[i] receives the returned pose; ownership and vps_contri/voxcount are updated.
// ------------------------------------------------------------------------------------------------
// rows of the last stage -> host (trace tap only); a sharded run first gathers the other ranks' blocks
static int fetch_rows_for_trace(fcvm_handle* h, DevBuf<uint32_t>& d_rows, int64_t n) {
  const int64_t W = h->W;
  if (h->world > 1) {
    const int64_t per = (n + h->world - 1) / h->world;
    int rc = do_comm(h, FCVM_COMM_ALLGATHER, d_rows.p, per * W * (int64_t)sizeof(uint32_t));
    if (rc) return rc;
  }
  CU(h->h_rows.reserve((size_t)n * W));
  CU(cudaMemcpyAsync(h->h_rows.p, d_rows.p, (size_t)n * W * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return FCVM_OK;
}

// per-viewpoint inputs of the pose sums (cpp:341-345): position, fov_sample_ray, fov_yaw_ray
struct PoseSeed { D3 vp, fs, fy; double init_yaw; };

// cpp:438-447: the averages of one viewpoint from its two ordered sums
static inline void finish_average(const fcvm_handle* h, double sum_pitch, double sum_yaw, long long free_num, double init_yaw, double& out_pitch, double& out_yaw) {
  const double pu = h->pitch_upper * M_PI / 180.0, pl = h->pitch_lower * M_PI / 180.0;
  double avg_pitch = sum_pitch / free_num;
  double avg_yaw = sum_yaw / free_num;
  avg_yaw = init_yaw + avg_yaw;
  warp_angle(avg_pitch);
  warp_angle(avg_yaw);
  if (avg_pitch > pu) avg_pitch = pu;
  if (avg_pitch < pl) avg_pitch = pl;
  out_pitch = avg_pitch;
  out_yaw = avg_yaw;
}

// the reference's loop body (cpp:398-421) for one viewpoint over its visible index list, all on the host (FCVM_HOST_SUMS=1)
static inline void host_pose_sums(const float* M, const PoseSeed& sd, const uint32_t* idx, long long cnt, double& sum_pitch, double& sum_yaw) {
  double avg_pitch = 0.0, avg_yaw = 0.0;
  for (long long k = 0; k < cnt; ++k) {
    const size_t j = idx[k];
    const D3 cand{M[3 * j], M[3 * j + 1], M[3 * j + 2]};
    const D3 d = sub3(cand, sd.vp);
    const D3 ray_dir = normalized3(d);
    // PitchYaw(ray_dir) (cpp:405-407): only the pitch is used here; the yaw (atan2) the reference also computes is dead
    double pitch = std::asin(ray_dir.z / (norm3(ray_dir) + 1e-3));
    warp_angle(pitch);
    avg_pitch += pitch;
    const double yaw_vec = std::min(std::max(dot3(sd.fs, ray_dir), -1.0), 1.0);
    double yawoff = std::acos(yaw_vec);
    // ray_dir_yaw.cross(fov_yaw_ray)[2] with the UNnormalised ray_dir_yaw (cpp:399-401, A.9)
    if (d.x * sd.fy.y - d.y * sd.fy.x < 0) yawoff = -yawoff;
    avg_yaw += yawoff;
  }
  sum_pitch = avg_pitch;
  sum_yaw = avg_yaw;
}

// E1 over rows [lo, hi) of a chunk (records and PoseAux already on the device) in sub-batches, each followed - on the
// high-priority aux stream, while the next sub-batch is being evaluated - by the device pose-sum pipeline:
//   rows -> popc -> offsets -> index lists -> k_pose_terms -> flagged arguments D2H -> glibc asin / acos on the host ->
//   H2D -> k_pose_patch -> k_pose_sum -> two sums per viewpoint D2H.
// free_num[i] / sums[2 i], sums[2 i + 1] are filled for i in [lo, hi).
static int e1_with_pose_sums(fcvm_handle* h, int64_t lo, int64_t hi, const std::vector<PoseSeed>& seed, std::vector<long long>& free_num,
                             std::vector<double>& sums) {
  const int64_t W = h->W, ln = hi - lo;
  if (ln <= 0) return FCVM_OK;
  int rc;
  if (ln * W <= kSmallRowsWords && !h->host_sums) {
    // A small batch (the <= 100 candidates of an uncovered-area round on the shipped scenes): the device pipeline's four
    // waits (offsets, flag counts, flagged arguments out and back, sums) cost more than its work.  The rows themselves go to
    // the host - at most 64 KB, one wait - and the sums are the reference's own loop there (host_pose_sums, every ray through
    // libm: identical by construction to what the device pipeline reproduces).
    rc = launch_stage(h, FCVM_STAGE_E1, h->d_recs.p + lo, ln, h->d_rows1.p + lo * W, nullptr, h->stream, true);
    if (rc) return rc;
    CU(h->h_rows.reserve((size_t)(ln * W)));
    CU(cudaMemcpyAsync(h->h_rows.p, h->d_rows1.p + lo * W, (size_t)(ln * W) * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->h_rows.note_use();
    Stopwatch sw(h->t_sums);
    const float* Mh = h->model.data();
    pfor(h, ln, 8, [&](int64_t i) {
      const uint32_t* row = h->h_rows.p + i * W;
      thread_local std::vector<uint32_t> idx;
      idx.clear();
      for_each_bit(row, W, [&](int j) { idx.push_back((uint32_t)j); });
      free_num[(size_t)(lo + i)] = (long long)idx.size();
      if (!idx.empty())
        host_pose_sums(Mh, seed[(size_t)(lo + i)], idx.data(), (long long)idx.size(), sums[(size_t)2 * (lo + i)], sums[(size_t)2 * (lo + i) + 1]);
    });
    return FCVM_OK;
  }
  const int levels = std::max(1, h->pri_lo - h->pri_hi);       // priority levels below the pose-sum pipeline's
  const int want = env().pipe_subs ? env().pipe_subs : (int)std::min<int64_t>(std::min(8, levels), std::max<int64_t>(1, ln / 6144));
  const int64_t sub = std::max<int64_t>(std::min<int64_t>(kPipeMinRows, ln), (ln + want - 1) / want);
  const size_t nsub = (size_t)((ln + sub - 1) / sub);
  CU(h->d_offsets.reserve((size_t)ln + nsub));
  CU(h->h_offsets.reserve((size_t)ln + nsub));
  CU(h->d_sums.reserve((size_t)2 * ln));
  CU(h->h_sums.reserve((size_t)2 * ln));
  CU(h->d_nflag.reserve(2));
  CU(h->h_nflag.reserve(2));
  while (h->pipe_ev.size() < nsub + 1) {
    cudaEvent_t e;
    CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->pipe_ev.push_back(e);
  }
  // everything of the evaluation side is queued up front, one priority stream per sub-batch (sub_stream): they finish in
  // order, the tail of one overlapping the head of the next
  size_t k = 0;
  CU(cudaEventRecord(h->sub_start, h->stream));
  for (int64_t s0 = 0; s0 < ln; s0 += sub, ++k) {
    const int64_t sn = std::min(sub, ln - s0), r0 = lo + s0;
    cudaStream_t ss;
    rc = sub_stream(h, k, ss);
    if (rc) return rc;
    CU(cudaStreamWaitEvent(ss, h->sub_start, 0));
    if ((int)k >= levels) CU(cudaStreamWaitEvent(ss, h->pipe_ev[k - 1], 0));     // out of priority levels: plain serialisation
    rc = launch_stage(h, FCVM_STAGE_E1, h->d_recs.p + r0, sn, h->d_rows1.p + r0 * W, nullptr, ss, true, nsub == 1 ? 0 : (k == 0 ? 1 : (k + 1 == nsub ? 3 : 2)));
    if (rc) return rc;
    CU(launch_popc_rows(h->d_rows1.p + r0 * W, sn, W, h->d_count.p + r0, ss));
    CU(launch_exscan(h->d_count.p + r0, sn, h->d_offsets.p + s0 + (int64_t)k, ss));
    h->counters[5] += 2;
    CU(cudaEventRecord(h->pipe_ev[k], ss));
  }
  for (k = 0; k < nsub; ++k) CU(cudaStreamWaitEvent(h->stream, h->pipe_ev[k], 0));      // the main stream continues after all of them
  const float* M = h->model.data();
  cudaStream_t ax = h->aux_stream;
  k = 0;
  for (int64_t s0 = 0; s0 < ln; s0 += sub, ++k) {
    const int64_t sn = std::min(sub, ln - s0), r0 = lo + s0;
    long long* h_off = h->h_offsets.p + s0 + (int64_t)k;
    const long long* d_off = h->d_offsets.p + s0 + (int64_t)k;
    CU(cudaStreamWaitEvent(ax, h->pipe_ev[k], 0));
    CU(cudaMemcpyAsync(h_off, d_off, ((size_t)sn + 1) * sizeof(long long), cudaMemcpyDeviceToHost, ax));
    CU(cudaStreamSynchronize(ax));                               // E1 of this sub-batch is done; the next one is already running
    Stopwatch sw(h->t_pose);
    const long long total = h_off[sn];
    for (int64_t i = 0; i < sn; ++i) free_num[(size_t)(r0 + i)] = h_off[i + 1] - h_off[i];
    if (total == 0) continue;
    const auto t_sub0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
      if (env().profile >= 3)
        std::fprintf(stderr, "[fcvm]     pose sums, sub-batch %zu (%lld rows, %lld rays): %s at %.2f ms\n", k, (long long)sn, total, what,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_sub0).count());
    };
    CU(h->d_pidx.reserve((size_t)total, ax));
    CU(launch_compact_rows(h->d_rows1.p + r0 * W, sn, W, d_off, h->d_pidx.p, ax));
    h->counters[5] += 1;
    if (h->host_sums) {
      // round-1 path, kept for A/B timing and as the cross-check of the device pipeline: index lists to the host, libm on every ray
      CU(h->h_indices.reserve((size_t)total));
      CU(cudaMemcpyAsync(h->h_indices.p, h->d_pidx.p, (size_t)total * sizeof(uint32_t), cudaMemcpyDeviceToHost, ax));
      CU(cudaStreamSynchronize(ax));
      h->h_indices.note_use();
      Stopwatch sw2(h->t_sums);
      pfor(h, sn, 4, [&](int64_t i) {
        const long long cnt = h_off[i + 1] - h_off[i];
        if (cnt > 0) host_pose_sums(M, seed[(size_t)(r0 + i)], h->h_indices.p + h_off[i], cnt, sums[(size_t)2 * (r0 + i)], sums[(size_t)2 * (r0 + i) + 1]);
      });
      continue;
    }
    CU(h->d_valp.reserve((size_t)total, ax));
    CU(h->d_valy.reserve((size_t)total, ax));
    unsigned long long cap = (unsigned long long)total / 10 + 4096;      // ~6.25 % of the rays are flagged per function
    unsigned long long np = 0, ny = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
      CU(h->d_keyp.reserve((size_t)cap, ax)); CU(h->d_argp.reserve((size_t)cap, ax));
      CU(h->d_keyy.reserve((size_t)cap, ax)); CU(h->d_argy.reserve((size_t)cap, ax));
      cap = std::min<unsigned long long>(std::min(h->d_keyp.cap, h->d_argp.cap), std::min(h->d_keyy.cap, h->d_argy.cap));
      CU(cudaMemsetAsync(h->d_nflag.p, 0, 2 * sizeof(unsigned long long), ax));
      CU(launch_pose_terms(h->d_points.p, h->d_aux.p + r0, d_off, h->d_pidx.p, (int)sn, h->d_valp.p, h->d_valy.p, h->d_keyp.p, h->d_argp.p,
                           h->d_keyy.p, h->d_argy.p, h->d_nflag.p, cap, ax));
      h->counters[5] += 1;
      CU(cudaMemcpyAsync(h->h_nflag.p, h->d_nflag.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ax));
      CU(cudaStreamSynchronize(ax));
      np = h->h_nflag.p[0]; ny = h->h_nflag.p[1];
      if (np <= cap && ny <= cap) break;
      if (attempt == 1) return fail(FCVM_EINTERNAL, "flagged-ray list overflow");
      cap = std::max(np, ny);                                             // an unusual batch (e.g. many coincident points): redo with room for all
    }
    mark("terms done");
    h->pose_rays += total;
    h->pose_flagged += (int64_t)(np + ny);
    if (np + ny > 0) {
      // The reference's own libm on exactly the arguments whose correctly rounded value glibc might not return.  Pieces of
      // kPiece arguments travel through a ring of page-locked slots: D2H of piece i + 1 (aux stream) and H2D of piece i - 1
      // (copy stream) run while the worker pool evaluates piece i; the values replace the arguments in place on the device.
      constexpr int R = fcvm_handle::kRing;
      const size_t PIECE = fcvm_handle::kPiece;
      if (!h->pool) h->pool.reset(new WorkerPool(h->threads));
      for (int q = 0; q < R; ++q) {
        CU(h->h_ring[q].reserve(std::min<size_t>(PIECE, (size_t)std::max(np, ny))));
        if (!h->ring_d2h[q]) { CU(cudaEventCreateWithFlags(&h->ring_d2h[q], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&h->ring_h2d[q], cudaEventDisableTiming)); }
      }
      struct Piece { int which; size_t off, cnt; };
      std::vector<Piece> pieces;
      for (size_t o = 0; o < (size_t)np; o += PIECE) pieces.push_back(Piece{0, o, std::min(PIECE, (size_t)np - o)});
      for (size_t o = 0; o < (size_t)ny; o += PIECE) pieces.push_back(Piece{1, o, std::min(PIECE, (size_t)ny - o)});
      std::vector<char> slot_used(R, 0);
      auto process = [&](size_t i) -> int {
        const Piece& pc = pieces[i];
        const int q = (int)(i % R);
        CU(cudaEventSynchronize(h->ring_d2h[q]));
        double* g = h->h_ring[q].p;
        {
          Stopwatch sw2(h->t_sums);
          const int64_t nb = ((int64_t)pc.cnt + 2047) / 2048;
          if (pc.which == 0)
            h->pool->run(nb, 1, [&](int64_t b) { const size_t e = std::min(pc.cnt, (size_t)(b + 1) * 2048); for (size_t j = (size_t)b * 2048; j < e; ++j) g[j] = std::asin(g[j]); });
          else
            h->pool->run(nb, 1, [&](int64_t b) { const size_t e = std::min(pc.cnt, (size_t)(b + 1) * 2048); for (size_t j = (size_t)b * 2048; j < e; ++j) g[j] = std::acos(g[j]); });
        }
        double* dst = (pc.which == 0 ? h->d_argp.p : h->d_argy.p) + pc.off;
        CU(cudaMemcpyAsync(dst, g, pc.cnt * sizeof(double), cudaMemcpyHostToDevice, h->copy_stream));
        CU(cudaEventRecord(h->ring_h2d[q], h->copy_stream));
        return FCVM_OK;
      };
      for (size_t i = 0; i < pieces.size(); ++i) {
        const Piece& pc = pieces[i];
        const int q = (int)(i % R);
        if (slot_used[q]) CU(cudaEventSynchronize(h->ring_h2d[q]));       // the slot's previous values have left
        slot_used[q] = 1;
        const double* src = (pc.which == 0 ? h->d_argp.p : h->d_argy.p) + pc.off;
        CU(cudaMemcpyAsync(h->h_ring[q].p, src, pc.cnt * sizeof(double), cudaMemcpyDeviceToHost, ax));
        CU(cudaEventRecord(h->ring_d2h[q], ax));
        if (i >= 1) { rc = process(i - 1); if (rc) return rc; }
      }
      rc = process(pieces.size() - 1);
      if (rc) return rc;
      for (int q = 0; q < R; ++q)
        if (slot_used[q]) CU(cudaStreamWaitEvent(ax, h->ring_h2d[q], 0));
      mark("libm values on their way back");
      CU(launch_pose_patch(h->d_keyp.p, h->d_argp.p, (long long)np, h->d_keyy.p, h->d_argy.p, (long long)ny, h->d_valp.p, h->d_valy.p, ax));
      h->counters[5] += 1;
    }
    CU(launch_pose_sum(d_off, (int)sn, h->d_valp.p, h->d_valy.p, h->d_sums.p + 2 * s0, ax));
    h->counters[5] += 1;
    CU(cudaMemcpyAsync(h->h_sums.p + 2 * s0, h->d_sums.p + 2 * s0, (size_t)2 * sn * sizeof(double), cudaMemcpyDeviceToHost, ax));
    CU(cudaStreamSynchronize(ax));
    mark("sums on the host");
    for (int64_t i = 0; i < sn; ++i) { sums[(size_t)2 * (r0 + i)] = h->h_sums.p[2 * (s0 + i)]; sums[(size_t)2 * (r0 + i) + 1] = h->h_sums.p[2 * (s0 + i) + 1]; }
  }
  return FCVM_OK;
}

static int get_pose_batch(fcvm_handle* h, const std::vector<D3>& pos, const std::vector<D3>& dir, int first_id, std::vector<Pose5>& poses_out) {
  const int64_t n = (int64_t)pos.size();
  poses_out.assign((size_t)n, Pose5());
  if (n == 0) return FCVM_OK;
  const int64_t W = h->W;
  int rc;
  const auto t_gp0 = std::chrono::steady_clock::now();
  auto gmark = [&](const char* what) {
    if (env().profile >= 2)
      std::fprintf(stderr, "[fcvm]   getPose batch n=%lld: %s at %.3f ms\n", (long long)n, what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_gp0).count());
  };
  for (int64_t c0 = 0; c0 < n; c0 += chunk_rows(h, n, true)) {
    const int64_t cn = std::min(chunk_rows(h, n, true), n - c0);
    std::vector<PoseSeed> seed((size_t)cn);
    std::vector<double> ip((size_t)cn);
    CU(h->h_recs.reserve((size_t)cn));
    CU(h->h_aux.reserve((size_t)cn));
    CU(h->h_count.reserve((size_t)cn));
    pfor(h, cn, 128, [&](int64_t i) {
      // fov_sample_ray = dir.normalized(); fov_yaw_ray = (x, y, 0).normalized(); init_py = PitchYaw
      PoseSeed& sd = seed[(size_t)i];
      sd.vp = pos[(size_t)(c0 + i)];
      sd.fs = normalized3(dir[(size_t)(c0 + i)]);
      sd.fy = normalized3(D3{sd.fs.x, sd.fs.y, 0.0});
      pitch_yaw(sd.fs, ip[(size_t)i], sd.init_yaw);
      h->cam.fill(h->h_recs.p[i], sd.vp.x, sd.vp.y, sd.vp.z, ip[(size_t)i], sd.init_yaw);
      h->h_aux.p[i] = PoseAux{sd.vp.x, sd.vp.y, sd.vp.z, sd.fs.x, sd.fs.y, sd.fs.z, sd.fy.x, sd.fy.y};
      poses_out[(size_t)(c0 + i)] = Pose5{{sd.vp.x, sd.vp.y, sd.vp.z, ip[(size_t)i], sd.init_yaw}};
    });
    std::vector<int> ids((size_t)cn), order((size_t)cn);
    for (int64_t i = 0; i < cn; ++i) { ids[(size_t)i] = first_id + (int)(c0 + i); order[(size_t)i] = (int)i; }
    int64_t lo = 0, hi = cn;
    if (h->world > 1) shard_of(h, cn, lo, hi);

    if (!h->prm.pose_update) {
      // getPose with pose_update == false delegates to updateViewpointInfo (cpp:347-353): E4 + ownership
      rc = eval_batch(h, FCVM_STAGE_E4, h->h_recs.p, cn, nullptr, nullptr, nullptr, h->d_rows1, false, nullptr, false);
      if (rc) return rc;
      if (h->trace_on) {
        rc = fetch_rows_for_trace(h, h->d_rows1, cn);
        if (rc) return rc;
        for (int64_t i = 0; i < cn; ++i) record(h, 4, ids[(size_t)i], poses_out[(size_t)(c0 + i)], h->h_rows.p + i * W);
      }
      rc = ownership(h, h->d_rows1, cn, lo, hi, order, ids, h->h_count.p);
      if (rc) return rc;
      continue;
    }

    // ---- E1 + the ordered pitch / yaw sums (cpp:355-447) ------------------------------------
    const int64_t per = (cn + h->world - 1) / h->world, cn_pad = per * h->world;
    {
      Stopwatch sw(h->t_eval);
      CU(h->d_recs.reserve((size_t)cn_pad));
      CU(h->d_aux.reserve((size_t)cn_pad));
      CU(h->d_rows1.reserve((size_t)cn_pad * W));
      CU(h->d_rows2.reserve((size_t)cn_pad * W));
      CU(h->d_count.reserve((size_t)cn_pad));
      CU(cudaMemcpyAsync(h->d_recs.p, h->h_recs.p, (size_t)cn * sizeof(VpRec), cudaMemcpyHostToDevice, h->stream));
      CU(cudaMemcpyAsync(h->d_aux.p, h->h_aux.p, (size_t)cn * sizeof(PoseAux), cudaMemcpyHostToDevice, h->stream));
    }
    std::vector<long long> free_num((size_t)cn, 0);
    std::vector<double> sums((size_t)2 * cn, 0.0);
    gmark("records built and queued for upload");
    {
      Stopwatch sw(h->t_eval);
      rc = e1_with_pose_sums(h, lo, hi, seed, free_num, sums);
      if (rc) return rc;
      rc = fetch_counters(h);      // synchronises the main stream
      if (rc) return rc;
    }
    gmark("E1 + pose sums");
    if (h->trace_on) {
      rc = fetch_rows_for_trace(h, h->d_rows1, cn);
      if (rc) return rc;
      for (int64_t i = 0; i < cn; ++i) record(h, 1, ids[(size_t)i], poses_out[(size_t)(c0 + i)], h->h_rows.p + i * W);
    }
    // averaged poses (cpp:427-447); (pitch, yaw, free_num) of every viewpoint reach every rank
    std::vector<double> avg3((size_t)cn * 3, 0.0);
    pfor(h, hi - lo, 1024, [&](int64_t r) {
      const int64_t i = lo + r;
      if (free_num[(size_t)i] == 0) return;       // free_num == 0: pose stays (pos, init_py), no E2, no ownership (cpp:427-436)
      finish_average(h, sums[(size_t)2 * i], sums[(size_t)2 * i + 1], free_num[(size_t)i], seed[(size_t)i].init_yaw, avg3[(size_t)3 * i], avg3[(size_t)3 * i + 1]);
      avg3[(size_t)3 * i + 2] = (double)free_num[(size_t)i];
    });
    rc = gather_doubles(h, avg3.data(), cn, 3);
    if (rc) return rc;
    std::vector<char> has_free((size_t)cn, 0);
    pfor(h, cn, 128, [&](int64_t i) {   // the E2 records (sin / cos of the averaged pose: libm, per viewpoint)
      if (avg3[(size_t)3 * i + 2] == 0.0) return;
      const D3& vp = seed[(size_t)i].vp;
      poses_out[(size_t)(c0 + i)].v[3] = avg3[(size_t)3 * i];
      poses_out[(size_t)(c0 + i)].v[4] = avg3[(size_t)3 * i + 1];
      has_free[(size_t)i] = 1;
      h->cam.fill(h->h_recs.p[i], vp.x, vp.y, vp.z, avg3[(size_t)3 * i], avg3[(size_t)3 * i + 1]);
    });

    gmark("averaged poses, E2 records");
    // ---- E2 on the E1 sets (cpp:457-498) ---------------------------------------------------
    // viewpoints with free_num == 0 have all-zero E1 rows, so they contribute nothing here
    {
      Stopwatch sw(h->t_eval);
      CU(cudaMemcpyAsync(h->d_recs.p, h->h_recs.p, (size_t)cn * sizeof(VpRec), cudaMemcpyHostToDevice, h->stream));
      if (h->world > 1) CU(cudaMemsetAsync(h->d_count.p, 0, (size_t)cn_pad * sizeof(int), h->stream));
      if (hi > lo) {
        rc = launch_stage(h, FCVM_STAGE_E2, h->d_recs.p + lo, hi - lo, h->d_rows2.p + lo * W, h->d_rows1.p + lo * W, h->stream, true);
        if (rc) return rc;
        CU(launch_popc_rows(h->d_rows2.p + lo * W, hi - lo, W, h->d_count.p + lo, h->stream));
        h->counters[5] += 1;
      }
      rc = fetch_counters(h);
      if (rc) return rc;
    }
    if (h->trace_on) {
      rc = fetch_rows_for_trace(h, h->d_rows2, cn);
      if (rc) return rc;
      for (int64_t i = 0; i < cn; ++i)
        if (has_free[(size_t)i]) record(h, 2, ids[(size_t)i], poses_out[(size_t)(c0 + i)], h->h_rows.p + i * W);
    }
    gmark("E2");
    rc = ownership(h, h->d_rows2, cn, lo, hi, order, ids, h->h_count.p);
    if (rc) return rc;
    gmark("ownership queued");
  }
  return FCVM_OK;
}

// updateViewpointInfo over a batch in processing order (viewpoint_manager.cpp:533-618)
static int update_info_batch(fcvm_handle* h, const std::vector<int>& ids, const std::vector<Pose5>& poses) {
  const int64_t n = (int64_t)ids.size();
  const int64_t W = h->W;
  int rc;
  for (int64_t c0 = 0; c0 < n; c0 += chunk_rows(h, n, true)) {
    const int64_t cn = std::min(chunk_rows(h, n, true), n - c0);
    CU(h->h_recs.reserve((size_t)cn));
    CU(h->h_count.reserve((size_t)cn));
    std::vector<int> cid((size_t)cn), order((size_t)cn);
    for (int64_t i = 0; i < cn; ++i) {
      const Pose5& p = poses[(size_t)(c0 + i)];
      h->cam.fill(h->h_recs.p[i], p.v[0], p.v[1], p.v[2], p.v[3], p.v[4]);
      cid[(size_t)i] = ids[(size_t)(c0 + i)];
      order[(size_t)i] = (int)i;
    }
    rc = eval_batch(h, FCVM_STAGE_E4, h->h_recs.p, cn, nullptr, nullptr, nullptr, h->d_rows1, false, nullptr, false);
    if (rc) return rc;
    if (h->trace_on) {
      rc = fetch_rows_for_trace(h, h->d_rows1, cn);
      if (rc) return rc;
      for (int64_t i = 0; i < cn; ++i) record(h, 4, cid[(size_t)i], poses[(size_t)(c0 + i)], h->h_rows.p + i * W);
    }
    int64_t lo = 0, hi = cn;
    if (h->world > 1) shard_of(h, cn, lo, hi);
    rc = ownership(h, h->d_rows1, cn, lo, hi, order, cid, h->h_count.p);
    if (rc) return rc;
  }
  return FCVM_OK;
}

// ------------------------------------------------------------------------------------------------
// viewpointsPrune (viewpoint_manager.cpp:620-710) with updatePoseGravitation (712-887) split into
// the sequential host flag chain and one batched E3 launch.
// ------------------------------------------------------------------------------------------------
// Stand-in for the reference's inverse_idx_ (unordered_map keyed by the candidate's float position, cpp:643: the LAST candidate
// inserted at a position owns it).  It is only ever looked up (cpp:739), so its iteration order is not
// observable: an open-addressing table over the slots' positions, with the map's key equality (operator== per component,
// so -0 == +0 and a NaN position equals nothing, itself included).
struct PosTable {
  std::vector<int> tab;
  const float* cloud = nullptr;
  uint32_t mask = 0;
  static inline uint32_t bits(float f) { if (f == 0.0f) return 0u; uint32_t u; std::memcpy(&u, &f, 4); return u; }
  static inline uint32_t hash3(const float* q) {
    uint64_t x = ((uint64_t)bits(q[0]) * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)bits(q[1]) * 0xC2B2AE3D27D4EB4Full) ^ ((uint64_t)bits(q[2]) * 0x165667B19E3779F9ull);
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    return (uint32_t)x;
  }
  void build(const float* c, int n) {
    cloud = c;
    uint32_t cap = 16;
    while (cap < 2u * (uint32_t)std::max(n, 1)) cap <<= 1;
    mask = cap - 1;
    tab.assign(cap, -1);
    for (int sl = 0; sl < n; ++sl) {
      const float* q = c + 3 * (size_t)sl;
      uint32_t i = hash3(q) & mask;
      for (;; i = (i + 1) & mask) {
        const int o = tab[i];
        if (o < 0) { tab[i] = sl; break; }
        const float* p = c + 3 * (size_t)o;
        if (p[0] == q[0] && p[1] == q[1] && p[2] == q[2]) { tab[i] = sl; break; }     // same key: the later insertion owns it
      }
    }
  }
  int find(const float* q, int self) const {          // the slot that owns position q (self if the key matches nothing: NaN)
    for (uint32_t i = hash3(q) & mask;; i = (i + 1) & mask) {
      const int o = tab[i];
      if (o < 0) return self;
      const float* p = cloud + 3 * (size_t)o;
      if (p[0] == q[0] && p[1] == q[1] && p[2] == q[2]) return o;
    }
  }
};

static int viewpoints_prune(fcvm_handle* h, const std::vector<Pose5>& vps_pose, const std::vector<int>& sub_voxcount) {
  const fcvm_params& prm = h->prm;
  const auto t_prune0 = std::chrono::steady_clock::now();
  auto pmark = [&](const char* what) {
    if (env().profile >= 2)
      std::fprintf(stderr, "[fcvm]   viewpointsPrune: %s at %.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_prune0).count());
  };
  auto& idx_slot = h->idx_slot;
  idx_slot.clear();
  std::vector<int> slot_id;
  std::vector<Pose5> slot_pose;
  std::vector<char> live, query;
  std::vector<float> vp_cloud;

  for (int i = h->last_vps_num; i < (int)vps_pose.size(); ++i) {
    if (sub_voxcount[(size_t)i] > 0) {
      const Pose5& q = vps_pose[(size_t)i];
      const float fx = (float)q.v[0], fyy = (float)q.v[1], fz = (float)q.v[2];
      vp_cloud.push_back(fx); vp_cloud.push_back(fyy); vp_cloud.push_back(fz);
      idx_slot[i] = (int)slot_id.size();
      slot_id.push_back(i);
      slot_pose.push_back(q);
      live.push_back(1);
      query.push_back(0);
    }
  }
  if (vp_cloud.empty()) return FCVM_OK;   // "No need prune viewpoints!" cpp:651-655

  // sort by ctrl voxels, descending, over the unordered_map's own iteration order (cpp:658-662).  The reference's
  // idx_ctrl_voxels_ receives the same keys in the same order, and is cleared at the same moments, as idx_viewpoints_ (idx_slot
  // here): libstdc++'s iteration order depends on nothing else, so one map stands for both.
  std::vector<std::pair<int, int>> ctrlVector;
  ctrlVector.reserve(idx_slot.size());
  for (const auto& kv : idx_slot) ctrlVector.emplace_back(kv.first, sub_voxcount[(size_t)kv.first]);
  std::sort(ctrlVector.begin(), ctrlVector.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.second > b.second; });

  CellIndex vps_tree;
  const double bubble_radius = 0.55 * prm.viewpoints_distance * std::tan(h->fov_base);
  vps_tree.build(vp_cloud.data(), (int64_t)slot_id.size(), bubble_radius);

  const double pitch_bound = 0.3 * prm.fov_h * M_PI / 180.0;
  const double yaw_bound = 0.3 * prm.fov_w * M_PI / 180.0;

  pmark("candidates collected, sorted, indexed");
  // nearCheck of every candidate's own position (cpp:678-682): the poses read here are never modified before their
  // own turn in the chain below, so the checks are independent and run as one device batch
  std::vector<uint8_t> near_ok;
  {
    int rc = safety_batch(h, vp_cloud, 1, near_ok);
    if (rc) return rc;
  }

  pmark("nearCheck batch done");
  // Tables that stand in for the reference's per-neighbour map look-ups inside the chain (cpp:733-745): the slot of an id,
  // the ctrl-voxel count of a slot, and - because inverse_idx_ is keyed by the float position, so coincident candidates
  // collapse to the LAST inserted id - the slot the reference actually reaches from cloud point s.
  // The radius searches (cpp:685) read only the original float cloud, so they are independent of the chain's state; their
  // results are only ever used for a candidate that is still alive at its turn (cpp:688) - a few hundred of 200 000 at
  // configs[4], the rest having been absorbed by a stronger neighbour.  So they run lazily, in waves: when the chain reaches a
  // live candidate without a list, the lists of the next kWave live candidates in chain order are computed on all host threads
  // (some of those die before their turn: wasted, harmless).
  const int nslot = (int)slot_id.size();
  const int id0 = h->last_vps_num;
  std::vector<int> slot_of_id(vps_pose.size() - (size_t)id0, -1), slot_vox((size_t)nslot), canon((size_t)nslot);
  for (int sl = 0; sl < nslot; ++sl) { slot_of_id[(size_t)(slot_id[(size_t)sl] - id0)] = sl; slot_vox[(size_t)sl] = sub_voxcount[(size_t)slot_id[(size_t)sl]]; }
  PosTable pos_table;
  pos_table.build(vp_cloud.data(), nslot);
  pfor(h, nslot, 2048, [&](int64_t sl) { canon[(size_t)sl] = pos_table.find(&vp_cloud[3 * (size_t)sl], (int)sl); });
  std::vector<std::vector<int>> lists;          // radiusSearch results, (distance, index)-sorted
  std::vector<int> list_of((size_t)nslot, -1);  // slot -> index into `lists`, -1 = not searched (yet)
  std::vector<int> wave;
  const size_t kWave = (size_t)std::max(32, 4 * h->threads);
  auto qualifies = [&](int sl) { return slot_vox[(size_t)sl] > 1 && near_ok[(size_t)sl]; };
  auto search_wave = [&](size_t from) {
    wave.clear();
    for (size_t j = from; j < ctrlVector.size() && wave.size() < kWave; ++j) {
      const int sl = slot_of_id[(size_t)(ctrlVector[j].first - id0)];
      if (live[(size_t)sl] && qualifies(sl) && list_of[(size_t)sl] < 0) wave.push_back(sl);
    }
    const size_t base = lists.size();
    lists.resize(base + wave.size());
    pfor(h, (int64_t)wave.size(), 1, [&](int64_t w) {
      std::vector<int> found;
      vps_tree.radius_search(&vp_cloud[3 * (size_t)wave[(size_t)w]], bubble_radius, found);
      lists[base + (size_t)w] = std::move(found);
    });
    for (size_t w = 0; w < wave.size(); ++w) list_of[(size_t)wave[w]] = (int)(base + w);
  };
  pmark("position table, canonical slots");

  struct Grav { int slot; D3 up; double upi, uy; };
  std::vector<Grav> grav;
  struct Pulled { int slot; D3 up; D3 cur; };
  std::vector<Pulled> pulled;                 // positions whose viewpointSafetyCheck (cpp:805-810) is resolved after the chain
  std::vector<int> updated;                   // slots absorbed by the current candidate
  for (size_t ci = 0; ci < ctrlVector.size(); ++ci) {
    const int c = ctrlVector[ci].first;
    const int cs = slot_of_id[(size_t)(c - id0)];
    const Pose5 cur_pose = slot_pose[(size_t)cs];
    const D3 cur_position{cur_pose.v[0], cur_pose.v[1], cur_pose.v[2]};
    const int cur_vox = slot_vox[(size_t)cs];
    if (!qualifies(cs)) { live[(size_t)cs] = 0; continue; }                           // nearCheck(cur_position, cur_vox)
    if (!live[(size_t)cs]) continue;                                                   // its search result would not be used (cpp:688)
    if (list_of[(size_t)cs] < 0) search_wave(ci);
    const std::vector<int>& nlist = lists[(size_t)list_of[(size_t)cs]];
    const int* neigh = nlist.data();
    const int n_neigh = (int)nlist.size();
    if (n_neigh <= 1) continue;

    // ---- updatePoseGravitation, host part (cpp:714-810) -----------------------------------
    query[(size_t)cs] = 1;
    const double cur_pitch = cur_pose.v[3], cur_yaw = cur_pose.v[4];
    updated.clear();
    for (int e = 0; e < n_neigh; ++e) {
      const int ms = canon[(size_t)neigh[e]];
      const Pose5& inner_pose = slot_pose[(size_t)ms];
      const int inner_vox = slot_vox[(size_t)ms];
      if (inner_vox > cur_vox) continue;
      double diff_yaw = std::fabs(cur_pose.v[4] - inner_pose.v[4]);
      warp_angle(diff_yaw);
      double diff_pitch = std::fabs(cur_pose.v[3] - inner_pose.v[3]);
      warp_angle(diff_pitch);
      const bool update_flag = std::fabs(diff_pitch) < pitch_bound && std::fabs(diff_yaw) < yaw_bound;
      if (update_flag && !query[(size_t)ms]) {
        live[(size_t)ms] = 0;
        updated.push_back(ms);
      }
    }
    D3 up = cur_position;
    double upi = cur_pitch, uy = cur_yaw;
    for (int ms : updated) {
      const double grav_factor = (double)slot_vox[(size_t)ms] / (double)cur_vox;
      const Pose5& o = slot_pose[(size_t)ms];
      const D3 pull{grav_factor * (o.v[0] - cur_position.x), grav_factor * (o.v[1] - cur_position.y), grav_factor * (o.v[2] - cur_position.z)};
      up = D3{up.x + pull.x, up.y + pull.y, up.z + pull.z};
      double off; int flag, dirn;
      if (prm.attitude == FCVM_ATTITUDE_ALL) {
        wrapped_diff(o.v[3], cur_pitch, off, flag, dirn);
        upi = upi + grav_factor * flag * dirn * off;
      }
      wrapped_diff(o.v[4], cur_yaw, off, flag, dirn);
      uy = uy + grav_factor * flag * dirn * off;
    }
    if (prm.z_ground && up.z < prm.ground_pos + prm.safe_height) up.z = prm.ground_pos + prm.safe_height + 0.1;
    Pose5& dst = slot_pose[(size_t)cs];
    // keep the pulled position only if it is safe (cpp:805-810).  Nothing later in the chain reads the position of a
    // viewpoint that has been queried, so all these checks are resolved together below.
    pulled.push_back(Pulled{cs, up, cur_position});
    if (!prm.pose_update) {
      dst.v[3] = upi; dst.v[4] = uy;
    } else {
      grav.push_back(Grav{cs, up, upi, uy});   // E3 uses `updated_position` even if it was reverted (cpp:821, 844)
    }
  }

  pmark("flag chain done");
  if (!pulled.empty()) {
    std::vector<float> q;
    q.reserve(3 * pulled.size());
    for (const Pulled& pl : pulled) { q.push_back((float)pl.up.x); q.push_back((float)pl.up.y); q.push_back((float)pl.up.z); }
    std::vector<uint8_t> safe;
    int rc = safety_batch(h, q, 0, safe);
    if (rc) return rc;
    for (size_t i = 0; i < pulled.size(); ++i) {
      const D3& src = safe[i] ? pulled[i].up : pulled[i].cur;
      Pose5& dst = slot_pose[(size_t)pulled[i].slot];
      dst.v[0] = src.x; dst.v[1] = src.y; dst.v[2] = src.z;
    }
  }

  pmark("pulled positions checked");
  // ---- E3 for every gravitated viewpoint in one batch (cpp:817-886) ---------------------------
  if (!grav.empty()) {
    const int64_t W = h->W;
    const float* M = h->model.data();
    const double pu = h->pitch_upper * M_PI / 180.0, pl = h->pitch_lower * M_PI / 180.0;
    const int64_t n = (int64_t)grav.size();
    for (int64_t c0 = 0; c0 < n; c0 += chunk_rows(h, n, true)) {
      const int64_t cn = std::min(chunk_rows(h, n, true), n - c0);
      CU(h->h_recs.reserve((size_t)cn));
      for (int64_t i = 0; i < cn; ++i) {
        const Grav& gq = grav[(size_t)(c0 + i)];
        h->cam.fill(h->h_recs.p[i], gq.up.x, gq.up.y, gq.up.z, gq.upi, gq.uy);
      }
      CsrOut csr;
      int64_t my_lo = 0, my_hi = cn;
      if (h->world > 1) { shard_of(h, cn, my_lo, my_hi); csr.lo = my_lo; csr.hi = my_hi; }
      int rc = eval_batch(h, FCVM_STAGE_E3, h->h_recs.p, cn, nullptr, nullptr, nullptr, h->d_rows1, false, &csr, false);
      if (rc) return rc;
      if (h->trace_on) {
        rc = fetch_rows_for_trace(h, h->d_rows1, cn);
        if (rc) return rc;
      }
      std::vector<double> upd((size_t)cn * 2, 0.0);
      const auto t_sums0 = std::chrono::steady_clock::now();
      pfor(h, my_hi - my_lo, 1, [&](int64_t r) {
        const int64_t i = my_lo + r;
        const Grav& gq = grav[(size_t)(c0 + i)];
        double avg_pitch = gq.upi, avg_yaw = gq.uy;
        for (long long k = csr.offs[(size_t)r]; k < csr.offs[(size_t)r + 1]; ++k) {
          const size_t j = csr.idx[k];
          const D3 cand{M[3 * j], M[3 * j + 1], M[3 * j + 2]};
          const D3 ray_dir = normalized3(sub3(cand, gq.up));
          double pitch, yaw, off; int flag, dirn;
          pitch_yaw(ray_dir, pitch, yaw);
          warp_angle(pitch);
          warp_angle(yaw);
          wrapped_diff(pitch, gq.upi, off, flag, dirn);
          avg_pitch = avg_pitch + flag * dirn * off;
          wrapped_diff(yaw, gq.uy, off, flag, dirn);
          avg_yaw = avg_yaw + flag * dirn * off;
        }
        double updated_pitch = avg_pitch, updated_yaw = avg_yaw;
        if (updated_pitch > pu) updated_pitch = pu;
        if (updated_pitch < pl) updated_pitch = pl;
        warp_angle(updated_pitch);
        warp_angle(updated_yaw);
        upd[(size_t)2 * i] = updated_pitch;
        upd[(size_t)2 * i + 1] = updated_yaw;
      });
      h->t_sums += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_sums0).count();
      rc = gather_doubles(h, upd.data(), cn, 2);
      if (rc) return rc;
      for (int64_t i = 0; i < cn; ++i) {
        const Grav& gq = grav[(size_t)(c0 + i)];
        slot_pose[(size_t)gq.slot].v[3] = upd[(size_t)2 * i];
        slot_pose[(size_t)gq.slot].v[4] = upd[(size_t)2 * i + 1];
      }
      if (h->trace_on)
        for (int64_t i = 0; i < cn; ++i) {
          const Grav& gq = grav[(size_t)(c0 + i)];
          record(h, 3, slot_id[(size_t)gq.slot], Pose5{{gq.up.x, gq.up.y, gq.up.z, gq.upi, gq.uy}}, h->h_rows.p + i * W);
        }
    }
  }

  // ---- E4 + ownership for the live ones, in unordered_map iteration order (cpp:693-697) -------
  std::vector<int> ids;
  std::vector<Pose5> poses;
  for (auto& kv : idx_slot)
    if (live[(size_t)kv.second]) { ids.push_back(kv.first); poses.push_back(slot_pose[(size_t)kv.second]); }
  pmark("E3 done");
  int rc = update_info_batch(h, ids, poses);
  if (rc) return rc;
  pmark("E4 + ownership done");
  rc = voxcount_download(h, h->vps_voxcount);
  if (rc) return rc;
  for (auto& kv : idx_slot)       // cpp:699-709
    if (live[(size_t)kv.second] && h->vps_voxcount[(size_t)kv.first] > 0)
      h->final_vps.push_back(FinalVp{kv.first, slot_pose[(size_t)kv.second], h->vps_voxcount[(size_t)kv.first]});
  return FCVM_OK;
}

// uncNeighVps (viewpoint_manager.cpp:948-1131) with injected seeds: engine k of call c is seeded
// with seed + 8*c + k (k = 0..3 pitch, 4..7 yaw)
// The 8 candidates are produced here; their viewpointSafetyCheck (cpp:1060-1128) runs on the device for all points at once.
static void unc_neigh_vps(const fcvm_handle* h, const float pt[3], const float nrm[3], PtNormal cand[8], int64_t call) {
  const fcvm_params& prm = h->prm;
  double max_pitch_range = 40.0, max_yaw_range = 40.0;
  if (prm.attitude == FCVM_ATTITUDE_YAW) max_pitch_range = 0;
  double op, oy;
  pitch_yaw(D3{nrm[0], nrm[1], nrm[2]}, op, oy);
  const uint64_t base = prm.seed + 8ull * (uint64_t)call;
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
  static const int sp[8] = {+1, -1, +1, -1, +1, -1, +1, -1};
  static const int sy[8] = {+1, -1, +1, -1, -1, +1, -1, +1};
  for (int q = 0; q < 8; ++q) {
    const int k = q & 3;
    const double pitch = sp[q] > 0 ? op + rand_p[k] : op - rand_p[k];
    const double yaw = sy[q] > 0 ? oy + rand_y[k] : oy - rand_y[k];
    const D3 vec = py_to_vec(pitch, yaw);
    cand[q] = PtNormal{(float)(pt[0] + prm.viewpoints_distance * vec.x), (float)(pt[1] + prm.viewpoints_distance * vec.y),
                       (float)(pt[2] + prm.viewpoints_distance * vec.z), (float)(-vec.x), (float)(-vec.y), (float)(-vec.z)};
  }
}

// stand-in for pcl::RandomSample (cpp:249-255): selection sampling driven by minstd_rand0(seed + phi + call#).
// Returns the (ascending) positions chosen out of a pool of `pool_size` elements; the pool itself need not exist.
static std::vector<size_t> random_sample_positions(fcvm_handle* h, size_t pool_size, int sample) {
  std::minstd_rand0 g((std::minstd_rand0::result_type)(h->prm.seed + 0x9e3779b9ull + (uint64_t)h->sample_calls++));
  std::vector<size_t> out;
  size_t N = pool_size, n = (size_t)sample, index = 0;
  while (n > 0 && index < pool_size) {
    const float U = (float)((double)g() / 2147483646.0);
    if ((float)N * U <= (float)n) { out.push_back(index); --n; }
    ++index; --N;
  }
  return out;
}

static int set_init_viewpoints_impl(fcvm_handle* h, const std::vector<PtNormal>& vps) {
  if (vps.empty()) return fail(FCVM_EINVAL, "Input viewpoints Fail!! Input viewpoint is empty!!!");
  if (!h->map_flag || !h->model_flag) return fail(FCVM_ENOTREADY, "setInitViewpoints needs the map and the model first");
  int rc = ensure_device(h);
  if (rc) return rc;
  const size_t n = vps.size();
  // cpp:129-131: resize (keeps earlier contents if called twice)
  rc = voxcount_resize(h, std::max(h->voxcount_n, n));
  if (rc) return rc;
  if (h->vps_contri.size() < n) h->vps_contri.resize(n, 0);
  h->vps_pose.resize(n);
  std::vector<D3> pos(n), dir(n);
  for (size_t i = 0; i < n; ++i) {
    pos[i] = D3{vps[i].x, vps[i].y, vps[i].z};
    dir[i] = D3{vps[i].nx, vps[i].ny, vps[i].nz};
  }
  std::vector<Pose5> poses;
  rc = get_pose_batch(h, pos, dir, 0, poses);
  if (rc) return rc;
  h->vps_pose = poses;
  h->viewpoints_flag = true;
  return FCVM_OK;
}

static int update_impl(fcvm_handle* h) {
  int rc = ensure_device(h);
  if (rc) return rc;
  const fcvm_params& prm = h->prm;
  const float* M = h->model.data();
  const auto t_upd0 = std::chrono::steady_clock::now();
  auto umark = [&](const char* what) {
    if (env().profile >= 2)
      std::fprintf(stderr, "[fcvm] updateViewpoints: %s at %.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upd0).count());
  };
  if (!h->viewpoints_flag) {
    // "No input viewpoints!! Use normals gengerate viewpoints!!" (cpp:145-171)
    if (!h->map_flag || !h->model_flag || !h->normal_flag) return fail(FCVM_ENOTREADY, "Not ready: map, model and normals are required");
    std::vector<PtNormal> all, vps;
    std::vector<float> q;
    for (int64_t i = 0; i < h->P; ++i) {
      auto it = h->pt_normal_pairs.find(D3{M[3 * i], M[3 * i + 1], M[3 * i + 2]});
      if (it == h->pt_normal_pairs.end()) continue;    // reference: unchecked dereference
      const D3 nd = it->second;
      const float vx = (float)(M[3 * i] + prm.viewpoints_distance * nd.x);
      const float vy = (float)(M[3 * i + 1] + prm.viewpoints_distance * nd.y);
      const float vz = (float)(M[3 * i + 2] + prm.viewpoints_distance * nd.z);
      all.push_back(PtNormal{vx, vy, vz, (float)(-nd.x), (float)(-nd.y), (float)(-nd.z)});
      q.push_back(vx); q.push_back(vy); q.push_back(vz);
    }
    std::vector<uint8_t> safe;
    umark("candidates from the normals");
    rc = safety_batch(h, q, 0, safe);                  // viewpointSafetyCheck of every candidate (cpp:159-160), one launch
    if (rc) return rc;
    for (size_t i = 0; i < all.size(); ++i)
      if (safe[i]) vps.push_back(all[i]);
    umark("their safety check");
    rc = set_init_viewpoints_impl(h, vps);
    if (rc && rc != FCVM_EINVAL) return rc;
    umark("setInitViewpoints (getPose of all)");
  }
  if (!(h->map_flag && h->model_flag && h->viewpoints_flag && h->normal_flag)) return fail(FCVM_ENOTREADY, "[ViewpointManager] Not ready");

  // cpp:176-187
  std::vector<int> temp_vps_voxcount;
  rc = cover_reset(h);
  if (rc) return rc;
  rc = voxcount_download(h, temp_vps_voxcount);
  if (rc) return rc;
  rc = voxcount_resize(h, 0);                                                                  // resize(P) in the reference; A.9:
  if (!rc) rc = voxcount_resize(h, std::max((size_t)h->P, h->vps_pose.size()));                 // all zeros, by a memset on the device
  if (rc) return rc;
  umark("cover reset, voxcount swap");
  rc = viewpoints_prune(h, h->vps_pose, temp_vps_voxcount);
  if (rc) return rc;
  umark("first prune");

  auto uncovered = [&](std::vector<int>& ids) -> int {
    std::vector<int> cid((size_t)h->P);
    CU(cudaMemcpyAsync(cid.data(), h->d_contrib_id.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    ids.clear();
    for (int i = 0; i < (int)h->P; ++i)
      if (cid[(size_t)i] < 0) ids.push_back(i);
    return FCVM_OK;
  };
  std::vector<int> uncoveredID;
  rc = uncovered(uncoveredID);
  if (rc) return rc;
  umark("uncovered ids");
  const std::vector<int> uncoveredID0 = uncoveredID;      // the library below is indexed by position in this first list

  // cpp:206-221: candidate library for the uncovered points.
  // The points are independent (candidate q of the c-th call draws from engines seeded seed + 8c + k), and a candidate is a
  // pure function of its point and its call number.  What the rounds below need of the library is, per uncovered point, which
  // of its 8 candidates passed viewpointSafetyCheck - one byte; the <= 100 candidates a round actually draws are regenerated on
  // the spot.  So the library is a byte per point: all host threads generate the candidates of this rank's block of the points
  // (RNG + libm), one k_safety launch checks them, and a sharded run all-gathers the bytes (round 2 before: 200-byte records).
  const int64_t n_lib = (int64_t)uncoveredID.size();
  const int64_t lib_per = (n_lib + h->world - 1) / h->world;
  std::vector<uint8_t> lib_mask((size_t)std::max<int64_t>(1, lib_per * h->world), 0);
  std::vector<int> lib_slot((size_t)h->P, -1);
  const int64_t call0 = h->unc_calls;
  h->unc_calls += n_lib;
  auto candidates_of = [&](int64_t k, PtNormal c8[8]) {       // the 8 candidates of the k-th uncovered point (uncNeighVps, cpp:948-1131)
    const int p = uncoveredID0[(size_t)k];
    auto it = h->pt_normal_pairs.find(D3{M[3 * p], M[3 * p + 1], M[3 * p + 2]});
    const D3 nd = it == h->pt_normal_pairs.end() ? D3{0, 0, 1} : it->second;
    const float nrm[3] = {(float)nd.x, (float)nd.y, (float)nd.z};
    unc_neigh_vps(h, &M[3 * p], nrm, c8, call0 + k);
  };
  {
    Stopwatch sw_unc(h->t_unc);
    const bool vprof = env().profile >= 3;
    const auto t_unc0 = std::chrono::steady_clock::now();
    auto tmark = [&](const char* what) { if (vprof) std::fprintf(stderr, "[fcvm]     unc: %s at %.2f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_unc0).count()); };
    tmark("start");
    for (int64_t k = 0; k < n_lib; ++k) lib_slot[(size_t)uncoveredID0[(size_t)k]] = (int)k;
    int64_t lo = 0, hi = n_lib;
    if (h->world > 1) shard_of(h, n_lib, lo, hi);
    std::unique_ptr<float[]> q(new float[(size_t)std::max<int64_t>(1, 24 * (hi - lo))]);
    tmark("buffers ready");
    pfor(h, hi - lo, 64, [&](int64_t r) {
      PtNormal c8[8];
      candidates_of(lo + r, c8);
      for (int c = 0; c < 8; ++c) { q[(size_t)24 * r + 3 * c] = c8[c].x; q[(size_t)24 * r + 3 * c + 1] = c8[c].y; q[(size_t)24 * r + 3 * c + 2] = c8[c].z; }
    });
    tmark("candidates generated");
    std::vector<uint8_t> safe;
    rc = safety_batch(h, q.get(), 8 * (hi - lo), 0, safe);
    if (rc) return rc;
    tmark("safety batch done");
    for (int64_t r = 0; r < hi - lo; ++r) {
      uint8_t m = 0;
      for (int c = 0; c < 8; ++c) m |= safe[(size_t)8 * r + c] ? (uint8_t)(1u << c) : (uint8_t)0;
      lib_mask[(size_t)(lo + r)] = m;
    }
    rc = gather_records(h, lib_mask.data(), n_lib, 1);
    if (rc) return rc;
    tmark("library complete");
  }

  umark("candidate library");
  int unc_iter = 0, last_unc_num = -1;
  while ((int)uncoveredID.size() > 0 && unc_iter < prm.max_iter_num && last_unc_num != (int)uncoveredID.size()) {
    last_unc_num = (int)uncoveredID.size();
    // pool = the safe candidates of the still-uncovered points, in point order (cpp:240-247); at most 100 of them are drawn
    // (cpp:249-255).  The pool is only counted and walked, never materialised.
    std::vector<PtNormal> cand;
    {
      size_t pool_size = 0;
      for (int p : uncoveredID) pool_size += (size_t)__builtin_popcount(lib_mask[(size_t)lib_slot[(size_t)p]]);
      std::vector<size_t> pick;
      const bool all = pool_size <= 100;
      if (!all) pick = random_sample_positions(h, pool_size, 100);
      size_t pos = 0, next = 0;
      for (int p : uncoveredID) {
        if (!all && next >= pick.size()) break;
        const int slot = lib_slot[(size_t)p];
        const unsigned m = lib_mask[(size_t)slot];
        const size_t cnt = (size_t)__builtin_popcount(m);
        if (!all && pick[next] >= pos + cnt) { pos += cnt; continue; }      // none of this point's candidates is drawn
        PtNormal c8[8];
        candidates_of(slot, c8);
        for (int c = 0; c < 8; ++c) {
          if (!((m >> c) & 1u)) continue;
          if (all) cand.push_back(c8[c]);
          else if (next < pick.size() && pick[next] == pos) { cand.push_back(c8[c]); ++next; }
          ++pos;
        }
      }
    }

    // snapshot of MCI (cpp:272-274)
    DevBuf<int>& snap_contrib = h->d_snap_contrib;
    DevBuf<int>& snap_id = h->d_snap_id;
    CU(snap_contrib.reserve((size_t)h->P));
    CU(snap_id.reserve((size_t)h->P));
    CU(cudaMemcpyAsync(snap_contrib.p, h->d_cover_contrib.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(snap_id.p, h->d_contrib_id.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));

    const int oriVpNum = (int)h->vps_pose.size();
    h->last_vps_num = oriVpNum;
    rc = voxcount_resize(h, h->voxcount_n + cand.size());
    if (rc) return rc;
    h->vps_contri.resize(h->vps_contri.size() + cand.size(), 0);
    if (h->vps_contri.size() < (size_t)oriVpNum + cand.size()) h->vps_contri.resize((size_t)oriVpNum + cand.size(), 0);
    // before_vps_voxcount (cpp:276): kept on the device, it only ever comes back as the vector's contents (cpp:307)
    const size_t before_n = h->voxcount_n;
    CU(h->d_snap_vox.reserve(std::max<size_t>(before_n, 1)));
    if (before_n) CU(cudaMemcpyAsync(h->d_snap_vox.p, h->d_voxcount.p, before_n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));

    std::vector<D3> pos(cand.size()), dir(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) {
      pos[i] = D3{cand[i].x, cand[i].y, cand[i].z};
      dir[i] = D3{cand[i].nx, cand[i].ny, cand[i].nz};
    }
    std::vector<Pose5> pose_set_unc;
    umark("  iteration: candidates drawn, snapshots");
    rc = get_pose_batch(h, pos, dir, oriVpNum, pose_set_unc);
    if (rc) return rc;      // the snapshots are released by their destructors
    umark("  iteration: getPose of the candidates");
    h->vps_pose.insert(h->vps_pose.end(), pose_set_unc.begin(), pose_set_unc.end());

    // restore MCI, swap the voxcounts (cpp:304-309)
    CU(cudaMemcpyAsync(h->d_cover_contrib.p, snap_contrib.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_contrib_id.p, snap_id.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    rc = voxcount_download(h, temp_vps_voxcount);
    if (rc) return rc;
    if (h->voxcount_n != before_n) return fail(FCVM_EINTERNAL, "vps_voxcount changed its size inside getPose");
    if (before_n) CU(cudaMemcpyAsync(h->d_voxcount.p, h->d_snap_vox.p, before_n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    umark("  iteration: state restored");
    rc = viewpoints_prune(h, h->vps_pose, temp_vps_voxcount);
    if (rc) return rc;
    umark("  iteration: prune");
    rc = uncovered(uncoveredID);
    if (rc) return rc;
    umark("  iteration: uncovered ids");
    unc_iter++;
  }
  return voxcount_download(h, h->vps_voxcount);
}

}  // namespace fcvm

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

int fcvm_abi_version(void) { return FCVM_ABI_VERSION; }
const char* fcvm_last_error(void) { return g_err.c_str(); }

fcvm_t* fcvm_create(const fcvm_params* p) {
  if (!p) { fail(FCVM_EINVAL, "null params"); return nullptr; }
  fcvm_handle* h = new fcvm_handle();
  h->prm = *p;
  h->pitch_upper = p->pitch_upper;
  h->pitch_lower = p->pitch_lower;
  if (p->attitude == FCVM_ATTITUDE_YAW) { h->pitch_upper = 0.0; h->pitch_lower = 0.0; }   // cpp:50-54
  h->fov_base = std::min(p->fov_h, p->fov_w) * M_PI / 360.0;                               // cpp:55
  h->cam.init(*p);
  int t = p->host_threads;
  if (t <= 0) {      // default: the CPUs this process may run on (a rank bound to its GPU's NUMA node gets that node's cores)
    cpu_set_t set;
    CPU_ZERO(&set);
    t = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
  }
  h->threads = std::max(1, std::min(t, 64));
  const EnvSettings e = env_now();
  h->host_sums = e.host_sums;
  h->own_sweep = e.own_sweep;
  h->use_sat = !e.no_pfree;
  return h;
}

void fcvm_destroy(fcvm_t* h) {
  if (!h) return;
  if (h->device_ok) {
    cudaSetDevice(h->prm.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->aux_stream) cudaStreamSynchronize(h->aux_stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
  }
  if (h->nccl) { nccl_comm_destroy(h->nccl); h->nccl = nullptr; }
  for (auto& se : h->stage_ev_pending) { cudaEventDestroy(se.a); cudaEventDestroy(se.b); }
  for (auto& se : h->stage_ev_free) { cudaEventDestroy(se.a); cudaEventDestroy(se.b); }
  for (int q = 0; q < fcvm_handle::kRing; ++q) { if (h->ring_d2h[q]) cudaEventDestroy(h->ring_d2h[q]); if (h->ring_h2d[q]) cudaEventDestroy(h->ring_h2d[q]); }
  h->pool.reset();
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  for (cudaStream_t st : h->sub_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (h->sub_start) cudaEventDestroy(h->sub_start);
  h->d_map_xyz.free(); h->d_map_sorted.free(); h->d_cell_start.free(); h->d_cell_count.free(); h->d_minmax.free();
  h->d_cand.free(); h->d_ok.free(); h->d_looks.free(); h->d_gather.free();
  h->d_bricks.free(); h->d_points.free(); h->d_cover_contrib.free(); h->d_contrib_id.free(); h->d_voxcount.free();
  h->d_recs.free(); h->d_rows1.free(); h->d_rows2.free(); h->d_count.free(); h->d_sched.free(); h->d_counters.free();
  h->h_rows.free(); h->h_count.free(); h->h_recs.free(); h->d_offsets.free(); h->d_indices.free(); h->h_indices.free(); h->h_offsets.free();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (cudaEvent_t e : h->pipe_ev) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int fcvm_reset(fcvm_t* h) {   // cpp:64-72
  if (!h) return fail(FCVM_EINVAL, "null handle");
  h->map_flag = h->model_flag = h->viewpoints_flag = h->normal_flag = false;
  h->vps_pose.clear(); h->vps_voxcount.clear(); h->vps_contri.clear();
  h->last_vps_num = 0; h->voxcount_n = 0;
  h->pt_normal_pairs.clear();
  h->final_vps.clear();
  h->idx_slot.clear();   // VPP.reset(): clear(), buckets kept
  h->n_map = 0;
  h->trace.clear();
  for (int k = 0; k < FCVM_NCOUNTERS; ++k) h->counters[k] = 0;
  h->unc_calls = 0; h->sample_calls = 0;
  h->t_eval = h->t_own = h->t_init = h->t_update = h->t_unc = h->t_sums = h->t_chain = h->t_safety = h->t_pose = h->t_comm = 0;
  for (int k = 0; k < 4; ++k) { h->stage_ms[k] = 0; h->comm_stats[k] = 0; }
  h->pose_rays = h->pose_flagged = 0;
  h->resident_n = 0;
  if (h->device_ok && h->P > 0) {
    int rc = ensure_device(h);
    if (rc) return rc;
    return cover_reset(h);
  }
  return FCVM_OK;
}

int fcvm_set_map_cloud(fcvm_t* h, const float* xyz, int64_t n) {   // cpp:74-89
  if (!h || !xyz || n <= 0) return fail(FCVM_EINVAL, "empty map cloud");
  double t_dev = 0, t_grid = 0, t_index = 0;
  int rc;
  { Stopwatch sw(t_dev); rc = ensure_device(h); }
  if (rc) return rc;
  cudaStream_t s = h->stream;
  {
    // initHCMap on the device (sdf_map.cpp:140-187): bounding box, geometry, one bit per occupied voxel
    Stopwatch sw(t_grid);
    CU(h->d_map_xyz.reserve((size_t)3 * n));
    CU(cudaMemcpyAsync(h->d_map_xyz.p, xyz, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(296, (n + 255) / 256));
    CU(h->d_minmax.reserve((size_t)6 * blocks));
    CU(launch_cloud_minmax(h->d_map_xyz.p, n, h->d_minmax.p, blocks, s));
    std::vector<float> part((size_t)6 * blocks);
    CU(cudaMemcpyAsync(part.data(), h->d_minmax.p, part.size() * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int b = 0; b < blocks; ++b)
      for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], part[(size_t)6 * b + c]); mx[c] = std::max(mx[c], part[(size_t)6 * b + 3 + c]); }
    h->grid.set_geometry(h->prm, mn, mx);
    if (h->grid.voxels() <= 0) return fail(FCVM_EINVAL, "degenerate grid");
    if (h->grid.voxels() >= (1ll << 31)) return fail(FCVM_EINVAL, "grid has 2^31 voxels or more: beyond the reference's int addressing (sdf_map.h:266-269)");
    MapGeom g;
    for (int c = 0; c < 3; ++c) { g.org[c] = h->grid.origin[c]; g.dims[c] = h->grid.dims[c]; g.nb[c] = h->grid.ns[c]; }
    g.res_inv = h->grid.res_inv;
    // Guard bands: a fast-path half-ray makes at most guard_steps voxel steps from a voxel inside the map, so even a
    // marcher that runs away (raycast.cpp has no termination guarantee) stays within guard_steps
    // voxels of the map on every axis; its brick index then falls into [-guard, nbricks + guard), which is allocated
    // and reads 'occupied'.  Longer rays (none pass the range test) take the literal, bounds-checked marcher.
    h->guard_steps = 3 * ((int)std::ceil(std::max(h->prm.max_dist, h->prm.visible_range) / (2 * h->grid.res)) + 2);
    h->guard_steps = std::min(h->guard_steps, EVAL_RCP_TABLE - 1);      // the fast path's reciprocal table (longer half-rays take the literal marcher)
    h->brick_guard = (size_t)brick_guard_words(h->guard_steps, h->grid.ns[1], h->grid.ns[2]);
    CU(h->d_bricks.reserve(h->grid.nbricks() + 2 * h->brick_guard));
    CU(cudaMemsetAsync(h->d_bricks.p, 0xFF, h->brick_guard * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(h->d_bricks.p + h->brick_guard, 0, h->grid.nbricks() * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(h->d_bricks.p + h->brick_guard + h->grid.nbricks(), 0xFF, h->brick_guard * sizeof(unsigned long long), s));
    CU(launch_voxelise(h->d_map_xyz.p, n, g, h->d_bricks.p + h->brick_guard, s));
    CU(launch_mark_padding(h->d_bricks.p + h->brick_guard, h->grid.dims, h->grid.ns, s));
    // the free-space table; a grid so large that its offsets would not fit 32 bits goes without (every ray is then marched)
    const size_t n_sat = (size_t)(h->grid.ns[0] / FCVM_SAT_BRICKS + 1) * (h->grid.ns[1] / FCVM_SAT_BRICKS + 1) * (h->grid.ns[2] / FCVM_SAT_BRICKS + 1);
    h->sat_built = n_sat < ((size_t)1 << 31);
    if (h->sat_built) {
      CU(h->d_sat.reserve(n_sat));
      CU(launch_build_sat(h->d_bricks.p + h->brick_guard, h->grid.ns, h->d_sat.p, s));
    }
    h->counters[5] += 7;

    // uniform cell index over the map cloud for nearCheck: cells well below the radius so that far cells are never visited
    Stopwatch sw2(t_index);
    CellGeom& cg = h->cells;
    cg.cell = std::max(std::max(h->prm.safe_radius / 4, 2 * h->prm.resolution), 1e-6);
    for (int c = 0; c < 3; ++c) cg.mn[c] = mn[c];
    for (;;) {
      double tot = 1;
      for (int c = 0; c < 3; ++c) { cg.nc[c] = (int)std::floor(((double)mx[c] - mn[c]) / cg.cell) + 1; tot *= cg.nc[c]; }
      if (tot <= 4.0e6) break;
      cg.cell *= 1.5;
    }
    const int64_t ncell = (int64_t)cg.nc[0] * cg.nc[1] * cg.nc[2];
    CU(h->d_cell_count.reserve((size_t)ncell));
    CU(h->d_cell_start.reserve((size_t)ncell + 1));
    CU(h->d_map_sorted.reserve((size_t)n));
    CU(cudaMemsetAsync(h->d_cell_count.p, 0, (size_t)ncell * sizeof(int), s));
    CU(launch_cell_count(h->d_map_xyz.p, n, cg, h->d_cell_count.p, s));
    CU(launch_exscan(h->d_cell_count.p, ncell, h->d_cell_start.p, s));
    CU(cudaMemsetAsync(h->d_cell_count.p, 0, (size_t)ncell * sizeof(int), s));
    CU(launch_cell_fill(h->d_map_xyz.p, n, cg, h->d_cell_start.p, h->d_cell_count.p, h->d_map_sorted.p, s));
    h->counters[5] += 3;
    CU(cudaStreamSynchronize(s));
    h->n_map = n;
  }
  if (env().profile)
    std::fprintf(stderr, "[fcvm] setMapPointCloud: device bring-up %.1f ms, upload + grid + cell index on the device %.1f ms\n", t_dev, t_grid);
  h->map_flag = true;
  return FCVM_OK;
}

int fcvm_set_model(fcvm_t* h, const float* xyz, int64_t n) {   // cpp:91-108
  if (!h || !xyz || n <= 0) return fail(FCVM_EINVAL, "[ViewpointManager] Input model is empty!!");
  int rc = ensure_device(h);
  if (rc) return rc;
  h->model.assign(xyz, xyz + 3 * n);
  const bool grown = n != h->P;
  h->resident_n = 0;          // a staged batch was sized for the previous model (row stride W)
  h->P = n;
  h->W = ((n + 31) / 32 + 3) / 4 * 4;
  // one extra tile of slack so the TMA bulk copy of the last tile never needs a tail case
  std::vector<float4> pts((size_t)n);
  for (int64_t i = 0; i < n; ++i) pts[(size_t)i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
  CU(h->d_points.reserve((size_t)n));
  CU(cudaMemcpyAsync(h->d_points.p, pts.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
  CU(h->d_pt_ends.reserve((size_t)n));
  CU(launch_ray_ends(nullptr, 0, h->d_points.p, n, h->prm.resolution, h->d_pt_ends.p, h->stream));   // resolution_ = hcmap/resolution (sdf_map.cpp:128)
  CU(cudaStreamSynchronize(h->stream));
  if (grown || !h->model_flag) {   // MCI resize (keeps values when the size is unchanged, like vector::resize)
    rc = cover_reset(h);
    if (rc) return rc;
  }
  if (n <= (1 << 16)) {
    // The page-locked staging a first updateViewpoints would otherwise allocate call by call (one candidate viewpoint per model
    // point is what the reference generates, cpp:145-171): cudaMallocHost costs 0.1 - 3 ms apiece, tens of ms on a loaded host.
    // Reserved here, where the model size becomes known; larger models allocate on demand (their first call is dominated by work).
    const size_t v = (size_t)n + 128;
    CU(h->h_recs.reserve(v)); CU(h->h_aux.reserve(v)); CU(h->h_count.reserve(v)); CU(h->h_offsets.reserve(v + 16));
    CU(h->h_own_ids.reserve(v)); CU(h->h_own_sched.reserve(v)); CU(h->h_sums.reserve(2 * v)); CU(h->h_nflag.reserve(2));
    CU(h->h_rows.reserve((size_t)kSmallRowsWords));
    CU(h->h_indices.reserve((size_t)64 * v));
    for (int q = 0; q < fcvm_handle::kRing; ++q) CU(h->h_ring[q].reserve((size_t)16 * v));
  }
  h->model_flag = true;
  return FCVM_OK;
}

int fcvm_set_normals(fcvm_t* h, const double* pt, const double* nrm, int64_t n) {   // cpp:110-120
  if (!h || !pt || !nrm || n <= 0) return fail(FCVM_EINVAL, "[ViewpointManager] Input normal Fail!! Input normal is empty!!!");
  h->pt_normal_pairs.clear();
  for (int64_t i = 0; i < n; ++i)
    h->pt_normal_pairs.insert(std::make_pair(D3{pt[3 * i], pt[3 * i + 1], pt[3 * i + 2]}, D3{nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]}));
  h->normal_flag = true;
  return FCVM_OK;
}

int fcvm_set_init_viewpoints(fcvm_t* h, const float* pos_dir, int64_t n) {   // cpp:122-141
  if (!h || !pos_dir || n <= 0) return fail(FCVM_EINVAL, "[ViewpointManager] Input viewpoints Fail!! Input viewpoint is empty!!!");
  std::vector<PtNormal> v((size_t)n);
  for (int64_t i = 0; i < n; ++i)
    v[(size_t)i] = PtNormal{pos_dir[6 * i], pos_dir[6 * i + 1], pos_dir[6 * i + 2], pos_dir[6 * i + 3], pos_dir[6 * i + 4], pos_dir[6 * i + 5]};
  Stopwatch sw(h->t_init);
  return set_init_viewpoints_impl(h, v);
}

int fcvm_update(fcvm_t* h) {   // cpp:143-330
  if (!h) return fail(FCVM_EINVAL, "null handle");
  int rc;
  {
    Stopwatch sw(h->t_update);
    rc = update_impl(h);
  }
  if (env().profile)
    std::fprintf(stderr, "[fcvm] rank %d/%d: setInitViewpoints %.1f ms, updateViewpoints %.1f ms; of which evaluator batches (H2D + kernels + D2H + sync, "
                 "incl. the pose-sum pipeline) %.1f ms, ownership %.1f ms, uncovered-candidate generation %.1f ms, host libm of the pose sums %.1f ms, "
                 "pose-sum pipeline %.1f ms, safety batches %.1f ms, collectives (enqueue) %.1f ms; device time E1 %.1f E2 %.1f E3 %.1f E4 %.1f ms; "
                 "rays summed on the device %lld, of which flagged for the host's libm %lld; kernels launched %lld\n",
                 h->rank, h->world, h->t_init, h->t_update, h->t_eval, h->t_own, h->t_unc, h->t_sums, h->t_pose, h->t_safety, h->t_comm, h->stage_ms[0],
                 h->stage_ms[1], h->stage_ms[2], h->stage_ms[3], (long long)h->pose_rays, (long long)h->pose_flagged, (long long)h->counters[5]);
  return rc;
}

int64_t fcvm_get_viewpoints(fcvm_t* h, int32_t* vp_id, double* pose5, int32_t* vox_count, int64_t cap) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  const int64_t n = (int64_t)h->final_vps.size();
  for (int64_t i = 0; i < n && i < cap; ++i) {
    if (vp_id) vp_id[i] = h->final_vps[(size_t)i].vp_id;
    if (pose5) for (int k = 0; k < 5; ++k) pose5[5 * i + k] = h->final_vps[(size_t)i].pose.v[k];
    if (vox_count) vox_count[i] = h->final_vps[(size_t)i].vox_count;
  }
  return n;
}

int64_t fcvm_get_uncovered(fcvm_t* h, float* xyz, int64_t cap) {   // cpp:1138-1144
  if (!h) return fail(FCVM_EINVAL, "null handle");
  if (!h->device_ok || h->P == 0) return 0;
  std::vector<int> cid((size_t)h->P);
  if (ensure_device(h)) return FCVM_ECUDA;
  if (cudaMemcpyAsync(cid.data(), h->d_contrib_id.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
      cudaStreamSynchronize(h->stream) != cudaSuccess)
    return fail(FCVM_ECUDA, "download of contrib_id failed");
  int64_t n = 0;
  for (int64_t i = 0; i < h->P; ++i)
    if (cid[(size_t)i] < 0) {
      if (xyz && n < cap) { xyz[3 * n] = h->model[3 * (size_t)i]; xyz[3 * n + 1] = h->model[3 * (size_t)i + 1]; xyz[3 * n + 2] = h->model[3 * (size_t)i + 2]; }
      ++n;
    }
  return n;
}

int64_t fcvm_get_vps_pose(fcvm_t* h, double* pose5, int64_t cap_rows) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  for (int64_t i = 0; i < (int64_t)h->vps_pose.size() && i < cap_rows; ++i)
    for (int k = 0; k < 5; ++k) pose5[5 * i + k] = h->vps_pose[(size_t)i].v[k];
  return (int64_t)h->vps_pose.size();
}
int64_t fcvm_get_vps_voxcount(fcvm_t* h, int32_t* out, int64_t cap) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  if (h->device_ok) {
    int rc = ensure_device(h);
    if (rc) return rc;
    rc = voxcount_download(h, h->vps_voxcount);
    if (rc) return rc;
  }
  for (int64_t i = 0; i < (int64_t)h->vps_voxcount.size() && i < cap; ++i) out[i] = h->vps_voxcount[(size_t)i];
  return (int64_t)h->vps_voxcount.size();
}
int64_t fcvm_get_vps_contri(fcvm_t* h, int32_t* out, int64_t cap) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  for (int64_t i = 0; i < (int64_t)h->vps_contri.size() && i < cap; ++i) out[i] = h->vps_contri[(size_t)i];
  return (int64_t)h->vps_contri.size();
}
int64_t fcvm_get_cover_state(fcvm_t* h, uint8_t* covered, int32_t* cover_contrib, int32_t* contrib_id, int64_t cap) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  if (!h->device_ok || h->P == 0 || cap <= 0) return h->P;
  if (ensure_device(h)) return FCVM_ECUDA;
  std::vector<int> cc((size_t)h->P), ci((size_t)h->P);
  if (cudaMemcpyAsync(cc.data(), h->d_cover_contrib.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
      cudaMemcpyAsync(ci.data(), h->d_contrib_id.p, (size_t)h->P * sizeof(int), cudaMemcpyDeviceToHost, h->stream) != cudaSuccess ||
      cudaStreamSynchronize(h->stream) != cudaSuccess)
    return fail(FCVM_ECUDA, "download of cover state failed");
  for (int64_t i = 0; i < h->P && i < cap; ++i) {
    if (covered) covered[i] = ci[(size_t)i] >= 0 ? 1 : 0;
    if (cover_contrib) cover_contrib[i] = cc[(size_t)i];
    if (contrib_id) contrib_id[i] = ci[(size_t)i];
  }
  return h->P;
}

int fcvm_get_grid(fcvm_t* h, double* origin3, int32_t* voxel_num3, uint8_t* bits, int64_t cap_bytes) {
  if (!h || !h->grid.ready) return fail(FCVM_ENOTREADY, "no map");
  for (int c = 0; c < 3; ++c) { if (origin3) origin3[c] = h->grid.origin[c]; if (voxel_num3) voxel_num3[c] = h->grid.dims[c]; }
  if (bits) {
    if (cap_bytes < (h->grid.voxels() + 7) / 8) return fail(FCVM_ECAP, "grid buffer too small");
    int rc = ensure_device(h);
    if (rc) return rc;
    h->grid.bricks.resize(h->grid.nbricks());
    CU(cudaMemcpyAsync(h->grid.bricks.data(), h->d_bricks.p + h->brick_guard, h->grid.nbricks() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->grid.export_xmajor(bits);
    h->grid.bricks.clear();
    h->grid.bricks.shrink_to_fit();
  }
  return FCVM_OK;
}

int fcvm_safety_check(fcvm_t* h, const float* xyz, int64_t n, uint8_t* ok) {
  if (!h || !xyz || n <= 0 || !ok) return fail(FCVM_EINVAL, "bad arguments");
  if (!h->map_flag) return fail(FCVM_ENOTREADY, "safety check needs the map");
  {
    int rc = ensure_device(h);
    if (rc) return rc;
  }
  std::vector<float> q(xyz, xyz + 3 * n);
  std::vector<uint8_t> out;
  int rc = safety_batch(h, q, 0, out);
  if (rc) return rc;
  std::memcpy(ok, out.data(), (size_t)n);
  return FCVM_OK;
}

int fcvm_trace_enable(fcvm_t* h, int32_t on) { if (!h) return FCVM_EINVAL; h->trace_on = on != 0; if (!on) h->trace.clear(); return FCVM_OK; }
int64_t fcvm_trace_count(fcvm_t* h) { return h ? (int64_t)h->trace.size() : 0; }
int64_t fcvm_trace_words(fcvm_t* h) { return h ? h->W : 0; }
int fcvm_trace_get(fcvm_t* h, int64_t i, int32_t* stage, int32_t* vp_id, double* pose5, uint32_t* row_words) {
  if (!h || i < 0 || i >= (int64_t)h->trace.size()) return fail(FCVM_EINVAL, "trace index out of range");
  const TraceEv& e = h->trace[(size_t)i];
  if (stage) *stage = e.stage;
  if (vp_id) *vp_id = e.vp_id;
  if (pose5) for (int k = 0; k < 5; ++k) pose5[k] = e.pose.v[k];
  if (row_words) std::memcpy(row_words, e.row.data(), e.row.size() * sizeof(uint32_t));
  return FCVM_OK;
}
int fcvm_get_counters(fcvm_t* h, int64_t* out8) {
  if (!h || !out8) return fail(FCVM_EINVAL, "null");
  for (int k = 0; k < FCVM_NCOUNTERS; ++k) out8[k] = h->counters[k];
  return FCVM_OK;
}

// ---------------------------------------------------------------- evaluator level
int fcvm_evaluate(fcvm_t* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows, int32_t* count, uint32_t* rows) {
  if (!h || !pose5 || n <= 0 || stage < 1 || stage > 4) return fail(FCVM_EINVAL, "bad arguments");
  if (!h->map_flag || !h->model_flag) return fail(FCVM_ENOTREADY, "evaluate needs the map and the model");
  int rc = ensure_device(h);
  if (rc) return rc;
  const int64_t W = h->W;
  for (int64_t c0 = 0; c0 < n; c0 += chunk_rows(h, n)) {
    const int64_t cn = std::min(chunk_rows(h, n), n - c0);
    CU(h->h_recs.reserve((size_t)cn));
    pfor(h, cn, 128, [&](int64_t i) {
      const double* p = pose5 + 5 * (c0 + i);
      h->cam.fill(h->h_recs.p[i], p[0], p[1], p[2], p[3], p[4]);
    });
    rc = eval_batch(h, stage, h->h_recs.p, cn, restrict_rows ? restrict_rows + c0 * W : nullptr, rows ? rows + c0 * W : nullptr,
                    count ? count + c0 : nullptr, h->d_rows1);
    if (rc) return rc;
  }
  return FCVM_OK;
}

int64_t fcvm_evaluate_csr(fcvm_t* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows, int32_t* count,
                          int64_t* offsets, uint32_t* indices, int64_t cap) {
  if (!h || !pose5 || n <= 0 || stage < 1 || stage > 4 || !offsets) return fail(FCVM_EINVAL, "bad arguments");
  if (!h->map_flag || !h->model_flag) return fail(FCVM_ENOTREADY, "evaluate needs the map and the model");
  int rc = ensure_device(h);
  if (rc) return rc;
  const int64_t W = h->W;
  int64_t total = 0;
  offsets[0] = 0;
  for (int64_t c0 = 0; c0 < n; c0 += chunk_rows(h, n)) {
    const int64_t cn = std::min(chunk_rows(h, n), n - c0);
    CU(h->h_recs.reserve((size_t)cn));
    pfor(h, cn, 128, [&](int64_t i) {
      const double* p = pose5 + 5 * (c0 + i);
      h->cam.fill(h->h_recs.p[i], p[0], p[1], p[2], p[3], p[4]);
    });
    CsrOut csr;
    if (indices && total < cap) { csr.user_idx = indices + total; csr.user_cap = cap - total; }
    rc = eval_batch(h, stage, h->h_recs.p, cn, restrict_rows ? restrict_rows + c0 * W : nullptr, nullptr, count ? count + c0 : nullptr,
                    h->d_rows1, false, &csr);
    if (rc) return rc;
    const long long ct = csr.offs[(size_t)cn];
    for (int64_t i = 0; i < cn; ++i) offsets[c0 + i + 1] = total + csr.offs[(size_t)i + 1];
    if (indices && total < cap && !csr.in_user) std::memcpy(indices + total, csr.idx, (size_t)std::min<long long>(ct, cap - total) * sizeof(uint32_t));
    total += ct;
  }
  return total;
}

int fcvm_stage_poses(fcvm_t* h, const double* pose5, int64_t n) {
  if (!h || !pose5 || n <= 0) return fail(FCVM_EINVAL, "bad arguments");
  if (!h->map_flag || !h->model_flag) return fail(FCVM_ENOTREADY, "needs the map and the model");
  int rc = ensure_device(h);
  if (rc) return rc;
  CU(h->h_recs.reserve((size_t)n));
  pfor(h, n, 128, [&](int64_t i) { h->cam.fill(h->h_recs.p[i], pose5[5 * i], pose5[5 * i + 1], pose5[5 * i + 2], pose5[5 * i + 3], pose5[5 * i + 4]); });
  CU(h->d_recs.reserve((size_t)n));
  CU(h->d_rows1.reserve((size_t)n * h->W));
  CU(h->d_count.reserve((size_t)n));
  CU(cudaMemcpyAsync(h->d_recs.p, h->h_recs.p, (size_t)n * sizeof(VpRec), cudaMemcpyHostToDevice, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->resident_n = n;
  h->resident_W = h->W;
  h->resident_e1 = false;
  return FCVM_OK;
}

int fcvm_evaluate_resident(fcvm_t* h, int32_t stage, void* stream, int32_t sync, int64_t* rays_out) {
  if (!h || stage < 1 || stage > 4) return fail(FCVM_EINVAL, "bad stage");
  if (h->resident_n <= 0 || h->resident_W != h->W) return fail(FCVM_ENOTREADY, "no resident batch: call fcvm_stage_poses first (after setModel)");
  if (stage == FCVM_STAGE_E2 && !h->resident_e1) return fail(FCVM_ENOTREADY, "E2 re-tests the visible sets of an E1 pass over the same resident batch: run stage 1 first");
  int rc = ensure_device(h);
  if (rc) return rc;
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  if (stage == FCVM_STAGE_E2) {
    // the E1 rows stay in the first rows buffer; E2's rows (and counts) go to the second one
    CU(h->d_rows2.reserve((size_t)h->resident_n * h->W));
    rc = launch_stage(h, stage, h->d_recs.p, h->resident_n, h->d_rows2.p, h->d_rows1.p, s, sync != 0);
    if (rc) return rc;
    CU(launch_popc_rows(h->d_rows2.p, h->resident_n, h->W, h->d_count.p, s));
  } else {
    rc = launch_stage(h, stage, h->d_recs.p, h->resident_n, h->d_rows1.p, nullptr, s, sync != 0);
    if (rc) return rc;
    CU(launch_popc_rows(h->d_rows1.p, h->resident_n, h->W, h->d_count.p, s));
    h->resident_e1 = stage == FCVM_STAGE_E1;
  }
  h->counters[5] += 1;
  if (sync) {
    CU(cudaStreamSynchronize(s));
    const int64_t rays_before = h->counters[1];
    if (s != h->stream) CU(cudaStreamSynchronize(h->stream));
    rc = fetch_counters(h);
    if (rc) return rc;
    if (rays_out) *rays_out = h->counters[1] - rays_before;
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_kernel_ms = ms;
  }
  return FCVM_OK;
}

int fcvm_resident_buffers(fcvm_t* h, void** rows_dev, void** count_dev, int64_t* n, int64_t* words) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  if (rows_dev) *rows_dev = h->d_rows1.p;
  if (count_dev) *count_dev = h->d_count.p;
  if (n) *n = h->resident_n;
  if (words) *words = h->W;
  return FCVM_OK;
}

int fcvm_own_resident(fcvm_t* h, const int32_t* vp_ids, const int32_t* order, int64_t n) {
  if (!h || !vp_ids || n <= 0 || n > h->resident_n) return fail(FCVM_EINVAL, "bad arguments");
  int rc = ensure_device(h);
  if (rc) return rc;
  if (h->world > 1) return fail(FCVM_EINVAL, "fcvm_own_resident works on an unsharded resident batch");
  std::vector<int> cnt((size_t)n), ids(vp_ids, vp_ids + n), ord((size_t)n);
  std::vector<char> seen((size_t)n, 0);
  int max_id = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int r = order ? order[i] : (int)i;
    if (r < 0 || r >= n || seen[(size_t)r]) return fail(FCVM_EINVAL, "order must be a permutation of the row indices 0..n-1");
    if (ids[(size_t)i] < 0) return fail(FCVM_EINVAL, "negative viewpoint id");
    seen[(size_t)r] = 1;
    ord[(size_t)i] = r;
    max_id = std::max(max_id, ids[(size_t)i]);
  }
  rc = voxcount_resize(h, std::max(h->voxcount_n, (size_t)max_id + 1));
  if (rc) return rc;
  if (h->vps_contri.size() < (size_t)max_id + 1) h->vps_contri.resize((size_t)max_id + 1, 0);
  return ownership(h, h->d_rows1, n, 0, n, ord, ids, cnt.data());
}

double fcvm_last_kernel_ms(fcvm_t* h) { return h ? h->last_kernel_ms : 0.0; }

int64_t fcvm_debug_bounds_violations(fcvm_t* h) {
#ifdef FCVM_DEBUG_BOUNDS
  return h ? h->bounds_violations : -1;
#else
  (void)h;
  return -1;    // not a checking build
#endif
}

int fcvm_set_shard(fcvm_t* h, int32_t rank, int32_t world, fcvm_comm_fn fn, void* user) {
  if (!h || world < 1 || rank < 0 || rank >= world) return fail(FCVM_EINVAL, "bad shard");
  if (world > 1 && !fn) return fail(FCVM_EINVAL, "world > 1 needs the collectives callback");
  if (h->nccl) { nccl_comm_destroy(h->nccl); h->nccl = nullptr; }
  h->rank = rank; h->world = world; h->comm = world > 1 ? fn : nullptr; h->comm_user = user;
  return FCVM_OK;
}

int fcvm_nccl_unique_id(uint8_t* id128) {
  if (!id128) return fail(FCVM_EINVAL, "null");
  return nccl_unique_id(id128);
}

int fcvm_set_shard_nccl(fcvm_t* h, const uint8_t* id128, int32_t rank, int32_t world) {
  if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(FCVM_EINVAL, "bad shard");
  int rc = ensure_device(h);          // the communicator binds to the handle's device
  if (rc) return rc;
  if (h->nccl) { nccl_comm_destroy(h->nccl); h->nccl = nullptr; }
  h->rank = 0; h->world = 1; h->comm = nullptr; h->comm_user = nullptr;
  if (world == 1) return FCVM_OK;
  NcclComm* c = nullptr;
  rc = nccl_comm_create(id128, rank, world, &c);
  if (rc) return rc;
  h->nccl = c;
  h->rank = rank; h->world = world; h->comm = nccl_comm_call; h->comm_user = c;
  return FCVM_OK;
}

int fcvm_comm_stats(fcvm_t* h, int64_t* out4) {
  if (!h || !out4) return fail(FCVM_EINVAL, "null");
  for (int k = 0; k < 4; ++k) out4[k] = h->comm_stats[k];
  return FCVM_OK;
}

int fcvm_get_timers(fcvm_t* h, double* out) {
  if (!h || !out) return fail(FCVM_EINVAL, "null");
  const double v[FCVM_NTIMERS] = {h->t_init, h->t_update, h->t_eval, h->t_own, h->t_unc, h->t_sums, h->t_safety, h->t_pose, h->t_comm,
                                  h->stage_ms[0], h->stage_ms[1], h->stage_ms[2], h->stage_ms[3]};
  for (int k = 0; k < FCVM_NTIMERS; ++k) out[k] = v[k];
  return FCVM_OK;
}

void* fcvm_host_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) { fail(FCVM_EINVAL, "bad size"); return nullptr; }
  cudaError_t e = cudaMallocHost(&p, (size_t)bytes);
  if (e != cudaSuccess) { fail(e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver ? FCVM_ENODEVICE : FCVM_ECUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); return nullptr; }
  return p;
}
void fcvm_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------- host-only helpers
int fcvm_host_fov_normals(const fcvm_params* p, double pitch, double yaw, double* out12) {
  if (!p || !out12) return fail(FCVM_EINVAL, "null");
  Camera c;
  c.init(*p);
  double n[4][3];
  c.world_normals(pitch, yaw, n);
  for (int q = 0; q < 4; ++q) for (int i = 0; i < 3; ++i) out12[3 * q + i] = n[q][i];
  return FCVM_OK;
}

int fcvm_host_build_grid(const fcvm_params* p, const float* xyz, int64_t n, double* origin3, int32_t* voxel_num3, uint8_t* bits, int64_t cap_bytes) {
  if (!p || !xyz || n <= 0) return fail(FCVM_EINVAL, "empty cloud");
  HostGrid g;
  g.build(*p, xyz, n);
  for (int c = 0; c < 3; ++c) { if (origin3) origin3[c] = g.origin[c]; if (voxel_num3) voxel_num3[c] = g.dims[c]; }
  if (bits) {
    if (cap_bytes < (g.voxels() + 7) / 8) return fail(FCVM_ECAP, "grid buffer too small");
    g.export_xmajor(bits);
  }
  return FCVM_OK;
}

int fcvm_shard_range(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi) {
  if (world < 1 || rank < 0 || rank >= world || n < 0) return fail(FCVM_EINVAL, "bad shard");
  const int64_t per = (n + world - 1) / world;
  if (lo) *lo = std::min(n, per * rank);
  if (hi) *hi = std::min(n, per * (rank + 1));
  return FCVM_OK;
}

}  // extern "C"
