/*
 * fcvm.h — C ABI of the B200-native viewpoint visibility / coverage-pruning path.
 *
 * This is the drop-in boundary behind FC-Planner's `predrecon::ViewpointManager`
 * (reference: FC-Planner/src/viewpoint_manager/include/viewpoint_manager/viewpoint_manager.h:180-193).
 * Every manager-level entry point below replaces exactly one public method of that class;
 * the file:line of the method it replaces is quoted next to it.  A thin C++ shim with the
 * reference's own signatures (fc-planner_b200/shim/viewpoint_manager/viewpoint_manager.h) and a ctypes
 * mirror (fc-planner_b200/manager.py) forward to these symbols.  See INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; host pointers in, host pointers out unless a name ends
 *     in `_dev` / a parameter is documented as a device pointer;
 *   - return 0 on success, a negative FCVM_E* code on failure; the message is available
 *     from fcvm_last_error(); nothing throws across the ABI (the reference's convention is
 *     "ROS_ERROR + return", viewpoint_manager.cpp:93-96, 112-116, 124-128, 173-174);
 *   - single-threaded contract per handle, as the reference (stateful RayCaster /
 *     PerceptionUtils members, viewpoint_manager.h:214-216);
 *   - there is NO CPU fallback: every call that evaluates visibility fails with
 *     FCVM_ENODEVICE when no CUDA device is usable.
 */
#ifndef FCVM_H_
#define FCVM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCVM_ABI_VERSION 2

/* error codes */
#define FCVM_OK          0
#define FCVM_EINVAL     -1   /* bad argument / empty input (reference: ROS_ERROR + return)   */
#define FCVM_ENOTREADY  -2   /* TaskState::isReady() false (viewpoint_manager.h:102-125)     */
#define FCVM_ENODEVICE  -3   /* no usable CUDA device: the product path has no CPU fallback  */
#define FCVM_ECUDA      -4   /* a CUDA runtime call or kernel failed                          */
#define FCVM_ECAP       -5   /* caller buffer too small; required size returned where stated */
#define FCVM_EINTERNAL  -6

/* attitude_type (viewpoint_manager.cpp:45, 50-54, 951-952, 771) */
#define FCVM_ATTITUDE_OTHER 0
#define FCVM_ATTITUDE_YAW   1
#define FCVM_ATTITUDE_ALL   2

/* evaluator kinds, SURVEY.md section 3.3 */
#define FCVM_STAGE_E1 1  /* getPose pass 1:  FOV + BiRC(discard first) + range            (cpp:355-425) */
#define FCVM_STAGE_E2 2  /* getPose pass 2:  E1 set ∧ FOV' + offset walk + BiRC(discard)   (cpp:457-498) */
#define FCVM_STAGE_E3 3  /* gravitation:     FOV + BiRC(discard first), no range re-check  (cpp:817-886) */
#define FCVM_STAGE_E4 4  /* updateViewpointInfo: FOV + offset walk + BiRC(no discard)      (cpp:533-618) */

/*
 * Everything the reference reads from the ROS parameter server on this path
 * (viewpoint_manager.cpp:35-47, perception_utils.cpp:32-37, sdf_map.cpp:128-136),
 * plus the injected seed that replaces its two wall-clock seeded random sites
 * (viewpoint_manager.cpp:962-977 and 249-255).
 */
typedef struct fcvm_params {
  double visible_range;        /* viewpoint_manager/visible_range (also HC map inflation)  */
  double viewpoints_distance;  /* viewpoint_manager/viewpoints_distance  (dist_vp)          */
  double fov_h, fov_w;         /* degrees                                                    */
  double pitch_upper, pitch_lower; /* degrees                                               */
  double ground_pos;           /* viewpoint_manager/GroundPos                                */
  double safe_height;          /* viewpoint_manager/safeHeight                               */
  double safe_radius;          /* viewpoint_manager/safe_radius                              */
  double top_angle, left_angle, right_angle; /* perception_utils/..., radians               */
  double max_dist;             /* perception_utils/max_dist                                  */
  double vis_dist;             /* perception_utils/vis_dist (visualisation only; unused)     */
  double resolution;           /* hcmap/resolution                                           */
  int32_t z_ground;            /* viewpoint_manager/zGround                                  */
  int32_t attitude;            /* FCVM_ATTITUDE_*                                            */
  int32_t max_iter_num;        /* viewpoint_manager/max_iter_num                             */
  int32_t pose_update;         /* viewpoint_manager/pose_update                              */
  uint64_t seed;               /* injected seed for uncNeighVps + the <=100 random sample    */
  int32_t device;              /* CUDA device ordinal                                        */
  int32_t host_threads;        /* threads for the per-viewpoint host reductions; 0 = auto    */
} fcvm_params;

typedef struct fcvm_handle fcvm_t;

/* ---------------------------------------------------------------- manager level ---- */

int         fcvm_abi_version(void);
const char* fcvm_last_error(void);

/* ViewpointManager() + init(nh)                 viewpoint_manager.cpp:25-62 */
fcvm_t* fcvm_create(const fcvm_params* p);
/* ~ViewpointManager()                           viewpoint_manager.cpp:28-30 */
void    fcvm_destroy(fcvm_t* h);
/* reset()                                       viewpoint_manager.cpp:64-72 */
int     fcvm_reset(fcvm_t* h);
/* setMapPointCloud(cloud)                       viewpoint_manager.cpp:74-89
 * xyz: n packed float triples.  Builds the occupancy grid (sdf_map.cpp:122-188),
 * uploads it bit-packed, builds the map-cloud nearest-neighbour index. */
int     fcvm_set_map_cloud(fcvm_t* h, const float* xyz, int64_t n);
/* setModel(cloud)                               viewpoint_manager.cpp:91-108 */
int     fcvm_set_model(fcvm_t* h, const float* xyz, int64_t n);
/* setNormals(map<Vector3d,Vector3d>)            viewpoint_manager.cpp:110-120
 * n (key, value) pairs; key = exact double coordinates, std::map insert semantics. */
int     fcvm_set_normals(fcvm_t* h, const double* pt_xyz, const double* normal_xyz, int64_t n);
/* setInitViewpoints(PointNormal cloud)          viewpoint_manager.cpp:122-141
 * pos_dir: n rows of 6 floats (x, y, z, normal_x, normal_y, normal_z). */
int     fcvm_set_init_viewpoints(fcvm_t* h, const float* pos_dir, int64_t n);
/* updateViewpoints()                            viewpoint_manager.cpp:143-330 */
int     fcvm_update(fcvm_t* h);
/* getUpdatedViewpoints(vector<SingleViewpoint>&) viewpoint_manager.cpp:1133-1136
 * returns the number of final viewpoints; fills at most cap entries
 * (vp_id[i], pose5[5*i..], vox_count[i]).  Any output pointer may be NULL. */
int64_t fcvm_get_viewpoints(fcvm_t* h, int32_t* vp_id, double* pose5, int32_t* vox_count, int64_t cap);
/* getUncoveredModelCloud(cloud&)                viewpoint_manager.cpp:1138-1144 */
int64_t fcvm_get_uncovered(fcvm_t* h, float* xyz, int64_t cap);

/* ---------------------------------------------------------------- parity taps ------- */

/* state mirrors of VPI / MCI (viewpoint_manager.h:128-158); each returns the element count */
int64_t fcvm_get_vps_pose(fcvm_t* h, double* pose5, int64_t cap_rows);
int64_t fcvm_get_vps_voxcount(fcvm_t* h, int32_t* out, int64_t cap);
int64_t fcvm_get_vps_contri(fcvm_t* h, int32_t* out, int64_t cap);
int64_t fcvm_get_cover_state(fcvm_t* h, uint8_t* covered, int32_t* cover_contrib, int32_t* contrib_id, int64_t cap);

/* grid geometry as initHCMap derives it: origin[3], voxel_num[3]; bits = occupancy in the
 * reference's x-major order, 1 bit per voxel (LSB first), cap_bytes >= ceil(G/8). */
int     fcvm_get_grid(fcvm_t* h, double* origin3, int32_t* voxel_num3, uint8_t* bits, int64_t cap_bytes);

/* viewpointSafetyCheck (viewpoint_manager.cpp:889-914, with nearCheck :916-946) for n candidate positions
 * (packed float triples, as pcl::PointXYZ holds them): ok[i] = 1 if the position is kept.  One kernel launch;
 * counter [7] advances by the safety_check probes the reference would have made. */
int     fcvm_safety_check(fcvm_t* h, const float* xyz, int64_t n, uint8_t* ok);

/* Event log of every evaluator launch, in execution order, for mask-level parity:
 * enable before the calls to be traced.  Entry i has stage[i], vp_id[i] and a bitset row of
 * fcvm_trace_words() 32-bit words (bit j of the row = model point j visible). */
int     fcvm_trace_enable(fcvm_t* h, int32_t on);
int64_t fcvm_trace_count(fcvm_t* h);
int64_t fcvm_trace_words(fcvm_t* h);
int     fcvm_trace_get(fcvm_t* h, int64_t i, int32_t* stage, int32_t* vp_id, double* pose5, uint32_t* row_words);

/* counters accumulated since create/reset (all evaluators):
 * [0] pair tests (insideFOV calls)   [1] rays (BiRC invocations)   [2] voxel lookups
 * [3] offset-walk lookups            [4] step-cap trips (must be 0) [5] kernels launched
 * [6] viewpoint evaluations          [7] safety_check probes of viewpointSafetyCheck (counted on the device, k_safety) */
#define FCVM_NCOUNTERS 8
int     fcvm_get_counters(fcvm_t* h, int64_t* out8);

/* ---------------------------------------------------------------- evaluator level ---- */

/*
 * One evaluator pass over a batch of poses against the resident model cloud and grid:
 * the batched form of the loops at viewpoint_manager.cpp:358-425 (E1), 457-498 (E2),
 * 822-875 (E3), 546-586 (E4).  pose5: n x (x, y, z, pitch, yaw) host doubles.
 * Outputs (host, each may be NULL): count[n] = visible points per pose;
 * rows = n x fcvm_trace_words() bitset words.  restrict_rows (E2 only, may be NULL otherwise)
 * = n x words input bitsets whose set bits are the only points considered.
 * Timed end to end by bench.py (host->device copy, kernel, device->host copy).
 */
int fcvm_evaluate(fcvm_t* h, int32_t stage, const double* pose5, int64_t n,
                  const uint32_t* restrict_rows, int32_t* count, uint32_t* rows);

/* The same pass with the visible sets returned as ascending index lists (CSR): offsets[n+1] and at most
 * cap indices; returns the total number of visible pairs (> cap: `indices` holds the first cap of them; call again with
 * a larger buffer).  Large batches are evaluated in sub-batches whose lists are extracted and copied straight into
 * `indices` while the next sub-batch runs; give it page-locked memory (fcvm_host_alloc) for full PCIe rate.
 * This is the form the per-viewpoint host reductions of getPose / updatePoseGravitation consume
 * (viewpoint_manager.cpp:398-425, 856-875): 4 bytes per visible ray instead of P/8 bytes per pose. */
int64_t fcvm_evaluate_csr(fcvm_t* h, int32_t stage, const double* pose5, int64_t n, const uint32_t* restrict_rows,
                          int32_t* count, int64_t* offsets, uint32_t* indices, int64_t cap);

/* Same pass on poses that are already resident: uploads nothing, downloads nothing, leaves the
 * bitset rows and counts on the device.  fcvm_stage_poses copies/derives the per-pose device
 * record (position + 4 world-frame FOV normals) once; fcvm_evaluate_resident can then be
 * launched repeatedly on `stream` (a cudaStream_t passed as void*, NULL = the handle's stream).
 * rays_out (host, may be NULL) is only read back when sync != 0.  Stage 2 re-tests the visible sets the most recent stage-1 pass
 * over the same resident batch left on the device (same poses; its rows go to a second buffer, the E1 rows stay). */
int fcvm_stage_poses(fcvm_t* h, const double* pose5, int64_t n);
int fcvm_evaluate_resident(fcvm_t* h, int32_t stage, void* stream, int32_t sync, int64_t* rays_out);
/* device pointers of the resident batch: rows (n x words uint32), counts (n int32) */
int fcvm_resident_buffers(fcvm_t* h, void** rows_dev, void** count_dev, int64_t* n, int64_t* words);

/* Ownership pass (viewpoint_manager.cpp:500-528, 589-617) over the resident batch rows in the
 * given processing order, updating the manager's cover state; vp_ids[n] are the ids the rows
 * stand for; order[n] = row indices in processing order (NULL = 0..n-1). */
int fcvm_own_resident(fcvm_t* h, const int32_t* vp_ids, const int32_t* order, int64_t n);

/* Page-locked host memory for the caller's input / output buffers of the evaluator-level calls:
 * copies to and from it run at full PCIe rate and overlap with the kernels (rows of a finished
 * sub-batch are copied out while the next sub-batch is evaluated).  Pageable buffers work too, slower. */
void* fcvm_host_alloc(int64_t bytes);
void  fcvm_host_free(void* p);

/* device time of the most recent visibility kernel (k_eval alone, CUDA events on its stream), in ms;
 * valid after a synchronising call */
double fcvm_last_kernel_ms(fcvm_t* h);

/* Index checks of the visibility kernel: a library built with -DFCVM_DEBUG_BOUNDS (libfcvm_dbg.so)
 * verifies every data-dependent index (queue slots, tile offsets, bitset words, brick indices) and
 * returns the number of violations seen since create; a release build returns -1. */
int64_t fcvm_debug_bounds_violations(fcvm_t* h);

/* --------------------------------------------------- multi-GPU ----------------------- */

/* Viewpoint sharding over the GPUs of one box (north_star (4)).  Every handle of a group is given the SAME
 * inputs through the SAME call sequence; the occupancy grid and the model cloud are replicated, and of every viewpoint
 * batch (setInitViewpoints, the E3 / E4 batches of viewpointsPrune, the round appends) a handle evaluates only the
 * contiguous block fcvm_shard_range(n, rank, world).  Bitset rows never leave the GPU that produced them.  Per stage the
 * ranks exchange
 *   - the per-viewpoint results of their blocks (visible counts: 4 B, averaged pitch / yaw: 24 B per viewpoint): all-gather;
 *   - the ownership rule (viewpoint_manager.cpp:500-528, 589-617) as ONE all-reduce MAX over per-point keys
 *     (contribute_voxel_num << 32 | ~position in processing order; 8 B per model point),
 * after which every rank applies the same winners to its replica of the cover state.  Every value is produced by exactly
 * one rank with the same code, so all ranks end with results bit-identical to the single-GPU run.
 *
 * fcvm_set_shard_nccl: the library drives NCCL itself (libnccl.so.2 is dlopen'ed; collectives are enqueued on the handle's
 * stream).  id128 = an ncclUniqueId made by fcvm_nccl_unique_id on one rank and handed to the others by any means (the
 * caller's launcher, a file, MPI, torch.distributed ...).  Works for one process per GPU and for one thread per GPU in a
 * single process (each handle on its own device, each driven by its own thread: the call blocks until all ranks joined).
 * fcvm_set_shard: the caller supplies the collectives instead (any backend that can run them on device memory). */
#define FCVM_COMM_ALLGATHER      0   /* buf_dev = world equal slices of n BYTES, slice `rank` filled; in place */
#define FCVM_COMM_ALLREDUCE_MAX  1   /* buf_dev = n int64 values; element-wise MAX over the ranks, in place */
typedef int (*fcvm_comm_fn)(void* user, int32_t op, void* buf_dev, int64_t n, void* stream);
int fcvm_set_shard(fcvm_t* h, int32_t rank, int32_t world, fcvm_comm_fn fn, void* user);
int fcvm_nccl_unique_id(uint8_t* id128);
int fcvm_set_shard_nccl(fcvm_t* h, const uint8_t* id128, int32_t rank, int32_t world);
/* collectives issued by this handle since create / reset: out4 = { all-gathers, bytes gathered (all slices),
 * all-reduces, bytes reduced } */
int fcvm_comm_stats(fcvm_t* h, int64_t* out4);

/* Wall-clock split of the manager-level calls since create / reset, in milliseconds (the timers FCVM_PROFILE=1 prints):
 * out[0] setInitViewpoints  [1] updateViewpoints  [2] evaluator batches (H2D + kernels + D2H + sync)  [3] ownership
 * [4] uncovered-candidate generation  [5] host libm on flagged / all rays of the pose sums  [6] safety batches
 * [7] device pose-sum pipeline (terms + patch + sum, incl. its copies)  [8] collectives (host-side wait included)
 * [9] E1 kernels  [10] E2 kernels  [11] E3 kernels  [12] E4 kernels (device time, CUDA events) */
#define FCVM_NTIMERS 13
int fcvm_get_timers(fcvm_t* h, double* out13);

/* --------------------------------------------------- pin-hole coverage evaluation -------
 * the brute-force loop that follows the viewpoint path in a real run -
 * every path / trajectory pose x every point of the full cloud: insideFOV, pin-hole projection, a
 * per-pixel nearest-point z-buffer (hierarchical_coverage_planner/src/hcplanner.cpp).
 * One handle per (camera, cloud); poses are evaluated in batches on the GPU.  No CPU fallback. */
typedef struct fcvm_camera {
  double fx, fy, cx, cy;   /* hcplanner/{fx, fy, cx, cy}            (hcplanner.cpp:56-59) */
  double fov_h, fov_w;     /* degrees, as viewpoint_manager/{fov_h, fov_w} (hcplanner.cpp:1040-1041 use them with M_PI) */
} fcvm_camera;
typedef struct fcvm_cov_handle fcvm_cov_t;

/* p supplies perception_utils/{top,left,right}_angle and max_dist (PerceptionUtils::insideFOV) and device */
fcvm_cov_t* fcvm_cov_create(const fcvm_params* p, const fcvm_camera* cam);
void        fcvm_cov_destroy(fcvm_cov_t* c);
/* the cloud the poses are checked against: `Fullmodel` (hcplanner.cpp:252) / `checkCloud` (:1129) */
int         fcvm_cov_set_cloud(fcvm_cov_t* c, const float* xyz, int64_t n);
/* PinHoleCamera (hcplanner.cpp:1128-1194) for m poses (x, y, z, pitch, yaw): the ids pose k covers are
 * ids[offsets[k] .. offsets[k+1]) ascending (the std::set of :1191); covered[n] (may be NULL) = union over
 * the poses, i.e. `CoverState` of the path-coverage loop (:253-261).  Returns the total number of ids
 * (> cap: fcvm_cov_fetch_ids has them all), negative on error. */
int64_t     fcvm_cov_pinhole(fcvm_cov_t* c, const double* pose5, int64_t m, int64_t* offsets, int32_t* ids, int64_t cap, uint8_t* covered);
/* CoverageEvaluation (hcplanner.cpp:1032-1126) over m trajectory poses: *rate = coverage rate; the
 * points pushed into pose k's `visibleCloud` (visible_group, :1102-1107) are
 * group_ids[group_offsets[k] .. group_offsets[k+1]); ever[n] (may be NULL) = visible in any pose.
 * Returns the total number of group ids (> cap: fcvm_cov_fetch_ids), negative on error. */
int64_t     fcvm_cov_evaluate(fcvm_cov_t* c, const double* pose5, int64_t m, double* rate, int64_t* group_offsets, int32_t* group_ids,
                              int64_t cap, uint8_t* ever);
/* the ids of the most recent fcvm_cov_pinhole / fcvm_cov_evaluate call (kept in the handle, so a caller that did
 * not know the size calls once with ids = NULL and then fetches): copies at most cap ids, returns their number */
int64_t     fcvm_cov_fetch_ids(fcvm_cov_t* c, int32_t* ids, int64_t cap);
/* counters since create: [0] pair tests (insideFOV) [1] projections [2] out-of-image projections (undefined
 * behaviour in the reference: unchecked Eigen index; skipped here) [3] kernels launched [4] chunks of
 * fcvm_cov_chunk_points() consecutive points scanned (the rest were culled whole) [5] points that shared their pixel with
 * another point of the same batch */
int         fcvm_cov_counters(fcvm_cov_t* c, int64_t* out6);
/* points per culling chunk (one bounding box each) */
int32_t     fcvm_cov_chunk_points(void);
/* device time of all kernels of the most recent pinhole / evaluate call, in ms (CUDA events) */
double      fcvm_cov_last_kernel_ms(fcvm_cov_t* c);
/* of which k_cov_scan (FOV + projection + z-buffer), the dominant kernel */
double      fcvm_cov_last_scan_ms(fcvm_cov_t* c);

/* --------------------------------------------------- planner-grid loops ---------------
 * SURVEY.md 8(f) rank 2 (outside the manager) and rank 4: three ray / probe loops of the planner over ITS OWN SDFMap
 * (the manager's private map has no inflate / internal marks), batched on the GPU.  No CPU fallback. */
typedef struct fcvm_grid_handle fcvm_grid_t;

/* The planner's SDFMap as initHCMap + InternalSpace / OuterCheck left it: map_origin_, resolution_, map_voxel_num_ and
 * the three voxel buffers occupancy_buffer_hc_, occupancy_inflate_buffer_hc_, occupancy_buffer_internal_
 * (sdf_map.h:217-227; x-major, sdf_map.h:266-269; any may be NULL = all zero).  Uploaded and bit-packed once. */
fcvm_grid_t* fcvm_grid_create(int32_t device, const double* origin3, double resolution, const int32_t* voxel_num3, const int8_t* occupancy_hc,
                              const int8_t* occupancy_inflate, const int8_t* occupancy_internal);
void         fcvm_grid_destroy(fcvm_grid_t* g);
/* lightStateDet (hcplanner.cpp:805-863) for n candidate viewpoints (FP64 triples) against the ROSA vertex cloud
 * (m float triples): det[i] = 1 if the candidate is outside the structure (odd number of occupied voxels crossed on the
 * way to its nearest vertex); cross[i], nearest[i] may be NULL.  `vpTest = vpCandi + 0.0 * res * normal` is vpCandi. */
int          fcvm_grid_light_state(fcvm_grid_t* g, const double* cand_xyz, int64_t n, const float* rosa_xyz, int64_t m, uint8_t* det,
                                   int32_t* cross, int32_t* nearest);
/* the safety box around sampled viewpoints (hcplanner.cpp:658-676), n float triples (pcl::PointNormal positions) */
int          fcvm_grid_safety_box(fcvm_grid_t* g, const float* vp_xyz, int64_t n, double safe_radius, uint8_t* ok);
/* HCSolver::search_Path's straight-line test (hcsolver.cpp:556-583) for every ordered pair of n positions:
 * safe[i * n + j] = 1 if the segment p_i -> p_j is free (then the path is {p_i, p_j} and its cost dist[i * n + j] =
 * (p_i - p_j).norm()); 0 = the caller falls back to A* as the reference does.  dist may be NULL. */
int          fcvm_grid_line_of_sight(fcvm_grid_t* g, const double* p_xyz, int64_t n, uint8_t* safe_nxn, double* dist_nxn);
/* The buffers can also be built on the device (pass NULL for all three to fcvm_grid_create = the zeroed buffers of
 * SDFMap::initHCMap, sdf_map.cpp:159-172):
 * fcvm_grid_set_occ         SDFMap::SetOcc (sdf_map.cpp:440-455) / the marking loop of initHCMap (:178-187): hc = 1 at the voxel of
 *                           every point that lies in the map.
 * fcvm_grid_internal_space  SDFMap::InternalSpace (sdf_map.cpp:263-395) for all skeleton segments at once: seg_xyz = the points of
 *                           the segment clouds concatenated (segment s = points seg_offsets[s] .. seg_offsets[s+1]), seg_ends6[6 s ..] =
 *                           vertices.row(segments[id][0]) and vertices.row(segments[id][1]); inflate_num, proj_interval, thickness =
 *                           hcmap/{inflateVoxel, interval, plane_thickness} (sdf_map.cpp:129-133).
 * fcvm_grid_outer_check     SDFMap::OuterCheck (sdf_map.cpp:457-473): outers6 = n x (position, direction), check_scale = hcmap/checkScale.
 * fcvm_grid_get_buffers     the three buffers back in the reference's x-major char layout (any pointer may be NULL).
 * The reference writes voxels outside the map unchecked (undefined behaviour); those writes are skipped and counted
 * (fcvm_grid_oob_writes). */
int          fcvm_grid_set_occ(fcvm_grid_t* g, const float* xyz, int64_t n);
int          fcvm_grid_internal_space(fcvm_grid_t* g, const float* seg_xyz, const int64_t* seg_offsets, int64_t nseg, const double* seg_ends6,
                                      int32_t inflate_num, double proj_interval, double thickness);
int          fcvm_grid_outer_check(fcvm_grid_t* g, const double* outers6, int64_t n, double check_scale);
int          fcvm_grid_get_buffers(fcvm_grid_t* g, int8_t* hc, int8_t* inflate, int8_t* internal);
int64_t      fcvm_grid_oob_writes(fcvm_grid_t* g);
/* [0] voxel lookups / probes made [1] marches that hit the step cap (the reference would not terminate) [2] kernels launched */
int          fcvm_grid_counters(fcvm_grid_t* g, int64_t* out3);
double       fcvm_grid_last_kernel_ms(fcvm_grid_t* g);

/* --------------------------------------------------- host-only helpers (CPU tests) --- */

/* PerceptionUtils::init + setPose_PY (perception_utils.cpp:29-96): 4 world-frame inward normals
 * (top, bottom, left, right) for (pitch, yaw); out12 = 4 x 3 doubles. */
int fcvm_host_fov_normals(const fcvm_params* p, double pitch, double yaw, double* out12);
/* initHCMap geometry + voxelisation on the host, no device needed (sdf_map.cpp:122-188) */
int fcvm_host_build_grid(const fcvm_params* p, const float* xyz, int64_t n,
                         double* origin3, int32_t* voxel_num3, uint8_t* bits, int64_t cap_bytes);
/* contiguous shard [lo, hi) of n items for (rank, world) */
int fcvm_shard_range(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi);

#ifdef __cplusplus
}
#endif
#endif /* FCVM_H_ */
