// fcvm_kernels.h — host-visible launch interface of the sm_100a kernels (fcvm_kernels.cu).
#pragma once

#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "fcvm_device.cuh"

namespace fcvm {

#ifndef EVAL_WARPS
#define EVAL_WARPS 8          // warps per k_eval CTA
#endif
#ifndef EVAL_MIN_CTAS
#define EVAL_MIN_CTAS 2       // resident CTAs per SM the register allocation is tuned for
#endif
#ifndef EVAL_KEEP
#define EVAL_KEEP 28          // march: suspend back to the FOV phase when fewer lanes than this are busy and the queue is empty
#endif
#ifndef EVAL_UNROLL
#define EVAL_UNROLL 2         // march: biNextId calls per refill / suspend check (with the hand-off build: 1 -> 24.1 ms per E1 step, 2 -> 23.1, 4 -> 22.9)
#endif
#ifndef EVAL_QUERY_PERIOD
#define EVAL_QUERY_PERIOD 2   // march: every this many rounds (of EVAL_UNROLL steps) the rays in flight test the box between their fronts (power of 2)
#endif
#ifndef EVAL_RCP_TABLE
#define EVAL_RCP_TABLE 1024   // reciprocals 1/k, k < this, for the DDA set-up of the fast path (half-rays of more voxel steps take the literal path)
#endif
#ifndef EVAL_WAVES
#define EVAL_WAVES 12         // CTA waves (over 148 SMs x EVAL_MIN_CTAS) the tiling aims at (3..96 measured: flat from 8 to 24, worse outside)
#endif
#ifndef EVAL_READYQ
#define EVAL_READYQ 64        // ray continuations per warp in shared memory (140 B each)
#endif
#ifndef EVAL_MAX_TILE
#define EVAL_MAX_TILE 1024    // points per CTA tile at large P (16 B each, staged by one TMA bulk copy)
#endif

struct EvalArgs {
  GridDesc grid;
  FovConst fov;
  const float4* points;            // model cloud, float4 (x, y, z, 0) like pcl::PointXYZ; P entries
  long long P;
  const VpRec* recs;               // V viewpoint records
  const RayEnd* vp_ends;           // V: the records' positions / resolution (k_ray_ends)
  const RayEnd* pt_ends;           // P: the model points / resolution (k_ray_ends, once per setModel)
  long long V;
  uint32_t* rows;                  // V x words bitset rows, zeroed by the caller
  const uint32_t* restrict_rows;   // E2 only: V x words rows of the E1 pass
  long long words;                 // row stride in 32-bit words (multiple of 4)
  int tile_pts;                    // points per CTA tile (multiple of 32, <= EVAL_MAX_TILE)
  int vp_per_cta;                  // viewpoints per CTA (grid.y = ceil(V / vp_per_cta))
  unsigned long long* counters;    // [0] pair tests [1] rays [2] lookups [3] walk lookups [4] cap trips; may be null
};

size_t eval_smem_bytes(int tile_pts);
// RayEnd of n positions: xyz = doubles with `stride_d` doubles between positions (VpRec: 16) or, with xyz == nullptr, float4 points
cudaError_t launch_ray_ends(const double* xyz, int stride_d, const float4* pts, long long n, double res, RayEnd* out, cudaStream_t s);
cudaError_t launch_eval(int stage, const EvalArgs& a, cudaStream_t s);
cudaError_t launch_popc_rows(const uint32_t* rows, long long n, long long words, int* count, cudaStream_t s);
cudaError_t launch_exscan(const int* count, long long n, long long* offsets, cudaStream_t s);
cudaError_t launch_compact_rows(const uint32_t* rows, long long n, long long words, const long long* offsets, uint32_t* indices, cudaStream_t s);
// sched[t] = {row index, viewpoint id, contribute_voxel_num, 0} in processing order, rows with 0 omitted
cudaError_t launch_own_sweep(const uint32_t* rows, long long words, const int4* sched, int nsched, int last_vps_num, long long P,
                             int* cover_contrib, int* contrib_id, int* voxcount, cudaStream_t s);
// the same rule as a per-point arg-max (shards over GPUs): best[P] (zeroed by the caller) = largest key over the rows listed in
// sched[t] = {row, contri, ~position in processing order, 0}, which must be sorted by (contri, ~position) descending; then one
// apply pass decodes the winners (id_at_pos[position] = viewpoint id)
cudaError_t launch_own_best(const uint32_t* rows, long long words, const int4* sched, int nsched, long long P, unsigned long long* best, cudaStream_t s);
cudaError_t launch_own_apply(const unsigned long long* best, const int* id_at_pos, int last_vps_num, long long P, int* cover_contrib, int* contrib_id,
                             int* voxcount, cudaStream_t s);

// ---- pitch / yaw sums of getPose on the device (fcvm_pose.cu) -------------------------------------
struct PoseAux {            // per viewpoint: position, fov_sample_ray = dir.normalized(), (x, y) of fov_yaw_ray (cpp:341-345)
  double vx, vy, vz, fsx, fsy, fsz, fyx, fyy;
};
// offsets[nrows + 1] / indices: the visible sets of the rows as ascending index lists; val_*: one double per ray;
// key_* / arg_*: compact lists of the rays whose asin (p) / acos (y) must come from the host's libm, nflag[2] their counts
cudaError_t launch_pose_terms(const float4* points, const PoseAux* aux, const long long* offsets, const uint32_t* indices, int nrows, double* val_p,
                              double* val_y, unsigned long long* key_p, double* arg_p, unsigned long long* key_y, double* arg_y,
                              unsigned long long* nflag, unsigned long long cap, cudaStream_t s);
cudaError_t launch_pose_patch(const unsigned long long* key_p, const double* g_p, long long np, const unsigned long long* key_y, const double* g_y,
                              long long ny, double* val_p, double* val_y, cudaStream_t s);
cudaError_t launch_pose_sum(const long long* offsets, int nrows, const double* val_p, const double* val_y, double* sums, cudaStream_t s);

// ---- map side (fcvm_map.cu) --------------------------------------------------------------------
struct MapGeom {            // SDFMap geometry as initHCMap derives it (sdf_map.cpp:144-157)
  double org[3];            // map_origin_
  double res_inv;           // 1 / resolution_
  int dims[3];              // map_voxel_num_
  int nb[3];                // bricks per axis
};
struct CellGeom {           // uniform cell index over the map cloud
  float mn[3];
  double cell;
  int nc[3];
};
struct SafetyArgs {
  MapGeom geom;
  const unsigned long long* bricks;
  const int* sat;                  // summed-area table of occupied cells (GridDesc::sat), may be null
  int sat_ny, sat_nz;
  CellGeom cells;
  const long long* cell_start;     // [ncell + 1]
  const float4* sorted;            // map points grouped by cell
  long long n_map;
  double sr, res;                  // safe_radius, resolution_
  int safe_voxel;                  // floor(0.5 * safe_radius / res)
  int z_ground;
  double ground_limit;             // GroundZPos + safeHeight
  int mode;                        // 0 = viewpointSafetyCheck (lattice + nearCheck), 1 = nearCheck only
  const float* cand;               // n x (x, y, z) float, as pcl::PointXYZ holds them
  long long n;
  uint8_t* ok;                     // [n]
  unsigned long long* lookups;     // += safety_check probes made (reference order, early exit included)
};
cudaError_t launch_cloud_minmax(const float* xyz, long long n, float* partial, int blocks, cudaStream_t s);
cudaError_t launch_voxelise(const float* xyz, long long n, const MapGeom& g, unsigned long long* bricks, cudaStream_t s);
// voxels of the padded brick array beyond map_voxel_num_ read 'occupied' (occCheck is false outside the map)
cudaError_t launch_mark_padding(unsigned long long* bricks, const int dims[3], const int nb[3], cudaStream_t s);
// summed-area table over cells of FCVM_SAT_BRICKS^3 bricks; sat must hold (nb[0]/B + 1) * (nb[1]/B + 1) * (nb[2]/B + 1) ints
cudaError_t launch_build_sat(const unsigned long long* bricks, const int nb[3], int* sat, cudaStream_t s);
cudaError_t launch_cell_count(const float* xyz, long long n, const CellGeom& cg, int* count, cudaStream_t s);
cudaError_t launch_cell_fill(const float* xyz, long long n, const CellGeom& cg, const long long* start, int* fill, float4* sorted, cudaStream_t s);
cudaError_t launch_safety(const SafetyArgs& a, cudaStream_t s);

// Loads every kernel of the viewpoint path into the current device's context (cudaFuncGetAttributes), so that the first
// updateViewpoints does not pay the driver's lazy module loading launch by launch (~0.2 ms per kernel); called once per device.
cudaError_t preload_eval_kernels();
cudaError_t preload_pose_kernels();
cudaError_t preload_map_kernels();

}  // namespace fcvm
