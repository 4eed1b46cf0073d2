// fcvm_plan.cu — three ray / probe loops of the planner that run on the planner's own SDFMap, batched on sm_100a
// (outside the ViewpointManager boundary, and rank 4): kernels + C ABI (fcvm_grid_*).
//
//   hierarchical_coverage_planner/src/hcplanner.cpp:805-863   lightStateDet      -> k_light_state
//   hierarchical_coverage_planner/src/hcplanner.cpp:658-692   candidate safety box -> k_safety_box
//   hierarchical_coverage_planner/src/hcsolver.cpp:556-583    search_Path, straight-line part -> k_los_ends, k_los_classify, k_los_march
//
// The planner's SDFMap holds three char buffers (sdf_map.h:217-227).  They are uploaded once and packed into the
// bricked 1-bit layout of fcvm_device.cuh as the three planes these loops test:
//   occ  = occupancy_buffer_hc_ == 1                              (SDFMap::occCheck, sdf_map.h:397-420)
//   sblk = occupancy_buffer_hc_ == 1 || occupancy_buffer_internal_ == 1      (SDFMap::safety_check, :375-395)
//   pblk = occupancy_inflate_buffer_hc_ == 1 || occupancy_buffer_internal_ == 1   (hcsolver.cpp:566)
// Marchers: Marcher of fcvm_device.cuh (raycast.cpp restated, FP64, -fmad=false), literal and bounds-checked.
//
// The buffers themselves can also be BUILT on the device, from an empty grid:
//   plan_env/src/sdf_map.cpp:440-455   SDFMap::SetOcc (and the marking loop of initHCMap, :178-187)  -> k_grid_mark_points
//   plan_env/src/sdf_map.cpp:263-395   SDFMap::InternalSpace  -> k_grid_mark_points (hc + inflate dilation), k_internal_cut,
//                                      k_internal_remain (rays from the segment's projection points, marking `internal`)
//   plan_env/src/sdf_map.cpp:457-473   SDFMap::OuterCheck     -> k_outer_check (clears `internal` along outward rays)
// Every write of those loops is "set to 1" (or "clear if 1"), so all rays of a call are independent: one thread per ray,
// atomicOr / atomicAnd on the bricked bit planes.
// No CPU fallback.
#include <cfloat>
#include <cstring>
#include <limits>
#include <vector>

#include "fcvm_hostutil.h"
#include "fcvm_kernels.h"

namespace fcvm {

struct PlanPlanes {
  GridDesc occ, sblk, pblk;     // same geometry, three brick arrays
};

// one thread per brick: 64 voxels of each buffer (x-major bytes, sdf_map.h:266-269) -> one word per raw plane
__global__ void __launch_bounds__(256) k_pack_planes(const signed char* __restrict__ hc, const signed char* __restrict__ inflate,
                                                     const signed char* __restrict__ internal, int nx, int ny, int nz, int nby, int nbz,
                                                     long long nbricks, unsigned long long* __restrict__ p_hc,
                                                     unsigned long long* __restrict__ p_inf, unsigned long long* __restrict__ p_int) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbricks) return;
  const int bz = (int)(b % nbz), by = (int)((b / nbz) % nby), bx = (int)(b / ((long long)nbz * nby));   // b enumerates brick coordinates
  const long long dst = brick_word_index(bx * 4, by * 4, bz * 4, nby, nbz);                                   // ... stored in line order
  unsigned long long wo = 0ull, wf = 0ull, wt = 0ull;
  for (int dx = 0; dx < 4; ++dx)
    for (int dy = 0; dy < 4; ++dy)
      for (int dz = 0; dz < 4; ++dz) {
        const int x = bx * 4 + dx, y = by * 4 + dy, z = bz * 4 + dz;
        if (x >= nx || y >= ny || z >= nz) continue;
        const long long a = ((long long)x * ny + y) * nz + z;
        const unsigned long long bit = 1ull << brick_bit_index(x, y, z);
        if (hc && hc[a] == 1) wo |= bit;
        if (inflate && inflate[a] == 1) wf |= bit;
        if (internal && internal[a] == 1) wt |= bit;
      }
  p_hc[dst] = wo; p_inf[dst] = wf; p_int[dst] = wt;
}
// the planes the three loops test: sblk = hc | internal (safety_check), pblk = inflate | internal (search_Path); occ = hc itself
__global__ void __launch_bounds__(256) k_derive_planes(const unsigned long long* __restrict__ p_hc, const unsigned long long* __restrict__ p_inf,
                                                       const unsigned long long* __restrict__ p_int, long long nbricks,
                                                       unsigned long long* __restrict__ sblk, unsigned long long* __restrict__ pblk) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbricks) return;
  const unsigned long long t = p_int[b];
  sblk[b] = p_hc[b] | t;
  pblk[b] = p_inf[b] | t;
}
// raw planes -> the reference's x-major char buffers (parity tap)
__global__ void __launch_bounds__(256) k_unpack_plane(const unsigned long long* __restrict__ plane, int nx, int ny, int nz, int nby, int nbz,
                                                      signed char* __restrict__ out) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= (long long)nx * ny * nz) return;
  const int z = (int)(a % nz), y = (int)((a / nz) % ny), x = (int)(a / ((long long)nz * ny));
  out[a] = (signed char)((plane[brick_word_index(x, y, z, nby, nbz)] >> brick_bit_index(x, y, z)) & 1ull);
}

struct PlaneGeom {          // SDFMap geometry for posToIndex_hc / isInMap_hc (sdf_map.h:234-237, 367-373)
  double org[3], res_inv;
  int dims[3], nby, nbz;
};
__device__ __forceinline__ bool plane_in_map(const PlaneGeom& g, int x, int y, int z) {
  return (unsigned)x < (unsigned)g.dims[0] && (unsigned)y < (unsigned)g.dims[1] && (unsigned)z < (unsigned)g.dims[2];
}
__device__ __forceinline__ void plane_set(unsigned long long* plane, const PlaneGeom& g, int x, int y, int z) {
  atomicOr(plane + brick_word_index(x, y, z, g.nby, g.nbz), 1ull << brick_bit_index(x, y, z));
}

// SetOcc (sdf_map.cpp:440-455; dilate < 0) or the first loop of InternalSpace (:282-297; dilate = inflateVoxel >= 0): the voxel of
// every point is marked in hc (and inflate), and inflate additionally in the (2 dilate + 1)^3 cube around it.  Writes outside
// the map (undefined behaviour in the reference, which does not test) are skipped and counted in counters[2].
__global__ void __launch_bounds__(256) k_grid_mark_points(PlaneGeom g, const float* __restrict__ xyz, long long n, int dilate,
                                                          unsigned long long* __restrict__ p_hc, unsigned long long* __restrict__ p_inf,
                                                          unsigned long long* __restrict__ counters) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)floor(((double)xyz[3 * i] - g.org[0]) * g.res_inv);
  const int y = (int)floor(((double)xyz[3 * i + 1] - g.org[1]) * g.res_inv);
  const int z = (int)floor(((double)xyz[3 * i + 2] - g.org[2]) * g.res_inv);
  unsigned long long oob = 0;
  if (plane_in_map(g, x, y, z)) {
    plane_set(p_hc, g, x, y, z);
    if (dilate >= 0) plane_set(p_inf, g, x, y, z);
  } else if (dilate >= 0) {
    oob++;
  }
  for (int a = -dilate; a <= dilate; ++a)
    for (int b = -dilate; b <= dilate; ++b)
      for (int c = -dilate; c <= dilate; ++c) {
        if (a == 0 && b == 0 && c == 0) continue;
        if (plane_in_map(g, x + a, y + b, z + c)) plane_set(p_inf, g, x + a, y + b, z + c);
        else oob++;
      }
  if (oob) atomicAdd(counters + 2, oob);
}

// `internal_cast_->input(a, b); nextId(idx); while (nextId(idx)) buffer[idx] = ...` (sdf_map.cpp:332-337, 379-384, 465-471): the
// voxels strictly between the first and the last one of the one-directional march a -> b.  SET: mark; !SET: clear.
template <bool SET>
__device__ __forceinline__ void cast_mark(const GridDesc& g, const PlaneGeom& pg, double ax, double ay, double az, double bx, double by, double bz,
                                          unsigned long long* __restrict__ plane, unsigned long long& looks, unsigned long long& trips, unsigned long long& oob) {
  const double res = g.res;
  Marcher m;
  m.init(ax / res, ay / res, az / res, (int)floor(bx / res), (int)floor(by / res), (int)floor(bz / res), g);
  if (m.at_end()) return;
  int budget = m.manhattan() + 8;
  m.step(g);                                    // the discarded first nextId
  --budget;
  for (;;) {
    const int ix = m.ix, iy = m.iy, iz = m.iz;  // nextId reports the voxel before stepping and returns false at the end voxel
    if (m.at_end()) break;
    if (--budget < 0) { trips++; break; }
    m.step(g);
    looks++;
    if (plane_in_map(pg, ix, iy, iz)) {
      const unsigned long long bit = 1ull << brick_bit_index(ix, iy, iz);
      if (SET) atomicOr(plane + brick_word_index(ix, iy, iz, pg.nby, pg.nbz), bit);
      else atomicAnd(plane + brick_word_index(ix, iy, iz, pg.nby, pg.nbz), ~bit);
    } else {
      oob++;
    }
  }
}
__device__ __forceinline__ void plan_flush(unsigned long long* counters, unsigned long long looks, unsigned long long trips, unsigned long long oob) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    looks += __shfl_xor_sync(0xffffffffu, looks, o); trips += __shfl_xor_sync(0xffffffffu, trips, o); oob += __shfl_xor_sync(0xffffffffu, oob, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (looks) atomicAdd(counters + 0, looks);
    if (trips) atomicAdd(counters + 1, trips);
    if (oob) atomicAdd(counters + 2, oob);
  }
}

// InternalSpace, cutting-plane part (sdf_map.cpp:315-341) of ONE segment: thread = (cloud point i, projection point k); the
// point lies in k's cutting plane when |(pt - proj_k) . seg_dir| <= thickness (points_in_plane, :475-491); then the ray
// proj_k -> pt marks `internal` and the point counts as visited.
__global__ void __launch_bounds__(128) k_internal_cut(GridDesc g, PlaneGeom pg, const float* __restrict__ cloud, long long n, const double* __restrict__ proj,
                                                      int K, double dirx, double diry, double dirz, double thickness, unsigned long long* __restrict__ p_int,
                                                      unsigned char* __restrict__ visited, unsigned long long* __restrict__ counters) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long looks = 0, trips = 0, oob = 0;
  if (t < n * K) {
    const long long i = t / K;
    const int k = (int)(t % K);
    const double px = (double)cloud[3 * i], py = (double)cloud[3 * i + 1], pz = (double)cloud[3 * i + 2];
    const double qx = proj[3 * k], qy = proj[3 * k + 1], qz = proj[3 * k + 2];
    const double d = ((px - qx) * dirx + (py - qy) * diry) + (pz - qz) * dirz;        // Eigen 3-term order
    if (fabs(d) <= thickness) {
      visited[i] = 1;
      cast_mark<true>(g, pg, qx, qy, qz, px, py, pz, p_int, looks, trips, oob);
    }
  }
  plan_flush(counters, looks, trips, oob);
}
// InternalSpace, the points no plane caught (sdf_map.cpp:343-384): ray from the nearest projection point (first minimum, strict <,
// starting from 100000)
__global__ void __launch_bounds__(128) k_internal_remain(GridDesc g, PlaneGeom pg, const float* __restrict__ cloud, long long n, const double* __restrict__ proj,
                                                         int K, const unsigned char* __restrict__ visited, unsigned long long* __restrict__ p_int,
                                                         unsigned long long* __restrict__ counters) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long looks = 0, trips = 0, oob = 0;
  if (i < n && !visited[i]) {
    const double px = (double)cloud[3 * i], py = (double)cloud[3 * i + 1], pz = (double)cloud[3 * i + 2];
    double dist_min = 100000.0;
    int id = -1;
    for (int k = 0; k < K; ++k) {
      const double dx = proj[3 * k] - px, dy = proj[3 * k + 1] - py, dz = proj[3 * k + 2] - pz;
      const double dist = sqrt((dx * dx + dy * dy) + dz * dz);
      if (dist < dist_min) { dist_min = dist; id = k; }
    }
    if (id >= 0) cast_mark<true>(g, pg, proj[3 * id], proj[3 * id + 1], proj[3 * id + 2], px, py, pz, p_int, looks, trips, oob);
  }
  plan_flush(counters, looks, trips, oob);
}
// OuterCheck (sdf_map.cpp:457-473): start = o.head(3), end = start + checkScale * o.tail(3)
__global__ void __launch_bounds__(128) k_outer_check(GridDesc g, PlaneGeom pg, const double* __restrict__ outers, long long n, double check_scale,
                                                     unsigned long long* __restrict__ p_int, unsigned long long* __restrict__ counters) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long looks = 0, trips = 0, oob = 0;
  if (i < n) {
    const double sx = outers[6 * i], sy = outers[6 * i + 1], sz = outers[6 * i + 2];
    cast_mark<false>(g, pg, sx, sy, sz, sx + check_scale * outers[6 * i + 3], sy + check_scale * outers[6 * i + 4], sz + check_scale * outers[6 * i + 5], p_int,
                     looks, trips, oob);
  }
  plan_flush(counters, looks, trips, oob);
}

// lightStateDet (hcplanner.cpp:805-863), K = 1: nearest ROSA vertex with FLANN's L2_Simple<float> (lowest index on ties),
// BiRC candidate -> vertex with the first biNextId discarded, every voxel of both half-rays tested with occCheck (no
// early exit), `idxFwd` once more after the loop; det = the number of blocked voxels is odd.
__global__ void __launch_bounds__(128) k_light_state(GridDesc g, const double* __restrict__ cand, long long n, const float* __restrict__ rosa,
                                                     long long m, uint8_t* __restrict__ det, int* __restrict__ cross, int* __restrict__ nearest,
                                                     unsigned long long* __restrict__ counters) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ax = cand[3 * i], ay = cand[3 * i + 1], az = cand[3 * i + 2];
  const float qx = (float)ax, qy = (float)ay, qz = (float)az;      // pcl::PointXYZ searchVP
  int best = -1;
  float bd = FLT_MAX;
  for (long long v = 0; v < m; ++v) {
    float result = 0.0f, diff;
    diff = qx - __ldg(rosa + 3 * v); result += diff * diff;
    diff = qy - __ldg(rosa + 3 * v + 1); result += diff * diff;
    diff = qz - __ldg(rosa + 3 * v + 2); result += diff * diff;
    if (result < bd) { bd = result; best = (int)v; }
  }
  int cc = 0;
  unsigned long long looks = 0, trips = 0;
  if (best >= 0) {
    const double bx = (double)__ldg(rosa + 3 * best), by = (double)__ldg(rosa + 3 * best + 1), bz = (double)__ldg(rosa + 3 * best + 2);
    const double res = g.res;
    const int mx = (int)floor((0.5 * (ax + bx)) / res), my = (int)floor((0.5 * (ay + by)) / res), mz = (int)floor((0.5 * (az + bz)) / res);
    Marcher p, q;
    p.init(ax / res, ay / res, az / res, mx, my, mz, g);
    q.init(bx / res, by / res, bz / res, mx, my, mz, g);
    const int kp = p.manhattan(), kn = q.manhattan();
    int budget = (kp > kn ? kp : kn) + 8;
    BrickCache cp, cn;
    cp.reset(); cn.reset();
    int pix, piy, piz, nix, niy, niz;
    bool more;
#define FCVM_BI_NEXT(ret)                        \
  {                                              \
    pix = p.ix; piy = p.iy; piz = p.iz;          \
    const bool pflag = !p.at_end();              \
    if (pflag) p.step(g);                        \
    nix = q.ix; niy = q.iy; niz = q.iz;          \
    const bool nflag = !q.at_end();              \
    if (nflag) q.step(g);                        \
    ret = pflag || nflag;                        \
  }
    FCVM_BI_NEXT(more);
    if (more) --budget;            // the discarded call counts against the step cap like any other (oracle: calls_)
    for (;;) {
      FCVM_BI_NEXT(more);
      if (!more) break;
      if (--budget < 0) { trips++; break; }
      if (!occ_free(g, pix, piy, piz, cp)) cc++;
      if (!occ_free(g, nix, niy, niz, cn)) cc++;
      looks += 2;
    }
#undef FCVM_BI_NEXT
    if (!occ_free(g, pix, piy, piz, cp)) cc++;
    looks += 1;
  }
  det[i] = (best >= 0 && (cc % 2) != 0) ? 1 : 0;
  if (cross) cross[i] = cc;
  if (nearest) nearest[i] = best;
  if (counters) { atomicAdd(counters + 0, looks); if (trips) atomicAdd(counters + 1, trips); }
}

// the candidate safety box (hcplanner.cpp:658-676).  `break` there leaves only the k loop and safe_flag is overwritten by
// every probe, so the verdict is that of the last (i, j) column: all of its probes (stride 2 in k) free.
__device__ __forceinline__ int box_steps(double v, double sr, double res) {
  int cnt = 0;
  for (int i = 0; v - sr + i * res < v + sr && cnt < 65536; i += 2) ++cnt;
  return cnt;
}
__global__ void __launch_bounds__(128) k_safety_box(GridDesc g, double res_inv, const float* __restrict__ vp, long long n, double sr,
                                                    uint8_t* __restrict__ ok, unsigned long long* __restrict__ counters) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const double x = (double)vp[3 * v], y = (double)vp[3 * v + 1], z = (double)vp[3 * v + 2];
  const int ni = box_steps(x, sr, g.res), nj = box_steps(y, sr, g.res);
  bool safe = true;
  unsigned long long looks = 0;
  if (ni > 0 && nj > 0) {
    const double px = x - sr + (2 * (ni - 1)) * g.res, py = y - sr + (2 * (nj - 1)) * g.res;
    const int ix = (int)floor((px - g.orgx) * res_inv), iy = (int)floor((py - g.orgy) * res_inv);
    BrickCache c;
    c.reset();
    for (int k = 0, cnt = 0; z - sr + k * g.res < z + sr && cnt < 65536; k += 2, ++cnt) {
      const int iz = (int)floor((z - sr + k * g.res - g.orgz) * res_inv);
      looks++;
      if (!occ_free(g, ix, iy, iz, c)) { safe = false; break; }
    }
  }
  ok[v] = safe ? 1 : 0;
  if (counters && looks) atomicAdd(counters + 0, looks);
}

// search_Path's straight-line test (hcsolver.cpp:556-583) for every ordered pair (i, j) of n positions:
//   raycaster_->input(p1, p2); while (raycaster_->nextId(idx)) if (inflate[idx] == 1 || internal[idx] == 1) -> not safe
// Thread = pair left 11 of 32 lanes busy (ncu r1n: rays of very different lengths share a warp), binning the pairs by length
// (round 2, first try) 15 of 32 at the price of a 0.9 ms scatter: a blocked ray ends anywhere.  Now three steps:
//   k_los_ends      per POSITION: p / resolution (the reference's own division), its voxel, its map index, whether it lies
//                   strictly inside its voxel - 6 FP64 divisions per pair become 3 per position
//   k_los_classify  per PAIR, full warps: rays on the fast path (below) whose whole box of voxels is free by the summed-area
//                   table of the pblk plane are safe without a march (8 loads); the other fast rays are appended to a list;
//                   the few rays off the fast path run the literal marcher on the spot
//   k_los_march     persistent warps over the list: a lane that finishes takes the next listed ray (refill in groups, so the
//                   set-up runs several lanes wide); the DDA set-up divides through a reciprocal table (div_exact), the step is
//                   predicated, and every 4 steps a ray asks the table whether the box from its front to its end is free.
// Fast path = both end voxels inside the map, the map-index conversion a pure shift between them, every |delta| below the
// reciprocal table's size, and the start strictly inside its voxel (by 1e-9) on every moving axis.  Such a marcher makes exactly
// |dx| + |dy| + |dz| steps and never leaves the box spanned by its end voxels (the direction is the INTEGER voxel delta,
// raycast.cpp:412-434, so the crossings it needs all have t < 1 and all others t > 1, by margins no rounding bridges; the
// reference's runaway marchers start ON a lattice plane).  So "the box is free" means every voxel nextId reports is free,
// whatever their exact sequence: the verdict and the look-up count (one per step) are the reference's.
constexpr int kLosRcpN = 1024;
struct LosEnd {
  double fx, fy, fz;     // position / resolution
  int vx, vy, vz;        // floor of it (world-anchored voxel)
  int ix, iy, iz;        // the voxel as a map index, `(int)(v + offset_)`
  int flags;             // bits 0-2: strictly inside the voxel on x / y / z; bit 3: finite and the voxel is inside the map
  int pad[3];
};
static_assert(sizeof(LosEnd) == 64, "LosEnd layout");

__global__ void __launch_bounds__(256) k_los_ends(GridDesc g, const double* __restrict__ pts, long long n, LosEnd* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  LosEnd e;
  e.fx = pts[3 * i] / g.res; e.fy = pts[3 * i + 1] / g.res; e.fz = pts[3 * i + 2] / g.res;
  const bool finite = (fabs(e.fx) + fabs(e.fy)) + fabs(e.fz) < 1.0e9;         // false for NaN / inf
  const double flx = floor(e.fx), fly = floor(e.fy), flz = floor(e.fz);
  e.vx = finite ? (int)flx : 0; e.vy = finite ? (int)fly : 0; e.vz = finite ? (int)flz : 0;
  e.ix = to_map_index(e.vx, g.offx); e.iy = to_map_index(e.vy, g.offy); e.iz = to_map_index(e.vz, g.offz);
  const double rx = e.fx - flx, ry = e.fy - fly, rz = e.fz - flz;              // exact
  e.flags = ((rx > 1e-9 && rx < 1.0 - 1e-9) ? 1 : 0) | ((ry > 1e-9 && ry < 1.0 - 1e-9) ? 2 : 0) | ((rz > 1e-9 && rz < 1.0 - 1e-9) ? 4 : 0);
  if (finite && (unsigned)e.ix < (unsigned)g.nx && (unsigned)e.iy < (unsigned)g.ny && (unsigned)e.iz < (unsigned)g.nz) e.flags |= 8;
  for (int k = 0; k < 3; ++k) e.pad[k] = 0;
  out[i] = e;
}

// the literal loop (every pair took it before round 2): Marcher of fcvm_device.cuh, bounds-checked look-ups, step cap
__device__ __noinline__ bool los_literal(const GridDesc& g, double ax, double ay, double az, double bx, double by, double bz,
                                         unsigned long long& looks, unsigned long long& trips) {
  const double res = g.res;
  Marcher m;
  m.init(ax / res, ay / res, az / res, (int)floor(bx / res), (int)floor(by / res), (int)floor(bz / res), g);
  int budget = m.manhattan() + 8;
  BrickCache c;
  c.reset();
  for (;;) {
    const int ix = m.ix, iy = m.iy, iz = m.iz;     // nextId reports the voxel before stepping ...
    if (m.at_end()) return true;                    // ... and returns false at the end voxel (which is never tested)
    if (--budget < 0) { trips++; return false; }
    m.step(g);
    looks++;
    if (!occ_free(g, ix, iy, iz, c)) return false; // blocked, or outside the map
  }
}

__device__ __forceinline__ bool los_fast(const LosEnd& a, const LosEnd& b, int& dx, int& dy, int& dz) {
  dx = b.vx - a.vx; dy = b.vy - a.vy; dz = b.vz - a.vz;
  const int inside = a.flags | (dx == 0 ? 1 : 0) | (dy == 0 ? 2 : 0) | (dz == 0 ? 4 : 0);     // idle axes do not matter
  return (a.flags & b.flags & 8) && (inside & 7) == 7 && (b.ix - a.ix == dx) && (b.iy - a.iy == dy) && (b.iz - a.iz == dz) &&
         abs(dx) < kLosRcpN && abs(dy) < kLosRcpN && abs(dz) < kLosRcpN;
}

__global__ void __launch_bounds__(256) k_los_classify(GridDesc g, const double* __restrict__ pts, const LosEnd* __restrict__ ends, long long n,
                                                      uint8_t* __restrict__ safe, double* __restrict__ dist, unsigned* __restrict__ list,
                                                      int* __restrict__ cursor, unsigned long long* __restrict__ counters) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = t < n * n;
  unsigned long long looks = 0, trips = 0;
  bool listed = false;
  unsigned packed = 0u;
  if (valid) {
    const unsigned i = (unsigned)(t / n), j = (unsigned)(t % n);
    const LosEnd a = ends[i], b = ends[j];
    int dx, dy, dz;
    if (los_fast(a, b, dx, dy, dz)) {
      const int steps = abs(dx) + abs(dy) + abs(dz);
      if (steps == 0 ||
          sat_box_count(g, min(a.ix, b.ix) >> FCVM_SAT_SHIFT, min(a.iy, b.iy) >> FCVM_SAT_SHIFT, min(a.iz, b.iz) >> FCVM_SAT_SHIFT,
                        max(a.ix, b.ix) >> FCVM_SAT_SHIFT, max(a.iy, b.iy) >> FCVM_SAT_SHIFT, max(a.iz, b.iz) >> FCVM_SAT_SHIFT) == 0) {
        safe[t] = 1;
        looks = (unsigned long long)steps;
      } else {
        listed = true;
        packed = (i << 16) | j;                       // n <= 46340 < 2^16
      }
    } else {
      safe[t] = los_literal(g, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], pts[3 * j], pts[3 * j + 1], pts[3 * j + 2], looks, trips) ? 1 : 0;
    }
    if (dist) {
      const double ex = pts[3 * i] - pts[3 * j], ey = pts[3 * i + 1] - pts[3 * j + 1], ez = pts[3 * i + 2] - pts[3 * j + 2];
      dist[t] = sqrt((ex * ex + ey * ey) + ez * ez);  // (p1 - p2).norm(), Eigen order
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, listed);
  const int lane = threadIdx.x & 31;
  if (m != 0u) {
    int base = 0;
    if (lane == 0) base = atomicAdd(cursor, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (listed) list[base + __popc(m & ((1u << lane) - 1u))] = packed;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { looks += __shfl_xor_sync(0xffffffffu, looks, o); trips += __shfl_xor_sync(0xffffffffu, trips, o); }
  if (lane == 0) { if (looks) atomicAdd(counters + 0, looks); if (trips) atomicAdd(counters + 1, trips); }
}

#ifndef LOS_REFILL
#define LOS_REFILL 8          // k_los_march: idle lanes of a warp take new rays once at least this many are idle
#endif
#ifndef LOS_QUERY_PERIOD
#define LOS_QUERY_PERIOD 8    // k_los_march: steps between two free-box queries of a ray
#endif
__global__ void __launch_bounds__(256) k_los_march(GridDesc g, const LosEnd* __restrict__ ends, long long n, const double* __restrict__ rcp,
                                                   const unsigned* __restrict__ list, int* __restrict__ cursor, uint8_t* __restrict__ safe,
                                                   unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int count = cursor[0];                       // rays listed by k_los_classify; cursor[1] = next one to hand out
  unsigned long long looks = 0;
  bool active = false, more = true;
  int ix = 0, iy = 0, iz = 0, ex = 0, ey = 0, ez = 0, sx = 0, sy = 0, sz = 0, rem = 0, key = -1;
  double tmx = 0, tmy = 0, tmz = 0, tdx = 0, tdy = 0, tdz = 0;
  unsigned long long word = 0ull;
  long long t = 0;
  for (;;) {
    const unsigned idle = __ballot_sync(0xffffffffu, !active);
    if (more && (__popc(idle) >= LOS_REFILL)) {
      int base = 0;
      if (lane == 0) base = atomicAdd(cursor + 1, __popc(idle));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + __popc(idle) >= count) more = false;
      const int mine = base + __popc(idle & lt_mask);
      if (!active && mine < count) {
        const unsigned pk = __ldg(list + mine);
        const unsigned i = pk >> 16, j = pk & 0xffffu;
        const LosEnd* a = ends + i;
        const LosEnd* b = ends + j;
        const int avx = __ldg(&a->vx), avy = __ldg(&a->vy), avz = __ldg(&a->vz);
        const int dx = __ldg(&b->vx) - avx, dy = __ldg(&b->vy) - avy, dz = __ldg(&b->vz) - avz;
        ix = __ldg(&a->ix); iy = __ldg(&a->iy); iz = __ldg(&a->iz);
        ex = __ldg(&b->ix); ey = __ldg(&b->iy); ez = __ldg(&b->iz);
        sx = d_signum(dx); sy = d_signum(dy); sz = d_signum(dz);
        rem = abs(dx) + abs(dy) + abs(dz);
        tmx = d_intbound_i(__ldg(&a->fx), dx, rcp); tmy = d_intbound_i(__ldg(&a->fy), dy, rcp); tmz = d_intbound_i(__ldg(&a->fz), dz, rcp);
        tdx = __ldg(rcp + abs(dx)); tdy = __ldg(rcp + abs(dy)); tdz = __ldg(rcp + abs(dz));     // rcp[0] = NaN = 0 / 0 on idle axes
        t = (long long)i * n + j;
        key = -1;
        active = true;
      }
    }
    if (__ballot_sync(0xffffffffu, active) == 0u) {
      if (!more) break;
      continue;
    }
    if (active) {
      bool blocked = false;
#pragma unroll
      for (int u = 0; u < LOS_QUERY_PERIOD; ++u) {
        if (rem > 0 && !blocked) {
          // nextId reports the voxel BEFORE it steps; the reference tests that voxel
          const int k = ((ix >> 2) * g.nby + (iy >> 2)) * g.nbz + (iz >> 2);
          if (k != key) { key = k; word = __ldg(g.bricks + k); }
          blocked = ((word >> brick_bit_index(ix, iy, iz)) & 1ull) != 0ull;
          looks++;
          // the branch ladder of nextId (raycast.cpp:455-476): strict '<', z on ties and NaN
          if (tmx < tmy) {
            if (tmx < tmz) { ix += sx; tmx += tdx; } else { iz += sz; tmz += tdz; }
          } else {
            if (tmy < tmz) { iy += sy; tmy += tdy; } else { iz += sz; tmz += tdz; }
          }
          rem--;
        }
      }
      if (blocked) {
        safe[t] = 0;
        active = false;
      } else if (rem == 0) {
        safe[t] = 1;
        active = false;
      } else if (sat_box_count(g, min(ix, ex) >> FCVM_SAT_SHIFT, min(iy, ey) >> FCVM_SAT_SHIFT, min(iz, ez) >> FCVM_SAT_SHIFT,
                               max(ix, ex) >> FCVM_SAT_SHIFT, max(iy, ey) >> FCVM_SAT_SHIFT, max(iz, ez) >> FCVM_SAT_SHIFT) == 0) {
        looks += (unsigned long long)rem;             // every voxel still to be reported is free: one look-up per remaining step
        safe[t] = 1;
        active = false;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) looks += __shfl_xor_sync(0xffffffffu, looks, o);
  if (lane == 0 && looks) atomicAdd(counters + 0, looks);
}

}  // namespace fcvm

using namespace fcvm;

struct fcvm_grid_handle {
  int device = 0;
  double origin[3], res = 0;
  int dims[3], nb[3];
  long long nbricks = 0;
  DevBuf<unsigned long long> d_occ, d_inf, d_int;     // raw planes: occupancy_buffer_hc_ (= occ), occupancy_inflate_buffer_hc_, occupancy_buffer_internal_
  DevBuf<unsigned long long> d_sblk, d_pblk;          // derived planes, refreshed after every change of the raw ones
  DevBuf<unsigned char> d_visited;
  DevBuf<int> d_sat;                  // summed-area table of non-empty bricks of the pblk plane (line-of-sight certificate)
  bool sat_ok = false;                //   ... rebuilt lazily after the planes changed
  DevBuf<double> d_rcp;               // 1 / k, k < kLosRcpN (the divisions of the DDA set-up)
  DevBuf<LosEnd> d_ends;
  DevBuf<unsigned> d_list;            // rays k_los_classify left for k_los_march, (i << 16) | j
  DevBuf<int> d_cursor;               // [0] rays listed, [1] next one handed out
  DevBuf<double> d_in;
  DevBuf<float> d_in_f, d_rosa;
  DevBuf<uint8_t> d_u8;
  DevBuf<int> d_i32a, d_i32b;
  DevBuf<double> d_dist;
  DevBuf<unsigned long long> d_counters;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_kernel_ms = 0;
  int64_t counters[3] = {0, 0, 0};    // lookups, cap trips, kernels launched
  int64_t oob_writes = 0;             // voxel writes of SetOcc / InternalSpace / OuterCheck that fell outside the map (skipped)
};

static void plan_desc(const fcvm_grid_handle* h, const unsigned long long* bricks, GridDesc& g) {
  g.bricks = bricks;
  g.nx = h->dims[0]; g.ny = h->dims[1]; g.nz = h->dims[2];
  g.nbx = h->nb[0]; g.nby = h->nb[1]; g.nbz = h->nb[2];
  g.nly = h->nb[1] >> 1; g.nlz = h->nb[2] >> 2;
  g.nbricks = (int)h->nbricks;
  g.guard_steps = 0;
  g.sat = nullptr; g.sat_ny = g.sat_nz = 0;
  g.res = h->res;
  g.res_rcp = 1.0 / h->res;
  g.offx = 0.5 - h->origin[0] / h->res;      // RayCaster::setParams (raycast.cpp:341-345; hcplanner.cpp:89, hcsolver.cpp:40)
  g.offy = 0.5 - h->origin[1] / h->res;
  g.offz = 0.5 - h->origin[2] / h->res;
  g.orgx = h->origin[0]; g.orgy = h->origin[1]; g.orgz = h->origin[2];
}

static void plane_geom(const fcvm_grid_handle* h, PlaneGeom& g) {
  for (int c = 0; c < 3; ++c) { g.org[c] = h->origin[c]; g.dims[c] = h->dims[c]; }
  g.res_inv = 1 / h->res;
  g.nby = h->nb[1]; g.nbz = h->nb[2];
}
static int plan_derive(fcvm_grid_handle* h) {
  k_derive_planes<<<(unsigned)((h->nbricks + 255) / 256), 256, 0, h->stream>>>(h->d_occ.p, h->d_inf.p, h->d_int.p, h->nbricks, h->d_sblk.p, h->d_pblk.p);
  CU(cudaGetLastError());
  h->sat_ok = false;
  return FCVM_OK;
}

static int plan_finish(fcvm_grid_handle* h) {
  unsigned long long c[3];
  CU(cudaMemcpyAsync(c, h->d_counters.p, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemsetAsync(h->d_counters.p, 0, sizeof(c), h->stream));
  CU(cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_kernel_ms = ms;
  h->counters[0] += (int64_t)c[0];
  h->counters[1] += (int64_t)c[1];
  h->oob_writes += (int64_t)c[2];
  h->counters[2] += 1;
  return FCVM_OK;
}

extern "C" {

fcvm_grid_t* fcvm_grid_create(int32_t device, const double* origin3, double resolution, const int32_t* voxel_num3, const int8_t* occupancy_hc,
                              const int8_t* occupancy_inflate, const int8_t* occupancy_internal) {
  if (!origin3 || !voxel_num3 || !(resolution > 0)) { fail(FCVM_EINVAL, "bad grid geometry"); return nullptr; }
  const long long G = (long long)voxel_num3[0] * voxel_num3[1] * voxel_num3[2];
  if (voxel_num3[0] <= 0 || voxel_num3[1] <= 0 || voxel_num3[2] <= 0 || G >= (1ll << 31)) { fail(FCVM_EINVAL, "bad voxel counts"); return nullptr; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fail(FCVM_ENODEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + " (this library has no CPU fallback)");
    return nullptr;
  }
  fcvm_grid_handle* h = new (std::nothrow) fcvm_grid_handle();
  if (!h) { fail(FCVM_EINTERNAL, "out of memory"); return nullptr; }
  h->device = device;
  h->res = resolution;
  for (int c = 0; c < 3; ++c) { h->origin[c] = origin3[c]; h->dims[c] = voxel_num3[c]; }
  pad_bricks(h->dims, h->nb);
  h->nbricks = (long long)h->nb[0] * h->nb[1] * h->nb[2];
  auto build = [&]() -> int {
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(h->d_counters.reserve(3));
    CU(cudaMemset(h->d_counters.p, 0, 3 * sizeof(unsigned long long)));
    CU(h->d_occ.reserve((size_t)h->nbricks));
    CU(h->d_inf.reserve((size_t)h->nbricks));
    CU(h->d_int.reserve((size_t)h->nbricks));
    CU(h->d_sblk.reserve((size_t)h->nbricks));
    CU(h->d_pblk.reserve((size_t)h->nbricks));
    DevBuf<signed char> b[3];
    const int8_t* src[3] = {occupancy_hc, occupancy_inflate, occupancy_internal};
    for (int k = 0; k < 3; ++k)
      if (src[k]) {
        CU(b[k].reserve((size_t)G));
        CU(cudaMemcpyAsync(b[k].p, src[k], (size_t)G, cudaMemcpyHostToDevice, h->stream));
      }
    k_pack_planes<<<(unsigned)((h->nbricks + 255) / 256), 256, 0, h->stream>>>(b[0].p, b[1].p, b[2].p, h->dims[0], h->dims[1], h->dims[2], h->nb[1],
                                                                                  h->nb[2], h->nbricks, h->d_occ.p, h->d_inf.p, h->d_int.p);
    CU(cudaGetLastError());
    if (plan_derive(h) != FCVM_OK) return FCVM_ECUDA;
    CU(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 3; ++k) b[k].free();
    h->counters[2] += 1;
    return FCVM_OK;
  };
  if (build() != FCVM_OK) { fcvm_grid_destroy(h); return nullptr; }
  return h;
}

void fcvm_grid_destroy(fcvm_grid_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->d_occ.free(); h->d_inf.free(); h->d_int.free(); h->d_sblk.free(); h->d_pblk.free(); h->d_visited.free(); h->d_in.free(); h->d_in_f.free(); h->d_rosa.free(); h->d_u8.free();
  h->d_i32a.free(); h->d_i32b.free(); h->d_dist.free(); h->d_counters.free(); h->d_sat.free(); h->d_rcp.free(); h->d_ends.free(); h->d_list.free(); h->d_cursor.free();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int fcvm_grid_light_state(fcvm_grid_t* h, const double* cand_xyz, int64_t n, const float* rosa_xyz, int64_t m, uint8_t* det, int32_t* cross,
                          int32_t* nearest) {
  if (!h || !cand_xyz || n <= 0 || !rosa_xyz || m < 0 || !det) return fail(FCVM_EINVAL, "bad arguments");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->d_in.reserve((size_t)3 * n));
  CU(h->d_rosa.reserve((size_t)3 * std::max<int64_t>(m, 1)));
  CU(h->d_u8.reserve((size_t)n));
  CU(h->d_i32a.reserve((size_t)n));
  CU(h->d_i32b.reserve((size_t)n));
  CU(cudaMemcpyAsync(h->d_in.p, cand_xyz, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
  if (m > 0) CU(cudaMemcpyAsync(h->d_rosa.p, rosa_xyz, (size_t)3 * m * sizeof(float), cudaMemcpyHostToDevice, s));
  GridDesc g;
  plan_desc(h, h->d_occ.p, g);
  CU(cudaEventRecord(h->ev0, s));
  k_light_state<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, h->d_in.p, n, h->d_rosa.p, m, h->d_u8.p, h->d_i32a.p, h->d_i32b.p, h->d_counters.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(h->ev1, s));      // ev0 .. ev1 bracket the kernel alone
  CU(cudaMemcpyAsync(det, h->d_u8.p, (size_t)n, cudaMemcpyDeviceToHost, s));
  if (cross) CU(cudaMemcpyAsync(cross, h->d_i32a.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  if (nearest) CU(cudaMemcpyAsync(nearest, h->d_i32b.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
  return plan_finish(h);
}

int fcvm_grid_safety_box(fcvm_grid_t* h, const float* vp_xyz, int64_t n, double safe_radius, uint8_t* ok) {
  if (!h || !vp_xyz || n <= 0 || !ok) return fail(FCVM_EINVAL, "bad arguments");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->d_in_f.reserve((size_t)3 * n));
  CU(h->d_u8.reserve((size_t)n));
  CU(cudaMemcpyAsync(h->d_in_f.p, vp_xyz, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  GridDesc g;
  plan_desc(h, h->d_sblk.p, g);
  CU(cudaEventRecord(h->ev0, s));
  k_safety_box<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, 1 / h->res, h->d_in_f.p, n, safe_radius, h->d_u8.p, h->d_counters.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(h->ev1, s));      // ev0 .. ev1 bracket the kernel alone
  CU(cudaMemcpyAsync(ok, h->d_u8.p, (size_t)n, cudaMemcpyDeviceToHost, s));
  return plan_finish(h);
}

int fcvm_grid_line_of_sight(fcvm_grid_t* h, const double* p_xyz, int64_t n, uint8_t* safe_nxn, double* dist_nxn) {
  if (!h || !p_xyz || n <= 0 || !safe_nxn) return fail(FCVM_EINVAL, "bad arguments");
  if (n > 46340) return fail(FCVM_EINVAL, "too many positions for one pair matrix");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->d_in.reserve((size_t)3 * n));
  CU(h->d_u8.reserve((size_t)n * n));
  if (dist_nxn) CU(h->d_dist.reserve((size_t)n * n));
  CU(cudaMemcpyAsync(h->d_in.p, p_xyz, (size_t)3 * n * sizeof(double), cudaMemcpyHostToDevice, s));
  GridDesc g;
  plan_desc(h, h->d_pblk.p, g);
  if (!h->d_rcp.p) {                   // 1 / k by the divider the reference uses (IEEE, same bits on host and device); 0 / 0 for k = 0
    std::vector<double> rcp(kLosRcpN);
    for (int k = 0; k < kLosRcpN; ++k) rcp[(size_t)k] = k == 0 ? std::numeric_limits<double>::quiet_NaN() : 1.0 / (double)k;
    CU(h->d_rcp.reserve(kLosRcpN));
    CU(cudaMemcpyAsync(h->d_rcp.p, rcp.data(), kLosRcpN * sizeof(double), cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));      // `rcp` goes out of scope
  }
  const int sat_dims[3] = {h->nb[0] / FCVM_SAT_BRICKS + 1, h->nb[1] / FCVM_SAT_BRICKS + 1, h->nb[2] / FCVM_SAT_BRICKS + 1};
  if ((long long)sat_dims[0] * sat_dims[1] * sat_dims[2] >= (1ll << 31)) return fail(FCVM_EINVAL, "grid too large for the free-box table");
  if (!h->sat_ok) {                    // the planes changed since the table was built
    CU(h->d_sat.reserve((size_t)sat_dims[0] * sat_dims[1] * sat_dims[2]));
    CU(launch_build_sat(h->d_pblk.p, h->nb, h->d_sat.p, s));
    h->sat_ok = true;
    h->counters[2] += 4;
  }
  g.sat = h->d_sat.p; g.sat_ny = sat_dims[1]; g.sat_nz = sat_dims[2];
  CU(h->d_ends.reserve((size_t)n));
  CU(h->d_list.reserve((size_t)n * n));
  CU(h->d_cursor.reserve(2));
  CU(cudaEventRecord(h->ev0, s));
  CU(cudaMemsetAsync(h->d_cursor.p, 0, 2 * sizeof(int), s));
  k_los_ends<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(g, h->d_in.p, n, h->d_ends.p);
  CU(cudaGetLastError());
  k_los_classify<<<(unsigned)((n * n + 255) / 256), 256, 0, s>>>(g, h->d_in.p, h->d_ends.p, n, h->d_u8.p, dist_nxn ? h->d_dist.p : nullptr, h->d_list.p,
                                                               h->d_cursor.p, h->d_counters.p);
  CU(cudaGetLastError());
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const long long want = (n * n + 255) / 256;                       // never more warps than rays
    k_los_march<<<(unsigned)std::max<long long>(1, std::min<long long>((long long)sms * 4, want)), 256, 0, s>>>(g, h->d_ends.p, n, h->d_rcp.p, h->d_list.p,
                                                                                                             h->d_cursor.p, h->d_u8.p, h->d_counters.p);
    CU(cudaGetLastError());
  }
  h->counters[2] += 3;
  CU(cudaEventRecord(h->ev1, s));      // ev0 .. ev1 bracket the kernel alone
  CU(cudaMemcpyAsync(safe_nxn, h->d_u8.p, (size_t)n * n, cudaMemcpyDeviceToHost, s));
  if (dist_nxn) CU(cudaMemcpyAsync(dist_nxn, h->d_dist.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, s));
  return plan_finish(h);
}

int fcvm_grid_set_occ(fcvm_grid_t* h, const float* xyz, int64_t n) {
  if (!h || !xyz || n <= 0) return fail(FCVM_EINVAL, "bad arguments");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->d_in_f.reserve((size_t)3 * n));
  CU(cudaMemcpyAsync(h->d_in_f.p, xyz, (size_t)3 * n * sizeof(float), cudaMemcpyHostToDevice, s));
  PlaneGeom pg;
  plane_geom(h, pg);
  CU(cudaEventRecord(h->ev0, s));
  k_grid_mark_points<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(pg, h->d_in_f.p, n, -1, h->d_occ.p, h->d_inf.p, h->d_counters.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(h->ev1, s));
  int rc = plan_derive(h);
  if (rc) return rc;
  return plan_finish(h);
}

int fcvm_grid_internal_space(fcvm_grid_t* h, const float* seg_xyz, const int64_t* seg_offsets, int64_t nseg, const double* seg_ends6, int32_t inflate_num,
                             double proj_interval, double thickness) {
  if (!h || !seg_xyz || !seg_offsets || nseg <= 0 || !seg_ends6 || inflate_num < 0 || !(proj_interval > 0)) return fail(FCVM_EINVAL, "bad arguments");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  const int64_t n_all = seg_offsets[nseg];
  if (n_all <= 0) return fail(FCVM_EINVAL, "empty segment clouds");
  CU(h->d_in_f.reserve((size_t)3 * n_all));
  CU(h->d_visited.reserve((size_t)n_all));
  CU(cudaMemcpyAsync(h->d_in_f.p, seg_xyz, (size_t)3 * n_all * sizeof(float), cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(h->d_visited.p, 0, (size_t)n_all, s));
  PlaneGeom pg;
  plane_geom(h, pg);
  GridDesc g;
  plan_desc(h, h->d_int.p, g);
  CU(cudaEventRecord(h->ev0, s));
  // first loop of every segment (:282-297): one launch over all points
  k_grid_mark_points<<<(unsigned)((n_all + 255) / 256), 256, 0, s>>>(pg, h->d_in_f.p, n_all, inflate_num, h->d_occ.p, h->d_inf.p, h->d_counters.p);
  CU(cudaGetLastError());
  // projection points of every segment (:318-324), on the host in the reference's arithmetic: seg_start + (interval * i) * seg_dir
  std::vector<double> proj;
  std::vector<int64_t> proj_off((size_t)nseg + 1, 0);
  std::vector<double> dirs((size_t)3 * nseg);
  for (int64_t sg = 0; sg < nseg; ++sg) {
    const double* e = seg_ends6 + 6 * sg;
    const double dx = e[3] - e[0], dy = e[4] - e[1], dz = e[5] - e[2];
    const double z = (dx * dx + dy * dy) + dz * dz;
    const double len = std::sqrt(z);
    double ux = dx, uy = dy, uz = dz;
    if (z > 0.0) { ux = dx / len; uy = dy / len; uz = dz / len; }           // normalized()
    dirs[3 * sg] = ux; dirs[3 * sg + 1] = uy; dirs[3 * sg + 2] = uz;
    for (int i = 0; proj_interval * i < len; ++i) {
      const double t = proj_interval * i;
      proj.push_back(e[0] + t * ux); proj.push_back(e[1] + t * uy); proj.push_back(e[2] + t * uz);
    }
    proj.push_back(e[3]); proj.push_back(e[4]); proj.push_back(e[5]);
    proj_off[(size_t)sg + 1] = (int64_t)proj.size() / 3;
  }
  CU(h->d_in.reserve(proj.size()));
  CU(cudaMemcpyAsync(h->d_in.p, proj.data(), proj.size() * sizeof(double), cudaMemcpyHostToDevice, s));
  for (int64_t sg = 0; sg < nseg; ++sg) {
    const int64_t p0 = seg_offsets[sg], np = seg_offsets[sg + 1] - p0;
    const int K = (int)(proj_off[(size_t)sg + 1] - proj_off[(size_t)sg]);
    if (np <= 0 || K <= 0) continue;
    const double* dproj = h->d_in.p + 3 * proj_off[(size_t)sg];
    k_internal_cut<<<(unsigned)((np * K + 127) / 128), 128, 0, s>>>(g, pg, h->d_in_f.p + 3 * p0, np, dproj, K, dirs[3 * sg], dirs[3 * sg + 1], dirs[3 * sg + 2],
                                                                      thickness, h->d_int.p, h->d_visited.p + p0, h->d_counters.p);
    CU(cudaGetLastError());
    k_internal_remain<<<(unsigned)((np + 127) / 128), 128, 0, s>>>(g, pg, h->d_in_f.p + 3 * p0, np, dproj, K, h->d_visited.p + p0, h->d_int.p, h->d_counters.p);
    CU(cudaGetLastError());
    h->counters[2] += 2;
  }
  CU(cudaEventRecord(h->ev1, s));
  int rc = plan_derive(h);
  if (rc) return rc;
  return plan_finish(h);      // synchronises: `proj` must outlive its copy
}

int fcvm_grid_outer_check(fcvm_grid_t* h, const double* outers6, int64_t n, double check_scale) {
  if (!h || !outers6 || n <= 0) return fail(FCVM_EINVAL, "bad arguments");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  CU(h->d_in.reserve((size_t)6 * n));
  CU(cudaMemcpyAsync(h->d_in.p, outers6, (size_t)6 * n * sizeof(double), cudaMemcpyHostToDevice, s));
  PlaneGeom pg;
  plane_geom(h, pg);
  GridDesc g;
  plan_desc(h, h->d_int.p, g);
  CU(cudaEventRecord(h->ev0, s));
  k_outer_check<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(g, pg, h->d_in.p, n, check_scale, h->d_int.p, h->d_counters.p);
  CU(cudaGetLastError());
  CU(cudaEventRecord(h->ev1, s));
  int rc = plan_derive(h);
  if (rc) return rc;
  return plan_finish(h);
}

int fcvm_grid_get_buffers(fcvm_grid_t* h, int8_t* hc, int8_t* inflate, int8_t* internal) {
  if (!h) return fail(FCVM_EINVAL, "null handle");
  cudaStream_t s = h->stream;
  CU(cudaSetDevice(h->device));
  const long long G = (long long)h->dims[0] * h->dims[1] * h->dims[2];
  DevBuf<signed char> tmp;
  CU(tmp.reserve((size_t)G));
  int8_t* dst[3] = {hc, inflate, internal};
  const unsigned long long* src[3] = {h->d_occ.p, h->d_inf.p, h->d_int.p};
  for (int k = 0; k < 3; ++k) {
    if (!dst[k]) continue;
    k_unpack_plane<<<(unsigned)((G + 255) / 256), 256, 0, s>>>(src[k], h->dims[0], h->dims[1], h->dims[2], h->nb[1], h->nb[2], tmp.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(dst[k], tmp.p, (size_t)G, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  return FCVM_OK;
}

int64_t fcvm_grid_oob_writes(fcvm_grid_t* h) { return h ? h->oob_writes : 0; }

int fcvm_grid_counters(fcvm_grid_t* h, int64_t* out3) {
  if (!h || !out3) return fail(FCVM_EINVAL, "bad arguments");
  for (int k = 0; k < 3; ++k) out3[k] = h->counters[k];
  return FCVM_OK;
}

double fcvm_grid_last_kernel_ms(fcvm_grid_t* h) { return h ? h->last_kernel_ms : 0.0; }

}  // extern "C"
