// fcvm_device.cuh — device-side primitives of the visibility path (sm_100a).
//
// Everything here computes in FP64 with the exact operation order of the reference and is
// compiled with -fmad=false, so results are bit-identical to the CPU build of the reference
// (x86-64 SSE2, no FMA contraction; FC-Planner/src/viewpoint_manager/CMakeLists.txt:4-8).
// Only + - * / sqrt floor fmod trunc are used on the device; every transcendental of the path
// (asin / atan2 / acos / sin / cos) runs on the host with the same libm as the reference.
//
// Reference statements restated here (paths relative to FC-Planner/src/):
//   plan_env/src/raycast.cpp:26-43 (signum, mod, intbound), 341-345 (setParams),
//   347-390 (input), 392-444 (biInput), 446-479 (nextId), 481-559 (biNextId);
//   plan_env/include/plan_env/sdf_map.h:244-247 (indexToPos_hc), 367-373 (isInMap_hc),
//   397-420 (occCheck);  active_perception/src/perception_utils.cpp:115-124 (insideFOV);
//   viewpoint_manager/src/viewpoint_manager.cpp:386-396, 463-477, 482-492, 553-580, 844-853.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace fcvm {

// ---------------------------------------------------------------------------------------------
// Occupancy grid in HBM/L2: 1 bit per voxel (set = occupied), stored as 4 x 4 x 4 bricks of one
// uint64 each, bricks laid out z-fastest.  Consecutive DDA steps stay inside the same 8-byte word
// ~3 steps out of 4 (the word is cached in registers), a 128-byte line holds a 4 x 4 x 64 column,
// and the address arithmetic is 3 shifts + 2 IMADs.  The reference stores one `char` per voxel,
// x-major (sdf_map.h:266-269: x*Ny*Nz + y*Nz + z).
// ---------------------------------------------------------------------------------------------
struct GridDesc {
  const unsigned long long* bricks;  // [nbx*nby*nbz]
  int nx, ny, nz;                    // map_voxel_num_
  int nbx, nby, nbz;                 // bricks per axis: ceil(n/4), padded (pad_bricks)
  int nly, nlz;                      // lines per axis in y and z: nby / 2, nbz / 4
  int nbricks;                       // nbx * nby * nbz (< 2^31: the reference addresses voxels with int)
  int guard_steps;                   // fast-path half-rays make at most this many steps; the allocation has guard words for them
  // summed-area table over 8 x 8 x 8-voxel cells (2 x 2 x 2 bricks): sat[(x * sat_ny + y) * sat_nz + z] = number of
  // non-empty cells with coordinates < (x, y, z).  Eight loads tell whether a box of voxels is entirely free; k_eval uses
  // it to certify the viewpoint-side half-ray, which then needs no marching (nullptr = feature off).
  const int* sat;
  int sat_ny, sat_nz;                // (nby / 2 + 1), (nbz / 2 + 1)
  double res;                        // resolution_
  double res_rcp;                    // RN(1 / resolution_), for div_exact
  double offx, offy, offz;           // RayCaster::offset_ = 0.5 - origin/res   (raycast.cpp:344)
  double orgx, orgy, orgz;           // map_origin_
};

#ifndef FCVM_BRICK_LINES
#define FCVM_BRICK_LINES 0      // measured: 88.2 ms per bench step with lines vs 84.0 ms flat (the extra index arithmetic costs more than the L1 hits return)
#endif
// Storage order of the bricks.  nby / nbz = bricks per axis, padded to a multiple of 2 (x, y) and 4 (z) (pad_bricks).
// FCVM_BRICK_LINES: the 16 bricks of a 128-byte line form a 2 x 2 x 4 block of bricks (8 x 8 x 16 voxels), lines z-fastest,
// so a ray stays inside one line for ~8-16 steps whatever its direction; flat (0): bricks z-fastest, a line is a
// 4 x 4 x 64-voxel column that x- and y-going rays leave every 4 steps.
__host__ __device__ __forceinline__ long long brick_word_index(int x, int y, int z, int nby, int nbz) {
#if FCVM_BRICK_LINES
  const long long line = ((long long)(x >> 3) * (nby >> 1) + (y >> 3)) * (nbz >> 2) + (z >> 4);
  return (line << 4) | (((x >> 2) & 1) << 3) | (((y >> 2) & 1) << 2) | ((z >> 2) & 3);
#else
  return ((long long)(x >> 2) * nby + (y >> 2)) * nbz + (z >> 2);
#endif
}
__host__ __device__ __forceinline__ void pad_bricks(const int dims[3], int nb[3]) {
  nb[0] = (((dims[0] + 3) >> 2) + 1) & ~1;
  nb[1] = (((dims[1] + 3) >> 2) + 1) & ~1;
  nb[2] = (((dims[2] + 3) >> 2) + 3) & ~3;
}
// guard words on each side of the brick array that cover every index reachable within `steps` voxel steps of the map
__host__ __device__ __forceinline__ long long brick_guard_words(int steps, int nby, int nbz) {
#if FCVM_BRICK_LINES
  const long long G = steps / 8 + 1;
  return 16 * G * ((long long)(nby >> 1) * (nbz >> 2) + (nbz >> 2) + 1);
#else
  const long long G = steps / 4 + 1;
  return G * ((long long)nby * nbz + nbz + 1);
#endif
}
__host__ __device__ __forceinline__ int brick_bit_index(int x, int y, int z) {
  return ((x & 3) << 4) | ((y & 3) << 2) | (z & 3);
}

// The summed-area table's cells are (1 << FCVM_SAT_SHIFT)^3 voxels: 3 = 2 x 2 x 2 bricks (6.4 MB at config 5), 2 = one brick, the default
// (51 MB there, but a ray leaves the cells of its own surface patch after half as many steps).
#ifndef FCVM_SAT_SHIFT
#define FCVM_SAT_SHIFT 2
#endif
#define FCVM_SAT_BRICKS (1 << (FCVM_SAT_SHIFT - 2))     // bricks per cell and axis
// number of non-empty cells in the inclusive cell box [lo, hi] (coordinates already >> FCVM_SAT_SHIFT)
__device__ __forceinline__ int sat_box_count(const int* __restrict__ s, int ny, int nz, int x0, int y0, int z0, int x1, int y1, int z1) {
  // the table has fewer than 2^31 entries (fcvm_set_map_cloud checks), so the eight offsets are 32-bit arithmetic: four row
  // bases, each read at z0 and z1 + 1
  const int r00 = (x0 * ny + y0) * nz, r01 = (x0 * ny + (y1 + 1)) * nz, r10 = ((x1 + 1) * ny + y0) * nz, r11 = ((x1 + 1) * ny + (y1 + 1)) * nz;
  const int* lo = s + z0;
  const int* hi = s + (z1 + 1);
  return ((__ldg(hi + r11) - __ldg(hi + r01)) - (__ldg(hi + r10) - __ldg(hi + r00))) - ((__ldg(lo + r11) - __ldg(lo + r01)) - (__ldg(lo + r10) - __ldg(lo + r00)));
}

__device__ __forceinline__ int sat_box_count(const GridDesc& g, int x0, int y0, int z0, int x1, int y1, int z1) {
  return sat_box_count(g.sat, g.sat_ny, g.sat_nz, x0, y0, z0, x1, y1, z1);
}

// One-entry register cache of the last brick a ray touched.
struct BrickCache {
  long long key;
  unsigned long long word;
  __device__ __forceinline__ void reset() { key = -1; word = 0ull; }
};

// SDFMap::occCheck (sdf_map.h:397-420): true = inside the map AND not occupied.
__device__ __forceinline__ bool occ_free(const GridDesc& g, int ix, int iy, int iz, BrickCache& c) {
  if ((unsigned)ix >= (unsigned)g.nx || (unsigned)iy >= (unsigned)g.ny || (unsigned)iz >= (unsigned)g.nz) return false;
  long long k = brick_word_index(ix, iy, iz, g.nby, g.nbz);
  if (k != c.key) {
    c.key = k;
    c.word = (unsigned long long)k < (unsigned long long)g.nbricks ? __ldg(g.bricks + k) : ~0ull;   // unreachable after the bounds test; belt and braces
  }
  return ((c.word >> brick_bit_index(ix, iy, iz)) & 1ull) == 0ull;
}

// `idx = (Vector3d(x,y,z) + offset_).cast<int>()` for one axis (raycast.cpp:447-448): int -> double,
// one add, truncation toward zero (so a value in (-1, 0) maps to voxel 0).
__device__ __forceinline__ int to_map_index(int v, double off) { return __double2int_rz((double)v + off); }

// raycast.cpp:26-43
// mod(value, 1) = fmod(fmod(value, 1) + 1, 1).  fmod(v, 1) is exact and equals v - trunc(v), and that
// subtraction is itself exact in FP64 for every finite v, so the two fmod calls are replaced by
// trunc/sub pairs; the "+ 1" in between is the same rounded addition the reference performs.
__device__ __forceinline__ double d_mod1(double value) {
  const double a = value - trunc(value);
  const double b = a + 1.0;
  return b - trunc(b);
}
// a / b, correctly rounded, from y = RN(1 / b): q0 = RN(a y), r = a - b q0 (exact, one fused multiply-add), q = RN(q0 + r y)
// (Markstein's theorem; explicit fma intrinsics are not subject to -fmad=false).  Exact for every finite a, b as long as the
// quotient is a normal number - which is all the marchers feed it (tests/ddmath_check.cpp re-checks it against the divider on
// 2 10^8 arguments; the only miss there is a subnormal numerator).  The compiler's own FP64 division is ~70 instructions per
// site, and the DDA set-up has 18 of them per ray.
__device__ __forceinline__ double div_exact(double a, double b, double y) {
  const double q0 = a * y;
  const double r = __fma_rn(-b, q0, a);
  return __fma_rn(r, y, q0);
}
// intbound (raycast.cpp:29-43) with the integer voxel delta the call sites pass ((double)(end - start)): the quotient
// (1 - s) / |delta| has 2^-53 <= 1 - s <= 1 and 1 <= |delta|, so div_exact applies with y from the reciprocal table;
// delta == 0 gives +inf like the division by zero it replaces.  |delta| must be below the table's size (the fast path's
// guard_steps is clamped to it on the host).
__device__ __forceinline__ double d_intbound_i(double s, int delta, const double* __restrict__ rcp_table) {
  if (delta < 0) s = -s;
  s = d_mod1(s);
  const int ad = abs(delta);
  const double a = 1 - s;
  const double q = div_exact(a, (double)ad, __ldg(rcp_table + ad));
  return ad == 0 ? __longlong_as_double(0x7ff0000000000000ll) : q;
}
__device__ __forceinline__ double d_intbound(double s, double ds) {
  if (ds < 0) { s = -s; ds = -ds; }       // intbound(-s, -ds): one level of recursion at most
  s = d_mod1(s);
  return (1 - s) / ds;                     // ds == 0 -> +inf, as on the host
}
__device__ __forceinline__ int d_signum(int x) { return x == 0 ? 0 : (x < 0 ? -1 : 1); }

// One half-ray (or the one-directional ray) of the Amanatides-Woo marcher: position in
// world-anchored voxel coordinates floor(p/res), target voxel, and the tMax/tDelta state derived
// from the INTEGER voxel deltas (raycast.cpp:412-434).
struct Marcher {
  int x, y, z;        // current voxel (world-anchored lattice)
  int ex, ey, ez;     // end / midpoint voxel
  int sx, sy, sz;     // step direction
  double tmx, tmy, tmz, tdx, tdy, tdz;
  int ix, iy, iz;     // current voxel as a map index (maintained incrementally)

  __device__ __forceinline__ void init(double fx, double fy, double fz, int tx, int ty, int tz, const GridDesc& g) {
    x = (int)floor(fx); y = (int)floor(fy); z = (int)floor(fz);
    ex = tx; ey = ty; ez = tz;
    double dx = (double)(ex - x), dy = (double)(ey - y), dz = (double)(ez - z);
    sx = d_signum(ex - x); sy = d_signum(ey - y); sz = d_signum(ez - z);
    tmx = d_intbound(fx, dx); tmy = d_intbound(fy, dy); tmz = d_intbound(fz, dz);
    tdx = ((double)sx) / dx; tdy = ((double)sy) / dy; tdz = ((double)sz) / dz;   // 0/0 = NaN on idle axes
    ix = to_map_index(x, g.offx); iy = to_map_index(y, g.offy); iz = to_map_index(z, g.offz);
  }
  // the same set-up with every division taken from the reciprocal table (d_intbound_i, 1 / |delta|): the caller guarantees
  // finite coordinates and |delta| < the table's size on every axis
  __device__ __forceinline__ void init_table(double fx, double fy, double fz, int tx, int ty, int tz, const GridDesc& g, const double* __restrict__ rcp) {
    x = (int)floor(fx); y = (int)floor(fy); z = (int)floor(fz);
    ex = tx; ey = ty; ez = tz;
    const int dx = ex - x, dy = ey - y, dz = ez - z;
    sx = d_signum(dx); sy = d_signum(dy); sz = d_signum(dz);
    tmx = d_intbound_i(fx, dx, rcp); tmy = d_intbound_i(fy, dy, rcp); tmz = d_intbound_i(fz, dz, rcp);
    tdx = __ldg(rcp + abs(dx)); tdy = __ldg(rcp + abs(dy)); tdz = __ldg(rcp + abs(dz));   // rcp[0] = NaN = 0 / 0 on idle axes
    ix = to_map_index(x, g.offx); iy = to_map_index(y, g.offy); iz = to_map_index(z, g.offz);
  }
  __device__ __forceinline__ bool at_end() const { return x == ex && y == ey && z == ez; }
  __device__ __forceinline__ int manhattan() const { return abs(ex - x) + abs(ey - y) + abs(ez - z); }
  // the branch ladder of nextId / biNextId (raycast.cpp:455-476): strict '<', z on ties and NaN
  __device__ __forceinline__ void step(const GridDesc& g) {
    if (tmx < tmy) {
      if (tmx < tmz) { x += sx; tmx += tdx; ix = to_map_index(x, g.offx); }
      else           { z += sz; tmz += tdz; iz = to_map_index(z, g.offz); }
    } else {
      if (tmy < tmz) { y += sy; tmy += tdy; iy = to_map_index(y, g.offy); }
      else           { z += sz; tmz += tdz; iz = to_map_index(z, g.offz); }
    }
  }
};

struct RayStats {
  unsigned int lookups;   // occCheck calls (same sequence as the reference makes)
  unsigned int walk;      // of which inside the surface-offset walk
  unsigned int trips;     // step-cap trips (the reference would not terminate); must stay 0
};

// Bidirectional ray cast (BiRC) exactly as the call sites drive it:
//   biInput(a, b); [biNextId once, discarded]; while (biNextId) { f = occ(idx); r = occ(idxR);
//   if (!f || !r) break; }  f = occ(idx);  visible = f && r.
// a = viewpoint side (p-ray), b = target side (n-ray); both march to the midpoint voxel in lock
// step, idx / idxR are the voxels BEFORE the step of that call.
static __device__ __noinline__ bool birc(const GridDesc& g, double ax, double ay, double az, double bx, double by, double bz,
                                     bool discard_first, RayStats& st) {
  const double res = g.res;
  // start_ = start / resolution_;  inter_ = (0.5*(start+end)) / resolution_;  end_ = end / resolution_
  const double sxf = ax / res, syf = ay / res, szf = az / res;
  const double mxf = (0.5 * (ax + bx)) / res, myf = (0.5 * (ay + by)) / res, mzf = (0.5 * (az + bz)) / res;
  const double exf = bx / res, eyf = by / res, ezf = bz / res;
  const int mx = (int)floor(mxf), my = (int)floor(myf), mz = (int)floor(mzf);
  Marcher p, n;
  p.init(sxf, syf, szf, mx, my, mz, g);
  n.init(exf, eyf, ezf, mx, my, mz, g);
  const int kp = p.manhattan(), kn = n.manhattan();
  int budget = (kp > kn ? kp : kn) + 8;      // the reference has no cap
  BrickCache cp, cn;
  cp.reset(); cn.reset();

  bool f = true, r = true;
  int pix, piy, piz, nix, niy, niz;
  // one biNextId call: report both voxels, then advance whichever side is not at the midpoint
#define FCVM_BI_NEXT(ret)                                     \
  {                                                           \
    pix = p.ix; piy = p.iy; piz = p.iz;                       \
    const bool pflag = !p.at_end();                           \
    if (pflag) p.step(g);                                     \
    nix = n.ix; niy = n.iy; niz = n.iz;                       \
    const bool nflag = !n.at_end();                           \
    if (nflag) n.step(g);                                     \
    ret = pflag || nflag;                                     \
  }
  bool more, tripped = false;
  if (discard_first) { FCVM_BI_NEXT(more); }
  for (;;) {
    FCVM_BI_NEXT(more);
    if (!more) break;
    // A marcher that has not arrived after Manhattan + 8 calls has run past its midpoint voxel and will never report "at
    // end": the reference then walks on until an occCheck fails (an occupied voxel or the map's edge) and returns NOT visible.
    if (--budget < 0) { st.trips++; tripped = true; break; }
    f = occ_free(g, pix, piy, piz, cp);
    r = occ_free(g, nix, niy, niz, cn);
    st.lookups += 2;
    if (!f || !r) break;
  }
#undef FCVM_BI_NEXT
  f = occ_free(g, pix, piy, piz, cp);
  st.lookups += 1;
  return f && r && !tripped;
}

// Surface-offset walk (viewpoint_manager.cpp:463-477 / 554-568): march point -> viewpoint with the
// one-directional caster, stop at the 2nd free voxel (or at the viewpoint's voxel), return that
// voxel's centre via indexToPos_hc (sdf_map.h:244-247).
// With `rcp` (a table of 1 / k, k < rcp_n) the quotients point / res and viewpoint / res arrive precomputed (sxf .. ezf, the
// very divisions, made once per point / per viewpoint) and the DDA set-up divides through the table (div_exact) whenever
// the ray is finite and shorter than the table on every axis; otherwise, and with rcp == nullptr, it divides literally.
__device__ __forceinline__ void offset_target(const GridDesc& g, double px, double py, double pz, double vx, double vy, double vz,
                                              double& tx, double& ty, double& tz, RayStats& st, const double* __restrict__ rcp = nullptr, int rcp_n = 0,
                                              double sxf = 0.0, double syf = 0.0, double szf = 0.0, double exf = 0.0, double eyf = 0.0, double ezf = 0.0) {
  const double res = g.res;
  if (rcp == nullptr) {
    sxf = px / res; syf = py / res; szf = pz / res;
    exf = vx / res; eyf = vy / res; ezf = vz / res;
  }
  Marcher m;
  const bool table = rcp != nullptr && ((fabs(sxf) + fabs(syf)) + fabs(szf)) + ((fabs(exf) + fabs(eyf)) + fabs(ezf)) < 1.0e9 &&
                     abs((int)floor(exf) - (int)floor(sxf)) < rcp_n && abs((int)floor(eyf) - (int)floor(syf)) < rcp_n &&
                     abs((int)floor(ezf) - (int)floor(szf)) < rcp_n;
  if (table) m.init_table(sxf, syf, szf, (int)floor(exf), (int)floor(eyf), (int)floor(ezf), g, rcp);
  else m.init(sxf, syf, szf, (int)floor(exf), (int)floor(eyf), (int)floor(ezf), g);
  int budget = m.manhattan() + 8;
  BrickCache c;
  c.reset();
  int count = 0;
  int ix = m.ix, iy = m.iy, iz = m.iz;
  for (;;) {
    ix = m.ix; iy = m.iy; iz = m.iz;              // nextId reports the voxel before stepping
    if (m.at_end()) break;                        // ... and returns false at the end voxel
    if (--budget < 0) { st.trips++; break; }
    m.step(g);
    st.lookups += 1; st.walk += 1;
    if (occ_free(g, ix, iy, iz, c)) {
      count++;
      if (count == 2) break;
    }
  }
  tx = ((double)ix + 0.5) * res + g.orgx;
  ty = ((double)iy + 0.5) * res + g.orgy;
  tz = ((double)iz + 0.5) * res + g.orgz;
}

// ---------------------------------------------------------------------------------------------
// One end of a ray as RayCaster::biInput sees it (raycast.cpp:394-396: start_ = start / resolution_): the three quotients,
// computed ONCE per viewpoint / per model point instead of once per ray (IEEE division: the same bits wherever it runs), and
// whether the end lies strictly inside its voxel on x / y / z (bits 0..2 of `inside`), which is what lets a half-ray be
// certified (k_eval, CONT_SANE / CONT_PFREE).  32 bytes = one sector.
// ---------------------------------------------------------------------------------------------
struct alignas(32) RayEnd {
  double sx, sy, sz;
  double inside;            // 0..7 as a double
};
__host__ __device__ __forceinline__ RayEnd make_ray_end(double x, double y, double z, double res) {
  RayEnd e;
  e.sx = x / res; e.sy = y / res; e.sz = z / res;
  const double fx = e.sx - floor(e.sx), fy = e.sy - floor(e.sy), fz = e.sz - floor(e.sz);
  const int bits = ((fx > 1e-9 && fx < 1.0 - 1e-9) ? 1 : 0) | ((fy > 1e-9 && fy < 1.0 - 1e-9) ? 2 : 0) | ((fz > 1e-9 && fz < 1.0 - 1e-9) ? 4 : 0);
  e.inside = (double)bits;
  return e;
}

// ---------------------------------------------------------------------------------------------
// Per-viewpoint record: position + the four world-frame inward FOV normals (top, bottom, left,
// right) computed on the host by PerceptionUtils::setPose_PY (perception_utils.cpp:76-96).
// 16 doubles = 128 bytes, one line.
// ---------------------------------------------------------------------------------------------
struct alignas(16) VpRec {
  double px, py, pz;
  double n[4][3];
  double pad;
};

struct FovConst {
  double max_dist;        // perception_utils/max_dist
  double md2_lo, md2_hi;  // max_dist^2 * (1 -/+ 1e-12): outside this band the sqrt is not needed
  double plane_eps;       // |d.n| below this (unnormalised) -> take the exact normalised path
  double visible_range;   // viewpoint_manager/visible_range (E1 only)
  double vr2_lo, vr2_hi;  // visible_range^2 * (1 -/+ 1e-12): outside this band `ray_length <= visible_range` needs no sqrt
};

// PerceptionUtils::insideFOV (perception_utils.cpp:115-124), Eigen order:
//   dir = p - pos; if (dir.norm() > max_dist) false; dir.normalize(); all four dir.dot(n) >= 0.
// Fast path: the verdict of the normalised test is decided by the sign of the UNNORMALISED dot
// product whenever that is farther from zero than any rounding error could move it (bound:
// |err| < 1e-14 * |d|, the band used is 1e-10 * max_dist); points inside the band, and points
// within 1e-12 relative of the range sphere, take the literal path with sqrt and three divisions.
__device__ __forceinline__ bool inside_fov(const VpRec& v, const FovConst& fc, double qx, double qy, double qz) {
  double dx = qx - v.px, dy = qy - v.py, dz = qz - v.pz;
  const double n2 = (dx * dx + dy * dy) + dz * dz;
  if (n2 > fc.md2_hi) return false;
  bool exact = !(n2 < fc.md2_lo);
  double s[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s[q] = (dx * v.n[q][0] + dy * v.n[q][1]) + dz * v.n[q][2];
    exact |= fabs(s[q]) <= fc.plane_eps;
  }
  if (!exact) return s[0] > 0.0 && s[1] > 0.0 && s[2] > 0.0 && s[3] > 0.0;
  // literal path
  const double nrm = sqrt(n2);
  if (nrm > fc.max_dist) return false;
  if (n2 > 0.0) { dx = dx / nrm; dy = dy / nrm; dz = dz / nrm; }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const double d = (dx * v.n[q][0] + dy * v.n[q][1]) + dz * v.n[q][2];
    if (d < 0.0) return false;
  }
  return true;
}

}  // namespace fcvm
