// fcvm_map.cu — the map side of the path on the device:
//   * HC occupancy grid built on the GPU from the uploaded map cloud
//       SDFMap::initHCMap, plan_env/src/sdf_map.cpp:140-187 (pcl::getMinMax3D, the per-point posToIndex_hc + mark)
//   * a uniform cell index over the map cloud for exact fixed-radius queries (replaces pcl::KdTreeFLANN at
//       viewpoint_manager.cpp:933, k = 1 nearestKSearch, of which nearCheck only uses "distance < safe_radius")
//   * batched ViewpointManager::viewpointSafetyCheck + nearCheck (viewpoint_manager.cpp:889-946), one warp per
//       candidate position: ground test, the lattice of SDFMap::safety_check probes (sdf_map.h:375-395) with the
//       reference's early exit (the number of probes made is reproduced exactly), nearest-point test.
// FP64 / FP32 arithmetic in the reference's order, compiled with -fmad=false (no contraction), IEEE sqrtf.
#include <cfloat>

#include "fcvm_kernels.h"

namespace fcvm {

// ---- bounding box of the cloud: per-block partial (min xyz, max xyz); the host folds the partials ----------------
__global__ void __launch_bounds__(256) k_cloud_minmax(const float* __restrict__ xyz, long long n, float* __restrict__ partial) {
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __ldg(xyz + 3 * i + c);
      mn[c] = fminf(mn[c], v);
      mx[c] = fmaxf(mx[c], v);
    }
  }
  __shared__ float red[8][6];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) { red[warp][c] = mn[c]; red[warp][3 + c] = mx[c]; }
  __syncthreads();
  if (threadIdx.x < 6) {
    float r = red[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) r = threadIdx.x < 3 ? fminf(r, red[w][threadIdx.x]) : fmaxf(r, red[w][threadIdx.x]);
    partial[blockIdx.x * 6 + threadIdx.x] = r;
  }
}

// ---- voxelisation: id = floor((p - origin) * res_inv) per axis (sdf_map.h:234-237), bit set in the bricked grid ---
__global__ void __launch_bounds__(256) k_voxelise(const float* __restrict__ xyz, long long n, MapGeom g, unsigned long long* __restrict__ bricks) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = (int)floor(((double)__ldg(xyz + 3 * i) - g.org[0]) * g.res_inv);
  const int y = (int)floor(((double)__ldg(xyz + 3 * i + 1) - g.org[1]) * g.res_inv);
  const int z = (int)floor(((double)__ldg(xyz + 3 * i + 2) - g.org[2]) * g.res_inv);
  if ((unsigned)x >= (unsigned)g.dims[0] || (unsigned)y >= (unsigned)g.dims[1] || (unsigned)z >= (unsigned)g.dims[2]) return;   // reference: unchecked write
  atomicOr(bricks + brick_word_index(x, y, z, g.nb[1], g.nb[2]), 1ull << brick_bit_index(x, y, z));
}

// ---- the voxels of the padded brick array that lie outside the map read 'occupied', like occCheck outside the map
// (sdf_map.h:397-420 returns false there): a fast-path ray that strays beyond the last voxel stops where the reference's would
__global__ void __launch_bounds__(256) k_mark_padding(unsigned long long* __restrict__ bricks, int nx, int ny, int nz, int nbx, int nby, int nbz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)nbx * nby * nbz) return;
  const int bz = (int)(i % nbz), by = (int)((i / nbz) % nby), bx = (int)(i / ((long long)nbz * nby));
  if (4 * bx + 3 < nx && 4 * by + 3 < ny && 4 * bz + 3 < nz) return;       // interior brick
  unsigned long long m = 0ull;
  for (int x = 0; x < 4; ++x)
    for (int y = 0; y < 4; ++y)
      for (int z = 0; z < 4; ++z)
        if (4 * bx + x >= nx || 4 * by + y >= ny || 4 * bz + z >= nz) m |= 1ull << brick_bit_index(x, y, z);
  bricks[i] |= m;
}

// ---- summed-area table of "cell has an occupied voxel" over cells of FCVM_SAT_BRICKS^3 bricks (GridDesc::sat) ------------------
// flags into sat[(x + 1, y + 1, z + 1)] (the table is zeroed first), then one inclusive prefix pass per axis.
__global__ void __launch_bounds__(256) k_sat_flags(const unsigned long long* __restrict__ bricks, int nby, int nbz, int cx, int cy, int cz, int* __restrict__ sat) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)cx * cy * cz) return;
  const int z = (int)(i % cz), y = (int)((i / cz) % cy), x = (int)(i / ((long long)cz * cy));
  unsigned long long any = 0ull;
#pragma unroll
  for (int a = 0; a < FCVM_SAT_BRICKS; ++a)
#pragma unroll
    for (int b = 0; b < FCVM_SAT_BRICKS; ++b)
#pragma unroll
      for (int c = 0; c < FCVM_SAT_BRICKS; ++c)
        any |= __ldg(bricks + ((long long)(FCVM_SAT_BRICKS * x + a) * nby + (FCVM_SAT_BRICKS * y + b)) * nbz + (FCVM_SAT_BRICKS * z + c));
  sat[((long long)(x + 1) * (cy + 1) + (y + 1)) * (cz + 1) + (z + 1)] = any ? 1 : 0;
}
// axis 0 / 1 / 2 = prefix along x / y / z; one thread per line of the other two axes
__global__ void __launch_bounds__(256) k_sat_scan(int* __restrict__ sat, int nx, int ny, int nz, int axis) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long base, stride;
  int len;
  if (axis == 2) { if (i >= (long long)nx * ny) return; base = i * nz; stride = 1; len = nz; }
  else if (axis == 1) { if (i >= (long long)nx * nz) return; base = (i / nz) * (long long)ny * nz + (i % nz); stride = nz; len = ny; }
  else { if (i >= (long long)ny * nz) return; base = i; stride = (long long)ny * nz; len = nx; }
  int run = 0;
  for (int k = 0; k < len; ++k) { run += sat[base + k * stride]; sat[base + k * stride] = run; }
}

// ---- cell index: count, (exclusive scan by k_exscan), fill ---------------------------------------------------------
__device__ __forceinline__ int cell_coord(const CellGeom& cg, double v, int c) {
  const int k = (int)floor((v - (double)cg.mn[c]) / cg.cell);
  return k < 0 ? 0 : (k >= cg.nc[c] ? cg.nc[c] - 1 : k);
}
__device__ __forceinline__ long long cell_of(const CellGeom& cg, float x, float y, float z) {
  return ((long long)cell_coord(cg, (double)x, 0) * cg.nc[1] + cell_coord(cg, (double)y, 1)) * cg.nc[2] + cell_coord(cg, (double)z, 2);
}
__global__ void __launch_bounds__(256) k_cell_count(const float* __restrict__ xyz, long long n, CellGeom cg, int* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  atomicAdd(count + cell_of(cg, __ldg(xyz + 3 * i), __ldg(xyz + 3 * i + 1), __ldg(xyz + 3 * i + 2)), 1);
}
__global__ void __launch_bounds__(256) k_cell_fill(const float* __restrict__ xyz, long long n, CellGeom cg, const long long* __restrict__ start,
                                                   int* __restrict__ fill, float4* __restrict__ sorted) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = __ldg(xyz + 3 * i), y = __ldg(xyz + 3 * i + 1), z = __ldg(xyz + 3 * i + 2);
  const long long c = cell_of(cg, x, y, z);
  sorted[start[c] + atomicAdd(fill + c, 1)] = make_float4(x, y, z, 0.f);
}

// ---- viewpointSafetyCheck + nearCheck, one warp per candidate ----------------------------------------------------
// SDFMap::safety_check (sdf_map.h:375-395): inside the map and not occupied (the internal buffer is all zero here)
__device__ __forceinline__ bool safety_probe(const SafetyArgs& a, double px, double py, double pz) {
  const int x = (int)floor((px - a.geom.org[0]) * a.geom.res_inv);
  const int y = (int)floor((py - a.geom.org[1]) * a.geom.res_inv);
  const int z = (int)floor((pz - a.geom.org[2]) * a.geom.res_inv);
  if ((unsigned)x >= (unsigned)a.geom.dims[0] || (unsigned)y >= (unsigned)a.geom.dims[1] || (unsigned)z >= (unsigned)a.geom.dims[2]) return false;
  const unsigned long long w = __ldg(a.bricks + brick_word_index(x, y, z, a.geom.nb[1], a.geom.nb[2]));
  return ((w >> brick_bit_index(x, y, z)) & 1ull) == 0ull;
}
// number of lattice steps of one axis: for (i = 0; v - sr + i * res < v + sr; i += safe_voxel)
__device__ __forceinline__ int lattice_steps(double v, double sr, double res, int safe_voxel) {
  int cnt = 0;
  for (int i = 0; v - sr + i * res < v + sr && cnt < 4096; i += safe_voxel) ++cnt;
  return cnt;
}

__global__ void __launch_bounds__(256) k_safety(const __grid_constant__ SafetyArgs a) {
  const long long v = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= a.n) return;
  const int lane = threadIdx.x & 31;
  const float fx = __ldg(a.cand + 3 * v), fy = __ldg(a.cand + 3 * v + 1), fz = __ldg(a.cand + 3 * v + 2);
  const double vx = (double)fx, vy = (double)fy, vz = (double)fz;
  bool ok = true;
  unsigned long long looks = 0;
  // Free-space certificate: the box of voxels around the candidate that reaches one voxel beyond +- safe_radius on every axis
  // lies inside the map and holds no occupied voxel (summed-area table, 8 loads).  Every probe of the lattice falls into that
  // box, so all of them are "inside the map and free"; every map point lies in an occupied voxel (k_voxelise marks the voxel
  // of each), so none is within safe_radius either.  Both loops are then only counted.  Candidates are generated
  // viewpoints_distance away from the surface, so this is the common case.
  bool certified = false;
  if (a.sat != nullptr && fabs(vx) < 1e15 && fabs(vy) < 1e15 && fabs(vz) < 1e15) {      // false for NaN
    const double r = a.sr * 1.0001 + 1e-6;
    const int x0 = (int)floor((vx - r - a.geom.org[0]) * a.geom.res_inv) - 1, x1 = (int)floor((vx + r - a.geom.org[0]) * a.geom.res_inv) + 1;
    const int y0 = (int)floor((vy - r - a.geom.org[1]) * a.geom.res_inv) - 1, y1 = (int)floor((vy + r - a.geom.org[1]) * a.geom.res_inv) + 1;
    const int z0 = (int)floor((vz - r - a.geom.org[2]) * a.geom.res_inv) - 1, z1 = (int)floor((vz + r - a.geom.org[2]) * a.geom.res_inv) + 1;
    if (x0 >= 0 && y0 >= 0 && z0 >= 0 && x1 < a.geom.dims[0] && y1 < a.geom.dims[1] && z1 < a.geom.dims[2])
      certified = sat_box_count(a.sat, a.sat_ny, a.sat_nz, x0 >> FCVM_SAT_SHIFT, y0 >> FCVM_SAT_SHIFT, z0 >> FCVM_SAT_SHIFT, x1 >> FCVM_SAT_SHIFT,
                                y1 >> FCVM_SAT_SHIFT, z1 >> FCVM_SAT_SHIFT) == 0;
  }
  if (a.mode == 0) {
    if (a.z_ground && vz < a.ground_limit) ok = false;                               // viewpoint_manager.cpp:891-892
    if (ok && a.safe_voxel <= 0) ok = false;                                          // reference: infinite loop
    if (ok) {
      const int nx = lattice_steps(vx, a.sr, a.res, a.safe_voxel), ny = lattice_steps(vy, a.sr, a.res, a.safe_voxel),
                nz = lattice_steps(vz, a.sr, a.res, a.safe_voxel);
      const long long total = (long long)nx * ny * nz;      // up to 4096^3: 64-bit
      looks = (unsigned long long)total;
      for (long long base = 0; base < total && !certified; base += 32) {
        const long long t = base + lane;
        bool bad = false;
        if (t < total) {
          const int k = (int)(t % nz), j = (int)((t / nz) % ny), i = (int)(t / ((long long)nz * ny));
          bad = !safety_probe(a, vx - a.sr + (i * a.safe_voxel) * a.res, vy - a.sr + (j * a.safe_voxel) * a.res, vz - a.sr + (k * a.safe_voxel) * a.res);
        }
        const unsigned m = __ballot_sync(0xffffffffu, bad);
        if (m) {                                   // the reference returns at the first blocked probe
          looks = (unsigned long long)(base + __ffs(m));
          ok = false;
          break;
        }
      }
    }
  }
  // nearCheck (viewpoint_manager.cpp:916-946) with vox_num > 1: fails when the nearest map point is closer than safe_radius
  if (ok && !(certified && a.n_map > 0)) {
    if (!(isfinite(vx) && isfinite(vy) && isfinite(vz)) || a.n_map == 0) {
      ok = false;
    } else if (fabs(vx) >= 1e15 || fabs(vy) >= 1e15 || fabs(vz) >= 1e15) {
      // Absurdly far positions: the float squared distances can overflow.  The k = 1 search only reports a neighbour whose
      // squared distance is below FLT_MAX; with none the reference's `dist` is 0 and the position is dropped.  Any
      // neighbour that is reported is farther than safe_radius.  Brute force over the cloud (never taken on real input).
      bool any = false;
      for (long long e = lane; e < a.n_map && !any; e += 32) {
        const float4 p = __ldg(a.sorted + e);
        float result = 0.0f, diff;
        diff = fx - p.x; result += diff * diff;
        diff = fy - p.y; result += diff * diff;
        diff = fz - p.z; result += diff * diff;
        any = result < FLT_MAX;
      }
      ok = __ballot_sync(0xffffffffu, any) != 0u;
    } else {
      const CellGeom& cg = a.cells;
      const double r = a.sr * 1.0001 + 1e-6;        // covers the float rounding of the distance test
      const int x0 = cell_coord(cg, vx - r, 0), x1 = cell_coord(cg, vx + r, 0);
      const int y0 = cell_coord(cg, vy - r, 1), y1 = cell_coord(cg, vy + r, 1);
      const int z0 = cell_coord(cg, vz - r, 2), z1 = cell_coord(cg, vz + r, 2);
      const int ny = y1 - y0 + 1, nz = z1 - z0 + 1;
      const int total = (x1 - x0 + 1) * ny * nz;
      for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        bool found = false;
        if (t < total) {
          const int cz = z0 + t % nz, cy = y0 + (t / nz) % ny, cx = x0 + t / (nz * ny);
          const long long c = ((long long)cx * cg.nc[1] + cy) * cg.nc[2] + cz;
          for (long long e = __ldg(a.cell_start + c), e1 = __ldg(a.cell_start + c + 1); e < e1 && !found; ++e) {
            const float4 p = __ldg(a.sorted + e);
            // flann::L2_Simple<float>: sequential float accumulation
            float result = 0.0f, diff;
            diff = fx - p.x; result += diff * diff;
            diff = fy - p.y; result += diff * diff;
            diff = fz - p.z; result += diff * diff;
            found = (double)sqrtf(result) < a.sr;
          }
        }
        if (__ballot_sync(0xffffffffu, found)) { ok = false; break; }
      }
    }
  }
  if (lane == 0) {
    a.ok[v] = ok ? 1 : 0;
    if (looks) atomicAdd(a.lookups, looks);
  }
}

// ---- launchers ---------------------------------------------------------------------------------------------------
cudaError_t preload_map_kernels() {
  cudaFuncAttributes fa;
  const void* fns[] = {(const void*)k_cloud_minmax, (const void*)k_voxelise, (const void*)k_mark_padding, (const void*)k_sat_flags, (const void*)k_sat_scan,
                       (const void*)k_cell_count, (const void*)k_cell_fill, (const void*)k_safety};
  for (const void* f : fns) {
    const cudaError_t e = cudaFuncGetAttributes(&fa, f);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_cloud_minmax(const float* xyz, long long n, float* partial, int blocks, cudaStream_t s) {
  k_cloud_minmax<<<blocks, 256, 0, s>>>(xyz, n, partial);
  return cudaGetLastError();
}
cudaError_t launch_voxelise(const float* xyz, long long n, const MapGeom& g, unsigned long long* bricks, cudaStream_t s) {
  k_voxelise<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, n, g, bricks);
  return cudaGetLastError();
}
cudaError_t launch_mark_padding(unsigned long long* bricks, const int dims[3], const int nb[3], cudaStream_t s) {
  const long long n = (long long)nb[0] * nb[1] * nb[2];
  k_mark_padding<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(bricks, dims[0], dims[1], dims[2], nb[0], nb[1], nb[2]);
  return cudaGetLastError();
}

cudaError_t launch_build_sat(const unsigned long long* bricks, const int nb[3], int* sat, cudaStream_t s) {
  const int cx = nb[0] / FCVM_SAT_BRICKS, cy = nb[1] / FCVM_SAT_BRICKS, cz = nb[2] / FCVM_SAT_BRICKS;
  const long long n = (long long)(cx + 1) * (cy + 1) * (cz + 1);
  cudaError_t e = cudaMemsetAsync(sat, 0, (size_t)n * sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_sat_flags<<<(unsigned)(((long long)cx * cy * cz + 255) / 256), 256, 0, s>>>(bricks, nb[1], nb[2], cx, cy, cz, sat);
  k_sat_scan<<<(unsigned)(((long long)(cx + 1) * (cy + 1) + 255) / 256), 256, 0, s>>>(sat, cx + 1, cy + 1, cz + 1, 2);
  k_sat_scan<<<(unsigned)(((long long)(cx + 1) * (cz + 1) + 255) / 256), 256, 0, s>>>(sat, cx + 1, cy + 1, cz + 1, 1);
  k_sat_scan<<<(unsigned)(((long long)(cy + 1) * (cz + 1) + 255) / 256), 256, 0, s>>>(sat, cx + 1, cy + 1, cz + 1, 0);
  return cudaGetLastError();
}

cudaError_t launch_cell_count(const float* xyz, long long n, const CellGeom& cg, int* count, cudaStream_t s) {
  k_cell_count<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, n, cg, count);
  return cudaGetLastError();
}
cudaError_t launch_cell_fill(const float* xyz, long long n, const CellGeom& cg, const long long* start, int* fill, float4* sorted, cudaStream_t s) {
  k_cell_fill<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(xyz, n, cg, start, fill, sorted);
  return cudaGetLastError();
}
cudaError_t launch_safety(const SafetyArgs& a, cudaStream_t s) {
  if (a.n <= 0) return cudaSuccess;
  k_safety<<<(unsigned)((a.n + 7) / 8), 256, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace fcvm
