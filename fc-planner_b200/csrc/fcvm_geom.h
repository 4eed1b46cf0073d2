// fcvm_geom.h — host-side geometry of the path: everything that depends on libm or on the
// ordering of libstdc++ containers stays on the host so that it is bit-identical to the reference
// (SURVEY.md section 7 "Hard parts", Appendix A).  Compiled with -ffp-contract=off.
//
// Reference statements (paths relative to FC-Planner/src/):
//   active_perception/src/perception_utils.cpp:29-74 (init), 76-96 (setPose_PY)
//   plan_env/src/sdf_map.cpp:122-188 (initHCMap); plan_env/include/plan_env/sdf_map.h:234-247,
//   266-273, 367-395 (index math, safety_check)
//   viewpoint_manager/include/viewpoint_manager/viewpoint_manager.h:238-267 (PitchYaw, pyToVec, warpAngle)
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/fcvm.h"
#include "fcvm_device.cuh"

namespace fcvm {

struct D3 { double x, y, z; };

// Eigen's 3-term reduction order under SSE2: (a0*b0 + a1*b1) + a2*b2   
inline double dot3(const D3& a, const D3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline double norm3(const D3& a) { return std::sqrt(dot3(a, a)); }
inline D3 sub3(const D3& a, const D3& b) { return D3{a.x - b.x, a.y - b.y, a.z - b.z}; }
// MatrixBase::normalized(): divide by sqrt(squaredNorm) only when squaredNorm > 0
inline D3 normalized3(const D3& a) {
  const double z = dot3(a, a);
  if (z > 0.0) { const double n = std::sqrt(z); return D3{a.x / n, a.y / n, a.z / n}; }
  return a;
}
// viewpoint_manager.h:238-248
inline void pitch_yaw(const D3& v, double& pitch, double& yaw) {
  pitch = std::asin(v.z / (norm3(v) + 1e-3));
  yaw = std::atan2(v.y, v.x);
}
// viewpoint_manager.h:250-259
inline D3 py_to_vec(double pitch, double yaw) {
  return D3{std::cos(pitch) * std::cos(yaw), std::cos(pitch) * std::sin(yaw), std::sin(pitch)};
}
// viewpoint_manager.h:261-267
inline void warp_angle(double& a) {
  while (a < -M_PI) a += 2 * M_PI;
  while (a > M_PI) a -= 2 * M_PI;
}
// the signed wrapped difference idiom of viewpoint_manager.cpp:777-788 and 864-872:
//   off = |a-b| > pi ? 2pi - |a-b| : |a-b|;  flag = a-b > 0 ? 1 : -1;  dir = |a-b| > pi ? -1 : 1
// returned as the three factors because the reference multiplies them in source order.
inline void wrapped_diff(double a, double b, double& off, int& flag, int& dir) {
  const double ad = std::abs(a - b);
  off = ad > M_PI ? (2 * M_PI - ad) : ad;
  flag = a - b > 0 ? 1 : -1;
  dir = ad > M_PI ? -1 : 1;
}

// Camera-frame FOV normals (perception_utils.cpp:39-43)
struct Camera {
  double st, ct, sl, cl, sr, cr;  // sin/cos of (pi/2 - top), (pi/2 - left), (pi/2 - right)
  void init(const fcvm_params& p) {
    st = std::sin(M_PI / 2 - p.top_angle); ct = std::cos(M_PI / 2 - p.top_angle);
    sl = std::sin(M_PI / 2 - p.left_angle); cl = std::cos(M_PI / 2 - p.left_angle);
    sr = std::sin(M_PI / 2 - p.right_angle); cr = std::cos(M_PI / 2 - p.right_angle);
  }
  // setPose_PY (perception_utils.cpp:76-96) in closed form.  R_wb = R_yaw * R_pitch has one non-zero
  // product per entry; T_bc is a signed permutation, so R_wc = [-R_wb(:,1), R_wb(:,2), R_wb(:,0)];
  // each world normal is a two-term sum (the third term is an exact zero).
  void world_normals(double pitch, double yaw, double out[4][3]) const {
    const double cy = std::cos(yaw), sy = std::sin(yaw), cp = std::cos(pitch), sp = std::sin(pitch);
    const double R[3][3] = {{cy * cp, -sy, cy * (-sp)}, {sy * cp, cy, sy * (-sp)}, {sp, 0.0, cp}};
    for (int i = 0; i < 3; ++i) {
      const double c0 = -R[i][1], c1 = R[i][2], c2 = R[i][0];   // R_wc(i, 0..2)
      out[0][i] = c1 * st + c2 * ct;        // top    (0,  st, ct)
      out[1][i] = c1 * (-st) + c2 * ct;     // bottom (0, -st, ct)
      out[2][i] = c0 * sl + c2 * cl;        // left   ( sl, 0, cl)
      out[3][i] = c0 * (-sr) + c2 * cr;     // right  (-sr, 0, cr)
    }
  }
  void fill(VpRec& r, double x, double y, double z, double pitch, double yaw) const {
    r.px = x; r.py = y; r.pz = z; r.pad = 0.0;
    world_normals(pitch, yaw, r.n);
  }
};

// Eigen::Matrix3d::inverse() (Eigen/src/LU/InverseImpl.h, size 3): cofactors times 1 / det, with
// det = (c00*m00 + c10*m10) + c20*m20.  Used by CameraModelProjection (hcplanner.cpp:1201-1204).
inline void inverse3(const double m[3][3], double r[3][3]) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const double det = (c0 * m[0][0] + c1 * m[1][0]) + c2 * m[2][0];
  const double invdet = 1.0 / det;
  r[0][0] = c0 * invdet; r[0][1] = c1 * invdet; r[0][2] = c2 * invdet;
  r[1][0] = cof(0, 1) * invdet; r[1][1] = cof(1, 1) * invdet; r[1][2] = cof(2, 1) * invdet;
  r[2][0] = cof(0, 2) * invdet; r[2][1] = cof(1, 2) * invdet; r[2][2] = cof(2, 2) * invdet;
}
// R_p.inverse() * R_y.inverse() of CameraModelProjection (hcplanner.cpp:1199-1204), row-major into M[9]
inline void camera_pose_matrix(double pitch, double yaw, double M[9]) {
  const double Ry[3][3] = {{std::cos(yaw), -std::sin(yaw), 0.0}, {std::sin(yaw), std::cos(yaw), 0.0}, {0.0, 0.0, 1.0}};
  const double Rp[3][3] = {{std::cos(pitch), 0.0, -std::sin(pitch)}, {0.0, 1.0, 0.0}, {std::sin(pitch), 0.0, std::cos(pitch)}};
  double Ryi[3][3], Rpi[3][3];
  inverse3(Ry, Ryi);
  inverse3(Rp, Rpi);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M[3 * i + j] = (Rpi[i][0] * Ryi[0][j] + Rpi[i][1] * Ryi[1][j]) + Rpi[i][2] * Ryi[2][j];
}

// ---------------------------------------------------------------------------------------------
// HC occupancy grid on the host: geometry exactly as initHCMap derives it, storage in the bricked
// 1-bit layout the kernels read (fcvm_device.cuh).  The host copy serves safety_check.
// ---------------------------------------------------------------------------------------------
struct HostGrid {
  double res = 0, res_inv = 0;
  double origin[3] = {0, 0, 0}, size[3] = {0, 0, 0};
  int dims[3] = {0, 0, 0};
  int ns[3] = {0, 0, 0};   // bricks per axis
  std::vector<unsigned long long> bricks;
  bool ready = false;

  // geometry from the float bounding box of the map cloud (pcl::getMinMax3D), sdf_map.cpp:144-157
  void set_geometry(const fcvm_params& p, const float mn[3], const float mx[3]) {
    res = p.resolution;
    const double inflate = p.visible_range;
    res_inv = 1 / res;
    for (int c = 0; c < 3; ++c) {
      size[c] = 4 * inflate + mx[c] - mn[c];        // sdf_map.cpp:146-148 (double + float - float)
      origin[c] = mn[c] - 2 * inflate;              // sdf_map.cpp:151
      dims[c] = (int)std::ceil(size[c] / res);      // sdf_map.cpp:156-157
    }
    pad_bricks(dims, ns);
    bricks.clear();
    ready = true;
  }
  // host-side build (the CPU-only helper fcvm_host_build_grid; the product path voxelises on the device)
  void build(const fcvm_params& p, const float* xyz, int64_t n) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};   // pcl::getMinMax3D
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        const float v = xyz[3 * i + c];
        mn[c] = std::min(mn[c], v);
        mx[c] = std::max(mx[c], v);
      }
    set_geometry(p, mn, mx);
    bricks.assign((size_t)ns[0] * ns[1] * ns[2], 0ull);
    for (int64_t i = 0; i < n; ++i) {               // sdf_map.cpp:178-187
      int id[3];
      for (int c = 0; c < 3; ++c) id[c] = (int)std::floor(((double)xyz[3 * i + c] - origin[c]) * res_inv);
      if (in_map(id[0], id[1], id[2]))              // reference: unchecked write
        bricks[(size_t)brick_word_index(id[0], id[1], id[2], ns[1], ns[2])] |= 1ull << brick_bit_index(id[0], id[1], id[2]);
    }
  }
  size_t nbricks() const { return (size_t)ns[0] * ns[1] * ns[2]; }
  bool in_map(int x, int y, int z) const {          // sdf_map.h:367-373
    return !(x < 0 || y < 0 || z < 0 || x > dims[0] - 1 || y > dims[1] - 1 || z > dims[2] - 1);
  }
  bool occupied(int x, int y, int z) const {
    return (bricks[(size_t)brick_word_index(x, y, z, ns[1], ns[2])] >> brick_bit_index(x, y, z)) & 1ull;
  }
  // SDFMap::safety_check(Vector3d&) (sdf_map.h:375-395); the internal buffer is all zero here
  bool safety_check(double px, double py, double pz) const {
    const int x = (int)std::floor((px - origin[0]) * res_inv);
    const int y = (int)std::floor((py - origin[1]) * res_inv);
    const int z = (int)std::floor((pz - origin[2]) * res_inv);
    if (!in_map(x, y, z)) return false;
    return !occupied(x, y, z);
  }
  int64_t voxels() const { return (int64_t)dims[0] * dims[1] * dims[2]; }
  // reference order (x-major, sdf_map.h:266-269), 1 bit per voxel, for the parity tap
  void export_xmajor(uint8_t* bits) const {
    std::memset(bits, 0, (size_t)((voxels() + 7) / 8));
    for (int x = 0; x < dims[0]; ++x)
      for (int y = 0; y < dims[1]; ++y)
        for (int z = 0; z < dims[2]; ++z)
          if (occupied(x, y, z)) {
            const int64_t a = ((int64_t)x * dims[1] + y) * dims[2] + z;
            bits[a >> 3] |= (uint8_t)(1u << (a & 7));
          }
  }
};

// ---------------------------------------------------------------------------------------------
// Exact fixed-radius / nearest queries on a float cloud with FLANN's L2_Simple<float> arithmetic
// (sequential float accumulation), accelerated by a uniform cell hash.  Replaces
// pcl::KdTreeFLANN at viewpoint_manager.cpp:685 (radiusSearch) and :933 (nearestKSearch, k = 1).
// ---------------------------------------------------------------------------------------------
struct CellIndex {
  std::vector<float> pts;
  std::vector<int> order;          // point ids sorted by cell
  std::vector<int64_t> cell_start; // per cell
  float mn[3] = {0, 0, 0};
  double cell = 1.0;
  int nc[3] = {1, 1, 1};

  static float dist2(const float* a, const float* b) {
    float result = 0.0f, diff;
    for (int i = 0; i < 3; ++i) { diff = a[i] - b[i]; result += diff * diff; }
    return result;
  }
  size_t size() const { return pts.size() / 3; }
  void build(const float* xyz, int64_t n, double cell_size) {
    pts.assign(xyz, xyz + 3 * n);
    cell = cell_size > 1e-6 ? cell_size : 1e-6;
    float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    mn[0] = mn[1] = mn[2] = FLT_MAX;
    for (int64_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) { mn[c] = std::min(mn[c], xyz[3 * i + c]); mx[c] = std::max(mx[c], xyz[3 * i + c]); }
    if (n == 0) { mn[0] = mn[1] = mn[2] = 0; mx[0] = mx[1] = mx[2] = 0; }
    // keep the table small: grow the cell until it has at most ~4M cells
    for (;;) {
      double tot = 1;
      for (int c = 0; c < 3; ++c) { nc[c] = (int)std::floor(((double)mx[c] - mn[c]) / cell) + 1; tot *= nc[c]; }
      if (tot <= 4.0e6) break;
      cell *= 1.5;
    }
    const int64_t ncell = (int64_t)nc[0] * nc[1] * nc[2];
    std::vector<int64_t> cnt((size_t)ncell + 1, 0);
    std::vector<int64_t> key((size_t)n);
    for (int64_t i = 0; i < n; ++i) { key[(size_t)i] = cell_of(&pts[3 * (size_t)i]); cnt[(size_t)key[(size_t)i] + 1]++; }
    for (int64_t k = 0; k < ncell; ++k) cnt[(size_t)k + 1] += cnt[(size_t)k];
    cell_start = cnt;
    order.assign((size_t)n, 0);
    std::vector<int64_t> fill(cell_start.begin(), cell_start.end() - 1);
    for (int64_t i = 0; i < n; ++i) order[(size_t)fill[(size_t)key[(size_t)i]]++] = (int)i;
  }
  int cell_coord(double v, int c) const {
    int k = (int)std::floor((v - mn[c]) / cell);
    return k < 0 ? 0 : (k >= nc[c] ? nc[c] - 1 : k);
  }
  int64_t cell_of(const float* p) const {
    return ((int64_t)cell_coord(p[0], 0) * nc[1] + cell_coord(p[1], 1)) * nc[2] + cell_coord(p[2], 2);
  }
  template <class F>
  void for_each_within(const float* q, double radius, F&& f) const {
    const double r = radius * 1.0001 + 1e-6;     // cover float rounding of the distance test
    int lo[3], hi[3];
    for (int c = 0; c < 3; ++c) { lo[c] = cell_coord(q[c] - r, c); hi[c] = cell_coord(q[c] + r, c); }
    const double r2 = r * r;
    auto gap2 = [&](int k, int c) {              // squared distance from q to the slab of cell k along axis c
      const double a = (double)mn[c] + k * cell, b = a + cell;
      const double d = q[c] < a ? a - q[c] : (q[c] > b ? q[c] - b : 0.0);
      return d * d;
    };
    for (int x = lo[0]; x <= hi[0]; ++x) {
      const double gx = gap2(x, 0);
      for (int y = lo[1]; y <= hi[1]; ++y) {
        const double gxy = gx + gap2(y, 1);
        if (gxy > r2) continue;
        for (int z = lo[2]; z <= hi[2]; ++z) {
          if (gxy + gap2(z, 2) > r2) continue;   // the whole cell is farther than the query radius
          const int64_t k = ((int64_t)x * nc[1] + y) * nc[2] + z;
          for (int64_t e = cell_start[(size_t)k]; e < cell_start[(size_t)k + 1]; ++e) {
            const int id = order[(size_t)e];
            if (!f(id, dist2(q, &pts[3 * (size_t)id]))) return;
          }
        }
      }
    }
  }
  // is the (float) nearest-neighbour distance sqrt(d2) < radius ?  This is all nearCheck needs
  // (viewpoint_manager.cpp:941-943): dist = sqrt(nn_squared_distance[0]) with the float overload.
  bool any_closer_than(const float* q, double radius) const {
    bool found = false;
    for_each_within(q, radius, [&](int, float d2) {
      if ((double)std::sqrt(d2) < radius) { found = true; return false; }
      return true;
    });
    return found;
  }
  // FLANN radiusSearch: d2 < (float)(radius*radius), sorted by (distance, index)
  void radius_search(const float* q, double radius, std::vector<int>& ids) const {
    const float r2 = static_cast<float>(radius * radius);
    std::vector<std::pair<float, int>> found;
    for_each_within(q, radius, [&](int id, float d2) {
      if (d2 < r2) found.emplace_back(d2, id);
      return true;
    });
    std::sort(found.begin(), found.end());
    ids.clear();
    for (auto& f : found) ids.push_back(f.second);
  }
};

}  // namespace fcvm
