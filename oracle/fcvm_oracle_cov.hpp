// fcvm_oracle_cov.hpp — CPU ORACLE of the pin-hole coverage evaluation.  TEST INFRASTRUCTURE, NOT
// PRODUCT CODE (same rules as fcvm_oracle.hpp: only tests/, bench.py's cpu_baseline leg and
// __graft_entry__.smoke() may use it, as the checker).
//
// Restates, statement by statement, from FC-Planner/src/hierarchical_coverage_planner/src/hcplanner.cpp:
//   CameraModelProjection  :1196-1211      PinHoleCamera  :1128-1194      CoverageEvaluation  :1032-1126
// and the path-coverage loop that calls PinHoleCamera (:250-271).
//
// PIN: hcplanner.cpp cannot be compiled here (it pulls in ROSA, HCSolver, PCL, ROS), but the reference ships the OUTPUT of
// this very function: solution/Traj/CloudInfo{MBS,Pipe,Christ}.txt is `visible_group` of its own CoverageEvaluation run over
// the poses printed in TrajInfo*.txt (hctraj.cpp:197-221).  tests/golden/make_cov_golden.py maps those points back to cloud
// indices; this restatement reproduces 82 581 of 82 589 ids on MBS, 26 512 of 26 513 on pipe, 105 717 of 105 747 on Christ,
// the residue being the 6-decimal rounding of the printed poses (tests/test_oracle_cov.py asserts >= 99.9 % per scene, plus
// hand-derived known answers).  Third-party arithmetic assumed (not in this image): Eigen's 3x3 inverse =
// cofactors * (1 / det) with det = (c00*m00 + c10*m10) + c20*m20 (Eigen/src/LU/InverseImpl.h, compute_inverse<.., 3>), and
// the SSE2 3-term reduction order of fcvm_oracle.hpp for the matrix products.
#pragma once

#include <set>

#include "fcvm_oracle.hpp"

namespace orc {

struct Camera {          // hcplanner/{fx, fy, cx, cy} (hcplanner.cpp:56-59) and viewpoint_manager/{fov_h, fov_w}
  double fx, fy, cx, cy, fov_h, fov_w;
};

// Eigen::Matrix3d::inverse(): cofactor matrix times the reciprocal of the determinant
inline void inverse3(const double m[3][3], double r[3][3]) {
  auto cof = [&](int i, int j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  const double det = redux3(c0 * m[0][0], c1 * m[1][0], c2 * m[2][0]);
  const double invdet = 1.0 / det;
  r[0][0] = c0 * invdet; r[0][1] = c1 * invdet; r[0][2] = c2 * invdet;
  r[1][0] = cof(0, 1) * invdet; r[1][1] = cof(1, 1) * invdet; r[1][2] = cof(2, 1) * invdet;
  r[2][0] = cof(0, 2) * invdet; r[2][1] = cof(1, 2) * invdet; r[2][2] = cof(2, 2) * invdet;
}

struct CovCounters {
  int64_t pair_tests = 0;    // insideFOV calls
  int64_t projections = 0;   // CameraModelProjection calls
  int64_t out_of_image = 0;  // xID / yID beyond the image: undefined behaviour in the reference (unchecked Eigen index), skipped here
};

class CoverageEvaluator {
 public:
  void init(const Params& p, const Camera& c) {
    prm = p; cam = c;
    percep_utils_.init(p);
    xPixel = 2 * (int)cam.cx; yPixel = 2 * (int)cam.cy;                         // :1039, :1131
    const double xRange = cam.fx * std::tan(0.5 * cam.fov_w * M_PI / 180.0) / 1000.0;   // :1040
    const double yRange = cam.fy * std::tan(0.5 * cam.fov_h * M_PI / 180.0) / 1000.0;   // :1041
    ldc[0] = -xRange; ldc[1] = -yRange;
    xRes = 2 * xRange / xPixel; yRes = 2 * yRange / yPixel;                       // :1044
  }
  void setCloud(const float* xyz, int64_t n) { cloud.assign(xyz, xyz + 3 * n); }
  int64_t size() const { return (int64_t)cloud.size() / 3; }

  // the pose-dependent part of CameraModelProjection: R_p.inverse() * R_y.inverse()  (:1201-1204)
  static void poseMatrix(const double pose[5], double M[3][3]) {
    const double pitch = pose[3], yaw = pose[4];
    const double Ry[3][3] = {{std::cos(yaw), -std::sin(yaw), 0.0}, {std::sin(yaw), std::cos(yaw), 0.0}, {0.0, 0.0, 1.0}};
    const double Rp[3][3] = {{std::cos(pitch), 0.0, -std::sin(pitch)}, {0.0, 1.0, 0.0}, {std::sin(pitch), 0.0, std::cos(pitch)}};
    double Ryi[3][3], Rpi[3][3];
    inverse3(Ry, Ryi);
    inverse3(Rp, Rpi);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) M[i][j] = redux3(Rpi[i][0] * Ryi[0][j], Rpi[i][1] * Ryi[1][j], Rpi[i][2] * Ryi[2][j]);
  }
  // CameraModelProjection :1196-1211 with the pose matrix hoisted (it does not depend on the point)
  void project(const double pose[5], const double M[3][3], const V3& point, double& distance, int inCam[2]) const {
    const V3 pos_(pose[0], pose[1], pose[2]);
    const V3 d = point - pos_;
    distance = norm(d);
    double pc[3];
    for (int i = 0; i < 3; ++i) pc[i] = redux3(M[i][0] * d.x, M[i][1] * d.y, M[i][2] * d.z);
    const double camX = cam.fx * pc[1] / (1000.0 * pc[0]) - ldc[0];
    const double camY = cam.fy * pc[2] / (1000.0 * pc[0]) - ldc[1];
    inCam[0] = (int)std::floor(camX / xRes);
    inCam[1] = (int)std::floor(camY / yRes);
  }

  // The z-buffer scan shared by PinHoleCamera (:1154-1182) and CoverageEvaluation (:1063-1092).
  // visible = ascending unique ids of the pixel occupants (the std::set of :1191 / :1101);
  // displaced (optional) = every point that lost a pixel it held, in scan order (:1085).
  void scan(const double pose[5], std::vector<int>& visible, std::vector<int>* displaced) {
    if ((int64_t)state.size() != (int64_t)xPixel * yPixel) {
      state.assign((size_t)xPixel * yPixel, 0);
      occupant.assign((size_t)xPixel * yPixel, 0);
      zdist.assign((size_t)xPixel * yPixel, 1000.0);
    }
    touched.clear();
    double M[3][3];
    poseMatrix(pose, M);
    percep_utils_.setPose_PY(V3(pose[0], pose[1], pose[2]), pose[3], pose[4]);
    const int64_t n = size();
    for (int64_t i = 0; i < n; ++i) {
      const V3 pt(cloud[3 * i], cloud[3 * i + 1], cloud[3 * i + 2]);
      ctr.pair_tests++;
      if (!percep_utils_.insideFOV(pt)) continue;
      double camDist;
      int cod[2];
      project(pose, M, pt, camDist, cod);
      ctr.projections++;
      if (cod[0] < 0 || cod[1] < 0) continue;
      if (cod[0] >= xPixel || cod[1] >= yPixel) { ctr.out_of_image++; continue; }
      const size_t px = (size_t)cod[0] * yPixel + cod[1];
      if (!state[px]) {
        state[px] = 1; occupant[px] = (int)i; zdist[px] = camDist;
        touched.push_back(px);
      } else if (camDist < zdist[px]) {
        if (displaced) displaced->push_back(occupant[px]);
        occupant[px] = (int)i; zdist[px] = camDist;
      }
    }
    std::set<int> visSet;
    for (size_t px : touched) { visSet.insert(occupant[px]); state[px] = 0; zdist[px] = 1000.0; }
    visible.assign(visSet.begin(), visSet.end());
  }

  // PinHoleCamera :1128-1194
  void PinHoleCamera(const double pose[5], std::vector<int>& covered_id) {
    std::vector<int> vis;
    scan(pose, vis, nullptr);
    covered_id.insert(covered_id.end(), vis.begin(), vis.end());
  }

  // CoverageEvaluation :1032-1126.  visible_group[k] = ids pushed into pose k's visibleCloud.
  double CoverageEvaluation(const double* pose5, int64_t m, std::vector<std::vector<int>>& visible_group, std::vector<char>& ever_visible) {
    const int64_t n = size();
    std::vector<char> cover_state((size_t)n, 0);
    ever_visible.assign((size_t)n, 0);
    visible_group.assign((size_t)m, std::vector<int>());
    std::vector<int> vis, displaced;
    for (int64_t k = 0; k < m; ++k) {
      displaced.clear();
      scan(pose5 + 5 * k, vis, &displaced);
      for (int p : displaced) cover_state[(size_t)p] = 0;          // :1085, in scan order (idempotent)
      for (int id : vis)
        if (!cover_state[(size_t)id]) visible_group[(size_t)k].push_back(id);    // :1102-1106
      for (int id : vis) { cover_state[(size_t)id] = 1; ever_visible[(size_t)id] = 1; }   // :1108-1111
    }
    int64_t truenum = 0;
    for (char c : ever_visible) truenum += c;
    return (double)truenum / (double)n;                            // :1113-1117 (uniqueSet of allVisible)
  }

  Params prm;
  Camera cam;
  PerceptionUtils percep_utils_;
  std::vector<float> cloud;
  int xPixel = 0, yPixel = 0;
  double xRes = 0, yRes = 0, ldc[2] = {0, 0};
  CovCounters ctr;

 private:
  std::vector<char> state;
  std::vector<int> occupant;
  std::vector<double> zdist;
  std::vector<size_t> touched;
};

}  // namespace orc
