// fcvm_pose.cu — the ordered pitch / yaw sums of getPose on the device (viewpoint_manager.cpp:398-447).
//
// For every ray that survives E1 the reference computes, in point order,
//     ray_dir = (pt - vp).normalized();  pitch = asin(ray_dir.z / (ray_dir.norm() + 1e-3));  avg_pitch += pitch;
//     yaw_off = acos(clamp(fov_sample_ray . ray_dir, -1, 1)), negated when (ray_dir_yaw x fov_yaw_ray).z < 0;  avg_yaw += yaw_off;
// Round 1 shipped the visible sets to the host as index lists and spent ~40 ns of host time per ray there (glibc's asin /
// acos are not correctly rounded, so a device libm cannot stand in for them).  Here:
//   k_pose_terms   one warp per viewpoint walks its index list; every lane evaluates one ray with exactly the host's
//                  operation order (FP64, no contraction) and obtains asin / acos from fcvm_ddmath.h: the correctly
//                  rounded value plus a flag when the true result is within 1/32 ulp of a rounding midpoint - the only
//                  place glibc can differ.  Values go to per-ray arrays, flagged arguments (~12 %) to a compact list.
//   host           calls the real glibc asin / acos on the flagged arguments only (fcvm_host.cu).
//   k_pose_patch   writes those values over the flagged rays.
//   k_pose_sum     one warp per viewpoint adds the per-ray values strictly in point order (the sums are not associative).
// The result is bit-identical to the reference's loop; tests/test_gpu_pose_sums.py compares it with the all-host path.
#include "fcvm_ddmath.h"
#include "fcvm_kernels.h"

namespace fcvm {

__global__ void __launch_bounds__(256) k_pose_terms(const float4* __restrict__ points, const PoseAux* __restrict__ aux, const long long* __restrict__ offsets,
                                                    const uint32_t* __restrict__ indices, int nrows, double* __restrict__ val_p, double* __restrict__ val_y,
                                                    unsigned long long* __restrict__ key_p, double* __restrict__ arg_p, unsigned long long* __restrict__ key_y,
                                                    double* __restrict__ arg_y, unsigned long long* __restrict__ nflag, unsigned long long cap) {
  const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (r >= nrows) return;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const PoseAux a = aux[r];
  const long long k0 = offsets[r], k1 = offsets[r + 1];
  for (long long kb = k0; kb < k1; kb += 32) {
    const long long k = kb + lane;
    const bool on = k < k1;
    bool fp = false, fy = false, neg = false;
    double a1 = 0.0, yv = 0.0;
    if (on) {
      const float4 q = __ldg(points + __ldg(indices + k));
      const double dx = (double)q.x - a.vx, dy = (double)q.y - a.vy, dz = (double)q.z - a.vz;      // cand - vp
      const double z = (dx * dx + dy * dy) + dz * dz;
      double rx = dx, ry = dy, rz = dz;
      if (z > 0.0) { const double n = sqrt(z); rx = dx / n; ry = dy / n; rz = dz / n; }            // normalized()
      const double nr = sqrt((rx * rx + ry * ry) + rz * rz);                                        // ray_dir.norm()
      a1 = rz / (nr + 1e-3);
      yv = (a.fsx * rx + a.fsy * ry) + a.fsz * rz;
      yv = (yv < -1.0) ? -1.0 : yv;                                                                 // std::max(yv, -1.0)
      yv = (1.0 < yv) ? 1.0 : yv;                                                                   // std::min(.., 1.0)
      neg = dx * a.fyy - dy * a.fyx < 0;                                                            // cross(ray_dir_yaw, fov_yaw_ray).z
      const double vp_ = ddm::asin_cr(a1, asin(fabs(a1)), &fp);
      const double vy_ = ddm::acos_cr(yv, acos(fabs(yv)), &fy);
      val_p[k] = vp_;
      val_y[k] = neg ? -vy_ : vy_;
    }
    const unsigned mp = __ballot_sync(0xffffffffu, fp), my = __ballot_sync(0xffffffffu, fy);
    if (mp | my) {
      unsigned long long bp = 0, by = 0;
      if (lane == 0) {
        if (mp) bp = atomicAdd(nflag + 0, (unsigned long long)__popc(mp));
        if (my) by = atomicAdd(nflag + 1, (unsigned long long)__popc(my));
      }
      bp = __shfl_sync(0xffffffffu, bp, 0);
      by = __shfl_sync(0xffffffffu, by, 0);
      if (fp) {
        const unsigned long long s = bp + __popc(mp & lt_mask);
        if (s < cap) { key_p[s] = (unsigned long long)k; arg_p[s] = a1; }
      }
      if (fy) {
        const unsigned long long s = by + __popc(my & lt_mask);
        if (s < cap) { key_y[s] = ((unsigned long long)k << 1) | (neg ? 1ull : 0ull); arg_y[s] = yv; }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_pose_patch(const unsigned long long* __restrict__ key_p, const double* __restrict__ g_p, long long np,
                                                    const unsigned long long* __restrict__ key_y, const double* __restrict__ g_y, long long ny,
                                                    double* __restrict__ val_p, double* __restrict__ val_y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) val_p[key_p[i]] = g_p[i];
  if (i < ny) {
    const unsigned long long key = key_y[i];
    const double g = g_y[i];
    val_y[key >> 1] = (key & 1ull) ? -g : g;
  }
}

// sums[2 r] = ((0 + v_0) + v_1) + ... of val_p over row r, sums[2 r + 1] likewise for val_y: strictly sequential adds in
// point order.  Coalesced 32-value loads; the chain itself runs in every lane (only the order matters, not who adds).
__global__ void __launch_bounds__(256) k_pose_sum(const long long* __restrict__ offsets, int nrows, const double* __restrict__ val_p,
                                                  const double* __restrict__ val_y, double* __restrict__ sums) {
  const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (r >= nrows) return;
  const int lane = threadIdx.x & 31;
  const long long k0 = offsets[r], k1 = offsets[r + 1];
  double sp = 0.0, sy = 0.0;
  for (long long kb = k0; kb < k1; kb += 32) {
    const long long k = kb + lane;
    const double vp = k < k1 ? val_p[k] : 0.0, vy = k < k1 ? val_y[k] : 0.0;
    const int m = (int)min(32ll, k1 - kb);
    if (m == 32) {
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        sp = sp + __shfl_sync(0xffffffffu, vp, l);
        sy = sy + __shfl_sync(0xffffffffu, vy, l);
      }
    } else {
      for (int l = 0; l < m; ++l) {
        sp = sp + __shfl_sync(0xffffffffu, vp, l);
        sy = sy + __shfl_sync(0xffffffffu, vy, l);
      }
    }
  }
  if (lane == 0) { sums[2 * (long long)r] = sp; sums[2 * (long long)r + 1] = sy; }
}

cudaError_t preload_pose_kernels() {
  cudaFuncAttributes fa;
  const void* fns[] = {(const void*)k_pose_terms, (const void*)k_pose_patch, (const void*)k_pose_sum};
  for (const void* f : fns) {
    const cudaError_t e = cudaFuncGetAttributes(&fa, f);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_pose_terms(const float4* points, const PoseAux* aux, const long long* offsets, const uint32_t* indices, int nrows, double* val_p,
                              double* val_y, unsigned long long* key_p, double* arg_p, unsigned long long* key_y, double* arg_y,
                              unsigned long long* nflag, unsigned long long cap, cudaStream_t s) {
  if (nrows <= 0) return cudaSuccess;
  const int wpb = 8;
  k_pose_terms<<<(unsigned)((nrows + wpb - 1) / wpb), wpb * 32, 0, s>>>(points, aux, offsets, indices, nrows, val_p, val_y, key_p, arg_p, key_y, arg_y, nflag, cap);
  return cudaGetLastError();
}

cudaError_t launch_pose_patch(const unsigned long long* key_p, const double* g_p, long long np, const unsigned long long* key_y, const double* g_y,
                              long long ny, double* val_p, double* val_y, cudaStream_t s) {
  const long long n = np > ny ? np : ny;
  if (n <= 0) return cudaSuccess;
  k_pose_patch<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(key_p, g_p, np, key_y, g_y, ny, val_p, val_y);
  return cudaGetLastError();
}

cudaError_t launch_pose_sum(const long long* offsets, int nrows, const double* val_p, const double* val_y, double* sums, cudaStream_t s) {
  if (nrows <= 0) return cudaSuccess;
  const int wpb = 8;
  k_pose_sum<<<(unsigned)((nrows + wpb - 1) / wpb), wpb * 32, 0, s>>>(offsets, nrows, val_p, val_y, sums);
  return cudaGetLastError();
}

}  // namespace fcvm
